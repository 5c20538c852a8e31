"""Debug probe for welsh_solo_kernel (run under gpurun with a timeout)."""
import faulthandler, os, sys, time
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from groove_b200 import Engine, workloads
voices = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
def p(*a):
    print(*a, flush=True)
cfg = workloads.cfg4_slice(voices, frames)
e = Engine(48000.0, device=0)
p("engine")
workloads.build_cfg4(e, cfg)
p("built")
faulthandler.dump_traceback_later(90, exit=True)
try:
    y = e.render(frames)
    p("rendered", float(np.abs(y).max()))
except Exception as ex:
    p("error:", ex)
st = e.stats()
p({k: (list(getattr(st, k)) if k == "solo_class_items" else getattr(st, k)) for k, _ in st._fields_})
e.close()
p("closed")
