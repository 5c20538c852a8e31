"""Ad-hoc GPU-vs-oracle parity probe (run under gpurun): prints max abs error per scene."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine
from tests.oracle_binding import OracleEngine
from tests import scenes

names = sys.argv[1:] or list(scenes.ALL_SCENES)
for name in names:
    fn = scenes.ALL_SCENES[name]
    o = OracleEngine(44100.0); n = fn(o); ref = o.render(n)
    for mb in (0, 1000):
        g = Engine(44100.0, max_block=mb); fn(g)
        t = time.time(); out = g.render(n); dt = time.time() - t
        err = np.abs(out - ref)
        i = int(err.argmax() // 2)
        print(f"{name:20s} max_block={mb:5d} frames={n} peak={np.abs(ref).max():.4f} maxerr={err.max():.3e} at frame {i} gpu_s={dt:.4f}", flush=True)
        g.close()
