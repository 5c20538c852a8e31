"""Wall-clock of host-buffer vs device-resident renders of config 4 (development probe, run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, workloads
frames = 2_880_000
cfg = workloads.Cfg4()
out = np.empty((frames, 2))
for mode in ("host", "device", "host", "host", "host", "host", "host", "host", "device", "host"):
    e = Engine(48000.0, max_block=1 << 16)
    e.set_timing(True)
    workloads.build_cfg4(e, cfg)
    t = time.perf_counter()
    if mode == "host":
        e.render(frames, out)
    else:
        e.render_device(frames)
    dt = time.perf_counter() - t
    st = e.stats()
    e.close()
    print(f"{mode:6s} wall {dt * 1e3:7.2f} ms  render_ms(events) {st.render_ms:7.2f}", flush=True)
