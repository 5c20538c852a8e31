"""Config 5 through the engine under several scheduler settings (development tool; run under gpurun).

    python tools/cfg5_sweep.py "GB_SOLO_SUB=2048" "GB_SOLO_SUB=1024 GB_SOLO_LOCKSTEP=0" ...

The settings are read by gb_create, so each configuration gets fresh engines in this one process."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from groove_b200 import Engine, workloads  # noqa: E402

variants = [workloads.cfg5_variant(i) for i in range(8192)]
KEYS = ("GB_SOLO_SUB", "GB_SOLO_LOCKSTEP", "GB_SOLO_WAVES", "GB_SOLO_MIN", "GB_SOLO_GRID")
for spec in sys.argv[1:] or [""]:
    for k in KEYS:
        os.environ.pop(k, None)
    for kv in spec.split():
        k, v = kv.split("=")
        os.environ[k] = v
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        e = Engine(48000.0, device=0, max_block=workloads.CFG5_FRAMES)
        e.set_timing(True)
        workloads.build_cfg5(e, 8192, variants=variants)
        t1 = time.perf_counter()
        e.render_device(workloads.CFG5_FRAMES)
        t2 = time.perf_counter()
        st = e.stats()
        e.close()
        res.append((st.render_ms, st.solo_kernel_ms, st.fm_kernel_ms, st.voice_kernel_ms, st.fx_kernel_ms, (t2 - t1) * 1e3, (t1 - t0) * 1e3))
    r = res[-1]
    print(json.dumps({"spec": spec, "render_ms": round(r[0], 3), "solo_ms": round(r[1], 3), "fm_ms": round(r[2], 3),
                      "voice_ms": round(r[3], 3), "fx_ms": round(r[4], 3), "render_wall_ms": round(r[5], 2),
                      "build_wall_ms": round(r[6], 1), "solo_ms_all": [round(x[1], 3) for x in res],
                      "classes": list(st.solo_class_items), "jobs": st.solo_jobs}), flush=True)
