"""One short run of a bench workload for `ncu` (run under gpurun): builds the engine once and renders it
`--reps` times (fresh engine each time).  Development tool; never part of the product path.

    python tools/prof_run.py cfg5 [--variants 8192]       # config 5: one-shot patch variants
    python tools/prof_run.py cfg4 [--voices 4096 --seconds 6]
    python tools/prof_run.py tv                            # config 4 recipe with the filter decay stretched
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from groove_b200 import Engine, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload", choices=["cfg4", "cfg5", "tv", "strong", "fx", "piano"])
ap.add_argument("--variants", type=int, default=8192)
ap.add_argument("--voices", type=int, default=4096)
ap.add_argument("--seconds", type=float, default=6.0)
ap.add_argument("--groups", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
a = ap.parse_args()
for _ in range(a.reps):
    if a.workload == "cfg5":
        frames = workloads.CFG5_FRAMES
        e = Engine(48000.0, device=0, max_block=frames)
        e.set_timing(True)
        workloads.build_cfg5(e, a.variants)
    elif a.workload == "fx":
        frames = 1 << 16
        e = Engine(48000.0, device=0, max_block=frames)
        e.set_timing(True)
        e.push_events(workloads.build_fx_chains(e, 1024, frames))
    else:
        frames = int(round(a.seconds * 48000))
        cfg = workloads.Cfg4(total_voices=a.voices, frames=frames, note_off_base=int(frames * 2_400_000 / 2_880_000),
                             groups=a.groups or min(128, a.voices), filter_decay=120.0 if a.workload == "tv" else 3.29)
        e = Engine(48000.0, device=0)
        e.set_timing(True)
        workloads.build_cfg4(e, cfg, params=workloads.piano_params if a.workload == "piano" else None)
    e.render_device(frames)
    st = e.stats()
    print(json.dumps({k: (list(getattr(st, k)) if k == "solo_class_items" else getattr(st, k)) for k, _ in st._fields_}))
    e.close()
