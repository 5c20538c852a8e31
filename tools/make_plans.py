"""Compiles reference project fixtures into plan fixtures under tests/golden/plans/ (run in the build
container, where /root/reference exists; the GPU box only sees the committed outputs).

Each plan is this repo's own compiled form (entities + cables + frame-stamped events) of a reference
project; the 707 drum samples it needs (CC0, assets/samples/elphnt.io/707/LICENSE.txt) are stored as
24-bit integers in samples707.npz.  Oracle renders of every plan are stored decimated for regression.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from groove_b200 import project
from tests.oracle_binding import OracleEngine

REF = "/root/reference"
PROJECTS = {
    "drums-filtered-24db": "projects/demos/effects/drums-filtered-24db.json",   # BASELINE config 1
    "perf-1": "test-data/perf-1.json",                                           # config 2
    "kitchen-sink": "test-data/kitchen-sink.json",                               # config 3
    "delay": "projects/demos/effects/delay.json",
    "chorus": "projects/demos/effects/chorus.json",
    "compressor": "projects/demos/effects/compressor.json",
    "drums-reverb": "projects/demos/effects/drums-reverb.json",
    "fm-synthesizer": "projects/demos/instruments/fm-synthesizer.json",
    "arpeggiator": "projects/demos/controllers/arpeggiator.json",
    "stereo-automation": "projects/demos/controllers/stereo-automation.json",
    "sidechain": "projects/demos/controllers/sidechain.json",
}
out_dir = os.path.join(ROOT, "tests", "golden", "plans")
os.makedirs(out_dir, exist_ok=True)
loader = project.ProjectLoader(os.path.join(REF, "assets"))
used = set()
renders = {}
for name, rel in PROJECTS.items():
    path = os.path.join(REF, rel)
    if not os.path.exists(path):
        print("missing", rel); continue
    plan = loader.load(path, 44100.0)
    keys = {ev[3] for ev in plan.events if ev[2] == 1}
    for e in plan.entities:
        if e.samples and e.kind == 4:
            e.samples = [s for s in e.samples if s[0] in keys]
        for s in e.samples:
            used.add(s[1])
    with open(os.path.join(out_dir, name + ".plan.json"), "w") as f:
        f.write(plan.to_json())
    o = OracleEngine(plan.sample_rate)
    project.build_plan(o, plan, loader.sample)
    y = o.render(plan.frames)
    renders[name + "/stride41"] = y[::41].copy()
    renders[name + "/stats"] = np.array([plan.frames, y.sum(), (y * y).sum(), np.abs(y).max()])
    print(f"{name:22s} frames={plan.frames:6d} events={len(plan.events):5d} peak={np.abs(y).max():.4f} skipped={len(plan.skipped)}")
samples = {}
for nm in sorted(used):
    x, sr = loader.sample(nm)
    assert sr == 44100.0 and x.ndim == 1
    samples[nm] = np.round(x * 8388608.0).astype(np.int32)
np.savez_compressed(os.path.join(out_dir, "samples707.npz"), **samples)
np.savez_compressed(os.path.join(out_dir, "oracle_renders.npz"), **renders)
print("samples:", sorted(used))
