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
    # bare oscillator / envelope devices driving the effect families (round 2)
    "drums-filtered-12db": "projects/demos/effects/drums-filtered-12db.json",
    "drums-filtered-q": "projects/demos/effects/drums-filtered-q.json",
    "drums-chorus": "projects/demos/effects/drums-chorus.json",
    "drums": "projects/demos/instruments/drums.json",
    "filter-lp12-noise": "projects/demos/effects/filter-low-pass-12db_noise_cutoff-1000_q-0.707.json",
    "filter-lp12-sine-q20": "projects/demos/effects/filter-low-pass-12db_sine_cutoff-1000_q-20.json",
    "filter-hp12-noise-q20": "projects/demos/effects/filter-high-pass-12db_noise_cutoff-1000_q-20.json",
    "filter-hp12-sine": "projects/demos/effects/filter-high-pass-12db_sine_cutoff-1000_q-0.707.json",
    "filter-bp12-noise-bw30": "projects/demos/effects/filter-band-pass-12db_noise_cutoff-1000_bandwidth-30.json",
    "filter-bp12-sine-bw2000": "projects/demos/effects/filter-band-pass-12db_sine_cutoff-1000_bandwidth-2000.json",
    "filter-bs12-noise-bw2": "projects/demos/effects/filter-band-stop-12db_noise_cutoff-1000_bandwidth-2.json",
    "filter-bs12-sine-bw30": "projects/demos/effects/filter-band-stop-12db_sine_cutoff-1000_bandwidth-30.json",
    "filter-ap12-noise-q20": "projects/demos/effects/filter-all-pass-12db_noise_cutoff-1000_q-20.json",
    "filter-ap12-sine": "projects/demos/effects/filter-all-pass-12db_sine_cutoff-1000_q-0.707.json",
    "filter-peak12-noise-30db": "projects/demos/effects/filter-peaking-eq-12db_noise_cutoff-1000_db-gain-30.json",
    "filter-peak12-sine-6db": "projects/demos/effects/filter-peaking-eq-12db_sine_cutoff-1000_db-gain-6.json",
    "filter-ls12-noise-6db": "projects/demos/effects/filter-low-shelf-12db_noise_cutoff-1000_db-gain-6.json",
    "filter-ls12-sine-30db": "projects/demos/effects/filter-low-shelf-12db_sine_cutoff-1000_db-gain-30.json",
    "filter-hs12-noise-30db": "projects/demos/effects/filter-high-shelf-12db_noise_cutoff-1000_db-gain-30.json",
    "filter-hs12-sine-6db": "projects/demos/effects/filter-high-shelf-12db_sine_cutoff-1000_db-gain-6.json",
    "filter-lp24-ripple-sweep": "projects/demos/effects/filter-low-pass-24db_noise_cutoff-1000_passband-ripple-sweep.json",
    "filter-lp12-sweep-down": "projects/demos/effects/filter-lpf-12db-noise-sweep-down.json",
    "gain-sine-0.5": "projects/demos/effects/gain_sine_ceiling-0.500.json",
    "gain-noise-0.1": "projects/demos/effects/gain_noise_ceiling-0.100.json",
    "limiter-noise": "projects/demos/effects/limiter_noise_min-0.400_max-0.600.json",
    "limiter-sine": "projects/demos/effects/limiter_sine_min-0.100_max-0.900.json",
    "bitcrusher-saw-8": "projects/demos/effects/bitcrusher_sawtooth_bits-to-crush-8.json",
    "bitcrusher-tri-13": "projects/demos/effects/bitcrusher_triangle_bits-to-crush-13.json",
    "oscillator-pw10": "projects/demos/instruments/oscillator-pulse-width-10-percent-a4.json",
    "oscillator-tri-1hz": "projects/demos/instruments/oscillator-triangle-1Hz.json",
    "oscillator-square-1k": "projects/demos/instruments/oscillator-square-1000Hz.json",
    "oscillator-noise": "projects/demos/instruments/oscillator-noise.json",
    "envelope-adsr": "projects/demos/instruments/envelope-adsr-linear.json",
    "fm-beta-10": "projects/demos/instruments/fm-synthesizer-beta-10.0.json",
    "welsh-piano": "projects/demos/instruments/welsh-piano.json",          # hard sync
    "welsh-cello": "projects/demos/instruments/welsh-cello.json",
    "welsh-lfo-pitch": "projects/demos/instruments/welsh-test-lfo-pitch.json",
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
