"""Regenerates tests/golden/scenes.npz from the CPU oracle.

The reference cannot be run (no Rust toolchain; DSP source absent — SURVEY.md §0), so these vectors
pin the ORACLE, not the reference: they make any later change of the oracle's arithmetic visible.
Stored per scene: the first 256 frames, every 61st frame, and the float64 sum / sum of squares.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tests import scenes
from tests.oracle_binding import OracleEngine

out = {}
for name, fn in scenes.ALL_SCENES.items():
    o = OracleEngine(44100.0)
    n = fn(o)
    y = o.render(n)
    out[name + "/head"] = y[:256].copy()
    out[name + "/stride61"] = y[::61].copy()
    out[name + "/stats"] = np.array([n, y.sum(), (y * y).sum(), np.abs(y).max()])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "scenes.npz"), **out)
print("wrote", len(out), "arrays")
