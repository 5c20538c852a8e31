"""Access to the committed plan fixtures (tests/golden/plans): compiled reference projects + 707 samples."""
import os

import numpy as np

from groove_b200 import project

PLAN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plans")
PLAN_NAMES = sorted(f[:-len(".plan.json")] for f in os.listdir(PLAN_DIR) if f.endswith(".plan.json"))
_samples = None


def sample(name):
    global _samples
    if _samples is None:
        _samples = np.load(os.path.join(PLAN_DIR, "samples707.npz"))
    return _samples[name].astype(np.float64) / 8388608.0, 44100.0


def load_plan(name) -> project.Plan:
    with open(os.path.join(PLAN_DIR, name + ".plan.json")) as f:
        return project.Plan.from_json(f.read())


def oracle_renders():
    return np.load(os.path.join(PLAN_DIR, "oracle_renders.npz"))
