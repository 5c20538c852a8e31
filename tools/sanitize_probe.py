"""Small end-to-end renders for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, workloads
from tests import scenes
for name in ("welsh_variants", "fm", "drums_and_sampler", "effects_rack"):
    g = Engine(44100.0, max_block=2048); n = scenes.ALL_SCENES[name](g); y = g.render(min(n, 6000)); g.close()
    print(name, float(np.abs(y).max()))
g = Engine(48000.0); n = workloads.build_cfg4(g, workloads.cfg4_slice(64, 3000)); y = g.render(n); g.close(); print("cfg4", float(np.abs(y).max()))
g = Engine(48000.0, max_block=3000); n, _ = workloads.build_cfg5(g, 24, frames=3000, note_off=1500); y = g.render(n); g.close(); print("cfg5", float(np.abs(y).max()))
