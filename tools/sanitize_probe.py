"""Small end-to-end renders for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, workloads
from tests import scenes
for name in ("welsh_variants", "welsh_sustain", "fm", "drums_and_sampler", "effects_rack", "sidechain"):
    g = Engine(44100.0, max_block=2048); n = scenes.ALL_SCENES[name](g); y = g.render(min(n, 6000)); g.close()
    print(name, float(np.abs(y).max()))
g = Engine(48000.0); n = workloads.build_cfg4(g, workloads.cfg4_slice(64, 3000)); y = g.render(n); g.close(); print("cfg4", float(np.abs(y).max()))
g = Engine(48000.0, max_block=3000); n, _ = workloads.build_cfg5(g, 24, frames=3000, note_off=1500); y = g.render(n); g.close(); print("cfg5", float(np.abs(y).max()))
# resting-voice kernel: 20 held voices with fast envelopes, 1024-frame chunks (pairs, singles, hand-over chunks)
from groove_b200 import abi
g = Engine(48000.0, max_block=1024)
u = g.add_instrument(abi.INST_WELSH, scenes.generic_welsh(w1=abi.WAVE_PULSE_WIDTH, pw1=0.1, w2=abi.WAVE_SQUARE, voices=20,
                     routing=abi.LFO_AMPLITUDE, depth=0.05, lfo_hz=7.5, filt=(0.0, 0.005, 0.6, 0.01), amp=(0.005, 0.0, 1.0, 0.0)))
g.patch(u, abi.MAIN_MIXER); g.finalize()
for v in range(20):
    g.note_on(3 + v, u, 40 + v); g.note_off(4200 + v, u, 40 + v)
y = g.render(6000); st = g.stats(); g.close(); print("rest", float(np.abs(y).max()), int(st.rest_kernel_launches))
