import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, abi
from tests.oracle_binding import OracleEngine
from tests import scenes

def build(r, nfx, fx_controls, inst_controls):
    src_p = scenes.generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_NOISE, mix=0.7, cutoff_end=0.8, voices=2, gain=0.9)
    u = r.add_instrument(abi.INST_WELSH, src_p)
    fx = [r.add_effect(abi.FX_GAIN, abi.GainParams(0.5 + 0.01 * i)) for i in range(nfx)]
    sub = r.add_effect(abi.FX_MIXER)
    for f in fx:
        r.patch_chain([u, f, sub])
    r.patch(sub, abi.MAIN_MIXER)
    r.finalize()
    ev = [(0, u, abi.EV_NOTE_ON, 50, 127, 0.0), (5000, u, abi.EV_NOTE_OFF, 50, 0, 0.0)]
    if fx_controls:
        for k, f in enumerate(range(64, 9000, 640)):
            ev.append((f + 64, fx[0], abi.EV_CONTROL, 0, 0, 0.5 + 0.03 * k))
    if inst_controls:
        ev.append((2048, u, abi.EV_CONTROL, 1, 0, 0.2))
        ev.append((4096, u, abi.EV_CONTROL, 0, 0, 0.6))
    r.push_events(sorted(ev, key=lambda e: e[0]))
    return 12345

for nfx in (4, 8, 9, 16):
    for fc, ic in ((0, 0), (1, 0), (0, 1)):
        o = OracleEngine(44100.0); n = build(o, nfx, fc, ic); ref = o.render(n)
        g = Engine(44100.0); build(g, nfx, fc, ic); out = g.render(n)
        err = np.abs(out - ref).max(axis=1)
        bad = np.nonzero(err > 1e-9)[0]
        print(f"nfx={nfx:2d} fxctl={fc} instctl={ic} maxerr={err.max():.3e} first_bad={bad[0] if len(bad) else -1} nbad={len(bad)}", flush=True)
        g.close()
