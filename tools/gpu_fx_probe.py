"""Per-effect GPU-vs-oracle probe (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from groove_b200 import Engine, abi
from tests.oracle_binding import OracleEngine
from tests import scenes

FX = {
    "gain": (abi.FX_GAIN, abi.GainParams(0.5)),
    "limiter": (abi.FX_LIMITER, abi.LimiterParams(0.02, 0.1)),
    "bitcrusher": (abi.FX_BITCRUSHER, abi.BitcrusherParams(8)),
    "compressor": (abi.FX_COMPRESSOR, abi.CompressorParams(0.05, 0.25, 0, 0)),
    "delay": (abi.FX_DELAY, abi.DelayParams(0.011)),
    "chorus": (abi.FX_CHORUS, abi.ChorusParams(4, 0.013, 0.6)),
    "reverb": (abi.FX_REVERB, abi.ReverbParams(0.8, 0.4)),
    "lpf12": (abi.FX_LOW_PASS_12DB, abi.BiquadParams(900.0, 0.9)),
    "hpf12": (abi.FX_HIGH_PASS_12DB, abi.BiquadParams(500.0, 2.0)),
    "bpf12": (abi.FX_BAND_PASS_12DB, abi.BiquadParams(700.0, 1.5)),
    "apf12": (abi.FX_ALL_PASS_12DB, abi.BiquadParams(1500.0, 20.0)),
    "lpf24": (abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(600.0, 2.0)),
    "mixer": (abi.FX_MIXER, None),
}

def build(r, name, controls):
    p = scenes.generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_NOISE, mix=0.7, cutoff_end=0.8, voices=2, gain=0.9)
    u = r.add_instrument(abi.INST_WELSH, p)
    kind, params = FX[name]
    f = r.add_effect(kind, params)
    r.patch_chain([u, f, abi.MAIN_MIXER]); r.finalize()
    ev = [(0, u, abi.EV_NOTE_ON, 50, 127, 0.0), (333, u, abi.EV_NOTE_ON, 57, 127, 0.0),
          (5000, u, abi.EV_NOTE_OFF, 50, 0, 0.0), (5600, u, abi.EV_NOTE_OFF, 57, 0, 0.0)]
    if controls:
        for k, fr in enumerate(range(64, 9000, 640)):
            ev.append((fr, f, abi.EV_CONTROL, 0, 0, 0.3 + 0.04 * k))
    r.push_events(sorted(ev, key=lambda e: e[0]))
    return 12345

for name in FX:
    for controls in (False, True):
        if controls and name in ("delay", "mixer", "chorus"): continue
        o = OracleEngine(44100.0); n = build(o, name, controls); ref = o.render(n)
        for mb in (0, 1000):
            g = Engine(44100.0, max_block=mb); build(g, name, controls); out = g.render(n)
            err = np.abs(out - ref); i = int(err.argmax() // 2)
            print(f"{name:11s} ctl={int(controls)} mb={mb:5d} peak={np.abs(ref).max():.3f} maxerr={err.max():.3e} @ {i}", flush=True)
            g.close()
