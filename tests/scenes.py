"""Deterministic scenes built through the block-render ABI.  Each scene function takes any
``abi.Renderer`` (the CUDA engine or the CPU oracle), builds the same graph and events on it and
returns the number of frames to render."""
from __future__ import annotations

import math

import numpy as np

from groove_b200 import abi, workloads

LOG800 = math.log(800.0)


def hz_to_pct(hz: float) -> float:
    """FrequencyHz::frequency_to_percent (settings/src/patches.rs:150-152): inverse of 25*800^pct."""
    return max(0.0, min(1.0, math.log(hz / 25.0) / LOG800))


def cello_params(voices: int = 8, gain: float = 1.0, pan: float = 0.0) -> abi.WelshParams:
    """assets/patches/welsh/cello.json through derive_welsh_synth_params (settings/src/patches.rs:87-170)."""
    p = abi.WelshParams()
    p.oscillator_1 = abi.osc(abi.WAVE_PULSE_WIDTH, 0.1)
    p.oscillator_2 = abi.osc(abi.WAVE_SQUARE)
    p.oscillator_2_sync = 0
    p.oscillator_mix = 0.5
    p.amp_envelope = abi.env(0.06, 0.0, 1.0, 0.0)           # release := decay (patches.rs:137)
    p.lfo = abi.osc(abi.WAVE_SINE, frequency=7.5)
    p.lfo_routing = abi.LFO_AMPLITUDE
    p.lfo_depth = 0.05
    p.filter_cutoff_hz = 40.0
    p.filter_passband_ripple = 0.707                         # denormalize_q(0)
    p.filter_cutoff_start = hz_to_pct(40.0)
    p.filter_cutoff_end = 0.9
    p.filter_envelope = abi.env(0.0, 3.29, 0.78, 3.29)      # release := decay (patches.rs:158)
    p.voice_dca = abi.DcaParams(1.0, 0.0)
    p.dca = abi.DcaParams(gain, pan)
    p.voices = voices
    return p


def generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_TRIANGLE, sync=0, routing=abi.LFO_NONE, depth=0.0,
                  lfo_wave=abi.WAVE_SINE, lfo_hz=5.0, cutoff_end=0.5, cutoff_start=0.3, ripple=0.707,
                  tune2=1.0, mix=0.6, pw1=0.5, pw2=0.5, amp=(0.01, 0.1, 0.7, 0.1), filt=(0.02, 0.2, 0.5, 0.2),
                  voices=4, cutoff_hz=900.0, gain=1.0, pan=0.0, fixed2=0.0) -> abi.WelshParams:
    p = abi.WelshParams()
    p.oscillator_1 = abi.osc(w1, pw1)
    p.oscillator_2 = abi.osc(w2, pw2, tune=tune2, fixed_frequency=fixed2)
    p.oscillator_2_sync = sync
    p.oscillator_mix = mix
    p.amp_envelope = abi.env(*amp)
    p.lfo = abi.osc(lfo_wave, frequency=lfo_hz)
    p.lfo_routing = routing
    p.lfo_depth = depth
    p.filter_cutoff_hz = cutoff_hz
    p.filter_passband_ripple = ripple
    p.filter_cutoff_start = cutoff_start
    p.filter_cutoff_end = cutoff_end
    p.filter_envelope = abi.env(*filt)
    p.voice_dca = abi.DcaParams(1.0, 0.0)
    p.dca = abi.DcaParams(gain, pan)
    p.voices = voices
    return p


def scene_cello_chord(r: abi.Renderer) -> int:
    u = r.add_instrument(abi.INST_WELSH, cello_params(voices=8))
    r.patch(u, abi.MAIN_MIXER)
    r.finalize()
    ev = []
    for i, key in enumerate((36, 43, 48, 52, 55)):
        ev.append((64 * i, u, abi.EV_NOTE_ON, key, 127, 0.0))
        ev.append((9000 + 640 * i, u, abi.EV_NOTE_OFF, key, 0, 0.0))
    r.push_events(ev)
    return 12000


def scene_welsh_variants(r: abi.Renderer) -> int:
    """Every waveform / LFO routing / sync combination, notes on odd frames, retriggers and steals."""
    cfgs = [
        dict(w1=abi.WAVE_SINE, w2=abi.WAVE_SAWTOOTH, routing=abi.LFO_NONE),
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SQUARE, sync=1, tune2=1.4983070768766815, routing=abi.LFO_AMPLITUDE, depth=0.3),
        dict(w1=abi.WAVE_PULSE_WIDTH, pw1=0.3, w2=abi.WAVE_PULSE_WIDTH, pw2=0.7, routing=abi.LFO_PULSE_WIDTH, depth=0.4,
             lfo_wave=abi.WAVE_TRIANGLE),
        dict(w1=abi.WAVE_TRIANGLE, w2=abi.WAVE_NOISE, routing=abi.LFO_FILTER_CUTOFF, depth=0.5, cutoff_end=0.0,
             cutoff_start=0.6),
        dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_SAWTOOTH, routing=abi.LFO_PITCH, depth=0.08, lfo_hz=6.0),
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SAWTOOTH, sync=1, tune2=2.7, routing=abi.LFO_PITCH, depth=0.2,
             lfo_wave=abi.WAVE_SAWTOOTH, lfo_hz=3.0),
        dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_NONE, mix=1.0, cutoff_end=0.0, routing=abi.LFO_NONE, cutoff_hz=1200.0),
        dict(w1=abi.WAVE_DEBUG_MAX, w2=abi.WAVE_DEBUG_MIN, mix=0.25, ripple=3.0, amp=(0.0, 0.0, 1.0, 0.0)),
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SINE, fixed2=220.0, ripple=10.707, cutoff_start=0.2, cutoff_end=1.0),
    ]
    uids = []
    for i, c in enumerate(cfgs):
        u = r.add_instrument(abi.INST_WELSH, generic_welsh(voices=3, gain=0.5, pan=-0.8 + 0.2 * i, **c))
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    ev = []
    for i, u in enumerate(uids):
        base = 37 * i
        ev += [(base + 1, u, abi.EV_NOTE_ON, 45 + i, 127, 0.0),
               (base + 700, u, abi.EV_NOTE_ON, 52 + i, 127, 0.0),
               (base + 1203, u, abi.EV_NOTE_OFF, 45 + i, 0, 0.0),
               (base + 1301, u, abi.EV_NOTE_ON, 45 + i, 127, 0.0),      # retrigger while releasing
               (base + 1302, u, abi.EV_NOTE_ON, 60 + i, 127, 0.0),
               (base + 1305, u, abi.EV_NOTE_ON, 64 + i, 127, 0.0),      # steals the oldest voice
               (base + 2500, u, abi.EV_NOTE_OFF, 52 + i, 0, 0.0),
               (base + 2500, u, abi.EV_NOTE_OFF, 60 + i, 0, 0.0),
               (base + 2501, u, abi.EV_NOTE_OFF, 64 + i, 0, 0.0),
               (base + 2600, u, abi.EV_NOTE_OFF, 45 + i, 0, 0.0),
               (base + 9000, u, abi.EV_NOTE_ON, 40 + i, 127, 0.0),      # restart from idle
               (base + 9900, u, abi.EV_NOTE_OFF, 40 + i, 0, 0.0)]
    r.push_events(ev)
    return 16000


def scene_welsh_sustain(r: abi.Renderer) -> int:
    """Long notes whose filter envelope reaches its sustain level (the cutoff rests: the engine's
    time-invariant block), fixed-cutoff filters, releases that set the cutoff moving again, and
    retriggers in the middle of a resting stretch."""
    fast_filt = (0.0, 0.01, 0.6, 0.05)
    cfgs = [
        dict(w1=abi.WAVE_PULSE_WIDTH, pw1=0.1, w2=abi.WAVE_SQUARE, mix=0.5, routing=abi.LFO_AMPLITUDE, depth=0.05,
             lfo_hz=7.5, filt=fast_filt, amp=(0.06, 0.0, 1.0, 0.0), cutoff_start=hz_to_pct(40.0), cutoff_end=0.9),
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_TRIANGLE, tune2=1.0029, routing=abi.LFO_NONE, filt=fast_filt,
             amp=(0.001, 0.05, 0.8, 0.02), ripple=2.5, cutoff_start=0.2, cutoff_end=0.4),
        dict(w1=abi.WAVE_TRIANGLE, w2=abi.WAVE_PULSE_WIDTH, pw2=0.8, mix=0.3, cutoff_end=0.0, cutoff_hz=700.0,
             routing=abi.LFO_AMPLITUDE, depth=0.5, lfo_hz=11.0),
        dict(w1=abi.WAVE_SQUARE, w2=abi.WAVE_SAWTOOTH, filt=(0.0, 0.0, 1.0, 0.0), cutoff_start=0.1, cutoff_end=1.0,
             ripple=10.707, amp=(0.0, 0.0, 1.0, 0.0)),
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SINE, filt=fast_filt, routing=abi.LFO_NONE),       # sine: not this path
        dict(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SQUARE, filt=(0.0, 0.005, 0.0, 0.005), cutoff_start=0.15, cutoff_end=0.7),
    ]
    uids = []
    for i, c in enumerate(cfgs):
        u = r.add_instrument(abi.INST_WELSH, generic_welsh(voices=3, gain=0.4, pan=-0.5 + 0.2 * i, **c))
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    ev = []
    for i, u in enumerate(uids):
        base = 53 * i
        ev += [(base + 3, u, abi.EV_NOTE_ON, 40 + 3 * i, 127, 0.0),
               (base + 900, u, abi.EV_NOTE_ON, 47 + 3 * i, 127, 0.0),
               (base + 6000, u, abi.EV_NOTE_ON, 40 + 3 * i, 127, 0.0),     # retrigger a resting voice
               (base + 9001, u, abi.EV_NOTE_OFF, 47 + 3 * i, 0, 0.0),
               (base + 15000, u, abi.EV_NOTE_OFF, 40 + 3 * i, 0, 0.0),
               (base + 15400, u, abi.EV_NOTE_ON, 52 + 3 * i, 127, 0.0),
               (base + 21000, u, abi.EV_NOTE_OFF, 52 + 3 * i, 0, 0.0)]
    r.push_events(ev)
    return 24000


def fm_params(ratio=2.0, depth=1.0, beta=1.0, car=(0.01, 0.1, 0.8, 0.2), mod=(0.0, 0.3, 0.4, 0.3),
              gain=1.0, pan=0.0, voices=4) -> abi.FmParams:
    p = abi.FmParams()
    p.ratio, p.depth, p.beta = ratio, depth, beta
    p.carrier_envelope = abi.env(*car)
    p.modulator_envelope = abi.env(*mod)
    p.dca = abi.DcaParams(gain, pan)
    p.voices = voices
    return p


def scene_fm(r: abi.Renderer) -> int:
    uids = []
    for i, (ratio, beta) in enumerate(((2.0, 1.0), (0.5, 10.0), (3.0, 0.1), (1.0, 100.0), (4.0, 0.0))):
        u = r.add_instrument(abi.INST_FM, fm_params(ratio=ratio, beta=beta, gain=0.4, pan=-0.5 + 0.25 * i, voices=2))
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    ev = []
    for i, u in enumerate(uids):
        ev += [(3 + i, u, abi.EV_NOTE_ON, 57 + 2 * i, 127, 0.0), (2000 + i, u, abi.EV_NOTE_ON, 64, 127, 0.0),
               (4000, u, abi.EV_NOTE_OFF, 57 + 2 * i, 0, 0.0), (4100, u, abi.EV_NOTE_ON, 57 + 2 * i, 127, 0.0),
               (4105, u, abi.EV_NOTE_ON, 70, 127, 0.0),
               (6000, u, abi.EV_NOTE_OFF, 64, 0, 0.0), (6000, u, abi.EV_NOTE_OFF, 70, 0, 0.0),
               (6500, u, abi.EV_NOTE_OFF, 57 + 2 * i, 0, 0.0)]
    r.push_events(ev)
    return 17000


def synthetic_sample(n: int, seed: int, stereo: bool = False) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 44100.0
    x = np.sin(2 * np.pi * (120.0 + 30 * seed) * t) * np.exp(-t * 18.0) + 0.2 * rng.uniform(-1, 1, n) * np.exp(-t * 40.0)
    x = np.clip(x * 0.8, -1.0, 1.0 - 2.0 ** -23)
    if stereo:
        return np.stack([x, np.roll(x, 7) * 0.9], axis=1)
    return x


def scene_drums_and_sampler(r: abi.Renderer) -> int:
    d = r.add_instrument(abi.INST_DRUMKIT, abi.DrumkitParams())
    for i, key in enumerate((35, 38, 42, 44)):
        r.load_sample(d, key, synthetic_sample(3000 + 900 * i, i), 44100.0)
    s = r.add_instrument(abi.INST_SAMPLER, abi.SamplerParams(261.6255653005986, 3, 0))
    r.load_sample(s, 0, synthetic_sample(5000, 9, stereo=True), 44100.0, 261.6255653005986)
    lp = r.add_effect(abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(1000.0, 0.8))
    r.patch_chain([d, lp, abi.MAIN_MIXER])
    r.patch(s, abi.MAIN_MIXER)
    r.finalize()
    ev = []
    step = 1291
    pat = [(42, 35), (44,), (42,), (44,), (42, 38, 35), (44,), (42,), (44,)]
    for i, keys in enumerate(pat * 2):
        for k in keys:
            ev.append((i * step, d, abi.EV_NOTE_ON, k, 127, 0.0))
            ev.append((i * step + step, d, abi.EV_NOTE_OFF, k, 0, 0.0))
    for i, k in enumerate((60, 64, 67, 72, 55)):
        ev.append((500 + 1500 * i, s, abi.EV_NOTE_ON, k, 127, 0.0))
        ev.append((500 + 1500 * i + 2200, s, abi.EV_NOTE_OFF, k, 0, 0.0))
    # exponential cutoff sweep on the filter, one control point per 64-frame buffer
    n = 16 * step + 4000
    for f in range(0, n, 64):
        ev.append((f, lp, abi.EV_CONTROL, 0, 0, f / n))
    r.push_events(sorted(ev, key=lambda e: e[0]))
    return n


def scene_effects_rack(r: abi.Renderer) -> int:
    """One noisy source through every effect type, in parallel chains, with automation."""
    src_p = generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_NOISE, mix=0.7, cutoff_end=0.8, voices=2, gain=0.9)
    u = r.add_instrument(abi.INST_WELSH, src_p)
    fx = []
    fx.append(r.add_effect(abi.FX_GAIN, abi.GainParams(0.5)))
    fx.append(r.add_effect(abi.FX_LIMITER, abi.LimiterParams(0.02, 0.1)))
    fx.append(r.add_effect(abi.FX_BITCRUSHER, abi.BitcrusherParams(8)))
    fx.append(r.add_effect(abi.FX_COMPRESSOR, abi.CompressorParams(0.05, 0.25, 0, 0)))
    fx.append(r.add_effect(abi.FX_DELAY, abi.DelayParams(0.011)))
    fx.append(r.add_effect(abi.FX_CHORUS, abi.ChorusParams(4, 0.013, 0.6)))
    fx.append(r.add_effect(abi.FX_REVERB, abi.ReverbParams(0.8, 0.4)))
    for kind, a, b in ((abi.FX_LOW_PASS_12DB, 900.0, 0.9), (abi.FX_HIGH_PASS_12DB, 500.0, 2.0),
                       (abi.FX_BAND_PASS_12DB, 700.0, 1.5), (abi.FX_BAND_STOP_12DB, 1000.0, 0.7),
                       (abi.FX_ALL_PASS_12DB, 1500.0, 20.0), (abi.FX_PEAKING_EQ_12DB, 800.0, 6.0),
                       (abi.FX_LOW_SHELF_12DB, 300.0, -4.0), (abi.FX_HIGH_SHELF_12DB, 3000.0, 5.0)):
        fx.append(r.add_effect(kind, abi.BiquadParams(a, b)))
    fx.append(r.add_effect(abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(600.0, 2.0)))
    sub = r.add_effect(abi.FX_MIXER)
    for f in fx:
        r.patch_chain([u, f, sub])
    out_gain = r.add_effect(abi.FX_GAIN, abi.GainParams(0.2))
    r.patch_chain([sub, out_gain, abi.MAIN_MIXER])
    r.finalize()
    ev = [(0, u, abi.EV_NOTE_ON, 50, 127, 0.0), (333, u, abi.EV_NOTE_ON, 57, 127, 0.0),
          (5000, u, abi.EV_NOTE_OFF, 50, 0, 0.0), (5600, u, abi.EV_NOTE_OFF, 57, 0, 0.0)]
    for k, f in enumerate(range(64, 9000, 640)):
        ev.append((f, fx[7], abi.EV_CONTROL, 0, 0, 0.3 + 0.04 * k))       # lpf12 cutoff
        ev.append((f + 64, fx[0], abi.EV_CONTROL, 0, 0, 0.5 + 0.03 * k))  # gain ceiling
        ev.append((f + 128, fx[15], abi.EV_SET_PARAM, 0, 0, 400.0 + 90 * k))  # lpf24 cutoff (Hz)
        ev.append((f + 192, fx[5], abi.EV_CONTROL, 2, 0, (k % 5) / 5.0))  # chorus wet/dry
        ev.append((f + 256, fx[6], abi.EV_CONTROL, 0, 0, 0.9 - 0.05 * k)) # reverb attenuation
    ev.append((2048, u, abi.EV_CONTROL, 1, 0, 0.2))                       # instrument pan
    ev.append((4096, u, abi.EV_CONTROL, 0, 0, 0.6))                       # instrument gain
    r.push_events(sorted(ev, key=lambda e: e[0]))
    return 12345


def scene_graph_toys(r: abi.Renderer) -> int:
    """orchestrator.rs:1444-1668 — gather_audio tests restated with ToyAudioSource levels."""
    a = r.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.1, 0.1))
    b = r.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.3, 0.3))
    c = r.add_instrument(abi.INST_TOY_SOURCE, abi.ToySourceParams(0.5, 0.5))
    g = r.add_effect(abi.FX_GAIN, abi.GainParams(0.5))
    r.patch(a, abi.MAIN_MIXER)
    r.patch_chain([b, g, abi.MAIN_MIXER])
    r.patch(c, g)
    r.finalize()
    return 100


def scene_sidechain(r: abi.Renderer) -> int:
    """projects/demos/controllers/sidechain.json in miniature: the magnitude of one chain's signal, sampled at
    every 64-frame control boundary, drives a compressor threshold, a gain ceiling and a limiter maximum
    of other chains (gb_link_control)."""
    drums = r.add_instrument(abi.INST_FM, fm_params(ratio=3.0, beta=4.0, car=(0.0, 0.05, 0.0, 0.05), gain=0.9))
    tap = r.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    tap2 = r.add_effect(abi.FX_SIGNAL_PASSTHROUGH)
    bass = r.add_instrument(abi.INST_WELSH, generic_welsh(w1=abi.WAVE_SAWTOOTH, w2=abi.WAVE_SQUARE, voices=2, gain=0.8))
    comp = r.add_effect(abi.FX_COMPRESSOR, abi.CompressorParams(1.0, 0.2, 0.0, 0.0))
    pad = r.add_instrument(abi.INST_WELSH, generic_welsh(w1=abi.WAVE_TRIANGLE, w2=abi.WAVE_SAWTOOTH, voices=2, gain=0.5, pan=0.4))
    gain = r.add_effect(abi.FX_GAIN, abi.GainParams(1.0))
    lim = r.add_effect(abi.FX_LIMITER, abi.LimiterParams(0.0, 1.0))
    dead = r.add_effect(abi.FX_SIGNAL_PASSTHROUGH)      # never patched: its link must stay silent
    g2 = r.add_effect(abi.FX_GAIN, abi.GainParams(0.7))
    r.patch_chain([drums, tap, abi.MAIN_MIXER])
    r.patch_chain([bass, comp, abi.MAIN_MIXER])
    r.patch_chain([pad, gain, lim, tap2, abi.MAIN_MIXER])
    r.patch_chain([bass, g2, abi.MAIN_MIXER])
    r.link_control(tap, comp, 0)
    r.link_control(tap, gain, 0)
    r.link_control(tap, lim, 1)
    r.link_control(dead, g2, 0)
    r.finalize()
    ev = []
    for k in range(6):
        ev.append((1 + 2500 * k, drums, abi.EV_NOTE_ON, 36 + k, 127, 0.0))
        ev.append((900 + 2500 * k, drums, abi.EV_NOTE_OFF, 36 + k, 0, 0.0))
    ev += [(0, bass, abi.EV_NOTE_ON, 40, 127, 0.0), (7000, bass, abi.EV_NOTE_ON, 47, 127, 0.0),
           (13000, bass, abi.EV_NOTE_OFF, 40, 0, 0.0), (13500, bass, abi.EV_NOTE_OFF, 47, 0, 0.0),
           (100, pad, abi.EV_NOTE_ON, 60, 127, 0.0), (12000, pad, abi.EV_NOTE_OFF, 60, 0, 0.0)]
    r.push_events(ev)
    return 15000


def scene_cello_held_chord(r: abi.Renderer) -> int:
    """Config 4's recipe in miniature: one 16-voice cello instrument (grouped CTAs), every note held from the
    first frames to near the end.  Rendered in 4096-frame chunks this walks the voices through the general
    kernel (note-ons, envelope stage boundaries, note-offs), welsh_sweep_kernel (filter decay, 3.29 s) and
    welsh_rest_kernel (both envelopes at sustain)."""
    u = r.add_instrument(abi.INST_WELSH, cello_params(voices=16, gain=1.0 / 16.0, pan=-0.3))
    r.patch(u, abi.MAIN_MIXER)
    r.finalize()
    ev = []
    for i in range(16):
        ev.append((64 * i, u, abi.EV_NOTE_ON, 36 + i, 127, 0.0))
        ev.append((180000 + 64 * i, u, abi.EV_NOTE_OFF, 36 + i, 0, 0.0))
    r.push_events(ev)
    return 200000


def scene_cello_ensemble(r: abi.Renderer) -> int:
    """Config 4's STRUCTURE in miniature: 12 cello instruments of 16 voices with their own pans, all patched to
    the main mixer (more than kMaxSources inputs: the mixer sums the instruments' CTA partials from a pointer
    table).  That is the layout the overlapped mixdown needs (voice kernels on their own stream, partial buffers
    alternating by chunk parity), and with GB_REST_VR=1 the all-resting chunks run over voice ranges of 14
    across instrument boundaries (welsh_rest_vr_kernel).  Filter decay shortened to 50 ms so that the voices
    rest from the second 4096-frame chunk on."""
    uids = []
    for q in range(12):
        u = r.add_instrument(abi.INST_WELSH, workloads.cello_params(16, 1.0 / 192.0, -1.0 + 2.0 * q / 11.0, 0.05))
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    ev = []
    for q, u in enumerate(uids):
        for j in range(16):
            ev.append((4 * q + 64 * (j % 4), u, abi.EV_NOTE_ON, 36 + (q + 12 * j) % 49, 127, 0.0))
            ev.append((36000 + 4 * q, u, abi.EV_NOTE_OFF, 36 + (q + 12 * j) % 49, 0, 0.0))
    ev.sort(key=lambda t: t[0])
    r.push_events(ev)
    return 45000


def scene_bare_sources(r: abi.Renderer) -> int:
    """Bare oscillator / envelope devices (the "oscillator" and "envelope" instrument types of the reference's
    filter-*, gain_*, bitcrusher_* and oscillator-* demo projects): free-running oscillators of every
    waveform into filters, and an envelope device retriggered inside its attack and inside its release."""
    def osc_dev(wave, hz, pw=0.5):
        return r.add_instrument(abi.INST_OSCILLATOR, abi.OscillatorSourceParams(abi.osc(wave, pw, frequency=hz)))
    noise = osc_dev(abi.WAVE_NOISE, 0.0)
    saw = osc_dev(abi.WAVE_SAWTOOTH, 440.0)
    sine = osc_dev(abi.WAVE_SINE, 1000.0)
    pulse = osc_dev(abi.WAVE_PULSE_WIDTH, 440.0, 0.1)
    tri = osc_dev(abi.WAVE_TRIANGLE, 1.0)
    env = r.add_instrument(abi.INST_ENVELOPE, abi.EnvelopeSourceParams(abi.env(0.1, 0.2, 0.6, 0.3)))
    lp = r.add_effect(abi.FX_LOW_PASS_12DB, abi.BiquadParams(1000.0, 0.707))
    bs = r.add_effect(abi.FX_BAND_STOP_12DB, abi.BiquadParams(1000.0, 2.0))
    crush = r.add_effect(abi.FX_BITCRUSHER, abi.BitcrusherParams(13.0))
    g1 = r.add_effect(abi.FX_GAIN, abi.GainParams(0.2))
    g2 = r.add_effect(abi.FX_GAIN, abi.GainParams(0.1))
    r.patch_chain([noise, lp, g1, abi.MAIN_MIXER])
    r.patch_chain([saw, crush, g2, abi.MAIN_MIXER])
    r.patch_chain([sine, bs, g2])
    r.patch_chain([pulse, g2])
    r.patch_chain([tri, g1])
    r.patch_chain([env, g1])
    r.finalize()
    ev = [(100, env, abi.EV_NOTE_ON, 60, 127, 0.0), (2000, env, abi.EV_NOTE_ON, 62, 127, 0.0),      # inside the attack
          (20000, env, abi.EV_NOTE_OFF, 62, 0, 0.0), (25000, env, abi.EV_NOTE_ON, 64, 127, 0.0),   # inside the release
          (26000, env, abi.EV_NOTE_OFF, 64, 0, 0.0), (26001, env, abi.EV_NOTE_OFF, 64, 0, 0.0),
          (50, saw, abi.EV_NOTE_ON, 60, 127, 0.0)]                                                  # oscillators ignore MIDI
    r.push_events(ev)
    return 45000


def scene_fx_chains(r: abi.Renderer) -> int:
    """Parallel effect chains (the shape of a batch of songs, or of config 1's chain copied): each an IIR
    stage followed by a memoryless effect, all automated.  On the GPU the IIR stages of one kind share a launch
    and the memoryless effect runs as the stage's post-op; chain 6's filter feeds TWO consumers (no fusion),
    chain 7's gain has a second source (no fusion either)."""
    def osc_dev(wave, hz, pw=0.5):
        return r.add_instrument(abi.INST_OSCILLATOR, abi.OscillatorSourceParams(abi.osc(wave, pw, frequency=hz)))
    chains = [
        (abi.WAVE_SAWTOOTH, 110.0, (abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(700.0, 0.8)), (abi.FX_GAIN, abi.GainParams(0.4))),
        (abi.WAVE_SQUARE, 220.0, (abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(1500.0, 2.5)), (abi.FX_LIMITER, abi.LimiterParams(0.0, 0.3))),
        (abi.WAVE_NOISE, 0.0, (abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(300.0, 1.2)), (abi.FX_COMPRESSOR, abi.CompressorParams(0.1, 0.5, 0, 0))),
        (abi.WAVE_TRIANGLE, 330.0, (abi.FX_BAND_PASS_12DB, abi.BiquadParams(900.0, 1.5)), (abi.FX_BITCRUSHER, abi.BitcrusherParams(9))),
        (abi.WAVE_PULSE_WIDTH, 82.0, (abi.FX_HIGH_PASS_12DB, abi.BiquadParams(400.0, 3.0)), (abi.FX_GAIN, abi.GainParams(0.3))),
        (abi.WAVE_SINE, 660.0, (abi.FX_PEAKING_EQ_12DB, abi.BiquadParams(660.0, 9.0)), (abi.FX_GAIN, abi.GainParams(0.2))),
    ]
    stages, posts = [], []
    for wave, hz, (fk, fp), (pk, pp) in chains:
        o = osc_dev(wave, hz, 0.3)
        f = r.add_effect(fk, fp)
        g = r.add_effect(pk, pp)
        r.patch_chain([o, f, g, abi.MAIN_MIXER])
        stages.append(f)
        posts.append(g)
    o6 = osc_dev(abi.WAVE_SAWTOOTH, 55.0)
    f6 = r.add_effect(abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(2000.0, 1.0))
    g6a = r.add_effect(abi.FX_GAIN, abi.GainParams(0.1))
    g6b = r.add_effect(abi.FX_LIMITER, abi.LimiterParams(0.0, 0.05))
    r.patch_chain([o6, f6, g6a, abi.MAIN_MIXER])
    r.patch_chain([f6, g6b, abi.MAIN_MIXER])
    o7 = osc_dev(abi.WAVE_SQUARE, 98.0)
    f7 = r.add_effect(abi.FX_LOW_SHELF_12DB, abi.BiquadParams(200.0, 5.0))
    g7 = r.add_effect(abi.FX_GAIN, abi.GainParams(0.15))
    r.patch_chain([o7, f7, g7, abi.MAIN_MIXER])
    r.patch(o6, g7)
    r.finalize()
    ev = []
    for k, f in enumerate(range(100, 20000, 1777)):
        ev.append((f, stages[0], abi.EV_SET_PARAM, 0, 0, 500.0 + 150.0 * k))     # lpf24 cutoff (Hz)
        ev.append((f + 64, stages[1], abi.EV_CONTROL, 0, 0, 0.2 + 0.05 * k))     # lpf24 cutoff (control value)
        ev.append((f + 128, stages[3], abi.EV_CONTROL, 0, 0, 0.3 + 0.03 * k))    # band-pass cutoff
        ev.append((f + 200, posts[0], abi.EV_CONTROL, 0, 0, 0.9 - 0.06 * k))     # fused gain's ceiling
        ev.append((f + 300, posts[3], abi.EV_CONTROL, 0, 0, (4 + k % 9) / 16.0))  # fused bitcrusher's bits
        ev.append((f + 400, g6b, abi.EV_CONTROL, 1, 0, 0.02 + 0.01 * k))         # unfused limiter's max
    r.push_events(sorted(ev, key=lambda e: e[0]))
    return 21000


ALL_SCENES = {
    "bare_sources": scene_bare_sources,
    "fx_chains": scene_fx_chains,
    "cello_chord": scene_cello_chord,
    "cello_held_chord": scene_cello_held_chord,
    "cello_ensemble": scene_cello_ensemble,
    "welsh_variants": scene_welsh_variants,
    "welsh_sustain": scene_welsh_sustain,
    "fm": scene_fm,
    "drums_and_sampler": scene_drums_and_sampler,
    "effects_rack": scene_effects_rack,
    "graph_toys": scene_graph_toys,
    "sidechain": scene_sidechain,
}
