"""Synthetic workloads of BASELINE.json, built through the block-render ABI.

Config 4: 4096 Welsh-cookbook `cello` voices, SR 48 kHz, 60 s stereo.  Voice i:
key 36 + (i mod 49), pan -1 + 2*(i mod 64)/63, gain 1/4096, note-on at frame 64*(i mod 128),
note-off at 2,400,000 + 64*(i mod 128).  Voices are grouped into 128 instruments of 32 voices
(instrument q holds voices i = q + 128*j), which keeps every key of an instrument distinct, so each
note-on allocates its own voice exactly as the recipe intends.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import abi

LOG800 = math.log(800.0)

# Algorithmic work per Welsh voice-sample: SURVEY.md §8(d) / BASELINE.md §3 counting rule.
W_VOICE_FLOP = 150.0
W_FM_FLOP = 67.0


def hz_to_pct(hz: float) -> float:
    return max(0.0, min(1.0, math.log(hz / 25.0) / LOG800))


def cello_params(voices: int, gain: float, pan: float, filter_decay: float = 3.29) -> abi.WelshParams:
    """assets/patches/welsh/cello.json mapped as settings/src/patches.rs:87-170 does (release := decay)."""
    p = abi.WelshParams()
    p.oscillator_1 = abi.osc(abi.WAVE_PULSE_WIDTH, 0.1)
    p.oscillator_2 = abi.osc(abi.WAVE_SQUARE)
    p.oscillator_2_sync = 0
    p.oscillator_mix = 0.5
    p.amp_envelope = abi.env(0.06, 0.0, 1.0, 0.0)
    p.lfo = abi.osc(abi.WAVE_SINE, frequency=7.5)
    p.lfo_routing = abi.LFO_AMPLITUDE
    p.lfo_depth = 0.05
    p.filter_cutoff_hz = 40.0
    p.filter_passband_ripple = 0.707
    p.filter_cutoff_start = hz_to_pct(40.0)
    p.filter_cutoff_end = 0.9
    p.filter_envelope = abi.env(0.0, filter_decay, 0.78, filter_decay)
    p.voice_dca = abi.DcaParams(1.0, 0.0)
    p.dca = abi.DcaParams(gain, pan)
    p.voices = voices
    return p


def piano_params(voices: int, gain: float, pan: float) -> abi.WelshParams:
    """assets/patches/welsh/piano.json through derive_welsh_synth_params (settings/src/patches.rs:87-170):
    sawtooth + 15 % pulse an octave and two semitones up, HARD SYNC, no LFO, amplitude A0/D0.67/S0.25,
    filter envelope A0/D5.22/S0 with weight 0.75 from 40 Hz (release := decay).  bench.py's `general` leg."""
    p = abi.WelshParams()
    p.oscillator_1 = abi.osc(abi.WAVE_SAWTOOTH)
    p.oscillator_2 = abi.osc(abi.WAVE_PULSE_WIDTH, 0.15, tune=2.244924096618746)
    p.oscillator_2_sync = 1
    p.oscillator_mix = 3.0 / 7.0
    p.amp_envelope = abi.env(0.0, 0.67, 0.25, 0.67)
    p.lfo = abi.osc(abi.WAVE_NONE, frequency=0.0)
    p.lfo_routing = abi.LFO_NONE
    p.lfo_depth = 0.0
    p.filter_cutoff_hz = 40.0
    p.filter_passband_ripple = 0.707
    p.filter_cutoff_start = hz_to_pct(40.0)
    p.filter_cutoff_end = 0.75
    p.filter_envelope = abi.env(0.0, 5.22, 0.0, 5.22)
    p.voice_dca = abi.DcaParams(1.0, 0.0)
    p.dca = abi.DcaParams(gain, pan)
    p.voices = voices
    return p


@dataclass
class Cfg4:
    sample_rate: float = 48000.0
    total_voices: int = 4096
    frames: int = 2_880_000
    note_off_base: int = 2_400_000
    groups: int = 128          # instruments; voice i belongs to instrument i mod groups
    voice_offset: int = 0      # first global voice index (multi-GPU shards use rank * total_voices)
    group_first: int = 0       # strong-scaling shards: this slice holds instruments group_first .. group_first + groups - 1
    group_total: int = 0       # ... of an ensemble of group_total instruments (0 = groups: the whole ensemble)
    filter_decay: float = 3.29 # cello.json's filter-envelope decay (s); bench.py's "time_varying" leg stretches it
                               # past the note length so that the cutoff never rests

    @property
    def voice_samples(self) -> int:
        return self.total_voices * self.frames


def build_cfg4(r: abi.Renderer, cfg: Cfg4, params=None) -> int:
    """Build config 4 (or a truncated slice of it) on ``r``; returns the frame count to render."""
    uids = build_cfg4_graph(r, cfg, params)
    r.push_events(cfg4_events(cfg, uids))
    return cfg.frames


def build_cfg4_graph(r: abi.Renderer, cfg: Cfg4, params=None):
    """The instruments and patch cables of config 4, finalized; returns the instrument uids.
    `params(voices, gain, pan)` replaces the cello recipe (bench.py's `general` leg: another patch)."""
    assert cfg.total_voices % cfg.groups == 0
    per = cfg.total_voices // cfg.groups
    uids = []
    for q in range(cfg.groups):
        i0 = cfg.voice_offset + cfg.group_first + q
        pan = -1.0 + 2.0 * (i0 % 64) / 63.0
        p = params(per, 1.0 / 4096.0, pan) if params else cello_params(per, 1.0 / 4096.0, pan, cfg.filter_decay)
        u = r.add_instrument(abi.INST_WELSH, p)
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    return uids


def cfg4_events(cfg: Cfg4, uids) -> np.ndarray:
    """Config 4's note events (frame-sorted) for the instruments `uids` of build_cfg4_graph."""
    per = cfg.total_voices // cfg.groups
    stride = cfg.group_total or cfg.groups
    ev = np.zeros(2 * cfg.total_voices, dtype=abi.EVENT_DTYPE)
    k = 0
    for j in range(per):
        for q in range(cfg.groups):
            i = cfg.voice_offset + cfg.group_first + q + stride * j
            on = 64 * (i % 128)
            off = cfg.note_off_base + 64 * (i % 128)
            key = 36 + (i % 49)
            ev[k] = (on, uids[q], abi.EV_NOTE_ON, key, 127, 0.0)
            ev[k + 1] = (off, uids[q], abi.EV_NOTE_OFF, key, 0, 0.0)
            k += 2
    return ev[np.argsort(ev["frame"], kind="stable")]


def cfg4_slice(voices: int, frames: int, voice_offset: int = 0) -> Cfg4:
    """The first ``voices`` voices / first ``frames`` frames of config 4 (CPU-baseline samples, tests)."""
    groups = min(128, voices)
    while voices % groups:
        groups -= 1
    return Cfg4(total_voices=voices, frames=frames, groups=groups, voice_offset=voice_offset)


# ---------------------------------------------------------------------------------------------------
# Config 5: 65 536 one-shot patch variants, SR 48 kHz, 2 s each, note-off at 1 s.
# Variant j draws from splitmix64 seeded with 0x5EED + j; even j = subtractive, odd j = FM.
CFG5_SAMPLE_RATE = 48000.0
CFG5_FRAMES = 96000
CFG5_NOTE_OFF = 48000
_M64 = (1 << 64) - 1


class _SplitMix:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def u(self) -> float:
        return (self.next() >> 11) * (1.0 / 9007199254740992.0)

    def log_u(self, lo: float, hi: float) -> float:
        return lo * (hi / lo) ** self.u()

    def pick(self, seq):
        return seq[self.next() % len(seq)]


def cfg5_variant(j: int, gain: float = 1.0 / 256.0):
    """(kind, params struct, MIDI key) of variant j."""
    g = _SplitMix(0x5EED + j)
    key = 36 + g.next() % 49
    env4 = lambda: (g.log_u(1e-3, 1.0), g.log_u(1e-3, 1.0), g.u(), g.log_u(1e-3, 1.0))
    pan = 2.0 * g.u() - 1.0
    if j % 2 == 0:
        shapes = (abi.WAVE_SAWTOOTH, abi.WAVE_SQUARE, abi.WAVE_PULSE_WIDTH, abi.WAVE_TRIANGLE)
        p = abi.WelshParams()
        p.oscillator_1 = abi.osc(g.pick(shapes), 0.1 + 0.8 * g.u())
        cents = (g.u() * 100.0 - 50.0) + 1200.0 * g.pick((-2, -1, 0, 1, 2))
        p.oscillator_2 = abi.osc(g.pick(shapes), 0.1 + 0.8 * g.u(), tune=2.0 ** (cents / 1200.0))
        p.oscillator_2_sync = 0
        p.oscillator_mix = g.u()
        p.amp_envelope = abi.env(*env4())
        p.lfo = abi.osc(abi.WAVE_SINE, frequency=0.0)
        p.lfo_routing = abi.LFO_NONE
        p.lfo_depth = 0.0
        cutoff = g.log_u(40.0, 8000.0)
        p.filter_cutoff_hz = cutoff
        res = g.u()
        p.filter_passband_ripple = res * res * 10.0 + 0.707
        p.filter_cutoff_start = hz_to_pct(cutoff)
        p.filter_cutoff_end = g.u()
        p.filter_envelope = abi.env(*env4())
        p.voice_dca = abi.DcaParams(1.0, 0.0)
        p.dca = abi.DcaParams(gain, pan)
        p.voices = 1
        return abi.INST_WELSH, p, key
    p = abi.FmParams()
    p.ratio = g.pick((0.5, 1.0, 2.0, 3.0, 4.0))
    p.beta = g.log_u(0.1, 20.0)
    p.depth = g.u()
    p.carrier_envelope = abi.env(*env4())
    p.modulator_envelope = abi.env(*env4())
    p.dca = abi.DcaParams(gain, pan)
    p.voices = 1
    return abi.INST_FM, p, key


def build_cfg5(r: abi.Renderer, n_variants: int, first: int = 0, frames: int = CFG5_FRAMES,
               note_off: int = CFG5_NOTE_OFF, variants=None, push: bool = True):
    """Variants first .. first+n-1, each its own one-voice instrument patched into the main mixer
    (its node buffer is the per-variant stereo output; the mixer's is the summed bus).
    `variants` = precomputed [cfg5_variant(j)] (bench.py draws them once); push=False returns the events
    instead of pushing them (bench.py pushes them inside its end-to-end span)."""
    uids = []
    ev = np.zeros(2 * n_variants, dtype=abi.EVENT_DTYPE)
    for i in range(n_variants):
        kind, p, key = variants[i] if variants is not None else cfg5_variant(first + i)
        u = r.add_instrument(kind, p)
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
        ev[2 * i] = (0, u, abi.EV_NOTE_ON, key, 127, 0.0)
        ev[2 * i + 1] = (note_off, u, abi.EV_NOTE_OFF, key, 0, 0.0)
    r.finalize()
    ev = ev[np.argsort(ev["frame"], kind="stable")]
    if not push:
        return frames, uids, ev
    r.push_events(ev)
    return frames, uids


# ---------------------------------------------------------------------------------------------------
# Stream-effect batch: N copies of BASELINE config 1's effect chain (a source through a 24 dB low-pass whose
# cutoff a control trip moves, projects/demos/effects/drums-filtered-24db.json) plus a gain, each chain its
# own patch-cable path into the main mixer.  The sources are bare oscillator devices (resident buffers after
# one cheap launch each); what the leg measures is the effect side: HBM-bound, 16 B read + 16 B written per
# frame per chain.
def build_fx_chains(r: abi.Renderer, n_chains: int, frames: int = 1 << 16, step: int = 4096) -> np.ndarray:
    """Returns the (frame-sorted) cutoff automation events; the graph is finalized."""
    shapes = (abi.WAVE_SAWTOOTH, abi.WAVE_SQUARE, abi.WAVE_TRIANGLE, abi.WAVE_PULSE_WIDTH)
    filters = []
    for i in range(n_chains):
        o = r.add_instrument(abi.INST_OSCILLATOR,
                             abi.OscillatorSourceParams(abi.osc(shapes[i % 4], 0.25, frequency=55.0 * (1 + i % 16))))
        f = r.add_effect(abi.FX_LOW_PASS_24DB, abi.Lowpass24Params(300.0 + 7.0 * (i % 256), 0.8))
        g = r.add_effect(abi.FX_GAIN, abi.GainParams(1.0 / n_chains))
        r.patch_chain([o, f, g, abi.MAIN_MIXER])
        filters.append(f)
    r.finalize()
    points = list(range(step, frames, step))
    ev = np.zeros(len(points) * n_chains, dtype=abi.EVENT_DTYPE)
    k = 0
    for j, t in enumerate(points):          # the trip: cutoff rising, one control step per `step` frames
        for i, f in enumerate(filters):
            ev[k] = (t, f, abi.EV_CONTROL, 0, 0, 0.3 + 0.6 * (j + 1) / (len(points) + 1))
            k += 1
    return ev
