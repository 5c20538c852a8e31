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


def cello_params(voices: int, gain: float, pan: float) -> abi.WelshParams:
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
    p.filter_envelope = abi.env(0.0, 3.29, 0.78, 3.29)
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

    @property
    def voice_samples(self) -> int:
        return self.total_voices * self.frames


def build_cfg4(r: abi.Renderer, cfg: Cfg4) -> int:
    """Build config 4 (or a truncated slice of it) on ``r``; returns the frame count to render."""
    assert cfg.total_voices % cfg.groups == 0
    per = cfg.total_voices // cfg.groups
    uids = []
    for q in range(cfg.groups):
        i0 = cfg.voice_offset + q
        pan = -1.0 + 2.0 * (i0 % 64) / 63.0
        u = r.add_instrument(abi.INST_WELSH, cello_params(per, 1.0 / 4096.0, pan))
        r.patch(u, abi.MAIN_MIXER)
        uids.append(u)
    r.finalize()
    ev = np.zeros(2 * cfg.total_voices, dtype=abi.EVENT_DTYPE)
    k = 0
    for j in range(per):
        for q in range(cfg.groups):
            i = cfg.voice_offset + q + cfg.groups * j
            on = 64 * (i % 128)
            off = cfg.note_off_base + 64 * (i % 128)
            key = 36 + (i % 49)
            ev[k] = (on, uids[q], abi.EV_NOTE_ON, key, 127, 0.0)
            ev[k + 1] = (off, uids[q], abi.EV_NOTE_OFF, key, 0, 0.0)
            k += 2
    ev = ev[np.argsort(ev["frame"], kind="stable")]
    r.push_events(ev)
    return cfg.frames


def cfg4_slice(voices: int, frames: int, voice_offset: int = 0) -> Cfg4:
    """The first ``voices`` voices / first ``frames`` frames of config 4 (CPU-baseline samples, tests)."""
    groups = min(128, voices)
    while voices % groups:
        groups -= 1
    return Cfg4(total_voices=voices, frames=frames, groups=groups, voice_offset=voice_offset)
