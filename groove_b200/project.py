"""Project loader + event compiler (host side).

Turns a Groove project (JSON / JSON5, `settings/src/songs.rs:17-56`) plus the Welsh patch files it
names into a *compiled plan*: entity list, patch cables and a frame-stamped event list, and builds
that plan on any block-render ABI engine.  Semantics kept from the reference:

  * devices / patch cables / tracks / trips           settings/src/songs.rs:91-306
  * Welsh patch -> voice params, incl. release:=decay settings/src/patches.rs:87-170
  * pattern note = velocity 127, one step long, 0 = rest  settings/src/lib.rs:55-77
  * controllers run once per 64-frame buffer, before the buffer's audio
                                                       orchestration/src/orchestrator.rs:631-708,856-877
  * render length ceil(beats * 60 / bpm * SR)          orchestration/src/orchestrator.rs:1723-1737

Schema drift across the reference's fixtures is accepted: `[4,4]` or `{top,bottom}`
time signatures, `flat: [v]` or `flat: {value: v}`, `min/max` or `minimum/maximum`, `bits` or
`bits-to-crush`, `delay` or `seconds`.  Event sources compiled: the pattern sequencer, control trips,
the LFO controller (through `controls` links) and the arpeggiator.  The signal-passthrough (sidechain)
controller is not an event source: it becomes a pass-through node in its patch chain plus a control link
(`Plan.links`) that the engine evaluates from the rendered signal (`gb_link_control`).
"""
from __future__ import annotations

import json
import math
import os
import re
import wave
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import abi

BUFFER_FRAMES = 64  # the reference CLI / tests / GUI all tick 64-frame buffers (orchestrator.rs:1696)
LOG800 = math.log(800.0)

NOTE_VALUE_DIVISOR = {  # settings/src/lib.rs:120-135 (value = 4096 * divisor / 4 ...): beats = 4 / divisor
    "octuple": 0.125, "quadruple": 0.25, "double": 0.5, "whole": 1, "half": 2, "quarter": 4, "eighth": 8,
    "sixteenth": 16, "thirty-second": 32, "sixty-fourth": 64, "one-hundred-twenty-eighth": 128,
    "two-hundred-fifty-sixth": 256, "five-hundred-twelfth": 512,
}

# GM percussion key -> 707 sample (doc/general-midi-percussion-numbers.csv names; the reference's own
# table is in the absent Drumkit source: parity unpinned)
KIT_707 = {
    35: "Kick 1 R1", 36: "Kick 2 R1", 37: "Rim R1", 38: "Snare 1 R1", 39: "Clap R1", 40: "Snare 2 R1",
    42: "Hat Closed R1", 44: "Hat Closed R2", 45: "Tom 1 R1", 46: "Hat Open R1", 47: "Tom 1 R2", 48: "Tom 2 R1",
    49: "Crash R1", 50: "Tom 3 R1", 51: "Ride R1", 53: "Cowbell R1", 54: "Tambourine R1", 57: "Crash R2",
    59: "Ride R2", 67: "Cowbell R3", 68: "Cowbell R4",
}

WAVEFORMS = {"none": abi.WAVE_NONE, "sine": abi.WAVE_SINE, "square": abi.WAVE_SQUARE, "triangle": abi.WAVE_TRIANGLE,
             "sawtooth": abi.WAVE_SAWTOOTH, "noise": abi.WAVE_NOISE, "debug-zero": abi.WAVE_DEBUG_ZERO,
             "debug-max": abi.WAVE_DEBUG_MAX, "debug-min": abi.WAVE_DEBUG_MIN}
LFO_ROUTINGS = {"none": abi.LFO_NONE, "amplitude": abi.LFO_AMPLITUDE, "pitch": abi.LFO_PITCH,
                "pulse-width": abi.LFO_PULSE_WIDTH, "filter-cutoff": abi.LFO_FILTER_CUTOFF}

EFFECT_KINDS = {
    "mixer": abi.FX_MIXER, "gain": abi.FX_GAIN, "limiter": abi.FX_LIMITER, "bitcrusher": abi.FX_BITCRUSHER,
    "compressor": abi.FX_COMPRESSOR, "delay": abi.FX_DELAY, "chorus": abi.FX_CHORUS, "reverb": abi.FX_REVERB,
    "filter-low-pass-12db": abi.FX_LOW_PASS_12DB, "filter-high-pass-12db": abi.FX_HIGH_PASS_12DB,
    "filter-band-pass-12db": abi.FX_BAND_PASS_12DB, "filter-band-stop-12db": abi.FX_BAND_STOP_12DB,
    "filter-all-pass-12db": abi.FX_ALL_PASS_12DB, "filter-peaking-eq-12db": abi.FX_PEAKING_EQ_12DB,
    "filter-low-shelf-12db": abi.FX_LOW_SHELF_12DB, "filter-high-shelf-12db": abi.FX_HIGH_SHELF_12DB,
    "filter-low-pass-24db": abi.FX_LOW_PASS_24DB,
}
# control-parameter name -> flattened control index (proc-macros/src/control.rs:126-130,210-226)
CONTROL_INDEX = {
    "ceiling": 0, "min": 0, "minimum": 0, "max": 1, "maximum": 1, "bits": 0, "bits-to-crush": 0, "threshold": 0,
    "ratio": 1, "cutoff": 0, "q": 1, "bandwidth": 1, "db-gain": 1, "passband-ripple": 1, "wet-dry-mix": 2,
    "attenuation": 0, "gain": 0, "dca-gain": 0, "pan": 1, "dca-pan": 1,
}


UNITS_IN_BEAT = 65536  # MusicalTime resolution (doc/designs/time.md:94-98)


def frames_to_units(frames: int, bpm: float, sample_rate: int) -> int:
    """MusicalTime::new_with_frames (src/mini/transport.rs:58-63): whole MusicalTime units elapsed after
    `frames` frames.  Integer arithmetic, so the per-frame deltas telescope exactly: one second of frames
    at 60 bpm is exactly UNITS_IN_BEAT units at any sample rate (transport.rs:157-188)."""
    bpm_milli = int(round(bpm * 1000.0))
    return (frames * bpm_milli * UNITS_IN_BEAT) // (60 * 1000 * int(sample_rate))


def song_frames(beats: float, bpm: float, sample_rate: float) -> int:
    """orchestrator.rs:1723-1737: a song of `beats` beats renders ceil(beats*60/bpm*SR) frames."""
    return int(math.ceil(beats * 60.0 / bpm * sample_rate - 1e-9))


def parse_json5(text: str):
    """Enough of JSON5 for the reference's fixtures: comments, trailing commas, bare keys."""
    try:
        return json.loads(text)
    except json.JSONDecodeError:
        pass
    out, i, n, in_str, q = [], 0, len(text), False, ""
    while i < n:
        c = text[i]
        if in_str:
            if c == "\\" and i + 1 < n:
                out.append(c)
                out.append(text[i + 1])
                i += 1
            elif c == q:
                out.append('"')
                in_str = False
            else:
                out.append('\\"' if c == '"' else c)
        elif c in "\"'":
            in_str, q = True, c
            out.append('"')
        elif text.startswith("//", i):
            while i < n and text[i] != "\n":
                i += 1
            continue
        elif text.startswith("/*", i):
            i = text.find("*/", i) + 2
            continue
        else:
            out.append(c)
        i += 1
    s = "".join(out)
    s = re.sub(r",(\s*[\]}])", r"\1", s)
    s = re.sub(r'([{,]\s*)([A-Za-z_][A-Za-z0-9_\-]*)(\s*:)', r'\1"\2"\3', s)
    return json.loads(s)


def read_wav(path: str) -> Tuple[np.ndarray, float]:
    """PCM WAV (8/16/24/32-bit) -> float64 in [-1, 1), shape (n,) or (n, 2)."""
    with wave.open(path, "rb") as w:
        ch, width, sr, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float64) / 32768.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float64) / 8388608.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0
    else:
        raise ValueError(f"unsupported sample width {width}")
    if ch == 2:
        x = x.reshape(-1, 2)
    elif ch != 1:
        x = x.reshape(-1, ch)[:, 0].copy()
    return x, float(sr)


def wav_root_note(path: str) -> Optional[int]:
    """MIDI root note from a WAV file's metadata, or None: the `smpl` chunk's MIDI unity note, else the
    `acid` chunk's root note when its flags say it is valid (bit 1).  README.md:82-84: "If it can figure
    out the root frequency from the WAV file's metadata, then it will play the sample at the right
    adjusted frequency"; test-data/samples/riff-acidized.wav carries an acid chunk with root note 57."""
    import struct
    with open(path, "rb") as f:
        b = f.read()
    if len(b) < 12 or b[:4] != b"RIFF" or b[8:12] != b"WAVE":
        return None
    smpl = acid = None
    i = 12
    while i + 8 <= len(b):
        cid, size = b[i:i + 4], struct.unpack("<I", b[i + 4:i + 8])[0]
        body = b[i + 8:i + 8 + size]
        if cid == b"smpl" and len(body) >= 16:
            note = struct.unpack("<I", body[12:16])[0]
            if 0 <= note < 128:
                smpl = int(note)
        elif cid == b"acid" and len(body) >= 6:
            flags, note = struct.unpack("<IH", body[:6])
            if flags & 0x02 and note < 128:
                acid = int(note)
        i += 8 + size + (size & 1)
    return smpl if smpl is not None else acid


def write_wav16(path: str, pcm: np.ndarray, sample_rate: float) -> None:
    """16-bit stereo PCM writer (orchestration/src/helpers.rs:74-97)."""
    with wave.open(path, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(int(sample_rate))
        w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())


# ------------------------------------------------------------------------------------------------
@dataclass
class Entity:
    uvid: str
    role: str                 # "instrument" | "effect"
    kind: int
    params: Dict[str, object] = field(default_factory=dict)   # plain numbers, JSON-serialisable
    midi_in: Optional[int] = None
    samples: List[Tuple[int, str, float]] = field(default_factory=list)  # (key, sample name, root_hz)


@dataclass
class Plan:
    title: str
    sample_rate: float
    bpm: float
    frames: int
    entities: List[Entity]
    cables: List[List[str]]
    events: List[Tuple[int, str, int, int, int, float]]   # (frame, uvid, type, a, b, value)
    skipped: List[str] = field(default_factory=list)
    links: List[Tuple[str, str, int]] = field(default_factory=list)   # (source uvid, target uvid, control index)

    def to_json(self) -> str:
        return json.dumps({
            "title": self.title, "sample_rate": self.sample_rate, "bpm": self.bpm, "frames": self.frames,
            "entities": [e.__dict__ for e in self.entities], "cables": self.cables, "events": self.events,
            "skipped": self.skipped, "links": self.links}, indent=None, separators=(",", ":"))

    @staticmethod
    def from_json(text: str) -> "Plan":
        d = json.loads(text)
        ents = [Entity(uvid=e["uvid"], role=e["role"], kind=e["kind"], params=e["params"], midi_in=e["midi_in"],
                       samples=[tuple(s) for s in e["samples"]]) for e in d["entities"]]
        return Plan(d["title"], d["sample_rate"], d["bpm"], d["frames"], ents, d["cables"],
                    [tuple(e) for e in d["events"]], d.get("skipped", []), [tuple(l) for l in d.get("links", [])])


def _osc_tune(tune) -> Tuple[float, Optional[int]]:
    """OscillatorTune -> (ratio, fixed note) — settings/src/patches.rs:211-234."""
    if isinstance(tune, dict):
        if "float" in tune:
            return float(tune["float"]), None
        if "note" in tune:
            return 1.0, int(tune["note"])
        if "osc" in tune:
            o = tune["osc"]
            semis = int(o.get("octave", 0)) * 12 + int(o.get("semi", 0))
            return 2.0 ** ((semis * 100.0 + float(o.get("cent", 0))) / 1200.0), None
    return 1.0, None


def _waveform(w) -> Tuple[int, float]:
    if isinstance(w, dict):
        if "pulse-width" in w:
            return abi.WAVE_PULSE_WIDTH, float(w["pulse-width"])
        raise ValueError(f"unknown waveform {w}")
    if w not in WAVEFORMS:  # e.g. "triangle-sine": a TODO in the reference (settings/src/patches.rs:188)
        raise ValueError(f"waveform {w!r} is not implemented")
    return WAVEFORMS[w], 0.5


def welsh_params_from_patch(patch: dict, voices: int = 8) -> dict:
    """WelshPatchSettings::derive_welsh_synth_params (settings/src/patches.rs:87-170) as plain numbers."""
    o1, o2 = patch["oscillator-1"], patch["oscillator-2"]
    w1, pw1 = _waveform(o1["waveform"])
    w2, pw2 = _waveform(o2["waveform"])
    t1, _ = _osc_tune(o1.get("tune", {}))
    t2, note2 = _osc_tune(o2.get("tune", {}))
    # oscillator-2-track=false: the reference sets the fixed frequency only on a throwaway Oscillator that
    # is used for the mix count (patches.rs:92-101); the WelshVoiceParams it returns carry
    # oscillator_2 = {waveform, frequency_tune (Note -> ratio 1.0), ..Default} (patches.rs:118-122), so
    # oscillator 2 still tracks the note at ratio 1.0.  Reproduced as is; the panic is kept.
    fixed2 = 0.0
    if w2 != abi.WAVE_NONE and not patch.get("oscillator-2-track", True) and note2 is None:
        raise ValueError("Patch configured without oscillator 2 tracking, but tune is not a note specification")
    n_osc = (w1 != abi.WAVE_NONE) + (w2 != abi.WAVE_NONE) + (float(patch.get("noise", 0.0)) > 0.0)
    m1, m2 = float(o1.get("mix-pct", 1.0)), float(o2.get("mix-pct", 1.0))
    if n_osc == 0:
        mix = 0.0
    elif n_osc == 1 or (m1 == 0.0 and m2 == 0.0):
        mix = 1.0
    else:
        mix = m1 / (m1 + m2)
    lfo = patch.get("lfo", {})
    routing = lfo.get("routing", "none")
    if routing not in LFO_ROUTINGS:
        raise ValueError(f"LFO routing {routing!r} does not deserialise (patches.rs:269-278)")
    depth = lfo.get("depth", {"pct": 0.0})
    if depth == "none":
        d = 0.0
    elif not isinstance(depth, dict):  # LfoDepth is an enum (patches.rs:295-303): a bare number does not deserialise
        raise ValueError(f"LFO depth {depth!r} does not deserialise")
    elif "pct" in depth:
        d = float(depth["pct"])
    else:  # cents: Normal::new(1 - 2^(c/1200)) — negative for c > 0, clamped by Normal (patches.rs:304-314)
        d = 1.0 - 2.0 ** (float(depth["cents"]) / 1200.0)
    d = min(1.0, max(0.0, d))
    lw, lpw = _waveform(lfo.get("waveform", "sine"))
    fe, ae = patch["filter-envelope"], patch["amp-envelope"]
    res = float(patch.get("filter-resonance", 0.0))
    f24 = patch.get("filter-type-24db", {"cutoff-hz": 0.0})
    f12 = patch.get("filter-type-12db", {"cutoff-hz": 0.0})
    hz12 = float(f12.get("cutoff-hz", 0.0))
    return {
        "w1": w1, "pw1": pw1, "tune1": t1, "w2": w2, "pw2": pw2, "tune2": t2, "fixed2": fixed2,
        "sync": int(bool(patch.get("oscillator-2-sync", False))), "mix": mix,
        "amp": [float(ae["attack"]), float(ae["decay"]), float(ae["sustain"]), float(ae["decay"])],   # release := decay
        "lfo_wave": lw, "lfo_pw": lpw, "lfo_hz": float(lfo.get("frequency", 0.0)), "lfo_routing": LFO_ROUTINGS[routing],
        "lfo_depth": d,
        "cutoff_hz": float(f24.get("cutoff-hz", 0.0)), "ripple": res * res * 10.0 + 0.707,            # denormalize_q
        "cutoff_start": max(0.0, min(1.0, math.log(max(hz12, 1e-9) / 25.0) / LOG800)) if hz12 > 0 else 0.0,
        "cutoff_end": float(patch.get("filter-envelope-weight", 0.0)),
        "filt": [float(fe["attack"]), float(fe["decay"]), float(fe["sustain"]), float(fe["decay"])],  # release := decay
        "gain": 1.0, "pan": 0.0, "voices": voices,
    }


def welsh_struct(p: dict) -> abi.WelshParams:
    s = abi.WelshParams()
    s.oscillator_1 = abi.osc(p["w1"], p["pw1"], tune=p["tune1"])
    s.oscillator_2 = abi.osc(p["w2"], p["pw2"], tune=p["tune2"], fixed_frequency=p["fixed2"])
    s.oscillator_2_sync = p["sync"]
    s.oscillator_mix = p["mix"]
    s.amp_envelope = abi.env(*p["amp"])
    s.lfo = abi.osc(p["lfo_wave"], p["lfo_pw"], frequency=p["lfo_hz"])
    s.lfo_routing = p["lfo_routing"]
    s.lfo_depth = p["lfo_depth"]
    s.filter_cutoff_hz = p["cutoff_hz"]
    s.filter_passband_ripple = p["ripple"]
    s.filter_cutoff_start = p["cutoff_start"]
    s.filter_cutoff_end = p["cutoff_end"]
    s.filter_envelope = abi.env(*p["filt"])
    s.voice_dca = abi.DcaParams(1.0, 0.0)
    s.dca = abi.DcaParams(p["gain"], p["pan"])
    s.voices = p["voices"]
    return s


def _env4(e: Optional[dict], default=(0.0, 0.0, 1.0, 0.0)) -> List[float]:
    if not e:
        return list(default)
    return [float(e.get("attack", 0.0)), float(e.get("decay", 0.0)), float(e.get("sustain", 1.0)), float(e.get("release", 0.0))]


def effect_struct(kind: int, p: dict):
    g = lambda *names, default=0.0: next((float(p[n]) for n in names if n in p), default)
    if kind in (abi.FX_MIXER, abi.FX_SIGNAL_PASSTHROUGH):
        return None
    if kind == abi.FX_GAIN:
        return abi.GainParams(g("ceiling", default=1.0))
    if kind == abi.FX_LIMITER:
        return abi.LimiterParams(g("min", "minimum", default=0.0), g("max", "maximum", default=1.0))
    if kind == abi.FX_BITCRUSHER:
        return abi.BitcrusherParams(g("bits", "bits-to-crush", default=8.0))
    if kind == abi.FX_COMPRESSOR:
        return abi.CompressorParams(g("threshold", default=1.0), g("ratio", default=1.0), g("attack"), g("release"))
    if kind == abi.FX_DELAY:
        return abi.DelayParams(g("seconds", "delay", default=0.0))
    if kind == abi.FX_CHORUS:
        return abi.ChorusParams(g("voices", default=1.0), g("delay-seconds", "delay-factor", default=0.0),
                                g("wet-dry-mix", default=1.0))
    if kind == abi.FX_REVERB:
        return abi.ReverbParams(g("attenuation", default=1.0), g("seconds", default=1.0))
    if kind == abi.FX_LOW_PASS_24DB:
        return abi.Lowpass24Params(g("cutoff", default=1000.0), g("passband-ripple", default=0.707))
    if kind in abi.BIQUAD_KINDS:
        return abi.BiquadParams(g("cutoff", default=1000.0), g("q", "bandwidth", "db-gain", default=0.707))
    raise ValueError(kind)


class ProjectLoader:
    def __init__(self, assets_dir: str):
        """assets_dir holds `patches/welsh/*.json` and `samples/elphnt.io/707/*.wav` (reference layout)."""
        self.assets = assets_dir

    # ---- devices -----------------------------------------------------------------------------
    def _instrument(self, uvid: str, spec: dict, plan: Plan) -> Optional[Entity]:
        (kind_name, body), = spec.items()
        midi = body[0].get("midi-in") if isinstance(body, list) and body else None
        args = body[1] if isinstance(body, list) and len(body) > 1 else {}
        if kind_name == "welsh":
            name = re.sub(r"(?<!^)(?=[A-Z])", "-", args["name"]).lower()
            with open(os.path.join(self.assets, "patches", "welsh", name + ".json")) as f:
                patch = json.load(f)
            return Entity(uvid, "instrument", abi.INST_WELSH, welsh_params_from_patch(patch), midi)
        if kind_name == "drumkit":
            e = Entity(uvid, "instrument", abi.INST_DRUMKIT, {"name": args.get("name", "707")}, midi)
            e.samples = [(k, v, 0.0) for k, v in sorted(KIT_707.items())]
            return e
        if kind_name == "sampler":
            # older fixtures (projects/tests/load-stereo-wav.json) keep midi-in and the parameters in ONE object
            if not args and isinstance(body, list) and body and "filename" in body[0]:
                args = body[0]
            # root: the project's value if positive, else the WAV file's own metadata (README.md:82-84), else the
            # engine default (440 Hz)
            root = float(args.get("root", 0.0))
            if not root > 0.0:
                root = self.sample_root_hz(args["filename"])
            e = Entity(uvid, "instrument", abi.INST_SAMPLER, {"root": root, "voices": 8}, midi)
            e.samples = [(0, args["filename"], root)]
            return e
        if kind_name == "fm-synthesizer":
            dca = args.get("dca", {})
            return Entity(uvid, "instrument", abi.INST_FM, {
                "ratio": float(args.get("ratio", 2.0)), "depth": float(args.get("depth", 1.0)),
                "beta": float(args.get("beta", 1.0)), "car": _env4(args.get("carrier-envelope")),
                "mod": _env4(args.get("modulator-envelope")), "gain": float(dca.get("gain", 1.0)),
                "pan": float(dca.get("pan", 0.0)), "voices": 8}, midi)
        # "oscillator" / "envelope": bare Oscillator / Envelope devices of the older fixtures (58 + 1 projects under
        # projects/demos/); the variants are gone from InstrumentSettings (settings/src/instruments.rs:24-39) but "a
        # loader must accept the union" (SURVEY.md §5).  The body is a one-element list: midi-in and the parameters
        # share one object.
        flat = body[0] if isinstance(body, list) and body else {}
        if kind_name == "oscillator":
            wf, pw = _waveform(flat.get("waveform", "sine"))
            return Entity(uvid, "instrument", abi.INST_OSCILLATOR,
                          {"waveform": wf, "pw": pw, "hz": float(flat.get("frequency", 0.0))}, midi)
        if kind_name == "envelope":
            return Entity(uvid, "instrument", abi.INST_ENVELOPE, {"env": _env4(flat)}, midi)
        plan.skipped.append(f"instrument {uvid}: type {kind_name!r} is not in InstrumentSettings (settings/src/instruments.rs:24-39)")
        return None

    def _effect(self, uvid: str, spec: dict, plan: Plan) -> Optional[Entity]:
        (kind_name, body), = spec.items()
        if kind_name not in EFFECT_KINDS:
            plan.skipped.append(f"effect {uvid}: unknown type {kind_name!r}")
            return None
        return Entity(uvid, "effect", EFFECT_KINDS[kind_name], dict(body or {}))

    # ---- compile -------------------------------------------------------------------------------
    def compile(self, project: dict, sample_rate: float = 44100.0) -> Plan:
        clock = project.get("clock", {})
        bpm = float(clock.get("bpm", 128.0))
        ts = clock.get("time-signature", [4, 4])
        top, bottom = (ts["top"], ts["bottom"]) if isinstance(ts, dict) else (ts[0], ts[1])
        plan = Plan(project.get("title") or "", sample_rate, bpm, 0, [], [], [])
        by_uvid: Dict[str, Entity] = {}
        controllers: Dict[str, tuple] = {}
        for dev in project.get("devices", []):
            (role, (uvid, spec)), = dev.items()
            ent = None
            if role == "instrument":
                ent = self._instrument(uvid, spec, plan)
            elif role == "effect":
                ent = self._effect(uvid, spec, plan)
            else:
                (ckind, cbody), = spec.items()
                midi = cbody[0] if isinstance(cbody, list) and cbody else {}
                args = cbody[1] if isinstance(cbody, list) and len(cbody) > 1 else {}
                if ckind in ("lfo", "arpeggiator"):
                    controllers[uvid] = (ckind, midi, args)
                elif ckind == "signal-passthrough-controller":
                    # patched into a chain like an effect (settings/src/controllers.rs:110-111,181-187); the
                    # engine evaluates its control links from the rendered signal (gb_link_control)
                    controllers[uvid] = (ckind, midi, args)
                    ent = Entity(uvid, "effect", abi.FX_SIGNAL_PASSTHROUGH, {})
                else:
                    plan.skipped.append(f"controller {uvid}: {ckind} is not compiled")
            if ent:
                plan.entities.append(ent)
                by_uvid[uvid] = ent
        for cable in project.get("patch-cables", []):
            if len(cable) >= 2:
                plan.cables.append([c for c in cable])
        frames_per_beat = 60.0 / bpm * sample_rate

        def quantise(beat: float) -> int:
            f = beat * frames_per_beat
            return int(math.floor(f / BUFFER_FRAMES + 1e-9)) * BUFFER_FRAMES

        # tracks: patterns laid end to end per track, each rounded up to whole measures
        patterns = {p["id"]: p for p in project.get("patterns", [])}
        end_beats = 0.0
        by_channel: Dict[int, List[Entity]] = {}
        for e in plan.entities:
            if e.role == "instrument" and e.midi_in is not None:
                by_channel.setdefault(int(e.midi_in), []).append(e)
        arps = {int(m.get("midi-in", -1)): int(m.get("midi-out", 0))
                for (k, m, _a) in controllers.values() if k == "arpeggiator"}
        for track in project.get("tracks", []):
            cursor = 0.0
            for pid in track.get("patterns", []):
                pat = patterns.get(pid)
                if pat is None:
                    continue
                divisor = NOTE_VALUE_DIVISOR[pat["note-value"]] if pat.get("note-value") else bottom
                step = bottom / divisor  # beats per pattern note in this time signature
                longest = 0
                for row in pat.get("notes", []):
                    longest = max(longest, len(row))
                    for i, key in enumerate(row):
                        if key == 0:
                            continue
                        ch = int(track["midi-channel"])
                        for ent in by_channel.get(ch, []):
                            plan.events.append((quantise(cursor + i * step), ent.uvid, abi.EV_NOTE_ON, int(key), 127, 0.0))
                            plan.events.append((quantise(cursor + (i + 1) * step), ent.uvid, abi.EV_NOTE_OFF, int(key), 0, 0.0))
                        if ch in arps:
                            # Arpeggiator (parity unpinned): while the input note is held it plays the major
                            # scale key + [0,2,4,5,7,9,11,12], one note per quarter beat, each 0.2 beat long,
                            # repeating every 2 beats, on its midi-out channel.
                            b0, b1 = cursor + i * step, cursor + (i + 1) * step
                            k = 0
                            while b0 + 0.25 * k < b1 - 1e-9:
                                nk = int(key) + (0, 2, 4, 5, 7, 9, 11, 12)[k % 8]
                                for ent in by_channel.get(arps[ch], []):
                                    plan.events.append((quantise(b0 + 0.25 * k), ent.uvid, abi.EV_NOTE_ON, nk, 127, 0.0))
                                    plan.events.append((quantise(b0 + 0.25 * k + 0.2), ent.uvid, abi.EV_NOTE_OFF, nk, 0, 0.0))
                                k += 1
                cursor += math.ceil(longest * step / top - 1e-9) * top
            end_beats = max(end_beats, cursor)
        # control trips: one control event per 64-frame buffer while a step is active
        paths = {p["id"]: p for p in project.get("paths", [])}
        for trip in project.get("trips", []):
            tgt = trip.get("target", {})
            ent = by_uvid.get(tgt.get("id"))
            idx = CONTROL_INDEX.get(tgt.get("param"))
            if ent is None or idx is None:
                plan.skipped.append(f"trip {trip.get('id')}: target {tgt} not controllable")
                continue
            cursor = 0.0
            for pid in trip.get("paths", []):
                path = paths.get(pid)
                if path is None:
                    continue
                divisor = NOTE_VALUE_DIVISOR[path["note-value"]] if path.get("note-value") else bottom
                step_beats = bottom / divisor
                for st in path.get("steps", []):
                    (shape, body), = st.items() if isinstance(st, dict) else (("flat", st),)
                    f0, f1 = quantise(cursor), quantise(cursor + step_beats)
                    if shape == "flat":
                        v = body[0] if isinstance(body, list) else body.get("value", 0.0)
                        plan.events.append((f0, ent.uvid, abi.EV_CONTROL, idx, 0, float(v)))
                    elif shape in ("slope", "logarithmic", "exponential"):
                        a, b = float(body["start"]), float(body["end"])
                        span = max((cursor + step_beats) * frames_per_beat - cursor * frames_per_beat, 1.0)
                        for f in range(f0, f1, BUFFER_FRAMES):
                            t = min(max((f - cursor * frames_per_beat) / span, 0.0), 1.0)
                            if shape == "logarithmic":      # fast first, slow later
                                t = math.log10(1.0 + 9.0 * t)
                            elif shape == "exponential":    # slow first, fast later
                                t = (10.0 ** t - 1.0) / 9.0
                            plan.events.append((f, ent.uvid, abi.EV_CONTROL, idx, 0, a + (b - a) * t))
                    cursor += step_beats
            end_beats = max(end_beats, cursor)
        plan.frames = song_frames(end_beats, bpm, sample_rate)
        # LFO controllers (`controls` links, settings/src/songs.rs:166-202): one control event per 64-frame
        # buffer carrying the oscillator's value at the buffer's first frame, mapped to 0..1
        for link in project.get("controls", []):
            src = controllers.get(link.get("source"))
            tgt = link.get("target", {})
            ent = by_uvid.get(tgt.get("id"))
            idx = CONTROL_INDEX.get(tgt.get("param"))
            if src is not None and src[0] == "signal-passthrough-controller" and ent is not None and idx is not None:
                plan.links.append((link.get("source"), ent.uvid, idx))
                continue
            if src is None or src[0] != "lfo" or ent is None or idx is None:
                plan.skipped.append(f"control {link.get('id')}: source/target not compiled")
                continue
            wf, pw = _waveform(src[2].get("waveform", "sine"))
            hz = float(src[2].get("frequency", 1.0))
            for f in range(0, plan.frames, BUFFER_FRAMES):
                ph = (f * hz / sample_rate) % 1.0
                if wf == abi.WAVE_SINE:
                    v = math.sin(2.0 * math.pi * ph)
                elif wf == abi.WAVE_TRIANGLE:
                    v = 4.0 * ph - 1.0 if ph < 0.5 else 3.0 - 4.0 * ph
                elif wf == abi.WAVE_SAWTOOTH:
                    v = 2.0 * ph if ph < 0.5 else 2.0 * ph - 2.0
                elif wf in (abi.WAVE_SQUARE, abi.WAVE_PULSE_WIDTH):
                    v = 1.0 if ph < (pw if wf == abi.WAVE_PULSE_WIDTH else 0.5) else -1.0
                else:
                    v = 0.0
                plan.events.append((f, ent.uvid, abi.EV_CONTROL, idx, 0, 0.5 * (v + 1.0)))
        plan.events.sort(key=lambda ev: ev[0])
        return plan

    def load(self, path: str, sample_rate: float = 44100.0) -> Plan:
        with open(path) as f:
            return self.compile(parse_json5(f.read()), sample_rate)

    def _sample_path(self, name: str) -> str:
        for cand in (os.path.join(self.assets, "samples", "elphnt.io", "707", name + ".wav"),
                     os.path.join(self.assets, "samples", name), os.path.join(self.assets, name)):
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(name)

    def sample(self, name: str) -> Tuple[np.ndarray, float]:
        return read_wav(self._sample_path(name))

    def sample_root_hz(self, name: str) -> float:
        """Root frequency from the file's metadata (smpl / acid chunk), 0.0 if it has none."""
        try:
            note = wav_root_note(self._sample_path(name))
        except FileNotFoundError:
            return 0.0
        return 440.0 * 2.0 ** ((note - 69) / 12.0) if note is not None else 0.0


def build_plan(r: abi.Renderer, plan: Plan, samples) -> Dict[str, int]:
    """Instantiate a compiled plan on an engine.  `samples(name) -> (float64 array, sample_rate)`."""
    uid: Dict[str, int] = {"main-mixer": abi.MAIN_MIXER}
    for e in plan.entities:
        if e.role == "instrument":
            if e.kind == abi.INST_WELSH:
                uid[e.uvid] = r.add_instrument(e.kind, welsh_struct(e.params))
            elif e.kind == abi.INST_FM:
                p = e.params
                s = abi.FmParams(p["ratio"], p["depth"], p["beta"], abi.env(*p["car"]), abi.env(*p["mod"]),
                                 abi.DcaParams(p["gain"], p["pan"]), p["voices"], 0)
                uid[e.uvid] = r.add_instrument(e.kind, s)
            elif e.kind == abi.INST_OSCILLATOR:
                p = e.params
                uid[e.uvid] = r.add_instrument(e.kind, abi.OscillatorSourceParams(abi.osc(p["waveform"], p["pw"], frequency=p["hz"])))
            elif e.kind == abi.INST_ENVELOPE:
                uid[e.uvid] = r.add_instrument(e.kind, abi.EnvelopeSourceParams(abi.env(*e.params["env"])))
            elif e.kind == abi.INST_DRUMKIT:
                uid[e.uvid] = r.add_instrument(e.kind, abi.DrumkitParams())
                for key, name, _ in e.samples:
                    data, sr = samples(name)
                    r.load_sample(uid[e.uvid], key, data, sr)
            elif e.kind == abi.INST_SAMPLER:
                uid[e.uvid] = r.add_instrument(e.kind, abi.SamplerParams(e.params.get("root", 0.0), e.params.get("voices", 8), 0))
                for key, name, root in e.samples:
                    data, sr = samples(name)
                    r.load_sample(uid[e.uvid], 0, data, sr, root if root > 0 else 440.0)
        else:
            uid[e.uvid] = r.add_effect(e.kind, effect_struct(e.kind, e.params))
    for cable in plan.cables:
        for a, b in zip(cable[:-1], cable[1:]):
            if a in uid and b in uid:
                r.patch(uid[a], uid[b])
    for src, dst, idx in plan.links:
        if src in uid and dst in uid:
            r.link_control(uid[src], uid[dst], idx)
    r.finalize()
    ev = np.zeros(len(plan.events), dtype=abi.EVENT_DTYPE)
    k = 0
    for frame, uvid, typ, a, b, value in plan.events:
        if uvid in uid:
            ev[k] = (frame, uid[uvid], typ, a, b, value)
            k += 1
    r.push_events(ev[:k])
    return uid
