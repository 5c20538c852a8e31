"""ctypes mirror of ``include/groove_b200.h`` and a thin object wrapper over it.

``Renderer`` binds a shared library that exports the block-render C ABI under a
symbol prefix.  The product library (``libgroove_b200.so``) uses ``gb_``; see
``groove_b200.engine.Engine``.  The host-side method names follow the
reference's Orchestrator API (``add``/``patch``/``tick``:
orchestration/src/orchestrator.rs:174-304,856-877).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

ABI_VERSION = 4
MAIN_MIXER = 1

# --- error codes -------------------------------------------------------------
OK, EINVAL, ENOENT, ESTATE, ENODEV, ECUDA, ENOMEM, EGRAPH = 0, -1, -2, -3, -4, -5, -6, -7

# --- kinds ---------------------------------------------------------------------
INST_WELSH, INST_FM, INST_SAMPLER, INST_DRUMKIT, INST_TOY_SOURCE = 1, 2, 3, 4, 5
INST_OSCILLATOR, INST_ENVELOPE = 6, 7
FX_MIXER, FX_GAIN, FX_LIMITER, FX_BITCRUSHER, FX_COMPRESSOR = 32, 33, 34, 35, 36
FX_DELAY, FX_CHORUS, FX_REVERB = 37, 38, 39
FX_LOW_PASS_12DB, FX_HIGH_PASS_12DB, FX_BAND_PASS_12DB, FX_BAND_STOP_12DB = 40, 41, 42, 43
FX_ALL_PASS_12DB, FX_PEAKING_EQ_12DB, FX_LOW_SHELF_12DB, FX_HIGH_SHELF_12DB = 44, 45, 46, 47
FX_LOW_PASS_24DB = 48
FX_SIGNAL_PASSTHROUGH = 49
CONTROL_PERIOD = 64
BIQUAD_KINDS = tuple(range(40, 48))

WAVE_NONE, WAVE_SINE, WAVE_SQUARE, WAVE_PULSE_WIDTH, WAVE_TRIANGLE = 0, 1, 2, 3, 4
WAVE_SAWTOOTH, WAVE_NOISE, WAVE_DEBUG_ZERO, WAVE_DEBUG_MAX, WAVE_DEBUG_MIN = 5, 6, 7, 8, 9
LFO_NONE, LFO_AMPLITUDE, LFO_PITCH, LFO_PULSE_WIDTH, LFO_FILTER_CUTOFF = 0, 1, 2, 3, 4
EV_NOTE_ON, EV_NOTE_OFF, EV_CONTROL, EV_SET_PARAM = 1, 2, 3, 4


class OscillatorParams(C.Structure):
    _fields_ = [("waveform", C.c_int32), ("_pad", C.c_int32), ("pulse_width", C.c_double),
                ("frequency", C.c_double), ("fixed_frequency", C.c_double), ("frequency_tune", C.c_double)]


class EnvelopeParams(C.Structure):
    _fields_ = [("attack", C.c_double), ("decay", C.c_double), ("sustain", C.c_double), ("release", C.c_double)]


class DcaParams(C.Structure):
    _fields_ = [("gain", C.c_double), ("pan", C.c_double)]


class WelshParams(C.Structure):
    _fields_ = [("oscillator_1", OscillatorParams), ("oscillator_2", OscillatorParams),
                ("oscillator_2_sync", C.c_int32), ("lfo_routing", C.c_int32),
                ("oscillator_mix", C.c_double), ("amp_envelope", EnvelopeParams),
                ("lfo", OscillatorParams), ("lfo_depth", C.c_double),
                ("filter_cutoff_hz", C.c_double), ("filter_passband_ripple", C.c_double),
                ("filter_cutoff_start", C.c_double), ("filter_cutoff_end", C.c_double),
                ("filter_envelope", EnvelopeParams), ("voice_dca", DcaParams), ("dca", DcaParams),
                ("voices", C.c_uint32), ("_pad", C.c_uint32)]


class FmParams(C.Structure):
    _fields_ = [("ratio", C.c_double), ("depth", C.c_double), ("beta", C.c_double),
                ("carrier_envelope", EnvelopeParams), ("modulator_envelope", EnvelopeParams),
                ("dca", DcaParams), ("voices", C.c_uint32), ("_pad", C.c_uint32)]


class SamplerParams(C.Structure):
    _fields_ = [("root_hz", C.c_double), ("voices", C.c_uint32), ("_pad", C.c_uint32)]


class DrumkitParams(C.Structure):
    _fields_ = [("_reserved", C.c_uint32), ("_pad", C.c_uint32)]


class ToySourceParams(C.Structure):
    _fields_ = [("level_left", C.c_double), ("level_right", C.c_double)]


class OscillatorSourceParams(C.Structure):
    _fields_ = [("oscillator", OscillatorParams)]


class EnvelopeSourceParams(C.Structure):
    _fields_ = [("envelope", EnvelopeParams)]


class GainParams(C.Structure):
    _fields_ = [("ceiling", C.c_double)]


class LimiterParams(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double)]


class BitcrusherParams(C.Structure):
    _fields_ = [("bits", C.c_double)]


class CompressorParams(C.Structure):
    _fields_ = [("threshold", C.c_double), ("ratio", C.c_double), ("attack", C.c_double), ("release", C.c_double)]


class DelayParams(C.Structure):
    _fields_ = [("seconds", C.c_double)]


class ChorusParams(C.Structure):
    _fields_ = [("voices", C.c_double), ("delay_seconds", C.c_double), ("wet_dry_mix", C.c_double)]


class ReverbParams(C.Structure):
    _fields_ = [("attenuation", C.c_double), ("seconds", C.c_double)]


class BiquadParams(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("param2", C.c_double)]


class Lowpass24Params(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("passband_ripple", C.c_double)]


class Event(C.Structure):
    _fields_ = [("frame", C.c_int64), ("uid", C.c_uint32), ("type", C.c_uint32),
                ("a", C.c_int32), ("b", C.c_int32), ("value", C.c_double)]


EVENT_DTYPE = np.dtype([("frame", "<i8"), ("uid", "<u4"), ("type", "<u4"), ("a", "<i4"), ("b", "<i4"),
                        ("value", "<f8")])
assert EVENT_DTYPE.itemsize == C.sizeof(Event)


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("sample_rate", C.c_double),
                ("max_block", C.c_uint32), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("voice_kernel_launches", C.c_uint64),
                ("voice_kernel_ms", C.c_double), ("fx_kernel_ms", C.c_double), ("render_ms", C.c_double),
                ("voice_samples", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("rest_kernel_launches", C.c_uint64), ("rest_kernel_ms", C.c_double), ("rest_voice_samples", C.c_uint64),
                ("sweep_kernel_launches", C.c_uint64), ("sweep_kernel_ms", C.c_double), ("sweep_voice_samples", C.c_uint64),
                ("solo_kernel_launches", C.c_uint64), ("solo_kernel_ms", C.c_double), ("solo_voice_samples", C.c_uint64),
                ("solo_jobs", C.c_uint64), ("solo_class_items", C.c_uint64 * 4),
                ("fm_kernel_launches", C.c_uint64), ("fm_kernel_ms", C.c_double), ("idle_voice_samples", C.c_uint64),
                ("rest_ctas", C.c_uint64), ("sweep_ctas", C.c_uint64), ("fx_batched_nodes", C.c_uint64), ("rest_tp_launches", C.c_uint64),
                ("rest_vr_launches", C.c_uint64), ("rest_vr16_launches", C.c_uint64)]


# every symbol include/groove_b200.h declares (suffix after the prefix)
ABI_SYMBOLS = (
    "create", "destroy", "last_error", "add_instrument", "add_effect", "load_sample", "patch", "finalize",
    "push_events", "render_block", "render_pcm16", "render_device", "last_device_buffer", "read_last", "position",
    "save_state",
    "restore_state", "get_stats", "reset_stats", "set_timing", "measure_fma_peak", "link_control", "set_lookahead",
    "bus_exchange_create", "bus_exchange_destroy", "bus_exchange_export", "bus_exchange_open", "bus_exchange_publish",
    "bus_exchange_reduce", "bus_exchange_result",
)


class GrooveError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


def osc(waveform=WAVE_SINE, pulse_width=0.5, frequency=0.0, fixed_frequency=0.0, tune=1.0) -> OscillatorParams:
    return OscillatorParams(waveform, 0, pulse_width, frequency, fixed_frequency, tune)


def env(attack=0.0, decay=0.0, sustain=1.0, release=0.0) -> EnvelopeParams:
    return EnvelopeParams(attack, decay, sustain, release)


class Renderer:
    """Object wrapper over one engine handle of a block-render C ABI library."""

    def __init__(self, lib: C.CDLL, prefix: str, sample_rate: float = 44100.0, device: int = 0,
                 max_block: int = 0):
        self._lib = lib
        self._p = prefix
        self._h = C.c_void_p()
        self.sample_rate = float(sample_rate)
        self._bind()
        cfg = Config(ABI_VERSION, device, self.sample_rate, max_block, 0)
        rc = self._f("create")(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self._f("last_error")(None)
            raise GrooveError(rc, (msg or b"").decode())

    # -- plumbing ---------------------------------------------------------------
    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def _bind(self):
        vp = C.c_void_p
        sig = {
            "create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
            "destroy": (None, [vp]),
            "last_error": (C.c_char_p, [vp]),
            "add_instrument": (C.c_int, [vp, C.c_int32, vp, C.c_size_t, C.POINTER(C.c_uint32)]),
            "add_effect": (C.c_int, [vp, C.c_int32, vp, C.c_size_t, C.POINTER(C.c_uint32)]),
            "load_sample": (C.c_int, [vp, C.c_uint32, C.c_uint8, vp, C.c_size_t, C.c_int32, C.c_double, C.c_double]),
            "patch": (C.c_int, [vp, C.c_uint32, C.c_uint32]),
            "finalize": (C.c_int, [vp]),
            "push_events": (C.c_int, [vp, vp, C.c_size_t]),
            "render_block": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
            "render_pcm16": (C.c_int, [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
            "render_device": (C.c_int, [vp, C.c_size_t, C.POINTER(C.c_size_t)]),
            "read_last": (C.c_int, [vp, vp, C.c_size_t]),
            "last_device_buffer": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
            "position": (C.c_int64, [vp]),
            "save_state": (C.c_int, [vp, vp, C.POINTER(C.c_size_t)]),
            "restore_state": (C.c_int, [vp, vp, C.c_size_t]),
            "get_stats": (C.c_int, [vp, C.POINTER(Stats)]),
            "reset_stats": (C.c_int, [vp]),
            "set_timing": (C.c_int, [vp, C.c_int32]),
            "measure_fma_peak": (C.c_int, [vp, C.c_int32, C.POINTER(C.c_double)]),
            "link_control": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_int32]),
            "set_lookahead": (C.c_int, [vp, C.c_size_t]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(self._lib, self._p + name, None)
            if fn is None:
                continue  # a checker library may implement a subset
            fn.restype = res
            fn.argtypes = args

    def _check(self, rc: int):
        if rc != 0:
            msg = self._f("last_error")(self._h)
            raise GrooveError(rc, (msg or b"").decode())

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- graph construction -------------------------------------------------------
    def add_instrument(self, kind: int, params: Optional[C.Structure]) -> int:
        uid = C.c_uint32()
        ptr = C.byref(params) if params is not None else None
        size = C.sizeof(params) if params is not None else 0
        self._check(self._f("add_instrument")(self._h, kind, ptr, size, C.byref(uid)))
        return uid.value

    def add_effect(self, kind: int, params: Optional[C.Structure] = None) -> int:
        uid = C.c_uint32()
        ptr = C.byref(params) if params is not None else None
        size = C.sizeof(params) if params is not None else 0
        self._check(self._f("add_effect")(self._h, kind, ptr, size, C.byref(uid)))
        return uid.value

    def load_sample(self, uid: int, key: int, frames: np.ndarray, sample_rate: float, root_hz: float = 0.0):
        a = np.ascontiguousarray(frames, dtype=np.float64)
        channels = 1 if a.ndim == 1 else a.shape[1]
        n = a.shape[0]
        self._check(self._f("load_sample")(self._h, uid, key, a.ctypes.data, n, channels, sample_rate, root_hz))

    def patch(self, src: int, dst: int):
        self._check(self._f("patch")(self._h, src, dst))

    def patch_chain(self, uids: Sequence[int]):
        """Patch a source -> ... -> sink cable (settings/src/songs.rs:134-164)."""
        for a, b in zip(uids[:-1], uids[1:]):
            self.patch(a, b)

    def link_control(self, source: int, target: int, control_index: int):
        """Control link from a signal-passthrough node (sidechain): include/groove_b200.h, gb_link_control."""
        self._check(self._f("link_control")(self._h, source, target, control_index))

    def finalize(self):
        self._check(self._f("finalize")(self._h))

    # -- events ---------------------------------------------------------------------
    def push_events(self, events) -> None:
        if isinstance(events, np.ndarray):
            arr = np.ascontiguousarray(events, dtype=EVENT_DTYPE)
        else:
            events = list(events)
            arr = np.zeros(len(events), dtype=EVENT_DTYPE)
            for i, ev in enumerate(events):
                arr[i] = tuple(ev)
        if arr.size:
            self._check(self._f("push_events")(self._h, arr.ctypes.data, arr.size))

    def note_on(self, frame: int, uid: int, key: int, velocity: int = 127):
        self.push_events([(frame, uid, EV_NOTE_ON, key, velocity, 0.0)])

    def note_off(self, frame: int, uid: int, key: int):
        self.push_events([(frame, uid, EV_NOTE_OFF, key, 0, 0.0)])

    def control(self, frame: int, uid: int, index: int, value: float):
        self.push_events([(frame, uid, EV_CONTROL, index, 0, value)])

    def set_param(self, frame: int, uid: int, index: int, value: float):
        self.push_events([(frame, uid, EV_SET_PARAM, index, 0, value)])

    # -- rendering --------------------------------------------------------------------
    def render(self, frames: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Render ``frames`` stereo frames; returns an array of shape (frames, 2), f64."""
        if out is None:
            out = np.empty((frames, 2), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size >= 2 * frames
        done = C.c_size_t()
        self._check(self._f("render_block")(self._h, out.ctypes.data, frames, C.byref(done)))
        return out

    def render_pcm16(self, frames: int) -> np.ndarray:
        out = np.empty((frames, 2), dtype=np.int16)
        done = C.c_size_t()
        self._check(self._f("render_pcm16")(self._h, out.ctypes.data, frames, C.byref(done)))
        return out

    def render_device(self, frames: int) -> int:
        done = C.c_size_t()
        self._check(self._f("render_device")(self._h, frames, C.byref(done)))
        return done.value

    def read_last(self, frames: int) -> np.ndarray:
        out = np.empty((frames, 2), dtype=np.float64)
        self._check(self._f("read_last")(self._h, out.ctypes.data, frames))
        return out

    def last_device_buffer(self):
        """(device pointer, frames) of the last ``render_device`` result (f64 L,R interleaved in HBM)."""
        ptr = C.c_void_p()
        n = C.c_size_t()
        self._check(self._f("last_device_buffer")(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def set_lookahead(self, frames: int):
        """Serve small render() calls from one big render every ``frames`` frames (include/groove_b200.h)."""
        self._check(self._f("set_lookahead")(self._h, frames))

    @property
    def position(self) -> int:
        return int(self._f("position")(self._h))

    def save_state(self) -> bytes:
        size = C.c_size_t()
        self._check(self._f("save_state")(self._h, None, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        self._check(self._f("save_state")(self._h, buf, C.byref(size)))
        return buf.raw[: size.value]

    def restore_state(self, blob: bytes):
        self._check(self._f("restore_state")(self._h, blob, len(blob)))

    # -- measurement ---------------------------------------------------------------------
    def stats(self) -> Stats:
        s = Stats()
        self._check(self._f("get_stats")(self._h, C.byref(s)))
        return s

    def reset_stats(self):
        self._check(self._f("reset_stats")(self._h))

    def set_timing(self, enabled: bool):
        self._check(self._f("set_timing")(self._h, 1 if enabled else 0))

    def measure_fma_peak(self, fp64: bool = True) -> float:
        v = C.c_double()
        self._check(self._f("measure_fma_peak")(self._h, 1 if fp64 else 0, C.byref(v)))
        return v.value
