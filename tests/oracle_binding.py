"""Test-side binding of the CPU oracle (oracle/libgroove_oracle.so, prefix ``go_``).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from groove_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgroove_oracle.so")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "groove_oracle.cpp")
    hdr = os.path.join(ROOT, "include", "groove_b200.h")
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(ORACLE_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"] + (["-B"] if force else []))
    return ORACLE_SO


_lib = None


def oracle_lib() -> C.CDLL:
    global _lib
    if _lib is None:
        # GROOVE_ORACLE_SO: bench.py's CPU legs point this at their -O3 -march=native build of the same source
        _lib = C.CDLL(os.environ.get("GROOVE_ORACLE_SO") or build_oracle())
        _lib.go_tune_ratio.restype = C.c_double
        _lib.go_tune_ratio.argtypes = [C.c_int, C.c_double]
        _lib.go_note_hz.restype = C.c_double
        _lib.go_note_hz.argtypes = [C.c_int]
        _lib.go_pct_to_hz.restype = C.c_double
        _lib.go_pct_to_hz.argtypes = [C.c_double]
        _lib.go_envelope_level.restype = C.c_double
        _lib.go_envelope_level.argtypes = [C.POINTER(abi.EnvelopeParams), C.c_double, C.c_int64, C.c_int64, C.c_int64]
        _lib.go_lp24_coefficients.restype = None
        _lib.go_lp24_coefficients.argtypes = [C.c_double, C.c_double, C.c_double, C.c_void_p]
        _lib.go_rbj_coefficients.restype = None
        _lib.go_rbj_coefficients.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p]
        for fn in (_lib.go_mma_concave, _lib.go_mma_convex):
            fn.restype = C.c_double
            fn.argtypes = [C.c_double]
        _lib.go_pcm16.restype = None
        _lib.go_pcm16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    return _lib


class OracleEngine(abi.Renderer):
    def __init__(self, sample_rate: float = 44100.0):
        super().__init__(oracle_lib(), "go_", sample_rate=sample_rate)


def pcm16(x: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(a.shape, dtype=np.int16)
    oracle_lib().go_pcm16(a.ctypes.data, out.ctypes.data, a.size)
    return out
