"""The B200 block renderer: loads ``libgroove_b200.so`` (hand-written CUDA for sm_100a) over ctypes.

There is no CPU fallback.  If the library has not been built, or no CUDA device is present,
construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgroove_b200.so")

_lib = None


def load_library() -> C.CDLL:
    """Load the CUDA library.  Raises if it is missing — the product never falls back to CPU code."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m groove_b200.build` "
                "(nvcc, sm_100a). groove_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        missing = [s for s in abi.ABI_SYMBOLS if not hasattr(_lib, "gb_" + s)]
        if missing:
            raise RuntimeError(f"{LIB_PATH} does not export: {missing}")
    return _lib


class Engine(abi.Renderer):
    """One engine = one Orchestrator-equivalent render graph living on one GPU."""

    def __init__(self, sample_rate: float = 44100.0, device: int = 0, max_block: int = 0):
        super().__init__(load_library(), "gb_", sample_rate=sample_rate, device=device, max_block=max_block)
