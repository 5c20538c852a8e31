"""groove_b200 — B200-native block renderer for Groove's synthesis + effects hot path."""
from . import abi  # noqa: F401
from .engine import Engine, load_library  # noqa: F401

__all__ = ["abi", "Engine", "load_library"]
