"""lrce_b200 — B200-native (sm_100a) forward hot path of LRCE behind the reference's E2E model API."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
