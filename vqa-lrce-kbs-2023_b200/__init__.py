"""lrce_b200 — B200-native (sm_100a) forward hot path of LRCE behind the reference's E2E model API."""
from . import _lib, dist, feed, ops  # noqa: F401
from ._lib import LrceError  # noqa: F401
from .e2e import E2ECount, E2EMultipleChoice, E2EOpenEnded, install  # noqa: F401
from .feature_extractor import SwinTransformer3D, TextExtractor, VideoExtractor  # noqa: F401
from .fusion import LRCECount, LRCEMultipleChoice, LRCEOpenEnded  # noqa: F401

__all__ = ["E2EOpenEnded", "E2EMultipleChoice", "E2ECount", "VideoExtractor", "TextExtractor", "SwinTransformer3D",
           "LRCEOpenEnded", "LRCEMultipleChoice", "LRCECount", "install", "ops", "dist", "feed", "LrceError"]
