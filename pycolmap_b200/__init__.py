"""pycolmap_b200 -- B200-native exhaustive feature matcher + two-view geometric verifier.

Drop-in for the hot path of pycolmap.match_exhaustive / match_sequential / verify_matches /
estimate_two_view_geometry (R:pipeline/match_features.h, R:estimators/two_view_geometry.h):
hand-written sm_100a CUDA behind the C ABI in include/b200match.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import Context, B2MError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
