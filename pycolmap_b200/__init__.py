"""pycolmap_b200 -- B200-native exhaustive feature matcher + two-view geometric verifier.

Drop-in for the hot path of pycolmap.match_exhaustive / match_sequential / verify_matches /
estimate_two_view_geometry (R:pipeline/match_features.h, R:estimators/two_view_geometry.h):
hand-written sm_100a CUDA behind the C ABI in include/b200match.h.  No CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import Context, B2MError, LIB_PATH  # noqa: F401
from .options import (Device, ExhaustiveMatchingOptions, RANSACOptions, SequentialMatchingOptions,  # noqa: F401
                      SiftMatchingOptions, TwoViewGeometryConfiguration, TwoViewGeometryOptions)
from .database import Database, image_pair_to_pair_id  # noqa: F401
from .pipeline import (Rigid3d, Rotation3d, TwoViewGeometry, essential_matrix_estimation, estimate_calibrated_two_view_geometry,  # noqa: F401
                       estimate_two_view_geometry, estimate_two_view_geometry_pose, fundamental_matrix_estimation, homography_matrix_estimation,
                       match_exhaustive, match_sequential, squared_sampson_error, verify_matches)

__version__ = "0.1.0"
has_cuda = True  # R:main.cc:98: this build has nothing but the CUDA path
