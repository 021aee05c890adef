"""The C++ host layer (pybind11 module `pycolmap_b200._core`, sources in pycolmap_b200/host/).

Same names and keyword arguments as `pycolmap_b200` itself (and as the reference:
R:pipeline/match_features.h:219-260, R:estimators/two_view_geometry.h:95-175), but the option
classes, the COLMAP database layer, the pair generators and the controllers are C++ and call the C ABI
of libb200match.so directly -- the shape a pycolmap maintainer would link.

    import pycolmap_b200.native as pycolmap
    pycolmap.match_exhaustive(database_path, sift_options={"max_ratio": 0.8})
"""
try:
    from ._core import *  # noqa: F401,F403
    from ._core import (Context, Database, Device, ExhaustiveMatchingOptions, RANSACOptions, Results,  # noqa: F401
                        SequentialMatchingOptions, SiftMatchingOptions, TwoViewGeometry,
                        TwoViewGeometryConfiguration, TwoViewGeometryOptions, abi_version,
                        essential_matrix_estimation, estimate_calibrated_two_view_geometry,
                        estimate_two_view_geometry, exhaustive_pair_blocks, fundamental_matrix_estimation,
                        has_cuda, homography_matrix_estimation, image_pair_to_pair_id, match_exhaustive,
                        match_sequential, pair_id_to_image_pair, sequential_pairs, sqlite_version,
                        squared_sampson_error, verify_matches)
except ImportError as e:  # never fall back silently: the C++ host is a build product
    raise ImportError(
        "pycolmap_b200._core is missing or does not load: build it with "
        "`python -c 'import __graft_entry__ as g; g.build()'` (needs libb200match.so next to it)") from e
