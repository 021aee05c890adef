"""CPU: the C-ABI library loads and exports every symbol include/b200match.h declares; without a
GPU it must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import pycolmap_b200 as pb
from pycolmap_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "b200match.h")).read()
    declared = set(re.findall(r"\b(b2m_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b2m_abi_version() == 1


def test_struct_sizes_match_defaults():
    lib = _lib.load()
    s = _lib.SiftOpts()
    lib.b2m_sift_opts_default(ctypes.byref(s))
    assert s.struct_size == ctypes.sizeof(_lib.SiftOpts)
    assert (round(s.max_ratio, 6), round(s.max_distance, 6), s.cross_check, s.max_num_matches) == (0.8, 0.7, 1, 32768)
    t = _lib.TvgOpts()
    lib.b2m_tvg_opts_default(ctypes.byref(t))
    assert t.struct_size == ctypes.sizeof(_lib.TvgOpts)
    assert t.ransac.struct_size == ctypes.sizeof(_lib.RansacOpts)
    assert (t.min_num_inliers, t.min_E_F_inlier_ratio, t.max_H_inlier_ratio) == (15, 0.95, 0.8)
    assert (t.ransac.max_error, t.ransac.confidence, t.ransac.min_num_trials, t.ransac.max_num_trials,
            t.ransac.min_inlier_ratio) == (4.0, 0.999, 100, 10000, 0.25)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.B2MError, match="no CPU fallback"):
        pb.Context()
