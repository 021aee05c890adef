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
    assert lib.b2m_abi_version() == 2


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


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The header compiles as plain C and every struct the ctypes binding mirrors has the C size."""
    import subprocess
    names = {"b2m_device_cfg": _lib.DeviceCfg, "b2m_sift_opts": _lib.SiftOpts, "b2m_ransac_opts": _lib.RansacOpts,
             "b2m_tvg_opts": _lib.TvgOpts, "b2m_camera": _lib.Camera, "b2m_pair_view": _lib.PairView,
             "b2m_tvg_result": _lib.TvgResult, "b2m_tvg_problem": _lib.TvgProblem, "b2m_stats": _lib.Stats}
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "b200match.h"\nint main(void) {\n' +
                   "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names) + "  return 0;\n}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for n, cls in names.items():
        assert int(sizes[n]) == ctypes.sizeof(cls), n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.B2MError, match="no CPU fallback"):
        pb.Context()
