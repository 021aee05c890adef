"""CPU: the C-ABI library loads and exports every symbol include/b200match.h declares, the header is plain C,
the option defaults are the reference's, and without a GPU the library fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pycolmap_b200", "libb200match.so")
HDR = os.path.join(ROOT, "include", "b200match.h")


def _declared():
    return set(re.findall(r"\b(b2m_[a-z0-9_]+)\s*\(", open(HDR).read()))


def test_header_symbols_exported():
    lib = ctypes.CDLL(LIB)
    declared = _declared()
    assert len(declared) >= 26
    for name in declared:
        assert hasattr(lib, name), name
    # ... and nothing else leaks out under the b2m_ prefix
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("b2m_")}
    assert exported == declared, exported ^ declared
    ver = int(re.search(r"#define\s+B2M_ABI_VERSION\s+(\d+)", open(HDR).read()).group(1))
    assert lib.b2m_abi_version() == ver


def test_header_is_plain_c_and_defaults_are_the_references(tmp_path):
    """The header compiles as C11 with -Wall -Werror; a C program reads the option defaults through the ABI
    (SiftMatchingOptions / TwoViewGeometryOptions / RANSAC C++-constructor defaults)."""
    src = tmp_path / "defaults.c"
    src.write_text(r'''
#include <stdio.h>
#include "b200match.h"
int main(void) {
  b2m_sift_opts s; b2m_tvg_opts t;
  b2m_sift_opts_default(&s); b2m_tvg_opts_default(&t);
  if (s.struct_size != sizeof(s) || t.struct_size != sizeof(t) || t.ransac.struct_size != sizeof(t.ransac)) return 2;
  printf("%.6f %.6f %d %d %d\n", s.max_ratio, s.max_distance, s.cross_check, s.max_num_matches, s.guided_matching);
  printf("%d %.3f %.3f %.3f %.3f %d %d %d %d %d\n", t.min_num_inliers, t.min_E_F_inlier_ratio, t.max_H_inlier_ratio,
         t.watermark_min_inlier_ratio, t.watermark_border_size, t.detect_watermark, t.multiple_ignore_watermark,
         t.force_H_use, t.compute_relative_pose, t.multiple_models);
  printf("%.3f %.4f %d %d %.3f %.1f\n", t.ransac.max_error, t.ransac.confidence, t.ransac.min_num_trials,
         t.ransac.max_num_trials, t.ransac.min_inlier_ratio, t.ransac.dyn_num_trials_multiplier);
  printf("%zu %zu %zu %zu\n", sizeof(b2m_camera), sizeof(b2m_pair_view), sizeof(b2m_tvg_result), sizeof(b2m_stats));
  return 0;
}
''')
    exe = tmp_path / "defaults"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           "-L", os.path.dirname(LIB), "-lb200match", "-Wl,-rpath," + os.path.dirname(LIB)])
    lines = subprocess.check_output([str(exe)], text=True).splitlines()
    assert lines[0] == "0.800000 0.700000 1 32768 0"
    assert lines[1] == "15 0.950 0.800 0.700 0.100 1 1 0 0 0"
    assert lines[2] == "4.000 0.9990 100 10000 0.250 3.0"
    assert all(int(x) % 8 == 0 for x in lines[3].split())


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pycolmap_b200 as pb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pb.Context()
    lib = ctypes.CDLL(LIB)
    lib.b2m_last_error.restype = ctypes.c_char_p
    ctx = ctypes.c_void_p()
    assert lib.b2m_create(None, ctypes.byref(ctx)) == -2 and not ctx.value          # B2M_ENODEV
    assert b"no CPU fallback" in lib.b2m_last_error(None)
