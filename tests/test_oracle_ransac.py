"""CPU: the numpy verifier oracle on planted scenes (known answers derived from the semantics;
the reference holds no golden vectors for this path -- parity unpinned), the product's solver
math (geom.h, host build) against the oracle's independent solvers, and the pair generators."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ransac as R
from helpers import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def geom():
    so = os.path.join(HERE, "helpers", "libgeom_host.so")
    src = os.path.join(HERE, "helpers", "geom_host.cpp")
    hdr = os.path.join(HERE, "..", "pycolmap_b200", "csrc", "geom.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.gh_sampson.restype = ctypes.c_double
    lib.gh_homography.restype = ctypes.c_double
    lib.gh_num_trials.restype = ctypes.c_double
    lib.gh_num_trials.argtypes = [ctypes.c_double] * 4 + [ctypes.c_int]
    return lib


def P(a):
    return np.ascontiguousarray(a, np.float64).ctypes.data_as(ctypes.c_void_p)


def _canon(M):
    M = np.asarray(M, np.float64).reshape(9)
    M = M / np.linalg.norm(M)
    return M * np.sign(M[np.abs(M).argmax()])


@pytest.mark.parametrize("kind,expect", [("general", R.CALIBRATED), ("planar", R.PLANAR_OR_PANORAMIC),
                                         ("rotation", R.PLANAR_OR_PANORAMIC)])
def test_oracle_config_and_mask(kind, expect):
    rng = np.random.default_rng(3)
    p1, p2, inl = scenes.two_view_scene(rng, 300, 0.3, kind)
    g = R.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2, seed=1)
    assert g.config == expect
    got = np.zeros(300, bool)
    got[g.inlier_matches[:, 0]] = True
    # planar scenes leave F under-constrained: a few outliers may satisfy the chosen epipolar model
    assert (got & inl).sum() >= 0.99 * inl.sum() and (got & ~inl).sum() <= (3 if kind == "general" else 9)


def test_oracle_uncalibrated_and_degenerate():
    rng = np.random.default_rng(4)
    p1, p2, inl = scenes.two_view_scene(rng, 300, 0.3, "general")
    g = R.estimate_two_view_geometry(scenes.CAM_NOPRIOR, p1, scenes.CAM_NOPRIOR, p2, seed=1)
    assert g.config == R.UNCALIBRATED and g.nE == 0 and abs(len(g.inlier_matches) - inl.sum()) <= 3
    g = R.estimate_two_view_geometry(scenes.CAM, p1[:10], scenes.CAM, p2[:10])
    assert g.config == R.DEGENERATE and len(g.inlier_matches) == 0
    g = R.estimate_two_view_geometry(scenes.CAM, rng.uniform(0, 1000, (40, 2)), scenes.CAM,
                                     rng.uniform(0, 1000, (40, 2)), seed=2)
    assert g.config == R.DEGENERATE


def test_oracle_watermark():
    rng = np.random.default_rng(5)
    n = 200
    p1 = np.c_[rng.uniform(0, 1600, n), rng.uniform(0, 100, n)]   # all in the top border strip
    p2 = p1 + [7.0, 3.0]
    g = R.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2, seed=1)
    assert g.config == R.WATERMARK


def test_sampson_closed_form(geom):
    rng = np.random.default_rng(6)
    E = rng.normal(size=(3, 3))
    p1, p2 = rng.normal(size=(50, 2)), rng.normal(size=(50, 2))
    r = R.squared_sampson_error(p1, p2, E)
    for i in range(50):
        x1, x2 = np.r_[p1[i], 1], np.r_[p2[i], 1]
        Ex1, Etx2 = E @ x1, E.T @ x2
        ref = (x2 @ Ex1) ** 2 / (Ex1[0] ** 2 + Ex1[1] ** 2 + Etx2[0] ** 2 + Etx2[1] ** 2)
        assert np.isclose(r[i], ref, rtol=1e-12)
        assert np.isclose(geom.gh_sampson(P(E), *map(ctypes.c_double, (p1[i, 0], p1[i, 1], p2[i, 0], p2[i, 1]))),
                          ref, rtol=1e-12)


def test_five_point_two_algorithms_agree(geom):
    """geom.h (hidden-variable elimination) vs oracle (action matrix): same solution sets."""
    rng = np.random.default_rng(7)
    agree = 0
    for _ in range(40):
        p1, p2, _ = scenes.two_view_scene(rng, 5, 0.0, "general")
        n1, n2 = R.cam_from_img(scenes.CAM, p1), R.cam_from_img(scenes.CAM, p2)
        a = sorted(tuple(np.round(_canon(E), 5)) for E in R.EssentialFivePoint.estimate(n1, n2))
        m = np.zeros(90)
        k = geom.gh_minimal_E5(P(n1[:, 0]), P(n1[:, 1]), P(n2[:, 0]), P(n2[:, 1]), P(m))
        b = sorted(tuple(np.round(_canon(m[9 * i:9 * i + 9]), 5)) for i in range(k))
        agree += len(a) == len(b) and np.allclose(a, b, atol=2e-4)
    assert agree >= 36


def test_seven_eight_dlt_recover_model(geom):
    rng = np.random.default_rng(8)
    for _ in range(20):
        p1, p2, _ = scenes.two_view_scene(rng, 30, 0.0, "general")
        F8 = R.FundamentalEightPoint.estimate(p1, p2)[0]
        assert R.squared_sampson_error(p1, p2, F8).max() < 1e-12
        m = np.zeros(9)
        geom.gh_estimate_F8(P(p1[:, 0]), P(p1[:, 1]), P(p2[:, 0]), P(p2[:, 1]), 30, P(m))
        assert np.allclose(_canon(m), _canon(F8), atol=1e-6)
        m = np.zeros(27)
        k = geom.gh_minimal_F7(P(p1[:7, 0]), P(p1[:7, 1]), P(p2[:7, 0]), P(p2[:7, 1]), P(m))
        assert any(np.allclose(_canon(m[9 * i:9 * i + 9]), _canon(F8), atol=1e-5) for i in range(k))
        assert any(np.allclose(_canon(F), _canon(F8), atol=1e-5) for F in R.FundamentalSevenPoint.estimate(p1[:7], p2[:7]))
        q1, q2, _ = scenes.two_view_scene(rng, 30, 0.0, "planar")
        H = R.Homography.estimate(q1, q2)[0]
        assert R.homography_residuals(q1, q2, H).max() < 1e-12
        m = np.zeros(9)
        geom.gh_minimal_H4(P(q1[:4, 0]), P(q1[:4, 1]), P(q2[:4, 0]), P(q2[:4, 1]), P(m))
        assert np.allclose(_canon(m), _canon(H), atol=1e-6)
        m = np.zeros(9)
        assert geom.gh_minimal_H4_closed(P(q1[4:8, 0]), P(q1[4:8, 1]), P(q2[4:8, 0]), P(q2[4:8, 1]), P(m)) == 1
        assert np.allclose(_canon(m), _canon(H), atol=1e-6)          # closed form == DLT on minimal samples


def test_inverse_iteration_null_space(geom):
    """LO refits: K smallest eigenvectors by inverse subspace iteration span numpy's eigh subspace."""
    rng = np.random.default_rng(9)
    for K, rows, noise in [(1, 40, 1e-3), (4, 60, 1e-4), (1, 8, 0.0), (4, 5, 0.0), (4, 200, 1e-2)]:
        for _ in range(10):
            basis = np.linalg.qr(rng.normal(size=(9, 9)))[0]
            null, rest = basis[:, :K], basis[:, K:]
            A = rng.normal(size=(rows, 9 - K)) @ rest.T + noise * rng.normal(size=(rows, 9))
            S = A.T @ A
            out = np.zeros((K, 9))
            geom.gh_invit(P(S), K, P(out))
            w, V = np.linalg.eigh(S)
            ref = V[:, :K]
            # principal angles between the two K-dim subspaces
            sv = np.linalg.svd(ref.T @ out.T, compute_uv=False)
            assert sv.min() > 1 - 1e-9, (K, rows, noise, sv)
            assert np.allclose(out @ out.T, np.eye(K), atol=1e-12)


def test_num_trials(geom):
    for ni, ns, k in [(100, 400, 5), (350, 400, 7), (0, 10, 4), (10, 10, 4), (25000, 100000, 4)]:
        a = R.compute_num_trials(ni, ns, 0.999, 3.0, k)
        b = geom.gh_num_trials(ni, ns, 0.999, 3.0, k)
        assert (a == float("inf") and b > 1e17) or a == b
    assert R.compute_num_trials(25000, 100000, 0.999, 3.0, 4) == 5295   # H: max_num_trials clipped


def test_pair_generators():
    for n, bs in [(1, 50), (2, 1), (49, 50), (50, 50), (51, 50), (101, 50), (7, 2)]:
        pr = R.exhaustive_pairs(range(n), bs)
        assert len(pr) == n * (n - 1) // 2 == len({(min(a, b), max(a, b)) for a, b in pr})
    # COLMAP 3.9.1: idx2 = idx1 + i, i in [0, overlap) -> overlap - 1 linear neighbours (i = 0 is the self pair)
    assert len(R.sequential_pairs(range(10000), 20, False)) == 10000 * 19 - 19 * 20 // 2
    s = R.sequential_pairs(range(100), 3, True)
    assert (0, 1) in s and (0, 2) in s and (0, 4) in s and (0, 3) not in s and (0, 5) not in s and (0, 0) not in s
    assert R.sequential_pairs(range(5), 1, False) == []


def _two_motion_scene(rng, na, nb, n_out):
    a1, a2, _ = scenes.two_view_scene(rng, na, 0.0, "general")
    b1, b2, _ = scenes.two_view_scene(rng, nb, 0.0, "general")     # a second, independent rigid motion
    o1 = np.c_[rng.uniform(0, 1600, n_out), rng.uniform(0, 1200, n_out)]
    o2 = np.c_[rng.uniform(0, 1600, n_out), rng.uniform(0, 1200, n_out)]
    return np.concatenate([a1, b1, o1]), np.concatenate([a2, b2, o2])


def test_multiple_models_finds_both_motions():
    """EstimateMultipleTwoViewGeometries: estimate, remove the inliers, repeat until DEGENERATE."""
    rng = np.random.default_rng(3)
    p1, p2 = _two_motion_scene(rng, 70, 50, 15)
    g = R.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2,
                                     options=R.TwoViewGeometryOptions(multiple_models=True), seed=1)
    im = np.asarray(g.inlier_matches)
    assert g.config == R.MULTIPLE and len(np.unique(im[:, 0])) == len(im)
    assert (im[:, 0] < 70).sum() >= 68 and ((im[:, 0] >= 70) & (im[:, 0] < 120)).sum() >= 48 and (im[:, 0] >= 120).sum() <= 4
    first = R.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2, seed=1)
    assert first.config == R.CALIBRATED and len(first.inlier_matches) < len(im)      # one model explains one motion
    # a single-motion scene comes back as itself, not as MULTIPLE
    q1, q2, pl = scenes.two_view_scene(rng, 90, 0.2, "general")
    g1 = R.estimate_two_view_geometry(scenes.CAM, q1, scenes.CAM, q2,
                                      options=R.TwoViewGeometryOptions(multiple_models=True), seed=2)
    assert g1.config == R.CALIBRATED and abs(len(g1.inlier_matches) - pl.sum()) <= 3


def test_sequential_cpp_oracle_agrees_with_numpy_oracle():
    """oracle/ransac_seq.cpp (the scalar fp64 sequential LO-RANSAC that bench.py times as the CPU arm) against the numpy
    restatement oracle/ransac.py: same configuration, inlier counts within +-1 % (RANSAC parity is statistical), on the
    planted scene types; the pair-parallel entry point returns what the single calls return."""
    from oracle import ransac_seq as S
    rng = np.random.default_rng(21)
    jobs, singles = [], []
    for kind, cam in (("general", scenes.CAM), ("planar", scenes.CAM), ("rotation", scenes.CAM),
                      ("general", scenes.CAM_NOPRIOR)):
        p1, p2, planted = scenes.two_view_scene(rng, 500, 0.3, kind, 0.3)
        g = R.estimate_two_view_geometry(cam, p1, cam, p2, seed=2)
        s = S.estimate_two_view_geometry(cam, p1, cam, p2, seed=2)
        assert s["config"] == g.config, (kind, s["config"], g.config)
        tol = max(3, int(0.01 * len(g.inlier_matches)))
        extra = max(3, int(0.03 * len(p1))) if kind != "general" else 0          # degenerate F: admitted outliers vary
        assert abs(len(s["inlier_matches"]) - len(g.inlier_matches)) <= max(tol, extra)
        assert abs(s["nE"] - g.nE) <= max(tol, extra) and abs(s["nH"] - g.nH) <= max(tol, int(0.03 * len(p1)))
        assert s["models_scored"] > 100
        m = np.stack([np.arange(len(p1), dtype=np.uint32)] * 2, 1)
        jobs.append((cam, p1, cam, p2, m))
        singles.append(s)
    few1, few2, _ = scenes.two_view_scene(rng, 10, 0.0)
    assert S.estimate_two_view_geometry(scenes.CAM, few1, scenes.CAM, few2)["config"] == R.DEGENERATE
    out, scored = S.verify_pairs(jobs, seed=2, n_threads=3)
    for k, s in enumerate(singles):   # job k runs with seed + k: configuration equal, counts statistically equal
        assert out[k, 0] == s["config"] and abs(out[k, 4] - len(s["inlier_matches"])) <= max(3, int(0.03 * 500))
    assert scored > 400
