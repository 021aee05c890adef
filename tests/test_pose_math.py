"""CPU: pycolmap_b200/csrc/pose.h (the header the pose kernel compiles) against the numpy oracle
(oracle.ransac: np.linalg.svd based) and against planted poses -- essential / homography decomposition,
triangulation, cheirality selection, median triangulation angle, quaternion conversion."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ransac as R

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ph():
    so = os.path.join(HERE, "helpers", "libpose_host.so")
    src = os.path.join(HERE, "helpers", "pose_host.cpp")
    hdrs = [os.path.join(HERE, "..", "pycolmap_b200", "csrc", h) for h in ("pose.h", "geom.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.ph_triangulate.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_double] * 4 + [ctypes.c_void_p]
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def C(a):
    return np.ascontiguousarray(a, np.float64)


def rot(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def scene(rng, Rm, t, n=60, planar=None):
    X = np.c_[rng.uniform(-2, 2, n), rng.uniform(-2, 2, n), rng.uniform(4, 9, n)]
    if planar is not None:
        nrm, d = planar
        X[:, 2] = (d - nrm[0] * X[:, 0] - nrm[1] * X[:, 1]) / nrm[2]
    Xc = X @ Rm.T + t
    return X, C(X[:, :2] / X[:, 2:]), C(Xc[:, :2] / Xc[:, 2:])


def test_svd3(ph):
    rng = np.random.default_rng(0)
    mats = [rng.normal(size=(3, 3)) for _ in range(30)] + [skew(rng.normal(size=3)) @ rot(rng.normal(size=3)) for _ in range(10)]
    mats += [np.outer(rng.normal(size=3), rng.normal(size=3)), np.zeros((3, 3)), np.eye(3) * 2.5]
    for A in mats:
        A = C(A)
        U, S, V = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        ph.ph_svd3(P(A), P(U), P(S), P(V))
        assert np.allclose(U @ np.diag(S) @ V.T, A, atol=1e-7 * max(1.0, np.abs(A).max()))
        # singular values come from the eigenvalues of A^T A: a zero one is accurate to sqrt(eps) * scale only
        assert np.allclose(np.abs(S), np.linalg.svd(A)[1], atol=1e-7 * max(1.0, np.abs(A).max()))
        assert np.allclose(U.T @ U, np.eye(3), atol=1e-10) and np.allclose(V.T @ V, np.eye(3), atol=1e-10)
        assert np.isclose(np.linalg.det(U), 1.0) and np.isclose(np.linalg.det(V), 1.0)


def test_essential_decomposition_and_selection(ph):
    rng = np.random.default_rng(1)
    for _ in range(25):
        Rm, t = rot(rng.normal(size=3) * 0.3), rng.normal(size=3)
        t /= np.linalg.norm(t)
        E = C(skew(t) @ Rm * rng.uniform(0.5, 2) * rng.choice([-1, 1]))
        X, x1, x2 = scene(rng, Rm, t)
        R1, R2, tt = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3)
        ph.ph_decompose_E(P(E), P(R1), P(R2), P(tt))
        o1, o2, ot = R.decompose_essential_matrix(E)
        # the same four candidate poses as the oracle's SVD gives (as a set)
        mine = [(R1, tt), (R2, tt), (R1, -tt), (R2, -tt)]
        theirs = [(o1, ot), (o2, ot), (o1, -ot), (o2, -ot)]
        for Rc, tc in theirs:
            assert min(np.abs(Rc - a).max() + np.abs(tc - b).max() for a, b in mine) < 1e-9
        Rs = C(np.stack([m[0] for m in mine]))
        ts = C(np.stack([m[1] for m in mine]))
        n_front, tri = ctypes.c_int(0), ctypes.c_double(0)
        best = ph.ph_select_pose(P(Rs), P(ts), 4, P(x1), P(x2), len(x1), ctypes.byref(n_front), ctypes.byref(tri))
        assert np.allclose(Rs[best], Rm, atol=1e-9) and np.allclose(ts[best], t, atol=1e-9) and n_front.value == len(x1)
        Ro, to, Xo = R.pose_from_essential_matrix(E, x1, x2)
        assert np.allclose(Ro, Rm, atol=1e-9) and len(Xo) == len(x1)
        want = R.median(R.triangulation_angles(np.zeros(3), -Rm.T @ t, X))
        assert abs(tri.value - want) < 1e-9
        Xt = np.zeros(3)
        assert ph.ph_triangulate(P(C(Rm)), P(C(t)), x1[0, 0], x1[0, 1], x2[0, 0], x2[0, 1], P(Xt)) == 1
        assert np.allclose(Xt, X[0], atol=1e-8)


def test_homography_decomposition(ph):
    rng = np.random.default_rng(2)
    K1 = C([1200.0, 1190.0, 800.0, 600.0])
    K2 = C([900.0, 905.0, 640.0, 480.0])
    Km = [np.array([[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1.0]]) for k in (K1, K2)]
    for _ in range(25):
        Rm, t = rot(rng.normal(size=3) * 0.2), rng.normal(size=3) * 0.5
        nrm = np.array([0.1, -0.2, 1.0]) + rng.normal(size=3) * 0.05
        nrm /= np.linalg.norm(nrm)
        d = 6.0
        H = C(Km[1] @ (Rm + np.outer(t, nrm) / d) @ np.linalg.inv(Km[0]) * rng.uniform(0.5, 2) * rng.choice([-1, 1]))
        Rs, ts, ns = np.zeros((4, 3, 3)), np.zeros((4, 3)), np.zeros((4, 3))
        assert ph.ph_decompose_H(P(H), P(K1), P(K2), P(Rs), P(ts), P(ns)) == 4
        oracle_c = R.decompose_homography_matrix(H, Km[0], Km[1])
        for k in range(4):                                        # same candidates in the same order
            assert np.allclose(Rs[k], oracle_c[k][0], atol=1e-8) and np.allclose(ts[k], oracle_c[k][1], atol=1e-8)
            assert np.allclose(ns[k], oracle_c[k][2], atol=1e-8)
        # the planted pose is among them (COLMAP's normal points towards the camera: n = -n_plane)
        assert min(np.abs(Rs[k] - Rm).max() + np.abs(ts[k] - t / d).max() + np.abs(ns[k] + nrm).max() for k in range(4)) < 1e-8
        X, x1, x2 = scene(rng, Rm, t, 40, planar=(nrm, d))
        n_front, tri = ctypes.c_int(0), ctypes.c_double(0)
        best = ph.ph_select_pose(P(C(Rs)), P(C(ts)), 4, P(x1), P(x2), len(x1), ctypes.byref(n_front), ctypes.byref(tri))
        Ro, to, _, Xo = R.pose_from_homography_matrix(H, Km[0], Km[1], x1, x2)
        assert np.allclose(Rs[best], Ro, atol=1e-8) and np.allclose(ts[best], to, atol=1e-8) and n_front.value == len(Xo) == 40
    Hr = C(Km[1] @ rot(np.array([0.1, 0.2, -0.1])) @ np.linalg.inv(Km[0]) * 3.0)      # pure rotation: one candidate
    Rs, ts, ns = np.zeros((4, 3, 3)), np.zeros((4, 3)), np.zeros((4, 3))
    assert ph.ph_decompose_H(P(Hr), P(K1), P(K2), P(Rs), P(ts), P(ns)) == 1
    assert np.allclose(Rs[0], rot(np.array([0.1, 0.2, -0.1])), atol=1e-9) and np.all(ts[0] == 0)


def test_quaternion(ph):
    rng = np.random.default_rng(3)
    mats = [rot(rng.normal(size=3) * s) for s in (0.1, 1.0, 2.5, 3.1) for _ in range(5)]
    mats += [np.diag([1.0, -1, -1]), np.diag([-1.0, 1, -1]), np.diag([-1.0, -1, 1]), np.eye(3)]
    for Rm in mats:
        q = np.zeros(4)
        ph.ph_quat(P(C(Rm)), P(q))
        assert np.allclose(q, R.rotation_to_quat(Rm), atol=1e-12) and np.isclose(np.linalg.norm(q), 1.0)
        w, x, y, z = q
        back = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.allclose(back, Rm, atol=1e-12)
