"""Planted two-view scenes for the verifier tests."""
import numpy as np

CAM = dict(model=0, width=1600, height=1200, params=[1200.0, 800.0, 600.0], has_prior_focal_length=1)
CAM_NOPRIOR = dict(CAM, has_prior_focal_length=0)


def _skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def _rot(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    K = _skew(w / th)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def two_view_scene(rng, n, outlier_ratio=0.3, kind="general", noise=0.0):
    """Returns p1, p2 [n x 2] pixel points and the planted inlier mask.

    kind: 'general' (3-D points, rotation + translation), 'planar' (points on a plane),
          'rotation' (pure rotation -> panoramic)."""
    f, cx, cy = CAM["params"]
    R = _rot(rng.normal(size=3) * 0.15)
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    if kind == "rotation":
        t = np.zeros(3)
    X = np.c_[rng.uniform(-2.5, 2.5, n), rng.uniform(-1.8, 1.8, n), rng.uniform(4, 9, n)]
    if kind == "planar":
        X[:, 2] = 6.0 + 0.2 * X[:, 0] - 0.1 * X[:, 1]
    x1 = X[:, :2] / X[:, 2:]
    Xc = X @ R.T + t
    x2 = Xc[:, :2] / Xc[:, 2:]
    p1 = x1 * f + [cx, cy]
    p2 = x2 * f + [cx, cy]
    if noise > 0:
        p1 = p1 + rng.normal(0, noise, p1.shape)
        p2 = p2 + rng.normal(0, noise, p2.shape)
    out = rng.random(n) < outlier_ratio
    p2 = p2.copy()
    p2[out] = np.c_[rng.uniform(0, 1600, out.sum()), rng.uniform(0, 1200, out.sum())]
    return p1, p2, ~out
