#!/usr/bin/env python
"""Regenerates tests/golden/geometry_golden.npz: camera models (row V9) and relative pose (rows B7 / B9).

As for match_golden.npz: the reference holds no vectors for this path (parity unpinned), so
these pin the ORACLE's restatement (oracle/ransac.py: numpy, np.linalg.svd) on small seeded inputs -- the host
builds of csrc/camera_models.h and csrc/pose.h (CPU tests) and the CUDA kernels (GPU tests) must reproduce them.
Run from the repo root:  python tests/golden/make_geometry_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ransac as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CAMS = {
    0: [1200.0, 800.0, 600.0],
    1: [1200.0, 1190.0, 800.0, 600.0],
    2: [1200.0, 800.0, 600.0, -0.12],
    3: [1200.0, 800.0, 600.0, -0.12, 0.03],
    4: [1200.0, 1190.0, 800.0, 600.0, -0.12, 0.03, 1e-3, -2e-3],
    5: [700.0, 705.0, 800.0, 600.0, 0.05, -0.01, 0.003, -0.001],
    6: [1200.0, 1190.0, 800.0, 600.0, -0.12, 0.03, 1e-3, -2e-3, 0.002, 0.01, -0.004, 0.0005],
    7: [1200.0, 1190.0, 800.0, 600.0, 0.7],
    8: [700.0, 800.0, 600.0, 0.05],
    9: [700.0, 800.0, 600.0, 0.05, -0.01],
    10: [700.0, 705.0, 800.0, 600.0, 0.05, -0.01, 1e-3, -2e-3, 0.003, -0.001, 2e-3, -1e-3],
}


def rot(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def main():
    rng = np.random.default_rng(20260923)
    out = {}
    uv = rng.uniform(-0.5, 0.5, (64, 2))
    uv[0] = 0.0
    out["cam_uv"] = uv
    for m, p in CAMS.items():
        cam = dict(model=m, params=p)
        px = R.img_from_cam(cam, uv)
        out[f"cam{m}_params"] = np.array(p)
        out[f"cam{m}_px"] = px
        out[f"cam{m}_norm"] = R.cam_from_img(cam, px)          # == uv up to the solver tolerance
    # relative pose from an essential matrix: planted pose, 48 normalised correspondences, 6 of them behind the camera
    Rm, t = rot(np.array([0.1, -0.2, 0.05])), np.array([0.8, -0.1, 0.3])
    t /= np.linalg.norm(t)
    X = np.c_[rng.uniform(-2, 2, 48), rng.uniform(-2, 2, 48), rng.uniform(4, 9, 48)]
    X[:6, 2] *= -1.0
    x1 = X[:, :2] / X[:, 2:]
    Xc = X @ Rm.T + t
    x2 = Xc[:, :2] / Xc[:, 2:]
    E = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ Rm * 1.7
    Re, te, Xe = R.pose_from_essential_matrix(E, x1, x2)
    out.update(pose_E=E, pose_x1=x1, pose_x2=x2, pose_R=Re, pose_t=te, pose_n_front=np.array(len(Xe)),
               pose_tri=np.array(R.median(R.triangulation_angles(np.zeros(3), -Re.T @ te, Xe))),
               pose_q=R.rotation_to_quat(Re))
    # homography decomposition: plane n.X = d seen from two calibrated cameras
    K1 = np.array([[1200.0, 0, 800], [0, 1190.0, 600], [0, 0, 1]])
    K2 = np.array([[900.0, 0, 640], [0, 905.0, 480], [0, 0, 1]])
    Rh, th_ = rot(np.array([-0.05, 0.15, 0.02])), np.array([0.3, 0.1, -0.2])
    n = np.array([0.1, -0.2, 1.0])
    n /= np.linalg.norm(n)
    H = K2 @ (Rh + np.outer(th_, n) / 6.0) @ np.linalg.inv(K1) * -2.5
    cands = R.decompose_homography_matrix(H, K1, K2)
    out.update(homog_H=H, homog_K1=np.array([1200.0, 1190.0, 800.0, 600.0]), homog_K2=np.array([900.0, 905.0, 640.0, 480.0]),
               homog_R=np.stack([c[0] for c in cands]), homog_t=np.stack([c[1] for c in cands]),
               homog_n=np.stack([c[2] for c in cands]))
    np.savez_compressed(os.path.join(HERE, "geometry_golden.npz"), **out)
    print("wrote geometry_golden.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
