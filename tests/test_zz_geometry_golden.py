"""Golden fixtures for the geometry stages (tests/golden/geometry_golden.npz, made by
tests/golden/make_geometry_golden.py from the numpy oracle; the reference holds none): the oracle itself, the host
builds of csrc/camera_models.h and csrc/pose.h, and -- on a GPU -- the CUDA kernels must reproduce them.  (The file name sorts after the GPU-validated test files: its GPU tests drive kernels
written after the last GPU session.)"""
import ctypes
import os

import numpy as np
import pytest

from oracle import ransac as R
from test_camera_models import camlib, P as CP, _params12  # noqa: F401  (fixture + helpers)
from test_pose_math import ph, P, C  # noqa: F401

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geometry_golden.npz"))
MODELS = sorted(int(k[3:-7]) for k in G.files if k.endswith("_params"))


def test_fixture_covers_all_models():
    assert MODELS == list(range(11))


@pytest.mark.parametrize("model", MODELS)
def test_camera_golden_oracle_and_host_header(camlib, model):
    cam = dict(model=model, params=G[f"cam{model}_params"].tolist())
    px, norm = G[f"cam{model}_px"], G[f"cam{model}_norm"]
    assert np.allclose(R.img_from_cam(cam, G["cam_uv"]), px, rtol=0, atol=1e-9)
    assert np.allclose(R.cam_from_img(cam, px), norm, rtol=0, atol=1e-12)
    assert np.abs(norm - G["cam_uv"]).max() < 1e-9
    out = np.zeros_like(norm)
    camlib.ch_cam_from_img(model, CP(_params12(cam["params"])), CP(px), len(px), CP(out))
    assert np.abs(out - norm).max() < 1e-9


def test_pose_golden_oracle_and_host_header(ph):
    E, x1, x2 = C(G["pose_E"]), C(G["pose_x1"]), C(G["pose_x2"])
    Ro, to, Xo = R.pose_from_essential_matrix(E, x1, x2)
    assert np.allclose(Ro, G["pose_R"], atol=1e-12) and np.allclose(to, G["pose_t"], atol=1e-12)
    assert len(Xo) == int(G["pose_n_front"]) == 42                         # 6 of 48 points are behind the first camera
    R1, R2, t = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros(3)
    ph.ph_decompose_E(P(E), P(R1), P(R2), P(t))
    Rs, ts = C(np.stack([R1, R2, R1, R2])), C(np.stack([t, t, -t, -t]))
    n_front, tri = ctypes.c_int(0), ctypes.c_double(0)
    best = ph.ph_select_pose(P(Rs), P(ts), 4, P(x1), P(x2), len(x1), ctypes.byref(n_front), ctypes.byref(tri))
    assert np.allclose(Rs[best], G["pose_R"], atol=1e-9) and np.allclose(ts[best], G["pose_t"], atol=1e-9)
    assert n_front.value == 42 and abs(tri.value - float(G["pose_tri"])) < 1e-9
    q = np.zeros(4)
    ph.ph_quat(P(C(Rs[best])), P(q))
    assert np.allclose(q, G["pose_q"], atol=1e-9)
    Rh, th, nh = np.zeros((4, 3, 3)), np.zeros((4, 3)), np.zeros((4, 3))
    assert ph.ph_decompose_H(P(C(G["homog_H"])), P(C(G["homog_K1"])), P(C(G["homog_K2"])), P(Rh), P(th), P(nh)) == 4
    assert np.allclose(Rh, G["homog_R"], atol=1e-8) and np.allclose(th, G["homog_t"], atol=1e-8)
    assert np.allclose(nh, G["homog_n"], atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("model", MODELS)
def test_gpu_cam_from_img_reproduces_golden(model):
    import pycolmap_b200 as pb
    cam = dict(model=model, width=1600, height=1200, params=G[f"cam{model}_params"].tolist())
    assert np.abs(pb.cam_from_img(cam, G[f"cam{model}_px"]) - G[f"cam{model}_norm"]).max() < 1e-9


@pytest.mark.gpu
def test_gpu_pose_reproduces_golden():
    import pycolmap_b200 as pb
    cam = dict(model=0, width=2, height=2, params=[1.0, 0.0, 0.0])           # identity calibration: points are normalised
    idx = np.arange(len(G["pose_x1"]), dtype=np.uint32)
    g = pb.TwoViewGeometry("CALIBRATED", E=G["pose_E"], inlier_matches=np.stack([idx, idx], 1))
    assert pb.estimate_two_view_geometry_pose(cam, G["pose_x1"], cam, G["pose_x2"], g) is True
    q = g.cam2_from_cam1.rotation.quat                                       # (x, y, z, w); the fixture holds (w, x, y, z)
    assert np.allclose([q[3], q[0], q[1], q[2]], G["pose_q"], atol=1e-9)
    assert np.allclose(g.cam2_from_cam1.translation, G["pose_t"], atol=1e-9) and abs(g.tri_angle - float(G["pose_tri"])) < 1e-9
