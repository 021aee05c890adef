"""CPU: camera models.  pycolmap_b200/csrc/camera_models.h -- the header the CUDA
kernels compile -- is built for the host and checked against the numpy oracle (oracle.ransac.cam_from_img
/ img_from_cam) and through the model-independent round trip ImgFromCam(CamFromImg(x)) == x."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import ransac as R

HERE = os.path.dirname(os.path.abspath(__file__))

# COLMAP model id -> a realistic parameter vector (U:sensor/models.h parameter orders)
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


@pytest.fixture(scope="module")
def camlib():
    so = os.path.join(HERE, "helpers", "libcamera_host.so")
    src = os.path.join(HERE, "helpers", "camera_host.cpp")
    hdr = os.path.join(HERE, "..", "pycolmap_b200", "csrc", "camera_models.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=off", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.ch_mean_focal_length.restype = ctypes.c_double
    return lib


def P(a):
    return np.ascontiguousarray(a, np.float64).ctypes.data_as(ctypes.c_void_p)


def _params12(p):
    return np.array(list(p) + [0.0] * (12 - len(p)), np.float64)


@pytest.mark.parametrize("model", sorted(CAMS))
def test_cam_from_img_matches_oracle_and_round_trips(camlib, model):
    rng = np.random.default_rng(model)
    cam = dict(model=model, params=CAMS[model])
    p12 = _params12(CAMS[model])
    assert camlib.ch_num_params(model) == len(CAMS[model]) == R.CAMERA_NUM_PARAMS[model]
    assert camlib.ch_mean_focal_length(model, P(p12)) == R.mean_focal_length(cam)
    uv = rng.uniform(-0.55, 0.55, (2000, 2))
    uv[0] = 0.0                                          # the principal point (fisheye r == 0 branch)
    px_oracle = R.img_from_cam(cam, uv)
    px = np.zeros_like(uv)
    camlib.ch_img_from_cam(model, P(p12), P(uv), len(uv), P(px))
    assert np.allclose(px, px_oracle, rtol=0, atol=1e-9)             # same distortion function
    back = np.zeros_like(uv)
    camlib.ch_cam_from_img(model, P(p12), P(px), len(px), P(back))
    assert np.abs(back - uv).max() < 1e-9                              # inverse of ImgFromCam
    assert np.abs(back - R.cam_from_img(cam, px)).max() < 1e-9         # and equal to the oracle's inverse
    if model in (0, 1):                                                # pinhole: closed form, exact
        f = np.array([p12[0], p12[0]] if model == 0 else [p12[0], p12[1]])
        c = p12[1:3] if model == 0 else p12[2:4]
        assert np.array_equal(back, (px - c) / f)


def test_known_distortion_values():
    """Hand-computed values of the published model definitions."""
    cam = dict(model=2, params=[1000.0, 500.0, 400.0, 0.1])             # SIMPLE_RADIAL: u (1 + k r^2)
    assert np.allclose(R.img_from_cam(cam, [[0.3, 0.4]]), [[500 + 1000 * 0.3 * 1.025, 400 + 1000 * 0.4 * 1.025]])
    cam = dict(model=4, params=[1000.0, 1000.0, 0.0, 0.0, 0.0, 0.0, 0.01, 0.02])   # OPENCV tangential only
    u, v = 0.3, 0.4
    du = 2 * 0.01 * u * v + 0.02 * (0.25 + 2 * u * u)
    dv = 2 * 0.02 * u * v + 0.01 * (0.25 + 2 * v * v)
    assert np.allclose(R.img_from_cam(cam, [[u, v]]), [[1000 * (u + du), 1000 * (v + dv)]])
    cam = dict(model=8, params=[500.0, 0.0, 0.0, 0.0])                   # equidistant fisheye: r -> atan(r)
    assert np.allclose(R.img_from_cam(cam, [[1.0, 0.0]]), [[500 * np.pi / 4, 0.0]])
    cam = dict(model=7, params=[1000.0, 1000.0, 0.0, 0.0, 0.8])         # FOV: r_d = atan(2 r tan(w / 2)) / w
    assert np.allclose(R.img_from_cam(cam, [[0.5, 0.0]]), [[1000 * np.arctan(2 * 0.5 * np.tan(0.4)) / 0.8, 0.0]])
    cam = dict(model=10, params=[500.0, 500.0, 0.0, 0.0] + [0.0] * 6 + [0.01, -0.02])   # thin prism terms on theta
    th = np.arctan(1.0)
    assert np.allclose(R.img_from_cam(cam, [[1.0, 0.0]]), [[500 * (th + 0.01 * th * th), 500 * (-0.02 * th * th)]])
    for bad in (11, -1):
        with pytest.raises(ValueError):
            R.cam_from_img(dict(model=bad, params=[1.0] * 5), np.zeros((1, 2)))
