"""ctypes front end of oracle/ransac_seq.cpp: the sequential scalar fp64 LO-RANSAC verifier used as the CPU arm
of bench.py (one image pair per host thread, like the reference's VerifierWorker pool).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_ransac.so")
        src = os.path.join(_HERE, "ransac_seq.cpp")
        hdr = os.path.join(_HERE, "..", "pycolmap_b200", "csrc", "geom.h")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_ransac.so"])
        _LIB = ctypes.CDLL(so)
    return _LIB


def _opts(o):
    r = o.ransac
    return np.array([o.min_num_inliers, o.min_E_F_inlier_ratio, o.max_H_inlier_ratio, o.watermark_min_inlier_ratio,
                     o.watermark_border_size, int(o.detect_watermark), int(o.force_H_use), r.max_error, r.min_inlier_ratio,
                     r.confidence, r.dyn_num_trials_multiplier, r.min_num_trials, r.max_num_trials], np.float64)


def _cam(c):
    p = [float(x) for x in c["params"]]
    model = c.get("model", 0)
    if model == 0:
        fx = fy = p[0]
        cx, cy = p[1], p[2]
    elif model == 1:
        fx, fy, cx, cy = p
    else:
        raise ValueError("oracle/ransac_seq.cpp takes the pinhole camera models only")
    return [fx, fy, cx, cy, c["width"], c["height"], int(bool(c.get("has_prior_focal_length")))]


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def estimate_two_view_geometry(cam1, points1, cam2, points2, matches=None, options=None, seed=0):
    """Same call as oracle.ransac.estimate_two_view_geometry; returns a dict(config, nE, nF, nH, inlier_matches, E, F, H,
    models_scored)."""
    from .ransac import TwoViewGeometryOptions
    o = options or TwoViewGeometryOptions()
    p1 = np.ascontiguousarray(points1, np.float64).reshape(-1, 2)
    p2 = np.ascontiguousarray(points2, np.float64).reshape(-1, 2)
    if matches is None:
        matches = np.stack([np.arange(len(p1))] * 2, 1)
    mm = np.ascontiguousarray(matches, np.uint32).reshape(-1, 2)
    out_i = np.zeros(5, np.int32)
    inl = np.zeros((max(1, len(mm)), 2), np.uint32)
    models = np.zeros(27)
    scored = ctypes.c_int64(0)
    c1, c2 = np.array(_cam(cam1), np.float64), np.array(_cam(cam2), np.float64)
    op = _opts(o)
    _lib().orc_estimate_two_view_geometry(_p(c1), _p(p1), _p(c2), _p(p2), _p(mm), ctypes.c_int64(len(mm)), _p(op),
                                          ctypes.c_uint32(seed), _p(out_i), _p(inl), _p(models), ctypes.byref(scored))
    return dict(config=int(out_i[0]), nE=int(out_i[1]), nF=int(out_i[2]), nH=int(out_i[3]),
                inlier_matches=inl[: out_i[4]].copy(), E=models[:9].reshape(3, 3), F=models[9:18].reshape(3, 3),
                H=models[18:].reshape(3, 3), models_scored=int(scored.value))


def verify_pairs(jobs, options=None, seed=0, n_threads=None):
    """jobs: list of (cam1, kpts1 [n x 2] float64, cam2, kpts2, matches [m x 2] uint32).  Returns (out [n_jobs x 5] int32:
    config, nE, nF, nH, inliers; models_scored) -- all jobs on `n_threads` host threads inside one C++ call."""
    from .ransac import TwoViewGeometryOptions
    o = options or TwoViewGeometryOptions()
    n = len(jobs)
    n_threads = n_threads or os.cpu_count() or 1
    cams1 = np.array([_cam(j[0]) for j in jobs], np.float64).reshape(-1, 7)
    cams2 = np.array([_cam(j[2]) for j in jobs], np.float64).reshape(-1, 7)
    k1 = [np.ascontiguousarray(j[1], np.float64) for j in jobs]
    k2 = [np.ascontiguousarray(j[3], np.float64) for j in jobs]
    mm = [np.ascontiguousarray(j[4], np.uint32).reshape(-1, 2) for j in jobs]
    P = ctypes.c_void_p * max(1, n)
    a1, a2, am = P(*[a.ctypes.data for a in k1]), P(*[a.ctypes.data for a in k2]), P(*[a.ctypes.data for a in mm])
    m = np.array([len(x) for x in mm], np.int64)
    out = np.zeros((n, 5), np.int32)
    scored = ctypes.c_int64(0)
    op = _opts(o)
    _lib().orc_verify_pairs(ctypes.c_int64(n), _p(cams1), a1, _p(cams2), a2, am, _p(m), _p(op), ctypes.c_uint32(seed),
                            ctypes.c_int(n_threads), _p(out), ctypes.byref(scored))
    return out, int(scored.value)
