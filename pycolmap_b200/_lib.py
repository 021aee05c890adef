"""ctypes binding of the C ABI declared in include/b200match.h (libb200match.so).

This is the stub a maintainer of the reference would add on the Python side; the pybind11
flavour is shown in INTEGRATION.md.  There is no CPU fallback: if the shared library is missing
or no B200 is visible the calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2M_LIB") or os.path.join(_HERE, "libb200match.so")  # B2M_LIB: experiment builds only

c_i32, c_i64, c_u32, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
c_f32, c_f64 = ctypes.c_float, ctypes.c_double

B2M_OK, B2M_EINVAL, B2M_ENODEV, B2M_ECUDA, B2M_ENOMEM, B2M_ESTOPPED, B2M_ESTATE = 0, -1, -2, -3, -4, -5, -6


class DeviceCfg(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("device", c_i32), ("seed", c_u64), ("pair_batch", c_i32),
                ("reserved", c_i32)]


class SiftOpts(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("max_ratio", c_f32), ("max_distance", c_f32), ("cross_check", c_i32),
                ("max_num_matches", c_i32), ("guided_matching", c_i32)]


class RansacOpts(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("min_num_trials", c_i32), ("max_num_trials", c_i32), ("reserved", c_i32),
                ("max_error", c_f64), ("min_inlier_ratio", c_f64), ("confidence", c_f64),
                ("dyn_num_trials_multiplier", c_f64)]


class TvgOpts(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("min_num_inliers", c_i32), ("min_E_F_inlier_ratio", c_f64),
                ("max_H_inlier_ratio", c_f64), ("watermark_min_inlier_ratio", c_f64),
                ("watermark_border_size", c_f64), ("detect_watermark", c_i32),
                ("multiple_ignore_watermark", c_i32), ("force_H_use", c_i32), ("compute_relative_pose", c_i32),
                ("multiple_models", c_i32), ("reserved", c_i32), ("ransac", RansacOpts)]


class Camera(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("model", c_i32), ("width", c_i32), ("height", c_i32),
                ("has_prior_focal_length", c_i32), ("reserved", c_i32), ("params", c_f64 * 12)]


class PairView(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("image1", c_i32), ("image2", c_i32), ("config", c_i32),
                ("n_matches", c_i64), ("matches", ctypes.POINTER(c_u32)), ("n_inliers", c_i64),
                ("inlier_matches", ctypes.POINTER(c_u32)), ("E", c_f64 * 9), ("F", c_f64 * 9), ("H", c_f64 * 9),
                ("qvec", c_f64 * 4), ("tvec", c_f64 * 3), ("tri_angle", c_f64), ("pose_valid", c_i32),
                ("reserved", c_i32)]


class TvgResult(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("config", c_i32), ("n_inliers", c_i64), ("E", c_f64 * 9),
                ("F", c_f64 * 9), ("H", c_f64 * 9), ("nE", c_i32), ("nF", c_i32), ("nH", c_i32),
                ("pose_valid", c_i32), ("qvec", c_f64 * 4), ("tvec", c_f64 * 3), ("tri_angle", c_f64)]


class TvgProblem(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("reserved", c_i32), ("cam1", Camera), ("cam2", Camera),
                ("points1", ctypes.c_void_p), ("n1", c_i64), ("points2", ctypes.c_void_p), ("n2", c_i64),
                ("matches", ctypes.c_void_p), ("m", c_i64)]


class Stats(ctypes.Structure):
    _fields_ = [("struct_size", c_u32), ("reserved", c_u32), ("kernel_launches", c_u64), ("match_tiles", c_u64),
                ("last_match_ms", c_f64), ("last_verify_ms", c_f64), ("last_total_ms", c_f64), ("last_k1_ms", c_f64),
                ("last_k1_launches", c_u64), ("k1_dir1_mode", c_u64)]


# every symbol include/b200match.h declares
EXPORTS = [
    "b2m_abi_version", "b2m_create", "b2m_destroy", "b2m_last_error", "b2m_request_stop",
    "b2m_sift_opts_default", "b2m_ransac_opts_default", "b2m_tvg_opts_default", "b2m_match_pair",
    "b2m_set_images", "b2m_set_images_device", "b2m_match_pairs", "b2m_match_verify", "b2m_results_num_pairs",
    "b2m_results_total_matches", "b2m_results_num_verified", "b2m_results_get", "b2m_results_free", "b2m_estimate_two_view_geometry",
    "b2m_estimate_two_view_geometry_batch", "b2m_estimate_two_view_geometry_pose",
    "b2m_ransac_model", "b2m_cam_from_img", "b2m_squared_sampson_error", "b2m_get_stats", "b2m_reset_stats",
]

_lib = None


def load():
    """dlopen libb200match.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(pycolmap_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.c_void_p
    lib.b2m_abi_version.restype = ctypes.c_int
    lib.b2m_create.argtypes = [ctypes.POINTER(DeviceCfg), ctypes.POINTER(P)]
    lib.b2m_destroy.argtypes = [P]
    lib.b2m_destroy.restype = None
    lib.b2m_last_error.argtypes = [P]
    lib.b2m_last_error.restype = ctypes.c_char_p
    lib.b2m_request_stop.argtypes = [P]
    for f, T in (("b2m_sift_opts_default", SiftOpts), ("b2m_ransac_opts_default", RansacOpts),
                 ("b2m_tvg_opts_default", TvgOpts)):
        getattr(lib, f).argtypes = [ctypes.POINTER(T)]
        getattr(lib, f).restype = None
    lib.b2m_match_pair.argtypes = [P, P, c_i32, P, c_i32, ctypes.POINTER(SiftOpts), P, c_i64,
                                   ctypes.POINTER(c_i64)]
    lib.b2m_set_images.argtypes = [P, c_i32, P, P, P, P]
    lib.b2m_set_images_device.argtypes = [P, c_i32, P, P, P, P]
    lib.b2m_match_pairs.argtypes = [P, P, c_i64, ctypes.POINTER(SiftOpts), ctypes.POINTER(TvgOpts),
                                    ctypes.POINTER(P)]
    lib.b2m_results_num_pairs.argtypes = [P]
    lib.b2m_results_num_pairs.restype = c_i64
    lib.b2m_results_total_matches.argtypes = [P]
    lib.b2m_results_total_matches.restype = c_i64
    lib.b2m_results_num_verified.argtypes = [P]
    lib.b2m_results_num_verified.restype = c_i64
    lib.b2m_results_get.argtypes = [P, c_i64, ctypes.POINTER(PairView)]
    lib.b2m_results_free.argtypes = [P]
    lib.b2m_results_free.restype = None
    lib.b2m_estimate_two_view_geometry.argtypes = [P, ctypes.POINTER(Camera), P, c_i64, ctypes.POINTER(Camera), P,
                                                   c_i64, P, c_i64, ctypes.POINTER(TvgOpts),
                                                   ctypes.POINTER(TvgResult), P]
    lib.b2m_estimate_two_view_geometry_batch.argtypes = [P, ctypes.POINTER(TvgProblem), c_i64,
                                                         ctypes.POINTER(TvgOpts), ctypes.POINTER(TvgResult), P]
    lib.b2m_estimate_two_view_geometry_pose.argtypes = [P, ctypes.POINTER(Camera), P, c_i64, ctypes.POINTER(Camera), P,
                                                        c_i64, P, c_i64, ctypes.POINTER(TvgResult)]
    lib.b2m_ransac_model.argtypes = [P, c_i32, P, P, c_i64, ctypes.POINTER(RansacOpts), P, P,
                                     ctypes.POINTER(c_i64), ctypes.POINTER(c_i32)]
    lib.b2m_squared_sampson_error.argtypes = [P, P, P, c_i64, P, P]
    lib.b2m_cam_from_img.argtypes = [P, ctypes.POINTER(Camera), P, c_i64, P]
    lib.b2m_get_stats.argtypes = [P, ctypes.POINTER(Stats)]
    lib.b2m_reset_stats.argtypes = [P]
    _lib = lib
    return lib


class B2MError(RuntimeError):
    pass


def _raise(lib, ctx, rc):
    msg = lib.b2m_last_error(ctx)
    msg = msg.decode() if msg else ""
    if rc == B2M_EINVAL:
        raise ValueError(msg)  # reference: THROW_CHECK -> ValueError (R:log_exceptions.h:114-147)
    if rc == B2M_ESTOPPED:
        raise KeyboardInterrupt(msg)  # reference: PyInterrupt (R:helpers.h:306-347)
    if rc == B2M_ENOMEM:
        raise MemoryError(msg)
    raise B2MError(f"b200match error {rc}: {msg}")


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    """Owns one b2m_ctx (one per process and GPU)."""

    def __init__(self, device=0, seed=0, pair_batch=0):
        self.lib = load()
        cfg = DeviceCfg(ctypes.sizeof(DeviceCfg), int(device), int(seed), int(pair_batch), 0)
        h = ctypes.c_void_p()
        rc = self.lib.b2m_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != B2M_OK:
            _raise(self.lib, None, rc)
        self.h = h
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2m_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != B2M_OK:
            _raise(self.lib, self.h, rc)

    # -- options ------------------------------------------------------------------------
    def sift_opts(self, **kw):
        o = SiftOpts()
        self.lib.b2m_sift_opts_default(ctypes.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def tvg_opts(self, ransac=None, **kw):
        o = TvgOpts()
        self.lib.b2m_tvg_opts_default(ctypes.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        for k, v in (ransac or {}).items():
            setattr(o.ransac, k, v)
        return o

    def ransac_opts(self, **kw):
        o = RansacOpts()
        self.lib.b2m_ransac_opts_default(ctypes.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    # -- single pair --------------------------------------------------------------------
    def match_pair(self, d1, d2, opts=None):
        d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 128)
        d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 128)
        opts = opts or self.sift_opts()
        out = np.zeros((max(1, len(d1)), 2), np.uint32)
        n = c_i64(0)
        self.check(self.lib.b2m_match_pair(self.h, ptr(d1), len(d1), ptr(d2), len(d2), ctypes.byref(opts), ptr(out),
                                           len(out), ctypes.byref(n)))
        return out[:n.value].copy()

    # -- image set ----------------------------------------------------------------------
    @staticmethod
    def make_cameras(cams):
        arr = (Camera * len(cams))()
        for i, c in enumerate(cams):
            arr[i].struct_size = ctypes.sizeof(Camera)
            arr[i].model = int(c.get("model", 0))
            arr[i].width = int(c["width"])
            arr[i].height = int(c["height"])
            arr[i].has_prior_focal_length = int(c.get("has_prior_focal_length", 0))
            for k, v in enumerate(c["params"]):
                arr[i].params[k] = float(v)
        return arr

    def set_images(self, descs, kpts=None, cams=None):
        n = len(descs)
        descs = [np.ascontiguousarray(d, np.uint8).reshape(-1, 128) for d in descs]
        nfeat = np.array([len(d) for d in descs], np.int32)
        dptr = (ctypes.c_void_p * max(1, n))(*[d.ctypes.data for d in descs])
        kp, kptr = None, None
        if kpts is not None:
            kp = [np.ascontiguousarray(k, np.float32).reshape(-1, 2) for k in kpts]
            assert all(len(a) == len(b) for a, b in zip(kp, descs))
            kptr = (ctypes.c_void_p * max(1, n))(*[k.ctypes.data for k in kp])
        carr = self.make_cameras(cams) if cams is not None else None
        self.check(self.lib.b2m_set_images(self.h, n, ptr(nfeat), dptr, kptr, carr))

    def set_images_device(self, n_feat, dev_desc_ptr, dev_kpts_ptr=None, cams=None):
        nfeat = np.ascontiguousarray(n_feat, np.int32)
        carr = self.make_cameras(cams) if cams is not None else None
        self.check(self.lib.b2m_set_images_device(self.h, len(nfeat), ptr(nfeat), ctypes.c_void_p(dev_desc_ptr),
                                                  ctypes.c_void_p(dev_kpts_ptr) if dev_kpts_ptr else None, carr))

    def match_pairs(self, pairs, sift=None, tvg=None):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1, 2)
        sift = sift or self.sift_opts()
        res = ctypes.c_void_p()
        self.check(self.lib.b2m_match_pairs(self.h, ptr(pairs), len(pairs), ctypes.byref(sift),
                                            ctypes.byref(tvg) if tvg is not None else None, ctypes.byref(res)))
        return Results(self.lib, res)

    # -- estimators ----------------------------------------------------------------------
    def estimate_two_view_geometry(self, cam1, points1, cam2, points2, matches=None, opts=None):
        """Returns (TvgResult, inlier_matches [k x 2] uint32)."""
        p1 = np.ascontiguousarray(points1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(points2, np.float64).reshape(-1, 2)
        opts = opts or self.tvg_opts()
        cams = self.make_cameras([cam1, cam2])
        if matches is not None:
            matches = np.ascontiguousarray(matches, np.uint32).reshape(-1, 2)
            m = len(matches)
        else:
            m = len(p1)
        out = TvgResult()
        inl = np.zeros((max(1, m), 2), np.uint32)
        self.check(self.lib.b2m_estimate_two_view_geometry(
            self.h, ctypes.byref(cams[0]), ptr(p1), len(p1), ctypes.byref(cams[1]), ptr(p2), len(p2),
            ptr(matches) if matches is not None else None, m, ctypes.byref(opts), ctypes.byref(out), ptr(inl)))
        return out, inl[:out.n_inliers].copy()

    def estimate_two_view_geometry_pose(self, cam1, points1, cam2, points2, config, E, H, inlier_matches):
        """Returns the TvgResult with qvec / tvec / tri_angle / pose_valid / (possibly resolved) config."""
        p1 = np.ascontiguousarray(points1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(points2, np.float64).reshape(-1, 2)
        inl = np.ascontiguousarray(inlier_matches, np.uint32).reshape(-1, 2)
        cams = self.make_cameras([cam1, cam2])
        g = TvgResult()
        g.struct_size = ctypes.sizeof(TvgResult)
        g.config = int(config)
        for k, v in enumerate(np.asarray(E, np.float64).reshape(9)):
            g.E[k] = v
        for k, v in enumerate(np.asarray(H, np.float64).reshape(9)):
            g.H[k] = v
        self.check(self.lib.b2m_estimate_two_view_geometry_pose(
            self.h, ctypes.byref(cams[0]), ptr(p1), len(p1), ctypes.byref(cams[1]), ptr(p2), len(p2), ptr(inl), len(inl),
            ctypes.byref(g)))
        return g

    def ransac_model(self, kind, points1, points2, opts=None):
        """kind 0 = E (normalised points), 1 = F, 2 = H.  Returns dict or None (reference: None on failure)."""
        p1 = np.ascontiguousarray(points1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(points2, np.float64).reshape(-1, 2)
        opts = opts or self.ransac_opts()
        model = np.zeros(9, np.float64)
        mask = np.zeros(max(1, len(p1)), np.uint8)
        n, ok = c_i64(0), c_i32(0)
        self.check(self.lib.b2m_ransac_model(self.h, int(kind), ptr(p1), ptr(p2), len(p1), ctypes.byref(opts),
                                             ptr(model), ptr(mask), ctypes.byref(n), ctypes.byref(ok)))
        if not ok.value:
            return None
        return {"model": model.reshape(3, 3), "num_inliers": int(n.value), "inliers": mask[:len(p1)].astype(bool)}

    def cam_from_img(self, cam, points):
        """Camera::CamFromImg on an [n x 2] point list (all supported camera models)."""
        p = np.ascontiguousarray(points, np.float64).reshape(-1, 2)
        out = np.zeros_like(p)
        cams = self.make_cameras([cam])
        self.check(self.lib.b2m_cam_from_img(self.h, ctypes.byref(cams[0]), ptr(p), len(p), ptr(out)))
        return out

    def squared_sampson_error(self, points1, points2, E):
        p1 = np.ascontiguousarray(points1, np.float64).reshape(-1, 2)
        p2 = np.ascontiguousarray(points2, np.float64).reshape(-1, 2)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        out = np.zeros(len(p1), np.float64)
        self.check(self.lib.b2m_squared_sampson_error(self.h, ptr(p1), ptr(p2), len(p1), ptr(E), ptr(out)))
        return out

    def stats(self):
        s = Stats()
        self.check(self.lib.b2m_get_stats(self.h, ctypes.byref(s)))
        return s


class Results:
    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def __len__(self):
        return int(self.lib.b2m_results_num_pairs(self.h))

    @property
    def total_matches(self):
        return int(self.lib.b2m_results_total_matches(self.h))

    @property
    def num_verified(self):
        return int(self.lib.b2m_results_num_verified(self.h))

    def view(self, k):
        v = PairView()
        rc = self.lib.b2m_results_get(self.h, k, ctypes.byref(v))
        if rc != B2M_OK:
            raise IndexError(k)
        return v

    def matches(self, k):
        v = self.view(k)
        if v.n_matches == 0:
            return np.zeros((0, 2), np.uint32)
        return np.ctypeslib.as_array(v.matches, shape=(v.n_matches, 2)).copy()

    def inlier_matches(self, k):
        v = self.view(k)
        if v.n_inliers == 0:
            return np.zeros((0, 2), np.uint32)
        return np.ctypeslib.as_array(v.inlier_matches, shape=(v.n_inliers, 2)).copy()

    def free(self):
        if self.h:
            self.lib.b2m_results_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
