"""CPU restatement (numpy, fp64) of the reference's two-view geometric verifier.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED: the reference
(/root/reference) forwards into COLMAP 3.9.1 which is not vendored and holds no golden vectors
for this path; this file restates the published algorithms, anchored on the reference call sites:

  estimate_two_view_geometry        <- R:estimators/two_view_geometry.h:95-151 -> U:estimators/two_view_geometry.cc
                                       EstimateTwoViewGeometry / EstimateCalibrated... / EstimateUncalibrated...  (V1, V2)
  loransac                          <- R:estimators/essential_matrix.h:48-52, fundamental_matrix.h:26-29,
                                       homography_matrix.h:25-27 -> U:optim/loransac.h LORANSAC::Estimate      (V3)
  EssentialFivePoint                <- U:estimators/essential_matrix.cc EssentialMatrixFivePointEstimator       (V4)
  FundamentalSevenPoint/EightPoint  <- U:estimators/fundamental_matrix.cc                                      (V5)
  Homography                        <- U:estimators/homography_matrix.cc                                       (V6)
  squared_sampson_error             <- R:estimators/two_view_geometry.h:161-175 -> U:estimators/utils.cc        (V7)
  detect_watermark                  <- U:estimators/two_view_geometry.cc DetectWatermark                       (V8)
  cam_from_img / cam_from_img_threshold <- R:scene/camera.h cam_from_img, cam_from_img_threshold               (V9)
  exhaustive_pairs / sequential_pairs   <- U:controllers/feature_matching.cc                                   (P1, P2)

The control flow is the sequential upstream one (one trial at a time, LO on every new best);
numpy's PCG64 stands in for upstream's thread-local mt19937, so results are statistically --
not bitwise -- comparable, exactly like two upstream runs with different thread counts.
The 5-point solver uses the action-matrix (Stewenius) formulation on purpose: it is a different
algorithm from the hidden-variable elimination the CUDA path uses, so the two check each other.
"""
import math

import numpy as np

UNDEFINED, DEGENERATE, CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC, WATERMARK, MULTIPLE = range(9)

# cost constants published for the roofline arithmetic
FLOPS_SAMPSON = 33
FLOPS_HOMOGRAPHY = 19


# ---------------------------------------------------------------------------------------------
# residuals
# ---------------------------------------------------------------------------------------------
def squared_sampson_error(points1, points2, E):
    x1, y1 = points1[:, 0], points1[:, 1]
    x2, y2 = points2[:, 0], points2[:, 1]
    Ex1_0 = E[0, 0] * x1 + E[0, 1] * y1 + E[0, 2]
    Ex1_1 = E[1, 0] * x1 + E[1, 1] * y1 + E[1, 2]
    Ex1_2 = E[2, 0] * x1 + E[2, 1] * y1 + E[2, 2]
    Etx2_0 = E[0, 0] * x2 + E[1, 0] * y2 + E[2, 0]
    Etx2_1 = E[0, 1] * x2 + E[1, 1] * y2 + E[2, 1]
    x2tEx1 = x2 * Ex1_0 + y2 * Ex1_1 + Ex1_2
    with np.errstate(divide="ignore", invalid="ignore"):
        return x2tEx1 * x2tEx1 / (Ex1_0 * Ex1_0 + Ex1_1 * Ex1_1 + Etx2_0 * Etx2_0 + Etx2_1 * Etx2_1)


def homography_residuals(points1, points2, H):
    s0, s1 = points1[:, 0], points1[:, 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        pd2 = H[2, 0] * s0 + H[2, 1] * s1 + H[2, 2]
        inv = 1.0 / pd2
        dd0 = points2[:, 0] - (H[0, 0] * s0 + H[0, 1] * s1 + H[0, 2]) * inv
        dd1 = points2[:, 1] - (H[1, 0] * s0 + H[1, 1] * s1 + H[1, 2]) * inv
        return dd0 * dd0 + dd1 * dd1


def center_and_normalize(points):
    c = points.mean(0)
    rms = math.sqrt(((points - c) ** 2).sum(1).mean())
    s = math.sqrt(2.0) / rms if rms > 0 else float("inf")
    T = np.array([[s, 0, -s * c[0]], [0, s, -s * c[1]], [0, 0, 1.0]])
    return (points - c) * s, T


def _epipolar_rows(p1, p2):
    x1, y1, x2, y2 = p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]
    o = np.ones_like(x1)
    return np.stack([x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, o], 1)


# ---------------------------------------------------------------------------------------------
# 5-point essential (action matrix)
# ---------------------------------------------------------------------------------------------
def _pmul(a, b):
    """Product of dense trivariate polynomials (coefficient tensors indexed by exponents)."""
    out = np.zeros(tuple(np.array(a.shape) + np.array(b.shape) - 1))
    for idx in np.ndindex(*a.shape):
        if a[idx] != 0.0:
            out[idx[0]:idx[0] + b.shape[0], idx[1]:idx[1] + b.shape[1], idx[2]:idx[2] + b.shape[2]] += a[idx] * b
    return out


_CUBICS = [(3, 0, 0), (2, 1, 0), (2, 0, 1), (1, 2, 0), (1, 1, 1), (1, 0, 2), (0, 3, 0), (0, 2, 1), (0, 1, 2), (0, 0, 3)]
_BASIS = [(2, 0, 0), (1, 1, 0), (1, 0, 1), (0, 2, 0), (0, 1, 1), (0, 0, 2), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 0)]


def _pad3(p):
    out = np.zeros((4, 4, 4))
    out[:p.shape[0], :p.shape[1], :p.shape[2]] = p
    return out


def five_point_from_nullspace(N):
    """N: [4, 9] rows X, Y, Z, W; E = x X + y Y + z Z + W.  Returns list of 3x3 E."""
    Ep = np.empty((3, 3), object)
    for i in range(3):
        for j in range(3):
            p = np.zeros((2, 2, 2))
            p[1, 0, 0], p[0, 1, 0], p[0, 0, 1], p[0, 0, 0] = N[0, 3 * i + j], N[1, 3 * i + j], N[2, 3 * i + j], N[3, 3 * i + j]
            Ep[i, j] = p
    eqs = []
    det = (_pmul(Ep[0, 0], _pmul(Ep[1, 1], Ep[2, 2]) - _pmul(Ep[1, 2], Ep[2, 1]))
           - _pmul(Ep[0, 1], _pmul(Ep[1, 0], Ep[2, 2]) - _pmul(Ep[1, 2], Ep[2, 0]))
           + _pmul(Ep[0, 2], _pmul(Ep[1, 0], Ep[2, 1]) - _pmul(Ep[1, 1], Ep[2, 0])))
    eqs.append(_pad3(det))
    EEt = np.empty((3, 3), object)
    for i in range(3):
        for j in range(3):
            EEt[i, j] = sum(_pmul(Ep[i, k], Ep[j, k]) for k in range(3))
    tr = EEt[0, 0] + EEt[1, 1] + EEt[2, 2]
    for i in range(3):
        for j in range(3):
            acc = np.zeros((4, 4, 4))
            for k in range(3):
                lam = EEt[i, k] - (0.5 * tr if i == k else 0.0)
                acc += _pad3(_pmul(lam, Ep[k, j]))
            eqs.append(acc)
    M = np.array([[e[m] for m in _CUBICS + _BASIS] for e in eqs])
    try:
        R = np.linalg.solve(M[:, :10], M[:, 10:])     # cubic_i = -R[i] . basis
    except np.linalg.LinAlgError:
        return []
    A = np.zeros((10, 10))
    # x * basis_i expressed in the basis
    for i, (a, b, c) in enumerate(_BASIS):
        m = (a + 1, b, c)
        if m in _BASIS:
            A[i, _BASIS.index(m)] = 1.0
        else:
            A[i, :] = -R[_CUBICS.index(m)]
    w, V = np.linalg.eig(A)
    out = []
    for k in range(10):
        if abs(w[k].imag) > 1e-9 * max(1.0, abs(w[k].real)):
            continue
        v = V[:, k].real
        if abs(v[9]) < 1e-14 * np.abs(v).max():
            continue
        x, y, z = v[6] / v[9], v[7] / v[9], v[8] / v[9]
        out.append((x * N[0] + y * N[1] + z * N[2] + N[3]).reshape(3, 3))
    return out


class EssentialFivePoint:
    kMinNumSamples = 5

    @staticmethod
    def estimate(p1, p2):
        Q = _epipolar_rows(p1, p2)
        _, _, Vt = np.linalg.svd(Q, full_matrices=True)
        return five_point_from_nullspace(Vt[5:9])

    residuals = staticmethod(squared_sampson_error)


class FundamentalSevenPoint:
    kMinNumSamples = 7

    @staticmethod
    def estimate(p1, p2):
        A = _epipolar_rows(p1, p2)
        _, _, Vt = np.linalg.svd(A, full_matrices=True)
        f1, f2 = Vt[7].reshape(3, 3), Vt[8].reshape(3, 3)
        # det(l f1 + (1-l) f2) sampled at 4 points -> cubic coefficients
        ls = np.array([-1.0, 0.0, 1.0, 2.0])
        d = [np.linalg.det(l * f1 + (1 - l) * f2) for l in ls]
        c = np.linalg.solve(np.vander(ls, 4), d)
        out = []
        for r in np.roots(c):
            if abs(r.imag) < 1e-10:
                out.append(r.real * f1 + (1 - r.real) * f2)
        return out

    residuals = staticmethod(squared_sampson_error)


class FundamentalEightPoint:
    kMinNumSamples = 8

    @staticmethod
    def estimate(p1, p2):
        n1, T1 = center_and_normalize(p1)
        n2, T2 = center_and_normalize(p2)
        A = _epipolar_rows(n1, n2)
        _, _, Vt = np.linalg.svd(A, full_matrices=True)
        F = Vt[8].reshape(3, 3)
        U, S, Vt2 = np.linalg.svd(F)
        S[2] = 0.0
        return [T2.T @ (U @ np.diag(S) @ Vt2) @ T1]

    residuals = staticmethod(squared_sampson_error)


class Homography:
    kMinNumSamples = 4

    @staticmethod
    def estimate(p1, p2):
        n1, T1 = center_and_normalize(p1)
        n2, T2 = center_and_normalize(p2)
        N = len(p1)
        A = np.zeros((2 * N, 9))
        s0, s1, d0, d1 = n1[:, 0], n1[:, 1], n2[:, 0], n2[:, 1]
        A[0::2, 0], A[0::2, 1], A[0::2, 2] = -s0, -s1, -1
        A[0::2, 6], A[0::2, 7], A[0::2, 8] = s0 * d0, s1 * d0, d0
        A[1::2, 3], A[1::2, 4], A[1::2, 5] = -s0, -s1, -1
        A[1::2, 6], A[1::2, 7], A[1::2, 8] = s0 * d1, s1 * d1, d1
        if not np.isfinite(A).all():
            return []
        _, _, Vt = np.linalg.svd(A, full_matrices=True)
        H = Vt[8].reshape(3, 3)
        return [np.linalg.inv(T2) @ H @ T1]

    residuals = staticmethod(homography_residuals)


class Translation2D:
    """U:estimators/translation_transform.h (used by DetectWatermark)."""
    kMinNumSamples = 1

    @staticmethod
    def estimate(p1, p2):
        return [(p2 - p1).mean(0)]

    @staticmethod
    def residuals(p1, p2, t):
        d = p2 - (p1 + t)
        return (d * d).sum(1)


# ---------------------------------------------------------------------------------------------
# LO-RANSAC
# ---------------------------------------------------------------------------------------------
class RansacOptions:
    def __init__(self, max_error=4.0, min_inlier_ratio=0.25, confidence=0.999, dyn_num_trials_multiplier=3.0,
                 min_num_trials=100, max_num_trials=10000):
        self.max_error, self.min_inlier_ratio, self.confidence = max_error, min_inlier_ratio, confidence
        self.dyn_num_trials_multiplier, self.min_num_trials, self.max_num_trials = (
            dyn_num_trials_multiplier, min_num_trials, max_num_trials)

    def copy(self, **kw):
        o = RansacOptions(**self.__dict__)
        o.__dict__.update(kw)
        return o


def compute_num_trials(num_inliers, num_samples, confidence, multiplier, k_min):
    ratio = num_inliers / float(num_samples)
    nom = 1.0 - confidence
    if nom <= 0:
        return float("inf")
    denom = 1.0 - ratio ** k_min
    if denom <= 0:
        return 1
    if denom == 1.0:
        return float("inf")
    return math.ceil(math.log(nom) / math.log(denom) * multiplier)


class Report:
    def __init__(self):
        self.success, self.num_trials, self.num_inliers, self.residual_sum = False, 0, 0, float("inf")
        self.model, self.inlier_mask = None, None
        self.num_models_scored = 0


def _support(res, max_residual):
    with np.errstate(invalid="ignore"):
        m = res <= max_residual
    return int(m.sum()), float(res[m].sum())


def _better(a, b):
    return a[0] > b[0] or (a[0] == b[0] and a[1] < b[1])


def loransac(est, local_est, X, Y, opt, rng):
    """LORANSAC<est, local_est>::Estimate (U:optim/loransac.h), sequential restatement."""
    rep = Report()
    n = len(X)
    rep.inlier_mask = np.zeros(n, bool)
    if n < est.kMinNumSamples:
        return rep
    # RANSAC ctor: clip max_num_trials by the trials needed at min_inlier_ratio
    max_trials = min(opt.max_num_trials,
                     compute_num_trials(int(opt.min_inlier_ratio * 100000), 100000, opt.confidence,
                                        opt.dyn_num_trials_multiplier, est.kMinNumSamples))
    dyn_max = max_trials
    max_residual = opt.max_error * opt.max_error
    best = (0, float("inf"))
    best_model, best_local = None, False
    idx = np.arange(n)
    abort = False
    trial = 0
    while trial < max_trials:
        if abort:
            trial += 1
            break
        # RandomSampler: partial Fisher-Yates on a persistent index vector
        k = est.kMinNumSamples
        for i in range(k):
            j = int(rng.integers(i, n))
            idx[i], idx[j] = idx[j], idx[i]
        s = idx[:k]
        for model in est.estimate(X[s], Y[s]):
            res = est.residuals(X, Y, model)
            rep.num_models_scored += 1
            sup = _support(res, max_residual)
            if _better(sup, best):
                best, best_model, best_local = sup, model, False
                if sup[0] > est.kMinNumSamples and sup[0] >= local_est.kMinNumSamples:
                    for _ in range(10):
                        with np.errstate(invalid="ignore"):
                            inl = res <= max_residual
                        prev = best[0]
                        best_local_res = None
                        for lm in local_est.estimate(X[inl], Y[inl]):
                            lres = local_est.residuals(X, Y, lm)
                            rep.num_models_scored += 1
                            lsup = _support(lres, max_residual)
                            if _better(lsup, best):
                                best, best_model, best_local, best_local_res = lsup, lm, True, lres
                        if best[0] <= prev:
                            break
                        res = best_local_res
                dyn_max = compute_num_trials(best[0], n, opt.confidence, opt.dyn_num_trials_multiplier,
                                             est.kMinNumSamples)
            if trial >= dyn_max and trial >= opt.min_num_trials:
                abort = True
                break
        trial += 1
    rep.num_trials = trial
    rep.num_inliers, rep.residual_sum, rep.model = best[0], best[1], best_model
    if best[0] < est.kMinNumSamples or best_model is None:
        return rep
    rep.success = True
    res = (local_est if best_local else est).residuals(X, Y, best_model)
    with np.errstate(invalid="ignore"):
        rep.inlier_mask = res <= max_residual
    return rep


# ---------------------------------------------------------------------------------------------
# cameras (U:sensor/models.h, COLMAP 3.9.1 ids and parameter orders)
#   0 SIMPLE_PINHOLE f cx cy | 1 PINHOLE fx fy cx cy | 2 SIMPLE_RADIAL f cx cy k | 3 RADIAL f cx cy k1 k2
#   4 OPENCV fx fy cx cy k1 k2 p1 p2 | 5 OPENCV_FISHEYE fx fy cx cy k1 k2 k3 k4
#   6 FULL_OPENCV fx fy cx cy k1 k2 p1 p2 k3 k4 k5 k6 | 7 FOV fx fy cx cy omega | 8 SIMPLE_RADIAL_FISHEYE f cx cy k
#   9 RADIAL_FISHEYE f cx cy k1 k2 | 10 THIN_PRISM_FISHEYE fx fy cx cy k1 k2 p1 p2 k3 k4 sx1 sy1
# ---------------------------------------------------------------------------------------------
CAMERA_NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}
_SINGLE_FOCAL = (0, 2, 3, 8, 9)


def _intrinsics(cam):
    p = [float(x) for x in cam["params"]]
    model = cam.get("model", 0)
    if model not in CAMERA_NUM_PARAMS or len(p) != CAMERA_NUM_PARAMS[model]:
        raise ValueError(f"camera model {model} with {len(p)} parameters")
    if model in _SINGLE_FOCAL:
        return model, np.array([p[0], p[0]]), np.array([p[1], p[2]]), p[3:]
    return model, np.array([p[0], p[1]]), np.array([p[2], p[3]]), p[4:]


def camera_distortion(model, k, uv):
    """d(u, v) of the model for an [n x 2] array of normalised points."""
    u, v = uv[:, 0], uv[:, 1]
    r2 = u * u + v * v
    if model in (0, 1):
        return np.zeros_like(uv)
    if model in (2, 3):
        radial = k[0] * r2 + (k[1] * r2 * r2 if model == 3 else 0.0)
        return np.stack([u * radial, v * radial], 1)
    if model in (4, 6):
        if model == 4:
            radial = k[0] * r2 + k[1] * r2 * r2
        else:
            radial = (1 + k[0] * r2 + k[1] * r2 ** 2 + k[4] * r2 ** 3) / (1 + k[5] * r2 + k[6] * r2 ** 2 + k[7] * r2 ** 3) - 1
        du = u * radial + 2 * k[2] * u * v + k[3] * (r2 + 2 * u * u)
        dv = v * radial + 2 * k[3] * u * v + k[2] * (r2 + 2 * v * v)
        return np.stack([du, dv], 1)
    if model == 10:   # thin prism, on equidistant fisheye coordinates
        r4 = r2 * r2
        radial = k[0] * r2 + k[1] * r4 + k[4] * r4 * r2 + k[5] * r4 * r4
        du = u * radial + 2 * k[2] * u * v + k[3] * (r2 + 2 * u * u) + k[6] * r2
        dv = v * radial + 2 * k[3] * u * v + k[2] * (r2 + 2 * v * v) + k[7] * r2
        return np.stack([du, dv], 1)
    # fisheye family: theta_d = theta * (1 + k1 theta^2 + ...), d = uv * theta_d / r - uv
    kk = (list(k) + [0.0, 0.0, 0.0])[:4]      # models 8 / 9 carry one / two coefficients
    r = np.sqrt(r2)
    safe = np.where(r > np.finfo(float).eps, r, 1.0)
    th = np.arctan(r)
    thd = th * (1 + kk[0] * th ** 2 + kk[1] * th ** 4 + kk[2] * th ** 6 + kk[3] * th ** 8)
    scale = np.where(r > np.finfo(float).eps, thd / safe - 1.0, 0.0)
    return np.stack([u * scale, v * scale], 1)


def img_from_cam(cam, uv):
    """Camera::ImgFromCam on normalised coordinates [n x 2]."""
    model, f, c, k = _intrinsics(cam)
    uv = np.asarray(uv, np.float64).reshape(-1, 2)
    if model == 7:    # FOV: r_d = atan(2 r tan(w / 2)) / w
        r = np.linalg.norm(uv, axis=1)
        w = k[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.where(r > 1e-9, np.arctan(2 * r * np.tan(w / 2)) / (r * w), 2 * np.tan(w / 2) / w)
        return uv * factor[:, None] * f + c
    if model == 10:   # equidistant fisheye first, then the thin-prism distortion
        r = np.linalg.norm(uv, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            scale = np.where(r > np.finfo(float).eps, np.arctan(r) / r, 1.0)
        uv = uv * scale[:, None]
    return (uv + camera_distortion(model, k, uv)) * f + c


def cam_from_img(cam, pts):
    """Camera::CamFromImg: closed form for the pinhole models, otherwise upstream's IterativeUndistortion
    (Newton on uv + d(uv) - uv0 with a central-difference Jacobian, <= 100 iterations, |step|^2 < 1e-10)."""
    model, f, c, k = _intrinsics(cam)
    uv0 = (np.asarray(pts, np.float64) - c) / f
    if model in (0, 1):
        return uv0
    if model == 7:    # closed-form inverse: r = tan(r_d w) / (2 tan(w / 2))
        rd = np.linalg.norm(uv0, axis=1)
        w = k[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            factor = np.where(rd > 1e-9, np.tan(rd * w) / (2 * rd * np.tan(w / 2)), w / (2 * np.tan(w / 2)))
        return uv0 * factor[:, None]
    x = uv0.copy()
    active = np.ones(len(x), bool)
    eps = np.finfo(float).eps
    for _ in range(100):
        if not active.any():
            break
        xa = x[active]
        h = np.maximum(eps, np.abs(1e-6 * xa))
        d = camera_distortion(model, k, xa)
        hx = np.stack([h[:, 0], np.zeros(len(xa))], 1)
        hy = np.stack([np.zeros(len(xa)), h[:, 1]], 1)
        ddx = (camera_distortion(model, k, xa + hx) - camera_distortion(model, k, xa - hx)) / (2 * h[:, :1])
        ddy = (camera_distortion(model, k, xa + hy) - camera_distortion(model, k, xa - hy)) / (2 * h[:, 1:])
        J = np.zeros((len(xa), 2, 2))
        J[:, :, 0] = ddx
        J[:, :, 1] = ddy
        J[:, 0, 0] += 1.0
        J[:, 1, 1] += 1.0
        step = np.linalg.solve(J, (xa + d - uv0[active])[:, :, None])[:, :, 0]
        x[active] = xa - step
        idx = np.flatnonzero(active)
        active[idx[(step ** 2).sum(1) < 1e-10]] = False
    if model == 10:   # fisheye coordinates -> normalised plane: scale tan(theta) / theta
        th = np.linalg.norm(x, axis=1)
        tc = th * np.cos(th)
        with np.errstate(divide="ignore", invalid="ignore"):
            x = x * np.where(tc > np.finfo(float).eps, np.sin(th) / tc, 1.0)[:, None]
    return x


def mean_focal_length(cam):
    _, f, _, _ = _intrinsics(cam)
    return 0.5 * (f[0] + f[1])


def cam_from_img_threshold(cam, thr):
    return thr / mean_focal_length(cam)


# ---------------------------------------------------------------------------------------------
# two-view geometry
# ---------------------------------------------------------------------------------------------
class TwoViewGeometryOptions:
    def __init__(self, **kw):
        self.min_num_inliers = 15
        self.min_E_F_inlier_ratio = 0.95
        self.max_H_inlier_ratio = 0.8
        self.watermark_min_inlier_ratio = 0.7
        self.watermark_border_size = 0.1
        self.detect_watermark = True
        self.multiple_ignore_watermark = True
        self.force_H_use = False
        self.compute_relative_pose = False
        self.multiple_models = False
        self.ransac = RansacOptions()
        for k, v in kw.items():
            assert hasattr(self, k), k
            setattr(self, k, v)


class TwoViewGeometry:
    def __init__(self):
        self.config = UNDEFINED
        self.E, self.F, self.H = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        self.inlier_matches = np.zeros((0, 2), np.uint32)
        self.nE = self.nF = self.nH = 0
        self.reports = {}
        self.qvec, self.tvec, self.tri_angle, self.pose_valid = np.array([1.0, 0, 0, 0]), np.zeros(3), 0.0, False


def detect_watermark(cam1, pts1, cam2, pts2, num_inliers, mask, opt, rng):
    """DetectWatermark (U:estimators/two_view_geometry.cc)."""
    if num_inliers == 0:
        return False
    d1 = opt.watermark_border_size * math.hypot(cam1["width"], cam1["height"])
    d2 = opt.watermark_border_size * math.hypot(cam2["width"], cam2["height"])
    a, b = pts1[mask], pts2[mask]

    def border(p, cam, d):
        return (p[:, 0] < d) | (p[:, 0] > cam["width"] - d) | (p[:, 1] < d) | (p[:, 1] > cam["height"] - d)
    sel = border(a, cam1, d1) & border(b, cam2, d2)
    if sel.sum() / float(num_inliers) < opt.watermark_min_inlier_ratio:
        return False
    # the translation model is fitted to ALL inlier points (upstream builds inlier_points1/2 from the whole
    # mask; the border test above only gates the attempt)
    ro = opt.ransac.copy(min_inlier_ratio=opt.watermark_min_inlier_ratio)
    rep = loransac(Translation2D, Translation2D, a, b, ro, rng)
    inl_ratio = rep.num_inliers / float(num_inliers)
    return rep.success and inl_ratio >= opt.watermark_min_inlier_ratio


def _decide(g, opt, repE, repF, repH, calibrated):
    """Decision step of EstimateCalibratedTwoViewGeometry / EstimateUncalibratedTwoViewGeometry
    (U:estimators/two_view_geometry.cc).  Returns (inlier mask, number of inliers) or (None, 0)."""
    nE = repE.num_inliers if calibrated else 0
    nF, nH = repF.num_inliers, repH.num_inliers
    mn = opt.min_num_inliers
    if not calibrated:
        # EstimateUncalibratedTwoViewGeometry: the configuration depends on nH / nF, but the inlier matches and
        # the watermark test ALWAYS come from F's mask (also when the pair is PLANAR_OR_PANORAMIC)
        if (not repF.success and not repH.success) or (nF < mn and nH < mn):
            g.config = DEGENERATE
            return None, 0
        with np.errstate(divide="ignore", invalid="ignore"):
            H_F = np.float64(nH) / np.float64(nF)
        g.config = PLANAR_OR_PANORAMIC if H_F > opt.max_H_inlier_ratio else UNCALIBRATED
        return repF.inlier_mask, nF
    okE = calibrated and repE.success
    if (not okE and not repF.success and not repH.success) or (nE < mn and nF < mn and nH < mn):
        g.config = DEGENERATE
        return None, 0
    with np.errstate(divide="ignore", invalid="ignore"):
        E_F = np.float64(nE) / np.float64(nF)
        H_F = np.float64(nH) / np.float64(nF)
        H_E = np.float64(nH) / np.float64(nE)
    if okE and E_F > opt.min_E_F_inlier_ratio and nE >= mn:
        if nE >= nF:
            num, mask = nE, repE.inlier_mask
        else:
            num, mask = nF, repF.inlier_mask
        if H_E > opt.max_H_inlier_ratio:
            g.config = PLANAR_OR_PANORAMIC
            if nH > num:
                num, mask = nH, repH.inlier_mask
        else:
            g.config = CALIBRATED
    elif repF.success and nF >= mn:
        num, mask = nF, repF.inlier_mask
        if H_F > opt.max_H_inlier_ratio:
            g.config = PLANAR_OR_PANORAMIC
            if nH > num:
                num, mask = nH, repH.inlier_mask
        else:
            g.config = UNCALIBRATED
    elif repH.success and nH >= mn:
        num, mask = nH, repH.inlier_mask
        g.config = PLANAR_OR_PANORAMIC
    else:
        g.config = DEGENERATE
        return None, 0
    return mask, num


def estimate_two_view_geometry(cam1, points1, cam2, points2, matches=None, options=None, seed=0):
    """EstimateTwoViewGeometry.  points: [n, 2] float64; matches: [m, 2] or None (identity)."""
    opt = options or TwoViewGeometryOptions()
    rng = np.random.default_rng(seed)
    points1, points2 = np.asarray(points1, np.float64), np.asarray(points2, np.float64)
    if matches is None:
        assert len(points1) == len(points2)
        matches = np.stack([np.arange(len(points1))] * 2, 1)
    matches = np.asarray(matches, np.int64).reshape(-1, 2)
    g = TwoViewGeometry()
    if opt.multiple_models:
        # EstimateMultipleTwoViewGeometries (U:estimators/two_view_geometry.cc): estimate on the remaining
        # matches, keep the geometry (WATERMARK only if not multiple_ignore_watermark), drop its inliers, repeat
        # until DEGENERATE; one geometry -> itself, several -> MULTIPLE with concatenated inlier matches.
        import copy
        single = copy.copy(opt)
        single.multiple_models = False
        remaining, found = matches, []
        while True:
            gi = estimate_two_view_geometry(cam1, points1, cam2, points2, remaining, single,
                                            seed=int(rng.integers(0, 2 ** 31)))
            if gi.config == DEGENERATE or len(gi.inlier_matches) == 0:
                break
            if not (opt.multiple_ignore_watermark and gi.config == WATERMARK):
                found.append(gi)
            inl = {tuple(x) for x in np.asarray(gi.inlier_matches, np.int64).tolist()}
            remaining = np.array([x for x in remaining.tolist() if tuple(x) not in inl], np.int64).reshape(-1, 2)
        if not found:
            g.config = DEGENERATE
        elif len(found) == 1:
            g = found[0]
        else:
            g.config = MULTIPLE
            g.inlier_matches = np.concatenate([np.asarray(x.inlier_matches).reshape(-1, 2) for x in found])
        return g
    if len(matches) < opt.min_num_inliers:
        g.config = DEGENERATE
        return g
    p1, p2 = points1[matches[:, 0]], points2[matches[:, 1]]
    calibrated = bool(cam1.get("has_prior_focal_length")) and bool(cam2.get("has_prior_focal_length"))
    if opt.force_H_use:
        calibrated = False
    repE = Report()
    if calibrated:
        n1, n2 = cam_from_img(cam1, p1), cam_from_img(cam2, p2)
        e_opt = opt.ransac.copy(max_error=0.5 * (cam_from_img_threshold(cam1, opt.ransac.max_error)
                                                 + cam_from_img_threshold(cam2, opt.ransac.max_error)))
        repE = loransac(EssentialFivePoint, EssentialFivePoint, n1, n2, e_opt, rng)
        if repE.model is not None:
            g.E = repE.model
    repF = loransac(FundamentalSevenPoint, FundamentalEightPoint, p1, p2, opt.ransac, rng)
    if repF.model is not None:
        g.F = repF.model
    repH = loransac(Homography, Homography, p1, p2, opt.ransac, rng)
    if repH.model is not None:
        g.H = repH.model
    g.nE, g.nF, g.nH = repE.num_inliers, repF.num_inliers, repH.num_inliers
    g.reports = {"E": repE, "F": repF, "H": repH}
    if opt.force_H_use:
        # EstimateCalibratedHomography: H only
        if not repH.success or repH.num_inliers < opt.min_num_inliers:
            g.config = DEGENERATE
            return g
        g.config = PLANAR_OR_PANORAMIC
        mask, num = repH.inlier_mask, repH.num_inliers
    else:
        mask, num = _decide(g, opt, repE, repF, repH, calibrated)
    if mask is None:
        return g
    g.inlier_matches = matches[mask].astype(np.uint32)
    if opt.detect_watermark and detect_watermark(cam1, p1, cam2, p2, num, mask, opt, rng):
        g.config = WATERMARK
    if opt.compute_relative_pose:
        estimate_two_view_geometry_pose(cam1, points1, cam2, points2, g)
    return g


# ---------------------------------------------------------------------------------------------
# relative pose (U:geometry/essential_matrix.cc, homography_matrix.cc, pose.cc, triangulation.cc;
# R:estimators/two_view_geometry.h:153-158).  numpy SVDs on purpose: the CUDA path uses Jacobi eigen-solvers.
# ---------------------------------------------------------------------------------------------
def decompose_essential_matrix(E):
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    t = U[:, 2] / np.linalg.norm(U[:, 2])
    return U @ W @ Vt, U @ W.T @ Vt, t


def triangulate_point(P1, P2, x1, x2):
    A = np.stack([x1[0] * P1[2] - P1[0], x1[1] * P1[2] - P1[1], x2[0] * P2[2] - P2[0], x2[1] * P2[2] - P2[1]])
    X = np.linalg.svd(A)[2][3]
    return X[:3] / X[3]


def check_cheirality(Rm, t, pts1, pts2):
    """Points triangulated in front of both cameras and closer than 1000 baselines."""
    P1 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P2 = np.hstack([Rm, np.asarray(t, np.float64).reshape(3, 1)])
    max_depth = 1000.0 * np.linalg.norm(Rm.T @ t)
    eps = np.finfo(float).eps
    out = []
    for a, b in zip(pts1, pts2):
        X = triangulate_point(P1, P2, a, b)
        d1 = P1[2] @ np.append(X, 1.0) * np.linalg.norm(P1[:, 2])
        if eps < d1 < max_depth:
            d2 = P2[2] @ np.append(X, 1.0) * np.linalg.norm(P2[:, 2])
            if eps < d2 < max_depth:
                out.append(X)
    return out


def pose_from_essential_matrix(E, pts1, pts2):
    R1, R2, t = decompose_essential_matrix(E)
    best = (None, None, [])
    first = True
    for Rm, tt in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        X = check_cheirality(Rm, tt, pts1, pts2)
        if first or len(X) >= len(best[2]):
            best = (Rm, tt, X)
        first = False
    return best


def decompose_homography_matrix(H, K1, K2):
    """Malis & Vargas analytical decomposition: [(R, t, n)], one entry for a pure rotation, else four."""
    Hn = np.linalg.inv(K2) @ H @ K1
    Hn = Hn / np.linalg.svd(Hn)[1][1]
    if np.linalg.det(Hn) < 0:
        Hn = -Hn
    S = Hn.T @ Hn - np.eye(3)
    if np.abs(S).max() < 1e-3:                      # lpNorm<Infinity> of the matrix: largest |coefficient|
        return [(Hn, np.zeros(3), np.zeros(3))]

    def minor(r, c):
        c0, c1 = (1 if c == 0 else 0), (1 if c == 2 else 2)
        r0, r1 = (1 if r == 0 else 0), (1 if r == 2 else 2)
        return S[r0, c1] * S[r1, c0] - S[r0, c0] * S[r1, c1]

    def sgn(v):
        return -1.0 if v < 0 else 1.0
    M00, M11, M22 = minor(0, 0), minor(1, 1), minor(2, 2)
    r00, r11, r22 = np.sqrt(max(M00, 0)), np.sqrt(max(M11, 0)), np.sqrt(max(M22, 0))
    e12, e02, e01 = sgn(minor(1, 2)), sgn(minor(0, 2)), sgn(minor(0, 1))
    nS = [abs(S[0, 0]), abs(S[1, 1]), abs(S[2, 2])]
    if nS[0] < nS[1]:
        idx = 2 if nS[1] < nS[2] else 1
    else:
        idx = 2 if nS[0] < nS[2] else 0
    if idx == 0:
        np1 = np.array([S[0, 0], S[0, 1] + r22, S[0, 2] + e12 * r11])
        np2 = np.array([S[0, 0], S[0, 1] - r22, S[0, 2] - e12 * r11])
    elif idx == 1:
        np1 = np.array([S[0, 1] + r22, S[1, 1], S[1, 2] - e02 * r00])
        np2 = np.array([S[0, 1] - r22, S[1, 1], S[1, 2] + e02 * r00])
    else:
        np1 = np.array([S[0, 2] + e01 * r11, S[1, 2] + r00, S[2, 2]])
        np2 = np.array([S[0, 2] - e01 * r11, S[1, 2] - r00, S[2, 2]])
    tr = np.trace(S)
    v = 2.0 * np.sqrt(max(1.0 + tr - M00 - M11 - M22, 0))
    es = sgn(S[idx, idx])
    r, n_t = np.sqrt(max(2.0 + tr + v, 0)), np.sqrt(max(2.0 + tr - v, 0))
    n1, n2 = np1 / np.linalg.norm(np1), np2 / np.linalg.norm(np2)
    t1s = 0.5 * n_t * (es * r * n2 - n_t * n1)
    t2s = 0.5 * n_t * (es * r * n1 - n_t * n2)
    Ra = Hn @ (np.eye(3) - (2.0 / v) * np.outer(t1s, n1))
    Rb = Hn @ (np.eye(3) - (2.0 / v) * np.outer(t2s, n2))
    ta, tb = Ra @ t1s, Rb @ t2s
    return [(Ra, ta, -n1), (Ra, -ta, n1), (Rb, tb, -n2), (Rb, -tb, n2)]


def calibration_matrix(cam):
    _, f, c, _ = _intrinsics(cam)
    return np.array([[f[0], 0, c[0]], [0, f[1], c[1]], [0, 0, 1.0]])


def pose_from_homography_matrix(H, K1, K2, pts1, pts2):
    best = (None, None, None, [])
    first = True
    for Rm, t, n in decompose_homography_matrix(H, K1, K2):
        X = check_cheirality(Rm, t, pts1, pts2)
        if first or len(X) >= len(best[3]):
            best = (Rm, t, n, X)
        first = False
    return best


def triangulation_angles(c1, c2, X):
    X = np.asarray(X, np.float64).reshape(-1, 3)
    b2 = ((c1 - c2) ** 2).sum()
    r1, r2 = ((X - c1) ** 2).sum(1), ((X - c2) ** 2).sum(1)
    den = 2.0 * np.sqrt(r1 * r2)
    ang = np.zeros(len(X))
    ok = den > 0
    ang[ok] = np.abs(np.arccos(np.clip((r1 + r2 - b2)[ok] / den[ok], -1.0, 1.0)))
    return np.minimum(ang, np.pi - ang)


def median(v):
    v = np.sort(np.asarray(v, np.float64))
    m = len(v) // 2
    return 0.5 * (v[m - 1] + v[m]) if len(v) % 2 == 0 else v[m]


def rotation_to_quat(Rm):
    """(w, x, y, z), the branch selection of Eigen's Quaterniond(Matrix3d)."""
    tr = np.trace(Rm)
    q = np.zeros(4)
    if tr > 0:
        s = np.sqrt(tr + 1.0)
        q[0] = 0.5 * s
        s = 0.5 / s
        q[1:] = [(Rm[2, 1] - Rm[1, 2]) * s, (Rm[0, 2] - Rm[2, 0]) * s, (Rm[1, 0] - Rm[0, 1]) * s]
    else:
        i = int(np.argmax(np.diag(Rm)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(Rm[i, i] - Rm[j, j] - Rm[k, k] + 1.0)
        q[1 + i] = 0.5 * s
        s = 0.5 / s
        q[0] = (Rm[k, j] - Rm[j, k]) * s
        q[1 + j] = (Rm[j, i] + Rm[i, j]) * s
        q[1 + k] = (Rm[k, i] + Rm[i, k]) * s
    return q


def estimate_two_view_geometry_pose(cam1, points1, cam2, points2, g):
    """EstimateTwoViewGeometryPose: fills g.qvec (w, x, y, z), g.tvec, g.tri_angle; may turn
    PLANAR_OR_PANORAMIC into PLANAR / PANORAMIC.  Returns False when no pose could be recovered."""
    g.qvec, g.tvec, g.tri_angle, g.pose_valid = np.array([1.0, 0, 0, 0]), np.zeros(3), 0.0, False
    if g.config not in (CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC):
        return False
    im = np.asarray(g.inlier_matches, np.int64).reshape(-1, 2)
    n1 = cam_from_img(cam1, np.asarray(points1, np.float64)[im[:, 0]])
    n2 = cam_from_img(cam2, np.asarray(points2, np.float64)[im[:, 1]])
    if g.config in (CALIBRATED, UNCALIBRATED):
        Rm, t, X = pose_from_essential_matrix(g.E, n1, n2)
        if len(X) == 0:
            return False
    else:
        Rm, t, _, X = pose_from_homography_matrix(g.H, calibration_matrix(cam1), calibration_matrix(cam2), n1, n2)
    g.qvec, g.tvec, g.pose_valid = rotation_to_quat(Rm), np.asarray(t, np.float64), True
    g.tri_angle = 0.0 if len(X) == 0 else float(median(triangulation_angles(np.zeros(3), -Rm.T @ t, X)))
    if g.config == PLANAR_OR_PANORAMIC:
        if np.linalg.norm(t) == 0:
            g.config, g.tri_angle = PANORAMIC, 0.0
        else:
            g.config = PLANAR
    return True


# ---------------------------------------------------------------------------------------------
# pair generators (U:controllers/feature_matching.cc)
# ---------------------------------------------------------------------------------------------
def exhaustive_pairs(image_ids, block_size=50):
    """ExhaustiveFeatureMatcher::Run visiting order (row P1)."""
    ids = list(image_ids)
    n = len(ids)
    out = []
    nb = int(math.ceil(n / block_size))
    for b1 in range(nb):
        s1, e1 = b1 * block_size, min(n, (b1 + 1) * block_size) - 1
        for b2 in range(nb):
            s2, e2 = b2 * block_size, min(n, (b2 + 1) * block_size) - 1
            for i1 in range(s1, e1 + 1):
                for i2 in range(s2, e2 + 1):
                    b_i1, b_i2 = i1 % block_size, i2 % block_size
                    if (i1 > i2 and b_i1 <= b_i2) or (i1 < i2 and b_i1 < b_i2):
                        out.append((ids[i1], ids[i2]))
    return out


def gps_to_ecef(lat_deg, lon_deg, alt):
    """GPSTransform::EllToXYZ on the WGS84 ellipsoid (U:geometry/gps.cc)."""
    a, f = 6378137.0, 1.0 / 298.257223563
    b = a * (1.0 - f)
    e2 = (a * a - b * b) / (a * a)
    lat, lon = np.radians(lat_deg), np.radians(lon_deg)
    N = a / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    return np.stack([(N + alt) * np.cos(lat) * np.cos(lon), (N + alt) * np.cos(lat) * np.sin(lon),
                     ((b * b) / (a * a) * N + alt) * np.sin(lat)], -1)


def spatial_pairs(prior_t, has_prior, is_gps=True, ignore_z=True, max_num_neighbors=50, max_distance=100.0):
    """SpatialFeatureMatcher::Run pair generation (U:controllers/feature_matching.cc): for every image with a location
    prior, the k = min(max_num_neighbors, #locations) nearest located images including itself, minus itself, up to
    max_distance; indices are positions in the image list."""
    prior_t = np.asarray(prior_t, np.float64).reshape(-1, 3)
    idx = [i for i, h in enumerate(has_prior) if h]
    if not idx:
        return []
    loc = prior_t[idx].copy()
    if is_gps:
        loc = gps_to_ecef(loc[:, 0], loc[:, 1], np.zeros(len(loc)) if ignore_z else loc[:, 2])
    elif ignore_z:
        loc[:, 2] = 0.0
    knn = min(max_num_neighbors, len(loc))
    out = []
    for i in range(len(loc)):
        d2 = ((loc - loc[i]) ** 2).sum(1)
        order = sorted(range(len(loc)), key=lambda j: (d2[j], j))[:knn]
        for j in order:
            if j == i:
                continue
            if d2[j] > max_distance ** 2:
                break
            out.append((idx[i], idx[j]))
    return out


def sequential_pairs(image_ids, overlap=10, quadratic_overlap=True):
    """SequentialFeatureMatcher::RunSequentialMatching (row P2); images ordered by name upstream."""
    ids = list(image_ids)
    n = len(ids)
    out = []
    # COLMAP 3.9.1: image_idx2 = image_idx1 + i, i in [0, overlap): i = 0 is the self pair (dropped by the
    # controller), so overlap - 1 linear neighbours; the quadratic partner sits inside the same in-range test
    for i1 in range(n):
        for k in range(overlap):
            i2 = i1 + k
            if i2 >= n:
                break
            if i2 != i1:
                out.append((ids[i1], ids[i2]))
            if quadratic_overlap:
                i2q = i1 + (1 << k)
                if i2q < n:
                    out.append((ids[i1], ids[i2q]))
    seen, uniq = set(), []
    for a, b in out:            # the controller drops duplicates
        key = (min(a, b), max(a, b))
        if key not in seen:
            seen.add(key)
            uniq.append((a, b))
    return uniq
