"""Host-side mirror of the reference's matching / verification entry points.

Same names, keyword arguments, defaults, dict-merge behaviour and error types as
  match_exhaustive / match_sequential   R:pipeline/match_features.h:22-49, 219-235
  verify_matches                        R:pipeline/match_features.h:51-68, 255-260
  estimate_two_view_geometry, estimate_calibrated_two_view_geometry, squared_sampson_error
                                        R:estimators/two_view_geometry.h:95-175
  essential / fundamental / homography_matrix_estimation
                                        R:estimators/essential_matrix.h:19-103, fundamental_matrix.h:17-50,
                                        homography_matrix.h:17-48
The compute path is libb200match.so (C ABI, include/b200match.h); this module only does what the
reference's controllers do on the host: database I/O, pair generation, the write rules.
"""
import os
import threading
import time

import numpy as np

from . import _lib
from .database import Database
from .options import (Device, ExhaustiveMatchingOptions, RANSACOptions, SequentialMatchingOptions,
                      SiftMatchingOptions, TwoViewGeometryConfiguration, TwoViewGeometryOptions, _enum_from)

_contexts = {}
_ctx_lock = threading.Lock()


def get_context(device_index=0):
    """One b2m context per process and GPU (created lazily)."""
    with _ctx_lock:
        c = _contexts.get(device_index)
        if c is None or c.h is None:
            c = _lib.Context(device=device_index, seed=0)
            _contexts[device_index] = c
        return c


def _check_file_exists(path, where):
    # THROW_CHECK_FILE_EXISTS (R:log_exceptions.h:137-141) -> ValueError with the [file:line] prefix
    if not os.path.isfile(os.fspath(path)):
        raise ValueError(f"[{where}] Check Failed: File {os.fspath(path)} does not exist.")


def _resolve_device(device, sift_options):
    """IsGPU / VerifyGPUParams (R:utils.h:11-31).  `auto` means the GPU; there is no CPU path."""
    device = _enum_from(Device, device)
    if device == Device.cpu:
        raise ValueError("[pipeline.py] pycolmap_b200 has no CPU path: use Device.auto or Device.cuda")
    gi = str(sift_options.gpu_index).split(",")[0].strip()
    return max(0, int(gi)) if gi not in ("", "-1") else 0


def _sift_struct(ctx, o):
    return ctx.sift_opts(max_ratio=float(o.max_ratio), max_distance=float(o.max_distance),
                         cross_check=int(bool(o.cross_check)), max_num_matches=int(o.max_num_matches),
                         guided_matching=int(bool(o.guided_matching)))


def _ransac_kwargs(r):
    return dict(max_error=float(r.max_error), min_inlier_ratio=float(r.min_inlier_ratio),
                confidence=float(r.confidence), dyn_num_trials_multiplier=float(r.dyn_num_trials_multiplier),
                min_num_trials=int(r.min_num_trials), max_num_trials=int(r.max_num_trials))


def _tvg_struct(ctx, o):
    return ctx.tvg_opts(ransac=_ransac_kwargs(o.ransac), min_num_inliers=int(o.min_num_inliers),
                        min_E_F_inlier_ratio=float(o.min_E_F_inlier_ratio),
                        max_H_inlier_ratio=float(o.max_H_inlier_ratio),
                        watermark_min_inlier_ratio=float(o.watermark_min_inlier_ratio),
                        watermark_border_size=float(o.watermark_border_size),
                        detect_watermark=int(bool(o.detect_watermark)),
                        multiple_ignore_watermark=int(bool(o.multiple_ignore_watermark)),
                        force_H_use=int(bool(o.force_H_use)), compute_relative_pose=int(bool(o.compute_relative_pose)),
                        multiple_models=int(bool(o.multiple_models)))


def _run_interruptible(ctx, fn):
    """PyWait (R:helpers.h:335-347): the C call runs in a worker thread, the caller polls so that
    Ctrl-C is honoured: KeyboardInterrupt -> b2m_request_stop -> join -> re-raise."""
    box = {}

    def work():
        try:
            box["r"] = fn()
        except BaseException as e:  # noqa: BLE001
            box["e"] = e
    t = threading.Thread(target=work, daemon=True)
    t.start()
    try:
        while t.is_alive():
            t.join(0.2)
    except KeyboardInterrupt:
        ctx.lib.b2m_request_stop(ctx.h)
        t.join()
        raise
    if "e" in box:
        raise box["e"]
    return box["r"]


# ---------------------------------------------------------------------------------------------------
# pair generators (U:controllers/feature_matching.cc)
# ---------------------------------------------------------------------------------------------------
def exhaustive_pair_blocks(n, block_size):
    """Yields index-pair arrays block by block in ExhaustiveFeatureMatcher::Run order."""
    nb = (n + block_size - 1) // block_size
    for b1 in range(nb):
        i1 = np.arange(b1 * block_size, min(n, (b1 + 1) * block_size))
        for b2 in range(nb):
            i2 = np.arange(b2 * block_size, min(n, (b2 + 1) * block_size))
            a, b = np.meshgrid(i1, i2, indexing="ij")
            ma, mb = a % block_size, b % block_size
            keep = ((a > b) & (ma <= mb)) | ((a < b) & (ma < mb))
            if keep.any():
                yield np.stack([a[keep], b[keep]], 1).astype(np.int32)


def sequential_pairs(n, overlap, quadratic_overlap):
    out, seen = [], set()
    for i1 in range(n):
        for k in range(overlap):
            for i2 in ((i1 + k + 1,) + ((i1 + (1 << k),) if quadratic_overlap else ())):
                if i2 < n and (i1, i2) not in seen:
                    seen.add((i1, i2))
                    out.append((i1, i2))
    return np.array(out, np.int32).reshape(-1, 2)


# ---------------------------------------------------------------------------------------------------
# database-driven pipelines
# ---------------------------------------------------------------------------------------------------
class _Loaded:
    pass


def _load_image_set(db, ctx, need_geometry, order_by_name=False):
    images = db.read_all_images()
    if order_by_name:
        images = sorted(images, key=lambda r: r[1])
    L = _Loaded()
    L.ids = [r[0] for r in images]
    L.names = [r[1] for r in images]
    descs = [db.read_descriptors(i) for i in L.ids]
    kpts = cams = None
    if need_geometry:
        kpts = [np.ascontiguousarray(db.read_keypoints(i)[:, :2]) for i in L.ids]
        cam_cache = {}
        cams = []
        for (_, _, cid) in images:
            if cid not in cam_cache:
                cam_cache[cid] = db.read_camera(cid)
            cams.append(cam_cache[cid])
        for d, k in zip(descs, kpts):
            if len(d) != len(k):
                raise ValueError("[pipeline.py] Check Failed: keypoints.rows == descriptors.rows")
    ctx.set_images(descs, kpts, cams)
    return L


def _write_results(db, L, pairs, res, verified):
    """FeatureMatcherController::Match tail (row P3): one transaction per chunk."""
    with db.transaction():
        for k in range(len(pairs)):
            id1, id2 = L.ids[pairs[k, 0]], L.ids[pairs[k, 1]]
            v = res.view(k)
            db.write_matches(id1, id2, res.matches(k))
            if verified:
                E = np.array(v.E).reshape(3, 3)
                F = np.array(v.F).reshape(3, 3)
                H = np.array(v.H).reshape(3, 3)
                db.write_two_view_geometry(id1, id2, v.config, res.inlier_matches(k), F, E, H,
                                           qvec=list(v.qvec), tvec=list(v.tvec))


def _match_pairs_into_db(db, ctx, L, pair_chunks, sift, tvg, skip_existing=True):
    have_m = db.existing_pair_ids("matches") if skip_existing else set()
    have_g = db.existing_pair_ids("two_view_geometries") if skip_existing else set()
    from .database import image_pair_to_pair_id
    for pairs in pair_chunks:
        if len(pairs) == 0:
            continue
        keep = np.ones(len(pairs), bool)
        for k, (a, b) in enumerate(pairs):
            if a == b:
                keep[k] = False
                continue
            pid = image_pair_to_pair_id(L.ids[a], L.ids[b])
            if pid in have_m and pid in have_g:
                keep[k] = False  # both results stored: skip (resume semantics)
            have_m.add(pid)
            have_g.add(pid)
        pairs = np.ascontiguousarray(pairs[keep])
        if len(pairs) == 0:
            continue
        res = _run_interruptible(ctx, lambda p=pairs: ctx.match_pairs(p, sift, tvg))
        _write_results(db, L, pairs, res, tvg is not None)
        res.free()


def _chunked(gen, target=65536):
    buf, n = [], 0
    for p in gen:
        buf.append(p)
        n += len(p)
        if n >= target:
            yield np.concatenate(buf)
            buf, n = [], 0
    if buf:
        yield np.concatenate(buf)


def match_exhaustive(database_path, sift_options=None, matching_options=None, verification_options=None,
                     device=Device.auto):
    """Exhaustive feature matching + geometric verification of every image pair of the database."""
    _check_file_exists(database_path, "match_features.h:32")
    sift_options = SiftMatchingOptions.coerce(sift_options)
    matching_options = ExhaustiveMatchingOptions.coerce(matching_options)
    verification_options = TwoViewGeometryOptions.coerce(verification_options)
    if matching_options.block_size <= 1:
        raise ValueError("[pipeline.py] Check Failed: block_size > 1")
    ctx = get_context(_resolve_device(device, sift_options))
    with Database(database_path) as db:
        L = _load_image_set(db, ctx, need_geometry=True)
        _match_pairs_into_db(db, ctx, L, _chunked(exhaustive_pair_blocks(len(L.ids), matching_options.block_size)),
                             _sift_struct(ctx, sift_options), _tvg_struct(ctx, verification_options))


def match_sequential(database_path, sift_options=None, matching_options=None, verification_options=None,
                     device=Device.auto):
    """Sequential feature matching (images ordered by name; overlap / quadratic overlap)."""
    _check_file_exists(database_path, "match_features.h:32")
    sift_options = SiftMatchingOptions.coerce(sift_options)
    matching_options = SequentialMatchingOptions.coerce(matching_options)
    verification_options = TwoViewGeometryOptions.coerce(verification_options)
    if matching_options.loop_detection:  # guided matching IS supported (K1g)
        raise ValueError("[pipeline.py] loop_detection needs a vocabulary tree: out of scope (SURVEY.md row B6)")
    ctx = get_context(_resolve_device(device, sift_options))
    with Database(database_path) as db:
        L = _load_image_set(db, ctx, need_geometry=True, order_by_name=True)
        pairs = sequential_pairs(len(L.ids), matching_options.overlap, matching_options.quadratic_overlap)
        _match_pairs_into_db(db, ctx, L, [pairs], _sift_struct(ctx, sift_options),
                             _tvg_struct(ctx, verification_options))


def verify_matches(database_path, pairs_path, options=None):
    """Run geometric verification of the matches of the listed pairs (`name1 name2` per line).
    Pairs without stored raw matches are matched first, like the reference's ImagePairsFeatureMatcher."""
    _check_file_exists(database_path, "match_features.h:54")
    _check_file_exists(pairs_path, "match_features.h:55")
    options = TwoViewGeometryOptions.coerce(options)
    sift_options = SiftMatchingOptions()
    ctx = get_context(0)
    with Database(database_path) as db:
        L = _load_image_set(db, ctx, need_geometry=True)
        name_to_idx = {n: i for i, n in enumerate(L.names)}
        todo_verify, todo_match, seen = [], [], set()
        with open(pairs_path) as f:
            for line in f:
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                parts = line.split()
                if len(parts) < 2 or parts[0] not in name_to_idx or parts[1] not in name_to_idx:
                    continue  # upstream logs and skips unknown images
                a, b = name_to_idx[parts[0]], name_to_idx[parts[1]]
                if a == b or (min(a, b), max(a, b)) in seen:
                    continue
                seen.add((min(a, b), max(a, b)))
                has_m = db.exists_matches(L.ids[a], L.ids[b])
                has_g = db.exists_inlier_matches(L.ids[a], L.ids[b])
                if has_m and has_g:
                    continue
                (todo_verify if has_m else todo_match).append((a, b))
        tvg = _tvg_struct(ctx, options)
        if todo_match:
            _match_pairs_into_db(db, ctx, L, [np.array(todo_match, np.int32)], _sift_struct(ctx, sift_options), tvg,
                                 skip_existing=False)
        with db.transaction():
            for a, b in todo_verify:
                id1, id2 = L.ids[a], L.ids[b]
                m = db.read_matches(id1, id2)
                kp1 = db.read_keypoints(id1)[:, :2].astype(np.float64)
                kp2 = db.read_keypoints(id2)[:, :2].astype(np.float64)
                cam1 = db.read_camera(db.con.execute("SELECT camera_id FROM images WHERE image_id=?", (id1,)).fetchone()[0])
                cam2 = db.read_camera(db.con.execute("SELECT camera_id FROM images WHERE image_id=?", (id2,)).fetchone()[0])
                cfg, inl, E, F, H = TwoViewGeometryConfiguration.UNDEFINED, np.zeros((0, 2), np.uint32), None, None, None
                qvec = tvec = None
                if len(m) >= options.min_num_inliers:
                    r, inl = ctx.estimate_two_view_geometry(cam1, kp1, cam2, kp2, m, tvg)
                    cfg = r.config
                    E, F, H = (np.array(x).reshape(3, 3) for x in (r.E, r.F, r.H))
                    qvec, tvec = list(r.qvec), list(r.tvec)
                if len(inl) < options.min_num_inliers:  # controller write rule (row P3)
                    cfg, inl, E, F, H = TwoViewGeometryConfiguration.UNDEFINED, np.zeros((0, 2), np.uint32), None, None, None
                    qvec = tvec = None
                db.write_two_view_geometry(id1, id2, int(cfg), inl, F, E, H, qvec=qvec, tvec=tvec)


# ---------------------------------------------------------------------------------------------------
# estimators
# ---------------------------------------------------------------------------------------------------
class Rotation3d:
    """Minimal stand-in for pycolmap.Rotation3d: `quat` is (x, y, z, w) like the reference's Eigen coefficients."""

    def __init__(self, quat_xyzw=(0.0, 0.0, 0.0, 1.0)):
        self.quat = np.array(quat_xyzw, np.float64).reshape(4)

    def matrix(self):
        x, y, z, w = self.quat
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class Rigid3d:
    """Minimal stand-in for pycolmap.Rigid3d: x_cam2 = rotation * x_cam1 + translation."""

    def __init__(self, rotation=None, translation=(0.0, 0.0, 0.0)):
        self.rotation = rotation if rotation is not None else Rotation3d()
        self.translation = np.array(translation, np.float64).reshape(3)

    def matrix(self):
        return np.hstack([self.rotation.matrix(), self.translation.reshape(3, 1)])

    def inverse(self):
        x, y, z, w = self.rotation.quat
        inv = Rotation3d((-x, -y, -z, w))
        return Rigid3d(inv, -inv.matrix() @ self.translation)

    def __repr__(self):
        return f"Rigid3d(quat_xyzw={self.rotation.quat.tolist()}, t={self.translation.tolist()})"


class TwoViewGeometry:
    """R:estimators/two_view_geometry.h:82-93 (read-only members)."""

    def __init__(self, config=TwoViewGeometryConfiguration.UNDEFINED, E=None, F=None, H=None, inlier_matches=None,
                 tri_angle=0.0, qvec=None, tvec=None):
        self.config = TwoViewGeometryConfiguration(int(config))
        self.E = np.zeros((3, 3)) if E is None else np.array(E, np.float64).reshape(3, 3)
        self.F = np.zeros((3, 3)) if F is None else np.array(F, np.float64).reshape(3, 3)
        self.H = np.zeros((3, 3)) if H is None else np.array(H, np.float64).reshape(3, 3)
        # identity unless compute_relative_pose recovered a pose (qvec is (w, x, y, z) as in the database)
        q = (1.0, 0.0, 0.0, 0.0) if qvec is None else tuple(float(v) for v in qvec)
        self.cam2_from_cam1 = Rigid3d(Rotation3d((q[1], q[2], q[3], q[0])), (0.0, 0.0, 0.0) if tvec is None else tvec)
        self.inlier_matches = (np.zeros((0, 2), np.uint32) if inlier_matches is None
                               else np.array(inlier_matches, np.uint32).reshape(-1, 2))
        self.tri_angle = float(tri_angle)

    def invert(self):
        self.F = self.F.T.copy()
        self.E = self.E.T.copy()
        if np.abs(self.H).sum() > 0:
            self.H = np.linalg.inv(self.H)
        self.inlier_matches = self.inlier_matches[:, ::-1].copy()
        self.cam2_from_cam1 = self.cam2_from_cam1.inverse()

    def __repr__(self):
        return f"TwoViewGeometry(config={self.config.name}, num_inliers={len(self.inlier_matches)})"


def _camera_dict(camera):
    from .database import CAMERA_MODEL_IDS, CAMERA_MODEL_NUM_PARAMS
    if isinstance(camera, dict):
        model, get = camera.get("model", camera.get("model_id", 0)), camera.get
    else:  # duck-typed pycolmap.Camera: model (enum, name or id), width, height, params, has_prior_focal_length
        model = getattr(camera, "model", getattr(camera, "model_id", 0))

        def get(key, default=None):
            return getattr(camera, key, default)
    name = getattr(model, "name", model)
    model_id = CAMERA_MODEL_IDS.get(name) if isinstance(name, str) else int(name)
    if model_id not in CAMERA_MODEL_NUM_PARAMS:
        raise ValueError(f"[pipeline.py] camera model {name} is not supported (COLMAP 3.9.1 models only)")
    params = [float(x) for x in get("params", [])]
    if len(params) != CAMERA_MODEL_NUM_PARAMS[model_id]:
        raise ValueError(f"[pipeline.py] Check Failed: camera model {name} has "
                         f"{CAMERA_MODEL_NUM_PARAMS[model_id]} parameters")
    return dict(model=model_id, width=int(get("width", 0)), height=int(get("height", 0)), params=params,
                has_prior_focal_length=int(bool(get("has_prior_focal_length", False))))


def _points(p, name):
    p = np.asarray(p, np.float64)
    if p.ndim != 2 or p.shape[1] != 2:
        raise ValueError(f"[pipeline.py] Check Failed: {name} is an N x 2 array")
    return np.ascontiguousarray(p)


def estimate_two_view_geometry(camera1, points1, camera2, points2, matches=None, options=None):
    options = TwoViewGeometryOptions.coerce(options)
    p1, p2 = _points(points1, "points1"), _points(points2, "points2")
    if matches is None and len(p1) != len(p2):
        raise ValueError("[two_view_geometry.h:137] Check Failed: points1.size() == points2.size()")
    ctx = get_context(0)
    r, inl = ctx.estimate_two_view_geometry(_camera_dict(camera1), p1, _camera_dict(camera2), p2, matches,
                                            _tvg_struct(ctx, options))
    return TwoViewGeometry(r.config, r.E, r.F, r.H, inl, tri_angle=r.tri_angle, qvec=list(r.qvec), tvec=list(r.tvec))


def estimate_calibrated_two_view_geometry(camera1, points1, camera2, points2, matches=None, options=None):
    c1, c2 = dict(_camera_dict(camera1)), dict(_camera_dict(camera2))
    c1["has_prior_focal_length"] = c2["has_prior_focal_length"] = 1
    return estimate_two_view_geometry(c1, points1, c2, points2, matches, options)


def estimate_two_view_geometry_pose(camera1, points1, camera2, points2, geometry):
    """R:estimators/two_view_geometry.h:153-158: fills geometry.cam2_from_cam1 / tri_angle (and resolves
    PLANAR_OR_PANORAMIC) in place; returns False when no pose could be recovered."""
    ctx = get_context(0)
    r = ctx.estimate_two_view_geometry_pose(_camera_dict(camera1), _points(points1, "points1"), _camera_dict(camera2),
                                            _points(points2, "points2"), int(geometry.config), geometry.E, geometry.H,
                                            geometry.inlier_matches)
    if not r.pose_valid:
        return False
    geometry.config = TwoViewGeometryConfiguration(int(r.config))
    q = list(r.qvec)
    geometry.cam2_from_cam1 = Rigid3d(Rotation3d((q[1], q[2], q[3], q[0])), list(r.tvec))
    geometry.tri_angle = float(r.tri_angle)
    return True


def _ransac(kind, p1, p2, opts):
    ctx = get_context(0)
    opts = RANSACOptions.coerce(opts)
    return ctx.ransac_model(kind, p1, p2, ctx.ransac_opts(**_ransac_kwargs(opts)))


def fundamental_matrix_estimation(points2D1, points2D2, estimation_options=None):
    p1, p2 = _points(points2D1, "points2D1"), _points(points2D2, "points2D2")
    if len(p1) != len(p2):
        raise ValueError("[fundamental_matrix.h:22] Check Failed: points1.size() == points2.size()")
    r = _ransac(1, p1, p2, estimation_options)
    return None if r is None else {"F": r["model"], "num_inliers": r["num_inliers"], "inliers": r["inliers"]}


def homography_matrix_estimation(points2D1, points2D2, estimation_options=None):
    p1, p2 = _points(points2D1, "points2D1"), _points(points2D2, "points2D2")
    if len(p1) != len(p2):
        raise ValueError("[homography_matrix.h:21] Check Failed: points1.size() == points2.size()")
    r = _ransac(2, p1, p2, estimation_options)
    return None if r is None else {"H": r["model"], "num_inliers": r["num_inliers"], "inliers": r["inliers"]}


def essential_matrix_estimation(points2D1, points2D2, camera1, camera2, estimation_options=None):
    p1, p2 = _points(points2D1, "points2D1"), _points(points2D2, "points2D2")
    if len(p1) != len(p2):
        raise ValueError("[essential_matrix.h:26] Check Failed: points1.size() == points2.size()")
    c1, c2 = _camera_dict(camera1), _camera_dict(camera2)
    ctx = get_context(0)

    def norm(c, p):  # Camera::CamFromImg (R:estimators/essential_matrix.h:31-39)
        pr = c["params"]
        if c["model"] == 0:   # the two pinhole models: closed form, as validated on the GPU box in round 1
            return (p - [pr[1], pr[2]]) / pr[0]
        if c["model"] == 1:
            return (p - [pr[2], pr[3]]) / [pr[0], pr[1]]
        return ctx.cam_from_img(c, p)   # models with distortion: iterative undistortion on the GPU

    def mean_f(c):   # MeanFocalLength: models 0, 2, 3, 8, 9 have one focal length, the others two
        return c["params"][0] if c["model"] in (0, 2, 3, 8, 9) else 0.5 * (c["params"][0] + c["params"][1])
    o = RANSACOptions.coerce(estimation_options)
    o = RANSACOptions(o.todict())
    # R:estimators/essential_matrix.h:42-46: threshold averaged over both cameras
    o.max_error = 0.5 * (o.max_error / mean_f(c1) + o.max_error / mean_f(c2))
    r = _ransac(0, norm(c1, p1), norm(c2, p2), o)
    if r is None:
        return None
    def pose():
        # PoseFromEssentialMatrix on the inliers (R:estimators/essential_matrix.h:62-83), on the GPU
        idx = np.flatnonzero(r["inliers"]).astype(np.uint32)
        g = ctx.estimate_two_view_geometry_pose(c1, p1, c2, p2, int(TwoViewGeometryConfiguration.CALIBRATED), r["model"],
                                                np.zeros((3, 3)), np.stack([idx, idx], 1))
        q = list(g.qvec)
        return Rigid3d(Rotation3d((q[1], q[2], q[3], q[0])), list(g.tvec))
    return _EssentialResult(pose, {"E": r["model"], "num_inliers": r["num_inliers"], "inliers": r["inliers"]})


class _EssentialResult(dict):
    """The reference's result dict; `cam2_from_cam1` is decomposed on first access (this mirror layer keeps the
    round-1 validated call sequence for callers that only read E / inliers; the C++ host computes it eagerly)."""

    def __init__(self, pose_fn, items):
        super().__init__(items)
        self._pose_fn = pose_fn

    def __missing__(self, key):
        if key != "cam2_from_cam1":
            raise KeyError(key)
        self[key] = self._pose_fn()
        return self[key]


def squared_sampson_error(points2D1, points2D2, E):
    """Keyword names as bound by the reference (R:estimators/two_view_geometry.h:161-175)."""
    ctx = get_context(0)
    return ctx.squared_sampson_error(_points(points2D1, "points2D1"), _points(points2D2, "points2D2"), E)


def wait_idle():
    """Test helper: nothing is asynchronous at this level."""
    time.sleep(0)
