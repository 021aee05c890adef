"""CPU: the C++ host layer (pycolmap_b200 = pybind11 module _core) -- option classes with the
reference's dataclass behaviour (R:helpers.h:40-283), the C++ COLMAP database layer (interoperable with
a database written by plain sqlite3 in upstream's column layout), pair generators, argument checks and
error types.  No compute call is made: without a GPU the C ABI refuses to create a context."""
import copy
import pickle
import sqlite3

import numpy as np
import pytest

from helpers.native_import import load_native

nat = load_native()
from oracle import ransac as R


def test_option_defaults_match_reference():
    # SURVEY.md rows B4-B8 (upstream defaults of the fields the reference exposes)
    assert nat.SiftMatchingOptions().todict() == dict(num_threads=-1, gpu_index="-1", max_ratio=0.8, max_distance=0.7,
                                                      cross_check=True, max_num_matches=32768, guided_matching=False)
    assert nat.ExhaustiveMatchingOptions().todict() == dict(block_size=50)
    sq = nat.SequentialMatchingOptions().todict()
    assert (sq["overlap"], sq["quadratic_overlap"], sq["loop_detection"]) == (10, True, False)
    tv = nat.TwoViewGeometryOptions().todict()
    assert {k: v for k, v in tv.items() if k != "ransac"} == dict(
        min_num_inliers=15, min_E_F_inlier_ratio=0.95, max_H_inlier_ratio=0.8, watermark_min_inlier_ratio=0.7,
        watermark_border_size=0.1, detect_watermark=True, multiple_ignore_watermark=True, force_H_use=False,
        compute_relative_pose=False, multiple_models=False)
    t = nat.TwoViewGeometryOptions()   # .ransac keeps the C++ ctor defaults, RANSACOptions() the binding's
    assert (t.ransac.max_error, t.ransac.confidence, t.ransac.min_num_trials, t.ransac.max_num_trials,
            t.ransac.min_inlier_ratio) == (4.0, 0.999, 100, 10000, 0.25)
    r = nat.RANSACOptions()
    assert (r.max_error, r.min_inlier_ratio, r.confidence, r.min_num_trials, r.max_num_trials) == (
        4.0, 0.01, 0.9999, 1000, 100000)


def test_dataclass_behaviour():
    t = nat.TwoViewGeometryOptions({"min_num_inliers": 30, "ransac": {"max_error": 2.0}})
    assert t.min_num_inliers == 30 and t.ransac.max_error == 2.0 and t.ransac.confidence == 0.999
    t2 = nat.TwoViewGeometryOptions(min_num_inliers=7)
    t2.mergedict({"ransac": {"min_num_trials": 5}})
    assert t2.todict()["ransac"]["min_num_trials"] == 5 and t2.todict()["min_num_inliers"] == 7
    assert isinstance(t2.todict(recursive=False)["ransac"], nat.RANSACOptions)
    assert "min_num_inliers = 7" in t2.summary() and "ransac: RANSACOptions:" in t2.summary()
    assert "min_num_inliers: int = 7" in t2.summary(write_type=True)
    with pytest.raises(AttributeError):
        t.mergedict({"no_such_field": 1})
    with pytest.raises(AttributeError):
        t.no_such_field = 3
    with pytest.raises(TypeError):
        t.min_num_inliers = "many"
    with pytest.raises(TypeError, match="max_ratio"):
        nat.SiftMatchingOptions(max_ratio="high")
    with pytest.raises(TypeError):
        nat.SiftMatchingOptions({1: 2})
    s = nat.SiftMatchingOptions(max_ratio=1)          # int -> float like pybind11
    assert isinstance(s.max_ratio, float) and s.gpu_index == "-1"
    t.ransac.max_error = 3.0                          # nested attribute access writes through
    assert t.todict()["ransac"]["max_error"] == 3.0
    t.ransac = {"max_error": 5.0}                     # implicit dict -> RANSACOptions (fresh Python defaults)
    assert t.ransac.max_error == 5.0 and t.ransac.min_num_trials == 1000
    c = copy.deepcopy(t)
    c.ransac.max_error = 9.0
    assert t.ransac.max_error == 5.0 and copy.copy(t) == t and c != t
    assert pickle.loads(pickle.dumps(t)) == t
    assert nat.Device("cuda") == nat.Device.cuda and nat.Device.auto.value == -1
    with pytest.raises(IndexError, match="Invalid string value tpu"):   # std::out_of_range in the reference
        nat.Device("tpu")
    cfg = nat.TwoViewGeometryConfiguration
    assert [cfg(i).name for i in range(9)] == ["UNDEFINED", "DEGENERATE", "CALIBRATED", "UNCALIBRATED", "PLANAR",
                                                "PANORAMIC", "PLANAR_OR_PANORAMIC", "WATERMARK", "MULTIPLE"]
    assert cfg("WATERMARK").value == 7 and nat.has_cuda is True


def test_mergedict_coerces_through_base_classes():
    """R:helpers.h:70-84: a value pybind11 refuses (numpy scalars for bool / int fields) is retried through the bases of
    the attribute's class and through the class itself; what cannot be coerced raises the reference's TypeError."""
    o = nat.SiftMatchingOptions({"max_ratio": np.float32(0.75), "max_num_matches": np.int64(100), "cross_check": np.bool_(False)})
    assert (o.max_ratio, o.max_num_matches, o.cross_check) == (0.75, 100, False)
    t = nat.TwoViewGeometryOptions()
    t.mergedict({"min_num_inliers": np.int32(20), "ransac": {"max_error": np.float64(2.0), "max_num_trials": np.uint16(500)}})
    assert (t.min_num_inliers, t.ransac.max_error, t.ransac.max_num_trials) == (20, 2.0, 500)
    with pytest.raises(TypeError, match="Failed to merge dict into class: Could not assign max_ratio"):
        nat.SiftMatchingOptions({"max_ratio": "abc"})


def test_pair_generators_match_oracle():
    for n, bs in [(1, 50), (2, 2), (49, 50), (50, 50), (51, 50), (101, 50), (7, 2), (130, 64), (0, 5)]:
        blocks = nat.exhaustive_pair_blocks(n, bs)
        got = np.concatenate(blocks) if blocks else np.zeros((0, 2), np.int32)
        want = np.array(R.exhaustive_pairs(range(n), bs), np.int32).reshape(-1, 2)
        assert got.dtype == np.int32 and np.array_equal(got, want)     # same pairs, same visiting order
        assert len({(min(a, b), max(a, b)) for a, b in got}) == n * (n - 1) // 2
    for n, ov, q in [(100, 3, True), (50, 10, False), (1000, 20, True), (3, 10, True), (0, 2, True)]:
        assert np.array_equal(nat.sequential_pairs(n, ov, q),
                              np.array(R.sequential_pairs(range(n), ov, q), np.int32).reshape(-1, 2))
    with pytest.raises(ValueError):
        nat.exhaustive_pair_blocks(5, 0)


def _fill(db, rng):
    cid = db.add_camera(0, 1600, 1200, [1200.0, 800.0, 600.0], True)
    ids = [db.add_image(f"img{i:03d}.jpg", cid) for i in range(3)]
    kp = rng.uniform(0, 1000, (10, 6)).astype(np.float32)
    d = rng.integers(0, 255, (10, 128)).astype(np.uint8)
    db.write_keypoints(ids[0], kp)
    db.write_descriptors(ids[0], d)
    return cid, ids, kp, d


def test_database_roundtrip_cxx(tmp_path):
    rng = np.random.default_rng(0)
    with nat.Database(tmp_path / "db.db") as db:
        cid, ids, kp, d = _fill(db, rng)
        assert np.array_equal(db.read_keypoints(ids[0]), kp) and np.array_equal(db.read_descriptors(ids[0]), d)
        assert db.read_keypoints(ids[1]).shape == (0, 2) and db.read_descriptors(ids[1]).shape == (0, 128)
        cam = db.read_camera(cid)
        assert cam["params"] == [1200.0, 800.0, 600.0] and cam["has_prior_focal_length"] == 1 and cam["model"] == 0
        assert db.read_all_images() == [(ids[i], f"img{i:03d}.jpg", cid) for i in range(3)]
        m = np.array([[0, 5], [3, 1]], np.uint32)
        db.write_matches(ids[2], ids[1], m)                     # stored swapped (id1 < id2)
        assert np.array_equal(db.read_matches(ids[2], ids[1]), m)
        assert np.array_equal(db.read_matches(ids[1], ids[2]), m[:, ::-1])
        assert db.exists_matches(ids[1], ids[2]) and not db.exists_matches(ids[0], ids[1])
        F, H = rng.normal(size=(3, 3)), rng.normal(size=(3, 3))
        db.write_two_view_geometry(ids[2], ids[1], nat.TwoViewGeometry("UNCALIBRATED", F=F, H=H, inlier_matches=m))
        g = db.read_two_view_geometry(ids[1], ids[2])
        assert g.config == nat.TwoViewGeometryConfiguration.UNCALIBRATED and np.allclose(g.F, F.T)
        assert np.allclose(g.H, np.linalg.inv(H)) and np.array_equal(g.inlier_matches, m[:, ::-1])
        g2 = db.read_two_view_geometry(ids[2], ids[1])          # read back in the written orientation
        assert np.allclose(g2.F, F) and np.allclose(g2.H, H) and np.array_equal(g2.inlier_matches, m)
        assert db.read_two_view_geometry(ids[0], ids[1]) is None
        assert (db.num_images, db.num_cameras, db.num_matches, db.num_inlier_matches) == (3, 1, 2, 2)
        assert (db.num_keypoints, db.num_descriptors) == (10, 10)
        assert db.num_matched_image_pairs == 1 and db.num_verified_image_pairs == 1
        with pytest.raises(ValueError):
            db.read_camera(99)
        with pytest.raises(ValueError):
            db.write_descriptors(ids[1], np.zeros((4, 64), np.uint8))
        # transactions: rollback drops the write
        db.begin()
        db.write_matches(ids[0], ids[1], m)
        db.rollback()
        assert not db.exists_matches(ids[0], ids[1])
    assert nat.image_pair_to_pair_id(7, 3) == 3 * 2147483647 + 7
    assert nat.pair_id_to_image_pair(nat.image_pair_to_pair_id(7, 3)) == (3, 7)
    inv = nat.TwoViewGeometry("PLANAR", H=H, F=F, inlier_matches=m)
    inv.invert()
    assert np.allclose(inv.H, np.linalg.inv(H)) and np.allclose(inv.F, F.T) and np.array_equal(inv.inlier_matches, m[:, ::-1])
    ident = inv.cam2_from_cam1                                   # no pose estimated: the identity, also after invert()
    assert np.array_equal(ident.rotation.quat, [0, 0, 0, 1]) and np.array_equal(ident.translation, [0, 0, 0])
    assert "PLANAR" in repr(inv)


def test_relative_pose_storage(tmp_path):
    """qvec (w, x, y, z) / tvec columns: stored in the id1 < id2 frame, inverted for the swapped pair; Rigid3d
    accessors; the raw columns hold what plain sqlite3 + numpy expects."""
    rng = np.random.default_rng(5)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    t = rng.normal(size=3)
    m = np.array([[0, 5], [3, 1]], np.uint32)
    g = nat.TwoViewGeometry("CALIBRATED", inlier_matches=m, qvec=q, tvec=t, tri_angle=0.25)
    pose = g.cam2_from_cam1
    Rm = pose.rotation.matrix()
    assert np.allclose(pose.rotation.quat, [q[1], q[2], q[3], q[0]]) and np.allclose(pose.translation, t)
    assert np.allclose(Rm @ Rm.T, np.eye(3)) and np.isclose(np.linalg.det(Rm), 1.0) and g.tri_angle == 0.25
    assert np.allclose(pose.matrix(), np.c_[Rm, t])
    inv = pose.inverse()
    assert np.allclose(inv.rotation.matrix(), Rm.T) and np.allclose(inv.translation, -Rm.T @ t)
    built = nat.Rigid3d(nat.Rotation3d([q[1], q[2], q[3], q[0]]), t)         # (x, y, z, w) like the reference
    assert np.allclose(built.matrix(), pose.matrix()) and np.allclose(built.inverse().matrix(), inv.matrix())
    with nat.Database(tmp_path / "a.db") as a:
        _, ids, _, _ = _fill(a, np.random.default_rng(2))
        a.write_two_view_geometry(ids[2], ids[0], g)             # swapped: stored as the inverse pose
        ga = a.read_two_view_geometry(ids[0], ids[2])
        assert np.allclose(ga.cam2_from_cam1.matrix(), inv.matrix())
        back = a.read_two_view_geometry(ids[2], ids[0])         # read in the written orientation
        assert np.allclose(back.cam2_from_cam1.matrix(), pose.matrix())
    (qa, ta), = sqlite3.connect(tmp_path / "a.db").execute("SELECT qvec, tvec FROM two_view_geometries").fetchall()
    assert np.allclose(np.frombuffer(qa), [q[0], -q[1], -q[2], -q[3]]) and np.allclose(np.frombuffer(ta), inv.translation)


# COLMAP 3.9.1's schema as upstream's own scripts/python/database.py creates it (plain sqlite3)
_UPSTREAM_SCHEMA = """
CREATE TABLE IF NOT EXISTS cameras (camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL,
    width INTEGER NOT NULL, height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL);
CREATE TABLE IF NOT EXISTS images (image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE,
    camera_id INTEGER NOT NULL, prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL, prior_tx REAL, prior_ty REAL,
    prior_tz REAL, CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < 2147483647),
    FOREIGN KEY(camera_id) REFERENCES cameras(camera_id));
CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name);
CREATE TABLE IF NOT EXISTS keypoints (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL,
    data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS descriptors (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL,
    data BLOB, FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL,
    cols INTEGER NOT NULL, data BLOB, config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB, qvec BLOB, tvec BLOB);
"""


def test_database_is_interoperable_with_plain_sqlite3(tmp_path):
    """Rows written through plain sqlite3 in upstream's column layout read back through the C++ layer, and rows the
    C++ layer writes decode with numpy exactly as upstream's scripts/python/database.py decodes them."""
    rng = np.random.default_rng(1)
    m = np.array([[0, 5], [3, 1], [7, 2]], np.uint32)
    E, F, H = (rng.normal(size=(3, 3)) for _ in range(3))
    kp = rng.uniform(0, 1000, (10, 4)).astype(np.float32)
    d = rng.integers(0, 255, (10, 128)).astype(np.uint8)
    up = tmp_path / "upstream.db"
    con = sqlite3.connect(up)
    con.executescript(_UPSTREAM_SCHEMA)
    con.execute("INSERT INTO cameras VALUES (?, ?, ?, ?, ?, ?)", (None, 2, 1600, 1200, np.array([1200.0, 800.0, 600.0, -0.1]).tobytes(), 1))
    for i in range(3):
        con.execute("INSERT INTO images VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)", (None, f"img{i}.jpg", 1) + (None,) * 7)
    con.execute("INSERT INTO keypoints VALUES (?, ?, ?, ?)", (1, 10, 4, kp.tobytes()))
    con.execute("INSERT INTO descriptors VALUES (?, ?, ?, ?)", (1, 10, 128, d.tobytes()))
    con.execute("INSERT INTO matches VALUES (?, ?, ?, ?)", (1 * 2147483647 + 2, 3, 2, m.tobytes()))
    con.execute("INSERT INTO two_view_geometries VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                (1 * 2147483647 + 2, 2, 2, m[:2].tobytes(), 2, F.tobytes(), E.tobytes(), H.tobytes(),
                 np.array([1.0, 0, 0, 0]).tobytes(), np.zeros(3).tobytes()))
    con.commit()
    con.close()
    with nat.Database(up) as a:
        assert a.read_all_images() == [(i + 1, f"img{i}.jpg", 1) for i in range(3)]
        cam = a.read_camera(1)
        assert cam["model"] == 2 and cam["params"] == [1200.0, 800.0, 600.0, -0.1] and cam["has_prior_focal_length"] == 1
        assert np.array_equal(a.read_keypoints(1), kp) and np.array_equal(a.read_descriptors(1), d)
        assert np.array_equal(a.read_matches(1, 2), m) and np.array_equal(a.read_matches(2, 1), m[:, ::-1])
        g = a.read_two_view_geometry(1, 2)
        assert g.config.value == 2 and np.array_equal(g.inlier_matches, m[:2])
        assert np.array_equal(g.E, E) and np.array_equal(g.F, F) and np.array_equal(g.H, H)
    # the other direction: written by the C++ layer, decoded with numpy from the raw columns
    mine = tmp_path / "cxx.db"
    with nat.Database(mine) as a:
        _, ids, kp2, d2 = _fill(a, np.random.default_rng(2))
        a.write_matches(ids[1], ids[0], m)                       # swapped on the way in
        a.write_matches(ids[1], ids[2], np.zeros((0, 2), np.uint32))
        a.write_two_view_geometry(ids[0], ids[1], nat.TwoViewGeometry("CALIBRATED", E=E, F=F, H=H, inlier_matches=m[:2]))
    con = sqlite3.connect(mine)
    schema = {r[0]: r[1] for r in con.execute("SELECT name, sql FROM sqlite_master WHERE type = 'table'")}
    up_con = sqlite3.connect(up)
    for table in ("cameras", "images", "keypoints", "descriptors", "matches", "two_view_geometries"):
        cols = [r[1:3] for r in con.execute(f"PRAGMA table_info({table})")]
        assert cols == [r[1:3] for r in up_con.execute(f"PRAGMA table_info({table})")], table   # same names and types
    up_con.close()
    rows = con.execute("SELECT pair_id, rows, cols, data FROM matches ORDER BY 1").fetchall()
    assert rows[0][:3] == (ids[0] * 2147483647 + ids[1], 3, 2)
    assert np.array_equal(np.frombuffer(rows[0][3], np.uint32).reshape(3, 2), m[:, ::-1])
    assert rows[1][1:] == (0, 2, b"") or rows[1][1:] == (0, 2, None)          # empty match list: zero rows
    r = con.execute("SELECT rows, cols, data, config, F, E, H, qvec, tvec FROM two_view_geometries").fetchone()
    assert r[:2] == (2, 2) and r[3] == 2 and np.array_equal(np.frombuffer(r[2], np.uint32).reshape(2, 2), m[:2])
    assert np.array_equal(np.frombuffer(r[4]).reshape(3, 3), F) and np.array_equal(np.frombuffer(r[5]).reshape(3, 3), E)
    assert np.array_equal(np.frombuffer(r[6]).reshape(3, 3), H)
    assert np.array_equal(np.frombuffer(r[7]), [1, 0, 0, 0]) and np.array_equal(np.frombuffer(r[8]), [0, 0, 0])
    kr = con.execute("SELECT rows, cols, data FROM keypoints WHERE image_id = ?", (ids[0],)).fetchone()
    assert kr[:2] == kp2.shape and np.array_equal(np.frombuffer(kr[2], np.float32).reshape(kp2.shape), kp2)
    con.close()
    assert "matches" in schema


def test_argument_checks_before_any_gpu_work(tmp_path):
    with pytest.raises(ValueError, match=r"\[match_features.h:32\] Check Failed: File .* does not exist"):
        nat.match_exhaustive(tmp_path / "missing.db")
    with pytest.raises(ValueError, match="does not exist"):
        nat.match_sequential(str(tmp_path / "missing.db"))
    with pytest.raises(ValueError, match="does not exist"):
        nat.verify_matches(tmp_path / "missing.db", tmp_path / "pairs.txt")
    db = tmp_path / "db.db"
    nat.Database(db).close()
    with pytest.raises(ValueError, match="does not exist"):
        nat.verify_matches(db, tmp_path / "pairs.txt")
    with pytest.raises(ValueError, match="no CPU path"):
        nat.match_exhaustive(db, device=nat.Device.cpu)
    with pytest.raises(ValueError, match="no CPU path"):
        nat.match_exhaustive(db, device="cpu")                  # enum from its name at the call site
    with pytest.raises(TypeError):
        nat.match_exhaustive(db, sift_options=3)
    with pytest.raises(TypeError):
        nat.match_exhaustive(db, sift_options={"max_ratio": "high"})
    with pytest.raises(TypeError):
        nat.match_exhaustive(database=db)                       # keyword names are part of the contract
    cam = dict(model=0, width=1, height=1, params=[1, 0, 0])
    with pytest.raises(ValueError):
        nat.estimate_two_view_geometry(cam, np.zeros((3, 3)), cam, np.zeros((3, 2)))
    with pytest.raises(ValueError, match="points1.size"):
        nat.fundamental_matrix_estimation(np.zeros((4, 2)), np.zeros((5, 2)))
    with pytest.raises(ValueError, match="not supported"):
        nat.estimate_two_view_geometry(dict(cam, model=11), np.zeros((3, 2)), cam, np.zeros((3, 2)))
    with pytest.raises(ValueError, match="8 parameters"):      # OPENCV needs fx fy cx cy k1 k2 p1 p2
        nat.estimate_two_view_geometry(dict(cam, model="OPENCV"), np.zeros((3, 2)), cam, np.zeros((3, 2)))
    with pytest.raises(ValueError, match="unknown camera model"):
        nat.estimate_two_view_geometry(dict(cam, model="LENSBABY"), np.zeros((3, 2)), cam, np.zeros((3, 2)))


def test_no_cpu_fallback():
    """Without a GPU every compute entry point fails loudly with the C ABI's B2M_ENODEV message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nat.Context()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nat.squared_sampson_error(np.zeros((2, 2)), np.zeros((2, 2)), np.eye(3))


def test_ctrl_c_surfaces_as_keyboard_interrupt():
    """PyWait behaviour (R:helpers.h:335-347): the blocking call runs on a worker thread with the GIL
    released while the caller polls for signals; SIGINT becomes KeyboardInterrupt once the worker is done."""
    import os
    import signal
    import threading
    from pycolmap_b200 import _core
    import time
    t0 = time.time()
    threading.Timer(0.15, lambda: os.kill(os.getpid(), signal.SIGINT)).start()
    with pytest.raises(KeyboardInterrupt):
        _core._sleep_interruptible(0.6)
    assert 0.5 < time.time() - t0 < 5.0
    # other Python threads keep running while the call blocks (GIL released)
    ticks = []
    th = threading.Thread(target=lambda: [ticks.append(time.time()) or time.sleep(0.01) for _ in range(20)])
    th.start()
    _core._sleep_interruptible(0.3)
    th.join()
    assert len(ticks) == 20


def test_gpu_index_list_and_pair_sharding():
    """SiftMatchingOptions.gpu_index "0,1,2,3" = one matcher per listed GPU (R:pipeline/match_features.h:76-81);
    a chunk of pairs is cut into contiguous slices of equal distance-matrix cost."""
    from pycolmap_b200 import _core
    assert _core.parse_gpu_indices("-1") == [0] and _core.parse_gpu_indices("") == [0]
    assert _core.parse_gpu_indices("0,1, 2 ,3") == [0, 1, 2, 3] and _core.parse_gpu_indices("3,3,1") == [3, 1]
    for bad in ("a", "0,x", "1.5", "-2"):
        with pytest.raises(ValueError):
            _core.parse_gpu_indices(bad)
    rng = np.random.default_rng(0)
    n_feat = rng.integers(100, 8192, 40).astype(np.int32).tolist()
    pairs = np.concatenate(nat.exhaustive_pair_blocks(40, 7))
    cost = np.array([n_feat[a] * n_feat[b] for a, b in pairs], np.float64)
    for parts in (1, 2, 3, 8):
        cut = _core.split_pairs_by_cost(pairs, n_feat, parts)
        assert len(cut) == parts + 1 and cut[0] == 0 and cut[-1] == len(pairs) and sorted(cut) == cut
        shares = np.array([cost[cut[d]:cut[d + 1]].sum() for d in range(parts)]) / cost.sum()
        assert np.all(np.abs(shares - 1.0 / parts) < 0.02), shares     # within one pair's cost of the ideal
    assert _core.split_pairs_by_cost(np.zeros((0, 2), np.int32), n_feat, 4) == [0, 0, 0, 0, 0]
    assert _core.split_pairs_by_cost(pairs[:2], n_feat, 4)[-1] == 2     # fewer pairs than GPUs: empty slices allowed
    empty = [0] * 40                                                      # images without features still get dealt out
    cut = _core.split_pairs_by_cost(pairs, empty, 4)
    assert np.all(np.diff(cut) >= len(pairs) // 4 - 1)
    with pytest.raises(ValueError):
        _core.split_pairs_by_cost(np.array([[0, 99]], np.int32), n_feat, 2)


def test_batched_estimator_argument_checks():
    cam = dict(model=0, width=1, height=1, params=[1, 0, 0])
    with pytest.raises(ValueError, match="points1.size"):
        nat.estimate_two_view_geometries([(cam, np.zeros((3, 2)), cam, np.zeros((4, 2)))])
    with pytest.raises(ValueError):
        nat.estimate_two_view_geometries([(cam, np.zeros((3, 2)), cam)])
    with pytest.raises(ValueError, match="N x 2"):
        nat.estimate_two_view_geometries([(cam, np.zeros((3, 3)), cam, np.zeros((3, 2)))])
    with pytest.raises(ValueError, match="matches"):
        nat.estimate_two_view_geometries([(cam, np.zeros((3, 2)), cam, np.zeros((3, 2)), np.zeros((3, 3), np.uint32))])


def test_database_reference_surface(tmp_path):
    """The members the reference binds on Database (R:scene/database.h:10-47) beyond the counters."""
    with nat.Database(tmp_path / "db.db") as db:
        cid = db.write_camera(dict(model="SIMPLE_RADIAL", width=640, height=480, params=[500.0, 320.0, 240.0, 0.01],
                                   has_prior_focal_length=True))
        iid = db.write_image("a.jpg", cid)
        db.write_keypoints(iid, np.zeros((7, 2), np.float32))
        db.write_descriptors(iid, np.zeros((7, 128), np.uint8))
        cams = db.read_all_cameras()
        assert len(cams) == 1 and cams[0]["camera_id"] == cid and cams[0]["model"] == 2 and cams[0]["params"][3] == 0.01
        assert db.read_image(iid) == (iid, "a.jpg", cid) == db.read_image_with_name("a.jpg")
        assert db.num_keypoints_for_image(iid) == 7 == db.num_descriptors_for_image(iid) and db.num_keypoints_for_image(99) == 0
        assert db.pair_id_to_image_pair(db.image_pair_to_pair_id(9, 4)) == (4, 9)
        with pytest.raises(ValueError):
            db.read_image(12345)
        with pytest.raises(ValueError, match="4 parameters"):
            db.write_camera(dict(model="SIMPLE_RADIAL", width=1, height=1, params=[1.0, 0, 0]))
        tx = nat.DatabaseTransaction(db)                 # BEGIN ... COMMIT on release
        db.write_matches(iid, iid + 1, np.array([[0, 1]], np.uint32))
        del tx
        assert db.exists_matches(iid, iid + 1)


def test_spatial_pairs_match_oracle():
    """match_spatial's pair generator (R:pipeline/match_features.h:154-174, 237-244): GPS priors through WGS84 -> ECEF,
    k nearest located images per query, distance cut-off; against the oracle restatement."""
    from pycolmap_b200 import _core
    assert nat.SpatialMatchingOptions().todict() == dict(is_gps=True, ignore_z=True, max_num_neighbors=50, max_distance=100.0)
    ecef = R.gps_to_ecef(np.array([0.0, 90.0, 47.3769]), np.array([0.0, 0.0, 8.5417]), np.array([0.0, 0.0, 408.0]))
    assert np.allclose(ecef[0], [6378137.0, 0, 0]) and np.allclose(ecef[1], [0, 0, 6356752.314245], atol=1e-3)
    assert np.allclose(ecef[2], [4278990.0, 642680.0, 4670540.0], atol=300.0)          # Zurich, to a few hundred metres
    rng = np.random.default_rng(9)
    n = 60
    # a ~400 m x 400 m patch around Zurich: 1e-3 deg of latitude is ~111 m
    gps = np.c_[47.3769 + rng.uniform(-2e-3, 2e-3, n), 8.5417 + rng.uniform(-3e-3, 3e-3, n), rng.uniform(400, 450, n)]
    has = rng.random(n) < 0.85
    for kw in (dict(), dict(max_num_neighbors=5), dict(max_distance=40.0), dict(ignore_z=False, max_num_neighbors=8),
               dict(is_gps=False, max_distance=2e-3, max_num_neighbors=6), dict(max_num_neighbors=1)):
        o = nat.SpatialMatchingOptions(**kw)
        got = _core.spatial_pairs(gps, has.tolist(), o)
        want = np.array(R.spatial_pairs(gps, has, **o.todict()), np.int32).reshape(-1, 2)
        assert np.array_equal(got, want), kw
        assert all(has[a] and has[b] and a != b for a, b in got)
    assert len(_core.spatial_pairs(gps, has.tolist(), nat.SpatialMatchingOptions(max_num_neighbors=1))) == 0   # only itself
    assert len(_core.spatial_pairs(gps, [False] * n, nat.SpatialMatchingOptions())) == 0
    with pytest.raises(ValueError, match="vocabulary tree"):
        nat.match_vocabtree("whatever.db")


def test_option_docstrings_carry_type_and_default():
    """R:helpers.h:217-241: every option field documents "(type, default: value)"."""
    assert nat.SiftMatchingOptions.max_ratio.__doc__.endswith("(float, default: 0.8)")
    assert "Maximum distance ratio" in nat.SiftMatchingOptions.max_ratio.__doc__
    assert nat.ExhaustiveMatchingOptions.block_size.__doc__ == "(int, default: 50)"
    assert nat.TwoViewGeometryOptions.detect_watermark.__doc__ == "(bool, default: True)"
    assert nat.TwoViewGeometryOptions.ransac.__doc__ == "(RANSACOptions, default: RANSACOptions())"
