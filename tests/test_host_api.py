"""CPU: host-side mirror of the reference's binding surface -- option classes (MakeDataclass
behaviour, R:helpers.h:40-283), database layer (COLMAP schema), pair generators, argument checks."""
import copy
import os
import pickle

import numpy as np
import pytest

import pycolmap_b200 as pb
from oracle import ransac as R
from pycolmap_b200 import pipeline
from pycolmap_b200.database import Database, image_pair_to_pair_id, pair_id_to_image_pair


def test_option_defaults_match_reference():
    s = pb.SiftMatchingOptions()
    assert (s.max_ratio, s.max_distance, s.cross_check, s.max_num_matches, s.guided_matching, s.num_threads,
            s.gpu_index) == (0.8, 0.7, True, 32768, False, -1, "-1")
    assert pb.ExhaustiveMatchingOptions().block_size == 50
    q = pb.SequentialMatchingOptions()
    assert (q.overlap, q.quadratic_overlap, q.loop_detection) == (10, True, False)
    r = pb.RANSACOptions()      # Python-side defaults of the binding (R:optim/bindings.h:10-18)
    assert (r.max_error, r.min_inlier_ratio, r.confidence, r.min_num_trials, r.max_num_trials) == (
        4.0, 0.01, 0.9999, 1000, 100000)
    t = pb.TwoViewGeometryOptions()  # .ransac keeps the C++ ctor defaults
    assert (t.min_num_inliers, t.min_E_F_inlier_ratio, t.max_H_inlier_ratio, t.detect_watermark) == (15, 0.95, 0.8, True)
    assert (t.ransac.max_error, t.ransac.confidence, t.ransac.min_num_trials, t.ransac.max_num_trials,
            t.ransac.min_inlier_ratio) == (4.0, 0.999, 100, 10000, 0.25)


def test_dataclass_behaviour():
    t = pb.TwoViewGeometryOptions({"min_num_inliers": 30, "ransac": {"max_error": 2.0}})
    assert t.min_num_inliers == 30 and t.ransac.max_error == 2.0 and t.ransac.confidence == 0.999
    t2 = pb.TwoViewGeometryOptions(min_num_inliers=7)
    t2.mergedict({"ransac": {"min_num_trials": 5}})
    assert t2.todict()["ransac"]["min_num_trials"] == 5 and t2.todict()["min_num_inliers"] == 7
    assert "min_num_inliers" in t2.summary() and "ransac" in t2.summary()
    with pytest.raises(AttributeError):
        t.mergedict({"no_such_field": 1})
    with pytest.raises(AttributeError):
        t.no_such_field = 3
    with pytest.raises(TypeError):
        t.min_num_inliers = "many"
    with pytest.raises(TypeError):
        pb.SiftMatchingOptions(max_ratio="high")
    s = pb.SiftMatchingOptions(max_ratio=1)          # int -> float coercion like pybind11
    assert isinstance(s.max_ratio, float)
    c = copy.deepcopy(t)
    c.ransac.max_error = 9.0
    assert t.ransac.max_error == 2.0
    assert pickle.loads(pickle.dumps(t)) == t
    assert pb.TwoViewGeometryOptions.coerce({"min_num_inliers": 3}).min_num_inliers == 3   # implicit dict -> Options
    assert pb.Device("auto") if False else pipeline._enum_from(pb.Device, "cuda") == pb.Device.cuda
    with pytest.raises(ValueError):
        pipeline._enum_from(pb.Device, "tpu")
    assert pb.TwoViewGeometryConfiguration.PLANAR_OR_PANORAMIC == 6 and pb.TwoViewGeometryConfiguration.WATERMARK == 7


def test_pair_generators_match_oracle():
    for n, bs in [(1, 50), (2, 2), (49, 50), (50, 50), (51, 50), (101, 50), (7, 2), (130, 64)]:
        got = np.concatenate(list(pipeline.exhaustive_pair_blocks(n, bs)) or [np.zeros((0, 2), np.int32)])
        want = np.array(R.exhaustive_pairs(range(n), bs), np.int32).reshape(-1, 2)
        assert np.array_equal(got, want)                       # same pairs in the same visiting order
    for n, ov, q in [(100, 3, True), (50, 10, False), (1000, 20, True)]:
        assert np.array_equal(pipeline.sequential_pairs(n, ov, q),
                              np.array(R.sequential_pairs(range(n), ov, q), np.int32).reshape(-1, 2))


def test_database_roundtrip(tmp_path):
    path = tmp_path / "db.db"
    rng = np.random.default_rng(0)
    with Database(path) as db:
        cid = db.add_camera(0, 1600, 1200, [1200.0, 800.0, 600.0], True)
        ids = [db.add_image(f"img{i:03d}.jpg", cid) for i in range(3)]
        kp = rng.uniform(0, 1000, (10, 6)).astype(np.float32)
        d = rng.integers(0, 255, (10, 128)).astype(np.uint8)
        db.write_keypoints(ids[0], kp)
        db.write_descriptors(ids[0], d)
        assert np.array_equal(db.read_keypoints(ids[0]), kp) and np.array_equal(db.read_descriptors(ids[0]), d)
        assert db.read_camera(cid)["params"] == [1200.0, 800.0, 600.0] and db.read_camera(cid)["has_prior_focal_length"] == 1
        m = np.array([[0, 5], [3, 1]], np.uint32)
        db.write_matches(ids[2], ids[1], m)                     # stored swapped (id1 < id2)
        assert np.array_equal(db.read_matches(ids[2], ids[1]), m)
        assert np.array_equal(db.read_matches(ids[1], ids[2]), m[:, ::-1])
        F = rng.normal(size=(3, 3))
        db.write_two_view_geometry(ids[2], ids[1], 3, m, F=F)
        g = db.read_two_view_geometry(ids[1], ids[2])
        assert g["config"] == 3 and np.allclose(g["F"], F.T) and np.array_equal(g["inlier_matches"], m[:, ::-1])
        assert db.num_images == 3 and db.num_cameras == 1 and db.num_matches == 2 and db.num_inlier_matches == 2
        assert db.num_matched_image_pairs == 1 and db.num_verified_image_pairs == 1
    pid = image_pair_to_pair_id(7, 3)
    assert pid == 3 * 2147483647 + 7 and pair_id_to_image_pair(pid) == (3, 7)


def test_argument_checks_before_any_gpu_work(tmp_path):
    with pytest.raises(ValueError, match="does not exist"):
        pb.match_exhaustive(tmp_path / "missing.db")
    with pytest.raises(ValueError, match="does not exist"):
        pb.verify_matches(tmp_path / "missing.db", tmp_path / "pairs.txt")
    db = tmp_path / "db.db"
    Database(db).close()
    with pytest.raises(ValueError, match="does not exist"):
        pb.verify_matches(db, tmp_path / "pairs.txt")
    with pytest.raises(ValueError, match="no CPU path"):
        pb.match_exhaustive(db, device=pb.Device.cpu)
    with pytest.raises(TypeError):
        pb.match_exhaustive(db, sift_options=3)
    with pytest.raises(ValueError):
        pb.estimate_two_view_geometry(dict(model=0, width=1, height=1, params=[1, 0, 0]), np.zeros((3, 3)),
                                      dict(model=0, width=1, height=1, params=[1, 0, 0]), np.zeros((3, 2)))
