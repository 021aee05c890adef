"""GPU: database-driven pipelines (match_exhaustive / match_sequential / verify_matches) on a
synthetic COLMAP database (BASELINE configs[0] shape, reduced) and the estimator functions."""
import numpy as np
import pytest

import oracle
import pycolmap_b200 as pb
from oracle import ransac as R
from helpers import scenes
import sqlite3

from pycolmap_b200 import synthetic as syn

Database = pb.Database


def _count(path, table):
    con = sqlite3.connect(path)
    try:
        return con.execute(f"SELECT COUNT(*) FROM {table}").fetchone()[0]
    finally:
        con.close()

pytestmark = pytest.mark.gpu


def _make_db(path, n_images=12, n_feat=1024, seed=2):
    scene = syn.make_scene(n_images, n_feat, seed=seed, window_images=2.5)
    with Database(path) as db:
        cid = db.add_camera(0, 1600, 1200, [1200.0, 800.0, 600.0], True)
        db.begin()
        for i in range(n_images):
            iid = db.add_image(f"frame{i:04d}.png", cid)
            kp = np.zeros((n_feat, 6), np.float32)
            kp[:, :2] = scene["kpts"][i].numpy()
            db.write_keypoints(iid, kp)
            db.write_descriptors(iid, scene["desc"][i].numpy())
        db.commit()
    return scene


def test_match_exhaustive_database(tmp_path):
    path = tmp_path / "scene.db"
    scene = _make_db(path)
    pb.match_exhaustive(path, matching_options={"block_size": 5})
    descs = [d.numpy() for d in scene["desc"]]
    with Database(path) as db:
        ids = [r[0] for r in db.read_all_images()]
        n_pairs = len(ids) * (len(ids) - 1) // 2
        assert _count(path, "matches") == n_pairs and _count(path, "two_view_geometries") == n_pairs
        verified = 0
        for a in range(len(ids)):
            for b in range(a + 1, len(ids)):
                want = oracle.fast_match_pair(descs[a], descs[b])
                got = db.read_matches(ids[a], ids[b])
                g = db.read_two_view_geometry(ids[a], ids[b])
                if len(want) < 15:       # write rule: stored empty, default geometry
                    assert len(got) == 0 and int(g.config) == 0 and len(g.inlier_matches) == 0
                else:
                    # ExhaustiveFeatureMatcher visits some pairs as (b, a): same match set, rows ordered
                    # by the other image's index (mutual nearest neighbours are symmetric under cross-check)
                    assert np.array_equal(got[np.argsort(got[:, 0], kind="stable")], want)
                    # CALIBRATED normally; UNCALIBRATED when E keeps < 95 % of F's inliers on a noisy pair
                    assert int(g.config) in (2, 3, 6) and len(g.inlier_matches) >= 15
                    verified += 1
        assert verified >= 10 and db.num_verified_image_pairs == verified
    # resume semantics: a second run finds everything stored and changes nothing
    before = open(path, "rb").read()
    pb.match_exhaustive(path)
    assert open(path, "rb").read() == before


def test_match_sequential_and_verify_matches(tmp_path):
    path = tmp_path / "seq.db"
    scene = _make_db(path, n_images=10, n_feat=768, seed=5)
    pb.match_sequential(path, matching_options=pb.SequentialMatchingOptions(overlap=2, quadratic_overlap=False))
    with Database(path) as db:
        ids = [r[0] for r in db.read_all_images()]
        assert _count(path, "matches") == 9      # COLMAP 3.9.1: overlap = 2 -> ONE linear neighbour (idx1 + 0 is the self pair)
        # drop the geometries, verify them again from the stored matches through a pair list
        db.clear_two_view_geometries()
        names = [r[1] for r in db.read_all_images()]
    pairs = tmp_path / "pairs.txt"
    pairs.write_text("# comment\n\n" + "\n".join(f"{names[i]} {names[i + 1]}" for i in range(9)) + "\n")
    pb.verify_matches(path, pairs)
    assert _count(path, "two_view_geometries") == 9
    with Database(path) as db:
        assert db.num_verified_image_pairs >= 5


def test_estimator_functions():
    rng = np.random.default_rng(11)
    p1, p2, planted = scenes.two_view_scene(rng, 400, 0.3, "general")
    g = pb.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2)
    assert g.config == pb.TwoViewGeometryConfiguration.CALIBRATED and abs(len(g.inlier_matches) - planted.sum()) <= 4
    g2 = pb.estimate_calibrated_two_view_geometry(scenes.CAM_NOPRIOR, p1, scenes.CAM_NOPRIOR, p2)
    assert g2.config == pb.TwoViewGeometryConfiguration.CALIBRATED
    g3 = pb.estimate_two_view_geometry(scenes.CAM_NOPRIOR, p1, scenes.CAM_NOPRIOR, p2,
                                       options={"ransac": {"max_error": 2.0}})
    assert g3.config == pb.TwoViewGeometryConfiguration.UNCALIBRATED
    f = pb.fundamental_matrix_estimation(p1, p2)
    assert f is not None and abs(f["num_inliers"] - planted.sum()) <= 5 and f["inliers"].dtype == bool
    e = pb.essential_matrix_estimation(p1, p2, scenes.CAM, scenes.CAM)
    assert e is not None and abs(e["num_inliers"] - planted.sum()) <= 5
    assert pb.homography_matrix_estimation(p1[:3], p2[:3]) is None
    q1, q2, pl = scenes.two_view_scene(rng, 300, 0.2, "planar")
    h = pb.homography_matrix_estimation(q1, q2, {"max_error": 4.0, "min_num_trials": 100, "confidence": 0.999})
    assert h is not None and abs(h["num_inliers"] - pl.sum()) <= 4
    E = rng.normal(size=(3, 3))
    assert np.allclose(pb.squared_sampson_error(p1, p2, E), R.squared_sampson_error(p1, p2, E), rtol=1e-12)
    inv = pb.TwoViewGeometry(g.config, g.E, g.F, g.H, g.inlier_matches)
    inv.invert()
    assert np.allclose(inv.F, g.F.T) and np.array_equal(inv.inlier_matches, g.inlier_matches[:, ::-1])
