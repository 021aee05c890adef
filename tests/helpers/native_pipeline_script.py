"""Runs INSIDE a subprocess of tests/test_native_pipeline_cpu.py with the mock libb200match.so in front of
the real one (LD_LIBRARY_PATH): drives the C++ host layer end to end on CPU and checks the databases."""
import os
import shutil
import sqlite3
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pycolmap_b200 import _core as nat  # noqa: E402
import oracle  # noqa: E402
from oracle import ransac as R  # noqa: E402
from pycolmap_b200 import synthetic as syn  # noqa: E402

tmp = sys.argv[1]
TABLES = ("cameras", "images", "keypoints", "descriptors", "matches", "two_view_geometries")
MIN = 15


def dump(path):
    con = sqlite3.connect(path)
    out = {t: con.execute(f"SELECT * FROM {t} ORDER BY 1").fetchall() for t in TABLES}
    con.close()
    return out


def fake_geometry(a, b, m):
    """tests/helpers/mock_b200match.cpp fake_geometry."""
    if len(m) < MIN or len(m[::2]) < MIN:
        return 0, np.zeros((0, 2), np.uint32), np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
    k = np.arange(9)
    E = (100.0 * a + b + 0.125 * k).reshape(3, 3)
    F = (-(100.0 * a + b) - 0.25 * k).reshape(3, 3)
    H = (np.where(k % 4 == 0, 2.0 + a, 0.0) + 0.0625 * k * (b + 1)).reshape(3, 3)
    return 2, m[::2], E, F, H


# the mock answers for "GPUs" 0..7; the real library would refuse without a device
c = nat.Context(device=3)
c.close()

n_img, n_feat = 9, 512
scene = syn.make_scene(n_img, n_feat, seed=7, window_images=2.5)
descs = [d.numpy() for d in scene["desc"]]


def make_db(path, names=None, camera=(0, [1200.0, 800.0, 600.0])):
    with nat.Database(path) as db:
        cid = db.add_camera(camera[0], 1600, 1200, camera[1], True)
        db.begin()
        ids = []
        for i in range(n_img):
            iid = db.add_image(names[i] if names else f"frame{i:04d}.png", cid)
            kp = np.zeros((n_feat, 4), np.float32)
            kp[:, :2] = scene["kpts"][i].numpy()
            db.write_keypoints(iid, kp)
            db.write_descriptors(iid, descs[i])
            ids.append(iid)
        db.commit()
    return ids


# ---- match_exhaustive: every pair once, visited in ExhaustiveFeatureMatcher order, write rules ----------
a_db = os.path.join(tmp, "a.db")
ids = make_db(a_db)
nat.match_exhaustive(a_db, matching_options={"block_size": 4})
visits = R.exhaustive_pairs(range(n_img), 4)
assert len(visits) == n_img * (n_img - 1) // 2
n_verified = 0
with nat.Database(a_db) as db:
    assert db.num_rows("matches") == db.num_rows("two_view_geometries") == len(visits)
    for i1, i2 in visits:
        want = oracle.fast_match_pair(descs[i1], descs[i2])          # in the VISITED orientation
        got = db.read_matches(ids[i1], ids[i2])
        g = db.read_two_view_geometry(ids[i1], ids[i2])
        cfg, inl, E, F, H = fake_geometry(i1, i2, want)
        if len(want) < MIN:                                           # stored empty, default geometry
            assert len(got) == 0 and g.config.value == 0 and len(g.inlier_matches) == 0
            continue
        assert np.array_equal(got, want), (i1, i2)
        assert g.config.value == cfg and np.array_equal(g.inlier_matches, inl)
        if cfg == 0:            # 15..29 matches: raw matches stored, too few placeholder inliers -> default geometry
            continue
        n_verified += 1
        assert np.array_equal(g.E, E) and np.array_equal(g.F, F) and np.allclose(g.H, H, rtol=1e-9)
        # and from the other side: columns swapped, F / E transposed, H inverted
        gi = db.read_two_view_geometry(ids[i2], ids[i1])
        assert np.array_equal(gi.inlier_matches, inl[:, ::-1]) and np.array_equal(gi.F, F.T)
        assert np.allclose(gi.H, np.linalg.inv(H), rtol=1e-9)
    assert n_verified >= 8 and db.num_verified_image_pairs == n_verified <= db.num_matched_image_pairs
assert any(i1 > i2 for i1, i2 in visits)                              # the reversed-visit branch was exercised

# ---- resume: everything stored -> nothing to do, file untouched --------------------------------------------
before = open(a_db, "rb").read()
nat.match_exhaustive(a_db, matching_options={"block_size": 4})
assert open(a_db, "rb").read() == before
ref_dump = dump(a_db)
# partial resume: a pair without a matches row is matched + verified again (same result); pairs whose matches are
# stored but whose geometry is missing are VERIFIED FROM THE STORED MATCHES, never re-matched (imported / custom
# matches survive a match_* run; the mock's estimator entry point marks its geometries with the pair (7, 9))
con = sqlite3.connect(a_db)
rows = [r[0] for r in con.execute("SELECT pair_id FROM two_view_geometries WHERE rows >= 15 ORDER BY pair_id")]
con.execute("DELETE FROM two_view_geometries WHERE pair_id IN (?, ?)", (rows[0], rows[3]))
custom = np.frombuffer(con.execute("SELECT data FROM matches WHERE pair_id = ?", (rows[3],)).fetchone()[0], np.uint32).reshape(-1, 2)[:-1]
con.execute("UPDATE matches SET rows = ?, data = ? WHERE pair_id = ?", (len(custom), custom.tobytes(), rows[3]))   # "imported"
con.execute("DELETE FROM matches WHERE pair_id = ?", (rows[5],))
con.commit()
con.close()
nat.match_exhaustive(a_db, matching_options={"block_size": 4})
resumed = dump(a_db)
assert [r for r in resumed["matches"] if r[0] != rows[3]] == [r for r in ref_dump["matches"] if r[0] != rows[3]]
assert np.array_equal(np.frombuffer([r for r in resumed["matches"] if r[0] == rows[3]][0][3], np.uint32).reshape(-1, 2), custom)
for ra, rb in zip(resumed["two_view_geometries"], ref_dump["two_view_geometries"]):
    if ra[0] in (rows[0], rows[3]):
        stored = np.frombuffer([r for r in resumed["matches"] if r[0] == ra[0]][0][3], np.uint32).reshape(-1, 2)
        cfg, inl, E, F, H = fake_geometry(7, 9, stored)
        assert ra[1] == len(inl) and ra[4] == cfg and np.array_equal(np.frombuffer(ra[3], np.uint32).reshape(-1, 2), inl)
        Es, Fs = np.frombuffer(ra[6]).reshape(3, 3), np.frombuffer(ra[5]).reshape(3, 3)   # transposed if visited as (id2, id1)
        assert (np.array_equal(Es, E) and np.array_equal(Fs, F)) or (np.array_equal(Es, E.T) and np.array_equal(Fs, F.T))
    else:
        assert ra == rb

# ---- gpu_index lists: same database whatever the number of GPUs and the block size -------------------------
for k, (gpus, bs) in enumerate((("0,1,2", 4), ("0, 1,2,3,4,5,6,7", 4), ("5", 4), ("-1", 3), ("1,0", 50))):
    p = os.path.join(tmp, f"g{k}.db")
    make_db(p)
    nat.match_exhaustive(p, sift_options={"gpu_index": gpus}, matching_options={"block_size": bs})
    d = dump(p)
    if bs == 4:
        assert d == ref_dump, gpus                                     # byte-identical tables
    else:  # another block size visits pairs in another orientation: same pair set, same match sets
        assert [r[:2] for r in d["matches"]] == [r[:2] for r in ref_dump["matches"]]
# the default "-1" = every visible GPU (upstream); without NCCL every GPU takes the whole set: same database
for k, env in enumerate(({"MOCK_B2M_DEVICES": "3"}, {"MOCK_B2M_DEVICES": "4", "MOCK_B2M_NO_NCCL": "1"})):
    os.environ.update(env)
    assert nat.parse_gpu_indices("-1") == list(range(int(env["MOCK_B2M_DEVICES"])))
    p = os.path.join(tmp, f"all{k}.db")
    make_db(p)
    nat.match_exhaustive(p, matching_options={"block_size": 4})
    assert dump(p) == ref_dump, env
    for key in env:
        del os.environ[key]
assert nat.parse_gpu_indices("-1") == [0] and nat.parse_gpu_indices("2,-1") == [2, 0]
try:
    nat.match_exhaustive(a_db, sift_options={"gpu_index": "0,9"})
    raise SystemExit("device 9 must be refused")
except RuntimeError:
    pass
try:
    nat.match_exhaustive(a_db, matching_options={"block_size": 1})
    raise SystemExit("block_size 1 must be refused")
except ValueError:
    pass

# ---- match_sequential: images ordered by NAME, overlap window ------------------------------------------------
s_db = os.path.join(tmp, "s.db")
names = [f"img{99 - i:03d}.png" for i in range(n_img)]                 # name order = reverse id order
sids = make_db(s_db, names)
nat.match_sequential(s_db, matching_options={"overlap": 3, "quadratic_overlap": False})
order = sorted(range(n_img), key=lambda i: names[i])
seq = R.sequential_pairs(range(n_img), 3, False)     # COLMAP 3.9.1: overlap 3 -> two linear neighbours
with nat.Database(s_db) as db:
    assert db.num_rows("matches") == len(seq) == (n_img - 1) + (n_img - 2)
    for k1, k2 in seq:
        i1, i2 = order[k1], order[k2]
        want = oracle.fast_match_pair(descs[i1], descs[i2])
        got = db.read_matches(sids[i1], sids[i2])
        assert np.array_equal(got, want if len(want) >= MIN else want[:0]), (k1, k2)
try:
    nat.match_sequential(s_db, matching_options={"loop_detection": True})
    raise SystemExit("loop detection must be refused")
except ValueError:
    pass

# ---- verify_matches: pair list; stored matches are verified, unknown pairs are matched first ---------------
con = sqlite3.connect(s_db)
con.execute("DELETE FROM two_view_geometries")
con.commit()
con.close()
listed = [(order[0], order[1]), (order[1], order[0]), (order[3], order[2]), (order[0], order[4]), (order[2], order[2])]
pairs_txt = os.path.join(tmp, "pairs.txt")
with open(pairs_txt, "w") as f:
    f.write("# header\n\n" + "".join(f"{names[a]} {names[b]}\n" for a, b in listed) + "ghost.png img000.png\nonlyone\n")
nat.verify_matches(s_db, pairs_txt)
with nat.Database(s_db) as db:
    assert db.num_rows("two_view_geometries") == 3                      # duplicate, self pair, ghost dropped
    for a, b in ((order[0], order[1]), (order[3], order[2])):         # had stored matches: verified from them
        m = db.read_matches(sids[a], sids[b])
        g = db.read_two_view_geometry(sids[a], sids[b])
        cfg, inl, E, F, H = fake_geometry(7, 9, m)                      # the mock's single-problem entry point
        assert g.config.value == cfg and np.array_equal(g.inlier_matches, inl), (a, b)
    a, b = order[0], order[4]                                          # not matched before (offset 4 > overlap)
    assert db.exists_matches(sids[a], sids[b]) and db.exists_inlier_matches(sids[a], sids[b])
    want = oracle.fast_match_pair(descs[a], descs[b])
    assert np.array_equal(db.read_matches(sids[a], sids[b]), want if len(want) >= MIN else want[:0])

# ---- verify_matches on a database WITHOUT descriptors (learned-feature matches imported by the user, e.g. hloc):
# only keypoints, cameras and the stored matches are needed
v_db = os.path.join(tmp, "nodesc.db")
vids = make_db(v_db)
con = sqlite3.connect(v_db)
con.execute("DELETE FROM descriptors")
con.commit()
con.close()
imported = {}
with nat.Database(v_db) as db:
    for a, b in ((0, 1), (2, 1), (3, 4)):
        m = np.stack([np.arange(40, dtype=np.uint32), np.arange(40, dtype=np.uint32)[::-1]], 1)
        m = m[: 40 if (a, b) != (3, 4) else 9]                         # (3, 4): fewer than min_num_inliers matches
        db.write_matches(vids[a], vids[b], m)
        imported[(a, b)] = m
vnames = [f"frame{i:04d}.png" for i in range(n_img)]
with open(pairs_txt, "w") as f:
    f.write("".join(f"{vnames[a]} {vnames[b]}\n" for a, b in imported))
nat.verify_matches(v_db, pairs_txt)
with nat.Database(v_db) as db:
    assert db.num_rows("two_view_geometries") == 3 and db.num_descriptors == 0
    for (a, b), m in imported.items():
        g = db.read_two_view_geometry(vids[a], vids[b])
        if len(m) < MIN:   # write rule: raw matches below min_num_inliers are rewritten empty, default geometry
            assert len(db.read_matches(vids[a], vids[b])) == 0 and g.config.value == 0 and len(g.inlier_matches) == 0
        else:
            cfg, inl, E, F, H = fake_geometry(7, 9, m)
            assert np.array_equal(db.read_matches(vids[a], vids[b]), m)
            assert g.config.value == cfg and np.array_equal(g.inlier_matches, inl)
with open(pairs_txt, "w") as f:
    f.write(f"{vnames[0]} {vnames[5]}\n")                                # no stored matches: now descriptors ARE needed
try:
    nat.verify_matches(v_db, pairs_txt)
    raise SystemExit("matching without descriptors must be refused")
except ValueError as e:
    assert "keypoints.rows == descriptors.rows" in str(e)

# ---- cameras: COLMAP's default SIMPLE_RADIAL is accepted, FOV is refused; inconsistent features are refused ----
r_db = os.path.join(tmp, "radial.db")
make_db(r_db, camera=(2, [1200.0, 800.0, 600.0, -0.1]))
nat.match_exhaustive(r_db, matching_options={"block_size": 4})
assert [r[:2] for r in dump(r_db)["matches"]] == [r[:2] for r in ref_dump["matches"]]
f_db = os.path.join(tmp, "unknown_model.db")
make_db(f_db, camera=(11, [1200.0, 1200.0, 800.0, 600.0, 0.5]))
try:
    nat.match_exhaustive(f_db)
    raise SystemExit("an unknown camera model id must be refused")
except ValueError as e:
    assert "not supported" in str(e)
bad = os.path.join(tmp, "bad.db")
shutil.copy(a_db, bad)
with nat.Database(bad) as db:
    db.write_keypoints(ids[2], np.zeros((n_feat - 1, 2), np.float32))
    db.clear_matches()
try:
    nat.match_exhaustive(bad)
    raise SystemExit("keypoints.rows != descriptors.rows must be refused")
except ValueError as e:
    assert "keypoints.rows == descriptors.rows" in str(e)

# ---- match_spatial: images on a line 30 m apart (Cartesian priors), 2 nearest neighbours within 70 m ---------------
sp_db = os.path.join(tmp, "spatial.db")
with nat.Database(sp_db) as db:
    cid = db.add_camera(0, 1600, 1200, [1200.0, 800.0, 600.0], True)
    db.begin()
    sp_ids = []
    for i in range(n_img):
        prior = None if i == 4 else [30.0 * i, 0.0, 5.0 * i]            # image 4 has no prior: never paired
        iid = db.add_image(f"frame{i:04d}.png", cid, prior)
        kp = np.zeros((n_feat, 2), np.float32)
        kp[:] = scene["kpts"][i].numpy()
        db.write_keypoints(iid, kp)
        db.write_descriptors(iid, descs[i])
        sp_ids.append(iid)
    db.commit()
opts = {"is_gps": False, "max_num_neighbors": 3, "max_distance": 70.0}
nat.match_spatial(sp_db, matching_options=opts)
prior = np.array([[30.0 * i, 0.0, 5.0 * i] for i in range(n_img)])
want_pairs = {(min(a, b), max(a, b)) for a, b in R.spatial_pairs(prior, [i != 4 for i in range(n_img)], **dict(nat.SpatialMatchingOptions(opts).todict()))}
assert want_pairs and all(4 not in p and abs(p[0] - p[1]) <= 2 for p in want_pairs)
con = sqlite3.connect(sp_db)
stored = {tuple(int(x) for x in nat.pair_id_to_image_pair(r[0])) for r in con.execute("SELECT pair_id FROM matches")}
con.close()
assert stored == {(sp_ids[a], sp_ids[b]) for a, b in want_pairs}

# ---- multiple_models in the pipelines: verified pairs are re-estimated through the estimator entry point ------------
m_db = os.path.join(tmp, "multi.db")
make_db(m_db)
nat.match_exhaustive(m_db, matching_options={"block_size": 4}, verification_options={"multiple_models": True})
n_multi = 0
with nat.Database(m_db) as db:
    for i1, i2 in visits:
        want = oracle.fast_match_pair(descs[i1], descs[i2])
        g = db.read_two_view_geometry(ids[i1], ids[i2])
        if fake_geometry(i1, i2, want)[0] == 0:                      # not verified by the batch: left alone
            assert g.config.value == 0
            continue
        cfg, inl, E, F, H = fake_geometry(7, 9, want)                  # the mock's estimator entry point signature
        assert g.config.value == cfg == 2 and np.array_equal(g.inlier_matches, inl) and np.array_equal(g.E, E), (i1, i2)
        n_multi += 1
assert n_multi == n_verified
assert [r[:3] for r in dump(m_db)["matches"]] == [r[:3] for r in ref_dump["matches"]]

# ---- max_num_matches: longer images are truncated to a prefix with a warning (WarnIfMaxNumMatchesReachedGPU) -----
t_db = os.path.join(tmp, "trunc.db")
make_db(t_db)
KEEP = 300
nat.match_exhaustive(t_db, sift_options={"max_num_matches": KEEP}, matching_options={"block_size": 4})
with nat.Database(t_db) as db:
    n_rows = 0
    for i1, i2 in visits:
        want = oracle.fast_match_pair(descs[i1][:KEEP], descs[i2][:KEEP])
        got = db.read_matches(ids[i1], ids[i2])
        if len(want) < MIN:
            assert len(got) == 0
            continue
        assert np.array_equal(got, want) and got.max() < KEEP, (i1, i2)
        n_rows += 1
    assert n_rows >= 5

# ---- low-level context through the mock: plumbing of options and result objects ------------------------------
c = nat.Context()
c.set_images(descs[:3])
res = c.match_pairs(np.array([(0, 1), (2, 1)], np.int32), nat.SiftMatchingOptions(cross_check=False, max_ratio=0.9))
assert len(res) == 2 and res.image_pair(1) == (2, 1)
assert np.array_equal(res.matches(1), oracle.fast_match_pair(descs[2], descs[1], max_ratio=0.9, cross_check=False))
assert res.total_matches == len(res.matches(0)) + len(res.matches(1)) and res.num_verified == 0
res.free()
c.close()
gs = nat.estimate_two_view_geometries([(dict(model=0, width=1, height=1, params=[1.0, 0, 0]), np.zeros((40, 2)),
                                        dict(model="PINHOLE", width=1, height=1, params=[1.0, 1.0, 0, 0]), np.zeros((40, 2)))])
assert len(gs) == 1 and gs[0].config.value == 2 and len(gs[0].inlier_matches) == 20
print("NATIVE-PIPELINE-OK")
