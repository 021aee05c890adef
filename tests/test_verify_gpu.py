"""GPU: batched LO-RANSAC verifier (K2/K3) vs the CPU oracle -- statistical parity (inlier counts
within +-1 % / mask Jaccard >= 0.99 on planted scenes, equal configuration), through the C ABI."""
import numpy as np
import pytest

import oracle
from oracle import ransac as R
from helpers import scenes
from pycolmap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _mask(inl, n):
    m = np.zeros(n, bool)
    m[inl[:, 0]] = True
    return m


@pytest.mark.parametrize("kind,cam,expect", [
    ("general", scenes.CAM, R.CALIBRATED), ("planar", scenes.CAM, R.PLANAR_OR_PANORAMIC),
    ("rotation", scenes.CAM, R.PLANAR_OR_PANORAMIC), ("general", scenes.CAM_NOPRIOR, R.UNCALIBRATED)])
@pytest.mark.parametrize("n,noise", [(60, 0.0), (400, 0.0), (1500, 0.5)])
def test_planted_scene(ctx, kind, cam, expect, n, noise):
    rng = np.random.default_rng(n + len(kind))
    p1, p2, planted = scenes.two_view_scene(rng, n, 0.3, kind, noise)
    res, inl = ctx.estimate_two_view_geometry(cam, p1, cam, p2)
    g = R.estimate_two_view_geometry(cam, p1, cam, p2, seed=3)
    assert res.config == expect == g.config
    got = _mask(inl, n)
    assert (np.diff(inl[:, 0].astype(np.int64)) > 0).all() and np.array_equal(inl[:, 0], inl[:, 1])
    tol = max(2, int(0.01 * planted.sum()))
    if noise == 0.0 and kind == "general":
        jac = (got & planted).sum() / max(1, (got | planted).sum())
        assert jac >= 0.99, jac
    if noise == 0.0:
        # planar / panoramic scenes leave F under-constrained (a few outliers may fit the chosen
        # epipolar model, in the reference too): every planted inlier must still be found
        assert (got & planted).sum() >= 0.99 * planted.sum()
        assert (got & ~planted).sum() <= (3 if kind == "general" else max(6, int(0.05 * n)))
    slack = 1 if kind == "general" else 4     # degenerate F on planes varies run to run upstream as well
    assert abs(int(res.n_inliers) - len(g.inlier_matches)) <= slack * max(tol, int(0.01 * len(g.inlier_matches))), (
        res.n_inliers, len(g.inlier_matches), planted.sum())
    # per-model inlier counts agree with the oracle's LO-RANSAC runs
    for a, b in ((res.nE, g.nE), (res.nF, g.nF)):
        assert abs(a - b) <= slack * max(3, int(0.02 * max(a, b))), (res.nE, res.nF, res.nH, g.nE, g.nF, g.nH)


def test_degenerate_and_random(ctx):
    rng = np.random.default_rng(1)
    p1, p2, _ = scenes.two_view_scene(rng, 10, 0.0)
    res, inl = ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2)
    assert res.config == R.DEGENERATE and len(inl) == 0
    a, b = rng.uniform(0, 1000, (40, 2)), rng.uniform(0, 1000, (40, 2))
    res, inl = ctx.estimate_two_view_geometry(scenes.CAM, a, scenes.CAM, b)
    assert res.config == R.DEGENERATE and len(inl) == 0
    res, inl = ctx.estimate_two_view_geometry(scenes.CAM, np.zeros((0, 2)), scenes.CAM, np.zeros((0, 2)))
    assert res.config == R.DEGENERATE


def test_matches_argument_and_errors(ctx):
    rng = np.random.default_rng(2)
    p1, p2, planted = scenes.two_view_scene(rng, 300, 0.2)
    perm = rng.permutation(300)
    matches = np.stack([np.arange(300), perm], 1).astype(np.uint32)
    p2s = np.empty_like(p2)
    p2s[perm] = p2
    res, inl = ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2s, matches)
    assert res.config == R.CALIBRATED and abs(int(res.n_inliers) - planted.sum()) <= 3
    assert all(perm[a] == b for a, b in inl)
    with pytest.raises(ValueError):
        ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2[:10])          # identity needs equal sizes
    with pytest.raises(ValueError):
        ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2, np.array([[0, 999]], np.uint32))


def test_watermark(ctx):
    rng = np.random.default_rng(5)
    n = 200
    p1 = np.c_[rng.uniform(0, 1600, n), rng.uniform(0, 100, n)]
    p2 = p1 + [7.0, 3.0]
    res, _ = ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2)
    assert res.config == R.WATERMARK
    res, _ = ctx.estimate_two_view_geometry(scenes.CAM, p1, scenes.CAM, p2, opts=ctx.tvg_opts(detect_watermark=0))
    assert res.config == R.PLANAR_OR_PANORAMIC


def test_single_model_estimators(ctx):
    rng = np.random.default_rng(6)
    p1, p2, planted = scenes.two_view_scene(rng, 500, 0.3, "general")
    r = ctx.ransac_model(1, p1, p2)
    assert r is not None and abs(r["num_inliers"] - planted.sum()) <= 5
    assert (r["inliers"] & planted).sum() >= 0.99 * planted.sum()
    assert (ctx.squared_sampson_error(p1[planted], p2[planted], r["model"]) < 16.0 + 1e-9).mean() > 0.99
    n1, n2 = R.cam_from_img(scenes.CAM, p1), R.cam_from_img(scenes.CAM, p2)
    r = ctx.ransac_model(0, n1, n2, ctx.ransac_opts(max_error=4.0 / 1200))
    assert r is not None and abs(r["num_inliers"] - planted.sum()) <= 5
    q1, q2, pl = scenes.two_view_scene(rng, 500, 0.4, "planar")
    r = ctx.ransac_model(2, q1, q2)
    assert r is not None and abs(r["num_inliers"] - pl.sum()) <= 5
    assert ctx.ransac_model(2, q1[:3], q2[:3]) is None          # fewer than the minimal sample -> None
    E = rng.normal(size=(3, 3))
    assert np.allclose(ctx.squared_sampson_error(p1, p2, E), R.squared_sampson_error(p1, p2, E), rtol=1e-12)


def test_pipeline_match_and_verify(ctx):
    """Image-set path: raw matches bit-exact vs the matcher oracle, verification vs the verifier oracle
    on the same raw matches, and the controller's write rule (row P3)."""
    scene = syn.make_scene(16, 1024, seed=1, window_images=3.0)
    descs = [d.numpy() for d in scene["desc"]]
    kpts = [k.numpy() for k in scene["kpts"]]
    cams = scene["cameras"]
    ctx.set_images(descs, kpts, cams)
    pairs = syn.exhaustive_pairs(16)
    res = ctx.match_pairs(pairs, tvg=ctx.tvg_opts())
    want = oracle.fast_match_pairs(np.concatenate(descs), [len(d) for d in descs], pairs)
    n_verified = n_checked = 0
    for k, (i, j) in enumerate(pairs):
        v = res.view(k)
        raw = want[k]
        if len(raw) < 15:
            assert v.n_matches == 0 and v.config == R.UNDEFINED and v.n_inliers == 0
            continue
        assert np.array_equal(res.matches(k), raw)
        n_verified += 1
        if n_checked < 6:
            n_checked += 1
            g = R.estimate_two_view_geometry(cams[i], kpts[i].astype(np.float64), cams[j],
                                             kpts[j].astype(np.float64), raw, seed=k)
            exp_cfg = g.config if len(g.inlier_matches) >= 15 else R.UNDEFINED
            assert v.config == exp_cfg, (k, v.config, g.config)
            assert abs(int(v.n_inliers) - (len(g.inlier_matches) if exp_cfg else 0)) <= max(
                2, int(0.01 * len(g.inlier_matches)))
            inl = res.inlier_matches(k)
            rawset = {tuple(x) for x in raw.tolist()}
            assert all(tuple(x) in rawset for x in inl.tolist())
    assert n_verified >= 10


def test_guided_matching_replaces_inliers(ctx):
    """K1g (row G1): with SiftMatchingOptions.guided_matching the inlier matches of a verified pair are
    the brute-force matches under the geometric filter of its F / H -- bit-exact vs the oracle given the
    same model (the model itself comes from the GPU verifier)."""
    scene = syn.make_scene(10, 1024, seed=4, window_images=2.5)
    descs = [d.numpy() for d in scene["desc"]]
    kpts = [k.numpy() for k in scene["kpts"]]
    ctx.set_images(descs, kpts, scene["cameras"])
    pairs = syn.exhaustive_pairs(10)
    plain = ctx.match_pairs(pairs, tvg=ctx.tvg_opts())
    guided = ctx.match_pairs(pairs, sift=ctx.sift_opts(guided_matching=1), tvg=ctx.tvg_opts())
    n_guided = 0
    for k, (i, j) in enumerate(pairs):
        vp, vg = plain.view(k), guided.view(k)
        assert np.array_equal(plain.matches(k), guided.matches(k))          # raw matches are untouched
        if vp.config in (R.CALIBRATED, R.UNCALIBRATED, R.PLANAR_OR_PANORAMIC) and vp.n_inliers >= 15:
            assert vg.config == vp.config
            kind = 1 if vp.config == R.PLANAR_OR_PANORAMIC else 0
            model = np.array(vg.H if kind else vg.F).reshape(3, 3)
            want = oracle.match_guided(descs[i], kpts[i], descs[j], kpts[j], kind, model, 4.0)
            got = guided.inlier_matches(k)
            if len(want) >= 15:
                assert np.array_equal(got, want), (k, len(got), len(want))
                n_guided += 1
                # the geometric filter removes ambiguous second-best candidates: never fewer than ~the inliers
                assert len(got) >= 0.9 * vp.n_inliers
        else:
            assert vg.n_inliers == vp.n_inliers
    assert n_guided >= 5
