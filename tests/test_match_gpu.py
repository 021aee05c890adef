"""GPU parity: K1 (tcgen05 int8 GEMM + fused top-2 / ratio / cross-check) vs the CPU oracle,
bit-exact on match indices, through the C ABI."""
import numpy as np
import pytest

import oracle
from pycolmap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _check(ctx, d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=1):
    got = ctx.match_pair(d1, d2, {"max_ratio": max_ratio, "max_distance": max_distance, "cross_check": bool(cross_check)})
    # SiftMatchingOptions holds doubles; the matcher takes them as float32 (double -> float at use, row B4)
    want = oracle.fast_match_pair(d1, d2, max_ratio=float(np.float32(max_ratio)), max_distance=float(np.float32(max_distance)),
                                  cross_check=bool(cross_check))
    assert got.dtype == np.uint32 and got.shape[1] == 2
    assert np.array_equal(got, want), (len(got), len(want))
    return got


@pytest.mark.parametrize("n1,n2,common", [
    (1, 1, 1), (5, 3, 2), (127, 129, 60), (128, 256, 100), (255, 257, 128), (300, 1000, 250),
    (1024, 1024, 512), (2000, 777, 400), (4096, 4096, 1500)])
def test_random_pairs_bit_exact(ctx, n1, n2, common):
    rng = np.random.default_rng(n1 * 31 + n2)
    d1, d2, _ = syn.matching_pair(rng, n1, n2, common)
    got = _check(ctx, d1, d2)
    if common >= 60:
        assert len(got) > common // 2
    _check(ctx, d1, d2, cross_check=0)
    _check(ctx, d1, d2, max_ratio=1.0, max_distance=float(np.pi))
    _check(ctx, d1, d2, max_ratio=0.6, max_distance=0.5)


@pytest.mark.parametrize("n1,n2,common", [(9000, 12000, 2000), (20000, 600, 300), (600, 20000, 300)])
def test_more_than_8192_features(ctx, n1, n2, common):
    # resolve kernel: staged slots up to 16384 features, direct global-memory scan beyond
    rng = np.random.default_rng(n1 + 7 * n2)
    d1, d2, _ = syn.matching_pair(rng, n1, n2, common)
    got = _check(ctx, d1, d2)
    assert len(got) > common // 2
    _check(ctx, d1, d2, cross_check=0)


def test_identity_and_reverse(ctx):
    rng = np.random.default_rng(0)
    d = syn.sift_like(rng, 700)
    m = _check(ctx, d, d)
    assert np.array_equal(m, np.stack([np.arange(700)] * 2, 1))
    m = _check(ctx, d, d[::-1].copy())
    assert np.array_equal(m[:, 1], 699 - m[:, 0])


def test_ties_zeros_duplicates(ctx):
    rng = np.random.default_rng(3)
    d1 = syn.sift_like(rng, 600)
    d2 = syn.sift_like(rng, 900)
    d2[10] = d1[5]
    d2[700] = d1[5]              # duplicated best across two column tiles
    d2[300] = d1[8]
    d2[301] = d1[8]              # duplicated best inside a tile
    d1[7] = 0
    d2[20] = 0
    d1[100:110] = d1[100]        # duplicate rows
    d2[500:505] = d1[100]
    for kw in ({}, {"cross_check": 0}, {"max_ratio": 1.0, "max_distance": float(np.pi), "cross_check": 0}):
        _check(ctx, d1, d2, **kw)


def test_low_entropy_many_ties(ctx):
    # few distinct values -> massive ties in dot products; exercises lowest-index tie-breaking
    rng = np.random.default_rng(4)
    d1 = (rng.integers(0, 2, (500, 128)) * 45).astype(np.uint8)
    d2 = (rng.integers(0, 2, (800, 128)) * 45).astype(np.uint8)
    for kw in ({"max_ratio": 1.0, "max_distance": float(np.pi), "cross_check": 0},
               {"max_ratio": 1.0, "max_distance": float(np.pi)}, {}):
        _check(ctx, d1, d2, **kw)


def test_saturated_descriptors(ctx):
    # dot products far above 2^18 (clamped in the acos) up to the 255^2*128 maximum
    rng = np.random.default_rng(5)
    d1 = rng.integers(200, 256, (300, 128)).astype(np.uint8)
    d2 = rng.integers(200, 256, (400, 128)).astype(np.uint8)
    d1[0] = 255
    d2[7] = 255
    for kw in ({}, {"max_ratio": 1.0, "max_distance": float(np.pi), "cross_check": 0}):
        _check(ctx, d1, d2, **kw)


def test_empty_inputs(ctx):
    rng = np.random.default_rng(6)
    d = syn.sift_like(rng, 10)
    e = np.zeros((0, 128), np.uint8)
    assert len(ctx.match_pair(e, d)) == 0 and len(ctx.match_pair(d, e)) == 0 and len(ctx.match_pair(e, e)) == 0


def test_image_set_batches_equal_oracle(ctx):
    rng = np.random.default_rng(7)
    nf = [512, 300, 1024, 129, 0, 768, 256, 1000]
    descs = [syn.sift_like(rng, n) for n in nf]
    for a, b, k in [(0, 2, 300), (1, 5, 200), (2, 7, 500), (3, 6, 100)]:
        descs[b][:k] = syn.perturb(rng, descs[a][:k])
    ctx.set_images(descs)
    pairs = syn.exhaustive_pairs(len(nf))
    # repeat the list so that several scheduler batches are exercised
    pairs = np.concatenate([pairs, pairs[::-1, ::-1]])
    res = ctx.match_pairs(pairs)
    want = oracle.fast_match_pairs(np.concatenate(descs), nf, pairs)
    assert len(res) == len(pairs)
    total = 0
    for k in range(len(pairs)):
        got = res.matches(k)
        assert np.array_equal(got, want[k]), (k, pairs[k])
        total += len(got)
    assert res.total_matches == total and total > 1000


def test_small_batch_scheduler(ctx):
    import pycolmap_b200 as pb
    c = pb.Context(device=0, pair_batch=3)
    rng = np.random.default_rng(8)
    descs = [syn.sift_like(rng, 200 + 10 * i) for i in range(6)]
    descs[1][:150] = syn.perturb(rng, descs[0][:150])
    c.set_images(descs)
    pairs = syn.exhaustive_pairs(6)
    res = c.match_pairs(pairs)
    want = oracle.fast_match_pairs(np.concatenate(descs), [len(d) for d in descs], pairs)
    for k in range(len(pairs)):
        assert np.array_equal(res.matches(k), want[k])
    c.close()


def test_large_pair_properties(ctx):
    # 8192 x 8192 (BASELINE config size): size-independent properties + oracle equality
    rng = np.random.default_rng(9)
    d1, d2, gt = syn.matching_pair(rng, 8192, 8192, 3000)
    m = _check(ctx, d1, d2)
    assert (np.diff(m[:, 0].astype(np.int64)) > 0).all()          # sorted, unique idx1
    assert len(np.unique(m[:, 1])) == len(m)                      # cross-check => injective
    mt = ctx.match_pair(d2, d1)
    assert np.array_equal(mt[np.argsort(mt[:, 1])][:, ::-1], m)   # symmetric under swapping images
    planted = {tuple(x) for x in gt.tolist()}
    assert len(planted & {tuple(x) for x in m.tolist()}) > 2500
