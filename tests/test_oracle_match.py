"""CPU: the oracle against the known-answer cases of SURVEY.md section 8(c) (derived from the
published semantics of U:feature/sift.cc -- the reference holds no golden vectors: parity unpinned)
and the literal vs streaming-SIMD implementations against each other."""
import numpy as np
import pytest

import oracle
from pycolmap_b200 import synthetic as syn


def test_identity_and_reverse():
    rng = np.random.default_rng(0)
    d = syn.sift_like(rng, 300)
    m = oracle.match_bruteforce(d, d)
    assert np.array_equal(m, np.stack([np.arange(300)] * 2, 1))
    m = oracle.match_bruteforce(d, d[::-1])
    assert np.array_equal(m[:, 1], 299 - m[:, 0]) and len(m) == 300


def test_duplicate_best_and_zero_rows():
    rng = np.random.default_rng(1)
    d1 = syn.sift_like(rng, 50)
    d2 = syn.sift_like(rng, 60)
    d2[10] = d1[5]
    d2[40] = d1[5]          # duplicated best -> ratio test a >= r*a rejects row 5
    d1[7] = 0               # all-zero descriptor never matches
    m = oracle.match_bruteforce(d1, d2, cross_check=False)
    assert 5 not in m[:, 0] and 7 not in m[:, 0]


def test_tie_break_lowest_index():
    # d2 rows 3 and 9 identical -> equal dot with every d1 row; second strictly smaller elsewhere
    d1 = np.zeros((1, 128), np.uint8)
    d1[0, :4] = 255
    d2 = np.zeros((12, 128), np.uint8)
    d2[:, 4:8] = 40
    d2[3, :4] = 200
    d2[9, :4] = 200
    m = oracle.match_bruteforce(d1, d2, max_ratio=1.0, max_distance=np.pi, cross_check=False)
    # duplicated maximum: second == best -> a >= 1.0*a -> rejected
    assert len(m) == 0
    d2[9, 0] = 199
    m = oracle.match_bruteforce(d1, d2, max_ratio=1.0, max_distance=np.pi, cross_check=False)
    assert m.tolist() == [[0, 3]]
    d2[3, 0] = 198          # now 9 is strictly better
    m = oracle.match_bruteforce(d1, d2, max_ratio=1.0, max_distance=np.pi, cross_check=False)
    assert m.tolist() == [[0, 9]]


def test_loose_thresholds_match_every_unique_max():
    rng = np.random.default_rng(2)
    d1, d2 = syn.sift_like(rng, 40), syn.sift_like(rng, 70)
    m = oracle.match_bruteforce(d1, d2, max_ratio=1.0, max_distance=np.pi, cross_check=False)
    dist = d1.astype(np.int64) @ d2.astype(np.int64).T
    for i in range(40):
        row = dist[i]
        uniq = (row == row.max()).sum() == 1
        assert (i in m[:, 0]) == bool(uniq)
        if uniq:
            assert m[m[:, 0] == i][0, 1] == row.argmax()


@pytest.mark.parametrize("n1,n2,common", [(1, 1, 1), (7, 300, 5), (513, 255, 100), (1000, 1200, 400)])
@pytest.mark.parametrize("cross", [True, False])
def test_fast_equals_literal(n1, n2, common, cross):
    rng = np.random.default_rng(n1 * 7 + n2)
    d1, d2, _ = syn.matching_pair(rng, n1, n2, common)
    a = oracle.match_bruteforce(d1, d2, cross_check=cross)
    b = oracle.fast_match_pair(d1, d2, cross_check=cross)
    assert np.array_equal(a, b)
    if common >= 100:
        assert len(a) > common // 2   # planted matches are found


def test_fast_batch_equals_single():
    rng = np.random.default_rng(5)
    nf = [300, 257, 128, 512]
    descs = [syn.sift_like(rng, n) for n in nf]
    descs[1][:100] = syn.perturb(rng, descs[0][:100])
    pairs = syn.exhaustive_pairs(4)
    res = oracle.fast_match_pairs(np.concatenate(descs), nf, pairs, n_threads=3)
    for (i, j), m in zip(pairs, res):
        assert np.array_equal(m, oracle.match_bruteforce(descs[i], descs[j]))


def test_guided_filter_keeps_only_consistent():
    rng = np.random.default_rng(6)
    n = 200
    d1 = syn.sift_like(rng, n)
    d2 = syn.perturb(rng, d1)
    kp1 = rng.uniform(0, 1000, (n, 2)).astype(np.float32)
    H = np.array([[1.0, 0.02, 5.0], [-0.01, 1.0, -3.0], [1e-5, 0, 1.0]])
    p = np.c_[kp1, np.ones(n)] @ H.T
    kp2 = (p[:, :2] / p[:, 2:]).astype(np.float32)
    kp2[:50] += 100.0       # first 50 violate the homography
    m = oracle.match_guided(d1, kp1, d2, kp2, 1, H, 4.0)
    assert set(m[:, 0]) == set(range(50, n))
    assert np.array_equal(m[:, 0], m[:, 1])


def test_acos_lut_monotone():
    lut = oracle.acos_lut()
    assert lut[0] == np.float32(np.pi / 2) and lut[-1] == 0.0
    assert (np.diff(lut) <= 0).all()
