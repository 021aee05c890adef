"""CPU property tests (hypothesis): the two oracle implementations against each other on adversarial
shapes and entropies, a pure-numpy restatement of rows M1-M3 as a third opinion, the size-independent
properties of the matcher (swap symmetry under cross-check, injectivity, sortedness, idempotence of the
column-direction skip), and host-side invariants (pair generators, cost slicing, option dict round trips)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from helpers.native_import import load_native

nat = load_native()
from oracle import ransac as R

LUT = oracle.acos_lut()


def numpy_match(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True):
    """Rows M1-M3 written with numpy on the full distance matrix (independent of oracle_match.c)."""
    if len(d1) == 0 or len(d2) == 0:
        return np.zeros((0, 2), np.uint32)
    dist = d1.astype(np.int32) @ d2.astype(np.int32).T

    def one_way(D):
        best = D.max(1)
        arg = D.argmax(1)                                   # lowest index attaining the maximum
        Dm = D.copy()
        Dm[np.arange(len(D)), arg] = 0                      # multiset second best, floor 0
        second = Dm.max(1)
        a, b = LUT[np.minimum(best, 262144)], LUT[np.minimum(second, 262144)]
        ok = (best > 0) & ~(a > np.float32(max_distance)) & ~(a >= np.float32(max_ratio) * b)
        return np.where(ok, arg, -1)
    m12 = one_way(dist)
    keep = m12 >= 0
    if cross_check:
        m21 = one_way(dist.T)
        keep &= m21[np.maximum(m12, 0)] == np.arange(len(d1))
    i = np.flatnonzero(keep)
    return np.stack([i, m12[i]], 1).astype(np.uint32)


@st.composite
def descriptor_pairs(draw):
    n1 = draw(st.integers(0, 70))
    n2 = draw(st.integers(0, 70))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    levels = draw(st.sampled_from([2, 3, 16, 256]))         # few levels -> massive ties
    scale = draw(st.sampled_from([1, 40, 255]))             # saturated / tiny dot products
    rng = np.random.default_rng(seed)
    d1 = (rng.integers(0, levels, (n1, 128)) * (scale // max(levels - 1, 1) or 1)).clip(0, 255).astype(np.uint8)
    d2 = (rng.integers(0, levels, (n2, 128)) * (scale // max(levels - 1, 1) or 1)).clip(0, 255).astype(np.uint8)
    n_copy = draw(st.integers(0, min(n1, n2)))
    if n_copy:                                              # planted (near-)duplicates: true matches and ties
        d2[:n_copy] = d1[rng.permutation(n1)[:n_copy]]
    if n2 > 2 and draw(st.booleans()):
        d2[-1] = d2[0]                                      # duplicated best column
    if n1 > 1 and draw(st.booleans()):
        d1[0] = 0                                           # all-zero row
    return d1, d2


@settings(max_examples=120, deadline=None)
@given(descriptor_pairs(), st.sampled_from([(0.8, 0.7), (1.0, 3.2), (0.6, 0.3), (0.95, 1.0)]), st.booleans())
def test_three_implementations_agree(pair, thr, cross):
    d1, d2 = pair
    a = oracle.match_bruteforce(d1, d2, thr[0], thr[1], cross)
    b = oracle.fast_match_pair(d1, d2, thr[0], thr[1], cross)
    c = numpy_match(d1, d2, thr[0], thr[1], cross)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # size-independent properties
    assert np.all(np.diff(a[:, 0].astype(np.int64)) > 0)                      # sorted by idx1, each row once
    if cross:
        assert len(np.unique(a[:, 1])) == len(a)                               # injective
        s = oracle.fast_match_pair(d2, d1, thr[0], thr[1], True)               # swap symmetry
        assert np.array_equal(s[np.argsort(s[:, 1], kind="stable")][:, ::-1], a)
        # the column direction matters only where the row direction has candidates: a pair without any
        # one-way match has no mutual match (what launch_k1_filter_skip relies on)
        one_way = oracle.fast_match_pair(d1, d2, thr[0], thr[1], False)
        assert set(map(tuple, a)) <= set(map(tuple, one_way))
        if len(one_way) == 0:
            assert len(a) == 0


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 130), st.integers(2, 70))
def test_exhaustive_blocks_cover_every_pair_once(n, bs):
    blocks = nat.exhaustive_pair_blocks(n, bs)
    got = np.concatenate(blocks) if blocks else np.zeros((0, 2), np.int32)
    assert np.array_equal(got, np.array(R.exhaustive_pairs(range(n), bs), np.int32).reshape(-1, 2))
    assert len({(min(a, b), max(a, b)) for a, b in got}) == len(got) == n * (n - 1) // 2
    assert all(a != b for a, b in got)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 200), st.integers(1, 25), st.booleans())
def test_sequential_pairs_properties(n, overlap, quadratic):
    got = nat.sequential_pairs(n, overlap, quadratic)
    assert np.array_equal(got, np.array(R.sequential_pairs(range(n), overlap, quadratic), np.int32).reshape(-1, 2))
    assert len({tuple(p) for p in got}) == len(got)                            # no duplicates
    assert np.all(got[:, 0] < got[:, 1]) and (len(got) == 0 or got.max() < n)
    offs = set((got[:, 1] - got[:, 0]).tolist())
    allowed = set(range(1, overlap + 1)) | ({1 << k for k in range(overlap)} if quadratic else set())
    assert offs <= allowed


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 60), st.integers(1, 9), st.integers(0, 2 ** 31 - 1))
def test_cost_slices_partition_the_pair_list(n_img, parts, seed):
    from pycolmap_b200 import _core
    rng = np.random.default_rng(seed)
    n_feat = rng.integers(0, 9000, n_img).astype(np.int32).tolist()
    blocks = nat.exhaustive_pair_blocks(n_img, 7)
    pairs = np.concatenate(blocks) if blocks else np.zeros((0, 2), np.int32)
    cut = _core.split_pairs_by_cost(pairs, n_feat, parts)
    assert len(cut) == parts + 1 and cut[0] == 0 and cut[-1] == len(pairs) and sorted(cut) == cut
    cost = np.array([max(1, n_feat[a] * n_feat[b]) for a, b in pairs], np.float64)
    if len(pairs) and parts > 1:
        worst = max(cost[cut[d]:cut[d + 1]].sum() for d in range(parts))
        assert worst <= cost.sum() / parts + cost.max() + 1e-6                 # within one pair of the ideal share


option_values = st.fixed_dictionaries({}, optional={
    "min_num_inliers": st.integers(0, 1000), "min_E_F_inlier_ratio": st.floats(0, 1), "max_H_inlier_ratio": st.floats(0, 1),
    "detect_watermark": st.booleans(), "force_H_use": st.booleans(),
    "ransac": st.fixed_dictionaries({}, optional={"max_error": st.floats(0.1, 50), "confidence": st.floats(0.5, 0.99999),
                                                  "min_num_trials": st.integers(0, 10 ** 5), "max_num_trials": st.integers(1, 10 ** 6)})})


@settings(max_examples=80, deadline=None)
@given(option_values)
def test_option_dict_round_trip(d):
    import pickle
    o = nat.TwoViewGeometryOptions(d)
    for k, v in d.items():
        if k != "ransac":
            assert getattr(o, k) == v
    for k, v in d.get("ransac", {}).items():
        assert getattr(o.ransac, k) == v
    assert nat.TwoViewGeometryOptions(o.todict()) == o and pickle.loads(pickle.dumps(o)) == o
    base = nat.TwoViewGeometryOptions()
    base.mergedict(d)
    assert base == o
