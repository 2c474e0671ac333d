"""GPU: edge cases the reference's tests exercise (tiny / degenerate inputs, no-split trees, full 256-slot
columns, config resets) and size-independent properties at a large size (conservation of gradient mass per
column, every row in exactly one leaf, parent = left + right)."""
import os

import numpy as np
import pytest

from helpers import compare_trees, synth_identity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built_lib):
    import lightgbm_b200 as lgb
    from oracle import oracle_py
    return lgb, oracle_py


def _run(lgb, orc, bins, g, h, num_bin=255, **cfg):
    lay = lgb.Layout.identity(bins, num_bin=num_bin)
    L = lgb.B200TreeLearner(lgb.Config(**cfg))
    L.init(lay)
    t = L.train(g, h)
    o = orc.train_tree(lay, g, h, **cfg)
    return L, lay, t, o


@pytest.mark.parametrize("n,f", [(1, 1), (2, 1), (31, 3), (33, 2), (64, 1)])
def test_tiny_inputs_grow_the_oracle_tree(mods, n, f):
    lgb, orc = mods
    rng = np.random.default_rng(n * 7 + f)
    bins = rng.integers(0, 8, (n, f), dtype=np.uint8)
    g = rng.normal(size=n).astype(np.float32); h = np.ones(n, np.float32)
    L, lay, t, o = _run(lgb, orc, bins, g, h, num_leaves=4, min_data_in_leaf=1, min_sum_hessian_in_leaf=0.5)
    compare_trees(t, o, 1e-5)
    lb, lc, idx = L.get_partition(t.num_leaves)
    assert sorted(idx.tolist()) == list(range(n))


def test_no_split_possible_returns_single_leaf(mods):
    lgb, orc = mods
    n = 500
    bins = np.full((n, 4), 7, np.uint8)                         # every feature constant
    g = np.random.default_rng(0).normal(size=n).astype(np.float32); h = np.ones(n, np.float32)
    L, lay, t, o = _run(lgb, orc, bins, g, h, num_leaves=8)
    assert t.num_leaves == o.num_leaves == 1
    np.testing.assert_allclose(t.leaf_value, o.leaf_value, rtol=1e-6)
    # min_data_in_leaf larger than half the data: BeforeFindBestSplit refuses
    bins2, y, g2, h2 = synth_identity(300, 5, seed=2)
    L2, _, t2, o2 = _run(lgb, orc, bins2, g2, h2, num_leaves=8, min_data_in_leaf=200)
    assert t2.num_leaves == o2.num_leaves == 1
    # all features masked out
    L3 = lgb.B200TreeLearner(lgb.Config(num_leaves=8)); L3.init(lgb.Layout.identity(bins2))
    L3.set_feature_mask(np.zeros(5, np.uint8))
    assert L3.train(g2, h2).num_leaves == 1


def test_full_256_slot_column_and_two_leaves(mods):
    lgb, orc = mods
    n = 6000
    rng = np.random.default_rng(5)
    bins = rng.integers(0, 256, (n, 3), dtype=np.uint8)           # stored values 0..255: lo=1, 255 entries + slot 0
    g = ((bins[:, 1] > 200) * 2.0 - 1 + 0.1 * rng.normal(size=n)).astype(np.float32); h = np.ones(n, np.float32)
    L, lay, t, o = _run(lgb, orc, bins, g, h, num_bin=256, num_leaves=2)
    matched, _ = compare_trees(t, o, 1e-5)
    assert matched == 1 and t.splits[0]["feature"] == 1


def test_reset_config_changes_tree_size_and_regularisation(mods):
    lgb, orc = mods
    bins, y, g, h = synth_identity(20000, 10, seed=9)
    lay = lgb.Layout.identity(bins)
    L = lgb.B200TreeLearner(lgb.Config(num_leaves=7)); L.init(lay)
    t7 = L.train(g, h)
    L.reset_config(lgb.Config(num_leaves=31, lambda_l2=5.0))
    t31 = L.train(g, h)
    o31 = orc.train_tree(lay, g, h, num_leaves=31, lambda_l2=5.0)
    assert t7.num_leaves == 7 and t31.num_leaves == 31
    compare_trees(t31, o31, 1e-5)


def test_large_size_invariants(mods):
    """Size-independent properties at a size the oracle would not finish in seconds (default 2M x 128; set
    LGBM_B200_FULL=1 for the BASELINE 10M x 1024 shape)."""
    lgb, _ = mods
    full = os.environ.get("LGBM_B200_FULL") == "1"
    n, f, leaves = (10_000_000, 1024, 127) if full else (2_000_000, 128, 63)
    rng = np.random.default_rng(123)
    bins = rng.integers(0, 255, (n, f), dtype=np.uint8)
    g = rng.normal(size=n).astype(np.float32)
    h = (0.5 + rng.random(n)).astype(np.float32)
    lay = lgb.Layout.identity(bins)
    L = lgb.B200TreeLearner(lgb.Config(num_leaves=leaves)); L.init(lay)
    hist, _ = L.construct_histogram(g, h)
    sg, sh = float(g.astype(np.float64).sum()), float(h.astype(np.float64).sum())
    # conservation: every column's bins add up to the totals (linearity of the scatter-add)
    np.testing.assert_allclose(hist[:, :, 1].sum(axis=1), sh, rtol=2e-6)
    np.testing.assert_allclose(hist[:, :, 0].sum(axis=1), sg, atol=2e-6 * np.abs(g).sum())
    t = L.train(g, h)
    assert t.num_leaves == leaves
    lb, lc, idx = L.get_partition(t.num_leaves)
    assert lc.sum() == n and np.array_equal(np.sort(idx), np.arange(n))          # a permutation: no row lost
    np.testing.assert_array_equal(lc, t.leaf_count)
    for s in t.splits:                                                             # parent = left + right
        assert s["left_count"] > 0 and s["right_count"] > 0
    np.testing.assert_allclose(t.root_sum_hessian, sh, rtol=1e-9)
    # children of the root carry the root's mass
    s0 = t.splits[0]
    np.testing.assert_allclose(s0["left_sum_hessian"] + s0["right_sum_hessian"], sh, rtol=1e-6)
    np.testing.assert_allclose(s0["left_sum_gradient"] + s0["right_sum_gradient"], sg, atol=1e-6 * np.abs(g).sum())
    # idempotence: the same call grows the same tree (graph replay, fixed-order accumulation)
    t2 = L.train(g, h)
    assert np.array_equal(t.splits, t2.splits) and np.array_equal(t.leaf_value, t2.leaf_value)
