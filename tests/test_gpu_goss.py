"""GPU: GOSS on the device (lightgbm_b200/csrc/goss_kernel.cuh) against the definition of GOSSStrategy::Helper
(reference src/boosting/goss.hpp:118-167): exact where the reference is deterministic (threshold, scaling, ordering),
statistical where it draws random numbers; and the tree grown on the sampled set equals the oracle's tree on the same
set and the same rescaled gradients."""
import numpy as np
import pytest

from helpers import compare_trees, synth_identity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,top,other", [(200_000, 0.2, 0.1), (50_001, 0.05, 0.3), (1000, 0.5, 0.5)])
def test_goss_sample_definition(built_lib, n, top, other):
    import lightgbm_b200 as lgb
    from lightgbm_b200.tree_learner import DeviceArray
    bins, y, g, h = synth_identity(n, 8, seed=n)
    rng = np.random.default_rng(5)
    h = (0.05 + rng.random(n)).astype(np.float32)
    L = lgb.B200TreeLearner(lgb.Config(num_leaves=15))
    L.init(lgb.Layout.identity(bins))
    dg, dh = DeviceArray(n * 4).upload(g), DeviceArray(n * 4).upload(h)
    cnt = L.goss_sample(dg, dh, top, other, seed=7, iteration=11)
    bag = L.get_bagging_data(cnt)
    g2, h2 = dg.download(np.float32, n), dh.download(np.float32, n)
    assert np.all(np.diff(bag) > 0)                                   # ascending, no duplicates
    key = np.abs(g * h)
    top_k, other_k = max(1, int(n * top)), int(n * other)
    thr = np.partition(key, n - top_k)[n - top_k]                     # ArgMaxAtK(top_k - 1): the k-th largest
    is_top = key >= thr
    in_bag = np.zeros(n, bool); in_bag[bag] = True
    assert np.all(in_bag[is_top])                                     # every row at or above the threshold is kept (goss.hpp:147)
    drawn = in_bag & ~is_top
    mult = np.float32((n - top_k) / other_k)
    np.testing.assert_array_equal(g2[drawn], g[drawn] * mult)         # goss.hpp:157-158
    np.testing.assert_array_equal(h2[drawn], h[drawn] * mult)
    np.testing.assert_array_equal(g2[~drawn], g[~drawn])              # top rows and dropped rows are untouched
    np.testing.assert_array_equal(h2[~drawn], h[~drawn])
    # number drawn ~ Binomial(n - #top, other_k / (n - top_k))
    m, p = int((~is_top).sum()), other_k / (n - top_k)
    assert abs(int(drawn.sum()) - m * p) <= 6 * np.sqrt(m * p * (1 - p)) + 2
    # a different iteration draws a different subset, the same iteration the same one
    dg.upload(g); dh.upload(h)
    cnt2 = L.goss_sample(dg, dh, top, other, seed=7, iteration=11)
    assert cnt2 == cnt and np.array_equal(L.get_bagging_data(cnt2), bag)
    dg.upload(g); dh.upload(h)
    cnt3 = L.goss_sample(dg, dh, top, other, seed=7, iteration=12)
    assert not np.array_equal(L.get_bagging_data(cnt3), bag) or cnt3 == n


def test_tree_on_goss_sample_matches_oracle(built_lib):
    import lightgbm_b200 as lgb
    from lightgbm_b200.tree_learner import DeviceArray
    from oracle import oracle_py
    n, f = 60000, 16
    bins, y, g, h = synth_identity(n, f, seed=31)
    h = (0.2 + np.random.default_rng(2).random(n)).astype(np.float32)
    lay = lgb.Layout.identity(bins)
    L = lgb.B200TreeLearner(lgb.Config(num_leaves=31))
    L.init(lay)
    dg, dh = DeviceArray(n * 4).upload(g), DeviceArray(n * 4).upload(h)
    cnt = L.goss_sample(dg, dh, 0.2, 0.1, seed=1, iteration=20)
    bag = L.get_bagging_data(cnt)
    t = L.train(dg, dh)
    o = oracle_py.train_tree(lay, dg.download(np.float32, n), dh.download(np.float32, n), bag_indices=bag, num_leaves=31)
    matched, diverged = compare_trees(t, o, 1e-5)
    assert diverged or matched == o.num_leaves - 1
    # graph replay with a bag of a different size (no re-capture needed: the count lives on the device)
    dg.upload(g); dh.upload(h)
    cnt2 = L.goss_sample(dg, dh, 0.1, 0.05, seed=1, iteration=21)
    bag2 = L.get_bagging_data(cnt2)
    assert cnt2 != cnt
    t2 = L.train(dg, dh)
    o2 = oracle_py.train_tree(lay, dg.download(np.float32, n), dh.download(np.float32, n), bag_indices=bag2, num_leaves=31)
    matched, diverged = compare_trees(t2, o2, 1e-5)
    assert diverged or matched == o2.num_leaves - 1


def test_goss_booster_learns(built_lib):
    """data_sample_strategy=goss through the booster mirror: the loss keeps falling once sampling starts."""
    import lightgbm_b200 as lgb
    n, f = 80000, 12
    bins, y, g, h = synth_identity(n, f, seed=9)
    B = lgb.B200Booster(lgb.Layout.identity(bins), y, lgb.Config(num_leaves=31), learning_rate=0.2, data_sample_strategy="goss",
                        top_rate=0.2, other_rate=0.1)
    losses = []
    for _ in range(12):
        B.update(); losses.append(B.l2())
    assert all(b < a for a, b in zip(losses, losses[1:])), losses
