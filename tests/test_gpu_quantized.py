"""GPU: quantized-gradient training (Config::use_quantized_grad) — the CUDA path against the oracle's restatement of
GradientDiscretizer + the integer split scan (oracle pinned to the reference by tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from helpers import compare_trees, synth_identity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built_lib):
    import lightgbm_b200 as lgb
    from oracle import oracle_py
    return lgb, oracle_py


def _orc_layout(lay):
    return lay        # oracle_py.make_layout reads the same attribute names


@pytest.mark.parametrize("bins_q,const_hess,leaves,n", [(4, False, 31, 40000), (16, False, 15, 25000), (4, True, 63, 60000),
                                                      (64, False, 31, 30000)])
def test_quantized_tree_matches_oracle(mods, bins_q, const_hess, leaves, n):
    lgb, orc = mods
    bins, y, g, h = synth_identity(n, 20, seed=5 + bins_q)
    if const_hess:
        h = np.ones_like(h)
    else:
        h = (0.2 + np.random.default_rng(1).random(n)).astype(np.float32)
    lay = lgb.Layout.identity(bins)
    cfg = lgb.Config(num_leaves=leaves, use_quantized_grad=True, num_grad_quant_bins=bins_q, stochastic_rounding=False)
    L = lgb.B200TreeLearner(cfg)
    L.init(lay, is_constant_hessian=const_hess)
    t = L.train(g, h)
    o = orc.train_tree_quant(_orc_layout(lay), g, h, num_grad_quant_bins=bins_q, is_constant_hessian=const_hess, num_leaves=leaves)
    assert t.grad_scale == o.grad_scale and t.hess_scale == o.hess_scale        # same fp64 expressions
    matched, diverged = compare_trees(t, o, 1e-11)
    assert not diverged and matched == o.num_leaves - 1 and t.num_leaves > 4
    # integer histograms: with splitting disabled the root's pool slot holds exactly the discretized sums
    R = lgb.B200TreeLearner(lgb.Config(num_leaves=2, min_gain_to_split=1e30, use_quantized_grad=True, num_grad_quant_bins=bins_q,
                                       stochastic_rounding=False))
    R.init(lay, is_constant_hessian=const_hess)
    assert R.train(g, h).num_leaves == 1
    qg, qh, gs, hs = orc.discretize(g, h, bins_q, const_hess)
    want = orc.construct_histogram(lay, None, qg.astype(np.float32), qh.astype(np.float32))
    np.testing.assert_array_equal(R.get_leaf_histogram(0), want)
    # a second tree from the same learner (new scales, graph replay) still matches
    g2 = (g * 0.5 + 0.01).astype(np.float32)
    t2 = L.train(g2, h)
    o2 = orc.train_tree_quant(_orc_layout(lay), g2, h, num_grad_quant_bins=bins_q, is_constant_hessian=const_hess, num_leaves=leaves)
    m2, d2 = compare_trees(t2, o2, 1e-11)
    assert not d2 and m2 == o2.num_leaves - 1


def test_quantized_bagging_renew_and_regularisation(mods):
    lgb, orc = mods
    n = 30000
    bins, y, g, h = synth_identity(n, 14, seed=77)
    h = (0.5 + np.random.default_rng(2).random(n)).astype(np.float32)
    lay = lgb.Layout.identity(bins)
    params = dict(num_leaves=24, lambda_l1=0.05, lambda_l2=1.5, max_delta_step=0.7, min_data_in_leaf=40, min_gain_to_split=0.01)
    cfg = lgb.Config(**params, use_quantized_grad=True, num_grad_quant_bins=8, stochastic_rounding=False, quant_train_renew_leaf=True)
    L = lgb.B200TreeLearner(cfg)
    L.init(lay)
    bag = np.sort(np.random.default_rng(3).choice(n, n // 2, replace=False)).astype(np.int32)
    L.set_bagging_data(bag)
    t = L.train(g, h)
    o = orc.train_tree_quant(_orc_layout(lay), g, h, num_grad_quant_bins=8, renew_leaf=True, bag_indices=bag, **params)
    matched, diverged = compare_trees(t, o, 1e-9)         # leaf values: renewed from fp32 gradients summed in another order
    assert not diverged and matched == o.num_leaves - 1
    np.testing.assert_allclose(t.leaf_value, o.leaf_value, rtol=1e-9, atol=1e-12)


def test_stochastic_rounding_is_seeded_and_unbiased(mods):
    lgb, orc = mods
    n = 200000
    bins, y, g, h = synth_identity(n, 8, seed=9)
    lay = lgb.Layout.identity(bins)

    def root_sums(seed, stochastic):
        L = lgb.B200TreeLearner(lgb.Config(num_leaves=7, use_quantized_grad=True, num_grad_quant_bins=4,
                                           stochastic_rounding=stochastic, seed=seed))
        L.init(lay, is_constant_hessian=True)
        t = L.train(g, np.ones_like(h))
        return t, t.root_sum_gradient

    t_a, s_a = root_sums(1, True)
    t_b, s_b = root_sums(1, True)
    t_c, s_c = root_sums(2, True)
    _, s_det = root_sums(1, False)
    assert s_a == s_b and np.array_equal(t_a.splits, t_b.splits)          # same seed -> same discretization
    assert s_a != s_c                                                      # another seed -> another stream
    true_sum = float(np.sum(g.astype(np.float64)))
    scale = t_a.grad_scale
    # E[int * scale] = g (unbiased): the error of the sum is ~ scale * sqrt(n) * 0.5; deterministic rounding is biased
    assert abs(s_a - true_sum) < 6 * scale * np.sqrt(n) * 0.5
    assert abs(s_c - true_sum) < 6 * scale * np.sqrt(n) * 0.5


def test_quantized_booster_learns(mods):
    lgb, _ = mods
    n = 50000
    bins, y, g, h = synth_identity(n, 16, seed=21)
    lay = lgb.Layout.identity(bins)
    cfg = lgb.Config(num_leaves=31, use_quantized_grad=True, num_grad_quant_bins=4, stochastic_rounding=True, seed=7,
                     quant_train_renew_leaf=True)
    a = lgb.B200Booster(lay, y, cfg, learning_rate=0.2, device_resident=True)
    b = lgb.B200Booster(lay, y, lgb.Config(num_leaves=31), learning_rate=0.2, device_resident=True)
    l0 = a.l2()
    for _ in range(10):
        a.update(); b.update()
    assert a.l2() < 0.5 * l0                     # quantized training converges ...
    assert a.l2() < 1.3 * b.l2()                 # ... close to full precision (the reference's own claim, config.h:626-631)
