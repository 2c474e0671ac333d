"""GPU: the CUDA path on the golden fixtures (missing values, most-frequent-bin elision, EFB bundles,
regularisation) against the reference's own trees and against the oracle."""
import numpy as np
import pytest

import golden_io
from helpers import compare_trees

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_io.names())
def test_cuda_path_on_golden_fixture(built_lib, name):
    import lightgbm_b200 as lgb
    from oracle import oracle_py
    g = golden_io.Golden(name)
    lay = lgb.Layout.from_attrs(g.layout)
    if g.quant is None:
        L = lgb.B200TreeLearner(lgb.Config(**g.params))
        L.init(lay)
        t = L.train(g.grad, g.hess)
        o = oracle_py.train_tree(g.layout, g.grad, g.hess, **g.params)
        matched, diverged = compare_trees(t, o, 1e-5)
    else:
        # use_quantized_grad fixtures (stochastic_rounding=false): integer histograms are exact, so the split sequence
        # must be IDENTICAL to the reference's and the values equal up to fp64 contraction (1e-12)
        q = g.quant
        L = lgb.B200TreeLearner(lgb.Config(**g.params, use_quantized_grad=True, num_grad_quant_bins=q["num_grad_quant_bins"],
                                           quant_train_renew_leaf=q["renew_leaf"], stochastic_rounding=False))
        L.init(lay, is_constant_hessian=q["is_constant_hessian"])
        t = L.train(g.grad, g.hess)
        o = oracle_py.train_tree_quant(g.layout, g.grad, g.hess, **q, **g.params)
        assert t.grad_scale == o.grad_scale and t.hess_scale == o.hess_scale
        matched, diverged = compare_trees(t, o, 1e-11)
        assert not diverged and matched == o.num_leaves - 1
        np.testing.assert_allclose(t.leaf_value, o.leaf_value, rtol=1e-11 if not q["renew_leaf"] else 1e-9, atol=1e-15)
        assert golden_io.check_against_reference(t, g, exact_values=False) == o.num_leaves - 1
    n_ref = golden_io.check_against_reference(t, g, exact_values=False)
    # the oracle reproduces the reference bit for bit, so a full match against it is a full match against the reference
    assert diverged or (matched == o.num_leaves - 1 and n_ref == o.num_leaves - 1)
    if g.kat_y is not None and name != "kat_missing_none":
        lb, lc, idx = L.get_partition(t.num_leaves)
        pred = golden_io.row_predictions(t, lb, lc, idx, lay.num_data)
        np.testing.assert_allclose(pred, g.kat_y, atol=1e-6)      # the reference's known-answer assertion
    if not diverged:
        lb, lc, idx = L.get_partition(t.num_leaves)
        for leaf in range(t.num_leaves):
            np.testing.assert_array_equal(idx[lb[leaf]:lb[leaf] + lc[leaf]],
                                          o.indices[o.leaf_begin[leaf]:o.leaf_begin[leaf] + o.leaf_count[leaf]])
