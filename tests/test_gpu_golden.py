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
    L = lgb.B200TreeLearner(lgb.Config(**g.params))
    L.init(lay)
    t = L.train(g.grad, g.hess)
    o = oracle_py.train_tree(g.layout, g.grad, g.hess, **g.params)
    matched, diverged = compare_trees(t, o, 1e-5)
    n_ref = golden_io.check_against_reference(t, g, exact_values=False)
    assert matched >= min(3, o.num_leaves - 1) and n_ref >= min(3, o.num_leaves - 1)
    if g.kat_y is not None and name != "kat_missing_none":
        lb, lc, idx = L.get_partition(t.num_leaves)
        pred = golden_io.row_predictions(t, lb, lc, idx, lay.num_data)
        np.testing.assert_allclose(pred, g.kat_y, atol=1e-6)      # the reference's known-answer assertion
    if not diverged:
        lb, lc, idx = L.get_partition(t.num_leaves)
        for leaf in range(t.num_leaves):
            np.testing.assert_array_equal(idx[lb[leaf]:lb[leaf] + lc[leaf]],
                                          o.indices[o.leaf_begin[leaf]:o.leaf_begin[leaf] + o.leaf_count[leaf]])
