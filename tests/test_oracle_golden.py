"""CPU: the oracle (oracle/lgbm_oracle.c) against the golden vectors produced by the real reference —
this is what pins the oracle.  Structure must be identical and leaf values BIT-identical (the reference
was run single-threaded/deterministic, the oracle restates its fp64 summation order)."""
import numpy as np
import pytest

import golden_io
from oracle import oracle_py


@pytest.mark.parametrize("name", golden_io.names())
def test_oracle_reproduces_reference_tree(name):
    g = golden_io.Golden(name)
    if g.quant is None:
        t = oracle_py.train_tree(g.layout, g.grad, g.hess, **g.params)
    else:
        # use_quantized_grad fixtures: integer histograms make every sum exact, so the bit-identity bar still holds
        # (leaf values after quant_train_renew_leaf are fp64 sums of the original gradients in row order)
        t = oracle_py.train_tree_quant(g.layout, g.grad, g.hess, **g.quant, **g.params)
    n = golden_io.check_against_reference(t, g, exact_values=True)
    assert n == int(g.ref["num_leaves"]) - 1
    if g.kat_y is not None and name in ("kat_missing_na", "kat_missing_zero", "kat_missing_handle", "kat_missing_more_na"):
        # the reference's known-answer assertion (test_engine.py:228,259): pred == y after one tree at lr 1
        pred = golden_io.row_predictions(t, t.leaf_begin, t.leaf_count, t.indices, g.layout.num_data)
        np.testing.assert_allclose(pred, g.kat_y, atol=1e-12)


def test_golden_set_covers_the_layout_contract():
    gs = [golden_io.Golden(n) for n in golden_io.names()]
    assert any((g.layout.feat_missing == 1).any() for g in gs)      # Zero-missing
    assert any((g.layout.feat_missing == 2).any() for g in gs)      # NaN-missing
    assert any((g.layout.feat_mfb > 0).any() for g in gs)           # FixHistogram path
    assert any((g.layout.feat_in_group > 1).any() for g in gs)      # EFB bundles
    assert any(g.params["lambda_l1"] > 0 for g in gs) and any(g.params["path_smooth"] > 0 for g in gs)


def test_oracle_histogram_and_partition_primitives():
    g = golden_io.Golden("efb_bundled")
    lay = g.layout
    rng = np.random.default_rng(0)
    idx = np.sort(rng.choice(lay.num_data, 1000, replace=False)).astype(np.int32)
    h = oracle_py.construct_histogram(lay, idx, g.grad, g.hess)
    # every row lands in exactly one slot per column
    np.testing.assert_allclose(h[:, :, 1].sum(axis=1), g.hess[idx].astype(np.float64).sum(), rtol=1e-12)
    f = 5
    out, nl = oracle_py.partition(lay, f, 3, 1, idx)
    assert sorted(out.tolist()) == idx.tolist() and 0 <= nl <= len(idx)
    assert np.all(np.diff(out[:nl]) > 0) and np.all(np.diff(out[nl:]) > 0)     # stable
