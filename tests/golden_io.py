"""Loads the golden fixtures of tests/golden/*.npz (written by tests/golden/make_golden.py from the real reference)."""
import glob
import os
import sys

import numpy as np

from oracle.refapi import Layout

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEARNER_KEYS = ("num_leaves", "max_depth", "min_data_in_leaf", "min_sum_hessian_in_leaf", "lambda_l1", "lambda_l2",
                "min_gain_to_split", "max_delta_step", "path_smooth")
INT_KEYS = ("num_leaves", "max_depth", "min_data_in_leaf")


def names():
    # binning_*.npz / model_*.npz are the fixtures of tests/test_binning.py / tests/test_model.py (different contents)
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) if not n.startswith(("binning_", "model_")))


class Golden:
    def __init__(self, name):
        d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
        self.name = name
        if "gen_efb4" in d:
            # inputs regenerated from their seeds (tests/golden/make_golden.py efb4_inputs) and checked against the
            # checksums taken when the reference produced the fixture
            sys.path.insert(0, os.path.join(GOLDEN_DIR))
            sys.path.insert(0, os.path.dirname(os.path.dirname(GOLDEN_DIR)))
            import make_golden
            assert [int(v) for v in d["gen_efb4"]] == [make_golden.EFB4_GEN[k] for k in ("rows", "cols", "seed", "grad_seed")]
            _, raw, _, g, h = make_golden.efb4_inputs()
            cols = make_golden.bundle_by_layout(raw, d["feat_column"], d["feat_lo"], d["feat_real_index"], int(d["dims"][1]))
            w = np.arange(cols.shape[1], dtype=np.uint64) + 1
            assert int(cols.astype(np.uint64).sum()) == int(d["bins_checksum"][0]) and int((cols.astype(np.uint64) * w).sum()) == int(d["bins_checksum"][1])
            assert float(g.astype(np.float64).sum()) == float(d["grad_checksum"][0]) and float(h.astype(np.float64).sum()) == float(d["grad_checksum"][1])
            d["bins"], d["grad"], d["hess"] = cols, g, h
        self.layout = Layout.from_npz_dict(d)
        self.grad, self.hess = d["grad"], d["hess"]
        self.params = {k: (int(v) if k in INT_KEYS else float(v)) for k, v in zip(LEARNER_KEYS, d["params"])}
        self.ref = {k[4:]: d[k] for k in d if k.startswith("ref_")}
        self.kat_y = d.get("kat_y")
        q = d.get("quant")        # use_quantized_grad fixtures: [num_grad_quant_bins, renew_leaf, is_constant_hessian]
        self.quant = None if q is None else dict(num_grad_quant_bins=int(q[0]), renew_leaf=bool(q[1]),
                                                 is_constant_hessian=bool(q[2]))


def check_against_reference(tree, g: Golden, exact_values: bool, rtol: float = 1e-5):
    """tree: anything with .num_leaves/.splits/.leaf_value/.leaf_count (oracle or CUDA path).
    Returns number of leading splits identical to the reference's."""
    ref = g.ref
    n_ref = int(ref["num_leaves"])
    ns = min(tree.num_leaves, n_ref) - 1
    for i in range(ns):
        s = tree.splits[i]
        same = (s["leaf"] == ref["split_leaf"][i] and s["feature"] == ref["split_feature_inner"][i]
                and s["threshold"] == ref["threshold_bin"][i] and s["default_left"] == ref["default_left"][i])
        if not same:
            if exact_values:
                raise AssertionError(f"{g.name}: split {i} differs from the reference: got {s}, want leaf="
                                     f"{ref['split_leaf'][i]} f={ref['split_feature_inner'][i]} t={ref['threshold_bin'][i]}")
            rel = abs(s["gain"] + g.params["min_gain_to_split"] - ref["split_gain"][i]) / max(abs(ref["split_gain"][i]), 1e-300)
            assert rel < 1e-4, f"{g.name}: split {i} mismatch is not a near tie ({rel})"
            return i
        assert s["left_count"] + s["right_count"] == ref["internal_count"][i]
        # the model text stores the gain as float32 printed with ~6 significant digits
        # Tree::Split stores float(gain + min_gain_to_split) (serial_tree_learner.cpp:811), printed with ~6 digits
        stored = s["gain"] + g.params["min_gain_to_split"]
        assert abs(stored - ref["split_gain"][i]) <= 2e-5 * abs(ref["split_gain"][i]) + 1e-7
    assert tree.num_leaves == n_ref
    np.testing.assert_array_equal(tree.leaf_count, ref["leaf_count"])
    if exact_values:
        np.testing.assert_array_equal(tree.leaf_value, ref["leaf_value"])
        np.testing.assert_array_equal(tree.leaf_weight, ref["leaf_weight"])
    else:
        np.testing.assert_allclose(tree.leaf_value, ref["leaf_value"], rtol=rtol * 10, atol=1e-9)
    return ns


def row_predictions(tree, leaf_begin, leaf_count, indices, n):
    out = np.zeros(n)
    for leaf in range(tree.num_leaves):
        out[indices[leaf_begin[leaf]:leaf_begin[leaf] + leaf_count[leaf]]] = tree.leaf_value[leaf]
    return out
