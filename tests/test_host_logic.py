"""CPU: host-side logic of the mirror classes (no GPU, no compute calls)."""
import numpy as np

import lightgbm_b200 as lgb
from lightgbm_b200.tree_learner import SPLIT_DTYPE, Tree
from oracle import refapi


def test_identity_layout_contract():
    bins = np.arange(12, dtype=np.uint8).reshape(3, 4)
    lay = lgb.Layout.identity(bins)
    assert lay.num_data == 3 and lay.num_columns == 4 and lay.num_features == 4
    assert (lay.feat_lo == 1).all() and (lay.feat_mfb == 0).all() and (lay.feat_num_bin == 255).all()


def test_column_slice_is_a_feature_shard():
    rng = np.random.default_rng(0)
    bins = rng.integers(0, 255, (10, 8), dtype=np.uint8)
    lay = lgb.Layout.identity(bins)
    s = lay.column_slice(4, 8)
    assert s.num_columns == 4 and s.num_features == 4
    assert np.array_equal(s.bins, bins[:, 4:8])
    assert np.array_equal(s.feat_real_index, np.arange(4, 8))      # global ids survive for the tie-break
    assert np.array_equal(s.feat_column, np.arange(4))


def test_tree_shrinkage_and_bias():
    t = Tree(2, np.zeros(1, SPLIT_DTYPE), np.array([1.0, -2.0]), np.ones(2), np.array([3, 4], np.int32),
             np.array([1, 1], np.int32), 0.0, 7.0)
    t.shrinkage(0.1)
    np.testing.assert_allclose(t.leaf_value, [0.1, -0.2])
    t.add_bias(1.0)
    np.testing.assert_allclose(t.leaf_value, [1.1, 0.8])


def test_config_maps_to_c_struct():
    c = lgb.Config(num_leaves=127, lambda_l2=1.5, use_cuda_graph=False).to_c()
    assert c.num_leaves == 127 and c.lambda_l2 == 1.5 and c.use_cuda_graph == 0


def test_model_text_parser_split_leaf_recovery():
    txt = ("tree\nTree=0\nnum_leaves=3\nnum_cat=0\nsplit_feature=1 0\nsplit_gain=10 5\nthreshold=0.5 1.5\n"
           "decision_type=2 0\nleft_child=1 -1\nright_child=-2 -3\nleaf_value=0.1 0.2 0.3\nleaf_weight=1 2 3\n"
           "leaf_count=1 2 3\ninternal_value=0 0\ninternal_weight=6 4\ninternal_count=6 4\nis_linear=0\nshrinkage=1\n\n"
           "end of trees\n")
    t = refapi.parse_model_trees(txt)[0]
    assert t.num_leaves == 3
    assert t.split_leaf().tolist() == [0, 0]          # node 1 is the left child of node 0 => it split leaf 0
    assert t.default_left.tolist() == [1, 0]


def test_packed_cell_arithmetic_of_the_quantized_histogram():
    """k_hist_q keeps (sum g << 16) + sum h in one int32 and adds one precomputed word per row.  Property the kernel
    relies on: for |g| <= Q/2, 0 <= h <= Q and at most floor(65535/Q) rows, H = P & 0xffff and G = (P - H) >> 16 recover
    the two sums exactly (worst cases included), and one more row of the maximum hessian may break it."""
    rng = np.random.default_rng(0)
    for Q in (2, 4, 5, 8, 15):
        R = (65535 // Q) // 32 * 32
        for trial in range(4):
            if trial == 0:
                g = np.full(R, Q // 2, np.int64); h = np.full(R, Q, np.int64)          # both fields at their maximum
            elif trial == 1:
                g = np.full(R, -(Q // 2), np.int64); h = np.zeros(R, np.int64)          # most negative gradient field
            else:
                g = rng.integers(-(Q // 2), Q // 2 + 1, R); h = rng.integers(0, Q + 1, R)
            words = (g * 65536 + h).astype(np.int64)
            P = int(np.sum(words))
            assert -2 ** 31 <= P < 2 ** 31                                              # fits the int32 cell
            H = P & 0xffff
            G = (P - H) >> 16
            assert H == int(h.sum()) and G == int(g.sum())
        # the bound is tight for the hessian field: R' = floor(65535/Q) + 1 rows of h = Q overflow 16 bits
        assert (65535 // Q + 1) * Q > 65535
