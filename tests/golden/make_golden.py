"""Generates the golden fixtures under tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/lib_lightgbm.so built by oracle/Makefile.ref from /root/reference) in this container.

Each fixture holds: the binned matrix + layout metadata read back from the reference Dataset (probe),
the fp32 gradients/hessians fed through LGBM_BoosterUpdateOneIterCustom, the learner parameters, and the
tree the reference CPU learner (serial, col-wise, num_threads=1, deterministic) grew from them, parsed from
LGBM_BoosterSaveModelToString.  The missing-value tables are the reference's own known-answer tests
(tests/python_package_test/test_engine.py:161-292).

Usage:  python tests/golden/make_golden.py      (needs oracle/_ref; the GPU box only reads the .npz)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refapi  # noqa: E402

BASE_DS = dict(verbosity=-1, num_threads=1, min_data_in_bin=1, feature_pre_filter="false", max_bin=255)
BASE_BOOST = dict(objective="custom", learning_rate=1.0, force_col_wise="true", deterministic="true", num_threads=1,
                  verbosity=-1)
LEARNER_KEYS = ("num_leaves", "max_depth", "min_data_in_leaf", "min_sum_hessian_in_leaf", "lambda_l1", "lambda_l2",
                "min_gain_to_split", "max_delta_step", "path_smooth")
DEFAULTS = dict(num_leaves=31, max_depth=-1, min_data_in_leaf=20, min_sum_hessian_in_leaf=1e-3, lambda_l1=0.0,
                lambda_l2=0.0, min_gain_to_split=0.0, max_delta_step=0.0, path_smooth=0.0)


def thresholds_to_bins(lay, tree):
    """real-valued thresholds of the model -> threshold_in_bin via the BinMapper upper bounds."""
    inner_of_real = {int(r): i for i, r in enumerate(lay.feat_real_index)}
    feats, bins = [], []
    for rf, th in zip(tree.split_feature, tree.threshold):
        f = inner_of_real[int(rf)]
        ub = lay.bin_upper_bound[f]
        ubc = np.clip(np.nan_to_num(ub, nan=1e300), -1e300, 1e300)      # Common::AvoidInf (model text writes +-1e300)
        thc = float(np.clip(th, -1e300, 1e300))
        b = int(np.argmin(np.abs(ubc - thc)))
        assert abs(ubc[b] - thc) <= 1e-12 * max(1.0, abs(thc)), (ub[b], th)
        feats.append(f); bins.append(b)
    return np.array(feats, np.int32), np.array(bins, np.int32)


def run_case(name, X, grad, hess, ds_params=None, learner=None, y=None, quant=None, label=None):
    """quant: dict(num_grad_quant_bins=, quant_train_renew_leaf=) -> the reference trains with use_quantized_grad=true,
    stochastic_rounding=false (the stochastic branch draws from per-thread mt19937 streams: not a function of the inputs
    alone).  label: train through objective=regression (boost_from_average=false, so g = -label, h = 1 and the learner
    is Init-ed with is_constant_hessian=true) instead of custom gradients."""
    ds_params = dict(BASE_DS, **(ds_params or {}))
    lp = dict(DEFAULTS, **(learner or {}))
    lab = np.zeros(len(X), np.float32) if label is None else np.asarray(label, np.float32)
    ds = refapi.RefDataset(np.asarray(X, dtype=np.float64), lab, ds_params)
    lay = ds.layout()
    bp = dict(ds_params, **BASE_BOOST, **lp)
    bp["device_type"] = "cpu"     # Dataset may be *constructed* with cuda rules (dense bundles); training is CPU
    if quant is not None:
        bp.update(use_quantized_grad="true", stochastic_rounding="false", num_grad_quant_bins=quant["num_grad_quant_bins"],
                  quant_train_renew_leaf="true" if quant.get("quant_train_renew_leaf") else "false")
    if label is not None:
        bp.update(objective="regression", boost_from_average="false")
    bst = refapi.RefBooster(ds, bp)
    g = np.ascontiguousarray(grad, np.float32); h = np.ascontiguousarray(hess, np.float32)
    if label is None:
        bst.update_custom(g, h)
    else:
        bst.update()
    t = bst.trees()[0]
    feats, tbins = (thresholds_to_bins(lay, t) if t.num_leaves > 1 else (np.zeros(0, np.int32), np.zeros(0, np.int32)))
    d = lay.to_npz_dict()
    d.update(grad=g, hess=h, params=np.array([lp[k] for k in LEARNER_KEYS], np.float64),
             ref_num_leaves=np.int64(t.num_leaves), ref_split_feature_inner=feats, ref_threshold_bin=tbins,
             ref_split_leaf=t.split_leaf() if t.num_leaves > 1 else np.zeros(0, np.int32),
             ref_default_left=t.default_left.astype(np.int32), ref_split_gain=t.split_gain,
             ref_internal_count=t.internal_count, ref_leaf_value=t.leaf_value, ref_leaf_count=t.leaf_count,
             ref_leaf_weight=t.leaf_weight, ref_threshold_real=t.threshold)
    if y is not None:
        d["kat_y"] = np.asarray(y, np.float64)
    if quant is not None:
        d["quant"] = np.array([quant["num_grad_quant_bins"], 1 if quant.get("quant_train_renew_leaf") else 0,
                               1 if label is not None else 0], np.int32)     # bins, renew_leaf, is_constant_hessian
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print(f"{name}: N={lay.num_data} C={lay.num_columns} F={lay.num_features} leaves={t.num_leaves} "
          f"missing={sorted(set(lay.feat_missing.tolist()))} mfb>0={int((lay.feat_mfb > 0).sum())} "
          f"multi_feature_cols={int((lay.feat_in_group > 1).sum())} -> {os.path.getsize(path)} B")
    bst.free(); ds.free()


def quantized_cases():
    """use_quantized_grad fixtures (own RNG stream, so the full-precision fixtures above stay byte-identical)."""
    rng = np.random.default_rng(4242)
    nan = np.nan
    n, f = 6000, 12
    bins = rng.integers(0, 255, (n, f))
    yv = (bins[:, :6] / 127.0 - 1) @ rng.normal(size=6) + 0.3 * rng.normal(size=n)
    # constant hessian through the L2 objective (g = -label, h = 1): int hessian == 1, hess_scale = 1
    run_case("quant_l2_consthess", bins, -yv, np.ones(n), dict(enable_bundle="false"), dict(num_leaves=31),
             quant=dict(num_grad_quant_bins=4), label=yv)
    # custom gradients: hessians are discretized too
    p = 1 / (1 + np.exp(-0.5 * yv)); yb = (rng.random(n) < p).astype(float); q = np.full(n, 0.5)
    run_case("quant_logistic_bins16", bins, q - yb + 0.1 * rng.normal(size=n), q * (1 - q) + 0.2 * rng.random(n),
             dict(enable_bundle="false"), dict(num_leaves=31, lambda_l2=1.0), quant=dict(num_grad_quant_bins=16))
    run_case("quant_renew_leaf", bins, -yv, 0.5 + rng.random(n), dict(enable_bundle="false"),
             dict(num_leaves=24, min_data_in_leaf=30, lambda_l1=0.1, max_delta_step=0.6),
             quant=dict(num_grad_quant_bins=6, quant_train_renew_leaf=True))
    # reference-binned features with NaN / zero-as-missing / most-frequent-bin elision, quantized
    n, f = 5000, 10
    X = rng.normal(size=(n, f))
    X[rng.random((n, f)) < 0.08] = nan
    X[:, 3] = np.where(rng.random(n) < 0.85, 0.0, X[:, 3])
    X[:, 4] = np.where(rng.random(n) < 0.8, 2.5, rng.normal(size=n))
    X[:, 6] = (rng.random(n) < 0.5).astype(float)
    logit = np.nan_to_num(X[:, 0]) - 0.7 * np.nan_to_num(X[:, 1]) + (X[:, 4] == 2.5) * 0.8 + np.isnan(X[:, 2]) * 1.0
    yb = (rng.random(n) < 1 / (1 + np.exp(-logit))).astype(float)
    pp = np.full(n, 0.5)
    run_case("quant_mixed_missing", X, pp - yb, pp * (1 - pp) + 0.05 * rng.random(n), dict(max_bin=63, device_type="cuda"),
             dict(num_leaves=31, min_data_in_leaf=10), quant=dict(num_grad_quant_bins=8))
    run_case("quant_mixed_zero_as_missing", X, pp - yb, pp * (1 - pp) + 0.05 * rng.random(n),
             dict(max_bin=63, zero_as_missing="true", device_type="cuda"), dict(num_leaves=31, min_data_in_leaf=10),
             quant=dict(num_grad_quant_bins=4))


def example_cases():
    """The reference's own example datasets (examples/regression, examples/binary_classification: 7000 x 28 real-valued
    features, the data behind tests/python_package_test/test_consistency.py) with the learner settings of their
    train.conf (num_leaves=31, min_data_in_leaf=100/50, min_sum_hessian_in_leaf=5.0, max_bin=255): real-data bins, ties
    and a most-frequent-bin pattern no synthetic generator produces.  Only the binned matrix travels in the fixture."""
    ex = "/root/reference/examples"
    reg = np.loadtxt(os.path.join(ex, "regression", "regression.train"))
    y, X = reg[:, 0], reg[:, 1:]
    run_case("example_regression", X, y.mean() - y, np.ones(len(y)), dict(max_bin=255),
             dict(num_leaves=31, min_data_in_leaf=100, min_sum_hessian_in_leaf=5.0))
    run_case("example_regression_quant", X, y.mean() - y, np.ones(len(y)), dict(max_bin=255),
             dict(num_leaves=31, min_data_in_leaf=100, min_sum_hessian_in_leaf=5.0), quant=dict(num_grad_quant_bins=4))
    bi = np.loadtxt(os.path.join(ex, "binary_classification", "binary.train"))
    yb, Xb = bi[:, 0], bi[:, 1:]
    p = np.full(len(yb), yb.mean())
    run_case("example_binary", Xb, p - yb, p * (1 - p), dict(max_bin=255),
             dict(num_leaves=63, min_data_in_leaf=50, min_sum_hessian_in_leaf=5.0))
    # second-iteration-like gradients (hessians vary per row) with quantization and leaf renewal
    rng = np.random.default_rng(7)
    s = 0.4 * (Xb[:, 0] - Xb[:, 0].mean()) + 0.1 * rng.normal(size=len(yb))
    p2 = 1 / (1 + np.exp(-s))
    run_case("example_binary_quant", Xb, p2 - yb, p2 * (1 - p2), dict(max_bin=255),
             dict(num_leaves=31, min_data_in_leaf=50, min_sum_hessian_in_leaf=5.0, lambda_l2=0.5),
             quant=dict(num_grad_quant_bins=8, quant_train_renew_leaf=True))


EFB4_GEN = dict(rows=200_000, cols=256, seed=45, grad_seed=9045)


def efb4_inputs():
    """The C4-shaped inputs of the efb4 fixture, a pure function of EFB4_GEN (bench.py generators): bundled columns,
    raw features, binary labels, logistic gradients at a seeded random score."""
    import bench
    wl = dict(bench.WORKLOADS["C4"], rows=EFB4_GEN["rows"], cols=EFB4_GEN["cols"], seed=EFB4_GEN["seed"])
    raw = bench.gen_efb4(wl["rows"], wl["cols"], wl["seed"], raw=True)
    y = bench.gen_label_wl(wl, bench.gen_columns(wl, 0, bench.LABEL_COLS["efb4"]))
    score = np.random.default_rng(EFB4_GEN["grad_seed"]).normal(size=wl["rows"]) * 0.8
    p = 1.0 / (1.0 + np.exp(-score))
    return wl, raw, y, (p - y).astype(np.float32), (p * (1.0 - p)).astype(np.float32)


def bundle_by_layout(raw, feat_column, feat_lo, feat_real_index, num_columns):
    """Stored group values (feature_group.h:253-267) of sparse raw features whose value v > 0 has bin v and whose most
    frequent bin 0 is elided, for the bundling (column, offset per feature) a reference Dataset chose."""
    out = np.zeros((len(raw), num_columns), dtype=np.uint8)
    for f in range(len(feat_column)):
        v = raw[:, feat_real_index[f]]
        nz = np.nonzero(v)[0]
        out[nz, feat_column[f]] = feat_lo[f] + v[nz] - 1
    return out


def efb4_case():
    """C4 at fixture scale: 200 000 x 256 sparse features, exclusive in blocks of 4, through the
    reference's OWN Dataset construction with the cuda rules (sampled-column API + PushRows: EFB bundles them into 64
    uint8 columns), binary-logloss gradients, the reference CPU learner's 63-leaf tree.  The 12.8 MB bin matrix is not
    stored: the fixture keeps the layout the reference produced, checksums of the matrix / gradients, and the tree; the
    loader regenerates the inputs from EFB4_GEN and checks them against the checksums."""
    import bench
    wl, raw, y, g, h = efb4_inputs()
    dsp = dict(BASE_DS, enable_bundle="true", device_type="cuda", num_threads=8)
    ds = refapi.RefDatasetStreamed(lambda lo, hi: bench.gen_raw_float(wl, row_lo=lo, row_hi=hi), wl["rows"], wl["cols"], y, dsp,
                                   block_rows=4 * bench.GEN_CHUNK, sample_rows=wl["rows"], sampled_columns=True)
    lay = ds.layout()
    cols = bundle_by_layout(raw, lay.feat_column, lay.feat_lo, lay.feat_real_index, lay.num_columns)
    assert lay.num_columns == wl["cols"] // 4 and np.array_equal(lay.bins, cols), "the reference stored the bundles differently"
    lp = dict(DEFAULTS, num_leaves=63, min_data_in_leaf=20)
    bp = dict(dsp, **BASE_BOOST, **lp); bp["device_type"] = "cpu"; bp["num_threads"] = 1
    bst = refapi.RefBooster(ds, bp)
    bst.update_custom(g, h)
    t = bst.trees()[0]
    feats, tbins = thresholds_to_bins(lay, t)
    d = lay.to_npz_dict(with_bins=False)
    d.update(gen_efb4=np.array([EFB4_GEN[k] for k in ("rows", "cols", "seed", "grad_seed")], np.int64),
             bins_checksum=np.array([int(cols.astype(np.uint64).sum()), int((cols.astype(np.uint64) * (np.arange(cols.shape[1], dtype=np.uint64) + 1)).sum())], np.uint64),
             grad_checksum=np.array([float(g.astype(np.float64).sum()), float(h.astype(np.float64).sum())]),
             params=np.array([lp[k] for k in LEARNER_KEYS], np.float64),
             ref_num_leaves=np.int64(t.num_leaves), ref_split_feature_inner=feats, ref_threshold_bin=tbins,
             ref_split_leaf=t.split_leaf(), ref_default_left=t.default_left.astype(np.int32), ref_split_gain=t.split_gain,
             ref_internal_count=t.internal_count, ref_leaf_value=t.leaf_value, ref_leaf_count=t.leaf_count,
             ref_leaf_weight=t.leaf_weight, ref_threshold_real=t.threshold)
    path = os.path.join(HERE, "efb4_200k_x256.npz")
    np.savez_compressed(path, **d)
    print(f"efb4_200k_x256: N={lay.num_data} C={lay.num_columns} F={lay.num_features} leaves={t.num_leaves} "
          f"multi_feature_cols={int((lay.feat_in_group > 1).sum())} -> {os.path.getsize(path)} B")
    bst.free(); ds.free()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "quant":
        quantized_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "efb4":
        efb4_case()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "examples":
        example_cases()
        return
    rng = np.random.default_rng(2024)
    nan = np.nan
    # --- the reference's own known-answer tables (test_engine.py:201-292): 1 tree, lr 1, pred == y
    kat = dict(num_leaves=2, min_data_in_leaf=1, min_sum_hessian_in_leaf=1e-3)
    x9 = np.array([0, 1, 2, 3, 4, 5, 6, 7, nan]).reshape(-1, 1)
    y_na = np.array([1, 1, 1, 1, 0, 0, 0, 0, 1.0])
    y_zero = np.array([0, 1, 1, 1, 0, 0, 0, 0, 0.0])
    run_case("kat_missing_na", x9, -y_na, np.ones(9), dict(zero_as_missing="false"), kat, y=y_na)
    run_case("kat_missing_zero", x9, -y_zero, np.ones(9), dict(zero_as_missing="true"), kat, y=y_zero)
    run_case("kat_missing_none", x9, -y_zero, np.ones(9), dict(use_missing="false"), kat, y=y_zero)
    # test_missing_value_handle (:161) / _more_na (:180): 100 rows, one feature, NaN marks the positives
    x = np.zeros((100, 1)); y = np.zeros(100); idx = rng.choice(100, 20, replace=False); x[idx, 0] = nan; y[idx] = 1
    run_case("kat_missing_handle", x, -y, np.ones(100), None, dict(num_leaves=31), y=y)
    x = np.ones((100, 1)); y = np.ones(100); idx = rng.choice(100, 80, replace=False); x[idx, 0] = nan; y[idx] = 0
    run_case("kat_missing_more_na", x, -y, np.ones(100), None, dict(num_leaves=31), y=y)

    # --- integer-valued features: bin == value
    n, f = 4000, 12
    bins = rng.integers(0, 255, (n, f))
    yv = (bins[:, :6] / 127.0 - 1) @ rng.normal(size=6) + 0.3 * rng.normal(size=n)
    run_case("identity_l2", bins, -yv, np.ones(n), dict(enable_bundle="false"), dict(num_leaves=31))
    run_case("identity_l2_reg", bins, -yv, np.ones(n), dict(enable_bundle="false"),
             dict(num_leaves=24, lambda_l1=0.3, lambda_l2=2.0, min_gain_to_split=1.0, max_depth=6))
    run_case("identity_l2_smooth", bins, -yv, np.ones(n), dict(enable_bundle="false"),
             dict(num_leaves=16, path_smooth=5.0, max_delta_step=0.4, min_data_in_leaf=50))

    # --- continuous features binned by the reference, NaN + zeros + a "mostly one value" feature
    n, f = 5000, 10
    X = rng.normal(size=(n, f))
    X[rng.random((n, f)) < 0.08] = nan                     # NaN-missing features
    X[:, 3] = np.where(rng.random(n) < 0.85, 0.0, X[:, 3])   # sparse: default bin is the most frequent
    X[:, 4] = np.where(rng.random(n) < 0.8, 2.5, rng.normal(size=n))   # most frequent bin != zero bin
    X[:, 5] = np.round(X[:, 5])                            # few distinct values
    X[:, 6] = (rng.random(n) < 0.5).astype(float)          # 2 bins
    logit = np.nan_to_num(X[:, 0]) - 0.7 * np.nan_to_num(X[:, 1]) + (X[:, 4] == 2.5) * 0.8 + np.isnan(X[:, 2]) * 1.0
    yb = (rng.random(n) < 1 / (1 + np.exp(-logit))).astype(float)
    p = np.full(n, 0.5)
    run_case("mixed_missing_binary", X, p - yb, p * (1 - p) + 0.05 * rng.random(n), dict(max_bin=63, device_type="cuda"), dict(num_leaves=31, min_data_in_leaf=10))
    run_case("mixed_zero_as_missing", X, p - yb, p * (1 - p) + 0.05 * rng.random(n),
             dict(max_bin=63, zero_as_missing="true", device_type="cuda"), dict(num_leaves=31, min_data_in_leaf=10))

    # --- EFB: mutually exclusive sparse features bundled into shared columns (dense storage as on cuda)
    n, f = 6000, 24
    X = np.zeros((n, f))
    for blk in range(0, f, 4):
        who = rng.integers(0, 6, n)                         # which of the 4 features (or none) is non-zero
        for j in range(4):
            m = who == j
            X[m, blk + j] = rng.integers(1, 40, m.sum())
    yv = X[:, 0] * 0.05 - X[:, 5] * 0.03 + (X[:, 9] > 20) * 1.0 + 0.2 * rng.normal(size=n)
    run_case("efb_bundled", X, -yv, np.ones(n), dict(enable_bundle="true", device_type="cuda", max_bin=63),
             dict(num_leaves=31, min_data_in_leaf=10))
    quantized_cases()
    example_cases()


if __name__ == "__main__":
    main()
