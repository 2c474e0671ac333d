"""Generates tests/golden/model_*.npz: model texts trained by the UNMODIFIED reference (CPU learner) on small float matrices,
a test matrix, and the reference's own predictions for it (LGBM_BoosterPredictForMat, raw and transformed).
Fixtures of tests/test_model.py.  Run in the build container: python tests/golden/make_model_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import refapi  # noqa: E402


def data(seed, n, f, nan_rate=0.0, zero_rate=0.0):
    r = np.random.default_rng(seed)
    X = r.normal(size=(n, f)).astype(np.float32)
    if zero_rate:
        X[r.random((n, f)) < zero_rate] = 0.0
    if nan_rate:
        X[r.random((n, f)) < nan_rate] = np.nan
    w = r.normal(size=f)
    s = np.nan_to_num(X) @ w + np.sin(np.nan_to_num(X[:, 0]) * 3)
    return X, s, r


CASES = {
    "regression_dense": dict(params=dict(objective="regression", num_leaves=31, learning_rate=0.1), iters=20, nan=0.0, zero=0.0, f=12),
    "regression_missing_nan": dict(params=dict(objective="regression", num_leaves=15, learning_rate=0.2), iters=12, nan=0.15, zero=0.0, f=8),
    "regression_zero_as_missing": dict(params=dict(objective="regression", num_leaves=15, zero_as_missing="true"), iters=10, nan=0.05, zero=0.3, f=8),
    "binary_sigmoid": dict(params=dict(objective="binary", num_leaves=31, sigmoid=1.5, learning_rate=0.1), iters=15, nan=0.02, zero=0.1, f=10),
    "regression_no_average_deep": dict(params=dict(objective="regression", num_leaves=127, boost_from_average="false", min_data_in_leaf=5), iters=6, nan=0.0, zero=0.0, f=6),
}

for i, (name, c) in enumerate(CASES.items()):
    X, s, r = data(100 + i, 6000, c["f"], c["nan"], c["zero"])
    y = (s + 0.3 * r.normal(size=len(s))).astype(np.float32)
    if c["params"]["objective"] == "binary":
        y = (s > np.median(s)).astype(np.float32)
    params = dict(c["params"], verbosity=-1, num_threads=4, deterministic="true", force_col_wise="true", min_data_in_bin=3)
    ds = refapi.RefDataset(X, y, params)
    bst = refapi.RefBooster(ds, params)
    for _ in range(c["iters"]):
        bst.update()
    text = bst.model_string()
    Xt, _, _ = data(900 + i, 3000, c["f"], max(c["nan"], 0.05), max(c["zero"], 0.05))
    Xt[:5] = 0.0; Xt[5:8] = np.nan; Xt[8] = 1e30; Xt[9] = -1e30
    loaded = refapi.RefLoadedBooster(text)
    raw = loaded.predict(Xt, raw_score=True)
    out = loaded.predict(Xt, raw_score=False)
    raw64 = loaded.predict(Xt.astype(np.float64), raw_score=True)
    assert np.array_equal(raw, raw64)
    path = os.path.join(HERE, f"model_{name}.npz")
    np.savez_compressed(path, model=np.frombuffer(text.encode(), np.uint8), X=Xt, raw=raw, out=out)
    print(name, len(text), "bytes of model text,", loaded.num_iterations, "iterations ->", os.path.getsize(path) // 1024, "KB")
    loaded.free(); bst.free(); ds.free()
