"""Runs in a subprocess with LGBM_REF_LIB pointing at ONE lib_lightgbm.so (the unmodified reference or the
drop-in of integration/): trains through the real LGBM_* C API and prints the model + train predictions as JSON."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import refapi  # noqa: E402


def main():
    device, n, f, iters, case = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    rng = np.random.default_rng(17)
    quant = case.endswith("_quant")        # Config::use_quantized_grad through the real C API
    case = case.replace("_quant", "")
    if case == "identity":
        X = rng.integers(0, 255, (n, f)).astype(np.float32)
        y = ((X[:, :6] / 127.0 - 1) @ rng.normal(size=6) + 0.3 * rng.normal(size=n)).astype(np.float32)
        extra = dict(enable_bundle="false")
        obj = "regression"
    else:   # continuous with NaN + sparse column, binary objective
        X = rng.normal(size=(n, f)).astype(np.float32)
        X[rng.random((n, f)) < 0.05] = np.nan
        X[:, 2] = np.where(rng.random(n) < 0.8, 0.0, X[:, 2])
        logit = np.nan_to_num(X[:, 0]) - 0.5 * np.nan_to_num(X[:, 1]) + np.isnan(X[:, 3]) * 0.7
        y = (rng.random(n) < 1 / (1 + np.exp(-logit))).astype(np.float32)
        extra = dict(max_bin=63)
        obj = "binary"
    dsp = dict(verbosity=-1, num_threads=4, min_data_in_bin=1, feature_pre_filter="false", device_type="cuda", **extra)
    ds = refapi.RefDataset(X, y, dsp)
    bp = dict(dsp, objective=obj, num_leaves=31, learning_rate=0.1, min_data_in_leaf=20, device_type=device,
              force_row_wise="true", deterministic="true", num_threads=1 if device == "cpu" else 4)
    if device == "cuda" and int(os.environ.get("DROPIN_NUM_GPU", "1")) > 1:
        bp["num_gpu"] = int(os.environ["DROPIN_NUM_GPU"])      # features sharded over the GPUs of this box, one process
    if quant:
        bp.update(use_quantized_grad="true", stochastic_rounding="false", num_grad_quant_bins=4 if case == "identity" else 8,
                  quant_train_renew_leaf="false" if case == "identity" else "true")
    bst = refapi.RefBooster(ds, bp)
    for _ in range(iters):
        bst.update()
    trees = bst.trees()
    out = dict(num_trees=len(trees), pred=bst.inner_predict().tolist()[:2000],
               trees=[dict(split_feature=t.split_feature.tolist(), threshold=t.threshold.tolist(),
                           leaf_count=t.leaf_count.tolist(), leaf_value=t.leaf_value.tolist(),
                           split_gain=t.split_gain.tolist()) for t in trees])
    print("JSON" + json.dumps(out))


if __name__ == "__main__":
    main()
