"""Generates tests/golden/binning_*.npz: float matrices (tests/binning_cases.py) together with the Dataset the UNMODIFIED
reference builds from them — LGBM_DatasetCreateFromMat with device_type=cuda in the Dataset parameters (dense storage,
bundles capped at 256 stored values: the Dataset a cuda tree learner is Init-ed with), read back through
oracle/ref_probe.cpp: per-feature layout, bin upper bounds, and every stored byte.

Run in the build container (needs /root/reference compiled into oracle/_ref): python tests/golden/make_binning_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from binning_cases import CASES  # noqa: E402
from oracle import refapi  # noqa: E402

for name, fn in CASES.items():
    X, params = fn()
    ds = refapi.RefDataset(X, None, dict(params, device_type="cuda", verbosity=-1))
    lay = ds.layout()
    ds.free()
    d = lay.to_npz_dict(with_bins=True)
    d["X"] = X
    d["params"] = np.frombuffer(json.dumps(params).encode(), dtype=np.uint8)
    path = os.path.join(HERE, f"binning_{name}.npz")
    np.savez_compressed(path, **d)
    print(f"{name}: X {X.shape} {X.dtype} -> {lay.num_columns} columns, {lay.num_features} features, {os.path.getsize(path) / 1024:.0f} KB")
