"""Micro-benchmark of the histogram kernel alone (root pass and gathered sub-leaf pass); used under ncu."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgbm_b200 as lgb

n = int(os.environ.get("HB_ROWS", 1_000_000)); f = int(os.environ.get("HB_COLS", 256)); reps = int(os.environ.get("HB_REPS", 5))
rng = np.random.default_rng(0)
bins = rng.integers(0, 255, (n, f), dtype=np.uint8)
g = rng.normal(size=n).astype(np.float32); h = np.ones(n, np.float32)
cfg = lgb.Config(num_leaves=4, use_cuda_graph=False)
const_h = os.environ.get("HB_CONST", "1") == "1"        # 1: constant-hessian (count-and-scale) kernel, 0: general hessians
L = lgb.B200TreeLearner(cfg)
L.init(lgb.Layout.identity(bins), is_constant_hessian=const_h)
print("constant_hessian", const_h)
from lightgbm_b200.tree_learner import DeviceArray
dg = DeviceArray(n * 4).upload(g); dh = DeviceArray(n * 4).upload(h)
for label, idx in (("root", None), ("gather50", np.sort(rng.choice(n, n // 2, replace=False)).astype(np.int32)),
                   ("gather5", np.sort(rng.choice(n, n // 20, replace=False)).astype(np.int32))):
    ts = []
    for _ in range(reps):
        _, ms = L.construct_histogram(dg, dh, idx, want_hist=False)
        ts.append(ms)
    rows = n if idx is None else len(idx)
    best = min(ts)
    print(f"{label}: rows={rows} cols={f} best {best:.4f} ms  median {np.median(ts):.4f} ms -> {rows*f/best/1e6:.1f} GB/s bin bytes, "
          f"{rows*f/best/1e6/148/1.9:.2f} cells/clk/SM@1.9GHz")
# fixed cost per pass: gathered passes over ever smaller leaves (HB_SWEEP=1)
if os.environ.get("HB_SWEEP"):
    for rows in (1_000_000, 300_000, 100_000, 30_000, 10_000, 3_000, 1_000, 100):
        if rows > n:
            continue
        idx = np.sort(rng.choice(n, rows, replace=False)).astype(np.int32)
        ts = [L.construct_histogram(dg, dh, idx, want_hist=False)[1] for _ in range(7)]
        print(f"sweep rows={rows:8d} cols={f} best {1e3 * min(ts):8.1f} us  median {1e3 * float(np.median(ts)):8.1f} us")
