"""Trains a few trees on an N x F synthetic matrix (for ncu captures of the non-histogram kernels)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgbm_b200 as lgb

n = int(os.environ.get("TB_ROWS", 2_000_000)); f = int(os.environ.get("TB_COLS", 1024)); leaves = int(os.environ.get("TB_LEAVES", 63))
trees = int(os.environ.get("TB_TREES", 3))
rng = np.random.default_rng(0)
bins = rng.integers(0, 255, (n, f), dtype=np.uint8)
y = ((bins[:, :32] / 127.0 - 1) @ rng.normal(size=32) + 0.5 * rng.normal(size=n)).astype(np.float32)
B = lgb.B200Booster(lgb.Layout.identity(bins), y, lgb.Config(num_leaves=leaves, use_cuda_graph=os.environ.get("TB_GRAPH", "1") == "1"), learning_rate=0.1)
for _ in range(trees):
    t0 = time.time(); B.update(); print(f"tree {time.time()-t0:.4f}s")
