"""Trains a few trees on an N x F synthetic matrix (ncu captures of the non-histogram kernels, A/B of launch
options).  Prints the device time per tree (CUDA events on the learner's stream) and a hash of the split records so
that two runs with different LGBMB200_DEBUG bits can be checked for identical trees."""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgbm_b200 as lgb

n = int(os.environ.get("TB_ROWS", 2_000_000)); f = int(os.environ.get("TB_COLS", 1024)); leaves = int(os.environ.get("TB_LEAVES", 63))
trees = int(os.environ.get("TB_TREES", 3))
rng = np.random.default_rng(0)
bins = rng.integers(0, 255, (n, f), dtype=np.uint8)
y = ((bins[:, :32] / 127.0 - 1) @ rng.normal(size=32) + 0.5 * rng.normal(size=n)).astype(np.float32)
qb = int(os.environ.get("TB_QUANT", 0))          # > 0: use_quantized_grad with that many num_grad_quant_bins
cfg = lgb.Config(num_leaves=leaves, use_cuda_graph=os.environ.get("TB_GRAPH", "1") == "1", use_quantized_grad=qb > 0,
                 num_grad_quant_bins=max(qb, 2), stochastic_rounding=os.environ.get("TB_STOCH", "0") == "1")
B = lgb.B200Booster(lgb.Layout.identity(bins), y, cfg, learning_rate=0.1)
ms, h = [], hashlib.sha1()
for _ in range(trees):
    B.learner.timer_start()
    t = B.update()
    ms.append(B.learner.timer_stop())
    h.update(t.splits.tobytes()); h.update(t.leaf_value.tobytes())
warm = ms[min(2, len(ms) - 1):]
if os.environ.get("TB_PROFILE"):
    B.learner.set_profiling(True); B.learner.hist_stats(reset=True); B.update(); B.learner.set_profiling(False)
    print("by_kind_ms", {k: round(v, 3) for k, v in B.learner.profile_by_kind().items()}, "l2", B.l2())
print(f"quant_bins={qb} debug={os.environ.get('LGBMB200_DEBUG', '0')} rows={n} cols={f} leaves={leaves} device_ms_per_tree mean={np.mean(warm):.3f} min={np.min(warm):.3f} "
      f"trees_sha1={h.hexdigest()[:16]}")
