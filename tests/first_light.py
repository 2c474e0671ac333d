"""Diagnostic script for the first GPU run (not a pytest file)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lightgbm_b200 as lgb
from oracle import oracle_py
from helpers import synth_identity
from test_gpu_parity import lay_for_oracle

n, f, leaves = 50000, 40, 31
bins, y, g, h = synth_identity(n, f, seed=1)
lay = lgb.Layout.identity(bins)
L = lgb.B200TreeLearner(lgb.Config(num_leaves=leaves, use_cuda_graph=False))
L.init(lay)
got, ms = L.construct_histogram(g, h)
want = oracle_py.construct_histogram(lay, None, g, h)
print("hist max abs diff", np.max(np.abs(got - want)), "max rel", np.max(np.abs(got - want) / np.maximum(np.abs(want), 1)), "ms", ms)
t = L.train(g, h)
o = oracle_py.train_tree(lay_for_oracle(lay), g, h, num_leaves=leaves)
print("leaves", t.num_leaves, o.num_leaves)
for i in range(min(t.num_leaves, o.num_leaves) - 1):
    a, b = t.splits[i], o.splits[i]
    ok = all(a[k] == b[k] for k in ("leaf", "feature", "threshold", "default_left", "left_count", "right_count"))
    print(i, "OK " if ok else "XX ", a["leaf"], a["feature"], a["threshold"], a["left_count"], a["right_count"], f"{a['gain']:.6f}",
          "|", b["leaf"], b["feature"], b["threshold"], b["left_count"], b["right_count"], f"{b['gain']:.6f}")
print("leaf_value maxdiff", np.max(np.abs(t.leaf_value[:o.num_leaves] - o.leaf_value[:t.num_leaves])))

# timing at a larger size
for (n, f, leaves) in [(1_000_000, 256, 63)]:
    rng = np.random.default_rng(0)
    bins = rng.integers(0, 255, (n, f), dtype=np.uint8)
    y = ((bins[:, :32] / 127.0 - 1) @ rng.normal(size=32) + 0.5 * rng.normal(size=n)).astype(np.float32)
    lay = lgb.Layout.identity(bins)
    for graph in (False, True):
        B = lgb.B200Booster(lay, y, lgb.Config(num_leaves=leaves, use_cuda_graph=graph), learning_rate=0.1)
        B.update(); B.update()
        t0 = time.time()
        for _ in range(10):
            B.update()
        dt = (time.time() - t0) / 10
        print(f"{n}x{f} leaves={leaves} graph={graph}: {dt*1e3:.2f} ms/iter  l2={B.l2():.5f}")
    B.learner.set_profiling(True)
    B.learner.hist_stats(reset=True)
    for _ in range(3):
        B.update()
    ms, rows, nl = B.learner.hist_stats()
    print(f"hist: {ms/3:.3f} ms/iter over {rows/3:.0f} rows/iter, {nl/3:.0f} launches/iter -> {rows*f/ms/1e6:.1f} GB/s bin bytes")
    _, ms1 = B.learner.construct_histogram(B.d_grad, B.d_hess, None, want_hist=False)
    print(f"root histogram alone: {ms1:.3f} ms -> {n*f/ms1/1e6:.1f} GB/s")
