"""Throughput of the device value->bin pass (k_value_to_bin): float32 matrix resident in HBM -> uint8 bins in HBM, and the
same through the host-facing call (pageable host matrix, chunked H2D inside the timed region)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lightgbm_b200 as lgb
from lightgbm_b200.tree_learner import DeviceArray

n = int(os.environ.get("BB_ROWS", 2_000_000)); f = int(os.environ.get("BB_COLS", 1024))
r = np.random.default_rng(0)
X = np.empty((n, f), np.float32)
for lo in range(0, n, 1 << 18):
    X[lo:lo + (1 << 18)] = r.standard_normal((min(1 << 18, n - lo), f), dtype=np.float32)
t0 = time.time()
b = lgb.Binner({}).fit(X)
print(f"fit (host: sample {len(b.sample_indices())} rows, {f} mappers): {time.time() - t0:.2f} s")
dx = DeviceArray(X.nbytes).upload(X)
for _ in range(3):
    out, ms = b.transform((dx, np.float32), to_device=True, data_rows=n)
m = b.layout_meta()
byts = n * (f * 4 + m["num_columns"])
print(f"device-resident {n} x {f} float32 -> {m['num_columns']} byte columns: {ms:.3f} ms, {byts / ms / 1e6:.0f} GB/s algorithmic (4 B read + 1 B written per cell)")
t0 = time.time()
out2, ms2 = b.transform(X, to_device=True)
print(f"from the host matrix (pageable, chunked H2D on two streams): {ms2:.1f} ms device time, {time.time() - t0:.2f} s wall, {X.nbytes / ms2 / 1e6:.1f} GB/s of input")
assert np.array_equal(out.download()[:100000], out2.download()[:100000])
