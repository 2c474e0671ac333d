"""CPU: bench.py's synthetic-data helpers (the matrix must not depend on thread count or rank count)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_gen_bins_is_slice_and_thread_invariant():
    a = bench.gen_bins(70000, 300, 44, threads=1)
    b = bench.gen_bins(70000, 300, 44, col_lo=96, col_hi=290, threads=5)
    assert a.dtype == np.uint8 and a.max() <= 254
    assert np.array_equal(a[:, 96:290], b)
    assert not np.array_equal(a, bench.gen_bins(70000, 300, 45))


def test_gen_bins_row_ranges_tile_the_matrix():
    a = bench.gen_bins(200000, 160, 7)
    b = bench.gen_bins(200000, 160, 7, row_lo=65536, row_hi=150000)
    assert np.array_equal(a[65536:150000], b)
    c = bench.gen_bins(200000, 160, 7, col_lo=128, col_hi=160, row_lo=131072)
    assert np.array_equal(a[131072:, 128:], c)


def test_effective_cores_is_sane():
    n = bench.effective_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "C2", "--rows", "20000",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "iters/sec" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0
    # the arm trains on exactly the workload it prints: no sampling, no scaling
    assert "20000 rows" in d["config"]["workload"] and "all 20000 rows" in d["cpu_baseline"]["sample"]
    assert abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-6 * 1e3
