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


def test_c4_generator_raw_and_bundled_views_agree():
    """C4: the raw sparse features (fed to the reference arms) and the EFB-bundled columns (this repo's arm) describe the
    same matrix, features of a block of 4 are mutually exclusive, and any 64K-aligned row range / column range of the
    bundled view equals the corresponding slice."""
    wl = dict(bench.WORKLOADS["C4"], rows=140000, cols=512)
    cols = bench.gen_columns(wl)
    raw = bench.gen_efb4(wl["rows"], wl["cols"], wl["seed"], raw=True)
    assert cols.shape == (140000, 128) and raw.shape == (140000, 512) and cols.max() <= 252
    assert (raw.reshape(len(raw), -1, 4) > 0).sum(axis=2).max() == 1
    r, c = np.nonzero(raw)
    assert np.array_equal(cols[r, c // 4], 1 + bench.EFB_VALUES * (c % 4) + raw[r, c] - 1)
    assert np.array_equal((cols > 0), (raw.reshape(len(raw), -1, 4) > 0).any(axis=2))
    assert abs((cols > 0).mean() - (1 - (1 - bench.EFB_P) ** 4)) < 3e-3
    part = bench.gen_columns(wl, 32, 96, row_lo=65536)
    assert np.array_equal(part, cols[65536:, 32:96])
    y = bench.gen_label_wl(wl, cols[:, :bench.LABEL_COLS["efb4"]])
    assert set(np.unique(y)) == {0.0, 1.0} and 0.3 < y.mean() < 0.7


def test_c5_generator_and_row_sliced_labels():
    wl = dict(bench.WORKLOADS["C5"], rows=150000)
    h = bench.gen_columns(wl)
    assert h.shape == (150000, 28) and h.max() <= 254
    assert np.array_equal(h[:, 21], ((h[:, 0].astype(int) + h[:, 1] + h[:, 2]) // 3).astype(np.uint8))
    y = bench.gen_label_wl(wl, h)
    assert np.array_equal(y[65536:], bench.gen_label_wl(wl, h[65536:], row_lo=65536))     # a rank's row slice gets the same labels


def test_c4_reference_dataset_bundles_like_the_generator():
    """The reference's own Dataset construction (cuda rules, sampled-column API + PushRows) turns the raw C4 features into
    exactly the bundled columns the generator writes directly (same bundles, offsets 1 + 63 j, most-frequent bin elided)."""
    from oracle import refapi
    if not refapi.available():
        import pytest
        pytest.skip("oracle/_ref not built")
    wl = dict(bench.WORKLOADS["C4"], rows=70000, cols=64)
    dsp, _ = bench._ref_params(wl, 2, "cpu")
    ds, _ = bench._ref_dataset(refapi, wl, wl["rows"], dsp, 2)
    lay = ds.layout()
    ds.free()
    assert lay.num_columns == 16 and lay.num_features == 64
    raw = bench.gen_efb4(wl["rows"], wl["cols"], wl["seed"], raw=True)
    for f in range(lay.num_features):
        rf = int(lay.feat_real_index[f])
        assert lay.feat_num_bin[f] == bench.EFB_VALUES + 1 and lay.feat_mfb[f] == 0
        nz = np.nonzero(raw[:, rf])[0]
        assert np.array_equal(lay.bins[nz, lay.feat_column[f]], lay.feat_lo[f] + raw[nz, rf] - 1)
