"""Seeded float matrices for the Dataset-construction parity tests.  Each case returns (X, params);
tests/golden/make_binning_golden.py runs them through the unmodified reference and stores X with the result."""
import numpy as np

BASE = dict(max_bin=255, min_data_in_bin=3, min_data_in_leaf=20, bin_construct_sample_cnt=200000, data_random_seed=1,
            feature_pre_filter="true", use_missing="true", zero_as_missing="false", enable_bundle="true")


def _p(**kw):
    d = dict(BASE); d.update(kw); return d


def dense_normal():
    """continuous values: > 255 distinct per column, the equal-count regime of the greedy bin finder"""
    r = np.random.default_rng(11)
    X = r.normal(size=(6000, 10)).astype(np.float32)
    X[:, 3] = np.abs(X[:, 3])                       # positive only
    X[:, 4] = -np.abs(X[:, 4])                      # negative only
    X[:, 5] *= 1e6
    return X, _p()


def sampled_rows():
    """num_data > bin_construct_sample_cnt: the reference's row sample (set-based branch of Random::Sample)"""
    r = np.random.default_rng(12)
    X = r.gamma(2.0, size=(20000, 6)).astype(np.float32) - 1.5
    return X, _p(bin_construct_sample_cnt=3000, data_random_seed=7)


def sampled_rows_dense_branch():
    """sample count close to num_data: the sequential-probability branch of Random::Sample"""
    r = np.random.default_rng(13)
    X = r.uniform(-3, 3, size=(5000, 5)).astype(np.float64)
    return X, _p(bin_construct_sample_cnt=2500, data_random_seed=3)


def missing_values():
    """NaN handling: NaN-type mappers (last bin), columns without NaN next to them, an all-NaN tail"""
    r = np.random.default_rng(14)
    X = r.normal(size=(5000, 8)).astype(np.float32)
    X[r.random(5000) < 0.1, 0] = np.nan
    X[r.random(5000) < 0.6, 1] = np.nan
    X[:, 2] = np.round(X[:, 2] * 2)                 # few distinct values incl. zero
    X[r.random(5000) < 0.05, 2] = np.nan
    X[r.random(5000) < 0.3, 3] = 0.0
    return X, _p()


def zero_as_missing():
    r = np.random.default_rng(15)
    X = r.normal(size=(5000, 6)).astype(np.float32)
    X[r.random(5000) < 0.4, 0] = 0.0
    X[r.random(5000) < 0.2, 1] = np.nan
    X[:, 2] = np.where(r.random(5000) < 0.5, 0.0, 1.0)
    return X, _p(zero_as_missing="true")


def no_missing_handling():
    r = np.random.default_rng(16)
    X = r.normal(size=(4000, 5)).astype(np.float32)
    X[r.random(4000) < 0.2, 0] = np.nan
    return X, _p(use_missing="false")


def few_values_and_trivial():
    """integer-valued and constant columns, a column that feature_pre_filter must drop, small max_bin"""
    r = np.random.default_rng(17)
    n = 5000
    X = np.zeros((n, 9), np.float32)
    X[:, 0] = r.integers(0, 5, n)
    X[:, 1] = 3.25                                   # constant: trivial
    X[:, 2] = r.integers(-3, 4, n)
    X[:, 3] = (r.random(n) < 0.001)                  # 5 rows set: cannot be split with min_data_in_leaf=20
    X[:, 4] = r.integers(0, 1000, n)
    X[:, 5] = r.integers(0, 40, n) * 0.5 - 7
    X[:, 6] = np.where(r.random(n) < 0.9, 0.0, r.normal(size=n))      # zero is the most frequent bin
    X[:, 7] = np.where(r.random(n) < 0.8, 5.0, r.normal(size=n))      # a non-zero value is the most frequent bin
    X[:, 8] = r.normal(size=n)
    return X, _p(max_bin=63, min_data_in_bin=5)


def big_count_values():
    """a few heavy values among continuous ones: the `is_big_count_value` path of the greedy finder"""
    r = np.random.default_rng(18)
    n = 8000
    X = r.normal(size=(n, 4)).astype(np.float32)
    X[r.random(n) < 0.3, 0] = 1.5
    X[r.random(n) < 0.2, 0] = -0.25
    X[r.random(n) < 0.5, 1] = 2.0
    X[:, 2] = np.round(X[:, 2], 1)
    return X, _p(max_bin=32, min_data_in_bin=10)


def efb_exclusive():
    """mutually exclusive sparse features: bundled into shared columns (EFB), 64 features -> 16 columns"""
    r = np.random.default_rng(19)
    n, f = 12000, 64
    X = np.zeros((n, f), np.float32)
    for b in range(f // 4):
        which = r.integers(0, 4, n)
        on = r.random(n) < 0.08
        vals = r.integers(1, 64, n).astype(np.float32)
        for k in range(4):
            m = on & (which == k)
            X[m, 4 * b + k] = vals[m]
    return X, _p(min_data_in_bin=1)


def efb_mixed_dense_sparse():
    """dense and sparse columns together; exclusive blocks whose features carry ~120 distinct values each, so the
    256-stored-values cap of a cuda Dataset splits a block over two columns; two conflict rows inside the budget
    (total_sample_cnt / 10000 = 3) that the bundle search must tolerate, and one feature whose conflicts exceed it"""
    r = np.random.default_rng(20)
    n = 30000
    X = np.zeros((n, 25), np.float32)
    X[:, :4] = r.normal(size=(n, 4))
    for b in range(5):
        which = r.integers(0, 4, n)
        on = r.random(n) < 0.1
        hi = 120 if b < 3 else 30
        vals = r.integers(1, hi, n).astype(np.float32)
        for k in range(4):
            m = on & (which == k)
            X[m, 4 + 4 * b + k] = vals[m]
    rows = np.nonzero(X[:, 20] != 0)[0][:2]         # block 4: two rows where features 20 and 21 are both set
    X[rows, 21] = 7.0
    m = r.random(n) < 0.05                           # feature 24 collides with everything
    X[m, 24] = r.integers(1, 10, int(m.sum()))
    return X, _p(min_data_in_bin=1)


def no_bundle():
    r = np.random.default_rng(21)
    n = 4000
    X = np.zeros((n, 12), np.float32)
    for j in range(12):
        m = r.random(n) < 0.1
        X[m, j] = r.normal(size=int(m.sum()))
    return X, _p(enable_bundle="false", feature_pre_filter="false")


CASES = dict(dense_normal=dense_normal, sampled_rows=sampled_rows, sampled_rows_dense_branch=sampled_rows_dense_branch,
             missing_values=missing_values, zero_as_missing=zero_as_missing, no_missing_handling=no_missing_handling,
             few_values_and_trivial=few_values_and_trivial, big_count_values=big_count_values, efb_exclusive=efb_exclusive,
             efb_mixed_dense_sparse=efb_mixed_dense_sparse, no_bundle=no_bundle)
