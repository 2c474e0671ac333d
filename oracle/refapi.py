"""ctypes harness over the UNMODIFIED reference library built by oracle/Makefile.ref
(oracle/_ref/lib_lightgbm.so) plus the probe helper (oracle/_ref/libref_probe.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, tests/golden/make_golden.py and bench.py's
reference / cpu_baseline legs.  Nothing under lightgbm_b200/ may import this module.

Mirrors the few LGBM_* entry points of include/LightGBM/c_api.h that python-package/lightgbm/basic.py
uses for `lgb.Dataset(np2d)` + `Booster.update` (SURVEY.md §8b "minimum export set").
"""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
# LGBM_REF_LIB lets the drop-in test point the very same harness at integration/_build/lib_lightgbm.so
REF_LIB = os.environ.get("LGBM_REF_LIB", os.path.join(REF_DIR, "lib_lightgbm.so"))
PROBE_LIB = os.path.join(REF_DIR, "libref_probe.so")

C_API_DTYPE_FLOAT32, C_API_DTYPE_FLOAT64, C_API_DTYPE_INT32 = 0, 1, 2


def available() -> bool:
    return os.path.exists(REF_LIB) and os.path.exists(PROBE_LIB)


_lib = None
_probe = None


def lib():
    global _lib, _probe
    if _lib is None:
        if not available():
            raise RuntimeError(f"reference library not built: run `make -C oracle -f Makefile.ref` ({REF_LIB})")
        _lib = C.CDLL(REF_LIB)
        _lib.LGBM_GetLastError.restype = C.c_char_p
        _probe = C.CDLL(PROBE_LIB)
    return _lib


def _check(ret: int):
    if ret != 0:
        raise RuntimeError("reference LightGBM error: " + lib().LGBM_GetLastError().decode())


def params_str(params: dict) -> bytes:
    return " ".join(f"{k}={v}" for k, v in params.items()).encode()


class RefDataset:
    """LGBM_DatasetCreateFromMat + LGBM_DatasetSetField (c_api.h:409, :552)."""

    def __init__(self, X: np.ndarray, label: np.ndarray | None, params: dict):
        L = lib()
        X = np.ascontiguousarray(X)
        assert X.dtype in (np.float32, np.float64) and X.ndim == 2
        dtype = C_API_DTYPE_FLOAT32 if X.dtype == np.float32 else C_API_DTYPE_FLOAT64
        self.handle = C.c_void_p()
        _check(L.LGBM_DatasetCreateFromMat(X.ctypes.data_as(C.c_void_p), C.c_int(dtype), C.c_int32(X.shape[0]),
                                           C.c_int32(X.shape[1]), C.c_int(1), C.c_char_p(params_str(params)),
                                           None, C.byref(self.handle)))
        if label is not None:
            lab = np.ascontiguousarray(label, dtype=np.float32)
            _check(L.LGBM_DatasetSetField(self.handle, b"label", lab.ctypes.data_as(C.c_void_p),
                                          C.c_int(len(lab)), C.c_int(C_API_DTYPE_FLOAT32)))
        self.num_data = X.shape[0]

    def set_label(self, label: np.ndarray):
        lab = np.ascontiguousarray(label, dtype=np.float32)
        _check(lib().LGBM_DatasetSetField(self.handle, b"label", lab.ctypes.data_as(C.c_void_p),
                                          C.c_int(len(lab)), C.c_int(C_API_DTYPE_FLOAT32)))

    def layout(self) -> "Layout":
        lib()
        n, c, f, tf = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _probe.RefProbe_Dims(self.handle, C.byref(n), C.byref(c), C.byref(f), C.byref(tf))
        F = f.value
        arrs = [np.zeros(F, dtype=np.int32) for _ in range(9)]
        r = _probe.RefProbe_Layout(self.handle, *[a.ctypes.data_as(C.c_void_p) for a in arrs])
        if r != 0:
            raise RuntimeError(f"RefProbe_Layout failed ({r}): multi-val group in dataset")
        bins = np.zeros((n.value, c.value), dtype=np.uint8)
        r = _probe.RefProbe_Bins(self.handle, bins.ctypes.data_as(C.c_void_p))
        if r != 0:
            raise RuntimeError(f"RefProbe_Bins failed ({r})")
        ub = []
        for i in range(F):
            o = np.zeros(int(arrs[2][i]), dtype=np.float64)
            _probe.RefProbe_BinUpperBounds(self.handle, C.c_int(i), o.ctypes.data_as(C.c_void_p))
            ub.append(o)
        if np.any(arrs[8] != 0):
            raise RuntimeError("categorical features are outside the hot-path contract")
        return Layout(num_data=n.value, num_columns=c.value, num_features=F, num_total_features=tf.value,
                      feat_column=arrs[0], feat_lo=arrs[1], feat_num_bin=arrs[2], feat_mfb=arrs[3],
                      feat_default_bin=arrs[4], feat_missing=arrs[5], feat_real_index=arrs[6],
                      feat_in_group=arrs[7], bins=bins, bin_upper_bound=ub)

    def free(self):
        if self.handle:
            lib().LGBM_DatasetFree(self.handle)
            self.handle = None


@dataclass
class Layout:
    """The layout contract of SURVEY.md §8 a15 (see oracle/lgbm_oracle.h OrcLayout)."""
    num_data: int
    num_columns: int
    num_features: int
    num_total_features: int
    feat_column: np.ndarray
    feat_lo: np.ndarray
    feat_num_bin: np.ndarray
    feat_mfb: np.ndarray
    feat_default_bin: np.ndarray
    feat_missing: np.ndarray
    feat_real_index: np.ndarray
    feat_in_group: np.ndarray
    bins: np.ndarray                      # [N, C] uint8 stored group values
    bin_upper_bound: list = field(default_factory=list)

    META = ("feat_column", "feat_lo", "feat_num_bin", "feat_mfb", "feat_default_bin", "feat_missing",
            "feat_real_index", "feat_in_group")

    def to_npz_dict(self, with_bins=True) -> dict:
        d = {k: getattr(self, k) for k in self.META}
        d["dims"] = np.array([self.num_data, self.num_columns, self.num_features, self.num_total_features], np.int64)
        if with_bins:
            d["bins"] = self.bins
        if self.bin_upper_bound:
            d["ub_concat"] = np.concatenate(self.bin_upper_bound)
        return d

    @staticmethod
    def from_npz_dict(d) -> "Layout":
        dims = d["dims"]
        lay = Layout(int(dims[0]), int(dims[1]), int(dims[2]), int(dims[3]),
                     *[np.ascontiguousarray(d[k], dtype=np.int32) for k in Layout.META],
                     bins=np.ascontiguousarray(d["bins"]) if "bins" in d else None)
        if "ub_concat" in d:
            ub, o = [], 0
            for nb in lay.feat_num_bin:
                ub.append(np.asarray(d["ub_concat"][o:o + nb])); o += nb
            lay.bin_upper_bound = ub
        return lay

    @staticmethod
    def identity(bins: np.ndarray, num_bin: int = 255) -> "Layout":
        """bin == stored value, one feature per column, most_freq_bin = default_bin = 0, no missing:
        what the reference produces for integer-valued features."""
        N, F = bins.shape
        z = np.zeros(F, np.int32)
        return Layout(N, F, F, F, np.arange(F, dtype=np.int32), np.ones(F, np.int32), np.full(F, num_bin, np.int32),
                      z.copy(), z.copy(), z.copy(), np.arange(F, dtype=np.int32), np.ones(F, np.int32),
                      bins=np.ascontiguousarray(bins, dtype=np.uint8))


@dataclass
class RefTree:
    """One tree parsed back from the model text (gbdt_model_text.cpp:314+, tree.cpp Tree::ToString)."""
    num_leaves: int
    split_feature: np.ndarray      # real feature index per node (node i == split i)
    split_gain: np.ndarray         # float32 as stored
    threshold: np.ndarray          # real-valued
    decision_type: np.ndarray
    left_child: np.ndarray
    right_child: np.ndarray
    leaf_value: np.ndarray
    leaf_weight: np.ndarray
    leaf_count: np.ndarray
    internal_value: np.ndarray
    internal_weight: np.ndarray
    internal_count: np.ndarray
    shrinkage: float = 1.0

    @property
    def default_left(self):
        return (self.decision_type.astype(np.int64) >> 1) & 1

    @property
    def missing_type(self):
        return (self.decision_type.astype(np.int64) >> 2) & 3

    def split_leaf(self) -> np.ndarray:
        """leaf id that node i split: the left child keeps the parent's leaf id (tree.h:543-585)."""
        out = np.zeros(self.num_leaves - 1, dtype=np.int32)
        for i in range(self.num_leaves - 2, -1, -1):
            lc = self.left_child[i]
            out[i] = ~lc if lc < 0 else out[lc]
        return out


def parse_model_trees(model_str: str) -> list:
    trees = []
    for blk in re.split(r"\nTree=\d+\n", "\n" + model_str)[1:]:
        kv = {}
        for line in blk.split("\n"):
            if "=" in line:
                k, v = line.split("=", 1)
                kv[k] = v
            if line.startswith("end of trees"):
                break
        nl = int(kv["num_leaves"])

        def arr(k, dt, n):
            if n == 0 or k not in kv or kv[k].strip() == "":
                return np.zeros(0, dtype=dt)
            return np.array(kv[k].split(" "), dtype=np.float64).astype(dt)
        trees.append(RefTree(
            num_leaves=nl,
            split_feature=arr("split_feature", np.int32, nl - 1), split_gain=arr("split_gain", np.float64, nl - 1),
            threshold=arr("threshold", np.float64, nl - 1), decision_type=arr("decision_type", np.int32, nl - 1),
            left_child=arr("left_child", np.int32, nl - 1), right_child=arr("right_child", np.int32, nl - 1),
            leaf_value=arr("leaf_value", np.float64, nl), leaf_weight=arr("leaf_weight", np.float64, nl),
            leaf_count=arr("leaf_count", np.int64, nl), internal_value=arr("internal_value", np.float64, nl - 1),
            internal_weight=arr("internal_weight", np.float64, nl - 1),
            internal_count=arr("internal_count", np.int64, nl - 1), shrinkage=float(kv.get("shrinkage", 1.0))))
    return trees


class RefBooster:
    """LGBM_BoosterCreate / UpdateOneIter / UpdateOneIterCustom / SaveModelToString (c_api.h:656,769,801)."""

    def __init__(self, train: RefDataset, params: dict):
        self.handle = C.c_void_p()
        self.train = train
        _check(lib().LGBM_BoosterCreate(train.handle, C.c_char_p(params_str(params)), C.byref(self.handle)))

    def update(self) -> bool:
        fin = C.c_int(0)
        _check(lib().LGBM_BoosterUpdateOneIter(self.handle, C.byref(fin)))
        return bool(fin.value)

    def update_custom(self, grad: np.ndarray, hess: np.ndarray) -> bool:
        g = np.ascontiguousarray(grad, dtype=np.float32)
        h = np.ascontiguousarray(hess, dtype=np.float32)
        fin = C.c_int(0)
        _check(lib().LGBM_BoosterUpdateOneIterCustom(self.handle, g.ctypes.data_as(C.c_void_p),
                                                     h.ctypes.data_as(C.c_void_p), C.byref(fin)))
        return bool(fin.value)

    def model_string(self) -> str:
        n = C.c_int64(0)
        buf = C.create_string_buffer(1 << 20)
        _check(lib().LGBM_BoosterSaveModelToString(self.handle, C.c_int(0), C.c_int(-1), C.c_int(0),
                                                   C.c_int64(len(buf)), C.byref(n), buf))
        if n.value > len(buf):
            buf = C.create_string_buffer(n.value)
            _check(lib().LGBM_BoosterSaveModelToString(self.handle, C.c_int(0), C.c_int(-1), C.c_int(0),
                                                       C.c_int64(len(buf)), C.byref(n), buf))
        return buf.value.decode()

    def trees(self) -> list:
        return parse_model_trees(self.model_string())

    def inner_predict(self) -> np.ndarray:
        """LGBM_BoosterGetPredict(data_idx=0): the training scores."""
        n = C.c_int64(0)
        _check(lib().LGBM_BoosterGetNumPredict(self.handle, C.c_int(0), C.byref(n)))
        out = np.zeros(n.value, dtype=np.float64)
        _check(lib().LGBM_BoosterGetPredict(self.handle, C.c_int(0), C.byref(n), out.ctypes.data_as(C.c_void_p)))
        return out

    def free(self):
        if self.handle:
            lib().LGBM_BoosterFree(self.handle)
            self.handle = None


class RefLoadedBooster:
    """LGBM_BoosterLoadModelFromString + LGBM_BoosterPredictForMat (c_api.h:690, :1283): the reference scoring a dense
    matrix with a model text — its own, or one written by lightgbm_b200/model.py."""

    def __init__(self, model_str: str):
        self.handle = C.c_void_p()
        it = C.c_int(0)
        _check(lib().LGBM_BoosterLoadModelFromString(C.c_char_p(model_str.encode()), C.byref(it), C.byref(self.handle)))
        self.num_iterations = it.value

    def predict(self, X: np.ndarray, raw_score: bool = True) -> np.ndarray:
        X = np.ascontiguousarray(X)
        assert X.dtype in (np.float32, np.float64) and X.ndim == 2
        out = np.zeros(X.shape[0], np.float64)
        n = C.c_int64(0)
        _check(lib().LGBM_BoosterPredictForMat(self.handle, X.ctypes.data_as(C.c_void_p),
                                               C.c_int(C_API_DTYPE_FLOAT32 if X.dtype == np.float32 else C_API_DTYPE_FLOAT64),
                                               C.c_int32(X.shape[0]), C.c_int32(X.shape[1]), C.c_int(1),
                                               C.c_int(1 if raw_score else 0), C.c_int(0), C.c_int(-1), C.c_char_p(b"verbosity=-1"),
                                               C.byref(n), out.ctypes.data_as(C.c_void_p)))
        assert n.value == X.shape[0]
        return out

    model_string = RefBooster.model_string

    def free(self):
        if self.handle:
            lib().LGBM_BoosterFree(self.handle)
            self.handle = None


class RefDatasetStreamed(RefDataset):
    """The reference's own streaming ingestion (c_api.h LGBM_DatasetCreateByReference + LGBM_DatasetPushRows): bin
    mappers come from a sample Dataset built with LGBM_DatasetCreateFromMat on the first `sample_rows` rows, then the
    full matrix is pushed in row blocks, so the fp32 copy of a 10M x 1024 matrix (41 GB) never exists.
    `block_fn(lo, hi)` returns the float32 [hi-lo, ncol] block of rows lo..hi; the next block is produced on a
    helper thread while the current one is being pushed."""

    def __init__(self, block_fn, nrow: int, ncol: int, label: np.ndarray | None, params: dict, block_rows: int = 250_000,
                 sample_rows: int = 65_536, sampled_columns: bool = False):
        """sampled_columns=True builds the Dataset skeleton with LGBM_DatasetCreateFromSampledColumn (bin mappers AND the
        EFB bundling decision from the sample, exactly as a file / matrix load would: DatasetLoader::ConstructFromSampleData)
        instead of LGBM_DatasetCreateByReference, whose CreateValid keeps one feature per group."""
        from concurrent.futures import ThreadPoolExecutor
        L = lib()
        first = np.ascontiguousarray(block_fn(0, min(nrow, block_rows)), dtype=np.float32)
        self.handle = C.c_void_p()
        sample = None
        if sampled_columns:
            smp = first[:min(len(first), sample_rows)]
            vals, idxs, cnts = [], [], np.zeros(ncol, dtype=np.int32)
            for c in range(ncol):
                nz = np.nonzero(smp[:, c])[0].astype(np.int32)          # zeros are implicit in the sampled-column format
                vals.append(np.ascontiguousarray(smp[nz, c], dtype=np.float64)); idxs.append(nz); cnts[c] = len(nz)
            pv = (C.POINTER(C.c_double) * ncol)(*[v.ctypes.data_as(C.POINTER(C.c_double)) for v in vals])
            pi = (C.POINTER(C.c_int) * ncol)(*[i.ctypes.data_as(C.POINTER(C.c_int)) for i in idxs])
            _check(L.LGBM_DatasetCreateFromSampledColumn(pv, pi, C.c_int32(ncol), cnts.ctypes.data_as(C.POINTER(C.c_int)),
                                                         C.c_int32(len(smp)), C.c_int32(nrow), C.c_int64(nrow),
                                                         C.c_char_p(params_str(params)), C.byref(self.handle)))
        else:
            sample = RefDataset(first[:min(len(first), sample_rows)], None, params)
            _check(L.LGBM_DatasetCreateByReference(sample.handle, C.c_int64(nrow), C.byref(self.handle)))
        with ThreadPoolExecutor(max_workers=1) as ex:
            lo, blk = 0, first
            while lo < nrow:
                hi = min(nrow, lo + block_rows)
                nxt = ex.submit(block_fn, hi, min(nrow, hi + block_rows)) if hi < nrow else None
                blk = np.ascontiguousarray(blk, dtype=np.float32)
                _check(L.LGBM_DatasetPushRows(self.handle, blk.ctypes.data_as(C.c_void_p), C.c_int(C_API_DTYPE_FLOAT32),
                                              C.c_int32(hi - lo), C.c_int32(ncol), C.c_int32(lo)))
                lo = hi
                blk = nxt.result() if nxt is not None else None
        del first
        if sample is not None:
            sample.free()
        if label is not None:
            self.set_label(label)
        self.num_data = nrow
