"""Dataset construction from a dense float matrix: the mirror of `lightgbm.Dataset(data, label, params)`
-> LGBM_DatasetCreateFromMat (reference python-package/lightgbm/basic.py, src/c_api.cpp:1296-1408) for numerical features.

The bin mappers and the feature bundles are found on the host from the sampled rows exactly as the reference finds them
(lightgbm_b200/csrc/binning.cuh); the N x F value -> bin pass runs on the device and leaves the stored-byte matrix in HBM,
where `B200TreeLearner.init` picks it up without a round trip through the host."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib
from .tree_learner import DeviceArray, Layout, _CLayout, _p

# Dataset parameters read by the construction, with the reference's defaults (include/LightGBM/config.h)
DATASET_DEFAULTS = dict(max_bin=255, min_data_in_bin=3, min_data_in_leaf=20, bin_construct_sample_cnt=200000,
                        data_random_seed=1, feature_pre_filter=True, use_missing=True, zero_as_missing=False,
                        enable_bundle=True, gpu_device_id=-1)


class _CBinConfig(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("max_bin", "min_data_in_bin", "min_data_in_leaf", "bin_construct_sample_cnt",
                                         "data_random_seed", "feature_pre_filter", "use_missing", "zero_as_missing",
                                         "enable_bundle", "gpu_device_id")]


def _truth(v) -> int:
    if isinstance(v, str):
        return 1 if v.strip().lower() in ("true", "1", "+", "yes") else 0
    return 1 if v else 0


class DeviceMatrix:
    """[rows x cols] uint8 matrix in HBM (the output of the device binning pass)."""

    def __init__(self, rows: int, cols: int):
        self.shape = (int(rows), int(cols))
        self.buf = DeviceArray(max(1, rows * cols))

    @property
    def ptr(self):
        return self.buf.ptr

    def download(self) -> np.ndarray:
        return self.buf.download(np.uint8, self.shape[0] * self.shape[1]).reshape(self.shape)


class Binner:
    """Thin handle over LGBMB200_Binner* (include/lgbm_b200.h)."""

    def __init__(self, params: dict | None = None):
        p = dict(DATASET_DEFAULTS)
        for k, v in (params or {}).items():
            if k in p:
                p[k] = v
        self.params = p
        c = _CBinConfig(int(p["max_bin"]), int(p["min_data_in_bin"]), int(p["min_data_in_leaf"]), int(p["bin_construct_sample_cnt"]),
                        int(p["data_random_seed"]), _truth(p["feature_pre_filter"]), _truth(p["use_missing"]),
                        _truth(p["zero_as_missing"]), _truth(p["enable_bundle"]), int(p["gpu_device_id"]))
        self.handle = C.c_void_p()
        check(lib().LGBMB200_BinnerCreate(C.byref(c), C.byref(self.handle)))
        self.ncol = 0

    @staticmethod
    def _dtype_code(a: np.ndarray) -> int:
        if a.dtype == np.float32:
            return 0
        if a.dtype == np.float64:
            return 1
        raise TypeError("data must be float32 or float64")

    def fit(self, data: np.ndarray) -> "Binner":
        """Row sample, bin mappers, bundles: host work only."""
        assert data.ndim == 2
        row_major = data.flags["C_CONTIGUOUS"]
        if not row_major and not data.flags["F_CONTIGUOUS"]:
            data = np.ascontiguousarray(data); row_major = True
        self.ncol = data.shape[1]
        check(lib().LGBMB200_BinnerFit(self.handle, _p(data), C.c_int32(self._dtype_code(data)), C.c_int32(data.shape[0]),
                                       C.c_int32(data.shape[1]), C.c_int32(1 if row_major else 0)))
        return self

    def layout_meta(self) -> dict:
        cl = _CLayout()
        check(lib().LGBMB200_BinnerGetLayout(self.handle, C.byref(cl)))
        F = cl.num_features
        out = dict(num_data=cl.num_data, num_columns=cl.num_columns, num_features=F)
        for k in ("feat_column", "feat_lo", "feat_num_bin", "feat_most_freq_bin", "feat_default_bin", "feat_missing_type", "feat_real_index"):
            out[k] = np.ctypeslib.as_array(C.cast(getattr(cl, k), C.POINTER(C.c_int32)), shape=(F,)).copy()
        return out

    def bin_upper_bounds(self) -> list:
        meta = self.layout_meta()
        out = []
        for f in range(meta["num_features"]):
            nb = int(meta["feat_num_bin"][f])
            ub = np.empty(nb, np.float64)
            check(lib().LGBMB200_BinnerGetFeatureBounds(self.handle, C.c_int32(f), _p(ub), None))
            out.append(ub)
        return out

    def sample_indices(self) -> np.ndarray:
        n = C.c_int32(0)
        check(lib().LGBMB200_BinnerGetSampleIndices(self.handle, None, C.byref(n)))
        out = np.empty(n.value, np.int32)
        check(lib().LGBMB200_BinnerGetSampleIndices(self.handle, _p(out), C.byref(n)))
        return out

    def transform(self, data, to_device: bool = True, data_rows: int | None = None):
        """The value -> bin pass on the device.  `data`: host float32/float64 row-major matrix, or a (DeviceArray, dtype)
        pair already in HBM (then pass data_rows).  Returns (DeviceMatrix | np.ndarray, device milliseconds)."""
        meta = self.layout_meta()
        Ccols = meta["num_columns"]
        ms = C.c_float(0)
        if isinstance(data, tuple):
            dev, dtype = data
            rows = int(data_rows)
            src, code, on_dev = dev.ptr, (0 if np.dtype(dtype) == np.float32 else 1), 1
        else:
            data = np.ascontiguousarray(data)
            assert data.ndim == 2 and data.shape[1] == self.ncol
            rows, src, code, on_dev = data.shape[0], _p(data), self._dtype_code(data), 0
        if to_device:
            out = DeviceMatrix(rows, Ccols)
            check(lib().LGBMB200_BinnerTransform(self.handle, src, C.c_int32(code), C.c_int32(rows), C.c_int32(on_dev), out.ptr, C.c_int32(1), C.byref(ms)))
        else:
            out = np.empty((rows, Ccols), np.uint8)
            check(lib().LGBMB200_BinnerTransform(self.handle, src, C.c_int32(code), C.c_int32(rows), C.c_int32(on_dev), _p(out), C.c_int32(0), C.byref(ms)))
        return out, float(ms.value)

    def layout(self, bins) -> Layout:
        m = self.layout_meta()
        return Layout(bins, m["feat_column"], m["feat_lo"], m["feat_num_bin"], m["feat_most_freq_bin"], m["feat_default_bin"],
                      m["feat_missing_type"], m["feat_real_index"], bin_upper_bound=self.bin_upper_bounds())

    def free(self):
        if self.handle:
            lib().LGBMB200_BinnerFree(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Dataset:
    """`lightgbm.Dataset(data, label=..., params=...)` for a dense numerical matrix: construct() bins it on the device."""

    def __init__(self, data: np.ndarray, label=None, params: dict | None = None, reference: "Dataset | None" = None):
        self.data, self.params, self.reference = data, dict(params or {}), reference
        self.label = None if label is None else np.ascontiguousarray(label, dtype=np.float32)
        self.binner: Binner | None = None
        self._layout: Layout | None = None
        self.transform_ms = 0.0

    def construct(self) -> "Dataset":
        if self._layout is None:
            if self.reference is not None:          # validation data: the training set's mappers (Dataset::CreateValid)
                self.binner = self.reference.construct().binner
            else:
                self.binner = Binner(self.params).fit(self.data)
            bins, self.transform_ms = self.binner.transform(self.data, to_device=True)
            self._layout = self.binner.layout(bins)
        return self

    @property
    def layout(self) -> Layout:
        return self.construct()._layout

    def num_data(self) -> int:
        return self.data.shape[0]

    def num_feature(self) -> int:
        return self.data.shape[1]

    def get_label(self):
        return self.label
