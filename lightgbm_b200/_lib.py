"""Loads the C-ABI library (include/lgbm_b200.h).  No CPU fallback: if the CUDA library is missing or
cannot be loaded the import of anything that computes fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

_lib = None

EXPORTS = [
    "LGBMB200_GetLastError", "LGBMB200_LearnerCreate", "LGBMB200_LearnerInit", "LGBMB200_LearnerResetConfig",
    "LGBMB200_LearnerSetFeatureMask", "LGBMB200_LearnerSetConstantHessian", "LGBMB200_LearnerSetBaggingData", "LGBMB200_LearnerGossSample", "LGBMB200_LearnerGetBaggingData", "LGBMB200_LearnerAddPredictionAllRows", "LGBMB200_LearnerTrain",
    "LGBMB200_LearnerAddPredictionToScore", "LGBMB200_LearnerGetPartition", "LGBMB200_LearnerGetLeafHistogram",
    "LGBMB200_LearnerConstructHistogram", "LGBMB200_L2Gradients", "LGBMB200_BinaryGradients", "LGBMB200_LearnerKernelLaunches",
    "LGBMB200_LearnerHistStats", "LGBMB200_LearnerSetProfiling", "LGBMB200_LearnerProfileByKind", "LGBMB200_DeviceAlloc", "LGBMB200_DeviceFree",
    "LGBMB200_MemcpyH2D", "LGBMB200_MemcpyD2H", "LGBMB200_LearnerFree", "LGBMB200_LearnerGetLeafIndex", "LGBMB200_LearnerGetLeafIndexRange8",
    "LGBMB200_LearnerCommExport", "LGBMB200_LearnerCommConnect", "LGBMB200_LearnerCommExportPool", "LGBMB200_LearnerCommConnectRows",
    "LGBMB200_LearnersConnectLocal", "LGBMB200_LearnerCommExportColumns", "LGBMB200_LearnerCommShareColumns", "LGBMB200_LearnerTimerStart", "LGBMB200_LearnerTimerStop", "LGBMB200_HostAllocPinned", "LGBMB200_HostFreePinned",
    "LGBMB200_BinnerCreate", "LGBMB200_BinnerFit", "LGBMB200_BinnerGetLayout", "LGBMB200_BinnerGetFeatureBounds",
    "LGBMB200_BinnerGetSampleIndices", "LGBMB200_BinnerTransform", "LGBMB200_BinnerFree",
    "LGBMB200_PredictorCreate", "LGBMB200_PredictorPredict", "LGBMB200_PredictorFree",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m lightgbm_b200.build` "
                               "(or __graft_entry__.build()).  There is no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.LGBMB200_GetLastError.restype = C.c_char_p
        _lib.LGBMB200_LearnerKernelLaunches.restype = C.c_int64
        _lib.LGBMB200_LearnerKernelLaunches.argtypes = [C.c_void_p]
    return _lib


def check(ret: int):
    if ret != 0:
        raise RuntimeError("lgbm_b200: " + lib().LGBMB200_GetLastError().decode())
