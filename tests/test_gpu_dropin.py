"""GPU: the drop-in.  integration/_build/lib_lightgbm.so = the unmodified reference host objects with only
the tree-learner factory replaced.  The REAL LGBM_* C API (DatasetCreateFromMat, BoosterCreate,
BoosterUpdateOneIter, SaveModelToString, GetPredict) is driven with device_type=cuda and compared with the
unmodified reference library run with device_type=cpu on the same data."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "integration", "_build", "lib_lightgbm.so")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "lib_lightgbm.so")


def run(lib, device, n, f, iters, case, num_gpu=1):
    env = dict(os.environ, LGBM_REF_LIB=lib, DROPIN_NUM_GPU=str(num_gpu))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), device, str(n), str(f), str(iters), case],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("JSON")][-1]
    return json.loads(line[4:])


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REFLIB)), reason="drop-in / reference library not built")
@pytest.mark.parametrize("case,n,f", [("identity", 20000, 12), ("mixed", 20000, 10)])
def test_lgbm_capi_device_cuda_matches_reference_cpu(case, n, f):
    iters = 5
    gpu = run(DROPIN, "cuda", n, f, iters, case)
    cpu = run(REFLIB, "cpu", n, f, iters, case)
    assert gpu["num_trees"] == cpu["num_trees"] == iters
    # first tree: identical structure (later trees inherit fp32-level score differences)
    g0, c0 = gpu["trees"][0], cpu["trees"][0]
    k = min(8, len(c0["split_feature"]))
    assert g0["split_feature"][:k] == c0["split_feature"][:k]
    np.testing.assert_allclose(g0["threshold"][:k], c0["threshold"][:k], rtol=1e-9)
    assert g0["leaf_count"] == c0["leaf_count"] or g0["split_feature"] != c0["split_feature"]
    # the reference's own CPU<->GPU criterion (test_dual.py:35-36): predictions agree
    np.testing.assert_allclose(gpu["pred"], cpu["pred"], rtol=2e-3, atol=2e-3)


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REFLIB)), reason="drop-in / reference library not built")
@pytest.mark.parametrize("case,n,f", [("identity_quant", 20000, 12), ("mixed_quant", 20000, 10)])
def test_lgbm_capi_quantized_training_matches_reference_cpu(case, n, f):
    """use_quantized_grad=true, stochastic_rounding=false through the real LGBM_* API: integer histograms are exact, so
    the first tree of the drop-in (device_type=cuda -> B200TreeLearner) must be the reference CPU learner's tree."""
    iters = 4
    gpu = run(DROPIN, "cuda", n, f, iters, case)
    cpu = run(REFLIB, "cpu", n, f, iters, case)
    assert gpu["num_trees"] == cpu["num_trees"] == iters
    g0, c0 = gpu["trees"][0], cpu["trees"][0]
    assert g0["split_feature"] == c0["split_feature"] and g0["leaf_count"] == c0["leaf_count"]
    np.testing.assert_allclose(g0["threshold"], c0["threshold"], rtol=1e-12)
    np.testing.assert_allclose(g0["split_gain"], c0["split_gain"], rtol=1e-6)      # stored as float32 text
    np.testing.assert_allclose(g0["leaf_value"], c0["leaf_value"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gpu["pred"], cpu["pred"], rtol=2e-3, atol=2e-3)


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(not (os.path.exists(DROPIN) and os.path.exists(REFLIB)), reason="drop-in / reference library not built")
@pytest.mark.parametrize("case,n,f", [("identity", 30000, 96), ("mixed", 20000, 70)])
def test_lgbm_capi_num_gpu_2_matches_single_gpu_and_reference(case, n, f):
    """device_type=cuda, num_gpu=2 through the real LGBM_* API: the adapter shards the feature groups over two library
    learners inside ONE process (LGBMB200_LearnersConnectLocal, one host thread per GPU) — the model must equal the
    single-GPU drop-in's bit for bit and the reference CPU learner's first tree."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    iters = 4
    two = run(DROPIN, "cuda", n, f, iters, case, num_gpu=2)
    one = run(DROPIN, "cuda", n, f, iters, case, num_gpu=1)
    cpu = run(REFLIB, "cpu", n, f, iters, case)
    assert two["num_trees"] == one["num_trees"] == iters
    for a, b in zip(two["trees"], one["trees"]):
        assert a == b                                   # integer histograms: sharding cannot change a single bit
    g0, c0 = two["trees"][0], cpu["trees"][0]
    k = min(8, len(c0["split_feature"]))
    assert g0["split_feature"][:k] == c0["split_feature"][:k]
    np.testing.assert_allclose(two["pred"], cpu["pred"], rtol=2e-3, atol=2e-3)


DROPIN_CUDA = os.path.join(ROOT, "integration", "_build_cuda", "lib_lightgbm.so")


@pytest.mark.skipif(not (os.path.exists(DROPIN_CUDA) and os.path.exists(REFLIB)), reason="device-resident drop-in / reference library not built")
@pytest.mark.parametrize("case,n,f", [("identity", 20000, 12), ("mixed", 20000, 10)])
def test_device_resident_dropin_matches_reference_cpu(case, n, f):
    """The adapter linked against the -DUSE_CUDA build of the reference host code (integration/Makefile `cuda`): the
    reference's own CUDA objective and CUDAScoreUpdater keep gradients and scores on the device and call
    B200TreeLearner::Train / AddPredictionToScore with DEVICE pointers (boosting_on_gpu_, gbdt.cpp:110-135)."""
    iters = 5
    gpu = run(DROPIN_CUDA, "cuda", n, f, iters, case)
    cpu = run(REFLIB, "cpu", n, f, iters, case)
    assert gpu["num_trees"] == cpu["num_trees"] == iters
    g0, c0 = gpu["trees"][0], cpu["trees"][0]
    k = min(8, len(c0["split_feature"]))
    assert g0["split_feature"][:k] == c0["split_feature"][:k]
    np.testing.assert_allclose(g0["threshold"][:k], c0["threshold"][:k], rtol=1e-9)
    np.testing.assert_allclose(gpu["pred"], cpu["pred"], rtol=2e-3, atol=2e-3)
