"""Multi-process paths.  CPU: world_size-2 gloo test of the host-side sharding / bootstrap logic.
GPU (needs >= 2 GPUs): the feature-sharded learner grows the same trees on every rank and the same trees as a
single-GPU learner on the full matrix."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(world, args, timeout=600):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "multi_worker.py")] + [str(a) for a in args],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, e[-3000:]
        outs.append(json.loads([l for l in o.splitlines() if l.startswith("JSON")][-1][4:]))
    return sorted(outs, key=lambda d: d["rank"])


def test_host_side_sharding_gloo_world2():
    outs = launch(2, ["host", 100])
    assert outs[0]["plan"] == outs[1]["plan"] == [[0, 64], [64, 100]]
    assert outs[0]["handles"] == outs[1]["handles"] == [0, 1]
    assert outs[0]["off"] == outs[1]["off"] == [0, 64, 100]


def test_shard_columns_properties():
    from lightgbm_b200.distributed import shard_columns
    for c in (1, 5, 28, 32, 33, 100, 256, 1024, 1000):
        for w in (1, 2, 3, 4, 8):
            plan = shard_columns(c, w)
            assert plan[0][0] == 0 and plan[-1][1] == c
            assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
            if (c + 31) // 32 >= w:
                assert all(lo % 32 == 0 for lo, _ in plan)


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["gpu", "gpupush", "gpuquant", "gpuquantpush"])   # replicated columns / owner pushes the go-left bits; + quantized gradients
@pytest.mark.parametrize("n,f,leaves", [(30000, 96, 31), (20000, 40, 15)])
def test_feature_shard_world2_matches_single_gpu(n, f, leaves, mode):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = launch(2, [mode, n, f, leaves])
    a, b = outs
    for ta, tb in zip(a["trees"], b["trees"]):
        assert ta == tb                      # every rank grows the identical tree, bit for bit
    assert a["part_hash"] == b["part_hash"]
    for ts, tm in zip(a["single"], a["trees"]):
        assert ts["n"] == tm["n"] and ts["feature"] == tm["feature"] and ts["leaf"] == tm["leaf"]
        assert ts["threshold"] == tm["threshold"] and ts["leaf_count"] == tm["leaf_count"]
        np.testing.assert_allclose(ts["gain"], tm["gain"], rtol=1e-5)
        np.testing.assert_allclose(ts["leaf_value"], tm["leaf_value"], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["golden", "goldenpush"])
@pytest.mark.parametrize("name", ["efb_bundled", "mixed_zero_as_missing", "mixed_missing_binary"])
def test_feature_shard_world2_reproduces_reference_tree(name, mode):
    """Column-sharded over 2 GPUs, a fixture with EFB bundles / elided most-frequent bins / missing types must still
    give the tree the unmodified reference grew (tests/golden/*.npz)."""
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    a, b = launch(2, [mode, name])
    assert a["feature"] == b["feature"] and a["threshold"] == b["threshold"] and a["num_leaves"] == b["num_leaves"]
    assert a["splits_checked"] >= min(3, a["num_leaves"] - 1)     # same bar as the single-GPU golden test (near ties may diverge later)


@pytest.mark.gpu
@pytest.mark.parametrize("n,f,leaves", [(30000, 28, 31), (20001, 40, 15)])
def test_row_shard_world2_matches_single_gpu(n, f, leaves):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = launch(2, ["rows", n, f, leaves])
    a, b = outs
    for ta, tb in zip(a["trees"], b["trees"]):
        assert ta == tb                      # identical decisions and identical (global) values on every rank
    # local leaf counts add up to the global leaf counts of the last tree
    glob = a["trees"][-1]["leaf_count"]
    assert [x + y for x, y in zip(a["local_leaf_count"], b["local_leaf_count"])] == glob
    for ts, tm in zip(a["single"], a["trees"]):
        assert ts["n"] == tm["n"] and ts["feature"] == tm["feature"] and ts["leaf"] == tm["leaf"]
        assert ts["threshold"] == tm["threshold"] and ts["leaf_count"] == tm["leaf_count"]
        np.testing.assert_allclose(ts["gain"], tm["gain"], rtol=1e-5)
        np.testing.assert_allclose(ts["leaf_value"], tm["leaf_value"], rtol=1e-5, atol=1e-9)


def test_shard_rows_contract():
    from lightgbm_b200.distributed import shard_rows
    assert shard_rows(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert shard_rows(11_000_000, 8)[-1] == (9_625_000, 11_000_000)
