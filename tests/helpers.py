"""Shared helpers for the parity tests: synthetic binned data + tree comparison with the near-tie rule
of SURVEY.md §8d (if two candidate splits are closer than the stated tolerance either choice is valid and
the comparison stops at the divergence point)."""
from __future__ import annotations

import numpy as np


def synth_identity(n, f, seed, informative=8, num_bin=255):
    """bins drawn uniformly; stored value == bin; L2 gradients at score 0."""
    rng = np.random.default_rng(seed)
    bins = rng.integers(0, num_bin, (n, f), dtype=np.uint8)
    k = min(informative, f)
    w = rng.normal(size=k)
    y = ((bins[:, :k] / (num_bin / 2.0) - 1) @ w + 0.5 * rng.normal(size=n)).astype(np.float32)
    grad = (0.0 - y).astype(np.float32)
    hess = np.ones(n, np.float32)
    return bins, y, grad, hess


def logistic_grad(y01, score=0.0):
    p = 1.0 / (1.0 + np.exp(-score))
    g = (p - y01).astype(np.float32)
    h = np.full(len(y01), p * (1 - p), np.float32) if np.isscalar(p) else (p * (1 - p)).astype(np.float32)
    return g, h


def compare_trees(gpu, orc, rtol=1e-5, min_prefix=None):
    """gpu: lightgbm_b200.Tree, orc: oracle_py.OracleTree.  Returns (#splits that matched exactly, diverged?).
    Structural fields must match exactly; float fields within rtol.  A divergence is tolerated only when it
    is a near tie (the two gains agree within rtol), per the reference's own CPU<->GPU tolerance
    (tests/python_package_test/test_dual.py:35-36)."""
    ns = min(gpu.num_leaves, orc.num_leaves) - 1
    matched = 0
    for i in range(ns):
        a, b = gpu.splits[i], orc.splits[i]
        same = (a["leaf"] == b["leaf"] and a["feature"] == b["feature"] and a["threshold"] == b["threshold"]
                and a["default_left"] == b["default_left"])
        if not same:
            rel = abs(a["gain"] - b["gain"]) / max(abs(b["gain"]), 1e-300)
            assert rel < rtol * 10, f"split {i}: structural mismatch that is not a near tie: gpu={a} oracle={b}"
            return matched, True
        assert a["left_count"] == b["left_count"] and a["right_count"] == b["right_count"], f"split {i} counts {a} {b}"
        for k in ("gain", "left_sum_gradient", "left_sum_hessian", "right_sum_gradient", "right_sum_hessian",
                  "left_output", "right_output"):
            scale = max(abs(b[k]), 1e-12)
            if k.endswith("gradient"):
                # gradient sums can cancel: compare against the magnitude of the leaf's hessian-weighted scale
                scale = max(scale, 1e-6 * max(abs(b["left_sum_hessian"]), abs(b["right_sum_hessian"])))
            assert abs(a[k] - b[k]) <= rtol * scale * 10 + 1e-9, f"split {i} field {k}: gpu={a[k]!r} oracle={b[k]!r}"
        matched += 1
    if min_prefix is not None:
        assert matched >= min_prefix
    assert gpu.num_leaves == orc.num_leaves, f"num_leaves gpu={gpu.num_leaves} oracle={orc.num_leaves}"
    np.testing.assert_array_equal(gpu.leaf_count, orc.leaf_count)
    np.testing.assert_allclose(gpu.leaf_value, orc.leaf_value, rtol=rtol * 10, atol=1e-9)
    return matched, False
