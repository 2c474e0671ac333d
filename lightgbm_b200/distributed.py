"""Feature-shard plumbing (one process per GPU).  torch.distributed is used ONLY for rendezvous: to all-gather
the 64-byte CUDA-IPC handles and the per-rank feature counts, and (in bench.py) for the barrier / max-reduce of
the timing.  All per-split traffic goes through NVLink peer memory inside the kernels (csrc/scan_kernel.cuh
k_select, csrc/partition_kernel.cuh k_part_flags / k_part_count).

Sharding rule (mirrors FeatureParallelTreeLearner's "every worker sees all rows", reference
src/treelearner/feature_parallel_tree_learner.cpp:37-78, but balanced by column groups instead of #bins):
columns are dealt to ranks in contiguous runs that are multiples of 32 (one histogram-kernel column group).
"""
from __future__ import annotations

import numpy as np

from .tree_learner import B200TreeLearner, Config, Layout

COLGROUP = 32


def shard_columns(num_columns: int, world: int) -> list[tuple[int, int]]:
    """[col_lo, col_hi) per rank.  Runs are multiples of 32 columns (one histogram column group) when there are at
    least `world` groups, else single columns are dealt out; ranks beyond the column count get an empty run."""
    groups = (num_columns + COLGROUP - 1) // COLGROUP
    unit = COLGROUP if groups >= world else 1
    units = (num_columns + unit - 1) // unit
    base, extra = divmod(units, world)
    out, g = [], 0
    for r in range(world):
        take = base + (1 if r < extra else 0)
        lo = min(num_columns, g * unit)
        hi = min(num_columns, (g + take) * unit)
        out.append((lo, hi))
        g += take
    return out


def empty_shard(num_data: int, rank: int) -> Layout:
    """A rank with no real column still takes part in every exchange: give it one constant column whose single
    2-bin feature can never split (all rows in bin 0)."""
    z = np.zeros(1, np.int32)
    return Layout(np.zeros((num_data, 1), np.uint8), z.copy(), np.ones(1, np.int32), np.full(1, 2, np.int32), z.copy(),
                  z.copy(), z.copy(), np.full(1, (1 << 30) + rank, np.int32))


def feature_offsets(feature_counts: list[int]) -> np.ndarray:
    off = np.zeros(len(feature_counts) + 1, np.int32)
    off[1:] = np.cumsum(feature_counts)
    return off


def gather_bytes(payload: bytes, rank: int, world: int) -> list[bytes]:
    """all-gather of a small byte string through torch.distributed (any backend)."""
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, payload)
    return out


def make_sharded_learner(shard_layout: Layout, config: Config, rank: int, world: int, gather=gather_bytes,
                         replicate_columns: bool = True, is_constant_hessian: bool = False) -> B200TreeLearner:
    """shard_layout: this rank's column slice (Layout.column_slice of the full layout, or generated directly).
    Returns a learner whose Train() grows the same tree on every rank, with global inner-feature ids.
    replicate_columns: also hold a column-major copy of EVERY rank's columns (total_columns x num_data bytes per
    GPU, filled over NVLink) so that the row partition of every split is computed locally — the reference's
    feature-parallel trade (every worker holds the full data, docs/Features.rst:109-125).  False keeps one copy of
    the matrix across the box; the split's owner then pushes the go-left bits to its peers."""
    L = B200TreeLearner(config)
    L.init(shard_layout, is_constant_hessian=is_constant_hessian)
    if world > 1:
        handle = L.comm_export()
        handles = gather(handle, rank, world)
        counts = [int(x) for x in gather(str(shard_layout.num_features).encode(), rank, world)]
        L.comm_connect(rank, world, b"".join(handles), feature_offsets(counts))
        if replicate_columns:
            cols = gather(L.comm_export_columns(), rank, world)
            L.comm_share_columns(b"".join(cols))
            gather(b"done", rank, world)      # barrier: nobody re-Inits while a peer still copies out of its buffer
    return L


def shard_rows(num_data: int, world: int) -> list[tuple[int, int]]:
    """[row_lo, row_hi) per rank: contiguous blocks of ceil(N/world) rows, the reference's own rule
    (src/boosting/cuda/nccl_gbdt_component.hpp:32-44)."""
    per = (num_data + world - 1) // world
    return [(min(num_data, r * per), min(num_data, (r + 1) * per)) for r in range(world)]


def make_row_sharded_learner(shard_layout: Layout, config: Config, rank: int, world: int, gather=gather_bytes) -> B200TreeLearner:
    """shard_layout: this rank's ROW slice (all columns).  Train() takes the gradients of those rows only and
    grows the same tree on every rank; leaf_count / split counts are global."""
    L = B200TreeLearner(config)
    L.init(shard_layout, is_constant_hessian=False)
    if world > 1:
        comm = gather(L.comm_export(), rank, world)
        pool = gather(L.comm_export_pool(), rank, world)
        L.comm_connect_rows(rank, world, b"".join(comm), b"".join(pool))
    return L
