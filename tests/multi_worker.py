"""Worker for the multi-GPU tests / CPU gloo test.  Launched once per rank with RANK/WORLD_SIZE/MASTER_* set."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    mode = sys.argv[1]
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = "gloo" if mode == "host" else "nccl"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group(backend, rank=rank, world_size=world)
    import lightgbm_b200 as lgb
    from lightgbm_b200 import distributed as D
    from helpers import synth_identity

    if mode == "host":
        # host-side logic only (no GPU): shard plan, handle gather, feature offsets
        plan = D.shard_columns(int(sys.argv[2]), world)
        handles = D.gather_bytes(bytes([rank]) * 64, rank, world)
        counts = [int(x) for x in D.gather_bytes(str(plan[rank][1] - plan[rank][0]).encode(), rank, world)]
        off = D.feature_offsets(counts)
        print("JSON" + json.dumps(dict(rank=rank, plan=plan, handles=[h[0] for h in handles], off=off.tolist())))
        dist.barrier()
        dist.destroy_process_group()
        return

    if mode.startswith("golden"):
        # a fixture read out of a real reference Dataset: EFB bundles, most-freq-bin elision, missing types —
        # the sharded learners must reproduce the reference's own tree (mode goldenpush: no column replication)
        import golden_io
        gd = golden_io.Golden(sys.argv[2])
        full = lgb.Layout.from_attrs(gd.layout)
        n, f = full.num_data, full.num_columns
        g, h = gd.grad.astype(np.float32), gd.hess.astype(np.float32)
        cfg = lgb.Config(**gd.params, gpu_device_id=int(os.environ.get("LOCAL_RANK", rank)))
        lo, hi = D.shard_columns(f, world)[rank]
        shard = full.column_slice(lo, hi) if hi > lo else D.empty_shard(n, rank)
        L = D.make_sharded_learner(shard, cfg, rank, world, replicate_columns=(mode == "golden"))
        t = L.train(g, h)
        ok = golden_io.check_against_reference(t, gd, exact_values=False)
        print("JSON" + json.dumps(dict(rank=rank, splits_checked=int(ok), num_leaves=int(t.num_leaves),
                                       feature=t.splits["feature"].tolist(), threshold=t.splits["threshold"].tolist())))
        dist.barrier()
        dist.destroy_process_group()
        return

    n, f, leaves = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    bins, y, g, h = synth_identity(n, f, seed=99)
    full = lgb.Layout.identity(bins)
    lo, hi = D.shard_columns(f, world)[rank]
    shard = full.column_slice(lo, hi) if hi > lo else D.empty_shard(n, rank)
    cfg = lgb.Config(num_leaves=leaves, gpu_device_id=int(os.environ.get("LOCAL_RANK", rank)),
                     use_cuda_graph=not os.environ.get("NO_GRAPH"),
                     use_quantized_grad=mode.startswith("gpuquant"), num_grad_quant_bins=4, stochastic_rounding=False)
    if mode == "rows":
        r0, r1 = D.shard_rows(n, world)[rank]
        shard = lgb.Layout.identity(bins[r0:r1])
        L = D.make_row_sharded_learner(shard, cfg, rank, world)
        gl, hl = g[r0:r1], h[r0:r1]
    else:
        L = D.make_sharded_learner(shard, cfg, rank, world, replicate_columns=(mode not in ("gpupush", "gpuquantpush")))
        gl, hl = g, h
    trees = []
    for it in range(3):
        t = L.train(gl * (1 + 0.1 * it), hl)
        trees.append(t)
    lb, lc, idx = L.get_partition(trees[-1].num_leaves)
    if mode == "rows":
        # local partition: every local row in exactly one leaf; report per-leaf local counts for a global check
        assert sorted(idx[idx >= 0].tolist()) == list(range(len(gl)))
    out = dict(rank=rank, trees=[dict(n=t.num_leaves, feature=t.splits["feature"].tolist(), leaf=t.splits["leaf"].tolist(),
                                      threshold=t.splits["threshold"].tolist(), gain=t.splits["gain"].tolist(),
                                      leaf_value=t.leaf_value.tolist(), leaf_count=t.leaf_count.tolist()) for t in trees],
               part_hash=int(np.bitwise_xor.reduce(idx.astype(np.int64) * (np.arange(len(idx)) + 1))),
               local_leaf_count=lc.tolist())
    if rank == 0:
        # single-GPU learner on the full matrix as the reference for the sharded result
        S = lgb.B200TreeLearner(cfg)
        S.init(full)
        single = [S.train(g * (1 + 0.1 * it), h) for it in range(3)]
        out["single"] = [dict(n=t.num_leaves, feature=t.splits["feature"].tolist(), leaf=t.splits["leaf"].tolist(),
                              threshold=t.splits["threshold"].tolist(), gain=t.splits["gain"].tolist(),
                              leaf_value=t.leaf_value.tolist(), leaf_count=t.leaf_count.tolist()) for t in single]
    print("JSON" + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
