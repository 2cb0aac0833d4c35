#!/usr/bin/env python
"""SVR on N GPUs (one process per GPU, NCCL) against the same run on one GPU: slices sharded as in bench.py (rank r takes
every N-th slice of every stack), accumulator all-reduce after K1 / K3, statistics and per-slice vectors exchanged
(pipeline.py::SVRPipeline with a Comm).  Rank 0 repeats the run on one GPU with all slices (in the same rank-major slice
order) and reports the difference.  Test tooling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/svr_multi_gpu.py OUT.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from fetalreconstruction_b200.phantom import make_dataset, shard_count, small_config
    from fetalreconstruction_b200.pipeline import Comm, SVRParams, SVRPipeline, upload_dataset
    from fetalreconstruction_b200.reconstruction import Reconstruction
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm(dist.group.WORLD if world > 1 else None, dev)
    cfg = small_config(seed=7, vol=96, n_stacks=8, slices=24, size=96, inplane=1.0, spacing=2.0)
    params = SVRParams(iterations=2, rec_iterations_first=3, rec_iterations_last=4)
    per_rank = [cfg.n_stacks * shard_count(cfg.slices_per_stack, r, world) for r in range(world)]
    b0 = sum(per_rank[:rank]); e0 = b0 + per_rank[rank]
    S = cfg.n_stacks * cfg.slices_per_stack
    ds = make_dataset(cfg, device=str(dev), shard=(rank, world))
    b = Reconstruction(local)
    upload_dataset(b, ds)
    acc = torch.as_tensor(b.accumulator(), device=dev)
    p = SVRPipeline(b, S, b0, e0, comm, params, accumulator_tensor=lambda: acc)
    p.InitializeEMGPU(ds.slices)
    vol = p.run()
    chk = torch.tensor([float(np.abs(vol.astype(np.float64)).sum())], dtype=torch.float64, device=dev)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        # the same slices in the same (rank-major) order on one GPU
        parts = [make_dataset(cfg, device=str(dev), shard=(r, world)) for r in range(world)]
        full = parts[0]
        for name in ("slices", "i2w", "w2i", "trans", "trans_inv", "dims", "stack_index"):
            setattr(full, name, np.concatenate([getattr(q, name) for q in parts]))
        b1 = Reconstruction(local)
        upload_dataset(b1, full)
        p1 = SVRPipeline(b1, S, 0, S, params=params)
        p1.InitializeEMGPU(full.slices)
        v1 = p1.run()
        m = full.mask.ravel() != 0
        sc = np.sqrt(np.mean(v1[m].astype(np.float64) ** 2))
        rep = {"gpus": world, "slices": S, "volume": list(cfg.vol_size), "replica_checksum_min_max": [float(lo), float(hi)],
               "volume_rel_max_diff_n_vs_one": float(np.abs(vol[m] - v1[m]).max() / sc),
               "volume_rel_rms_diff_n_vs_one": float(np.sqrt(np.mean((vol[m].astype(np.float64) - v1[m]) ** 2)) / sc),
               "slice_weights_max_diff": float(np.abs(p._slice_weight - p1._slice_weight).max()),
               "em_n": [p._sigma, p._mix, p._m], "em_one": [p1._sigma, p1._mix, p1._m]}
        json.dump(rep, open(sys.argv[1], "w"), indent=1)
        print(json.dumps(rep))
    if world > 1:
        comm.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
