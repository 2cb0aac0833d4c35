#!/usr/bin/env python
"""PVR on N GPUs (one process per GPU, NCCL): patches sharded (every N-th patch of every stack per rank), the volume
accumulator all-reduced after P1 and P3, the robust-statistics sums and per-patch vectors exchanged (pvr.py::PVRPipeline
with a Comm).  Rank 0 then repeats the run on one GPU with all patches and reports the difference and both times.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/pvr_multi_gpu.py OUT.json
Test tooling."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from fetalreconstruction_b200.pipeline import Comm
    from fetalreconstruction_b200.pvr import PatchReconstruction, PVRParams, PVRPipeline
    from pvr_case import make_pvr_case, setup_backend, shard_case
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    comm = Comm(dist.group.WORLD if world > 1 else None, dev)
    vol, slices, size = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (128, 24, 128)
    case = make_pvr_case(seed=43, vol=vol, n_stacks=3, slices=slices, size=size, inplane=1.0, spacing=2.5, pbb=(64, 64), stride=(32, 32))
    ds = case["ds"]
    full = setup_backend(PatchReconstruction(local), case)              # device-side patch extraction (masked values)
    case["cube"] = full.patches_copyToHost()
    params = PVRParams(iterations=1, rec_iterations=3)

    def timed(pipe):
        comm.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        v = pipe.run()
        comm.barrier(); torch.cuda.synchronize()
        return v, time.perf_counter() - t0

    sub, gidx = shard_case(case, rank, world)
    b = setup_backend(PatchReconstruction(local), sub, device_patch_init=False)
    p = PVRPipeline(b, ds.min_intensity, ds.max_intensity, params, comm=comm, global_index=gidx, patches_per_stack_global=case["per_stack"])
    p.run()                                                              # warm-up
    p = PVRPipeline(b, ds.min_intensity, ds.max_intensity, params, comm=comm, global_index=gidx, patches_per_stack_global=case["per_stack"])
    v_multi, t_multi = timed(p)
    # diagnostics: is the accumulator view zero-copy, and do the ranks end with the same replica?
    buf = b.accumulators("psf")[0]
    tview = torch.as_tensor(buf, device=dev)
    chk = torch.tensor([float(np.abs(v_multi.astype(np.float64)).sum())], dtype=torch.float64, device=dev)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("zero-copy view:", tview.data_ptr() == buf.__cuda_array_interface__["data"][0], "replica checksum min/max over ranks:", float(lo), float(hi), flush=True)
    if rank == 0:
        b1 = setup_backend(PatchReconstruction(local), case, device_patch_init=False)
        PVRPipeline(b1, ds.min_intensity, ds.max_intensity, params).run()
        p1 = PVRPipeline(b1, ds.min_intensity, ds.max_intensity, params)
        torch.cuda.synchronize(); t0 = time.perf_counter(); v_one = p1.run(); torch.cuda.synchronize(); t_one = time.perf_counter() - t0
        sc = np.sqrt(np.mean(v_one[v_one != 0].astype(np.float64) ** 2))
        n_patches = len(case["attrs"])
        proj = 2 * (1 + (1 + 3) + 3)                                     # per pass: 1 P1 + 4 P2 + 3 P3; 2 passes (iterations = 1)
        rep = {"gpus": world, "patches": n_patches, "patch_size": [64, 64], "volume": [vol] * 3, "passes": 2, "rec_iterations": 3,
               "seconds": {"n_gpus": t_multi, "one_gpu": t_one}, "patch_projections_per_s": {"n_gpus": n_patches * proj / t_multi, "one_gpu": n_patches * proj / t_one},
               "volume_rel_max_diff_n_vs_one": float(np.abs(v_multi - v_one).max() / sc),
               "em_n": [p.sigma, p.mix, p.m], "em_one": [p1.sigma, p1.mix, p1.m]}
        json.dump(rep, open(sys.argv[1], "w"), indent=1)
        print(json.dumps(rep))
    if world > 1:
        comm.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
