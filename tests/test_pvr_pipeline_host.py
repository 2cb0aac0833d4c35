"""CPU: the N > 1 path of the PVR host loop (fetalreconstruction_b200/pvr.py::PVRPipeline with a Comm) on the oracle twin,
world_size 2 over gloo: patches sharded (every second patch of every stack per rank), accumulator all-reduce after P1
and P3, partial sums of the robust statistics, gathered per-patch vectors for the patch-level EM."""
import os
import socket

import numpy as np

from fetalreconstruction_b200.pipeline import Comm
from fetalreconstruction_b200.pvr import PVRParams, PVRPipeline
from oracle.oracle_backend_pvr import OraclePatchReconstruction
from pvr_case import make_pvr_case, setup_backend, shard_case

CASE = dict(seed=41, vol=32, n_stacks=2, slices=4, size=32, pbb=(16, 16), stride=(8, 8))
PARAMS = dict(iterations=0, rec_iterations=2)        # one pass: every exchange (P1, P3, EM sums, patch vectors) happens in it


def _patch_cube(case):
    """patch values as the device-side extraction leaves them (masked), so every rank starts from the same numbers"""
    b = setup_backend(OraclePatchReconstruction(), case)
    return b.patches_copyToHost()


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = make_pvr_case(**CASE)
    case["cube"] = _patch_cube(case)
    sub, gidx = shard_case(case, rank, world)
    b = setup_backend(OraclePatchReconstruction(), sub, device_patch_init=False)
    ds = case["ds"]
    p = PVRPipeline(b, ds.min_intensity, ds.max_intensity, PVRParams(**PARAMS), comm=Comm(dist.group.WORLD, "cpu"),
                    global_index=gidx, patches_per_stack_global=case["per_stack"])
    vol = p.run()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vol=vol, em=np.array([p.sigma, p.mix, p.m, p.sigma_s, p.mix_s]),
             pot=p.patch_potential, w=b.rs_get_scales_weights()[1], gidx=gidx)
    dist.destroy_process_group()


def test_pvr_two_rank_gloo_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    case = make_pvr_case(**CASE)
    case["cube"] = _patch_cube(case)
    b = setup_backend(OraclePatchReconstruction(), case, device_patch_init=False)
    ds = case["ds"]
    p = PVRPipeline(b, ds.min_intensity, ds.max_intensity, PVRParams(**PARAMS))
    vol = p.run()
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    np.testing.assert_array_equal(r0["vol"], r1["vol"])                 # every rank ends with the same replica
    np.testing.assert_array_equal(r0["pot"], r1["pot"])
    scale = np.sqrt(np.mean(vol[vol != 0].astype(np.float64) ** 2))
    assert np.abs(r0["vol"] - vol).max() / scale < 1e-4                 # = single rank up to the rounding of the partial sums
    np.testing.assert_allclose(r0["em"], [p.sigma, p.mix, p.m, p.sigma_s, p.mix_s], rtol=1e-4)
    np.testing.assert_allclose(r0["pot"], p.patch_potential, rtol=1e-4, atol=1e-6)
    w = b.rs_get_scales_weights()[1]
    np.testing.assert_allclose(r0["w"], w[r0["gidx"]], atol=1e-4)
    np.testing.assert_allclose(r1["w"], w[r1["gidx"]], atol=1e-4)
    assert len(set(r0["gidx"]) & set(r1["gidx"])) == 0 and len(r0["gidx"]) + len(r1["gidx"]) == sum(case["per_stack"])
