"""CPU: the host orchestration (pipeline.py) run on the oracle backend -- single rank, and world_size 2
over gloo (the N > 1 path: stack sharding, accumulator all-reduce, statistics all-reduce)."""
import os
import socket
import sys

import numpy as np
import pytest

from fetalreconstruction_b200.phantom import make_dataset, small_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, Comm, upload_dataset
from fetalreconstruction_b200.reconstruction import host_partition
from oracle.oracle_backend import OracleReconstruction

CFG = dict(seed=11, vol=24, n_stacks=2, slices=5, size=20, inplane=1.2, spacing=2.5)
PARAMS = dict(iterations=2, rec_iterations_first=2, rec_iterations_last=3)


def _single():
    ds = make_dataset(small_config(**CFG))
    b = OracleReconstruction()
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(**PARAMS))
    p.InitializeEMGPU(ds.slices)
    vol = p.run()
    return ds, p, vol


def test_smoothing_schedule_matches_reference_defaults():
    """iterations=4, levels=3: lambda 0.08, 0.04, 0.02 then lastIterLambda 0.01 (reconstruction.cc:900-911);
    _lambda := lambda*delta^2, alpha := min(1, 0.05/lambda) (irtkReconstructionGPU.h:605-612)."""
    p = SVRPipeline(OracleReconstruction(), 1, 0, 1, params=SVRParams())
    seen = []
    for it in range(4):
        p.set_schedule(it)
        seen.append((p._lambda / 150.0 ** 2, p._alpha))
    lam = [s[0] for s in seen]
    np.testing.assert_allclose(lam, [0.08, 0.04, 0.02, 0.01], rtol=1e-12)
    np.testing.assert_allclose([s[1] for s in seen], [0.625, 1.0, 1.0, 1.0], rtol=1e-12)


def test_single_rank_pipeline_runs_and_downweights_corrupted_slices():
    cfg = small_config()                                 # 40^3 volume: enough fully-covered pixels for the statistics
    cfg.corrupt_fraction = 0.0
    ds = make_dataset(cfg)
    ds.slices[14][ds.slices[14] > 0] *= 0.3              # one badly corrupted slice
    b = OracleReconstruction()
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(iterations=1, rec_iterations_last=3, intensity_matching=False))
    p.InitializeEMGPU(ds.slices)
    vol = p.run()
    assert np.isfinite(vol).all()
    assert np.all(vol[ds.mask.ravel() == 0] == -1)
    w = p._slice_weight
    assert np.all((w >= 0) & (w <= 1))
    assert w[14] < 0.05 and np.median(np.delete(w, 14)) > 0.95


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = small_config(**CFG)
    full = make_dataset(cfg)
    b0, e0 = host_partition([cfg.slices_per_stack] * cfg.n_stacks, world, rank)
    shard = make_dataset(cfg, stacks=range(b0 // cfg.slices_per_stack, e0 // cfg.slices_per_stack))
    np.testing.assert_array_equal(shard.slices, full.slices[b0:e0])     # sharded generation == slicing
    backend = OracleReconstruction()
    upload_dataset(backend, shard)
    comm = Comm(dist.group.WORLD, "cpu")
    p = SVRPipeline(backend, full.S, b0, e0, comm, SVRParams(**PARAMS),
                    accumulator_tensor=lambda: torch.from_numpy(backend.accumulator()))
    p.InitializeEMGPU(shard.slices)
    vol = p.run()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vol=vol, w=p._slice_weight, s=p._scale,
             stats=np.array([p._sigma, p._mix, p._m]))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ds, p, vol = _single()
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # every rank ends with the same replica
    np.testing.assert_array_equal(r0["vol"], r1["vol"])
    np.testing.assert_array_equal(r0["w"], r1["w"])
    # and it is the single-rank answer up to the float32 rounding of the per-rank partial sums
    m = ds.mask.ravel() != 0
    assert np.array_equal(r0["vol"] == -1, vol == -1)
    scale = np.sqrt(np.mean(vol[m].astype(np.float64) ** 2))
    assert np.abs(r0["vol"][m] - vol[m]).max() / scale < 1e-4
    np.testing.assert_allclose(r0["w"], p._slice_weight, atol=1e-4)
    np.testing.assert_allclose(r0["s"], p._scale, rtol=1e-5)
    np.testing.assert_allclose(r0["stats"], [p._sigma, p._mix, p._m], rtol=1e-4)
