"""Live comparison helper (TEST INFRASTRUCTURE): one outer iteration of the SVR loop at the C3 geometry on either the
reference's own CUDA path (oracle/_ref/libref_cuda2.so, the unmodified reconstruction_cuda2.cu recompiled for sm_100a)
or libsvr_b200.so, with the stage outputs captured.  The reference constructor calls cudaDeviceReset(), so its arm runs
in a process of its own:

    python tests/ref_live.py gen DATASET.pt 0,2,7 [slices_per_stack]
    python tests/ref_live.py ref DATASET.pt OUT.npz
    python tests/ref_live.py cuda DATASET.pt OUT.npz

Stages follow reconstruction.cc:929-1138 (pipeline.SVRPipeline.outer_iteration) with one super-resolution iteration.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gen(path, stacks, slices_per_stack=None):
    import torch
    from fetalreconstruction_b200.phantom import c3_config, make_dataset
    cfg = c3_config()
    if slices_per_stack:
        cfg.slices_per_stack = int(slices_per_stack)
    parts = [make_dataset(cfg, device="cuda" if torch.cuda.is_available() else "cpu", stacks=[st]) for st in stacks]
    ds = parts[0]
    for name in ("slices", "i2w", "w2i", "trans", "trans_inv", "dims", "stack_index", "true_trans"):
        setattr(ds, name, np.concatenate([getattr(p, name) for p in parts]))
    ds.slice_attrs = sum((p.slice_attrs for p in parts), [])
    ds.stack_attrs = sum((p.stack_attrs for p in parts), [])
    torch.save(ds, path)
    return ds


def run_arm(arm, ds_path, out_path=None):
    import torch
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    ds = torch.load(ds_path, weights_only=False)
    if arm == "ref":
        from oracle.ref_backend import RefReconstruction
        from oracle.ref_runner import ref_pipeline_cls
        b, cls = RefReconstruction(0), ref_pipeline_cls()
    else:
        from fetalreconstruction_b200.reconstruction import Reconstruction
        b, cls = Reconstruction(0), SVRPipeline
    upload_dataset(b, ds)
    p = cls(b, ds.S, 0, ds.S, params=SVRParams(iterations=4, rec_iterations_first=1))
    p.InitializeEMGPU(ds.slices)
    res = {}
    p.set_schedule(0)
    p.InitializeEMValuesGPU()
    res["k1_voxel_num"] = np.asarray(p.GaussianReconstructionGPU())
    res["k1_volume"] = b.syncCPU().copy(); res["k1_psf"] = b.debugv_PSF_sums().copy()
    p.SimulateSlicesGPU()
    res["k2_sim"] = b.debugSimslices().copy(); res["k2_simw"] = b.debugSimweights().copy(); res["k2_inside"] = b.debugSiminside().copy()
    p.InitializeRobustStatisticsGPU()
    res["k11_sigma"] = np.float64(p._sigma)
    p.EStepGPU()
    res["e_weights"] = b.debugWeights().copy(); res["e_potential"] = p.slice_potential.copy(); res["e_slice_weight"] = p._slice_weight.copy()
    p.ScaleGPU()
    res["k10_scale"] = p._scale.copy()
    p.SuperresolutionGPU(1)
    res["k3_addon"] = b.debugAddon().copy(); res["k3_cmap"] = b.debugConfidenceMap().copy(); res["k5_volume"] = b.syncCPU().copy()
    p.SimulateSlicesGPU()
    p.MStepGPU(1)
    res["k9_em"] = np.array([p._sigma, p._mix, p._m], np.float64)
    res["mask"] = ds.mask.ravel() != 0
    if out_path:
        np.savez(out_path, **res)
    return res


def field_stats(ours, ref, flip_threshold):
    """Relative to the RMS of the reference's non-zeros: (rms, max, number of elements further apart than flip_threshold)."""
    a, r = np.asarray(ours, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    fin = np.isfinite(a) & np.isfinite(r)
    nz = r[(r != 0) & np.isfinite(r)]
    sc = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a[fin] - r[fin]) / sc
    return float(np.sqrt(np.mean(d ** 2))), float(d.max()) if d.size else 0.0, int(np.count_nonzero(d > flip_threshold)), int((~fin).sum())


if __name__ == "__main__":
    if sys.argv[1] == "gen":
        gen(sys.argv[2], [int(v) for v in sys.argv[3].split(",")], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        run_arm(sys.argv[1], sys.argv[2], sys.argv[3])
