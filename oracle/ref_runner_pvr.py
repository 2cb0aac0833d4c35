#!/usr/bin/env python
"""Runs the reference's own PVR CUDA path (oracle/_ref/libref_pvr.so, oracle/ref_backend_pvr.py) on the seeded PVR case
of tests/golden/make_golden.py stage by stage and writes its outputs (TEST INFRASTRUCTURE ONLY; GPU box):

    python -m oracle.ref_runner_pvr --out DIR          ->  DIR/ref_pvr_small.npz (committed under tests/golden/)
The patch list and patch values come from our own enumeration / extraction (oracle twin), see ref_backend_pvr.py."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF_PVR_CASE = dict(seed=41, vol=32, n_stacks=2, slices=4, size=32, pbb=(16, 16), stride=(8, 8))   # volume size a multiple of 8


def pvr_stages(backend, pipeline_cls, patch_cube=None, rec_iterations=2, case=None):
    """One PVR iteration (irtkPatchBasedReconstruction.cpp:482-560) with every stage output captured."""
    from fetalreconstruction_b200.pvr import PVRParams
    from pvr_case import make_pvr_case, setup_backend
    case = case or make_pvr_case(**REF_PVR_CASE)
    ds = case["ds"]
    if patch_cube is None:
        b = setup_backend(backend, case)
    else:
        case2 = dict(case); case2["cube"] = patch_cube
        b = setup_backend(backend, case2, device_patch_init=False)
    p = pipeline_cls(b, ds.min_intensity, ds.max_intensity, PVRParams(iterations=0, rec_iterations=rec_iterations))
    out = dict(per_stack=np.array(case["per_stack"], np.int32), patches=b.patches_copyToHost().astype(np.float32))
    b.rs_initializeEMValues()
    b.recon_reset()
    b.patchBasedPSFReconstruction_gpu()
    out["p1_recon_raw"] = b.recon_copyToHost(); out["p1_volw"] = b.getVolWeights(); out["p1_psf_sums"] = b.debugPSFsums()
    b.recon_equalize()
    out["p1_recon"] = b.recon_copyToHost()
    b.patchBasedSimulatePatches_gpu()
    out["p2_sim"] = b.debugSimpatches(); out["p2_simw"] = b.debugSimweights(); out["p2_inside"] = b.debugSiminside()
    p.InitializeRobustStatistics()
    out["rs_init"] = np.array([p.sigma, p.mix, p.m], np.float32)
    p.EStep()
    out["e0_weights"] = b.debugWeights()
    out["e0_patch_scale"], out["e0_patch_weight"] = b.rs_get_scales_weights()
    out["e0_state"] = np.array([p.sigma_s, p.mix_s], np.float32)
    for i in range(rec_iterations):
        p.Scale()
        out[f"r{i}_patch_scale"] = b.rs_get_scales_weights()[0]
        b.recon_resetAddonCmap()
        b.superresolution_run()
        out[f"r{i}_addon"] = b.debugAddon(); out[f"r{i}_cmap"] = b.debugConfidenceMap()
        b.superresolution_regularize(p.p.adaptive, p.alpha, p.min_intensity, p.max_intensity, p.p.delta, p.p.lambda_)
        out[f"r{i}_recon"] = b.recon_copyToHost()
        b.patchBasedSimulatePatches_gpu()
        out[f"r{i}_sim"] = b.debugSimpatches()
        p.MStep(i + 1)
        out[f"r{i}_mstep"] = np.array([p.sigma, p.mix, p.m], np.float32)
        p.EStep()
        out[f"r{i}_weights"] = b.debugWeights()
        out[f"r{i}_patch_weight"] = b.rs_get_scales_weights()[1]
    out["volume"] = b.recon_copyToHost()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    from fetalreconstruction_b200.pvr import PVRPipeline
    from oracle.oracle_backend_pvr import OraclePatchReconstruction
    from oracle.ref_backend_pvr import RefPatchReconstruction, ref_pvr_pipeline_cls
    orc = pvr_stages(OraclePatchReconstruction(), PVRPipeline)
    ref = pvr_stages(RefPatchReconstruction(0), ref_pvr_pipeline_cls(), patch_cube=orc["patches"])
    np.savez_compressed(os.path.join(a.out, "ref_pvr_small.npz"), **ref)
    print("wrote", os.path.join(a.out, "ref_pvr_small.npz"))
    arms = {"oracle": orc}
    try:
        import torch
        if torch.cuda.is_available():
            from fetalreconstruction_b200.pvr import PatchReconstruction
            arms["cuda"] = pvr_stages(PatchReconstruction(0), PVRPipeline)
    except Exception as e:                                    # report what we have
        print("cuda arm failed:", e)
    for k in ref:
        if k in ("per_stack",):
            continue
        line = f"{k:18s}"
        for name, arm in arms.items():
            r, x = np.asarray(ref[k], np.float64).ravel(), np.asarray(arm[k], np.float64).ravel()
            nz = r[(r != 0) & np.isfinite(r)]
            sc = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
            fin = np.isfinite(r) & np.isfinite(x)
            d = np.abs(r[fin] - x[fin]) / sc
            line += f"  {name}: rms {np.sqrt(np.mean(d ** 2)) if d.size else 0:.2e} max {d.max() if d.size else 0:.2e} nonfinite {int((~fin).sum())}"
        print(line)


if __name__ == "__main__":
    main()
