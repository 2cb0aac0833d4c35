#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/*.c) on seeded phantoms.

STATUS: the reference (bkainz/fetalReconstruction) ships no golden vectors, tests or expected outputs for this
path and cannot be built here, so these vectors pin the ORACLE against regressions and give
the CUDA path a fixed target; the vectors that pin both against the reference itself are ref_*.npz (written by
oracle/ref_runner.py / ref_runner_pvr.py from the reference's own CUDA code on a B200; tests/test_ref_golden.py).

    python tests/golden/make_golden.py        # rewrites the fixtures (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


# The reference kernels have no x/y bounds check: grid-overhang threads of its (8,8,z) blocks alias
# into the next row / slice.  Cases that are compared with the REFERENCE's own output therefore use slice sizes
# that are multiples of 8, where the reference has no overhang threads.
REF_SVR_SIZE = (24, 24)
REF_STEPS_SIZE = (40, 32)
# The reference's volume kernels (AdaptiveRegularizationPrep / AdaptiveRegularizationKernel, cuda2.cu:1944-2117) have no
# x/y bounds check either and test `pos.z > size.z`: for a volume whose dimensions are not multiples of the (8,8,8) block,
# overhang threads alias onto x,y < 4 of the next row / plane and apply the update twice.  Reference-compared cases use a
# volume size that is a multiple of 8.
REF_SVR_VOL = 32


def svr_case(backend=None, pipeline_cls=None, slice_size=None):
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    if backend is None:
        from oracle.oracle_backend import OracleReconstruction
        backend = OracleReconstruction()
    # the reference-generated variant: Nx, Ny and the volume size multiples of 8 (see REF_SVR_SIZE, REF_SVR_VOL)
    cfg = small_config(seed=21, vol=28 if slice_size is None else REF_SVR_VOL, n_stacks=2, slices=5, size=24, inplane=1.1,
                       spacing=2.2)
    if slice_size is not None:
        cfg.slice_size = tuple(slice_size)
    ds = make_dataset(cfg)
    b = backend
    upload_dataset(b, ds)
    p = (pipeline_cls or SVRPipeline)(b, ds.S, 0, ds.S, params=SVRParams(iterations=1, rec_iterations_last=2))
    p.InitializeEMGPU(ds.slices)
    p.set_schedule(0)
    p.InitializeEMValuesGPU()
    voxel_num = p.GaussianReconstructionGPU()
    out = dict(config=np.array([21, 28, 2, 5, 24], np.int32), inplane=np.float32(1.1), spacing=np.float32(2.2),
               voxel_num=voxel_num, psf_sums=b.debugv_PSF_sums().astype(np.float16),
               gaussian_recon=b.syncCPU().astype(np.float32), volweights=b.getVolWeights().astype(np.float32))
    p.SimulateSlicesGPU()
    out["simslices"] = b.debugSimslices().astype(np.float32)
    out["siminside"] = b.debugSiminside().astype(np.int8)
    p.InitializeRobustStatisticsGPU()
    p.EStepGPU()
    out["sigma"] = np.float32(p._sigma)
    out["slice_potential"] = p.slice_potential.astype(np.float32)
    for i in range(2):
        p.reconstruction_iteration(i)
    p.MaskVolumeGPU()
    out["volume"] = b.syncCPU().astype(np.float32)
    out["scale"] = p._scale.astype(np.float32)
    out["slice_weight"] = p._slice_weight.astype(np.float32)
    out["em"] = np.array([p._sigma, p._mix, p._m], np.float32)
    return out


def reg_case(backend=None):
    from fetalreconstruction_b200.geometry import rigid_matrix
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    from fetalreconstruction_b200.registration import RegistrationFrontEnd
    oracle_backend = backend is None
    if backend is None:
        from oracle.oracle_backend import OracleReconstruction
        backend = OracleReconstruction()
    cfg = small_config(seed=31, vol=30, n_stacks=2, slices=4, size=26, inplane=1.2, spacing=2.0)
    cfg.noise = 2.0
    cfg.corrupt_fraction = 0.0
    ds = make_dataset(cfg)
    b = backend
    vx, vy, vz = cfg.vol_size
    vol = np.where(ds.mask > 0, ds.truth, -1.0).astype(np.float32)
    b.InitReconstructionVolume((vx, vy, vz), (cfg.vol_voxel,) * 3, vol.ravel())
    if oracle_backend:
        b.recon_w2i = ds.recon_w2i
    else:
        b.setMask((vx, vy, vz), (cfg.vol_voxel,) * 3, ds.mask.ravel())
        b.initStorageVolumes((ds.slices.shape[2], ds.slices.shape[1], ds.S))
        b.setSliceDims(ds.dims)
        b.SetSliceMatrices(ds.trans, ds.trans_inv, ds.i2w, ds.w2i, ds.i2w, ds.w2i, ds.recon_i2w, ds.recon_w2i)
    fe = RegistrationFrontEnd(b, ds.slices, ds.slice_attrs, cfg.vol_voxel)
    rng = np.random.default_rng(31)
    pert = np.stack([(ds.true_trans[k].reshape(4, 4).astype(np.float64)
                      @ rigid_matrix(*rng.normal(0, 0.6, 3), *rng.normal(0, 0.6, 3))).ravel() for k in range(ds.S)])
    b.updateResampledSlicesI2W(fe.ofs)
    b.prepareSliceToVolumeReg()
    t0 = fe.pack_transforms(pert)
    out = dict(config=np.array([31, 30, 2, 4, 26], np.int32), inplane=np.float32(1.2), spacing=np.float32(2.0),
               transforms_in=t0, sim_level0=b.evaluateCostsMultipleSlices(t0, 0), sim_level1=b.evaluateCostsMultipleSlices(t0, 1),
               resampled_checksum=np.float64(fe.cube.astype(np.float64).sum()))
    b.setRegSchedule(2, 2, 4)
    out["transforms_out"] = b.registerSlicesToVolume(t0)
    out["evaluations"] = np.int64(b.reg_evaluations)
    return out


def pvr_case(backend=None):
    from fetalreconstruction_b200.pvr import PVRParams, PVRPipeline
    from pvr_case import make_pvr_case, setup_backend
    if backend is None:
        from oracle.oracle_backend_pvr import OraclePatchReconstruction
        backend = OraclePatchReconstruction()
    case = make_pvr_case(seed=41, vol=32, n_stacks=2, slices=4, size=32, pbb=(16, 16), stride=(8, 8))
    ds = case["ds"]
    b = setup_backend(backend, case)
    pipe = PVRPipeline(b, ds.min_intensity, ds.max_intensity, PVRParams(iterations=0, rec_iterations=2))
    vol = pipe.run()
    return dict(config=np.array([41, 32, 2, 4, 32, 16, 8], np.int32), per_stack=np.array(case["per_stack"], np.int32),
                patches_checksum=np.float64(b.patches_copyToHost().astype(np.float64).sum()), volume=vol.astype(np.float32),
                psf_sums=b.debugPSFsums().astype(np.float16), em=np.array([pipe.sigma, pipe.mix, pipe.m], np.float32),
                patch_potential=pipe.patch_potential.astype(np.float32))


if __name__ == "__main__":
    for name, fn in (("svr_small", svr_case), ("reg_small", reg_case), ("pvr_small", pvr_case)):
        d = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path), "bytes")
