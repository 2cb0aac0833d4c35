#!/usr/bin/env python
"""Runs the reference's own CUDA path (oracle/_ref/libref_cuda2.so, see oracle/ref_backend.py) on the seeded
golden cases of tests/golden/make_golden.py and writes what it produced to .npz files (TEST INFRASTRUCTURE ONLY).

    python -m oracle.ref_runner --out DIR [--cases svr,reg,steps]

Must run in its own process on a GPU box (the reference constructor resets the device).  Outputs:
  DIR/ref_svr_small.npz    the SVR golden case (Gaussian reconstruction, simulate, EM, 2 SR iterations, mask)
  DIR/ref_reg_small.npz    the registration golden case (NCC per slice at both levels, optimised transforms)
  DIR/ref_steps_small.npz  per-kernel outputs on the `small_ds` case of tests/conftest.py with fixed parameters
The files written on a B200 are committed under tests/golden/ and pin BOTH the CPU oracle (tests, not gpu) and the
CUDA path (tests, gpu) against the reference itself.
"""
import argparse
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def ref_pipeline_cls():
    from fetalreconstruction_b200.pipeline import SVRPipeline

    class RefPipeline(SVRPipeline):
        """The reference finishes M-step and volume scaling inside the class (cuda2.cu:3056-3072,3458-3470)."""

        def MStepGPU(self, it):
            self._sigma, self._mix, self._m = self.b.MStep(it, self._step, self._sigma, self._mix, self._m)

        def ScaleVolumeGPU(self):
            return self.b.ScaleVolume()

    return RefPipeline


def steps_case(b, slice_size=None):
    """The call sequence of tests/test_gpu_parity.py (`pair` fixture + test_em_steps_and_superresolution) with every
    parameter fixed, so each kernel's output can be compared on its own."""
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    from fetalreconstruction_b200.pipeline import upload_dataset
    cfg = small_config()
    cfg.slice_size = tuple(slice_size or _mg().REF_STEPS_SIZE)
    ds = make_dataset(cfg)
    out = {}
    upload_dataset(b, ds)
    b.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
    b.InitializeEMValues()
    out["voxel_num"] = b.GaussianReconstruction()
    out["recon0"] = b.syncCPU()
    out["volw"] = b.getVolWeights()
    out["psf"] = b.debugv_PSF_sums()
    out["inside"] = b.SimulateSlices()
    out["sim"] = b.debugSimslices()
    out["simw"] = b.debugSimweights()
    out["simi"] = b.debugSiminside()
    out["sigma0"] = np.float32(b.InitializeRobustStatistics())
    pos = ds.slices[ds.slices > 0]
    m = 1.0 / (2.1 * pos.max() - 1.9 * pos.min())
    sigma = np.float32(STEPS_SIGMA)
    out["m"], out["sigma"] = np.float32(m), sigma
    out["potential"] = b.EStep(m, float(sigma), 0.9)
    out["weights"] = b.debugWeights()
    out["scale1"] = b.CalculateScaleVector()
    sw = np.ones(ds.S, np.float32); sw[1] = 0.0; sw[4] = 0.37
    lam, delta = 0.02, 150.0
    b.Superresolution(1, sw, False, min(1.0, 0.05 / lam), float(pos.min()), float(pos.max()), delta, lam * delta * delta)
    out["cmap"] = b.debugConfidenceMap()
    out["addon"] = b.debugAddon()
    out["recon1"] = b.syncCPU()
    b.SimulateSlices()
    out["sim2"] = b.debugSimslices()
    out["mstep"] = np.array(b.MStep(2, 1e-4, float(sigma), 0.9, m), np.float32)
    out["scale2"] = b.CalculateScaleVector()
    b.maskVolume()
    out["recon_masked"] = b.syncCPU()
    return out


STEPS_SIGMA = 400.0     # fixed E-step variance for steps_case (any value of the right order; identical on all backends)


def compact(d, f16=()):
    return {k: (np.asarray(v).astype(np.float16) if k in f16 else np.asarray(v)) for k, v in d.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--cases", default="svr,reg,steps")
    ap.add_argument("--single-threaded", action="store_true", help="Reconstruction(dev, multiThreadedGPU=false)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    from oracle.ref_backend import RefReconstruction
    mg = _mg()
    mt = not a.single_threaded
    for case in a.cases.split(","):
        b = RefReconstruction(0, multithreaded=mt)
        if case == "svr":
            d = mg.svr_case(b, pipeline_cls=ref_pipeline_cls(), slice_size=mg.REF_SVR_SIZE)
        elif case == "reg":
            d = mg.reg_case(b)
        elif case == "steps":
            d = compact(steps_case(b), f16=())
        else:
            raise SystemExit(f"unknown case {case}")
        path = os.path.join(a.out, f"ref_{case}_small.npz")
        np.savez_compressed(path, **d)
        print(case, "->", path, os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    main()
