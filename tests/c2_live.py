"""BASELINE.json configs[1] ("C2") helper (TEST INFRASTRUCTURE): the bundled 4-stack 3T data after the set-up pipeline
(tests/golden/c2_setup.npz, see tests/golden/make_c2_setup.py), full default schedule (4 outer iterations x 4,4,4,13
super-resolution iterations, transformations fixed), on libsvr_b200.so or on the reference's own CUDA path
(oracle/_ref/libref_cuda2.so; it resets the device, so it runs in a process of its own):

    python tests/c2_live.py ref|cuda OUT.npz
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FIXTURE = os.path.join(ROOT, "tests", "golden", "c2_setup.npz")


from fetalreconstruction_b200.fixtures import load_c2_setup as load_setup  # noqa: E402


def run(arm, out=None, iterations=4):
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    ds = load_setup()
    if arm == "ref":
        from oracle.ref_backend import RefReconstruction
        from oracle.ref_runner import ref_pipeline_cls
        b, cls = RefReconstruction(0), ref_pipeline_cls()
    else:
        from fetalreconstruction_b200.reconstruction import Reconstruction
        b, cls = Reconstruction(0), SVRPipeline
    t0 = time.perf_counter()
    upload_dataset(b, ds)
    p = cls(b, ds.S, 0, ds.S, params=SVRParams(iterations=iterations))
    p.InitializeEMGPU(ds.slices)
    res = {}
    for it in range(iterations):
        p.outer_iteration(it)
        res[f"image{it}"] = b.syncCPU().astype(np.float32)
    b.RestoreSliceIntensities(ds.stack_factor, ds.stack_index)
    p.ScaleVolumeGPU()
    res["volume"] = b.syncCPU().astype(np.float32)
    res.update(scale=p._scale, slice_weight=p._slice_weight, em=np.array([p._sigma, p._mix, p._m], np.float64),
               total_s=time.perf_counter() - t0, shape=np.array(ds.mask.shape))
    if out:
        np.savez(out, **res)
    return res


def volume_stats(a, ref):
    """RMSE and max-abs of a against ref, relative to the RMS of ref's non-zero voxels; also in intensity units."""
    a, r = np.asarray(a, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    nz = r[(r != 0) & np.isfinite(r)]
    sc = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - r)
    return {"rmse_rel": float(np.sqrt(np.mean(d ** 2))) / sc, "max_abs_rel": float(d.max()) / sc, "p999_rel": float(np.quantile(d, 0.999)) / sc,
            "rmse": float(np.sqrt(np.mean(d ** 2))), "max_abs": float(d.max()), "intensity_rms": sc}


if __name__ == "__main__":
    run(sys.argv[1], sys.argv[2])
