"""Prints (does not assert) the CUDA-vs-oracle deviation of every hot-path output on the small seeded
dataset, so the tolerances in tests/test_gpu_parity.py can be read against measured numbers.
Run on the GPU box:  python tools/parity_report.py > gpurun_out/parity_report.txt"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fetalreconstruction_b200.phantom import make_dataset, small_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from fetalreconstruction_b200.reconstruction import Reconstruction
from oracle.oracle_backend import OracleReconstruction


def stats(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    nz = b[b != 0]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - b) / scale
    return dict(rms=float(np.sqrt(np.mean(d ** 2))), p999=float(np.quantile(d, 0.999)), max=float(d.max()),
                frac_gt_1e3=float(np.mean(d > 1e-3)), support_mismatch=int(np.sum((a != 0) != (b != 0))))


def main():
    ds = make_dataset(small_config())
    g, o = Reconstruction(0), OracleReconstruction()
    rep = {}
    for b in (g, o):
        upload_dataset(b, ds)
        b.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
        b.InitializeEMValues()
    vg, vo = g.GaussianReconstruction(), o.GaussianReconstruction()
    rep["voxel_num_maxdiff"] = int(np.abs(vg.astype(int) - vo.astype(int)).max())
    rep["psf_sums"] = stats(g.debugv_PSF_sums(), o.debugv_PSF_sums())
    rep["volweights"] = stats(g.getVolWeights(), o.getVolWeights())
    rep["gaussian_recon"] = stats(g.syncCPU(), o.syncCPU())
    g.SimulateSlices(); o.SimulateSlices()
    rep["simslices"] = stats(g.debugSimslices(), o.debugSimslices())
    rep["simweights"] = stats(g.debugSimweights(), o.debugSimweights())
    sg, so = g.InitializeRobustStatistics(), o.InitializeRobustStatistics()
    rep["sigma"] = (sg, so)
    pos = ds.slices[ds.slices > 0]
    m = 1.0 / (2.1 * pos.max() - 1.9 * pos.min())
    pg, po = g.EStep(m, so, 0.9), o.EStep(m, so, 0.9)
    rep["potential_maxabs"] = float(np.abs(pg - po).max())
    rep["weights"] = stats(g.debugWeights(), o.debugWeights())
    rep["scale_maxrel"] = float(np.abs(g.CalculateScaleVector() / o.CalculateScaleVector() - 1).max())
    sw = np.ones(ds.S, np.float32)
    args = (1, sw, False, 1.0, float(pos.min()), float(pos.max()), 150.0, 0.02 * 150 * 150)
    g.Superresolution(*args); o.Superresolution(*args)
    rep["cmap_norm"] = stats(g.debugConfidenceMap(), o.debugConfidenceMap())
    rep["addon_norm"] = stats(g.debugAddon(), o.debugAddon())
    rep["superres_volume"] = stats(g.syncCPU(), o.syncCPU())
    g.SimulateSlices(); o.SimulateSlices()
    rep["mstep"] = (g.MStep(2, 1e-4, so, 0.9, m), o.MStep(2, 1e-4, so, 0.9, m))
    # full pipeline
    vols = {}
    for name, b in (("gpu", Reconstruction(0)), ("orc", OracleReconstruction())):
        upload_dataset(b, ds)
        p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(iterations=2, rec_iterations_first=3, rec_iterations_last=5))
        p.InitializeEMGPU(ds.slices)
        t = time.time(); vols[name] = p.run(); rep["time_" + name] = time.time() - t
        vols[name + "_w"] = p._slice_weight.copy()
    inm = ds.mask.ravel() != 0
    rep["pipeline_volume_inmask"] = stats(vols["gpu"][inm], vols["orc"][inm])
    rep["pipeline_slice_weight_maxabs"] = float(np.abs(vols["gpu_w"] - vols["orc_w"]).max())
    print(json.dumps(rep, indent=1, default=lambda x: [float(v) for v in x]))


if __name__ == "__main__":
    main()
