"""A/B of the PSF kernel variants (svr_set_tuning) at the C3 geometry, one stack (= one orientation) at a time:
per-launch device times and the agreement of the variants' outputs.  GPU-box tooling (tools/README.md).

  python tools/ww_ab.py [stacks...]      -> JSON lines on stdout
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fetalreconstruction_b200.phantom import make_dataset, c3_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from fetalreconstruction_b200.reconstruction import Reconstruction

stacks = [int(a) for a in sys.argv[1:]] or list(range(8))
REP = 3


def timed(b, kind, fn):
    b.profile_reset(); b.profile_enable(True)
    for _ in range(REP):
        fn()
    b.profile_enable(False)
    ms, n = b.profile_read()[kind]
    return ms / max(n, 1)


def rel(a, ref):
    sc = float(np.sqrt(np.mean(ref.astype(np.float64) ** 2))) or 1.0
    d = a.astype(np.float64) - ref.astype(np.float64)
    return {"rms": float(np.sqrt(np.mean(d ** 2))) / sc, "max": float(np.abs(d).max()) / sc}


for st in stacks:
    ds = make_dataset(c3_config(), device="cuda", stacks=[st])
    b = Reconstruction(0)
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
    p.InitializeEMGPU(ds.slices)
    b.set_tuning(b.TUNE_SCATTER, 0); b.set_tuning(b.TUNE_SIMULATE, 0)
    p.outer_iteration(0)                       # a realistic state: weights, simulated slices, scales
    sw = p._local(p._slice_weight)
    a = ds.stack_attrs[0]
    out = {"stack": st, "slice_x_axis": [round(float(v), 3) for v in a.xaxis], "slice_y_axis": [round(float(v), 3) for v in a.yaxis],
           "pixels": int(np.count_nonzero(b.debugv_PSF_sums()))}
    ref_acc = ref_vol = ref_sim = None
    for v in (0, 1, 2, 3):
        b.set_tuning(b.TUNE_SCATTER, v)
        out[f"K3_scatter{v}_ms"] = round(timed(b, "superres", lambda: b.superresolution_local(sw)), 3)
        addon, cmap = b.debugAddon(), b.debugConfidenceMap()
        if v in (1, 2):
            st_ = b.debugWindowStats()
            out[f"K3_scatter{v}_plan"] = {"S1_2_4_8_16": st_[0:5].tolist(), "no_window": int(st_[5]), "degree1to8": st_[8:16].tolist(),
                                          "tma": int(st_[16]), "general": int(st_[17]), "shared_centre_tiles": int(st_[18]), "fallback_px": int(st_[19])}
        if v == 0:
            ref_acc = (addon, cmap)
        else:
            out[f"K3_scatter{v}_vs0"] = {"addon": rel(addon, ref_acc[0]), "cmap": rel(cmap, ref_acc[1])}
    # K1 runs on a copy of the state (it resets weights / simulated slices): do it last
    sims = {}
    for v in (0, 1, 2):
        b.set_tuning(b.TUNE_SIMULATE, v)
        out[f"K2_sim{v}_ms"] = round(timed(b, "simulate", lambda: b.SimulateSlices()), 3)
        sims[v] = (b.debugSimslices(), b.debugSimweights(), b.debugSiminside())
        if v:
            out[f"K2_sim{v}_vs0"] = {"sim": rel(sims[v][0], sims[0][0]), "simw": rel(sims[v][1], sims[0][1]),
                                     "inside_diff": int(np.count_nonzero(sims[v][2] != sims[0][2]))}
    b.set_tuning(b.TUNE_SIMULATE, 0)
    for v in (0, 1, 2, 3):
        b.set_tuning(b.TUNE_SCATTER, v)
        out[f"K1_scatter{v}_ms"] = round(timed(b, "gaussian", lambda: b.gaussian_reconstruction_local()), 3)
        vn = b.gaussian_reconstruction_finish()
        vol, vw = b.syncCPU(), b.getVolWeights()
        if v == 0:
            ref_vol = (vol, vw, vn)
        else:
            out[f"K1_scatter{v}_vs0"] = {"volume": rel(vol, ref_vol[0]), "volw": rel(vw, ref_vol[1]),
                                         "voxel_num_diff": int(np.count_nonzero(vn != ref_vol[2]))}
    print(json.dumps(out), flush=True)
    b.close()
