#!/usr/bin/env python
"""The reference's own PVR CUDA path (oracle/_ref/libref_pvr.so) and ours on the same, larger PVR case: per-call wall times
and the parity of every stage.  Test tooling; GPU box:   python tools/ref_bench_pvr.py OUT.json [vol slices size]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    from ref_bench import Timed
    from fetalreconstruction_b200.pvr import PatchReconstruction, PVRPipeline
    from oracle.ref_backend_pvr import RefPatchReconstruction, ref_pvr_pipeline_cls
    from oracle.ref_runner_pvr import pvr_stages
    from pvr_case import make_pvr_case
    out_path = sys.argv[1]
    vol, slices, size = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (128, 24, 128)
    t0 = time.perf_counter()
    case = make_pvr_case(seed=43, vol=vol, n_stacks=3, slices=slices, size=size, inplane=1.0, spacing=2.5, pbb=(64, 64), stride=(32, 32))
    print("case: %d patches of 64x64, volume %d^3, enumeration %.1f s" % (len(case["attrs"]), vol, time.perf_counter() - t0), flush=True)
    ours = Timed(PatchReconstruction(0), torch.cuda.synchronize)
    a = pvr_stages(ours, PVRPipeline, case=case)
    ref = Timed(RefPatchReconstruction(0))
    b = pvr_stages(ref, ref_pvr_pipeline_cls(), patch_cube=a["patches"], case=case)
    rep = {"patches": int(len(case["attrs"])), "patch_size": [64, 64], "volume": [vol] * 3, "parity_ours_vs_reference": {}, "ms_per_call": {}}
    inside = case["mask"].ravel() != 0
    for k in b:
        if k in ("per_stack",):
            continue
        r, x = np.asarray(b[k], np.float64).ravel(), np.asarray(a[k], np.float64).ravel()
        if k.endswith("_addon") or k.endswith("_cmap"):
            r, x = r[inside], x[inside]
        nz = r[(r != 0) & np.isfinite(r)]
        sc = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
        d = np.abs(r - x) / sc
        rep["parity_ours_vs_reference"][k] = {"rel_rms": float(np.sqrt(np.mean(d ** 2))), "rel_max": float(d.max())}
    names = {"patchBasedPSFReconstruction_gpu": "P1 PSF reconstruction", "patchBasedSimulatePatches_gpu": "P2 simulate patches",
             "superresolution_run": "P3 super-resolution scatter", "superresolution_regularize": "P4 regularise"}
    for k, label in names.items():
        ra, oa = ref.times.get(k), ours.times.get(k)
        if ra and oa:
            rep["ms_per_call"][label] = {"reference_cuda": 1e3 * ra[0] / ra[1], "ours": 1e3 * oa[0] / oa[1], "speedup": (ra[0] / ra[1]) / (oa[0] / oa[1])}
    em_ref = sum(v[0] for k, v in ref.times.items() if k.startswith("ref_")) * 1e3
    em_ours = sum(v[0] for k, v in ours.times.items() if k.startswith("rs_")) * 1e3
    rep["ms_per_call"]["robust statistics, whole iteration (init, 3 E-steps, 2 M-steps, 2 scales)"] = {"reference_cuda": em_ref, "ours": em_ours,
                                                                                                      "speedup": em_ref / em_ours if em_ours else None}
    json.dump(rep, open(out_path, "w"), indent=1)
    for k, v in rep["ms_per_call"].items():
        print(k, {a_: round(b_, 3) for a_, b_ in v.items()})
    print("volume parity", rep["parity_ours_vs_reference"]["volume"], "worst stage", max(rep["parity_ours_vs_reference"].items(), key=lambda kv: kv[1]["rel_rms"]))


if __name__ == "__main__":
    main()
