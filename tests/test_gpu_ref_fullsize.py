"""Full-size parity against the reference's own CUDA code, LIVE on the GPU box.

SVR: C3 geometry (256x256 slices into 256^3 at 0.75 mm), stacks 0 (aligned), 2 (through-plane: K2's staged-row path) and
7 (the most oblique), one outer iteration with one super-resolution step, ours against oracle/_ref/libref_cuda2.so
(the unmodified reconstruction_cuda2.cu compiled for sm_100a) run in a process of its own on the same inputs.
PVR: 64x64 patches at stride 32 against oracle/_ref/libref_pvr.so.

Bounds = about three times what was measured on B200 for the same comparison (profiles/r01_v6_refbench_c3_2stacks_ours_vs_
reference.json, profiles/r02_ref_live_*.json), as RMS relative to the RMS of the reference's non-zeros, plus a COUNT bound
on elements further off than a threshold: one flipped epsilon-skip decision (|old - psf| against 1e-5, a 1-ulp effect,
SURVEY quirk Q1) moves a whole tap, so the maximum is not a meaningful bound but the number of such elements is.
Where the reference itself is not reproducible run to run (its regulariser updates the volume in place while
neighbouring threads read it, deviation D3) the bound says so.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_cuda2.so")
REF_PVR_LIB = os.path.join(ROOT, "oracle", "_ref", "libref_pvr.so")

pytestmark = pytest.mark.gpu

# field: (rms bound, flip threshold, allowed fraction of elements beyond the threshold)
# measured on B200 (profiles/r02_ref_live_svr.json): rms / elements beyond the threshold
#   k1_volume 1.3e-5 / 432 of 16.8 M, k1_psf 1.0e-4 / 1463 of 12.6 M, k2_sim 7.3e-6 / 97, k2_simw 5.8e-6 / 2, e_weights 6.7e-6 / 0,
#   k3_addon 2.5e-4 / 418 of 5.2 M, k5_volume 3.0e-4 / 27
SVR_BOUNDS = {
    "k1_volume": (5e-5, 1e-3, 1e-4),
    "k1_psf": (3e-4, 5e-3, 4e-4),
    "k2_sim": (3e-5, 1e-3, 5e-5),
    "k2_simw": (3e-5, 5e-3, 1e-5),
    "e_weights": (3e-5, 1e-2, 1e-5),
    "k3_addon": (8e-4, 1e-2, 3e-4),
    "k3_cmap": (1e-6, 1e-2, 1e-6),            # normalised to exactly 1 where non-zero (non-adaptive regularisation)
    "k5_volume": (1e-3, 1e-2, 1e-4),          # the reference alone: 1.4e-4 rms / 6e-3 max between two of its own runs (D3)
}
SVR_SCALARS = {"k11_sigma": 1e-5, "e_potential": 2e-5, "k10_scale": 2e-5, "k9_em": 5e-3}


def _report(name, rep):
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, name), "w") as f:
            json.dump(rep, f, indent=1)


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libref_cuda2.so not built (make -C oracle ref)")
def test_c3_stacks_against_reference_cuda_live(tmp_path):
    import ref_live
    ds_path, ref_out = str(tmp_path / "c3_sub.pt"), str(tmp_path / "ref.npz")
    ref_live.gen(ds_path, [0, 2, 7], 64)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_live.py"), "ref", ds_path, ref_out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ref = dict(np.load(ref_out))
    ours = ref_live.run_arm("cuda", ds_path)
    rep, bad = {}, []
    assert np.array_equal(ours["k2_inside"], ref["k2_inside"]), "siminside flags differ"
    inside = ref["mask"]
    for k, (rms_b, thr, frac) in SVR_BOUNDS.items():
        a, b = ours[k], ref[k]
        if k in ("k3_addon", "k3_cmap"):        # the reference leaves the accumulators of out-of-mask voxels untouched; we zero them
            a, b = a[inside], b[inside]
        rms, mx, nflip, nonfinite = ref_live.field_stats(a, b, thr)
        rep[k] = {"rms": rms, "max": mx, "beyond_threshold": nflip, "threshold": thr, "n": int(np.size(b)), "nonfinite": nonfinite,
                  "bound_rms": rms_b, "bound_count": int(frac * np.size(b))}
        if not (rms <= rms_b and nflip <= frac * np.size(b)):
            bad.append(k)
    for k, tol in SVR_SCALARS.items():
        a, b = np.asarray(ours[k], np.float64), np.asarray(ref[k], np.float64)
        d = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))) if k != "e_potential" else float(np.max(np.abs(a - b)))
        rep[k] = {"max_rel": d, "bound": tol}
        if not d <= tol:
            bad.append(k)
    sw = float(np.max(np.abs(ours["e_slice_weight"] - ref["e_slice_weight"])))
    rep["e_slice_weight"] = {"max_abs": sw, "bound": 1e-4}
    if sw > 1e-4:
        bad.append("e_slice_weight")
    _report("r02_ref_live_svr.json", rep)
    assert not bad, json.dumps({k: rep[k] for k in bad}, indent=1)


@pytest.mark.skipif(not os.path.exists(REF_PVR_LIB), reason="oracle/_ref/libref_pvr.so not built (make -C oracle ref)")
def test_pvr_64x64_against_reference_cuda_live():
    from fetalreconstruction_b200.pvr import PatchReconstruction, PVRPipeline
    from oracle.ref_backend_pvr import RefPatchReconstruction, ref_pvr_pipeline_cls
    from oracle.ref_runner_pvr import pvr_stages
    from pvr_case import make_pvr_case
    case = make_pvr_case(seed=43, vol=128, n_stacks=3, slices=24, size=128, inplane=1.0, spacing=2.5, pbb=(64, 64), stride=(32, 32))
    a = pvr_stages(PatchReconstruction(0), PVRPipeline, case=case)
    b = pvr_stages(RefPatchReconstruction(0), ref_pvr_pipeline_cls(), patch_cube=a["patches"], case=case)
    inside = case["mask"].ravel() != 0
    rep, bad = {"patches": int(len(case["attrs"]))}, []
    for k in b:
        if k == "per_stack":
            continue
        integral = np.asarray(b[k]).dtype.kind in "iub"
        r, x = np.asarray(b[k], np.float64).ravel(), np.asarray(a[k], np.float64).ravel()
        if k.endswith("_addon") or k.endswith("_cmap"):
            r, x = r[inside], x[inside]
        if integral or k.endswith("inside"):
            rep[k] = {"mismatches": int(np.count_nonzero(r != x)), "n": int(r.size)}
            if rep[k]["mismatches"] > 1e-5 * r.size:       # flags of pixels whose only accepted tap sits on the skip threshold (measured: 2 of 2.6 M)
                bad.append(k)
            continue
        rms, mx, nflip, _ = __import__("ref_live").field_stats(x, r, 1e-2)
        rep[k] = {"rms": rms, "max": mx, "beyond_1e-2": nflip, "n": int(r.size)}
        # measured on B200 (profiles/r02_ref_live_pvr.json): PSF sums 3.1e-4, addon 5.9e-4 / 7.1e-4, cmap 3.1e-4 (sums of few taps per
        # voxel at 12^3 support: a flipped tap weighs more than in SVR), everything else <= 1.7e-4; <= 6e-4 of the elements beyond 1e-2
        rms_b = 2e-3 if "addon" in k else (1e-3 if ("cmap" in k or "psf" in k) else 5e-4)
        if not (rms <= rms_b and nflip <= 2e-3 * max(r.size, 1000)):
            bad.append(k)
    _report("r02_ref_live_pvr.json", rep)
    assert not bad, json.dumps({k: rep[k] for k in bad}, indent=1)


# ---- C2 = BASELINE.json configs[1]: the bundled 4-stack 3T data, full iteration count ------------------------------------------
# Stated tolerance (north_star: "voxel max-abs and RMSE"), final volume against the reference's own CUDA output on the same
# inputs, relative to the RMS of the reference volume's non-zero voxels:
#     RMSE <= 2.5 %,   99.9th percentile of |diff| <= 30 %,   max-abs <= 300 %  (isolated voxels at the mask boundary)
# Two runs of the REFERENCE itself on these inputs differ by ~1 % RMSE / ~13 % p99.9 (measured on B200,
# profiles/r01_c2_parity_noreg_reference_vs_reference.json: its regulariser updates the volume in place while neighbouring
# threads read it, and the volume's y-size 105 is not a multiple of its 8x8x8 block), so the bound is ~2.5x the reference's own
# spread -- "about 1.5x the reference's spread" is what we measure (1.3e-2 against 0.8-1.1e-2), not "inside" it.
C2_RMSE, C2_P999, C2_MAX = 2.5e-2, 0.30, 3.0


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libref_cuda2.so not built (make -C oracle ref)")
def test_c2_bundled_data_against_reference_cuda_live(tmp_path):
    import c2_live
    outs = []
    for i in range(2):                                      # the reference twice: its own run-to-run spread, stated beside ours
        o = str(tmp_path / f"c2_ref{i}.npz")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "c2_live.py"), "ref", o], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        outs.append(dict(np.load(o)))
    ours = c2_live.run("cuda")
    rep = {"ours_vs_reference": {}, "reference_vs_reference": {}, "seconds": {"ours": float(ours["total_s"]), "reference": float(outs[1]["total_s"])}}
    for k in ("image0", "image1", "image2", "image3", "volume"):
        rep["ours_vs_reference"][k] = c2_live.volume_stats(ours[k], outs[0][k])
        rep["reference_vs_reference"][k] = c2_live.volume_stats(outs[1][k], outs[0][k])
    both = (outs[0]["slice_weight"] >= 0.5) == (ours["slice_weight"] >= 0.5)
    rep["slices_classified_alike"] = int(both.sum()); rep["slices"] = int(both.size)
    rep["scale_max_abs_diff"] = float(np.abs(ours["scale"] - outs[0]["scale"]).max())
    rep["scale_max_abs_diff_reference_vs_reference"] = float(np.abs(outs[1]["scale"] - outs[0]["scale"]).max())
    _report("r02_c2_live.json", rep)
    v = rep["ours_vs_reference"]["volume"]
    assert np.isfinite(ours["volume"]).all()
    assert v["rmse_rel"] <= C2_RMSE and v["p999_rel"] <= C2_P999 and v["max_abs_rel"] <= C2_MAX, json.dumps(rep, indent=1)
    assert rep["ours_vs_reference"]["image0"]["rmse_rel"] <= 1.5e-2, json.dumps(rep["ours_vs_reference"]["image0"])
    assert rep["slices_classified_alike"] >= 0.97 * rep["slices"], rep
    # per-slice intensity scales: the reference alone moves them by 1.6e-2 between two runs (measured); ours-vs-reference 2.1e-2
    assert rep["scale_max_abs_diff"] <= 5e-2, rep
