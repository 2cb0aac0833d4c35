#!/usr/bin/env python
"""BASELINE.json configs[1] ("C2"): the bundled 4-stack 3T data at --resolution 1.0, full iteration count, our CUDA path
against the reference's own CUDA path on identical inputs.

The inputs are what `host/SVRreconstructionGPU ... --dump_setup DIR` writes (packed slices, mask, matrices after the
reference's set-up pipeline: crop, template, mask, intensity matching, slice masking).  Each arm runs in its own
process (the reference resets the device):

    python tools/c2_parity.py run ref|cuda DIR OUT.npz [--iterations 4] [--register]
    python tools/c2_parity.py cmp A.npz B.npz OUT.json
Test tooling; data_local/ is not part of the repository (the bundled data belongs to the reference)."""
import argparse
import json
import os
import sys
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_setup(d):
    idx = dict(line.split() for line in open(os.path.join(d, "index.txt")))
    S, Nx, Ny = int(idx["S"]), int(idx["Nx"]), int(idx["Ny"])
    vx, vy, vz, voxel = int(idx["vx"]), int(idx["vy"]), int(idx["vz"]), float(idx["voxel"])
    f = lambda name, dt: np.fromfile(os.path.join(d, name), dt)
    from fetalreconstruction_b200.geometry import ImageAttributes
    attrs = f("slice_attrs.f64", np.float64).reshape(S, 18)
    slice_attrs = [ImageAttributes(int(a[0]), int(a[1]), int(a[2]), a[3], a[4], a[5], a[6:9].copy(), a[9:12].copy(), a[12:15].copy(),
                                   a[15:18].copy()) for a in attrs]
    ds = SimpleNamespace(S=S, cfg=SimpleNamespace(vol_voxel=voxel, vol_size=(vx, vy, vz)), slices=f("slices.f32", np.float32).reshape(S, Ny, Nx),
                         mask=f("mask.f32", np.float32).reshape(vz, vy, vx), dims=f("dims.f32", np.float32).reshape(S, 3),
                         trans=f("T.f32", np.float32).reshape(S, 16), trans_inv=f("Tinv.f32", np.float32).reshape(S, 16),
                         i2w=f("I2W.f32", np.float32).reshape(S, 16), w2i=f("W2I.f32", np.float32).reshape(S, 16),
                         recon_i2w=f("recon_i2w.f32", np.float32), recon_w2i=f("recon_w2i.f32", np.float32),
                         stack_index=f("stack_index.i32", np.int32), stack_factor=f("stack_factor.f32", np.float32),
                         slice_attrs=slice_attrs, sizes=f("sizes.i32", np.int32).reshape(S, 2))
    return ds


def run(arm, d, out, iterations, register):
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    ds = load_setup(d)
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
    trans = ds.trans.astype(np.float64).copy()
    reg = None
    if register:
        from fetalreconstruction_b200.registration import RegistrationFrontEnd
        fe = RegistrationFrontEnd(b, ds.slices, ds.slice_attrs, ds.cfg.vol_voxel)

        def reg(it):
            nonlocal trans
            trans = fe.SliceToVolumeRegistrationGPU(trans)

    def update_matrices():
        t = trans.reshape(-1, 4, 4)
        b.SetSliceMatrices(t.astype(np.float32).reshape(-1, 16), np.linalg.inv(t).astype(np.float32).reshape(-1, 16), ds.i2w, ds.w2i,
                           ds.i2w, ds.w2i, ds.recon_i2w, ds.recon_w2i)
    res = {}
    per_iter = []
    for it in range(iterations):
        if it > 0 and reg is not None:
            reg(it)
        p.outer_iteration(it, update_matrices)
        per_iter.append(b.syncCPU().astype(np.float32))
    b.RestoreSliceIntensities(ds.stack_factor, ds.stack_index) if hasattr(b, "RestoreSliceIntensities") else None
    p.ScaleVolumeGPU()
    total_s = time.perf_counter() - t0
    res["volume"] = b.syncCPU().astype(np.float32)
    for i, v in enumerate(per_iter):
        res[f"image{i}"] = v
    res.update(scale=p._scale, slice_weight=p._slice_weight, em=np.array([p._sigma, p._mix, p._m], np.float64), total_s=total_s,
               transforms=trans, S=ds.S, shape=np.array(ds.mask.shape))
    np.savez(out, **res)
    print(arm, "S", ds.S, "volume", ds.mask.shape, "iterations", iterations, "register", register, "total %.2f s" % total_s)


def stats(a, b, mask=None):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    nz = b[(b != 0) & np.isfinite(b)]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - b) / scale
    if mask is not None:
        d = d[mask.ravel()]
    return {"rel_rms": float(np.sqrt(np.mean(d ** 2))), "rel_max": float(d.max()), "rel_p999": float(np.quantile(d, 0.999)),
            "scale": scale, "n": int(d.size)}


def tre(setup_dir, Ta, Tb):
    """Per-slice registration difference in mm: mean distance between the positions the two transforms give to the four
    corners and the centre of the slice (parameter differences are meaningless here: rotations are about the world origin)."""
    ds = load_setup(setup_dir)
    out = np.zeros(ds.S)
    for k in range(ds.S):
        sx, sy = ds.sizes[k]
        pts = np.array([[0, 0, 0, 1], [sx - 1, 0, 0, 1], [0, sy - 1, 0, 1], [sx - 1, sy - 1, 0, 1], [(sx - 1) / 2, (sy - 1) / 2, 0, 1]], np.float64)
        w = pts @ ds.i2w[k].reshape(4, 4).astype(np.float64).T
        pa = w @ np.asarray(Ta[k], np.float64).reshape(4, 4).T
        pb = w @ np.asarray(Tb[k], np.float64).reshape(4, 4).T
        out[k] = np.linalg.norm(pa[:, :3] - pb[:, :3], axis=1).mean()
    return out


def compare(pa, pb, out, setup_dir=None):
    a, b = dict(np.load(pa)), dict(np.load(pb))
    vz, vy, vx = (int(v) for v in a["shape"])
    # the reference's regulariser kernels have no x/y bounds check: when a volume dimension is not a multiple of its
    # 8x8x8 block, overhang threads alias onto x < 8 - vx%8 / y < 8 - vy%8 of the next row / plane and apply the update
    # twice (tests/golden/make_golden.py REF_SVR_VOL); report the volume statistics with and without those voxels
    zz, yy, xx = np.meshgrid(np.arange(vz), np.arange(vy), np.arange(vx), indexing="ij")
    ox = (8 - vx % 8) % 8; oy = (8 - vy % 8) % 8
    clean = (xx >= ox) & (yy >= oy)
    rep = {"a": os.path.basename(pa), "b": os.path.basename(pb), "S": int(a["S"]), "volume_shape_zyx": [vz, vy, vx],
           "total_s": {"a": float(a["total_s"]), "b": float(b["total_s"])}, "stats": {}}
    for k in sorted(a.keys()):
        if k in ("total_s", "S", "shape", "transforms"):
            continue
        rep["stats"][k] = stats(b[k], a[k])
        if a[k].size == vz * vy * vx:
            rep["stats"][k + "_without_reference_overhang_voxels"] = stats(b[k], a[k], clean)
    rep["transforms_max_abs_diff"] = float(np.abs(a["transforms"] - b["transforms"]).max())
    from fetalreconstruction_b200.geometry import rigid_parameters
    pa_ = np.stack([rigid_parameters(m.reshape(4, 4)) for m in a["transforms"]]); pb_ = np.stack([rigid_parameters(m.reshape(4, 4)) for m in b["transforms"]])
    dpar = np.abs(pa_ - pb_).max(1)
    rep["transform_param_diff_mm_deg"] = {"median": float(np.median(dpar)), "p90": float(np.quantile(dpar, 0.9)), "max": float(dpar.max()),
                                          "slices_over_1": int((dpar > 1).sum()), "slices_over_5": int((dpar > 5).sum())}
    if setup_dir:
        ident = np.tile(np.eye(4).ravel(), (int(a["S"]), 1))
        both = (a["slice_weight"] >= 0.5) & (b["slice_weight"] >= 0.5)
        q = lambda v: {"median": float(np.median(v)), "p90": float(np.quantile(v, 0.9)), "max": float(v.max())} if v.size else {}
        d_ab = tre(setup_dir, a["transforms"], b["transforms"])
        rep["tre_mm"] = {"a_vs_b_all_slices": q(d_ab), "a_vs_b_slices_included_by_both": q(d_ab[both]),
                         "a_vs_initial": q(tre(setup_dir, a["transforms"], ident)), "b_vs_initial": q(tre(setup_dir, b["transforms"], ident))}
        print("TRE (mm):", json.dumps(rep["tre_mm"]))
    for name, arr in (("a", a), ("b", b)):
        par = np.stack([rigid_parameters(m.reshape(4, 4)) for m in arr["transforms"]])     # initial stack transforms are identity here
        mv = np.abs(par).max(1)
        rep.setdefault("moved_from_initial_mm_deg", {})[name] = {"median": float(np.median(mv)), "p90": float(np.quantile(mv, 0.9)),
            "max": float(mv.max()), "over_5": int((mv > 5).sum()), "over_20": int((mv > 20).sum())}
        w = arr["slice_weight"]
        rep.setdefault("slice_weight_summary", {})[name] = {"included": int((w >= 0.5).sum()), "excluded": int((w < 0.5).sum())}
    print("transform param diff (mm/deg):", rep["transform_param_diff_mm_deg"], "slice weights:", rep["slice_weight_summary"])
    print("moved from initial:", rep["moved_from_initial_mm_deg"])
    with open(out, "w") as f:
        json.dump(rep, f, indent=1)
    for k, v in rep["stats"].items():
        print(f"{k:60s} rms {v['rel_rms']:.2e}  p99.9 {v['rel_p999']:.2e}  max {v['rel_max']:.2e}")
    print("transforms max abs diff", rep["transforms_max_abs_diff"], "total_s", rep["total_s"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["run", "cmp"])
    ap.add_argument("args", nargs="+")
    ap.add_argument("--iterations", type=int, default=4)
    ap.add_argument("--register", action="store_true")
    ap.add_argument("--setup", default=None, help="cmp: set-up directory, adds the per-slice registration difference in mm")
    a = ap.parse_args()
    if a.cmd == "run":
        run(a.args[0], a.args[1], a.args[2], a.iterations, a.register)
    else:
        compare(*a.args, setup_dir=a.setup)


if __name__ == "__main__":
    main()
