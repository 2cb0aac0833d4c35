#!/usr/bin/env python
"""The reference's own CUDA path (oracle/_ref, recompiled for sm_100a) and ours on the SAME C3-geometry stacks: per-call
wall times (every call is synchronous) and the volume after one outer iteration, for a like-for-like speed and parity
figure at full slice/volume size.  Test tooling.  Each arm runs in its own process (the reference resets the device):

    python tools/ref_bench.py gen  --stacks 2                       # synthetic stacks -> /tmp/c3_sub.pt (torch, GPU)
    python tools/ref_bench.py ref  --out gpurun_out/refbench_ref.npz
    python tools/ref_bench.py cuda --out gpurun_out/refbench_cuda.npz
    python tools/ref_bench.py cmp  gpurun_out/refbench_ref.npz gpurun_out/refbench_cuda.npz gpurun_out/refbench.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
CACHE = "/tmp/c3_sub.pt"


class Timed:
    """Wraps a backend: wall time of every (synchronous) method call, summed per method name."""

    def __init__(self, b, sync=None):
        object.__setattr__(self, "_b", b)
        object.__setattr__(self, "_sync", sync)
        object.__setattr__(self, "times", {})

    def __getattr__(self, name):
        a = getattr(self._b, name)
        if not callable(a):
            return a

        def f(*args, **kw):
            t = time.perf_counter()
            r = a(*args, **kw)
            if self._sync:
                self._sync()
            dt = time.perf_counter() - t
            e = self.times.setdefault(name, [0.0, 0])
            e[0] += dt; e[1] += 1
            return r
        return f

    def __setattr__(self, k, v):
        setattr(self._b, k, v)


def run_arm(arm, out, rec_iters, timing_only=False, reps=3):
    import torch
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    ds = torch.load(CACHE, weights_only=False)
    if arm == "ref":
        from oracle.ref_backend import RefReconstruction
        from oracle.ref_runner import ref_pipeline_cls
        raw = RefReconstruction(0)
        b = Timed(raw)
        cls = ref_pipeline_cls()
    else:
        from fetalreconstruction_b200.reconstruction import Reconstruction
        raw = Reconstruction(0)
        b = Timed(raw, raw.synchronize)
        cls = SVRPipeline
    t0 = time.perf_counter()
    upload_dataset(b, ds)
    p = cls(b, ds.S, 0, ds.S, params=SVRParams(iterations=4, rec_iterations_first=rec_iters))
    p.InitializeEMGPU(ds.slices)
    setup_s = time.perf_counter() - t0
    if timing_only:
        # bench.py's reference_cuda leg: one untimed outer iteration (first-call costs of the process: module load, lazy
        # allocations), then `reps` timed ones; the MEDIAN is reported (the reference allocates and frees three volumes in every
        # Superresolution call, whose cost varies from run to run, so all timings are listed too)
        p.outer_iteration(0)
        runs = []
        for _ in range(max(reps, 1)):
            b.times.clear()
            t0 = time.perf_counter()
            p.outer_iteration(0)
            runs.append((time.perf_counter() - t0, {k: 1e3 * t / n for k, (t, n) in b.times.items()}))
        order = sorted(range(len(runs)), key=lambda i: runs[i][0])
        med = runs[order[len(order) // 2]]
        print("REFBENCH_JSON " + json.dumps({"arm": arm, "S": int(ds.S), "rec_iters": rec_iters, "iteration_s": med[0], "setup_s": setup_s,
                                             "iterations_s": [r[0] for r in runs], "ms_per_call": med[1]}))
        return
    b.times.clear()
    res = {}
    f16 = lambda a: np.asarray(a).astype(np.float16)
    t0 = time.perf_counter()
    # outer_iteration(0) of fetalreconstruction_b200/pipeline.py, with the stage outputs captured (untimed reads)
    p.set_schedule(0)
    p.InitializeEMValuesGPU()
    p.GaussianReconstructionGPU()
    tcap = time.perf_counter()
    res["k1_volume"] = raw.syncCPU(); res["k1_psf"] = f16(raw.debugv_PSF_sums()); res["k1_small_slices"] = np.asarray(p._small_slices)
    cap = time.perf_counter() - tcap
    p.SimulateSlicesGPU()
    p.InitializeRobustStatisticsGPU()
    p.EStepGPU()
    tcap = time.perf_counter()
    res["e0_sim"] = f16(raw.debugSimslices()); res["e0_sigma"] = np.float64(p._sigma); res["e0_potential"] = p.slice_potential.copy()
    res["e0_weights"] = f16(raw.debugWeights()); res["e0_slice_weight"] = p._slice_weight.copy()
    cap += time.perf_counter() - tcap
    for i in range(rec_iters):
        p.reconstruction_iteration(i)
        if i == 0:
            tcap = time.perf_counter()
            res["r1_scale"] = p._scale.copy(); res["r1_volume"] = raw.syncCPU(); res["r1_sim"] = f16(raw.debugSimslices())
            res["r1_em"] = np.array([p._sigma, p._mix, p._m], np.float64); res["r1_slice_weight"] = p._slice_weight.copy()
            res["r1_addon"] = raw.debugAddon(); res["r1_weights"] = f16(raw.debugWeights())
            cap += time.perf_counter() - tcap
    p.MaskVolumeGPU()
    iter_s = time.perf_counter() - t0 - cap
    res.update(volume=raw.syncCPU(), scale=p._scale, slice_weight=p._slice_weight, em=np.array([p._sigma, p._mix, p._m], np.float64),
               times=json.dumps({k: v for k, v in b.times.items()}), setup_s=setup_s, iter_s=iter_s, S=ds.S)
    if out:
        np.savez(out, **res)
    print(arm, "S", ds.S, "setup %.2fs iteration %.2fs" % (setup_s, iter_s))
    for k, (t, n) in sorted(b.times.items(), key=lambda kv: -kv[1][0]):
        print(f"  {k:40s} {n:3d} calls {1e3 * t / n:10.2f} ms/call")
    # one machine-readable line for bench.py
    print("REFBENCH_JSON " + json.dumps({"arm": arm, "S": int(ds.S), "rec_iters": rec_iters, "iteration_s": iter_s, "setup_s": setup_s,
                                         "ms_per_call": {k: 1e3 * t / n for k, (t, n) in b.times.items()}}))


def stats(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    nz = b[(b != 0) & np.isfinite(b)]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    fin = np.isfinite(a) & np.isfinite(b)
    d = np.abs(a[fin] - b[fin]) / scale
    return {"rel_rms": float(np.sqrt(np.mean(d ** 2))), "rel_max": float(d.max()), "rel_p999": float(np.quantile(d, 0.999)),
            "scale": scale, "n": int(a.size)}


def compare(ref_path, cuda_path, out_path):
    r, c = dict(np.load(ref_path)), dict(np.load(cuda_path))
    rep = {"S": int(r["S"]), "parity_cuda_vs_reference": {}, "times_ms_per_call": {}}
    for k in sorted(r.keys()):
        if k in ("times", "setup_s", "iter_s", "S"):
            continue
        rep["parity_cuda_vs_reference"][k] = stats(c[k], r[k])
    rep["values"] = {k: {"reference": np.asarray(r[k], np.float64).ravel()[:8].tolist(), "ours": np.asarray(c[k], np.float64).ravel()[:8].tolist()}
                     for k in ("e0_sigma", "r1_em", "em")}
    d = np.abs(np.asarray(c["scale"], np.float64) - np.asarray(r["scale"], np.float64))
    worst = np.argsort(d)[::-1][:6]
    rep["values"]["scale_worst"] = [{"slice": int(i), "reference": float(r["scale"][i]), "ours": float(c["scale"][i]),
                                     "slice_weight": float(r["slice_weight"][i])} for i in worst]
    tr, tc = json.loads(str(r["times"])), json.loads(str(c["times"]))
    groups = {"GaussianReconstruction (K1+K6)": ["GaussianReconstruction", "gaussian_reconstruction_local", "gaussian_reconstruction_finish"],
              "SimulateSlices (K2)": ["SimulateSlices"],
              "Superresolution (K3+K4/K5)": ["Superresolution", "superresolution_local", "superresolution_finish"],
              "EStep (K7/K8)": ["EStep"], "MStep (K9)": ["MStep", "mstep_local"],
              "CalculateScaleVector (K10)": ["CalculateScaleVector"]}
    for name, keys in groups.items():
        def per_call(t):
            hit = [t[k] for k in keys if k in t]
            n = max([h[1] for h in hit], default=1)
            return sum(h[0] for h in hit) / max(n, 1) * 1e3, n
        (a, n), (b, _) = per_call(tr), per_call(tc)
        rep["times_ms_per_call"][name] = {"reference_cuda": a, "ours": b, "speedup": a / b if b > 0 else None, "calls": n}
    rep["outer_iteration_s"] = {"reference_cuda": float(r["iter_s"]), "ours": float(c["iter_s"]),
                                "speedup": float(r["iter_s"]) / float(c["iter_s"])}
    with open(out_path, "w") as f:
        json.dump(rep, f, indent=1)
    print(json.dumps(rep, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("arm", choices=["gen", "ref", "cuda", "cmp"])
    ap.add_argument("paths", nargs="*")
    ap.add_argument("--stacks", type=int, default=2)
    ap.add_argument("--out", default=None)
    ap.add_argument("--rec-iters", type=int, default=2)
    ap.add_argument("--timing-only", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--stack-list", default=None, help="gen: comma-separated stack numbers instead of the first --stacks")
    a = ap.parse_args()
    if a.arm == "gen":
        import torch
        from fetalreconstruction_b200.phantom import c3_config, make_dataset
        if a.stack_list:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import ref_live
            ds = ref_live.gen(CACHE, [int(v) for v in a.stack_list.split(",")])
        else:
            ds = make_dataset(c3_config(), device="cuda" if torch.cuda.is_available() else "cpu", stacks=range(a.stacks))
            torch.save(ds, CACHE)
        print("generated", ds.S, "slices", ds.slices.shape)
    elif a.arm == "cmp":
        compare(*a.paths)
    else:
        run_arm(a.arm, a.out, a.rec_iters, a.timing_only, a.reps)


if __name__ == "__main__":
    main()
