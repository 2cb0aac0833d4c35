#!/usr/bin/env python
"""bench.py -- SVR / PVR hot-path throughput on B200 (BASELINE.json metric: slice-projections/s, volumes/hour,
PSF-kernel HBM GB/s against the measured peak).

Workloads (config.workload), one JSON line per run:
  C3 (default)  BASELINE.json configs[2]: synthetic 8 stacks x 128 slices of 256x256 into a 256^3 volume at 0.75 mm.
  C2            BASELINE.json configs[1]: the reference's bundled 4-stack 3T data at --resolution 1.0 after the set-up pipeline
                (tests/golden/c2_setup.npz: 317 slices <= 99x95 into 88x105x120), full iteration count in the e2e leg.
  C4            BASELINE.json configs[3]: PVRreconstructionGPU --patchSize 64 64 --patchStride 32 32 on a synthetic 6-stack
                whole-uterus volume (6 x 96 slices of 320x320 at 1 mm, 2.5 mm spacing, into 320x320x240 at 1 mm); the unit is a
                PATCH-projection (one 64x64 patch through one PSF pass).
  C5            BASELINE.json configs[4]: PVR with superpixel patches + the patch-to-volume registration similarity kernel.
  tiny          a seconds-long SVR smoke workload.
One *step*: SVR = one outer iteration of the reference loop (reconstruction.cc:929-1138) over ALL slices: InitializeEMValues,
1 Gaussian reconstruction (K1), 1+4 SimulateSlices (K2), 4 Superresolution (K3 + regulariser), robust statistics, MaskVolume
= 10 slice-projections per slice; PVR = one pass of irtkPatchBasedReconstruction<T>::run()'s loop body
(irtkPatchBasedReconstruction.cpp:490-542): 1 P1 + 8 P2 + 7 P3 = 16 patch-projections per patch.
N > 1: slices / patches are sharded over the ranks (rank r: every N-th of every stack; strong scaling), the interleaved
volume accumulator is all-reduced with NCCL after K1 / P1 and after every K3 / P3.

  python bench.py --gpus 1 --steps 3 --warmup 3 [--workload C3]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus 8 --steps 3 --warmup 3
  python bench.py --impl reference ...      # the CPU arm (restatement of the reference's --useCPU path, oracle/cpu_path.c)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "slice_projections_per_sec"
UNIT = "slice-projections/s"
PROJ_PER_SLICE_STEP = 10          # 1 K1 + 5 K2 + 4 K3 (rec_iterations_first = 4)
PROJ_PER_SLICE_VOLUME = 58        # 4 K1 + 29 K2 + 25 K3
PVR_REC_ITER = 7                  # patchBasedReconMain.cpp default
PROJ_PER_PATCH_STEP = 1 + (1 + PVR_REC_ITER) + PVR_REC_ITER       # 1 P1 + 8 P2 + 7 P3


# ----------------------------------------------------------------------------------------------------
# workloads
def svr_workload(args):
    """(S_global, workload description, phantom config or None, full C2 dataset or None) -- no slice data is generated here."""
    from fetalreconstruction_b200.phantom import c3_config, small_config
    if args.workload == "C2":
        from fetalreconstruction_b200.fixtures import load_c2_setup
        full = load_c2_setup()
        vx, vy, vz = full.cfg.vol_size
        desc = (f"C2: the reference's bundled 4-stack 3T data (14, 10, 21, 23 _3T_nody_001 + mask_10_3T_brain_smooth) at --resolution 1.0 "
                f"after the set-up pipeline: {full.S} slices <= {full.slices.shape[2]}x{full.slices.shape[1]} into {vx}x{vy}x{vz} @ 1.0 mm "
                "(BASELINE.json configs[1]; tests/golden/c2_setup.npz)")
        return full.S, desc, None, full
    if args.workload == "C3":
        cfg = c3_config()
    else:
        cfg = small_config(seed=7, vol=64, n_stacks=8, slices=16, size=64, inplane=1.0, spacing=2.0)
        cfg.name = "tiny"
    vx, vy, vz = cfg.vol_size
    desc = (f"{cfg.name}: synthetic {cfg.n_stacks} stacks x {cfg.slices_per_stack} slices of {cfg.slice_size[0]}x{cfg.slice_size[1]} into "
            f"{vx}x{vy}x{vz} @ {cfg.vol_voxel} mm" + (" (BASELINE.json configs[2])" if cfg.name == "C3" else ""))
    return cfg.n_stacks * cfg.slices_per_stack, desc, cfg, None


def svr_dataset(args, rank=0, world=1, device="cpu"):
    """(dataset of this rank, S_global, workload description, phantom config or None)."""
    from fetalreconstruction_b200.phantom import make_dataset
    S_global, desc, cfg, full = svr_workload(args)
    if full is not None:
        from fetalreconstruction_b200.fixtures import shard_dataset
        return shard_dataset(full, rank, world), S_global, desc, None
    return make_dataset(cfg, device=device, shard=(rank, world)), S_global, desc, cfg


def c4_config():
    """BASELINE.json configs[3] / SURVEY.md section 8d: 6 stacks x 96 slices x 320x320 @ 1.0 mm in plane, 2.5 mm spacing, into a
    320x320x240 volume @ 1.0 mm ("whole uterus")."""
    from fetalreconstruction_b200.phantom import PhantomConfig
    return PhantomConfig(vol_size=(320, 320, 240), vol_voxel=1.0, n_stacks=6, slices_per_stack=96, slice_size=(320, 320), inplane=1.0,
                         spacing=2.5, thickness=None, motion_mm=1.0, motion_deg=1.0, noise=10.0, corrupt_fraction=0.0,
                         mask_semi_axis=0.42, seed=20240601, name="C4")


def config_json(desc, n_gpus, kind="svr"):
    unit = "slices" if kind == "svr" else "patches"
    return {
        "workload": desc,
        "step": ("one outer iteration: 1 K1 + 5 K2 + 4 K3 + regulariser + EM = 10 slice-projections per slice" if kind == "svr" else
                 f"one pass of the PVR loop body: 1 P1 + {1 + PVR_REC_ITER} P2 + {PVR_REC_ITER} P3 + regulariser + EM = "
                 f"{PROJ_PER_PATCH_STEP} patch-projections per patch"),
        "parallelism": f"{unit} sharded over {n_gpus} rank(s) (rank r: every N-th of every stack), NCCL all-reduce of the volume accumulator",
        "l2": "inputs larger than L2: the slice-side arrays of a rank (>1 GB at C3/C4, N=1) and the volume-side buffers (0.6 GB) are "
              "streamed every step" if "C2" not in desc and "tiny" not in desc else
              "working set smaller than L2 at this size (the reference's own test data): an L2 flush (256 MB memset) runs between timed steps",
    }


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------------
# CPU arm: the restated --useCPU path (oracle/cpu_path.c) on the host cores.  It never loads libsvr_b200.so: the host helpers of
# the loop (slice-level EM, small-slice rule, M-step finish) come from the oracle's own copies (SVRPipeline(host=...)).
def cpu_sample_dataset(args, cfg):
    """A bounded sample of the same workload: C3 -> two WHOLE stacks (256 slices: the axis-aligned stack 0 and the oblique
    stack 4) against the full-size volume and mask; C2 / tiny -> the whole workload."""
    from fetalreconstruction_b200.phantom import make_dataset
    if cfg is None or cfg.name != "C3":
        ds, _, _, _ = svr_dataset(args)
        return ds, "the whole workload"
    ds = make_dataset(cfg, device="cpu", stacks=[0, 4])
    return ds, f"{ds.S} slices = two whole stacks (the axis-aligned stack 0 and the oblique stack 4) of the C3 workload, full 256^3 volume"


class _Timed:
    """Sums the wall time of a backend's method calls by name."""

    def __init__(self, b):
        object.__setattr__(self, "_b", b)
        object.__setattr__(self, "times", {})

    def __getattr__(self, name):
        a = getattr(self._b, name)
        if not callable(a):
            return a

        def f(*args, **kw):
            t = time.perf_counter()
            r = a(*args, **kw)
            self.times[name] = self.times.get(name, 0.0) + time.perf_counter() - t
            return r
        return f

    def __setattr__(self, k, v):
        setattr(self._b, k, v)


def cpu_step(ds):
    """One outer iteration (the same 10 slice-projections per slice) of the reference's CPU path as restated in
    oracle/cpu_path.c + cpu_backend.py: CoeffInit, Gaussian reconstruction, 5 SimulateSlices, 4 Superresolution, EM."""
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    from oracle import oracle as orc
    from oracle.cpu_backend import CpuPathReconstruction
    b = _Timed(CpuPathReconstruction())
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(), host=orc)
    p.InitializeEMGPU(ds.slices)
    b.times.clear()
    t0 = time.perf_counter()
    p.outer_iteration(0)
    total = time.perf_counter() - t0
    volume_side = sum(b.times.get(k, 0.0) for k in ("superresolution_finish", "gaussian_reconstruction_finish", "maskVolume"))
    return total, volume_side


def attrs18(a):
    """ImageAttributes -> the 18 doubles irtkImageAttributes is built from (oracle/ref_irtk.py)."""
    return np.concatenate([[a.x, a.y, a.z, a.dx, a.dy, a.dz], np.asarray(a.origin, float), np.asarray(a.xaxis, float),
                           np.asarray(a.yaxis, float), np.asarray(a.zaxis, float)])


def reference_cpu_step(ds, threads):
    """One outer iteration of the REFERENCE'S OWN CPU (--useCPU) path: class irtkReconstruction of irtkReconstructionGPU.cc and the
    vendored IRTK, compiled unmodified into oracle/_ref/libref_irtk.so (TBB stand-in on std::thread).  The call order is the
    useCPU branch of reconstruction.cc:800-1138: InitializeEM, InitializeEMValues, CoeffInit, GaussianReconstruction,
    SimulateSlices, InitializeRobustStatistics, EStep, 4 x (Scale, Superresolution, SimulateSlices, MStep, EStep), MaskVolume."""
    from oracle import ref_irtk as ri
    ri.set_threads(threads)
    r = ri.Reconstruction()
    vol = attrs18(ds.vol_attr) if hasattr(ds, "vol_attr") else None
    if vol is None:                                          # C2 fixture: the volume grid from its matrices
        from fetalreconstruction_b200.geometry import ImageAttributes
        vz, vy, vx = ds.mask.shape
        i2w = np.asarray(ds.recon_i2w, np.float64).reshape(4, 4)
        d = float(ds.cfg.vol_voxel)
        ax = [i2w[:3, c] / d for c in range(3)]
        centre = i2w @ np.array([(vx - 1) / 2.0, (vy - 1) / 2.0, (vz - 1) / 2.0, 1.0])
        vol = attrs18(ImageAttributes(vx, vy, vz, d, d, d, centre[:3], ax[0], ax[1], ax[2]))
    r.set_reconstructed(ri.Image.new(vol))
    r.set_mask(ri.Image.new(vol, ds.mask.astype(np.float64)), 0.0)
    imgs = []
    for k in range(ds.S):
        a = ds.slice_attrs[k]
        sx, sy = a.x, a.y
        imgs.append(ri.Image.new(attrs18(a), ds.slices[k, :sy, :sx].astype(np.float64)))
    dofs = np.stack([ri.rigid_from_matrix(np.asarray(t, np.float64).reshape(4, 4)) for t in ds.trans])
    r.set_slices(imgs, dofs, np.asarray(ds.stack_index, np.int32), np.asarray(ds.dims[:, 2], np.float64))
    r.cpu_step("InitializeEM")
    times = {}

    def step(name, it=0):
        t = time.perf_counter()
        r.cpu_step(name, it)
        times[name] = times.get(name, 0.0) + time.perf_counter() - t
    t0 = time.perf_counter()
    r.call("set_smoothing_parameters", 150.0, 0.08)         # iteration 0 of the default schedule (reconstruction.cc:900-911)
    r.call("speedup", 1)
    step("InitializeEMValues"); step("CoeffInit"); step("GaussianReconstruction"); step("SimulateSlices")
    step("InitializeRobustStatistics"); step("EStep")
    for i in range(4):
        step("Scale"); step("Superresolution", i + 1); step("SimulateSlices"); step("MStep", i + 1); step("EStep")
    step("MaskVolume")
    total = time.perf_counter() - t0
    vol_out = r.reconstructed().data
    return total, times, bool(np.isfinite(vol_out).all() and (vol_out > 0).any())


def run_cpu_baseline(args, cfg, S_full, steps=1, warmup=0):
    """The CPU arm: the reference's own CPU path when oracle/_ref/libref_irtk.so is there (kind "reference"), else our restatement
    of it (kind "port", oracle/cpu_path.c)."""
    from oracle import ref_irtk as ri
    if ri.available() and not os.environ.get("SVR_BENCH_CPU_PORT"):
        threads = host_threads()
        ds, sample = cpu_sample_dataset(args, cfg)
        for _ in range(warmup):
            reference_cpu_step(ds, threads)
        runs = [reference_cpu_step(ds, threads) for _ in range(max(steps, 1))]
        t = float(np.mean([r[0] for r in runs]))
        times = runs[-1][1]
        # volume-side work that does not grow with the slices: the regulariser inside Superresolution cannot be separated by timing
        # the calls, so the extrapolation scales EVERYTHING with the slices (conservative for the CPU: it over-counts its time)
        t_full = t * S_full / ds.S
        sampled = ds.S != S_full
        return {"value": S_full * PROJ_PER_SLICE_STEP / t_full, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": f"{sample}; one outer iteration (10 slice-projections per slice) of the reference's OWN CPU (--useCPU) path -- class "
                          "irtkReconstruction (irtkReconstructionGPU.cc: CoeffInit, GaussianReconstruction, SimulateSlices, EStep, Scale, "
                          "Superresolution, MStep) + the vendored IRTK, compiled unmodified into oracle/_ref/libref_irtk.so; its TBB "
                          f"parallel_for / parallel_reduce run on a std::thread stand-in with {threads} threads -- took {t:.1f} s"
                          + (f"; `value` is the full-workload figure, time x {S_full}/{ds.S} = {t_full:.1f} s per step" if sampled else ""),
                "seconds_per_step_sample": t, "seconds_per_step_full_workload": t_full, "finite": runs[-1][2],
                "seconds_by_call": {k: round(v, 2) for k, v in times.items()}}, t_full
    from oracle import oracle as orc
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its thread count itself
    orc.set_num_threads(host_threads())
    ds, sample = cpu_sample_dataset(args, cfg)
    for _ in range(warmup):
        cpu_step(ds)
    runs = [cpu_step(ds) for _ in range(max(steps, 1))]
    t = float(np.mean([r[0] for r in runs]))
    tv = float(np.mean([r[1] for r in runs]))
    t_full = tv + (t - tv) * S_full / ds.S               # volume-side work (regulariser, masking) does not grow with the slices
    sampled = ds.S != S_full
    value = S_full * PROJ_PER_SLICE_STEP / t_full
    return {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{sample}; one outer iteration (10 slice-projections per slice) took {t:.1f} s, of which {tv:.1f} s volume-side "
                      "(regulariser, masking) that does not grow with the number of slices"
                      + (f"; `value` is the FULL-workload figure: volume-side time + slice-side time x {S_full}/{ds.S} = {t_full:.1f} s per step "
                         f"(the sample alone: {ds.S * PROJ_PER_SLICE_STEP / t:.1f} {UNIT})" if sampled else "")
                      + ".  The port is the reference's CPU (--useCPU) formulation -- sparse slice-to-volume matrix of CoeffInit (Gaussian PSF, "
                      "trilinear splat, irtkReconstructionGPU.cc:2305-2673) + the functors applying it, OpenMP where the reference uses TBB "
                      "(oracle/cpu_path.c); its IRTK/TBB original cannot be built here.  Like the original's parallel_reduce, the "
                      "super-resolution keeps one addon + confidence-map volume per thread and joins them, and Gaussian reconstruction and the "
                      "regulariser's prep are serial: the arm scales ~4x on 16 threads, by construction of the reference's algorithm",
            "seconds_per_step_sample": t, "seconds_per_step_full_workload": t_full}, t_full


def run_reference_cuda(cfg, reps=3):
    """The reference's OWN CUDA path (oracle/_ref: its unmodified sources recompiled for sm_100a) timed on this box's GPU on the
    FULL C3 workload (all 8 stacks, 1024 slices), one outer iteration (the same 10 slice-projections per slice), median of `reps`
    after one untimed warm-up iteration; then one stack (= one orientation) at a time.  Runs in a subprocess (the reference resets
    the device and keeps process-global state).  A reported baseline."""
    if cfg is None or cfg.name != "C3" or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cuda2.so")):
        return None
    tool = os.path.join(ROOT, "tools", "ref_bench.py")

    def one(stack_list, reps_):
        subprocess.run([sys.executable, tool, "gen", "--stack-list", stack_list], check=True, capture_output=True, text=True, timeout=600)
        r = subprocess.run([sys.executable, tool, "ref", "--rec-iters", "4", "--timing-only", "--reps", str(reps_)], check=True,
                           capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("REFBENCH_JSON ")][-1]
        return json.loads(line[len("REFBENCH_JSON "):])
    try:
        d = one(",".join(str(i) for i in range(cfg.n_stacks)), reps)
        out = {"value": d["S"] * PROJ_PER_SLICE_STEP / d["iteration_s"], "unit": UNIT, "kind": "reference CUDA path on this GPU",
               "sample": f"the full workload: {d['S']} slices (all {cfg.n_stacks} stacks), full 256^3 volume; median of {reps} outer iterations after "
                         f"one untimed warm-up iteration: {d['iteration_s']:.3f} s wall (all: {', '.join('%.3f' % v for v in d['iterations_s'])} s); "
                         "every call synchronous as in the reference",
               "ms_per_call": {k: round(v, 3) for k, v in d["ms_per_call"].items() if v >= 0.05}}
    except Exception as e:                                   # the baseline is optional; the bench line is not
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}
    try:
        per = {}
        for st in range(cfg.n_stacks):
            ds_ = one(str(st), 1)
            mc = ds_["ms_per_call"]
            grp = lambda *keys: round(sum(mc.get(k, 0.0) for k in keys), 2)
            per[str(st)] = {"seconds_per_outer_iteration": round(ds_["iteration_s"], 4),
                            "GaussianReconstruction_ms": grp("GaussianReconstruction", "gaussian_reconstruction_local", "gaussian_reconstruction_finish"),
                            "SimulateSlices_ms": grp("SimulateSlices"),
                            "Superresolution_ms": grp("Superresolution", "superresolution_local", "superresolution_finish")}
        out["per_stack"] = per
    except Exception as e:
        out["per_stack"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return out


def ours_per_stack(cfg):
    """Our per-call wall times one stack (= one orientation) at a time, every call synchronous: the like-for-like twin of
    reference_cuda.per_stack."""
    import torch
    from fetalreconstruction_b200.phantom import make_dataset
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    from fetalreconstruction_b200.reconstruction import Reconstruction
    out = {}
    for st in range(cfg.n_stacks):
        ds = make_dataset(cfg, device="cuda", stacks=[st])
        b = Reconstruction(torch.cuda.current_device())
        upload_dataset(b, ds)
        p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
        p.InitializeEMGPU(ds.slices)
        p.outer_iteration(0)
        b.profile_reset(); b.profile_enable(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        p.outer_iteration(0)
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        b.profile_enable(False)
        prof = b.profile_read()
        a = ds.stack_attrs[0]
        out[str(st)] = {"slice_x_axis_in_volume": [round(float(v), 3) for v in a.xaxis], "slice_y_axis_in_volume": [round(float(v), 3) for v in a.yaxis],
                        "seconds_per_outer_iteration": round(wall, 4),
                        "K1_ms": round(prof["gaussian"][0] / max(prof["gaussian"][1], 1), 3),
                        "K2_ms": round(prof["simulate"][0] / max(prof["simulate"][1], 1), 3),
                        "K3_ms": round(prof["superres"][0] / max(prof["superres"][1], 1), 3)}
        b.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload in ("C4", "C5"):
        print(json.dumps({"impl": "reference", "metric": "patch_projections_per_sec", "value": None, "unit": "patch-projections/s",
                          "unavailable": "the reference has no CPU path for PVR (PVRreconstructionGPU is GPU-only, patchBasedReconMain.cpp:177-179); "
                                         "bench.py --workload C4 reports the reference's own PVR CUDA code on this GPU as `reference_cuda`"}))
        return
    S_full, desc, cfg, _ = svr_workload(args)
    base, t = run_cpu_baseline(args, cfg, S_full, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic" if cfg is not None else "bundled 3T stacks (fixture)",
            "config": config_json(desc, args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's IRTK/TBB CPU path cannot be compiled here (Boost/TBB/GSL absent, SURVEY.md 8c): this arm "
                    "times our restatement of that path (oracle/cpu_path.c: CoeffInit's sparse matrix + the functors applying "
                    f"it) on {base['cores']} host threads (set explicitly; the launcher's OMP_NUM_THREADS is ignored); warm-up capped at 1 step "
                    "(deterministic CPU code); ms_per_step is the full-workload figure `value` is computed from.  The reference's own CUDA "
                    "path does compile (oracle/_ref) and is reported by the main arm as `reference_cuda`."}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    return world, rank, local, dev, group


def allreduce_scalar(x, op, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=op)
    return float(t.item())


def kernel_table(prof, alg_bytes, names, ms_total):
    return {names.get(k, k): {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1],
                              "share_of_step": v[0] / ms_total if ms_total else 0.0,
                              "alg_bytes_per_launch": alg_bytes.get(k),
                              "alg_GBps": (alg_bytes[k] / (v[0] / max(v[1], 1) * 1e-3) / 1e9) if k in alg_bytes and v[0] > 0 else None}
            for k, v in prof.items() if v[1] > 0}


def timed_steps(args, step, comm, stream, backend, rank, local, small_working_set):
    """W warm-up steps, then K timed ones between CUDA events on the launching stream (max over ranks taken by the caller)."""
    import torch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}") if small_working_set else None
    for _ in range(max(args.warmup, 0)):
        step()
    comm.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    backend.profile_reset()
    backend.profile_enable(True)
    launches0 = backend.launch_count
    ms_total = 0.0
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        comm.barrier()
        torch.cuda.synchronize()
        ms_total = ev0.elapsed_time(ev1)
    else:
        for _ in range(args.steps):                       # working set < L2: flush L2 between the timed steps (outside the events)
            flush.zero_()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            step()
            ev1.record(stream)
            torch.cuda.synchronize()
            ms_total += ev0.elapsed_time(ev1)
        comm.barrier()
    clocks = sampler.stop() if rank == 0 else None
    backend.profile_enable(False)
    return ms_total, backend.launch_count - launches0, backend.profile_read(), clocks


# ----------------------------------------------------------------------------------------------------
def bench_svr(args):
    import torch
    import torch.distributed as dist
    from fetalreconstruction_b200 import build
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, Comm, upload_dataset
    from fetalreconstruction_b200.reconstruction import Reconstruction

    world, rank, local, dev, group = init_dist()
    comm = Comm(group, dev, stream_ordered=True)          # the library runs on the torch stream below (svr_set_stream)
    build.build()

    # Sharding: rank r takes every N-th slice (j % N == r) of EVERY stack, so all ranks see the same mix of stack
    # orientations and of positions along the stacks -- the per-slice cost of the PSF kernels depends on the orientation
    # (whole stacks per rank left rank 0 22 % slower than rank 1 at N=2) and the number of valid pixels on the distance
    # from the stack's ends (contiguous quarters of every stack: 60 % efficiency at N=4).  Global slice order = rank-major
    # (a permutation of the acquisition order; the slice-level EM is order-independent).
    ds, S_global, desc, cfg = svr_dataset(args, rank, world, device=str(dev))
    counts = [0] * world
    counts[rank] = ds.S
    if world > 1:
        t = torch.tensor(counts, dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        counts = [int(v) for v in t.tolist()]
    assert sum(counts) == S_global, (counts, S_global)
    b0 = sum(counts[:rank]); e0 = b0 + counts[rank]
    small_ws = args.workload in ("C2", "tiny")

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        backend = Reconstruction(local)
        backend.set_stream(stream.cuda_stream)
        backend.set_async(True)                           # calls that return no host data only enqueue (svr_set_async)
        for key, val in args.tune:
            backend.set_tuning(key, val)
        upload_dataset(backend, ds)
        acc = torch.as_tensor(backend.accumulator(), device=dev)
        pipe = SVRPipeline(backend, S_global, b0, e0, comm, SVRParams(), accumulator_tensor=lambda: acc)
        pipe.InitializeEMGPU(ds.slices)

        ms_total, launches, prof, clocks = timed_steps(args, lambda: pipe.outer_iteration(0), comm, stream, backend, rank, local, small_ws)

        # every rank must hold the same volume replica after a step (the accumulators were all-reduced, the rest is redundant)
        chk = float(np.abs(backend.syncCPU().astype(np.float64)).sum())
        chk_lo = allreduce_scalar(chk, dist.ReduceOp.MIN, dev, world)
        chk_hi = allreduce_scalar(chk, dist.ReduceOp.MAX, dev, world)
        replicas_identical = bool(chk_lo == chk_hi) and bool(np.isfinite(chk))
        ms_total = allreduce_scalar(ms_total, dist.ReduceOp.MAX, dev, world)
        n_launch = allreduce_scalar(launches, dist.ReduceOp.SUM, dev, world)
        ms_step = ms_total / args.steps
        value = S_global * PROJ_PER_SLICE_STEP / (ms_step * 1e-3)

        # ---- roofline of the dominant kernel (algorithmic bytes) --------------
        NP, V = backend.NP, backend.V
        alg_bytes = {"gaussian": 12 * NP + 20 * V, "simulate": 17 * NP + 8 * V, "superres": 16 * NP + 20 * V, "regularize": 28 * V,
                     "estep": 16 * NP, "mstep": 16 * NP, "scale": 16 * NP, "robust_init": 13 * NP}
        names = {"gaussian": "gaussian_sume_kernel + gaussian_scatter_kernel (K1)", "simulate": "simulate_kernel (K2)",
                 "superres": "superres_scatter_kernel (K3)", "regularize": "regularize_fused_kernel (K4+K5)", "estep": "estep_kernel (K7+K8)",
                 "mstep": "mstep_kernel (K9)", "scale": "scale_kernel (K10)", "robust_init": "robust_init_kernel (K11)"}
        dom = max(("gaussian", "simulate", "superres"), key=lambda k: prof[k][0])
        dom_ms = prof[dom][0] / max(prof[dom][1], 1)
        peak, peak_src = measured_peak()
        achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = recorded_traffic()
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": dom_ms,
                    "limiter": "fp32 ALU + MUFU (4096 sinc^2*gauss taps per pixel); HBM is not binding -- see DESIGN.md",
                    "kernels": kernel_table(prof, alg_bytes, names, ms_total),
                    "note": "alg_GBps uses SURVEY.md 8d's bytes over the PADDED slice cube; the robust-statistics kernels only touch the "
                            "pixels that are not padding (27 % at C3), so their figure can exceed the HBM peak"}
        # what does bound the PSF kernels: the warp schedulers' issue slots.  Tap count of one K2 launch (pixels that carry a
        # PSF sum x 16^3 taps) x the SASS instructions per tap of the per-pixel loop (349 per 16-tap row incl. the 16 loads and 32
        # FFMAs of the forward projection, profiles/r02_sass_k2_row.txt) against 4 issue slots per SM per clock.
        k2_ms = prof["simulate"][0] / max(prof["simulate"][1], 1)
        if k2_ms > 0:
            n_px = int(np.count_nonzero(backend.debugv_PSF_sums()))
            sm_mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
            warp_instr_per_s = n_px * 4096.0 * (349.0 / 16.0) / 32.0 / (k2_ms * 1e-3)
            issue_peak = 148 * 4 * sm_mhz * 1e6
            roofline["issue_slots"] = {"kernel": "simulate_kernel (K2)", "pixels_per_launch": n_px, "taps_per_s": n_px * 4096.0 / (k2_ms * 1e-3),
                                       "sass_instr_per_tap": 349.0 / 16.0, "warp_instr_per_s": warp_instr_per_s,
                                       "peak_warp_instr_per_s": issue_peak, "frac": warp_instr_per_s / issue_peak,
                                       "sm_mhz": sm_mhz}

        # ---- e2e: the full default schedule through host buffers ---------------------------------------
        e2e = None
        vph = None
        if not args.no_e2e:
            pinned = torch.from_numpy(np.ascontiguousarray(ds.slices)).pin_memory()
            cube = pinned.numpy()
            vol_host = torch.empty(backend.V, dtype=torch.float32).pin_memory().numpy()       # the caller's (pinned) output buffer
            h2d = d2h = 0
            comm.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            backend.FillSlices(cube.ravel())
            h2d += cube.nbytes
            pipe2 = SVRPipeline(backend, S_global, b0, e0, comm, SVRParams(), accumulator_tensor=lambda: acc)
            pipe2.InitializeEMGPU(ds.slices)
            for it in range(pipe2.p.iterations):
                backend.SetSliceMatrices(ds.trans, ds.trans_inv, ds.i2w, ds.w2i, ds.i2w, ds.w2i, ds.recon_i2w, ds.recon_w2i)
                h2d += 4 * ds.trans.nbytes
                pipe2.outer_iteration(it)
                vol = backend.syncCPU(out=vol_host)           # image<iter>_GPU.nii.gz (reconstruction.cc:1189-1193)
                d2h += vol.nbytes
            pipe2.ScaleVolumeGPU()
            vol = backend.syncCPU(out=vol_host)
            d2h += vol.nbytes
            comm.barrier()
            torch.cuda.synchronize()
            wall = allreduce_scalar(time.perf_counter() - t0, dist.ReduceOp.MAX, dev, world)
            n_outer = pipe2.p.iterations
            e2e = {"value": S_global * PROJ_PER_SLICE_VOLUME / wall, "unit": UNIT,
                   "h2d_bytes_per_step": int(h2d / n_outer), "d2h_bytes_per_step": int(d2h / n_outer),
                   "what": "full default schedule (4 outer iterations, 58 slice-projections per slice) through the C ABI "
                           "with pinned HOST buffers: FillSlices + per-iteration SetSliceMatrices / syncCPU inside the "
                           "timed region; a 'step' here is one outer iteration",
                   "seconds_per_volume": wall, "finite": bool(np.isfinite(vol).all())}
            vph = 3600.0 / wall

        # ---- slice-to-volume registration (a10): similarity-kernel throughput and one full registration call ----
        registration = None
        if not args.no_registration:
            from fetalreconstruction_b200.registration import RegistrationFrontEnd
            t0 = time.perf_counter()
            fe = RegistrationFrontEnd(backend, ds.slices, ds.slice_attrs, ds.cfg.vol_voxel)
            prep_s = time.perf_counter() - t0
            backend.updateResampledSlicesI2W(fe.ofs)
            backend.prepareSliceToVolumeReg()
            tr = fe.pack_transforms(ds.trans)
            backend.evaluateCostsMultipleSlices(tr, 0)                 # warm-up
            backend.profile_reset(); backend.profile_enable(True)
            n_eval = 5
            for _ in range(n_eval):
                backend.evaluateCostsMultipleSlices(tr, 0)
            backend.profile_enable(False)
            ev_ms = backend.profile_read()["reg_eval"][0] / n_eval
            W, H, Sl = backend.regW, backend.regH, backend.regS
            alg = 3 * 8 * W * H * Sl                                   # 8*W*H per cost evaluation per z-offset (fused)
            backend.setRegSchedule(2, 4, 20)
            comm.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            backend.registerSlicesToVolume(tr)
            torch.cuda.synchronize()
            reg_s = allreduce_scalar(time.perf_counter() - t0, dist.ReduceOp.MAX, dev, world)
            registration = {"kernel": "reg_eval_kernel (fused sample + blur + NCC moments)", "ms_per_cost_evaluation": ev_ms,
                            "slices": Sl, "slice_size": [W, H], "alg_bytes_per_evaluation": alg,
                            "alg_GBps": alg / (ev_ms * 1e-3) / 1e9 if ev_ms > 0 else None,
                            "full_registration_s": reg_s, "cost_evaluations_slice_offsets": backend.reg_evaluations,
                            "slices_registered_per_s": S_global / reg_s, "host_prep_s": prep_s,
                            "schedule": "reference default: 2 levels x 4 steps x <=20 iterations, epsilon 1e-4"}
        backend.close()

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = run_cpu_baseline(args, cfg, S_global, steps=1, warmup=0)
    ref_cuda = None
    per_stack = None
    if rank == 0 and world == 1 and not args.no_reference_cuda and cfg is not None and cfg.name == "C3":
        ref_cuda = run_reference_cuda(cfg)
        if ref_cuda and "per_stack" in ref_cuda:
            per_stack = ours_per_stack(cfg)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic" if cfg is not None else "bundled 3T stacks (fixture)",
                "config": config_json(desc, world), "volumes_per_hour": vph, "roofline": roofline, "cpu_baseline": base,
                "reference_cuda": ref_cuda, "ours_per_stack": per_stack, "e2e": e2e,
                "registration": registration, "gpu_launches": int(n_launch), "clocks": clocks,
                "replicas_identical": replicas_identical, "tuning": dict(args.tune)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def bench_pvr(args):
    """C4: PVR patches 64x64 / stride 32 (BASELINE.json configs[3]); C5 adds superpixel masks and the patch registration kernel."""
    import torch
    import torch.distributed as dist
    from fetalreconstruction_b200 import build
    from fetalreconstruction_b200.pipeline import Comm
    from fetalreconstruction_b200.pvr import PatchReconstruction, PVRParams, PVRPipeline
    from fetalreconstruction_b200.pvr_case import make_pvr_case, setup_backend, shard_case

    world, rank, local, dev, group = init_dist()
    comm = Comm(group, dev)
    build.build()
    cfg = c4_config()
    if args.pvr_slices:
        cfg.slices_per_stack = int(args.pvr_slices)
    t0 = time.perf_counter()
    case = make_pvr_case(cfg=cfg, pbb=(64, 64), stride=(32, 32), device=str(dev))
    enum_s = time.perf_counter() - t0
    ds = case["ds"]
    n_global = len(case["attrs"])
    sub, gidx = shard_case(case, rank, world)
    vx, vy, vz = cfg.vol_size
    desc = (f"C4: PVR, synthetic {cfg.n_stacks} stacks x {cfg.slices_per_stack} slices of {cfg.slice_size[0]}x{cfg.slice_size[1]} @ {cfg.inplane} mm, "
            f"{cfg.spacing} mm spacing, into {vx}x{vy}x{vz} @ {cfg.vol_voxel} mm; --patchSize 64 64 --patchStride 32 32: {n_global} patches after the "
            "1/3-coverage rule (BASELINE.json configs[3])")
    params = PVRParams(iterations=0, rec_iterations=PVR_REC_ITER)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        b = PatchReconstruction(local)
        b.set_stream(stream.cuda_stream)
        setup_backend(b, sub, device_patch_init=False)
        pipe = PVRPipeline(b, ds.min_intensity, ds.max_intensity, params, comm=comm if world > 1 else None, global_index=gidx,
                           patches_per_stack_global=case["per_stack"])
        ms_total, launches, prof, clocks = timed_steps(args, pipe.iteration, comm, stream, b, rank, local, False)
        chk = float(np.abs(b.recon_copyToHost().astype(np.float64)).sum())
        replicas_identical = allreduce_scalar(chk, dist.ReduceOp.MIN, dev, world) == allreduce_scalar(chk, dist.ReduceOp.MAX, dev, world) and np.isfinite(chk)
        ms_total = allreduce_scalar(ms_total, dist.ReduceOp.MAX, dev, world)
        n_launch = allreduce_scalar(launches, dist.ReduceOp.SUM, dev, world)
        ms_step = ms_total / args.steps
        value = n_global * PROJ_PER_PATCH_STEP / (ms_step * 1e-3)
        NP, V = b.NP, b.V
        alg_bytes = {"gaussian": 12 * NP + 20 * V, "simulate": 17 * NP + 8 * V, "superres": 16 * NP + 20 * V, "regularize": 28 * V,
                     "estep": 16 * NP, "mstep": 16 * NP, "scale": 16 * NP, "robust_init": 13 * NP}
        names = {"gaussian": "P1 gaussian_sume_kernel + gaussian_scatter_kernel <PvrTraits>", "simulate": "P2 simulate_kernel <PvrTraits>",
                 "superres": "P3 superres_scatter_kernel <PvrTraits>", "regularize": "P4 regularize_fused_kernel", "estep": "P6 estep_kernel",
                 "mstep": "P7 mstep_kernel", "scale": "P7 scale_kernel", "robust_init": "P7 robust_init_kernel"}
        dom = max(("gaussian", "simulate", "superres"), key=lambda k: prof[k][0])
        dom_ms = prof[dom][0] / max(prof[dom][1], 1)
        peak, peak_src = measured_peak()
        achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": dom_ms,
                    "limiter": "fp32 ALU + MUFU (1728 sinc^2*gauss taps per patch pixel); HBM is not binding -- see DESIGN.md",
                    "kernels": kernel_table(prof, alg_bytes, names, ms_total)}
        # ---- e2e: patch cube from pinned host memory, one pass, volume back to the host --------------------------------------
        e2e = None
        if not args.no_e2e:
            cube = torch.from_numpy(np.ascontiguousarray(sub["cube"])).pin_memory().numpy()
            comm.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            b.patches_copyFromHost(cube)
            b.patches_set_matrices(sub["i2w"], sub["w2i"], sub["T"], sub["Tinv"])
            pipe2 = PVRPipeline(b, ds.min_intensity, ds.max_intensity, params, comm=comm if world > 1 else None, global_index=gidx,
                                patches_per_stack_global=case["per_stack"])
            pipe2.iteration()
            vol = b.recon_copyToHost()
            comm.barrier(); torch.cuda.synchronize()
            wall = allreduce_scalar(time.perf_counter() - t0, dist.ReduceOp.MAX, dev, world)
            e2e = {"value": n_global * PROJ_PER_PATCH_STEP / wall, "unit": "patch-projections/s",
                   "h2d_bytes_per_step": int(cube.nbytes + 4 * sub["i2w"].nbytes), "d2h_bytes_per_step": int(vol.nbytes),
                   "what": "one pass of the PVR loop through the C ABI with HOST buffers: the patch cube (pinned) and the four matrix arrays go "
                           "up, the reconstructed volume comes back, inside the timed region", "seconds_per_pass": wall,
                   "finite": bool(np.isfinite(vol).all())}
        b.close()
    if rank == 0:
        line = {"metric": "patch_projections_per_sec", "value": value, "unit": "patch-projections/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_json(desc, world, "pvr"), "roofline": roofline, "cpu_baseline": None,
                "e2e": e2e, "gpu_launches": int(n_launch), "clocks": clocks, "replicas_identical": bool(replicas_identical),
                "patches": n_global, "patch_enumeration_s": enum_s,
                "note": "the reference has no CPU path for PVR (patchBasedReconMain.cpp:177-179 is GPU-only): no cpu_baseline; per-call times "
                        "of the reference's own PVR CUDA code on this GPU: profiles/r01_v6_refbench_pvr_ours_vs_reference.json"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "C2", "C4", "tiny"])
    ap.add_argument("--pvr-slices", type=int, default=0, help="C4: slices per stack (default: the full 96)")
    ap.add_argument("--tune", action="append", default=[], help="KEY=VALUE for svr_set_tuning (A/B runs), e.g. 0=2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-registration", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    args = ap.parse_args()
    args.tune = [tuple(int(v) for v in t.split("=")) for t in args.tune]
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "C4":
        return bench_pvr(args)
    return bench_svr(args)


if __name__ == "__main__":
    main()
