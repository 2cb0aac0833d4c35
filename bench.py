#!/usr/bin/env python
"""bench.py -- SVR hot-path throughput on B200 (BASELINE.json metric: slice-projections/s, volumes/hour,
PSF-kernel HBM GB/s against the measured peak).

Workload (config.workload): BASELINE.json configs[2] = "C3": synthetic 8 stacks x 128 slices of 256x256
into a 256^3 volume at 0.75 mm.  One *step* = one outer iteration of the
reference loop (reconstruction.cc:929-1138) over ALL slices: InitializeEMValues, 1 Gaussian
reconstruction (K1), 1+4 SimulateSlices (K2), 4 Superresolution (K3 + regulariser), robust statistics,
MaskVolume -> 10 slice-projections per slice.  N > 1: stacks are sharded over ranks (strong scaling),
the interleaved accumulator is all-reduced with NCCL after K1 and after every K3.

  python bench.py --gpus 1 --steps 3 --warmup 3
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


def workload_config(args):
    from fetalreconstruction_b200.phantom import c3_config, small_config
    if args.workload == "C3":
        return c3_config()
    cfg = small_config(seed=7, vol=64, n_stacks=8, slices=16, size=64, inplane=1.0, spacing=2.0)
    cfg.name = "tiny"
    return cfg


def config_json(cfg, n_gpus):
    vx, vy, vz = cfg.vol_size
    return {
        "workload": f"{cfg.name}: synthetic {cfg.n_stacks} stacks x {cfg.slices_per_stack} slices of "
                    f"{cfg.slice_size[0]}x{cfg.slice_size[1]} into {vx}x{vy}x{vz} @ {cfg.vol_voxel} mm "
                    "(BASELINE.json configs[2])",
        "step": "one outer iteration: 1 K1 + 5 K2 + 4 K3 + regulariser + EM = 10 slice-projections per slice",
        "slices": cfg.n_stacks * cfg.slices_per_stack,
        "parallelism": f"slices sharded over {n_gpus} rank(s) (rank r: every N-th slice of every stack), NCCL all-reduce of "
                       "the volume accumulator",
        "l2": "inputs larger than L2: per rank the slice-side arrays are >1 GB at N=1 and the volume-side "
              "buffers 0.6 GB, all streamed every step",
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


# ----------------------------------------------------------------------------------------------------
def cpu_sample_dataset(cfg, per_stack=32):
    """A bounded sample of the same workload for the CPU arm: `per_stack` mid-stack slices of one axis-aligned and one
    oblique stack against the full-size volume and mask."""
    from fetalreconstruction_b200.phantom import make_dataset
    st = [0, min(4, cfg.n_stacks - 1)]
    ds = make_dataset(cfg, device="cpu", stacks=st)
    per_stack = min(per_stack, cfg.slices_per_stack)
    lo = (cfg.slices_per_stack - per_stack) // 2
    keep = np.concatenate([np.arange(lo, lo + per_stack), cfg.slices_per_stack + np.arange(lo, lo + per_stack)])
    for name in ("slices", "i2w", "w2i", "trans", "trans_inv", "dims", "stack_index"):
        setattr(ds, name, np.ascontiguousarray(getattr(ds, name)[keep]))
    return ds


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
    from oracle.cpu_backend import CpuPathReconstruction
    b = _Timed(CpuPathReconstruction())
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
    p.InitializeEMGPU(ds.slices)
    b.times.clear()
    t0 = time.perf_counter()
    p.outer_iteration(0)
    total = time.perf_counter() - t0
    volume_side = sum(b.times.get(k, 0.0) for k in ("superresolution_finish", "gaussian_reconstruction_finish", "maskVolume"))
    return total, volume_side


def run_cpu_baseline(cfg, steps=1, warmup=0):
    from oracle import oracle as orc
    ds = cpu_sample_dataset(cfg)
    for _ in range(warmup):
        cpu_step(ds)
    runs = [cpu_step(ds) for _ in range(max(steps, 1))]
    t = float(np.mean([r[0] for r in runs]))
    tv = float(np.mean([r[1] for r in runs]))
    S_full = cfg.n_stacks * cfg.slices_per_stack
    t_full = tv + (t - tv) * S_full / ds.S               # volume-side work (regulariser, masking) does not grow with the slices
    return {"value": ds.S * PROJ_PER_SLICE_STEP / t, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
            "sample": f"{ds.S} mid-stack slices ({ds.S // 2} of an axis-aligned + {ds.S // 2} of an oblique stack) of the {cfg.name} workload, full "
                      f"{cfg.vol_size[0]}^3 volume, one outer iteration (10 slice-projections per slice), {t:.1f} s per step, of which "
                      f"{tv:.1f} s volume-side (regulariser) that does not grow with the number of slices.  The port is the reference's "
                      "CPU (--useCPU) formulation -- sparse slice-to-volume matrix of CoeffInit (Gaussian PSF, trilinear splat, "
                      "irtkReconstructionGPU.cc:2305-2673) + the functors applying it, OpenMP where the reference uses TBB "
                      "(oracle/cpu_path.c); its IRTK/TBB original cannot be built here",
            "extrapolated_full_workload": {"value": S_full * PROJ_PER_SLICE_STEP / t_full, "unit": UNIT,
                                           "how": f"volume-side time + slice-side time x {S_full}/{ds.S}"}}, t


def run_reference_cuda(cfg, stacks=1):
    """The reference's OWN CUDA path (oracle/_ref: its unmodified sources recompiled for sm_100a) timed on this box's GPU on a
    bounded sample of the workload: `stacks` whole stacks, one outer iteration (the same 10 slice-projections per slice).
    Runs in a subprocess (the reference resets the device and keeps process-global state).  A reported baseline."""
    if cfg.name != "C3" or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_cuda2.so")):
        return None
    tool = os.path.join(ROOT, "tools", "ref_bench.py")
    try:
        subprocess.run([sys.executable, tool, "gen", "--stacks", str(stacks)], check=True, capture_output=True, text=True, timeout=300)
        r = subprocess.run([sys.executable, tool, "ref", "--rec-iters", "4", "--timing-only"], check=True, capture_output=True, text=True,
                           timeout=600)
        line = [l for l in r.stdout.splitlines() if l.startswith("REFBENCH_JSON ")][-1]
        d = json.loads(line[len("REFBENCH_JSON "):])
    except Exception as e:                                   # the baseline is optional; the bench line is not
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return {"value": d["S"] * PROJ_PER_SLICE_STEP / d["iteration_s"], "unit": UNIT, "kind": "reference CUDA path on this GPU",
            "sample": f"{d['S']} slices ({stacks} stack(s) of the C3 workload, full 256^3 volume), the fastest of three outer iterations "
                      f"after one untimed warm-up iteration, {d['iteration_s']:.3f} s wall (all: "
                      f"{', '.join('%.3f' % v for v in d.get('iterations_s', [d['iteration_s']]))} s; the reference's per-call "
                      "cudaMalloc/cudaFree make its Superresolution vary); every call synchronous as in the reference",
            "ms_per_call": {k: round(v, 3) for k, v in d["ms_per_call"].items() if v >= 0.05}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args)
    base, t = run_cpu_baseline(cfg, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_json(cfg, args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's IRTK/TBB CPU path cannot be compiled here (Boost/TBB/GSL absent, SURVEY.md 8c): this arm "
                    "times our restatement of that path (oracle/cpu_path.c: CoeffInit's sparse matrix + the functors applying "
                    "it) on the host cores; warm-up capped at 1 step (deterministic CPU code).  The reference's own CUDA path "
                    "does compile (oracle/_ref) and is reported by the main arm as `reference_cuda`."}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=["C3", "tiny"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-registration", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fetalreconstruction_b200 import build
    from fetalreconstruction_b200.phantom import make_dataset
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, Comm, upload_dataset
    from fetalreconstruction_b200.reconstruction import Reconstruction, host_partition

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
    comm = Comm(group, dev)
    build.build()

    cfg = workload_config(args)
    S_global = cfg.n_stacks * cfg.slices_per_stack
    # Sharding: rank r takes every N-th slice (j % N == r) of EVERY stack, so all ranks see the same mix of stack
    # orientations and of positions along the stacks -- the per-slice cost of the PSF kernels depends on the orientation
    # (whole stacks per rank left rank 0 22 % slower than rank 1 at N=2) and the number of valid pixels on the distance
    # from the stack's ends (contiguous quarters of every stack: 60 % efficiency at N=4).  Global slice order = rank-major
    # (a permutation of the acquisition order; the slice-level EM is order-independent).
    from fetalreconstruction_b200.phantom import shard_count
    per_rank = [cfg.n_stacks * shard_count(cfg.slices_per_stack, r, world) for r in range(world)]
    b0 = sum(per_rank[:rank]); e0 = b0 + per_rank[rank]
    ds = make_dataset(cfg, device=str(dev), shard=(rank, world))
    assert ds.S == e0 - b0

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        backend = Reconstruction(local)
        backend.set_stream(stream.cuda_stream)
        upload_dataset(backend, ds)
        acc = torch.as_tensor(backend.accumulator(), device=dev)
        pipe = SVRPipeline(backend, S_global, b0, e0, comm, SVRParams(), accumulator_tensor=lambda: acc)
        pipe.InitializeEMGPU(ds.slices)

        def step():
            pipe.outer_iteration(0)

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
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        comm.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if rank == 0 else None
        backend.profile_enable(False)
        ms_total = ev0.elapsed_time(ev1)
        launches = backend.launch_count - launches0
        prof = backend.profile_read()

        # every rank must hold the same volume replica after a step (the accumulators were all-reduced, the rest is redundant)
        chk = torch.tensor([float(np.abs(backend.syncCPU().astype(np.float64)).sum())], dtype=torch.float64, device=dev)
        chk_lo, chk_hi = chk.clone(), chk.clone()
        if world > 1:
            dist.all_reduce(chk_lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(chk_hi, op=dist.ReduceOp.MAX)
        replicas_identical = bool(chk_lo.item() == chk_hi.item()) and bool(np.isfinite(chk.item()))

        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        ln = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        ms_total = float(t.item())
        ms_step = ms_total / args.steps
        value = S_global * PROJ_PER_SLICE_STEP / (ms_step * 1e-3)

        # ---- roofline of the dominant kernel (algorithmic bytes) --------------
        NP, V = backend.NP, backend.V
        alg_bytes = {"gaussian": 12 * NP + 20 * V, "simulate": 17 * NP + 8 * V, "superres": 16 * NP + 20 * V}
        names = {"gaussian": "gaussian_scatter_kernel (K1)", "simulate": "simulate_kernel (K2)",
                 "superres": "superres_scatter_kernel (K3)"}
        dom = max(alg_bytes, key=lambda k: prof[k][0])
        dom_ms = prof[dom][0] / max(prof[dom][1], 1)
        peak, peak_src = measured_peak()
        achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = recorded_traffic()
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic.get(dom), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes[dom], "ms_per_launch": dom_ms,
                    "limiter": "fp32 ALU + MUFU (4096 sinc^2*gauss taps per pixel); HBM is not binding -- see DESIGN.md",
                    "kernels": {names.get(k, k): {"ms_per_launch": v[0] / max(v[1], 1), "launches": v[1],
                                                  "share_of_step": v[0] / ms_total if ms_total else 0.0,
                                                  "alg_GBps": (alg_bytes[k] / (v[0] / max(v[1], 1) * 1e-3) / 1e9)
                                                  if k in alg_bytes and v[0] > 0 else None}
                                for k, v in prof.items()}}
        # what does bound these kernels: the warp schedulers' issue slots.  Tap count of one K2 launch (pixels that carry a
        # PSF sum x 16^3 taps) x the SASS instructions per tap of the per-pixel loop (349 per 16-tap row)
        # against 4 issue slots per SM per clock.  ncu's smsp__issue_active is ~87 % on stacks whose pixel rows run along
        # the volume's x; stacks in other orientations pay extra (L1 wavefronts; staged rows), which this figure shows.
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
            pinned = torch.from_numpy(ds.slices).pin_memory()
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
            wall = time.perf_counter() - t0
            tw = torch.tensor([wall], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            wall = float(tw.item())
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
            fe = RegistrationFrontEnd(backend, ds.slices, ds.slice_attrs, cfg.vol_voxel)
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
            reg_s = time.perf_counter() - t0
            treg = torch.tensor([reg_s], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(treg, op=dist.ReduceOp.MAX)
            registration = {"kernel": "reg_eval_kernel (fused sample + blur + NCC moments)", "ms_per_cost_evaluation": ev_ms,
                            "slices": Sl, "slice_size": [W, H], "alg_bytes_per_evaluation": alg,
                            "alg_GBps": alg / (ev_ms * 1e-3) / 1e9 if ev_ms > 0 else None,
                            "full_registration_s": float(treg.item()), "cost_evaluations_slice_offsets": backend.reg_evaluations,
                            "slices_registered_per_s": S_global / float(treg.item()), "host_prep_s": prep_s,
                            "schedule": "reference default: 2 levels x 4 steps x <=20 iterations, epsilon 1e-4"}

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _ = run_cpu_baseline(cfg, steps=1, warmup=0)
    ref_cuda = None
    if rank == 0 and world == 1 and not args.no_reference_cuda:
        ref_cuda = run_reference_cuda(cfg)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_json(cfg, world),
                "volumes_per_hour": vph, "roofline": roofline, "cpu_baseline": base, "reference_cuda": ref_cuda, "e2e": e2e,
                "registration": registration, "gpu_launches": int(ln.item()), "clocks": clocks,
                "replicas_identical": replicas_identical}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
