"""Host orchestration of the SVR hot path: the `...GPU()` shims of class irtkReconstruction
(source/reconstructionGPU2/irtkReconstructionGPU.cc) and the interleaved reconstruction loop of
SVRreconstructionGPU's main (source/reconstructionGPU2/reconstruction.cc:800-1276), restated over
the device library, for one rank or for N ranks (one process per GPU).

Multi-rank: slices are sharded by whole stacks (svr_host_partition); every rank
keeps a full volume replica; the only data-path exchanges are
  * one all-reduce(sum) of the interleaved accumulator [2V floats] after K1 and after every K3
    (replaces the reference's reduce-to-device-0 + broadcast, reconstruction_cuda2.cu:2225-2239,
    2445-2460, 2175-2180) -- every rank then runs the regulariser on identical data, so no broadcast;
  * tiny all-reduces of the robust statistics (sum / min / max) and of the per-slice vectors.
The collectives are torch.distributed calls (NCCL over NVLink on GPUs, gloo in the CPU tests) on the
library's own device buffers, ordered on the stream the library runs on.

`backend` is any object with the method set of fetalreconstruction_b200.reconstruction.Reconstruction;
tests/ substitutes an oracle-backed twin to exercise this file on CPU with gloo.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import reconstruction as R


@dataclass
class SVRParams:
    """Defaults of reconstruction.cc:92-116,171-186 and irtkReconstructionGPU.cc:159-187."""
    iterations: int = 4
    rec_iterations_first: int = 4
    rec_iterations_last: int = 13
    delta: float = 150.0
    lambda_: float = 0.02
    last_iter_lambda: float = 0.01
    levels: int = 3
    step: float = 0.0001
    adaptive: bool = False
    intensity_matching: bool = True
    force_excluded: list = field(default_factory=list)


class Comm:
    """Thin wrapper over a torch.distributed process group (or nothing for one rank)."""

    def __init__(self, group=None, device=None, stream_ordered=False):
        """stream_ordered: the library runs on torch's current stream (svr_set_stream), so a collective issued through torch is
        ordered with the library's kernels on the device and the host does not have to wait for it."""
        self.group = group
        self.device = device
        self.stream_ordered = stream_ordered
        if group is not None:
            import torch.distributed as dist
            self.dist = dist
            self.rank = dist.get_rank(group)
            self.size = dist.get_world_size(group)
        else:
            self.rank, self.size = 0, 1

    @property
    def active(self):
        return self.group is not None and self.size > 1

    def _reduce_np(self, arr, op):
        if not self.active:
            return arr
        import torch
        t = torch.as_tensor(np.ascontiguousarray(arr), device=self.device)
        self.dist.all_reduce(t, op=op, group=self.group)
        return t.cpu().numpy()

    def sum(self, arr):
        return self._reduce_np(arr, self.dist.ReduceOp.SUM) if self.active else arr

    def min(self, arr):
        return self._reduce_np(arr, self.dist.ReduceOp.MIN) if self.active else arr

    def max(self, arr):
        return self._reduce_np(arr, self.dist.ReduceOp.MAX) if self.active else arr

    def sum_device(self, tensor):
        """In-place all-reduce of a device buffer of the library.  The collective is ordered on torch's current stream;
        the library may run on a stream of its own (svr_set_stream not called), so the host waits for the result before
        the next library call can touch the buffer (a few microseconds when the streams are the same)."""
        if self.active:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
            if tensor.is_cuda and not self.stream_ordered:
                import torch
                torch.cuda.current_stream(tensor.device).synchronize()

    def gather_rows(self, arr):
        """Every rank's copy of a small host vector, stacked [size, n]: one collective for statistics that need different
        reductions (the M-step's sums, minimum and maximum)."""
        arr = np.ascontiguousarray(arr, np.float64)
        if not self.active:
            return arr[None]
        import torch
        t = torch.as_tensor(arr, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t, group=self.group)
        return torch.stack(out).cpu().numpy()

    def barrier(self):
        if self.active:
            self.dist.barrier(group=self.group)


class SVRPipeline:
    """State and steps of irtkReconstruction's GPU path for the slices [begin, end) of this rank."""

    def __init__(self, backend, S_global: int, begin: int, end: int, comm: Comm | None = None,
                 params: SVRParams | None = None, accumulator_tensor=None, host=None):
        self.b = backend
        # the pure host helpers (slice-level EM, small-slice rule, M-step finish): by default the ones the C ABI exports
        # (libsvr_b200.so); the CPU baseline arm of bench.py passes the oracle's own copies so that it never loads the GPU library
        self.host = host if host is not None else R
        self.S = S_global
        self.begin, self.end = begin, end
        self.comm = comm or Comm()
        self.p = params or SVRParams()
        self._acc_tensor = accumulator_tensor      # callable returning the torch view of the accumulator
        # irtkReconstruction members (irtkReconstructionGPU.cc:159-187)
        self._step = self.p.step
        self._sigma_s, self._mix_s, self._mix = 0.025, 0.9, 0.9
        self._mean_s = self._mean_s2 = self._sigma_s2 = 0.0
        self._sigma = self._m = 0.0
        self._delta, self._lambda = 1.0, 0.1
        self._alpha = (0.05 / self._lambda) * self._delta * self._delta
        self._adaptive = self.p.adaptive
        self._scale = np.ones(self.S, np.float32)
        self._slice_weight = np.ones(self.S, np.float32)
        self._slice_inside = np.zeros(self.S, bool)
        self._small_slices = np.zeros(0, np.int32)
        self._min_intensity = self._max_intensity = 0.0
        self.slice_potential = np.zeros(self.S, np.float32)
        self.timers = {}

    # -- helpers --------------------------------------------------------------------------------
    def _local(self, v):
        return np.ascontiguousarray(v[self.begin:self.end])

    def _gather(self, local, dtype=np.float32):
        """Per-slice vector of this rank -> global vector (zero-filled all-reduce)."""
        g = np.zeros(self.S, np.float64)
        g[self.begin:self.end] = local
        return self.comm.sum(g).astype(dtype)

    def _allreduce_accumulator(self):
        if self.comm.active:
            self.comm.sum_device(self._acc_tensor())

    # -- irtkReconstruction::SetSmoothingParameters (irtkReconstructionGPU.h:605-612) ------------
    def SetSmoothingParameters(self, delta, lambda_):
        self._delta = delta
        self._lambda = lambda_ * delta * delta
        self._alpha = 0.05 / lambda_
        if self._alpha > 1:
            self._alpha = 1

    # -- InitializeEMGPU (irtkReconstructionGPU.cc:2921-2953) -------------------------------------
    def InitializeEMGPU(self, local_slices):
        self.InitializeEMValuesGPU()
        pos = local_slices[local_slices > 0]
        mx = np.array([pos.max() if pos.size else -np.inf], np.float64)
        mn = np.array([pos.min() if pos.size else np.inf], np.float64)
        self._max_intensity = float(self.comm.max(mx)[0])
        self._min_intensity = float(self.comm.min(mn)[0])

    # -- InitializeEMValuesGPU (irtkReconstructionGPU.cc:2905-2919) --------------------------------
    def InitializeEMValuesGPU(self):
        self._slice_weight = np.ones(self.S, np.float32)
        self._scale = np.ones(self.S, np.float32)
        self.b.UpdateScaleVector(self._local(self._scale), self._local(self._slice_weight))
        self.b.InitializeEMValues()

    # -- GaussianReconstructionGPU (irtkReconstructionGPU.cc:2695-2762) ----------------------------
    def GaussianReconstructionGPU(self):
        self.b.gaussian_reconstruction_local()
        self._allreduce_accumulator()
        vn_local = self.b.gaussian_reconstruction_finish()
        voxel_num = self._gather(vn_local, np.int32)
        self._small_slices = self.host.host_small_slices(voxel_num)
        return voxel_num

    # -- SimulateSlicesGPU (irtkReconstructionGPU.cc:1163-1203) ------------------------------------
    def SimulateSlicesGPU(self, need_inside=True):
        """need_inside: fetch and exchange the per-slice "inside" flags (only InitializeRobustStatisticsGPU and the evaluation
        listing read them: the inner loop skips the read-back and its tiny all-reduce)."""
        try:
            inside = self.b.SimulateSlices(fetch_inside=need_inside)
        except TypeError:                                   # backend twins without the keyword (oracle / reference)
            inside = self.b.SimulateSlices()
        if need_inside:
            self._slice_inside = self._gather(inside.astype(np.float64), np.float64) > 0

    # -- InitializeRobustStatisticsGPU (irtkReconstructionGPU.cc:2988-3020) ------------------------
    def InitializeRobustStatisticsGPU(self):
        s2 = self.comm.sum(self.b.initialize_robust_statistics_local())
        self._sigma = float(np.float32(s2[0]) / np.float32(s2[1]))
        self._slice_weight[~self._slice_inside] = 0
        for i in self.p.force_excluded:
            self._slice_weight[i] = 0
        self._sigma_s = 0.025
        self._mix = 0.9
        self._mix_s = 0.9
        # (float)(1.0f / (2.1f * _max_intensity - 1.9f * _min_intensity)) with double intensities
        self._m = float(np.float32(1.0 / (float(np.float32(2.1)) * self._max_intensity
                                          - float(np.float32(1.9)) * self._min_intensity)))
        self.b.UpdateScaleVector(self._local(self._scale), self._local(self._slice_weight))

    # -- EStepGPU (irtkReconstructionGPU.cc:3184-3440) ---------------------------------------------
    def EStepGPU(self):
        pot_local = self.b.EStep(self._m, self._sigma, self._mix)
        pot = self._gather(pot_local)
        state = np.array([self._sigma_s, self._mix_s, self._mean_s, self._mean_s2, self._sigma_s2], np.float32)
        sw = self._slice_weight.astype(np.float32).copy()
        self.host.host_slice_em(pot, self._scale, sw, np.asarray(self.p.force_excluded, np.int32), self._small_slices,
                        self._step, state)
        self.slice_potential = pot
        self._slice_weight = sw
        self._sigma_s, self._mix_s, self._mean_s, self._mean_s2, self._sigma_s2 = (float(v) for v in state)
        self.b.UpdateSliceWeights(self._local(self._slice_weight))

    # -- ScaleGPU (irtkReconstructionGPU.cc:3751-3765) ---------------------------------------------
    def ScaleGPU(self):
        self._scale = self._gather(self.b.CalculateScaleVector())

    # -- SuperresolutionGPU (irtkReconstructionGPU.cc:4024-4053) -----------------------------------
    def SuperresolutionGPU(self, it):
        self.b.superresolution_local(self._local(self._slice_weight))
        self._allreduce_accumulator()
        self.b.superresolution_finish(self._adaptive, self._alpha, self._min_intensity, self._max_intensity,
                                      self._delta, self._lambda)

    # -- MStepGPU (irtkReconstructionGPU.cc:4214-4224) ---------------------------------------------
    def MStepGPU(self, it):
        rows = self.comm.gather_rows(self.b.mstep_local())       # one exchange; sums / min / max folded on the host
        s5 = np.concatenate([rows[:, :3].sum(0), rows[:, 3:4].min(0), rows[:, 4:5].max(0)])
        self._sigma, self._mix, self._m = self.host.mstep_finish(s5, it, self._step, self._sigma, self._mix, self._m)

    def MaskVolumeGPU(self):
        self.b.maskVolume()

    def ScaleVolumeGPU(self):
        s2 = self.comm.sum(self.b.scale_volume_local())
        scale = float(np.float32(s2[0] / s2[1]))
        self.b.scale_volume_apply(scale)
        return scale

    # -- one outer iteration of reconstruction.cc:817-1237 (registration excluded) ----------------
    def set_schedule(self, it):
        p = self.p
        if it == p.iterations - 1:
            self.SetSmoothingParameters(p.delta, p.last_iter_lambda)
        else:
            l = p.lambda_
            for i in range(p.levels):
                if it == p.iterations * (p.levels - i - 1) // p.levels:
                    self.SetSmoothingParameters(p.delta, l)
                l *= 2

    def reconstruction_iteration(self, i):
        """Body of the inner loop, reconstruction.cc:1013-1126."""
        if self.p.intensity_matching:
            self.ScaleGPU()
        self.SuperresolutionGPU(i + 1)
        self.SimulateSlicesGPU(need_inside=False)
        self.MStepGPU(i + 1)
        self.EStepGPU()

    def outer_iteration(self, it, update_matrices=None):
        p = self.p
        self.set_schedule(it)
        self.InitializeEMValuesGPU()
        if update_matrices is not None:
            update_matrices()                      # UpdateGPUTranformationMatrices, reconstruction.cc:945
        self.GaussianReconstructionGPU()
        self.SimulateSlicesGPU()
        self.InitializeRobustStatisticsGPU()
        self.EStepGPU()
        rec = p.rec_iterations_last if it == p.iterations - 1 else p.rec_iterations_first
        for i in range(rec):
            self.reconstruction_iteration(i)
        self.MaskVolumeGPU()

    def run(self, update_matrices=None, register=None):
        for it in range(self.p.iterations):
            if it > 0 and register is not None:
                register(it)                       # SliceToVolumeRegistration*, reconstruction.cc:826-886
            self.outer_iteration(it, update_matrices)
        self.ScaleVolumeGPU()
        return self.b.syncCPU()


def upload_dataset(backend, ds, begin=0, end=None):
    """irtkReconstruction::SyncGPU + UpdateGPUTranformationMatrices + generatePSFVolume
    (irtkReconstructionGPU.cc:249-401,1496-1610) for the slices [begin, end) of a phantom.Dataset."""
    from .geometry import ImageAttributes
    end = ds.S if end is None else end
    vz, vy, vx = ds.mask.shape
    vox = ds.cfg.vol_voxel
    backend.InitReconstructionVolume((vx, vy, vz), (vox, vox, vox), None, 0.0)
    backend.setMask((vx, vy, vz), (vox, vox, vox), ds.mask.ravel(), 0.0)
    Ny, Nx = ds.slices.shape[1:]
    backend.initStorageVolumes((Nx, Ny, end - begin), ds.dims[0] if ds.S else (1, 1, 1))
    backend.FillSlices(ds.slices[begin:end].ravel())
    backend.setSliceDims(ds.dims[begin:end], 1.0)
    backend.SetSliceMatrices(ds.trans[begin:end], ds.trans_inv[begin:end], ds.i2w[begin:end], ds.w2i[begin:end],
                             ds.i2w[begin:end], ds.w2i[begin:end], ds.recon_i2w, ds.recon_w2i)
    psf = ImageAttributes(128, 128, 128, vox, vox, vox)           # PSF_SIZE, irtkReconstructionGPU.cc:1512-1544
    backend.generatePSFVolume(None, (128, 128, 128), ds.dims[0] if ds.S else (1, 1, 1), (vox, vox, vox),
                              psf.image_to_world().astype(np.float32).ravel(),
                              psf.world_to_image().astype(np.float32).ravel(), 1.0)
