"""Host side of the PVR (patch-to-volume) path on top of include/pvr_abi.h.

`PatchReconstruction` is the ctypes binding (method names follow the reference objects they replace:
ReconVolume<T>, PatchBasedVolume<T>, patchBasedPSFReconstruction_gpu, patchBasedSimulatePatches_gpu,
patchBasedSuperresolution_gpu<T>, patchBasedRobustStatistics_gpu<T>); `PVRPipeline` restates the loop of
irtkPatchBasedReconstruction<T>::run() (source/reconstructionGPU2/irtkPatchBasedReconstruction.cpp:445-587) and
the scalar state of patchBasedRobustStatistics_gpu (patchBasedRobustStatistics_gpu.cu:746-851, 572-640);
`generate_2d_patches` restates the CPU patch enumeration PatchBasedObject::generate2DPatches
(include/patchBasedObject.cuh:176-342).  No arithmetic of the kernels lives here; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .geometry import ImageAttributes
from .reconstruction import SVRError, _f32, _p


class PatchReconstruction:
    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        if self._lib.pvr_create(C.byref(h), int(device)) != 0:
            raise SVRError(self._lib.svr_last_error(None).decode())
        self._h = h
        self.n = self.pbx = self.pby = 0
        self.vol_shape = (0, 0, 0)
        self.patches_per_stack = []

    def _ck(self, rc):
        if rc != 0:
            raise SVRError(self._lib.svr_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.svr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def V(self):
        return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]

    @property
    def NP(self):
        return self.n * self.pbx * self.pby

    @property
    def launch_count(self):
        return int(self._lib.svr_launch_count(self._h))

    # a PVR context is an svr_context with the PVR constants: stream, tuning and per-kernel timing are the SVR entry points
    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.svr_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def synchronize(self):
        self._ck(self._lib.svr_synchronize(self._h))

    def profile_enable(self, on=True):
        self._ck(self._lib.svr_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._ck(self._lib.svr_profile_reset(self._h))

    def profile_read(self):
        from .reconstruction import Reconstruction
        out = {}
        for i, name in enumerate(Reconstruction.KERNEL_KINDS):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self._lib.svr_profile_read(self._h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    # ReconVolume<T>
    def recon_init(self, size, dim, recon_w2i, recon_i2w):
        sx, sy, sz = (int(v) for v in size)
        a, b = _f32(np.asarray(recon_w2i).ravel(), (16,)), _f32(np.asarray(recon_i2w).ravel(), (16,))
        self._ck(self._lib.pvr_recon_init(self._h, sx, sy, sz, float(dim[0]), float(dim[1]), float(dim[2]), _p(a), _p(b)))
        self.vol_shape = (sx, sy, sz)

    def recon_setMask(self, mask):
        m = np.ascontiguousarray(np.asarray(mask).ravel(), np.int8)
        assert m.size == self.V
        self._ck(self._lib.pvr_recon_set_mask(self._h, _p(m)))

    def recon_reset(self): self._ck(self._lib.pvr_recon_reset(self._h))
    def recon_resetAddonCmap(self): self._ck(self._lib.pvr_recon_reset_addon_cmap(self._h))
    def recon_equalize(self): self._ck(self._lib.pvr_recon_equalize(self._h))

    def recon_copyFromHost(self, data):
        d = _f32(np.asarray(data).ravel(), (self.V,))
        self._ck(self._lib.pvr_recon_copy_from_host(self._h, _p(d)))

    def recon_copyToHost(self):
        out = np.empty(self.V, np.float32)
        self._ck(self._lib.pvr_recon_copy_to_host(self._h, _p(out)))
        return out

    # PatchBasedVolume<T>
    def patches_init(self, pbx, pby, patches_per_stack, stack_dims):
        pps = np.ascontiguousarray(patches_per_stack, np.int32)
        sd = _f32(stack_dims, (pps.size, 3))
        self._ck(self._lib.pvr_patches_init(self._h, int(pbx), int(pby), int(pps.size), _p(pps), _p(sd)))
        self.pbx, self.pby, self.n = int(pbx), int(pby), int(pps.sum())
        self.patches_per_stack = [int(v) for v in pps]

    def patches_set_matrices(self, i2w, w2i, transformation, inv_transformation):
        a, b = _f32(i2w, (self.n, 16)), _f32(w2i, (self.n, 16))
        t, ti = _f32(transformation, (self.n, 16)), _f32(inv_transformation, (self.n, 16))
        self._ck(self._lib.pvr_patches_set_matrices(self._h, _p(a), _p(b), _p(t), _p(ti)))

    def patches_set_spx(self, masks, use_spx=True):
        m = None if masks is None else np.ascontiguousarray(masks, "S1").reshape(self.n, 4096)
        self._ck(self._lib.pvr_patches_set_spx_masks(self._h, _p(m), int(bool(use_spx))))

    def patches_copyFromHost(self, cube):
        c = _f32(np.asarray(cube).ravel(), (self.NP,))
        self._ck(self._lib.pvr_patches_copy_from_host(self._h, _p(c)))

    def patches_copyToHost(self):
        out = np.empty(max(self.NP, 1), np.float32)
        self._ck(self._lib.pvr_patches_copy_to_host(self._h, _p(out)))
        return out[:self.NP].reshape(self.n, self.pby, self.pbx)

    def set_psf(self, psf_size, psf_i2w, quality_factor=1.0):
        sz = np.ascontiguousarray(psf_size, np.int32)
        m = _f32(np.asarray(psf_i2w).ravel(), (16,))
        self._ck(self._lib.pvr_set_psf(self._h, _p(sz), _p(m), float(quality_factor)))

    def initPatchBasedRecon_gpu(self, stack, stack_data, stack_w2i):
        sz, sy, sx = stack_data.shape
        d = _f32(np.asarray(stack_data).ravel())
        m = _f32(np.asarray(stack_w2i).ravel(), (16,))
        self._ck(self._lib.pvr_init_patch_based_recon(self._h, int(stack), _p(d), sx, sy, sz, _p(m)))

    # hot path
    def patchBasedPSFReconstruction_gpu(self): self._ck(self._lib.pvr_psf_reconstruction(self._h))
    def patchBasedSimulatePatches_gpu(self): self._ck(self._lib.pvr_simulate_patches(self._h))

    # N ranks: split phases around the all-reduce of the accumulator (include/pvr_abi.h)
    def psf_reconstruction_local(self): self._ck(self._lib.pvr_psf_reconstruction_local(self._h))
    def psf_reconstruction_finish(self): self._ck(self._lib.pvr_psf_reconstruction_finish(self._h))

    def accumulators(self, phase):
        """Device buffers the ranks must sum after `phase` ("psf" or "superresolution"): the interleaved accumulator."""
        from .reconstruction import _DeviceArray
        ptr, nb = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.svr_device_buffer(self._h, 0, C.byref(ptr), C.byref(nb)))
        return [_DeviceArray(ptr.value, nb.value // 4, self)]

    def rs_initialize_robust_statistics_local(self):
        s2 = (C.c_double * 2)()
        self._ck(self._lib.svr_initialize_robust_statistics_local(self._h, s2))
        return np.array([s2[0], s2[1]], np.float64)

    def rs_mstep_local(self):
        s5 = (C.c_double * 5)()
        self._ck(self._lib.svr_mstep_local(self._h, s5))
        return np.array(list(s5), np.float64)
    def superresolution_run(self): self._ck(self._lib.pvr_superresolution_run(self._h))

    def superresolution_regularize(self, adaptive, alpha, min_intensity, max_intensity, delta, lambda_):
        self._ck(self._lib.pvr_superresolution_regularize(self._h, int(bool(adaptive)), float(alpha), float(min_intensity),
                                                          float(max_intensity), float(delta), float(lambda_)))

    # robust statistics
    def rs_initializeEMValues(self): self._ck(self._lib.pvr_rs_initialize_em_values(self._h))

    def rs_InitializeRobustStatistics(self):
        s = C.c_float()
        self._ck(self._lib.pvr_rs_initialize_robust_statistics(self._h, C.byref(s)))
        return s.value

    def rs_estep_device(self, m, sigma, mix):
        pot = np.zeros(max(self.n, 1), np.float32)
        self._ck(self._lib.pvr_rs_estep_device(self._h, float(m), float(sigma), float(mix), _p(pot)))
        return pot[:self.n]

    def rs_get_scales_weights(self):
        s, w = np.zeros(max(self.n, 1), np.float32), np.zeros(max(self.n, 1), np.float32)
        self._ck(self._lib.pvr_rs_get_scales_weights(self._h, _p(s), _p(w)))
        return s[:self.n], w[:self.n]

    def rs_set_scales_weights(self, scales, weights):
        s, w = _f32(scales, (self.n,)), _f32(weights, (self.n,))
        self._ck(self._lib.pvr_rs_set_scales_weights(self._h, _p(s), _p(w)))

    def rs_MStep(self, it, step, sigma, mix, m):
        s, mi, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
        self._ck(self._lib.pvr_rs_mstep(self._h, int(it), float(step), C.byref(s), C.byref(mi), C.byref(mm)))
        return s.value, mi.value, mm.value

    def rs_Scale(self):
        sc = np.zeros(max(self.n, 1), np.float32)
        self._ck(self._lib.pvr_rs_scale(self._h, _p(sc)))
        return sc[:self.n]

    def _debug(self, kind, n, dtype):
        out = np.empty(max(n, 1), dtype)
        self._ck(self._lib.pvr_debug_get(self._h, kind, _p(out)))
        return out[:n]

    def debugWeights(self): return self._debug(0, self.NP, np.float32)
    def debugSimpatches(self): return self._debug(1, self.NP, np.float32)
    def debugSimweights(self): return self._debug(2, self.NP, np.float32)
    def debugSiminside(self): return self._debug(3, self.NP, np.int8)
    def debugConfidenceMap(self): return self._debug(4, self.V, np.float32)
    def debugAddon(self): return self._debug(5, self.V, np.float32)
    def debugPSFsums(self): return self._debug(6, self.NP, np.float32)

    def getVolWeights(self):
        out = np.empty(self.V, np.float32)
        self._ck(self._lib.svr_get_vol_weights(self._h, _p(out)))
        return out


def host_patch_em(patches_per_stack, patch_potential, scale, patch_weight, step, state5):
    """EStep host part (patchBasedRobustStatistics_gpu.cu:277-520), literal; returns the potentials the reference used."""
    lib = _lib.load()
    pps = np.ascontiguousarray(patches_per_stack, np.int32)
    pot, sc = _f32(patch_potential), _f32(scale)
    assert patch_weight.dtype == np.float32 and state5.dtype == np.float32
    used = np.zeros(max(int(pps.sum()), 1), np.float32)
    if lib.pvr_host_patch_em(int(pps.size), _p(pps), _p(pot), _p(sc), _p(patch_weight), float(step), _p(state5), _p(used)) != 0:
        raise SVRError("pvr_host_patch_em: bad argument")
    return used[:int(pps.sum())]


@dataclass
class PVRParams:
    """Defaults of patchBasedReconMain.cpp:110-144 / irtkPatchBasedReconstruction.cpp:415, patchBasedSuperresolution_gpu.cu:293-295."""
    iterations: int = 7            # the loop runs iterations + 1 passes (irtkPatchBasedReconstruction.cpp:445)
    rec_iterations: int = 7
    delta: float = 1.0
    lambda_: float = 0.1
    adaptive: bool = False
    step: float = 0.0001           # m_step of patchBasedRobustStatistics_gpu (:877); __step of the kernels is 1e-5


class PVRPipeline:
    """irtkPatchBasedReconstruction<T>::run() from the first iteration on, registration excluded."""

    def __init__(self, backend, min_intensity, max_intensity, params: PVRParams | None = None, comm=None, global_index=None,
                 patches_per_stack_global=None):
        """N ranks (one process per GPU): `backend` holds this rank's patches, `global_index[i]` is the position of its
        i-th patch in the stack-major order of ALL patches and `patches_per_stack_global` their count per stack; the
        ranks exchange the volume accumulators (all-reduce after P1 and P3), the partial sums of the robust statistics and
        the per-patch vectors, and every rank runs the identical patch-level EM."""
        self.b = backend
        self.comm = comm
        self.gidx = None if global_index is None else np.asarray(global_index, np.int64)
        self.pps_global = patches_per_stack_global
        self.p = params or PVRParams()
        self.min_intensity, self.max_intensity = float(min_intensity), float(max_intensity)
        self.alpha = (0.05 / self.p.lambda_) * self.p.delta * self.p.delta
        self.sigma = self.mix = self.m = 0.0
        self.sigma_s, self.mix_s, self.mean_s, self.mean_s2, self.sigma_s2 = 0.025, 0.9, 0.0, 0.0, 0.0
        self.patch_potential = None

    @property
    def distributed(self):
        return self.comm is not None and self.comm.active

    def _allreduce(self, phase):
        for buf in self.b.accumulators(phase):
            if isinstance(buf, np.ndarray):
                buf[...] = self.comm.sum(buf)                     # CPU twin (gloo)
            else:
                import torch
                self.comm.sum_device(torch.as_tensor(buf, device=self.comm.device))

    def _gather(self, local):
        """Per-patch vector of this rank -> the global stack-major vector (zero-filled all-reduce)."""
        g = np.zeros(int(np.sum(self.pps_global)), np.float64)
        g[self.gidx] = local
        return self.comm.sum(g).astype(np.float32)

    # patchBasedRobustStatistics_gpu<T>
    def InitializeRobustStatistics(self):
        if self.distributed:
            s2 = self.comm.sum(self.b.rs_initialize_robust_statistics_local())
            self.sigma = float(np.float32(s2[0]) / np.float32(s2[1]))
        else:
            self.sigma = self.b.rs_InitializeRobustStatistics()
        self.sigma_s, self.mix, self.mix_s = 0.025, 0.9, 0.9
        self.m = float(np.float32(1.0) / (np.float32(2.1) * np.float32(self.max_intensity)
                                          - np.float32(1.9) * np.float32(self.min_intensity)))

    def EStep(self):
        pot = self.b.rs_estep_device(self.m, self.sigma, self.mix)
        scales, weights = self.b.rs_get_scales_weights()
        weights = weights.astype(np.float32).copy()
        state = np.array([self.sigma_s, self.mix_s, self.mean_s, self.mean_s2, self.sigma_s2], np.float32)
        if self.distributed:
            g_pot, g_sc, g_w = self._gather(pot), self._gather(scales), np.ascontiguousarray(self._gather(weights), np.float32)
            self.patch_potential = host_patch_em(self.pps_global, g_pot, g_sc, g_w, self.p.step, state)
            weights = np.ascontiguousarray(g_w[self.gidx], np.float32)
        else:
            self.patch_potential = host_patch_em(self.b.patches_per_stack, pot, scales, weights, self.p.step, state)
        self.sigma_s, self.mix_s, self.mean_s, self.mean_s2, self.sigma_s2 = (float(v) for v in state)
        self.b.rs_set_scales_weights(scales, weights)

    def MStep(self, it):
        if self.distributed:
            from .reconstruction import mstep_finish
            s5 = self.b.rs_mstep_local()
            sums = self.comm.sum(s5[:3].copy()); mn = self.comm.min(s5[3:4].copy()); mx = self.comm.max(s5[4:5].copy())
            self.sigma, self.mix, self.m = mstep_finish(np.concatenate([sums, mn, mx]), it, self.p.step, self.sigma, self.mix, self.m)
        else:
            self.sigma, self.mix, self.m = self.b.rs_MStep(it, self.p.step, self.sigma, self.mix, self.m)

    def Scale(self):
        return self.b.rs_Scale()

    def iteration(self):
        b, p = self.b, self.p
        b.rs_initializeEMValues()
        b.recon_reset()
        if self.distributed:
            b.psf_reconstruction_local()
            self._allreduce("psf")
            b.psf_reconstruction_finish()
        else:
            b.patchBasedPSFReconstruction_gpu()
        b.recon_equalize()
        b.patchBasedSimulatePatches_gpu()
        self.InitializeRobustStatistics()
        self.EStep()
        for i in range(p.rec_iterations):
            self.Scale()
            b.recon_resetAddonCmap()
            b.superresolution_run()
            if self.distributed:
                self._allreduce("superresolution")
            b.superresolution_regularize(p.adaptive, self.alpha, self.min_intensity, self.max_intensity, p.delta, p.lambda_)
            b.patchBasedSimulatePatches_gpu()
            self.MStep(i + 1)
            self.EStep()

    def run(self, register=None):
        for it in range(self.p.iterations + 1):
            if it > 0 and register is not None:
                register(it)
            self.iteration()
        return self.b.recon_copyToHost()


class PatchRegistration:
    """The patch-to-volume registration PVRreconstructionGPU runs between iterations: patchBased2D3DRegistration<T>::runHybrid
    (patchBased2D3DRegistration.cpp:88-223) registers every patch with IRTK's irtkImageRigidRegistrationWithPadding on the CPU
    (GuessParameterSliceToVolume, target padding -1) against the current reconstruction.  Here all patches of the rank go through
    the device engine in one call (rreg.register, kind SLICE_TO_VOLUME): the patch (irtkGreyImage cast, origin reset into the
    transformation like ResetOrigin does) is the target, the reconstruction cast to short the shared source.
    (The reference resamples a patch to the volume's voxel size into a SHADOWED variable and registers the unresampled patch,
    patchBased2D3DRegistration.cpp:116-125: reproduced by not resampling.)"""

    def __init__(self, backend, patch_attrs, patch_cube, transformations, vol_attr):
        from . import rreg
        self.b = backend
        self.rreg = rreg
        self.cube = patch_cube
        self.T = [np.asarray(t, np.float64).reshape(4, 4).copy() for t in transformations]
        self.vol_attr18 = rreg.attrs18(vol_attr)
        self.attrs, self.mo = [], []
        for a in patch_attrs:
            a18 = rreg.attrs18(a)
            mo = np.eye(4)
            mo[:3, 3] = a18[6:9]
            a18[6:9] = 0.0
            self.attrs.append(a18); self.mo.append(mo)
        self.i2w = np.stack([a.image_to_world().astype(np.float32).ravel() for a in patch_attrs]) if patch_attrs else np.zeros((0, 16), np.float32)
        self.w2i = np.stack([a.world_to_image().astype(np.float32).ravel() for a in patch_attrs]) if patch_attrs else np.zeros((0, 16), np.float32)
        self.evaluations = 0

    def __call__(self, it=0):
        from .geometry import rigid_matrix, rigid_parameters
        n = len(self.attrs)
        if n == 0:
            return
        vol = self.b.recon_copyToHost()
        images = [self.rreg.to_grey(vol)] + [self.rreg.to_grey(self.cube[k]) for k in range(n)]
        attrs = [self.vol_attr18] + self.attrs
        starts = [rigid_parameters(self.T[k] @ self.mo[k]) for k in range(n)]
        dofs, _, ev = self.rreg.register(self.b, images, attrs, list(range(1, n + 1)), [0] * n, self.rreg.SLICE_TO_VOLUME, starts)
        self.evaluations += ev
        for k in range(n):
            self.T[k] = rigid_matrix(*dofs[k]) @ np.linalg.inv(self.mo[k])
        T = np.stack([t.astype(np.float32).ravel() for t in self.T])
        Ti = np.stack([np.linalg.inv(t).astype(np.float32).ravel() for t in self.T])
        self.b.patches_set_matrices(self.i2w, self.w2i, T, Ti)


def generate_2d_patches(stack: np.ndarray, attr: ImageAttributes, mask: np.ndarray, mask_attr: ImageAttributes,
                        pbb=(64, 64), stride=(32, 32), thickness=2.5):
    """PatchBasedObject::generate2DPatches (include/patchBasedObject.cuh:176-342): enumerate pbb-sized boxes on a
    stride grid over every slice (the grid runs to size + pbb), keep those where more than 1/3 of the pixels are
    inside the mask and neither 0 nor -1.  Returns (list of patch ImageAttributes, CPU-extracted patch cube)."""
    pbx, pby = pbb
    w2i_mask = mask_attr.world_to_image()
    attrs, cubes = [], []
    jj, ii = np.meshgrid(np.arange(pby, dtype=np.float64), np.arange(pbx, dtype=np.float64), indexing="ij")
    for z in range(attr.z):
        sattr = attr.slice_attributes(z, thickness * 2)
        s_i2w, s_w2i = sattr.image_to_world(), sattr.world_to_image()
        sl = stack[z]
        for y in range(0, attr.y + pby, stride[1]):
            for x in range(0, attr.x + pbx, stride[0]):
                pa = ImageAttributes(pbx, pby, 1, attr.dx, attr.dy, thickness * 2, np.zeros(3), sattr.xaxis, sattr.yaxis, sattr.zaxis)
                p1 = s_i2w @ np.array([x, y, 0, 1.0])
                p2 = pa.image_to_world() @ np.array([0, 0, 0, 1.0])
                pa.origin = (p1 - p2)[:3]
                i2w = pa.image_to_world()
                m_s = s_w2i @ i2w
                m_m = w2i_mask @ i2w
                xs = m_s[0, 0] * ii + m_s[0, 1] * jj + m_s[0, 3]
                ys = m_s[1, 0] * ii + m_s[1, 1] * jj + m_s[1, 3]
                xm = m_m[0, 0] * ii + m_m[0, 1] * jj + m_m[0, 3]
                ym = m_m[1, 0] * ii + m_m[1, 1] * jj + m_m[1, 3]
                zm = m_m[2, 0] * ii + m_m[2, 1] * jj + m_m[2, 3]
                ok = (xs >= 0) & (ys >= 0) & (xs < attr.x) & (ys < attr.y)
                ok &= (xm >= 0) & (ym >= 0) & (zm >= 0) & (xm < mask.shape[2]) & (ym < mask.shape[1]) & (zm < mask.shape[0])
                # irtkGenericImage::Get(double...) truncates to int
                xi, yi = np.clip(xs.astype(int), 0, attr.x - 1), np.clip(ys.astype(int), 0, attr.y - 1)
                mx = np.clip(xm.astype(int), 0, mask.shape[2] - 1)
                my = np.clip(ym.astype(int), 0, mask.shape[1] - 1)
                mz = np.clip(zm.astype(int), 0, mask.shape[0] - 1)
                ok &= mask[mz, my, mx] > 0
                patch = np.where(ok, sl[yi, xi], 0.0).astype(np.float32)
                set_count = int(np.count_nonzero(ok & (patch != 0) & (patch != -1)))
                if set_count > 1.0 / 3.0 * pbx * pby:
                    attrs.append(pa)
                    cubes.append(patch)
    cube = np.stack(cubes) if cubes else np.zeros((0, pby, pbx), np.float32)
    return attrs, cube
