"""The reference's own PVR CUDA path behind the method set of fetalreconstruction_b200.pvr.PatchReconstruction
(TEST INFRASTRUCTURE ONLY): oracle/_ref/libref_pvr.so = the unmodified reconVolume.cu, patchBased{PSFReconstruction,
SimulatePatches,Superresolution,RobustStatistics}_gpu.cu of /root/reference compiled for sm_100a (oracle/Makefile,
oracle/ref_shim/ref_pvr_capi.cu).  Needs a GPU.

The reference keeps the robust-statistics state (sigma, mix, m, patch-level EM) inside patchBasedRobustStatistics_gpu,
so RefPVRPipeline replaces the split device/host steps of PVRPipeline with the reference's monolithic calls and reads
the state back.  Patch enumeration and extraction are host (IRTK) code in the reference and are not part of this
harness: it receives the patch list and the patch-value cube (as our patch extraction produced them)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_pvr.so")


def available() -> bool:
    return os.path.exists(LIB)


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, np.float32)
    return a if shape is None else a.reshape(shape)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefPatchReconstruction:
    def __init__(self, device: int = 0):
        self.lib = C.CDLL(LIB)
        self.lib.refpvr_create.restype = C.c_void_p
        self.device = device
        self.h = None
        self.spx = None
        self.use_spx = False
        self.launch_count = 0

    @property
    def V(self): return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]
    @property
    def NP(self): return self.n * self.pbx * self.pby

    # ---- set-up: collected, the reference objects are created when the patch values arrive -------------------------
    def recon_init(self, size, dim, recon_w2i, recon_i2w):
        self.vol_shape = tuple(int(v) for v in size)
        self.vol_dim = tuple(float(v) for v in dim)
        self.recon_w2i, self.recon_i2w = _f32(np.asarray(recon_w2i).ravel()), _f32(np.asarray(recon_i2w).ravel())

    def recon_setMask(self, mask): self.mask = np.ascontiguousarray(np.asarray(mask).ravel(), np.int8)

    def patches_init(self, pbx, pby, patches_per_stack, stack_dims):
        self.pbx, self.pby = int(pbx), int(pby)
        self.patches_per_stack = [int(v) for v in patches_per_stack]
        self.n = int(sum(self.patches_per_stack))
        self.stack_dims = np.asarray(stack_dims, np.float32).reshape(-1, 3)

    def patches_set_matrices(self, i2w, w2i, transformation, inv_transformation):
        self.m = [_f32(x, (self.n, 16)) for x in (i2w, w2i, transformation, inv_transformation)]

    def patches_set_spx(self, masks, use_spx=True):
        self.spx = None if masks is None else np.ascontiguousarray(masks, "S1").reshape(self.n, 4096)
        self.use_spx = bool(use_spx)

    def set_psf(self, psf_size, psf_i2w, quality_factor=1.0):
        self.psf = (tuple(int(v) for v in psf_size), _f32(np.asarray(psf_i2w).ravel()), float(quality_factor))

    def patches_copyFromHost(self, cube):
        cube = _f32(np.asarray(cube).reshape(self.n, self.pby, self.pbx))
        vx, vy, vz = self.vol_shape
        f = C.c_float
        self.h = C.c_void_p(self.lib.refpvr_create(C.c_int(self.device), vx, vy, vz, f(self.vol_dim[0]), f(self.vol_dim[1]), f(self.vol_dim[2]),
                                                   _p(self.recon_w2i), _p(self.recon_i2w), _p(self.mask)))
        o = 0
        for st, n in enumerate(self.patches_per_stack):
            sl = slice(o, o + n)
            d = self.stack_dims[st]
            spx = None if self.spx is None else _p(np.ascontiguousarray(self.spx[sl]))
            parts = [np.ascontiguousarray(m[sl]) for m in self.m]
            self.lib.refpvr_add_stack(self.h, self.pbx, self.pby, n, f(d[0]), f(d[1]), f(d[2]), f(float(d[2])), _p(np.ascontiguousarray(cube[sl])),
                                      _p(parts[0]), _p(parts[1]), _p(parts[2]), _p(parts[3]), spx)
            o += n
        (sx, sy, sz), i2w, q = self.psf
        w2i = _f32(np.linalg.inv(i2w.reshape(4, 4).astype(np.float64)).ravel())
        self.lib.refpvr_set_psf(self.h, sx, sy, sz, f(self.vol_dim[0]), f(self.vol_dim[1]), f(self.vol_dim[2]), _p(i2w), _p(w2i), f(q))

    def begin(self, min_intensity, max_intensity, adaptive=False):
        self.lib.refpvr_begin(self.h, C.c_float(min_intensity), C.c_float(max_intensity), int(bool(adaptive)), int(self.use_spx))
        self.min_intensity, self.max_intensity = float(min_intensity), float(max_intensity)

    # ---- the reference's calls ---------------------------------------------------------------------------------------
    def rs_initializeEMValues(self): self.lib.refpvr_rs_initialize_em_values(self.h)
    def recon_reset(self): self.lib.refpvr_recon_reset(self.h)
    def recon_resetAddonCmap(self): self.lib.refpvr_recon_reset_addon_cmap(self.h)
    def recon_equalize(self): self.lib.refpvr_recon_equalize(self.h)
    def patchBasedPSFReconstruction_gpu(self): self.lib.refpvr_psf_reconstruction(self.h)
    def patchBasedSimulatePatches_gpu(self): self.lib.refpvr_simulate_patches(self.h)
    def superresolution_run(self): self.lib.refpvr_superresolution_run(self.h)
    def superresolution_regularize(self, *a): self.lib.refpvr_superresolution_regularize(self.h)   # alpha, delta, lambda fixed in its ctor
    def ref_InitializeRobustStatistics(self):
        self.lib.refpvr_rs_initialize_robust_statistics(self.h, C.c_float(self.min_intensity), C.c_float(self.max_intensity))
    def ref_EStep(self): self.lib.refpvr_rs_estep(self.h)
    def ref_MStep(self, it): self.lib.refpvr_rs_mstep(self.h, int(it))
    def ref_Scale(self): self.lib.refpvr_rs_scale(self.h)

    def em_state(self):
        out = np.zeros(8, np.float32)
        self.lib.refpvr_get_em_state(self.h, _p(out))
        return out

    # ---- read-backs --------------------------------------------------------------------------------------------------
    def _vol(self, kind):
        out = np.zeros(self.V, np.float32)
        assert self.lib.refpvr_get_volume(self.h, kind, _p(out)) == 0
        return out

    def _patch(self, kind, dtype=np.float32):
        parts = []
        for st, n in enumerate(self.patches_per_stack):
            out = np.zeros(max(n * self.pbx * self.pby, 1), dtype)
            assert self.lib.refpvr_get_patch_buffer(self.h, st, kind, _p(out)) == 0
            parts.append(out[:n * self.pbx * self.pby])
        return np.concatenate(parts) if parts else np.zeros(0, dtype)

    def recon_copyToHost(self): return self._vol(0)
    def getVolWeights(self): return self._vol(1)
    def debugAddon(self): return self._vol(2)
    def debugConfidenceMap(self): return self._vol(3)
    def debugWeights(self): return self._patch(0)
    def debugSimpatches(self): return self._patch(1)
    def debugSimweights(self): return self._patch(2)
    def debugSiminside(self): return self._patch(3, np.int8)
    def debugPSFsums(self): return self._patch(4)
    def patches_copyToHost(self): return self._patch(5).reshape(self.n, self.pby, self.pbx)

    def rs_get_scales_weights(self):
        sc, w = [], []
        for st, n in enumerate(self.patches_per_stack):
            a, b = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
            assert self.lib.refpvr_get_patch_scales_weights(self.h, st, _p(a), _p(b)) == 0
            sc.append(a[:n]); w.append(b[:n])
        return np.concatenate(sc), np.concatenate(w)


def ref_pvr_pipeline_cls():
    from fetalreconstruction_b200.pvr import PVRPipeline

    class RefPVRPipeline(PVRPipeline):
        """The reference finishes every robust-statistics step inside patchBasedRobustStatistics_gpu."""

        def __init__(self, backend, min_intensity, max_intensity, params=None):
            super().__init__(backend, min_intensity, max_intensity, params)
            backend.begin(min_intensity, max_intensity, self.p.adaptive)

        def _sync(self):
            s = self.b.em_state()
            self.sigma, self.mix, self.m, self.sigma_s, self.mix_s, self.mean_s, self.mean_s2, self.sigma_s2 = (float(v) for v in s)

        def InitializeRobustStatistics(self):
            self.b.ref_InitializeRobustStatistics(); self._sync()

        def EStep(self):
            self.b.ref_EStep(); self._sync()
            self.patch_potential = np.zeros(self.b.n, np.float32)      # kept inside the reference's EStep

        def MStep(self, it):
            self.b.ref_MStep(it); self._sync()

        def Scale(self):
            self.b.ref_Scale()

    return RefPVRPipeline
