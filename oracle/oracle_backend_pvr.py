"""Oracle-backed twin of fetalreconstruction_b200.pvr.PatchReconstruction (TEST INFRASTRUCTURE ONLY).

Same method set, implemented with oracle/pvr_oracle.c on numpy arrays; tests drive the CUDA backend and this
twin through the same PVRPipeline and compare.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import oracle as orc
from oracle.oracle import _ptr


class _DS:
    pass


class OraclePatchReconstruction:
    def __init__(self, device=0):
        self.n = self.pbx = self.pby = 0
        self.vol_shape = (0, 0, 0)
        self.patches_per_stack = []
        self.psf_c = np.zeros(3, np.float32)
        self.spx = None
        self.use_spx = False
        self.launch_count = 0

    @property
    def V(self): return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]
    @property
    def NP(self): return self.n * self.pbx * self.pby

    def recon_init(self, size, dim, recon_w2i, recon_i2w):
        self.vol_shape = tuple(int(v) for v in size)
        self.recon_w2i = np.ascontiguousarray(recon_w2i, np.float32).ravel()
        self.recon_i2w = np.ascontiguousarray(recon_i2w, np.float32).ravel()
        self.recon = np.zeros(self.V, np.float32)
        self.volw = np.zeros(self.V, np.float32)
        self.addon = np.zeros(self.V, np.float32)
        self.cmap = np.zeros(self.V, np.float32)

    def recon_setMask(self, mask): self.mask = np.ascontiguousarray(np.asarray(mask).ravel(), np.int8)
    def recon_reset(self): self.recon[:] = 0; self.volw[:] = 0
    def recon_resetAddonCmap(self): self.addon[:] = 0; self.cmap[:] = 0
    def recon_equalize(self): orc.lib().pvr_equalize(C.c_size_t(self.V), _ptr(self.recon), _ptr(self.volw))
    def recon_copyFromHost(self, data): self.recon = np.array(data, np.float32).ravel()
    def recon_copyToHost(self): return self.recon.copy()
    def getVolWeights(self): return self.volw.copy()

    def patches_init(self, pbx, pby, patches_per_stack, stack_dims):
        self.pbx, self.pby = int(pbx), int(pby)
        self.patches_per_stack = [int(v) for v in patches_per_stack]
        self.n = int(sum(self.patches_per_stack))
        sd = np.asarray(stack_dims, np.float32).reshape(-1, 3)
        self.dims = np.ascontiguousarray(np.repeat(sd, self.patches_per_stack, axis=0), np.float32)
        n = self.NP
        self.patches = np.zeros(n, np.float32)
        self.weights = np.zeros(n, np.float32)
        self.simpatches = np.zeros(n, np.float32)
        self.simweights = np.zeros(n, np.float32)
        self.siminside = np.zeros(n, np.int8)
        self.psf_sums = np.zeros(n, np.float32)
        self.scales = np.ones(self.n, np.float32)
        self.patch_weights = np.ones(self.n, np.float32)
        self.spx = np.full((self.n, 4096), b"0", "S1")

    def patches_set_matrices(self, i2w, w2i, transformation, inv_transformation):
        self.i2w = np.ascontiguousarray(i2w, np.float32)
        self.w2i = np.ascontiguousarray(w2i, np.float32)
        self.trans = np.ascontiguousarray(transformation, np.float32)
        self.trans_inv = np.ascontiguousarray(inv_transformation, np.float32)

    def patches_set_spx(self, masks, use_spx=True):
        if masks is not None:
            self.spx = np.ascontiguousarray(masks, "S1").reshape(self.n, 4096)
        self.use_spx = bool(use_spx)

    def patches_copyFromHost(self, cube): self.patches = np.array(cube, np.float32).ravel()
    def patches_copyToHost(self): return self.patches.reshape(self.n, self.pby, self.pbx).copy()

    def set_psf(self, psf_size, psf_i2w, quality_factor=1.0):
        m = np.asarray(psf_i2w, np.float32).reshape(4, 4)
        c = [np.float32((s - 1) * 0.5) for s in psf_size]
        for r in range(3):
            self.psf_c[r] = np.float32(np.float32(np.float32(m[r, 0] * c[0]) + np.float32(m[r, 1] * c[1]))
                                       + np.float32(m[r, 2] * c[2])) + m[r, 3]

    def _geom(self):
        ds = _DS()
        ds.i2w, ds.w2i, ds.trans, ds.trans_inv, ds.dims = self.i2w, self.w2i, self.trans, self.trans_inv, self.dims
        ds.slices = np.empty((self.n, self.pby, self.pbx), np.float32)
        vx, vy, vz = self.vol_shape
        ds.mask = np.empty((vz, vy, vx), np.float32)
        ds.recon_i2w, ds.recon_w2i, ds.psf_c = self.recon_i2w, self.recon_w2i, self.psf_c
        return orc.Geometry(ds)

    def _spx_ptr(self):
        return _ptr(self.spx) if self.use_spx else None

    def initPatchBasedRecon_gpu(self, stack, stack_data, stack_w2i):
        sz, sy, sx = stack_data.shape
        g = self._geom()
        p0 = int(sum(self.patches_per_stack[:stack]))
        d = np.ascontiguousarray(stack_data, np.float32)
        m = np.ascontiguousarray(stack_w2i, np.float32).ravel()
        orc.lib().pvr_patch_init(C.byref(g.c), p0, self.patches_per_stack[stack], _ptr(d), sx, sy, sz, _ptr(m), _ptr(self.mask),
                                 self._spx_ptr(), _ptr(self.patches))

    def patchBasedPSFReconstruction_gpu(self):
        g = self._geom()
        orc.lib().pvr_psf_reconstruction(C.byref(g.c), _ptr(self.patches), _ptr(self.scales), _ptr(self.mask), self._spx_ptr(),
                                         _ptr(self.recon), _ptr(self.volw), _ptr(self.psf_sums))

    # N ranks (PVRPipeline with a Comm): the twin keeps {recon, volw} / {addon, cmap} as separate arrays
    def psf_reconstruction_local(self): self.patchBasedPSFReconstruction_gpu()
    def psf_reconstruction_finish(self): pass
    def accumulators(self, phase): return [self.recon, self.volw] if phase == "psf" else [self.addon, self.cmap]

    def rs_initialize_robust_statistics_local(self):
        sa, sb = C.c_double(), C.c_double()
        f = orc.lib().pvr_initialize_robust_statistics
        f.restype = C.c_float
        f(C.c_size_t(self.NP), _ptr(self.patches), _ptr(self.siminside), _ptr(self.simpatches), _ptr(self.simweights), C.byref(sa), C.byref(sb))
        return np.array([sa.value, sb.value], np.float64)

    def rs_mstep_local(self):
        s5 = np.zeros(5, np.float64)
        orc.lib().pvr_mstep_sums(self.n, self.pbx, self.pby, _ptr(self.patches), _ptr(self.weights), _ptr(self.simpatches),
                                 _ptr(self.simweights), _ptr(self.scales), _ptr(s5))
        return s5

    def patchBasedSimulatePatches_gpu(self):
        g = self._geom()
        orc.lib().pvr_simulate_patches(C.byref(g.c), _ptr(self.patches), _ptr(self.psf_sums), _ptr(self.recon), _ptr(self.mask),
                                       _ptr(self.simpatches), _ptr(self.simweights), _ptr(self.siminside))

    def superresolution_run(self):
        g = self._geom()
        orc.lib().pvr_superresolution(C.byref(g.c), _ptr(self.patches), _ptr(self.weights), _ptr(self.simpatches),
                                      _ptr(self.patch_weights), _ptr(self.scales), _ptr(self.mask), _ptr(self.psf_sums),
                                      _ptr(self.addon), _ptr(self.cmap))

    def superresolution_regularize(self, adaptive, alpha, min_intensity, max_intensity, delta, lambda_):
        vx, vy, vz = self.vol_shape
        orc.lib().pvr_regularize(vx, vy, vz, _ptr(self.recon), _ptr(self.addon), _ptr(self.cmap), C.c_int(int(adaptive)),
                                 C.c_float(alpha), C.c_float(min_intensity), C.c_float(max_intensity), C.c_float(delta),
                                 C.c_float(lambda_))

    def rs_initializeEMValues(self):
        self.scales[:] = 1
        self.patch_weights[:] = 1
        orc.lib().pvr_initialize_em_values(C.c_size_t(self.NP), _ptr(self.patches), _ptr(self.weights))

    def rs_InitializeRobustStatistics(self):
        f = orc.lib().pvr_initialize_robust_statistics
        f.restype = C.c_float
        return float(f(C.c_size_t(self.NP), _ptr(self.patches), _ptr(self.siminside), _ptr(self.simpatches), _ptr(self.simweights),
                       None, None))

    def rs_estep_device(self, m, sigma, mix):
        pot = np.zeros(max(self.n, 1), np.float32)
        orc.lib().pvr_estep(self.n, self.pbx, self.pby, _ptr(self.patches), _ptr(self.simpatches), _ptr(self.simweights),
                            _ptr(self.scales), C.c_float(m), C.c_float(sigma), C.c_float(mix), _ptr(self.weights), _ptr(pot))
        return pot[:self.n]

    def rs_get_scales_weights(self): return self.scales.copy(), self.patch_weights.copy()

    def rs_set_scales_weights(self, scales, weights):
        self.scales = np.array(scales, np.float32)
        self.patch_weights = np.array(weights, np.float32)

    def rs_MStep(self, it, step, sigma, mix, m):
        s5 = np.zeros(5, np.float64)
        orc.lib().pvr_mstep_sums(self.n, self.pbx, self.pby, _ptr(self.patches), _ptr(self.weights), _ptr(self.simpatches),
                                 _ptr(self.simweights), _ptr(self.scales), _ptr(s5))
        s, mi, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
        orc.lib().pvr_mstep_finish(_ptr(s5), C.c_int(it), C.c_float(step), C.byref(s), C.byref(mi), C.byref(mm))
        return s.value, mi.value, mm.value

    def rs_Scale(self):
        sc = np.zeros(max(self.n, 1), np.float32)
        orc.lib().pvr_scale(self.n, self.pbx, self.pby, _ptr(self.patches), _ptr(self.weights), _ptr(self.simpatches),
                            _ptr(self.simweights), _ptr(sc))
        self.scales = sc[:self.n].copy()
        return self.scales.copy()

    def debugWeights(self): return self.weights.copy()
    def debugSimpatches(self): return self.simpatches.copy()
    def debugSimweights(self): return self.simweights.copy()
    def debugSiminside(self): return self.siminside.copy()
    def debugConfidenceMap(self): return self.cmap.copy()
    def debugAddon(self): return self.addon.copy()
    def debugPSFsums(self): return self.psf_sums.copy()


def host_patch_em(patches_per_stack, patch_potential, scale, patch_weight, step, state5):
    pps = np.ascontiguousarray(patches_per_stack, np.int32)
    pot = np.ascontiguousarray(patch_potential, np.float32)
    sc = np.ascontiguousarray(scale, np.float32)
    used = np.zeros(max(int(pps.sum()), 1), np.float32)
    orc.lib().pvr_host_patch_em(C.c_int(pps.size), _ptr(pps), _ptr(pot), _ptr(sc), _ptr(patch_weight), C.c_float(step),
                                _ptr(state5), _ptr(used))
    return used[:int(pps.sum())]
