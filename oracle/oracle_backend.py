"""Oracle-backed twin of fetalreconstruction_b200.reconstruction.Reconstruction (TEST INFRASTRUCTURE ONLY: tests/, smoke() and bench.py's CPU legs).

Same method set and call semantics as the CUDA-backed class, implemented with oracle/svr_oracle.c on
numpy arrays.  Used (a) as the checker the CUDA path is compared against through the SAME host
pipeline, and (b) to run the multi-rank host logic of pipeline.py on CPU with gloo.
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as orc


class _DS:
    """The subset of phantom.Dataset that oracle.Geometry reads."""


class OracleReconstruction:
    def __init__(self, device=0):
        self.S = self.Nx = self.Ny = 0
        self.vol_shape = (0, 0, 0)
        self.psf_c = np.zeros(3, np.float32)
        self.launch_count = 0

    @property
    def V(self):
        return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]

    @property
    def NP(self):
        return self.S * self.Nx * self.Ny

    # uploads
    def InitReconstructionVolume(self, size, dim, data=None, sigma_bias=0.0):
        self.vol_shape = tuple(int(v) for v in size)
        self.reg_voxel = float(dim[0])
        self.recon = np.zeros(self.V, np.float32) if data is None else np.array(data, np.float32).ravel()
        self.volw = np.zeros(self.V, np.float32)
        self.addon = np.zeros(self.V, np.float32)
        self.cmap = np.zeros(self.V, np.float32)
        self.acc = np.zeros(2 * self.V, np.float32)

    def setMask(self, size, dim, data, sigma_bias=0.0):
        self.mask = np.array(data, np.float32).ravel()

    def initStorageVolumes(self, size, dim=None):
        self.Nx, self.Ny, self.S = (int(v) for v in size)
        n = self.NP
        self.slices = np.zeros(n, np.float32)
        self.weights = np.zeros(n, np.float32)
        self.simslices = np.zeros(n, np.float32)
        self.simweights = np.zeros(n, np.float32)
        self.siminside = np.zeros(n, np.int8)
        self.psf_sums = np.zeros(n, np.float32)
        self.voxel_count = np.zeros(n, np.int32)
        self.scales = np.ones(self.S, np.float32)          # device scales
        self.h_scales = np.ones(self.S, np.float32)        # host h_scales
        self.slice_weights = np.ones(self.S, np.float32)

    def FillSlices(self, sdata, sizesX=None, sizesY=None):
        self.slices = np.array(sdata, np.float32).ravel()
        self.slices_restore = self.slices.copy()

    def setSliceDims(self, slice_dims, quality_factor=1.0):
        self.dims = np.ascontiguousarray(slice_dims, np.float32)

    def SetSliceMatrices(self, T, Tinv, I2Winit, W2Iinit, I2W, W2I, reconI2W, reconW2I):
        self.trans = np.ascontiguousarray(T, np.float32)
        self.trans_inv = np.ascontiguousarray(Tinv, np.float32)
        self.i2w = np.ascontiguousarray(I2W, np.float32)
        self.w2i = np.ascontiguousarray(W2I, np.float32)
        self.recon_i2w = np.ascontiguousarray(reconI2W, np.float32).ravel()
        self.recon_w2i = np.ascontiguousarray(reconW2I, np.float32).ravel()

    def generatePSFVolume(self, CPUPSF, PSFsize, sliceVoxelDim, PSFdim, PSFI2W, PSFW2I, quality_factor):
        m = np.asarray(PSFI2W, np.float32).reshape(4, 4)
        c = [np.float32((s - 1) * 0.5) for s in PSFsize]
        for r in range(3):
            self.psf_c[r] = np.float32(np.float32(np.float32(m[r, 0] * c[0]) + np.float32(m[r, 1] * c[1]))
                                       + np.float32(m[r, 2] * c[2])) + m[r, 3]

    def _geom(self):
        ds = _DS()
        ds.i2w, ds.w2i, ds.trans, ds.trans_inv, ds.dims = self.i2w, self.w2i, self.trans, self.trans_inv, self.dims
        ds.slices = np.empty((self.S, self.Ny, self.Nx), np.float32)
        vx, vy, vz = self.vol_shape
        ds.mask = np.empty((vz, vy, vx), np.float32)
        ds.recon_i2w, ds.recon_w2i, ds.psf_c = self.recon_i2w, self.recon_w2i, self.psf_c
        return orc.Geometry(ds)

    def UpdateScaleVector(self, scales, slices_weights):
        self.scales = np.array(scales, np.float32)
        self.h_scales = self.scales.copy()
        self.slice_weights = np.array(slices_weights, np.float32)

    def UpdateSliceWeights(self, w):
        self.slice_weights = np.array(w, np.float32)

    # hot path
    def InitializeEMValues(self):
        self.weights = orc.initialize_em_values(self.slices)

    def gaussian_reconstruction_local(self):
        n = self.NP
        self.weights = np.zeros(n, np.float32)
        self.simweights = np.zeros(n, np.float32)
        self.simslices = np.zeros(n, np.float32)
        self.siminside = np.zeros(n, np.int8)
        rec, vw, cnt, num = orc.gaussian_reconstruction(self._geom(), self.slices, self.scales, self.mask, self.psf_sums,
                                                        equalize=False)
        self.voxel_count, self._voxel_num = cnt, num
        self.acc = np.stack([rec, vw], 1).ravel().copy()

    def gaussian_reconstruction_finish(self):
        a = self.acc.reshape(-1, 2)
        num, den = a[:, 0], a[:, 1]
        self.volw = den.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            self.recon = np.where(den != 0, num / np.where(den != 0, den, 1), num).astype(np.float32)
        return self._voxel_num

    def GaussianReconstruction(self):
        self.gaussian_reconstruction_local()
        return self.gaussian_reconstruction_finish()

    def SimulateSlices(self):
        inside = orc.simulate_slices(self._geom(), self.slices, self.psf_sums, self.recon, self.mask, self.simslices,
                                     self.simweights, self.siminside)
        return inside.astype(bool)

    def initialize_robust_statistics_local(self):
        _, s, n = orc.initialize_robust_statistics(self.slices, self.siminside, self.simslices, self.simweights)
        return np.array([s, n], np.float64)

    def InitializeRobustStatistics(self):
        s2 = self.initialize_robust_statistics_local()
        return float(np.float32(s2[0]) / np.float32(s2[1]))

    def EStep(self, m, sigma, mix):
        self.weights, pot = orc.estep(self._geom(), self.slices, self.simslices, self.simweights, self.scales, m, sigma, mix)
        return pot

    def mstep_local(self):
        return orc.mstep_sums(self._geom(), self.slices, self.weights, self.simslices, self.simweights, self.h_scales)

    def MStep(self, it, step, sigma, mix, m):
        return orc.mstep_finish(self.mstep_local(), it, step, sigma, mix, m)

    def CalculateScaleVector(self):
        new = orc.calculate_scale_vector(self._geom(), self.slices, self.weights, self.simslices, self.simweights)
        self.scales = self.h_scales.copy()        # the device lags one call behind (cuda2.cu:3238 vs 3195)
        self.h_scales = new.copy()
        return new

    def superresolution_local(self, slice_weight=None):
        if slice_weight is not None:
            self.UpdateSliceWeights(slice_weight)
        addon, cmap = orc.superresolution_backproject(self._geom(), self.slices, self.weights, self.simslices,
                                                      self.slice_weights, self.scales, self.mask, self.psf_sums)
        self.acc = np.stack([addon, cmap], 1).ravel().copy()

    def superresolution_finish(self, adaptive, alpha, min_i, max_i, delta, lambda_):
        a = self.acc.reshape(-1, 2)
        self.addon, self.cmap = a[:, 0].copy(), a[:, 1].copy()
        orc.regularize(self._geom(), self.recon, self.addon, self.cmap, adaptive, alpha, min_i, max_i, delta, lambda_)

    def Superresolution(self, it, slice_weight, adaptive, alpha, min_i, max_i, delta, lambda_, *a):
        self.superresolution_local(slice_weight)
        self.superresolution_finish(adaptive, alpha, min_i, max_i, delta, lambda_)

    def maskVolume(self):
        orc.mask_volume(self.recon, self.mask)

    def scale_volume_local(self):
        g = self._geom()
        P = self.Nx * self.Ny
        s, sw, ss, w = self.slices, self.simweights, self.simslices, self.weights
        ok = (s != -1) & (sw.astype(np.float64) > 0.99)
        slw = np.repeat(self.slice_weights, P)
        num = (w * slw * s * ss)[ok].astype(np.float64).sum()
        den = (w * slw * ss * ss)[ok].astype(np.float64).sum()
        return np.array([num, den], np.float64)

    def scale_volume_apply(self, scale):
        self.recon = np.where(self.recon > 0, self.recon * np.float32(scale), self.recon).astype(np.float32)

    def ScaleVolume(self):
        s2 = self.scale_volume_local()
        sc = float(np.float32(s2[0] / s2[1]))
        self.scale_volume_apply(sc)
        return sc

    # registration (oracle/reg_oracle.c)
    def initRegStorageVolumes(self, size, dim):
        self.regW, self.regH, self.regS = (int(v) for v in size)
        self.reg_schedule = (2, 4, 20)
        self.reg_evaluations = 0

    def FillRegSlices(self, sdata, slices_resampledI2W=None):
        self.reg_slices = np.array(sdata, np.float32).reshape(self.regS, self.regH, self.regW)

    def updateResampledSlicesI2W(self, ofsSlice):
        self.reg_ofs = np.ascontiguousarray(ofsSlice, np.float32)

    def prepareSliceToVolumeReg(self):
        vx, vy, vz = self.vol_shape
        self.reg_vol = self.recon.reshape(vz, vy, vx).copy()
        self.reg_schedule = (2, 4, 20)

    def setRegSchedule(self, n_levels=2, n_steps=4, n_iterations=20):
        self.reg_schedule = (n_levels, n_steps, n_iterations)

    def registerSlicesToVolume(self, transf):
        out, ev = orc.reg_register_slices(self.reg_slices, self.reg_ofs, self.reg_vol, self.reg_voxel, self.recon_w2i,
                                          transf, *self.reg_schedule)
        self.reg_evaluations = ev
        return out

    def evaluateCostsMultipleSlices(self, transf, level=0):
        return orc.reg_evaluate(self.reg_slices, self.reg_ofs, self.reg_vol, self.reg_voxel, self.recon_w2i, transf, level)

    def syncCPU(self):
        return self.recon.copy()

    def getVolWeights(self):
        return self.volw.copy()

    def debugWeights(self): return self.weights.copy()
    def debugSimslices(self): return self.simslices.copy()
    def debugSimweights(self): return self.simweights.copy()
    def debugSiminside(self): return self.siminside.copy()
    def debugv_PSF_sums(self): return self.psf_sums.copy()
    def debugAddon(self): return self.addon.copy()
    def debugConfidenceMap(self): return self.cmap.copy()
    def debugVoxelCount(self): return self.voxel_count.copy()

    def accumulator(self):
        return self.acc
