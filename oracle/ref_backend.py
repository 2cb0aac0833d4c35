"""Reference-backed twin of fetalreconstruction_b200.reconstruction.Reconstruction (TEST INFRASTRUCTURE ONLY).

Drives the reference's OWN CUDA path — the unmodified `class Reconstruction` of
/root/reference/source/reconstructionGPU2/reconstruction_cuda2.cu, compiled for sm_100a into
oracle/_ref/libref_cuda2.so by `make -C oracle ref` (see oracle/ref_shim/) — through the same method set as the
CUDA-backed class and the oracle-backed twin, so one driver (tests/golden/make_golden.py) produces comparable
outputs from all three.  Needs a GPU.

The reference constructor calls cudaDeviceReset(): import and use this module ONLY in a process of its own
(oracle/ref_runner.py), never next to torch.cuda or libsvr_b200.so contexts.

Reference quirk handled here so that all real slices are processed: `initStorageVolumes` drops the last slice
(`end = size.z - 1`, cuda2.cu:1439-1452), so one all-padding dummy slice is appended to every
per-slice input and trimmed from every output.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_cuda2.so")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def available() -> bool:
    return os.path.exists(LIB)


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


IDENT = np.eye(4, dtype=np.float32).ravel()


class RefReconstruction:
    one_shot = True

    def __init__(self, device=0, multithreaded=True):
        if not available():
            raise RuntimeError(f"{LIB} is missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(LIB)
        self.lib.ref_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.ref_create(C.c_int(device), C.c_int(int(multithreaded)), C.c_int(0)))
        self.S = self.Nx = self.Ny = 0
        self.vol_shape = (0, 0, 0)
        self.launch_count = 0

    # ---- helpers ----------------------------------------------------------------------------------
    @property
    def V(self):
        return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]

    @property
    def P(self):
        return self.Nx * self.Ny

    def _pad_vec(self, v, fill=1.0):
        return np.concatenate([_f32(v).ravel(), np.full(1, fill, np.float32)])

    def _pad_mats(self, m):
        return np.concatenate([_f32(m).reshape(-1, 16), IDENT[None]]).copy()

    def _get(self, kind, n, dtype=np.float32):
        out = np.zeros(n, dtype)
        rc = self.lib.ref_get(self.h, C.c_int(kind), out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    # ---- uploads -----------------------------------------------------------------------------------
    def InitReconstructionVolume(self, size, dim, data=None, sigma_bias=0.0):
        self.vol_shape = tuple(int(v) for v in size)
        self.vol_dim = tuple(float(v) for v in dim)
        d = None if data is None else _f32(data).ravel().copy()
        self.lib.ref_init_reconstruction_volume(self.h, *[C.c_int(v) for v in self.vol_shape],
                                                *[C.c_float(v) for v in self.vol_dim],
                                                None if d is None else _fp(d), C.c_float(sigma_bias))

    def setMask(self, size, dim, data, sigma_bias=0.0):
        m = _f32(data).ravel().copy()
        self.lib.ref_set_mask(self.h, *[C.c_int(int(v)) for v in size], *[C.c_float(float(v)) for v in dim], _fp(m),
                              C.c_float(sigma_bias))

    def initStorageVolumes(self, size, dim=(1, 1, 1)):
        self.Nx, self.Ny, self.S = (int(v) for v in size)
        self.lib.ref_init_storage_volumes(self.h, C.c_int(self.Nx), C.c_int(self.Ny), C.c_int(self.S + 1),
                                          *[C.c_float(float(v)) for v in dim])
        assert self.lib.ref_device_slices(self.h) == self.S

    def FillSlices(self, sdata, sizesX=None, sizesY=None):
        cube = np.concatenate([_f32(sdata).ravel(), np.full(self.P, -1.0, np.float32)])
        sx = np.full(self.S + 1, self.Nx, np.int32) if sizesX is None else np.append(np.asarray(sizesX, np.int32), self.Nx)
        sy = np.full(self.S + 1, self.Ny, np.int32) if sizesY is None else np.append(np.asarray(sizesY, np.int32), self.Ny)
        sx, sy = np.ascontiguousarray(sx, np.int32), np.ascontiguousarray(sy, np.int32)
        self.lib.ref_fill_slices(self.h, _fp(cube), _ip(sx), _ip(sy))

    def setSliceDims(self, slice_dims, quality_factor=1.0):
        d = _f32(slice_dims).reshape(-1, 3)
        d = np.concatenate([d, d[-1:]]).copy()
        self.lib.ref_set_slice_dims(self.h, _fp(d), C.c_float(quality_factor))

    def SetSliceMatrices(self, T, Tinv, a3, a4, a5, a6, reconI2W, reconW2I):
        args = [self._pad_mats(m) for m in (T, Tinv, a3, a4, a5, a6)]
        ri, rw = _f32(reconI2W).ravel().copy(), _f32(reconW2I).ravel().copy()
        self.lib.ref_set_slice_matrices(self.h, *[_fp(a) for a in args], _fp(ri), _fp(rw))

    def generatePSFVolume(self, CPUPSF, PSFsize, sliceVoxelDim, PSFdim, PSFI2W, PSFW2I, quality_factor):
        sz = np.asarray(PSFsize, np.int32).copy()
        sd, pd = _f32(sliceVoxelDim).copy(), _f32(PSFdim).copy()
        a, b = _f32(PSFI2W).ravel().copy(), _f32(PSFW2I).ravel().copy()
        # the reference uploads the PSF texture volume too; with USE_INFINITE_PSF_SUPPORT the kernels never
        # read it (SURVEY 8b "uploads constants only"), so a zero volume of the stated size is passed
        psf = np.zeros(int(sz[0]) * int(sz[1]) * int(sz[2]), np.float32) if CPUPSF is None else _f32(CPUPSF).ravel().copy()
        self.lib.ref_generate_psf_volume(self.h, _fp(psf), _ip(sz), _fp(sd), _fp(pd), _fp(a), _fp(b),
                                         C.c_float(quality_factor))

    def UpdateScaleVector(self, scales, slices_weights):
        s, w = self._pad_vec(scales), self._pad_vec(slices_weights)
        self.lib.ref_update_scale_vector(self.h, _fp(s), _fp(w))

    def UpdateSliceWeights(self, w):
        w = self._pad_vec(w)
        self.lib.ref_update_slice_weights(self.h, _fp(w))

    # ---- hot path ----------------------------------------------------------------------------------
    def InitializeEMValues(self):
        self.lib.ref_initialize_em_values(self.h)

    def GaussianReconstruction(self):
        vn = np.zeros(8, np.int32)
        n = self.lib.ref_gaussian_reconstruction(self.h, _ip(vn), C.c_int(8))
        self.voxel_num_per_device = vn[:n].copy()          # what the reference returns (one number per device, Q10)
        cnt = self.debugVoxelCount().reshape(self.S, self.P)
        return (cnt > 0).sum(1).astype(np.int32)           # per-slice counts from sliceVoxel_count (deviation D4)

    def gaussian_reconstruction_local(self):
        pass

    def gaussian_reconstruction_finish(self):
        return self.GaussianReconstruction()

    def SimulateSlices(self):
        out = np.zeros(self.S + 1, np.uint8)
        self.lib.ref_simulate_slices(self.h, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return out[:self.S].astype(bool)

    def InitializeRobustStatistics(self):
        s = C.c_float(0)
        self.lib.ref_initialize_robust_statistics(self.h, C.byref(s))
        return float(s.value)

    def initialize_robust_statistics_local(self):
        return np.array([self.InitializeRobustStatistics(), 1.0], np.float64)

    def EStep(self, m, sigma, mix):
        pot = np.zeros(self.S + 1, np.float32)
        self.lib.ref_estep(self.h, C.c_float(m), C.c_float(sigma), C.c_float(mix), _fp(pot))
        return pot[:self.S].copy()

    def MStep(self, it, step, sigma, mix, m):
        s, x, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
        self.lib.ref_mstep(self.h, C.c_int(it), C.c_float(step), C.byref(s), C.byref(x), C.byref(mm))
        return float(s.value), float(x.value), float(mm.value)

    def CalculateScaleVector(self):
        sc = np.ones(self.S + 1, np.float32)
        self.lib.ref_calculate_scale_vector(self.h, _fp(sc))
        return sc[:self.S].copy()

    def Superresolution(self, it, slice_weight, adaptive, alpha, min_i, max_i, delta, lambda_, *a):
        w = self._pad_vec(slice_weight)
        self.lib.ref_superresolution(self.h, C.c_int(it), _fp(w), C.c_int(int(adaptive)), C.c_float(alpha),
                                     C.c_float(min_i), C.c_float(max_i), C.c_float(delta), C.c_float(lambda_),
                                     C.c_int(0), C.c_float(0.0), C.c_float(0.0))

    def superresolution_local(self, slice_weight=None):
        self._sr_weight = slice_weight

    def superresolution_finish(self, adaptive, alpha, min_i, max_i, delta, lambda_):
        self.Superresolution(1, self._sr_weight, adaptive, alpha, min_i, max_i, delta, lambda_)

    def maskVolume(self):
        self.lib.ref_mask_volume(self.h)

    def ScaleVolume(self):
        self.lib.ref_scale_volume(self.h)
        return float("nan")                                # the reference does not report the factor

    def RestoreSliceIntensities(self, stack_factors, stack_index):
        f = _f32(stack_factors).copy()
        idx = np.ascontiguousarray(np.append(np.asarray(stack_index, np.int32), 0), np.int32)
        self.lib.ref_restore_slice_intensities(self.h, _fp(f), C.c_int(len(f)), _ip(idx))

    # ---- registration ------------------------------------------------------------------------------
    def initRegStorageVolumes(self, size, dim):
        self.regW, self.regH, self.regS = (int(v) for v in size)
        assert self.regS == self.S, "registration slices follow initStorageVolumes (shared slice ranges)"
        self.lib.ref_reg_init_storage(self.h, C.c_int(self.regW), C.c_int(self.regH), C.c_int(self.regS + 1),
                                      *[C.c_float(float(v)) for v in dim])
        self.reg_evaluations = 0

    def FillRegSlices(self, sdata, slices_resampledI2W=None):
        cube = np.concatenate([_f32(sdata).ravel(), np.full(self.regW * self.regH, -1.0, np.float32)])
        m = self._pad_mats(np.tile(IDENT, (self.regS, 1)) if slices_resampledI2W is None else slices_resampledI2W)
        self.lib.ref_reg_fill_slices(self.h, _fp(cube), _fp(m))

    def updateResampledSlicesI2W(self, ofsSlice):
        m = self._pad_mats(ofsSlice)
        self.lib.ref_reg_update_slices_i2w(self.h, _fp(m))

    def prepareSliceToVolumeReg(self):
        self.lib.ref_reg_prepare(self.h)

    def setRegSchedule(self, n_levels=2, n_steps=4, n_iterations=20):
        self.lib.ref_reg_set_schedule(self.h, C.c_int(n_levels), C.c_int(n_steps), C.c_int(n_iterations))

    def registerSlicesToVolume(self, transf):
        t = self._pad_mats(transf)
        self.lib.ref_reg_register(self.h, _fp(t))
        return t[:self.S].copy()

    def evaluateCostsMultipleSlices(self, transf, level=0):
        t = self._pad_mats(transf)
        sim = np.zeros(self.S + 1, np.float32)
        self.lib.ref_reg_evaluate(self.h, _fp(t), C.c_int(level), _fp(sim))
        return sim[:self.S].copy()

    # ---- read-backs --------------------------------------------------------------------------------
    def syncCPU(self):
        out = np.zeros(self.V, np.float32)
        self.lib.ref_sync_cpu(self.h, _fp(out))
        return out

    def getVolWeights(self):
        out = np.zeros(self.V, np.float32)
        self.lib.ref_get_vol_weights(self.h, _fp(out))
        return out

    def _slice_buf(self, kind, dtype=np.float32):
        return self._get(kind, self.S * self.P, dtype)

    def debugWeights(self): return self._slice_buf(0)
    def debugSimslices(self): return self._slice_buf(1)
    def debugSimweights(self): return self._slice_buf(2)
    def debugSiminside(self): return self._slice_buf(3, np.int8)
    def debugv_PSF_sums(self): return self._slice_buf(4)
    def debugVoxelCount(self): return self._slice_buf(5, np.int32)
    def debugAddon(self): return self._get(6, self.V)
    def debugConfidenceMap(self): return self._get(7, self.V)
    def debugSlices(self): return self._slice_buf(8)
