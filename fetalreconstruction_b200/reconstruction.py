"""Host-side mirror of the reference's device-library interface, `class Reconstruction`
(source/reconstructionGPU2/include/reconstruction_cuda2.cuh:92-341), on top of the C ABI.

Method names, argument meaning and call order are the reference's; every method is a thin
marshalling shim over one svr_* entry point of libsvr_b200.so.  Vectors are numpy arrays, Matrix4
is a (16,) / (S,16) float32 row-major array.  Errors raise `SVRError` (the reference prints and
exit()s, reconstruction_cuda2.cuh:78-86).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class SVRError(RuntimeError):
    pass


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise SVRError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class _DeviceArray:
    """Exposes a context-owned device buffer through __cuda_array_interface__ (for torch / NCCL)."""

    def __init__(self, ptr: int, n: int, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2,
                                         "strides": None}


class Reconstruction:
    """One GPU = one rank.  ref: Reconstruction(std::vector<int> dev, bool multiThreadedGPU)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.svr_create(C.byref(h), int(device))
        if rc != 0:
            raise SVRError(self._lib.svr_last_error(None).decode())
        self._h = h
        self.device = device
        self.S = self.Nx = self.Ny = 0
        self.vol_shape = (0, 0, 0)

    # -- plumbing -------------------------------------------------------------------------------
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

    def set_stream(self, cuda_stream_ptr: int | None):
        self._ck(self._lib.svr_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    TUNE_SCATTER, TUNE_SIMULATE, TUNE_REGULARIZE = 0, 1, 2

    def set_tuning(self, key: int, value: int):
        """Kernel-variant selection for A/B measurements (svr_set_tuning)."""
        self._ck(self._lib.svr_set_tuning(self._h, int(key), int(value)))

    def set_async(self, on: bool = True):
        """Entry points that return no host data only enqueue their work (svr_set_async)."""
        self._ck(self._lib.svr_set_async(self._h, int(bool(on))))

    def synchronize(self):
        self._ck(self._lib.svr_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.svr_launch_count(self._h))

    KERNEL_KINDS = ("gaussian", "simulate", "superres", "regularize", "em", "reg_eval", "estep", "mstep", "scale", "robust_init")

    def profile_enable(self, on: bool = True):
        self._ck(self._lib.svr_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self._ck(self._lib.svr_profile_reset(self._h))

    def profile_read(self) -> dict:
        """{kind: (total device ms, launches)} measured with CUDA events on the launching stream."""
        out = {}
        for i, name in enumerate(self.KERNEL_KINDS):
            ms, n = C.c_double(), C.c_int64()
            self._ck(self._lib.svr_profile_read(self._h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    @property
    def V(self):
        return self.vol_shape[0] * self.vol_shape[1] * self.vol_shape[2]

    @property
    def NP(self):
        return self.S * self.Nx * self.Ny

    # -- uploads --------------------------------------------------------------------------------
    def InitReconstructionVolume(self, size, dim, data=None, sigma_bias=0.0):
        sx, sy, sz = (int(v) for v in size)
        d = None if data is None else _f32(np.asarray(data).ravel(), (sx * sy * sz,))
        self._ck(self._lib.svr_init_reconstruction_volume(self._h, sx, sy, sz, float(dim[0]), float(dim[1]),
                                                          float(dim[2]), _p(d)))
        self.vol_shape = (sx, sy, sz)

    def setMask(self, size, dim, data, sigma_bias=0.0):
        sx, sy, sz = (int(v) for v in size)
        m = _f32(np.asarray(data).ravel(), (sx * sy * sz,))
        self._ck(self._lib.svr_set_mask(self._h, sx, sy, sz, _p(m)))

    def initStorageVolumes(self, size, dim=None):
        Nx, Ny, S = (int(v) for v in size)
        self._ck(self._lib.svr_init_storage_volumes(self._h, Nx, Ny, S))
        self.Nx, self.Ny, self.S = Nx, Ny, S

    def FillSlices(self, sdata, sizesX=None, sizesY=None):
        cube = _f32(np.asarray(sdata).ravel(), (self.NP,))
        sx = None if sizesX is None else np.ascontiguousarray(sizesX, np.int32)
        sy = None if sizesY is None else np.ascontiguousarray(sizesY, np.int32)
        self._ck(self._lib.svr_fill_slices(self._h, _p(cube), _p(sx), _p(sy)))

    def setSliceDims(self, slice_dims, quality_factor=1.0):
        d = _f32(slice_dims, (self.S, 3))
        self._ck(self._lib.svr_set_slice_dims(self._h, _p(d), float(quality_factor)))

    def SetSliceMatrices(self, matSliceTransforms, matInvSliceTransforms, matsI2Winit, matsW2Iinit, matsI2W, matsW2I,
                         reconI2W, reconW2I):
        # argument order of the reference (the *init pairs are unused by its kernels, cuda2.cu:870-907)
        t = _f32(matSliceTransforms, (self.S, 16))
        ti = _f32(matInvSliceTransforms, (self.S, 16))
        a = _f32(matsI2W, (self.S, 16))
        b = _f32(matsW2I, (self.S, 16))
        ri, rw = _f32(np.asarray(reconI2W).ravel(), (16,)), _f32(np.asarray(reconW2I).ravel(), (16,))
        self._ck(self._lib.svr_set_slice_matrices(self._h, _p(t), _p(ti), _p(a), _p(b), _p(ri), _p(rw)))

    def generatePSFVolume(self, CPUPSF, PSFsize, sliceVoxelDim, PSFdim, PSFI2W, PSFW2I, quality_factor):
        sz = np.ascontiguousarray(PSFsize, np.int32)
        m = _f32(np.asarray(PSFI2W).ravel(), (16,))
        self._ck(self._lib.svr_generate_psf_volume(self._h, _p(sz), _p(m), float(quality_factor)))

    def UpdateScaleVector(self, scales, slices_weights):
        s, w = _f32(scales, (self.S,)), _f32(slices_weights, (self.S,))
        self._ck(self._lib.svr_update_scale_vector(self._h, _p(s), _p(w)))

    def UpdateSliceWeights(self, slices_weights):
        w = _f32(slices_weights, (self.S,))
        self._ck(self._lib.svr_update_slice_weights(self._h, _p(w)))

    def UpdateReconstructed(self, vsize, data):
        d = _f32(np.asarray(data).ravel(), (self.V,))
        self._ck(self._lib.svr_update_reconstructed(self._h, _p(d)))

    # -- hot path -------------------------------------------------------------------------------
    def InitializeEMValues(self):
        self._ck(self._lib.svr_initialize_em_values(self._h))

    def GaussianReconstruction(self):
        """Returns voxel_num (per slice, deviation D4)."""
        vn = np.zeros(max(self.S, 1), np.int32)
        self._ck(self._lib.svr_gaussian_reconstruction(self._h, _p(vn)))
        return vn[:self.S]

    def SimulateSlices(self, fetch_inside=True):
        """Returns slice_inside (bool per slice), or None without the read-back (fetch_inside=False)."""
        if not fetch_inside:
            self._ck(self._lib.svr_simulate_slices(self._h, None))
            return None
        si = np.zeros(max(self.S, 1), np.uint8)
        self._ck(self._lib.svr_simulate_slices(self._h, _p(si)))
        return si[:self.S].astype(bool)

    def InitializeRobustStatistics(self) -> float:
        s = C.c_float()
        self._ck(self._lib.svr_initialize_robust_statistics(self._h, C.byref(s)))
        return s.value

    def EStep(self, m, sigma, mix):
        """Returns slice_potential."""
        pot = np.zeros(max(self.S, 1), np.float32)
        self._ck(self._lib.svr_estep(self._h, float(m), float(sigma), float(mix), _p(pot)))
        return pot[:self.S]

    def MStep(self, iter, step, sigma, mix, m):
        s, mi, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
        self._ck(self._lib.svr_mstep(self._h, int(iter), float(step), C.byref(s), C.byref(mi), C.byref(mm)))
        return s.value, mi.value, mm.value

    def CalculateScaleVector(self):
        sc = np.zeros(max(self.S, 1), np.float32)
        self._ck(self._lib.svr_calculate_scale_vector(self._h, _p(sc)))
        return sc[:self.S]

    def Superresolution(self, iter, slice_weight, adaptive, alpha, min_intensity, max_intensity, delta, lambda_,
                        global_bias_correction=False, sigma_bias=0.0, low_intensity_cutoff=0.0):
        w = None if slice_weight is None else _f32(slice_weight, (self.S,))
        self._ck(self._lib.svr_superresolution(self._h, int(iter), _p(w), int(bool(adaptive)), float(alpha),
                                               float(min_intensity), float(max_intensity), float(delta),
                                               float(lambda_)))

    def maskVolume(self):
        self._ck(self._lib.svr_mask_volume(self._h))

    def ScaleVolume(self) -> float:
        s = C.c_float()
        self._ck(self._lib.svr_scale_volume(self._h, C.byref(s)))
        return s.value

    def RestoreSliceIntensities(self, stack_factors, stack_index):
        f = _f32(stack_factors)
        i = np.ascontiguousarray(stack_index, np.int32)
        self._ck(self._lib.svr_restore_slice_intensities(self._h, _p(f), int(f.size), _p(i)))

    # -- slice-to-volume registration (--useGPUReg), reconstruction_cuda2.cuh:326-338 ------------
    def initRegStorageVolumes(self, size, dim):
        W, H, S = (int(v) for v in size)
        self._ck(self._lib.svr_reg_init_storage(self._h, W, H, S, float(dim[0]), float(dim[1]), float(dim[2])))
        self.regW, self.regH, self.regS = W, H, S

    def FillRegSlices(self, sdata, slices_resampledI2W=None):
        cube = _f32(np.asarray(sdata).ravel(), (self.regW * self.regH * self.regS,))
        m = None if slices_resampledI2W is None else _f32(slices_resampledI2W, (self.regS, 16))
        self._ck(self._lib.svr_reg_fill_slices(self._h, _p(cube), _p(m)))

    def resampleRegSlices(self, out_i2w, in_w2i, in_sizes, out_sizes, slices_resampledI2W=None):
        """Device form of PrepareRegistrationSlices' resampling (svr_reg_resample_slices): the slices uploaded by
        FillSlices are resampled on the device into the registration cube.  out_i2w / in_w2i [S,4,4] float64: the resampled
        slice's image-to-world and the slice's world-to-image matrices."""
        m = np.ascontiguousarray(np.concatenate([np.asarray(out_i2w, np.float64).reshape(self.regS, 4, 4)[:, :3, :].reshape(self.regS, 12),
                                                 np.asarray(in_w2i, np.float64).reshape(self.regS, 4, 4)[:, :3, :].reshape(self.regS, 12)], 1))
        a = np.ascontiguousarray(in_sizes, np.int32).reshape(self.regS, 2)
        o = np.ascontiguousarray(out_sizes, np.int32).reshape(self.regS, 2)
        i2w = None if slices_resampledI2W is None else _f32(slices_resampledI2W, (self.regS, 16))
        self._ck(self._lib.svr_reg_resample_slices(self._h, _p(m), _p(a), _p(o), _p(i2w)))

    def updateResampledSlicesI2W(self, ofsSlice):
        m = _f32(ofsSlice, (self.regS, 16))
        self._ck(self._lib.svr_reg_update_slices_i2w(self._h, _p(m)))

    def prepareSliceToVolumeReg(self):
        self._ck(self._lib.svr_reg_prepare(self._h))

    def setRegSchedule(self, n_levels=2, n_steps=4, n_iterations=20):
        self._ck(self._lib.svr_reg_set_schedule(self._h, int(n_levels), int(n_steps), int(n_iterations)))

    def registerSlicesToVolume(self, transf):
        """transf [S,16] -> registered [S,16] (the reference updates its argument in place)."""
        t = _f32(transf, (self.regS, 16)).copy()
        self._ck(self._lib.svr_reg_register(self._h, _p(t)))
        return t

    def evaluateCostsMultipleSlices(self, transf, level=0):
        """Similarity of every slice for the given transforms (all slices active)."""
        t = _f32(transf, (self.regS, 16))
        out = np.zeros(max(self.regS, 1), np.float32)
        self._ck(self._lib.svr_reg_evaluate(self._h, _p(t), int(level), _p(out)))
        return out[:self.regS]

    @property
    def reg_evaluations(self) -> int:
        return int(self._lib.svr_reg_evaluations(self._h))

    def debugRegSlices(self, blurred=False):
        out = np.empty(max(self.regW * self.regH * self.regS, 1), np.float32)
        self._ck(self._lib.svr_reg_debug_get(self._h, 1 if blurred else 0, _p(out)))
        return out[:self.regW * self.regH * self.regS].reshape(self.regS, self.regH, self.regW)

    # -- downloads ------------------------------------------------------------------------------
    def syncCPU(self, out=None):
        """Volume -> host.  `out` may be a caller-owned float32 buffer of V elements (e.g. pinned memory: the copy then runs
        at PCIe speed instead of through the driver's staging buffer)."""
        if out is None:
            out = np.empty(self.V, np.float32)
        assert out.dtype == np.float32 and out.size >= self.V and out.flags["C_CONTIGUOUS"]
        self._ck(self._lib.svr_sync_cpu(self._h, _p(out)))
        return out

    def getVolWeights(self):
        out = np.empty(self.V, np.float32)
        self._ck(self._lib.svr_get_vol_weights(self._h, _p(out)))
        return out

    def _debug(self, kind, n, dtype):
        out = np.empty(max(n, 1), dtype)
        self._ck(self._lib.svr_debug_get(self._h, kind, _p(out)))
        return out[:n]

    def debugWeights(self): return self._debug(0, self.NP, np.float32)
    def debugSimslices(self): return self._debug(1, self.NP, np.float32)
    def debugSimweights(self): return self._debug(2, self.NP, np.float32)
    def debugSiminside(self): return self._debug(3, self.NP, np.int8)
    def debugConfidenceMap(self): return self._debug(4, self.V, np.float32)
    def debugAddon(self): return self._debug(5, self.V, np.float32)
    def debugv_PSF_sums(self): return self._debug(6, self.NP, np.float32)
    def getSlicesVol_debug(self): return self._debug(7, self.NP, np.float32)
    def debugVoxelCount(self): return self._debug(8, self.NP, np.int32)
    def debugSlicesRestored(self): return self._debug(9, self.NP, np.float32)
    def debugScalesDevice(self): return self._debug(10, self.S, np.float32)
    def debugWindowStats(self): return self._debug(12, 32, np.uint32)

    # -- multi-rank split phases (include/svr_abi.h, section "multi-rank") ----------------------
    def gaussian_reconstruction_local(self):
        self._ck(self._lib.svr_gaussian_reconstruction_local(self._h))

    def gaussian_reconstruction_finish(self):
        vn = np.zeros(max(self.S, 1), np.int32)
        self._ck(self._lib.svr_gaussian_reconstruction_finish(self._h, _p(vn)))
        return vn[:self.S]

    def superresolution_local(self, slice_weight=None):
        w = None if slice_weight is None else _f32(slice_weight, (self.S,))
        self._ck(self._lib.svr_superresolution_local(self._h, _p(w)))

    def superresolution_finish(self, adaptive, alpha, min_intensity, max_intensity, delta, lambda_):
        self._ck(self._lib.svr_superresolution_finish(self._h, int(bool(adaptive)), float(alpha), float(min_intensity),
                                                      float(max_intensity), float(delta), float(lambda_)))

    def mstep_local(self):
        out = (C.c_double * 5)()
        self._ck(self._lib.svr_mstep_local(self._h, out))
        return np.array(out[:], np.float64)

    def initialize_robust_statistics_local(self):
        out = (C.c_double * 2)()
        self._ck(self._lib.svr_initialize_robust_statistics_local(self._h, out))
        return np.array(out[:], np.float64)

    def scale_volume_local(self):
        out = (C.c_double * 2)()
        self._ck(self._lib.svr_scale_volume_local(self._h, out))
        return np.array(out[:], np.float64)

    def scale_volume_apply(self, scale):
        self._ck(self._lib.svr_scale_volume_apply(self._h, float(scale)))

    def accumulator(self):
        """float32 view [2V] of the interleaved scatter accumulator, for the caller's all-reduce."""
        ptr, nb = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.svr_device_buffer(self._h, 0, C.byref(ptr), C.byref(nb)))
        return _DeviceArray(ptr.value, nb.value // 4, self)


def mstep_finish(sums5, iter, step, sigma, mix, m):
    """Pure host part of Reconstruction::MStep (reconstruction_cuda2.cu:3056-3071)."""
    lib = _lib.load()
    arr = (C.c_double * 5)(*[float(v) for v in sums5])
    s, mi, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
    if lib.svr_mstep_finish(arr, int(iter), float(step), C.byref(s), C.byref(mi), C.byref(mm)) != 0:
        raise SVRError("svr_mstep_finish: bad argument")
    return s.value, mi.value, mm.value


def host_slice_em(slice_potential, scale, slice_weight, force_excluded, small_slices, step, state5):
    """irtkReconstruction::EStepGPU's slice-level EM (irtkReconstructionGPU.cc:3203-3420); in place."""
    lib = _lib.load()
    fe = np.ascontiguousarray(force_excluded, np.int32)
    sm = np.ascontiguousarray(small_slices, np.int32)
    assert slice_potential.dtype == np.float32 and slice_weight.dtype == np.float32 and state5.dtype == np.float32
    sc = _f32(scale)
    rc = lib.svr_host_slice_em(int(slice_potential.size), _p(slice_potential), _p(sc), _p(slice_weight), _p(fe),
                               int(fe.size), _p(sm), int(sm.size), float(step), _p(state5))
    if rc != 0:
        raise SVRError("svr_host_slice_em: bad argument")


def host_small_slices(voxel_num):
    lib = _lib.load()
    vn = np.ascontiguousarray(voxel_num, np.int32)
    out = np.zeros(max(vn.size, 1), np.int32)
    n = C.c_int()
    if lib.svr_host_small_slices(int(vn.size), _p(vn), _p(out), C.byref(n)) != 0:
        raise SVRError("svr_host_small_slices: bad argument")
    return out[:n.value].copy()


def host_partition_strided(slices_per_stack, nranks, rank):
    """Indices (stack-major global order) of the slices of `rank`: every nranks-th slice of every stack."""
    lib = _lib.load()
    sps = np.ascontiguousarray(slices_per_stack, np.int32)
    out = np.zeros(max(int(sps.sum()), 1), np.int32)
    n = C.c_int()
    if lib.svr_host_partition_strided(int(sps.size), _p(sps), int(nranks), int(rank), _p(out), C.byref(n)) != 0:
        raise SVRError("svr_host_partition_strided: bad argument")
    return out[:n.value].copy()


def host_partition(slices_per_stack, nranks, rank):
    lib = _lib.load()
    sp = np.ascontiguousarray(slices_per_stack, np.int32)
    b, e = C.c_int(), C.c_int()
    if lib.svr_host_partition(int(sp.size), _p(sp), int(nranks), int(rank), C.byref(b), C.byref(e)) != 0:
        raise SVRError("svr_host_partition: bad argument")
    return b.value, e.value
