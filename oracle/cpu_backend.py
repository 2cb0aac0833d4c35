"""The reference's CPU (--useCPU) projections behind the backend method set (TEST / BASELINE INFRASTRUCTURE ONLY): the
sparse slice-to-volume matrix of irtkReconstruction::CoeffInit and the functors that apply it (oracle/cpu_path.c; Gaussian
PSF + trilinear splat, irtkReconstructionGPU.cc:2305-2860, 1090-1144, 3940-4022), with the robust statistics and the
regulariser of the shared restatement (oracle/svr_oracle.c) around them.  bench.py's cpu_baseline / --impl reference legs
time one outer iteration of fetalreconstruction_b200.pipeline.SVRPipeline on it: CoeffInit (recomputed every outer
iteration, as reconstruction.cc:941 does), 1 Gaussian reconstruction, 5 SimulateSlices, 4 Superresolution + EM.
Not pinned against the reference (its CPU path cannot be built here); see tests/test_cpu_path.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from oracle import oracle as orc
from oracle.oracle import _ptr
from oracle.oracle_backend import OracleReconstruction


class CpuPathReconstruction(OracleReconstruction):
    quality_factor = 1.0               # SpeedupOn (all but the last outer iteration, reconstruction.cc:916-923)

    def __init__(self, device=0):
        super().__init__(device)
        self._coeffs = None
        lib = orc.lib()
        lib.cpu_coeff_init.restype = C.c_void_p
        lib.cpu_coeff_nnz.restype = C.c_long

    def __del__(self):
        try:
            self._free()
        except Exception:              # interpreter shutdown
            pass

    def _free(self):
        if getattr(self, "_coeffs", None):
            orc.lib().cpu_coeff_free(C.c_void_p(self._coeffs))
            self._coeffs = None

    def coeff_init(self):
        """irtkReconstruction::CoeffInit (irtkReconstructionGPU.cc:2614-2673)."""
        self._free()
        vx, vy, vz = self.vol_shape
        inside = np.zeros(max(self.S, 1), np.uint8)
        self._coeffs = orc.lib().cpu_coeff_init(self.S, self.Nx, self.Ny, _ptr(self.slices), _ptr(self.i2w), _ptr(self.trans), _ptr(self.dims),
                                                _ptr(self.recon_w2i), vx, vy, vz, _ptr(self.mask), C.c_double(self.vol_dim[0]),
                                                C.c_double(self.quality_factor), _ptr(inside))
        self._volw = np.zeros(self.V, np.float64)
        orc.lib().cpu_volume_weights(C.c_void_p(self._coeffs), _ptr(self._volw))
        self.volw = self._volw.astype(np.float32)
        return inside[:self.S].astype(bool)

    @property
    def nnz(self):
        return int(orc.lib().cpu_coeff_nnz(C.c_void_p(self._coeffs)))

    def InitReconstructionVolume(self, size, dim, data=None, sigma_bias=0.0):
        super().InitReconstructionVolume(size, dim, data, sigma_bias)
        self.vol_dim = tuple(float(v) for v in dim)

    # ---- the three projections ------------------------------------------------------------------------------------
    def gaussian_reconstruction_local(self):
        n = self.NP
        self.weights = np.zeros(n, np.float32)
        self.simweights = np.zeros(n, np.float32)
        self.simslices = np.zeros(n, np.float32)
        self.siminside = np.zeros(n, np.int8)
        self.coeff_init()
        rec = np.zeros(self.V, np.float32)
        num = np.zeros(max(self.S, 1), np.int32)
        orc.lib().cpu_gaussian_reconstruction(C.c_void_p(self._coeffs), _ptr(self.slices), _ptr(self.scales), _ptr(self._volw), _ptr(rec), _ptr(num))
        self._rec, self._voxel_num = rec, num[:self.S].copy()
        # single rank only: the accumulator of the split-phase API is not formed (the CPU path is one process in the reference)
        self.acc = np.zeros(2 * self.V, np.float32)

    def gaussian_reconstruction_finish(self):
        self.recon = self._rec
        return self._voxel_num

    def SimulateSlices(self):
        inside = np.zeros(max(self.S, 1), np.int32)
        orc.lib().cpu_simulate_slices(C.c_void_p(self._coeffs), _ptr(self.slices), _ptr(self.recon), _ptr(self.mask), _ptr(self.simslices),
                                      _ptr(self.simweights), _ptr(self.siminside), _ptr(inside))
        return inside[:self.S].astype(bool)

    def superresolution_local(self, slice_weight=None):
        if slice_weight is not None:
            self.UpdateSliceWeights(slice_weight)
        addon, cmap = np.zeros(self.V, np.float32), np.zeros(self.V, np.float32)
        orc.lib().cpu_superresolution(C.c_void_p(self._coeffs), _ptr(self.slices), _ptr(self.weights), _ptr(self.simslices), _ptr(self.slice_weights),
                                      _ptr(self.scales), _ptr(addon), _ptr(cmap))
        self.acc = np.stack([addon, cmap], 1).ravel().copy()
