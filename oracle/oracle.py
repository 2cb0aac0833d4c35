"""ctypes front-end of the CPU oracle (oracle/svr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.  Pinned against the reference's
own CUDA path (see the header of svr_oracle.c and tests/test_ref_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsvr_oracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class _Mat4(C.Structure):
    _fields_ = [("m", C.c_float * 16)]


class _F3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class _Geom(C.Structure):
    _fields_ = [("S", C.c_int), ("Nx", C.c_int), ("Ny", C.c_int),
                ("vx", C.c_int), ("vy", C.c_int), ("vz", C.c_int),
                ("I2W", C.c_void_p), ("W2I", C.c_void_p), ("T", C.c_void_p), ("Tinv", C.c_void_p),
                ("dims", C.c_void_p), ("RI2W", _Mat4), ("RW2I", _Mat4), ("psf_c", _F3)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("svr_oracle.c", "reg_oracle.c", "pvr_oracle.c", "cpu_path.c", "Makefile")]
    newest = max(os.path.getmtime(f) for f in srcs if os.path.exists(f))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE, "clean", "all"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_psf_value.restype = C.c_float
        _lib.orc_psf_value.argtypes = [C.c_float] * 6
        _lib.orc_initialize_robust_statistics.restype = C.c_float
        _lib.orc_scale_volume.restype = C.c_float
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Geometry:
    """Keeps the numpy arrays alive next to the C struct."""

    def __init__(self, ds, sl=None):
        sl = slice(None) if sl is None else sl
        self.i2w = np.ascontiguousarray(ds.i2w[sl])
        self.w2i = np.ascontiguousarray(ds.w2i[sl])
        self.trans = np.ascontiguousarray(ds.trans[sl])
        self.trans_inv = np.ascontiguousarray(ds.trans_inv[sl])
        self.dims = np.ascontiguousarray(ds.dims[sl])
        self.S = self.i2w.shape[0]
        Ny, Nx = ds.slices.shape[1:]
        vz, vy, vx = ds.mask.shape
        self.Nx, self.Ny, self.vx, self.vy, self.vz = Nx, Ny, vx, vy, vz
        g = _Geom()
        g.S, g.Nx, g.Ny, g.vx, g.vy, g.vz = self.S, Nx, Ny, vx, vy, vz
        g.I2W, g.W2I, g.T, g.Tinv = _ptr(self.i2w), _ptr(self.w2i), _ptr(self.trans), _ptr(self.trans_inv)
        g.dims = _ptr(self.dims)
        g.RI2W = _Mat4((C.c_float * 16)(*ds.recon_i2w.tolist()))
        g.RW2I = _Mat4((C.c_float * 16)(*ds.recon_w2i.tolist()))
        g.psf_c = _F3(*[float(v) for v in ds.psf_c])
        self.c = g

    def set_transforms(self, trans, trans_inv):
        self.trans = np.ascontiguousarray(trans, np.float32)
        self.trans_inv = np.ascontiguousarray(trans_inv, np.float32)
        self.c.T, self.c.Tinv = _ptr(self.trans), _ptr(self.trans_inv)

    @property
    def V(self):
        return self.vx * self.vy * self.vz

    @property
    def npix(self):
        return self.S * self.Nx * self.Ny


def psf_value(p, d):
    return float(lib().orc_psf_value(*[float(v) for v in p], *[float(v) for v in d]))


def gaussian_reconstruction(g: Geometry, slices, scales, mask, psf_sums, equalize=True):
    recon = np.zeros(g.V, np.float32)
    volw = np.zeros(g.V, np.float32)
    count = np.zeros(g.npix, np.int32)
    num = np.zeros(g.S, np.int32)
    lib().orc_gaussian_reconstruction(C.byref(g.c), _ptr(slices), _ptr(scales), _ptr(mask), _ptr(recon), _ptr(volw),
                                      _ptr(psf_sums), _ptr(count), _ptr(num), C.c_int(int(equalize)))
    return recon, volw, count, num


def simulate_slices(g: Geometry, slices, psf_sums, recon, mask, simslices, simweights, siminside):
    inside = np.zeros(g.S, np.uint8)
    lib().orc_simulate_slices(C.byref(g.c), _ptr(slices), _ptr(psf_sums), _ptr(recon), _ptr(mask), _ptr(simslices),
                              _ptr(simweights), _ptr(siminside), _ptr(inside))
    return inside


def superresolution_backproject(g: Geometry, slices, weights, simslices, slice_weights, scales, mask, psf_sums):
    addon = np.zeros(g.V, np.float32)
    cmap = np.zeros(g.V, np.float32)
    lib().orc_superresolution_backproject(C.byref(g.c), _ptr(slices), _ptr(weights), _ptr(simslices),
                                          _ptr(slice_weights), _ptr(scales), _ptr(mask), _ptr(psf_sums), _ptr(addon),
                                          _ptr(cmap))
    return addon, cmap


def regularize(g: Geometry, recon, addon, cmap, adaptive, alpha, min_i, max_i, delta, lam):
    lib().orc_regularize(g.vx, g.vy, g.vz, _ptr(recon), _ptr(addon), _ptr(cmap), C.c_int(int(adaptive)),
                         C.c_float(alpha), C.c_float(min_i), C.c_float(max_i), C.c_float(delta), C.c_float(lam))


def initialize_em_values(slices):
    w = np.zeros(slices.size, np.float32)
    lib().orc_initialize_em_values(C.c_size_t(slices.size), _ptr(slices), _ptr(w))
    return w


def initialize_robust_statistics(slices, siminside, simslices, simweights):
    s, n = C.c_double(), C.c_double()
    sig = lib().orc_initialize_robust_statistics(C.c_size_t(slices.size), _ptr(slices), _ptr(siminside), _ptr(simslices),
                                                 _ptr(simweights), C.byref(s), C.byref(n))
    return float(sig), s.value, n.value


def estep(g: Geometry, slices, simslices, simweights, scales, m, sigma, mix):
    w = np.zeros(g.npix, np.float32)
    pot = np.zeros(g.S, np.float32)
    lib().orc_estep(g.S, g.Nx, g.Ny, _ptr(slices), _ptr(simslices), _ptr(simweights), _ptr(scales), C.c_float(m),
                    C.c_float(sigma), C.c_float(mix), _ptr(w), _ptr(pot))
    return w, pot


def mstep_sums(g: Geometry, slices, weights, simslices, simweights, mstep_scales):
    out = np.zeros(5, np.float64)
    lib().orc_mstep_sums(g.S, g.Nx, g.Ny, _ptr(slices), _ptr(weights), _ptr(simslices), _ptr(simweights),
                         _ptr(mstep_scales), _ptr(out))
    return out


def mstep_finish(sums5, it, step, sigma, mix, m):
    s, mi, mm = C.c_float(sigma), C.c_float(mix), C.c_float(m)
    sums5 = np.ascontiguousarray(sums5, np.float64)
    lib().orc_mstep_finish(_ptr(sums5), C.c_int(it), C.c_float(step), C.byref(s), C.byref(mi), C.byref(mm))
    return s.value, mi.value, mm.value


def calculate_scale_vector(g: Geometry, slices, weights, simslices, simweights):
    sc = np.zeros(g.S, np.float32)
    lib().orc_calculate_scale_vector(g.S, g.Nx, g.Ny, _ptr(slices), _ptr(weights), _ptr(simslices), _ptr(simweights),
                                     _ptr(sc))
    return sc


def mask_volume(recon, mask):
    lib().orc_mask_volume(C.c_size_t(recon.size), _ptr(recon), _ptr(mask))


def scale_volume(g: Geometry, slices, weights, simslices, simweights, slice_weights, recon):
    return float(lib().orc_scale_volume(g.S, g.Nx, g.Ny, C.c_size_t(recon.size), _ptr(slices), _ptr(weights),
                                        _ptr(simslices), _ptr(simweights), _ptr(slice_weights), _ptr(recon)))


def host_slice_em(slice_potential, scale, slice_weight, force_excluded, small_slices, step, state5):
    fe = np.ascontiguousarray(force_excluded, np.int32)
    sm = np.ascontiguousarray(small_slices, np.int32)
    lib().orc_host_slice_em(C.c_int(slice_potential.size), _ptr(slice_potential), _ptr(scale), _ptr(slice_weight),
                            _ptr(fe), C.c_int(fe.size), _ptr(sm), C.c_int(sm.size), C.c_double(step), _ptr(state5))


def host_small_slices(voxel_num):
    """irtkReconstructionGPU.cc:2714-2726 on per-slice counts (deviation D4): slices with fewer than 10 % of the median count."""
    vn = np.asarray(voxel_num, np.int32)
    if vn.size == 0:
        return np.zeros(0, np.int32)
    tmp = np.sort(vn)
    mid = min(int(np.floor(tmp.size * 0.5 + 0.5)), tmp.size - 1)
    return np.nonzero(vn < 0.1 * tmp[mid])[0].astype(np.int32)


# ---- registration (oracle/reg_oracle.c) -------------------------------------------------------------
def reg_gauss_kernel(sigma):
    half = np.zeros(32, np.float32)
    lib().reg_gauss_kernel.restype = C.c_int
    k = lib().reg_gauss_kernel(C.c_float(sigma), _ptr(half))
    return int(k), half[:(k + 1) // 2].copy()


def reg_filter_gauss_stack(stack, sigma):
    """stack [n, H, W] float32, filtered in place (FilterGaussStack)."""
    n, H, W = stack.shape
    lib().reg_filter_gauss_stack(_ptr(stack), C.c_int(W), C.c_int(H), C.c_int(n), C.c_float(sigma))
    return stack


def reg_tex3d(vol, pos):
    vz, vy, vx = vol.shape
    lib().reg_tex3d.restype = C.c_float
    return float(lib().reg_tex3d(_ptr(vol), vx, vy, vz, C.c_float(pos[0]), C.c_float(pos[1]), C.c_float(pos[2])))


def reg_evaluate(resampled, ofs, vol, voxel, recon_w2i, transforms, level):
    S, H, W = resampled.shape
    vz, vy, vx = vol.shape
    sim = np.zeros(max(S, 1), np.float32)
    ofs = np.ascontiguousarray(ofs, np.float32); tr = np.ascontiguousarray(transforms, np.float32)
    rw = np.ascontiguousarray(recon_w2i, np.float32)
    lib().reg_evaluate(W, H, S, _ptr(resampled), _ptr(ofs), _ptr(vol), vx, vy, vz, C.c_float(voxel), _ptr(rw), _ptr(tr),
                       C.c_int(level), _ptr(sim))
    return sim[:S]


def reg_register_slices(resampled, ofs, vol, voxel, recon_w2i, transforms, n_levels=2, n_steps=4, n_iterations=20):
    S, H, W = resampled.shape
    vz, vy, vx = vol.shape
    ofs = np.ascontiguousarray(ofs, np.float32)
    tr = np.array(transforms, np.float32, copy=True)
    rw = np.ascontiguousarray(recon_w2i, np.float32)
    lib().reg_register_slices.restype = C.c_longlong
    ev = lib().reg_register_slices(W, H, S, _ptr(resampled), _ptr(ofs), _ptr(vol), vx, vy, vz, C.c_float(voxel), _ptr(rw),
                                   _ptr(tr), C.c_int(n_levels), C.c_int(n_steps), C.c_int(n_iterations))
    return tr, int(ev)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))
