"""Host binding of the IRTK-style rigid registration engine (csrc/svr_rreg.cu, include/svr_abi.h: svr_rreg_*): the reference's
default stack-to-template, slice-to-volume and patch-to-volume registrations (irtkImageRigidRegistrationWithPadding on the CPU,
irtkReconstructionGPU.cc:849-1001,1992-2059, patchBased2D3DRegistration.cpp:88-168), batched on the device.

Images are the reference's irtkGreyImage: int16 voxels [z][y][x] + 18 attribute numbers
{x, y, z, dx, dy, dz, origin, xaxis, yaxis, zaxis}.  `to_grey` is the reference's irtkRealImage -> irtkGreyImage conversion
(static_cast<short>: truncation towards zero, image++/src/irtkGenericImage.cc:699-713)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .reconstruction import SVRError

STACK, SLICE_TO_VOLUME = 0, 1


def to_grey(real):
    return np.ascontiguousarray(np.trunc(np.asarray(real, np.float64)).astype(np.int16))


def attrs18(a):
    """geometry.ImageAttributes -> the 18 numbers of irtkImageAttributes."""
    return np.concatenate([[a.x, a.y, a.z, a.dx, a.dy, a.dz], np.asarray(a.origin, float), np.asarray(a.xaxis, float),
                           np.asarray(a.yaxis, float), np.asarray(a.zaxis, float)]).astype(np.float64)


def _ctx(backend):
    return backend._h


def register(backend, images, attrs, target_of_item, source_of_item, kind, dofs, level_only=-1, want_prepared=False):
    """images: list of int16 arrays; attrs: [n_images, 18]; dofs: [n_items, 6] start parameters.
    Returns (dofs, similarity, evaluations) or, with level_only >= 0, (similarity, prepared target, its attrs, prepared source, attrs)."""
    lib = _lib.load()
    imgs = [np.ascontiguousarray(i, np.int16) for i in images]
    n_images, n_items = len(imgs), len(target_of_item)
    ptrs = (C.c_void_p * n_images)(*[i.ctypes.data for i in imgs])
    at = np.ascontiguousarray(attrs, np.float64).reshape(n_images, 18)
    ti = np.ascontiguousarray(target_of_item, np.int32)
    si = np.ascontiguousarray(source_of_item, np.int32)
    d = np.ascontiguousarray(dofs, np.float64).reshape(n_items, 6).copy()
    sim = np.zeros(max(n_items, 1), np.float64)
    ev = C.c_int64(0)
    pt = ps = pta = psa = None
    cap = 0
    if level_only >= 0 and want_prepared:
        cap = 4 * max(imgs[ti[0]].size, imgs[si[0]].size) + 1024      # resampling can enlarge an image (finer target resolution)
        pt = np.zeros(cap, np.int16); ps = np.zeros(cap, np.int16)
        pta = np.zeros(18); psa = np.zeros(18)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.svr_rreg_register(_ctx(backend), n_items, n_images, ptrs, p(at), p(ti), p(si), int(kind), p(d), p(sim), C.byref(ev), int(level_only),
                               cap, p(pt), p(pta), p(ps), p(psa))
    if rc != 0:
        raise SVRError(lib.svr_last_error(_ctx(backend)).decode())
    if level_only >= 0:
        if want_prepared:
            shp = lambda a18: (int(a18[2]), int(a18[1]), int(a18[0]))
            return sim[:n_items], pt[:int(np.prod(shp(pta)))].reshape(shp(pta)), pta, ps[:int(np.prod(shp(psa)))].reshape(shp(psa)), psa
        return sim[:n_items]
    return d, sim[:n_items], int(ev.value)


def blur_with_padding(backend, image, attr18, sigma, padding):
    lib = _lib.load()
    im = np.ascontiguousarray(image, np.int16)
    a = np.ascontiguousarray(attr18, np.float64)
    out = np.zeros_like(im)
    if lib.svr_rreg_blur_with_padding(_ctx(backend), im.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), float(sigma), int(padding),
                                      out.ctypes.data_as(C.c_void_p)) != 0:
        raise SVRError(lib.svr_last_error(_ctx(backend)).decode())
    return out


def resample_with_padding(backend, image, attr18, dx, dy, dz, padding):
    lib = _lib.load()
    im = np.ascontiguousarray(image, np.int16)
    a = np.ascontiguousarray(attr18, np.float64)
    cap = int(im.size * max(1.0, a[3] / dx) * max(1.0, a[4] / dy) * max(1.0, a[5] / dz)) + 16
    out = np.zeros(cap, np.int16)
    oa = np.zeros(18)
    if lib.svr_rreg_resample_with_padding(_ctx(backend), im.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_void_p), float(dx), float(dy), float(dz),
                                          int(padding), out.ctypes.data_as(C.c_void_p), cap, oa.ctypes.data_as(C.c_void_p)) != 0:
        raise SVRError(lib.svr_last_error(_ctx(backend)).decode())
    shape = (int(oa[2]), int(oa[1]), int(oa[0]))
    return out[:int(np.prod(shape))].reshape(shape), oa
