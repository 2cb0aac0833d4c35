"""ctypes binding of oracle/_ref/libref_irtk.so (TEST / BASELINE INFRASTRUCTURE ONLY): the reference's OWN host code -- the vendored
IRTK (images, NIfTI I/O, resampling / blurring with padding, rigid transformations, irtkImageRigidRegistrationWithPadding) and
class irtkReconstruction (set-up pipeline, stack / slice registration, the CPU --useCPU reconstruction path), compiled unmodified
from /root/reference by `make -C oracle ref_irtk` behind stand-ins for GSL / TBB / Boost (oracle/ref_shim/irtk/).  Pure CPU.

Used to pin (tests/) our NIfTI reader, set-up pipeline and IRTK-style registration engine against the reference itself, and as the
`kind: "reference"` CPU baseline of bench.py.  The reference prints its progress to stdout: `quiet()` silences it at the fd level.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_irtk.so")
_lib = None
vp, dp, ip = C.c_void_p, C.c_double, C.c_int


def available() -> bool:
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB} is missing: run `make -C oracle ref_irtk` where /root/reference exists")
        l = C.CDLL(LIB)
        sig = {
            "rirtk_set_threads": (None, [ip]),
            "rirtk_image_read": (vp, [C.c_char_p]), "rirtk_image_new": (vp, [vp, vp]), "rirtk_image_free": (None, [vp]),
            "rirtk_image_attrs": (None, [vp, vp]), "rirtk_image_size": (C.c_long, [vp]), "rirtk_image_data": (None, [vp, vp]),
            "rirtk_image_write": (None, [vp, C.c_char_p]), "rirtk_image_matrices": (None, [vp, vp, vp]),
            "rirtk_image_resample_with_padding": (vp, [vp, dp, dp, dp, dp]), "rirtk_image_blur_with_padding": (vp, [vp, dp, dp]),
            "rirtk_rigid_matrix": (None, [vp, vp]), "rirtk_rigid_from_matrix": (None, [vp, vp]),
            "rirtk_dof_write": (None, [vp, C.c_char_p]), "rirtk_dof_read": (ip, [C.c_char_p, vp]),
            "rirtk_rigid_register": (dp, [vp, vp, ip, vp]),
            "rirtk_reg_probe": (dp, [vp, vp, ip, ip, vp, C.POINTER(vp), C.POINTER(vp)]),
            "rirtk_create": (vp, []), "rirtk_debug": (None, [vp, ip]),
            "rirtk_add_stack": (ip, [vp, vp, vp, dp]), "rirtk_get_stack": (vp, [vp, ip]), "rirtk_get_stack_dof": (None, [vp, ip, vp]),
            "rirtk_invert_stack_transformations": (None, [vp]), "rirtk_create_template": (dp, [vp, ip, dp]),
            "rirtk_set_mask": (None, [vp, vp, dp, dp]), "rirtk_get_mask": (vp, [vp]), "rirtk_get_reconstructed": (vp, [vp]),
            "rirtk_set_reconstructed": (None, [vp, vp]), "rirtk_crop_stack_to_mask": (None, [vp, ip, vp]),
            "rirtk_stack_registrations": (None, [vp, ip]), "rirtk_match_stack_intensities_with_masking": (None, [vp, dp, ip]),
            "rirtk_create_slices_and_transformations": (None, [vp]), "rirtk_mask_slices": (None, [vp]),
            "rirtk_set_slices": (None, [vp, ip, vp, vp, vp, vp]), "rirtk_num_slices": (ip, [vp]), "rirtk_get_slice": (vp, [vp, ip]),
            "rirtk_get_transformations": (None, [vp, vp]), "rirtk_set_transformations": (None, [vp, ip, vp]),
            "rirtk_slice_to_volume_registration": (None, [vp]), "rirtk_set_smoothing_parameters": (None, [vp, dp, dp]),
            "rirtk_speedup": (None, [vp, ip]), "rirtk_cpu_step": (ip, [vp, C.c_char_p, ip]),
        }
        for name, (res, args) in sig.items():
            f = getattr(l, name)
            f.restype, f.argtypes = res, args
        _lib = l
    return _lib


@contextlib.contextmanager
def quiet(enabled=True):
    """Silences the reference's cout / printf chatter (fd 1 and 2) for the duration."""
    if not enabled:
        yield
        return
    sys.stdout.flush(); sys.stderr.flush()
    saved = os.dup(1), os.dup(2)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1); os.dup2(null, 2)
        yield
    finally:
        os.dup2(saved[0], 1); os.dup2(saved[1], 2)
        for fd in (*saved, null):
            os.close(fd)


def _p(a):
    return a.ctypes.data_as(vp)


def set_threads(n: int):
    lib().rirtk_set_threads(int(n))


class Image:
    """An irtkRealImage held by the reference library.  attrs = [x, y, z, dx, dy, dz, origin(3), xaxis(3), yaxis(3), zaxis(3)]."""

    def __init__(self, handle):
        self.h = vp(handle) if not isinstance(handle, vp) else handle

    @classmethod
    def read(cls, path):
        return cls(lib().rirtk_image_read(os.fsencode(path)))

    @classmethod
    def new(cls, attrs18, data=None):
        a = np.ascontiguousarray(attrs18, np.float64)
        d = None if data is None else np.ascontiguousarray(data, np.float64)
        return cls(lib().rirtk_image_new(_p(a), None if d is None else _p(d)))

    def __del__(self):
        try:
            if self.h:
                lib().rirtk_image_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def attrs(self):
        a = np.zeros(18)
        lib().rirtk_image_attrs(self.h, _p(a))
        return a

    @property
    def data(self):
        a = self.attrs
        d = np.zeros(int(lib().rirtk_image_size(self.h)))
        lib().rirtk_image_data(self.h, _p(d))
        return d.reshape(int(a[2]), int(a[1]), int(a[0]))

    def matrices(self):
        i2w, w2i = np.zeros(16), np.zeros(16)
        lib().rirtk_image_matrices(self.h, _p(i2w), _p(w2i))
        return i2w.reshape(4, 4), w2i.reshape(4, 4)

    def write(self, path):
        lib().rirtk_image_write(self.h, os.fsencode(path))

    def resample_with_padding(self, dx, dy, dz, padding):
        return Image(lib().rirtk_image_resample_with_padding(self.h, dx, dy, dz, padding))

    def blur_with_padding(self, sigma, padding):
        return Image(lib().rirtk_image_blur_with_padding(self.h, sigma, padding))


def rigid_matrix(dof6):
    d, m = np.ascontiguousarray(dof6, np.float64), np.zeros(16)
    lib().rirtk_rigid_matrix(_p(d), _p(m))
    return m.reshape(4, 4)


def rigid_from_matrix(m):
    mm, d = np.ascontiguousarray(m, np.float64).ravel(), np.zeros(6)
    lib().rirtk_rigid_from_matrix(_p(mm), _p(d))
    return d


def rigid_register(target: Image, source: Image, kind: int, dof6, silent=True):
    """irtkImageRigidRegistrationWithPadding as the reference configures it: kind 0 = stack registration
    (GuessParameterThickSlices, target padding 0), kind 1 = slice / patch to volume (GuessParameterSliceToVolume, padding -1)."""
    d = np.ascontiguousarray(dof6, np.float64).copy()
    with quiet(silent):
        lib().rirtk_rigid_register(target.h, source.h, int(kind), _p(d))
    return d


def reg_probe(target: Image, source: Image, kind: int, level: int, dof6, silent=True):
    """(similarity, prepared target, prepared source) of the reference's registration object at one resolution level: what
    irtkImageRegistrationWithPadding::Initialize(level) builds and irtkImageRigidRegistrationWithPadding::Evaluate() returns."""
    d = np.ascontiguousarray(dof6, np.float64)
    t, s = vp(), vp()
    with quiet(silent):
        sim = lib().rirtk_reg_probe(target.h, source.h, int(kind), int(level), _p(d), C.byref(t), C.byref(s))
    return float(sim), Image(t), Image(s)


class Reconstruction:
    """class irtkReconstruction (source/reconstructionGPU2/irtkReconstructionGPU.cc) with its GPU member stubbed out."""

    def __init__(self, silent=True):
        self.silent = silent
        with quiet(silent):
            self.h = vp(lib().rirtk_create())
        self.n_stacks = 0

    def add_stack(self, img: Image, dof6=None, thickness=0.0):
        d = np.zeros(6) if dof6 is None else np.ascontiguousarray(dof6, np.float64)
        self.n_stacks += 1
        return lib().rirtk_add_stack(self.h, img.h, _p(d), float(thickness))

    def stack(self, i):
        return Image(lib().rirtk_get_stack(self.h, i))

    def stack_dof(self, i):
        d = np.zeros(6)
        lib().rirtk_get_stack_dof(self.h, i, _p(d))
        return d

    def call(self, name, *args):
        with quiet(self.silent):
            return getattr(lib(), "rirtk_" + name)(self.h, *args)

    def create_template(self, stack, resolution):
        return self.call("create_template", int(stack), float(resolution))

    def set_mask(self, mask: Image, sigma, threshold=0.5):
        self.call("set_mask", mask.h, float(sigma), float(threshold))

    def mask(self):
        return Image(lib().rirtk_get_mask(self.h))

    def reconstructed(self):
        return Image(lib().rirtk_get_reconstructed(self.h))

    def set_reconstructed(self, img: Image):
        self.call("set_reconstructed", img.h)

    def set_slices(self, images, dofs6, stack_ids, thickness):
        n = len(images)
        arr = (vp * n)(*[im.h for im in images])
        d = np.ascontiguousarray(dofs6, np.float64)
        ids = np.ascontiguousarray(stack_ids, np.int32)
        th = np.ascontiguousarray(thickness, np.float64)
        with quiet(self.silent):
            lib().rirtk_set_slices(self.h, n, arr, _p(d), _p(ids), _p(th))

    def num_slices(self):
        return int(lib().rirtk_num_slices(self.h))

    def slice(self, i):
        return Image(lib().rirtk_get_slice(self.h, i))

    def transformations(self):
        d = np.zeros((self.num_slices(), 6))
        lib().rirtk_get_transformations(self.h, _p(d))
        return d

    def set_transformations(self, dofs6):
        d = np.ascontiguousarray(dofs6, np.float64)
        lib().rirtk_set_transformations(self.h, len(d), _p(d))

    def cpu_step(self, name, it=0):
        with quiet(self.silent):
            rc = lib().rirtk_cpu_step(self.h, name.encode(), int(it))
        if rc:
            raise ValueError(name)
