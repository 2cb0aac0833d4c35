"""Golden SLICO label images written by the REFERENCE's own code (oracle/_ref/libref_slic.so = its runStackSLIC.cpp compiled
where it lies behind inert IRTK stubs, `make -C oracle ref`): run in the build container (needs /root/reference),
    python tests/golden/make_golden_slic.py
writes tests/golden/ref_slic_small.npz.  tests/test_slic_ref.py pins host/pvr_slic.cc against it."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CASES = [(48, 48, 12, 0), (64, 80, 16, 1), (100, 73, 10, 2), (128, 128, 32, 3), (57, 91, 8, 4), (96, 96, 16, 5)]


def slice_image(Y, X, seed):
    """A seeded test slice: smooth texture, a bright disc with a sharp edge, a dark background corner, noise."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:Y, 0:X]
    img = 200 + 80 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 150 * ((xx - X / 2) ** 2 + (yy - Y / 2) ** 2 < (min(X, Y) / 3) ** 2)
    img = img + rng.normal(0, 8, (Y, X))
    img[: Y // 5, : X // 5] = 0.0
    return np.ascontiguousarray(img, np.float32)


def labels(lib, fn, img, spx):
    Y, X = img.shape
    out = np.zeros((Y, X), np.int32)
    f = getattr(lib, fn)
    f.restype = C.c_int
    n = f(img.ctypes.data_as(C.c_void_p), C.c_int(X), C.c_int(Y), C.c_float(float(img.min())), C.c_float(float(img.max())),
          C.c_uint(spx), C.c_uint(spx), out.ctypes.data_as(C.c_void_p))
    return n, out


if __name__ == "__main__":
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_slic.so"))
    out = {"cases": np.array(CASES, np.int32)}
    for i, (Y, X, spx, seed) in enumerate(CASES):
        n, lab = labels(ref, "refslic_labels", slice_image(Y, X, seed), spx)
        out[f"labels{i}"] = lab.astype(np.int16)
        out[f"n{i}"] = np.int32(n)
        print((Y, X, spx), "labels", n)
    np.savez_compressed(os.path.join(HERE, "ref_slic_small.npz"), **out)
