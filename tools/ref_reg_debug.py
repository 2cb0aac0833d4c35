#!/usr/bin/env python
"""Dumps the reference's registration intermediates (resampled targets, blurred targets, last sampled+blurred
volume slices) for the golden registration case, so oracle/reg_oracle.c can be diffed stage by stage.
Test tooling; own process on a GPU box:  python tools/ref_reg_debug.py gpurun_out/ref/ref_reg_debug.npz"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_runner  # noqa: E402
from oracle.ref_backend import RefReconstruction, _fp  # noqa: E402


def main():
    mg = ref_runner._mg()
    b = RefReconstruction(0)
    out = {}
    orig_eval = b.evaluateCostsMultipleSlices

    def dump(tag):
        n = (b.regS + 1) * b.regH * b.regW
        for kind, name in ((0, "reg"), (1, "blurred"), (2, "resampled")):
            buf = np.zeros(n, np.float32)
            b.lib.ref_reg_get(b.h, C.c_int(kind), _fp(buf))
            out[f"{name}_{tag}"] = buf.reshape(b.regS + 1, b.regH, b.regW)

    def ev(t, level=0):
        r = orig_eval(t, level)
        dump(f"l{level}")
        return r

    b.evaluateCostsMultipleSlices = ev
    d = mg.reg_case(b)
    out.update({k: v for k, v in d.items()})
    np.savez_compressed(sys.argv[1], **out)
    print("wrote", sys.argv[1], {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
