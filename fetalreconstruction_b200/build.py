"""Builds libsvr_b200.so in-tree with nvcc for sm_100a (and nothing else).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  `python -m
fetalreconstruction_b200.build` or `__graft_entry__.build()` both end up here.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libsvr_b200.so")
SOURCES = ["svr_abi.cu", "svr_psf.cu", "svr_window.cu", "svr_em.cu", "svr_reg.cu", "svr_rreg.cu", "pvr_abi.cu"]
HEADERS = ["svr_common.cuh", "svr_context.h", "svr_psf_pixel.cuh", os.path.join("..", "..", "include", "svr_abi.h"),
           os.path.join("..", "..", "include", "pvr_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsvr_b200 cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    out = os.environ.get("SVR_B200_LIB_OUT", LIB)       # experiments: build variants side by side
    extra = os.environ.get("SVR_NVCC_EXTRA", "").split()
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    env = dict(os.environ)
    # the image's $CC/$CXX wrappers are not a usable nvcc host compiler; use the system g++
    env.pop("CC", None)
    env.pop("CXX", None)
    cmd[1:1] = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = res.stdout + res.stderr
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
