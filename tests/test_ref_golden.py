"""Golden vectors written by the REFERENCE ITSELF: tests/golden/ref_{svr,reg,steps}_small.npz are outputs of the
unmodified reference CUDA path (source/reconstructionGPU2/reconstruction_cuda2.cu, GPUWorker.cpp, GPUGauss/gaussfilter.cu),
compiled for sm_100a by oracle/Makefile (`make ref`) and run on a B200 by oracle/ref_runner.py on the seeded cases of
tests/golden/make_golden.py / oracle/ref_runner.steps_case.  They pin

  * (CPU, not gpu)  the oracle restatement (oracle/*.c) against the reference, stage by stage;
  * (gpu)           the CUDA path (libsvr_b200.so through the C ABI) against the reference.

Reference-compared cases use slice sizes and volume sizes that are multiples of 8: the reference kernels have no x/y
bounds checks (make_golden.REF_SVR_SIZE / REF_SVR_VOL), we do (deviation D1).

Tolerances (relative to the RMS of the reference's non-zeros; measured values in profiles/r01_ref_parity.json):
  FIELD   volume- and slice-sized float fields: RMS <= 3e-4 (measured <= 8e-5 oracle, <= 2.9e-4 CUDA), max <= 6e-2.
          The max bound is loose because one flipped epsilon-skip decision (|old - psf| vs 1e-5, a 1-ulp effect of the
          reference's own float mat-vec) moves one pixel's PSF mass by up to a few per cent.
  SCALAR  per-slice / global statistics: relative 5e-4 (measured <= 2.3e-4).
  exact   integer outputs (voxel_num, slice_inside, siminside, slice weights in {0,1}).
Registration: similarities within 2e-3 for the oracle (software model of the texture filter, deviation D6) and 2e-5
for the CUDA path (same texture unit as the reference).
"""
import os

import numpy as np
import pytest

from conftest import rel_stats
from test_golden import mg

HERE = os.path.dirname(os.path.abspath(__file__))
FIELD = (3e-4, 6e-2)
SCALAR = (5e-4, 5e-4)


def _ref(name):
    return dict(np.load(os.path.join(HERE, "golden", f"ref_{name}_small.npz")))


# flipped epsilon-skip decisions are bounded by COUNT: at most FLIP_FRAC of a field's elements (never fewer than 2 allowed)
# may be further than FLIP_THR from the reference; the max bound of FIELD is what one flipped tap can move.
FLIP_THR, FLIP_FRAC = 3e-3, 1.5e-3       # measured: at most 33 of 64 000 (the K3 accumulator of the kernel-by-kernel case)


def _check(got, ref, names, tol):
    for n in names:
        a, b = np.asarray(got[n], np.float64), np.asarray(ref[n], np.float64)
        m, r = rel_stats(a, b)
        assert r <= tol[0] and m <= tol[1], f"{n}: rms/max = {(r, m)} > {tol}"
        if tol is FIELD and a.size >= 1000:
            nz = b[b != 0]
            sc = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
            off = int(np.count_nonzero(np.abs(a - b) / sc > FLIP_THR))
            assert off <= max(2, int(FLIP_FRAC * a.size)), f"{n}: {off} of {a.size} elements beyond {FLIP_THR}"


def _steps(backend):
    from oracle.ref_runner import steps_case
    return steps_case(backend)


STEPS_EXACT = ["voxel_num", "inside", "simi"]
STEPS_FIELDS = ["recon0", "volw", "psf", "sim", "simw", "weights", "addon", "cmap", "recon1", "sim2", "recon_masked"]
STEPS_SCALARS = ["sigma0", "potential", "scale1", "mstep", "scale2"]
SVR_FIELDS = ["gaussian_recon", "volweights", "simslices", "volume"]
SVR_SCALARS = ["slice_potential", "scale", "em"]


def _check_steps(got, ref):
    for n in STEPS_EXACT:
        assert np.array_equal(np.asarray(got[n]).astype(np.int64), np.asarray(ref[n]).astype(np.int64)), n
    _check(got, ref, STEPS_FIELDS, FIELD)
    _check(got, ref, STEPS_SCALARS, SCALAR)


def _check_svr(got, ref):
    assert np.array_equal(got["voxel_num"], ref["voxel_num"])
    assert np.array_equal(got["siminside"], ref["siminside"])
    assert np.array_equal(got["slice_weight"], ref["slice_weight"])
    _check(got, ref, SVR_FIELDS, FIELD)
    _check(got, ref, SVR_SCALARS, SCALAR)
    assert float(got["sigma"]) == pytest.approx(float(ref["sigma"]), rel=5e-4)
    m, r = rel_stats(got["psf_sums"].astype(np.float64), ref["psf_sums"].astype(np.float64))    # stored as float16
    assert r <= 1e-3 and m <= 6e-2, (r, m)


# ---- the oracle against the reference (CPU) ----------------------------------------------------------------------
def test_oracle_matches_reference_kernel_by_kernel():
    from oracle.oracle_backend import OracleReconstruction
    _check_steps(_steps(OracleReconstruction()), _ref("steps"))


def test_oracle_matches_reference_svr_loop():
    _check_svr(mg.svr_case(None, slice_size=mg.REF_SVR_SIZE), _ref("svr"))


def test_oracle_matches_reference_registration():
    got, ref = mg.reg_case(), _ref("reg")
    # the cube the reference was run on came from round 1's resampling (pre-multiplied matrices); the reference-order resampling of
    # round 2 moves a handful of its float32 voxels by an ulp (4e-11 of the sum)
    assert float(got["resampled_checksum"]) == pytest.approx(float(ref["resampled_checksum"]), rel=1e-9)
    for k in ("sim_level0", "sim_level1"):
        assert np.abs(got[k] - ref[k]).max() <= 2e-3, (k, np.abs(got[k] - ref[k]).max())
    # the optimiser walks a similarity staircase whose steps depend on the texture filter's arithmetic: the oracle
    # (software filter) ends near, not on, the reference's transforms; the CUDA path is held to a tight bound below
    from fetalreconstruction_b200.geometry import rigid_parameters
    pg = np.stack([rigid_parameters(m.reshape(4, 4).astype(np.float64)) for m in got["transforms_out"]])
    pr = np.stack([rigid_parameters(m.reshape(4, 4).astype(np.float64)) for m in ref["transforms_out"]])
    p0 = np.stack([rigid_parameters(m.reshape(4, 4).astype(np.float64)) for m in ref["transforms_in"]])
    moved = np.abs(pr - p0).max(1)
    assert np.median(np.abs(pg - pr).max(1)) <= 0.5 * np.median(moved), (np.abs(pg - pr).max(1), moved)


# ---- the CUDA path against the reference (GPU) -------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_matches_reference_kernel_by_kernel():
    from fetalreconstruction_b200.reconstruction import Reconstruction
    _check_steps(_steps(Reconstruction(0)), _ref("steps"))


@pytest.mark.gpu
def test_cuda_matches_reference_svr_loop():
    from fetalreconstruction_b200.reconstruction import Reconstruction
    _check_svr(mg.svr_case(Reconstruction(0), slice_size=mg.REF_SVR_SIZE), _ref("svr"))


@pytest.mark.gpu
def test_cuda_matches_reference_registration():
    """Same texture unit, same similarity: <= 2e-5 (measured 1.5e-6).  The optimiser is a chain of discrete decisions
    (line-search length, `similarity > previous + 1e-4`) on similarities that agree to ~1e-6 but not bitwise (the
    reference sums in float with a tree + atomics, we in double): after ONE iteration the parameters agree to 1e-3
    mm/degrees (measured 1e-4), then the trajectories separate (tools/ref_reg_trace.py: 7e-4 after 2 iterations, 3e-2
    after 4).  The full schedule is therefore held to the quality of its optimum, not to its path."""
    from fetalreconstruction_b200.geometry import rigid_parameters
    from fetalreconstruction_b200.reconstruction import Reconstruction
    b = Reconstruction(0)
    got, ref = mg.reg_case(b), _ref("reg")
    for k in ("sim_level0", "sim_level1"):
        assert np.abs(got[k] - ref[k]).max() <= 2e-5, (k, np.abs(got[k] - ref[k]).max())
    s_ours = b.evaluateCostsMultipleSlices(got["transforms_out"], 0)
    s_ref = b.evaluateCostsMultipleSlices(ref["transforms_out"], 0)
    s_in = b.evaluateCostsMultipleSlices(ref["transforms_in"], 0)
    assert s_ref.mean() > s_in.mean()
    assert s_ours.mean() >= s_ref.mean() - 0.01, (s_ours, s_ref)
    # one optimiser iteration (1 level, 1 step size, 1 iteration) against the reference's
    trace = dict(np.load(os.path.join(HERE, "golden", "ref_regtrace.npz")))
    b2 = Reconstruction(0)
    orig = b2.setRegSchedule
    b2.setRegSchedule = lambda *a: orig(1, 1, 1)
    one = mg.reg_case(b2)["transforms_out"]
    pg = np.stack([rigid_parameters(m.reshape(4, 4).astype(np.float64)) for m in one])
    pr = np.stack([rigid_parameters(m.reshape(4, 4).astype(np.float64)) for m in trace["T_111"]])
    assert np.abs(pg - pr).max() <= 1e-3, np.abs(pg - pr).max()


# ---- PVR: the reference's patch-based CUDA path (oracle/_ref/libref_pvr.so, oracle/ref_runner_pvr.py) -------------
# tests/golden/ref_pvr_small.npz = every stage of one PVR iteration (PSF reconstruction, equalize, simulate, robust
# statistics init, E-step, 2 x {scale, super-resolution, regularise, simulate, M-step, E-step}) written by the UNMODIFIED
# patchBased*_gpu.cu / reconVolume.cu of the reference on a B200, on the patch list and patch values of our enumeration.
PVR_FIELDS = ["p1_recon_raw", "p1_volw", "p1_psf_sums", "p1_recon", "p2_sim", "p2_simw", "e0_weights", "r0_recon", "r0_sim",
              "r0_weights", "r1_recon", "r1_sim", "r1_weights", "volume"]
PVR_SCALARS = ["rs_init", "e0_patch_scale", "e0_patch_weight", "e0_state", "r0_patch_scale", "r0_mstep", "r0_patch_weight",
               "r1_patch_scale", "r1_mstep", "r1_patch_weight"]
PVR_MASKED = ["r0_addon", "r0_cmap", "r1_addon", "r1_cmap"]     # compared inside the mask: the CUDA path applies the per-tap
                                                                # mask test once per voxel afterwards


def _check_pvr(got, ref):
    from oracle.ref_runner_pvr import REF_PVR_CASE
    from pvr_case import make_pvr_case
    assert np.array_equal(got["per_stack"], ref["per_stack"])
    assert np.array_equal(got["patches"], ref["patches"])
    assert np.array_equal(np.asarray(got["p2_inside"]).astype(np.int8), np.asarray(ref["p2_inside"]).astype(np.int8))
    _check(got, ref, PVR_FIELDS, FIELD)
    _check(got, ref, PVR_SCALARS, SCALAR)
    inside = make_pvr_case(**REF_PVR_CASE)["mask"].ravel() != 0
    for k in PVR_MASKED:
        m, r = rel_stats(np.asarray(got[k], np.float64)[inside], np.asarray(ref[k], np.float64)[inside])
        assert r <= FIELD[0] and m <= FIELD[1], f"{k}: rms/max = {(r, m)}"


def test_oracle_matches_reference_pvr():
    from fetalreconstruction_b200.pvr import PVRPipeline
    from oracle.oracle_backend_pvr import OraclePatchReconstruction
    from oracle.ref_runner_pvr import pvr_stages
    _check_pvr(pvr_stages(OraclePatchReconstruction(), PVRPipeline), _ref("pvr"))


@pytest.mark.gpu
def test_cuda_matches_reference_pvr():
    from fetalreconstruction_b200.pvr import PatchReconstruction, PVRPipeline
    from oracle.ref_runner_pvr import pvr_stages
    _check_pvr(pvr_stages(PatchReconstruction(0), PVRPipeline), _ref("pvr"))
