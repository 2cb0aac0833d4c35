"""CPU: the PVR oracle (oracle/pvr_oracle.c) against analytic known answers, the patch enumeration, the pure-host
patch-level EM of the C ABI against the oracle, and the PVR pipeline on the oracle twin.
The reference holds no golden vectors for this path; the vectors its own PVR CUDA code produces are pinned in
tests/test_ref_golden.py."""
import ctypes as C

import numpy as np
import pytest

from fetalreconstruction_b200.pvr import PVRParams, PVRPipeline, host_patch_em
from oracle import oracle as orc
from oracle.oracle import _ptr
from oracle.oracle_backend_pvr import OraclePatchReconstruction, host_patch_em as orc_host_patch_em
from pvr_case import make_pvr_case, setup_backend


def _psf(p, d):
    f = orc.lib().pvr_psf_value
    f.restype = C.c_float
    return float(f(*[C.c_float(v) for v in p], *[C.c_float(v) for v in d]))


def test_pvr_psf_constants():
    d = (1.2, 1.2, 2.5)
    assert _psf((0, 0, 0), d) == pytest.approx(1.0)               # sinc_pi Taylor branch: no NaN at 0
    # sigma_z = dim.z (no 2.3548): half maximum at z = sqrt(2 ln 2) * dz
    z = np.sqrt(2 * np.log(2)) * d[2]
    assert _psf((0, 0, z), d) == pytest.approx(0.5, rel=1e-5)
    # first in-plane zero where (x * dx / 2.3548) = 1
    assert _psf((2.3548 / d[0], 0, 0), d) < 1e-10
    # differs from the SVR PSF (sigma_z = dz / 2.3548)
    assert abs(_psf((0, 0, 1.0), d) - orc.psf_value((0, 0, 1.0), d)) > 0.05


def test_pvr_texture_read_is_eight_voxel_mean():
    rng = np.random.default_rng(0)
    vol = rng.uniform(1, 2, (5, 6, 7)).astype(np.float32)
    tex = np.zeros_like(vol)
    orc.lib().pvr_texture_volume(7, 6, 5, _ptr(vol), _ptr(tex))
    assert tex[2, 3, 4] == pytest.approx(vol[1:3, 2:4, 3:5].mean(), rel=1e-6)
    assert tex[0, 0, 0] == pytest.approx(vol[0, 0, 0] / 8, rel=1e-6)        # border texels read 0


def test_pvr_init_em_zeroes_padding_and_zero_pixels():
    p = np.array([-1, 0, 5, 0.001], np.float32)
    w = np.zeros(4, np.float32)
    orc.lib().pvr_initialize_em_values(C.c_size_t(4), _ptr(p), _ptr(w))
    assert w.tolist() == [0, 0, 1, 1]


def test_host_patch_em_matches_oracle_and_keeps_the_stack_offset_quirk(built_lib):
    rng = np.random.default_rng(1)
    per_stack = [5, 3, 4]
    n = sum(per_stack)
    pot = rng.uniform(0.05, 0.4, n).astype(np.float32)
    pot[6] = -1
    scale = rng.uniform(0.8, 1.2, n).astype(np.float32)
    scale[2] = 7.0
    w1 = rng.uniform(0.3, 1.0, n).astype(np.float32)
    w2 = w1.copy()
    s1 = np.array([0.025, 0.9, 0, 0, 0], np.float32)
    s2 = s1.copy()
    used1 = host_patch_em(per_stack, pot, scale, w1, 1e-4, s1)
    used2 = orc_host_patch_em(per_stack, pot, scale, w2, 1e-4, s2)
    assert np.array_equal(used1, used2) and np.allclose(w1, w2, atol=1e-6) and np.allclose(s1, s2, rtol=1e-6)
    # the reference indexes patch_potential[j] without the stack offset: the last stack's values overwrite the first
    # entries, entries beyond the largest stack keep their initial 0
    assert np.allclose(used1[[0, 1, 3]], pot[[8, 9, 11]]) and used1[4] == pot[4]
    assert np.all(used1[5:][scale[5:] <= 5] == 0)
    assert used1[2] == -1 and w1[2] == 0                       # unrealistic scale -> excluded


def test_patch_enumeration_rule():
    case = make_pvr_case()
    assert sum(case["per_stack"]) == len(case["attrs"]) > 4
    cube = case["cube"]
    good = (cube != 0) & (cube != -1)
    assert np.all(good.reshape(len(cube), -1).sum(1) > 16 * 16 / 3)
    # a patch pixel equals the stack pixel it was cut from
    a = case["attrs"][0]
    sattr = case["ds"].stack_attrs[0]
    p = sattr.world_to_image() @ a.image_to_world() @ np.array([3, 4, 0, 1.0])
    x, y, z = (int(round(v)) for v in p[:3])
    assert cube[0, 4, 3] in (0.0, case["stacks"][0][z, y, x])


def test_pvr_pipeline_on_oracle_reconstructs_the_phantom():
    case = make_pvr_case()
    b = setup_backend(OraclePatchReconstruction(), case, device_patch_init=True)
    # P0 on the oracle reproduces the CPU enumeration wherever the mask test agrees (truncation vs. (int) cast)
    dev = b.patches_copyToHost()
    cpu = case["cube"]
    assert np.mean(dev == cpu) > 0.97
    ds = case["ds"]
    pipe = PVRPipeline(b, ds.min_intensity, ds.max_intensity, PVRParams(iterations=0, rec_iterations=3))
    vol = pipe.run()
    m = case["mask"].ravel() > 0
    assert np.isfinite(vol).all()
    err = np.abs(vol[m] - ds.truth.ravel()[m]).mean() / ds.truth.ravel()[m].mean()
    assert err < 0.25, err
    assert 0 < pipe.sigma and 0 < pipe.mix <= 1
