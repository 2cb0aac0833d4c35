"""CPU: the registration oracle (oracle/reg_oracle.c) against analytic known answers, and the host
front-end (fetalreconstruction_b200/registration.py) against properties of irtkResamplingWithPadding.

The reference holds no golden vectors for this path (its own outputs are pinned in tests/test_ref_golden.py); these
tests pin the restatement against known answers.
"""
import numpy as np
import pytest

from fetalreconstruction_b200.geometry import ImageAttributes, rigid_matrix
from fetalreconstruction_b200.phantom import make_dataset, small_config
from fetalreconstruction_b200.registration import (RegistrationFrontEnd, resample_plane0_with_padding,
                                                   resampled_attributes)
from oracle import oracle as orc
from oracle.oracle_backend import OracleReconstruction


def test_gauss_kernel_length_rule_and_normalisation():
    # klength = max(min(int(5 sigma), 63), 7) made odd (gaussfilter.cu:189-192); weights sum to 1
    for sigma, klen in ((0.375, 7), (0.75, 7), (1.0, 7), (1.9, 9), (2.0, 9), (2.4, 11), (20.0, 63)):
        k, half = orc.reg_gauss_kernel(sigma)
        assert k == klen
        assert abs(half[0] + 2 * half[1:].sum() - 1.0) < 1e-6
        assert np.all(np.diff(half) <= 0)


def test_blur_keeps_constants_and_padding():
    img = np.full((2, 12, 15), 5.0, np.float32)
    img[0, 3, 4] = -1.0                       # a padded centre passes through, and counts as 0 for its neighbours
    out = orc.reg_filter_gauss_stack(img.copy(), 0.75)
    assert out[0, 3, 4] == -1.0
    assert np.allclose(out[1], 5.0, atol=1e-5)          # clamped reads keep a constant image constant
    assert out[0, 3, 5] < 5.0 and out[0, 3, 5] > 0.0
    assert np.allclose(out[0, 9:, 10:], 5.0, atol=1e-5)


def test_texture_read_has_half_voxel_shift_and_zero_border():
    rng = np.random.default_rng(0)
    vol = rng.uniform(1, 2, (6, 7, 8)).astype(np.float32)
    # quirk G4: position i + 0.5 reads voxel i exactly; position i reads the mean of voxels i-1 and i
    assert orc.reg_tex3d(vol, (3.5, 2.5, 4.5)) == pytest.approx(vol[4, 2, 3], rel=1e-6)
    assert orc.reg_tex3d(vol, (3.0, 2.5, 4.5)) == pytest.approx(0.5 * (vol[4, 2, 2] + vol[4, 2, 3]), rel=1e-6)
    # border mode: texels outside the array are 0
    assert orc.reg_tex3d(vol, (0.0, 2.5, 4.5)) == pytest.approx(0.5 * vol[4, 2, 0], rel=1e-6)
    assert orc.reg_tex3d(vol, (-3.0, 2.5, 4.5)) == 0.0
    assert orc.reg_tex3d(vol, (1e9, 2.5, 4.5)) == 0.0


def _plane_case(S=1, n=24, W=20, H=18, seed=3):
    """A volume that is constant along z and slices in the z = const plane that were cut from it."""
    rng = np.random.default_rng(seed)
    vattr = ImageAttributes(n, n, n, 1.0, 1.0, 1.0)
    plane = rng.uniform(100, 900, (n, n)).astype(np.float32)
    from scipy.ndimage import gaussian_filter
    plane = gaussian_filter(plane, 2.0).astype(np.float32)
    vol = np.repeat(plane[None], n, 0).copy()
    sattr = ImageAttributes(W, H, 1, 1.0, 1.0, 1.0)
    ofs = np.tile(sattr.image_to_world().astype(np.float32).ravel(), (S, 1))
    T = np.tile(np.eye(4, dtype=np.float32).ravel(), (S, 1))
    rw2i = vattr.world_to_image().astype(np.float32).ravel()
    return vol, ofs, T, rw2i, sattr, vattr


def test_similarity_of_a_slice_cut_from_the_volume():
    vol, ofs, T, rw2i, sattr, vattr = _plane_case(S=1)
    # the slice is what the (shifted) texture read returns -> NCC = 1 per in-slice offset
    from oracle.oracle import lib, _ptr
    import ctypes as C
    H, W = sattr.y, sattr.x
    sl = np.zeros((1, H, W), np.float32)
    lib().reg_generate_slice(_ptr(vol), 24, 24, 24, _ptr(rw2i), _ptr(T[0].copy()), _ptr(ofs[0].copy()), W, H, 0, _ptr(sl))
    sim = orc.reg_evaluate(sl, ofs, vol, 1.0, rw2i, T, 0)
    # literal buffer choreography (quirk G1) with a == slices == 1: the NCC accumulator and the first two
    # triplet entries are cleared before every offset, the third (sum of squares of the sampled slice)
    # is not -> 1 / sqrt(3) instead of the intended 3.0
    assert sim[0] == pytest.approx(1.0 / np.sqrt(3.0), abs=2e-3)


def test_similarity_choreography_depends_on_active_position():
    # with 3 identical slices all active (a = S = 3) the cleared window [2S,5S) = [6,15) covers the NCC
    # accumulators [6,9) and the triplets of t = 0,1 ([9,15)) but not t = 2 ([15,18))
    vol, ofs, T, rw2i, sattr, vattr = _plane_case(S=3)
    from oracle.oracle import lib, _ptr
    H, W = sattr.y, sattr.x
    one = np.zeros((1, H, W), np.float32)
    lib().reg_generate_slice(_ptr(vol), 24, 24, 24, _ptr(rw2i), _ptr(T[0].copy()), _ptr(ofs[0].copy()), W, H, 0, _ptr(one))
    sl = np.repeat(one, 3, 0).copy()
    sim = orc.reg_evaluate(sl, ofs, vol, 1.0, rw2i, T, 0)
    assert sim[0] == pytest.approx(1.0, abs=2e-3) and sim[1] == pytest.approx(1.0, abs=2e-3)
    assert sim[2] == pytest.approx(1.0, abs=2e-3)      # pooled triplets of three identical offsets: still 1
    # a mismatching slice scores lower
    bad = sl.copy()
    bad[1] = np.random.default_rng(1).uniform(100, 900, one.shape[1:]).astype(np.float32)
    sim2 = orc.reg_evaluate(bad, ofs, vol, 1.0, rw2i, T, 0)
    assert sim2[1] < 0.5 and sim2[0] == pytest.approx(sim[0], abs=1e-6)


def test_parameter_kernels_round_trip():
    import ctypes as C
    from oracle.oracle import lib, _ptr
    m = rigid_matrix(1.0, -2.0, 0.5, 10.0, -5.0, 20.0).astype(np.float32).ravel()
    out = np.zeros(16, np.float32)
    for part in range(6):
        lib().reg_adjust_matrix(_ptr(m), _ptr(out), C.c_int(part), C.c_float(0.5))
        p0 = np.array([1.0, -2.0, 0.5, 10.0, -5.0, 20.0])
        p0[part] += 0.5
        assert np.allclose(out.reshape(4, 4)[:3], rigid_matrix(*p0)[:3], atol=2e-6)
    g = np.array([0.1, 0.2, -0.3, 0.4, -0.5, 0.6], np.float32)
    mm = m.copy()
    lib().reg_gradient_step(_ptr(mm), _ptr(g), C.c_float(2.0))
    p1 = np.array([1.0, -2.0, 0.5, 10.0, -5.0, 20.0]) + 2.0 * g
    assert np.allclose(mm.reshape(4, 4)[:3], rigid_matrix(*p1)[:3], atol=2e-6)


def test_resampling_with_padding_matches_the_reference_irtk():
    """registration.resampled_attributes / resample_plane0_with_padding against irtkResamplingWithPadding<irtkRealPixel> of the
    reference's own IRTK (oracle/_ref/libref_irtk.so): grid (rounded size), geometry and plane 0 of the values."""
    from oracle import ref_irtk as ri
    if not ri.available():
        pytest.skip("oracle/_ref/libref_irtk.so not built")
    rng = np.random.default_rng(4)
    for (nx, ny, dx, dz, d) in ((10, 8, 1.2, 2.5, 1.0), (33, 27, 1.1765, 2.5, 1.0), (16, 16, 0.75, 3.0, 0.75), (21, 19, 1.3, 4.0, 0.9)):
        a = ImageAttributes(nx, ny, 1, dx, dx, dz, rng.uniform(-20, 20, 3))
        img = rng.uniform(1, 900, (ny, nx))
        img[:, : nx // 3] = -1.0
        img[ny // 2, nx // 2] = -1.0
        ra = resampled_attributes(a, d)
        ref = ri.Image.new(np.concatenate([[nx, ny, 1, dx, dx, dz], a.origin, a.xaxis, a.yaxis, a.zaxis]), img[None]).resample_with_padding(d, d, d, -1)
        got_attr = np.concatenate([[ra.x, ra.y, ra.z, ra.dx, ra.dy, ra.dz], ra.origin, ra.xaxis, ra.yaxis, ra.zaxis])
        np.testing.assert_allclose(got_attr, ref.attrs, rtol=0, atol=1e-12)
        out = resample_plane0_with_padding(img, a, ra)
        np.testing.assert_array_equal(out == -1, ref.data[0] == -1)
        np.testing.assert_allclose(out, ref.data[0], rtol=2e-7, atol=0)      # the restatement returns float32 (what the device is fed)


def test_resampling_with_padding_rules():
    a = ImageAttributes(10, 8, 1, 1.2, 1.2, 2.5)
    ra = resampled_attributes(a, 1.0)
    assert (ra.x, ra.y, ra.z) == (12, 10, 3)              # round(n * old / new): irtkResamplingWithPadding.cc:217-219
    img = np.full((8, 10), 7.0, np.float32)
    out = resample_plane0_with_padding(img, a, ra)
    assert out.shape == (10, 12)
    inner = out[1:-1, 1:-1]
    assert np.allclose(inner[inner != -1], 7.0)           # renormalised weights keep a constant
    img[:, :5] = -1.0
    out = resample_plane0_with_padding(img, a, ra)
    assert (out[:, :4] == -1).all() and np.allclose(out[2:-2, 8:-1], 7.0)
    # same grid, same voxel size: identity
    b = ImageAttributes(9, 7, 1, 1.0, 1.0, 1.0)
    rb = resampled_attributes(b, 1.0)
    rng = np.random.default_rng(2)
    im = rng.uniform(1, 2, (7, 9)).astype(np.float32)
    assert np.allclose(resample_plane0_with_padding(im, b, rb), im, atol=1e-6)


def _reg_dataset():
    cfg = small_config(seed=5, vol=36, n_stacks=2, slices=6, size=30, inplane=1.0, spacing=2.0)
    cfg.noise = 2.0
    cfg.corrupt_fraction = 0.0
    return make_dataset(cfg)


def test_front_end_and_registration_improve_similarity():
    ds = _reg_dataset()
    b = OracleReconstruction()
    vx, vy, vz = ds.cfg.vol_size
    vol = np.where(ds.mask > 0, ds.truth, -1.0).astype(np.float32)      # what MaskVolume leaves behind
    b.InitReconstructionVolume((vx, vy, vz), (ds.cfg.vol_voxel,) * 3, vol.ravel())
    b.recon_w2i = ds.recon_w2i
    fe = RegistrationFrontEnd(b, ds.slices, ds.slice_attrs, ds.cfg.vol_voxel)
    assert fe.cube.shape[0] == ds.S
    # perturb the true transforms by a known offset and register
    rng = np.random.default_rng(0)
    pert = np.stack([(ds.true_trans[k].reshape(4, 4).astype(np.float64) @ rigid_matrix(*(rng.normal(0, 0.6, 3)), *(rng.normal(0, 0.6, 3)))).ravel()
                     for k in range(ds.S)])
    b.updateResampledSlicesI2W(fe.ofs)
    b.prepareSliceToVolumeReg()
    b.setRegSchedule(2, 2, 6)
    s0 = b.evaluateCostsMultipleSlices(fe.pack_transforms(pert), 0)
    out = b.registerSlicesToVolume(fe.pack_transforms(pert))
    s1 = b.evaluateCostsMultipleSlices(out, 0)
    assert b.reg_evaluations > 0
    valid = s0 != 0
    # (the literal similarity depends on the position in the active list, quirk G1, so individual slices may
    # score lower afterwards even though the optimiser only accepted improving steps)
    assert np.mean(s1[valid] >= s0[valid] - 1e-4) >= 0.5
    assert np.mean(s1[valid]) > np.mean(s0[valid])
    # pack / unpack are inverse
    assert np.allclose(fe.unpack_transforms(fe.pack_transforms(pert)), pert, atol=1e-4)
