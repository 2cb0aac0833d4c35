"""GPU: the IRTK-style rigid registration engine (csrc/svr_rreg.cu) against the reference's OWN IRTK
(oracle/_ref/libref_irtk.so = the vendored IRTKSimple2 compiled unmodified, oracle/Makefile `make ref_irtk`).

The engine is integer work on short images with every floating-point operation rounded like the CPU code (no FMA contraction,
the reference's operation order), so the bar is EXACT: prepared images identical voxel for voxel, similarities equal to the last
bit, and with them the optimiser's path -- the final parameters equal the reference's."""
import os

import numpy as np
import pytest

from oracle import ref_irtk as ri

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ri.available(), reason="oracle/_ref/libref_irtk.so not built (make -C oracle ref_irtk)")]


def _gpu():
    from fetalreconstruction_b200.reconstruction import Reconstruction
    return Reconstruction(0)


def attrs(n, d, org=(0, 0, 0), axes=None):
    a = np.zeros(18)
    a[0:3] = n; a[3:6] = d; a[6:9] = org
    ax = np.eye(3) if axes is None else np.asarray(axes, float)
    a[9:12], a[12:15], a[15:18] = ax[0], ax[1], ax[2]
    return a


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = [f(np.deg2rad(v)) for v in (rx, ry, rz) for f in (np.cos, np.sin)]
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def blob_volume(shape, seed, amp=900.0):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    v = np.zeros(shape)
    for _ in range(6):
        c = [rng.uniform(0.3, 0.7) * s for s in shape]
        r = [rng.uniform(0.12, 0.3) * s for s in shape]
        v += rng.uniform(0.3, 1.0) * np.exp(-(((z - c[0]) / r[0]) ** 2 + ((y - c[1]) / r[1]) ** 2 + ((x - c[2]) / r[2]) ** 2))
    v = amp * v / v.max() + rng.uniform(0, 25, shape)
    return v


def stack_case():
    """Two thick-slice stacks of the same object: different grids, orientation, voxel sizes (dx != dy on the source)."""
    vol = blob_volume((24, 40, 44), 1)
    vol[vol < 60] = 0                                       # background 0 = the target padding of StackRegistrations
    ta = attrs((44, 40, 24), (1.1, 1.1, 2.5))
    R = rot(4, -3, 6)
    sa = attrs((44, 40, 24), (1.0, 1.2, 2.5), org=(1.5, -2.0, 1.0), axes=R.T)
    return vol, ta, vol[::-1, :, :].copy() * 0.8 + 3, sa


def slice_case():
    """A thick slice (z = 1, padding -1) against an isotropic volume with padding -1 outside a mask."""
    vol = blob_volume((40, 48, 48), 2)
    zz, yy, xx = np.meshgrid(np.arange(40), np.arange(48), np.arange(48), indexing="ij")
    vol[((zz - 20) / 18.0) ** 2 + ((yy - 24) / 21.0) ** 2 + ((xx - 24) / 21.0) ** 2 > 1] = -1
    va = attrs((48, 48, 40), (1.0, 1.0, 1.0))
    sl = vol[20].copy()
    sl[:, :6] = -1
    R = rot(2, 3, -4)
    sa = attrs((48, 48, 1), (1.0, 1.0, 2.5), org=(0.8, -0.6, 0.4), axes=R.T)
    return sl[None], sa, vol, va


@pytest.mark.parametrize("kind,case", [(0, "stack"), (1, "slice")])
def test_prepared_images_and_similarity_are_identical(kind, case):
    from fetalreconstruction_b200 import rreg
    t, ta, s, sa = stack_case() if case == "stack" else slice_case()
    b = _gpu()
    T, S = ri.Image.new(ta, t), ri.Image.new(sa, s)
    for level in (2, 1, 0):
        for dof in (np.zeros(6), np.array([1.3, -0.7, 0.4, 2.0, -1.5, 3.0])):
            ref_sim, rt, rs = ri.reg_probe(T, S, kind, level, dof)
            sim, pt, pta, ps, psa = rreg.register(b, [rreg.to_grey(t), rreg.to_grey(s)], [ta, sa], [0], [1], kind, [dof], level_only=level, want_prepared=True)
            np.testing.assert_array_equal(pta, rt.attrs, err_msg=f"target attributes, level {level}")
            np.testing.assert_array_equal(psa, rs.attrs, err_msg=f"source attributes, level {level}")
            np.testing.assert_array_equal(pt, rt.data.astype(np.int16), err_msg=f"prepared target, level {level}")
            np.testing.assert_array_equal(ps, rs.data.astype(np.int16), err_msg=f"prepared source, level {level}")
            assert sim[0] == ref_sim, (level, dof, sim[0], ref_sim, sim[0] - ref_sim)


@pytest.mark.parametrize("kind,case", [(0, "stack"), (1, "slice")])
def test_registration_follows_the_reference_path(kind, case):
    from fetalreconstruction_b200 import rreg
    t, ta, s, sa = stack_case() if case == "stack" else slice_case()
    b = _gpu()
    start = np.array([0.5, -0.5, 0.25, 1.0, 0.0, -1.0])
    want = ri.rigid_register(ri.Image.new(ta, t), ri.Image.new(sa, s), kind, start)
    got, sim, evals = rreg.register(b, [rreg.to_grey(t), rreg.to_grey(s)], [ta, sa], [0], [1], kind, [start])
    assert evals > 30
    np.testing.assert_array_equal(got[0], want)
    assert np.abs(want - start).max() > 1e-3                  # the optimiser did move


def test_batched_items_equal_one_at_a_time():
    """Many slices against one volume in one call = each of them alone (lockstep rounds do not couple the items)."""
    from fetalreconstruction_b200 import rreg
    sl, sa, vol, va = slice_case()
    b = _gpu()
    images, at = [rreg.to_grey(vol)], [va]
    starts = []
    rng = np.random.default_rng(5)
    for k in range(5):
        a = sa.copy(); a[6:9] += rng.uniform(-1, 1, 3)
        images.append(rreg.to_grey(np.where(sl >= 0, sl * (0.8 + 0.1 * k), -1))); at.append(a)
        starts.append(rng.uniform(-0.5, 0.5, 6))
    got, _, _ = rreg.register(b, images, at, list(range(1, 6)), [0] * 5, 1, starts)
    for k in range(5):
        one, _, _ = rreg.register(b, [images[0], images[k + 1]], [at[0], at[k + 1]], [1], [0], 1, [starts[k]])
        np.testing.assert_array_equal(got[k], one[0])


def test_pvr_patch_registration_equals_the_reference_per_patch():
    """pvr.PatchRegistration (all patches of a case in one device call) against the reference's own registration object run patch
    by patch as patchBased2D3DRegistration<T>::runHybrid does (ParallelPatchToVolumeRegistration, patchBased2D3DRegistration.cpp:88-168):
    target = the patch cast to irtkGreyImage with its origin moved into the transformation, source = the reconstruction cast to
    irtkGreyImage, GuessParameterSliceToVolume, target padding -1."""
    from fetalreconstruction_b200 import rreg
    from fetalreconstruction_b200.geometry import rigid_matrix, rigid_parameters
    from fetalreconstruction_b200.pvr import PatchReconstruction, PatchRegistration
    from pvr_case import make_pvr_case, setup_backend
    case = make_pvr_case(seed=21, vol=40, n_stacks=2, slices=5, size=40, pbb=(16, 16), stride=(8, 8))
    ds = case["ds"]
    n = min(len(case["attrs"]), 24)
    sub = dict(case)
    sub["per_stack"] = [min(case["per_stack"][0], n), n - min(case["per_stack"][0], n)]
    sub["attrs"] = case["attrs"][:n]
    for k in ("cube", "i2w", "w2i", "T", "Tinv"):
        sub[k] = np.ascontiguousarray(case[k][:n])
    rng = np.random.default_rng(3)
    start = [case["trans"][k] @ rigid_matrix(*(rng.uniform(-1, 1, 3)), *(rng.uniform(-2, 2, 3))) for k in range(n)]
    b = setup_backend(PatchReconstruction(0), sub, device_patch_init=False)
    # a smooth "reconstruction" to register against: the phantom itself, -1 outside the mask (what MaskVolume leaves)
    vol = np.where(ds.mask > 0, ds.truth, -1.0).astype(np.float32)
    b.recon_copyFromHost(vol.ravel())
    reg = PatchRegistration(b, sub["attrs"], sub["cube"], start, ds.vol_attr)
    reg()
    src = ri.Image.new(rreg.attrs18(ds.vol_attr), vol.astype(np.float64))
    moved = 0.0
    for k in range(n):
        a18 = rreg.attrs18(sub["attrs"][k])
        mo = np.eye(4); mo[:3, 3] = a18[6:9]; a18[6:9] = 0
        want = ri.rigid_register(ri.Image.new(a18, sub["cube"][k][None].astype(np.float64)), src, 1, rigid_parameters(start[k] @ mo))
        np.testing.assert_array_equal(rigid_parameters(reg.T[k] @ mo), rigid_parameters(rigid_matrix(*want)), err_msg=f"patch {k}")
        moved = max(moved, float(np.abs(want - rigid_parameters(start[k] @ mo)).max()))
    assert moved > 0.05 and reg.evaluations > 100 * n
