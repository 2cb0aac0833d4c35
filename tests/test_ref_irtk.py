"""CPU: our host-side restatements of the IRTK substrate against the reference's OWN IRTK (oracle/_ref/libref_irtk.so: the vendored
IRTKSimple2 compiled unmodified, `make -C oracle ref_irtk`): image <-> world matrices, rigid parameters <-> matrix, the NIfTI
geometry the reference's reader derives from an sform header, .dof files, and the registration engine's parameter guesses."""
import os

import numpy as np
import pytest

from oracle import ref_irtk as ri

pytestmark = pytest.mark.skipif(not ri.available(), reason="oracle/_ref/libref_irtk.so not built (make -C oracle ref_irtk)")


def random_attrs(rng):
    from fetalreconstruction_b200.geometry import ImageAttributes, rigid_matrix
    R = rigid_matrix(0, 0, 0, *rng.uniform(-40, 40, 3))[:3, :3]
    return ImageAttributes(int(rng.integers(2, 60)), int(rng.integers(2, 60)), int(rng.integers(1, 40)), *rng.uniform(0.5, 3.0, 3),
                           rng.uniform(-50, 50, 3), R[:, 0].copy(), R[:, 1].copy(), R[:, 2].copy())


def a18(a):
    return np.concatenate([[a.x, a.y, a.z, a.dx, a.dy, a.dz], a.origin, a.xaxis, a.yaxis, a.zaxis])


def test_image_world_matrices_match_irtk():
    """geometry.ImageAttributes (and with it host/svr_image.cc, tested against it in test_host_cli.py) vs irtkBaseImage.cc:79-147."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = random_attrs(rng)
        i2w, w2i = ri.Image.new(a18(a)).matrices()
        np.testing.assert_allclose(a.image_to_world(), i2w, rtol=0, atol=1e-12)
        np.testing.assert_allclose(a.world_to_image(), w2i, rtol=0, atol=1e-12)


def test_rigid_parameters_match_irtk():
    """geometry.rigid_matrix / rigid_parameters vs irtkRigidTransformation.cc:26-149 (degrees, Rx Ry Rz layout, decomposition)."""
    from fetalreconstruction_b200.geometry import rigid_matrix, rigid_parameters
    rng = np.random.default_rng(1)
    for _ in range(30):
        d = np.concatenate([rng.uniform(-30, 30, 3), rng.uniform(-80, 80, 3)])
        m = ri.rigid_matrix(d)
        np.testing.assert_allclose(rigid_matrix(*d), m, rtol=0, atol=1e-14)
        np.testing.assert_allclose(rigid_parameters(m), ri.rigid_from_matrix(m), rtol=0, atol=1e-9)
        np.testing.assert_allclose(ri.rigid_from_matrix(m), d, rtol=0, atol=1e-9)


def test_nifti_sform_geometry_as_the_reference_reads_it(tmp_path):
    """A NIfTI file written with an sform (the writer of tests/test_host_cli.py) read back by the reference's own reader
    (irtkFileNIFTIToImage.cc:227-373): centre-of-image origin, axes = matrix columns / voxel size, voxel sizes, data."""
    from test_host_cli import write_nifti
    rng = np.random.default_rng(2)
    for k in range(4):
        a = random_attrs(rng)
        a.z = max(a.z, 2)
        data = rng.uniform(0, 1000, (a.z, a.y, a.x)).astype(np.float32)
        p = str(tmp_path / f"img{k}.nii")
        write_nifti(p, data, a.image_to_world(), (a.dx, a.dy, a.dz))
        im = ri.Image.read(p)
        got = im.attrs
        np.testing.assert_array_equal(got[:3], [a.x, a.y, a.z])
        np.testing.assert_allclose(got[3:6], [a.dx, a.dy, a.dz], rtol=1e-6)            # pixdim is float32 in the header
        np.testing.assert_allclose(got[6:9], a.origin, atol=2e-4)                       # srow is float32
        np.testing.assert_allclose(got[9:18], np.concatenate([a.xaxis, a.yaxis, a.zaxis]), atol=1e-6)
        np.testing.assert_array_equal(im.data.astype(np.float32), data)


def test_dof_file_round_trip_through_irtk(tmp_path):
    """.dof written by IRTK (irtkRigidTransformation.cc:392-451: big endian, magic 815007, type 2, 6 doubles) parsed by hand, and back."""
    import struct
    d = np.array([1.5, -2.25, 3.0, 10.0, -20.0, 30.0])
    p = str(tmp_path / "t.dof")
    ri.lib().rirtk_dof_write(d.ctypes.data_as(ri.vp), p.encode())
    raw = open(p, "rb").read()
    magic, typ, n = struct.unpack(">III", raw[:12])
    assert (magic, typ, n) == (815007, 2, 6) and len(raw) == 12 + 48
    np.testing.assert_array_equal(np.array(struct.unpack(">6d", raw[12:])), d)
    out = np.zeros(6)
    assert ri.lib().rirtk_dof_read(p.encode(), out.ctypes.data_as(ri.vp)) == 0
    np.testing.assert_array_equal(out, d)
