"""The C++ host + SVRreconstructionGPU command line (host/), as far as it runs without a GPU: option parsing and error
behaviour, NIfTI in -> NIfTI out geometry, and the reference's set-up pipeline (crop to mask, template, mask resampling,
intensity matching, slice creation and masking; reconstruction.cc:160-815 / irtkReconstructionGPU.cc cites in
host/svr_reconstruction.h) checked through `--dump_setup` on a seeded synthetic acquisition.  The device part of the CLI
is exercised on the GPU box (tools/gpu_c2.sh, bundled 3T data)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
CLI = os.path.join(HOST, "SVRreconstructionGPU")


@pytest.fixture(scope="module")
def cli(built_lib):
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["make", "-C", HOST, "CXX=/usr/bin/g++"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return CLI


def write_nifti(path, data_zyx, affine, pixdim):
    """Minimal NIfTI-1 writer (float32, sform only) for the tests."""
    z, y, x = data_zyx.shape
    h = bytearray(348)
    struct.pack_into("<i", h, 0, 348)
    struct.pack_into("<8h", h, 40, 3, x, y, z, 1, 1, 1, 1)
    struct.pack_into("<hh", h, 70, 16, 32)
    struct.pack_into("<8f", h, 76, 1.0, pixdim[0], pixdim[1], pixdim[2], 1, 1, 1, 1)
    struct.pack_into("<f", h, 108, 352.0)
    struct.pack_into("<ff", h, 112, 1.0, 0.0)
    struct.pack_into("<hh", h, 252, 0, 1)                      # qform_code 0, sform_code 1
    for r in range(3):
        struct.pack_into("<4f", h, 280 + 16 * r, *[float(v) for v in affine[r]])
    h[344:348] = b"n+1\0"
    with open(path, "wb") as f:
        f.write(bytes(h) + b"\0\0\0\0" + np.ascontiguousarray(data_zyx, "<f4").tobytes())


def read_nifti(path):
    raw = open(path, "rb").read()
    dim = struct.unpack_from("<8h", raw, 40)
    dt = struct.unpack_from("<h", raw, 70)[0]
    off = int(struct.unpack_from("<f", raw, 108)[0])
    aff = np.eye(4)
    for r in range(3):
        aff[r] = struct.unpack_from("<4f", raw, 280 + 16 * r)
    n = dim[1] * dim[2] * dim[3]
    data = np.frombuffer(raw, {16: "<f4", 64: "<f8"}[dt], n, off).reshape(dim[3], dim[2], dim[1])
    qb, qc, qd, qx, qy, qz = struct.unpack_from("<6f", raw, 256)
    return data, aff, dict(qform_code=struct.unpack_from("<h", raw, 252)[0], quatern=(qb, qc, qd), qoffset=(qx, qy, qz), datatype=dt)


def run(cli, args, cwd):
    return subprocess.run([cli] + args, cwd=cwd, capture_output=True, text=True, timeout=600)


def test_cli_option_errors(cli, tmp_path):
    assert run(cli, ["--help"], tmp_path).returncode == 0
    r = run(cli, ["-i", "a.nii"], tmp_path)                        # -o is required
    assert r.returncode != 0 and "--output" in r.stderr
    r = run(cli, ["-o", "out.nii", "-i", "a.nii", "--bogus"], tmp_path)
    assert r.returncode != 0 and "unrecognised option" in r.stderr
    r = run(cli, ["-o", "out.nii", "-i", "a.nii", "--useCPU"], tmp_path)
    assert r.returncode != 0 and "no CPU" in r.stderr              # no CPU fallback by design
    r = run(cli, ["-o", "out.nii", "-i", "missing.nii"], tmp_path)
    assert r.returncode != 0 and "cannot read" in r.stderr


@pytest.fixture(scope="module")
def acquisition(tmp_path_factory):
    """Oblique synthetic stacks written as NIfTI + the mask on the volume grid (stack 0 is the template: identity)."""
    from fetalreconstruction_b200.phantom import make_dataset, small_config
    d = tmp_path_factory.mktemp("acq")
    cfg = small_config(seed=3, vol=48, n_stacks=3, slices=20, size=44, inplane=1.1, spacing=2.0)
    cfg.motion_mm = cfg.motion_deg = 0.0
    cfg.noise = 0.0
    cfg.corrupt_fraction = 0.0
    cfg.mask_semi_axis = 0.36
    ds = make_dataset(cfg)
    Ny, Nx = ds.slices.shape[1:]
    names, stacks, affs = [], [], []
    rng = np.random.default_rng(0)
    for s, attr in enumerate(ds.stack_attrs):
        # un-masked stack content: a smooth positive field so every voxel is data (the CLI does the masking)
        zz, yy, xx = np.meshgrid(np.arange(cfg.slices_per_stack), np.arange(Ny), np.arange(Nx), indexing="ij")
        vol = (100.0 + 10 * s + 3.0 * xx + 2.0 * yy + 5.0 * zz + rng.uniform(0, 1, xx.shape)).astype(np.float32)
        aff = attr.image_to_world()
        p = os.path.join(d, f"stack_{s}.nii")
        write_nifti(p, vol, aff, (attr.dx, attr.dy, attr.dz))
        names.append(p); stacks.append(vol); affs.append(aff)
    mp = os.path.join(d, "mask.nii")
    write_nifti(mp, ds.mask.astype(np.float32), ds.vol_attr.image_to_world(), (cfg.vol_voxel,) * 3)
    return dict(dir=str(d), names=names, stacks=stacks, affs=affs, mask_path=mp, ds=ds, cfg=cfg)


def test_setup_pipeline_on_synthetic_stacks(cli, acquisition, tmp_path):
    a = acquisition
    out = tmp_path / "dump"
    out.mkdir()
    r = run(cli, ["-o", "recon.nii.gz", "-i"] + a["names"] + ["-m", a["mask_path"], "--resolution", "1.0", "--smooth_mask", "0", "--noStackRegistration",
                  "--dump_setup", str(out)], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    idx = dict(line.split() for line in open(out / "index.txt"))
    S, Nx, Ny = int(idx["S"]), int(idx["Nx"]), int(idx["Ny"])
    slices = np.fromfile(out / "slices.f32", np.float32).reshape(S, Ny, Nx)
    sizes = np.fromfile(out / "sizes.i32", np.int32).reshape(S, 2)
    I2W = np.fromfile(out / "I2W.f32", np.float32).reshape(S, 4, 4).astype(np.float64)
    W2I = np.fromfile(out / "W2I.f32", np.float32).reshape(S, 4, 4).astype(np.float64)
    T = np.fromfile(out / "T.f32", np.float32).reshape(S, 4, 4)
    dims = np.fromfile(out / "dims.f32", np.float32).reshape(S, 3)
    sidx = np.fromfile(out / "stack_index.i32", np.int32)
    factor = np.fromfile(out / "stack_factor.f32", np.float32)
    mask = np.fromfile(out / "mask.f32", np.float32).reshape(int(idx["vz"]), int(idx["vy"]), int(idx["vx"]))
    rw2i = np.fromfile(out / "recon_w2i.f32", np.float32).reshape(4, 4).astype(np.float64)

    # stacks were cropped to the mask: fewer slices than acquired, all stacks represented, identity stack transforms
    assert 0 < S < a["cfg"].n_stacks * a["cfg"].slices_per_stack and set(sidx) == {0, 1, 2}
    assert np.allclose(T, np.eye(4)[None])
    assert np.allclose(dims[:, :2], 1.1, atol=1e-6) and np.allclose(dims[:, 2], 4.0)            # thickness = 2 * dz (reconstruction.cc:423-434)
    assert np.allclose(np.einsum("sij,sjk->sik", I2W, W2I), np.eye(4)[None], atol=1e-4)
    assert set(np.unique(mask)) <= {0.0, 1.0} and mask.sum() > 0
    assert factor.shape == (3,) and np.all(factor > 0)

    checked = 0
    for n in range(S):
        sx, sy = sizes[n]
        s = slices[n]
        assert np.all(s[sy:, :] == -1) and np.all(s[:, sx:] == -1)                                # top-left packing, -1 padding
        valid = np.argwhere(s[:sy, :sx] != -1)
        if not len(valid):
            continue
        st = sidx[n]
        pix = np.concatenate([valid[:, ::-1].astype(np.float64), np.zeros((len(valid), 1)), np.ones((len(valid), 1))], 1)
        world = pix @ I2W[n].T
        # (1) the value is the stack voxel at that world position times the stack's intensity factor
        ijk = world @ np.linalg.inv(a["affs"][st]).T
        r_ijk = np.rint(ijk[:, :3]).astype(int)
        assert np.abs(ijk[:, :3] - r_ijk).max() < 1e-3, "cropped slice grid is not aligned with its stack"
        ref = a["stacks"][st][r_ijk[:, 2], r_ijk[:, 1], r_ijk[:, 0]] * factor[st]
        got = s[valid[:, 0], valid[:, 1]]
        assert np.allclose(got, ref, rtol=2e-6), (n, np.abs(got - ref).max())
        # (2) MaskSlices: the pixel centre maps (rounded) onto a mask voxel (irtkReconstructionGPU.cc:1940-1988)
        v = np.floor((world @ rw2i.T)[:, :3] + 0.5).astype(int)              # round(): halves away from zero, as the C++ / IRTK code
        assert np.all(mask[v[:, 2], v[:, 1], v[:, 0]] == 1)
        checked += len(valid)
    assert checked > 1000
    # intensity matching: the in-mask mean of every stack is --average (700) after scaling (irtkReconstructionGPU.cc:1375-1493)
    for st in range(3):
        vals = np.concatenate([slices[n][slices[n] != -1] for n in range(S) if sidx[n] == st])
        assert abs(vals.mean() - 700.0) < 0.05 * 700.0

    # NIfTI out: stack<i>.nii is the stack as read (FLOAT64, reconstruction.cc:320-325) with the same voxel -> world map
    data, aff, meta = read_nifti(tmp_path / "stack1.nii")
    assert meta["datatype"] == 64 and meta["qform_code"] == 1
    assert np.allclose(aff, a["affs"][1], atol=1e-4)
    assert np.allclose(data, a["stacks"][1])


def test_dof_roundtrip_and_transformation_option(cli, acquisition, tmp_path):
    """-t: .dof files are big-endian {magic 815007, type, 6 dofs} and are inverted on input (reconstruction.cc:329-353,399)."""
    a = acquisition
    dof = tmp_path / "t1.dof"
    params = [1.5, -2.0, 0.5, 2.0, -1.0, 3.0]
    # IRTK_MAGIC 815007, rigid = 2 (irtkTransformation.h:24-27), 6 doubles
    dof.write_bytes(struct.pack(">III", 815007, 2, 6) + struct.pack(">6d", *params))
    out = tmp_path / "dump"
    out.mkdir()
    r = run(cli, ["-o", "recon.nii.gz", "-i"] + a["names"][:2] + ["-t", "id", str(dof), "-m", a["mask_path"], "--resolution", "1.0",
                  "--smooth_mask", "0", "--noStackRegistration", "--dump_setup", str(out)], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    S = int(dict(line.split() for line in open(out / "index.txt"))["S"])
    T = np.fromfile(out / "T.f32", np.float32).reshape(S, 4, 4).astype(np.float64)
    sidx = np.fromfile(out / "stack_index.i32", np.int32)
    from fetalreconstruction_b200.geometry import rigid_matrix
    want = np.linalg.inv(rigid_matrix(*params))
    assert np.allclose(T[sidx == 0], np.eye(4)[None], atol=1e-6)
    assert np.allclose(T[sidx == 1], want[None], atol=1e-5)
