"""GPU: the C++ host (host/SVRreconstructionGPU) end to end -- NIfTI stacks in, reconstructed NIfTI volume out -- against
the Python host driving the same library on the inputs the CLI itself prepared (`--dump_setup`).  One outer iteration
(the reference runs no registration in iteration 0), so both sides execute the same device calls in the same order and
the comparison is tight: differences come from the float atomics of the scatter only."""
import os
import subprocess

import numpy as np
import pytest

from test_host_cli import acquisition, cli, read_nifti, run  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def test_cli_end_to_end_matches_python_host(cli, acquisition, tmp_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from c2_parity import load_setup
    from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
    from fetalreconstruction_b200.reconstruction import Reconstruction
    a = acquisition
    common = ["-i"] + a["names"] + ["-m", a["mask_path"], "--resolution", "1.0", "--smooth_mask", "0", "--iterations", "1",
                                    "--rec_iterations_last", "3"]
    dump = tmp_path / "dump"
    dump.mkdir()
    r = run(cli, ["-o", "recon.nii.gz"] + common + ["--dump_setup", str(dump)], tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    r = run(cli, ["-o", "recon.nii.gz"] + common, tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    # the reference's output files
    for name in ("stack0.nii", "GaussianReconstruction_GPU0.nii", "image0_GPU.nii.gz", "log-reconstruction.txt", "log-evaluation.txt",
                 "log-registration.txt"):
        assert (tmp_path / name).exists(), name
    assert any(f.startswith("performance_GPU_") for f in os.listdir(tmp_path))
    assert "Included slices GPU:" in (tmp_path / "log-evaluation.txt").read_text()
    import gzip
    raw = gzip.open(tmp_path / "recon.nii.gz").read()
    (tmp_path / "recon.nii").write_bytes(raw)
    vol, aff, meta = read_nifti(tmp_path / "recon.nii")
    assert meta["datatype"] == 64 and np.isfinite(vol).all() and (vol != 0).any()

    ds = load_setup(str(dump))
    assert vol.shape == ds.mask.shape
    b = Reconstruction(0)
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(iterations=1, rec_iterations_last=3))
    p.InitializeEMGPU(ds.slices)
    p.outer_iteration(0)
    b.RestoreSliceIntensities(ds.stack_factor, ds.stack_index)
    p.ScaleVolumeGPU()
    want = b.syncCPU().reshape(ds.mask.shape)
    scale = np.sqrt(np.mean(want[want != 0] ** 2))
    d = np.abs(vol - want) / scale
    assert np.sqrt(np.mean(d ** 2)) <= 1e-5 and d.max() <= 1e-3, (np.sqrt(np.mean(d ** 2)), d.max())
    # the volume's voxel -> world map is the template's (stack 0 cropped to the mask, +2 slices, 1 mm isotropic)
    assert np.allclose(np.abs(np.linalg.det(aff[:3, :3])), 1.0, atol=1e-4)
