"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle).

CPU: the oracle still reproduces them (regression pin; sums formed with OpenMP double atomics may differ in the
last float bit).  GPU: the CUDA path reproduces them within the parity tolerances of test_gpu_parity.py.
The reference itself holds no golden vectors for this path; the vectors its own CUDA code produces on a B200 are in
tests/golden/ref_*.npz and pinned by tests/test_ref_golden.py."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import rel_stats

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def _cmp(got, gold, vol_rms, vol_max, names):
    for n in names:
        m, r = rel_stats(np.asarray(got[n], np.float64), np.asarray(gold[n], np.float64))
        assert r <= vol_rms and m <= vol_max, f"{n}: rms/max = {(r, m)}"


SVR_FLOAT = ["gaussian_recon", "volweights", "simslices", "slice_potential", "volume", "scale", "slice_weight", "em"]


def test_oracle_reproduces_svr_golden():
    got, gold = mg.svr_case(), _load("svr_small")
    assert np.array_equal(got["voxel_num"], gold["voxel_num"]) and np.array_equal(got["siminside"], gold["siminside"])
    _cmp(got, gold, 1e-6, 1e-5, SVR_FLOAT)
    assert float(got["sigma"]) == pytest.approx(float(gold["sigma"]), rel=1e-6)


def test_oracle_reproduces_registration_golden():
    got, gold = mg.reg_case(), _load("reg_small")
    assert float(got["resampled_checksum"]) == pytest.approx(float(gold["resampled_checksum"]), rel=1e-12)
    assert np.abs(got["sim_level0"] - gold["sim_level0"]).max() <= 1e-6
    assert np.abs(got["sim_level1"] - gold["sim_level1"]).max() <= 1e-6
    assert np.abs(got["transforms_out"] - gold["transforms_out"]).max() <= 1e-5
    assert int(got["evaluations"]) == int(gold["evaluations"])


def test_oracle_reproduces_pvr_golden():
    got, gold = mg.pvr_case(), _load("pvr_small")
    assert np.array_equal(got["per_stack"], gold["per_stack"])
    assert float(got["patches_checksum"]) == pytest.approx(float(gold["patches_checksum"]), rel=1e-12)
    _cmp(got, gold, 1e-6, 1e-5, ["volume", "em", "patch_potential"])


@pytest.mark.gpu
def test_cuda_matches_svr_golden():
    from fetalreconstruction_b200.reconstruction import Reconstruction
    got, gold = mg.svr_case(Reconstruction(0)), _load("svr_small")
    assert np.array_equal(got["voxel_num"], gold["voxel_num"])
    assert np.mean(got["siminside"] == gold["siminside"]) > 0.9999
    _cmp(got, gold, 3e-4, 3e-2, ["gaussian_recon", "volweights", "volume"])
    _cmp(got, gold, 5e-4, 8e-2, ["simslices"])
    _cmp(got, gold, 1e-3, 1e-2, ["slice_potential", "scale", "slice_weight", "em"])


@pytest.mark.gpu
def test_cuda_matches_registration_golden():
    """Oracle-generated golden: the oracle filters the volume in software, the CUDA path through the texture unit (as
    the reference; tolerances explained in test_gpu_registration.py).  The tight comparison is the one against the
    reference's own output in test_ref_golden.py."""
    from fetalreconstruction_b200.reconstruction import Reconstruction
    b = Reconstruction(0)
    got, gold = mg.reg_case(b), _load("reg_small")
    assert np.abs(got["sim_level0"] - gold["sim_level0"]).max() <= 3e-2
    assert np.abs(got["sim_level1"] - gold["sim_level1"]).max() <= 3e-2
    s_ours = b.evaluateCostsMultipleSlices(got["transforms_out"], 0)
    s_gold = b.evaluateCostsMultipleSlices(gold["transforms_out"], 0)
    assert s_ours.mean() >= s_gold.mean() - 0.01, (s_ours, s_gold)


@pytest.mark.gpu
def test_cuda_matches_pvr_golden():
    from fetalreconstruction_b200.pvr import PatchReconstruction
    got, gold = mg.pvr_case(PatchReconstruction(0)), _load("pvr_small")
    assert np.array_equal(got["per_stack"], gold["per_stack"])
    assert float(got["patches_checksum"]) == pytest.approx(float(gold["patches_checksum"]), rel=1e-5)
    _cmp(got, gold, 3e-4, 3e-2, ["volume"])
    _cmp(got, gold, 1e-3, 1e-2, ["em", "patch_potential"])
