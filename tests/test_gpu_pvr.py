"""GPU: the PVR path (include/pvr_abi.h) against oracle/pvr_oracle.c, step by step and through the pipeline.
Tolerances as in test_gpu_parity.py (same tap loop, same epsilon-skip discontinuity), relative to the RMS of the
oracle's non-zeros."""
import numpy as np
import pytest

from conftest import rel_stats
from fetalreconstruction_b200.pvr import PVRParams, PVRPipeline
from oracle.oracle_backend_pvr import OraclePatchReconstruction
from pvr_case import make_pvr_case, setup_backend

pytestmark = pytest.mark.gpu

VOL_RMS, VOL_MAX = 3e-4, 3e-2
# RMS + a count bound on flipped pixels (see test_gpu_parity.py) + what one flipped tap can move
PIX_RMS, PIX_THR, PIX_FRAC, PIX_MAX = 5e-4, 2e-3, 1e-3, 8e-2


def _gpu():
    import torch
    assert torch.cuda.is_available()
    from fetalreconstruction_b200.pvr import PatchReconstruction
    return PatchReconstruction(0)


def check_pixels(a, b, name):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    nz = b[b != 0]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - b) / scale
    off = int(np.count_nonzero(d > PIX_THR))
    st = (float(np.sqrt(np.mean(d ** 2))), off, float(d.max()))
    assert st[0] <= PIX_RMS and off <= max(2, int(PIX_FRAC * d.size)) and st[2] <= PIX_MAX, f"{name}: rms/beyond {PIX_THR}/max = {st} of {d.size}"


def check_volume(a, b, name):
    m, r = rel_stats(a, b)
    assert r <= VOL_RMS and m <= VOL_MAX, f"{name}: rms/max = {(r, m)}"


@pytest.fixture(scope="module")
def pair():
    case = make_pvr_case()
    g = setup_backend(_gpu(), case)
    o = setup_backend(OraclePatchReconstruction(), case)
    return case, g, o


def test_patch_init_matches_oracle(pair):
    case, g, o = pair
    a, b = g.patches_copyToHost(), o.patches_copyToHost()
    assert a.shape == b.shape and a.shape[0] == sum(case["per_stack"])
    assert np.mean(a == b) > 0.9999


def test_psf_reconstruction_simulate_and_em(pair):
    case, g, o = pair
    ds = case["ds"]
    for b in (g, o):
        b.rs_initializeEMValues()
        b.recon_reset()
        b.patchBasedPSFReconstruction_gpu()
    check_pixels(g.debugPSFsums(), o.debugPSFsums(), "psf sums")
    check_volume(g.getVolWeights(), o.getVolWeights(), "volume weights")
    check_volume(g.recon_copyToHost(), o.recon_copyToHost(), "un-equalised volume")
    for b in (g, o):
        b.recon_equalize()
    check_volume(g.recon_copyToHost(), o.recon_copyToHost(), "equalised volume")
    for b in (g, o):
        b.patchBasedSimulatePatches_gpu()
    check_pixels(g.debugSimpatches(), o.debugSimpatches(), "simulated patches")
    check_pixels(g.debugSimweights(), o.debugSimweights(), "simulated weights")
    assert np.mean(g.debugSiminside() == o.debugSiminside()) > 0.9999
    sg, so = g.rs_InitializeRobustStatistics(), o.rs_InitializeRobustStatistics()
    assert sg == pytest.approx(so, rel=1e-4)
    m = 1.0 / (2.1 * ds.max_intensity - 1.9 * ds.min_intensity)
    pg, po = g.rs_estep_device(m, so, 0.9), o.rs_estep_device(m, so, 0.9)
    assert np.abs(pg - po).max() < 1e-4
    check_pixels(g.debugWeights(), o.debugWeights(), "voxel weights")
    scg, sco = g.rs_Scale(), o.rs_Scale()
    assert np.abs(scg - sco).max() < 1e-4
    for b in (g, o):
        b.recon_resetAddonCmap()
        b.superresolution_run()
    # the CUDA accumulator also holds voxels outside the mask (the per-tap mask test is applied per voxel by the
    # regulariser prep / equalize): compare inside the mask
    mk = case["mask"].ravel() > 0
    check_volume(g.debugAddon()[mk], o.debugAddon()[mk], "addon")
    check_volume(g.debugConfidenceMap()[mk], o.debugConfidenceMap()[mk], "confidence map")
    for b in (g, o):
        b.superresolution_regularize(False, 0.5, ds.min_intensity, ds.max_intensity, 1.0, 0.1)
    check_volume(g.recon_copyToHost(), o.recon_copyToHost(), "regularised volume")
    mg, mo = g.rs_MStep(2, 1e-4, so, 0.9, m), o.rs_MStep(2, 1e-4, so, 0.9, m)
    assert np.allclose(mg, mo, rtol=2e-4)


def test_pvr_pipeline_matches_oracle():
    case = make_pvr_case(seed=5)
    ds = case["ds"]
    vols = []
    for b in (_gpu(), OraclePatchReconstruction()):
        setup_backend(b, case)
        pipe = PVRPipeline(b, ds.min_intensity, ds.max_intensity, PVRParams(iterations=1, rec_iterations=3))
        vols.append(pipe.run())
    m = case["mask"].ravel() > 0
    check_volume(vols[0][m], vols[1][m], "PVR pipeline volume (in mask)")


def test_superpixel_masks_gate_pixels():
    case = make_pvr_case(seed=7)
    n = sum(case["per_stack"])
    rng = np.random.default_rng(0)
    spx = np.where(rng.uniform(size=(n, 4096)) < 0.7, b"1", b"0").astype("S1")
    g = setup_backend(_gpu(), case, spx=spx)
    o = setup_backend(OraclePatchReconstruction(), case, spx=spx)
    assert np.mean(g.patches_copyToHost() == o.patches_copyToHost()) > 0.9999
    for b in (g, o):
        b.rs_initializeEMValues()
        b.recon_reset()
        b.patchBasedPSFReconstruction_gpu()
    check_pixels(g.debugPSFsums(), o.debugPSFsums(), "psf sums (superpixels)")
    check_volume(g.recon_copyToHost(), o.recon_copyToHost(), "volume (superpixels)")


def test_pvr_errors_and_empty():
    from fetalreconstruction_b200.reconstruction import SVRError
    g = _gpu()
    with pytest.raises(SVRError):
        g.patchBasedPSFReconstruction_gpu()
    g.recon_init((8, 8, 8), (1, 1, 1), np.eye(4, dtype=np.float32).ravel(), np.eye(4, dtype=np.float32).ravel())
    g.recon_setMask(np.ones(512, np.int8))
    g.patches_init(16, 16, [0, 0], [(1, 1, 2), (1, 1, 2)])
    g.patches_set_matrices(np.zeros((0, 16)), np.zeros((0, 16)), np.zeros((0, 16)), np.zeros((0, 16)))
    g.rs_initializeEMValues()
    g.recon_reset()
    g.patchBasedPSFReconstruction_gpu()
    g.recon_equalize()
    assert np.all(g.recon_copyToHost() == 0)
