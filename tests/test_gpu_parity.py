"""GPU: parity of the CUDA path (through the C ABI) against the oracle on seeded inputs.

Tolerances.  The path is floating point; the reference itself is built with --use_fast_math and
accumulates with float atomics in arbitrary order, so its own output is only defined up to a
tolerance.  Two mechanisms set the floor (DESIGN.md "Parity"):
  (1) rounding: fast-math MUFU (sin/ex2/rsqrt.approx) and the factored tap position differ from the
      oracle's libm evaluation by ~1e-6 relative per tap;
  (2) the reference's discontinuous epsilon-skip rule (a tap is dropped when it differs from the
      previously accepted tap of its row by < 1e-5): a 1-ulp change flips the decision for taps that
      sit on the threshold, which changes ONE tap of ONE pixel (up to a few % of that pixel's PSF
      mass).  Flips hit a small fraction of pixels, so slice-level outputs are checked with a bulk
      bound (RMS, 99.9th percentile) plus a loose max bound, and volume-level outputs (sums over many
      pixels) with tight RMS / max-abs bounds.
All bounds are relative to the RMS of the oracle's non-zero values.
"""
import numpy as np
import pytest

from conftest import rel_stats
from fetalreconstruction_b200.phantom import make_dataset, small_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from oracle import oracle as orc
from oracle.oracle_backend import OracleReconstruction

pytestmark = pytest.mark.gpu

# Bounds are ~3-5x the deviations measured on B200 (profiles/r01_*_parity_report.json), e.g. v_PSF_sums
# rms 9e-5 / max 1.6e-2, GaussianReconstruction rms 1.6e-5 / max 2.9e-3, simulated slices rms 8.6e-6 /
# max 1.2e-3, final volume of the full loop rms 4.6e-5 / max 4.7e-3.  The max bounds are loose because a
# single flipped epsilon-skip decision moves one pixel's PSF mass by up to a few per cent.
# volume-level (many-pixel sums)
VOL_RMS, VOL_MAX = 3e-4, 3e-2
# slice-level: RMS, plus a COUNT bound on flipped pixels (at most PIX_FRAC of the pixels, and never fewer than 2 allowed, may be
# further off than PIX_THR) instead of an open maximum; the maximum itself is bounded by what ONE flipped tap can move
# (one tap of at most 1.0 against PSF sums of ~40-80: < 5e-2 of the RMS).
PIX_RMS, PIX_THR, PIX_FRAC, PIX_MAX = 3e-4, 1e-3, 3e-4, 5e-2

_STATS = {}


def _dump_stats():
    import json, os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if _STATS and os.path.isdir(out):
        with open(os.path.join(out, "r02_parity_stats.json"), "w") as f:
            json.dump(_STATS, f, indent=1)


import atexit
atexit.register(_dump_stats)


def _gpu():
    import torch
    assert torch.cuda.is_available(), "the gpu-marked tests need a CUDA device"
    from fetalreconstruction_b200.reconstruction import Reconstruction
    return Reconstruction(0)


def check_pixels(a, b, name, rms=PIX_RMS, thr=PIX_THR, frac=PIX_FRAC, mx=PIX_MAX):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    nz = b[b != 0]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    d = np.abs(a - b) / scale
    n_off = int(np.count_nonzero(d > thr))
    allowed = max(2, int(frac * d.size))
    stats = {"rms": float(np.sqrt(np.mean(d ** 2))), "beyond_thr": n_off, "thr": thr, "allowed": allowed, "max": float(d.max()), "n": int(d.size)}
    _STATS[name] = stats
    assert stats["rms"] <= rms and n_off <= allowed and stats["max"] <= mx, f"{name}: {stats} (bounds: rms {rms}, max {mx})"
    return stats


def check_volume(a, b, name, rms=VOL_RMS, mx=VOL_MAX):
    m, r = rel_stats(a, b)
    _STATS[name] = {"rms": r, "max": m}
    assert r <= rms and m <= mx, f"{name}: rms/max = {(r, m)}"
    return r, m


@pytest.fixture(scope="module")
def pair(small_ds):
    """The same call sequence on both backends up to the first E-step."""
    ds = small_ds
    g, o = _gpu(), OracleReconstruction()
    out = {}
    for name, b in (("gpu", g), ("orc", o)):
        upload_dataset(b, ds)
        b.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
        b.InitializeEMValues()
        out[name + "_voxel_num"] = b.GaussianReconstruction()
        out[name + "_recon0"] = b.syncCPU()
        out[name + "_volw"] = b.getVolWeights()
        out[name + "_psf"] = b.debugv_PSF_sums()
        out[name + "_inside"] = b.SimulateSlices()
        out[name + "_sim"] = b.debugSimslices()
        out[name + "_simw"] = b.debugSimweights()
        out[name + "_simi"] = b.debugSiminside()
    out["g"], out["o"], out["ds"] = g, o, ds
    return out


def test_psf_sums_and_voxel_counts(pair):
    check_pixels(pair["gpu_psf"], pair["orc_psf"], "v_PSF_sums")
    assert np.array_equal(pair["gpu_psf"] != 0, pair["orc_psf"] != 0)
    # per-slice pixel counts are integers: allow the odd threshold pixel
    assert np.abs(pair["gpu_voxel_num"].astype(int) - pair["orc_voxel_num"].astype(int)).max() <= 1


def test_gaussian_reconstruction_volume(pair):
    check_volume(pair["gpu_volw"], pair["orc_volw"], "volWeights", mx=6e-2)
    check_volume(pair["gpu_recon0"], pair["orc_recon0"], "GaussianReconstruction")
    mask = pair["ds"].mask.ravel()
    assert np.all(pair["gpu_recon0"][mask == 0] == 0)


def test_simulate_slices(pair):
    check_pixels(pair["gpu_sim"], pair["orc_sim"], "simulated slices")
    check_pixels(pair["gpu_simw"], pair["orc_simw"], "simulated weights")
    assert np.mean(pair["gpu_simi"] != pair["orc_simi"]) < 1e-4
    assert np.array_equal(pair["gpu_inside"], pair["orc_inside"])


def test_em_steps_and_superresolution(pair):
    g, o, ds = pair["g"], pair["o"], pair["ds"]
    sig_g, sig_o = g.InitializeRobustStatistics(), o.InitializeRobustStatistics()
    assert sig_g == pytest.approx(sig_o, rel=2e-3)
    pos = ds.slices[ds.slices > 0]
    m = 1.0 / (2.1 * pos.max() - 1.9 * pos.min())
    # identical parameters on both sides from here on, so each kernel is compared on its own
    pg, po = g.EStep(m, sig_o, 0.9), o.EStep(m, sig_o, 0.9)
    np.testing.assert_allclose(pg, po, rtol=5e-3, atol=1e-4)
    # measured on B200: rms 1.1e-6, max 1.2e-4, no pixel beyond 1e-3 (profiles/r01_v9_parity_report.txt)
    check_pixels(g.debugWeights(), o.debugWeights(), "EStep weights", rms=2e-5, thr=1e-3, frac=1e-4, mx=2e-2)
    sg, so = g.CalculateScaleVector(), o.CalculateScaleVector()
    np.testing.assert_allclose(sg, so, rtol=1e-3)
    np.testing.assert_array_equal(g.debugScalesDevice(), 1.0)       # the device lags one call behind
    sw = np.ones(ds.S, np.float32); sw[1] = 0.0; sw[4] = 0.37
    lam, delta = 0.02, 150.0
    args = (1, sw, False, min(1.0, 0.05 / lam), float(pos.min()), float(pos.max()), delta, lam * delta * delta)
    g.Superresolution(*args); o.Superresolution(*args)
    check_volume(g.debugConfidenceMap(), o.debugConfidenceMap(), "confidence map (normalised)")
    check_volume(g.debugAddon(), o.debugAddon(), "addon (normalised)", rms=5e-4, mx=6e-2)
    check_volume(g.syncCPU(), o.syncCPU(), "Superresolution volume")
    g.SimulateSlices(); o.SimulateSlices()
    a, b = g.MStep(2, 1e-4, sig_o, 0.9, m), o.MStep(2, 1e-4, sig_o, 0.9, m)
    np.testing.assert_allclose(a, b, rtol=5e-3)
    # second scale call: now the device holds the first result
    g.CalculateScaleVector(); o.CalculateScaleVector()
    np.testing.assert_allclose(g.debugScalesDevice(), so, rtol=1e-3)
    g.maskVolume(); o.maskVolume()
    assert np.array_equal(g.syncCPU() == -1, o.syncCPU() == -1)
    assert g.ScaleVolume() == pytest.approx(o.ScaleVolume(), rel=1e-3)


def test_full_pipeline_matches_oracle(small_ds):
    """The complete loop (outer iterations with Gaussian reconstruction, simulate, robust statistics, scale,
    super-resolution + regulariser, M-step, E-step, masking, final scaling) through the SAME host pipeline on
    both backends (schedule shortened to 2 outer x 3/5 inner iterations to keep the oracle at ~10 s)."""
    ds = small_ds
    vols = {}
    for name, b in (("gpu", _gpu()), ("orc", OracleReconstruction())):
        upload_dataset(b, ds)
        p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams(iterations=2, rec_iterations_first=3, rec_iterations_last=5))
        p.InitializeEMGPU(ds.slices)
        vols[name] = p.run()
        vols[name + "_w"] = p._slice_weight.copy()
        vols[name + "_s"] = p._scale.copy()
    inmask = ds.mask.ravel() != 0
    assert np.array_equal(vols["gpu"] == -1, vols["orc"] == -1)
    check_volume(vols["gpu"][inmask], vols["orc"][inmask], "final volume", rms=1e-3, mx=2e-2)
    np.testing.assert_allclose(vols["gpu_s"], vols["orc_s"], rtol=2e-3)
    np.testing.assert_allclose(vols["gpu_w"], vols["orc_w"], atol=2e-2)


# ---- edge cases ------------------------------------------------------------------------------------
def test_empty_and_all_padding_inputs():
    ds = make_dataset(small_config(seed=2, vol=16, n_stacks=1, slices=3, size=12))
    g = _gpu()
    # S = 0: every call is a no-op that leaves an all-zero volume
    ds0 = make_dataset(small_config(seed=2, vol=16, n_stacks=1, slices=3, size=12))
    upload_dataset(g, ds0, 0, 0)
    assert g.GaussianReconstruction().size == 0
    assert np.all(g.syncCPU() == 0)
    assert g.SimulateSlices().size == 0
    # all-padding slices: nothing is written anywhere
    ds.slices[:] = -1
    upload_dataset(g, ds)
    g.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
    g.InitializeEMValues()
    assert np.all(g.GaussianReconstruction() == 0)
    assert np.all(g.syncCPU() == 0) and np.all(g.debugv_PSF_sums() == 0)
    assert not g.SimulateSlices().any()
    assert np.all(g.EStep(1e-3, 1.0, 0.9) == -1)
    assert np.all(g.CalculateScaleVector() == 1.0)


def test_slices_overhanging_the_volume_and_identity_alignment():
    """Slices larger than the volume (taps outside, negative coordinates -> quirk Q4) and a grid-aligned
    stack where tap offsets are exactly 0 (the reference's sin(0)/0; deviation D5)."""
    cfg = small_config(seed=9, vol=20, n_stacks=2, slices=6, size=40, inplane=1.0, spacing=2.0)
    cfg.motion_mm = 0.0; cfg.motion_deg = 0.0; cfg.mask_semi_axis = 0.7
    ds = make_dataset(cfg)
    g, o = _gpu(), OracleReconstruction()
    res = {}
    for name, b in (("g", g), ("o", o)):
        upload_dataset(b, ds)
        b.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
        b.InitializeEMValues()
        b.GaussianReconstruction()
        res[name] = (b.syncCPU(), b.debugv_PSF_sums())
        b.SimulateSlices()
        res[name] += (b.debugSimslices(),)
    assert np.isfinite(res["g"][0]).all() and np.isfinite(res["g"][2]).all()
    check_volume(res["g"][0], res["o"][0], "overhang volume", rms=5e-4, mx=2e-2)
    # exactly aligned grids put many taps ON the skip threshold (mirror-image twins), so allow more flips
    check_pixels(res["g"][1], res["o"][1], "overhang psf sums", rms=2e-2, thr=5e-2, frac=2e-2, mx=0.2)


def test_stale_psf_sums_persist_across_calls(small_ds):
    """v_PSF_sums is only written when sume > 0.5 (cuda2.cu:251-258) and never cleared: after the transforms
    move a slice out of the volume its old sums stay (reproduced quirk)."""
    ds = small_ds
    g = _gpu()
    upload_dataset(g, ds)
    g.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
    g.GaussianReconstruction()
    before = g.debugv_PSF_sums()
    far = ds.trans.copy(); far[:, 3] += 500.0            # translate every slice 500 mm away
    far_inv = np.stack([np.linalg.inv(t.reshape(4, 4)).ravel() for t in far]).astype(np.float32)
    g.SetSliceMatrices(far, far_inv, ds.i2w, ds.w2i, ds.i2w, ds.w2i, ds.recon_i2w, ds.recon_w2i)
    assert np.all(g.GaussianReconstruction() == 0)
    np.testing.assert_array_equal(g.debugv_PSF_sums(), before)
    assert np.all(g.syncCPU() == 0)


def test_error_reporting():
    from fetalreconstruction_b200.reconstruction import SVRError
    g = _gpu()
    with pytest.raises(SVRError, match="not initialised"):
        g.GaussianReconstruction()
    g.InitReconstructionVolume((8, 8, 8), (1, 1, 1))
    with pytest.raises(SVRError, match="differs"):
        g.setMask((4, 4, 4), (1, 1, 1), np.ones(64, np.float32))


# ---- the paired scatter (svr_psf.cu: scatter_pair) on the geometries that exercise each of its paths ------------------
@pytest.mark.parametrize("size,inplane,voxel,vol,why", [
    (35, 1.0, 1.0, 40, "odd Nx: every slice row ends in a lone pixel (partner switched off)"),
    (20, 2.6, 1.0, 44, "pixels 2.6x coarser than the voxels: centres > 2 voxels apart in x -> one-pixel fallback"),
    (30, 0.55, 1.0, 28, "pixels finer than the voxels: both pixels of a pair round to the same voxel (d = 0)"),
    (36, 1.1, 1.0, 24, "volume smaller than the slices: most supports touch the volume faces (non-interior path)"),
])
def test_paired_scatter_paths(size, inplane, voxel, vol, why):
    cfg = small_config(seed=17, vol=vol, n_stacks=3, slices=5, size=size, voxel=voxel, inplane=inplane, spacing=2.0)
    ds = make_dataset(cfg)
    res = {}
    for name, b in (("gpu", _gpu()), ("orc", OracleReconstruction())):
        upload_dataset(b, ds)
        b.UpdateScaleVector(np.ones(ds.S, np.float32), np.ones(ds.S, np.float32))
        b.InitializeEMValues()
        vn = b.GaussianReconstruction()
        recon0, volw, psf = b.syncCPU(), b.getVolWeights(), b.debugv_PSF_sums()
        b.SimulateSlices()
        sw = np.ones(ds.S, np.float32); sw[0] = 0.0; sw[2] = 0.4
        pos = ds.slices[ds.slices > 0]
        # voxel weights: GaussianReconstruction cleared them (cuda2.cu:2402-2411); without an E-step K3 would scatter zeros
        b.EStep(1.0 / (2.1 * float(pos.max()) - 1.9 * float(pos.min())), 2.0e4, 0.9)
        b.Superresolution(1, sw, False, 1.0, float(pos.min()), float(pos.max()), 150.0, 0.02 * 150.0 * 150.0)
        res[name] = dict(vn=vn, recon0=recon0, volw=volw, psf=psf, addon=b.debugAddon(), cmap=b.debugConfidenceMap(), recon1=b.syncCPU())
    g, o = res["gpu"], res["orc"]
    assert np.array_equal(g["vn"], o["vn"]), why
    check_volume(g["recon0"], o["recon0"], "K1 volume: " + why)
    check_volume(g["volw"], o["volw"], "K1 volume weights: " + why)
    check_pixels(g["psf"], o["psf"], "v_PSF_sums: " + why)
    assert np.count_nonzero(o["addon"]) > 100, "K3 must scatter something: " + why
    check_volume(g["addon"], o["addon"], "K3 addon: " + why, rms=1e-3, mx=8e-2)
    check_volume(g["cmap"], o["cmap"], "K3 confidence map: " + why)
    check_volume(g["recon1"], o["recon1"], "volume after one SR step: " + why)
