"""CPU: the committed C2 set-up fixture (tests/golden/c2_setup.npz = BASELINE.json configs[1] after the reference's set-up
pipeline, see tests/golden/make_c2_setup.py) is self-consistent with the host rules the CLI restates:
MaskSlices (irtkReconstructionGPU.cc:1940-1988) and the slice packing of SyncGPU (irtkReconstructionGPU.cc:269-314)."""
import numpy as np

import c2_live


def test_c2_fixture_shapes_and_packing():
    ds = c2_live.load_setup()
    assert ds.S == 317 and ds.slices.shape == (317, 95, 99) and ds.mask.shape == (120, 105, 88)
    assert ds.cfg.vol_voxel == 1.0 and len(ds.stack_factor) == 4 and set(np.unique(ds.stack_index)) == {0, 1, 2, 3}
    # slices are top-left aligned in a cube pre-filled with -1
    for k in (0, 100, 316):
        sx, sy = ds.sizes[k]
        assert np.all(ds.slices[k, sy:, :] == -1) and np.all(ds.slices[k, :, sx:] == -1)
    valid = ds.slices != -1
    assert 0.05 < valid.mean() < 0.6 and np.all(ds.slices[valid] >= 0.01)
    # Tinv really is the inverse; I2W / W2I likewise
    for k in (3, 200):
        np.testing.assert_allclose(ds.trans[k].reshape(4, 4) @ ds.trans_inv[k].reshape(4, 4), np.eye(4), atol=1e-4)
        np.testing.assert_allclose(ds.i2w[k].reshape(4, 4) @ ds.w2i[k].reshape(4, 4), np.eye(4), atol=1e-3)


def test_c2_fixture_obeys_mask_slices_rule():
    """Every pixel that is not padding maps (rounded) onto a mask voxel != 0 (MaskSlices)."""
    ds = c2_live.load_setup()
    rw2i = ds.recon_w2i.reshape(4, 4).astype(np.float64)
    vz, vy, vx = ds.mask.shape
    bad = 0
    total = 0
    for k in range(0, ds.S, 7):
        ys, xs = np.nonzero(ds.slices[k] != -1)
        if not len(xs):
            continue
        p = np.stack([xs, ys, np.zeros_like(xs), np.ones_like(xs)]).astype(np.float64)
        w = rw2i @ (ds.trans[k].reshape(4, 4).astype(np.float64) @ (ds.i2w[k].reshape(4, 4).astype(np.float64) @ p))
        v = np.round(w[:3]).astype(int)
        inb = (v[0] >= 0) & (v[0] < vx) & (v[1] >= 0) & (v[1] < vy) & (v[2] >= 0) & (v[2] < vz)
        m = np.zeros(len(xs), bool)
        m[inb] = ds.mask[v[2][inb], v[1][inb], v[0][inb]] != 0
        bad += int((~m).sum()); total += len(xs)
    assert total > 10000 and bad <= 1e-3 * total, (bad, total)      # pixels exactly on a rounding boundary may differ (float vs double)
