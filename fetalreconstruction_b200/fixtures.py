"""Loader of the committed C2 set-up fixture (BASELINE.json configs[1]: the reference's bundled 4-stack 3T data at
--resolution 1.0 after the set-up pipeline; tests/golden/c2_setup.npz, written by tests/golden/make_c2_setup.py from
`SVRreconstructionGPU --dump_setup`).  Used by the C2 parity test and by `bench.py --workload C2`."""
import os
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURE = os.path.join(ROOT, "tests", "golden", "c2_setup.npz")


def load_c2_setup(path=FIXTURE):
    z = np.load(path)
    idx = dict(line.split() for line in str(z["index"]).splitlines() if line.strip())
    S, Nx, Ny = int(idx["S"]), int(idx["Nx"]), int(idx["Ny"])
    vx, vy, vz, voxel = int(idx["vx"]), int(idx["vy"]), int(idx["vz"]), float(idx["voxel"])
    from .geometry import ImageAttributes
    attrs = z["slice_attrs"].reshape(S, 18)
    slice_attrs = [ImageAttributes(int(a[0]), int(a[1]), int(a[2]), a[3], a[4], a[5], a[6:9].copy(), a[9:12].copy(), a[12:15].copy(),
                                   a[15:18].copy()) for a in attrs]
    return SimpleNamespace(S=S, cfg=SimpleNamespace(vol_voxel=voxel, vol_size=(vx, vy, vz), name="C2"),
                           slices=z["slices"].reshape(S, Ny, Nx), mask=z["mask"].reshape(vz, vy, vx), dims=z["dims"].reshape(S, 3),
                           trans=z["T"].reshape(S, 16), trans_inv=z["Tinv"].reshape(S, 16), i2w=z["I2W"].reshape(S, 16),
                           w2i=z["W2I"].reshape(S, 16), recon_i2w=z["recon_i2w"], recon_w2i=z["recon_w2i"],
                           stack_index=z["stack_index"], stack_factor=z["stack_factor"], slice_attrs=slice_attrs,
                           sizes=z["sizes"].reshape(S, 2))


def shard_dataset(ds, rank, world):
    """Every world-th slice (global index i % world == rank) of a dataset namespace: the rank's share, balanced over stacks and
    positions along the stacks like phantom.make_dataset(shard=...)."""
    if world == 1:
        return ds
    idx = np.arange(ds.S)[rank::world]
    sub = SimpleNamespace(**vars(ds))
    for name in ("slices", "dims", "trans", "trans_inv", "i2w", "w2i", "stack_index", "sizes"):
        if hasattr(ds, name):
            setattr(sub, name, np.ascontiguousarray(getattr(ds, name)[idx]))
    sub.slice_attrs = [ds.slice_attrs[i] for i in idx]
    sub.S = len(idx)
    return sub
