"""Builder of a synthetic PVR case (stacks -> 64x64 patches on a stride grid) from a phantom configuration: shared by the PVR
tests (tests/pvr_case.py re-exports it) and bench.py --workload C4 (BASELINE.json configs[3])."""
import numpy as np

from .geometry import ImageAttributes
from .phantom import make_dataset, small_config
from .pvr import generate_2d_patches


def make_pvr_case(seed=3, vol=40, n_stacks=2, slices=6, size=40, inplane=1.0, spacing=2.5, pbb=(16, 16), stride=(8, 8), cfg=None,
                  device="cpu"):
    if cfg is None:
        cfg = small_config(seed=seed, vol=vol, n_stacks=n_stacks, slices=slices, size=size, inplane=inplane, spacing=spacing)
        cfg.noise = 2.0
        cfg.corrupt_fraction = 0.0
    else:
        n_stacks, slices, spacing = cfg.n_stacks, cfg.slices_per_stack, cfg.spacing
    ds = make_dataset(cfg, device=device)
    Ny, Nx = ds.slices.shape[1:]
    mask_i8 = (ds.mask > 0).astype(np.int8)
    case = dict(ds=ds, cfg=cfg, mask=mask_i8, pbb=pbb, stacks=[], stack_w2i=[], per_stack=[], attrs=[], cube=[], trans=[],
                stack_dims=[])
    for st in range(n_stacks):
        sl = ds.slices[st * slices:(st + 1) * slices]
        stack = np.where(sl == -1, 0.0, sl).astype(np.float32)            # PVR stacks carry 0 background
        sattr = ds.stack_attrs[st]
        attrs, cube = generate_2d_patches(stack, sattr, mask_i8, ds.vol_attr, pbb, stride, thickness=spacing)
        case["stacks"].append(stack)
        case["stack_w2i"].append(sattr.world_to_image().astype(np.float32).ravel())
        case["per_stack"].append(len(attrs))
        case["attrs"] += attrs
        case["cube"].append(cube)
        case["stack_dims"].append((sattr.dx, sattr.dy, sattr.dz))
        # every patch carries the transformation of its stack (identity here: the phantom's motion is per slice,
        # so use the true per-slice motion of the slice the patch was cut from)
        for a in attrs:
            z = int(round((sattr.world_to_image() @ np.append(a.image_to_world() @ np.array([0, 0, 0, 1.0]), [])[:4])[2]))
            z = min(max(z, 0), slices - 1)
            case["trans"].append(ds.true_trans[st * slices + z].astype(np.float64).reshape(4, 4))
    case["cube"] = np.concatenate(case["cube"]) if case["cube"] else np.zeros((0, pbb[1], pbb[0]), np.float32)
    n = len(case["attrs"])
    case["i2w"] = np.stack([a.image_to_world().astype(np.float32).ravel() for a in case["attrs"]]) if n else np.zeros((0, 16), np.float32)
    case["w2i"] = np.stack([a.world_to_image().astype(np.float32).ravel() for a in case["attrs"]]) if n else np.zeros((0, 16), np.float32)
    case["T"] = np.stack([t.astype(np.float32).ravel() for t in case["trans"]]) if n else np.zeros((0, 16), np.float32)
    case["Tinv"] = np.stack([np.linalg.inv(t).astype(np.float32).ravel() for t in case["trans"]]) if n else np.zeros((0, 16), np.float32)
    return case


def setup_backend(b, case, device_patch_init=True, spx=None):
    ds, cfg = case["ds"], case["cfg"]
    vx, vy, vz = cfg.vol_size
    b.recon_init((vx, vy, vz), (cfg.vol_voxel,) * 3, ds.recon_w2i, ds.recon_i2w)
    b.recon_setMask(case["mask"].ravel())
    b.patches_init(case["pbb"][0], case["pbb"][1], case["per_stack"], case["stack_dims"])
    b.patches_set_matrices(case["i2w"], case["w2i"], case["T"], case["Tinv"])
    psf = ImageAttributes(128, 128, 128, cfg.vol_voxel, cfg.vol_voxel, cfg.vol_voxel)
    b.set_psf((128, 128, 128), psf.image_to_world().astype(np.float32).ravel(), 1.0)
    if spx is not None:
        b.patches_set_spx(spx, True)
    if device_patch_init:
        for st, stack in enumerate(case["stacks"]):
            b.initPatchBasedRecon_gpu(st, stack, case["stack_w2i"][st])
    else:
        b.patches_copyFromHost(case["cube"])
    return b


def shard_case(case, rank, world):
    """The rank's share of a PVR case: every world-th patch of every stack (balanced by construction, as bench.py shards
    slices).  Returns (sub-case, global_index of its patches in the stack-major order of all patches)."""
    idx, per_stack, o = [], [], 0
    for n in case["per_stack"]:
        mine = [o + j for j in range(n) if j % world == rank]
        idx += mine
        per_stack.append(len(mine))
        o += n
    idx = np.asarray(idx, np.int64)
    sub = dict(case)
    sub["per_stack"] = per_stack
    sub["attrs"] = [case["attrs"][i] for i in idx]
    sub["trans"] = [case["trans"][i] for i in idx]
    for k in ("cube", "i2w", "w2i", "T", "Tinv"):
        sub[k] = np.ascontiguousarray(case[k][idx])
    return sub, idx
