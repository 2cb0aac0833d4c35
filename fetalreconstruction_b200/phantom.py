"""Seeded synthetic stacks for parity tests and bench.py (SURVEY.md section 8d, configs C3 and small
variants).  Produces exactly what irtkReconstruction::SyncGPU hands to the device library
(irtkReconstructionGPU.cc:249-328): one padded float slice cube (-1 = padding), per-slice
matrices and voxel sizes, the isotropic volume grid and its float mask.

The phantom is a sum of seeded ellipsoids evaluated analytically at the (motion-corrupted) world
position of every slice pixel; pixels whose centre maps outside the mask are set to -1 like
MaskSlices does (irtkReconstructionGPU.cc:1940-1988).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .geometry import ImageAttributes, rigid_matrix, rotation_axes, psf_centre_offset


@dataclass
class PhantomConfig:
    vol_size: tuple[int, int, int] = (256, 256, 256)
    vol_voxel: float = 0.75
    n_stacks: int = 8
    slices_per_stack: int = 128
    slice_size: tuple[int, int] = (256, 256)      # (Nx, Ny)
    inplane: float = 0.75
    spacing: float = 1.5
    thickness: float | None = None                # default 2*spacing (reconstruction.cc:423-434)
    motion_mm: float = 1.0
    motion_deg: float = 1.0
    noise: float = 10.0
    corrupt_fraction: float = 0.02
    mask_semi_axis: float = 0.42                  # fraction of the FOV
    seed: int = 20240601
    name: str = "C3"


def c3_config() -> PhantomConfig:
    """BASELINE.json configs[2]: 8 stacks x 128 slices of 256x256 into 256^3 at 0.75 mm."""
    return PhantomConfig()


def small_config(seed: int = 7, vol: int = 40, n_stacks: int = 3, slices: int = 10, size: int = 36,
                 voxel: float = 1.0, inplane: float = 1.1, spacing: float = 2.0) -> PhantomConfig:
    """A few-second CPU-oracle case with oblique stacks, motion and volume-edge overhang."""
    return PhantomConfig((vol, vol, vol), voxel, n_stacks, slices, (size, size - 4), inplane, spacing,
                         None, 0.7, 1.0, 5.0, 0.1, 0.42, seed, "small")


@dataclass
class Dataset:
    cfg: PhantomConfig
    vol_attr: ImageAttributes
    recon_i2w: np.ndarray          # float32 [16]
    recon_w2i: np.ndarray
    mask: np.ndarray               # float32 [vz, vy, vx]
    truth: np.ndarray              # float32 [vz, vy, vx]
    slices: np.ndarray             # float32 [S, Ny, Nx]
    i2w: np.ndarray                # float32 [S, 16]
    w2i: np.ndarray
    trans: np.ndarray              # slice -> volume (world) transform used for reconstruction
    trans_inv: np.ndarray
    dims: np.ndarray               # float32 [S, 3]
    stack_index: np.ndarray        # int32 [S]
    psf_c: np.ndarray              # float32 [3]
    min_intensity: float
    max_intensity: float
    slice_attrs: list | None = None    # ImageAttributes of every slice (for the registration front-end)
    true_trans: np.ndarray | None = None   # the motion each slice was acquired with [S,16]
    stack_attrs: list | None = None    # ImageAttributes of every stack (PVR patch enumeration)

    @property
    def S(self) -> int:
        return self.slices.shape[0]


_STACK_ANGLES = [(0, 0, 0), (90, 0, 0), (0, 90, 0)]


def _ellipsoids(rng: np.random.Generator, fov: np.ndarray, n: int = 10):
    centres = (rng.uniform(-0.22, 0.22, (n, 3)) * fov).astype(np.float32)
    radii = (rng.uniform(0.06, 0.3, (n, 3)) * fov).astype(np.float32)
    amps = rng.uniform(0.2, 1.0, n).astype(np.float32)
    centres[0] = 0
    radii[0] = 0.38 * fov
    amps[0] = 1.0
    return centres, radii, amps


def _phantom_at(world: torch.Tensor, centres, radii, amps) -> torch.Tensor:
    out = torch.zeros(world.shape[:-1], dtype=torch.float32, device=world.device)
    for c, r, a in zip(centres, radii, amps):
        c_t = torch.as_tensor(c, device=world.device)
        r_t = torch.as_tensor(r, device=world.device)
        q = (((world - c_t) / r_t) ** 2).sum(-1)
        out += float(a) * (q < 1.0)
    return out


def shard_count(n: int, rank: int, world: int) -> int:
    """Number of items j in [0, n) with j % world == rank."""
    return (n - rank + world - 1) // world if rank < n else 0


def make_dataset(cfg: PhantomConfig, device: str = "cpu", perfect_registration: bool = True,
                 stacks: range | None = None, shard: tuple[int, int] | None = None) -> Dataset:
    """`stacks` restricts generation to a contiguous range of stacks (a rank's shard); every stack is
    seeded on its own, so the slices are identical however the stacks are sharded.
    `shard = (rank, world)` instead gives the rank every world-th slice of EVERY stack (j % world == rank): every rank
    then sees the same mix of orientations AND of positions along the stack -- the per-slice cost of the PSF kernels
    depends on the orientation, and the number of valid pixels on how close a slice is to the end of the stack -- so
    the ranks are balanced by construction.  Every slice is seeded by its own (stack, index) either way."""
    rng = np.random.default_rng(cfg.seed)
    vx, vy, vz = cfg.vol_size
    vol_attr = ImageAttributes(vx, vy, vz, cfg.vol_voxel, cfg.vol_voxel, cfg.vol_voxel)
    ri2w, rw2i = vol_attr.image_to_world(), vol_attr.world_to_image()
    fov = np.array([vx, vy, vz], np.float64) * cfg.vol_voxel
    centres, radii, amps = _ellipsoids(rng, fov)
    dev = torch.device(device)

    # volume-grid world coordinates -> truth + mask
    gz, gy, gx = torch.meshgrid(torch.arange(vz, device=dev), torch.arange(vy, device=dev),
                                torch.arange(vx, device=dev), indexing="ij")
    gi = torch.stack([gx, gy, gz], -1).to(torch.float32)
    m = torch.as_tensor(ri2w, dtype=torch.float32, device=dev)
    gw = gi @ m[:3, :3].T + m[:3, 3]
    truth = _phantom_at(gw, centres, radii, amps)
    semi = torch.as_tensor(cfg.mask_semi_axis * fov, dtype=torch.float32, device=dev)
    mask = (((gw / semi) ** 2).sum(-1) < 1.0).to(torch.float32)
    inmask_mean = float((truth * mask).sum() / mask.sum().clamp(min=1))
    gain = 700.0 / max(inmask_mean, 1e-6)              # --average 700 (reconstruction.cc:175)
    truth = truth * gain
    del gz, gy, gx, gi, gw

    Nx, Ny = cfg.slice_size
    thickness = cfg.thickness if cfg.thickness is not None else 2.0 * cfg.spacing
    stacks = range(cfg.n_stacks) if stacks is None else stacks
    s_rank, s_world = shard if shard is not None else (0, 1)
    S = len(stacks) * shard_count(cfg.slices_per_stack, s_rank, s_world)
    slices = np.empty((S, Ny, Nx), np.float32)
    i2w = np.empty((S, 16), np.float32)
    w2i = np.empty((S, 16), np.float32)
    trans = np.empty((S, 16), np.float32)
    trans_inv = np.empty((S, 16), np.float32)
    dims = np.empty((S, 3), np.float32)
    stack_index = np.empty(S, np.int32)
    slice_attrs = []
    stack_attrs = []
    true_trans = np.empty((S, 16), np.float32)

    py, px = torch.meshgrid(torch.arange(Ny, device=dev), torch.arange(Nx, device=dev), indexing="ij")
    pix = torch.stack([px, py, torch.zeros_like(px)], -1).to(torch.float32)
    rw2i_t = torch.as_tensor(rw2i, dtype=torch.float32, device=dev)
    sizes = torch.tensor([vx, vy, vz], device=dev)

    k = 0
    for st in stacks:
        srng = np.random.default_rng(cfg.seed + st)
        if st < len(_STACK_ANGLES):
            ang = _STACK_ANGLES[st]
        else:
            ang = tuple(srng.uniform(-30, 30, 3))
        xa, ya, za = rotation_axes(*ang)
        stack_attr = ImageAttributes(Nx, Ny, cfg.slices_per_stack, cfg.inplane, cfg.inplane, cfg.spacing,
                                     np.zeros(3), xa, ya, za)
        stack_attrs.append(stack_attr)
        mrng = np.random.default_rng(cfg.seed + 1000 + st)
        for j in range(cfg.slices_per_stack):
            sa = stack_attr.slice_attributes(j, thickness)
            A, Ainv = sa.image_to_world(), sa.world_to_image()
            tpar = np.concatenate([mrng.normal(0, cfg.motion_mm, 3), mrng.normal(0, cfg.motion_deg, 3)])
            corrupt = srng.uniform() < cfg.corrupt_fraction
            if j % s_world != s_rank:                  # the random streams advance for every slice of the stack
                continue
            T_true = rigid_matrix(*tpar)
            T_used = T_true if perfect_registration else np.eye(4)
            # acquisition: sample the phantom where the (moved) slice really was
            M = torch.as_tensor(T_true @ A, dtype=torch.float32, device=dev)
            world = pix @ M[:3, :3].T + M[:3, 3]
            val = _phantom_at(world, centres, radii, amps) * gain
            if cfg.noise > 0:
                g = torch.Generator(device="cpu").manual_seed(cfg.seed * 131 + st * cfg.slices_per_stack + j)
                val = val + cfg.noise * torch.randn(val.shape, generator=g).to(dev)
            if corrupt:
                val = val * 0.3
            # MaskSlices: centre maps (rounded) outside the mask / onto mask 0 -> -1; <0.01 -> -1
            Mu = torch.as_tensor(T_used @ A, dtype=torch.float32, device=dev)
            wu = pix @ Mu[:3, :3].T + Mu[:3, 3]
            vi = torch.round(wu @ rw2i_t[:3, :3].T + rw2i_t[:3, 3]).to(torch.long)
            inb = ((vi >= 0) & (vi < sizes)).all(-1)
            vic = torch.minimum(torch.clamp(vi, min=0), sizes - 1)
            mv = mask[vic[..., 2], vic[..., 1], vic[..., 0]]
            keep = inb & (mv != 0) & (val >= 0.01)
            val = torch.where(keep, val, torch.full_like(val, -1.0))
            slices[k] = val.cpu().numpy()
            i2w[k] = A.astype(np.float32).ravel()
            w2i[k] = Ainv.astype(np.float32).ravel()
            trans[k] = T_used.astype(np.float32).ravel()
            trans_inv[k] = np.linalg.inv(T_used).astype(np.float32).ravel()
            dims[k] = (cfg.inplane, cfg.inplane, thickness)
            stack_index[k] = st
            slice_attrs.append(sa)
            true_trans[k] = T_true.astype(np.float32).ravel()
            k += 1

    pos = slices[slices > 0]
    return Dataset(cfg, vol_attr, ri2w.astype(np.float32).ravel(), rw2i.astype(np.float32).ravel(),
                   mask.cpu().numpy(), truth.cpu().numpy(), slices, i2w, w2i, trans, trans_inv, dims,
                   stack_index, psf_centre_offset(cfg.vol_voxel),
                   float(pos.min()) if pos.size else 0.0, float(pos.max()) if pos.size else 0.0, slice_attrs, true_trans, stack_attrs)
