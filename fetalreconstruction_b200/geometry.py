"""Host-side image geometry, restating the IRTK substrate rules the hot path depends on.

All matrices are 4x4 float64 numpy arrays (row-major); they are cast to float32 at the C-ABI
boundary exactly where the reference casts irtkMatrix -> Matrix4 (irtkReconstructionGPU.cc:330-342).

Reference rules:
  image -> world   IRTKSimple2/image++/src/irtkBaseImage.cc:79-112
  world -> image   IRTKSimple2/image++/src/irtkBaseImage.cc:114-147
  rigid params     IRTKSimple2/packages/transformation/src/irtkRigidTransformation.cc:26-149
  GetRegion        IRTKSimple2/image++/src/irtkGenericImage.cc:570-625
  slice creation   source/reconstructionGPU2/irtkReconstructionGPU.cc:1814-1850
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class ImageAttributes:
    """irtkImageAttributes: grid size, voxel size, centre-of-image origin, axis directions."""

    x: int
    y: int
    z: int
    dx: float = 1.0
    dy: float = 1.0
    dz: float = 1.0
    origin: np.ndarray = field(default_factory=lambda: np.zeros(3))
    xaxis: np.ndarray = field(default_factory=lambda: np.array([1.0, 0.0, 0.0]))
    yaxis: np.ndarray = field(default_factory=lambda: np.array([0.0, 1.0, 0.0]))
    zaxis: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 1.0]))

    def image_to_world(self) -> np.ndarray:
        """irtkBaseImage::GetImageToWorldMatrix (irtkBaseImage.cc:79-112)."""
        t1 = np.eye(4)
        t1[:3, 3] = [-(self.x - 1) / 2.0, -(self.y - 1) / 2.0, -(self.z - 1) / 2.0]
        sc = np.diag([self.dx, self.dy, self.dz, 1.0])
        rot = np.eye(4)
        rot[:3, 0] = self.xaxis
        rot[:3, 1] = self.yaxis
        rot[:3, 2] = self.zaxis
        t2 = np.eye(4)
        t2[:3, 3] = self.origin
        return t2 @ (rot @ (sc @ t1))

    def world_to_image(self) -> np.ndarray:
        """irtkBaseImage::GetWorldToImageMatrix (irtkBaseImage.cc:114-147)."""
        t1 = np.eye(4)
        t1[:3, 3] = -np.asarray(self.origin)
        rot = np.eye(4)
        rot[0, :3] = self.xaxis
        rot[1, :3] = self.yaxis
        rot[2, :3] = self.zaxis
        sc = np.diag([1.0 / self.dx, 1.0 / self.dy, 1.0 / self.dz, 1.0])
        t2 = np.eye(4)
        t2[:3, 3] = [(self.x - 1) / 2.0, (self.y - 1) / 2.0, (self.z - 1) / 2.0]
        return t2 @ (sc @ (rot @ t1))

    def slice_attributes(self, j: int, thickness: float) -> "ImageAttributes":
        """Slice j of a stack = GetRegion(0,0,j,X,Y,j+1) with dz := thickness
        (irtkReconstructionGPU.cc:1814-1850, irtkGenericImage.cc:570-625): axes and in-plane
        voxel size kept, origin moved to the world position of the slice centre."""
        centre = self.image_to_world() @ np.array([(self.x - 1) / 2.0, (self.y - 1) / 2.0, float(j), 1.0])
        return ImageAttributes(self.x, self.y, 1, self.dx, self.dy, thickness, centre[:3].copy(),
                               self.xaxis.copy(), self.yaxis.copy(), self.zaxis.copy())


def rigid_matrix(tx: float, ty: float, tz: float, rx: float, ry: float, rz: float) -> np.ndarray:
    """irtkRigidTransformation::Parameters2Matrix (irtkRigidTransformation.cc:55-95); degrees."""
    cx, cy, cz = np.cos(np.deg2rad([rx, ry, rz]))
    sx, sy, sz = np.sin(np.deg2rad([rx, ry, rz]))
    m = np.eye(4)
    m[0, :] = [cy * cz, cy * sz, -sy, tx]
    m[1, :] = [sx * sy * cz - cx * sz, sx * sy * sz + cx * cz, sx * cy, ty]
    m[2, :] = [cx * sy * cz + sx * sz, cx * sy * sz - sx * cz, cx * cy, tz]
    return m


def rigid_parameters(m: np.ndarray) -> np.ndarray:
    """irtkRigidTransformation::Matrix2Parameters (irtkRigidTransformation.cc:119-149)."""
    tol = 0.000001
    p = np.zeros(6)
    p[0:3] = m[0:3, 3]
    tmp = np.arcsin(-1 * m[0, 2])
    if abs(np.cos(tmp)) > tol:
        p[3] = np.arctan2(m[1, 2], m[2, 2])
        p[4] = tmp
        p[5] = np.arctan2(m[0, 1], m[0, 0])
    else:
        p[3] = np.arctan2(-1.0 * m[0, 2] * m[1, 0], -1.0 * m[0, 2] * m[2, 0])
        p[4] = tmp
        p[5] = 0
    p[3:] *= 180.0 / np.pi
    return p


def rotation_axes(rx: float, ry: float, rz: float) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Orthonormal stack axes from three Euler angles (degrees); used only to build synthetic stacks."""
    r = rigid_matrix(0, 0, 0, rx, ry, rz)[:3, :3]
    return r[:, 0].copy(), r[:, 1].copy(), r[:, 2].copy()


def psf_centre_offset(recon_voxel: float, psf_size: int = 128) -> np.ndarray:
    """d_PSFI2W * ((PSFsize-1)/2) evaluated in float32 like the kernels do
    (reconstruction_cuda2.cu:172; PSF image built at irtkReconstructionGPU.cc:1534-1551,
    PSF_SIZE = 128 at include/reconstruction_cuda2.cuh:56).  Exactly 0 in real arithmetic."""
    attr = ImageAttributes(psf_size, psf_size, psf_size, recon_voxel, recon_voxel, recon_voxel)
    m = attr.image_to_world().astype(np.float32)
    c = np.float32((psf_size - 1) * 0.5)
    out = np.zeros(3, np.float32)
    for r in range(3):
        out[r] = np.float32(np.float32(np.float32(m[r, 0] * c) + np.float32(m[r, 1] * c)) + np.float32(m[r, 2] * c)) + m[r, 3]
    return out
