"""Host front-end of the GPU slice-to-volume registration (--useGPUReg): what
irtkReconstruction::PrepareRegistrationSlices and irtkReconstruction::SliceToVolumeRegistrationGPU
(source/reconstructionGPU2/irtkReconstructionGPU.cc:2105-2179, 2218-2288) do around the device library.

  * every slice is resampled to the isotropic volume resolution with the padding-aware trilinear
    interpolator irtkResamplingWithPadding (IRTKSimple2/image++/src/irtkResamplingWithPadding.cc:19-183;
    output grid of irtkResampling: new_n = int(n * old / new), origin and axes unchanged,
    IRTKSimple2/image++/src/irtkResampling.cc:92-131) and z-plane 0 of the result is packed, top-left
    aligned, into one float[S][H][W] cube pre-filled with -1;
  * for the registration call the origin of each resampled slice is moved to 0 and folded into the
    transform (m = T * mo), the device optimises m, and T = m * mo^-1 is read back.

`backend` is a Reconstruction (CUDA) or its oracle twin.  No arithmetic of the cost function lives here.
"""
from __future__ import annotations

import numpy as np

from .geometry import ImageAttributes


def resampled_attributes(attr: ImageAttributes, d: float) -> ImageAttributes:
    """irtkResamplingWithPadding::Initialize (irtkResamplingWithPadding.cc:203-262): the grid size is ROUNDED (irtkResampling's own
    Initialize truncates), a dimension below 1 stays one voxel of the old size; same origin / axes.  (Round 1 truncated here, which
    differs from the reference whenever thickness / voxel size is not an integer, e.g. 2.5 mm slices into a 1 mm volume.)"""
    rnd = lambda v: int(v + 0.5) if v > 0 else int(v - 0.5)
    n = [rnd(attr.x * attr.dx / d), rnd(attr.y * attr.dy / d), rnd(attr.z * attr.dz / d)]
    old = [attr.dx, attr.dy, attr.dz]
    size = [d if n[i] >= 1 else old[i] for i in range(3)]
    n = [max(v, 1) for v in n]
    return ImageAttributes(n[0], n[1], n[2], size[0], size[1], size[2], np.array(attr.origin, float), attr.xaxis.copy(), attr.yaxis.copy(),
                           attr.zaxis.copy())


def resample_plane0_with_padding(img: np.ndarray, attr: ImageAttributes, out_attr: ImageAttributes,
                                 padding: float = -1.0) -> np.ndarray:
    """z-plane 0 of irtkResamplingWithPadding::Run for a single-plane input image img[Ny][Nx]
    (irtkResamplingWithPadding.cc:36-183): neighbours equal to the padding value are dropped and the
    remaining weights renormalised; out-of-bounds neighbours count as not padded but add nothing; the
    result is padding when >= 4 of the 8 neighbours are padding or the weight sum is 0."""
    # two matrix-vector products in the reference's order (ImageToWorld of the output, then WorldToImage of the input): on
    # coinciding grids the coordinates are integers up to rounding and floor() has to fall as it does there
    a, b = out_attr.image_to_world(), attr.world_to_image()
    j, i = np.meshgrid(np.arange(out_attr.y, dtype=np.float64), np.arange(out_attr.x, dtype=np.float64), indexing="ij")
    k = np.zeros_like(i)
    wx = a[0, 0] * i + a[0, 1] * j + a[0, 2] * k + a[0, 3]
    wy = a[1, 0] * i + a[1, 1] * j + a[1, 2] * k + a[1, 3]
    wz = a[2, 0] * i + a[2, 1] * j + a[2, 2] * k + a[2, 3]
    x = b[0, 0] * wx + b[0, 1] * wy + b[0, 2] * wz + b[0, 3]
    y = b[1, 0] * wx + b[1, 1] * wy + b[1, 2] * wz + b[1, 3]
    z = b[2, 0] * wx + b[2, 1] * wy + b[2, 2] * wz + b[2, 3]
    u, v, w = np.floor(x).astype(int), np.floor(y).astype(int), np.floor(z).astype(int)
    dx, dy, dz = x - u, y - v, z - w
    val = np.zeros_like(x)
    wsum = np.zeros_like(x)
    pad = np.full(x.shape, 8, int)
    img64 = img.astype(np.float64)
    for du, wx in ((0, 1 - dx), (1, dx)):
        for dv, wy in ((0, 1 - dy), (1, dy)):
            for dw, wz in ((0, 1 - dz), (1, dz)):
                uu, vv, ww = u + du, v + dv, w + dw
                inb = (uu >= 0) & (uu < attr.x) & (vv >= 0) & (vv < attr.y) & (ww >= 0) & (ww < attr.z)
                g = img64[np.clip(vv, 0, attr.y - 1), np.clip(uu, 0, attr.x - 1)]
                good = inb & (g != padding)
                wt = wx * wy * wz
                val += np.where(good, g * wt, 0.0)
                wsum += np.where(good, wt, 0.0)
                pad -= (~inb | good).astype(int)
    out = np.where((pad < 4) & (wsum > 0), val / np.where(wsum > 0, wsum, 1.0), padding)
    return out.astype(np.float32)


class RegistrationFrontEnd:
    """PrepareRegistrationSlices + SliceToVolumeRegistrationGPU for the slices of one rank."""

    def __init__(self, backend, slices: np.ndarray, slice_attrs: list[ImageAttributes], recon_voxel: float,
                 slices_resident: bool = False):
        """`slices_resident`: the CUDA backend already holds exactly these slices (FillSlices); otherwise they are
        uploaded first (the reference resamples its host copies of the slices it uploaded, irtkReconstructionGPU.cc:1992-2059)."""
        self.b = backend
        self.d = float(recon_voxel)
        self.res_attrs = [resampled_attributes(a, self.d) for a in slice_attrs]
        S = len(slice_attrs)
        W = max([a.x for a in self.res_attrs], default=1)
        H = max([a.y for a in self.res_attrs], default=1)
        i2w = np.stack([ra.image_to_world().astype(np.float32).ravel() for ra in self.res_attrs]) if S else np.zeros((0, 16), np.float32)
        d0 = (self.d, self.d, self.d)
        backend.initRegStorageVolumes((W, H, S), d0)
        self._cube = None
        if hasattr(backend, "resampleRegSlices"):
            # the CUDA backend resamples the slices it already holds (FillSlices) on the device
            if not slices_resident and S:
                backend.FillSlices(np.asarray(slices).ravel())
            mo = np.stack([ra.image_to_world() for ra in self.res_attrs]) if S else np.zeros((0, 4, 4))
            mi = np.stack([a.world_to_image() for a in slice_attrs]) if S else np.zeros((0, 4, 4))
            backend.resampleRegSlices(mo, mi, [(a.x, a.y) for a in slice_attrs], [(ra.x, ra.y) for ra in self.res_attrs], i2w)
        else:
            # the CPU twins (oracle, reference adapter: test infrastructure) take the host-resampled cube
            self._cube = self.resample_on_host(slices, slice_attrs)
            backend.FillRegSlices(self._cube.ravel(), i2w)
        # origin reset (irtkReconstructionGPU.cc:2226-2250)
        self.mo = []
        ofs = np.zeros((S, 16), np.float32)
        for k, ra in enumerate(self.res_attrs):
            mo = np.eye(4)
            mo[:3, 3] = ra.origin
            self.mo.append(mo)
            z = ImageAttributes(ra.x, ra.y, ra.z, ra.dx, ra.dy, ra.dz, np.zeros(3), ra.xaxis, ra.yaxis, ra.zaxis)
            ofs[k] = z.image_to_world().astype(np.float32).ravel()
        self.ofs = ofs

    def resample_on_host(self, slices, slice_attrs) -> np.ndarray:
        S = len(slice_attrs)
        W = max([a.x for a in self.res_attrs], default=1)
        H = max([a.y for a in self.res_attrs], default=1)
        cube = np.full((S, H, W), -1.0, np.float32)
        for k, (a, ra) in enumerate(zip(slice_attrs, self.res_attrs)):
            cube[k, :ra.y, :ra.x] = resample_plane0_with_padding(slices[k, :a.y, :a.x], a, ra)
        return cube

    @property
    def cube(self) -> np.ndarray:
        """The resampled slices [S][H][W] (read back from the device when they were resampled there)."""
        if self._cube is None:
            self._cube = self.b.debugRegSlices(blurred=False).copy()
        return self._cube

    def pack_transforms(self, transformations: np.ndarray) -> np.ndarray:
        """_transf[i] = toMatrix4(T_i * mo_i)"""
        out = np.zeros((len(self.mo), 16), np.float32)
        for k, mo in enumerate(self.mo):
            out[k] = (np.asarray(transformations[k], np.float64).reshape(4, 4) @ mo).astype(np.float32).ravel()
        return out

    def unpack_transforms(self, transf: np.ndarray) -> np.ndarray:
        """T_i = fromMatrix4(_transf[i]) * mo_i^-1"""
        out = np.zeros((len(self.mo), 16), np.float64)
        for k, mo in enumerate(self.mo):
            out[k] = (np.asarray(transf[k], np.float64).reshape(4, 4) @ np.linalg.inv(mo)).ravel()
        return out

    def SliceToVolumeRegistrationGPU(self, transformations: np.ndarray) -> np.ndarray:
        """transformations [S,16] (slice -> volume, world) -> registered transformations [S,16] float64."""
        self.b.updateResampledSlicesI2W(self.ofs)
        self.b.prepareSliceToVolumeReg()
        t = self.b.registerSlicesToVolume(self.pack_transforms(transformations))
        return self.unpack_transforms(t)
