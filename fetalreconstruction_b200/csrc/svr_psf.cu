// svr_psf.cu -- the PSF forward / adjoint projection kernels (K1, K2, K3) for sm_100a.
//
// Work decomposition (all three kernels): one thread per VALID slice pixel (pixels != -1 are
// compacted once in svr_fill_slices, so warps are dense: the reference launches a thread per
// padded pixel and returns early).  A thread walks the 16^3 PSF support with the per-tap
// position factored as  e + ox*b0 + oy*b1 + oz*b2  in pre-scaled "PSF units" (3 FFMA per tap
// instead of the reference's 3x4 mat-vec + scalings) and evaluates sinc^2 * gauss with three MUFU
// ops (rsqrt, sin, ex2).  Pixels whose support lies inside the volume (all of them unless the
// mask touches the volume faces) take a path without clamps or bounds checks where a row's 16 taps
// are addressed with immediate offsets from one row pointer.
//
// Scatter (K1 pass 2, K3): the per-SM rate of global reductions (~1 lane-op/clk/SM) is the
// bottleneck of a tap-per-RED scatter, so a row's 16 contributions are kept in registers and
// flushed as nine 128-bit vector reductions (red.global.add.v4.f32 on the interleaved
// {numerator, denominator} accumulator = two voxels per lane-op).  The mask test of the reference
// (`mask[v] != 0` per tap) is applied once per voxel afterwards (equalize / regulariser prep zero
// the masked-out voxels), which is exact: a masked voxel's sum is discarded either way.
//
// Reference kernels restated here: reconstruction_cuda2.cu:176-295 (K1), 298-404 (K2), 408-522 (K3).
#include "svr_context.h"

// Resident CTAs per SM the PSF kernels are compiled for (register budget = 65536 / (128 * SVR_MINB)).
#ifndef SVR_MINB
#define SVR_MINB 5
#endif

// ---------------------------------------------------------------------------------------------
// Per-slice geometry (runs when matrices or voxel sizes change; S threads).
// comb = (W2I * Tinv) * reconI2W in float, product order of reconstruction_cuda2.cu:223.
__device__ __forceinline__ void mat44_mul(const float* A, const float* B, float* C)
{   // operator*(Matrix4, Matrix4), recon_volumeHelper.cuh:134-145
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            C[4 * i + j] = A[4 * i + 0] * B[0 + j] + A[4 * i + 1] * B[4 + j] + A[4 * i + 2] * B[8 + j] + A[4 * i + 3] * B[12 + j];
}

struct Mat16 { float m[16]; };

__global__ void build_geom_kernel(int S, const float* __restrict__ T, const float* __restrict__ Tinv,
                                  const float* __restrict__ I2W, const float* __restrict__ W2I,
                                  const float* __restrict__ dims, Mat16 reconI2W, SliceGeom* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    float t1[16], comb[16];
    mat44_mul(W2I + 16 * k, Tinv + 16 * k, t1);
    mat44_mul(t1, reconI2W.m, comb);
    SliceGeom g;
    for (int i = 0; i < 12; ++i) { g.i2w[i] = I2W[16 * k + i]; g.t[i] = T[16 * k + i]; g.a[i] = comb[i]; }
    const float dx = dims[3 * k + 0], dy = dims[3 * k + 1], dz = dims[3 * k + 2];
    g.dimx = dx; g.dimy = dy; g.dimz = dz;
    const float sigmaz = dz / 2.3548f;
    g.kx = dx / 2.3548f * 3.14159265359f;           // sPos.x * dim.x / 2.3548, then R = pi * x (cuda2.cu:125-128)
    g.ky = dy / 2.3548f * 3.14159265359f;
    g.kz = 0.84932180028801907f / sigmaz;           // sqrt(log2(e) / 2) / sigma_z: exp(-z^2/(2 s^2)) = 2^(-(kz z)^2)
    for (int j = 0; j < 3; ++j) {
        g.bx[j] = comb[0 + j] * dx * g.kx;
        g.by[j] = comb[4 + j] * dy * g.ky;
        g.bz[j] = comb[8 + j] * dz * g.kz;
    }
    g.two_b = 2.0f * g.bz[0];
    g.bb = g.bz[0] * g.bz[0];
    g.kappa = exp2f(-2.0f * g.bb);
    g.recur = (fabsf(g.bz[0]) <= 1.0f && fabsf(g.bz[0]) + fabsf(g.bz[1]) + fabsf(g.bz[2]) <= 4.0f) ? 1 : 0;
    out[k] = g;
}

int svr_launch_build_geom(svr_context* c)
{
    if (!c->have_mats || !c->have_dims || c->S == 0) return 0;
    Mat16 ri2w;
    for (int i = 0; i < 16; ++i) ri2w.m[i] = c->recon_i2w[i];
    const size_t n = (size_t)c->S * 16;
    build_geom_kernel<<<divup_i(c->S, 128), 128, 0, c->stream>>>(c->S, c->mats, c->mats + n, c->mats + 2 * n, c->mats + 3 * n,
                                                                 c->dims, ri2w, c->geom);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Flush one interior x-row of contributions p[0..15] (voxels v0 .. v0+15) scaled by (a, c) as paired
// 128-bit reductions.  The pairs must be 16-byte aligned, so an odd v0 shifts the row by one voxel
// (17 selects); the accumulator is allocated with 2 voxels of slack for the zero half of the last pair.
__device__ __forceinline__ void red_row_paired(float2* __restrict__ acc2, int v0, const float (&p)[SVR_PSF_SUPPORT],
                                               float a, float c)
{
    const bool odd = (v0 & 1) != 0;
    float4* base = reinterpret_cast<float4*>(acc2 + (v0 - (odd ? 1 : 0)));
    float q[SVR_PSF_SUPPORT + 2];
    q[0] = odd ? 0.0f : p[0];
#pragma unroll
    for (int j = 1; j < SVR_PSF_SUPPORT; ++j) q[j] = odd ? p[j - 1] : p[j];
    q[SVR_PSF_SUPPORT] = odd ? p[SVR_PSF_SUPPORT - 1] : 0.0f;
    q[SVR_PSF_SUPPORT + 1] = 0.0f;
#pragma unroll
    for (int m = 0; m < SVR_PSF_SUPPORT / 2 + 1; ++m) {
        const float u = q[2 * m], w = q[2 * m + 1];
        if (u + w > 0.0f)                                  // psf >= 0: skip all-zero pairs (and NaNs)
            atomicAdd(base + m, make_float4(u * a, u * c, w * a, w * c));
    }
}

// ---------------------------------------------------------------------------------------------
// K1: gaussianReconstructionKernel3D_tex (reconstruction_cuda2.cu:176-295).
// Pass 1: sume = sum of accepted in-volume taps (mask ignored, quirk Q3); stored only if > 0.5.
// Pass 2: scatter psf/sume * {s*scale, 1}; flag the pixel if any accepted tap landed on a masked voxel.
__global__ void __launch_bounds__(128, SVR_MINB)
gaussian_scatter_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int Nx, int P,
                        const float* __restrict__ slices, const float* __restrict__ scales,
                        const SliceGeom* __restrict__ geom, VolGeom vg, const unsigned char* __restrict__ mask,
                        float2* __restrict__ acc2, float* __restrict__ psf_sums, unsigned char* __restrict__ voxel_flag,
                        int* __restrict__ slice_count)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid) return;
    const uint32_t idx = valid_idx[t];
    const int k = idx / P, pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    const SliceGeom& g = geom[k];
    const float s = slices[idx] * scales[k];
    const PixelSetup ps = pixel_setup(g, vg, x, y);

    float sume = 0.f;
    psf_rows_dispatch(g, vg, ps, [&](int, float psf, bool, int) { sume += psf; }, [](int) {});
    if (!(sume > 0.5f)) return;
    psf_sums[idx] = sume;

    const float inv = 1.0f / sume;
    const float sv = s * inv;
    bool any = false;
    if (ps.interior) {
        float p[SVR_PSF_SUPPORT];
        auto tap = [&](int i, float psf, bool ok, int v) { p[i] = psf; if (ok && mask[v]) any = true; };
        auto row = [&](int v0) { red_row_paired(acc2, v0, p, sv, inv); };
        if (g.recur) psf_rows<true, true>(g, vg, ps, tap, row);
        else psf_rows<true, false>(g, vg, ps, tap, row);
    } else {
        psf_rows_dispatch(g, vg, ps,
            [&](int, float psf, bool ok, int v) {
                if (ok) {
                    atomicAdd(&acc2[v], make_float2(psf * sv, psf * inv));
                    if (mask[v]) any = true;
                }
            },
            [](int) {});
    }
    if (any) {
        voxel_flag[idx] = 1;
        atomicAdd(&slice_count[k], 1);
    }
}

int svr_launch_gaussian_scatter(svr_context* c)
{
    if (c->n_valid == 0) return 0;
    ProfScope prof(c, 0);
    gaussian_scatter_kernel<<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
        c->n_valid, c->valid_idx, c->Nx, c->Nx * c->Ny, c->slices, c->scales, c->geom, c->vg, c->mask_u8, c->acc2,
        c->psf_sums, c->voxel_flag, c->slice_count);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K2: simulateSlicesKernel3D_tex (reconstruction_cuda2.cu:298-404).
// pack2[v] = {recon[v]*m, m} with m = (mask != 0), so a tap is one predicated 64-bit load + 2 FFMA.
__global__ void __launch_bounds__(128, SVR_MINB)
simulate_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int Nx, int P,
                const SliceGeom* __restrict__ geom, VolGeom vg, const float2* __restrict__ pack2,
                const float* __restrict__ psf_sums, float* __restrict__ simslices, float* __restrict__ simweights,
                unsigned char* __restrict__ siminside, int* __restrict__ slice_inside)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid) return;
    const uint32_t idx = valid_idx[t];
    const float sume = psf_sums[idx];
    if (sume == 0.0f) return;
    const int k = idx / P, pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    const SliceGeom& g = geom[k];
    const PixelSetup ps = pixel_setup(g, vg, x, y);

    float sim = 0.f, wsum = 0.f;
    auto tap = [&](int, float psf, bool ok, int v) {
        if (ok) {
            const float2 pm = __ldg(&pack2[v]);
            sim = fmaf(psf, pm.x, sim);
            wsum = fmaf(psf, pm.y, wsum);
        }
    };
    psf_rows_dispatch(g, vg, ps, tap, [](int) {});
    const float weight = wsum / sume;
    if (weight > 0.f) {
        simslices[idx] = sim / wsum;              // (sum psf/sume * x) / (sum psf/sume)
        simweights[idx] = weight;
        siminside[idx] = 1;
        slice_inside[k] = 1;                       // benign race: every writer stores 1
    }
}

int svr_launch_simulate(svr_context* c)
{
    if (c->n_valid == 0) return 0;
    ProfScope prof(c, 1);
    simulate_kernel<<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx, c->Nx * c->Ny, c->geom,
                                                                     c->vg, c->pack2, c->psf_sums, c->simslices,
                                                                     c->simweights, c->siminside, c->slice_inside);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K3: SuperresolutionKernel3D_tex (reconstruction_cuda2.cu:408-522).
__global__ void __launch_bounds__(128, SVR_MINB)
superres_scatter_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int Nx, int P,
                        const float* __restrict__ slices, const float* __restrict__ weights,
                        const float* __restrict__ simslices, const float* __restrict__ slice_weights,
                        const float* __restrict__ scales, const SliceGeom* __restrict__ geom, VolGeom vg,
                        const float* __restrict__ psf_sums, float2* __restrict__ acc2)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid) return;
    const uint32_t idx = valid_idx[t];
    const float sume = psf_sums[idx];
    if (sume == 0.0f) return;
    const int k = idx / P, pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    const SliceGeom& g = geom[k];

    const float w = weights[idx];
    const float ss = simslices[idx];
    float sliceVal = slices[idx] * scales[k];
    sliceVal = (ss > 0.0f) ? (sliceVal - ss) : 0.0f;
    const float cw = w * slice_weights[k] / sume;      // psf/sume * w * slice_weight
    const float aw = cw * sliceVal;
    // A pixel with zero weight adds exact zeros everywhere: skip its 4096 taps.
    if (cw == 0.0f) return;
    const PixelSetup ps = pixel_setup(g, vg, x, y);
    if (ps.interior) {
        float p[SVR_PSF_SUPPORT];
        auto tap = [&](int i, float psf, bool, int) { p[i] = psf; };
        auto row = [&](int v0) { red_row_paired(acc2, v0, p, aw, cw); };
        if (g.recur) psf_rows<true, true>(g, vg, ps, tap, row);
        else psf_rows<true, false>(g, vg, ps, tap, row);
    } else {
        psf_rows_dispatch(g, vg, ps,
            [&](int, float psf, bool ok, int v) { if (ok) atomicAdd(&acc2[v], make_float2(psf * aw, psf * cw)); },
            [](int) {});
    }
}

int svr_launch_superres_scatter(svr_context* c)
{
    if (c->n_valid == 0) return 0;
    ProfScope prof(c, 2);
    superres_scatter_kernel<<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
        c->n_valid, c->valid_idx, c->Nx, c->Nx * c->Ny, c->slices, c->weights, c->simslices, c->slice_weights, c->scales,
        c->geom, c->vg, c->psf_sums, c->acc2);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Volume-side elementwise helpers.
__global__ void pack_volume_kernel(size_t V, const float* __restrict__ recon, const unsigned char* __restrict__ mask,
                                   float2* __restrict__ pack2)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        const bool m = mask[v] != 0;
        pack2[v] = m ? make_float2(recon[v], 1.0f) : make_float2(0.f, 0.f);
    }
}

int svr_launch_pack_volume(svr_context* c)
{
    pack_volume_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->recon, c->mask_u8, c->pack2);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// equalizeVol (reconstruction_cuda2.cu:2312-2327) on the interleaved accumulator.
__global__ void equalize_kernel(size_t V, const float2* __restrict__ acc2, const unsigned char* __restrict__ mask,
                                float* __restrict__ recon, float* __restrict__ volw)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        float2 a = acc2[v];
        if (!mask[v]) a = make_float2(0.f, 0.f);           // the per-tap mask test of cuda2.cu:275-276, applied per voxel
        volw[v] = a.y;
        recon[v] = (a.y != 0.f) ? a.x / a.y : a.x;
    }
}

int svr_launch_equalize(svr_context* c)
{
    equalize_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->acc2, c->mask_u8, c->recon, c->volw);
    SVR_KERNEL_CHECK(c);
    return 0;
}

__global__ void deinterleave_kernel(size_t V, const float2* __restrict__ src, float* __restrict__ dst, int comp)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x)
        dst[v] = comp ? src[v].y : src[v].x;
}

int svr_launch_deinterleave(svr_context* c, const float2* src, float* dst, int component)
{
    deinterleave_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, src, dst, component);
    SVR_KERNEL_CHECK(c);
    return 0;
}
