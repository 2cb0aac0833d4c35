// svr_psf_pixel.cuh -- per-pixel helpers shared by the PSF scatter kernels (svr_psf.cu: paired scatter;
// svr_window.cu: warp-window scatter): the inputs of one pixel of K1 pass 2 / K3 and the one-pixel scatter paths.
#pragma once
#include "svr_context.h"

// ---------------------------------------------------------------------------------------------
// Flush one interior x-row of contributions p[0..15] (voxels v0 .. v0+15) scaled by (a, c) as paired
// 128-bit reductions.  The pairs must be 16-byte aligned, so an odd v0 shifts the row by one voxel
// (17 selects); the accumulator is allocated with 2 voxels of slack for the zero half of the last pair.
template <int SUP>
__device__ __forceinline__ void red_row_paired(float2* __restrict__ acc2, int v0, const float (&p)[SUP], float a, float c)
{
    const bool odd = (v0 & 1) != 0;
    float4* base = reinterpret_cast<float4*>(acc2 + (v0 - (odd ? 1 : 0)));
    float q[SUP + 2];
    q[0] = odd ? 0.0f : p[0];
#pragma unroll
    for (int j = 1; j < SUP; ++j) q[j] = odd ? p[j - 1] : p[j];
    q[SUP] = odd ? p[SUP - 1] : 0.0f;
    q[SUP + 1] = 0.0f;
#pragma unroll
    for (int m = 0; m < SUP / 2 + 1; ++m) {
        const float u = q[2 * m], w = q[2 * m + 1];
        if (u + w > 0.0f)                                  // psf >= 0: skip all-zero pairs (and NaNs)
            atomicAdd(base + m, make_float4(u * a, u * c, w * a, w * c));
    }
}

// Pass 2 inputs of one pixel: false when pass 1 did not mark it.
template <class TR>
__device__ __forceinline__ bool gaussian_pixel(uint32_t idx, int Nx, int P, const float* __restrict__ slices,
                                               const float* __restrict__ scales, const SliceGeom* __restrict__ geom,
                                               const VolGeom& vg, const float* __restrict__ psf_sums,
                                               const unsigned char* __restrict__ voxel_flag, int& k, PixelSetup& ps, float& sv, float& inv)
{
    if (voxel_flag[idx] != 2) return false;
    k = idx / P;
    const int pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    ps = pixel_setup<TR>(geom[k], vg, x, y);
    inv = 1.0f / psf_sums[idx];
    sv = slices[idx] * scales[k] * inv;
    return true;
}

template <class TR>
__device__ __forceinline__ bool gaussian_single(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps, float sv, float inv,
                                                const unsigned char* __restrict__ mask, float2* __restrict__ acc2)
{
    bool any = false;
    if (ps.interior) {
        float p[TR::SUP];
        auto tap = [&](int i, float psf, bool ok, int v) { p[i] = psf; if (ok && mask[v]) any = true; };
        auto row = [&](int v0) { red_row_paired<TR::SUP>(acc2, v0, p, sv, inv); };
        if (g.recur) psf_rows<TR, true, true>(g, vg, ps, tap, row);
        else psf_rows<TR, true, false>(g, vg, ps, tap, row);
    } else {
        psf_rows_dispatch<TR>(g, vg, ps,
            [&](int, float psf, bool ok, int v) {
                if (ok) {
                    atomicAdd(&acc2[v], make_float2(psf * sv, psf * inv));
                    if (mask[v]) any = true;
                }
            },
            [](int) {});
    }
    return any;
}

// Per-pixel inputs of K3: false when the pixel contributes nothing (no PSF mass or zero weight).
template <class TR>
__device__ __forceinline__ bool superres_pixel(uint32_t idx, int Nx, int P, const float* __restrict__ slices,
                                               const float* __restrict__ weights, const float* __restrict__ simslices,
                                               const float* __restrict__ slice_weights, const float* __restrict__ scales,
                                               const float* __restrict__ psf_sums, int& k, int& x, int& y, float& aw, float& cw)
{
    const float sume = psf_sums[idx];
    if (sume == 0.0f) return false;
    k = idx / P;
    const int pix = idx - k * P;
    y = pix / Nx; x = pix - y * Nx;
    const float w = weights[idx];
    const float ss = simslices[idx];
    float sliceVal = slices[idx] * scales[k];
    sliceVal = (ss > 0.0f) ? (sliceVal - ss) : 0.0f;
    cw = w * slice_weights[k] / sume;               // psf/sume * w * slice_weight
    aw = cw * sliceVal;
    return cw != 0.0f;                              // a pixel with zero weight adds exact zeros everywhere: skip its taps
}

template <class TR>
__device__ __forceinline__ void superres_single(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps, float aw, float cw,
                                                float2* __restrict__ acc2)
{
    if (ps.interior) {
        float p[TR::SUP];
        auto tap = [&](int i, float psf, bool, int) { p[i] = psf; };
        auto row = [&](int v0) { red_row_paired<TR::SUP>(acc2, v0, p, aw, cw); };
        if (g.recur) psf_rows<TR, true, true>(g, vg, ps, tap, row);
        else psf_rows<TR, true, false>(g, vg, ps, tap, row);
    } else {
        psf_rows_dispatch<TR>(g, vg, ps,
            [&](int, float psf, bool ok, int v) { if (ok) atomicAdd(&acc2[v], make_float2(psf * aw, psf * cw)); },
            [](int) {});
    }
}

