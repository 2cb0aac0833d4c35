// svr_common.cuh -- shared device structs and helpers of libsvr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define SVR_PSF_SUPPORT 16     // MAX_PSF_SUPPORT, include/reconstruction_cuda2.cuh:74
#define SVR_PSF_CENTRE 7       // (MAX_PSF_SUPPORT - 1) / 2, reconstruction_cuda2.cu:219
#define SVR_STEP 0.0001f       // __step, include/reconstruction_cuda2.cuh:54

// Flavour traits: SVR (reconstruction_cuda2.cu) and PVR (patchBased*_gpu.cu, include/reconConfig.cuh:119-140,
// include/pointSpreadFunction.cuh) run the same tap loop with different constants.
struct SvrTraits {
    static constexpr int SUP = 16;      // MAX_PSF_SUPPORT, reconstruction_cuda2.cuh:74 -> offsets -7..+8
    static constexpr int CEN = 7;
    // abs(oldPSF - psfval) < PSF_EPSILON with a DOUBLE 1e-5 (cuda2.cu:238): skipped iff the float difference is
    // <= 1e-5f (the largest float below the double literal)
    static __device__ __forceinline__ bool accept(float old, float psf) { return !(fabsf(old - psf) <= 1.0e-5f); }
    static __device__ __forceinline__ bool sume_ok(float s) { return s > 0.5f; }          // cuda2.cu:251
};
struct PvrTraits {
    static constexpr int SUP = 12;      // MAX_PSF_SUPPORT, reconConfig.cuh:140 -> offsets -5..+6
    static constexpr int CEN = 5;
    // PSF_EPSILON is a FLOAT literal here (reconConfig.cuh:138)
    static __device__ __forceinline__ bool accept(float old, float psf) { return !(fabsf(old - psf) < 1.0e-5f); }
    static __device__ __forceinline__ bool sume_ok(float s) { return s > 1.0e-5f; }       // patchBasedPSFReconstruction_gpu.cu:110
};

// Per-slice geometry, rebuilt on the device whenever the matrices / voxel sizes change
// (svr_set_slice_matrices, svr_set_slice_dims).  16-byte aligned.
//
// "PSF units": the tap position is carried pre-scaled so that calcPSF (reconstruction_cuda2.cu:112-131)
// needs no per-tap constant multiplies:
//   in-plane rows x,y are scaled by  dim * (dim / 2.3548) * pi   -> u = ux^2 + uy^2 = (pi r)^2
//   the through-plane row z by       dim.z * sqrt(log2(e)/2) / sigma_z, sigma_z = dim.z / 2.3548
//                                    -> gauss = 2^(-dz^2)
struct __align__(16) SliceGeom {
    float i2w[12];   // slice image -> world, rows 0..2
    float t[12];     // slice -> volume transform (world), rows 0..2
    float a[12];     // comb = W2I * Tinv * reconI2W, rows 0..2 (reconstruction_cuda2.cu:223)
    float bx[3], by[3], bz[3];   // column j of comb in PSF units (tap-offset basis)
    float kx, ky, kz;            // PSF-unit scale per mm for rows x, y, z
    float dimx, dimy, dimz;      // slice voxel size
    // Gaussian forward differencing along an x-row (dz_i = dz_0 + i*bz0): g_{i+1} = g_i * rho_i,
    // rho_{i+1} = rho_i * kappa with rho_i = 2^-(2 dz_i bz0 + bz0^2), kappa = 2^-(2 bz0^2).
    float two_b, bb, kappa;
    int recur;                   // 1 when the recurrence cannot overflow (|bz0| <= 1, |bz0|+|bz1|+|bz2| <= 4)
    int through_plane_rows;      // 1 when a step along the slice's x moves less than half a voxel along the volume's x:
                                 // the tap rows (runs along volume x) of a warp's pixels then share no cache lines
    int win_class;               // 1: the slice's x / y axes are not both within ~10 degrees of the volume's x / y axes: its scatter
                                 // goes through the warp-window kernel under SVR_TUNE_SCATTER = 3 (svr_window.cu)
};

struct VolGeom {
    int vx, vy, vz;
    float rw2i[12];          // recon world -> image rows 0..2
    float psf_c[3];          // d_PSFI2W * ((PSFsize-1)/2)  (cuda2.cu:172), ~0
};

__device__ __forceinline__ float3 mat_pt(const float* __restrict__ m, float3 v)
{   // operator*(Matrix4, float3), recon_volumeHelper.cuh:106-117
    return make_float3(m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3],
                       m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7],
                       m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11]);
}

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// sinc^2(pi r) * exp(-z^2 / (2 sigma_z^2)) for a tap position in PSF units: u = (pi r)^2, gauss = 2^(-dz^2).
// Restates calcPSF (reconstruction_cuda2.cu:112-131, USE_SINC_PSF); the reference is built with
// --use_fast_math, hence the approx MUFU forms (rsqrt, sin for the sinc; ex2 for the Gaussian).
// sin.approx has an ABSOLUTE error (~2^-21), which sin(x)/x amplifies by 1/x: below x^2 = 1e-3 the
// two-term series 1 - x^2/6 (error < 1e-8) is used instead.  It also covers x == 0, where the
// reference evaluates sin(0)/0 = NaN (deviation D5).
__device__ __forceinline__ float sinc2_eval(float ux, float uy)
{
    const float u = fmaf(ux, ux, uy * uy);
    const float rinv = rsqrt_approx(u);
    float si = sin_approx(u * rinv) * rinv;            // sin(pi r) / (pi r)
    si = (u < 1.0e-3f) ? fmaf(u, -0.16666667f, 1.0f) : si;
    return si * si;
}

// Per-pixel constants of the tap loop.
struct PixelSetup {
    int cx, cy, cz;          // rounded volume voxel the pixel centre maps to (cuda2.cu:225-226)
    float ex, ey, ez;        // PSF-unit offset of that voxel from the pixel: k * ((A c - p) * dim - psf_c)
    bool interior;           // the whole 16^3 support lies inside the volume: no clamping, no bounds checks
};

__device__ __forceinline__ int f2i_clamped(float f)
{   // f is already integer-valued (roundf result); NaN -> -1e6 (fmaxf returns the non-NaN operand)
    f = fminf(fmaxf(f, -1.0e6f), 1.0e6f);
    return __float2int_rn(f);
}

template <class TR>
__device__ __forceinline__ PixelSetup pixel_setup(const SliceGeom& g, const VolGeom& vg, int x, int y)
{
    PixelSetup ps;
    const float3 p = make_float3((float)x, (float)y, 0.0f);
    // d_reconstructedW2I * (T * (I2W * slicePos)), evaluated in that order (cuda2.cu:225)
    const float3 w = mat_pt(vg.rw2i, mat_pt(g.t, mat_pt(g.i2w, p)));
    const float3 c = make_float3(roundf(w.x), roundf(w.y), roundf(w.z));
    ps.cx = f2i_clamped(c.x); ps.cy = f2i_clamped(c.y); ps.cz = f2i_clamped(c.z);
    const float3 cc = make_float3((float)ps.cx, (float)ps.cy, (float)ps.cz);
    const float3 q = mat_pt(g.a, cc);              // comb * centre voxel -> slice pixel coordinates
    ps.ex = ((q.x - p.x) * g.dimx - vg.psf_c[0]) * g.kx;
    ps.ey = ((q.y - p.y) * g.dimy - vg.psf_c[1]) * g.ky;
    ps.ez = ((q.z - p.z) * g.dimz - vg.psf_c[2]) * g.kz;
    const int lo = TR::CEN, hi = TR::SUP - 1 - TR::CEN;
    ps.interior = ps.cx - lo >= 0 && ps.cx + hi < vg.vx && ps.cy - lo >= 0 && ps.cy + hi < vg.vy &&
                  ps.cz - lo >= 0 && ps.cz + hi < vg.vz;
    return ps;
}

// The shared tap loop of K1/K2/K3 (reconstruction_cuda2.cu:229-247, 262-288, 372-394, 498-520).
// Walks the 16^3 support (x innermost, fully unrolled) and calls, per tap,
//     tap(i, psf, ok, v)      i = ox + 7 (compile-time after unrolling), v = linear voxel index
// where ok means the tap (a) survives the epsilon-skip against the last ACCEPTED tap of its x-row
// (quirks Q1/Q2: the first tap of a row is always accepted) and (b) lies inside the volume after the
// unsigned saturation of negative coordinates (Q4); psf is 0 when !ok.  After each x-row it calls
//     row_end(v0)             v0 = linear index of the row's first tap (meaningful when INTERIOR).
// INTERIOR (the whole support inside the volume) drops every clamp and bounds check, so v = v0 + i
// and the loads / reductions of a row use immediate offsets from one row pointer.  Otherwise rows whose
// (clamped) y or z lies outside the volume are skipped whole: the skip state is per row, so this is exact.
//
// RECUR evaluates the through-plane Gaussian of a row by forward differencing in two segments of 8 taps
// (2 ex2 per segment instead of 1 per tap: the MUFU unit is the busiest pipe of these kernels); the
// accumulated rounding (<= ~2.5e-6 relative after 7 steps) is of the size of the ex2.approx argument
// rounding of the direct form.  Slices with extreme through-plane scaling use the direct form.
template <class TR, bool INTERIOR, bool RECUR, class Tap, class RowEnd>
__device__ __forceinline__ void psf_rows(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps, Tap&& tap,
                                         RowEnd&& row_end)
{
    const int vx = vg.vx, vy = vg.vy, vz = vg.vz;
    const float bx0 = g.bx[0], by0 = g.by[0], bz0 = g.bz[0];
    const float bx1 = g.bx[1], by1 = g.by[1], bz1 = g.bz[1];
    const float bx2 = g.bx[2], by2 = g.by[2], bz2 = g.bz[2];
    const float two_b = g.two_b, bb = g.bb, kappa = g.kappa;
    constexpr int HALF = TR::SUP / 2;
#pragma unroll 1
    for (int oz = -TR::CEN; oz <= TR::SUP - 1 - TR::CEN; ++oz) {
        const int zi = INTERIOR ? ps.cz + oz : max(ps.cz + oz, 0);
        if (!INTERIOR && zi >= vz) continue;
        const float foz = (float)oz;
        const float zx = fmaf(foz, bx2, ps.ex), zy = fmaf(foz, by2, ps.ey), zz = fmaf(foz, bz2, ps.ez);
#pragma unroll 1
        for (int oy = -TR::CEN; oy <= TR::SUP - 1 - TR::CEN; ++oy) {
            const int yi = INTERIOR ? ps.cy + oy : max(ps.cy + oy, 0);
            if (!INTERIOR && yi >= vy) continue;
            const float foy = (float)oy;
            const float rx = fmaf(foy, bx1, zx), ry = fmaf(foy, by1, zy), rz = fmaf(foy, bz1, zz);
            const int rowbase = (zi * vy + yi) * vx;
            const int v0 = rowbase + ps.cx - TR::CEN;
            float old = FLT_MAX;
            float gz = 0.f, rho = 0.f;
#pragma unroll
            for (int i = 0; i < TR::SUP; ++i) {
                const float fox = (float)(i - TR::CEN);
                float gauss;
                if (RECUR) {
                    if (i % HALF == 0) {                  // (re)start a segment with direct evaluations
                        const float dz = fmaf(fox, bz0, rz);
                        gz = ex2_approx(-(dz * dz));
                        rho = ex2_approx(-fmaf(two_b, dz, bb));
                    }
                    gauss = gz;
                    if (i % HALF != HALF - 1) { gz *= rho; rho *= kappa; }
                } else {
                    const float dz = fmaf(fox, bz0, rz);
                    gauss = ex2_approx(-(dz * dz));
                }
                const float psf = sinc2_eval(fmaf(fox, bx0, rx), fmaf(fox, by0, ry)) * gauss;
                const bool accept = TR::accept(old, psf);
                old = accept ? psf : old;
                if (INTERIOR) {
                    tap(i, accept ? psf : 0.0f, accept, v0 + i);
                } else {
                    const int xi = max(ps.cx + i - TR::CEN, 0);
                    const bool ok = accept && xi < vx;
                    tap(i, ok ? psf : 0.0f, ok, rowbase + xi);
                }
            }
            row_end(v0);
        }
    }
}

// One x-row of accepted PSF values (0 where the epsilon-skip rejects a tap): the inner loop of psf_rows for callers
// that place the row themselves (the paired scatter).  (rx, ry, rz) = PSF-unit position of the row's tap with ox = 0.
template <class TR, bool RECUR>
__device__ __forceinline__ void psf_row_values(const SliceGeom& g, float rx, float ry, float rz, float (&p)[TR::SUP])
{
    const float bx0 = g.bx[0], by0 = g.by[0], bz0 = g.bz[0];
    const float two_b = g.two_b, bb = g.bb, kappa = g.kappa;
    constexpr int HALF = TR::SUP / 2;
    float old = FLT_MAX;
    float gz = 0.f, rho = 0.f;
#pragma unroll
    for (int i = 0; i < TR::SUP; ++i) {
        const float fox = (float)(i - TR::CEN);
        float gauss;
        if (RECUR) {
            if (i % HALF == 0) {
                const float dz = fmaf(fox, bz0, rz);
                gz = ex2_approx(-(dz * dz));
                rho = ex2_approx(-fmaf(two_b, dz, bb));
            }
            gauss = gz;
            if (i % HALF != HALF - 1) { gz *= rho; rho *= kappa; }
        } else {
            const float dz = fmaf(fox, bz0, rz);
            gauss = ex2_approx(-(dz * dz));
        }
        const float psf = sinc2_eval(fmaf(fox, bx0, rx), fmaf(fox, by0, ry)) * gauss;
        const bool accept = TR::accept(old, psf);
        old = accept ? psf : old;
        p[i] = accept ? psf : 0.0f;
    }
}

// Dispatch on the (per-thread) interior flag and the (per-slice) recurrence flag.
template <class TR, class Tap, class RowEnd>
__device__ __forceinline__ void psf_rows_dispatch(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps, Tap&& tap,
                                                  RowEnd&& row_end)
{
    if (g.recur) {
        if (ps.interior) psf_rows<TR, true, true>(g, vg, ps, tap, row_end);
        else psf_rows<TR, false, true>(g, vg, ps, tap, row_end);
    } else {
        if (ps.interior) psf_rows<TR, true, false>(g, vg, ps, tap, row_end);
        else psf_rows<TR, false, false>(g, vg, ps, tap, row_end);
    }
}

static inline int divup_i(long long a, long long b) { return (int)((a + b - 1) / b); }
