// svr_common.cuh -- shared device structs and helpers of libsvr_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

#define SVR_PSF_SUPPORT 16     // MAX_PSF_SUPPORT, include/reconstruction_cuda2.cuh:74
#define SVR_PSF_CENTRE 7       // (MAX_PSF_SUPPORT - 1) / 2, reconstruction_cuda2.cu:219
#define SVR_STEP 0.0001f       // __step, include/reconstruction_cuda2.cuh:54

// Per-slice geometry, rebuilt on the device whenever the matrices / voxel sizes change
// (svr_set_slice_matrices, svr_set_slice_dims).  176 bytes, 16-byte aligned.
struct __align__(16) SliceGeom {
    float i2w[12];   // slice image -> world, rows 0..2
    float t[12];     // slice -> volume transform (world), rows 0..2
    float a[12];     // comb = W2I * Tinv * reconI2W, rows 0..2 (reconstruction_cuda2.cu:223)
    // tap-offset basis in "PSF units": column j of comb scaled per row by
    //   row x: dim.x * kx   row y: dim.y * ky   row z: dim.z      (cuda2.cu:125-126,158)
    float bx[3], by[3], bz[3];
    float kx, ky, kpad;      // dim.x / 2.3548, dim.y / 2.3548 (in-plane scale applied inside calcPSF)
    float gz;                // -log2(e) / (2 sigma_z^2), sigma_z = dim.z / 2.3548 (cuda2.cu:114,130)
    float dimx, dimy, dimz;  // slice voxel size
    float pad0;
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

// sinc^2(pi r) * exp(-dz^2 / (2 sigma_z^2)) with (ux,uy) already in the sinc's units and
// gz = -log2(e)/(2 sigma_z^2).  Restates calcPSF (reconstruction_cuda2.cu:112-131, USE_SINC_PSF);
// the reference is built with --use_fast_math, hence the approx MUFU forms.  sinc(0) = 1 (deviation D5).
__device__ __forceinline__ float psf_eval(float ux, float uy, float dz, float gz)
{
    const float u = fmaf(ux, ux, uy * uy);
    const float rinv = rsqrt_approx(u);           // +inf at u == 0
    const float r = u * rinv;                      // NaN at u == 0
    const float sn = sin_approx(3.14159265359f * r);
    float si = sn * (rinv * 0.31830988618f);       // sin(pi r) / (pi r)
    si = (u > 0.0f) ? si : 1.0f;
    const float g = ex2_approx(dz * dz * gz);
    return si * si * g;
}

// Per-pixel constants of the tap loop.
struct PixelSetup {
    int cx, cy, cz;          // rounded volume voxel the pixel centre maps to (cuda2.cu:225-226)
    float ex, ey, ez;        // PSF-unit offset of that voxel from the pixel: scale*((A c - p)*dim - psf_c)
};

__device__ __forceinline__ int f2i_clamped(float f)
{   // round-to-nearest-even is irrelevant here: f is already integer-valued (roundf result)
    f = fminf(fmaxf(f, -1.0e6f), 1.0e6f);         // NaN -> -1e6 (fmaxf returns the non-NaN operand)
    return __float2int_rn(f);
}

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
    ps.ez = ((q.z - p.z) * g.dimz - vg.psf_c[2]);
    return ps;
}

// The shared tap loop of K1/K2/K3 (reconstruction_cuda2.cu:229-247, 262-288, 372-394, 498-520).
// Walks the 16^3 support (x innermost) and calls body(psf, voxel_linear_index) for every tap that
//   (a) survives the epsilon-skip against the last ACCEPTED tap of its x-row (quirks Q1/Q2), and
//   (b) lies inside the volume after the unsigned saturation of negative coordinates (Q4).
// Rows whose (clamped) y or z lies outside the volume contribute nothing and are skipped whole:
// the skip state is per row, so this is exact.
template <class Body>
__device__ __forceinline__ void psf_tap_loop(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps, Body&& body)
{
    const int vx = vg.vx, vy = vg.vy, vz = vg.vz;
    const float gz = g.gz;
    const float bx0 = g.bx[0], by0 = g.by[0], bz0 = g.bz[0];
#pragma unroll 1
    for (int oz = -SVR_PSF_CENTRE; oz <= SVR_PSF_SUPPORT - 1 - SVR_PSF_CENTRE; ++oz) {
        const int zi = max(ps.cz + oz, 0);
        if (zi >= vz) continue;
        const float foz = (float)oz;
        const float zx = fmaf(foz, g.bx[2], ps.ex), zy = fmaf(foz, g.by[2], ps.ey), zz = fmaf(foz, g.bz[2], ps.ez);
#pragma unroll 1
        for (int oy = -SVR_PSF_CENTRE; oy <= SVR_PSF_SUPPORT - 1 - SVR_PSF_CENTRE; ++oy) {
            const int yi = max(ps.cy + oy, 0);
            if (yi >= vy) continue;
            const float foy = (float)oy;
            const float rx = fmaf(foy, g.bx[1], zx), ry = fmaf(foy, g.by[1], zy), rz = fmaf(foy, g.bz[1], zz);
            const int rowbase = (zi * vy + yi) * vx;
            float old = FLT_MAX;
#pragma unroll
            for (int ox = -SVR_PSF_CENTRE; ox <= SVR_PSF_SUPPORT - 1 - SVR_PSF_CENTRE; ++ox) {
                const float fox = (float)ox;
                const float psf = psf_eval(fmaf(fox, bx0, rx), fmaf(fox, by0, ry), fmaf(fox, bz0, rz), gz);
                // abs(oldPSF - psfval) < PSF_EPSILON with a double 1e-5 (cuda2.cu:238): true iff the float
                // difference is <= 1e-5f (the largest float below the double literal).
                if (fabsf(old - psf) <= 1.0e-5f) continue;
                old = psf;
                const int xi = max(ps.cx + ox, 0);
                if (xi < vx) body(psf, rowbase + xi);
            }
        }
    }
}

static inline int divup_i(long long a, long long b) { return (int)((a + b - 1) / b); }
