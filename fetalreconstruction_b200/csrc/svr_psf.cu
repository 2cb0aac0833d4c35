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
#include "svr_psf_pixel.cuh"

// Resident CTAs per SM the PSF kernels are compiled for (register budget = 65536 / (128 * SVR_MINB)).
#ifndef SVR_MINB
#define SVR_MINB 5
#endif
// The paired scatter keeps two pixel states and a 2 x 18-voxel row accumulator in registers.
#ifndef SVR_MINB_PAIR
#define SVR_MINB_PAIR 4
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
                                  const float* __restrict__ dims, Mat16 reconI2W, SliceGeom* __restrict__ out, int flavor)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    float t1[16], comb[16];
    if (flavor == 0) {          // SVR: (W2I * Tinv) * reconI2W, cuda2.cu:223
        mat44_mul(W2I + 16 * k, Tinv + 16 * k, t1);
        mat44_mul(t1, reconI2W.m, comb);
    } else {                    // PVR: W2I * (InvTransformation * reconstructedI2W), patchBasedPSFReconstruction_gpu.cu:78
        mat44_mul(Tinv + 16 * k, reconI2W.m, t1);
        mat44_mul(W2I + 16 * k, t1, comb);
    }
    SliceGeom g;
    for (int i = 0; i < 12; ++i) { g.i2w[i] = I2W[16 * k + i]; g.t[i] = T[16 * k + i]; g.a[i] = comb[i]; }
    const float dx = dims[3 * k + 0], dy = dims[3 * k + 1];
    // PVR: sigma_z = dim.z and the through-plane offset is additionally divided by 2.5
    // (pointSpreadFunction.cuh:76,112); SVR: sigma_z = dim.z / 2.3548 (cuda2.cu:114)
    const float sigmaz = flavor == 0 ? dims[3 * k + 2] / 2.3548f : dims[3 * k + 2];
    const float dz = flavor == 0 ? dims[3 * k + 2] : dims[3 * k + 2] / 2.5f;
    g.dimx = dx; g.dimy = dy; g.dimz = dz;
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
    {   // row 0 of comb = d(slice x)/d(voxel): for a rigid map into an isotropic volume it is also, up to scale, the
        // direction of a slice-x step in voxel space
        const float nx = sqrtf(comb[0] * comb[0] + comb[1] * comb[1] + comb[2] * comb[2]);
        g.through_plane_rows = (nx > 0.f && fabsf(comb[0]) < 0.5f * nx) ? 1 : 0;
        const float ny = sqrtf(comb[4] * comb[4] + comb[5] * comb[5] + comb[6] * comb[6]);
        g.win_class = (fabsf(comb[0]) > 0.985f * nx && fabsf(comb[5]) > 0.985f * ny) ? 0 : 1;
    }
    out[k] = g;
}

int svr_launch_build_geom(svr_context* c)
{
    if (!c->have_mats || !c->have_dims || c->S == 0) return 0;
    Mat16 ri2w;
    for (int i = 0; i < 16; ++i) ri2w.m[i] = c->recon_i2w[i];
    const size_t n = (size_t)c->S * 16;
    build_geom_kernel<<<divup_i(c->S, 128), 128, 0, c->stream>>>(c->S, c->mats, c->mats + n, c->mats + 2 * n, c->mats + 3 * n,
                                                                 c->dims, ri2w, c->geom, c->flavor);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Paired scatter.  The scatter kernels are bound by the rate of global reductions (measured: half the REDs = half the
// kernel time), and two neighbouring pixels of a slice write almost the same voxels: their 16^3 supports are offset by
// d = centre(B) - centre(A), usually one voxel along one axis.  A thread therefore takes TWO consecutive valid pixels,
// walks the UNION of their supports volume-row by volume-row, sums both pixels' contributions to a row in registers
// (20 voxels: 16 + |d.x| + alignment) and flushes the row once: ~10 vector reductions per row pair instead of 18.
// Exact: the per-pixel tap values, the epsilon-skip chains and the set of (voxel, value) contributions are unchanged,
// only the order in which floats meet in the accumulator differs (as with any atomic scatter, Q8).
// The pairs are the pixels (x, x + 1) with x even of one slice row (svr_launch_compact_valid builds the list of even
// positions where at least one of the two is valid): a pair never straddles rows or slices, and a lone pixel runs
// the same code with its partner switched off, so a warp has ONE code path.  Handled here: both pixels interior and
// |d.x| <= 2 (any d.y, d.z); anything else (mask touching the volume faces, very coarse pixels) takes the one-pixel path.
struct PairWeights { float aA, cA, aB, cB; };     // {numerator, denominator} scale of pixel A and of pixel B

constexpr int PAIR_SLACK = 4;                     // row accumulator = SUP + 4 voxels: |d.x| <= 2 plus one for the alignment

// Adds p[0..SUP) * (a, c) into (qn, qd) at element offset s in {0, 1, 2, 3}.
template <int SUP>
__device__ __forceinline__ void pair_accumulate(float (&qn)[SUP + PAIR_SLACK], float (&qd)[SUP + PAIR_SLACK], const float (&p)[SUP], int s,
                                                float a, float c)
{
    const bool s1 = (s & 1) != 0, s2 = (s & 2) != 0;
    float t[SUP + 1];                              // shift by the low bit
    t[0] = s1 ? 0.0f : p[0];
#pragma unroll
    for (int j = 1; j < SUP; ++j) t[j] = s1 ? p[j - 1] : p[j];
    t[SUP] = s1 ? p[SUP - 1] : 0.0f;
#pragma unroll
    for (int j = 0; j < SUP + 3; ++j) {            // shift by two
        const float lo = j < SUP + 1 ? t[j] : 0.0f;
        const float hi = j >= 2 ? t[j - 2] : 0.0f;
        const float v = s2 ? hi : lo;
        qn[j] = fmaf(v, a, qn[j]);
        qd[j] = fmaf(v, c, qd[j]);
    }
}

// MASKFLAG: also report, per pixel, whether an accepted tap landed on a voxel with mask != 0 (K1's sliceVoxel_count).
template <class TR, bool RECUR, bool MASKFLAG>
__device__ __forceinline__ void scatter_pair(const SliceGeom& g, const VolGeom& vg, const PixelSetup& A, const PixelSetup& B,
                                             bool liveA, bool liveB, const PairWeights w, float2* __restrict__ acc2,
                                             const unsigned char* __restrict__ mask, bool& anyA, bool& anyB)
{
    constexpr int SUP = TR::SUP, CEN = TR::CEN, HI = TR::SUP - 1 - TR::CEN;
    const int vx = vg.vx, vy = vg.vy;
    const float bx1 = g.bx[1], by1 = g.by[1], bz1 = g.bz[1];
    const float bx2 = g.bx[2], by2 = g.by[2], bz2 = g.bz[2];
    // bounds of the union of the live supports (a dead pixel takes its partner's centre)
    const int czA = liveA ? A.cz : B.cz, czB = liveB ? B.cz : A.cz;
    const int cyA = liveA ? A.cy : B.cy, cyB = liveB ? B.cy : A.cy;
    const int cxA = liveA ? A.cx : B.cx, cxB = liveB ? B.cx : A.cx;
    const int zlo = min(czA, czB) - CEN, zhi = max(czA, czB) + HI;
    const int ylo = min(cyA, cyB) - CEN, yhi = max(cyA, cyB) + HI;
    const int xmin = min(cxA, cxB) - CEN;
    if (MASKFLAG) {
        // "did an accepted tap land on a masked voxel" is almost always decided by the row through the pixel's own
        // centre voxel: test that row first, and only pixels it leaves undecided pay the 16 mask loads per row below
        float p[SUP];
        if (liveA) {
            psf_row_values<TR, RECUR>(g, A.ex, A.ey, A.ez, p);
            const int v0 = (A.cz * vy + A.cy) * vx + A.cx - CEN;
#pragma unroll
            for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && mask[v0 + i]) anyA = true;
        }
        if (liveB) {
            psf_row_values<TR, RECUR>(g, B.ex, B.ey, B.ez, p);
            const int v0 = (B.cz * vy + B.cy) * vx + B.cx - CEN;
#pragma unroll
            for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && mask[v0 + i]) anyB = true;
        }
    }
#pragma unroll 1
    for (int Z = zlo; Z <= zhi; ++Z) {
        const int ozA = Z - A.cz, ozB = Z - B.cz;
        const bool zA = liveA && (unsigned)(ozA + CEN) < (unsigned)SUP, zB = liveB && (unsigned)(ozB + CEN) < (unsigned)SUP;
        const float fzA = (float)ozA, fzB = (float)ozB;
        const float zxA = fmaf(fzA, bx2, A.ex), zyA = fmaf(fzA, by2, A.ey), zzA = fmaf(fzA, bz2, A.ez);
        const float zxB = fmaf(fzB, bx2, B.ex), zyB = fmaf(fzB, by2, B.ey), zzB = fmaf(fzB, bz2, B.ez);
#pragma unroll 1
        for (int Y = ylo; Y <= yhi; ++Y) {
            const int oyA = Y - A.cy, oyB = Y - B.cy;
            const bool okA = zA && (unsigned)(oyA + CEN) < (unsigned)SUP;
            const bool okB = zB && (unsigned)(oyB + CEN) < (unsigned)SUP;
            if (!okA && !okB) continue;
            const int rowbase = (Z * vy + Y) * vx;
            const int x0 = xmin - ((rowbase + xmin) & 1);          // even accumulator index: 16-byte aligned pairs
            float qn[SUP + PAIR_SLACK], qd[SUP + PAIR_SLACK];      // voxels x0 .. x0 + SUP + 3
#pragma unroll
            for (int j = 0; j < SUP + PAIR_SLACK; ++j) { qn[j] = 0.f; qd[j] = 0.f; }
            float p[SUP];
            if (okA) {
                const float foy = (float)oyA;
                psf_row_values<TR, RECUR>(g, fmaf(foy, bx1, zxA), fmaf(foy, by1, zyA), fmaf(foy, bz1, zzA), p);
                if (MASKFLAG && !anyA) {
                    const int v0 = rowbase + A.cx - CEN;
#pragma unroll
                    for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && mask[v0 + i]) anyA = true;
                }
                pair_accumulate<SUP>(qn, qd, p, A.cx - CEN - x0, w.aA, w.cA);
            }
            if (okB) {
                const float foy = (float)oyB;
                psf_row_values<TR, RECUR>(g, fmaf(foy, bx1, zxB), fmaf(foy, by1, zyB), fmaf(foy, bz1, zzB), p);
                if (MASKFLAG && !anyB) {
                    const int v0 = rowbase + B.cx - CEN;
#pragma unroll
                    for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && mask[v0 + i]) anyB = true;
                }
                pair_accumulate<SUP>(qn, qd, p, B.cx - CEN - x0, w.aB, w.cB);
            }
            float4* base = reinterpret_cast<float4*>(acc2 + (rowbase + x0));
#pragma unroll
            for (int m = 0; m < (SUP + PAIR_SLACK) / 2; ++m) {
                const float4 r = make_float4(qn[2 * m], qd[2 * m], qn[2 * m + 1], qd[2 * m + 1]);
                if (r.y + r.w > 0.0f) atomicAdd(base + m, r);       // denominators are psf * c with c > 0: skips all-zero pairs (and NaNs)
            }
        }
    }
}

// Which path a pair takes: 0 = nothing live, 1 = scatter_pair, 2 = one-pixel path(s).
__device__ __forceinline__ int pair_mode(const PixelSetup& a, const PixelSetup& b, bool liveA, bool liveB)
{
    if (!liveA && !liveB) return 0;
    if ((liveA && !a.interior) || (liveB && !b.interior)) return 2;
    if (liveA && liveB && abs(a.cx - b.cx) > 2) return 2;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// K1: gaussianReconstructionKernel3D_tex (reconstruction_cuda2.cu:176-295).
// Pass 1: sume = sum of accepted in-volume taps (mask ignored, quirk Q3); stored only if > 0.5.
// Pass 2: scatter psf/sume * {s*scale, 1}; flag the pixel if any accepted tap landed on a masked voxel.
// Pass 1 (one thread per valid pixel, the occupancy of the compute-bound kernels): sume; pixels that take part are
// marked voxel_flag = 2 for pass 2 (which leaves 1 / 0 there, the flag the reference keeps).
template <class TR>
__global__ void __launch_bounds__(128, SVR_MINB)
gaussian_sume_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int Nx, int P,
                     const SliceGeom* __restrict__ geom, VolGeom vg, float* __restrict__ psf_sums,
                     unsigned char* __restrict__ voxel_flag, const char* __restrict__ spx)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_valid) return;
    const uint32_t idx = valid_idx[t];
    const int k = idx / P, pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    // PVR superpixels: sume only accumulates when the pixel's own flag is '1' (patchBasedPSFReconstruction_gpu.cu:99),
    // i.e. other pixels end with sume = 0 and return; the mask is char[64*64] per patch, indexed x + 64*y.
    if (spx && spx[(size_t)k * 4096 + x + 64 * y] != '1') return;
    const SliceGeom& g = geom[k];
    const PixelSetup ps = pixel_setup<TR>(g, vg, x, y);
    float sume = 0.f;
    psf_rows_dispatch<TR>(g, vg, ps, [&](int, float psf, bool, int) { sume += psf; }, [](int) {});
    if (!TR::sume_ok(sume)) return;
    psf_sums[idx] = sume;
    voxel_flag[idx] = 2;
}

// Pass 2: one thread = one pixel pair (see scatter_pair).
template <class TR>
__global__ void __launch_bounds__(128, SVR_MINB_PAIR)
gaussian_scatter_kernel(uint32_t n_pairs, const uint32_t* __restrict__ pair_idx, int Nx, int P,
                        const float* __restrict__ slices, const float* __restrict__ scales,
                        const SliceGeom* __restrict__ geom, VolGeom vg, const unsigned char* __restrict__ mask,
                        float2* __restrict__ acc2, const float* __restrict__ psf_sums, unsigned char* __restrict__ voxel_flag,
                        int* __restrict__ slice_count, int skip_win)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const uint32_t ia = pair_idx[t];
    if (skip_win && geom[ia / (uint32_t)P].win_class) return;      // left to the warp-window kernel
    const int xe = (int)(ia % (uint32_t)Nx);
    int k = 0, kB = 0;
    PixelSetup psA = {}, psB = {};
    float svA = 0.f, invA = 0.f, svB = 0.f, invB = 0.f;
    const bool liveA = gaussian_pixel<TR>(ia, Nx, P, slices, scales, geom, vg, psf_sums, voxel_flag, k, psA, svA, invA);
    const bool liveB = xe + 1 < Nx && gaussian_pixel<TR>(ia + 1, Nx, P, slices, scales, geom, vg, psf_sums, voxel_flag, kB, psB, svB, invB);
    if (!liveA && !liveB) return;
    if (!liveA) k = kB;
    const SliceGeom& g = geom[k];
    bool anyA = false, anyB = false;
    const int mode = pair_mode(psA, psB, liveA, liveB);
    if (mode == 1) {
        const PairWeights w = { svA, invA, svB, invB };
        if (g.recur) scatter_pair<TR, true, true>(g, vg, psA, psB, liveA, liveB, w, acc2, mask, anyA, anyB);
        else scatter_pair<TR, false, true>(g, vg, psA, psB, liveA, liveB, w, acc2, mask, anyA, anyB);
    } else {
        if (liveA) anyA = gaussian_single<TR>(g, vg, psA, svA, invA, mask, acc2);
        if (liveB) anyB = gaussian_single<TR>(g, vg, psB, svB, invB, mask, acc2);
    }
    if (liveA) voxel_flag[ia] = anyA ? 1 : 0;
    if (liveB) voxel_flag[ia + 1] = anyB ? 1 : 0;
    const int n = (anyA ? 1 : 0) + (anyB ? 1 : 0);
    if (n) atomicAdd(&slice_count[k], n);
}

int svr_launch_gaussian_scatter(svr_context* c)
{
    if (c->n_valid == 0) return 0;
    ProfScope prof(c, 0);
    const int P = c->Nx * c->Ny;
    const bool avail = svr_window_scatter_available(c);
    const bool window = (c->tune_scatter == 1 || c->tune_scatter == 2) && avail;
    const int split = (c->tune_scatter == 3 && avail) ? 1 : 0;     // aligned slices: paired scatter, the others: warp windows
    if (c->flavor == 0) {
        gaussian_sume_kernel<SvrTraits><<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
            c->n_valid, c->valid_idx, c->Nx, P, c->geom, c->vg, c->psf_sums, c->voxel_flag, nullptr);
        SVR_KERNEL_CHECK(c);
        if (window) return svr_launch_window_scatter(c, 1, -1);
        gaussian_scatter_kernel<SvrTraits><<<divup_i(c->n_pairs, 128), 128, 0, c->stream>>>(
            c->n_pairs, c->pair_idx, c->Nx, P, c->slices, c->scales, c->geom, c->vg, c->mask_u8, c->acc2, c->psf_sums,
            c->voxel_flag, c->slice_count, split);
    } else {
        gaussian_sume_kernel<PvrTraits><<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
            c->n_valid, c->valid_idx, c->Nx, P, c->geom, c->vg, c->psf_sums, c->voxel_flag, c->use_spx ? c->spx : nullptr);
        SVR_KERNEL_CHECK(c);
        if (window) return svr_launch_window_scatter(c, 1, -1);
        gaussian_scatter_kernel<PvrTraits><<<divup_i(c->n_pairs, 128), 128, 0, c->stream>>>(
            c->n_pairs, c->pair_idx, c->Nx, P, c->slices, c->scales, c->geom, c->vg, c->mask_u8, c->acc2, c->psf_sums,
            c->voxel_flag, c->slice_count, split);
    }
    SVR_KERNEL_CHECK(c);
    if (split) return svr_launch_window_scatter(c, 1, 1);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K2: simulateSlicesKernel3D_tex (reconstruction_cuda2.cu:298-404).
// pack2[v] = {recon[v]*m, m} with m = (mask != 0), so a tap is one predicated 64-bit load + 2 FFMA.
// Cooperative row staging for K2.  When a slice's pixel rows do not run along the volume's x (sagittal / coronal stacks
// against an axial template), the 32 pixels of a warp sit on 32 different volume rows: every per-tap load touches 32
// cache lines and the kernel becomes bound by L1 wavefronts (measured 4.3x the time of an aligned stack).  Here the warp
// fetches the 32 tap rows of one (oy, oz) step together -- 16 lanes read one pixel's 128-byte row, two rows per load
// instruction, fully coalesced -- parks them in shared memory (row stride 34 floats: the 64-bit reads of a half-warp
// hit 16 distinct bank pairs) and every lane then reads its own row from there.  Same taps, same epsilon-skip chain,
// same order of the per-pixel sums; rejected taps contribute psf = 0 instead of being skipped.
#ifndef SVR_COOP
#define SVR_COOP 1
#endif
constexpr int COOP_STRIDE = 34;

// ZINNER walks oz innermost.  The staged rows of step (oy, oz + 1) are the neighbours' rows of step (oy, oz) when the
// warp's pixels are neighbours along the volume's z, so with oz innermost they are still in L1 (the ncu capture of the
// oy-innermost order shows a 3 % L1 hit rate on such a stack: every row comes from L2).  It changes the order of the
// per-pixel sum, i.e. results agree within tolerance, not bit for bit: compiled only with -DSVR_COOP_ZINNER=1 (an
// experiment prepared at the end of round 1, not yet measured).
#ifndef SVR_COOP_ZINNER
#define SVR_COOP_ZINNER 0
#endif

template <class TR, bool RECUR, bool ZINNER>
__device__ __forceinline__ void simulate_rows_coop(const SliceGeom& g, const VolGeom& vg, const PixelSetup& ps,
                                                   const float2* __restrict__ pack2, float* __restrict__ stage, float& sim, float& wsum)
{
    static_assert(TR::SUP == 16, "two 16-tap rows per warp-wide load");
    const int vx = vg.vx, vy = vg.vy;
    const float bx1 = g.bx[1], by1 = g.by[1], bz1 = g.bz[1];
    const float bx2 = g.bx[2], by2 = g.by[2], bz2 = g.bz[2];
    const int lane = threadIdx.x & 31, half = lane >> 4, el = lane & 15;
    float2* mine = reinterpret_cast<float2*>(stage + lane * COOP_STRIDE);
#pragma unroll 1
    for (int a = -TR::CEN; a <= TR::SUP - 1 - TR::CEN; ++a) {
#pragma unroll 1
        for (int b = -TR::CEN; b <= TR::SUP - 1 - TR::CEN; ++b) {
            const int oz = ZINNER ? b : a, oy = ZINNER ? a : b;
            const float foz = (float)oz, foy = (float)oy;
            const float zx = fmaf(foz, bx2, ps.ex), zy = fmaf(foz, by2, ps.ey), zz = fmaf(foz, bz2, ps.ez);
            const int v0 = ((ps.cz + oz) * vy + (ps.cy + oy)) * vx + ps.cx - TR::CEN;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int src = 2 * j + half;
                const int v0s = __shfl_sync(0xffffffffu, v0, src);
                *reinterpret_cast<float2*>(stage + src * COOP_STRIDE + 2 * el) = __ldg(&pack2[v0s + el]);
            }
            __syncwarp();
            float p[TR::SUP];
            psf_row_values<TR, RECUR>(g, fmaf(foy, bx1, zx), fmaf(foy, by1, zy), fmaf(foy, bz1, zz), p);
#pragma unroll
            for (int i = 0; i < TR::SUP; ++i) {
                const float2 pm = mine[i];
                sim = fmaf(p[i], pm.x, sim);
                wsum = fmaf(p[i], pm.y, wsum);
            }
            __syncwarp();
        }
    }
}

template <class TR>
__global__ void __launch_bounds__(128, SVR_MINB)
simulate_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int Nx, int P,
                const SliceGeom* __restrict__ geom, VolGeom vg, const float2* __restrict__ pack2,
                const float* __restrict__ psf_sums, float* __restrict__ simslices, float* __restrict__ simweights,
                unsigned char* __restrict__ siminside, int* __restrict__ slice_inside, int skip_class)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t idx = 0;
    float sume = 0.0f;
    if (t < n_valid) { idx = valid_idx[t]; sume = psf_sums[idx]; }
    // skip_class >= 0: slices with through_plane_rows == skip_class are left to the window kernel (svr_window.cu)
    if (skip_class >= 0 && sume != 0.0f && geom[idx / P].through_plane_rows == skip_class) sume = 0.0f;
    const bool alive = sume != 0.0f;
#if SVR_COOP
    if (TR::SUP != 16 && !alive) return;           // the cooperative path (SVR only) needs whole warps
#else
    if (!alive) return;
#endif
    const int k = idx / P, pix = idx - k * P;
    const int y = pix / Nx, x = pix - y * Nx;
    const SliceGeom& g = geom[k];
    const PixelSetup ps = pixel_setup<TR>(g, vg, x, y);

    float sim = 0.f, wsum = 0.f;
    bool done = false;
#if SVR_COOP
    if constexpr (TR::SUP == 16) {
        __shared__ __align__(16) float stage_all[4][32 * COOP_STRIDE];
        const bool ok = alive && ps.interior && g.through_plane_rows && g.recur;
        if (__all_sync(0xffffffffu, ok)) {
#if SVR_COOP_ZINNER
            // the warp's pixels are neighbours along the slice's x: which volume axis that is decides the loop order
            if (__all_sync(0xffffffffu, fabsf(g.bx[2]) > fabsf(g.bx[1])))
                simulate_rows_coop<TR, true, true>(g, vg, ps, pack2, stage_all[threadIdx.x >> 5], sim, wsum);
            else
#endif
            simulate_rows_coop<TR, true, false>(g, vg, ps, pack2, stage_all[threadIdx.x >> 5], sim, wsum);
            done = true;
        }
        if (!alive) return;
    }
#endif
    if (!done) {
        auto tap = [&](int, float psf, bool ok, int v) {
            if (ok) {
                const float2 pm = __ldg(&pack2[v]);
                sim = fmaf(psf, pm.x, sim);
                wsum = fmaf(psf, pm.y, wsum);
            }
        };
        psf_rows_dispatch<TR>(g, vg, ps, tap, [](int) {});
    }
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
    const int mode = svr_window_simulate_available(c) ? c->tune_simulate : 0;
    if (mode == 1) return svr_launch_window_simulate(c, -1);
    const int skip = mode == 2 ? 1 : -1;
    if (c->flavor == 0)
        simulate_kernel<SvrTraits><<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
            c->n_valid, c->valid_idx, c->Nx, c->Nx * c->Ny, c->geom, c->vg, c->pack2, c->psf_sums, c->simslices, c->simweights,
            c->siminside, c->slice_inside, skip);
    else
        simulate_kernel<PvrTraits><<<divup_i(c->n_valid, 128), 128, 0, c->stream>>>(
            c->n_valid, c->valid_idx, c->Nx, c->Nx * c->Ny, c->geom, c->vg, c->pack2, c->psf_sums, c->simslices, c->simweights,
            c->siminside, c->slice_inside, skip);
    SVR_KERNEL_CHECK(c);
    if (mode == 2) return svr_launch_window_simulate(c, 1);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K3: SuperresolutionKernel3D_tex (reconstruction_cuda2.cu:408-522).
// One thread = one pixel pair (see scatter_pair); pair_idx[t] = linear index of the pair's even-x pixel.
template <class TR>
__global__ void __launch_bounds__(128, SVR_MINB_PAIR)
superres_scatter_kernel(uint32_t n_pairs, const uint32_t* __restrict__ pair_idx, int Nx, int P,
                        const float* __restrict__ slices, const float* __restrict__ weights,
                        const float* __restrict__ simslices, const float* __restrict__ slice_weights,
                        const float* __restrict__ scales, const SliceGeom* __restrict__ geom, VolGeom vg,
                        const float* __restrict__ psf_sums, float2* __restrict__ acc2, int skip_win)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const uint32_t ia = pair_idx[t];
    if (skip_win && geom[ia / (uint32_t)P].win_class) return;      // left to the warp-window kernel
    int k = 0, xA = 0, yA = 0, kB = 0, xB = 0, yB = 0;
    float awA = 0.f, cwA = 0.f, awB = 0.f, cwB = 0.f;
    const int xe = (int)(ia % (uint32_t)Nx);
    const bool liveA = superres_pixel<TR>(ia, Nx, P, slices, weights, simslices, slice_weights, scales, psf_sums, k, xA, yA, awA, cwA);
    const bool liveB = xe + 1 < Nx &&
                       superres_pixel<TR>(ia + 1, Nx, P, slices, weights, simslices, slice_weights, scales, psf_sums, kB, xB, yB, awB, cwB);
    if (!liveA && !liveB) return;
    if (!liveA) k = kB;
    const SliceGeom& g = geom[k];
    PixelSetup psA = {}, psB = {};
    if (liveA) psA = pixel_setup<TR>(g, vg, xA, yA);
    if (liveB) psB = pixel_setup<TR>(g, vg, xB, yB);
    const int mode = pair_mode(psA, psB, liveA, liveB);
    if (mode == 1) {
        bool d0 = false, d1 = false;
        const PairWeights w = { awA, cwA, awB, cwB };
        if (g.recur) scatter_pair<TR, true, false>(g, vg, psA, psB, liveA, liveB, w, acc2, nullptr, d0, d1);
        else scatter_pair<TR, false, false>(g, vg, psA, psB, liveA, liveB, w, acc2, nullptr, d0, d1);
    } else if (mode == 2) {
        if (liveA) superres_single<TR>(g, vg, psA, awA, cwA, acc2);
        if (liveB) superres_single<TR>(g, vg, psB, awB, cwB, acc2);
    }
}

int svr_launch_superres_scatter(svr_context* c)
{
    if (c->n_pairs == 0) return 0;
    ProfScope prof(c, 2);
    const bool avail = svr_window_scatter_available(c);
    if ((c->tune_scatter == 1 || c->tune_scatter == 2) && avail) return svr_launch_window_scatter(c, 0, -1);
    const int split = (c->tune_scatter == 3 && avail) ? 1 : 0;
    if (c->flavor == 0)
        superres_scatter_kernel<SvrTraits><<<divup_i(c->n_pairs, 128), 128, 0, c->stream>>>(
            c->n_pairs, c->pair_idx, c->Nx, c->Nx * c->Ny, c->slices, c->weights, c->simslices, c->slice_weights, c->scales,
            c->geom, c->vg, c->psf_sums, c->acc2, split);
    else
        superres_scatter_kernel<PvrTraits><<<divup_i(c->n_pairs, 128), 128, 0, c->stream>>>(
            c->n_pairs, c->pair_idx, c->Nx, c->Nx * c->Ny, c->slices, c->weights, c->simslices, c->slice_weights, c->scales,
            c->geom, c->vg, c->psf_sums, c->acc2, split);
    SVR_KERNEL_CHECK(c);
    if (split) return svr_launch_window_scatter(c, 0, 1);
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

// PVR reads the volume through a linear-filtered, border-mode 3D texture at UN-OFFSET normalised coordinates
// (reconVolume.cu:169-187, patchBasedSimulatePatches_gpu.cu:106): texel coordinate pos - 0.5, i.e. the mean of the
// 8 voxels pos + {-1,0}^3 with out-of-volume voxels reading 0.  The averaged volume is formed once per call here
// (the reference copies the whole volume into a cudaArray per call instead, patchBasedSimulatePatches_gpu.cu:135).
__global__ void pack_volume_tex_kernel(int vx, int vy, int vz, const float* __restrict__ recon,
                                       const unsigned char* __restrict__ mask, float2* __restrict__ pack2)
{
    const size_t V = (size_t)vx * vy * vz;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        if (!mask[v]) { pack2[v] = make_float2(0.f, 0.f); continue; }
        const int x = (int)(v % vx), y = (int)((v / vx) % vy), z = (int)(v / ((size_t)vx * vy));
        float s = 0.f;
        for (int dz = -1; dz <= 0; ++dz)
            for (int dy = -1; dy <= 0; ++dy)
                for (int dx = -1; dx <= 0; ++dx) {
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx >= 0 && yy >= 0 && zz >= 0) s += 0.125f * recon[xx + (size_t)yy * vx + (size_t)zz * vx * vy];
                }
        pack2[v] = make_float2(s, 1.0f);
    }
}

int svr_launch_pack_volume(svr_context* c)
{
    if (c->flavor == 1) {
        pack_volume_tex_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->vx, c->vy, c->vz, c->recon, c->mask_u8, c->pack2);
        SVR_KERNEL_CHECK(c);
        return 0;
    }
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

// PVR: P1 accumulates into recon / volWeights across stacks (ReconVolume::reset once per iteration,
// irtkPatchBasedReconstruction.cpp:492-497) and equalize is a separate call (reconVolume.cu:59-100).
__global__ void unpack_acc_kernel(size_t V, const float2* __restrict__ acc2, const unsigned char* __restrict__ mask,
                                  float* __restrict__ recon, float* __restrict__ volw)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        float2 a = acc2[v];
        if (!mask[v]) a = make_float2(0.f, 0.f);
        recon[v] = a.x;
        volw[v] = a.y;
    }
}
int svr_launch_unpack_acc(svr_context* c)
{
    unpack_acc_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->acc2, c->mask_u8, c->recon, c->volw);
    SVR_KERNEL_CHECK(c);
    return 0;
}
__global__ void equalize_inplace_kernel(size_t V, float* __restrict__ recon, const float* __restrict__ volw)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        const float a = recon[v], b = volw[v];
        recon[v] = (b != 0.f) ? a / b : a;
    }
}
int svr_launch_equalize_inplace(svr_context* c)
{
    equalize_inplace_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->recon, c->volw);
    SVR_KERNEL_CHECK(c);
    return 0;
}
