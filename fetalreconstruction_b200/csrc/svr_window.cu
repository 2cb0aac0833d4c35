// svr_window.cu -- warp-window PSF kernels for sm_100a (round 2): the adjoint scatter (K1 pass 2, K3) and the
// forward projection (K2) with a warp-private sub-volume ("window") of the accumulator / the packed volume in
// shared memory, moved between shared memory and HBM by the TMA unit:
//
//   scatter:  UTMAREDG (cp.reduce.async.bulk.tensor.3d ... add.f32) flushes the window into the interleaved
//             {numerator, denominator} accumulator: one instruction instead of per-lane 128-bit reductions;
//   forward:  UTMALDG  (cp.async.bulk.tensor.3d + mbarrier) brings the window of pack2 into shared memory once per
//             16 tap-row steps; the taps then read shared memory.
//
// Why.  Round 1's scatter (svr_psf.cu: scatter_pair) sends every voxel-row of every pixel pair to L2 as vector
// reductions.  Its per-stack cost (7.7 - 14.2 ms per 128 C3 slices, by orientation) tracks the number of 32-byte sector
// updates the L2 has to perform (700 - 1300 per pixel), not the tap arithmetic (K2 does the same taps in 6.5 ms).
// Neighbouring pixels of a slice write almost the same voxels, so the sums are merged BEFORE they leave the SM:
// one warp takes a tile of 8 x 4 pixels of one slice and walks the 16 x 16 tap rows with one of the two row axes
// (volume y or z) innermost; all rows of the inner loop land in a window of (ex + 15) x ey x (ez + 15) voxels
// (ex, ey, ez = extent of the tile's 32 centre voxels) that lives in the warp's shared memory.  Within one tap
// instruction the 32 lanes touch 32 different voxels when their centre voxels differ (checked per tile; lanes that
// share a centre voxel take turns), so the accumulation is a plain shared-memory read-modify-write, not an atomic
// (float atomics on shared memory are CAS loops on this architecture).  After the inner loop the window is added to
// the accumulator and cleared.  Sector updates per pixel drop to ~60 (aligned stacks) / ~190 (through-plane stacks).
// When the window of a full inner loop does not fit (oblique tiles), the inner loop is flushed in chunks of
// S = 8, 4, 2, 1 steps; tiles that touch the volume faces (reference quirk Q4 needs the clamped indices) or whose
// window cannot be formed fall back to the one-pixel paths of svr_psf_pixel.cuh.
//
// Exactness: the tap values, epsilon-skip chains and the set of (voxel, value) contributions are those of the
// reference kernels (reconstruction_cuda2.cu:176-295, 298-404, 408-522); only the order in which floats are summed
// differs (as with any atomic scatter, quirk Q8; for K2: oy/oz loop order).
#include <cuda.h>                     // CUtensorMap and its enums (types only: the encoder comes from cudaGetDriverEntryPoint)
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "svr_context.h"
#include "svr_psf_pixel.cuh"

constexpr int WW_TW = 8, WW_TH = 4;          // pixel tile of one warp
constexpr int WW_WARPS = 4;                  // warps per CTA (independent: no CTA-wide barrier)
constexpr int WW_WIN_VOX = 1536;             // float2 voxels of window per warp (12 KB)
constexpr int WW_SMEM = WW_WARPS * WW_WIN_VOX * 8 + WW_WARPS * 8;     // windows + one mbarrier per warp

// Window dimensions are quantised so that a fixed menu of TMA tensor maps (box shape = window shape) covers them.
constexpr int WW_NQX = 10, WW_NQ = 19;
__host__ __device__ __forceinline__ int ww_qx_val(int i)
{
    return i <= 5 ? 18 + 2 * i : (i == 6 ? 32 : (i == 7 ? 36 : (i == 8 ? 40 : 48)));
}
__host__ __device__ __forceinline__ int ww_q_val(int i) { return i < 8 ? i + 1 : (i < 18 ? 10 + 2 * (i - 8) : 32); }
__device__ __forceinline__ int ww_qx_idx(int n)
{   // smallest menu entry >= n, -1 when none
    if (n <= 18) return 0;
    if (n <= 28) return (n - 18 + 1) >> 1;
    if (n <= 32) return 6;
    if (n <= 36) return 7;
    if (n <= 40) return 8;
    if (n <= 48) return 9;
    return -1;
}
__device__ __forceinline__ int ww_q_idx(int n)
{
    if (n <= 8) return n - 1;
    if (n <= 28) return 8 + ((n - 10 + 1) >> 1);
    if (n <= 32) return 18;
    return -1;
}

struct WinPlan {
    int S;                 // inner steps per flush (0: no window)
    int inner_z;           // 1: z is the inner tap-row axis, 0: y
    int bx, by, bz;        // window layout (voxels): row stride, rows per plane, planes
    int map;               // index into the tensor-map menu
};

// ex, ey, ez: extents of the tile's centre voxels; need_x = ex + SUP - 1 (+1 when the origin was rounded down to even).
__device__ __forceinline__ WinPlan ww_plan(int need_x, int ey, int ez, int SUP)
{
    WinPlan w; w.S = 0; w.inner_z = 0; w.bx = w.by = w.bz = 0; w.map = 0;
    const int ix = ww_qx_idx(need_x);
    if (ix < 0) return w;
    const int bx = ww_qx_val(ix);
    for (int S = SUP; S >= 1; S >>= 1) {
        const int iyY = ww_q_idx(ey + S - 1), izY = ww_q_idx(ez);         // inner = y
        const int iyZ = ww_q_idx(ey), izZ = ww_q_idx(ez + S - 1);         // inner = z
        const int volY = (iyY < 0 || izY < 0) ? (1 << 30) : bx * ww_q_val(iyY) * ww_q_val(izY);
        const int volZ = (iyZ < 0 || izZ < 0) ? (1 << 30) : bx * ww_q_val(iyZ) * ww_q_val(izZ);
        const bool z = volZ < volY;
        const int vol = z ? volZ : volY;
        if (vol <= WW_WIN_VOX) {
            const int iy = z ? iyZ : iyY, iz = z ? izZ : izY;
            w.S = S; w.inner_z = z ? 1 : 0; w.bx = bx; w.by = ww_q_val(iy); w.bz = ww_q_val(iz);
            w.map = (ix * WW_NQ + iy) * WW_NQ + iz;
            return w;
        }
    }
    return w;
}

__device__ __forceinline__ uint32_t ww_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------------------------------------
// Window -> accumulator.
// TMA form: one UTMAREDG adds the whole window box at volume coordinates (X0, Y0, Z0); parts of the box outside the
// volume are clipped by the unit.  The window is cleared once the unit has read it.
__device__ __forceinline__ void ww_flush_tma(float2* win, const WinPlan& w, const CUtensorMap* maps, int X0, int Y0, int Z0, int lane)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // the lanes' generic-proxy writes -> visible to the TMA unit
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                     :: "l"(maps + w.map), "r"(2 * X0), "r"(Y0), "r"(Z0), "r"(ww_smem_u32(win)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory may be reused (the adds may still be in flight)
    }
    __syncwarp();
    float4* w4 = reinterpret_cast<float4*>(win);
    const int n4 = (w.bx * w.by * w.bz) >> 1;
    for (int q = lane; q < n4; q += 32) w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
}

// SIMT form: consecutive lanes read consecutive 16-byte chunks (two voxels) of the window, clear them and send the
// non-zero ones as 128-bit reductions.  used_z planes of by rows each carry data.
__device__ __forceinline__ void ww_flush_simt(float2* win, const WinPlan& w, float2* __restrict__ acc2, int vx, int vy,
                                              int X0, int Y0, int Z0, int used_z, int lane)
{
    __syncwarp();
    float4* w4 = reinterpret_cast<float4*>(win);
    const int hb = w.bx >> 1;
    const int n4 = hb * w.by * used_z;
    const uint32_t m_hb = 0xffffffffu / (uint32_t)hb + 1u, m_by = 0xffffffffu / (uint32_t)w.by + 1u;   // exact for q < 2^16
    for (int q = lane; q < n4; q += 32) {
        const float4 v = w4[q];
        if (v.y + v.w > 0.0f) {                          // denominators are sums of psf * c with c > 0
            w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row = (int)__umulhi((uint32_t)q, m_hb), ch = q - row * hb;
            const int wz = (int)__umulhi((uint32_t)row, m_by), wy = row - wz * w.by;
            float4* dst = reinterpret_cast<float4*>(acc2 + ((size_t)((Z0 + wz) * vy + (Y0 + wy)) * vx + X0)) + ch;
            atomicAdd(dst, v);
        } else if (v.x != 0.0f || v.z != 0.0f || v.y != 0.0f || v.w != 0.0f) {
            w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);     // NaN / non-positive denominators are dropped like red_row_paired does
        }
    }
    __syncwarp();
}

struct WWScatterArgs {
    uint32_t n_tiles;
    const uint32_t* tile_idx;
    unsigned int* counter;             // dynamic tile queue
    int Nx, Ny, P, tilesX, tilesPerSlice;
    const float* slices; const float* weights; const float* simslices; const float* slice_weights; const float* scales;
    const float* psf_sums;
    const SliceGeom* geom;
    float2* acc2;
    const unsigned char* mask;
    unsigned char* voxel_flag;
    int* slice_count;
    const CUtensorMap* maps;           // menu over acc2 (nullptr: SIMT flush only)
    int flush_tma;
};

// One tile.  MODE 0: K3 (superresolution), MODE 1: K1 pass 2 (Gaussian reconstruction, with the per-pixel mask flag).
template <class TR, int MODE, bool RECUR>
__device__ __forceinline__ void ww_scatter_tile(const WWScatterArgs& a, const VolGeom& vg, const SliceGeom& g, int k, int x, int y,
                                                bool inb, float2* win, int lane)
{
    constexpr int SUP = TR::SUP, CEN = TR::CEN;
    const unsigned FULL = 0xffffffffu;
    const uint32_t idx = (uint32_t)k * (uint32_t)a.P + (uint32_t)(y * a.Nx + x);
    float aw = 0.f, cw = 0.f;
    bool live = false;
    if (inb) {
        if (MODE == 0) {
            int kk, xx, yy;
            live = superres_pixel<TR>(idx, a.Nx, a.P, a.slices, a.weights, a.simslices, a.slice_weights, a.scales, a.psf_sums, kk, xx, yy, aw, cw);
        } else if (a.voxel_flag[idx] == 2) {
            live = true;
            cw = 1.0f / a.psf_sums[idx];
            aw = a.slices[idx] * a.scales[k] * cw;
        }
    }
    PixelSetup ps = {};
    if (live) ps = pixel_setup<TR>(g, vg, x, y);
    bool any = false;                                        // MODE 1: an accepted tap landed on a masked voxel
    // pixels whose support touches the volume faces: one-pixel path with the reference's clamped indices (quirk Q4)
    bool use = live && ps.interior;
    const unsigned usemask = __ballot_sync(FULL, use);
    WinPlan w; w.S = 0;
    int x0 = 0, y0 = 0, z0 = 0, X0 = 0;
    if (usemask) {
        x0 = __reduce_min_sync(FULL, use ? ps.cx : 0x7fffffff);
        y0 = __reduce_min_sync(FULL, use ? ps.cy : 0x7fffffff);
        z0 = __reduce_min_sync(FULL, use ? ps.cz : 0x7fffffff);
        const int x1 = __reduce_max_sync(FULL, use ? ps.cx : -0x7fffffff);
        const int y1 = __reduce_max_sync(FULL, use ? ps.cy : -0x7fffffff);
        const int z1 = __reduce_max_sync(FULL, use ? ps.cz : -0x7fffffff);
        X0 = (x0 - CEN) & ~1;                                // even: 16-byte aligned rows in the accumulator (vx is even)
        w = ww_plan(x1 + (SUP - 1 - CEN) - X0 + 1, y1 - y0 + 1, z1 - z0 + 1, SUP);
    }
    if (w.S == 0) use = false;
    if (live && !use) {
        if (MODE == 0) superres_single<TR>(g, vg, ps, aw, cw, a.acc2);
        else any = gaussian_single<TR>(g, vg, ps, aw, cw, a.mask, a.acc2);
    }
    if (w.S != 0) {
        // lanes sharing a centre voxel would collide inside one tap instruction: they take turns
        const unsigned key = use ? (unsigned)((ps.cx - x0) | ((ps.cy - y0) << 8) | ((ps.cz - z0) << 16)) : (0x80000000u | (unsigned)lane);
        const unsigned grp = __match_any_sync(FULL, key);
        const int rank = __popc(grp & ((1u << lane) - 1u));
        const int nrounds = __reduce_max_sync(FULL, rank) + 1;
        const int vx = vg.vx, vy = vg.vy;
        if (MODE == 1 && use) {
            // the row through the pixel's own centre voxel almost always decides the mask flag (see scatter_pair)
            float p[SUP];
            psf_row_values<TR, RECUR>(g, ps.ex, ps.ey, ps.ez, p);
            const int v0 = (ps.cz * vy + ps.cy) * vx + ps.cx - CEN;
#pragma unroll
            for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && a.mask[v0 + i]) any = true;
        }
        // outer / inner tap-row axes
        const float bOx = w.inner_z ? g.bx[1] : g.bx[2], bOy = w.inner_z ? g.by[1] : g.by[2], bOz = w.inner_z ? g.bz[1] : g.bz[2];
        const float bIx = w.inner_z ? g.bx[2] : g.bx[1], bIy = w.inner_z ? g.by[2] : g.by[1], bIz = w.inner_z ? g.bz[2] : g.bz[1];
        const int strideI = w.inner_z ? w.by * w.bx : w.bx;
        const int lane_base = ((ps.cz - z0) * w.by + (ps.cy - y0)) * w.bx + (ps.cx - CEN - X0);
#pragma unroll 1
        for (int o = 0; o < SUP; ++o) {
            const float fo = (float)(o - CEN);
            const float ox = fmaf(fo, bOx, ps.ex), oy = fmaf(fo, bOy, ps.ey), oz = fmaf(fo, bOz, ps.ez);
#pragma unroll 1
            for (int c0 = 0; c0 < SUP; c0 += w.S) {
#pragma unroll 1
                for (int s = 0; s < w.S; ++s) {
                    const float fi = (float)(c0 + s - CEN);
                    float p[SUP];
                    if (use) {
                        psf_row_values<TR, RECUR>(g, fmaf(fi, bIx, ox), fmaf(fi, bIy, oy), fmaf(fi, bIz, oz), p);
                        if (MODE == 1 && !any) {
                            const int yy = ps.cy + (w.inner_z ? o : c0 + s) - CEN, zz = ps.cz + (w.inner_z ? c0 + s : o) - CEN;
                            const int v0 = (zz * vy + yy) * vx + ps.cx - CEN;
#pragma unroll
                            for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && a.mask[v0 + i]) any = true;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < SUP; ++i) p[i] = 0.f;
                    }
                    float2* wp = win + lane_base + s * strideI;
#pragma unroll 1
                    for (int r = 0; r < nrounds; ++r) {
                        const bool doit = use && rank == r;
#pragma unroll
                        for (int i = 0; i < SUP; ++i) {
                            if (doit) {
                                float2 v = wp[i];
                                v.x = fmaf(p[i], aw, v.x);
                                v.y = fmaf(p[i], cw, v.y);
                                wp[i] = v;
                            }
                            __syncwarp();                    // orders the lanes' shared-memory accesses tap by tap
                        }
                    }
                }
                // window origin in the volume for this (outer offset, chunk)
                const int Y0 = w.inner_z ? y0 + o - CEN : y0 - CEN + c0;
                const int Z0 = w.inner_z ? z0 - CEN + c0 : z0 + o - CEN;
                if (a.flush_tma) ww_flush_tma(win, w, a.maps, X0, Y0, Z0, lane);
                else ww_flush_simt(win, w, a.acc2, vx, vy, X0, Y0, Z0, w.bz, lane);
            }
        }
    }
    if (MODE == 1) {
        if (live) a.voxel_flag[idx] = any ? 1 : 0;
        const unsigned anymask = __ballot_sync(FULL, live && any);
        if (lane == 0 && anymask) atomicAdd(&a.slice_count[k], __popc(anymask));
    }
}

#ifndef WW_MINB
#define WW_MINB 4
#endif

template <class TR, int MODE>
__global__ void __launch_bounds__(32 * WW_WARPS, WW_MINB)
window_scatter_kernel(const __grid_constant__ WWScatterArgs a, const __grid_constant__ VolGeom vg)
{
    extern __shared__ __align__(128) unsigned char ww_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* win = reinterpret_cast<float2*>(ww_smem) + warp * WW_WIN_VOX;
    {
        float4* w4 = reinterpret_cast<float4*>(win);
        for (int q = lane; q < WW_WIN_VOX / 2; q += 32) w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
    }
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= a.n_tiles) break;
        const uint32_t tile = a.tile_idx[t];
        const int k = (int)(tile / (uint32_t)a.tilesPerSlice);
        const int rem = (int)(tile - (uint32_t)k * (uint32_t)a.tilesPerSlice);
        const int ty = rem / a.tilesX, tx = rem - ty * a.tilesX;
        const int x = tx * WW_TW + (lane & (WW_TW - 1)), y = ty * WW_TH + (lane / WW_TW);
        const bool inb = x < a.Nx && y < a.Ny;
        const SliceGeom& g = a.geom[k];
        if (g.recur) ww_scatter_tile<TR, MODE, true>(a, vg, g, k, x, y, inb, win, lane);
        else ww_scatter_tile<TR, MODE, false>(a, vg, g, k, x, y, inb, win, lane);
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");           // the unit's adds are complete before the CTA retires
}

// ---------------------------------------------------------------------------------------------
// K2 with the window of pack2 = {recon * m, m} staged by the TMA unit.
struct WWSimArgs {
    uint32_t n_tiles;
    const uint32_t* tile_idx;
    unsigned int* counter;
    int Nx, Ny, P, tilesX, tilesPerSlice;
    const float* slices;
    const float* psf_sums;
    const SliceGeom* geom;
    const float2* pack2;
    const CUtensorMap* maps;           // menu over pack2
    float* simslices; float* simweights; unsigned char* siminside; int* slice_inside;
    int only_class;                    // -1: every tile; 0 / 1: only tiles of slices with through_plane_rows == only_class ... see launcher
};

template <class TR, bool RECUR>
__device__ __forceinline__ void ww_simulate_tile(const WWSimArgs& a, const VolGeom& vg, const SliceGeom& g, int k, int x, int y, bool inb,
                                                 float2* win, uint64_t* bar, uint32_t& phase, int lane)
{
    constexpr int SUP = TR::SUP, CEN = TR::CEN;
    const unsigned FULL = 0xffffffffu;
    const uint32_t idx = (uint32_t)k * (uint32_t)a.P + (uint32_t)(y * a.Nx + x);
    float sume = 0.f;
    if (inb && a.slices[idx] != -1.0f) sume = a.psf_sums[idx];
    const bool live = sume != 0.0f;
    PixelSetup ps = {};
    if (live) ps = pixel_setup<TR>(g, vg, x, y);
    bool use = live && ps.interior;
    const unsigned usemask = __ballot_sync(FULL, use);
    WinPlan w; w.S = 0;
    int x0 = 0, y0 = 0, z0 = 0, X0 = 0;
    if (usemask) {
        x0 = __reduce_min_sync(FULL, use ? ps.cx : 0x7fffffff);
        y0 = __reduce_min_sync(FULL, use ? ps.cy : 0x7fffffff);
        z0 = __reduce_min_sync(FULL, use ? ps.cz : 0x7fffffff);
        const int x1 = __reduce_max_sync(FULL, use ? ps.cx : -0x7fffffff);
        const int y1 = __reduce_max_sync(FULL, use ? ps.cy : -0x7fffffff);
        const int z1 = __reduce_max_sync(FULL, use ? ps.cz : -0x7fffffff);
        X0 = (x0 - CEN) & ~1;
        w = ww_plan(x1 + (SUP - 1 - CEN) - X0 + 1, y1 - y0 + 1, z1 - z0 + 1, SUP);
    }
    if (w.S == 0) use = false;
    float sim = 0.f, wsum = 0.f;
    if (live && !use) {
        auto tap = [&](int, float psf, bool ok, int v) {
            if (ok) {
                const float2 pm = __ldg(&a.pack2[v]);
                sim = fmaf(psf, pm.x, sim);
                wsum = fmaf(psf, pm.y, wsum);
            }
        };
        psf_rows_dispatch<TR>(g, vg, ps, tap, [](int) {});
    }
    if (w.S != 0) {
        const float bOx = w.inner_z ? g.bx[1] : g.bx[2], bOy = w.inner_z ? g.by[1] : g.by[2], bOz = w.inner_z ? g.bz[1] : g.bz[2];
        const float bIx = w.inner_z ? g.bx[2] : g.bx[1], bIy = w.inner_z ? g.by[2] : g.by[1], bIz = w.inner_z ? g.bz[2] : g.bz[1];
        const int strideI = w.inner_z ? w.by * w.bx : w.bx;
        const int lane_base = ((ps.cz - z0) * w.by + (ps.cy - y0)) * w.bx + (ps.cx - CEN - X0);
        const uint32_t bytes = (uint32_t)(w.bx * w.by * w.bz) * 8u;
#pragma unroll 1
        for (int o = 0; o < SUP; ++o) {
            const float fo = (float)(o - CEN);
            const float ox = fmaf(fo, bOx, ps.ex), oy = fmaf(fo, bOy, ps.ey), oz = fmaf(fo, bOz, ps.ez);
#pragma unroll 1
            for (int c0 = 0; c0 < SUP; c0 += w.S) {
                const int Y0 = w.inner_z ? y0 + o - CEN : y0 - CEN + c0;
                const int Z0 = w.inner_z ? z0 - CEN + c0 : z0 + o - CEN;
                __syncwarp();                                   // every lane is done with the previous window
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ww_smem_u32(bar)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 :: "r"(ww_smem_u32(win)), "l"(a.maps + w.map), "r"(2 * X0), "r"(Y0), "r"(Z0), "r"(ww_smem_u32(bar)) : "memory");
                }
                uint32_t ok = 0, spins = 0;
                while (!ok) {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(ww_smem_u32(bar)), "r"(phase) : "memory");
                    if (!ok && ++spins > (1u << 22)) __trap();      // a lost TMA transaction must not hang the device
                }
                phase ^= 1u;
                if (use) {
#pragma unroll 1
                    for (int s = 0; s < w.S; ++s) {
                        const float fi = (float)(c0 + s - CEN);
                        float p[SUP];
                        psf_row_values<TR, RECUR>(g, fmaf(fi, bIx, ox), fmaf(fi, bIy, oy), fmaf(fi, bIz, oz), p);
                        const float2* wp = win + lane_base + s * strideI;
#pragma unroll
                        for (int i = 0; i < SUP; ++i) {
                            const float2 pm = wp[i];
                            sim = fmaf(p[i], pm.x, sim);
                            wsum = fmaf(p[i], pm.y, wsum);
                        }
                    }
                }
            }
        }
    }
    if (live) {
        const float weight = wsum / sume;
        if (weight > 0.f) {
            a.simslices[idx] = sim / wsum;
            a.simweights[idx] = weight;
            a.siminside[idx] = 1;
            a.slice_inside[k] = 1;
        }
    }
}

template <class TR>
__global__ void __launch_bounds__(32 * WW_WARPS, WW_MINB)
window_simulate_kernel(const __grid_constant__ WWSimArgs a, const __grid_constant__ VolGeom vg)
{
    extern __shared__ __align__(128) unsigned char ww_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* win = reinterpret_cast<float2*>(ww_smem) + warp * WW_WIN_VOX;
    uint64_t* bar = reinterpret_cast<uint64_t*>(ww_smem + WW_WARPS * WW_WIN_VOX * 8) + warp;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ww_smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(a.counter, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= a.n_tiles) break;
        const uint32_t tile = a.tile_idx[t];
        const int k = (int)(tile / (uint32_t)a.tilesPerSlice);
        const SliceGeom& g = a.geom[k];
        if (a.only_class >= 0 && g.through_plane_rows != a.only_class) continue;
        const int rem = (int)(tile - (uint32_t)k * (uint32_t)a.tilesPerSlice);
        const int ty = rem / a.tilesX, tx = rem - ty * a.tilesX;
        const int x = tx * WW_TW + (lane & (WW_TW - 1)), y = ty * WW_TH + (lane / WW_TW);
        const bool inb = x < a.Nx && y < a.Ny;
        if (g.recur) ww_simulate_tile<TR, true>(a, vg, g, k, x, y, inb, win, bar, phase, lane);
        else ww_simulate_tile<TR, false>(a, vg, g, k, x, y, inb, win, bar, phase, lane);
    }
}

// ---------------------------------------------------------------------------------------------
// Host side: tile list, tensor-map menus, launchers.
struct TileHead {
    const float* s;
    int Nx, Ny, tilesX, tilesPerSlice, P;
    __device__ bool operator()(uint32_t tile) const
    {
        const int k = (int)(tile / (uint32_t)tilesPerSlice);
        const int rem = (int)(tile - (uint32_t)k * (uint32_t)tilesPerSlice);
        const int ty = rem / tilesX, tx = rem - ty * tilesX;
        const float* base = s + (size_t)k * P;
        for (int j = 0; j < WW_TH; ++j) {
            const int y = ty * WW_TH + j;
            if (y >= Ny) break;
            for (int i = 0; i < WW_TW; ++i) {
                const int x = tx * WW_TW + i;
                if (x < Nx && base[y * Nx + x] != -1.0f) return true;
            }
        }
        return false;
    }
};

int svr_window_build_tiles(svr_context* c)
{
    c->tilesX = divup_i(c->Nx, WW_TW);
    c->tilesPerSlice = c->tilesX * divup_i(c->Ny, WW_TH);
    const size_t n = (size_t)c->tilesPerSlice * (size_t)c->S;
    c->n_tiles = 0;
    if (n == 0) return 0;
    if (c->tile_cap < n) {
        if (c->tile_idx) cudaFree(c->tile_idx);
        c->tile_idx = nullptr;
        SVR_CUDA(c, cudaMalloc(&c->tile_idx, n * sizeof(uint32_t)));
        c->tile_cap = n;
    }
    if (!c->ww_counter) SVR_CUDA(c, cudaMalloc(&c->ww_counter, sizeof(unsigned int)));
    thrust::counting_iterator<uint32_t> it(0);
    TileHead pred{ c->slices, c->Nx, c->Ny, c->tilesX, c->tilesPerSlice, c->Nx * c->Ny };
    int* d_num = (int*)(c->partials);
    size_t need = 0;
    SVR_CUDA(c, cub::DeviceSelect::If(nullptr, need, it, c->tile_idx, d_num, (int)n, pred, c->stream));
    if (need > c->cub_tmp_bytes) {
        if (c->cub_tmp) cudaFree(c->cub_tmp);
        SVR_CUDA(c, cudaMalloc(&c->cub_tmp, need));
        c->cub_tmp_bytes = need;
    }
    SVR_CUDA(c, cub::DeviceSelect::If(c->cub_tmp, need, it, c->tile_idx, d_num, (int)n, pred, c->stream));
    c->launches += 2;
    int h = 0;
    SVR_CUDA(c, cudaMemcpyAsync(&h, d_num, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_tiles = (uint32_t)h;
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// One menu = a tensor map per window shape over an interleaved float2 volume, viewed as float[vz][vy][2 vx].
static int build_menu(svr_context* c, void* base, CUtensorMap** dev_menu)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (c->vx & 1)) {                 // rows of the float2 volume must be 16-byte multiples for the unit
        if (*dev_menu) { cudaFree(*dev_menu); *dev_menu = nullptr; }
        return 0;
    }
    const int n = WW_NQX * WW_NQ * WW_NQ;
    std::vector<CUtensorMap> host(n);
    memset(host.data(), 0, n * sizeof(CUtensorMap));
    const cuuint64_t gdim[3] = { (cuuint64_t)2 * c->vx, (cuuint64_t)c->vy, (cuuint64_t)c->vz };
    const cuuint64_t gstr[2] = { (cuuint64_t)8 * c->vx, (cuuint64_t)8 * c->vx * c->vy };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    for (int ix = 0; ix < WW_NQX; ++ix)
        for (int iy = 0; iy < WW_NQ; ++iy)
            for (int iz = 0; iz < WW_NQ; ++iz) {
                const int bx = ww_qx_val(ix), by = ww_q_val(iy), bz = ww_q_val(iz);
                if (bx * by * bz > WW_WIN_VOX) continue;
                const cuuint32_t box[3] = { (cuuint32_t)2 * bx, (cuuint32_t)by, (cuuint32_t)bz };
                const CUresult r = enc(&host[(ix * WW_NQ + iy) * WW_NQ + iz], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstr, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) {
                    char buf[160];
                    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) for window %dx%dx%d", (int)r, bx, by, bz);
                    c->err = buf;
                    return 2;
                }
            }
    if (!*dev_menu) SVR_CUDA(c, cudaMalloc((void**)dev_menu, n * sizeof(CUtensorMap)));
    SVR_CUDA(c, cudaMemcpyAsync(*dev_menu, host.data(), n * sizeof(CUtensorMap), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_window_build_maps(svr_context* c)
{
    if (int r = build_menu(c, c->acc2, (CUtensorMap**)&c->maps_acc)) return r;
    return build_menu(c, c->pack2, (CUtensorMap**)&c->maps_pack);
}

void svr_window_free(svr_context* c)
{
    if (c->tile_idx) cudaFree(c->tile_idx);
    if (c->ww_counter) cudaFree(c->ww_counter);
    if (c->maps_acc) cudaFree(c->maps_acc);
    if (c->maps_pack) cudaFree(c->maps_pack);
    c->tile_idx = nullptr; c->ww_counter = nullptr; c->maps_acc = nullptr; c->maps_pack = nullptr;
}

bool svr_window_scatter_available(const svr_context* c) { return c->n_tiles > 0 && (c->vx & 1) == 0; }
bool svr_window_simulate_available(const svr_context* c) { return c->n_tiles > 0 && c->maps_pack != nullptr; }

template <class K>
static int ww_config(svr_context* c, K kernel)
{
    SVR_CUDA(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WW_SMEM));
    return 0;
}

static int ww_grid(const svr_context* c, uint32_t n_tiles)
{
    const long long want = (long long)c->sm_count * WW_MINB;
    const long long need = ((long long)n_tiles + WW_WARPS - 1) / WW_WARPS;
    return (int)(need < want ? need : want);
}

int svr_launch_window_scatter(svr_context* c, int mode)
{
    if (c->n_tiles == 0) return 0;
    WWScatterArgs a{};
    a.n_tiles = c->n_tiles; a.tile_idx = c->tile_idx; a.counter = c->ww_counter;
    a.Nx = c->Nx; a.Ny = c->Ny; a.P = c->Nx * c->Ny; a.tilesX = c->tilesX; a.tilesPerSlice = c->tilesPerSlice;
    a.slices = c->slices; a.weights = c->weights; a.simslices = c->simslices; a.slice_weights = c->slice_weights; a.scales = c->scales;
    a.psf_sums = c->psf_sums; a.geom = c->geom; a.acc2 = c->acc2; a.mask = c->mask_u8; a.voxel_flag = c->voxel_flag;
    a.slice_count = c->slice_count; a.maps = (const CUtensorMap*)c->maps_acc;
    a.flush_tma = (c->tune_scatter == 2 && c->maps_acc) ? 1 : 0;
    SVR_CUDA(c, cudaMemsetAsync(c->ww_counter, 0, sizeof(unsigned int), c->stream));
    const int grid = ww_grid(c, c->n_tiles);
    if (c->flavor == 0) {
        if (mode == 0) {
            if (ww_config(c, window_scatter_kernel<SvrTraits, 0>)) return 1;
            window_scatter_kernel<SvrTraits, 0><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
        } else {
            if (ww_config(c, window_scatter_kernel<SvrTraits, 1>)) return 1;
            window_scatter_kernel<SvrTraits, 1><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
        }
    } else {
        if (mode == 0) {
            if (ww_config(c, window_scatter_kernel<PvrTraits, 0>)) return 1;
            window_scatter_kernel<PvrTraits, 0><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
        } else {
            if (ww_config(c, window_scatter_kernel<PvrTraits, 1>)) return 1;
            window_scatter_kernel<PvrTraits, 1><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
        }
    }
    SVR_KERNEL_CHECK(c);
    return 0;
}

int svr_launch_window_simulate(svr_context* c, int only_class)
{
    if (c->n_tiles == 0) return 0;
    WWSimArgs a{};
    a.n_tiles = c->n_tiles; a.tile_idx = c->tile_idx; a.counter = c->ww_counter;
    a.Nx = c->Nx; a.Ny = c->Ny; a.P = c->Nx * c->Ny; a.tilesX = c->tilesX; a.tilesPerSlice = c->tilesPerSlice;
    a.slices = c->slices; a.psf_sums = c->psf_sums; a.geom = c->geom; a.pack2 = c->pack2; a.maps = (const CUtensorMap*)c->maps_pack;
    a.simslices = c->simslices; a.simweights = c->simweights; a.siminside = c->siminside; a.slice_inside = c->slice_inside;
    a.only_class = only_class;
    SVR_CUDA(c, cudaMemsetAsync(c->ww_counter, 0, sizeof(unsigned int), c->stream));
    const int grid = ww_grid(c, c->n_tiles);
    if (c->flavor == 0) {
        if (ww_config(c, window_simulate_kernel<SvrTraits>)) return 1;
        window_simulate_kernel<SvrTraits><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
    } else {
        if (ww_config(c, window_simulate_kernel<PvrTraits>)) return 1;
        window_simulate_kernel<PvrTraits><<<grid, 32 * WW_WARPS, WW_SMEM, c->stream>>>(a, c->vg);
    }
    SVR_KERNEL_CHECK(c);
    return 0;
}
