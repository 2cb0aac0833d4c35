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

// Window layout.  A window holds nz planes of ny rows of the volume; voxel (x, y, z) relative to the window origin lives
// at float2 index x + RS * y + PS * z.  The strides are padded per tile so that the 16 lanes of a half-warp (one
// shared-memory phase of a 64-bit access) fall into different 8-byte bank pairs: every tap instruction of the scatter is
// a read-modify-write of 512 B per warp, i.e. the shared-memory pipe is ~60 % busy when conflict-free and the bound of
// the kernel as soon as it is not (measured: oblique stacks at 3-4 way conflicts ran 1.7x slower than aligned ones).
// TMA-compatible layouts are dense boxes (RS even, PS = RS * by) drawn from a menu of tensor maps, one per box shape.
constexpr int WW_MENU_X0 = 12, WW_MENU_NX = 23;     // box widths 12, 14, ..., 56 voxels
constexpr int WW_MENU_N = 32;                       // box heights / depths 1 .. 32
__host__ __device__ __forceinline__ int ww_menu_index(int rs, int by, int bz)
{
    return (((rs - WW_MENU_X0) >> 1) * WW_MENU_N + (by - 1)) * WW_MENU_N + (bz - 1);
}

struct WinPlan {
    int S;                 // inner steps per flush (0: no window)
    int inner_z;           // 1: z is the inner tap-row axis, 0: y
    int RS, PS;            // row / plane stride (voxels)
    int nx, ny, nz;        // used extent (voxels): nx <= RS, ny * RS <= PS
    int tma;               // 1: dense box from the menu (PS = RS * by): the TMA unit can move it
    int map;               // menu index when tma
    int deg;               // bank-conflict degree of the layout (1 = conflict-free)
};

// Conflict degree of a layout: the largest number of lanes of one half-warp whose tap-0 address falls into the same
// bank pair.  (lx, ly, lz) = the lane's centre voxel relative to the tile's bounding box; the inner step and the tap index
// shift every lane's address by the same amount, so this is the degree of every tap instruction of the tile.
__device__ __forceinline__ int ww_conflict(bool use, int lx, int ly, int lz, int RS, int PS, int lane)
{
    const unsigned b = (unsigned)(lx + RS * ly + PS * lz) & 15u;
    const unsigned key = use ? (b | ((unsigned)(lane >> 4) << 4)) : (64u + (unsigned)lane);
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    return __reduce_max_sync(0xffffffffu, use ? __popc(grp) : 1);
}

// need_x: voxels per window row; ey, ez: extents of the tile's centre voxels.  want: 0 = any layout, 1 = TMA layouts only,
// 2 = TMA unless a general layout has fewer conflicts.
__device__ __forceinline__ WinPlan ww_plan(int need_x, int ey, int ez, int SUP, int want, bool use, int lx, int ly, int lz, int lane)
{
    WinPlan w; w.S = 0; w.inner_z = 0; w.RS = w.PS = 0; w.nx = need_x; w.ny = w.nz = 0; w.tma = 0; w.map = 0; w.deg = 0;
    const int nxe = (need_x + 1) & ~1;
    for (int S = SUP; S >= 1; S >>= 1) {
        const int volY = nxe * (ey + S - 1) * ez, volZ = nxe * ey * (ez + S - 1);
        const bool z = volZ < volY;
        if ((z ? volZ : volY) > WW_WIN_VOX) continue;
        w.S = S; w.inner_z = z ? 1 : 0;
        w.ny = z ? ey : ey + S - 1;
        w.nz = z ? ez + S - 1 : ez;
        break;
    }
    if (w.S == 0) return w;
    int best = 99;
    if (want >= 1 && w.nz <= WW_MENU_N) {          // dense boxes: widths nxe + 2k, heights ny + j
        for (int k = 0; k < 4 && best > 1; ++k) {
            const int RS = nxe + 2 * k;
            if (RS < WW_MENU_X0 || RS > WW_MENU_X0 + 2 * (WW_MENU_NX - 1)) continue;
            for (int j = 0; j < 4 && best > 1; ++j) {
                const int by = w.ny + j;
                if (by > WW_MENU_N || RS * by * w.nz > WW_WIN_VOX) break;
                const int d = ww_conflict(use, lx, ly, lz, RS, RS * by, lane);
                if (d < best) { best = d; w.RS = RS; w.PS = RS * by; w.tma = 1; w.map = ww_menu_index(RS, by, w.nz); }
            }
        }
        if (want == 1 || best <= 1) {
            if (best == 99) w.S = 0;
            w.deg = best;
            return w;
        }
    }
    for (int dr = 0; dr < 8 && best > 1; ++dr) {    // general strides
        const int RS = need_x + dr;
        for (int dp = 0; dp < 16 && best > 1; ++dp) {
            const int PS = RS * w.ny + dp;
            if (PS * w.nz > WW_WIN_VOX) break;
            const int d = ww_conflict(use, lx, ly, lz, RS, PS, lane);
            if (d < best) { best = d; w.RS = RS; w.PS = PS; w.tma = 0; }
        }
    }
    if (best == 99) w.S = 0;
    w.deg = best;
    return w;
}

// Plan statistics (svr_debug_get SVR_DBG_WW_STATS): [0..4] tiles by log2(S) (S = 1, 2, 4, 8, 16; S = 12, 6, 3 of the 12^3
// support count as 8, 4, 2), [5] tiles without a window, [8..15] tiles by conflict degree 1 .. 8+, [16] TMA layouts,
// [17] general layouts, [18] tiles with lanes sharing a centre voxel, [19] one-pixel fallbacks (pixels).
__device__ __forceinline__ void ww_count(unsigned int* stats, const WinPlan& w, int nrounds, int fallback_px, int lane)
{
    if (!stats || lane != 0) return;
    if (w.S == 0) atomicAdd(stats + 5, 1u);
    else {
        atomicAdd(stats + (31 - __clz(w.S)), 1u);
        atomicAdd(stats + 8 + min(max(w.deg, 1), 8) - 1, 1u);
        atomicAdd(stats + (w.tma ? 16 : 17), 1u);
        if (nrounds > 1) atomicAdd(stats + 18, 1u);
    }
    if (fallback_px) atomicAdd(stats + 19, (unsigned)fallback_px);
}

__device__ __forceinline__ uint32_t ww_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ww_magic(int d) { return d > 1 ? 0xffffffffu / (uint32_t)d + 1u : 0u; }   // q / d = umulhi(q, magic), q < 2^16
__device__ __forceinline__ int ww_div(int q, int d, uint32_t m) { return d > 1 ? (int)__umulhi((uint32_t)q, m) : q; }

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
    const int n4 = (w.PS * w.nz + 1) >> 1;
    for (int q = lane; q < n4; q += 32) w4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
}

// SIMT form: consecutive lanes read consecutive pieces of the used part of the window, clear them and send the non-zero
// ones as vector reductions: 16-byte pieces (two voxels) when the strides keep every row 16-byte aligned, 8-byte otherwise.
__device__ __forceinline__ void ww_flush_simt(float2* win, const WinPlan& w, float2* __restrict__ acc2, int vx, int vy,
                                              int X0, int Y0, int Z0, int lane)
{
    __syncwarp();
    if (((w.RS | w.PS) & 1) == 0) {
        const int hx = (w.nx + 1) >> 1;
        const int n = hx * w.ny * w.nz;
        const uint32_t m_hx = ww_magic(hx), m_ny = ww_magic(w.ny);
        for (int q = lane; q < n; q += 32) {
            const int row = ww_div(q, hx, m_hx), ch = q - row * hx;
            const int wz = ww_div(row, w.ny, m_ny), wy = row - wz * w.ny;
            float4* src = reinterpret_cast<float4*>(win + wz * w.PS + wy * w.RS) + ch;
            const float4 v = *src;
            if (v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
                *src = make_float4(0.f, 0.f, 0.f, 0.f);
                if (v.y + v.w > 0.0f)                      // denominators are sums of psf * c, c > 0 (NaNs are dropped, as red_row_paired does)
                    atomicAdd(reinterpret_cast<float4*>(acc2 + ((size_t)((Z0 + wz) * vy + (Y0 + wy)) * vx + X0)) + ch, v);
            }
        }
    } else {
        const int n = w.nx * w.ny * w.nz;
        const uint32_t m_nx = ww_magic(w.nx), m_ny = ww_magic(w.ny);
        for (int q = lane; q < n; q += 32) {
            const int row = ww_div(q, w.nx, m_nx), x = q - row * w.nx;
            const int wz = ww_div(row, w.ny, m_ny), wy = row - wz * w.ny;
            float2* src = win + wz * w.PS + wy * w.RS + x;
            const float2 v = *src;
            if (v.x != 0.0f || v.y != 0.0f) {
                *src = make_float2(0.f, 0.f);
                if (v.y > 0.0f) atomicAdd(acc2 + ((size_t)((Z0 + wz) * vy + (Y0 + wy)) * vx + X0 + x), v);
            }
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
    int only_class;                    // -1: every tile, else only tiles of slices with win_class == only_class
    unsigned int* stats;               // plan statistics (may be nullptr)
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
        w = ww_plan(x1 + (SUP - 1 - CEN) - X0 + 1, y1 - y0 + 1, z1 - z0 + 1, SUP, a.flush_tma ? 2 : 0, use,
                    ps.cx - CEN - X0, ps.cy - y0, ps.cz - z0, lane);
    }
    if (w.S == 0) use = false;
    {
        const unsigned fb = __ballot_sync(FULL, live && !use);
        if (usemask || fb) ww_count(a.stats, w, 1, __popc(fb), lane);
    }
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
        if (nrounds > 1 && a.stats && lane == 0) atomicAdd(a.stats + 18, 1u);
        const int vx = vg.vx, vy = vg.vy;
        if (MODE == 1 && use) {
            // the row through the pixel's own centre voxel almost always decides the mask flag (see scatter_pair)
            float p[SUP];
            psf_row_values<TR, RECUR>(g, ps.ex, ps.ey, ps.ez, p);
            const int v0 = (ps.cz * vy + ps.cy) * vx + ps.cx - CEN;
#pragma unroll
            for (int i = 0; i < SUP; ++i) if (p[i] != 0.0f && a.mask[v0 + i]) any = true;
        }
        const float bx1 = g.bx[1], by1 = g.by[1], bz1 = g.bz[1];
        const float bx2 = g.bx[2], by2 = g.by[2], bz2 = g.bz[2];
        const int strideI = w.inner_z ? w.PS : w.RS;
        const int lane_base = (ps.cz - z0) * w.PS + (ps.cy - y0) * w.RS + (ps.cx - CEN - X0);
#pragma unroll 1
        for (int o = 0; o < SUP; ++o) {
#pragma unroll 1
            for (int c0 = 0; c0 < SUP; c0 += w.S) {
#pragma unroll 1
                for (int s = 0; s < w.S; ++s) {
                    const int oy = (w.inner_z ? o : c0 + s) - CEN, oz = (w.inner_z ? c0 + s : o) - CEN;
                    float p[SUP];
                    if (use) {
                        // tap-row position in the operation order of psf_rows (z term first): the tap values, and with them the
                        // epsilon-skip decisions, are bit-identical to those of the other PSF kernels whatever the loop order
                        const float foz = (float)oz, foy = (float)oy;
                        const float zx = fmaf(foz, bx2, ps.ex), zy = fmaf(foz, by2, ps.ey), zz = fmaf(foz, bz2, ps.ez);
                        psf_row_values<TR, RECUR>(g, fmaf(foy, bx1, zx), fmaf(foy, by1, zy), fmaf(foy, bz1, zz), p);
                        if (MODE == 1 && !any) {
                            const int v0 = ((ps.cz + oz) * vy + ps.cy + oy) * vx + ps.cx - CEN;
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
                if (w.tma) ww_flush_tma(win, w, a.maps, X0, Y0, Z0, lane);
                else ww_flush_simt(win, w, a.acc2, vx, vy, X0, Y0, Z0, lane);
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
        if (a.only_class >= 0 && g.win_class != a.only_class) continue;
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
        w = ww_plan(x1 + (SUP - 1 - CEN) - X0 + 1, y1 - y0 + 1, z1 - z0 + 1, SUP, 1, use, ps.cx - CEN - X0, ps.cy - y0, ps.cz - z0, lane);
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
        const float bx1 = g.bx[1], by1 = g.by[1], bz1 = g.bz[1];
        const float bx2 = g.bx[2], by2 = g.by[2], bz2 = g.bz[2];
        const int strideI = w.inner_z ? w.PS : w.RS;
        const int lane_base = (ps.cz - z0) * w.PS + (ps.cy - y0) * w.RS + (ps.cx - CEN - X0);
        const uint32_t bytes = (uint32_t)(w.PS * w.nz) * 8u;
#pragma unroll 1
        for (int o = 0; o < SUP; ++o) {
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
                        const int oy = (w.inner_z ? o : c0 + s) - CEN, oz = (w.inner_z ? c0 + s : o) - CEN;
                        const float foz = (float)oz, foy = (float)oy;
                        const float zx = fmaf(foz, bx2, ps.ex), zy = fmaf(foz, by2, ps.ey), zz = fmaf(foz, bz2, ps.ez);
                        float p[SUP];
                        psf_row_values<TR, RECUR>(g, fmaf(foy, bx1, zx), fmaf(foy, by1, zy), fmaf(foy, bz1, zz), p);
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
    if (!c->ww_counter) {
        SVR_CUDA(c, cudaMalloc(&c->ww_counter, 40 * sizeof(unsigned int)));      // [0] tile queue, [8..39] plan statistics of the last scatter
        SVR_CUDA(c, cudaMemsetAsync(c->ww_counter, 0, 40 * sizeof(unsigned int), c->stream));
    }
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
    const int n = WW_MENU_NX * WW_MENU_N * WW_MENU_N;
    std::vector<CUtensorMap> host(n);
    memset(host.data(), 0, n * sizeof(CUtensorMap));
    const cuuint64_t gdim[3] = { (cuuint64_t)2 * c->vx, (cuuint64_t)c->vy, (cuuint64_t)c->vz };
    const cuuint64_t gstr[2] = { (cuuint64_t)8 * c->vx, (cuuint64_t)8 * c->vx * c->vy };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    for (int ix = 0; ix < WW_MENU_NX; ++ix)
        for (int by = 1; by <= WW_MENU_N; ++by)
            for (int bz = 1; bz <= WW_MENU_N; ++bz) {
                const int bx = WW_MENU_X0 + 2 * ix;
                if (bx * by * bz > WW_WIN_VOX) continue;
                const cuuint32_t box[3] = { (cuuint32_t)2 * bx, (cuuint32_t)by, (cuuint32_t)bz };
                const CUresult r = enc(&host[ww_menu_index(bx, by, bz)], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, gdim, gstr, box, estr,
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

int svr_launch_window_scatter(svr_context* c, int mode, int only_class)
{
    if (c->n_tiles == 0) return 0;
    WWScatterArgs a{};
    a.n_tiles = c->n_tiles; a.tile_idx = c->tile_idx; a.counter = c->ww_counter;
    a.Nx = c->Nx; a.Ny = c->Ny; a.P = c->Nx * c->Ny; a.tilesX = c->tilesX; a.tilesPerSlice = c->tilesPerSlice;
    a.slices = c->slices; a.weights = c->weights; a.simslices = c->simslices; a.slice_weights = c->slice_weights; a.scales = c->scales;
    a.psf_sums = c->psf_sums; a.geom = c->geom; a.acc2 = c->acc2; a.mask = c->mask_u8; a.voxel_flag = c->voxel_flag;
    a.slice_count = c->slice_count; a.maps = (const CUtensorMap*)c->maps_acc;
    a.flush_tma = (c->tune_scatter != 1 && c->maps_acc) ? 1 : 0;
    a.only_class = only_class;
    a.stats = c->ww_counter + 8;
    SVR_CUDA(c, cudaMemsetAsync(c->ww_counter, 0, 40 * sizeof(unsigned int), c->stream));
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
