// svr_reg.cu -- batched 6-DOF slice-to-volume registration on blurred NCC (--useGPUReg) for sm_100a.
//
// Replaces Reconstruction::{initRegStorageVolumes, FillRegSlices, updateResampledSlicesI2W,
// prepareSliceToVolumeReg, registerSlicesToVolume} (reconstruction_cuda2.cu:3800-3959, 4001-4575,
// 4707-5088) and GPUGauss/gaussfilter.cu.
//
// One cost evaluation in the reference = per in-slice offset: resample kernel -> 2 Gauss kernels ->
// averageIf -> memset -> NCC kernel -> addNcc, i.e. 19 launches and 5 full passes over a materialised
// slice stack (~24 B per pixel per offset).  Here it is ONE kernel over (tile, active slice, offset):
//   sample the volume at a 32x32 pixel tile plus the blur halo through the TEXTURE UNIT (3D cudaArray, linear
//   filter, border addressing, normalised coordinates without the half-texel offset: the reference's own fetch,
//   so the samples are the reference's bit for bit up to the float position) -> separable Gauss in shared memory
//   with the reference's padding rules -> raw moments {n, Sa, Sb, Sab, Saa, Sbb} over the NCC domain plus {count,
//   sum} of the sampled slice, accumulated in double, warp-shuffle + block reduce, one slot of 8 doubles per CTA
//   summed in a fixed order by the finishing kernel (bit-reproducible similarities).
// The sampled slice is never written to memory: algorithmic traffic is the 4 B read of the blurred input
// slice per pixel per offset plus the (L2-resident) volume gather.  A finishing kernel (1 thread per
// active slice) turns the raw moments into the reference's mean-subtracted sums and replays the
// reference's temp-buffer choreography literally (quirks G1/G2 of oracle/reg_oracle.c), so the
// similarity values are the ones the reference produces.
// The optimiser (central differences, normalise, line search with compaction of still-improving
// slices) follows registerMultipleSlicesToVolume literally, including quirk G3; only the count of active
// slices crosses to the host (one pinned int per line-search step).
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <vector>
#include "../../include/svr_abi.h"
#include "svr_context.h"

#define REG_TILE 32
#define REG_THREADS 256
#define REG_MAXK 32          // (klength + 1) / 2 <= 32 for klength <= 63

struct RegKernel { float w[REG_MAXK]; int K; };   // half Gaussian kernel, w[0] = centre

struct RegState {
    int W = 0, H = 0, S = 0;
    float voxel = 1.f;
    float* resampled = nullptr;   // dev_v_slices_resampled        [S][H][W]
    float* blurred = nullptr;     // dev_v_slices_resampled_float
    float* tmp = nullptr;         // dev_temp_slices (X pass of the per-level blur)
    float* vol = nullptr;         // unused since the texture path (kept so the free list stays simple)
    cudaArray_t vol_array = nullptr;      // snapshot of the volume (dev_reconstructed_array, cuda2.cu:3931-3947)
    cudaTextureObject_t vol_tex = 0;      // linear filter, border addressing, normalised coordinates (cuda2.cu:3950-3956)
    int ax = 0, ay = 0, az = 0;           // extent of vol_array
    float* ofs = nullptr;         // dev_d_slicesOfs               [S][16]
    float* res_i2w = nullptr;     // dev_d_slicesResampledI2W      [S][16] (kept; unused by the kernels, as in the reference)
    float* M = nullptr;           // dev_recon_matrices            [S][16]
    float* Morig = nullptr;       // dev_recon_matrices_orig
    float* sim = nullptr;         // dev_recon_similarities        [5][S]
    float* grad = nullptr;        // dev_recon_gradient            [7][S]
    int* active = nullptr;        // dev_active_slices
    int* active2 = nullptr;
    int* active_prev = nullptr;
    double* moments = nullptr;    // [S][3][tiles][8]: per-CTA partial moments, summed in a fixed order (deterministic)
    float* slice_sum = nullptr;   // per-level sum / count of the blurred input slices (averageIf on layersA)
    int* slice_cnt = nullptr;
    int* d_count = nullptr;       // dev_active_slice_count
    int* h_count = nullptr;       // pinned
    bool have_slices = false, have_ofs = false, prepared = false;
    int n_levels = 2, n_steps = 4, n_iterations = 20;
    float epsilon = 0.0001f;
    long long evals = 0;          // (slice, offset) cost evaluations of the last register call
    double eval_ms = 0;           // device time of the fused kernel (when profiling is on)
};

static int reg_fail(svr_context* c, const char* msg) { c->err = msg; return 2; }
#define REG_REQUIRE(c, cond, msg) do { if (!(cond)) return reg_fail((c), (msg)); } while (0)

template <class T>
static int reg_alloc(svr_context* c, T** p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    SVR_CUDA(c, cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}

void svr_reg_free(svr_context* c)
{
    RegState* r = (RegState*)c->reg;
    if (!r) return;
    void* ptrs[] = { r->resampled, r->blurred, r->tmp, r->vol, r->ofs, r->res_i2w, r->M, r->Morig, r->sim, r->grad,
                     r->active, r->active2, r->active_prev, r->moments, r->slice_sum, r->slice_cnt, r->d_count };
    for (void* p : ptrs) if (p) cudaFree(p);
    if (r->vol_tex) cudaDestroyTextureObject(r->vol_tex);
    if (r->vol_array) cudaFreeArray(r->vol_array);
    if (r->h_count) cudaFreeHost(r->h_count);
    delete r;
    c->reg = nullptr;
}

// ---------------------------------------------------------------------------------------------
// generateGaussianKernel, gaussfilter.cu:56-88 + klength rule :189-192
static RegKernel make_kernel(float sigma)
{
    RegKernel k;
    int klength = std::max(std::min((int)(sigma * 5), 63), 7);
    klength -= 1 - klength % 2;
    float kernel[64];
    float sum = 0;
    const int mid = (int)floorf(klength / 2.0f);
    for (int i = 0; i < klength; i++) {
        kernel[i] = (float)exp(-(float)abs(i - mid) * (float)abs(i - mid) / (2 * sigma * sigma));
        sum += kernel[i];
    }
    for (int i = 0; i < klength; i++) kernel[i] /= sum;
    k.K = (klength + 1) / 2;
    for (int i = 0; i < REG_MAXK; ++i) k.w[i] = i < k.K ? kernel[klength / 2 + i] : 0.f;
    return k;
}

// GaussX/YKernel, gaussfilter.cu:92-173 (clamped reads, -1 centres pass through, neighbours clamped to >= 0)
__global__ void reg_blur_x_kernel(const float* __restrict__ in, float* __restrict__ out, int W, int H, RegKernel k)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* row = in + ((size_t)blockIdx.z * H + y) * W;
    float v = row[x];
    if (v != -1.f) {
        v = v * k.w[0];
        for (int i = 1; i < k.K; ++i)
            v = v + k.w[i] * (fmaxf(0.f, row[min(x + i, W - 1)]) + fmaxf(0.f, row[max(x - i, 0)]));
    }
    out[((size_t)blockIdx.z * H + y) * W + x] = v;
}
__global__ void reg_blur_y_kernel(const float* __restrict__ in, float* __restrict__ out, int W, int H, RegKernel k)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float* img = in + (size_t)blockIdx.z * H * W;
    float v = img[(size_t)y * W + x];
    if (v != -1.f) {
        v = v * k.w[0];
        for (int i = 1; i < k.K; ++i)
            v = v + k.w[i] * (fmaxf(0.f, img[(size_t)min(y + i, H - 1) * W + x]) + fmaxf(0.f, img[(size_t)max(y - i, 0) * W + x]));
    }
    out[((size_t)blockIdx.z * H + y) * W + x] = v;
}

// averageIf on the (blurred) input slices, cuda2.cu:4459-4496: one CTA per slice.
__global__ void __launch_bounds__(REG_THREADS)
reg_slice_mean_kernel(const float* __restrict__ img, int P, float* __restrict__ sum, int* __restrict__ cnt)
{
    const float* p = img + (size_t)blockIdx.x * P;
    double s = 0; int n = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const float v = p[i];
        if (v > -1.0f) { ++n; s += v; }
    }
    __shared__ double sh_s[REG_THREADS / 32];
    __shared__ int sh_n[REG_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); n += __shfl_down_sync(0xffffffffu, n, o); }
    if ((threadIdx.x & 31) == 0) { sh_s[threadIdx.x >> 5] = s; sh_n[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < REG_THREADS / 32; ++w) { s += sh_s[w]; n += sh_n[w]; }
        sum[blockIdx.x] = (float)s;
        cnt[blockIdx.x] = n;
    }
}

// ---------------------------------------------------------------------------------------------
// tex3D(reconstructedTex_, pos / size) (cuda2.cu:3525): the volume is sampled by the TEXTURE UNIT, exactly as the
// reference does it -- linear filter, normalised coordinates, border addressing, no half-texel offset (quirk G4).
// Hardware filtering uses 1.8 fixed-point weights and its own arithmetic; going through the same unit makes the
// sampled slices (and through them the similarity staircase the optimiser walks on) those of the reference, which a
// software trilinear with float weights cannot reproduce (measured: 3e-4 relative on the samples, enough to change
// which side of the `val < 0` padding test mask-border pixels fall on).  The reference divides with --use_fast_math.
__device__ __forceinline__ float reg_tex3d(cudaTextureObject_t tex, float sx, float sy, float sz, float px, float py, float pz)
{
    return tex3D<float>(tex, __fdividef(px, sx), __fdividef(py, sy), __fdividef(pz, sz));
}

__device__ __forceinline__ float3 reg_mul_pt(const float* __restrict__ m, float x, float y, float z)
{
    return make_float3(m[0] * x + m[1] * y + m[2] * z + m[3], m[4] * x + m[5] * y + m[6] * z + m[7],
                       m[8] * x + m[9] * y + m[10] * z + m[11]);
}

// The fused cost kernel: R1 (cuda2.cu:3504-3529) + R2 (gaussfilter.cu) + R3 (averageIf on the sampled slice)
// + the raw moments R4 (computeNCCAndReduce, cuda2.cu:4498-4550) needs.
// grid = (tilesX*tilesY, active slices, 3 in-slice offsets); dynamic smem = raw[(T+2h)^2] + xpass[(T+2h)*T].
__global__ void __launch_bounds__(REG_THREADS)
reg_eval_kernel(cudaTextureObject_t vol, int vx, int vy, int vz, VolGeom vg, const float* __restrict__ blurred,
                const int* __restrict__ active, const float* __restrict__ M, const float* __restrict__ ofs, int W, int H,
                int tilesX, int level_mod, RegKernel k, double* __restrict__ moments)
{
    extern __shared__ float smem[];
    const int h = k.K - 1, TW = REG_TILE + 2 * h;
    float* raw = smem;                    // [TW][TW]
    float* xp = smem + TW * TW;           // [TW][REG_TILE]
    const int t = blockIdx.y, o = blockIdx.z;
    const int slice = active[t];
    const int tx0 = (blockIdx.x % tilesX) * REG_TILE, ty0 = (blockIdx.x / tilesX) * REG_TILE;
    __shared__ float sT[12], sO[12];
    if (threadIdx.x < 12) { sT[threadIdx.x] = M[16 * slice + threadIdx.x]; sO[threadIdx.x] = ofs[16 * slice + threadIdx.x]; }
    __syncthreads();
    const float zofs = (float)((o - 1) * 2);

    // 1. sample tile + halo at clamped pixel coordinates (cudaBoundaryModeClamp of the Gauss kernels)
    for (int i = threadIdx.x; i < TW * TW; i += REG_THREADS) {
        const int ly = i / TW, lx = i - ly * TW;
        const int gx = min(max(tx0 + lx - h, 0), W - 1), gy = min(max(ty0 + ly - h, 0), H - 1);
        float3 w = reg_mul_pt(sO, (float)gx, (float)gy, zofs);
        w = reg_mul_pt(sT, w.x, w.y, w.z);
        const float3 p = reg_mul_pt(vg.rw2i, w.x, w.y, w.z);
        float val = reg_tex3d(vol, (float)vx, (float)vy, (float)vz, p.x, p.y, p.z);
        if (val < 0) val = -1.0f;
        raw[i] = val;
    }
    __syncthreads();
    // 2. X pass for all TW rows, REG_TILE columns
    for (int i = threadIdx.x; i < TW * REG_TILE; i += REG_THREADS) {
        const int ly = i / REG_TILE, lx = i - ly * REG_TILE;
        const float* row = raw + ly * TW + lx + h;
        float v = row[0];
        if (v != -1.f) {
            v = v * k.w[0];
            for (int j = 1; j < k.K; ++j) v = v + k.w[j] * (fmaxf(0.f, row[j]) + fmaxf(0.f, row[-j]));
        }
        xp[i] = v;
    }
    __syncthreads();
    // 3. Y pass + moments
    double n = 0, sa = 0, sb = 0, sab = 0, saa = 0, sbb = 0, nb = 0, sball = 0;
    for (int i = threadIdx.x; i < REG_TILE * REG_TILE; i += REG_THREADS) {
        const int ly = i / REG_TILE, lx = i - ly * REG_TILE;
        const int gx = tx0 + lx, gy = ty0 + ly;
        if (gx >= W || gy >= H) continue;
        const float* col = xp + (ly + h) * REG_TILE + lx;
        float b = col[0];
        if (b != -1.f) {
            b = b * k.w[0];
            for (int j = 1; j < k.K; ++j) b = b + k.w[j] * (fmaxf(0.f, col[j * REG_TILE]) + fmaxf(0.f, col[-j * REG_TILE]));
        }
        if (b > -1.0f) { nb += 1.0; sball += b; }
        const int lin = gy * W + gx;
        const float a = blurred[(size_t)slice * W * H + lin];
        if (a >= 0.0f && b >= 0.0f && lin % level_mod == 0) {
            n += 1.0; sa += a; sb += b;
            sab += (double)a * b; saa += (double)a * a; sbb += (double)b * b;
        }
    }
    // 4. block reduce (warp shuffles, then one warp) and 8 double atomics
    double vals[8] = { n, sa, sb, sab, saa, sbb, nb, sball };
    __shared__ double red[REG_THREADS / 32][8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        double v = vals[q];
        for (int s = 16; s > 0; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = 0;
        for (int w = 0; w < REG_THREADS / 32; ++w) v += red[w][threadIdx.x];
        // one slot per CTA, reduced in a fixed order by reg_finish_kernel: the similarity (and so the optimiser's
        // trajectory) is bit-reproducible from run to run, which a float/double atomic accumulation is not
        moments[(((size_t)t * 3 + o) * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = v;
    }
}

// addNccValues / writeSimilarities (cuda2.cu:4552-4575) with the reference's temp-buffer layout replayed
// literally: per active index t the NCC accumulator lives at tf[2a+t], the triplet at tf[3a+3t..], and
// between offsets the reference clears tf[2S, 5S) (quirk G1); sum/count of the sampled slice accumulate
// over the offsets (quirk G2).
__global__ void reg_finish_kernel(int a, int S, int tiles, const int* __restrict__ active, const double* __restrict__ moments,
                                  const float* __restrict__ slice_sum, const int* __restrict__ slice_cnt,
                                  float* __restrict__ sim, int writeoffset, int writestep, int writenum)
{   // one warp per active slice
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= a) return;
    // fixed-order sum of the per-CTA partial moments: lanes stride over the tiles, then a butterfly
    double mo[3][8];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double v = 0.0;
            for (int tl = lane; tl < tiles; tl += 32) v += moments[(((size_t)t * 3 + o) * tiles + tl) * 8 + q];
            for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
            mo[o][q] = v;
        }
    if (lane != 0) return;
    const int slice = active[t];
    float avg_a = slice_sum[slice];
    if (avg_a != 0) avg_a /= slice_cnt[slice];
    float cum_sb = 0.f; int cum_nb = 0;
    float trip[3] = { 0.f, 0.f, 0.f };
    float acc = 0.f;
    const int zlo = 2 * S, zhi = 5 * S;
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        const double* m = mo[o];
        cum_sb += (float)m[7];
        cum_nb += (int)m[6];
        if (2 * a + t >= zlo && 2 * a + t < zhi) acc = 0.f;
        for (int q = 0; q < 3; ++q) { const int idx = 3 * a + 3 * t + q; if (idx >= zlo && idx < zhi) trip[q] = 0.f; }
        float avg_b = cum_sb;
        if (avg_b != 0) avg_b /= cum_nb;
        const double A = avg_a, B = avg_b, n = m[0], sa = m[1], sb = m[2];
        trip[0] += (float)(m[3] - A * sb - B * sa + n * A * B);
        trip[1] += (float)(m[4] - 2.0 * A * sa + n * A * A);
        trip[2] += (float)(m[5] - 2.0 * B * sb + n * B * B);
        const float norm = trip[1] * trip[2];
        float res = 0;
        if (norm > 0) res = trip[0] / sqrtf(norm);
        acc += res;
    }
    for (int i = 0; i < writenum; ++i) sim[(size_t)writeoffset * S + (size_t)S * writestep * i + slice] = acc;
}

// ---------------------------------------------------------------------------------------------
// Parameter kernels (one thread per active slice), cuda2.cu:4224-4391
__device__ __forceinline__ void reg_rot_params(const float* in, float p_rot[3])
{
    const float TOL = 0.000001f;
    const float tmp = asinf(-1.0f * in[2]);
    if (fabsf(__cosf(tmp)) > TOL) {
        p_rot[0] = atan2f(in[6], in[10]);
        p_rot[1] = tmp;
        p_rot[2] = atan2f(in[1], in[0]);
    } else {
        p_rot[0] = atan2f(-in[2] * in[4], -in[2] * in[8]);
        p_rot[1] = tmp;
        p_rot[2] = 0;
    }
}
__device__ __forceinline__ void reg_set_rotation(float* out, const float p_rot[3])
{
    // the reference is built with --use_fast_math: cos/sin are the MUFU intrinsics there
    const float cosrx = __cosf(p_rot[0]), cosry = __cosf(p_rot[1]), cosrz = __cosf(p_rot[2]);
    const float sinrx = __sinf(p_rot[0]), sinry = __sinf(p_rot[1]), sinrz = __sinf(p_rot[2]);
    out[0] = cosry * cosrz; out[1] = cosry * sinrz; out[2] = -sinry;
    out[4] = (sinrx * sinry * cosrz - cosrx * sinrz);
    out[5] = (sinrx * sinry * sinrz + cosrx * cosrz);
    out[6] = sinrx * cosry;
    out[8] = (cosrx * sinry * cosrz + sinrx * sinrz);
    out[9] = (cosrx * sinry * sinrz - sinrx * cosrz);
    out[10] = cosrx * cosry;
}

__global__ void reg_init_active_kernel(int* buffer, int num)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i < num) buffer[i] = i;
}

__global__ void reg_adjust_kernel(const float* __restrict__ inM, float* __restrict__ outM, const int* __restrict__ active,
                                  int a, int part, float step)
{
    const float pi = 3.14159265358979323846f;
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= a) return;
    const int slice = active[i];
    float in[16], out[16];
    for (int q = 0; q < 16; ++q) { in[q] = inM[16 * slice + q]; out[q] = in[q]; }
    if (part < 3) {
        out[4 * part + 3] = in[4 * part + 3] + step;
    } else {
        float p_rot[3];
        reg_rot_params(in, p_rot);
        p_rot[part - 3] += step * pi / 180.0f;
        reg_set_rotation(out, p_rot);
    }
    for (int q = 0; q < 16; ++q) outM[16 * slice + q] = out[q];
}

__global__ void reg_gradient_kernel(const float* __restrict__ sim34, float* __restrict__ grad, const int* __restrict__ active,
                                    int a, int S, int p)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= a) return;
    const int slice = active[i];
    const float dx = sim34[slice] - sim34[S + slice];
    grad[(size_t)p * S + slice] = dx;
    if (p == 0) grad[(size_t)6 * S + slice] = dx * dx;
    else grad[(size_t)6 * S + slice] += dx * dx;
}

__global__ void reg_normalize_kernel(float* __restrict__ grad, const int* __restrict__ active, int a, int S)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= a) return;
    const int slice = active[i];
    float norm = grad[(size_t)6 * S + slice];
    if (norm > 0) norm = 1.0f / sqrtf(norm);
    for (int j = 0; j < 6; ++j) grad[(size_t)j * S + slice] *= norm;
}

__global__ void reg_copy_sim_kernel(float* __restrict__ sim, int a, int S, const int* __restrict__ active, int target, int source)
{
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= a) return;
    const int slice = active[i];
    sim[(size_t)target * S + slice] = sim[(size_t)source * S + slice];
}

__global__ void reg_step_kernel(float* __restrict__ M, const float* __restrict__ grad, int a, int S,
                                const int* __restrict__ active, float step)
{
    const float pi = 3.14159265358979323846f;
    const int i = threadIdx.x + blockIdx.x * blockDim.x;
    if (i >= a) return;
    const int slice = active[i];
    float m[16];
    for (int q = 0; q < 16; ++q) m[q] = M[16 * slice + q];
    for (int p = 0; p < 3; ++p) m[4 * p + 3] = m[4 * p + 3] + step * grad[(size_t)p * S + slice];
    float p_rot[3];
    reg_rot_params(m, p_rot);
    for (int p = 0; p < 3; ++p) p_rot[p] += grad[(size_t)(p + 3) * S + slice] * step * pi / 180.0f;
    reg_set_rotation(m, p_rot);
    for (int q = 0; q < 16; ++q) M[16 * slice + q] = m[q];
}

// checkImprovement (cuda2.cu:4393-4457) as a single-CTA STABLE compaction (the reference's order across
// 512-thread blocks is decided by an atomicAdd race).
__global__ void __launch_bounds__(1024)
reg_compact_kernel(int* __restrict__ newActive, int a, int S, const int* __restrict__ active, const float* __restrict__ sim,
                   int cursim, int prev, float eps, int* __restrict__ count)
{
    __shared__ int warp_tot[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < a; start += blockDim.x) {
        const int tid = start + threadIdx.x;
        int slice = -1, flag = 0;
        if (tid < a) {
            slice = active[tid];
            flag = sim[(size_t)cursim * S + slice] > sim[(size_t)prev * S + slice] + eps ? 1 : 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int within = __popc(bal & ((1u << lane) - 1));
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < wid; ++w) before += warp_tot[w];
        if (flag) newActive[base + before + within] = slice;
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_tot[w]; base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}

// ---------------------------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

static int reg_blur_level(svr_context* c, RegState* r, float blur)
{   // copyFromOther + FilterGaussStack + (hoisted) averageIf of the input slices, cuda2.cu:4018-4026,4164
    const size_t n = (size_t)r->W * r->H * r->S;
    if (r->S == 0) return 0;
    const RegKernel k = make_kernel(blur);
    dim3 b(32, 8), g(cdiv(r->W, 32), cdiv(r->H, 8), r->S);
    reg_blur_x_kernel<<<g, b, 0, c->stream>>>(r->resampled, r->tmp, r->W, r->H, k);
    SVR_KERNEL_CHECK(c);
    reg_blur_y_kernel<<<g, b, 0, c->stream>>>(r->tmp, r->blurred, r->W, r->H, k);
    SVR_KERNEL_CHECK(c);
    reg_slice_mean_kernel<<<r->S, REG_THREADS, 0, c->stream>>>(r->blurred, r->W * r->H, r->slice_sum, r->slice_cnt);
    SVR_KERNEL_CHECK(c);
    (void)n;
    return 0;
}

static int reg_evaluate_costs(svr_context* c, RegState* r, int a, int level, const RegKernel& k, int writeoffset,
                              int writestep, int writenum)
{   // evaluateCostsMultipleSlices, cuda2.cu:4150-4221
    if (a == 0) return 0;
    const int tilesX = cdiv(r->W, REG_TILE), tilesY = cdiv(r->H, REG_TILE);
    const int h = k.K - 1, TW = REG_TILE + 2 * h;
    const size_t smem = sizeof(float) * ((size_t)TW * TW + (size_t)TW * REG_TILE);
    {
        ProfScope prof(c, 5);
        dim3 grid(tilesX * tilesY, a, 3);
        reg_eval_kernel<<<grid, REG_THREADS, smem, c->stream>>>(r->vol_tex, c->vx, c->vy, c->vz, c->vg, r->blurred, r->active,
                                                                 r->M, r->ofs, r->W, r->H, tilesX, level + 1, k, r->moments);
        SVR_KERNEL_CHECK(c);
    }
    reg_finish_kernel<<<cdiv(a, 4), 128, 0, c->stream>>>(a, r->S, tilesX * tilesY, r->active, r->moments, r->slice_sum, r->slice_cnt, r->sim,
                                                           writeoffset, writestep, writenum);
    SVR_KERNEL_CHECK(c);
    r->evals += 3LL * a;
    return 0;
}

static int reg_read_count(svr_context* c, RegState* r, int* out)
{
    SVR_CUDA(c, cudaMemcpyAsync(r->h_count, r->d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    *out = *r->h_count;
    return 0;
}

static int reg_check_ready(svr_context* c, RegState* r, const char* who)
{
    (void)who;
    REG_REQUIRE(c, r && r->resampled, "registration: call svr_reg_init_storage first");
    REG_REQUIRE(c, r->have_slices, "registration: call svr_reg_fill_slices first");
    REG_REQUIRE(c, r->have_ofs, "registration: call svr_reg_update_slices_i2w first");
    REG_REQUIRE(c, r->prepared, "registration: call svr_reg_prepare first");
    return 0;
}

extern "C" {

int svr_reg_init_storage(svr_context* c, int W, int H, int S, float dx, float dy, float dz)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    (void)dy; (void)dz;
    REG_REQUIRE(c, W > 0 && H > 0 && S >= 0, "svr_reg_init_storage: bad size");
    REG_REQUIRE(c, (double)W * H * std::max(S, 1) < 2147483647.0, "svr_reg_init_storage: slice cube too large");
    SVR_CUDA(c, cudaSetDevice(c->device));
    svr_reg_free(c);
    RegState* r = new RegState();
    c->reg = r;
    r->W = W; r->H = H; r->S = S; r->voxel = dx;
    const size_t n = (size_t)W * H * S, Sn = (size_t)std::max(S, 1);
    if (reg_alloc(c, &r->resampled, n) || reg_alloc(c, &r->blurred, n) || reg_alloc(c, &r->tmp, n) ||
        reg_alloc(c, &r->ofs, 16 * Sn) || reg_alloc(c, &r->res_i2w, 16 * Sn) || reg_alloc(c, &r->M, 16 * Sn) ||
        reg_alloc(c, &r->Morig, 16 * Sn) || reg_alloc(c, &r->sim, 5 * Sn) || reg_alloc(c, &r->grad, 7 * Sn) ||
        reg_alloc(c, &r->active, Sn) || reg_alloc(c, &r->active2, Sn) || reg_alloc(c, &r->active_prev, Sn) ||
        reg_alloc(c, &r->moments, 24 * Sn * (size_t)(cdiv(W, REG_TILE) * cdiv(H, REG_TILE))) || reg_alloc(c, &r->slice_sum, Sn) || reg_alloc(c, &r->slice_cnt, Sn) ||
        reg_alloc(c, &r->d_count, 1))
        return 1;
    SVR_CUDA(c, cudaMallocHost((void**)&r->h_count, sizeof(int)));
    SVR_CUDA(c, cudaMemsetAsync(r->sim, 0, sizeof(float) * 5 * Sn, c->stream));
    SVR_CUDA(c, cudaMemsetAsync(r->grad, 0, sizeof(float) * 7 * Sn, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_reg_fill_slices(svr_context* c, const float* cube, const float* slices_resampled_i2w)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r, "svr_reg_fill_slices: call svr_reg_init_storage first");
    const size_t n = (size_t)r->W * r->H * r->S;
    if (n == 0) { r->have_slices = true; return 0; }
    REG_REQUIRE(c, cube, "svr_reg_fill_slices: cube is NULL");
    SVR_CUDA(c, cudaMemcpyAsync(r->resampled, cube, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (slices_resampled_i2w)
        SVR_CUDA(c, cudaMemcpyAsync(r->res_i2w, slices_resampled_i2w, sizeof(float) * 16 * r->S, cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    r->have_slices = true;
    return 0;
}

// irtkResamplingWithPadding::Run for single-plane slices (irtkResamplingWithPadding.cc:36-183), one thread per output
// pixel, double precision, no FMA contraction and the host's operation order (bit-identical to the host front-end):
// neighbours equal to the padding value are dropped and the remaining weights renormalised; out-of-bounds neighbours
// count as not padded but add nothing; the result is padding when >= 4 of the 8 neighbours are padding or the weight
// sum is 0.
__global__ void __launch_bounds__(256)
reg_resample_kernel(int S, int W, int H, int Nx, int Ny, const float* __restrict__ slices, const double* __restrict__ m,
                    const int* __restrict__ in_sizes, const int* __restrict__ out_sizes, float* __restrict__ out)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t P = (size_t)W * H;
    if (idx >= P * S) return;
    const int s = (int)(idx / P), pix = (int)(idx - (size_t)s * P);
    const int j = pix / W, i = pix - j * W;
    if (i >= out_sizes[2 * s] || j >= out_sizes[2 * s + 1]) { out[idx] = -1.0f; return; }
    const int ax = in_sizes[2 * s], ay = in_sizes[2 * s + 1];
    // the reference maps the output voxel to world (ImageToWorld of the resampled slice) and then into the input image
    // (WorldToImage of the slice), two separately rounded matrix-vector products (image++/include/irtkBaseImage.h:425-468):
    // on grids that coincide the coordinates are integers up to rounding, and floor() must fall as it does there
    const double* q = m + 24 * (size_t)s;
    const double* p = q + 12;
    const double di = (double)i, dj = (double)j, dk = 0.0;
    const double wx_ = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[0], di), __dmul_rn(q[1], dj)), __dmul_rn(q[2], dk)), q[3]);
    const double wy_ = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[4], di), __dmul_rn(q[5], dj)), __dmul_rn(q[6], dk)), q[7]);
    const double wz_ = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(q[8], di), __dmul_rn(q[9], dj)), __dmul_rn(q[10], dk)), q[11]);
    const double x = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p[0], wx_), __dmul_rn(p[1], wy_)), __dmul_rn(p[2], wz_)), p[3]);
    const double y = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p[4], wx_), __dmul_rn(p[5], wy_)), __dmul_rn(p[6], wz_)), p[7]);
    const double z = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(p[8], wx_), __dmul_rn(p[9], wy_)), __dmul_rn(p[10], wz_)), p[11]);
    const double fu = floor(x), fv = floor(y), fw = floor(z);
    const int u = (int)fu, v = (int)fv, w = (int)fw;
    const double dx = __dadd_rn(x, -fu), dy = __dadd_rn(y, -fv), dz = __dadd_rn(z, -fw);
    const float* img = slices + (size_t)s * Nx * Ny;
    double val = 0.0, wsum = 0.0;
    int pad = 8;
#pragma unroll
    for (int du = 0; du < 2; ++du)
#pragma unroll
        for (int dv = 0; dv < 2; ++dv)
#pragma unroll
            for (int dw = 0; dw < 2; ++dw) {
                const int uu = u + du, vv = v + dv, ww = w + dw;
                const bool inb = uu >= 0 && uu < ax && vv >= 0 && vv < ay && ww >= 0 && ww < 1;
                const double g = inb ? (double)img[(size_t)vv * Nx + uu] : -1.0;
                const bool good = inb && g != -1.0;
                const double wx = du ? dx : __dadd_rn(1.0, -dx), wy = dv ? dy : __dadd_rn(1.0, -dy), wz = dw ? dz : __dadd_rn(1.0, -dz);
                const double wt = __dmul_rn(__dmul_rn(wx, wy), wz);
                if (good) { val = __dadd_rn(val, __dmul_rn(g, wt)); wsum = __dadd_rn(wsum, wt); }
                if (!inb || good) --pad;
            }
    out[idx] = (pad < 4 && wsum > 0.0) ? (float)__ddiv_rn(val, wsum) : -1.0f;
}

int svr_reg_resample_slices(svr_context* c, const double* src_from_out, const int* in_sizes, const int* out_sizes,
                            const float* slices_resampled_i2w)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r, "svr_reg_resample_slices: call svr_reg_init_storage first");
    REG_REQUIRE(c, r->S == c->S, "svr_reg_resample_slices: the registration cube and the slice cube hold different numbers of slices");
    const int S = r->S;
    if (S == 0 || (size_t)r->W * r->H == 0) { r->have_slices = true; return 0; }
    REG_REQUIRE(c, c->slices, "svr_reg_resample_slices: call svr_init_storage_volumes / svr_fill_slices first");
    REG_REQUIRE(c, src_from_out && in_sizes && out_sizes, "svr_reg_resample_slices: NULL argument");
    for (int s = 0; s < S; ++s) {
        REG_REQUIRE(c, in_sizes[2 * s] >= 0 && in_sizes[2 * s] <= c->Nx && in_sizes[2 * s + 1] >= 0 && in_sizes[2 * s + 1] <= c->Ny,
                    "svr_reg_resample_slices: slice extent outside the slice cube");
        REG_REQUIRE(c, out_sizes[2 * s] >= 0 && out_sizes[2 * s] <= r->W && out_sizes[2 * s + 1] >= 0 && out_sizes[2 * s + 1] <= r->H,
                    "svr_reg_resample_slices: resampled extent outside the registration cube");
    }
    SVR_CUDA(c, cudaSetDevice(c->device));
    double* d_m = nullptr;
    int* d_sz = nullptr;
    SVR_CUDA(c, cudaMalloc(&d_m, sizeof(double) * 24 * S));
    if (cudaMalloc(&d_sz, sizeof(int) * 4 * S) != cudaSuccess) { cudaFree(d_m); c->err = "svr_reg_resample_slices: out of device memory"; return 1; }
    cudaMemcpyAsync(d_m, src_from_out, sizeof(double) * 24 * S, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_sz, in_sizes, sizeof(int) * 2 * S, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_sz + 2 * S, out_sizes, sizeof(int) * 2 * S, cudaMemcpyHostToDevice, c->stream);
    if (slices_resampled_i2w)
        cudaMemcpyAsync(r->res_i2w, slices_resampled_i2w, sizeof(float) * 16 * S, cudaMemcpyHostToDevice, c->stream);
    const size_t n = (size_t)r->W * r->H * S;
    reg_resample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(S, r->W, r->H, c->Nx, c->Ny, c->slices, d_m, d_sz, d_sz + 2 * S,
                                                                           r->resampled);
    c->launches++;
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_m);
    cudaFree(d_sz);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { c->err = "svr_reg_resample_slices: kernel failed"; return 1; }
    r->have_slices = true;
    return 0;
}

int svr_reg_update_slices_i2w(svr_context* c, const float* ofs)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r, "svr_reg_update_slices_i2w: call svr_reg_init_storage first");
    if (r->S) {
        REG_REQUIRE(c, ofs, "svr_reg_update_slices_i2w: matrices are NULL");
        SVR_CUDA(c, cudaMemcpyAsync(r->ofs, ofs, sizeof(float) * 16 * r->S, cudaMemcpyHostToDevice, c->stream));
        SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    r->have_ofs = true;
    return 0;
}

int svr_reg_prepare(svr_context* c)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r, "svr_reg_prepare: call svr_reg_init_storage first");
    REG_REQUIRE(c, c->V > 0 && c->recon, "svr_reg_prepare: no reconstruction volume");
    // snapshot of the volume in a 3D cudaArray behind a texture object (cuda2.cu:3931-3956)
    if (!r->vol_array || r->ax != c->vx || r->ay != c->vy || r->az != c->vz) {
        if (r->vol_tex) { cudaDestroyTextureObject(r->vol_tex); r->vol_tex = 0; }
        if (r->vol_array) { cudaFreeArray(r->vol_array); r->vol_array = nullptr; }
        cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
        SVR_CUDA(c, cudaMalloc3DArray(&r->vol_array, &desc, make_cudaExtent(c->vx, c->vy, c->vz)));
        r->ax = c->vx; r->ay = c->vy; r->az = c->vz;
        cudaResourceDesc res;
        memset(&res, 0, sizeof(res));
        res.resType = cudaResourceTypeArray;
        res.res.array.array = r->vol_array;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 1;
        SVR_CUDA(c, cudaCreateTextureObject(&r->vol_tex, &res, &td, nullptr));
    }
    {
        cudaMemcpy3DParms cp;
        memset(&cp, 0, sizeof(cp));
        cp.srcPtr = make_cudaPitchedPtr((void*)c->recon, (size_t)c->vx * sizeof(float), c->vx, c->vy);
        cp.dstArray = r->vol_array;
        cp.extent = make_cudaExtent(c->vx, c->vy, c->vz);
        cp.kind = cudaMemcpyDeviceToDevice;
        SVR_CUDA(c, cudaMemcpy3DAsync(&cp, c->stream));
    }
    r->voxel = c->vdx;                 // _Blurring[0] = dev_reconstructed_.dim.x / 2, cuda2.cu:3892
    r->n_levels = 2; r->n_steps = 4; r->n_iterations = 20; r->epsilon = 0.0001f;
    r->prepared = true;
    return 0;
}

int svr_reg_set_schedule(svr_context* c, int n_levels, int n_steps, int n_iterations)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r, "svr_reg_set_schedule: call svr_reg_init_storage first");
    REG_REQUIRE(c, n_levels >= 1 && n_levels <= 8 && n_steps >= 1 && n_iterations >= 1, "svr_reg_set_schedule: bad schedule");
    r->n_levels = n_levels; r->n_steps = n_steps; r->n_iterations = n_iterations;
    return 0;
}

int svr_reg_evaluate(svr_context* c, const float* transforms, int level, float* similarity)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    if (reg_check_ready(c, r, "svr_reg_evaluate")) return 2;
    REG_REQUIRE(c, level >= 0 && level < 8, "svr_reg_evaluate: bad level");
    if (r->S == 0) return 0;
    REG_REQUIRE(c, transforms && similarity, "svr_reg_evaluate: NULL argument");
    SVR_CUDA(c, cudaMemcpyAsync(r->M, transforms, sizeof(float) * 16 * r->S, cudaMemcpyHostToDevice, c->stream));
    float blur = r->voxel / 2.0f;
    for (int i = 0; i < level; ++i) blur *= 2;
    if (reg_blur_level(c, r, blur)) return 1;
    reg_init_active_kernel<<<cdiv(r->S, 512), 512, 0, c->stream>>>(r->active, r->S);
    SVR_KERNEL_CHECK(c);
    const RegKernel k = make_kernel(blur);
    if (reg_evaluate_costs(c, r, r->S, level, k, 0, 1, 1)) return 1;
    SVR_CUDA(c, cudaMemcpyAsync(similarity, r->sim, sizeof(float) * r->S, cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_reg_register(svr_context* c, float* transforms)
{   // registerMultipleSlicesToVolume, cuda2.cu:4001-4141
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    if (reg_check_ready(c, r, "svr_reg_register")) return 2;
    const int S = r->S;
    r->evals = 0;
    if (S == 0) return 0;
    REG_REQUIRE(c, transforms, "svr_reg_register: transforms is NULL");
    SVR_CUDA(c, cudaMemcpyAsync(r->M, transforms, sizeof(float) * 16 * S, cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(r->Morig, r->M, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, c->stream));

    float Blurring[8], LengthOfSteps[8];
    Blurring[0] = r->voxel / 2.0f;
    for (int i = 0; i < r->n_levels; i++) LengthOfSteps[i] = (float)(0.1 * pow(2.0f, i));
    for (int i = 1; i < r->n_levels; i++) Blurring[i] = Blurring[i - 1] * 2;
    const int T = 512;

    for (int level = r->n_levels - 1; level >= 0; --level) {
        const float blur = Blurring[level];
        float step = LengthOfSteps[level];
        if (reg_blur_level(c, r, blur)) return 1;
        const RegKernel k = make_kernel(blur);
        for (int st = 0; st < r->n_steps; st++) {
            reg_init_active_kernel<<<cdiv(S, T), T, 0, c->stream>>>(r->active, S);
            SVR_KERNEL_CHECK(c);
            int a = S;
            for (int iter = 0; iter < r->n_iterations; iter++) {
                if (reg_evaluate_costs(c, r, a, level, k, 0, 1, 3)) return 1;
                for (int p = 0; p < 6; ++p) {
                    reg_adjust_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->Morig, r->M, r->active, a, p, step);
                    SVR_KERNEL_CHECK(c);
                    if (reg_evaluate_costs(c, r, a, level, k, 3, 0, 1)) return 1;
                    reg_adjust_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->Morig, r->M, r->active, a, p, -step);
                    SVR_KERNEL_CHECK(c);
                    if (reg_evaluate_costs(c, r, a, level, k, 4, 0, 1)) return 1;
                    reg_gradient_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->sim + 3 * (size_t)S, r->grad, r->active, a, S, p);
                    SVR_KERNEL_CHECK(c);
                }
                reg_normalize_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->grad, r->active, a, S);
                SVR_KERNEL_CHECK(c);
                const int prevActive = a;
                SVR_CUDA(c, cudaMemcpyAsync(r->active_prev, r->active, sizeof(int) * a, cudaMemcpyDeviceToDevice, c->stream));
                do {
                    reg_copy_sim_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->sim, a, S, r->active, 2, 0);
                    SVR_KERNEL_CHECK(c);
                    reg_step_kernel<<<cdiv(a, T), T, 0, c->stream>>>(r->M, r->grad, a, S, r->active, step);   // quirk G3
                    SVR_KERNEL_CHECK(c);
                    if (reg_evaluate_costs(c, r, a, level, k, 0, 1, 1)) return 1;
                    reg_compact_kernel<<<1, 1024, 0, c->stream>>>(r->active2, a, S, r->active, r->sim, 0, 2, r->epsilon, r->d_count);
                    SVR_KERNEL_CHECK(c);
                    std::swap(r->active, r->active2);
                    if (reg_read_count(c, r, &a)) return 1;
                } while (a > 0);
                reg_step_kernel<<<cdiv(prevActive, T), T, 0, c->stream>>>(r->M, r->grad, prevActive, S, r->active_prev, -step);
                SVR_KERNEL_CHECK(c);
                SVR_CUDA(c, cudaMemcpyAsync(r->Morig, r->M, sizeof(float) * 16 * S, cudaMemcpyDeviceToDevice, c->stream));
                reg_compact_kernel<<<1, 1024, 0, c->stream>>>(r->active, prevActive, S, r->active_prev, r->sim, 2, 1, r->epsilon, r->d_count);
                SVR_KERNEL_CHECK(c);
                if (reg_read_count(c, r, &a)) return 1;
                if (a == 0) break;
            }
            step /= 2.0f;
        }
    }
    SVR_CUDA(c, cudaMemcpyAsync(transforms, r->M, sizeof(float) * 16 * S, cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int64_t svr_reg_evaluations(const svr_context* c)
{
    SVR_ENTRY(c);
    const RegState* r = c ? (const RegState*)c->reg : nullptr;
    return r ? r->evals : 0;
}

int svr_reg_debug_get(svr_context* c, int kind, void* out)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    RegState* r = (RegState*)c->reg;
    REG_REQUIRE(c, r && out, "svr_reg_debug_get: no registration storage");
    const size_t n = (size_t)r->W * r->H * r->S;
    const void* src = nullptr; size_t bytes = 0;
    switch (kind) {
    case 0: src = r->resampled; bytes = n * sizeof(float); break;
    case 1: src = r->blurred; bytes = n * sizeof(float); break;
    case 2: src = r->sim; bytes = sizeof(float) * 5 * r->S; break;
    case 3: src = r->grad; bytes = sizeof(float) * 7 * r->S; break;
    default: return reg_fail(c, "svr_reg_debug_get: unknown kind");
    }
    if (bytes == 0) return 0;
    SVR_CUDA(c, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
