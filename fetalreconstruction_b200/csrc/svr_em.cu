// svr_em.cu -- robust-statistics (EM) kernels, regulariser and volume elementwise kernels for sm_100a.
//
// The reference issues one thrust::transform_reduce (+ a host sync) PER SLICE for the slice
// potentials, the scale vector and the inside flags (reconstruction_cuda2.cu:2892-2911, 3206-3237,
// 2742-2752) and materialises a per-pixel scale buffer with one thrust::fill per slice for the
// M-step (cuda2.cu:3091-3094).  Here every statistic is ONE launch: per-slice sums use a
// (chunk, slice) grid with cub::BlockReduce and one double atomicAdd per CTA; global sums use a
// two-stage deterministic reduction.
#include <cub/block/block_reduce.cuh>
#include "svr_context.h"

#define EM_THREADS 256

// ---- K12: InitializeEMValuesKernel (reconstruction_cuda2.cu:3241-3267) ------------------------
// PVR (patchBasedRobustStatistics_gpu.cu:57-78) also zeroes the weight of pixels that are exactly 0.
__global__ void init_em_kernel(size_t n, const float* __restrict__ slices, float* __restrict__ weights, int flavor)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float s = slices[i];
        weights[i] = (s != -1.0f && (flavor == 0 || s != 0.0f)) ? 1.0f : 0.0f;
    }
}
int svr_launch_init_em(svr_context* c)
{
    init_em_kernel<<<c->sm_count * 8, EM_THREADS, 0, c->stream>>>(c->NP, c->slices, c->weights, c->flavor);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// G_ / M_ (reconstruction_cuda2.cu:62-70)
// (PVR: __step = 1e-5, include/reconConfig.cuh:120)
__device__ __forceinline__ float G_(float x, float s, float step) { return step * __expf(-x * x / (2.0f * s)) / (sqrtf(6.28f * s)); }
__device__ __forceinline__ float M_(float m, float step) { return m * step; }

// ---- iteration over the compacted list of pixels != -1 -----------------------------------------------
// Every robust-statistics kernel only looks at pixels that are not padding (s != -1), 27 % of the padded slice cube at C3.
// They therefore walk valid_idx (built once per svr_fill_slices, sorted, slice-major) instead of the cube: a thread takes
// four consecutive list entries with one 128-bit load and gathers the per-pixel values; neighbouring threads read
// neighbouring pixels, so every 32-byte sector that is fetched is used.  Padding pixels keep weight 0 (init_em_kernel and
// the memset of svr_gaussian_reconstruction_local put it there and nothing else writes them).
#define EM_ITEMS 4
template <class Body>
__device__ __forceinline__ void em_for_each(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, uint32_t t, Body&& body)
{
    const uint32_t base = t * EM_ITEMS;
    if (base >= n_valid) return;
    const uint4 q = *reinterpret_cast<const uint4*>(valid_idx + base);       // the list is allocated with 4 entries of slack
    const uint32_t ids[EM_ITEMS] = { q.x, q.y, q.z, q.w };
    const uint32_t cnt = min((uint32_t)EM_ITEMS, n_valid - base);
    uint32_t k = ids[0] / (uint32_t)P, k_end = (k + 1u) * (uint32_t)P;
#pragma unroll
    for (int j = 0; j < EM_ITEMS; ++j) {
        if ((uint32_t)j < cnt) {
            if (ids[j] >= k_end) { k = ids[j] / (uint32_t)P; k_end = (k + 1u) * (uint32_t)P; }
            body(ids[j], (int)k);
        }
    }
}

// Per-slice sums of NV values: a thread keeps the sums of its current slice and sends them to slice_acc when the slice
// changes; at the end a block whose threads all sit in one slice (all but the ~S blocks that straddle a slice boundary)
// folds them with one block reduction and NV double atomics.
template <int NV>
struct SliceSums {
    float v[NV];
    int k;
    __device__ __forceinline__ SliceSums() : k(-1) { for (int i = 0; i < NV; ++i) v[i] = 0.f; }
    __device__ __forceinline__ void flush(double* __restrict__ slice_acc)
    {
        if (k >= 0)
            for (int i = 0; i < NV; ++i)
                if (v[i] != 0.f) atomicAdd(&slice_acc[NV * k + i], (double)v[i]);
        for (int i = 0; i < NV; ++i) v[i] = 0.f;
    }
    __device__ __forceinline__ void enter(int kk, double* __restrict__ slice_acc)
    {
        if (kk != k) { flush(slice_acc); k = kk; }
    }
    __device__ __forceinline__ void finish(double* __restrict__ slice_acc)
    {
        __shared__ int s_k;
        if (threadIdx.x == 0) s_k = k;
        __syncthreads();
        const int k0 = s_k;
        const bool uniform = __syncthreads_and(k == k0 || k < 0) != 0;
        if (uniform && k0 >= 0) {
            typedef cub::BlockReduce<float, EM_THREADS> BR;
            __shared__ typename BR::TempStorage tmp;
            for (int i = 0; i < NV; ++i) {
                const float r = BR(tmp).Sum(v[i]);
                if (threadIdx.x == 0 && r != 0.f) atomicAdd(&slice_acc[NV * k0 + i], (double)r);
                __syncthreads();
            }
        } else {
            flush(slice_acc);
        }
    }
};

static inline int valid_grid(const svr_context* c) { return divup_i(c->n_valid, EM_THREADS * EM_ITEMS); }

// ---- K7 + K8 fused: EStepKernel3D_tex + slice potentials (cuda2.cu:2766-2813, 2816-2911) ------
// slice_acc[2k] += sum (1-w)^2, slice_acc[2k+1] += n over pixels with simweight > 0.99.
__global__ void __launch_bounds__(EM_THREADS)
estep_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, const float* __restrict__ slices,
             const float* __restrict__ simslices, const float* __restrict__ simweights, const float* __restrict__ scales,
             float m_, float sigma_, float mix_, float* __restrict__ weights, double* __restrict__ slice_acc, int flavor)
{
    const float step = flavor == 0 ? SVR_STEP : 0.00001f;
    const float m = M_(m_, step);
    SliceSums<2> acc;
    em_for_each(n_valid, valid_idx, P, blockIdx.x * EM_THREADS + threadIdx.x, [&](uint32_t i, int k) {
        acc.enter(k, slice_acc);
        const float s = slices[i];
        const float sw = simweights[i];
        float w;
        if (flavor == 0) {
            w = 0.0f;                                     // cudaMemsetAsync(weights, 0) at cuda2.cu:2881
            if (!(sw <= 0.0f)) {
                const float e = s * scales[k] - simslices[i];
                const float g = G_(e, sigma_, step);
                w = (g * mix_) / (g * mix_ + m * (1.0f - mix_));
            }
        } else {
            // PVR EStepKernel gates on the PREVIOUS weight and leaves gated pixels untouched
            // (patchBasedRobustStatistics_gpu.cu:121-124)
            w = weights[i];
            if (!(w <= 0.0f)) {
                const float e = s * scales[k] - simslices[i];
                const float g = G_(e, sigma_, step);
                w = (float)((double)(g * mix_) / ((double)(g * mix_) + (double)m * (1.0 - (double)mix_)));
            }
        }
        weights[i] = w;
        if ((double)sw > 0.99) {                          // transformSlicePotential compares against a double literal
            const float d = 1.0f - w;
            acc.v[0] += d * d;
            acc.v[1] += 1.0f;
        }
    });
    acc.finish(slice_acc);
}
__global__ void potential_finish_kernel(int S, const double* __restrict__ slice_acc, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    const float s = (float)slice_acc[2 * k], n = (float)slice_acc[2 * k + 1];
    out[k] = (n > 0.f) ? sqrtf(s / n) : -1.0f;            // cuda2.cu:2903-2910
}

int svr_launch_estep(svr_context* c, float m, float sigma, float mix)
{
    ProfScope prof(c, 6);
    double* acc = c->partials;                            // [2*S] doubles, zeroed
    SVR_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c->S, c->stream));
    if (c->n_valid) {
        estep_kernel<<<valid_grid(c), EM_THREADS, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx * c->Ny, c->slices, c->simslices,
                                                                  c->simweights, c->scales, m, sigma, mix, c->weights, acc, c->flavor);
        SVR_KERNEL_CHECK(c);
    }
    potential_finish_kernel<<<divup_i(c->S, 128), 128, 0, c->stream>>>(c->S, acc, c->slice_tmp);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- K10: CalculateScaleVector (cuda2.cu:3142-3239) --------------------------------------------
__global__ void __launch_bounds__(EM_THREADS)
scale_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, const float* __restrict__ slices,
             const float* __restrict__ weights, const float* __restrict__ simslices, const float* __restrict__ simweights,
             double* __restrict__ slice_acc)
{
    SliceSums<2> acc;
    em_for_each(n_valid, valid_idx, P, blockIdx.x * EM_THREADS + threadIdx.x, [&](uint32_t i, int k) {
        acc.enter(k, slice_acc);
        if (simweights[i] <= 0.99f) return;
        const float s = slices[i], w = weights[i], ss = simslices[i];
        acc.v[0] += w * s * ss;
        acc.v[1] += w * s * s;
    });
    acc.finish(slice_acc);
}
__global__ void scale_finish_kernel(int S, const double* __restrict__ slice_acc, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    const float n = (float)slice_acc[2 * k], d = (float)slice_acc[2 * k + 1];
    out[k] = (d != 0.0f) ? n / d : 1.0f;                  // cuda2.cu:3229-3236
}
int svr_launch_scale(svr_context* c)
{
    ProfScope prof(c, 8);
    double* acc = c->partials;
    SVR_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c->S, c->stream));
    if (c->n_valid) {
        scale_kernel<<<valid_grid(c), EM_THREADS, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx * c->Ny, c->slices, c->weights,
                                                                  c->simslices, c->simweights, acc);
        SVR_KERNEL_CHECK(c);
    }
    scale_finish_kernel<<<divup_i(c->S, 128), 128, 0, c->stream>>>(c->S, acc, c->slice_tmp);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- global two-stage reductions --------------------------------------------------------------
// Stage 1 writes one partial tuple per CTA; stage 2 (one CTA) folds them in a fixed order.
template <int NV>
__global__ void __launch_bounds__(EM_THREADS)
fold_partials_kernel(int nblocks, const double* __restrict__ partials, double* __restrict__ out, int minmax_from)
{
    // values [0, minmax_from) are sums; minmax_from = min, minmax_from + 1 = max (if present)
    double v[NV];
    for (int j = 0; j < NV; ++j) v[j] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        for (int j = 0; j < NV; ++j) {
            const double p = partials[(size_t)b * NV + j];
            if (j < minmax_from) v[j] += p;
            else if (j == minmax_from) v[j] = fmin(v[j], p);
            else v[j] = fmax(v[j], p);
        }
    typedef cub::BlockReduce<double, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    for (int j = 0; j < NV; ++j) {
        double r;
        if (j < minmax_from) r = BR(tmp).Sum(v[j]);
        else if (j == minmax_from) r = BR(tmp).Reduce(v[j], cub::Min());
        else r = BR(tmp).Reduce(v[j], cub::Max());
        if (threadIdx.x == 0) out[j] = r;
        __syncthreads();
    }
}

// ---- K9: MStep statistics (cuda2.cu:2966-3112): {sum e^2 w, sum w, n, min e, max e}; min/max seeded with 0.
__global__ void __launch_bounds__(EM_THREADS)
mstep_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, const float* __restrict__ slices,
             const float* __restrict__ weights, const float* __restrict__ simslices, const float* __restrict__ simweights,
             const float* __restrict__ mstep_scales, double* __restrict__ partials)
{
    float sigma = 0.f, mix = 0.f, num = 0.f, mn = 0.f, mx = 0.f;
    const uint32_t nthreads = gridDim.x * EM_THREADS;
    for (uint32_t t = blockIdx.x * EM_THREADS + threadIdx.x; t * EM_ITEMS < n_valid; t += nthreads)
        em_for_each(n_valid, valid_idx, P, t, [&](uint32_t i, int k) {
            if (simweights[i] > 0.99f) {
                const float e = slices[i] * mstep_scales[k] - simslices[i];
                const float w = weights[i];
                sigma += e * e * w;
                mix += w;
                num += 1.0f;
                mn = fminf(mn, e);
                mx = fmaxf(mx, e);
            }
        });
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float r0 = BR(tmp).Sum(sigma); __syncthreads();
    const float r1 = BR(tmp).Sum(mix); __syncthreads();
    const float r2 = BR(tmp).Sum(num); __syncthreads();
    const float r3 = BR(tmp).Reduce(mn, cub::Min()); __syncthreads();
    const float r4 = BR(tmp).Reduce(mx, cub::Max());
    if (threadIdx.x == 0) {
        double* p = partials + (size_t)blockIdx.x * 5;
        p[0] = r0; p[1] = r1; p[2] = r2; p[3] = r3; p[4] = r4;
    }
}

static int read_back(svr_context* c, const double* dsrc, double* out, int n)
{
    SVR_CUDA(c, cudaMemcpyAsync(c->pinned, dsrc, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) out[i] = ((const double*)c->pinned)[i];
    return 0;
}

int svr_launch_mstep(svr_context* c, double out5[5])
{
    ProfScope prof(c, 7);
    const int nb = c->sm_count * 4;
    double* part = c->partials;                           // [nb*5] + result [8]
    double* res = c->partials + (size_t)nb * 5;
    mstep_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx * c->Ny, c->slices, c->weights, c->simslices,
                                                   c->simweights, c->scales_mstep, part);
    SVR_KERNEL_CHECK(c);
    fold_partials_kernel<5><<<1, EM_THREADS, 0, c->stream>>>(nb, part, res, 3);
    SVR_KERNEL_CHECK(c);
    return read_back(c, res, out5, 5);
}

// ---- K11: InitializeRobustStatistics (cuda2.cu:2243-2308): {sum (s - sim)^2, n} -----------------
__global__ void __launch_bounds__(EM_THREADS)
robust_init_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, const float* __restrict__ slices,
                   const unsigned char* __restrict__ siminside, const float* __restrict__ simslices,
                   const float* __restrict__ simweights, double* __restrict__ partials)
{
    float sa = 0.f, sb = 0.f;
    const uint32_t nthreads = gridDim.x * EM_THREADS;
    for (uint32_t t = blockIdx.x * EM_THREADS + threadIdx.x; t * EM_ITEMS < n_valid; t += nthreads)
        em_for_each(n_valid, valid_idx, P, t, [&](uint32_t i, int) {
            if (siminside[i] == 1 && (double)simweights[i] > 0.99) {
                const float d = slices[i] - simslices[i];
                sa += d * d;
                sb += 1.0f;
            }
        });
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float r0 = BR(tmp).Sum(sa); __syncthreads();
    const float r1 = BR(tmp).Sum(sb);
    if (threadIdx.x == 0) { partials[(size_t)blockIdx.x * 2] = r0; partials[(size_t)blockIdx.x * 2 + 1] = r1; }
}
int svr_launch_robust_init(svr_context* c, double out2[2])
{
    ProfScope prof(c, 9);
    const int nb = c->sm_count * 4;
    double* part = c->partials;
    double* res = c->partials + (size_t)nb * 5;
    robust_init_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx * c->Ny, c->slices, c->siminside, c->simslices,
                                                         c->simweights, part);
    SVR_KERNEL_CHECK(c);
    fold_partials_kernel<2><<<1, EM_THREADS, 0, c->stream>>>(nb, part, res, 2);
    SVR_KERNEL_CHECK(c);
    return read_back(c, res, out2, 2);
}

// ---- K16: ScaleVolume sums (cuda2.cu:3386-3413, 3451-3454) --------------------------------------
__global__ void __launch_bounds__(EM_THREADS)
scale_volume_sums_kernel(uint32_t n_valid, const uint32_t* __restrict__ valid_idx, int P, const float* __restrict__ slices,
                         const float* __restrict__ weights, const float* __restrict__ simslices,
                         const float* __restrict__ simweights, const float* __restrict__ slice_weights, double* __restrict__ partials)
{
    float num = 0.f, den = 0.f;
    const uint32_t nthreads = gridDim.x * EM_THREADS;
    for (uint32_t t = blockIdx.x * EM_THREADS + threadIdx.x; t * EM_ITEMS < n_valid; t += nthreads)
        em_for_each(n_valid, valid_idx, P, t, [&](uint32_t i, int k) {
            if ((double)simweights[i] <= 0.99) return;
            const float s = slices[i], ss = simslices[i], w = weights[i], sw = slice_weights[k];
            num += w * sw * s * ss;
            den += w * sw * ss * ss;
        });
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float r0 = BR(tmp).Sum(num); __syncthreads();
    const float r1 = BR(tmp).Sum(den);
    if (threadIdx.x == 0) { partials[(size_t)blockIdx.x * 2] = r0; partials[(size_t)blockIdx.x * 2 + 1] = r1; }
}
int svr_launch_scale_volume_sums(svr_context* c, double out2[2])
{
    const int nb = c->sm_count * 4;
    double* part = c->partials;
    double* res = c->partials + (size_t)nb * 5;
    scale_volume_sums_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->n_valid, c->valid_idx, c->Nx * c->Ny, c->slices, c->weights, c->simslices,
                                                               c->simweights, c->slice_weights, part);
    SVR_KERNEL_CHECK(c);
    fold_partials_kernel<2><<<1, EM_THREADS, 0, c->stream>>>(nb, part, res, 2);
    SVR_KERNEL_CHECK(c);
    return read_back(c, res, out2, 2);
}

__global__ void scale_volume_apply_kernel(size_t V, float* __restrict__ recon, float scale)
{   // scaleVolumeKernel (cuda2.cu:3415-3423)
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        const float r = recon[v];
        if (r > 0.f) recon[v] = r * scale;
    }
}
int svr_launch_scale_volume_apply(svr_context* c, float scale)
{
    scale_volume_apply_kernel<<<c->sm_count * 8, EM_THREADS, 0, c->stream>>>(c->V, c->recon, scale);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- K15: maskVolumeKernel (cuda2.cu:3313-3326) --------------------------------------------------
__global__ void mask_volume_kernel(size_t V, float* __restrict__ recon, const unsigned char* __restrict__ mask)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x)
        if (mask[v] == 0) recon[v] = -1.0f;
}
int svr_launch_mask_volume(svr_context* c)
{
    mask_volume_kernel<<<c->sm_count * 8, EM_THREADS, 0, c->stream>>>(c->V, c->recon, c->mask_u8);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- K17: RestoreSliceIntensitiesKernel (cuda2.cu:3349-3367) on the restore copy -----------------
__global__ void restore_kernel(size_t NP, int P, float* __restrict__ slices, const float* __restrict__ factors,
                               const int* __restrict__ stack_index)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < NP; i += (size_t)gridDim.x * blockDim.x) {
        const float s = slices[i];
        if (s > 0.f) slices[i] = s / factors[stack_index[i / P]];
    }
}
int svr_launch_restore(svr_context* c, const float* d_factors, const int* d_index)
{
    restore_kernel<<<c->sm_count * 8, EM_THREADS, 0, c->stream>>>(c->NP, c->Nx * c->Ny, c->slices_restore, d_factors, d_index);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- K4 + K5: regulariser (cuda2.cu:1944-1969, 2046-2117; constants cuda2.cu:666-695) -------------
// K4 reads recon (= `original`, untouched) and the accumulator, writes the post-step volume to
// recon_tmp1 and normalises the accumulator in place.  K5 reads original / post-step / cmap and
// writes recon_tmp2; the caller swaps recon <-> recon_tmp2.  (Deviation D3: neighbours come from
// the frozen post-step copy, like the reference's CPU twin.)
__constant__ int c_dirs[13][3] = {
    { 1, 0, -1 }, { 0, 1, -1 }, { 1, 1, -1 }, { 1, -1, -1 }, { 1, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 },
    { 1, -1, 0 }, { 1, 0, 1 }, { 0, 1, 1 }, { 1, 1, 1 }, { 1, -1, 1 }, { 0, 0, 1 } };

__global__ void reg_prep_kernel(size_t V, const float* __restrict__ recon, float2* __restrict__ acc2,
                                const unsigned char* __restrict__ mask, float* __restrict__ post, int adaptive,
                                float alpha, float min_i, float max_i, int flavor)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) {
        float2 a = acc2[v];
        if (!mask[v]) a = make_float2(0.f, 0.f);           // the per-tap mask test of cuda2.cu:509-510, applied per voxel
        if (!adaptive && a.y != 0.f) { a.x = a.x / a.y; a.y = 1.0f; }
        acc2[v] = a;
        float r = recon[v] + a.x * alpha;
        if (flavor == 0) {     // double literals in cuda2.cu:1962-1965, float literals in patchBasedSuperresolution_gpu.cu:176-179
            if ((double)r < (double)min_i * 0.9) r = (float)((double)min_i * 0.9);
            if ((double)r > (double)max_i * 1.1) r = (float)((double)max_i * 1.1);
        } else {
            if (r < min_i * 0.9f) r = min_i * 0.9f;
            if (r > max_i * 1.1f) r = max_i * 1.1f;
        }
        post[v] = r;
    }
}

__global__ void __launch_bounds__(256)
reg_kernel(int vx, int vy, int vz, const float* __restrict__ original, const float* __restrict__ post,
           const float2* __restrict__ acc2, float* __restrict__ out, float delta, float alpha, float lambda)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int z = blockIdx.z;
    if (x >= vx || y >= vy || z >= vz) return;
    const size_t p = x + (size_t)y * vx + (size_t)z * vx * vy;
    const float cp = acc2[p].y, rp = post[p], op = original[p];
    float val = 0.f, valW = 0.f, sum = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const int dx = c_dirs[i][0], dy = c_dirs[i][1], dz = c_dirs[i][2];
        const float f = 1.0f / (float)(abs(dx) + abs(dy) + abs(dz));
        const float sf = sqrtf(f);
        const int x2 = x + dx, y2 = y + dy, z2 = z + dz;
        const bool in2 = x2 >= 0 && x2 < vx && y2 >= 0 && y2 < vy && z2 >= 0 && z2 < vz;
        if (!in2) continue;                               // the mirrored neighbour also needs pos2 inside (cuda2.cu:2089-2093)
        const size_t p2 = x2 + (size_t)y2 * vx + (size_t)z2 * vx * vy;
        const float c2 = acc2[p2].y, o2 = original[p2];
        {
            float bi = 0.f;
            if (!(cp <= 0.f || c2 <= 0.f)) {
                const float diff = (o2 - op) * sf / delta;
                bi = f / sqrtf(1.0f + diff * diff);
            }
            val += bi * post[p2] * c2;
            valW += bi * c2;
            sum += bi;
        }
        const int x3 = x - dx, y3 = y - dy, z3 = z - dz;
        const bool in3 = x3 >= 0 && x3 < vx && y3 >= 0 && y3 < vy && z3 >= 0 && z3 < vz;
        if (in3) {
            const size_t p3 = x3 + (size_t)y3 * vx + (size_t)z3 * vx * vy;
            const float c3 = acc2[p3].y;
            float bi = 0.f;
            if (!(c3 <= 0.f || c2 <= 0.f)) {              // AdaptiveRegularization1(i, pos3, pos2): cuda2.cu:2095
                const float diff = (o2 - original[p3]) * sf / delta;
                bi = f / sqrtf(1.0f + diff * diff);
            }
            val += bi * post[p3] * c3;
            valW += bi * c3;
            sum += bi;
        }
    }
    val -= sum * rp * cp;
    valW -= sum * cp;
    const float k = alpha * lambda / (delta * delta);
    val = rp * cp + k * val;
    valW = cp + k * valW;
    out[p] = (valW > 0.0f) ? val / valW : 0.0f;
}

// Fused K4 + K5 (round 2).  reg_prep_kernel + reg_kernel above read, per voxel, 26 neighbours of three volumes through L1/L2
// (measured 1.22 ms at 256^3 = 6 % of the HBM rate for its 28 V algorithmic bytes).  Here a CTA stages a tile of
// 32 x 8 x 8 voxels plus a one-voxel halo in shared memory -- computing the K4 update (mask, normalisation, gradient step,
// clamp) on the fly for tile AND halo, so the post-step volume is never written to HBM -- and evaluates the 13-direction
// smoothing from shared memory.  Same arithmetic and operation order per voxel as the two kernels (they remain the
// fallback and the parity reference: tests/test_gpu_parity.py compares both against the oracle).
constexpr int RT_X = 32, RT_Y = 8, RT_Z = 8;
constexpr int RH_X = RT_X + 2, RH_Y = RT_Y + 2, RH_Z = RT_Z + 2;

__global__ void __launch_bounds__(256)
regularize_fused_kernel(int vx, int vy, int vz, const float* __restrict__ recon, float2* __restrict__ acc2,
                        const unsigned char* __restrict__ mask, float* __restrict__ out, int adaptive, float alpha, float min_i,
                        float max_i, float delta, float lambda, int flavor)
{
    __shared__ float s_org[RH_Z][RH_Y][RH_X];      // original (untouched) volume
    __shared__ float s_post[RH_Z][RH_Y][RH_X];     // after the gradient step + clamp (K4)
    __shared__ float s_cm[RH_Z][RH_Y][RH_X];       // confidence map after K4 (-1 marks voxels outside the volume)
    const int bx0 = blockIdx.x * RT_X - 1, by0 = blockIdx.y * RT_Y - 1, bz0 = blockIdx.z * RT_Z - 1;
    for (int i = threadIdx.x; i < RH_X * RH_Y * RH_Z; i += blockDim.x) {
        const int lx = i % RH_X, ly = (i / RH_X) % RH_Y, lz = i / (RH_X * RH_Y);
        const int x = bx0 + lx, y = by0 + ly, z = bz0 + lz;
        float org = 0.f, post = 0.f, cm = -1.f;
        if (x >= 0 && x < vx && y >= 0 && y < vy && z >= 0 && z < vz) {
            const size_t v = x + (size_t)y * vx + (size_t)z * vx * vy;
            float2 a = acc2[v];
            if (!mask[v]) a = make_float2(0.f, 0.f);
            if (!adaptive && a.y != 0.f) { a.x = a.x / a.y; a.y = 1.0f; }
            org = recon[v];
            float r = org + a.x * alpha;
            if (flavor == 0) {
                if ((double)r < (double)min_i * 0.9) r = (float)((double)min_i * 0.9);
                if ((double)r > (double)max_i * 1.1) r = (float)((double)max_i * 1.1);
            } else {
                if (r < min_i * 0.9f) r = min_i * 0.9f;
                if (r > max_i * 1.1f) r = max_i * 1.1f;
            }
            post = r; cm = a.y;
            // the normalised accumulator is part of the interface (debug taps, adaptive mode): the tile's own voxels write it back
            if (lx >= 1 && lx <= RT_X && ly >= 1 && ly <= RT_Y && lz >= 1 && lz <= RT_Z) acc2[v] = a;
        }
        s_org[lz][ly][lx] = org; s_post[lz][ly][lx] = post; s_cm[lz][ly][lx] = cm;
    }
    __syncthreads();
    const float kreg = alpha * lambda / (delta * delta);
    // b = f / sqrt(1 + (d sqrt(f) / delta)^2) as f * rsqrt.approx(...) with the constant factor folded: the reference is built
    // with --use_fast_math (div.approx, sqrt.approx), so neither form is "the" rounding; 26 evaluations per voxel made the
    // IEEE division + square root of the two-kernel version the bound of this kernel (1.0 ms at 256^3 against 0.1 ms of HBM time)
    const float inv_delta = 1.0f / delta;
    for (int i = threadIdx.x; i < RT_X * RT_Y * RT_Z; i += blockDim.x) {
        const int lx = 1 + i % RT_X, ly = 1 + (i / RT_X) % RT_Y, lz = 1 + i / (RT_X * RT_Y);
        const int x = bx0 + lx, y = by0 + ly, z = bz0 + lz;
        if (x >= vx || y >= vy || z >= vz) continue;
        const float cp = s_cm[lz][ly][lx], rp = s_post[lz][ly][lx], op = s_org[lz][ly][lx];
        float val = 0.f, valW = 0.f, sum = 0.f;
#pragma unroll
        for (int d = 0; d < 13; ++d) {
            const int dx = c_dirs[d][0], dy = c_dirs[d][1], dz = c_dirs[d][2];
            const int nn = abs(dx) + abs(dy) + abs(dz);
            const float f = nn == 1 ? 1.0f : (nn == 2 ? 0.5f : 0.33333334f);
            const float sfd = (nn == 1 ? 1.0f : (nn == 2 ? 0.70710677f : 0.57735026f)) * inv_delta;      // sqrt(f) / delta
            const float c2 = s_cm[lz + dz][ly + dy][lx + dx];
            if (c2 < 0.f) continue;                       // pos2 outside the volume: neither term counts (cuda2.cu:2089-2093)
            const float o2 = s_org[lz + dz][ly + dy][lx + dx];
            {
                float bi = 0.f;
                if (!(cp <= 0.f || c2 <= 0.f)) {
                    const float diff = (o2 - op) * sfd;
                    bi = f * rsqrt_approx(fmaf(diff, diff, 1.0f));
                }
                val += bi * s_post[lz + dz][ly + dy][lx + dx] * c2;
                valW += bi * c2;
                sum += bi;
            }
            const float c3 = s_cm[lz - dz][ly - dy][lx - dx];
            if (c3 >= 0.f) {
                float bi = 0.f;
                if (!(c3 <= 0.f || c2 <= 0.f)) {          // AdaptiveRegularization1(i, pos3, pos2): cuda2.cu:2095
                    const float diff = (o2 - s_org[lz - dz][ly - dy][lx - dx]) * sfd;
                    bi = f * rsqrt_approx(fmaf(diff, diff, 1.0f));
                }
                val += bi * s_post[lz - dz][ly - dy][lx - dx] * c3;
                valW += bi * c3;
                sum += bi;
            }
        }
        val -= sum * rp * cp;
        valW -= sum * cp;
        val = rp * cp + kreg * val;
        valW = cp + kreg * valW;
        out[x + (size_t)y * vx + (size_t)z * vx * vy] = (valW > 0.0f) ? val / valW : 0.0f;
    }
}

int svr_launch_regularize(svr_context* c, int adaptive, float alpha, float min_i, float max_i, float delta, float lambda)
{
    ProfScope prof(c, 3);
    if (c->tune_regularize == 0) {
        reg_prep_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->recon, c->acc2, c->mask_u8, c->recon_tmp1, adaptive, alpha, min_i, max_i, c->flavor);
        SVR_KERNEL_CHECK(c);
        dim3 block(64, 4, 1), grid(divup_i(c->vx, 64), divup_i(c->vy, 4), c->vz);
        reg_kernel<<<grid, block, 0, c->stream>>>(c->vx, c->vy, c->vz, c->recon, c->recon_tmp1, c->acc2, c->recon_tmp2, delta, alpha, lambda);
        SVR_KERNEL_CHECK(c);
    } else {
        dim3 grid(divup_i(c->vx, RT_X), divup_i(c->vy, RT_Y), divup_i(c->vz, RT_Z));
        regularize_fused_kernel<<<grid, 256, 0, c->stream>>>(c->vx, c->vy, c->vz, c->recon, c->acc2, c->mask_u8, c->recon_tmp2, adaptive, alpha,
                                                             min_i, max_i, delta, lambda, c->flavor);
        SVR_KERNEL_CHECK(c);
    }
    float* t = c->recon; c->recon = c->recon_tmp2; c->recon_tmp2 = t;
    return 0;
}

// ---- valid-pixel compaction (runs once per svr_fill_slices) ---------------------------------------
// Two passes with a per-CTA count + exclusive scan on the host side of the small count vector would be
// overkill: a single ordered pass with a decoupled look-back is what cub::DeviceSelect does.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
struct NotPadding {
    const float* s;
    __device__ bool operator()(uint32_t i) const { return s[i] != -1.0f; }
};
// even-x pixel e such that e or e + 1 (same slice row) is valid: one unit of the paired scatter (svr_psf.cu)
struct PairHead {
    const float* s;
    uint32_t Nx;
    __device__ bool operator()(uint32_t i) const
    {
        const uint32_t x = i % Nx;
        if (x & 1u) return false;
        return s[i] != -1.0f || (x + 1 < Nx && s[i + 1] != -1.0f);
    }
};
int svr_launch_compact_valid(svr_context* c)
{
    thrust::counting_iterator<uint32_t> it(0);
    NotPadding pred{ c->slices };
    int* d_num = (int*)(c->partials);
    size_t need = 0;
    SVR_CUDA(c, cub::DeviceSelect::If(nullptr, need, it, c->valid_idx, d_num, (int)c->NP, pred, c->stream));
    if (need > c->cub_tmp_bytes) {
        if (c->cub_tmp) cudaFree(c->cub_tmp);
        SVR_CUDA(c, cudaMalloc(&c->cub_tmp, need));
        c->cub_tmp_bytes = need;
    }
    SVR_CUDA(c, cub::DeviceSelect::If(c->cub_tmp, need, it, c->valid_idx, d_num, (int)c->NP, pred, c->stream));
    c->launches += 2;
    int h = 0;
    SVR_CUDA(c, cudaMemcpyAsync(&h, d_num, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_valid = (uint32_t)h;
    PairHead ph{ c->slices, (uint32_t)c->Nx };
    SVR_CUDA(c, cub::DeviceSelect::If(nullptr, need, it, c->pair_idx, d_num, (int)c->NP, ph, c->stream));
    if (need > c->cub_tmp_bytes) {
        if (c->cub_tmp) cudaFree(c->cub_tmp);
        SVR_CUDA(c, cudaMalloc(&c->cub_tmp, need));
        c->cub_tmp_bytes = need;
    }
    SVR_CUDA(c, cub::DeviceSelect::If(c->cub_tmp, need, it, c->pair_idx, d_num, (int)c->NP, ph, c->stream));
    c->launches += 2;
    SVR_CUDA(c, cudaMemcpyAsync(&h, d_num, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_pairs = (uint32_t)h;
    return 0;
}

__global__ void flags_to_int_kernel(size_t n, const unsigned char* __restrict__ src, int* __restrict__ dst)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int svr_launch_flags_to_int(svr_context* c, const unsigned char* src, int* dst, size_t n)
{
    flags_to_int_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(n, src, dst);
    SVR_KERNEL_CHECK(c);
    return 0;
}
