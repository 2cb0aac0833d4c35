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

// ---- K7 + K8 fused: EStepKernel3D_tex + slice potentials (cuda2.cu:2766-2813, 2816-2911) ------
// grid = (chunks, S).  slice_acc[2k] += sum (1-w)^2, slice_acc[2k+1] += n over pixels with simweight > 0.99.
__global__ void __launch_bounds__(EM_THREADS)
estep_kernel(int P, const float* __restrict__ slices, const float* __restrict__ simslices,
             const float* __restrict__ simweights, const float* __restrict__ scales, float m_, float sigma_, float mix_,
             float* __restrict__ weights, double* __restrict__ slice_acc, int flavor)
{
    const int k = blockIdx.y;
    const size_t base = (size_t)k * P;
    const float scale = scales[k];
    const float step = flavor == 0 ? SVR_STEP : 0.00001f;
    const float m = M_(m_, step);
    float sum = 0.f, num = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float s = slices[base + i];
        const float sw = simweights[base + i];
        float w;
        if (flavor == 0) {
            w = 0.0f;                                     // cudaMemsetAsync(weights, 0) at cuda2.cu:2881
            if (!((s == -1.0f) || sw <= 0.0f)) {
                const float e = s * scale - simslices[base + i];
                const float g = G_(e, sigma_, step);
                w = (g * mix_) / (g * mix_ + m * (1.0f - mix_));
            }
        } else {
            // PVR EStepKernel gates on the PREVIOUS weight and leaves gated pixels untouched
            // (patchBasedRobustStatistics_gpu.cu:121-124)
            w = weights[base + i];
            if (!((s == -1.0f) || w <= 0.0f)) {
                const float e = s * scale - simslices[base + i];
                const float g = G_(e, sigma_, step);
                w = (float)((double)(g * mix_) / ((double)(g * mix_) + (double)m * (1.0 - (double)mix_)));
            }
        }
        weights[base + i] = w;
        if ((double)sw > 0.99) {                          // transformSlicePotential compares against a double literal
            const float d = 1.0f - w;
            sum += d * d;
            num += 1.0f;
        }
    }
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float bs = BR(tmp).Sum(sum);
    __syncthreads();
    const float bn = BR(tmp).Sum(num);
    if (threadIdx.x == 0 && bn > 0.f) {
        atomicAdd(&slice_acc[2 * k], (double)bs);
        atomicAdd(&slice_acc[2 * k + 1], (double)bn);
    }
}
__global__ void potential_finish_kernel(int S, const double* __restrict__ slice_acc, float* __restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S) return;
    const float s = (float)slice_acc[2 * k], n = (float)slice_acc[2 * k + 1];
    out[k] = (n > 0.f) ? sqrtf(s / n) : -1.0f;            // cuda2.cu:2903-2910
}

static inline dim3 slice_grid(const svr_context* c)
{
    const int P = c->Nx * c->Ny;
    int chunks = divup_i(P, EM_THREADS * 4);
    const int want = divup_i(c->sm_count * 4, c->S > 0 ? c->S : 1);
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    return dim3(chunks, c->S);
}

int svr_launch_estep(svr_context* c, float m, float sigma, float mix)
{
    ProfScope prof(c, 4);
    double* acc = c->partials;                            // [2*S] doubles, zeroed
    SVR_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c->S, c->stream));
    estep_kernel<<<slice_grid(c), EM_THREADS, 0, c->stream>>>(c->Nx * c->Ny, c->slices, c->simslices, c->simweights,
                                                              c->scales, m, sigma, mix, c->weights, acc, c->flavor);
    SVR_KERNEL_CHECK(c);
    potential_finish_kernel<<<divup_i(c->S, 128), 128, 0, c->stream>>>(c->S, acc, c->slice_tmp);
    SVR_KERNEL_CHECK(c);
    return 0;
}

// ---- K10: CalculateScaleVector (cuda2.cu:3142-3239) --------------------------------------------
__global__ void __launch_bounds__(EM_THREADS)
scale_kernel(int P, const float* __restrict__ slices, const float* __restrict__ weights,
             const float* __restrict__ simslices, const float* __restrict__ simweights, double* __restrict__ slice_acc)
{
    const int k = blockIdx.y;
    const size_t base = (size_t)k * P;
    float num = 0.f, den = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
        const float s = slices[base + i];
        const float sw = simweights[base + i];
        if ((s == -1.0f) || sw <= 0.99f) continue;
        const float w = weights[base + i], ss = simslices[base + i];
        num += w * s * ss;
        den += w * s * s;
    }
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float bn = BR(tmp).Sum(num);
    __syncthreads();
    const float bd = BR(tmp).Sum(den);
    if (threadIdx.x == 0 && (bn != 0.f || bd != 0.f)) {
        atomicAdd(&slice_acc[2 * k], (double)bn);
        atomicAdd(&slice_acc[2 * k + 1], (double)bd);
    }
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
    ProfScope prof(c, 4);
    double* acc = c->partials;
    SVR_CUDA(c, cudaMemsetAsync(acc, 0, sizeof(double) * 2 * c->S, c->stream));
    scale_kernel<<<slice_grid(c), EM_THREADS, 0, c->stream>>>(c->Nx * c->Ny, c->slices, c->weights, c->simslices,
                                                              c->simweights, acc);
    SVR_KERNEL_CHECK(c);
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
mstep_kernel(size_t NP, int P, const float* __restrict__ slices, const float* __restrict__ weights,
             const float* __restrict__ simslices, const float* __restrict__ simweights,
             const float* __restrict__ mstep_scales, double* __restrict__ partials)
{
    float sigma = 0.f, mix = 0.f, num = 0.f, mn = 0.f, mx = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < NP; i += (size_t)gridDim.x * blockDim.x) {
        const float s = slices[i], sw = simweights[i];
        if (s != -1.0f && sw > 0.99f) {
            const float e = s * mstep_scales[i / P] - simslices[i];
            const float w = weights[i];
            sigma += e * e * w;
            mix += w;
            num += 1.0f;
            mn = fminf(mn, e);
            mx = fmaxf(mx, e);
        }
    }
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
    ProfScope prof(c, 4);
    const int nb = c->sm_count * 4;
    double* part = c->partials;                           // [nb*5] + result [8]
    double* res = c->partials + (size_t)nb * 5;
    mstep_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->NP, c->Nx * c->Ny, c->slices, c->weights, c->simslices,
                                                   c->simweights, c->scales_mstep, part);
    SVR_KERNEL_CHECK(c);
    fold_partials_kernel<5><<<1, EM_THREADS, 0, c->stream>>>(nb, part, res, 3);
    SVR_KERNEL_CHECK(c);
    return read_back(c, res, out5, 5);
}

// ---- K11: InitializeRobustStatistics (cuda2.cu:2243-2308): {sum (s - sim)^2, n} -----------------
__global__ void __launch_bounds__(EM_THREADS)
robust_init_kernel(size_t NP, const float* __restrict__ slices, const unsigned char* __restrict__ siminside,
                   const float* __restrict__ simslices, const float* __restrict__ simweights, double* __restrict__ partials)
{
    float sa = 0.f, sb = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < NP; i += (size_t)gridDim.x * blockDim.x) {
        const float s = slices[i];
        if (s != -1.0f && siminside[i] == 1 && (double)simweights[i] > 0.99) {
            const float d = s - simslices[i];
            sa += d * d;
            sb += 1.0f;
        }
    }
    typedef cub::BlockReduce<float, EM_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    const float r0 = BR(tmp).Sum(sa); __syncthreads();
    const float r1 = BR(tmp).Sum(sb);
    if (threadIdx.x == 0) { partials[(size_t)blockIdx.x * 2] = r0; partials[(size_t)blockIdx.x * 2 + 1] = r1; }
}
int svr_launch_robust_init(svr_context* c, double out2[2])
{
    ProfScope prof(c, 4);
    const int nb = c->sm_count * 4;
    double* part = c->partials;
    double* res = c->partials + (size_t)nb * 5;
    robust_init_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->NP, c->slices, c->siminside, c->simslices, c->simweights, part);
    SVR_KERNEL_CHECK(c);
    fold_partials_kernel<2><<<1, EM_THREADS, 0, c->stream>>>(nb, part, res, 2);
    SVR_KERNEL_CHECK(c);
    return read_back(c, res, out2, 2);
}

// ---- K16: ScaleVolume sums (cuda2.cu:3386-3413, 3451-3454) --------------------------------------
__global__ void __launch_bounds__(EM_THREADS)
scale_volume_sums_kernel(size_t NP, int P, const float* __restrict__ slices, const float* __restrict__ weights,
                         const float* __restrict__ simslices, const float* __restrict__ simweights,
                         const float* __restrict__ slice_weights, double* __restrict__ partials)
{
    float num = 0.f, den = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < NP; i += (size_t)gridDim.x * blockDim.x) {
        const float s = slices[i];
        if (s == -1.0f) continue;
        if ((double)simweights[i] <= 0.99) continue;
        const float ss = simslices[i], w = weights[i], sw = slice_weights[i / P];
        num += w * sw * s * ss;
        den += w * sw * ss * ss;
    }
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
    scale_volume_sums_kernel<<<nb, EM_THREADS, 0, c->stream>>>(c->NP, c->Nx * c->Ny, c->slices, c->weights, c->simslices,
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

int svr_launch_regularize(svr_context* c, int adaptive, float alpha, float min_i, float max_i, float delta, float lambda)
{
    ProfScope prof(c, 3);
    reg_prep_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->recon, c->acc2, c->mask_u8, c->recon_tmp1, adaptive, alpha, min_i, max_i, c->flavor);
    SVR_KERNEL_CHECK(c);
    dim3 block(64, 4, 1), grid(divup_i(c->vx, 64), divup_i(c->vy, 4), c->vz);
    reg_kernel<<<grid, block, 0, c->stream>>>(c->vx, c->vy, c->vz, c->recon, c->recon_tmp1, c->acc2, c->recon_tmp2, delta, alpha, lambda);
    SVR_KERNEL_CHECK(c);
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
