// pvr_abi.cu -- the PVR entry points of libsvr_b200.so (include/pvr_abi.h).
//
// A PVR context is an svr_context in flavour 1: the PSF / EM kernels of svr_psf.cu / svr_em.cu are
// instantiated with PvrTraits (12^3 support, PVR constants) and the patch grids of all stacks are one
// "slice cube" of nPatches slices of pbx x pby pixels.  This file adds what only PVR has: the patch
// extraction kernel P0, the accumulate-then-equalize split of the initial reconstruction, the un-offset
// texture read (as an 8-voxel mean, svr_psf.cu) and the host part of the patch-level EM.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <vector>
#include "../../include/pvr_abi.h"
#include "svr_context.h"

struct PvrState {
    int n_stacks = 0;
    std::vector<int> per_stack;       // patches per stack
    std::vector<int> offset;          // first patch of each stack
};

static int pvr_fail(svr_context* c, const char* msg) { if (c) c->err = msg; return 2; }
#define PVR_REQUIRE(c, cond, msg) do { if (!(cond)) return pvr_fail((c), (msg)); } while (0)

void svr_pvr_free(svr_context* c)
{
    if (c->pvr) { delete (PvrState*)c->pvr; c->pvr = nullptr; }
    if (c->spx) { cudaFree(c->spx); c->spx = nullptr; }
}

// ---- P0: patchBasedPatchInitKernel (initPatchBasedRecon_gpu.cu:44-86) for the patches [p0, p0 + n) of one stack.
__device__ __forceinline__ unsigned int pvr_f2u(float f) { return (unsigned int)f; }   // cvt.rzi.u32.f32 saturates like the reference's

__global__ void pvr_patch_init_kernel(int p0, int n, int Nx, int Ny, const float* __restrict__ stack, int sx, int sy, int sz,
                                      const float* __restrict__ stackW2I, const float* __restrict__ mats, size_t matstride,
                                      VolGeom vg, const float* __restrict__ mask_f, const char* __restrict__ spx,
                                      float* __restrict__ patches)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = p0 + blockIdx.z;
    if (x >= Nx || y >= Ny || (int)blockIdx.z >= n) return;
    const float* T = mats + 16 * (size_t)k;                       // Transformation
    const float* I2W = mats + 2 * matstride + 16 * (size_t)k;     // I2W
    // getValueFromPatchCoords (patchBasedVolume.cuh:233-249): stackW2I * p.I2W * scoord, Matrix4 product first
    float m[12];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
            m[4 * i + j] = stackW2I[4 * i + 0] * I2W[0 + j] + stackW2I[4 * i + 1] * I2W[4 + j] + stackW2I[4 * i + 2] * I2W[8 + j] +
                           stackW2I[4 * i + 3] * I2W[12 + j];
    const float3 pos = make_float3((float)x, (float)y, 0.0f);
    const float3 sc = mat_pt(m, pos);
    float s = 0.0f;
    if (sc.x >= 0 && sc.x < sx && sc.y >= 0 && sc.y < sy && sc.z >= 0 && sc.z < sz) {
        const unsigned int idx = pvr_f2u(sc.x + sc.y * sx + sc.z * sx * sy);
        s = stack[idx];
    }
    if (s == -1.0f) return;
    float ti[12];                                                  // patch.Transformation * patch.I2W
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
            ti[4 * i + j] = T[4 * i + 0] * I2W[0 + j] + T[4 * i + 1] * I2W[4 + j] + T[4 * i + 2] * I2W[8 + j] + T[4 * i + 3] * I2W[12 + j];
    const float3 vp = mat_pt(vg.rw2i, mat_pt(ti, pos));
    const unsigned int ax = pvr_f2u(vp.x), ay = pvr_f2u(vp.y), az = pvr_f2u(vp.z);
    bool masked = false;                                           // ReconVolume::isMasked, reconVolume.cuh:241-254
    if (ax < (unsigned)vg.vx && ay < (unsigned)vg.vy && az < (unsigned)vg.vz) {
        const float mv = mask_f[ax + (size_t)ay * vg.vx + (size_t)az * vg.vx * vg.vy];
        masked = !(mv == -1.0f || mv == 0.0f);
    }
    const size_t idx = ((size_t)k * Ny + y) * Nx + x;
    if (spx) {
        const bool on = spx[(size_t)k * 4096 + x + 64 * y] == '1';
        if (masked && on) patches[idx] = s;
        else if (masked && !on) patches[idx] = -1.0f;
    } else if (masked) patches[idx] = s;
}

static int pvr_ready(svr_context* c, const char* who)
{
    if (!c) return 2;
    if (c->flavor != 1) { c->err = std::string(who) + ": not a PVR context (use pvr_create)"; return 2; }
    if (!c->recon || !c->have_mask) { c->err = std::string(who) + ": volume / mask not initialised"; return 2; }
    if (!c->pvr) { c->err = std::string(who) + ": call pvr_patches_init first"; return 2; }
    if (c->S > 0 && (!c->have_mats || !c->have_dims)) { c->err = std::string(who) + ": patch matrices not set"; return 2; }
    return 0;
}

static int sync_(svr_context* c) { SVR_CUDA(c, cudaStreamSynchronize(c->stream)); return 0; }

extern "C" {

int pvr_create(svr_context** out, int device)
{
    const int rc = svr_create(out, device);
    if (rc) return rc;
    (*out)->flavor = 1;
    return 0;
}

int pvr_recon_init(svr_context* c, int sx, int sy, int sz, float dx, float dy, float dz, const float recon_w2i[16],
                   const float recon_i2w[16])
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && recon_w2i && recon_i2w, "pvr_recon_init: NULL argument");
    if (int r = svr_init_reconstruction_volume(c, sx, sy, sz, dx, dy, dz, nullptr)) return r;
    memcpy(c->recon_i2w, recon_i2w, 16 * sizeof(float));
    memcpy(c->recon_w2i, recon_w2i, 16 * sizeof(float));
    memcpy(c->vg.rw2i, recon_w2i, 12 * sizeof(float));
    return 0;
}

int pvr_recon_set_mask(svr_context* c, const signed char* mask)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && mask && c->V > 0, "pvr_recon_set_mask: volume not initialised or NULL mask");
    std::vector<float> f(c->V);
    for (size_t i = 0; i < c->V; ++i) f[i] = (float)mask[i];
    return svr_set_mask(c, c->vx, c->vy, c->vz, f.data());
}

int pvr_recon_reset(svr_context* c)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->recon, "pvr_recon_reset: volume not initialised");
    SVR_CUDA(c, cudaMemsetAsync(c->recon, 0, c->V * sizeof(float), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->volw, 0, c->V * sizeof(float), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->acc2, 0, c->V * sizeof(float2), c->stream));
    return sync_(c);
}

int pvr_recon_reset_addon_cmap(svr_context* c)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->recon, "pvr_recon_reset_addon_cmap: volume not initialised");
    SVR_CUDA(c, cudaMemsetAsync(c->acc2, 0, c->V * sizeof(float2), c->stream));
    return sync_(c);
}

int pvr_recon_equalize(svr_context* c)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->recon, "pvr_recon_equalize: volume not initialised");
    if (svr_launch_equalize_inplace(c)) return 1;
    return sync_(c);
}

int pvr_recon_copy_from_host(svr_context* c, const float* data) { return svr_update_reconstructed(c, data); }
int pvr_recon_copy_to_host(svr_context* c, float* data) { return svr_sync_cpu(c, data); }

int pvr_patches_init(svr_context* c, int pbx, int pby, int n_stacks, const int* patches_per_stack, const float* stack_dims)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->flavor == 1, "pvr_patches_init: not a PVR context");
    PVR_REQUIRE(c, pbx > 0 && pby > 0 && n_stacks >= 0 && (n_stacks == 0 || (patches_per_stack && stack_dims)),
                "pvr_patches_init: bad argument");
    PVR_REQUIRE(c, pbx <= 64 && pby <= 64, "pvr_patches_init: patches larger than 64x64 are not supported (spxMask is char[64*64])");
    svr_pvr_free(c);
    PvrState* p = new PvrState();
    c->pvr = p;
    p->n_stacks = n_stacks;
    int total = 0;
    for (int i = 0; i < n_stacks; ++i) {
        PVR_REQUIRE(c, patches_per_stack[i] >= 0, "pvr_patches_init: negative patch count");
        p->per_stack.push_back(patches_per_stack[i]);
        p->offset.push_back(total);
        total += patches_per_stack[i];
    }
    if (int r = svr_init_storage_volumes(c, pbx, pby, total)) return r;
    std::vector<float> dims((size_t)std::max(total, 1) * 3);
    for (int i = 0; i < n_stacks; ++i)
        for (int j = 0; j < patches_per_stack[i]; ++j)
            for (int q = 0; q < 3; ++q) dims[(size_t)(p->offset[i] + j) * 3 + q] = stack_dims[3 * i + q];
    if (int r = svr_set_slice_dims(c, dims.data(), 1.0f)) return r;
    if (total) {
        SVR_CUDA(c, cudaMalloc((void**)&c->spx, (size_t)total * 4096));
        SVR_CUDA(c, cudaMemsetAsync(c->spx, '0', (size_t)total * 4096, c->stream));     // char spxMask[64*64] = {'0'}
    }
    c->use_spx = 0;
    // the patch buffer starts at 0 (PatchBasedVolume::reset), psf sums at 0; scale = weight = 1
    std::vector<float> ones((size_t)std::max(total, 1), 1.0f);
    if (int r = svr_update_scale_vector(c, ones.data(), ones.data())) return r;
    return sync_(c);
}

int pvr_patches_set_matrices(svr_context* c, const float* i2w, const float* w2i, const float* transformation,
                             const float* inv_transformation)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->pvr, "pvr_patches_set_matrices: call pvr_patches_init first");
    return svr_set_slice_matrices(c, transformation, inv_transformation, i2w, w2i, c->recon_i2w, c->recon_w2i);
}

int pvr_patches_set_spx_masks(svr_context* c, const char* masks, int use_spx)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->pvr, "pvr_patches_set_spx_masks: call pvr_patches_init first");
    if (masks && c->S) {
        SVR_CUDA(c, cudaMemcpyAsync(c->spx, masks, (size_t)c->S * 4096, cudaMemcpyHostToDevice, c->stream));
        if (int r = sync_(c)) return r;
    }
    c->use_spx = use_spx ? 1 : 0;
    return 0;
}

int pvr_patches_copy_from_host(svr_context* c, const float* cube)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->pvr, "pvr_patches_copy_from_host: call pvr_patches_init first");
    return svr_fill_slices(c, cube, nullptr, nullptr);
}

int pvr_patches_copy_to_host(svr_context* c, float* cube) { return svr_debug_get(c, SVR_DBG_SLICES, cube); }

int pvr_set_psf(svr_context* c, const int psf_size[3], const float psf_i2w[16], float quality_factor)
{
    SVR_ENTRY(c);
    return svr_generate_psf_volume(c, psf_size, psf_i2w, quality_factor);
}

int pvr_init_patch_based_recon(svr_context* c, int stack, const float* stack_data, int sx, int sy, int sz, const float stack_w2i[16])
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_init_patch_based_recon")) return r;
    PvrState* p = (PvrState*)c->pvr;
    PVR_REQUIRE(c, stack >= 0 && stack < p->n_stacks && stack_data && stack_w2i && sx > 0 && sy > 0 && sz > 0,
                "pvr_init_patch_based_recon: bad argument");
    const int n = p->per_stack[stack], p0 = p->offset[stack];
    if (n == 0) return 0;
    float* d_stack = nullptr; float* d_w2i = nullptr;
    const size_t nvox = (size_t)sx * sy * sz;
    SVR_CUDA(c, cudaMalloc((void**)&d_stack, nvox * sizeof(float)));
    SVR_CUDA(c, cudaMalloc((void**)&d_w2i, 16 * sizeof(float)));
    SVR_CUDA(c, cudaMemcpyAsync(d_stack, stack_data, nvox * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(d_w2i, stack_w2i, 16 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    dim3 b(32, 8), g(divup_i(c->Nx, 32), divup_i(c->Ny, 8), n);
    pvr_patch_init_kernel<<<g, b, 0, c->stream>>>(p0, n, c->Nx, c->Ny, d_stack, sx, sy, sz, d_w2i, c->mats, (size_t)c->S * 16, c->vg,
                                                   c->mask_f, c->use_spx ? c->spx : nullptr, c->slices);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(c->stream);
    cudaFree(d_stack); cudaFree(d_w2i);
    if (e != cudaSuccess) return svr_fail(c, "pvr_patch_init_kernel", e, __FILE__, __LINE__);
    // the valid-pixel list depends on the patch values
    SVR_CUDA(c, cudaMemcpyAsync(c->slices_restore, c->slices, c->NP * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    return svr_launch_compact_valid(c);
}

int pvr_psf_reconstruction(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_psf_reconstruction")) return r;
    if (svr_launch_gaussian_scatter(c)) return 1;
    if (svr_launch_unpack_acc(c)) return 1;
    return sync_(c);
}

// Split-phase P1 for N ranks (one process per GPU, patches sharded): scatter this rank's patches into the interleaved
// accumulator, let the host all-reduce svr_device_buffer(SVR_BUF_ACCUMULATOR) over the ranks, then unpack it.
int pvr_psf_reconstruction_local(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_psf_reconstruction_local")) return r;
    if (svr_launch_gaussian_scatter(c)) return 1;
    return sync_(c);
}

int pvr_psf_reconstruction_finish(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_psf_reconstruction_finish")) return r;
    if (svr_launch_unpack_acc(c)) return 1;
    return sync_(c);
}

int pvr_simulate_patches(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_simulate_patches")) return r;
    if (svr_launch_pack_volume(c)) return 1;
    if (svr_launch_simulate(c)) return 1;
    return sync_(c);
}

int pvr_superresolution_run(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_superresolution_run")) return r;
    if (svr_launch_superres_scatter(c)) return 1;
    return sync_(c);
}

int pvr_superresolution_regularize(svr_context* c, int adaptive, float alpha, float min_i, float max_i, float delta, float lambda)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_superresolution_regularize")) return r;
    if (svr_launch_regularize(c, adaptive, alpha, min_i, max_i, delta, lambda)) return 1;
    return sync_(c);
}

int pvr_rs_initialize_em_values(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_rs_initialize_em_values")) return r;
    std::vector<float> ones((size_t)std::max(c->S, 1), 1.0f);       // resetScaleAndWeights
    if (int r = svr_update_scale_vector(c, ones.data(), ones.data())) return r;
    return svr_initialize_em_values(c);
}

int pvr_rs_initialize_robust_statistics(svr_context* c, float* sigma)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_rs_initialize_robust_statistics")) return r;
    return svr_initialize_robust_statistics(c, sigma);
}

int pvr_rs_estep_device(svr_context* c, float m, float sigma, float mix, float* patch_potential)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_rs_estep_device")) return r;
    return svr_estep(c, m, sigma, mix, patch_potential);
}

int pvr_rs_get_scales_weights(svr_context* c, float* scales, float* patch_weights)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->pvr && scales && patch_weights, "pvr_rs_get_scales_weights: bad argument");
    for (int i = 0; i < c->S; ++i) { scales[i] = c->h_scales[i]; patch_weights[i] = c->h_slice_weights[i]; }
    return 0;
}

int pvr_rs_set_scales_weights(svr_context* c, const float* scales, const float* patch_weights)
{
    SVR_ENTRY(c);
    PVR_REQUIRE(c, c && c->pvr, "pvr_rs_set_scales_weights: call pvr_patches_init first");
    return svr_update_scale_vector(c, scales, patch_weights);
}

static float Gf(float x, float s) { return 0.00001f * expf(-x * x / (2.0f * s)) / (sqrtf(6.28f * s)); }

int pvr_host_patch_em(int n_stacks, const int* patches_per_stack, const float* potentials_per_patch, const float* scale,
                      float* patch_weight, float step, float state5[5], float* potential_used)
{
    // patchBasedRobustStatistics_gpu<T>::EStep host part, patchBasedRobustStatistics_gpu.cu:277-520 (T = float)
    if (n_stacks < 0 || (n_stacks > 0 && (!patches_per_stack || !potentials_per_patch || !scale || !patch_weight)) || !state5) return 2;
    int numPatches = 0;
    for (int i = 0; i < n_stacks; ++i) numPatches += patches_per_stack[i];
    std::vector<float> pot((size_t)std::max(numPatches, 1), 0.0f);
    int ofs = 0;
    for (int i = 0; i < n_stacks; ++i) {       // patch_potential[j] = ... : no stack offset (:268,272)
        for (int j = 0; j < patches_per_stack[i]; ++j) pot[j] = potentials_per_patch[ofs + j];
        ofs += patches_per_stack[i];
    }
    float sigma_s = state5[0], mix_s = state5[1], mean_s, mean_s2, sigma_s2;
    for (int i = 0; i < numPatches; ++i)
        if ((scale[i] < 0.2) || (scale[i] > 5)) pot[i] = -1;
    double sum = 0, den = 0, sum2 = 0, den2 = 0, maxs = 0, mins = 1;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) {
            sum += pot[i] * patch_weight[i];
            den += patch_weight[i];
            sum2 += pot[i] * (1.0 - patch_weight[i]);
            den2 += (1.0 - patch_weight[i]);
            if (pot[i] > maxs) maxs = pot[i];
            if (pot[i] < mins) mins = pot[i];
        }
    mean_s = (den > 0) ? (float)(sum / den) : (float)mins;
    mean_s2 = (den2 > 0) ? (float)(sum2 / den2) : (float)((maxs + mean_s) / 2.0);
    sum = den = sum2 = den2 = 0;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) {
            sum += (pot[i] - mean_s) * (pot[i] - mean_s) * patch_weight[i];
            den += patch_weight[i];
            sum2 += (pot[i] - mean_s2) * (pot[i] - mean_s2) * (1 - patch_weight[i]);
            den2 += (1 - patch_weight[i]);
        }
    if ((sum > 0) && (den > 0)) {
        sigma_s = (float)(sum / den);
        if (sigma_s < step * step / 6.28) sigma_s = (float)(step * step / 6.28);
    } else sigma_s = 0.025f;
    if ((sum2 > 0) && (den2 > 0)) {
        sigma_s2 = (float)(sum2 / den2);
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    } else {
        sigma_s2 = (mean_s2 - mean_s) * (mean_s2 - mean_s) / 4;
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    }
    for (int i = 0; i < numPatches; ++i) {
        if (pot[i] == -1) { patch_weight[i] = 0; continue; }
        if ((den <= 0) || (mean_s2 <= mean_s)) { patch_weight[i] = 1; continue; }
        const double gs1 = (pot[i] < mean_s2) ? Gf(pot[i] - mean_s, sigma_s) : 0;
        const double gs2 = (pot[i] > mean_s) ? Gf(pot[i] - mean_s2, sigma_s2) : 0;
        const double likelihood = gs1 * mix_s + gs2 * (1 - mix_s);
        if (likelihood > 0) patch_weight[i] = (float)(gs1 * mix_s / likelihood);
        else {
            if (pot[i] <= mean_s) patch_weight[i] = 1;
            if (pot[i] >= mean_s2) patch_weight[i] = 0;
            if ((pot[i] < mean_s2) && (pot[i] > mean_s)) patch_weight[i] = 1;
        }
    }
    sum = 0; int num = 0;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) { sum += patch_weight[i]; num++; }
    mix_s = (num > 0) ? (float)(sum / num) : 0.9f;
    state5[0] = sigma_s; state5[1] = mix_s; state5[2] = mean_s; state5[3] = mean_s2; state5[4] = sigma_s2;
    if (potential_used) memcpy(potential_used, pot.data(), sizeof(float) * numPatches);
    return 0;
}

int pvr_rs_mstep(svr_context* c, int iter, float step, float* sigma, float* mix, float* m)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_rs_mstep")) return r;
    return svr_mstep(c, iter, step, sigma, mix, m);       // same host arithmetic (FLT_MAX / FLT_MIN seeds, 6.28f floor)
}

int pvr_rs_scale(svr_context* c, float* scale_vec)
{
    SVR_ENTRY(c);
    if (int r = pvr_ready(c, "pvr_rs_scale")) return r;
    PVR_REQUIRE(c, scale_vec, "pvr_rs_scale: NULL output");
    if (c->S == 0) return 0;
    if (svr_launch_scale(c)) return 1;
    SVR_CUDA(c, cudaMemcpyAsync(scale_vec, c->slice_tmp, c->S * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (int r = sync_(c)) return r;
    // copyToScales (:724-731): the patches carry the new scale immediately (no one-call lag as in SVR)
    c->h_scales.assign(scale_vec, scale_vec + c->S);
    SVR_CUDA(c, cudaMemcpyAsync(c->scales, scale_vec, c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->scales_mstep, scale_vec, c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return sync_(c);
}

int pvr_debug_get(svr_context* c, int kind, void* out) { return svr_debug_get(c, kind, out); }

}  // extern "C"
