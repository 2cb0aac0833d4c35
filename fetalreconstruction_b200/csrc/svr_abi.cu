// svr_abi.cu -- the C ABI of libsvr_b200.so (see include/svr_abi.h for the reference citations).
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <new>
#include "../../include/svr_abi.h"
#include "svr_context.h"

static thread_local std::string g_create_error;

int svr_fail(svr_context* ctx, const char* what, cudaError_t e, const char* file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return 1;
}
static int fail_msg(svr_context* ctx, const char* msg)
{
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return 2;
}
#define REQUIRE(ctx, cond, msg) do { if (!(cond)) return fail_msg((ctx), (msg)); } while (0)

static cudaEvent_t prof_event(svr_context* c)
{
    cudaEvent_t e = nullptr;
    if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}
ProfScope::ProfScope(svr_context* ctx, int kind) : c(ctx), idx(-1)
{
    if (!c->prof_on) return;
    svr_context::ProfPair p{ prof_event(c), prof_event(c), kind };
    cudaEventRecord(p.a, c->stream);
    c->prof_pending.push_back(p);
    idx = (int)c->prof_pending.size() - 1;
}
ProfScope::~ProfScope()
{
    if (idx >= 0) cudaEventRecord(c->prof_pending[idx].b, c->stream);
}
static void prof_fold(svr_context* c)
{
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->prof_ms[p.kind] += ms; c->prof_n[p.kind] += 1; }
        c->prof_pool.push_back(p.a); c->prof_pool.push_back(p.b);
    }
    c->prof_pending.clear();
}

template <class T>
static int dev_alloc(svr_context* c, T** p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) return 0;
    SVR_CUDA(c, cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}
template <class T>
static void dev_free(T** p) { if (*p) { cudaFree(*p); *p = nullptr; } }

extern "C" {

int svr_abi_version(void) { return SVR_ABI_VERSION; }

const char* svr_last_error(const svr_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int svr_create(svr_context** out, int device)
{
    if (!out) return fail_msg(nullptr, "svr_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail_msg(nullptr, "svr_create: no CUDA device (libsvr_b200 has no CPU fallback)");
    if (device < 0 || device >= n) return fail_msg(nullptr, "svr_create: device index out of range");
    svr_context* c = new (std::nothrow) svr_context();
    if (!c) return fail_msg(nullptr, "svr_create: out of host memory");
    c->device = device;
    SVR_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    SVR_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    SVR_CUDA(nullptr, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    if (const char* e = getenv("SVR_TUNE_SCATTER")) { const int v = atoi(e); if (v >= 0 && v <= 3) c->tune_scatter = v; }
    if (const char* e = getenv("SVR_TUNE_REGULARIZE")) { const int v = atoi(e); if (v == 0 || v == 1) c->tune_regularize = v; }
    if (const char* e = getenv("SVR_TUNE_SIMULATE")) { const int v = atoi(e); if (v >= 0 && v <= 2) c->tune_simulate = v; }
    c->pinned_bytes = 4096;
    SVR_CUDA(nullptr, cudaMallocHost(&c->pinned, c->pinned_bytes));
    *out = c;
    return 0;
}

int svr_destroy(svr_context* c)
{
    SVR_ENTRY(c);
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    dev_free(&c->recon); dev_free(&c->recon_tmp1); dev_free(&c->recon_tmp2); dev_free(&c->volw);
    dev_free(&c->mask_f); dev_free(&c->mask_u8); dev_free(&c->acc2); dev_free(&c->pack2);
    dev_free(&c->slices); dev_free(&c->slices_restore); dev_free(&c->weights); dev_free(&c->simslices);
    dev_free(&c->simweights); dev_free(&c->siminside); dev_free(&c->psf_sums); dev_free(&c->voxel_flag);
    dev_free(&c->valid_idx); dev_free(&c->pair_idx); dev_free(&c->slice_count); dev_free(&c->slice_inside); dev_free(&c->scales);
    dev_free(&c->scales_mstep); dev_free(&c->slice_weights); dev_free(&c->slice_tmp); dev_free(&c->geom);
    dev_free(&c->mats); dev_free(&c->dims); dev_free(&c->partials);
    svr_reg_free(c);
    svr_pvr_free(c);
    svr_window_free(c);
    prof_fold(c);
    for (auto e : c->prof_pool) cudaEventDestroy(e);
    if (c->cub_tmp) cudaFree(c->cub_tmp);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return 0;
}

int svr_set_stream(svr_context* c, void* cuda_stream)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return 0;
}

int svr_synchronize(svr_context* c)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_get_stream(svr_context* c, void** cuda_stream)
{
    REQUIRE(c, c && cuda_stream, "svr_get_stream: NULL argument");
    *cuda_stream = (void*)c->stream;
    return 0;
}

int svr_set_async(svr_context* c, int on)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    if (!on) SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->async = on != 0;
    return 0;
}

int svr_make_current(svr_context* c)
{
    REQUIRE(c, c, "null context");
    SVR_CUDA(c, cudaSetDevice(c->device));
    return 0;
}

int64_t svr_launch_count(const svr_context* c) { return c ? c->launches : 0; }

int svr_set_tuning(svr_context* c, int key, int value)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    switch (key) {
    case SVR_TUNE_SCATTER: REQUIRE(c, value >= 0 && value <= 3, "svr_set_tuning: SVR_TUNE_SCATTER takes 0 .. 3"); c->tune_scatter = value; return 0;
    case SVR_TUNE_SIMULATE: REQUIRE(c, value >= 0 && value <= 2, "svr_set_tuning: SVR_TUNE_SIMULATE takes 0, 1 or 2"); c->tune_simulate = value; return 0;
    case SVR_TUNE_REGULARIZE: REQUIRE(c, value == 0 || value == 1, "svr_set_tuning: SVR_TUNE_REGULARIZE takes 0 or 1"); c->tune_regularize = value; return 0;
    default: return fail_msg(c, "svr_set_tuning: unknown key");
    }
}

int svr_profile_enable(svr_context* c, int on)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    if (!on) prof_fold(c);
    c->prof_on = on != 0;
    return 0;
}
int svr_profile_read(svr_context* c, int kind, double* total_ms, int64_t* launches)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && kind >= 0 && kind < SVR_K_COUNT, "svr_profile_read: bad argument");
    prof_fold(c);
    if (total_ms) *total_ms = c->prof_ms[kind];
    if (launches) *launches = c->prof_n[kind];
    return 0;
}
int svr_profile_reset(svr_context* c)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    prof_fold(c);
    for (int i = 0; i < 16; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int upload(svr_context* c, void* dst, const void* src, size_t bytes)
{
    SVR_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    SVR_SYNC(c);                                     // the caller may free its buffer on return (reference semantics)
    return 0;
}
static int download(svr_context* c, void* dst, const void* src, size_t bytes)
{
    SVR_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_init_reconstruction_volume(svr_context* c, int sx, int sy, int sz, float dx, float dy, float dz, const float* data)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, sx > 0 && sy > 0 && sz > 0, "svr_init_reconstruction_volume: bad size");
    REQUIRE(c, (double)sx * sy * sz < 2147483647.0, "svr_init_reconstruction_volume: volume too large for 32-bit voxel indices");
    SVR_CUDA(c, cudaSetDevice(c->device));
    c->vx = sx; c->vy = sy; c->vz = sz; c->vdx = dx; c->vdy = dy; c->vdz = dz;
    c->V = (size_t)sx * sy * sz;
    c->vg.vx = sx; c->vg.vy = sy; c->vg.vz = sz;
    if (dev_alloc(c, &c->recon, c->V) || dev_alloc(c, &c->recon_tmp1, c->V) || dev_alloc(c, &c->recon_tmp2, c->V) ||
        dev_alloc(c, &c->volw, c->V) || dev_alloc(c, &c->mask_f, c->V) || dev_alloc(c, &c->mask_u8, c->V) ||
        dev_alloc(c, &c->acc2, c->V + 2) || dev_alloc(c, &c->pack2, c->V))
        return 1;
    SVR_CUDA(c, cudaMemsetAsync(c->recon, 0, c->V * sizeof(float), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->volw, 0, c->V * sizeof(float), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->acc2, 0, (c->V + 2) * sizeof(float2), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->mask_f, 0, c->V * sizeof(float), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->mask_u8, 0, c->V, c->stream));
    c->have_mask = false;
    if (int r = svr_window_build_maps(c)) return r;      // tensor-map menus over acc2 / pack2 (svr_window.cu)
    if (data) return upload(c, c->recon, data, c->V * sizeof(float));
    SVR_SYNC(c);
    return 0;
}

__global__ void mask_to_u8_kernel(size_t V, const float* __restrict__ m, unsigned char* __restrict__ o)
{
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < V; v += (size_t)gridDim.x * blockDim.x) o[v] = m[v] != 0.f;
}

int svr_set_mask(svr_context* c, int sx, int sy, int sz, const float* mask)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, c->V > 0, "svr_set_mask: call svr_init_reconstruction_volume first");
    REQUIRE(c, sx == c->vx && sy == c->vy && sz == c->vz, "svr_set_mask: mask grid differs from the volume grid");
    REQUIRE(c, mask, "svr_set_mask: mask is NULL");
    if (upload(c, c->mask_f, mask, c->V * sizeof(float))) return 1;
    mask_to_u8_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(c->V, c->mask_f, c->mask_u8);
    SVR_KERNEL_CHECK(c);
    c->have_mask = true;
    return 0;
}

int svr_init_storage_volumes(svr_context* c, int Nx, int Ny, int S)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, Nx > 0 && Ny > 0 && S >= 0, "svr_init_storage_volumes: bad size");
    REQUIRE(c, (double)Nx * Ny * (double)std::max(S, 1) < 2147483647.0, "svr_init_storage_volumes: slice cube too large for 32-bit pixel indices");
    // the registration and PVR patch kernels put the slice / patch index on gridDim.y / gridDim.z (limit 65535)
    REQUIRE(c, S <= 65535, "svr_init_storage_volumes: more than 65535 slices / patches per context (shard them over more ranks)");
    SVR_CUDA(c, cudaSetDevice(c->device));
    c->Nx = Nx; c->Ny = Ny; c->S = S; c->NP = (size_t)Nx * Ny * S;
    const size_t NP = c->NP, Sn = (size_t)std::max(S, 1);
    if (dev_alloc(c, &c->slices, NP) || dev_alloc(c, &c->slices_restore, NP) || dev_alloc(c, &c->weights, NP) ||
        dev_alloc(c, &c->simslices, NP) || dev_alloc(c, &c->simweights, NP) || dev_alloc(c, &c->siminside, NP) ||
        dev_alloc(c, &c->psf_sums, NP) || dev_alloc(c, &c->voxel_flag, NP) || dev_alloc(c, &c->valid_idx, NP + 4) || dev_alloc(c, &c->pair_idx, (size_t)std::max(S, 1) * Ny * ((Nx + 1) / 2)) ||
        dev_alloc(c, &c->slice_count, Sn) || dev_alloc(c, &c->slice_inside, Sn) || dev_alloc(c, &c->scales, Sn) ||
        dev_alloc(c, &c->scales_mstep, Sn) || dev_alloc(c, &c->slice_weights, Sn) || dev_alloc(c, &c->slice_tmp, 4 * Sn) ||
        dev_alloc(c, &c->geom, Sn) || dev_alloc(c, &c->mats, 4 * 16 * Sn) || dev_alloc(c, &c->dims, 3 * Sn))
        return 1;
    c->n_partials = (int)std::max<size_t>(2 * Sn, (size_t)c->sm_count * 4 * 5 + 16);
    if (dev_alloc(c, &c->partials, (size_t)c->n_partials)) return 1;
    if (NP) {
        // cudaMemset list of initStorageVolumesOnX, cuda2.cu:1555-1568
        SVR_CUDA(c, cudaMemsetAsync(c->slices, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->weights, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->simslices, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->simweights, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->siminside, 0, NP, c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->psf_sums, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->voxel_flag, 0, NP, c->stream));
    }
    SVR_CUDA(c, cudaMemsetAsync(c->slice_count, 0, Sn * sizeof(int), c->stream));
    SVR_CUDA(c, cudaMemsetAsync(c->slice_inside, 0, Sn * sizeof(int), c->stream));
    c->h_scales.assign(S, 1.0f);
    c->h_slice_weights.assign(S, 1.0f);
    c->n_valid = 0;
    c->n_pairs = 0;
    c->have_mats = c->have_dims = false;
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_fill_slices(svr_context* c, const float* cube, const int* sizesX, const int* sizesY)
{
    SVR_ENTRY(c);
    (void)sizesX; (void)sizesY;
    REQUIRE(c, c, "null context");
    REQUIRE(c, c->slices || c->NP == 0, "svr_fill_slices: call svr_init_storage_volumes first");
    if (c->NP == 0) return 0;
    REQUIRE(c, cube, "svr_fill_slices: cube is NULL");
    if (upload(c, c->slices, cube, c->NP * sizeof(float))) return 1;
    // "needed for stack-wise restore slice intensity" (cuda2.cu:1654-1655)
    SVR_CUDA(c, cudaMemcpyAsync(c->slices_restore, c->slices, c->NP * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    if (int r = svr_launch_compact_valid(c)) return r;
    return svr_window_build_tiles(c);
}

int svr_set_slice_dims(svr_context* c, const float* dims_xyz, float quality_factor)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, c->dims || c->S == 0, "svr_set_slice_dims: call svr_init_storage_volumes first");
    c->quality_factor = quality_factor;     // only stored: the PSF support is fixed at 16 (USE_INFINITE_PSF_SUPPORT)
    if (c->S == 0) return 0;
    REQUIRE(c, dims_xyz, "svr_set_slice_dims: dims is NULL");
    if (upload(c, c->dims, dims_xyz, (size_t)c->S * 3 * sizeof(float))) return 1;
    c->have_dims = true;
    if (svr_launch_build_geom(c)) return 1;
    return 0;
}

int svr_set_slice_matrices(svr_context* c, const float* T, const float* Tinv, const float* I2W, const float* W2I,
                           const float recon_i2w[16], const float recon_w2i[16])
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, c->mats || c->S == 0, "svr_set_slice_matrices: call svr_init_storage_volumes first");
    REQUIRE(c, recon_i2w && recon_w2i, "svr_set_slice_matrices: volume matrices are NULL");
    memcpy(c->recon_i2w, recon_i2w, 16 * sizeof(float));
    memcpy(c->recon_w2i, recon_w2i, 16 * sizeof(float));
    memcpy(c->vg.rw2i, recon_w2i, 12 * sizeof(float));
    if (c->S == 0) return 0;
    REQUIRE(c, T && Tinv && I2W && W2I, "svr_set_slice_matrices: a matrix array is NULL");
    const size_t n = (size_t)c->S * 16;
    SVR_CUDA(c, cudaMemcpyAsync(c->mats, T, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->mats + n, Tinv, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->mats + 2 * n, I2W, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->mats + 3 * n, W2I, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->have_mats = true;
    if (svr_launch_build_geom(c)) return 1;
    return 0;
}

int svr_generate_psf_volume(svr_context* c, const int psf_size[3], const float psf_i2w[16], float quality_factor)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, psf_size && psf_i2w, "svr_generate_psf_volume: NULL argument");
    c->quality_factor = quality_factor;
    // d_PSFI2W * ((d_PSFsize - 1) * 0.5f), the constant second term of cuda2.cu:172, in float like the kernel.
    const float cx = (psf_size[0] - 1) * 0.5f, cy = (psf_size[1] - 1) * 0.5f, cz = (psf_size[2] - 1) * 0.5f;
    for (int r = 0; r < 3; ++r)
        c->vg.psf_c[r] = psf_i2w[4 * r + 0] * cx + psf_i2w[4 * r + 1] * cy + psf_i2w[4 * r + 2] * cz + psf_i2w[4 * r + 3];
    return 0;
}

int svr_update_scale_vector(svr_context* c, const float* scales, const float* slice_weights)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    REQUIRE(c, c->scales || c->S == 0, "svr_update_scale_vector: call svr_init_storage_volumes first");
    if (c->S == 0) return 0;
    REQUIRE(c, scales && slice_weights, "svr_update_scale_vector: NULL argument");
    c->h_scales.assign(scales, scales + c->S);
    c->h_slice_weights.assign(slice_weights, slice_weights + c->S);
    SVR_CUDA(c, cudaMemcpyAsync(c->scales, scales, c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->scales_mstep, scales, c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaMemcpyAsync(c->slice_weights, slice_weights, c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_SYNC(c);
    return 0;
}

int svr_update_slice_weights(svr_context* c, const float* slice_weights)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    if (c->S == 0) return 0;
    REQUIRE(c, slice_weights, "svr_update_slice_weights: NULL argument");
    c->h_slice_weights.assign(slice_weights, slice_weights + c->S);
    return upload(c, c->slice_weights, slice_weights, c->S * sizeof(float));
}

int svr_update_reconstructed(svr_context* c, const float* data)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && c->recon && data, "svr_update_reconstructed: volume not initialised or NULL data");
    return upload(c, c->recon, data, c->V * sizeof(float));
}

// ---------------------------------------------------------------------------------------------
static int ready(svr_context* c, const char* who)
{
    if (!c) return fail_msg(nullptr, "null context");
    if (!c->recon || !c->have_mask) { c->err = std::string(who) + ": volume / mask not initialised"; return 2; }
    if (c->S > 0 && (!c->have_mats || !c->have_dims)) { c->err = std::string(who) + ": slice matrices / dims not set"; return 2; }
    return 0;
}

int svr_initialize_em_values(svr_context* c)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    if (c->NP == 0) return 0;
    if (svr_launch_init_em(c)) return 1;
    SVR_SYNC(c);
    return 0;
}

int svr_gaussian_reconstruction_local(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_gaussian_reconstruction")) return r;
    const size_t NP = c->NP;
    if (NP) {
        // memset list of GaussianReconstructionOnX1, cuda2.cu:2402-2411 (psf_sums is NOT in it)
        SVR_CUDA(c, cudaMemsetAsync(c->weights, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->simweights, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->simslices, 0, NP * sizeof(float), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->siminside, 0, NP, c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->voxel_flag, 0, NP, c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->slice_count, 0, c->S * sizeof(int), c->stream));
        SVR_CUDA(c, cudaMemsetAsync(c->slice_inside, 0, c->S * sizeof(int), c->stream));
    }
    SVR_CUDA(c, cudaMemsetAsync(c->acc2, 0, c->V * sizeof(float2), c->stream));
    if (svr_launch_gaussian_scatter(c)) return 1;
    // the caller all-reduces the accumulator next, possibly on another stream: like every entry point, return when done
    SVR_SYNC(c);
    return 0;
}

int svr_gaussian_reconstruction_finish(svr_context* c, int* voxel_num)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_gaussian_reconstruction")) return r;
    if (svr_launch_equalize(c)) return 1;
    if (voxel_num && c->S) return download(c, voxel_num, c->slice_count, c->S * sizeof(int));
    SVR_SYNC(c);
    return 0;
}

int svr_gaussian_reconstruction(svr_context* c, int* voxel_num)
{
    SVR_ENTRY(c);
    if (int r = svr_gaussian_reconstruction_local(c)) return r;
    return svr_gaussian_reconstruction_finish(c, voxel_num);
}

int svr_simulate_slices(svr_context* c, unsigned char* slice_inside)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_simulate_slices")) return r;
    if (svr_launch_pack_volume(c)) return 1;
    if (svr_launch_simulate(c)) return 1;
    if (slice_inside && c->S) {
        std::vector<int> tmp(c->S);
        if (download(c, tmp.data(), c->slice_inside, c->S * sizeof(int))) return 1;
        for (int i = 0; i < c->S; ++i) slice_inside[i] = tmp[i] != 0;
        return 0;
    }
    SVR_SYNC(c);
    return 0;
}

int svr_initialize_robust_statistics_local(svr_context* c, double sums2[2])
{
    SVR_ENTRY(c);
    REQUIRE(c, c && sums2, "svr_initialize_robust_statistics: NULL argument");
    sums2[0] = sums2[1] = 0;
    if (c->NP == 0) return 0;
    return svr_launch_robust_init(c, sums2);
}

int svr_initialize_robust_statistics(svr_context* c, float* sigma)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && sigma, "svr_initialize_robust_statistics: NULL argument");
    double s2[2];
    if (int r = svr_initialize_robust_statistics_local(c, s2)) return r;
    *sigma = (float)s2[0] / (float)s2[1];               // _sigma = sa / sb, cuda2.cu:2305
    return 0;
}

int svr_estep(svr_context* c, float m, float sigma, float mix, float* slice_potential)
{
    SVR_ENTRY(c);
    REQUIRE(c, c, "null context");
    if (c->NP == 0) return 0;
    if (svr_launch_estep(c, m, sigma, mix)) return 1;
    if (slice_potential) return download(c, slice_potential, c->slice_tmp, c->S * sizeof(float));
    SVR_SYNC(c);
    return 0;
}

int svr_mstep_local(svr_context* c, double sums5[5])
{
    SVR_ENTRY(c);
    REQUIRE(c, c && sums5, "svr_mstep: NULL argument");
    for (int i = 0; i < 5; ++i) sums5[i] = 0;
    if (c->NP == 0) return 0;
    return svr_launch_mstep(c, sums5);
}

int svr_mstep_finish(const double sums5[5], int iter, float step, float* sigma_, float* mix_, float* m_)
{
    // Reconstruction::MStep host part, cuda2.cu:3016-3071
    if (!sums5 || !sigma_ || !mix_ || !m_) return 2;
    const float sigma = (float)sums5[0], mix = (float)sums5[1], num = (float)sums5[2];
    float min_ = FLT_MAX, max_ = FLT_MIN;
    min_ = std::min(min_, (float)sums5[3]);
    max_ = std::max(max_, (float)sums5[4]);
    if (mix > 0) *sigma_ = sigma / mix;
    if (*sigma_ < step * step / 6.28f) *sigma_ = step * step / 6.28f;
    if (iter > 1) *mix_ = mix / num;
    *m_ = 1.0f / (max_ - min_);
    return 0;
}

int svr_mstep(svr_context* c, int iter, float step, float* sigma, float* mix, float* m)
{
    SVR_ENTRY(c);
    double s5[5];
    if (int r = svr_mstep_local(c, s5)) return r;
    return svr_mstep_finish(s5, iter, step, sigma, mix, m);
}

int svr_calculate_scale_vector(svr_context* c, float* scale_vec)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && scale_vec, "svr_calculate_scale_vector: NULL argument");
    if (c->S == 0) return 0;
    if (svr_launch_scale(c)) return 1;
    if (download(c, scale_vec, c->slice_tmp, c->S * sizeof(float))) return 1;
    // cuda2.cu:3238 uploads the OLD h_scales to the device, then cuda2.cu:3195 sets h_scales = scale_vec:
    // the kernels lag one call behind, the M-step (which fills its buffer from h_scales) does not.
    SVR_CUDA(c, cudaMemcpyAsync(c->scales, c->h_scales.data(), c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    c->h_scales.assign(scale_vec, scale_vec + c->S);
    SVR_CUDA(c, cudaMemcpyAsync(c->scales_mstep, c->h_scales.data(), c->S * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_superresolution_local(svr_context* c, const float* slice_weight)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_superresolution")) return r;
    if (slice_weight && c->S) if (int r = svr_update_slice_weights(c, slice_weight)) return r;
    SVR_CUDA(c, cudaMemsetAsync(c->acc2, 0, c->V * sizeof(float2), c->stream));
    if (svr_launch_superres_scatter(c)) return 1;
    SVR_SYNC(c);                                         // see svr_gaussian_reconstruction_local
    return 0;
}

int svr_superresolution_finish(svr_context* c, int adaptive, float alpha, float min_i, float max_i, float delta, float lambda)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_superresolution")) return r;
    if (svr_launch_regularize(c, adaptive, alpha, min_i, max_i, delta, lambda)) return 1;
    SVR_SYNC(c);
    return 0;
}

int svr_superresolution(svr_context* c, int iter, const float* slice_weight, int adaptive, float alpha, float min_i,
                        float max_i, float delta, float lambda)
{
    SVR_ENTRY(c);
    (void)iter;
    if (int r = svr_superresolution_local(c, slice_weight)) return r;
    return svr_superresolution_finish(c, adaptive, alpha, min_i, max_i, delta, lambda);
}

int svr_mask_volume(svr_context* c)
{
    SVR_ENTRY(c);
    if (int r = ready(c, "svr_mask_volume")) return r;
    if (svr_launch_mask_volume(c)) return 1;
    SVR_SYNC(c);
    return 0;
}

int svr_scale_volume_local(svr_context* c, double sums2[2])
{
    SVR_ENTRY(c);
    REQUIRE(c, c && sums2, "svr_scale_volume: NULL argument");
    sums2[0] = sums2[1] = 0;
    if (c->NP == 0) return 0;
    return svr_launch_scale_volume_sums(c, sums2);
}

int svr_scale_volume_apply(svr_context* c, float scale)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && c->recon, "svr_scale_volume: volume not initialised");
    if (svr_launch_scale_volume_apply(c, scale)) return 1;
    SVR_SYNC(c);
    return 0;
}

int svr_scale_volume(svr_context* c, float* scale_out)
{
    SVR_ENTRY(c);
    double s2[2];
    if (int r = svr_scale_volume_local(c, s2)) return r;
    const float scale = (float)(s2[0] / s2[1]);          // cuda2.cu:3459
    if (scale_out) *scale_out = scale;
    return svr_scale_volume_apply(c, scale);
}

int svr_restore_slice_intensities(svr_context* c, const float* stack_factors, int n_stacks, const int* stack_index)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && stack_factors && stack_index && n_stacks > 0, "svr_restore_slice_intensities: bad argument");
    if (c->NP == 0) return 0;
    float* d_f = nullptr; int* d_i = nullptr;
    SVR_CUDA(c, cudaMalloc(&d_f, n_stacks * sizeof(float)));
    SVR_CUDA(c, cudaMalloc(&d_i, c->S * sizeof(int)));
    int rc = upload(c, d_f, stack_factors, n_stacks * sizeof(float)) || upload(c, d_i, stack_index, c->S * sizeof(int)) ||
             svr_launch_restore(c, d_f, d_i);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_f); cudaFree(d_i);
    return rc;
}

int svr_sync_cpu(svr_context* c, float* reconstructed)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && c->recon && reconstructed, "svr_sync_cpu: volume not initialised or NULL output");
    return download(c, reconstructed, c->recon, c->V * sizeof(float));
}

int svr_get_vol_weights(svr_context* c, float* weights)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && c->volw && weights, "svr_get_vol_weights: volume not initialised or NULL output");
    return download(c, weights, c->volw, c->V * sizeof(float));
}

int svr_debug_get(svr_context* c, int kind, void* out)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && out, "svr_debug_get: NULL argument");
    switch (kind) {
    case SVR_DBG_WEIGHTS: return download(c, out, c->weights, c->NP * sizeof(float));
    case SVR_DBG_SIMSLICES: return download(c, out, c->simslices, c->NP * sizeof(float));
    case SVR_DBG_SIMWEIGHTS: return download(c, out, c->simweights, c->NP * sizeof(float));
    case SVR_DBG_SIMINSIDE: return download(c, out, c->siminside, c->NP);
    case SVR_DBG_PSF_SUMS: return download(c, out, c->psf_sums, c->NP * sizeof(float));
    case SVR_DBG_SLICES: return download(c, out, c->slices, c->NP * sizeof(float));
    case SVR_DBG_SLICES_RESTORED: return download(c, out, c->slices_restore, c->NP * sizeof(float));
    case SVR_DBG_SCALES_DEVICE: return download(c, out, c->scales, c->S * sizeof(float));
    case SVR_DBG_MASK: return download(c, out, c->mask_f, c->V * sizeof(float));
    case SVR_DBG_CONFIDENCE_MAP:
    case SVR_DBG_ADDON:
        // recon_tmp1 is free outside svr_superresolution_finish: use it to de-interleave the accumulator
        if (svr_launch_deinterleave(c, c->acc2, c->recon_tmp1, kind == SVR_DBG_CONFIDENCE_MAP)) return 1;
        return download(c, out, c->recon_tmp1, c->V * sizeof(float));
    case SVR_DBG_VOXEL_COUNT: {
        int* tmp = nullptr;
        SVR_CUDA(c, cudaMalloc(&tmp, c->NP * sizeof(int)));
        int rc = svr_launch_flags_to_int(c, c->voxel_flag, tmp, c->NP) || download(c, out, tmp, c->NP * sizeof(int));
        cudaFree(tmp);
        return rc;
    }
    case SVR_DBG_WW_STATS:
        if (!c->ww_counter) { memset(out, 0, 32 * sizeof(unsigned int)); return 0; }
        return download(c, out, c->ww_counter + 8, 32 * sizeof(unsigned int));
    default: return fail_msg(c, "svr_debug_get: unknown kind");
    }
}

int svr_device_buffer(svr_context* c, int kind, void** dev_ptr, size_t* nbytes)
{
    SVR_ENTRY(c);
    REQUIRE(c, c && dev_ptr && nbytes, "svr_device_buffer: NULL argument");
    switch (kind) {
    case SVR_BUF_ACCUMULATOR: *dev_ptr = c->acc2; *nbytes = c->V * sizeof(float2); return 0;
    case SVR_BUF_RECON: *dev_ptr = c->recon; *nbytes = c->V * sizeof(float); return 0;
    default: return fail_msg(c, "svr_device_buffer: unknown kind");
    }
}

// ---------------------------------------------------------------------------------------------
// Pure host helpers.
static double Gd(double x, double s, double step) { return step * exp(-x * x / (2 * s)) / (sqrt(6.28 * s)); }

int svr_host_slice_em(int S, float* pot, const float* scale, float* sw, const int* force_excluded, int n_force,
                      const int* small_slices, int n_small, double step, float state5[5])
{
    // irtkReconstruction::EStepGPU, irtkReconstructionGPU.cc:3203-3420
    if (S < 0 || (S > 0 && (!pot || !scale || !sw)) || !state5) return 2;
    float sigma_s = state5[0], mix_s = state5[1], mean_s = state5[2], mean_s2 = state5[3], sigma_s2 = state5[4];
    for (int i = 0; i < n_force; i++) if (force_excluded[i] >= 0 && force_excluded[i] < S) pot[force_excluded[i]] = -1;
    for (int i = 0; i < n_small; i++) if (small_slices[i] >= 0 && small_slices[i] < S) pot[small_slices[i]] = -1;
    for (int i = 0; i < S; i++) if ((scale[i] < 0.2) || (scale[i] > 5)) pot[i] = -1;

    double sum = 0, den = 0, sum2 = 0, den2 = 0, maxs = 0, mins = 1;
    for (int i = 0; i < S; i++)
        if (pot[i] >= 0) {
            sum += pot[i] * sw[i];
            den += sw[i];
            sum2 += pot[i] * (1.0 - sw[i]);
            den2 += (1.0 - sw[i]);
            if (pot[i] > maxs) maxs = pot[i];
            if (pot[i] < mins) mins = pot[i];
        }
    mean_s = (den > 0) ? (float)(sum / den) : (float)mins;
    mean_s2 = (den2 > 0) ? (float)(sum2 / den2) : (float)((maxs + mean_s) / 2.0);

    sum = 0; den = 0; sum2 = 0; den2 = 0;
    for (int i = 0; i < S; i++)
        if (pot[i] >= 0) {
            sum += (pot[i] - mean_s) * (pot[i] - mean_s) * sw[i];
            den += sw[i];
            sum2 += (pot[i] - mean_s2) * (pot[i] - mean_s2) * (1 - sw[i]);
            den2 += (1 - sw[i]);
        }
    const double floor_ = step * step / 6.28;
    if ((sum > 0) && (den > 0)) {
        sigma_s = (float)(sum / den);
        if (sigma_s < floor_) sigma_s = (float)floor_;
    } else
        sigma_s = 0.025f;
    if ((sum2 > 0) && (den2 > 0)) {
        sigma_s2 = (float)(sum2 / den2);
        if (sigma_s2 < floor_) sigma_s2 = (float)floor_;
    } else {
        sigma_s2 = (mean_s2 - mean_s) * (mean_s2 - mean_s) / 4;
        if (sigma_s2 < floor_) sigma_s2 = (float)floor_;
    }
    for (int i = 0; i < S; i++) {
        if (pot[i] == -1) { sw[i] = 0; continue; }
        if ((den <= 0) || (mean_s2 <= mean_s)) { sw[i] = 1; continue; }
        const double gs1 = (pot[i] < mean_s2) ? Gd(pot[i] - mean_s, sigma_s, step) : 0;
        const double gs2 = (pot[i] > mean_s) ? Gd(pot[i] - mean_s2, sigma_s2, step) : 0;
        const double likelihood = gs1 * mix_s + gs2 * (1 - mix_s);
        if (likelihood > 0)
            sw[i] = (float)(gs1 * mix_s / likelihood);
        else {
            if (pot[i] <= mean_s) sw[i] = 1;
            if (pot[i] >= mean_s2) sw[i] = 0;
            if ((pot[i] < mean_s2) && (pot[i] > mean_s)) sw[i] = 1;
        }
    }
    sum = 0; int num = 0;
    for (int i = 0; i < S; i++)
        if (pot[i] >= 0) { sum += sw[i]; num++; }
    mix_s = (num > 0) ? (float)(sum / num) : 0.9f;
    state5[0] = sigma_s; state5[1] = mix_s; state5[2] = mean_s; state5[3] = mean_s2; state5[4] = sigma_s2;
    return 0;
}

int svr_host_small_slices(int S, const int* voxel_num, int* out_small, int* n_out)
{
    // irtkReconstructionGPU.cc:2714-2726 on per-slice counts (deviation D4)
    if (S < 0 || !n_out || (S > 0 && (!voxel_num || !out_small))) return 2;
    *n_out = 0;
    if (S == 0) return 0;
    std::vector<int> tmp(voxel_num, voxel_num + S);
    std::sort(tmp.begin(), tmp.end());
    size_t mid = (size_t)llround(tmp.size() * 0.5);
    if (mid >= tmp.size()) mid = tmp.size() - 1;
    const int median = tmp[mid];
    for (int i = 0; i < S; i++)
        if (voxel_num[i] < 0.1 * median) out_small[(*n_out)++] = i;
    return 0;
}

int svr_host_partition_strided(int n_stacks, const int* slices_per_stack, int nranks, int rank, int* out_indices, int* out_count)
{
    if (n_stacks < 0 || nranks <= 0 || rank < 0 || rank >= nranks || !out_count || (n_stacks > 0 && (!slices_per_stack || !out_indices)))
        return 2;
    int n = 0, base = 0;
    for (int st = 0; st < n_stacks; ++st) {
        if (slices_per_stack[st] < 0) return 2;
        for (int j = rank; j < slices_per_stack[st]; j += nranks) out_indices[n++] = base + j;
        base += slices_per_stack[st];
    }
    *out_count = n;
    return 0;
}

int svr_host_partition(int n_stacks, const int* slices_per_stack, int nranks, int rank, int* out_begin, int* out_end)
{
    if (n_stacks < 0 || nranks <= 0 || rank < 0 || rank >= nranks || !out_begin || !out_end || (n_stacks > 0 && !slices_per_stack))
        return 2;
    long long total = 0;
    for (int i = 0; i < n_stacks; ++i) { if (slices_per_stack[i] < 0) return 2; total += slices_per_stack[i]; }
    if (n_stacks >= nranks) {
        // whole stacks, contiguous, greedy by slice count: cut after the stack that brings the running
        // total closest to (r+1)*total/nranks while leaving at least one stack per remaining rank
        std::vector<long long> cut(nranks + 1, 0);
        int st = 0; long long run = 0;
        for (int r = 0; r < nranks; ++r) {
            cut[r] = run;
            const double target = (double)total * (r + 1) / nranks;
            const int must_leave = nranks - r - 1;
            bool took = false;
            while (st < n_stacks - must_leave) {
                const long long nxt = run + slices_per_stack[st];
                if (took && std::fabs((double)nxt - target) > std::fabs((double)run - target)) break;
                run = nxt; ++st; took = true;
            }
            if (r == nranks - 1) { while (st < n_stacks) run += slices_per_stack[st++]; }
        }
        cut[nranks] = run;
        *out_begin = (int)cut[rank]; *out_end = (int)cut[rank + 1];
    } else {
        *out_begin = (int)(total * rank / nranks);
        *out_end = (int)(total * (rank + 1) / nranks);
    }
    return 0;
}

}  // extern "C"
