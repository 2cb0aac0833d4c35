// svr_context.h -- the state behind the opaque svr_context handle (one GPU, one rank).
// Replaces the member set of class Reconstruction (include/reconstruction_cuda2.cuh:97-341).
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "svr_common.cuh"

struct svr_context {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;
    bool async = false;            // svr_set_async: entry points that return no host data do not wait for the stream

    // ---- volume side (V voxels) ----
    int vx = 0, vy = 0, vz = 0;
    float vdx = 1, vdy = 1, vdz = 1;
    size_t V = 0;
    float* recon = nullptr;        // reconstructed volume (dev_reconstructed_)
    float* recon_tmp1 = nullptr;   // post-gradient-step volume, frozen for the regulariser (deviation D3)
    float* recon_tmp2 = nullptr;   // regulariser output (swapped with recon)
    float* volw = nullptr;         // dev_reconstructed_volWeigths
    float* mask_f = nullptr;       // dev_mask_ as uploaded
    unsigned char* mask_u8 = nullptr;
    float2* acc2 = nullptr;        // interleaved scatter accumulator {numerator, denominator}
    float2* pack2 = nullptr;       // {recon*m, m}, m in {0,1}: one 64-bit load per forward tap
    bool have_mask = false;

    // ---- slice side (S slices of Nx x Ny) ----
    int Nx = 0, Ny = 0, S = 0;
    size_t NP = 0;                 // S*Nx*Ny
    float* slices = nullptr;       // dev_v_slices
    float* slices_restore = nullptr; // v_slices: the copy RestoreSliceIntensities rescales (cuda2.cu:1655)
    float* weights = nullptr;      // dev_v_weights
    float* simslices = nullptr;    // dev_v_simulated_slices
    float* simweights = nullptr;   // dev_v_simulated_weights
    unsigned char* siminside = nullptr; // dev_v_simulated_inside
    float* psf_sums = nullptr;     // dev_v_PSF_sums_
    unsigned char* voxel_flag = nullptr; // dev_sliceVoxel_count_ (0/1)
    uint32_t* valid_idx = nullptr; // compacted indices of pixels != -1 (static after FillSlices)
    uint32_t n_valid = 0;
    uint32_t* pair_idx = nullptr;  // even-x pixels (x, x+1) of which at least one is != -1: the units of the paired scatter
    uint32_t n_pairs = 0;
    // warp-window kernels (svr_window.cu): tiles of 8 x 4 pixels holding at least one pixel != -1, the dynamic tile queue,
    // and the tensor-map menus (one CUtensorMap per window shape) over acc2 and pack2
    uint32_t* tile_idx = nullptr;
    size_t tile_cap = 0;
    uint32_t n_tiles = 0;
    int tilesX = 0, tilesPerSlice = 0;
    unsigned int* ww_counter = nullptr;
    void* maps_acc = nullptr;
    void* maps_pack = nullptr;
    int tune_scatter = 0;          // 0: paired scatter (round 1), 1: warp windows + SIMT flush, 2: warp windows + TMA reduce flush,
                                   // 3: paired for slices aligned with the volume axes, warp windows (TMA flush) for the others
    int tune_regularize = 1;       // 0: reg_prep_kernel + reg_kernel (round 1), 1: fused K4 + K5 with a shared-memory halo tile
    int tune_simulate = 0;         // 0: per-tap loads (+ staged rows), 1: TMA-staged windows for every tile, 2: windows for through-plane slices only
    int* slice_count = nullptr;    // [S] per-slice voxel_num (deviation D4)
    int* slice_inside = nullptr;   // [S] OR of siminside since the last Gaussian reconstruction
    float* scales = nullptr;       // dev_d_scales: what the kernels see
    float* scales_mstep = nullptr; // device copy of h_scales: what the M-step sees (cuda2.cu:3093)
    float* slice_weights = nullptr;// dev_d_slice_weights
    float* slice_tmp = nullptr;    // [4*S] per-slice reduction outputs
    SliceGeom* geom = nullptr;     // [S]
    float* mats = nullptr;         // [4][S][16] staging of the uploaded matrices: T, Tinv, I2W, W2I
    float* dims = nullptr;         // [S][3]
    bool have_mats = false, have_dims = false;
    std::vector<float> h_scales;   // Reconstruction::h_scales
    std::vector<float> h_slice_weights;

    double* partials = nullptr;    // device scratch for grid reductions
    int n_partials = 0;
    void* pinned = nullptr;        // small pinned staging buffer for scalar read-backs
    size_t pinned_bytes = 0;
    void* cub_tmp = nullptr;
    size_t cub_tmp_bytes = 0;

    // per-kernel device timing (svr_profile_*): event pairs recorded on the launching stream
    struct ProfPair { cudaEvent_t a, b; int kind; };
    bool prof_on = false;
    std::vector<ProfPair> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {0};
    long long prof_n[16] = {0};

    VolGeom vg{};
    float recon_i2w[16]{}, recon_w2i[16]{};
    float quality_factor = 1.0f;
    int sm_count = 148;
    int flavor = 0;                // 0 = SVR constants, 1 = PVR constants (svr_common.cuh traits)
    char* spx = nullptr;           // PVR superpixel masks, char[S][64*64] ('1' = inside)
    int use_spx = 0;
    void* reg = nullptr;           // RegState (svr_reg.cu), owned
    void* pvr = nullptr;           // PvrState (pvr.cu), owned
};
void svr_reg_free(svr_context* c);
void svr_pvr_free(svr_context* c);

// error helper used by every translation unit
int svr_fail(svr_context* ctx, const char* what, cudaError_t e, const char* file, int line);
#define SVR_CUDA(ctx, call)                                                        \
    do {                                                                           \
        cudaError_t _e = (call);                                                   \
        if (_e != cudaSuccess) return svr_fail((ctx), #call, _e, __FILE__, __LINE__); \
    } while (0)
// end of an entry point that returns no host data: wait for the stream unless the context is asynchronous
#define SVR_SYNC(ctx)                                                              \
    do {                                                                           \
        if (!(ctx)->async) SVR_CUDA((ctx), cudaStreamSynchronize((ctx)->stream));   \
    } while (0)
#define SVR_KERNEL_CHECK(ctx)                                                      \
    do {                                                                           \
        (ctx)->launches++;                                                         \
        cudaError_t _e = cudaGetLastError();                                       \
        if (_e != cudaSuccess) return svr_fail((ctx), "kernel launch", _e, __FILE__, __LINE__); \
    } while (0)

// Every extern "C" entry point that takes a context runs on the context's device whatever device is current in the calling
// thread, and restores the caller's device on return (a host that drives several contexts from one thread, as the
// reference's own host loop over devicesToUse does).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const svr_context* c)
    {
        if (!c) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != c->device) { prev = cur; cudaSetDevice(c->device); }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define SVR_ENTRY(ctx) DeviceGuard _svr_device_guard(ctx)

// RAII event bracket around a launch (no-op unless svr_profile_enable(ctx, 1))
struct ProfScope {
    svr_context* c; int idx;
    ProfScope(svr_context* ctx, int kind);
    ~ProfScope();
};

// launchers implemented in svr_psf.cu / svr_em.cu
int svr_launch_build_geom(svr_context* c);
int svr_launch_gaussian_scatter(svr_context* c);
int svr_launch_simulate(svr_context* c);
int svr_launch_superres_scatter(svr_context* c);
int svr_launch_pack_volume(svr_context* c);
int svr_launch_equalize(svr_context* c);
int svr_launch_regularize(svr_context* c, int adaptive, float alpha, float min_i, float max_i, float delta, float lambda);
int svr_launch_init_em(svr_context* c);
int svr_launch_estep(svr_context* c, float m, float sigma, float mix);
int svr_launch_mstep(svr_context* c, double out5[5]);
int svr_launch_scale(svr_context* c);
int svr_launch_robust_init(svr_context* c, double out2[2]);
int svr_launch_mask_volume(svr_context* c);
int svr_launch_scale_volume_sums(svr_context* c, double out2[2]);
int svr_launch_scale_volume_apply(svr_context* c, float scale);
int svr_launch_restore(svr_context* c, const float* d_factors, const int* d_index);
int svr_launch_compact_valid(svr_context* c);
int svr_launch_unpack_acc(svr_context* c);
int svr_launch_equalize_inplace(svr_context* c);
int svr_launch_deinterleave(svr_context* c, const float2* src, float* dst, int component);
int svr_launch_flags_to_int(svr_context* c, const unsigned char* src, int* dst, size_t n);
// svr_window.cu
int svr_window_build_tiles(svr_context* c);
int svr_window_build_maps(svr_context* c);
void svr_window_free(svr_context* c);
bool svr_window_scatter_available(const svr_context* c);
bool svr_window_simulate_available(const svr_context* c);
int svr_launch_window_scatter(svr_context* c, int mode, int only_class);   // mode 0: K3, 1: K1 pass 2; only_class -1: all slices
int svr_launch_window_simulate(svr_context* c, int only_class);
