/*
 * svr_abi.h -- C ABI of libsvr_b200.so: the B200-native SVR (slice-to-volume reconstruction) device
 * library.  It is a drop-in for the public method set of `class Reconstruction` of
 * bkainz/fetalReconstruction (source/reconstructionGPU2/include/reconstruction_cuda2.cuh:92-341),
 * i.e. the boundary `irtkReconstruction::*GPU()` calls into (irtkReconstructionGPU.cc).
 *
 * Conventions
 *   - plain C types only: host pointers, sizes, scalars; matrices are float[16] row-major
 *     (the reference's Matrix4 = float4 data[4], recon_volumeHelper.cuh:41-46).
 *   - every call returns 0 on success, non-zero on failure; svr_last_error() gives the text.
 *     (The reference prints and exit()s, reconstruction_cuda2.cuh:78-86; a CLI built on this
 *     library does the exit itself.)
 *   - one context = one GPU = one rank.  Calls are synchronous with respect to the host like the
 *     reference's (each returns after its stream work is complete) unless stated otherwise.
 *   - volumes are x-fastest: idx = x + y*sx + z*sx*sy; the slice cube is float[S][Ny][Nx] with
 *     -1 marking padding (irtkReconstructionGPU.cc:269-314).
 *   - there is no CPU fallback: without a CUDA device svr_create fails.
 *
 * "ref:" cites the reference interface each entry replaces (paths relative to
 * source/reconstructionGPU2/; cuda2.cu = reconstruction_cuda2.cu, .cuh = include/reconstruction_cuda2.cuh).
 */
#ifndef SVR_ABI_H
#define SVR_ABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVR_ABI_VERSION 2

typedef struct svr_context svr_context;

/* ---- life cycle ------------------------------------------------------------------------- */
/* ref: Reconstruction::Reconstruction(std::vector<int> dev, bool) .cuh:94, cuda2.cu:616-706.
 * One device per context (multi-GPU = one context per rank, section "multi-rank" below). */
int svr_create(svr_context **out, int device);
/* ref: Reconstruction::~Reconstruction cuda2.cu:1232-1310 */
int svr_destroy(svr_context *ctx);
const char *svr_last_error(const svr_context *ctx);   /* ctx may be NULL: error of a failed svr_create */
int svr_abi_version(void);
/* Run all work of this context on a caller-owned cudaStream_t (NULL = the context's own stream).
 * ref: cudaStream_t streams[] .cuh:141 */
int svr_set_stream(svr_context *ctx, void *cuda_stream);
/* Block until all queued work of the context is complete.  ref: cudaDeviceSynchronize calls. */
int svr_synchronize(svr_context *ctx);
/* The cudaStream_t the context launches on (its own, or the one given to svr_set_stream): a caller that issues collectives
 * (ncclAllReduce on svr_device_buffer) enqueues them here, between the *_local and *_finish calls, with no host round trip. */
int svr_get_stream(svr_context *ctx, void **cuda_stream);
/* Asynchronous mode.  The reference's calls all return after a device synchronisation
 * (cuda2.cu: cudaDeviceSynchronize / per-GPU cudaStreamSynchronize), and so do ours by default.  With on != 0 an entry point
 * that returns no host data only ENQUEUES its work on the context's stream and returns; entry points that return host data
 * (slice potentials, scale vector, statistics, volumes) still complete before they return.  Host input buffers must then stay
 * valid until svr_synchronize (pageable memory is staged by the driver before the call returns; pinned memory is not). */
int svr_set_async(svr_context *ctx, int on);
/* Makes the context's device the calling thread's current device (and leaves it so): for a host thread that drives one rank
 * and calls NCCL next to the library.  ref: cudaSetDevice(dev) at the top of every *OnX(dev), e.g. cuda2.cu:2385. */
int svr_make_current(svr_context *ctx);
/* Number of kernels this context has launched since creation (for bench.py's gpu_launches). */
int64_t svr_launch_count(const svr_context *ctx);
/* Kernel-variant selection for A/B measurements (tools/, bench.py --tune); results agree within float summation order.
 * No reference counterpart (the reference has compile-time switches only, .cuh:53-75). */
enum {
    SVR_TUNE_SCATTER = 0,   /* K1 pass 2 / K3: 0 = paired per-lane reductions (default: fastest on every orientation measured,
                               profiles/r02_ww_ab.txt), 1 = warp windows + SIMT flush, 2 = warp windows + TMA reduce flush,
                               3 = paired for slices aligned with the volume axes and warp windows for the others */
    SVR_TUNE_SIMULATE = 1,  /* K2: 0 = per-tap loads + staged rows (default), 1 = TMA-staged windows for every tile,
                               2 = windows for through-plane slices only */
    SVR_TUNE_REGULARIZE = 2 /* K4 + K5: 0 = two kernels through L1/L2, 1 = one fused kernel with a shared-memory halo tile (default) */
};
int svr_set_tuning(svr_context *ctx, int key, int value);

/* ---- uploads ---------------------------------------------------------------------------- */
/* ref: InitReconstructionVolume(uint3 s, float3 dim, float* data, float sigma_bias) .cuh:230,
 * cuda2.cu:1160-1230.  data may be NULL (zeros).  dim = voxel size in mm. */
int svr_init_reconstruction_volume(svr_context *ctx, int sx, int sy, int sz, float dx, float dy, float dz,
                                   const float *data);
/* ref: setMask(uint3, float3, float* data, float sigma_bias) .cuh:240, cuda2.cu:1095-1158.
 * (The smoothed copy maskC_ only feeds the unreachable bias correction and is not built.) */
int svr_set_mask(svr_context *ctx, int sx, int sy, int sz, const float *mask);
/* ref: initStorageVolumes(uint3 size, float3 dim) .cuh:214, cuda2.cu:1408-1573.  Allocates the
 * per-pixel buffers for S slices of Nx x Ny pixels.  All S slices are kept (deviation D2). */
int svr_init_storage_volumes(svr_context *ctx, int Nx, int Ny, int S);
/* ref: FillSlices(float* sdata, vector<int> sizesX, vector<int> sizesY) .cuh:216, cuda2.cu:1574-1656.
 * sizesX/sizesY may be NULL (they are informational in the reference too). */
int svr_fill_slices(svr_context *ctx, const float *cube, const int *sizesX, const int *sizesY);
/* ref: setSliceDims(vector<float3> slice_dims, float quality_factor) .cuh:220, cuda2.cu:772-833. */
int svr_set_slice_dims(svr_context *ctx, const float *dims_xyz /* [S][3] */, float quality_factor);
/* ref: SetSliceMatrices(T, Tinv, I2Winit, W2Iinit, I2W, W2I, reconI2W, reconW2I) .cuh:222-224,
 * cuda2.cu:835-907.  All per-slice arrays are [S][16]. */
int svr_set_slice_matrices(svr_context *ctx, const float *transforms, const float *inv_transforms,
                           const float *slices_i2w, const float *slices_w2i, const float recon_i2w[16],
                           const float recon_w2i[16]);
/* ref: generatePSFVolume(float* psf, uint3, float3, float3, Matrix4 PSFI2W, Matrix4 PSFW2I, float q)
 * .cuh:218, cuda2.cu:718-750 (uploads constants only; the sampled PSF volume is never read). */
int svr_generate_psf_volume(svr_context *ctx, const int psf_size[3], const float psf_i2w[16], float quality_factor);
/* ref: UpdateScaleVector(vector<float> scales, vector<float> slice_weights) .cuh:236, cuda2.cu:1342-1393 */
int svr_update_scale_vector(svr_context *ctx, const float *scales, const float *slice_weights);
/* ref: UpdateSliceWeights(vector<float>) .cuh:228, cuda2.cu:1312-1340 */
int svr_update_slice_weights(svr_context *ctx, const float *slice_weights);
/* ref: UpdateReconstructed(uint3, float* data) .cuh:233, cuda2.cu:1658-1668 */
int svr_update_reconstructed(svr_context *ctx, const float *data);

/* ---- the hot path ----------------------------------------------------------------------- */
/* ref: InitializeEMValues() .cuh:211, cuda2.cu:3241-3310 (K12) */
int svr_initialize_em_values(svr_context *ctx);
/* ref: GaussianReconstruction(vector<int>& voxel_num) .cuh:274, cuda2.cu:2329-2493 (K1, K6, K14).
 * voxel_num[S] (may be NULL): per-SLICE count of pixels that touched the masked volume (deviation D4). */
int svr_gaussian_reconstruction(svr_context *ctx, int *voxel_num);
/* ref: SimulateSlices(vector<bool>& slice_inside) .cuh:257, cuda2.cu:2654-2763 (K2, K13). slice_inside[S] may be NULL. */
int svr_simulate_slices(svr_context *ctx, unsigned char *slice_inside);
/* ref: InitializeRobustStatistics(float& sigma) .cuh:260, cuda2.cu:2243-2308 (K11) */
int svr_initialize_robust_statistics(svr_context *ctx, float *sigma);
/* ref: EStep(float m, float sigma, float mix, vector<float>& slice_potential) .cuh:253, cuda2.cu:2766-2925 (K7, K8) */
int svr_estep(svr_context *ctx, float m, float sigma, float mix, float *slice_potential);
/* ref: MStep(int iter, float step, float& sigma, float& mix, float& m) .cuh:255, cuda2.cu:2927-3112 (K9) */
int svr_mstep(svr_context *ctx, int iter, float step, float *sigma, float *mix, float *m);
/* ref: CalculateScaleVector(vector<float>& scale_vec) .cuh:238, cuda2.cu:3114-3239 (K10).
 * Reproduces the reference's upload order: the device keeps the PREVIOUS scale vector for the
 * kernels (cuda2.cu:3238 copies h_scales before cuda2.cu:3195 updates it); the M-step uses the new one. */
int svr_calculate_scale_vector(svr_context *ctx, float *scale_vec);
/* ref: Superresolution(int iter, vector<float> slice_weight, bool adaptive, float alpha, float min_intensity,
 * float max_intensity, float delta, float lambda, bool global_bias, float sigma_bias, float low_cutoff)
 * .cuh:263-265, cuda2.cu:2119-2241 (K3, K4, K5).  slice_weight may be NULL (keep the uploaded vector). */
int svr_superresolution(svr_context *ctx, int iter, const float *slice_weight, int adaptive, float alpha,
                        float min_intensity, float max_intensity, float delta, float lambda);
/* ref: maskVolume() .cuh:270, cuda2.cu:3313-3347 (K15) */
int svr_mask_volume(svr_context *ctx);
/* ref: ScaleVolume() .cuh:271, cuda2.cu:3386-3470 (K16).  scale_out may be NULL. */
int svr_scale_volume(svr_context *ctx, float *scale_out);
/* ref: RestoreSliceIntensities(vector<float> stack_factors, vector<int> stack_index) .cuh:272, cuda2.cu:3349-3383 (K17).
 * As in the reference this rescales the restore-copy of the slices (v_slices), not the working set. */
int svr_restore_slice_intensities(svr_context *ctx, const float *stack_factors, int n_stacks, const int *stack_index);

/* ---- slice-to-volume registration (--useGPUReg) ------------------------------------------
 * Batched 6-DOF rigid registration of every slice to the current volume, maximising the blurred NCC
 * summed over three in-slice offsets by finite-difference gradient ascent (2 levels x 4 step sizes x
 * <= 20 iterations).  The similarity and the optimiser reproduce the reference literally (DESIGN.md
 * section 6 lists the quirks G1-G4 that are part of the reference's numbers). */
/* ref: initRegStorageVolumes(uint3 size, float3 dim) .cuh:326, cuda2.cu:4892-5018.  W x H = padded size of the
 * slices resampled to the volume resolution, S slices; dim = their voxel size. */
int svr_reg_init_storage(svr_context *ctx, int W, int H, int S, float dx, float dy, float dz);
/* ref: FillRegSlices(float* sdata, vector<Matrix4> slices_resampledI2W) .cuh:328, cuda2.cu:5023-5088.
 * cube = float[S][H][W], -1 = padding.  The matrices are stored but, as in the reference, not used. */
int svr_reg_fill_slices(svr_context *ctx, const float *cube, const float *slices_resampled_i2w /* [S][16] or NULL */);
/* Device form of the resampling that feeds FillRegSlices: irtkReconstruction::PrepareRegistrationSlices
 * (irtkReconstructionGPU.cc:1992-2059) resamples every slice to the reconstruction's voxel size with
 * irtkResamplingWithPadding (IRTKSimple2/image++/src/irtkResamplingWithPadding.cc:36-183, padding -1) on the host and
 * uploads the result.  Here the slices uploaded by svr_fill_slices are resampled in place on the device, in double
 * precision and in the reference's order of operations, into the registration cube of svr_reg_init_storage:
 *   src_from_out [S][24]  rows 0..2 of the resampled slice's image-to-world matrix, then rows 0..2 of the slice's world-to-image
 *                         matrix (double): applied one after the other like the reference's ImageToWorld / WorldToImage calls
 *   in_sizes     [S][2]   valid extent (x, y) of slice s inside the packed slice cube
 *   out_sizes    [S][2]   extent (x, y) of the resampled slice; the rest of its [H][W] plane is padding
 * Equivalent to svr_reg_fill_slices with the host-resampled cube. */
int svr_reg_resample_slices(svr_context *ctx, const double *src_from_out, const int *in_sizes, const int *out_sizes,
                            const float *slices_resampled_i2w /* [S][16] or NULL */);
/* ref: updateResampledSlicesI2W(vector<Matrix4> ofsSlice) .cuh:330, cuda2.cu:4707-4757: image-to-world of
 * the resampled slices with their origin reset to 0 (irtkReconstructionGPU.cc:2226-2250). */
int svr_reg_update_slices_i2w(svr_context *ctx, const float *ofs_slice /* [S][16] */);
/* ref: prepareSliceToVolumeReg() .cuh:332, cuda2.cu:3800-3959: snapshots the current volume for sampling and
 * sets levels = 2, steps = 4, iterations = 20, epsilon = 1e-4, blurring = voxel/2 * 2^level, step = 0.1 * 2^level. */
int svr_reg_prepare(svr_context *ctx);
/* Overrides the schedule set by svr_reg_prepare (tests; the reference hard-codes it at cuda2.cu:3884-3887). */
int svr_reg_set_schedule(svr_context *ctx, int n_levels, int n_steps, int n_iterations);
/* ref: registerSlicesToVolume(vector<Matrix4>& transf_) .cuh:338, cuda2.cu:4760-4867 -> registerMultipleSlicesToVolume
 * cuda2.cu:4001-4141.  transforms [S][16] in/out (slice -> world transform times the origin offset,
 * irtkReconstructionGPU.cc:2243-2247). */
int svr_reg_register(svr_context *ctx, float *transforms);
/* ref: evaluateCostsMultipleSlices(...) cuda2.cu:4150-4221 with all slices active: similarity[S] of the given
 * transforms at `level` (0 = fine).  This is the "registration similarity kernel" of BASELINE.json config 5. */
int svr_reg_evaluate(svr_context *ctx, const float *transforms, int level, float *similarity);
/* Number of (slice, in-slice offset) cost evaluations of the last svr_reg_register / svr_reg_evaluate call. */
int64_t svr_reg_evaluations(const svr_context *ctx);
/* Debug taps: 0 = resampled slices float[S*H*W], 1 = blurred slices of the last level, 2 = similarities
 * float[5*S], 3 = gradient float[7*S].  ref: getRegSlicesVol_debug / debugRegSlicesVolume .cuh:198-205 */
int svr_reg_debug_get(svr_context *ctx, int kind, void *out);

/* ---- downloads -------------------------------------------------------------------------- */
/* ref: syncCPU(float* reconstructed) .cuh:213, cuda2.cu:2031-2037 */
int svr_sync_cpu(svr_context *ctx, float *reconstructed);
/* ref: getVolWeights(float*) .cuh:206 */
int svr_get_vol_weights(svr_context *ctx, float *weights);

/* Debug taps.  ref: debugWeights/Simslices/Simweights/Siminside/ConfidenceMap/Addon/v_PSF_sums .cuh:192-207
 * (+ the slice cube, the restore copy and the K1 voxel counts). */
enum svr_debug_kind {
    SVR_DBG_WEIGHTS = 0,        /* float[S*Ny*Nx] */
    SVR_DBG_SIMSLICES = 1,      /* float[S*Ny*Nx] */
    SVR_DBG_SIMWEIGHTS = 2,     /* float[S*Ny*Nx] */
    SVR_DBG_SIMINSIDE = 3,      /* char [S*Ny*Nx] */
    SVR_DBG_CONFIDENCE_MAP = 4, /* float[V] */
    SVR_DBG_ADDON = 5,          /* float[V] */
    SVR_DBG_PSF_SUMS = 6,       /* float[S*Ny*Nx] */
    SVR_DBG_SLICES = 7,         /* float[S*Ny*Nx] */
    SVR_DBG_VOXEL_COUNT = 8,    /* int  [S*Ny*Nx] */
    SVR_DBG_SLICES_RESTORED = 9,/* float[S*Ny*Nx] */
    SVR_DBG_SCALES_DEVICE = 10, /* float[S]: the scale vector the kernels see */
    SVR_DBG_MASK = 11,          /* float[V] */
    SVR_DBG_WW_STATS = 12       /* unsigned[32]: window-plan statistics of the last scatter launch (svr_window.cu: ww_count) */
};
int svr_debug_get(svr_context *ctx, int kind, void *out);

/* ---- multi-rank (one process per GPU; stacks sharded; volume accumulators all-reduced) ----
 * The reference reduces to device 0 and re-broadcasts (cuda2.cu:2225-2239, 2445-2460, 2175-2180).
 * Here every rank holds its own slices and a full replica of the volume; the split-phase calls
 * below expose the exchange points so the caller can all-reduce the accumulator (NCCL over
 * NVLink: ncclAllReduce or torch.distributed on svr_device_buffer()) between the two halves.
 * svr_gaussian_reconstruction == _local + _finish; svr_superresolution == _local + _finish. */
int svr_gaussian_reconstruction_local(svr_context *ctx);
int svr_gaussian_reconstruction_finish(svr_context *ctx, int *voxel_num);
int svr_superresolution_local(svr_context *ctx, const float *slice_weight);
int svr_superresolution_finish(svr_context *ctx, int adaptive, float alpha, float min_intensity, float max_intensity,
                               float delta, float lambda);
/* Partial statistics (this rank's slices); sum / min / max them over ranks, then *_finish.
 * sums5 = {sum e^2 w, sum w, n, min e, max e}   (ref: MStepOnX cuda2.cu:3074-3112) */
int svr_mstep_local(svr_context *ctx, double sums5[5]);
/* Pure host arithmetic of Reconstruction::MStep, cuda2.cu:3056-3071. */
int svr_mstep_finish(const double sums5[5], int iter, float step, float *sigma, float *mix, float *m);
/* sums2 = {sum (s-sim)^2, n}  (ref: cuda2.cu:2296-2305) */
int svr_initialize_robust_statistics_local(svr_context *ctx, double sums2[2]);
/* sums2 = {scalenum, scaleden}  (ref: cuda2.cu:3446-3454); then svr_scale_volume_apply on every rank. */
int svr_scale_volume_local(svr_context *ctx, double sums2[2]);
int svr_scale_volume_apply(svr_context *ctx, float scale);

enum svr_buffer_kind {
    SVR_BUF_ACCUMULATOR = 0, /* float2[V], interleaved {numerator, denominator}: {sum psf*s, sum psf} after
                                svr_gaussian_reconstruction_local, {addon, cmap} after svr_superresolution_local */
    SVR_BUF_RECON = 1        /* float[V] */
};
/* Device address + byte size of a context-owned buffer (for collectives issued by the caller on the
 * context's stream).  SVR_BUF_ACCUMULATOR is valid until svr_destroy / re-initialisation of the volume.
 * SVR_BUF_RECON must be re-queried after every svr_superresolution / svr_superresolution_finish call: the regulariser
 * writes into a second buffer and the two are swapped (double-buffered, deviation D3), so a cached pointer then refers to
 * the scratch copy. */
int svr_device_buffer(svr_context *ctx, int kind, void **dev_ptr, size_t *nbytes);

/* ---- per-kernel device timing (CUDA events on the launching stream; for bench.py's roofline) ---- */
enum svr_kernel_kind {
    SVR_K_GAUSSIAN = 0,   /* K1  gaussian_scatter_kernel  */
    SVR_K_SIMULATE = 1,   /* K2  simulate_kernel          */
    SVR_K_SUPERRES = 2,   /* K3  superres_scatter_kernel  */
    SVR_K_REGULARIZE = 3, /* K4+K5                        */
    SVR_K_EM = 4,         /* (round 1: all robust-statistics reductions; now split into the four kinds below) */
    SVR_K_REG_EVAL = 5,   /* reg_eval_kernel: fused sample + blur + NCC moments of the registration */
    SVR_K_ESTEP = 6,      /* K7+K8 estep_kernel (+ potential_finish_kernel) */
    SVR_K_MSTEP = 7,      /* K9 mstep_kernel (+ fold) */
    SVR_K_SCALE = 8,      /* K10 scale_kernel (+ scale_finish_kernel) */
    SVR_K_ROBUST_INIT = 9,/* K11 robust_init_kernel (+ fold) */
    SVR_K_COUNT = 10
};
/* When enabled every launch of the kinds above is bracketed by a cudaEvent pair (no host sync). */
int svr_profile_enable(svr_context *ctx, int on);
/* Synchronises the stream, folds all pending event pairs and returns the accumulated device time and
 * launch count of one kind since the last svr_profile_reset. */
int svr_profile_read(svr_context *ctx, int kind, double *total_ms, int64_t *launches);
int svr_profile_reset(svr_context *ctx);

/* ---- pure host helpers (no device work; usable without a GPU) ---------------------------- */
/* Slice-level EM of irtkReconstruction::EStepGPU, irtkReconstructionGPU.cc:3184-3440.
 * slice_potential[S] in/out (exclusions set to -1), scale[S], slice_weight[S] in/out,
 * state5 = {sigma_s, mix_s, mean_s, mean_s2, sigma_s2} in/out. */
int svr_host_slice_em(int S, float *slice_potential, const float *scale, float *slice_weight, const int *force_excluded,
                      int n_force_excluded, const int *small_slices, int n_small_slices, double step, float state5[5]);
/* "Small slices" rule of GaussianReconstructionGPU, irtkReconstructionGPU.cc:2714-2726, applied to
 * per-slice counts (deviation D4).  out_small[S] receives indices; returns their number via n_out. */
int svr_host_small_slices(int S, const int *voxel_num, int *out_small, int *n_out);
/* Contiguous, balanced assignment of whole stacks (or of slices when n_stacks < nranks) to ranks.
 * slices_per_stack[n_stacks]; out_begin/out_end = this rank's [begin,end) in global slice order.
 * ref (what it replaces): floor(N/D)-sized ranges that drop the remainder, cuda2.cu:1413-1457. */
int svr_host_partition(int n_stacks, const int *slices_per_stack, int nranks, int rank, int *out_begin, int *out_end);
/* The balanced split bench.py and the C++ host use: rank r takes every nranks-th slice of EVERY stack (slice j of a stack with
 * j % nranks == rank), so every rank sees the same mix of stack orientations and of positions along the stacks (the cost of the
 * PSF kernels depends on both: whole stacks per rank measured 22 % imbalance at N = 2, contiguous ranges 60 % efficiency at
 * N = 4).  out_indices receives the rank's slices as indices into the stack-major global order (capacity: sum of
 * slices_per_stack), out_count their number.  ref: the contiguous split of cuda2.cu:1408-1457 (which also drops slices, Q6). */
int svr_host_partition_strided(int n_stacks, const int *slices_per_stack, int nranks, int rank, int *out_indices, int *out_count);

/* ---- IRTK-style rigid registration on the device (csrc/svr_rreg.cu) ---------------------------------------------------
 * The engine behind the reference's DEFAULT registrations, which all run irtkImageRigidRegistrationWithPadding on the CPU:
 *   kind 0  stack-to-template      ref: irtkReconstruction::StackRegistrations, irtkReconstructionGPU.cc:849-1001
 *                                       (GuessParameterThickSlices, SetTargetPadding(0))
 *   kind 1  slice / patch to volume ref: ParallelSliceToVolumeRegistration, irtkReconstructionGPU.cc:1992-2059, and
 *                                       ParallelPatchToVolumeRegistration, patchBased2D3DRegistration.cpp:88-168
 *                                       (GuessParameterSliceToVolume(false), SetTargetPadding(-1))
 * Images are the reference's irtkGreyImage: short voxels (x fastest) + the 18 numbers of irtkImageAttributes
 * {x, y, z, dx, dy, dz, origin[3], xaxis[3], yaxis[3], zaxis[3]}.  voxels[n_images] / attrs18[n_images][18] list the images;
 * item i registers target images[target_of_item[i]] to source images[source_of_item[i]], starting from and returning
 * dofs[6 i ..] = {tx, ty, tz (mm), rx, ry, rz (degrees)} (irtkRigidTransformation).  similarity (may be NULL) receives the final
 * cross-correlation per item, evaluations (may be NULL) the number of similarity evaluations.
 * level_only >= 0 (test tap): no optimisation; that resolution level is prepared, similarity[i] = the metric at dofs, and item 0's
 * prepared target / source images are copied out (buffers of prepared_capacity voxels each; NULL to skip). */
int svr_rreg_register(svr_context *ctx, int n_items, int n_images, const short *const *voxels, const double *attrs18,
                      const int *target_of_item, const int *source_of_item, int kind, double *dofs, double *similarity,
                      int64_t *evaluations, int level_only, size_t prepared_capacity, short *prepared_target, double *prepared_target_attr18,
                      short *prepared_source, double *prepared_source_attr18);
/* The reference's padded image filters on the device, on their own:
 * ref: irtkGaussianBlurringWithPadding<irtkGreyPixel>::Run (image++/src/irtkGaussianBlurringWithPadding.cc:36-117) */
int svr_rreg_blur_with_padding(svr_context *ctx, const short *voxels, const double attr18[18], double sigma, int padding, short *out);
/* ref: irtkResamplingWithPadding<irtkGreyPixel>::Run (image++/src/irtkResamplingWithPadding.cc:36-183, irtkResampling.cc:74-131) */
int svr_rreg_resample_with_padding(svr_context *ctx, const short *voxels, const double attr18[18], double dx, double dy, double dz,
                                   int padding, short *out, size_t out_capacity, double out_attr18[18]);

#ifdef __cplusplus
}
#endif
#endif /* SVR_ABI_H */
