/*
 * pvr_abi.h -- C ABI of the PVR (patch-to-volume reconstruction) path of libsvr_b200.so.
 *
 * Drop-in for what irtkPatchBasedReconstruction<T>::run() (source/reconstructionGPU2/irtkPatchBasedReconstruction.cpp:194-593)
 * calls on the device side: ReconVolume<T> (include/reconVolume.cuh, reconVolume.cu), the device part of
 * PatchBasedVolume<T> (include/patchBasedVolume.cuh), the free functions initPatchBasedRecon_gpu,
 * patchBasedPSFReconstruction_gpu, patchBasedSimulatePatches_gpu and the classes patchBasedSuperresolution_gpu<T>,
 * patchBasedRobustStatistics_gpu<T>.  Same conventions as svr_abi.h (plain C types, status returns,
 * svr_last_error, one context = one GPU); the context type is shared: a PVR context is an svr_context created in
 * the PVR flavour (12^3 PSF support, PVR PSF constants, __step = 1e-5, E-step gated on the previous weight, the
 * forward read through the reference's un-offset linear texture = 8-voxel mean).
 *
 * Data model: the reference keeps one PatchBasedVolume per stack; here the patch grids of all stacks are
 * concatenated into one float[nPatches][pby][pbx] cube (stack 0 first) with per-patch voxel sizes and matrices,
 * so every step is ONE launch over all stacks.  "ref:" paths are relative to source/reconstructionGPU2/.
 */
#ifndef PVR_ABI_H
#define PVR_ABI_H

#include "svr_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ref: cudaSetDevice + object construction in irtkPatchBasedReconstruction.cpp:300-430 */
int pvr_create(svr_context **out, int device);

/* ---- ReconVolume<T> ------------------------------------------------------------------------------ */
/* ref: ReconVolume::init(int dev, uint3 s, float3 d, const Matrix4& W2I, const Matrix4& I2W) include/reconVolume.cuh:46-72 */
int pvr_recon_init(svr_context *ctx, int sx, int sy, int sz, float dx, float dy, float dz, const float recon_w2i[16],
                   const float recon_i2w[16]);
/* ref: ReconVolume::setMask(char*) include/reconVolume.cuh:103-106 */
int pvr_recon_set_mask(svr_context *ctx, const signed char *mask);
/* ref: ReconVolume::reset() include/reconVolume.cuh:88-92 (volume + volume weights := 0) */
int pvr_recon_reset(svr_context *ctx);
/* ref: ReconVolume::resetAddonCmap() include/reconVolume.cuh:94-98 */
int pvr_recon_reset_addon_cmap(svr_context *ctx);
/* ref: ReconVolume::equalize() reconVolume.cu:77-100 (P5) */
int pvr_recon_equalize(svr_context *ctx);
/* ref: Volume::copyFromHost / copyToHost include/volume.cuh; irtkPatchBasedReconstruction.cpp:456,585 */
int pvr_recon_copy_from_host(svr_context *ctx, const float *data);
int pvr_recon_copy_to_host(svr_context *ctx, float *data);

/* ---- PatchBasedVolume<T>, all stacks at once -------------------------------------------------------- */
/* ref: PatchBasedVolume::init (device allocations + reset) include/patchBasedVolume.cuh:108-210.
 * pbx x pby = patch bounding box (--patchSize), patches_per_stack[n_stacks], stack_dims[n_stacks][3] = voxel size
 * of each stack (patch dim.z = stack dz, patchBasedVolume.cuh:111). */
int pvr_patches_init(svr_context *ctx, int pbx, int pby, int n_stacks, const int *patches_per_stack, const float *stack_dims);
/* ref: PatchBasedVolume::updateTransformationMatrices + the ImagePatch2D members I2W, W2I, Transformation,
 * InvTransformation (include/ImagePatch2D.cuh:33-53); all [nPatches][16]. */
int pvr_patches_set_matrices(svr_context *ctx, const float *i2w, const float *w2i, const float *transformation,
                             const float *inv_transformation);
/* ref: ImagePatch2D::spxMask char[64*64] ('1' = inside the superpixel), include/ImagePatch2D.cuh:51; masks =
 * char[nPatches][4096] or NULL; use_spx = the m_superpixel flag passed to the kernels. */
int pvr_patches_set_spx_masks(svr_context *ctx, const char *masks, int use_spx);
/* ref: patch buffer d_m_PatchesPtr copy in / out (PatchBasedVolume::copyFromHost, getPatchesPtr) */
int pvr_patches_copy_from_host(svr_context *ctx, const float *cube);
int pvr_patches_copy_to_host(svr_context *ctx, float *cube);
/* ref: cudaMemcpyToSymbol(_PSF, ...) initPatchBasedRecon_gpu.cu:95 with the PointSpreadFunction built at
 * irtkPatchBasedReconstruction.cpp:401-415 */
int pvr_set_psf(svr_context *ctx, const int psf_size[3], const float psf_i2w[16], float quality_factor);
/* ref: initPatchBasedRecon_gpu(dev, PatchBasedVolume&, ReconVolume&, PSF&, useSpx) initPatchBasedRecon_gpu.cu:44-133 (P0):
 * fills the patch buffers of one stack from the stack volume stack_data[sz][sy][sx]. */
int pvr_init_patch_based_recon(svr_context *ctx, int stack, const float *stack_data, int sx, int sy, int sz,
                               const float stack_w2i[16]);

/* ---- the hot path (every call covers all stacks) ---------------------------------------------------- */
/* ref: patchBasedPSFReconstruction_gpu patchBasedPSFReconstruction_gpu.cu:41-160 (P1); accumulates into the volume */
int pvr_psf_reconstruction(svr_context *ctx);
/* Multi-GPU (one process per GPU, patches sharded over the ranks; the reference is single-device here): P1 in two
 * phases around the host's all-reduce of svr_device_buffer(ctx, SVR_BUF_ACCUMULATOR, ...); P3 needs no extra entry point
 * (all-reduce the same buffer between pvr_superresolution_run and pvr_superresolution_regularize); the robust-statistics
 * partial sums come from svr_initialize_robust_statistics_local / svr_mstep_local + svr_mstep_finish (same context type). */
int pvr_psf_reconstruction_local(svr_context *ctx);
int pvr_psf_reconstruction_finish(svr_context *ctx);
/* ref: patchBasedSimulatePatches_gpu patchBasedSimulatePatches_gpu.cu:41-146 (P2) incl. updateReconTex */
int pvr_simulate_patches(svr_context *ctx);
/* ref: patchBasedSuperresolution_gpu<T>::run patchBasedSuperresolution_gpu.cu:34-144 (P3); accumulates into addon / cmap */
int pvr_superresolution_run(svr_context *ctx);
/* ref: patchBasedSuperresolution_gpu<T>::regularize patchBasedSuperresolution_gpu.cu:152-287 (P4); the reference
 * hard-codes delta = 1, lambda = 0.1, alpha = 0.5 (:293-295) */
int pvr_superresolution_regularize(svr_context *ctx, int adaptive, float alpha, float min_intensity, float max_intensity,
                                   float delta, float lambda);

/* ---- patchBasedRobustStatistics_gpu<T> ------------------------------------------------------------- */
/* ref: initializeEMValues patchBasedRobustStatistics_gpu.cu:41-95 (scale = weight = 1 per patch; voxel weights) */
int pvr_rs_initialize_em_values(svr_context *ctx);
/* ref: InitializeRobustStatistics :746-851 (device sums; the scalar initialisation is host arithmetic) */
int pvr_rs_initialize_robust_statistics(svr_context *ctx, float *sigma);
/* ref: EStep :226-275 (EStepKernel + per-patch potentials).  patch_potential[nPatches], correctly indexed. */
int pvr_rs_estep_device(svr_context *ctx, float m, float sigma, float mix, float *patch_potential);
/* ref: copyFromWeightsAndScales / copyToWeightsAndScales :152-196 */
int pvr_rs_get_scales_weights(svr_context *ctx, float *scales, float *patch_weights);
int pvr_rs_set_scales_weights(svr_context *ctx, const float *scales, const float *patch_weights);
/* ref: EStep host part :277-520, LITERAL: potentials are replayed through the reference's
 * `patch_potential[j]` indexing without the stack offset (:268,272).  state5 = {sigma_s, mix_s, mean_s, mean_s2,
 * sigma_s2} in/out; patch_weight[nPatches] in/out; potential_used[nPatches] (may be NULL) receives the vector the
 * reference actually used.  Pure host code. */
int pvr_host_patch_em(int n_stacks, const int *patches_per_stack, const float *patch_potential, const float *scale,
                      float *patch_weight, float step, float state5[5], float *potential_used);
/* ref: MStep :572-640 */
int pvr_rs_mstep(svr_context *ctx, int iter, float step, float *sigma, float *mix, float *m);
/* ref: Scale :642-744 (per-patch scale, stored into the patches) */
int pvr_rs_scale(svr_context *ctx, float *scale_vec);

/* Debug taps: same kinds as svr_debug_get (weights, simulated patches / weights / inside, addon, confidence map,
 * PSF sums, patches). */
int pvr_debug_get(svr_context *ctx, int kind, void *out);

#ifdef __cplusplus
}
#endif
#endif /* PVR_ABI_H */
