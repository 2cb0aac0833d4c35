/* TEST INFRASTRUCTURE (oracle/): C entry points around the UNMODIFIED reference `class Reconstruction`
 * (/root/reference/source/reconstructionGPU2/include/reconstruction_cuda2.cuh:92-341), so the reference's own
 * CUDA path can be driven from Python (ctypes) on a B200 and its outputs compared with ours and with the CPU
 * oracle.  Built into oracle/_ref/libref_cuda2.so by oracle/Makefile (`make ref`).  This file contains no
 * arithmetic: every function forwards to the reference method named in its comment, or copies one of the
 * class's public device buffers back to the host.  Never linked into the product library.
 *
 * NB the reference constructor calls cudaDeviceReset() (cuda2.cu:630): load this library only in a process of
 * its own (oracle/ref_runner.py), never next to torch or libsvr_b200.so. */
/* included at the end of ref_cuda2_tu.cu: same translation unit as the reference file, so its kernels and
 * file-local helpers (initActiveSlices, FilterGaussStack, ...) are visible without relocatable device code. */

#include <cstring>
#include <vector>

namespace {
struct RefHandle {
  Reconstruction* r;
  int S;
};
inline Matrix4 m4(const float* p) {
  Matrix4 m;
  for (int i = 0; i < 4; ++i) m.data[i] = make_float4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
  return m;
}
inline std::vector<Matrix4> m4v(const float* p, int n) {
  std::vector<Matrix4> v(n);
  for (int i = 0; i < n; ++i) v[i] = m4(p + 16 * i);
  return v;
}
inline void m4out(const Matrix4& m, float* p) {
  for (int i = 0; i < 4; ++i) {
    p[4 * i] = m.data[i].x; p[4 * i + 1] = m.data[i].y; p[4 * i + 2] = m.data[i].z; p[4 * i + 3] = m.data[i].w;
  }
}
template <class T>
inline size_t vcount(const Volume<T>& v) { return (size_t)v.size.x * v.size.y * v.size.z; }
}  // namespace

extern "C" {

/* Reconstruction::Reconstruction (cuda2.cu:620) + the flags GPU.cc:220-237 sets on it */
void* ref_create(int device, int multithreaded, int debug_gpu) {
  std::vector<int> dev(1, device);
  RefHandle* h = new RefHandle;
  h->r = new Reconstruction(dev, multithreaded != 0);
  h->r->_useCPUReg = false;
  h->r->_debugGPU = debug_gpu != 0;
  h->r->_disableBiasC = true; /* CLI default: --disableBiasCorrection defaults to true (reconstruction.cc:121,202) */
  h->S = 0;
  return h;
}
/* InitReconstructionVolume (cuda2.cu:1159), reconstructedVoxelSize (GPU.cc:317) */
void ref_init_reconstruction_volume(void* hv, int sx, int sy, int sz, float dx, float dy, float dz, float* data,
                                    float sigma_bias) {
  RefHandle* h = (RefHandle*)hv;
  h->r->InitReconstructionVolume(make_uint3(sx, sy, sz), make_float3(dx, dy, dz), data, sigma_bias);
  h->r->reconstructedVoxelSize = dx;
}
/* setMask (cuda2.cu:1095) */
void ref_set_mask(void* hv, int sx, int sy, int sz, float dx, float dy, float dz, float* mask, float sigma_bias) {
  ((RefHandle*)hv)->r->setMask(make_uint3(sx, sy, sz), make_float3(dx, dy, dz), mask, sigma_bias);
}
/* initStorageVolumes (cuda2.cu:1408) */
void ref_init_storage_volumes(void* hv, int Nx, int Ny, int S, float dx, float dy, float dz) {
  RefHandle* h = (RefHandle*)hv;
  h->S = S;
  h->r->initStorageVolumes(make_uint3(Nx, Ny, S), make_float3(dx, dy, dz));
}
/* FillSlices (cuda2.cu:1574) */
void ref_fill_slices(void* hv, float* cube, const int* sizesX, const int* sizesY) {
  RefHandle* h = (RefHandle*)hv;
  h->r->FillSlices(cube, std::vector<int>(sizesX, sizesX + h->S), std::vector<int>(sizesY, sizesY + h->S));
}
/* setSliceDims */
void ref_set_slice_dims(void* hv, const float* dims, float quality) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<float3> d(h->S);
  for (int i = 0; i < h->S; ++i) d[i] = make_float3(dims[3 * i], dims[3 * i + 1], dims[3 * i + 2]);
  h->r->setSliceDims(d, quality);
}
/* SetSliceMatrices, argument order as called from GPU.cc:397-398 */
void ref_set_slice_matrices(void* hv, const float* T, const float* Tinv, const float* a3, const float* a4,
                            const float* a5, const float* a6, const float* reconI2W, const float* reconW2I) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<Matrix4> v3 = m4v(a3, h->S), v4 = m4v(a4, h->S), v5 = m4v(a5, h->S), v6 = m4v(a6, h->S);
  h->r->SetSliceMatrices(m4v(T, h->S), m4v(Tinv, h->S), v3, v4, v5, v6, m4(reconI2W), m4(reconW2I));
}
/* generatePSFVolume */
void ref_generate_psf_volume(void* hv, float* psf, const int* psf_size, const float* slice_dim, const float* psf_dim,
                             const float* psf_i2w, const float* psf_w2i, float quality) {
  ((RefHandle*)hv)->r->generatePSFVolume(psf, make_uint3(psf_size[0], psf_size[1], psf_size[2]),
                                   make_float3(slice_dim[0], slice_dim[1], slice_dim[2]),
                                   make_float3(psf_dim[0], psf_dim[1], psf_dim[2]), m4(psf_i2w), m4(psf_w2i), quality);
}
void ref_update_scale_vector(void* hv, const float* scales, const float* weights) {
  RefHandle* h = (RefHandle*)hv;
  h->r->UpdateScaleVector(std::vector<float>(scales, scales + h->S), std::vector<float>(weights, weights + h->S));
}
void ref_update_slice_weights(void* hv, const float* weights) {
  RefHandle* h = (RefHandle*)hv;
  h->r->UpdateSliceWeights(std::vector<float>(weights, weights + h->S));
}
void ref_update_reconstructed(void* hv, int sx, int sy, int sz, float* data) {
  ((RefHandle*)hv)->r->UpdateReconstructed(make_uint3(sx, sy, sz), data);
}
void ref_initialize_em_values(void* hv) { ((RefHandle*)hv)->r->InitializeEMValues(); }
/* GaussianReconstruction: returns the per-DEVICE voxel_num vector (Q10) */
int ref_gaussian_reconstruction(void* hv, int* voxel_num, int cap) {
  std::vector<int> v;
  ((RefHandle*)hv)->r->GaussianReconstruction(v);
  for (int i = 0; i < (int)v.size() && i < cap; ++i) voxel_num[i] = v[i];
  return (int)v.size();
}
void ref_simulate_slices(void* hv, unsigned char* slice_inside) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<bool> v(h->S, false);
  h->r->SimulateSlices(v);
  for (int i = 0; i < h->S; ++i) slice_inside[i] = v[i] ? 1 : 0;
}
void ref_initialize_robust_statistics(void* hv, float* sigma) { ((RefHandle*)hv)->r->InitializeRobustStatistics(*sigma); }
void ref_estep(void* hv, float m, float sigma, float mix, float* slice_potential) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<float> v(slice_potential, slice_potential + h->S);
  h->r->EStep(m, sigma, mix, v);
  for (int i = 0; i < h->S; ++i) slice_potential[i] = v[i];
}
void ref_mstep(void* hv, int iter, float step, float* sigma, float* mix, float* m) {
  ((RefHandle*)hv)->r->MStep(iter, step, *sigma, *mix, *m);
}
void ref_calculate_scale_vector(void* hv, float* scale) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<float> v(scale, scale + h->S);
  h->r->CalculateScaleVector(v);
  for (int i = 0; i < h->S; ++i) scale[i] = v[i];
}
void ref_superresolution(void* hv, int iter, const float* slice_weight, int adaptive, float alpha, float min_i,
                         float max_i, float delta, float lambda, int global_bias, float sigma_bias, float low_cut) {
  RefHandle* h = (RefHandle*)hv;
  h->r->Superresolution(iter, std::vector<float>(slice_weight, slice_weight + h->S), adaptive != 0, alpha, min_i, max_i,
                        delta, lambda, global_bias != 0, sigma_bias, low_cut);
}
void ref_mask_volume(void* hv) { ((RefHandle*)hv)->r->maskVolume(); }
void ref_scale_volume(void* hv) { ((RefHandle*)hv)->r->ScaleVolume(); }
void ref_restore_slice_intensities(void* hv, const float* factors, int n_stacks, const int* stack_index) {
  RefHandle* h = (RefHandle*)hv;
  h->r->RestoreSliceIntensities(std::vector<float>(factors, factors + n_stacks),
                                std::vector<int>(stack_index, stack_index + h->S));
}
void ref_sync_cpu(void* hv, float* out) { ((RefHandle*)hv)->r->syncCPU(out); }
void ref_get_vol_weights(void* hv, float* out) { ((RefHandle*)hv)->r->getVolWeights(out); }

/* Copies of the class's public device buffers (device 0).  kind: 0 weights, 1 simslices, 2 simweights,
 * 3 siminside (char), 4 v_PSF_sums, 5 sliceVoxel_count (int), 6 addon, 7 confidence map, 8 slices,
 * 9 volume weights (dev_volume_weights_), 10 reconstructed. */
int ref_get(void* hv, int kind, void* out) {
  Reconstruction* r = ((RefHandle*)hv)->r;
  const int d = r->devicesToUse[0];
  cudaDeviceSynchronize();
  const void* src = nullptr;
  size_t bytes = 0;
  switch (kind) {
    case 0: src = r->dev_v_weights[d].data; bytes = vcount(r->dev_v_weights[d]) * 4; break;
    case 1: src = r->dev_v_simulated_slices[d].data; bytes = vcount(r->dev_v_simulated_slices[d]) * 4; break;
    case 2: src = r->dev_v_simulated_weights[d].data; bytes = vcount(r->dev_v_simulated_weights[d]) * 4; break;
    case 3: src = r->dev_v_simulated_inside[d].data; bytes = vcount(r->dev_v_simulated_inside[d]); break;
    case 4: src = r->dev_v_PSF_sums_[d].data; bytes = vcount(r->dev_v_PSF_sums_[d]) * 4; break;
    case 5: src = r->dev_sliceVoxel_count_[d].data; bytes = vcount(r->dev_sliceVoxel_count_[d]) * 4; break;
    case 6: src = r->dev_addon_[d].data; bytes = vcount(r->dev_addon_[d]) * 4; break;
    case 7: src = r->dev_confidence_map_[d].data; bytes = vcount(r->dev_confidence_map_[d]) * 4; break;
    case 8: src = r->dev_v_slices[d].data; bytes = vcount(r->dev_v_slices[d]) * 4; break;
    case 9: src = r->dev_volume_weights_[d].data; bytes = vcount(r->dev_volume_weights_[d]) * 4; break;
    case 10: src = r->dev_reconstructed_[d].data; bytes = vcount(r->dev_reconstructed_[d]) * 4; break;
    default: return -1;
  }
  return cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
/* number of slices the reference actually stores on device 0 (Q6: it drops the last one) */
int ref_device_slices(void* hv) {
  Reconstruction* r = ((RefHandle*)hv)->r;
  return (int)r->dev_v_slices[r->devicesToUse[0]].size.z;
}

/* registration (cuda2.cu:3800-4141) */
void ref_reg_init_storage(void* hv, int W, int H, int S, float dx, float dy, float dz) {
  ((RefHandle*)hv)->r->initRegStorageVolumes(make_uint3(W, H, S), make_float3(dx, dy, dz));
}
void ref_reg_fill_slices(void* hv, float* cube, const float* i2w) {
  RefHandle* h = (RefHandle*)hv;
  h->r->FillRegSlices(cube, m4v(i2w, h->S));
}
void ref_reg_update_slices_i2w(void* hv, const float* ofs) {
  RefHandle* h = (RefHandle*)hv;
  h->r->updateResampledSlicesI2W(m4v(ofs, h->S));
}
void ref_reg_prepare(void* hv) { ((RefHandle*)hv)->r->prepareSliceToVolumeReg(); }
void ref_reg_set_schedule(void* hv, int levels, int steps, int iterations) {
  /* prepareSliceToVolumeReg() sets 2/4/20 (cuda2.cu:3884-3886); tests may shorten the schedule afterwards */
  Reconstruction* r = ((RefHandle*)hv)->r;
  if (levels > 0 && levels <= r->_NumberOfLevels) r->_NumberOfLevels = levels;
  if (steps > 0) r->_NumberOfSteps = steps;
  if (iterations > 0) r->_NumberOfIterations = iterations;
}
void ref_reg_register(void* hv, float* T) {
  RefHandle* h = (RefHandle*)hv;
  std::vector<Matrix4> v = m4v(T, h->S);
  h->r->registerSlicesToVolume(v);
  for (int i = 0; i < h->S; ++i) m4out(v[i], T + 16 * i);
}
/* One cost evaluation for given transforms at a given level: the prologue of registerMultipleSlicesToVolume
 * (cuda2.cu:4001-4050: upload matrices, blur the target slices, activate all slices) followed by the reference's
 * evaluateCostsMultipleSlices (cuda2.cu:4149); returns the NCC sum over the three in-slice z-offsets per slice. */
void ref_reg_evaluate(void* hv, const float* T, int level, float* similarity) {
  RefHandle* h = (RefHandle*)hv;
  Reconstruction* r = h->r;
  const int dev = r->devicesToUse[0];
  const int n = (int)r->dev_v_slices_resampled_float[dev].size.z;
  std::vector<Matrix4> v = m4v(T, h->S);
  checkCudaErrors(cudaMemcpy(r->dev_recon_matrices[dev], &v[0], sizeof(Matrix4) * n, cudaMemcpyHostToDevice));
  checkCudaErrors(cudaMemcpy(r->dev_recon_matrices_orig[dev], &v[0], sizeof(Matrix4) * n, cudaMemcpyHostToDevice));
  const float blur = r->_Blurring[level];
  r->dev_v_slices_resampled_float[dev].copyFromOther(r->dev_v_slices_resampled[dev]);
  FilterGaussStack(r->dev_v_slices_resampled_float[dev].surface, r->dev_v_slices_resampled_float[dev].surface,
                   r->dev_temp_slices[dev].surface, r->dev_regSlices[dev].size.x, r->dev_regSlices[dev].size.y, n, blur);
  initActiveSlices<<<divup(n, 512), 512>>>(r->dev_active_slices[dev], n);
  r->evaluateCostsMultipleSlices(n, n, level, blur, 0, 1, 3, dev);
  checkCudaErrors(cudaDeviceSynchronize());
  checkCudaErrors(cudaMemcpy(similarity, r->dev_recon_similarities[dev], sizeof(float) * n, cudaMemcpyDeviceToHost));
}
/* the last batch of resampled+blurred volume slices (dev_regSlices) and blurred targets, for diffing R1/R2 */
int ref_reg_get(void* hv, int kind, float* out) {
  Reconstruction* r = ((RefHandle*)hv)->r;
  const int dev = r->devicesToUse[0];
  cudaDeviceSynchronize();
  if (kind == 0) r->dev_regSlices[dev].copyToHost(out);
  else if (kind == 1) r->dev_v_slices_resampled_float[dev].copyToHost(out);
  else if (kind == 2) r->dev_v_slices_resampled[dev].copyToHost(out);
  else return -1;
  return 0;
}
}  // extern "C"
