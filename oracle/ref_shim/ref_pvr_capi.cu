/* TEST INFRASTRUCTURE (oracle/): C entry points around the UNMODIFIED reference PVR CUDA path -- ReconVolume<float>,
 * PatchBasedVolume<float>, initPatchBasedRecon_gpu's constant upload, patchBasedPSFReconstruction_gpu,
 * patchBasedSimulatePatches_gpu, patchBasedSuperresolution_gpu<float>::{run,regularize},
 * patchBasedRobustStatistics_gpu<float>::{initializeEMValues,InitializeRobustStatistics,EStep,MStep,Scale}
 * (source/reconstructionGPU2/*.cu, include/*.cuh) -- in the call order of irtkPatchBasedReconstruction<T>::run()
 * (irtkPatchBasedReconstruction.cpp:445-560).  Built into oracle/_ref/libref_pvr.so (`make ref`).
 *
 * The reference fills a PatchBasedVolume from an IRTK stack (PatchBasedVolume::init -> generate2DPatches, host code on
 * irtkGenericImage / irtkRigidTransformation, not available here).  This harness fills the same members itself from the
 * patch list our own enumeration produced (patch matrices + the patch-value cube) and then only calls reference code.
 * Included at the end of ref_pvr_tu.cu (same translation unit as the reference sources).
 * No arithmetic of the path. */
#include <cstring>
#include <vector>


/* runStackSLIC<T>::segmentSLIC (runStackSLIC.cpp, superpixel segmentation on IRTK images) is referenced by the virtual
 * superpixel patch generator of PatchBasedObject; the harness never reaches it (patches arrive enumerated). */
template <typename T>
void runStackSLIC<T>::segmentSLIC(irtkGenericImage<T>&, irtkGenericImage<T>&, unsigned int&, unsigned int&, unsigned int, bool) {
  fprintf(stderr, "[ref] runStackSLIC::segmentSLIC is not part of the parity harness\n");
  abort();
}
template class runStackSLIC<float>;
template class runStackSLIC<double>;

namespace {

struct RefPvr {
  int dev = 0;
  ReconVolume<float> recon;
  std::vector<PatchBasedVolume<float> > stacks;
  patchBasedSuperresolution_gpu<float>* sr = nullptr;
  patchBasedRobustStatistics_gpu<float>* rs = nullptr;
  bool spx = false;
  size_t V = 0;
};

Matrix4<float> pm4(const float* p) {
  Matrix4<float> m;
  for (int i = 0; i < 4; ++i) { m.data[i].x = p[4 * i]; m.data[i].y = p[4 * i + 1]; m.data[i].z = p[4 * i + 2]; m.data[i].w = p[4 * i + 3]; }
  return m;
}
template <class T> T* dalloc(size_t n) {
  T* p = nullptr;
  checkCudaErrors(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)));
  checkCudaErrors(cudaMemset(p, 0, (n ? n : 1) * sizeof(T)));
  return p;
}
size_t np_of(PatchBasedVolume<float>& s) { return (size_t)s.m_XYZPatchGridSize.x * s.m_XYZPatchGridSize.y * s.m_XYZPatchGridSize.z; }

}  // namespace

extern "C" {

/* ReconVolume<T>::init + setMask (reconVolume.cuh:46-114) */
void* refpvr_create(int dev, int vx, int vy, int vz, float dx, float dy, float dz, const float* w2i, const float* i2w, const char* mask) {
  RefPvr* h = new RefPvr;
  h->dev = dev;
  h->V = (size_t)vx * vy * vz;
  h->recon.init(dev, make_uint3(vx, vy, vz), make_float3(dx, dy, dz), pm4(w2i), pm4(i2w));
  h->recon.setMask(const_cast<char*>(mask));
  return h;
}

/* One PatchBasedVolume per stack: the members PatchBasedVolume::init / generate2DPatches would set
 * (patchBasedVolume.cuh:108-195, patchBasedObject.cuh:176-342), filled from the caller's patch list. */
int refpvr_add_stack(void* hv, int pbx, int pby, int n, float sdx, float sdy, float sdz, float thickness, const float* patches,
                     const float* I2W, const float* W2I, const float* T, const float* Tinv, const char* spx) {
  RefPvr* h = (RefPvr*)hv;
  PatchBasedVolume<float> s;
  s.m_size = make_uint3(pbx, pby, n);              /* only the patch buffer is used: the stack itself is never read */
  s.m_dim = make_float3(sdx, sdy, sdz);
  s.m_pbbsize = make_uint2(pbx, pby);
  s.m_stride = make_uint2(pbx, pby);
  s.m_thickness = thickness;
  s.m_numPatches = n;
  s.m_XYZPatchGridSize = make_uint3(pbx, pby, n);
  const size_t N = (size_t)pbx * pby * n;
  s.numDemonElems = (int)N;
  std::vector<ImagePatch2D<float> > hp(n ? n : 1);
  for (int i = 0; i < n; ++i) {
    ImagePatch2D<float>& p = hp[i];
    p.I2W = pm4(I2W + 16 * i); p.W2I = pm4(W2I + 16 * i);
    p.Transformation = pm4(T + 16 * i); p.InvTransformation = pm4(Tinv + 16 * i);
    p.RI2W = p.I2W; p.Mo = pm4(I2W + 16 * i); p.InvMo = p.Mo;   /* registration scratch, unused by the reconstruction */
    p.scale = 1.0f; p.patchWeight = 1.0f;
    if (spx) memcpy(p.spxMask, spx + (size_t)4096 * i, 4096); else memset(p.spxMask, '1', 4096);
  }
  s.d_patches = dalloc<ImagePatch2D<float> >(n);
  checkCudaErrors(cudaMemcpy(s.d_patches, hp.data(), sizeof(ImagePatch2D<float>) * n, cudaMemcpyHostToDevice));
  s.m_d_data = dalloc<float>(N);
  s.d_m_weightsPtr = dalloc<float>(N);
  s.d_m_simulated_weightsPtr = dalloc<float>(N);
  s.d_m_simulated_patchesPtr = dalloc<float>(N);
  s.d_m_bufferPtr = dalloc<float>(N);
  s.d_m_simulated_insidePtr = dalloc<char>(N);
  s.d_m_patchVoxel_count_Ptr = dalloc<int>(N);
  s.d_m_PSF_sums_Ptr = dalloc<float>(N);
  s.d_m_regPatchesPtr = dalloc<float>(N);
  s.d_m_PatchesPtr = dalloc<float>(N);
  checkCudaErrors(cudaMemcpy(s.d_m_PatchesPtr, patches, sizeof(float) * N, cudaMemcpyHostToDevice));
  h->stacks.push_back(s);
  return (int)h->stacks.size() - 1;
}

/* the constant upload of initPatchBasedRecon_gpu (initPatchBasedRecon_gpu.cu:92) with the PointSpreadFunction the host
 * builds in irtkPatchBasedReconstruction.cpp (PSF image attributes of the volume grid) */
void refpvr_set_psf(void* hv, int sx, int sy, int sz, float pdx, float pdy, float pdz, const float* psf_i2w, const float* psf_w2i, float q) {
  RefPvr* h = (RefPvr*)hv;
  checkCudaErrors(cudaSetDevice(h->dev));
  PointSpreadFunction<float> p;
  p.m_PSFdim = make_float3(pdx, pdy, pdz);
  p.m_PSFsize = make_uint3(sx, sy, sz);
  p.m_PSFI2W = pm4(psf_i2w);
  p.m_PSFW2I = pm4(psf_w2i);
  p.m_quality_factor = q;
  checkCudaErrors(cudaMemcpyToSymbol(_PSF, &p, sizeof(PointSpreadFunction<float>)));
}

/* patchBasedSuperresolution_gpu / patchBasedRobustStatistics_gpu objects (irtkPatchBasedReconstruction.cpp:415-428) */
void refpvr_begin(void* hv, float min_intensity, float max_intensity, int adaptive, int use_spx) {
  RefPvr* h = (RefPvr*)hv;
  h->spx = use_spx != 0;
  h->sr = new patchBasedSuperresolution_gpu<float>(min_intensity, max_intensity, adaptive != 0);
  h->rs = new patchBasedRobustStatistics_gpu<float>(h->stacks);
}

void refpvr_rs_initialize_em_values(void* hv) { ((RefPvr*)hv)->rs->initializeEMValues(); }
void refpvr_recon_reset(void* hv) { ((RefPvr*)hv)->recon.reset(); }
void refpvr_recon_reset_addon_cmap(void* hv) { ((RefPvr*)hv)->recon.resetAddonCmap(); }
void refpvr_recon_equalize(void* hv) { ((RefPvr*)hv)->recon.equalize(); }
void refpvr_psf_reconstruction(void* hv) {
  RefPvr* h = (RefPvr*)hv;
  for (size_t i = 0; i < h->stacks.size(); ++i) patchBasedPSFReconstruction_gpu<float>(h->dev, h->stacks[i], h->recon, h->spx);
}
void refpvr_simulate_patches(void* hv) {
  RefPvr* h = (RefPvr*)hv;
  for (size_t i = 0; i < h->stacks.size(); ++i) patchBasedSimulatePatches_gpu<float>(h->dev, h->stacks[i], h->recon);
}
void refpvr_rs_initialize_robust_statistics(void* hv, float min_intensity, float max_intensity) {
  RefPvr* h = (RefPvr*)hv;
  h->rs->InitializeRobustStatistics(min_intensity, max_intensity, h->dev);
}
void refpvr_rs_estep(void* hv) { ((RefPvr*)hv)->rs->EStep(); }
void refpvr_rs_mstep(void* hv, int iter) { ((RefPvr*)hv)->rs->MStep(iter); }
void refpvr_rs_scale(void* hv) { ((RefPvr*)hv)->rs->Scale(); }
void refpvr_superresolution_run(void* hv) {
  RefPvr* h = (RefPvr*)hv;
  for (size_t i = 0; i < h->stacks.size(); ++i) h->sr->run(h->dev, &h->stacks[i], &h->recon);
}
void refpvr_superresolution_regularize(void* hv) { RefPvr* h = (RefPvr*)hv; h->sr->regularize(h->dev, &h->recon); }

/* read-backs.  volume kinds: 0 reconstruction, 1 volume weights, 2 addon, 3 confidence map */
int refpvr_get_volume(void* hv, int kind, float* out) {
  RefPvr* h = (RefPvr*)hv;
  cudaDeviceSynchronize();
  const float* src = kind == 0 ? h->recon.getDataPtr() : kind == 1 ? h->recon.getReconstructed_volWeigthsPtr()
                   : kind == 2 ? h->recon.getAddonPtr() : kind == 3 ? h->recon.getCMapPtr() : nullptr;
  if (!src) return -1;
  return cudaMemcpy(out, src, h->V * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
/* per-stack patch buffers: 0 weights, 1 simulated patches, 2 simulated weights, 3 simulated inside (char), 4 PSF sums,
 * 5 patch values, 6 patch voxel count (int) */
int refpvr_get_patch_buffer(void* hv, int stack, int kind, void* out) {
  RefPvr* h = (RefPvr*)hv;
  cudaDeviceSynchronize();
  PatchBasedVolume<float>& s = h->stacks[stack];
  const size_t N = np_of(s);
  const void* src = nullptr; size_t es = 4;
  switch (kind) {
    case 0: src = s.d_m_weightsPtr; break;
    case 1: src = s.d_m_simulated_patchesPtr; break;
    case 2: src = s.d_m_simulated_weightsPtr; break;
    case 3: src = s.d_m_simulated_insidePtr; es = 1; break;
    case 4: src = s.d_m_PSF_sums_Ptr; break;
    case 5: src = s.d_m_PatchesPtr; break;
    case 6: src = s.d_m_patchVoxel_count_Ptr; break;
    default: return -1;
  }
  return cudaMemcpy(out, src, N * es, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
/* per-patch EM outputs kept in ImagePatch2D: scale[n], patchWeight[n] */
int refpvr_get_patch_scales_weights(void* hv, int stack, float* scale, float* weight) {
  RefPvr* h = (RefPvr*)hv;
  cudaDeviceSynchronize();
  PatchBasedVolume<float>& s = h->stacks[stack];
  std::vector<ImagePatch2D<float> > hp(s.m_numPatches ? s.m_numPatches : 1);
  if (cudaMemcpy(hp.data(), s.d_patches, sizeof(ImagePatch2D<float>) * s.m_numPatches, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  for (unsigned i = 0; i < s.m_numPatches; ++i) { scale[i] = hp[i].scale; weight[i] = hp[i].patchWeight; }
  return 0;
}
/* {sigma, mix, m, sigma_s, mix_s, mean_s, mean_s2, sigma_s2} of patchBasedRobustStatistics_gpu */
void refpvr_get_em_state(void* hv, float* out8) {
  patchBasedRobustStatistics_gpu<float>* r = ((RefPvr*)hv)->rs;
  out8[0] = r->m_sigma_gpu; out8[1] = r->m_mix_gpu; out8[2] = r->m_m_gpu; out8[3] = r->m_sigma_s_gpu; out8[4] = r->m_mix_s_gpu;
  out8[5] = r->m_mean_s_gpu; out8[6] = r->m_mean_s2_gpu; out8[7] = r->m_sigma_s2_gpu;
}

}  /* extern "C" */
