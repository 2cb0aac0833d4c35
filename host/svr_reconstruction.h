// svr_reconstruction.h -- host orchestration of the SVR GPU path over the C ABI of libsvr_b200.so.
// Mirrors the members of class irtkReconstruction that SVRreconstructionGPU's main uses on the GPU path
// (source/reconstructionGPU2/irtkReconstructionGPU.cc; method names and call order are the reference's):
//   CreateTemplate :648-694, SetMask :750-803, TransformMask :805-822, CropImage :5205-5306,
//   MatchStackIntensitiesWithMasking :1375-1493, CreateSlicesAndTransformations :1814-1850, MaskSlices :1940-1988,
//   SyncGPU :249-328, UpdateGPUTranformationMatrices :372-401, generatePSFVolume :1496-1610 (constants only),
//   InitializeEMGPU / InitializeEMValuesGPU :2905-2953, GaussianReconstructionGPU :2695-2762, SimulateSlicesGPU :1163-1203,
//   InitializeRobustStatisticsGPU :2988-3020, EStepGPU :3162-3440, ScaleGPU :3751-3765, SuperresolutionGPU :4024-4053,
//   MStepGPU :4214-4224, MaskVolumeGPU, ScaleVolumeGPU, RestoreSliceIntensitiesGPU :1003-1032, SyncCPU,
//   PrepareRegistrationSlices :2105-2179, SliceToVolumeRegistrationGPU :2218-2288, SaveTransformations :4884-4919,
//   ReadTransformation :4733-4765, EvaluateGPU :4503-4538.
// Not restated: the IRTK CPU registrations (StackRegistrations, the default CPU
// slice-to-volume registration) -- stacks enter with the transformations given by -t (or identity) and slices are
// registered with the GPU registration (--useGPUReg behaviour).
#pragma once
#include <ostream>
#include <string>
#include <vector>
#include "../include/svr_abi.h"
#include "svr_image.h"

namespace svr {

// One rank = one GPU = one svr_context holding a share of the slices and a full replica of the volume
// (replaces class GPUWorker + the per-device vectors of class Reconstruction, GPUWorker.cpp:104-257, .cuh:103-139).
struct Rank {
    int device = 0;
    svr_context* c = nullptr;
    std::vector<int> idx;            // global indices of this rank's slices (svr_host_partition_strided)
    void* stream = nullptr;          // the context's cudaStream_t: the NCCL all-reduce is enqueued here
    void* comm = nullptr;            // ncclComm_t
};

class Reconstruction {
public:
    explicit Reconstruction(int device);
    // -d d0 d1 ...: one rank per device, driven by one host thread each; the volume accumulators are summed with
    // ncclAllReduce over NVLink (replaces the reduce-to-device-0 + broadcast of cuda2.cu:2225-2239,2445-2460,2175-2180).
    explicit Reconstruction(const std::vector<int>& devices);
    ~Reconstruction();
    size_t NumberOfRanks() const { return devices_.size(); }

    double CreateTemplate(const Image& stack, double resolution);
    void SetMask(Image* mask, double sigma, double threshold = 0.5);
    void TransformMask(const Image& image, Image& mask, const Rigid& transformation);
    void CropImage(Image& image, const Image& mask);
    void MatchStackIntensitiesWithMasking(std::vector<Image>& stacks, const std::vector<Rigid>& t, double averageValue, bool together = false);
    void CreateSlicesAndTransformations(const std::vector<Image>& stacks, const std::vector<Rigid>& t, const std::vector<double>& thickness);
    void MaskSlices();
    void SetForceExcludedSlices(const std::vector<int>& f) { force_excluded_ = f; }
    void SetSmoothingParameters(double delta, double lambda);
    void ReadTransformation(const std::string& folder);
    void SaveTransformations(const std::string& prefix = "");

    void SyncGPU();
    void UpdateGPUTranformationMatrices();
    void generatePSFVolume();
    void InitializeEMGPU();
    void InitializeEMValuesGPU();
    void GaussianReconstructionGPU();
    void SimulateSlicesGPU();
    void InitializeRobustStatisticsGPU();
    void EStepGPU();
    void ScaleGPU();
    void SuperresolutionGPU(int iter);
    void MStepGPU(int iter);
    void MaskVolumeGPU();
    void ScaleVolumeGPU();
    void RestoreSliceIntensitiesGPU();
    void SyncCPU();
    void PrepareRegistrationSlices();
    void SliceToVolumeRegistrationGPU();
    // The reference's DEFAULT registrations (IRTK's irtkImageRigidRegistrationWithPadding on the CPU through TBB), on the device
    // engine svr_rreg_register: StackRegistrations irtkReconstructionGPU.cc:940-1001, SliceToVolumeRegistration :1992-2059, :2291-2303
    void StackRegistrations(std::vector<Image>& stacks, std::vector<Rigid>& stack_transformations, int templateNumber);
    void SliceToVolumeRegistration();
    void InvertStackTransformations(std::vector<Rigid>& t) const { for (Rigid& r : t) r.invert(); }
    void EvaluateGPU(int iter, std::ostream& os);

    const Image& GetReconstructedGPU() const { return reconstructed_; }
    const Image& GetMask() const { return mask_; }
    size_t NumberOfSlices() const { return slices_.size(); }
    double device_ms(int kind, long long* launches) const;
    void profile(bool on);
    void DumpSetup(const std::string& dir) const;     // the inputs SyncGPU would upload, as raw arrays (tools/c2_parity.py)
    bool debug = false;

private:
    void ck(int rc, const char* what) const;
    void ensure_context();
    void allreduce_accumulator(Rank& r);
    template <class F> void each_rank(F&& f);
    template <class T> std::vector<T> take(const Rank& r, const std::vector<T>& global, int stride = 1) const;
    template <class T> void put(const Rank& r, const std::vector<T>& local, std::vector<T>& global, int stride = 1) const;
    std::vector<int> devices_;
    std::vector<Rank> ranks_;
    svr_context* c_ = nullptr;       // rank 0's context (single-rank calls: profile read-outs, registration debug taps)
    Image reconstructed_, mask_;
    bool template_created_ = false, have_mask_ = false;
    std::vector<Image> slices_;                    // single-plane images
    std::vector<Rigid> transformations_;
    std::vector<int> stack_index_;
    std::vector<float> stack_factor_;
    std::vector<int> force_excluded_, small_slices_;
    std::vector<float> scale_, slice_weight_, slice_potential_;
    std::vector<unsigned char> slice_inside_;
    int Nx_ = 0, Ny_ = 0;
    // irtkReconstruction members (irtkReconstructionGPU.cc:159-187)
    double step_ = 0.0001, delta_ = 1, lambda_ = 0.1, alpha_ = 0.5, average_value_ = 700;
    float sigma_ = 0, mix_ = 0.9f, m_ = 0;
    float state5_[5] = { 0.025f, 0.9f, 0, 0, 0 };   // sigma_s, mix_s, mean_s, mean_s2, sigma_s2
    double min_intensity_ = 0, max_intensity_ = 0;
    bool adaptive_ = false;
    // registration front-end
    std::vector<ImageAttr> res_attrs_;
    int regW_ = 0, regH_ = 0;
    bool reg_prepared_ = false;
};

}  // namespace svr
