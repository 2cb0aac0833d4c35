// svr_reconstruction.cc -- host orchestration of the SVR GPU path over the C ABI (see svr_reconstruction.h for the
// reference lines each member restates).  No arithmetic of the hot path lives here: every device step is one svr_* call.
#include "svr_reconstruction.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <thread>
#ifdef SVR_WITH_NCCL
#include <nccl.h>
#endif

namespace svr {

Reconstruction::Reconstruction(int device) : devices_(1, device) {}
Reconstruction::Reconstruction(const std::vector<int>& devices) : devices_(devices.empty() ? std::vector<int>(1, 0) : devices) {}

// The device contexts are created on the first device call (SyncGPU): the stack/mask/slice set-up above it is host work
// and can be run (and tested) on a machine without a GPU; everything from SyncGPU on fails loudly without one.
void Reconstruction::ensure_context()
{
    if (!ranks_.empty()) return;
    ranks_.resize(devices_.size());
    for (size_t r = 0; r < devices_.size(); ++r) {
        ranks_[r].device = devices_[r];
        if (svr_create(&ranks_[r].c, devices_[r]) != 0) throw std::runtime_error(std::string("svr_create: ") + svr_last_error(nullptr));
        if (svr_get_stream(ranks_[r].c, &ranks_[r].stream) != 0) throw std::runtime_error("svr_get_stream failed");
    }
    c_ = ranks_[0].c;
    if (ranks_.size() > 1) {
#ifdef SVR_WITH_NCCL
        std::vector<ncclComm_t> comms(ranks_.size());
        const ncclResult_t rc = ncclCommInitAll(comms.data(), (int)ranks_.size(), devices_.data());
        if (rc != ncclSuccess) throw std::runtime_error(std::string("ncclCommInitAll: ") + ncclGetErrorString(rc));
        for (size_t r = 0; r < ranks_.size(); ++r) ranks_[r].comm = comms[r];
        // between *_local and *_finish the all-reduce is enqueued on the context's stream: no host round trip is needed
        for (Rank& r : ranks_) svr_set_async(r.c, 1);
        std::cout << "NCCL: " << ranks_.size() << " ranks (one per GPU), all-reduce of the volume accumulator over NVLink" << std::endl;
#else
        throw std::runtime_error("more than one device requested, but this binary was built without NCCL (nccl.h not found at build time)");
#endif
    }
}

Reconstruction::~Reconstruction()
{
    for (Rank& r : ranks_) {
#ifdef SVR_WITH_NCCL
        if (r.comm) ncclCommDestroy((ncclComm_t)r.comm);
#endif
        if (r.c) svr_destroy(r.c);
    }
}

void Reconstruction::ck(int rc, const char* what) const
{
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + svr_last_error(c_));
}

// Runs f(rank) for every rank: inline for one rank, one host thread per rank otherwise (each makes its device current: the
// library calls carry a device guard of their own, the NCCL call next to them needs the right current device).
template <class F>
void Reconstruction::each_rank(F&& f)
{
    if (ranks_.size() == 1) { f(ranks_[0]); return; }
    std::vector<std::thread> th;
    std::vector<std::string> err(ranks_.size());
    for (size_t i = 0; i < ranks_.size(); ++i)
        th.emplace_back([&, i] {
            try {
                if (svr_make_current(ranks_[i].c) != 0) throw std::runtime_error(svr_last_error(ranks_[i].c));
                f(ranks_[i]);
            } catch (const std::exception& e) { err[i] = e.what()[0] ? e.what() : "error"; }
        });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < err.size(); ++i)
        if (!err[i].empty()) throw std::runtime_error("rank " + std::to_string(i) + " (device " + std::to_string(ranks_[i].device) + "): " + err[i]);
}

// this rank's share of a per-slice vector (stride values per slice) / the way back
template <class T>
std::vector<T> Reconstruction::take(const Rank& r, const std::vector<T>& global, int stride) const
{
    std::vector<T> out(r.idx.size() * (size_t)stride);
    for (size_t j = 0; j < r.idx.size(); ++j)
        for (int q = 0; q < stride; ++q) out[j * stride + q] = global[(size_t)r.idx[j] * stride + q];
    return out;
}
template <class T>
void Reconstruction::put(const Rank& r, const std::vector<T>& local, std::vector<T>& global, int stride) const
{
    for (size_t j = 0; j < r.idx.size(); ++j)
        for (int q = 0; q < stride; ++q) global[(size_t)r.idx[j] * stride + q] = local[j * stride + q];
}

// Sum of the interleaved {numerator, denominator} accumulator over the ranks, in place, on the context's stream.
void Reconstruction::allreduce_accumulator(Rank& r)
{
    if (ranks_.size() == 1) return;
#ifdef SVR_WITH_NCCL
    void* buf = nullptr; size_t bytes = 0;
    if (svr_device_buffer(r.c, SVR_BUF_ACCUMULATOR, &buf, &bytes) != 0) throw std::runtime_error(svr_last_error(r.c));
    const ncclResult_t rc = ncclAllReduce(buf, buf, bytes / sizeof(float), ncclFloat, ncclSum, (ncclComm_t)r.comm, (cudaStream_t)r.stream);
    if (rc != ncclSuccess) throw std::runtime_error(std::string("ncclAllReduce: ") + ncclGetErrorString(rc));
#else
    (void)r;
#endif
}

double Reconstruction::device_ms(int kind, long long* launches) const
{
    double ms = 0; int64_t n = 0;
    svr_profile_read(c_, kind, &ms, &n);
    if (launches) *launches = n;
    return ms;
}

void Reconstruction::profile(bool on) { ensure_context(); for (Rank& r : ranks_) svr_profile_enable(r.c, on ? 1 : 0); }

// Writes what SyncGPU / UpdateGPUTranformationMatrices would upload, as raw little-endian arrays plus a text index, so
// the same inputs can be fed to other backends (tools/c2_parity.py drives the reference's CUDA path with them).
void Reconstruction::DumpSetup(const std::string& dir) const
{
    auto wr = [&](const std::string& name, const void* p, size_t bytes) {
        FILE* f = std::fopen((dir + "/" + name).c_str(), "wb");
        if (!f || std::fwrite(p, 1, bytes, f) != bytes) throw std::runtime_error("cannot write " + dir + "/" + name);
        std::fclose(f);
    };
    const int S = (int)slices_.size();
    int Nx = 0, Ny = 0;
    for (const Image& s : slices_) { Nx = std::max(Nx, s.a.x); Ny = std::max(Ny, s.a.y); }
    std::vector<float> cube((size_t)Nx * Ny * S, -1.0f), dims(3 * (size_t)S), T(16 * (size_t)S), Ti(16 * (size_t)S), I2W(16 * (size_t)S), W2I(16 * (size_t)S);
    std::vector<int> sizes(2 * (size_t)S);
    std::vector<double> attrs(18 * (size_t)S);
    for (int n = 0; n < S; ++n) {
        const Image& s = slices_[n];
        for (int y = 0; y < s.a.y; ++y) for (int x = 0; x < s.a.x; ++x) cube[((size_t)n * Ny + y) * Nx + x] = (float)s.at(x, y, 0);
        sizes[2 * n] = s.a.x; sizes[2 * n + 1] = s.a.y;
        dims[3 * n] = (float)s.a.dx; dims[3 * n + 1] = (float)s.a.dy; dims[3 * n + 2] = (float)s.a.dz;
        const Mat4 t = transformations_[n].matrix();
        t.to_float16(&T[16 * n]); t.inverse().to_float16(&Ti[16 * n]);
        s.a.image_to_world().to_float16(&I2W[16 * n]); s.a.world_to_image().to_float16(&W2I[16 * n]);
        double* a = &attrs[18 * n];
        a[0] = s.a.x; a[1] = s.a.y; a[2] = s.a.z; a[3] = s.a.dx; a[4] = s.a.dy; a[5] = s.a.dz;
        for (int q = 0; q < 3; ++q) { a[6 + q] = s.a.origin[q]; a[9 + q] = s.a.xaxis[q]; a[12 + q] = s.a.yaxis[q]; a[15 + q] = s.a.zaxis[q]; }
    }
    std::vector<float> mask(mask_.n());
    for (size_t i = 0; i < mask.size(); ++i) mask[i] = (float)mask_.v[i];
    float ri2w[16], rw2i[16];
    reconstructed_.a.image_to_world().to_float16(ri2w);
    reconstructed_.a.world_to_image().to_float16(rw2i);
    wr("slices.f32", cube.data(), cube.size() * 4); wr("sizes.i32", sizes.data(), sizes.size() * 4); wr("dims.f32", dims.data(), dims.size() * 4);
    wr("T.f32", T.data(), T.size() * 4); wr("Tinv.f32", Ti.data(), Ti.size() * 4); wr("I2W.f32", I2W.data(), I2W.size() * 4); wr("W2I.f32", W2I.data(), W2I.size() * 4);
    wr("slice_attrs.f64", attrs.data(), attrs.size() * 8);
    wr("mask.f32", mask.data(), mask.size() * 4); wr("recon_i2w.f32", ri2w, 64); wr("recon_w2i.f32", rw2i, 64);
    wr("stack_index.i32", stack_index_.data(), stack_index_.size() * 4); wr("stack_factor.f32", stack_factor_.data(), stack_factor_.size() * 4);
    FILE* f = std::fopen((dir + "/index.txt").c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + dir + "/index.txt");
    std::fprintf(f, "S %d\nNx %d\nNy %d\nvx %d\nvy %d\nvz %d\nvoxel %.9g\nstacks %d\n", S, Nx, Ny, reconstructed_.a.x, reconstructed_.a.y,
                 reconstructed_.a.z, reconstructed_.a.dx, (int)stack_factor_.size());
    std::fclose(f);
}

// ---- irtkReconstruction::CreateTemplate, irtkReconstructionGPU.cc:648-694 -----------------------------------------
double Reconstruction::CreateTemplate(const Image& stack, double resolution)
{
    ImageAttr attr = stack.a;
    attr.z += 2;                                  // enlarge in z in case the top of the head is cut off
    double d;
    if (resolution <= 0) {
        const double dx = stack.a.dx, dy = stack.a.dy, dz = stack.a.dz;
        if (dx <= dy && dx <= dz) d = dx; else if (dy <= dz) d = dy; else d = dz;
    } else d = resolution;
    std::cout << "Constructing volume with isotropic voxel size " << d << std::endl;
    reconstructed_ = Image(resampled_attr(attr, d, d, d), 0.0);   // resampling of an empty image: only the grid matters
    template_created_ = true;
    return d;
}

// ---- SetMask, irtkReconstructionGPU.cc:750-803 ----------------------------------------------------------------------
void Reconstruction::SetMask(Image* mask, double sigma, double threshold)
{
    if (!template_created_) throw std::runtime_error("Please create the template before setting the mask");
    mask_ = reconstructed_;
    if (mask) {
        if (sigma > 0) {
            mask->gaussian_blur(sigma);
            for (double& v : mask->v) v = v > threshold ? 1 : 0;
        }
        Rigid id;
        transform_image_nn(*mask, id, mask_, -1, 0);
    } else {
        std::fill(mask_.v.begin(), mask_.v.end(), 1.0);
    }
    have_mask_ = true;
}

// ---- TransformMask, irtkReconstructionGPU.cc:805-822 ---------------------------------------------------------------
void Reconstruction::TransformMask(const Image& image, Image& mask, const Rigid& transformation)
{
    Image m = image;
    transform_image_nn(mask, transformation, m, -1, 0);
    mask = m;
}

// ---- CropImage, irtkReconstructionGPU.cc:5205-5306 -----------------------------------------------------------------
void Reconstruction::CropImage(Image& image, const Image& mask)
{
    const int X = image.a.x, Y = image.a.y, Z = image.a.z;
    auto plane_z = [&](int k) { for (int j = 0; j < Y; ++j) for (int i = 0; i < X; ++i) if (mask.at(i, j, k) > 0) return true; return false; };
    auto plane_y = [&](int j) { for (int k = 0; k < Z; ++k) for (int i = 0; i < X; ++i) if (mask.at(i, j, k) > 0) return true; return false; };
    auto plane_x = [&](int i) { for (int k = 0; k < Z; ++k) for (int j = 0; j < Y; ++j) if (mask.at(i, j, k) > 0) return true; return false; };
    int x1, x2, y1, y2, z1, z2;
    for (z2 = Z - 1; z2 >= 0 && !plane_z(z2); --z2) {}
    for (z1 = 0; z1 <= Z - 1 && !plane_z(z1); ++z1) {}
    for (y2 = Y - 1; y2 >= 0 && !plane_y(y2); --y2) {}
    for (y1 = 0; y1 <= Y - 1 && !plane_y(y1); ++y1) {}
    for (x2 = X - 1; x2 >= 0 && !plane_x(x2); --x2) {}
    for (x1 = 0; x1 <= X - 1 && !plane_x(x1); ++x1) {}
    if (debug) std::cout << "Region of interest is " << x1 << " " << y1 << " " << z1 << " " << x2 << " " << y2 << " " << z2 << std::endl;
    if (x2 < x1 || y2 < y1 || z2 < z1) throw std::runtime_error("CropImage: the mask does not overlap the image");
    image = image.get_region(x1, y1, z1, x2 + 1, y2 + 1, z2 + 1);
}

// ---- MatchStackIntensitiesWithMasking, irtkReconstructionGPU.cc:1375-1493 -------------------------------------------
void Reconstruction::MatchStackIntensitiesWithMasking(std::vector<Image>& stacks, const std::vector<Rigid>& t, double averageValue, bool together)
{
    average_value_ = averageValue;
    std::vector<double> stack_average;
    const Mat4 mw2i = mask_.a.world_to_image();
    for (size_t ind = 0; ind < stacks.size(); ++ind) {
        const Image& s = stacks[ind];
        const Mat4 m = mw2i * (t[ind].matrix() * s.a.image_to_world());
        double sum = 0, num = 0;
        for (int i = 0; i < s.a.x; ++i) for (int j = 0; j < s.a.y; ++j) for (int k = 0; k < s.a.z; ++k) {
            double x = i, y = j, z = k;
            m.apply(x, y, z);
            x = std::round(x); y = std::round(y); z = std::round(z);
            if (x >= 0 && x < mask_.a.x && y >= 0 && y < mask_.a.y && z >= 0 && z < mask_.a.z && mask_.at((int)x, (int)y, (int)z) == 1) {
                sum += s.at(i, j, k);
                num++;
            }
        }
        if (num > 0) stack_average.push_back(sum / num);
        else throw std::runtime_error("Stack " + std::to_string(ind) + " has no overlap with ROI");
    }
    double global_average = 0;
    if (together) {
        for (double a : stack_average) global_average += a;
        global_average /= stack_average.size();
    }
    for (size_t ind = 0; ind < stacks.size(); ++ind) {
        const double factor = together ? averageValue / global_average : averageValue / stack_average[ind];
        stack_factor_.push_back((float)factor);
        for (double& v : stacks[ind].v) if (v > 0) v *= factor;
    }
    if (debug) {
        std::cout << "Stack average intensities are ";
        for (double a : stack_average) std::cout << a << " ";
        std::cout << std::endl << "The new average value is " << averageValue << std::endl;
    }
}

// ---- CreateSlicesAndTransformations, irtkReconstructionGPU.cc:1814-1850 --------------------------------------------
void Reconstruction::CreateSlicesAndTransformations(const std::vector<Image>& stacks, const std::vector<Rigid>& t, const std::vector<double>& thickness)
{
    for (size_t i = 0; i < stacks.size(); ++i) {
        const ImageAttr& attr = stacks[i].a;
        for (int j = 0; j < attr.z; ++j) {
            Image slice = stacks[i].get_region(0, 0, j, attr.x, attr.y, j + 1);
            slice.a.dz = thickness[i];             // z size of a slice = slice thickness
            slices_.push_back(slice);
            stack_index_.push_back((int)i);
            transformations_.push_back(t[i]);
        }
    }
    std::cout << "Number of slices: " << slices_.size() << std::endl;
}

// ---- MaskSlices, irtkReconstructionGPU.cc:1940-1988 ----------------------------------------------------------------
void Reconstruction::MaskSlices()
{
    std::cout << "Masking slices ... ";
    if (!have_mask_) { std::cout << "Could not mask slices because no mask has been set." << std::endl; return; }
    const Mat4 mw2i = mask_.a.world_to_image();
    for (size_t n = 0; n < slices_.size(); ++n) {
        Image& s = slices_[n];
        const Mat4 m = mw2i * (transformations_[n].matrix() * s.a.image_to_world());
        for (int i = 0; i < s.a.x; ++i) for (int j = 0; j < s.a.y; ++j) {
            if (s.at(i, j, 0) < 0.01) s.at(i, j, 0) = -1;
            double x = i, y = j, z = 0;
            m.apply(x, y, z);
            x = std::round(x); y = std::round(y); z = std::round(z);
            if (x >= 0 && x < mask_.a.x && y >= 0 && y < mask_.a.y && z >= 0 && z < mask_.a.z) {
                if (mask_.at((int)x, (int)y, (int)z) == 0) s.at(i, j, 0) = -1;
            } else s.at(i, j, 0) = -1;
        }
    }
    std::cout << "done." << std::endl;
}

// ---- SetSmoothingParameters, irtkReconstructionGPU.h:605-612 --------------------------------------------------------
void Reconstruction::SetSmoothingParameters(double delta, double lambda)
{
    delta_ = delta;
    lambda_ = lambda * delta * delta;
    alpha_ = 0.05 / lambda;
    if (alpha_ > 1) alpha_ = 1;
}

// ---- ReadTransformation / SaveTransformations, irtkReconstructionGPU.cc:4733-4765, 4884-4919 -------------------------
void Reconstruction::ReadTransformation(const std::string& folder)
{
    if (slices_.empty()) throw std::runtime_error("Please create slices before reading transformations!");
    std::cout << "Reading transformations:" << std::endl;
    for (size_t i = 0; i < slices_.size(); ++i) {
        const std::string path = (folder.empty() ? std::string() : folder + "/") + "transformation" + std::to_string(i) + ".dof";
        Rigid r;
        if (!r.read_dof(path)) throw std::runtime_error("cannot read " + path);
        transformations_[i] = r;
        std::cout << path << std::endl;
    }
}

void Reconstruction::SaveTransformations(const std::string& prefix)
{
    const Mat4 rw2i = reconstructed_.a.world_to_image();
    for (size_t i = 0; i < slices_.size(); ++i) {
        transformations_[i].write_dof(prefix + "croppedSliceTransformation" + std::to_string(i) + ".dof");
        const Rigid t = Rigid::from_matrix(rw2i * (transformations_[i].matrix() * slices_[i].a.image_to_world()));
        t.write_dof(prefix + "croppedSliceToVolumeTransformation" + std::to_string(i) + ".dof");
    }
}

// ---- SyncGPU, irtkReconstructionGPU.cc:249-328 ---------------------------------------------------------------------
#define RCK(r, call, what) do { if ((call) != 0) throw std::runtime_error(std::string(what) + ": " + svr_last_error((r).c)); } while (0)

void Reconstruction::SyncGPU()
{
    std::cout << "SyncGPU()" << std::endl;
    ensure_context();
    const ImageAttr& ra = reconstructed_.a;
    std::vector<float> vol(reconstructed_.n());
    for (size_t i = 0; i < vol.size(); ++i) vol[i] = (float)reconstructed_.v[i];
    std::vector<float> mask(mask_.n());
    for (size_t i = 0; i < mask.size(); ++i) mask[i] = (float)mask_.v[i];

    int minx = INT_MAX, miny = INT_MAX;
    Nx_ = Ny_ = 0;
    for (const Image& s : slices_) {
        Nx_ = std::max(Nx_, s.a.x); Ny_ = std::max(Ny_, s.a.y);
        minx = std::min(minx, s.a.x); miny = std::min(miny, s.a.y);
    }
    const int S = (int)slices_.size();
    const double waste = ((double)(Nx_ - minx) * (Ny_ - miny) * S) * sizeof(double) * 5.0 / 1024.0;
    std::printf("GPU memory waste approx: %f KB with %d %d %d %d\n", waste, Nx_, Ny_, minx, miny);

    // which slices go where: every D-th slice of every stack (svr_host_partition_strided)
    std::vector<int> per_stack;
    for (int n = 0; n < S; ++n) {
        if (stack_index_[n] >= (int)per_stack.size()) per_stack.resize(stack_index_[n] + 1, 0);
        per_stack[stack_index_[n]]++;
    }
    for (size_t r = 0; r < ranks_.size(); ++r) {
        ranks_[r].idx.assign(std::max(S, 1), 0);
        int n = 0;
        ck(svr_host_partition_strided((int)per_stack.size(), per_stack.data(), (int)ranks_.size(), (int)r, ranks_[r].idx.data(), &n), "partition");
        ranks_[r].idx.resize(n);
        if (ranks_.size() > 1) std::cout << "rank " << r << " (device " << ranks_[r].device << "): " << n << " slices" << std::endl;
    }
    each_rank([&](Rank& r) {
        RCK(r, svr_init_reconstruction_volume(r.c, ra.x, ra.y, ra.z, (float)ra.dx, (float)ra.dy, (float)ra.dz, vol.data()), "InitReconstructionVolume");
        RCK(r, svr_set_mask(r.c, mask_.a.x, mask_.a.y, mask_.a.z, mask.data()), "setMask");
        const int Sl = (int)r.idx.size();
        RCK(r, svr_init_storage_volumes(r.c, Nx_, Ny_, Sl), "initStorageVolumes");
        std::vector<float> cube((size_t)Nx_ * Ny_ * std::max(Sl, 1), -1.0f);       // top-left aligned, pre-filled with the padding value
        std::vector<int> sx(std::max(Sl, 1)), sy(std::max(Sl, 1));
        std::vector<float> dims(3 * (size_t)std::max(Sl, 1));
        for (int j = 0; j < Sl; ++j) {
            const Image& s = slices_[r.idx[j]];
            for (int y = 0; y < s.a.y; ++y) for (int x = 0; x < s.a.x; ++x) cube[((size_t)j * Ny_ + y) * Nx_ + x] = (float)s.at(x, y, 0);
            sx[j] = s.a.x; sy[j] = s.a.y;
            dims[3 * j] = (float)s.a.dx; dims[3 * j + 1] = (float)s.a.dy; dims[3 * j + 2] = (float)s.a.dz;
        }
        RCK(r, svr_fill_slices(r.c, cube.data(), sx.data(), sy.data()), "FillSlices");
        RCK(r, svr_set_slice_dims(r.c, dims.data(), 1.0f), "setSliceDims");
        RCK(r, svr_synchronize(r.c), "synchronize");
    });
    scale_.assign(S, 1.0f);
    slice_weight_.assign(S, 1.0f);
    slice_potential_.assign(S, 0.0f);
    slice_inside_.assign(S, 0);
}

// ---- UpdateGPUTranformationMatrices, irtkReconstructionGPU.cc:372-401 ----------------------------------------------
void Reconstruction::UpdateGPUTranformationMatrices()
{
    const size_t S = slices_.size();
    std::vector<float> T(16 * S), Ti(16 * S), I2W(16 * S), W2I(16 * S);
    for (size_t i = 0; i < S; ++i) {
        const Mat4 t = transformations_[i].matrix();
        t.to_float16(&T[16 * i]);
        t.inverse().to_float16(&Ti[16 * i]);
        slices_[i].a.image_to_world().to_float16(&I2W[16 * i]);
        slices_[i].a.world_to_image().to_float16(&W2I[16 * i]);
    }
    float ri2w[16], rw2i[16];
    reconstructed_.a.image_to_world().to_float16(ri2w);
    reconstructed_.a.world_to_image().to_float16(rw2i);
    each_rank([&](Rank& r) {
        const std::vector<float> t = take(r, T, 16), ti = take(r, Ti, 16), a = take(r, I2W, 16), b = take(r, W2I, 16);
        RCK(r, svr_set_slice_matrices(r.c, t.data(), ti.data(), a.data(), b.data(), ri2w, rw2i), "SetSliceMatrices");
    });
}

// ---- generatePSFVolume, irtkReconstructionGPU.cc:1496-1610: only the PSF image attributes reach the device -----------
void Reconstruction::generatePSFVolume()
{
    ImageAttr attr;
    attr.x = attr.y = attr.z = 128;               // PSF_SIZE
    attr.dx = reconstructed_.a.dx; attr.dy = reconstructed_.a.dy; attr.dz = reconstructed_.a.dz;
    const int size[3] = { 128, 128, 128 };
    float i2w[16];
    attr.image_to_world().to_float16(i2w);
    for (Rank& r : ranks_) RCK(r, svr_generate_psf_volume(r.c, size, i2w, 1.0f), "generatePSFVolume");
}

// ---- InitializeEMGPU / InitializeEMValuesGPU, irtkReconstructionGPU.cc:2905-2953 -------------------------------------
void Reconstruction::InitializeEMValuesGPU()
{
    std::fill(slice_weight_.begin(), slice_weight_.end(), 1.0f);
    std::fill(scale_.begin(), scale_.end(), 1.0f);
    each_rank([&](Rank& r) {
        const std::vector<float> sc = take(r, scale_), sw = take(r, slice_weight_);
        RCK(r, svr_update_scale_vector(r.c, sc.data(), sw.data()), "UpdateScaleVector");
        RCK(r, svr_initialize_em_values(r.c), "InitializeEMValues");
    });
}

void Reconstruction::InitializeEMGPU()
{
    InitializeEMValuesGPU();
    max_intensity_ = -1e300; min_intensity_ = 1e300;
    for (const Image& s : slices_) for (double v : s.v) if (v > 0) { max_intensity_ = std::max(max_intensity_, v); min_intensity_ = std::min(min_intensity_, v); }
}

// ---- GaussianReconstructionGPU, irtkReconstructionGPU.cc:2695-2762 --------------------------------------------------
void Reconstruction::GaussianReconstructionGPU()
{
    std::cout << "Gaussian reconstruction ... ";
    const int S = (int)slices_.size();
    std::vector<int> voxel_num(std::max(S, 1));
    each_rank([&](Rank& r) {
        std::vector<int> vn(std::max<size_t>(r.idx.size(), 1));
        RCK(r, svr_gaussian_reconstruction_local(r.c), "GaussianReconstruction");
        allreduce_accumulator(r);                       // C1: numerator / denominator of the PSF-weighted splat
        RCK(r, svr_gaussian_reconstruction_finish(r.c, vn.data()), "GaussianReconstruction");
        put(r, vn, voxel_num);
    });
    std::cout << "done." << std::endl;
    small_slices_.assign(std::max(S, 1), 0);
    int n = 0;
    ck(svr_host_small_slices(S, voxel_num.data(), small_slices_.data(), &n), "small slices");
    small_slices_.resize(n);
    if (debug) {
        std::cout << "Small slices GPU:";
        for (int i : small_slices_) std::cout << " " << i;
        std::cout << std::endl;
    }
}

// ---- SimulateSlicesGPU, irtkReconstructionGPU.cc:1163-1203 ----------------------------------------------------------
void Reconstruction::SimulateSlicesGPU()
{
    each_rank([&](Rank& r) {
        std::vector<unsigned char> in(std::max<size_t>(r.idx.size(), 1));
        RCK(r, svr_simulate_slices(r.c, in.data()), "SimulateSlices");
        put(r, in, slice_inside_);
    });
}

// ---- InitializeRobustStatisticsGPU, irtkReconstructionGPU.cc:2988-3020 ----------------------------------------------
void Reconstruction::InitializeRobustStatisticsGPU()
{
    // C4: the partial sums of the ranks are added on the host (one process drives all ranks: no collective is needed)
    std::vector<double> part(2 * ranks_.size(), 0.0);
    each_rank([&](Rank& r) { RCK(r, svr_initialize_robust_statistics_local(r.c, &part[2 * (&r - ranks_.data())]), "InitializeRobustStatistics"); });
    double s0 = 0, s1 = 0;
    for (size_t r = 0; r < ranks_.size(); ++r) { s0 += part[2 * r]; s1 += part[2 * r + 1]; }
    sigma_ = (float)s0 / (float)s1;                 // _sigma = sa / sb, cuda2.cu:2305
    for (size_t i = 0; i < slices_.size(); ++i) if (!slice_inside_[i]) slice_weight_[i] = 0;
    for (int i : force_excluded_) if (i >= 0 && (size_t)i < slice_weight_.size()) slice_weight_[i] = 0;
    state5_[0] = 0.025f;                           // sigma_s
    mix_ = 0.9f;
    state5_[1] = 0.9f;                             // mix_s
    m_ = (float)(1.0f / (2.1f * max_intensity_ - 1.9f * min_intensity_));
    if (debug) std::cout << "Initializing robust statistics GPU: sigma=" << std::sqrt(sigma_) << " m=" << m_ << " mix=" << mix_ << " mix_s=" << state5_[1] << std::endl;
    each_rank([&](Rank& r) {
        const std::vector<float> sc = take(r, scale_), sw = take(r, slice_weight_);
        RCK(r, svr_update_scale_vector(r.c, sc.data(), sw.data()), "UpdateScaleVector");
    });
}

// ---- EStepGPU, irtkReconstructionGPU.cc:3162-3440 (device part + host slice-level EM) --------------------------------
void Reconstruction::EStepGPU()
{
    const int S = (int)slices_.size();
    each_rank([&](Rank& r) {
        std::vector<float> pot(std::max<size_t>(r.idx.size(), 1));
        RCK(r, svr_estep(r.c, m_, sigma_, mix_, pot.data()), "EStep");
        put(r, pot, slice_potential_);
    });
    ck(svr_host_slice_em(S, slice_potential_.data(), scale_.data(), slice_weight_.data(), force_excluded_.data(), (int)force_excluded_.size(),
                         small_slices_.data(), (int)small_slices_.size(), step_, state5_), "slice EM");
    each_rank([&](Rank& r) {
        const std::vector<float> sw = take(r, slice_weight_);
        RCK(r, svr_update_slice_weights(r.c, sw.data()), "UpdateSliceWeights");
    });
}

// ---- ScaleGPU, irtkReconstructionGPU.cc:3751-3765 -------------------------------------------------------------------
void Reconstruction::ScaleGPU()
{
    each_rank([&](Rank& r) {
        std::vector<float> sc(std::max<size_t>(r.idx.size(), 1));
        RCK(r, svr_calculate_scale_vector(r.c, sc.data()), "CalculateScaleVector");
        put(r, sc, scale_);
    });
}

// ---- SuperresolutionGPU, irtkReconstructionGPU.cc:4024-4053 ---------------------------------------------------------
void Reconstruction::SuperresolutionGPU(int iter)
{
    (void)iter;
    each_rank([&](Rank& r) {
        const std::vector<float> sw = take(r, slice_weight_);
        RCK(r, svr_superresolution_local(r.c, sw.data()), "Superresolution");
        allreduce_accumulator(r);                       // C2: addon / confidence map; every rank then regularises its replica (no C3 broadcast)
        RCK(r, svr_superresolution_finish(r.c, adaptive_ ? 1 : 0, (float)alpha_, (float)min_intensity_, (float)max_intensity_, (float)delta_,
                                          (float)lambda_), "Superresolution");
    });
}

// ---- MStepGPU, irtkReconstructionGPU.cc:4214-4224 -------------------------------------------------------------------
void Reconstruction::MStepGPU(int iter)
{
    std::vector<double> part(5 * ranks_.size(), 0.0);
    each_rank([&](Rank& r) { RCK(r, svr_mstep_local(r.c, &part[5 * (&r - ranks_.data())]), "MStep"); });
    double s5[5] = { 0, 0, 0, 0, 0 };               // sums; min / max are seeded with 0 by the kernel (cuda2.cu:3103,3110)
    for (size_t r = 0; r < ranks_.size(); ++r) {
        for (int q = 0; q < 3; ++q) s5[q] += part[5 * r + q];
        s5[3] = std::min(s5[3], part[5 * r + 3]);
        s5[4] = std::max(s5[4], part[5 * r + 4]);
    }
    ck(svr_mstep_finish(s5, iter, (float)step_, &sigma_, &mix_, &m_), "MStep");
    if (debug) std::cout << "Voxel-wise robust statistics parameters GPU: sigma = " << std::sqrt(sigma_) << " mix = " << mix_ << " m = " << m_ << std::endl;
}

void Reconstruction::MaskVolumeGPU() { each_rank([&](Rank& r) { RCK(r, svr_mask_volume(r.c), "maskVolume"); }); }

void Reconstruction::ScaleVolumeGPU()
{
    std::vector<double> part(2 * ranks_.size(), 0.0);
    each_rank([&](Rank& r) { RCK(r, svr_scale_volume_local(r.c, &part[2 * (&r - ranks_.data())]), "ScaleVolume"); });
    double s0 = 0, s1 = 0;
    for (size_t r = 0; r < ranks_.size(); ++r) { s0 += part[2 * r]; s1 += part[2 * r + 1]; }
    const float scale = (float)(s0 / s1);          // cuda2.cu:3459
    each_rank([&](Rank& r) { RCK(r, svr_scale_volume_apply(r.c, scale), "ScaleVolume"); });
}

// ---- RestoreSliceIntensitiesGPU, irtkReconstructionGPU.cc:1026-1032 -------------------------------------------------
void Reconstruction::RestoreSliceIntensitiesGPU()
{
    each_rank([&](Rank& r) {
        const std::vector<int> si = take(r, stack_index_);
        if (si.empty()) return;
        RCK(r, svr_restore_slice_intensities(r.c, stack_factor_.data(), (int)stack_factor_.size(), si.data()), "RestoreSliceIntensities");
    });
}

// ---- SyncCPU, irtkReconstructionGPU.cc:2675-2683 --------------------------------------------------------------------
void Reconstruction::SyncCPU()
{
    std::vector<float> vol(reconstructed_.n());
    ck(svr_sync_cpu(c_, vol.data()), "syncCPU");        // every rank holds the same replica: rank 0's is read
    for (size_t i = 0; i < vol.size(); ++i) reconstructed_.v[i] = vol[i];
}

// ---- irtkResamplingWithPadding (image++/src/irtkResamplingWithPadding.cc:36-183), z-plane 0 of a one-plane slice ------
static Image resample_slice_with_padding(const Image& in, double d, double padding)
{
    Image out(resampled_attr_with_padding(in.a, d, d, d), padding);
    const Mat4 mo = out.a.image_to_world(), mi = in.a.world_to_image();     // applied one after the other, as the reference does
    for (int k = 0; k < out.a.z; ++k) for (int j = 0; j < out.a.y; ++j) for (int i = 0; i < out.a.x; ++i) {
        double x = i, y = j, z = k;
        mo.apply(x, y, z);
        mi.apply(x, y, z);
        const int u = (int)std::floor(x), v = (int)std::floor(y), w = (int)std::floor(z);
        const double fx = x - u, fy = y - v, fz = z - w;
        double val = 0, wsum = 0;
        int pad = 8;
        for (int du = 0; du < 2; ++du) for (int dv = 0; dv < 2; ++dv) for (int dw = 0; dw < 2; ++dw) {
            const int uu = u + du, vv = v + dv, ww = w + dw;
            const bool inb = uu >= 0 && uu < in.a.x && vv >= 0 && vv < in.a.y && ww >= 0 && ww < in.a.z;
            const double wt = (du ? fx : 1 - fx) * (dv ? fy : 1 - fy) * (dw ? fz : 1 - fz);
            if (inb) {
                const double g = in.at(uu, vv, ww);
                if (g != padding) { val += g * wt; wsum += wt; pad--; }
            } else pad--;                           // out-of-bounds neighbours count as not padded but add nothing
        }
        out.at(i, j, k) = (pad < 4 && wsum > 0) ? val / wsum : padding;
    }
    return out;
}

// ---- PrepareRegistrationSlices, irtkReconstructionGPU.cc:2105-2179 --------------------------------------------------
// The reference resamples its host copies of the slices and uploads them (FillRegSlices); the slices are already on the
// device here (SyncGPU), so the resampling runs there (svr_reg_resample_slices: same rules, double precision, on the
// float32 slices the device holds: within one float ulp of resampling the host's double copies).  With --debug the host
// form above is evaluated as well and the largest difference is printed.
void Reconstruction::PrepareRegistrationSlices()
{
    const int S = (int)slices_.size();
    const double d = reconstructed_.a.dx;
    res_attrs_.clear();
    regW_ = regH_ = 0;
    int minx = INT_MAX, miny = INT_MAX;
    for (const Image& s : slices_) {
        res_attrs_.push_back(resampled_attr_with_padding(s.a, d, d, d));
        const ImageAttr& ra = res_attrs_.back();
        regW_ = std::max(regW_, ra.x); regH_ = std::max(regH_, ra.y);
        minx = std::min(minx, ra.x); miny = std::min(miny, ra.y);
    }
    const double waste = ((double)(regW_ - minx) * (regH_ - miny) * S) * sizeof(double) * 5.0 / 1024.0;
    std::printf("GPU memory waste approx RegSlices: %f KB with %d %d %d %d\n", waste, regW_, regH_, minx, miny);
    std::vector<double> m(24 * (size_t)S);
    std::vector<int> in_sizes(2 * (size_t)S), out_sizes(2 * (size_t)S);
    std::vector<float> i2w(16 * (size_t)S);
    for (int n = 0; n < S; ++n) {
        const Mat4 mo = res_attrs_[n].image_to_world(), mi = slices_[n].a.world_to_image();
        for (int r = 0; r < 3; ++r) for (int q = 0; q < 4; ++q) { m[24 * (size_t)n + 4 * r + q] = mo.m[r][q]; m[24 * (size_t)n + 12 + 4 * r + q] = mi.m[r][q]; }
        in_sizes[2 * n] = slices_[n].a.x; in_sizes[2 * n + 1] = slices_[n].a.y;
        out_sizes[2 * n] = res_attrs_[n].x; out_sizes[2 * n + 1] = res_attrs_[n].y;
        res_attrs_[n].image_to_world().to_float16(&i2w[16 * n]);
    }
    each_rank([&](Rank& r) {                          // every rank resamples (and later registers) its own slices: no exchange
        const int Sl = (int)r.idx.size();
        RCK(r, svr_reg_init_storage(r.c, regW_, regH_, Sl, (float)d, (float)d, (float)d), "initRegStorageVolumes");
        if (Sl == 0) return;
        const std::vector<double> ml = take(r, m, 24);
        const std::vector<int> is = take(r, in_sizes, 2), os = take(r, out_sizes, 2);
        const std::vector<float> il = take(r, i2w, 16);
        RCK(r, svr_reg_resample_slices(r.c, ml.data(), is.data(), os.data(), il.data()), "resampleRegSlices");
    });
    if (debug && S > 0 && ranks_.size() == 1) {
        std::vector<float> cube((size_t)regW_ * regH_ * S);
        ck(svr_reg_debug_get(c_, 0, cube.data()), "debugRegSlices");
        // the device resamples the float32 slices it was given, the host form its double copies: up to one float ulp apart
        double worst = 0;
        for (int n = 0; n < S; n += std::max(1, S / 8)) {          // a sample of the slices
            const Image h = resample_slice_with_padding(slices_[n], d, -1);
            for (int y = 0; y < h.a.y; ++y) for (int x = 0; x < h.a.x; ++x) {
                const double a = (double)(float)h.at(x, y, 0), b = (double)cube[((size_t)n * regH_ + y) * regW_ + x];
                worst = std::max(worst, std::fabs(a - b) / std::max(1.0, std::fabs(a)));
            }
        }
        std::cout << "registration slices: device vs host resampling, max relative difference " << worst << std::endl;
    }
    reg_prepared_ = true;
}

// ---- SliceToVolumeRegistrationGPU, irtkReconstructionGPU.cc:2218-2288 -----------------------------------------------
void Reconstruction::SliceToVolumeRegistrationGPU()
{
    if (!reg_prepared_) PrepareRegistrationSlices();
    const size_t S = slices_.size();
    std::vector<Mat4> mos(S);
    std::vector<float> transf(16 * S), ofs(16 * S);
    for (size_t i = 0; i < S; ++i) {
        ImageAttr a = res_attrs_[i];
        Mat4 mo = Mat4::identity();                  // offset: translation by the slice origin
        for (int q = 0; q < 3; ++q) { mo.m[q][3] = a.origin[q]; a.origin[q] = 0; }
        mos[i] = mo;
        (transformations_[i].matrix() * mo).to_float16(&transf[16 * i]);
        a.image_to_world().to_float16(&ofs[16 * i]);
    }
    each_rank([&](Rank& r) {
        if (r.idx.empty()) return;
        const std::vector<float> ol = take(r, ofs, 16);
        std::vector<float> tl = take(r, transf, 16);
        RCK(r, svr_reg_update_slices_i2w(r.c, ol.data()), "updateResampledSlicesI2W");
        RCK(r, svr_reg_prepare(r.c), "prepareSliceToVolumeReg");
        RCK(r, svr_reg_register(r.c, tl.data()), "registerSlicesToVolume");
        put(r, tl, transf, 16);
    });
    for (size_t i = 0; i < S; ++i) {
        const Mat4 mat = Mat4::from_float16(&transf[16 * i]) * mos[i].inverse();
        transformations_[i] = Rigid::from_matrix(mat);
    }
}

// ---- the reference's default (IRTK) registrations on the device engine -------------------------------------------------------
static void attr18(const ImageAttr& a, double* o)
{
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.dx; o[4] = a.dy; o[5] = a.dz;
    for (int i = 0; i < 3; ++i) { o[6 + i] = a.origin[i]; o[9 + i] = a.xaxis[i]; o[12 + i] = a.yaxis[i]; o[15 + i] = a.zaxis[i]; }
}
// irtkGreyImage = irtkRealImage: static_cast<short> per voxel (image++/src/irtkGenericImage.cc:699-713)
static std::vector<short> to_grey(const Image& im)
{
    std::vector<short> g(im.n());
    for (size_t i = 0; i < g.size(); ++i) g[i] = static_cast<short>(im.v[i]);
    return g;
}
// ResetOrigin (irtkReconstructionGPU.cc:823-834): the image origin moves into a translation
static Mat4 reset_origin(ImageAttr& a)
{
    Mat4 mo = Mat4::identity();
    for (int q = 0; q < 3; ++q) { mo.m[q][3] = a.origin[q]; a.origin[q] = 0; }
    return mo;
}

// StackRegistrations, irtkReconstructionGPU.cc:940-1001 + ParallelStackRegistrations :849-938
void Reconstruction::StackRegistrations(std::vector<Image>& stacks, std::vector<Rigid>& stack_transformations, int templateNumber)
{
    ensure_context();
    if (stacks.size() < 2) return;
    InvertStackTransformations(stack_transformations);
    Image target = stacks[templateNumber];
    std::vector<short> tgrey = to_grey(target);
    if (have_mask_) {                                  // the target is masked before registration (:957-984)
        const Mat4 t_i2w = target.a.image_to_world(), m_w2i = mask_.a.world_to_image();
        for (int k = 0; k < target.a.z; ++k) for (int j = 0; j < target.a.y; ++j) for (int i = 0; i < target.a.x; ++i) {
            double x = i, y = j, z = k;
            t_i2w.apply(x, y, z);
            m_w2i.apply(x, y, z);
            const double rx = std::round(x), ry = std::round(y), rz = std::round(z);
            bool keep = false;
            if (rx >= 0 && rx < mask_.a.x && ry >= 0 && ry < mask_.a.y && rz >= 0 && rz < mask_.a.z) keep = mask_.at((int)rx, (int)ry, (int)rz) != 0;
            if (!keep) tgrey[((size_t)k * target.a.y + j) * target.a.x + i] = 0;
        }
    }
    ImageAttr ta = target.a;
    const Mat4 mo = reset_origin(ta);
    std::vector<std::vector<short>> grey;
    std::vector<const short*> vox;
    std::vector<double> attrs;
    std::vector<int> tgt, src, which;
    std::vector<double> dofs;
    grey.push_back(tgrey);
    attrs.resize(18); attr18(ta, attrs.data());
    for (size_t i = 0; i < stacks.size(); ++i) {
        if ((int)i == templateNumber) continue;        // "do not perform registration for template"
        grey.push_back(to_grey(stacks[i]));
        attrs.resize(attrs.size() + 18); attr18(stacks[i].a, attrs.data() + attrs.size() - 18);
        stack_transformations[i] = Rigid::from_matrix(stack_transformations[i].matrix() * mo);     // include the offset (:889-892)
        tgt.push_back(0); src.push_back((int)grey.size() - 1); which.push_back((int)i);
        for (int q = 0; q < 6; ++q) dofs.push_back(stack_transformations[i].p[q]);
    }
    for (auto& g : grey) vox.push_back(g.data());
    int64_t evals = 0;
    ck(svr_rreg_register(c_, (int)tgt.size(), (int)vox.size(), vox.data(), attrs.data(), tgt.data(), src.data(), 0, dofs.data(), nullptr, &evals, -1, 0,
                         nullptr, nullptr, nullptr, nullptr), "StackRegistrations");
    const Mat4 moi = mo.inverse();
    for (size_t a = 0; a < which.size(); ++a) {
        Rigid t; for (int q = 0; q < 6; ++q) t.p[q] = dofs[6 * a + q];
        stack_transformations[which[a]] = Rigid::from_matrix(t.matrix() * moi);
        std::cout << "stack " << which[a] << " registered to the template: " << t.p[0] << " " << t.p[1] << " " << t.p[2] << " " << t.p[3] << " " << t.p[4] << " "
                  << t.p[5] << std::endl;
    }
    std::cout << "StackRegistrations: " << evals << " similarity evaluations on the device" << std::endl;
    InvertStackTransformations(stack_transformations);
}

// SliceToVolumeRegistration (the reference's default, CPU / IRTK), irtkReconstructionGPU.cc:1992-2059, on the device engine:
// every slice resampled to the volume's voxel size with padding (host, double, as there), cast to short, origin reset; the
// volume (SyncCPU copy) cast to short is the shared source.
void Reconstruction::SliceToVolumeRegistration()
{
    const double d = reconstructed_.a.dx;
    const std::vector<short> source = to_grey(reconstructed_);
    each_rank([&](Rank& r) {
        std::vector<std::vector<short>> grey;
        std::vector<const short*> vox;
        std::vector<double> attrs(18), dofs;
        std::vector<int> tgt, src, which;
        std::vector<Mat4> mos;
        attr18(reconstructed_.a, attrs.data());
        for (int gi : r.idx) {
            Image t = resample_slice_with_padding(slices_[gi], d, -1);
            std::vector<short> g = to_grey(t);
            short smax = -32768;
            for (short v : g) smax = std::max(smax, v);
            if (!(smax > -1)) continue;                  // "if (smax > -1)": empty slices keep their transformation
            ImageAttr a = t.a;
            const Mat4 mo = reset_origin(a);
            const Rigid start = Rigid::from_matrix(transformations_[gi].matrix() * mo);
            grey.push_back(std::move(g));
            attrs.resize(attrs.size() + 18); attr18(a, attrs.data() + attrs.size() - 18);
            tgt.push_back((int)grey.size()); src.push_back(0); which.push_back(gi); mos.push_back(mo);
            for (int q = 0; q < 6; ++q) dofs.push_back(start.p[q]);
        }
        if (which.empty()) return;
        vox.push_back(source.data());
        for (auto& g : grey) vox.push_back(g.data());
        RCK(r, svr_rreg_register(r.c, (int)which.size(), (int)vox.size(), vox.data(), attrs.data(), tgt.data(), src.data(), 1, dofs.data(), nullptr, nullptr,
                                 -1, 0, nullptr, nullptr, nullptr, nullptr), "SliceToVolumeRegistration");
        for (size_t a = 0; a < which.size(); ++a) {
            Rigid t; for (int q = 0; q < 6; ++q) t.p[q] = dofs[6 * a + q];
            transformations_[which[a]] = Rigid::from_matrix(t.matrix() * mos[a].inverse());          // undo the offset
        }
    });
}

// ---- EvaluateGPU, irtkReconstructionGPU.cc:4503-4538 ----------------------------------------------------------------
void Reconstruction::EvaluateGPU(int iter, std::ostream& os)
{
    os << "Iteration " << iter << ": " << std::endl;
    auto list = [&](const char* title, auto pred) {
        os << title;
        int sum = 0;
        for (size_t i = 0; i < slices_.size(); ++i) if (pred(i)) { os << i << " "; sum++; }
        os << std::endl << "Total GPU: " << sum << std::endl;
    };
    list("Included slices GPU: ", [&](size_t i) { return slice_weight_[i] >= 0.5f && slice_inside_[i]; });
    list("Excluded slices  GPU: ", [&](size_t i) { return slice_weight_[i] < 0.5f && slice_inside_[i]; });
    list("Outside slices GPU: ", [&](size_t i) { return !slice_inside_[i]; });
}

}  // namespace svr
