// pvr_main.cc -- PVRreconstructionGPU: the reference's patch-to-volume command line
// (source/reconstructionGPU2/patchBasedReconMain.cpp:51-470 + irtkPatchBasedReconstruction<T>::run(),
// irtkPatchBasedReconstruction.cpp:194-593) over the PVR entry points of libsvr_b200 (include/pvr_abi.h).
// Same option names and defaults, same set-up order (mask binarise / dilate, per-stack TransformMask + CropImage, mask
// resampled to the isotropic grid, intensity matching, template, patch enumeration with the 1/3-coverage rule of
// include/patchBasedObject.cuh:176-342), same iteration loop and the same output files
// (reconimage<iter>_<patchSize>_<patchStride>.nii.gz every iteration, -o at the end).
// What differs (printed at start-up):
//   * the patch-to-volume registration between iterations (patchBased2D3DRegistration::runHybrid, IRTK on the CPU) and
//     the 3D stack-to-stack registration are not restated: patches keep the -t transformation of
//     their stack, so every pass of the iteration loop reconstructs from the same geometry;
//   * --hierarchical, --resample (B-spline), --useFullSlices and the evaluation options are refused.
// --superpixel runs SLICO per slice on the host (pvr_slic.cc) and cuts one 64 x 64 patch per superpixel with its
// char[64*64] mask (PatchBasedVolume::generate2DSuperpixelPatches, include/patchBasedObject.cuh:433-797).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/pvr_abi.h"
#include "pvr_slic.h"
#include "svr_image.h"
#include "svr_reconstruction.h"

namespace {

using svr::Image;
using svr::ImageAttr;
using svr::Mat4;
using svr::Rigid;

struct Options {
    std::string output, mask, existing_target, dump_patches;
    std::vector<std::string> input, transformation;
    std::vector<unsigned> patchSize = { 32, 32 }, patchStride = { 16, 16 };   // patchBasedReconMain.cpp:99-102
    std::vector<double> thickness;
    std::vector<int> devices, packages;
    double resolution = 0.75;
    int iterations = 7, sr_iterations = 7, dilateMask = 0;
    bool noMatchIntensities = false, debug = false, superpixel = false;
    unsigned spxSize = 16, spxExtend = 50;                                     // patchBasedReconMain.cpp:105-106
};

void usage()
{
    std::cout << "Application to perform reconstruction of volumetric MRI from thick patches.\nOptions:\n"
        "  -h [ --help ]                    Print usage messages\n"
        "  -o [ --output ] arg              Name for the reconstructed volume. Nifti format.\n"
        "  -m [ --mask ] arg                Binary mask to define the region od interest. [Default: the overlap of the stacks]\n"
        "  -i [ --input ] arg               [stack_1] .. [stack_N]  The input stacks.\n"
        "  -e [ --existingReconTarget ] arg Set an existing reconstruction as target image.\n"
        "  --patchSize arg                  size of the 2D patches [Default: 32 32]\n"
        "  --patchStride arg                stride of the 2D patches [Default: 16 16]\n"
        "  --resolution arg (=0.75)         Isotropic output resolution of the volume.\n"
        "  -t [ --transformation ] arg      The transformations of the input stacks to the template in 'dof' format; 'id' = identity.\n"
        "  --noMatchIntensities             Skip match intensities between the input stacks\n"
        "  --dilateMask arg (=0)            Dilate reconstruction mask n-iterations.\n"
        "  --iterations arg (=7)            number of registration iterations.\n"
        "  --sr_iterations arg (=7)         number of Super-resolution iterations.\n"
        "  -d [ --devices ] arg             GPU to use (one device per process)\n"
        "  --thickness arg                  [th_1] .. [th_N] patch thickness. [Default: twice voxel size in z direction]\n"
        "  --debug                          Write debug images.\n"
        "  -s [ --superpixel ]              Turn on superpixel-based reconstruction. [Default: false]\n"
        "  --spxSize arg                    initial size (<=64) of the 2D superpixels [Default: 16]\n"
        "  --spxExtend arg                  ratio [0-100]% of a superpixel's size for dilation / overlap [Default: 50%]\n"
        "  -p [ --packages ] arg            Number of packages used during acquisition for each stack: every stack is split\n"
        "                                   into that many interleaved sub-stacks, which the reconstruction treats as stacks.\n"
        "  --hierarchical / --resample / --useFullSlices   (not supported by this build)\n"
        "  --dump_patches arg               (this build only) write the enumerated patches (matrices, counts, cropped stacks,\n"
        "                                   resampled mask) into directory arg and exit; needs no GPU.\n";
}

bool parse(int argc, char** argv, Options& o, bool& help)
{
    std::map<std::string, std::string> alias = { { "-h", "--help" }, { "-o", "--output" }, { "-m", "--mask" }, { "-i", "--input" },
        { "-e", "--existingReconTarget" }, { "-t", "--transformation" }, { "-p", "--packages" }, { "-d", "--devices" }, { "-s", "--superpixel" },
        { "-v", "--evaluation" } };
    auto is_option = [&](const char* a) {
        if (a[0] != '-' || a[1] == 0) return false;
        if (a[1] == '-') return true;
        return alias.count(a) > 0;
    };
    int i = 1;
    auto values = [&](std::vector<std::string>& out) { while (i + 1 < argc && !is_option(argv[i + 1])) out.push_back(argv[++i]); };
    auto one = [&](const std::string& name, std::string& out) {
        if (i + 1 >= argc) { std::cerr << "ERROR: the required argument for option '" << name << "' is missing" << std::endl; return false; }
        out = argv[++i];
        return true;
    };
    auto unsupported = [&](const std::string& a) {
        std::cerr << "ERROR: option '" << a << "' is not supported by this build (see --help)" << std::endl;
        return false;
    };
    for (; i < argc; ++i) {
        std::string a = argv[i];
        if (alias.count(a)) a = alias[a];
        std::string v;
        std::vector<std::string> vs;
        if (a == "--help") { help = true; return true; }
        else if (a == "--output") { if (!one(a, o.output)) return false; }
        else if (a == "--mask") { if (!one(a, o.mask)) return false; }
        else if (a == "--input") values(o.input);
        else if (a == "--existingReconTarget") { if (!one(a, o.existing_target)) return false; }
        else if (a == "--patchSize") { values(vs); if (!vs.empty()) { o.patchSize.clear(); for (auto& s : vs) o.patchSize.push_back((unsigned)atoi(s.c_str())); } }
        else if (a == "--patchStride") { values(vs); if (!vs.empty()) { o.patchStride.clear(); for (auto& s : vs) o.patchStride.push_back((unsigned)atoi(s.c_str())); } }
        else if (a == "--resolution") { if (!one(a, v)) return false; o.resolution = atof(v.c_str()); }
        else if (a == "--transformation") values(o.transformation);
        else if (a == "--noMatchIntensities") o.noMatchIntensities = true;
        else if (a == "--debug") o.debug = true;
        else if (a == "--dilateMask") { if (!one(a, v)) return false; o.dilateMask = atoi(v.c_str()); }
        else if (a == "--iterations") { if (!one(a, v)) return false; o.iterations = atoi(v.c_str()); }
        else if (a == "--sr_iterations") { if (!one(a, v)) return false; o.sr_iterations = atoi(v.c_str()); }
        else if (a == "--devices") { values(vs); for (auto& s : vs) o.devices.push_back(atoi(s.c_str())); }
        else if (a == "--thickness") { values(vs); for (auto& s : vs) o.thickness.push_back(atof(s.c_str())); }
        else if (a == "--dump_patches") { if (!one(a, o.dump_patches)) return false; }
        else if (a == "--superpixel") o.superpixel = true;
        else if (a == "--spxSize") { if (!one(a, v)) return false; o.spxSize = (unsigned)atoi(v.c_str()); }
        else if (a == "--spxExtend") { if (!one(a, v)) return false; o.spxExtend = (unsigned)atoi(v.c_str()); }
        else if (a == "--packages") { values(vs); for (auto& q : vs) o.packages.push_back(atoi(q.c_str())); }
        else if (a == "--hierarchical" || a == "--resample" ||
                 a == "--useFullSlices" || a == "--evaluateGt" || a == "--evaluation" || a == "--evaluateBaseline" ||
                 a == "--patchExtraction")
            return unsupported(a);
        else { std::cerr << "ERROR: unrecognised option '" << argv[i] << "'" << std::endl; return false; }
    }
    if (o.output.empty()) { std::cerr << "ERROR: the option '--output' is required but missing" << std::endl; return false; }
    if (o.patchSize.size() < 2 || o.patchStride.size() < 2 || o.patchSize[0] == 0 || o.patchSize[1] == 0 || o.patchStride[0] == 0 ||
        o.patchStride[1] == 0) {
        std::cerr << "ERROR: --patchSize and --patchStride take two positive values" << std::endl;
        return false;
    }
    if (o.superpixel) {                                  // patchBasedReconMain.cpp:289-296
        if (o.spxSize == 0 || o.spxSize > 64 || o.spxExtend > 100) { std::cerr << "ERROR: --spxSize is 1..64 and --spxExtend 0..100" << std::endl; return false; }
        o.patchSize = { o.spxSize, o.spxSize };
        o.patchStride = { o.spxExtend, o.spxExtend };
    }
    if (o.patchSize[0] > 64 || o.patchSize[1] > 64) {      // the device patch cube and the 64x64 superpixel mask layout (reconConfig.cuh)
        std::cerr << "ERROR: patches are at most 64 x 64" << std::endl;
        return false;
    }
    return true;
}

svr_context* g_ctx = nullptr;
void ck(int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + svr_last_error(g_ctx));
}

// 26-connected dilation, one iteration (irtkDilation<T>, CONNECTIVITY_26)
void dilate26(Image& m)
{
    const Image src = m;
    for (int k = 0; k < m.a.z; ++k) for (int j = 0; j < m.a.y; ++j) for (int i = 0; i < m.a.x; ++i) {
        double best = src.at(i, j, k);
        for (int c = std::max(k - 1, 0); c <= std::min(k + 1, m.a.z - 1); ++c)
            for (int b = std::max(j - 1, 0); b <= std::min(j + 1, m.a.y - 1); ++b)
                for (int a = std::max(i - 1, 0); a <= std::min(i + 1, m.a.x - 1); ++a) best = std::max(best, src.at(a, b, c));
        m.at(i, j, k) = best;
    }
}

// computeMinMaxIntensities, irtkPatchBasedReconstruction.cpp:792-814 (positive voxels only)
void min_max(const std::vector<Image>& stacks, double& mn, double& mx)
{
    mx = -std::numeric_limits<float>::max();
    mn = std::numeric_limits<float>::max();
    for (const Image& s : stacks) for (double v : s.v) if (v > 0) { mx = std::max(mx, v); mn = std::min(mn, v); }
}

// MatchStackIntensitiesWithMasking, irtkPatchBasedReconstruction.cpp:656-789: the target average is the mean of all positive
// voxels; stack averages are taken over positive voxels whose centre maps (rounded) onto a mask voxel equal to 1.
void match_intensities(std::vector<Image>& stacks, const std::vector<Rigid>& t, const Image& mask, bool debug)
{
    double average = 0;
    unsigned long long count = 0;
    for (const Image& s : stacks) for (double v : s.v) if (v > 0) { average += v; ++count; }
    if (count) average /= (double)count;
    std::printf("Average value: %f \n", average);
    const Mat4 mw2i = mask.a.world_to_image();
    std::vector<double> stack_average;
    for (size_t ind = 0; ind < stacks.size(); ++ind) {
        const Image& s = stacks[ind];
        const Mat4 m = mw2i * (t[ind].matrix() * s.a.image_to_world());
        double sum = 0, num = 0;
        for (int i = 0; i < s.a.x; ++i) for (int j = 0; j < s.a.y; ++j) for (int k = 0; k < s.a.z; ++k) {
            double x = i, y = j, z = k;
            m.apply(x, y, z);
            x = std::round(x); y = std::round(y); z = std::round(z);
            if (x >= 0 && x < mask.a.x && y >= 0 && y < mask.a.y && z >= 0 && z < mask.a.z && mask.at((int)x, (int)y, (int)z) == 1 &&
                s.at(i, j, k) > 0) {
                sum += s.at(i, j, k);
                num++;
            }
        }
        if (num > 0) stack_average.push_back(sum / num);
        else throw std::runtime_error("Stack " + std::to_string(ind) + " has no overlap with ROI");
    }
    for (size_t ind = 0; ind < stacks.size(); ++ind) {
        const double factor = average / stack_average[ind];
        for (double& v : stacks[ind].v) if (v > 0) v *= factor;
    }
    if (debug) {
        std::cout << "Stack average intensities are ";
        for (double a : stack_average) std::cout << a << " ";
        std::cout << std::endl;
    }
}

// irtkResampling<T>(d, d, d) with the nearest-neighbour interpolator: the grid of resampled_attr, each voxel taking the
// source voxel nearest to its centre.
Image resample_nn(const Image& src, double d)
{
    Image out(svr::resampled_attr(src.a, d, d, d), 1.0);
    svr::transform_image_nn(src, Rigid(), out, /*target_padding=*/-1, /*source_padding=*/0);
    return out;
}

struct Patches {                      // one stack
    std::vector<float> i2w, w2i;      // 16 floats per patch
    std::vector<char> spx;            // superpixel mode: char[64*64] per patch, '1' = pixel belongs to the patch
    int pbx = 0, pby = 0;             // superpixel mode: the patch size used (64 x 64 clamped to the stack)
    std::vector<int> labels;          // superpixel mode: SLICO labels [z][y][x] (for --dump_patches)
    std::vector<int> label_of_patch;  // superpixel mode: (slice, label) per patch
    int count = 0;
    long long total_pixels = 0;
};

// PatchBasedVolume::generate2DPatches, include/patchBasedObject.cuh:176-342: boxes of pbb pixels on a stride grid that runs
// to size + pbb over every slice; a patch is kept when more than 1/3 of its pixels lie inside the slice and the mask with
// a value that is neither 0 nor -1.  irtkGenericImage::Get(double ...) truncates its coordinates.
Patches generate_2d_patches(const Image& stack, const Image& mask, int pbx, int pby, int stride_x, int stride_y, double thickness)
{
    Patches out;
    const ImageAttr& attr = stack.a;
    const Mat4 stack_i2w = attr.image_to_world();
    const Mat4 mask_w2i = mask.a.world_to_image();
    for (int z = 0; z < attr.z; ++z) {
        ImageAttr sattr = attr;                         // GetRegion(0, 0, z, X, Y, z + 1) + PutPixelSize(dx, dy, 2 * thickness)
        sattr.z = 1;
        sattr.dz = thickness * 2;
        double cx = (attr.x - 1) / 2.0, cy = (attr.y - 1) / 2.0, cz = z;
        stack_i2w.apply(cx, cy, cz);
        sattr.origin[0] = cx; sattr.origin[1] = cy; sattr.origin[2] = cz;
        const Mat4 s_i2w = sattr.image_to_world(), s_w2i = sattr.world_to_image();
        for (int y = 0; y < attr.y + pby; y += stride_y) {
            for (int x = 0; x < attr.x + pbx; x += stride_x) {
                ImageAttr pa = sattr;
                pa.x = pbx; pa.y = pby;
                pa.origin[0] = pa.origin[1] = pa.origin[2] = 0;
                double x1 = x, y1 = y, z1 = 0, x2 = 0, y2 = 0, z2 = 0;
                s_i2w.apply(x1, y1, z1);
                pa.image_to_world().apply(x2, y2, z2);
                pa.origin[0] = x1 - x2; pa.origin[1] = y1 - y2; pa.origin[2] = z1 - z2;
                const Mat4 p_i2w = pa.image_to_world();
                const Mat4 m_s = s_w2i * p_i2w, m_m = mask_w2i * p_i2w;
                int set_count = 0;
                for (int j = 0; j < pby; ++j) for (int i = 0; i < pbx; ++i) {
                    double xx = i, yy = j, zz = 0, xm = i, ym = j, zm = 0;
                    m_s.apply(xx, yy, zz);
                    m_m.apply(xm, ym, zm);
                    if (!(xx >= 0 && yy >= 0 && xx < attr.x && yy < attr.y)) continue;
                    if (!(xm >= 0 && ym >= 0 && zm >= 0 && xm < mask.a.x && ym < mask.a.y && zm < mask.a.z)) continue;
                    if (!(mask.at((int)xm, (int)ym, (int)zm) > 0)) continue;
                    const float v = (float)stack.at((int)xx, (int)yy, z);
                    if (v != 0.0f && v != -1.0f) ++set_count;
                }
                if (set_count > 1.0f / 3.0f * pby * pbx) {
                    float m[16];
                    p_i2w.to_float16(m); out.i2w.insert(out.i2w.end(), m, m + 16);
                    pa.world_to_image().to_float16(m); out.w2i.insert(out.w2i.end(), m, m + 16);
                    out.count++;
                    out.total_pixels += set_count;
                }
            }
        }
    }
    return out;
}

// PatchBasedVolume::generate2DSuperpixelPatches, include/patchBasedObject.cuh:433-797: SLICO labels per slice, then one
// patch per label: the label's bounding box grown to the (fixed 64 x 64, clamped to the stack) patch window, the label's
// pixels inside the mask dilated by spxExtend % of the larger box side (4-neighbourhood), and the result stored as the
// patch's char[64*64] mask.  Quirks kept: the largest label of a slice is never visited (`idxLbl < int(maxLbl)`, :510),
// labels with fewer than max(2, spx area / 4) masked pixels are dropped (:668-669), dilated pixels whose centre falls
// outside the slice or the mask grid stay in the mask (:708-721).
Patches generate_superpixel_patches(const Image& stack, const Image& mask, unsigned spx_x, unsigned spx_y, unsigned extend_pct, double thickness)
{
    Patches out;
    const ImageAttr& attr = stack.a;
    if ((int)spx_x > attr.x) { std::printf("WARNING: spxSize.x is bigger than imageSize.x - force spxSize.x = 0.5 * imageSize.x \n"); spx_x = attr.x / 2; }
    if ((int)spx_y > attr.y) { std::printf("WARNING: spxSize.y is bigger than imageSize.y - force spxSize.y = 0.5 * imageSize.y \n"); spx_y = attr.y / 2; }
    float vmin = std::numeric_limits<float>::max(), vmax = -std::numeric_limits<float>::max();
    for (double v : stack.v) { vmin = std::min(vmin, (float)v); vmax = std::max(vmax, (float)v); }
    std::printf("min: %f \nmax: %f \n", vmin, vmax);
    const float dilate_ratio = (float)extend_pct / 100.f;
    const int pbx = std::min(64, attr.x), pby = std::min(64, attr.y);
    out.pbx = pbx; out.pby = pby;
    std::printf("Dilation ratio: %d %% \nMax patch size - after dilation - %d x %d \n", extend_pct, pbx, pby);
    const Mat4 stack_i2w = attr.image_to_world();
    const Mat4 mask_w2i = mask.a.world_to_image();
    std::vector<float> slice((size_t)attr.x * attr.y);
    std::vector<int> cell((size_t)pbx * pby), grown;
    for (int z = 0; z < attr.z; ++z) {
        for (int y = 0; y < attr.y; ++y) for (int x = 0; x < attr.x; ++x) slice[(size_t)y * attr.x + x] = (float)stack.at(x, y, z);
        int n_labels = 0;
        const std::vector<int> labels = svr::slico_labels(slice.data(), attr.x, attr.y, vmin, vmax, spx_x, spx_y, &n_labels);
        out.labels.insert(out.labels.end(), labels.begin(), labels.end());
        int min_lbl = INT_MAX, max_lbl = INT_MIN;
        for (int l : labels) { min_lbl = std::min(min_lbl, l); max_lbl = std::max(max_lbl, l); }
        for (int lbl = min_lbl; lbl < max_lbl; ++lbl) {
            int x_min = INT_MAX, y_min = INT_MAX, x_max = INT_MIN, y_max = INT_MIN;
            for (int y = 0; y < attr.y; ++y) for (int x = 0; x < attr.x; ++x)
                if (labels[(size_t)y * attr.x + x] == lbl) { x_min = std::min(x_min, x); x_max = std::max(x_max, x); y_min = std::min(y_min, y); y_max = std::max(y_max, y); }
            if (x_min == INT_MAX) continue;
            const int box_x = x_max - x_min, box_y = y_max - y_min;
            const int diter = (int)(dilate_ratio * (float)std::max(box_x, box_y));
            const int ext_x = (int)std::round(((float)pbx - (float)box_x) / 2.f), ext_y = (int)std::round(((float)pby - (float)box_y) / 2.f);
            if (x_min - ext_x < 0) { x_max = pbx; x_min = 0; }
            else if (x_max + ext_x > attr.x) { x_max = attr.x; x_min = x_max - pbx; }
            else { x_min -= ext_x; x_max = x_min + pbx; }
            if (y_min - ext_y < 0) { y_max = pby; y_min = 0; }
            else if (y_max + ext_y > attr.y) { y_max = attr.y; y_min = y_max - pby; }
            else { y_min -= ext_y; y_max = y_min + pby; }

            // the patch image: GetRegion(x_min, y_min, z, x_max, y_max, z + 1) with dz = 2 * thickness (:590-591)
            ImageAttr pa = attr;
            pa.x = x_max - x_min; pa.y = y_max - y_min; pa.z = 1;
            pa.dz = thickness * 2;
            double cx = (x_min + x_max - 1) / 2.0, cy = (y_min + y_max - 1) / 2.0, cz = z;
            stack_i2w.apply(cx, cy, cz);
            pa.origin[0] = cx; pa.origin[1] = cy; pa.origin[2] = cz;
            const Mat4 p_i2w = pa.image_to_world();
            const Mat4 m_m = mask_w2i * p_i2w;
            // per patch pixel: 0 = outside the slice / the mask grid (stays what dilation makes of it), 1 = usable
            auto in_grids = [&](int i, int j, bool& masked) {
                const int xx = x_min + i, yy = y_min + j;           // the patch is an axis-aligned window of its slice
                double xm = i, ym = j, zm = 0;
                m_m.apply(xm, ym, zm);
                xm = std::round(xm); ym = std::round(ym); zm = std::round(zm);
                masked = false;
                if (!(xx >= 0 && yy >= 0 && xx < attr.x && yy < attr.y)) return false;
                if (!(xm >= 0 && ym >= 0 && zm >= 0 && xm < mask.a.x && ym < mask.a.y && zm < mask.a.z)) return false;
                masked = mask.at((int)xm, (int)ym, (int)zm) > 0;
                return true;
            };
            int set_count = 0;
            for (int j = 0; j < pby; ++j) for (int i = 0; i < pbx; ++i) {
                bool masked;
                int v = 0;
                if (i < pa.x && j < pa.y && in_grids(i, j, masked) && masked) v = labels[(size_t)(y_min + j) * attr.x + (x_min + i)] == lbl ? 1 : 0;
                cell[(size_t)j * pbx + i] = v;
                set_count += v;
            }
            if (set_count < 2) continue;
            if (set_count < 1.0f / 4.0f * spx_y * spx_x) continue;
            for (int it = 0; it < diter; ++it) {                      // dilatePatch (:347-368): 4-neighbourhood
                grown = cell;
                for (int j = 0; j < pby; ++j) for (int i = 0; i < pbx; ++i) {
                    if (cell[(size_t)j * pbx + i] != 1) continue;
                    if (i > 0) grown[(size_t)j * pbx + i - 1] = 1;
                    if (j > 0) grown[(size_t)(j - 1) * pbx + i] = 1;
                    if (i + 1 < pbx) grown[(size_t)j * pbx + i + 1] = 1;
                    if (j + 1 < pby) grown[(size_t)(j + 1) * pbx + i] = 1;
                }
                cell.swap(grown);
            }
            std::vector<char> spx(4096, '0');
            for (int j = 0; j < pby; ++j) for (int i = 0; i < pbx; ++i) {
                if (cell[(size_t)j * pbx + i] == 0) continue;
                bool masked;
                const bool inside = i < pa.x && j < pa.y && in_grids(i, j, masked);
                if (inside && !masked) continue;                       // inside the grids but outside the mask: -1
                spx[i + 64 * j] = '1';
                out.total_pixels++;
            }
            float m[16];
            p_i2w.to_float16(m); out.i2w.insert(out.i2w.end(), m, m + 16);
            pa.world_to_image().to_float16(m); out.w2i.insert(out.w2i.end(), m, m + 16);
            out.spx.insert(out.spx.end(), spx.begin(), spx.end());
            out.label_of_patch.push_back(z); out.label_of_patch.push_back(lbl);
            out.count++;
        }
    }
    return out;
}

// patchBasedPackageSplitter<T>::makePackageVolumes (patchBasedPackageSplitter.cpp:78-148): package l of a stack acquired in
// N interleaved packages holds its slices l, l + N, l + 2N, ... with N times the slice spacing, placed so that its first
// voxel coincides with voxel (0, 0, l) of the stack; transformation and thickness are inherited.
void split_into_packages(std::vector<Image>& stacks, std::vector<Rigid>& transformations, std::vector<double>& thickness,
                         const std::vector<int>& packages)
{
    std::vector<Image> out;
    std::vector<Rigid> out_t;
    std::vector<double> out_th;
    for (size_t s = 0; s < stacks.size(); ++s) {
        const Image& image = stacks[s];
        const int n = packages[s];
        const int pkg_z = image.a.z / n;
        for (int l = 0; l < n; ++l) {
            ImageAttr attr = image.a;
            attr.z = (pkg_z * n + l < image.a.z) ? pkg_z + 1 : pkg_z;
            attr.dz = image.a.dz * n;
            Image pkg(attr);
            for (int k = 0; k < attr.z; ++k) for (int j = 0; j < attr.y; ++j) for (int i = 0; i < attr.x; ++i) pkg.at(i, j, k) = image.at(i, j, k * n + l);
            double x = 0, y = 0, z = l, sx = 0, sy = 0, sz = 0;
            image.a.image_to_world().apply(x, y, z);
            pkg.a.image_to_world().apply(sx, sy, sz);
            pkg.a.origin[0] += x - sx; pkg.a.origin[1] += y - sy; pkg.a.origin[2] += z - sz;
            out.push_back(std::move(pkg));
            out_t.push_back(transformations[s]);
            out_th.push_back(thickness[s]);
        }
    }
    stacks.swap(out); transformations.swap(out_t); thickness.swap(out_th);
}

void write_raw(const std::string& path, const void* p, size_t bytes)
{
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f || std::fwrite(p, 1, bytes, f) != bytes) throw std::runtime_error("cannot write " + path);
    std::fclose(f);
}

void write_or_die(const std::string& path, const Image& img, bool f32)
{
    std::string err;
    if (!svr::write_nifti(path, img, f32, &err)) { std::cerr << "cannot write " << path << ": " << err << std::endl; exit(1); }
}

}  // namespace

int main(int argc, char** argv)
{
    Options o;
    bool help = false;
    if (!parse(argc, argv, o, help)) { std::cerr << std::endl; return EXIT_FAILURE; }
    if (help) { usage(); return EXIT_SUCCESS; }
    try {
        const int device = o.devices.empty() ? 0 : o.devices[0];
        std::cout << "Reconstructed volume name ... " << o.output << std::endl;
        std::cout << "Number of stacks ... " << o.input.size() << std::endl;
        std::cout << "NOTE: patch-to-volume and stack-to-stack registration are not part of this build; patches keep the -t transformation of their stack"
                  << std::endl;
        if (o.input.empty()) throw std::runtime_error("no input stacks (-i)");
        if (!o.transformation.empty() && o.transformation.size() != o.input.size())
            throw std::runtime_error("-t needs one transformation (or 'id') per input stack");
        const bool set_thickness = o.thickness.empty();
        if (!set_thickness && o.thickness.size() != o.input.size()) throw std::runtime_error("--thickness needs one value per input stack");

        // ---- stacks, thickness, transformations (patchBasedReconMain.cpp:190-258, setImageStacks :97-146) ---------------
        std::vector<Image> stacks;
        std::vector<double> thickness;
        std::vector<Rigid> transformations;
        int template_num = -1;
        for (size_t i = 0; i < o.input.size(); ++i) {
            Image stack;
            int frames = 1;
            std::string err;
            if (!svr::read_nifti(o.input[i], stack, &frames, &err)) throw std::runtime_error("cannot read " + o.input[i] + ": " + err);
            const double th = set_thickness ? stack.a.dz : o.thickness[i] / 2.0;
            std::cout << "Reading stack ... " << o.input[i] << " thickness " << th << " * 2 " << std::endl;
            Rigid t;
            if (!o.transformation.empty()) {
                if (o.transformation[i] == "id") { if (template_num < 0) template_num = (int)stacks.size(); }
                else if (!t.read_dof(o.transformation[i])) throw std::runtime_error("cannot read transformation " + o.transformation[i]);
            } else if (template_num < 0) template_num = 0;
            const size_t nvox = (size_t)stack.a.x * stack.a.y * stack.a.z;
            for (int f = 0; f < frames; ++f) {                        // 4D volumes are split into their frames
                Image s(stack.a);
                std::copy(stack.v.begin() + (size_t)f * nvox, stack.v.begin() + (size_t)(f + 1) * nvox, s.v.begin());
                stacks.push_back(std::move(s));
                thickness.push_back(th);
                transformations.push_back(t);
            }
        }
        if (template_num < 0) throw std::runtime_error("at least one stack needs the 'id' transformation (the template)");
        if (!o.packages.empty()) {
            if (o.packages.size() != stacks.size()) throw std::runtime_error("-p needs one package count per input stack");
            for (int n : o.packages) if (n < 1) throw std::runtime_error("-p: package counts are >= 1");
            std::printf("splitting volumes into Packages...\n");
            split_into_packages(stacks, transformations, thickness, o.packages);
            // as in the reference the template index is the one found before the split (setImageStacks, :97-146)
        }
        std::cout << (o.superpixel ? "superpixel-based ON" : "patch-based ON") << std::endl << "cuda_dev = " << device << std::endl;

        // ---- run(): mask, crop, resample (irtkPatchBasedReconstruction.cpp:197-266) ----------------------------------------
        Image mask;
        if (o.mask.empty()) {
            // CreateMaskFromOverlap (:968-1004): on the grid of the first stack, 1 where the voxel's world position lies
            // inside the grid of every stack (the stack transformations are not applied, as in the reference)
            std::cout << "creating mask from overlap " << std::endl;
            mask = Image(stacks[0].a, 0.0);
            const Mat4 i2w = mask.a.image_to_world();
            std::vector<Mat4> w2i;
            for (const Image& s : stacks) w2i.push_back(s.a.world_to_image());
            for (int z = 0; z < mask.a.z; ++z) for (int y = 0; y < mask.a.y; ++y) for (int x = 0; x < mask.a.x; ++x) {
                double wx = x, wy = y, wz = z;
                i2w.apply(wx, wy, wz);
                bool inside = true;
                for (size_t i = 0; i < stacks.size() && inside; ++i) {
                    double a = wx, b = wy, c = wz;
                    w2i[i].apply(a, b, c);
                    inside = a >= 0 && b >= 0 && c >= 0 && a < stacks[i].a.x && b < stacks[i].a.y && c < stacks[i].a.z;
                }
                if (inside) mask.at(x, y, z) = 1;
            }
        } else {
            std::string err;
            if (!svr::read_nifti(o.mask, mask, nullptr, &err)) throw std::runtime_error("cannot read " + o.mask + ": " + err);
            for (double& v : mask.v) v = ((unsigned)(signed char)(int)v == 0) ? 0 : 1;      // the mask is read as char (:201-207)
        }
        if (o.dilateMask) {
            std::printf("Dilate reconstruction mask %d-iterations \n", o.dilateMask);
            for (int i = 0; i < o.dilateMask; ++i) dilate26(mask);
        }
        svr::Reconstruction helper(device);            // TransformMask / CropImage of the shared host code; creates no device context
        helper.debug = o.debug;
        for (size_t i = 0; i < stacks.size(); ++i) {
            Image m = mask;
            helper.TransformMask(stacks[i], m, transformations[i]);
            helper.CropImage(stacks[i], m);
            if (o.debug) write_or_die("stack" + std::to_string(i) + ".nii.gz", stacks[i], true);
        }
        mask = resample_nn(mask, o.resolution);
        if (o.debug) write_or_die("mask.nii.gz", mask, true);
        double min_intensity, max_intensity;
        min_max(stacks, min_intensity, max_intensity);
        std::cout << "----------------------------------------------------------------------------------------------------------------" << std::endl;
        if (!o.noMatchIntensities) {
            std::printf("Intensities before -- min: %f max: %f \n", min_intensity, max_intensity);
            match_intensities(stacks, transformations, mask, o.debug);
            min_max(stacks, min_intensity, max_intensity);
            std::printf("After min: %f max: %f \n", min_intensity, max_intensity);
        }

        // ---- template and reconstruction mask (:296-308) --------------------------------------------------------------------
        Image recon;
        if (!o.existing_target.empty()) {
            std::string err;
            if (!svr::read_nifti(o.existing_target, recon, nullptr, &err)) throw std::runtime_error("cannot read " + o.existing_target + ": " + err);
        } else {
            std::cout << "Constructing volume with isotropic voxel size " << o.resolution << " mm." << std::endl;
            recon = Image(svr::resampled_attr(stacks[template_num].a, o.resolution, o.resolution, o.resolution), 0.0);
        }
        Image reconmask = mask;
        helper.TransformMask(recon, reconmask, transformations[template_num]);

        // ---- patches (:391-399, patchBasedObject.cuh:176-342) ---------------------------------------------------------------
        int pbx = (int)o.patchSize[0], pby = (int)o.patchSize[1];
        std::vector<char> spx_masks;
        std::vector<Patches> patches(stacks.size());
        std::vector<int> pps(stacks.size());
        std::vector<float> stack_dims(3 * stacks.size()), i2w, w2i, T, Ti;
        for (size_t i = 0; i < stacks.size(); ++i) {
            std::cout << "stack [" << i << "] -------------------------- " << std::endl;
            std::printf("Thickness %f \n", thickness[i]);
            if (o.superpixel) {
                patches[i] = generate_superpixel_patches(stacks[i], mask, o.patchSize[0], o.patchSize[1], o.patchStride[0], thickness[i]);
                if (i > 0 && (patches[i].pbx != pbx || patches[i].pby != pby))
                    throw std::runtime_error("superpixel mode: the stacks clamp the 64 x 64 patch window differently (a stack is smaller than 64 pixels)");
                pbx = patches[i].pbx; pby = patches[i].pby;
                spx_masks.insert(spx_masks.end(), patches[i].spx.begin(), patches[i].spx.end());
            } else
                patches[i] = generate_2d_patches(stacks[i], mask, pbx, pby, (int)o.patchStride[0], (int)o.patchStride[1], thickness[i]);
            std::printf("m_patches GPU size: %d ... \n", patches[i].count);
            pps[i] = patches[i].count;
            stack_dims[3 * i] = (float)stacks[i].a.dx; stack_dims[3 * i + 1] = (float)stacks[i].a.dy; stack_dims[3 * i + 2] = (float)stacks[i].a.dz;
            i2w.insert(i2w.end(), patches[i].i2w.begin(), patches[i].i2w.end());
            w2i.insert(w2i.end(), patches[i].w2i.begin(), patches[i].w2i.end());
            float m[16], mi[16];
            const Mat4 t = transformations[i].matrix();
            t.to_float16(m); t.inverse().to_float16(mi);
            for (int p = 0; p < patches[i].count; ++p) { T.insert(T.end(), m, m + 16); Ti.insert(Ti.end(), mi, mi + 16); }
        }
        const size_t n_patches = i2w.size() / 16;
        if (n_patches == 0) throw std::runtime_error("no patch passed the coverage rule: check the mask and the transformations");

        float rw2i[16], ri2w[16];
        recon.a.world_to_image().to_float16(rw2i);
        recon.a.image_to_world().to_float16(ri2w);
        std::vector<signed char> mask8(reconmask.n());
        for (size_t i = 0; i < mask8.size(); ++i) mask8[i] = (signed char)reconmask.v[i];

        if (!o.dump_patches.empty()) {
            const std::string& d = o.dump_patches;
            std::ofstream idx(d + "/index.txt");
            idx.precision(17);
            idx << "stacks " << stacks.size() << "\npatches " << n_patches << "\npbx " << pbx << "\npby " << pby << "\nvx " << recon.a.x << "\nvy "
                << recon.a.y << "\nvz " << recon.a.z << "\nmx " << mask.a.x << "\nmy " << mask.a.y << "\nmz " << mask.a.z << "\nmin " << min_intensity
                << "\nmax " << max_intensity << "\n";
            write_raw(d + "/per_stack.i32", pps.data(), pps.size() * 4);
            write_raw(d + "/i2w.f32", i2w.data(), i2w.size() * 4);
            write_raw(d + "/w2i.f32", w2i.data(), w2i.size() * 4);
            write_raw(d + "/T.f32", T.data(), T.size() * 4);
            if (o.superpixel) write_raw(d + "/spx.i8", spx_masks.data(), spx_masks.size());
            write_raw(d + "/recon_w2i.f32", rw2i, 64);
            write_raw(d + "/recon_i2w.f32", ri2w, 64);
            write_raw(d + "/recon_mask.i8", mask8.data(), mask8.size());
            std::vector<double> ma = { (double)mask.a.x, (double)mask.a.y, (double)mask.a.z, mask.a.dx, mask.a.dy, mask.a.dz };
            for (int q = 0; q < 3; ++q) ma.push_back(mask.a.origin[q]);
            for (int q = 0; q < 3; ++q) ma.push_back(mask.a.xaxis[q]);
            for (int q = 0; q < 3; ++q) ma.push_back(mask.a.yaxis[q]);
            for (int q = 0; q < 3; ++q) ma.push_back(mask.a.zaxis[q]);
            write_raw(d + "/mask_attr.f64", ma.data(), ma.size() * 8);
            write_raw(d + "/mask.f64", mask.v.data(), mask.v.size() * 8);
            for (size_t i = 0; i < stacks.size(); ++i) {
                const ImageAttr& a = stacks[i].a;
                std::vector<double> sa = { (double)a.x, (double)a.y, (double)a.z, a.dx, a.dy, a.dz };
                for (int q = 0; q < 3; ++q) sa.push_back(a.origin[q]);
                for (int q = 0; q < 3; ++q) sa.push_back(a.xaxis[q]);
                for (int q = 0; q < 3; ++q) sa.push_back(a.yaxis[q]);
                for (int q = 0; q < 3; ++q) sa.push_back(a.zaxis[q]);
                sa.push_back(thickness[i]);
                write_raw(d + "/stack" + std::to_string(i) + "_attr.f64", sa.data(), sa.size() * 8);
                write_raw(d + "/stack" + std::to_string(i) + ".f64", stacks[i].v.data(), stacks[i].v.size() * 8);
                if (o.superpixel) {
                    write_raw(d + "/labels" + std::to_string(i) + ".i32", patches[i].labels.data(), patches[i].labels.size() * 4);
                    write_raw(d + "/label_of_patch" + std::to_string(i) + ".i32", patches[i].label_of_patch.data(), patches[i].label_of_patch.size() * 4);
                }
            }
            std::cout << "patches written to " << d << std::endl;
            return EXIT_SUCCESS;
        }

        // ---- device set-up (:307-430) -------------------------------------------------------------------------------------------
        std::cout << "m_cuda_device " << device << std::endl;
        svr_context* c = nullptr;
        ck(pvr_create(&c, device), "pvr_create");
        g_ctx = c;
        ck(pvr_recon_init(c, recon.a.x, recon.a.y, recon.a.z, (float)recon.a.dx, (float)recon.a.dy, (float)recon.a.dz, rw2i, ri2w), "ReconVolume::init");
        ck(pvr_recon_set_mask(c, mask8.data()), "ReconVolume::setMask");
        if (!o.existing_target.empty()) {
            std::vector<float> r(recon.n());
            for (size_t i = 0; i < r.size(); ++i) r[i] = (float)recon.v[i];
            ck(pvr_recon_copy_from_host(c, r.data()), "ReconVolume::copyFromHost");
        }
        ck(pvr_patches_init(c, pbx, pby, (int)stacks.size(), pps.data(), stack_dims.data()), "PatchBasedVolume::init");
        ck(pvr_patches_set_matrices(c, i2w.data(), w2i.data(), T.data(), Ti.data()), "patch matrices");
        if (o.superpixel) ck(pvr_patches_set_spx_masks(c, spx_masks.data(), 1), "superpixel masks");
        {
            ImageAttr pa;                                   // PSF image: PSF_SIZE^3 voxels of the reconstruction's size (:404-419)
            pa.x = pa.y = pa.z = 128;
            pa.dx = pa.dy = pa.dz = o.resolution;
            const int size[3] = { 128, 128, 128 };
            float pi2w[16];
            pa.image_to_world().to_float16(pi2w);
            ck(pvr_set_psf(c, size, pi2w, 1.0f), "PointSpreadFunction");
        }
        for (size_t i = 0; i < stacks.size(); ++i) {       // initPatchBasedRecon_gpu: the device cuts the patches out of the stack
            std::vector<float> data(stacks[i].n());
            for (size_t q = 0; q < data.size(); ++q) data[q] = (float)stacks[i].v[q];
            float sw2i[16];
            stacks[i].a.world_to_image().to_float16(sw2i);
            ck(pvr_init_patch_based_recon(c, (int)i, data.data(), stacks[i].a.x, stacks[i].a.y, stacks[i].a.z, sw2i), "initPatchBasedRecon_gpu");
        }

        // ---- the iteration loop (:445-567); constants of patchBasedSuperresolution_gpu.cu:293-295 and
        // patchBasedRobustStatistics_gpu.cu:877 ----------------------------------------------------------------------------------------
        const float delta = 1.0f, lambda = 0.1f, step = 0.0001f;
        const float alpha = (0.05f / lambda) * delta * delta;
        std::vector<float> pot(n_patches), scales(n_patches), weights(n_patches), used(n_patches), out(recon.n());
        float sigma = 0, mix = 0, m = 0;
        float state5[5] = { 0.025f, 0.9f, 0, 0, 0 };     // sigma_s, mix_s, mean_s, mean_s2, sigma_s2
        auto estep = [&]() {
            ck(pvr_rs_estep_device(c, m, sigma, mix, pot.data()), "EStep");
            ck(pvr_rs_get_scales_weights(c, scales.data(), weights.data()), "EStep");
            if (pvr_host_patch_em((int)pps.size(), pps.data(), pot.data(), scales.data(), weights.data(), step, state5, used.data()) != 0)
                throw std::runtime_error("pvr_host_patch_em: bad argument");
            ck(pvr_rs_set_scales_weights(c, scales.data(), weights.data()), "EStep");
        };
        const auto t0 = std::chrono::steady_clock::now();
        for (int iter = 0; iter <= o.iterations; ++iter) {
            std::printf("iteration %d \n", iter);
            ck(pvr_rs_initialize_em_values(c), "initializeEMValues");
            ck(pvr_recon_reset(c), "reset");
            ck(pvr_psf_reconstruction(c), "patchBasedPSFReconstruction_gpu");
            ck(pvr_recon_equalize(c), "equalize");
            ck(pvr_simulate_patches(c), "patchBasedSimulatePatches_gpu");
            ck(pvr_rs_initialize_robust_statistics(c, &sigma), "InitializeRobustStatistics");
            state5[0] = 0.025f; mix = 0.9f; state5[1] = 0.9f;
            m = 1.0f / (2.1f * (float)max_intensity - 1.9f * (float)min_intensity);
            estep();
            for (int i = 0; i < o.sr_iterations; ++i) {
                ck(pvr_rs_scale(c, scales.data()), "Scale");
                ck(pvr_recon_reset_addon_cmap(c), "resetAddonCmap");
                ck(pvr_superresolution_run(c), "superresolution.run");
                ck(pvr_superresolution_regularize(c, 0, alpha, (float)min_intensity, (float)max_intensity, delta, lambda), "regularize");
                ck(pvr_simulate_patches(c), "patchBasedSimulatePatches_gpu");
                ck(pvr_rs_mstep(c, i + 1, step, &sigma, &mix, &m), "MStep");
                estep();
            }
            ck(pvr_recon_copy_to_host(c, out.data()), "copyToHost");
            Image reconimage(recon.a);
            for (size_t q = 0; q < out.size(); ++q) reconimage.v[q] = out[q];
            char buffer[256];
            std::snprintf(buffer, sizeof(buffer), "reconimage%i_%i_%i.nii.gz", iter, (int)o.patchSize[0], (int)o.patchStride[0]);
            write_or_die(buffer, reconimage, true);
            std::cout << "----------------------------------------------------------------------------------------------------------------" << std::endl;
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("reconstruction took %f s (%d passes, %zu patches)\n", secs, o.iterations + 1, n_patches);
        ck(pvr_recon_copy_to_host(c, out.data()), "copyToHost");
        for (size_t q = 0; q < out.size(); ++q) recon.v[q] = out[q];
        write_or_die(o.output, recon, true);
        svr_destroy(c);
    } catch (const std::exception& e) {
        std::cerr << "ERROR: " << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
}
