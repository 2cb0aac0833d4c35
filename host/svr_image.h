// svr_image.h -- the slice of the IRTK substrate the SVRreconstructionGPU drop-in needs: 4x4 double matrices, image
// attributes <-> image/world matrices, rigid transformations (+ .dof I/O), a double-valued image, NIfTI-1 I/O.
// Rules restated from the reference (paths relative to source/IRTKSimple2/):
//   image <-> world          image++/src/irtkBaseImage.cc:79-147
//   GetRegion                image++/src/irtkGenericImage.cc:570-625
//   NIfTI -> attributes      image++/src/irtkFileNIFTIToImage.cc:227-373   (qform preferred over sform)
//   NIfTI write              image++/src/irtkImageToFileNIFTI.cc:66-147     (irtkRealImage = double -> FLOAT64)
//   rigid parameters         packages/transformation/src/irtkRigidTransformation.cc:26-149
//   .dof files               packages/transformation/src/irtkRigidTransformation.cc:392-451 (big endian, magic 815007)
//   Gaussian blurring        image++/src/irtkGaussianBlurring.cc:40-124
//   resampling grid          image++/src/irtkResampling.cc:92-131
//   image transformation     packages/transformation/src/irtkImageTransformation.cc:200-300 (nearest neighbour)
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace svr {

struct Mat4 {
    double m[4][4];
    static Mat4 identity() { Mat4 r{}; for (int i = 0; i < 4; ++i) r.m[i][i] = 1; return r; }
    Mat4 operator*(const Mat4& b) const {
        Mat4 r{};
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double s = 0; for (int k = 0; k < 4; ++k) s += m[i][k] * b.m[k][j]; r.m[i][j] = s; }
        return r;
    }
    void apply(double& x, double& y, double& z) const {
        const double a = m[0][0] * x + m[0][1] * y + m[0][2] * z + m[0][3];
        const double b = m[1][0] * x + m[1][1] * y + m[1][2] * z + m[1][3];
        const double c = m[2][0] * x + m[2][1] * y + m[2][2] * z + m[2][3];
        x = a; y = b; z = c;
    }
    Mat4 inverse() const;                       // general 4x4 inverse (Gauss-Jordan, double)
    void to_float16(float* out) const { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = (float)m[i][j]; }
    static Mat4 from_float16(const float* in) { Mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = in[4 * i + j]; return r; }
};

struct ImageAttr {                               // irtkImageAttributes
    int x = 0, y = 0, z = 0;
    double dx = 1, dy = 1, dz = 1;
    double origin[3] = { 0, 0, 0 };               // world position of the image CENTRE
    double xaxis[3] = { 1, 0, 0 }, yaxis[3] = { 0, 1, 0 }, zaxis[3] = { 0, 0, 1 };
    Mat4 image_to_world() const;
    Mat4 world_to_image() const;
};

struct Image {                                    // irtkRealImage (double voxels), x fastest
    ImageAttr a;
    std::vector<double> v;
    Image() {}
    explicit Image(const ImageAttr& attr, double fill = 0.0) : a(attr), v((size_t)attr.x * attr.y * attr.z, fill) {}
    size_t n() const { return v.size(); }
    double& at(int i, int j, int k) { return v[(size_t)k * a.x * a.y + (size_t)j * a.x + i]; }
    double at(int i, int j, int k) const { return v[(size_t)k * a.x * a.y + (size_t)j * a.x + i]; }
    Image get_region(int x1, int y1, int z1, int x2, int y2, int z2) const;
    void gaussian_blur(double sigma_mm);
};

struct Rigid {                                    // irtkRigidTransformation: tx ty tz [mm], rx ry rz [degrees]
    double p[6] = { 0, 0, 0, 0, 0, 0 };
    Mat4 matrix() const;
    static Rigid from_matrix(const Mat4& m);
    void invert() { *this = from_matrix(matrix().inverse()); }
    bool read_dof(const std::string& path);       // returns false on failure
    bool write_dof(const std::string& path) const;
};

// irtkResampling output grid: new_n = int(n * old / new) (min 1); origin and axes unchanged
ImageAttr resampled_attr(const ImageAttr& a, double dx, double dy, double dz);
// irtkResamplingWithPadding output grid (image++/src/irtkResamplingWithPadding.cc:203-262): new_n = round(n * old / new); a
// dimension that would fall below 1 stays 1 voxel of the OLD size; origin and axes unchanged
ImageAttr resampled_attr_with_padding(const ImageAttr& a, double dx, double dy, double dz);
// irtkImageTransformation::Run with a nearest-neighbour interpolator, target padding -1, source padding 0:
// resamples `source` onto the grid of `target` (whose values select the voxels to fill) through `t`
void transform_image_nn(const Image& source, const Rigid& t, Image& target, double target_padding = -1, double source_padding = 0);

// NIfTI-1 (.nii / .nii.gz), single file.  Reads uint8/int8/int16/uint16/int32/uint32/float32/float64 with scl_slope/inter;
// 4D files return the number of frames in `frames` and all frames concatenated in the image data (z = dim3 * frames).
bool read_nifti(const std::string& path, Image& out, int* frames, std::string* err);
bool write_nifti(const std::string& path, const Image& img, bool as_float32, std::string* err);

}  // namespace svr
