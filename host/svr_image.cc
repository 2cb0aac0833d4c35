#include "svr_image.h"
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstring>

namespace svr {

Mat4 Mat4::inverse() const
{
    double a[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = m[i][j]; a[i][4 + j] = i == j; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        if (piv != c) for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
        const double d = a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != c) { const double f = a[r][c]; if (f != 0) for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    Mat4 r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = a[i][4 + j];
    return r;
}

Mat4 ImageAttr::image_to_world() const
{   // irtkBaseImage.cc:79-112: T(origin) * R(axes as columns) * S(dx,dy,dz) * T(-(n-1)/2)
    Mat4 t1 = Mat4::identity(), sc = Mat4::identity(), rot = Mat4::identity(), t2 = Mat4::identity();
    t1.m[0][3] = -(x - 1) / 2.0; t1.m[1][3] = -(y - 1) / 2.0; t1.m[2][3] = -(z - 1) / 2.0;
    sc.m[0][0] = dx; sc.m[1][1] = dy; sc.m[2][2] = dz;
    for (int i = 0; i < 3; ++i) { rot.m[i][0] = xaxis[i]; rot.m[i][1] = yaxis[i]; rot.m[i][2] = zaxis[i]; t2.m[i][3] = origin[i]; }
    return t2 * (rot * (sc * t1));
}

Mat4 ImageAttr::world_to_image() const
{   // irtkBaseImage.cc:114-147
    Mat4 t1 = Mat4::identity(), sc = Mat4::identity(), rot = Mat4::identity(), t2 = Mat4::identity();
    for (int i = 0; i < 3; ++i) { t1.m[i][3] = -origin[i]; rot.m[0][i] = xaxis[i]; rot.m[1][i] = yaxis[i]; rot.m[2][i] = zaxis[i]; }
    sc.m[0][0] = 1.0 / dx; sc.m[1][1] = 1.0 / dy; sc.m[2][2] = 1.0 / dz;
    t2.m[0][3] = (x - 1) / 2.0; t2.m[1][3] = (y - 1) / 2.0; t2.m[2][3] = (z - 1) / 2.0;
    return t2 * (sc * (rot * t1));
}

Image Image::get_region(int x1, int y1, int z1, int x2, int y2, int z2) const
{   // irtkGenericImage::GetRegion: axes / voxel size kept, origin shifted so world positions are preserved
    ImageAttr b = a;
    b.x = x2 - x1; b.y = y2 - y1; b.z = z2 - z1;
    b.origin[0] = b.origin[1] = b.origin[2] = 0;
    Image r(b);
    double px = x1, py = y1, pz = z1;             // first voxel of the ROI in the original image
    a.image_to_world().apply(px, py, pz);
    double qx = 0, qy = 0, qz = 0;                // first voxel of the new image
    b.image_to_world().apply(qx, qy, qz);
    r.a.origin[0] = px - qx; r.a.origin[1] = py - qy; r.a.origin[2] = pz - qz;
    for (int k = z1; k < z2; ++k) for (int j = y1; j < y2; ++j) for (int i = x1; i < x2; ++i) r.at(i - x1, j - y1, k - z1) = at(i, j, k);
    return r;
}

static void blur_axis(Image& img, int axis, double sigma_vox)
{   // irtkConvolution_1D with normalisation: kernel of 2*round(4 sigma)+1 taps, weights renormalised over the in-image taps
    const int half = (int)std::lround(4 * sigma_vox);
    if (half < 1) return;
    std::vector<double> k(2 * half + 1);
    for (int i = -half; i <= half; ++i) k[i + half] = std::exp(-(double)i * i / (2 * sigma_vox * sigma_vox));
    const int n[3] = { img.a.x, img.a.y, img.a.z };
    const size_t stride[3] = { 1, (size_t)img.a.x, (size_t)img.a.x * img.a.y };
    const int len = n[axis];
    if (len == 1) return;
    std::vector<double> line(len), out(len);
    const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    for (int u = 0; u < n[a1]; ++u) for (int w = 0; w < n[a2]; ++w) {
        const size_t base = u * stride[a1] + w * stride[a2];
        for (int i = 0; i < len; ++i) line[i] = img.v[base + i * stride[axis]];
        for (int i = 0; i < len; ++i) {
            double s = 0, ws = 0;
            for (int t = -half; t <= half; ++t) { const int j = i + t; if (j < 0 || j >= len) continue; s += k[t + half] * line[j]; ws += k[t + half]; }
            out[i] = ws > 0 ? s / ws : line[i];
        }
        for (int i = 0; i < len; ++i) img.v[base + i * stride[axis]] = out[i];
    }
}

void Image::gaussian_blur(double sigma_mm)
{
    blur_axis(*this, 0, sigma_mm / a.dx);
    blur_axis(*this, 1, sigma_mm / a.dy);
    blur_axis(*this, 2, sigma_mm / a.dz);
}

Mat4 Rigid::matrix() const
{   // irtkRigidTransformation::UpdateMatrix
    const double d2r = M_PI / 180.0;
    const double cx = std::cos(p[3] * d2r), cy = std::cos(p[4] * d2r), cz = std::cos(p[5] * d2r);
    const double sx = std::sin(p[3] * d2r), sy = std::sin(p[4] * d2r), sz = std::sin(p[5] * d2r);
    Mat4 r = Mat4::identity();
    r.m[0][0] = cy * cz; r.m[0][1] = cy * sz; r.m[0][2] = -sy; r.m[0][3] = p[0];
    r.m[1][0] = sx * sy * cz - cx * sz; r.m[1][1] = sx * sy * sz + cx * cz; r.m[1][2] = sx * cy; r.m[1][3] = p[1];
    r.m[2][0] = cx * sy * cz + sx * sz; r.m[2][1] = cx * sy * sz - sx * cz; r.m[2][2] = cx * cy; r.m[2][3] = p[2];
    return r;
}

Rigid Rigid::from_matrix(const Mat4& m)
{   // irtkRigidTransformation::Matrix2Parameters
    Rigid r;
    const double TOL = 0.000001;
    r.p[0] = m.m[0][3]; r.p[1] = m.m[1][3]; r.p[2] = m.m[2][3];
    const double tmp = std::asin(-1 * m.m[0][2]);
    if (std::fabs(std::cos(tmp)) > TOL) {
        r.p[3] = std::atan2(m.m[1][2], m.m[2][2]); r.p[4] = tmp; r.p[5] = std::atan2(m.m[0][1], m.m[0][0]);
    } else {
        r.p[3] = std::atan2(-1.0 * m.m[0][2] * m.m[1][0], -1.0 * m.m[0][2] * m.m[2][0]); r.p[4] = tmp; r.p[5] = 0;
    }
    for (int i = 3; i < 6; ++i) r.p[i] *= 180.0 / M_PI;
    return r;
}

static uint32_t be32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xff00) | ((v << 8) & 0xff0000) | (v << 24); }
static void be64(const void* src, void* dst) { const unsigned char* s = (const unsigned char*)src; unsigned char* d = (unsigned char*)dst; for (int i = 0; i < 8; ++i) d[i] = s[7 - i]; }

bool Rigid::read_dof(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    uint32_t h[3];
    bool ok = fread(h, 4, 3, f) == 3 && be32(h[0]) == 815007u && be32(h[2]) >= 6;
    for (int i = 0; ok && i < 6; ++i) { unsigned char b[8]; ok = fread(b, 1, 8, f) == 8; if (ok) be64(b, &p[i]); }
    fclose(f);
    return ok;
}

bool Rigid::write_dof(const std::string& path) const
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const uint32_t h[3] = { be32(815007u), be32(2u), be32(6u) };     // magic, IRTKTRANSFORMATION_RIGID, 6 dofs
    fwrite(h, 4, 3, f);
    for (int i = 0; i < 6; ++i) { unsigned char b[8]; be64(&p[i], b); fwrite(b, 1, 8, f); }
    fclose(f);
    return true;
}

ImageAttr resampled_attr(const ImageAttr& a, double dx, double dy, double dz)
{
    ImageAttr r = a;
    r.x = std::max(1, (int)(a.x * a.dx / dx)); r.y = std::max(1, (int)(a.y * a.dy / dy)); r.z = std::max(1, (int)(a.z * a.dz / dz));
    r.dx = dx; r.dy = dy; r.dz = dz;
    return r;
}

static int irtk_round(double x) { return x > 0 ? int(x + 0.5) : int(x - 0.5); }     // common++/include/irtkCommon.h:85-88
ImageAttr resampled_attr_with_padding(const ImageAttr& a, double dx, double dy, double dz)
{
    ImageAttr r = a;
    const int nx = irtk_round(a.x * a.dx / dx), ny = irtk_round(a.y * a.dy / dy), nz = irtk_round(a.z * a.dz / dz);
    if (nx < 1) { r.x = 1; r.dx = a.dx; } else { r.x = nx; r.dx = dx; }
    if (ny < 1) { r.y = 1; r.dy = a.dy; } else { r.y = ny; r.dy = dy; }
    if (nz < 1) { r.z = 1; r.dz = a.dz; } else { r.z = nz; r.dz = dz; }
    return r;
}

void transform_image_nn(const Image& source, const Rigid& t, Image& target, double target_padding, double source_padding)
{
    const Mat4 m = source.a.world_to_image() * (t.matrix() * target.a.image_to_world());
    for (int k = 0; k < target.a.z; ++k) for (int j = 0; j < target.a.y; ++j) for (int i = 0; i < target.a.x; ++i) {
        double& o = target.at(i, j, k);
        if (!(o > target_padding)) { o = source_padding; continue; }
        double x = i, y = j, z = k;
        m.apply(x, y, z);
        if (x > -0.5 && x < source.a.x - 0.5 && y > -0.5 && y < source.a.y - 0.5 && z > -0.5 && z < source.a.z - 0.5)
            o = source.at((int)std::lround(x), (int)std::lround(y), (int)std::lround(z));
        else o = source_padding;
    }
}

// ---------------------------------------------------------------------------------------------------------------
#pragma pack(push, 1)
struct Nifti1Header {
    int32_t sizeof_hdr; char data_type[10]; char db_name[18]; int32_t extents; int16_t session_error; char regular; char dim_info;
    int16_t dim[8]; float intent_p1, intent_p2, intent_p3; int16_t intent_code; int16_t datatype; int16_t bitpix; int16_t slice_start;
    float pixdim[8]; float vox_offset; float scl_slope; float scl_inter; int16_t slice_end; char slice_code; char xyzt_units;
    float cal_max, cal_min, slice_duration, toffset; int32_t glmax, glmin; char descrip[80]; char aux_file[24];
    int16_t qform_code, sform_code; float quatern_b, quatern_c, quatern_d, qoffset_x, qoffset_y, qoffset_z;
    float srow_x[4], srow_y[4], srow_z[4]; char intent_name[16]; char magic[4];
};
#pragma pack(pop)
static_assert(sizeof(Nifti1Header) == 348, "NIfTI-1 header must be 348 bytes");

static bool read_all(const std::string& path, std::vector<unsigned char>& buf)
{
    gzFile g = gzopen(path.c_str(), "rb");      // transparently reads uncompressed files too
    if (!g) return false;
    unsigned char tmp[1 << 16];
    int n;
    while ((n = gzread(g, tmp, sizeof tmp)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    gzclose(g);
    return n == 0;
}

template <class T> static void bswap(T& v) { unsigned char* b = (unsigned char*)&v; std::reverse(b, b + sizeof(T)); }

bool read_nifti(const std::string& path, Image& out, int* frames, std::string* err)
{
    std::vector<unsigned char> buf;
    if (!read_all(path, buf) || buf.size() < 352) { if (err) *err = "cannot read " + path; return false; }
    Nifti1Header h;
    memcpy(&h, buf.data(), 348);
    const bool swap = h.sizeof_hdr != 348;
    if (swap) {
        bswap(h.sizeof_hdr); for (auto& d : h.dim) bswap(d); bswap(h.datatype); bswap(h.bitpix); for (auto& p : h.pixdim) bswap(p);
        bswap(h.vox_offset); bswap(h.scl_slope); bswap(h.scl_inter); bswap(h.qform_code); bswap(h.sform_code);
        bswap(h.quatern_b); bswap(h.quatern_c); bswap(h.quatern_d); bswap(h.qoffset_x); bswap(h.qoffset_y); bswap(h.qoffset_z);
        for (int i = 0; i < 4; ++i) { bswap(h.srow_x[i]); bswap(h.srow_y[i]); bswap(h.srow_z[i]); }
    }
    if (h.sizeof_hdr != 348 || h.dim[0] < 1 || h.dim[0] > 7) { if (err) *err = path + ": not a NIfTI-1 file"; return false; }
    ImageAttr a;
    a.x = h.dim[1]; a.y = h.dim[0] >= 2 ? h.dim[2] : 1; a.z = h.dim[0] >= 3 ? h.dim[3] : 1;
    a.dx = std::fabs(h.pixdim[1]); a.dy = std::fabs(h.pixdim[2]); a.dz = std::fabs(h.pixdim[3]);
    const int t = h.dim[0] == 4 ? h.dim[4] : (h.dim[0] == 5 ? h.dim[5] : 1);
    double M[4][4] = { { 0 } };
    M[3][3] = 1;
    if (h.qform_code > 0) {              // nifti_quatern_to_mat44
        double b = h.quatern_b, c = h.quatern_c, d = h.quatern_d, aa = 1.0 - (b * b + c * c + d * d);
        if (aa < 1.e-7) { aa = 1.0 / std::sqrt(b * b + c * c + d * d); b *= aa; c *= aa; d *= aa; aa = 0.0; } else aa = std::sqrt(aa);
        const double xd = h.pixdim[1] > 0 ? h.pixdim[1] : 1, yd = h.pixdim[2] > 0 ? h.pixdim[2] : 1;
        double zd = h.pixdim[3] > 0 ? h.pixdim[3] : 1;
        if (h.pixdim[0] < 0) zd = -zd;
        M[0][0] = (aa * aa + b * b - c * c - d * d) * xd; M[0][1] = 2 * (b * c - aa * d) * yd; M[0][2] = 2 * (b * d + aa * c) * zd;
        M[1][0] = 2 * (b * c + aa * d) * xd; M[1][1] = (aa * aa + c * c - b * b - d * d) * yd; M[1][2] = 2 * (c * d - aa * b) * zd;
        M[2][0] = 2 * (b * d - aa * c) * xd; M[2][1] = 2 * (c * d + aa * b) * yd; M[2][2] = (aa * aa + d * d - c * c - b * b) * zd;
        M[0][3] = h.qoffset_x; M[1][3] = h.qoffset_y; M[2][3] = h.qoffset_z;
    } else if (h.sform_code > 0) {
        for (int j = 0; j < 4; ++j) { M[0][j] = h.srow_x[j]; M[1][j] = h.srow_y[j]; M[2][j] = h.srow_z[j]; }
    } else {
        M[0][0] = -a.dx; M[1][1] = a.dy; M[2][2] = a.dz;
        M[0][3] = a.dx * (a.x - 1) / 2.0; M[1][3] = -a.dy * (a.y - 1) / 2.0; M[2][3] = -a.dz * (a.z - 1) / 2.0;
    }
    for (int i = 0; i < 3; ++i) { a.xaxis[i] = M[i][0] / a.dx; a.yaxis[i] = M[i][1] / a.dy; a.zaxis[i] = M[i][2] / a.dz; }
    const double c[3] = { (a.x - 1) / 2.0, (a.y - 1) / 2.0, (a.z - 1) / 2.0 };      // origin = D * centre
    for (int i = 0; i < 3; ++i) a.origin[i] = M[i][0] * c[0] + M[i][1] * c[1] + M[i][2] * c[2] + M[i][3];
    const size_t nvox = (size_t)a.x * a.y * a.z * t;
    const size_t off = (size_t)h.vox_offset;
    const int bytes = h.bitpix / 8;
    if (buf.size() < off + nvox * bytes) { if (err) *err = path + ": truncated image data"; return false; }
    ImageAttr full = a; full.z = a.z * t;
    out = Image(full);
    out.a = a; out.a.z = a.z * t;                 // frames concatenated along z; the caller splits
    const double slope = h.scl_slope != 0 ? h.scl_slope : 1.0, inter = h.scl_inter;
    const unsigned char* p = buf.data() + off;
    for (size_t i = 0; i < nvox; ++i) {
        double val = 0;
#define RD(T) { T q; memcpy(&q, p + i * sizeof(T), sizeof(T)); if (swap) bswap(q); val = (double)q; }
        switch (h.datatype) {
        case 2: RD(uint8_t) break; case 256: RD(int8_t) break; case 4: RD(int16_t) break; case 512: RD(uint16_t) break;
        case 8: RD(int32_t) break; case 768: RD(uint32_t) break; case 16: RD(float) break; case 64: RD(double) break;
        default: if (err) *err = path + ": unsupported NIfTI datatype"; return false;
        }
#undef RD
        out.v[i] = val * slope + inter;
    }
    if (frames) *frames = t;
    return true;
}

bool write_nifti(const std::string& path, const Image& img, bool as_float32, std::string* err)
{
    Nifti1Header h;
    memset(&h, 0, sizeof h);
    h.sizeof_hdr = 348; h.regular = 'r';
    h.dim[0] = 3; h.dim[1] = (int16_t)img.a.x; h.dim[2] = (int16_t)img.a.y; h.dim[3] = (int16_t)img.a.z; h.dim[4] = h.dim[5] = h.dim[6] = h.dim[7] = 1;
    h.datatype = as_float32 ? 16 : 64; h.bitpix = as_float32 ? 32 : 64;
    h.pixdim[1] = (float)img.a.dx; h.pixdim[2] = (float)img.a.dy; h.pixdim[3] = (float)img.a.dz; h.pixdim[4] = 1;
    h.vox_offset = 352; h.scl_slope = 1; h.scl_inter = 0; h.xyzt_units = 2 | 8;     // mm, s
    const Mat4 m = img.a.image_to_world();
    // qform from the rotation part (nifti_mat44_to_quatern), sform = the full matrix
    double r[3][3];
    const double d[3] = { img.a.dx, img.a.dy, img.a.dz };
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = m.m[i][j] / d[j];
    double det = r[0][0] * (r[1][1] * r[2][2] - r[1][2] * r[2][1]) - r[0][1] * (r[1][0] * r[2][2] - r[1][2] * r[2][0]) + r[0][2] * (r[1][0] * r[2][1] - r[1][1] * r[2][0]);
    double qfac = 1;
    if (det < 0) { for (int i = 0; i < 3; ++i) r[i][2] = -r[i][2]; qfac = -1; }
    double a = r[0][0] + r[1][1] + r[2][2] + 1.0, b, c, dd;
    if (a > 0.5) { a = 0.5 * std::sqrt(a); b = 0.25 * (r[2][1] - r[1][2]) / a; c = 0.25 * (r[0][2] - r[2][0]) / a; dd = 0.25 * (r[1][0] - r[0][1]) / a; }
    else {
        const double xd = 1.0 + r[0][0] - (r[1][1] + r[2][2]), yd = 1.0 + r[1][1] - (r[0][0] + r[2][2]), zd = 1.0 + r[2][2] - (r[0][0] + r[1][1]);
        if (xd > 1.0) { b = 0.5 * std::sqrt(xd); c = 0.25 * (r[0][1] + r[1][0]) / b; dd = 0.25 * (r[0][2] + r[2][0]) / b; a = 0.25 * (r[2][1] - r[1][2]) / b; }
        else if (yd > 1.0) { c = 0.5 * std::sqrt(yd); b = 0.25 * (r[0][1] + r[1][0]) / c; dd = 0.25 * (r[1][2] + r[2][1]) / c; a = 0.25 * (r[0][2] - r[2][0]) / c; }
        else { dd = 0.5 * std::sqrt(zd); b = 0.25 * (r[0][2] + r[2][0]) / dd; c = 0.25 * (r[1][2] + r[2][1]) / dd; a = 0.25 * (r[1][0] - r[0][1]) / dd; }
        if (a < 0.0) { b = -b; c = -c; dd = -dd; }
    }
    h.qform_code = 1; h.sform_code = 1; h.pixdim[0] = (float)qfac;
    h.quatern_b = (float)b; h.quatern_c = (float)c; h.quatern_d = (float)dd;
    h.qoffset_x = (float)m.m[0][3]; h.qoffset_y = (float)m.m[1][3]; h.qoffset_z = (float)m.m[2][3];
    for (int j = 0; j < 4; ++j) { h.srow_x[j] = (float)m.m[0][j]; h.srow_y[j] = (float)m.m[1][j]; h.srow_z[j] = (float)m.m[2][j]; }
    memcpy(h.magic, "n+1", 4);
    std::vector<unsigned char> buf(352 + img.n() * (as_float32 ? 4 : 8), 0);
    memcpy(buf.data(), &h, 348);
    if (as_float32) { float* o = (float*)(buf.data() + 352); for (size_t i = 0; i < img.n(); ++i) o[i] = (float)img.v[i]; }
    else memcpy(buf.data() + 352, img.v.data(), img.n() * 8);
    const bool gz = path.size() > 3 && path.substr(path.size() - 3) == ".gz";
    bool ok;
    if (gz) { gzFile g = gzopen(path.c_str(), "wb1"); ok = g && gzwrite(g, buf.data(), (unsigned)buf.size()) == (int)buf.size(); if (g) gzclose(g); }
    else { FILE* f = fopen(path.c_str(), "wb"); ok = f && fwrite(buf.data(), 1, buf.size(), f) == buf.size(); if (f) fclose(f); }
    if (!ok && err) *err = "cannot write " + path;
    return ok;
}

}  // namespace svr
