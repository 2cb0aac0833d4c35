// svr_rreg.cu -- IRTK-style rigid image registration on the device, batched (SURVEY.md section 8f n2 / n3 and the PVR patch
// registration): the engine behind the reference's DEFAULT registrations, all of which run irtkImageRigidRegistrationWithPadding
// on the CPU through TBB:
//   * StackRegistrations            irtkReconstructionGPU.cc:849-1001    (GuessParameterThickSlices, target padding 0)
//   * SliceToVolumeRegistration     irtkReconstructionGPU.cc:1992-2059   (GuessParameterSliceToVolume, target padding -1)
//   * PVR patch-to-volume (runHybrid) patchBased2D3DRegistration.cpp:88-168 (the same call with a 64x64 patch as target)
// What is restated (IRTKSimple2/...):
//   packages/registration/src/irtkImageRegistration.cc:414-528 (Run: 3 levels x 4 step sizes x <= 20 iterations),
//   :530-568 (EvaluateGradient: central differences, normalised), irtkGradientDescentOptimizer.cc:25-69 (line search + back-track),
//   irtkImageRegistrationWithPadding.cc:27-334 (per-level preparation: blur, resample, shift to >= 0, padding -> -1),
//   irtkImageRigidRegistrationWithPadding.cc:110-205,304-402 (parameter guesses), :534-610 (Evaluate),
//   include/irtkCrossCorrelationSimilarityMetric.h (CC on integer samples), image++/src/irtkGaussianBlurringWithPadding.cc,
//   irtkConvolutionWithPadding_1D.cc:38-88, irtkResamplingWithPadding.cc:36-183,203-262,
//   irtkLinearInterpolateImageFunction.cc:59-99, irtkBaseImage.cc:79-147, irtkHomogeneousTransformationIterator.h,
//   packages/transformation/src/irtkRigidTransformation.cc:26-53.
//
// Design.  The optimiser of every item (stack / slice / patch) is a small state machine on the host; all items advance in
// lockstep rounds, and ONE kernel launch per round evaluates the similarity of every item that asked for one.  A similarity
// evaluation is integer work: the target voxel (short) and the rounded trilinear sample of the source (short) feed six integer
// sums (n, x, y, xx, yy, xy), accumulated exactly in 64-bit integers -- so the result does not depend on the summation order,
// and equals the reference's double-precision sums bit for bit.  The sample positions reproduce the reference's
// irtkHomogeneousTransformationIterator (positions advanced by repeated addition per row / plane, run-length jumps over padded
// target voxels) with one thread per target row walking its voxels in order, and every floating-point operation is a separately
// rounded IEEE operation (__dmul_rn / __dadd_rn: no FMA contraction), so positions, interpolation weights and roundings are those
// of the CPU code.  The per-level image preparation (separable Gaussian with padding, trilinear resampling with padding, both
// truncating to short after every pass like the reference's in-place filters) runs on the device under the same rule.
// tests/test_rreg.py compares every stage with the reference's own IRTK (oracle/_ref/libref_irtk.so): prepared images and
// similarities are required to be IDENTICAL, and with them the optimiser's path and the final parameters.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>
#include "../../include/svr_abi.h"
#include "svr_context.h"

namespace {

struct Attr { int x, y, z; double dx, dy, dz, o[3], ax[3], ay[3], az[3]; };
struct M4 { double m[4][4]; };

Attr attr_from18(const double* a)
{
    Attr t;
    memset(&t, 0, sizeof t);                   // padding bytes too: Attr is compared bytewise when images are grouped
    t.x = (int)a[0]; t.y = (int)a[1]; t.z = (int)a[2]; t.dx = a[3]; t.dy = a[4]; t.dz = a[5];
    for (int i = 0; i < 3; ++i) { t.o[i] = a[6 + i]; t.ax[i] = a[9 + i]; t.ay[i] = a[12 + i]; t.az[i] = a[15 + i]; }
    return t;
}
void attr_to18(const Attr& t, double* a)
{
    a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.dx; a[4] = t.dy; a[5] = t.dz;
    for (int i = 0; i < 3; ++i) { a[6 + i] = t.o[i]; a[9 + i] = t.ax[i]; a[12 + i] = t.ay[i]; a[15 + i] = t.az[i]; }
}
M4 zero4() { M4 r; memset(&r, 0, sizeof r); return r; }
M4 ident4() { M4 r = zero4(); for (int i = 0; i < 4; ++i) r.m[i][i] = 1.0; return r; }
// irtkMatrix::operator* (geometry++/src/irtkMatrix.cc:222-242): tmp(i,j) = 0; for k: tmp(i,j) += a(i,k) * b(k,j)
M4 mul(const M4& a, const M4& b)
{
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            volatile double s = 0.0;                           // volatile: one rounding per operation, whatever the host compiler flags
            for (int k = 0; k < 4; ++k) { volatile double p = a.m[i][k] * b.m[k][j]; s = s + p; }
            r.m[i][j] = s;
        }
    return r;
}
// irtkBaseImage::GetImageToWorldMatrix / GetWorldToImageMatrix (image++/src/irtkBaseImage.cc:79-147)
M4 i2w_of(const Attr& a)
{
    M4 t1 = ident4(), sc = zero4(), rot = zero4(), t2 = ident4();
    t1.m[0][3] = -(a.x - 1) / 2.0; t1.m[1][3] = -(a.y - 1) / 2.0; t1.m[2][3] = -(a.z - 1) / 2.0;
    sc.m[0][0] = a.dx; sc.m[1][1] = a.dy; sc.m[2][2] = a.dz; sc.m[3][3] = 1.0;
    for (int i = 0; i < 3; ++i) { rot.m[i][0] = a.ax[i]; rot.m[i][1] = a.ay[i]; rot.m[i][2] = a.az[i]; }
    rot.m[3][3] = 1.0;
    t2.m[0][3] = a.o[0]; t2.m[1][3] = a.o[1]; t2.m[2][3] = a.o[2];
    return mul(t2, mul(rot, mul(sc, t1)));
}
M4 w2i_of(const Attr& a)
{
    M4 t1 = ident4(), rot = zero4(), sc = zero4(), t2 = ident4();
    t1.m[0][3] = -a.o[0]; t1.m[1][3] = -a.o[1]; t1.m[2][3] = -a.o[2];
    for (int i = 0; i < 3; ++i) { rot.m[0][i] = a.ax[i]; rot.m[1][i] = a.ay[i]; rot.m[2][i] = a.az[i]; }
    rot.m[3][3] = 1.0;
    sc.m[0][0] = 1.0 / a.dx; sc.m[1][1] = 1.0 / a.dy; sc.m[2][2] = 1.0 / a.dz; sc.m[3][3] = 1.0;
    t2.m[0][3] = (a.x - 1) / 2.0; t2.m[1][3] = (a.y - 1) / 2.0; t2.m[2][3] = (a.z - 1) / 2.0;
    return mul(t2, mul(sc, mul(rot, t1)));
}
// irtkRigidTransformation::UpdateMatrix (packages/transformation/src/irtkRigidTransformation.cc:26-53)
M4 rigid_of(const double* d)
{
    const double cosrx = cos(d[3] * (M_PI / 180.0)), cosry = cos(d[4] * (M_PI / 180.0)), cosrz = cos(d[5] * (M_PI / 180.0));
    const double sinrx = sin(d[3] * (M_PI / 180.0)), sinry = sin(d[4] * (M_PI / 180.0)), sinrz = sin(d[5] * (M_PI / 180.0));
    M4 r = ident4();
    volatile double a, b;
    r.m[0][0] = cosry * cosrz; r.m[0][1] = cosry * sinrz; r.m[0][2] = -sinry; r.m[0][3] = d[0];
    a = sinrx * sinry; a = a * cosrz; b = cosrx * sinrz; r.m[1][0] = a - b;
    a = sinrx * sinry; a = a * sinrz; b = cosrx * cosrz; r.m[1][1] = a + b;
    r.m[1][2] = sinrx * cosry; r.m[1][3] = d[1];
    a = cosrx * sinry; a = a * cosrz; b = sinrx * sinrz; r.m[2][0] = a + b;
    a = cosrx * sinry; a = a * sinrz; b = sinrx * cosrz; r.m[2][1] = a - b;
    r.m[2][2] = cosrx * cosry; r.m[2][3] = d[2];
    return r;
}
inline int iround(double x) { return x > 0 ? int(x + 0.5) : int(x - 0.5); }      // common++/include/irtkCommon.h:85-88

struct DevImg {                      // nb images of identical geometry, back to back
    short* d = nullptr;
    Attr a{};
    int nb = 1;
    size_t n() const { return (size_t)a.x * a.y * a.z; }
    size_t total() const { return n() * (size_t)nb; }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// Device kernels.  Every floating-point operation below is an explicitly rounded IEEE operation.
#define DMUL(a, b) __dmul_rn((a), (b))
#define DADD(a, b) __dadd_rn((a), (b))
#define DSUB(a, b) __dadd_rn((a), -(b))

__device__ __forceinline__ short put_as_double_short(double v)
{   // irtkGenericImage<short>::PutAsDouble (image++/include/irtkGenericImage.h:303-333): clamp, then static_cast (truncation)
    if (v > 32767.0) v = 32767.0;
    if (v < -32768.0) v = -32768.0;
    return (short)v;
}

// irtkConvolutionWithPadding_1D<short>::Run(x, y, z, t) with normalisation, along `axis` (image++/src/irtkConvolutionWithPadding_1D.cc:38-88)
__global__ void rreg_blur_kernel(const short* __restrict__ in, short* __restrict__ out, int X, int Y, int Z, int nb, int axis,
                                 const double* __restrict__ kern, int n, int padding)
{
    // nb images of X x Y x Z voxels back to back: voxel index within the batch = image * XYZ + (z * Y + y) * X + x
    const size_t N = (size_t)X * Y * Z * nb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < N; idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % X), y = (int)((idx / X) % Y), z = (int)((idx / ((size_t)X * Y)) % Z);
        if ((int)in[idx] <= padding) { out[idx] = (short)padding; continue; }
        const int c0 = axis == 0 ? x : (axis == 1 ? y : z), dim = axis == 0 ? X : (axis == 1 ? Y : Z);
        const size_t stride = axis == 0 ? 1 : (axis == 1 ? (size_t)X : (size_t)X * Y);
        const size_t base = idx - (size_t)c0 * stride;
        double val = 0.0, sum = 0.0;
        const int c1 = c0 - n / 2;
        for (int t = 0; t < n; ++t) {
            const int c = c1 + t;
            if (c >= 0 && c < dim) {
                const int s = in[base + (size_t)c * stride];
                if (s > padding) { val = DADD(val, DMUL(kern[t], (double)s)); sum = DADD(sum, kern[t]); }
            }
        }
        out[idx] = put_as_double_short(sum > 0 ? __ddiv_rn(val, sum) : 0.0);
    }
}

struct Mat34 { double m[12]; };
__device__ __forceinline__ void apply34(const Mat34& M, double& x, double& y, double& z)
{   // irtkBaseImage::ImageToWorld / WorldToImage (image++/include/irtkBaseImage.h:425-468): ((m0 x + m1 y) + m2 z) + m3
    const double a = DADD(DADD(DADD(DMUL(M.m[0], x), DMUL(M.m[1], y)), DMUL(M.m[2], z)), M.m[3]);
    const double b = DADD(DADD(DADD(DMUL(M.m[4], x), DMUL(M.m[5], y)), DMUL(M.m[6], z)), M.m[7]);
    const double c = DADD(DADD(DADD(DMUL(M.m[8], x), DMUL(M.m[9], y)), DMUL(M.m[10], z)), M.m[11]);
    x = a; y = b; z = c;
}

// irtkMultiThreadedResamplingWithPadding (image++/src/irtkResamplingWithPadding.cc:36-183)
__global__ void rreg_resample_kernel(const short* __restrict__ in_all, int X, int Y, int Z, short* __restrict__ out, int OX, int OY, int OZ,
                                     int nb, Mat34 out_i2w, Mat34 in_w2i, int padding)
{
    const size_t ON = (size_t)OX * OY * OZ, N = ON * nb;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < N; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t loc = idx % ON;
        const short* in = in_all + (idx / ON) * ((size_t)X * Y * Z);
        const int i = (int)(loc % OX), j = (int)((loc / OX) % OY), k = (int)(loc / ((size_t)OX * OY));
        double x = i, y = j, z = k;
        apply34(out_i2w, x, y, z);
        apply34(in_w2i, x, y, z);
        const int u = (int)floor(x), v = (int)floor(y), w = (int)floor(z);
        const double dx = DSUB(x, (double)u), dy = DSUB(y, (double)v), dz = DSUB(z, (double)w);
        const double ax = DSUB(1.0, dx), ay = DSUB(1.0, dy), az = DSUB(1.0, dz);
        const double wt[8] = { DMUL(DMUL(ax, ay), az), DMUL(DMUL(ax, ay), dz), DMUL(DMUL(ax, dy), az), DMUL(DMUL(ax, dy), dz),
                               DMUL(DMUL(dx, ay), az), DMUL(DMUL(dx, ay), dz), DMUL(DMUL(dx, dy), az), DMUL(DMUL(dx, dy), dz) };
        double val = 0.0, sum = 0.0;
        int pad = 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int uu = u + (q >> 2), vv = v + ((q >> 1) & 1), ww = w + (q & 1);
            if (uu >= 0 && uu < X && vv >= 0 && vv < Y && ww >= 0 && ww < Z) {
                const int g = in[((size_t)ww * Y + vv) * X + uu];
                if (g != padding) { pad--; val = DADD(val, DMUL((double)g, wt[q])); sum = DADD(sum, wt[q]); }
            } else {
                pad--;
            }
        }
        short r = (short)padding;
        if (pad < 4 && sum > 0) r = put_as_double_short(__ddiv_rn(val, sum));
        out[idx] = r;
    }
}

// per-image range over voxels > padding: grid = (chunks, images)
__global__ void rreg_minmax_kernel(const short* __restrict__ in, size_t N, int padding, int* __restrict__ mm)
{
    const short* img = in + (size_t)blockIdx.y * N;
    int mn = INT_MAX, mx = INT_MIN;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < N; idx += (size_t)gridDim.x * blockDim.x) {
        const int v = img[idx];
        if (v > padding) { mn = min(mn, v); mx = max(mx, v); }
    }
    for (int o = 16; o; o >>= 1) { mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(&mm[2 * blockIdx.y], mn); atomicMax(&mm[2 * blockIdx.y + 1], mx); }
}
__global__ void rreg_fill_minmax_kernel(int* __restrict__ mm, int nb)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nb) { mm[2 * i] = INT_MAX; mm[2 * i + 1] = INT_MIN; }
}
// irtkImageRegistrationWithPadding::Initialize(level): voxels > padding -> value - min, others -> -1
__global__ void rreg_shift_kernel(short* __restrict__ in, size_t N, int padding, const int* __restrict__ mm)
{
    short* img = in + (size_t)blockIdx.y * N;
    const int mn = mm[2 * blockIdx.y];
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < N; idx += (size_t)gridDim.x * blockDim.x) {
        const int v = img[idx];
        img[idx] = v > padding ? (short)(v - mn) : (short)-1;
    }
}

// One similarity evaluation of many items (irtkImageRigidRegistrationWithPadding::Evaluate + CC metric Add), one thread per target row.
struct EvalItem {
    const short* tgt; const short* src;
    int tx, ty, tz, sx, sy, sz;
    int row0;                 // first row of this item in the launch
    double m[12];             // source W2I * T * target I2W, rows 0..2
};
__global__ void rreg_eval_kernel(const EvalItem* __restrict__ items, int n_items, int total_rows, unsigned long long* __restrict__ sums)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    long long n = 0, sx_ = 0, sy_ = 0, sxx = 0, syy = 0, sxy = 0;
    int it = -1;
    if (row < total_rows) {
        int lo = 0, hi = n_items - 1;                            // the item whose rows contain `row`
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (items[mid].row0 <= row) lo = mid; else hi = mid - 1; }
        it = lo;
        const EvalItem& e = items[it];
        const int r = row - e.row0, k = r / e.ty, j = r - k * e.ty;
        // irtkHomogeneousTransformationIterator: start = (m03, m13, m23); NextZ k times, then NextY j times (repeated additions)
        double zx = e.m[3], zy = e.m[7], zz = e.m[11];
        for (int q = 0; q < k; ++q) { zx = DADD(zx, e.m[2]); zy = DADD(zy, e.m[6]); zz = DADD(zz, e.m[10]); }
        double x = zx, y = zy, z = zz;
        for (int q = 0; q < j; ++q) { x = DADD(x, e.m[1]); y = DADD(y, e.m[5]); z = DADD(z, e.m[9]); }
        const short* trow = e.tgt + ((size_t)k * e.ty + j) * e.tx;
        const double x2 = e.sx - 1, y2 = e.sy - 1, z2 = e.sz - 1;
        const size_t o3 = e.sx, o5 = (size_t)e.sx * e.sy;
        int i = 0;
        while (i < e.tx) {
            const int tv = trow[i];
            if (tv >= 0) {
                if (x > 0 && x < x2 && y > 0 && y < y2 && z > 0 && z < z2) {
                    // irtkLinearInterpolateImageFunction::EvaluateInside (short voxels)
                    const int ii = (int)x, jj = (int)y, kk = (int)z;
                    const double t1 = DSUB(x, (double)ii), u1 = DSUB(y, (double)jj), v1 = DSUB(z, (double)kk);
                    const double t2 = DSUB(1.0, t1), u2 = DSUB(1.0, u1), v2 = DSUB(1.0, v1);
                    const short* p = e.src + ((size_t)kk * e.sy + jj) * e.sx + ii;
                    const double p1 = p[0], p2 = p[1], p3 = p[o3], p4 = p[o3 + 1], p5 = p[o5], p6 = p[o5 + 1], p7 = p[o5 + o3], p8 = p[o5 + o3 + 1];
                    const double a = DMUL(t1, DADD(DMUL(u2, DADD(DMUL(v2, p2), DMUL(v1, p6))), DMUL(u1, DADD(DMUL(v2, p4), DMUL(v1, p8)))));
                    const double b = DMUL(t2, DADD(DMUL(u2, DADD(DMUL(v2, p1), DMUL(v1, p5))), DMUL(u1, DADD(DMUL(v2, p3), DMUL(v1, p7)))));
                    const double value = DADD(a, b);
                    if (value >= 0) {
                        const long long sv = value > 0 ? (long long)(int)DADD(value, 0.5) : (long long)(int)DSUB(value, 0.5);
                        n += 1; sx_ += tv; sy_ += sv; sxx += (long long)tv * tv; syy += sv * sv; sxy += (long long)tv * sv;
                    }
                }
                x = DADD(x, e.m[0]); y = DADD(y, e.m[4]); z = DADD(z, e.m[8]);
                ++i;
            } else {
                int l = i + 1;                                   // irtkPadding run: jump to its end with ONE multiply-add
                while (l < e.tx && trow[l] < 0) ++l;
                const double off = (double)(l - i);
                x = DADD(x, DMUL(e.m[0], off)); y = DADD(y, DMUL(e.m[4], off)); z = DADD(z, DMUL(e.m[8], off));
                i = l;
            }
        }
    }
    // per-warp segmented fold: lanes of one item add up, one atomic per (warp, item, sum)
    const unsigned full = 0xffffffffu;
    const unsigned peers = __match_any_sync(full, it);
    const int leader = __ffs(peers) - 1;
    long long v[6] = { n, sx_, sy_, sxx, syy, sxy };
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        long long acc = 0;
        for (unsigned mset = peers; mset; mset &= mset - 1) acc += __shfl_sync(peers, v[q], __ffs(mset) - 1);
        if ((int)(threadIdx.x & 31) == leader && it >= 0 && acc != 0) atomicAdd(&sums[(size_t)it * 6 + q], (unsigned long long)acc);
    }
}

// ---------------------------------------------------------------------------------------------
namespace {

struct RegParams {
    int levels = 3;
    double tblur[3], sblur[3], tres[3][3], sres[3][3];
    int iterations[3], steps[3];
    double step_len[3];
    double epsilon = 0.0001;
    int tpad = 0, spad = -32768;
};

int corner_padding(const short* d, const Attr& a)
{   // the "guess padding from the eight corners" block of the GuessParameter* functions
    auto g = [&](int x, int y, int z) { return (int)d[((size_t)z * a.y + y) * a.x + x]; };
    const int X = a.x - 1, Y = a.y - 1, Z = a.z - 1, c = g(0, 0, 0);
    if (g(X, 0, 0) == c && g(0, Y, 0) == c && g(0, 0, Z) == c && g(X, Y, 0) == c && g(0, Y, Z) == c && g(X, 0, Z) == c && g(X, Y, Z) == c) return c;
    return -32768;
}

// GuessParameterThickSlices (kind 0) / GuessParameterSliceToVolume(false) (kind 1) + the SetTargetPadding of the call sites
RegParams guess(int kind, const Attr& t, const Attr& s, const short* sdata)
{
    RegParams p;
    double size = t.dy < t.dx ? t.dy : t.dx;
    p.tblur[0] = size / 2.0; p.tres[0][0] = size; p.tres[0][1] = size; p.tres[0][2] = t.dz;
    for (int i = 1; i < 3; ++i) {
        p.tblur[i] = p.tblur[i - 1] * 2; p.tres[i][0] = p.tres[i - 1][0] * 2; p.tres[i][1] = p.tres[i - 1][1] * 2; p.tres[i][2] = p.tres[i - 1][2];
    }
    size = s.dy < s.dx ? s.dy : s.dx;
    if (kind == 1 && s.dz < size) size = s.dz;
    p.sblur[0] = size / 2.0; p.sres[0][0] = size; p.sres[0][1] = size; p.sres[0][2] = kind == 1 ? size : s.dz;
    for (int i = 1; i < 3; ++i) {
        p.sblur[i] = p.sblur[i - 1] * 2; p.sres[i][0] = p.sres[i - 1][0] * 2; p.sres[i][1] = p.sres[i - 1][1] * 2;
        p.sres[i][2] = kind == 1 ? p.sres[i - 1][2] * 2 : p.sres[i - 1][2];
    }
    for (int i = 0; i < 3; ++i) { p.iterations[i] = 20; p.steps[i] = 4; p.step_len[i] = 2 * pow(2.0, i); }
    p.tpad = kind == 0 ? 0 : -1;
    p.spad = corner_padding(sdata, s);
    return p;
}

Mat34 to34(const M4& m) { Mat34 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) r.m[4 * i + j] = m.m[i][j]; return r; }

struct Engine {
    svr_context* c;
    std::vector<void*> owned;
    ~Engine() { for (void* p : owned) cudaFree(p); }
    template <class T> int alloc(T** p, size_t n)
    {
        *p = nullptr;
        if (cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) return svr_fail(c, "cudaMalloc (svr_rreg)", cudaGetLastError(), __FILE__, __LINE__);
        owned.push_back(*p);
        return 0;
    }
    void release(void* p) { auto it = std::find(owned.begin(), owned.end(), p); if (it != owned.end()) { owned.erase(it); cudaFree(p); } }
    int grid(size_t n) const { return (int)std::min<size_t>((n + 255) / 256, (size_t)c->sm_count * 16); }

    // irtkGaussianBlurringWithPadding<short>::Run (image++/src/irtkGaussianBlurringWithPadding.cc:36-117): X, Y, Z passes, short after each
    int blur(DevImg& img, double sigma, int padding)
    {
        short* tmp = nullptr;
        if (alloc(&tmp, img.total())) return 1;
        const double vox[3] = { img.a.dx, img.a.dy, img.a.dz };
        const int dims[3] = { img.a.x, img.a.y, img.a.z };
        for (int axis = 0; axis < 3; ++axis) {
            if (axis == 2 && dims[2] == 1) break;              // "if (this->_output->GetX() != 1)" after the flips
            const double s = sigma / vox[axis];
            const int n = 2 * iround(4 * sigma / vox[axis]) + 1;
            std::vector<double> k(n);
            // irtkScalarGaussian(s, 1, 1, 0, 0, 0) sampled at the world coordinates of an n x 1 x 1 unit-voxel image centred on 0
            const double norm = 1.0 / (sqrt(2.0 * M_PI) * s * sqrt(2.0 * M_PI) * 1 * sqrt(2.0 * M_PI) * 1);
            for (int i = 0; i < n; ++i) {
                const double x = 1.0 * i + 0.0 * 0 + 0.0 * 0 + (-(n - 1) / 2.0);
                double v = norm * exp(-((x - 0) * (x - 0)) / (2.0 * s * s) - ((0.0 - 0) * (0.0 - 0)) / (2.0 * 1 * 1) - ((0.0 - 0) * (0.0 - 0)) / (2.0 * 1 * 1));
                if (fabs(v) < FLT_MIN) v = 0;
                k[i] = v;
            }
            double* dk = nullptr;
            if (alloc(&dk, n)) return 1;
            SVR_CUDA(c, cudaMemcpyAsync(dk, k.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            SVR_CUDA(c, cudaStreamSynchronize(c->stream));
            rreg_blur_kernel<<<grid(img.total()), 256, 0, c->stream>>>(img.d, tmp, img.a.x, img.a.y, img.a.z, img.nb, axis, dk, n, padding);
            SVR_KERNEL_CHECK(c);
            std::swap(img.d, tmp);
            SVR_CUDA(c, cudaStreamSynchronize(c->stream));
            release(dk);
        }
        release(tmp);
        return 0;
    }
    // irtkResamplingWithPadding<short>::Initialize + Run (image++/src/irtkResamplingWithPadding.cc:203-262): the new grid size is
    // ROUNDED here (irtkResampling::Initialize truncates)
    int resample(DevImg& img, double rx, double ry, double rz, int padding)
    {
        Attr o = img.a;
        int nx = iround(img.a.x * img.a.dx / rx), ny = iround(img.a.y * img.a.dy / ry), nz = iround(img.a.z * img.a.dz / rz);
        if (nx < 1) { nx = 1; o.dx = img.a.dx; } else o.dx = rx;
        if (ny < 1) { ny = 1; o.dy = img.a.dy; } else o.dy = ry;
        if (nz < 1) { nz = 1; o.dz = img.a.dz; } else o.dz = rz;
        o.x = nx; o.y = ny; o.z = nz;
        short* out = nullptr;
        if (alloc(&out, (size_t)nx * ny * nz * img.nb)) return 1;
        rreg_resample_kernel<<<grid((size_t)nx * ny * nz * img.nb), 256, 0, c->stream>>>(img.d, img.a.x, img.a.y, img.a.z, out, nx, ny, nz, img.nb,
                                                                                          to34(i2w_of(o)), to34(w2i_of(img.a)), padding);
        SVR_KERNEL_CHECK(c);
        SVR_CUDA(c, cudaStreamSynchronize(c->stream));
        release(img.d);
        img.d = out; img.a = o;
        return 0;
    }
    // the rest of irtkImageRegistrationWithPadding::Initialize(level): range over voxels > padding, shift to >= 0, padding -> -1
    int shift(DevImg& img, int padding)
    {
        int* mm = nullptr;
        if (alloc(&mm, 2 * (size_t)img.nb)) return 1;
        rreg_fill_minmax_kernel<<<(img.nb + 255) / 256, 256, 0, c->stream>>>(mm, img.nb);
        SVR_KERNEL_CHECK(c);
        const int chunks = (int)std::min<size_t>((img.n() + 1023) / 1024, 64);
        for (int b0 = 0; b0 < img.nb; b0 += 65535) {            // gridDim.y limit
            const dim3 g(std::max(chunks, 1), std::min(img.nb - b0, 65535));
            rreg_minmax_kernel<<<g, 256, 0, c->stream>>>(img.d + (size_t)b0 * img.n(), img.n(), padding, mm + 2 * b0);
            SVR_KERNEL_CHECK(c);
            rreg_shift_kernel<<<g, 256, 0, c->stream>>>(img.d + (size_t)b0 * img.n(), img.n(), padding, mm + 2 * b0);
            SVR_KERNEL_CHECK(c);
        }
        SVR_KERNEL_CHECK(c);
        SVR_CUDA(c, cudaStreamSynchronize(c->stream));
        release(mm);
        return 0;
    }
    // the images of one group (identical geometry and parameters) at one level: copies of the originals, blurred, resampled when
    // the level asks for it, shifted -- one set of launches for the whole group
    int prepare(const std::vector<const short*>& host, const Attr& a, double blur_sigma, const double res0[3], const double res[3], int level, int padding,
                DevImg& out)
    {
        out.a = a; out.nb = (int)host.size();
        if (alloc(&out.d, out.total())) return 1;
        for (size_t q = 0; q < host.size(); ++q)
            SVR_CUDA(c, cudaMemcpyAsync(out.d + q * out.n(), host[q], out.n() * sizeof(short), cudaMemcpyHostToDevice, c->stream));
        SVR_CUDA(c, cudaStreamSynchronize(c->stream));
        if (blur_sigma > 0 && blur(out, blur_sigma, padding)) return 1;
        const double temp = fabs(res0[0] - a.dx) + fabs(res0[1] - a.dy) + fabs(res0[2] - a.dz);
        if ((level > 0 || temp > 0.000001) && resample(out, res[0], res[1], res[2], padding)) return 1;
        return shift(out, padding);
    }
};

// Per-item optimiser: irtkImageRegistration::Run's step / iteration loops around irtkGradientDescentOptimizer::Run, as a state
// machine that asks for one similarity evaluation at a time.
struct ItemOpt {
    double dof[6];
    double step = 0;
    int istep = 0, iter = 0;
    enum Phase { BASE, GRAD, LINE, LEVEL_DONE } phase = LEVEL_DONE;
    double old_sim = 0, new_sim = 0, sim = 0;
    int g = 0;                       // gradient probe index 0..11: dof g/2, + step for even, - step for odd
    double saved = 0, s1 = 0;
    float dx[6];
    double before[6];                // parameters at the start of the optimiser iteration (maxChange)
    long long evaluations = 0;

    void start_level(double step0) { step = step0; istep = 0; iter = 0; begin_iteration(); }
    void begin_iteration() { for (int i = 0; i < 6; ++i) before[i] = dof[i]; phase = BASE; }
    // feed the similarity of the current parameters; afterwards `dof` holds the next parameters to evaluate (unless LEVEL_DONE)
    void feed(double s, int n_steps, int n_iterations, double epsilon)
    {
        ++evaluations;
        if (phase == BASE) {
            old_sim = new_sim = sim = s;
            g = 0; saved = dof[0];
            dof[0] = saved + (double)(float)step;                  // EvaluateGradient(float step, ...): Put(i, value + step)
            phase = GRAD;
        } else if (phase == GRAD) {
            const int i = g >> 1;
            if ((g & 1) == 0) { s1 = s; dof[i] = saved - (double)(float)step; ++g; }
            else {
                dx[i] = (float)(s1 - s);
                dof[i] = saved;
                ++g;
                if (g < 12) { saved = dof[g >> 1]; dof[g >> 1] = saved + (double)(float)step; }
                else {
                    double norm = 0;
                    for (int q = 0; q < 6; ++q) norm += dx[q] * dx[q];
                    norm = sqrt(norm);
                    for (int q = 0; q < 6; ++q) dx[q] = norm > 0 ? (float)(dx[q] / norm) : 0.f;
                    new_sim = sim;                             // first pass of the do { } while
                    for (int q = 0; q < 6; ++q) dof[q] = dof[q] + step * dx[q];
                    phase = LINE;
                }
            }
        } else if (phase == LINE) {
            sim = s;
            if (sim > new_sim + epsilon) {
                new_sim = sim;
                for (int q = 0; q < 6; ++q) dof[q] = dof[q] + step * dx[q];
            } else {
                for (int q = 0; q < 6; ++q) dof[q] = dof[q] - step * dx[q];      // last step was no improvement: back-track
                const double eps = new_sim > old_sim ? new_sim - old_sim : 0;
                double max_change = 0;
                for (int q = 0; q < 6; ++q) max_change = std::max(max_change, fabs(dof[q] - before[q]));
                const bool improved = eps > epsilon && max_change > 0;          // _Delta[level] is 0
                ++iter;
                if (!improved || iter >= n_iterations) {
                    step = step / 2; ++istep; iter = 0;
                    if (istep >= n_steps) { phase = LEVEL_DONE; return; }
                }
                begin_iteration();
            }
        }
    }
};

}  // namespace

extern "C" {

// Blur / resample building blocks on host images, for the tests (and for callers that need IRTK's padded filters on the device).
int svr_rreg_blur_with_padding(svr_context* c, const short* voxels, const double attr18[18], double sigma, int padding, short* out)
{
    SVR_ENTRY(c);
    if (!c || !voxels || !attr18 || !out) return 2;
    Engine e{ c };
    DevImg img; img.a = attr_from18(attr18);
    if (e.alloc(&img.d, img.n())) return 1;
    SVR_CUDA(c, cudaMemcpyAsync(img.d, voxels, img.n() * sizeof(short), cudaMemcpyHostToDevice, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    if (e.blur(img, sigma, padding)) return 1;
    SVR_CUDA(c, cudaMemcpyAsync(out, img.d, img.n() * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int svr_rreg_resample_with_padding(svr_context* c, const short* voxels, const double attr18[18], double dx, double dy, double dz, int padding,
                                   short* out, size_t out_capacity, double out_attr18[18])
{
    SVR_ENTRY(c);
    if (!c || !voxels || !attr18 || !out || !out_attr18) return 2;
    Engine e{ c };
    DevImg img; img.a = attr_from18(attr18);
    if (e.alloc(&img.d, img.n())) return 1;
    SVR_CUDA(c, cudaMemcpyAsync(img.d, voxels, img.n() * sizeof(short), cudaMemcpyHostToDevice, c->stream));
    if (e.resample(img, dx, dy, dz, padding)) return 1;
    if (img.n() > out_capacity) { c->err = "svr_rreg_resample_with_padding: output buffer too small"; return 2; }
    SVR_CUDA(c, cudaMemcpyAsync(out, img.d, img.n() * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
    SVR_CUDA(c, cudaStreamSynchronize(c->stream));
    attr_to18(img.a, out_attr18);
    return 0;
}

// The registration.  images[n_images]: host images (short voxels + 18 attribute doubles each, given as two parallel arrays);
// item i registers images[target_of_item[i]] (target) to images[source_of_item[i]] (source) starting from dofs[6 i .. 6 i + 5]
// (tx, ty, tz in mm, rx, ry, rz in degrees), which receive the result.  kind 0 / 1: see the header.  level_only >= 0: prepare only
// that level and return the similarity of the given parameters in similarity[i] without optimising (test tap);
// prepared_target / prepared_source (may be NULL): receive item 0's prepared images of that level (prepared_capacity voxels each).
int svr_rreg_register(svr_context* c, int n_items, int n_images, const short* const* voxels, const double* attrs18, const int* target_of_item,
                      const int* source_of_item, int kind, double* dofs, double* similarity, int64_t* evaluations, int level_only, size_t prepared_capacity,
                      short* prepared_target, double* prepared_target_attr18, short* prepared_source, double* prepared_source_attr18)
{
    SVR_ENTRY(c);
    if (!c) return 2;
    if (n_items < 0 || n_images <= 0 || !voxels || !attrs18 || !target_of_item || !source_of_item || !dofs || (kind != 0 && kind != 1)) {
        c->err = "svr_rreg_register: bad argument";
        return 2;
    }
    if (n_items == 0) return 0;
    ProfScope prof(c, 5);
    std::vector<Attr> attr(n_images);
    for (int i = 0; i < n_images; ++i) attr[i] = attr_from18(attrs18 + 18 * i);
    for (int i = 0; i < n_items; ++i)
        if (target_of_item[i] < 0 || target_of_item[i] >= n_images || source_of_item[i] < 0 || source_of_item[i] >= n_images) {
            c->err = "svr_rreg_register: image index out of range";
            return 2;
        }
    // the reference guesses the parameters per registration object; a batch shares them per (target, source) geometry class, so
    // every item carries its own
    std::vector<RegParams> params(n_items);
    for (int i = 0; i < n_items; ++i) params[i] = guess(kind, attr[target_of_item[i]], attr[source_of_item[i]], voxels[source_of_item[i]]);
    std::vector<ItemOpt> opt(n_items);
    for (int i = 0; i < n_items; ++i) for (int q = 0; q < 6; ++q) opt[i].dof[q] = dofs[6 * i + q];
    long long total_evals = 0;

    for (int level = 2; level >= 0; --level) {
        if (level_only >= 0 && level != level_only) continue;
        Engine e{ c };
        // Prepare every distinct (image, role, parameters) once; images that share geometry and parameters (the patches of a stack,
        // the slices of a stack) form a GROUP that is filtered by one set of launches.
        struct Key { int img, role, pad; double blur, r0, r1, r2, q0, q1, q2; bool operator<(const Key& o) const { return memcmp(this, &o, sizeof(Key)) < 0; } };
        struct GKey { int role, pad; double blur, r0, r1, r2, q0, q1, q2; Attr a; bool operator<(const GKey& o) const { return memcmp(this, &o, sizeof(GKey)) < 0; } };
        std::map<Key, std::pair<int, int>> index;            // -> (group, position in the group)
        std::map<GKey, int> gindex;
        std::vector<GKey> gkeys;
        std::vector<std::vector<const short*>> ghost;
        std::vector<std::pair<int, int>> tprep(n_items), sprep(n_items);
        for (int i = 0; i < n_items; ++i) {
            const RegParams& p = params[i];
            for (int role = 0; role < 2; ++role) {
                const int img = role == 0 ? target_of_item[i] : source_of_item[i];
                Key k; memset(&k, 0, sizeof k);
                k.img = img; k.role = role; k.pad = role == 0 ? p.tpad : p.spad; k.blur = role == 0 ? p.tblur[level] : p.sblur[level];
                const double* r = role == 0 ? p.tres[level] : p.sres[level];
                const double* r0 = role == 0 ? p.tres[0] : p.sres[0];
                k.r0 = r[0]; k.r1 = r[1]; k.r2 = r[2]; k.q0 = r0[0]; k.q1 = r0[1]; k.q2 = r0[2];
                auto f = index.find(k);
                if (f == index.end()) {
                    GKey g; memset(&g, 0, sizeof g);
                    g.role = role; g.pad = k.pad; g.blur = k.blur; g.r0 = k.r0; g.r1 = k.r1; g.r2 = k.r2; g.q0 = k.q0; g.q1 = k.q1; g.q2 = k.q2;
                    memcpy(&g.a, &attr[img], sizeof(Attr));
                    auto gf = gindex.find(g);
                    int gid;
                    if (gf == gindex.end()) { gid = (int)gkeys.size(); gindex[g] = gid; gkeys.push_back(g); ghost.emplace_back(); }
                    else gid = gf->second;
                    ghost[gid].push_back(voxels[img]);
                    f = index.emplace(k, std::make_pair(gid, (int)ghost[gid].size() - 1)).first;
                }
                (role == 0 ? tprep : sprep)[i] = f->second;
            }
        }
        std::vector<DevImg> prepared(gkeys.size());
        for (size_t g = 0; g < gkeys.size(); ++g) {
            const GKey& k = gkeys[g];
            const double res0[3] = { k.q0, k.q1, k.q2 }, res[3] = { k.r0, k.r1, k.r2 };
            if (e.prepare(ghost[g], k.a, k.blur, res0, res, level, k.pad, prepared[g])) return 1;
        }
        if (level_only >= 0) {
            auto dump = [&](std::pair<int, int> id, short* out, double* a18) -> int {
                if (!out) return 0;
                const DevImg& d = prepared[id.first];
                if (d.n() > prepared_capacity) { c->err = "svr_rreg_register: prepared-image buffer too small"; return 2; }
                SVR_CUDA(c, cudaMemcpyAsync(out, d.d + (size_t)id.second * d.n(), d.n() * sizeof(short), cudaMemcpyDeviceToHost, c->stream));
                SVR_CUDA(c, cudaStreamSynchronize(c->stream));
                if (a18) attr_to18(d.a, a18);
                return 0;
            };
            if (dump(tprep[0], prepared_target, prepared_target_attr18) || dump(sprep[0], prepared_source, prepared_source_attr18)) return 1;
        }
        std::vector<M4> ti2w(prepared.size()), sw2i(prepared.size());
        for (size_t q = 0; q < prepared.size(); ++q) { ti2w[q] = i2w_of(prepared[q].a); sw2i[q] = w2i_of(prepared[q].a); }

        EvalItem* d_items = nullptr; unsigned long long* d_sums = nullptr;
        if (e.alloc(&d_items, n_items) || e.alloc(&d_sums, (size_t)n_items * 6)) return 1;
        std::vector<EvalItem> h_items(n_items);
        std::vector<unsigned long long> h_sums((size_t)n_items * 6);
        std::vector<int> active;
        for (int i = 0; i < n_items; ++i) {
            if (level_only >= 0) opt[i].phase = ItemOpt::BASE;
            else opt[i].start_level(params[i].step_len[level]);
            active.push_back(i);
        }
        while (!active.empty()) {
            int rows = 0;
            for (size_t a = 0; a < active.size(); ++a) {
                const int i = active[a];
                const DevImg& t = prepared[tprep[i].first]; const DevImg& s = prepared[sprep[i].first];
                EvalItem& it = h_items[a];
                it.tgt = t.d + (size_t)tprep[i].second * t.n(); it.src = s.d + (size_t)sprep[i].second * s.n(); it.tx = t.a.x; it.ty = t.a.y; it.tz = t.a.z; it.sx = s.a.x; it.sy = s.a.y; it.sz = s.a.z;
                it.row0 = rows; rows += t.a.y * t.a.z;
                const M4 m = mul(mul(sw2i[sprep[i].first], rigid_of(opt[i].dof)), ti2w[tprep[i].first]);      // (W2I * T) * I2W
                for (int r = 0; r < 3; ++r) for (int q = 0; q < 4; ++q) it.m[4 * r + q] = m.m[r][q];
            }
            const int na = (int)active.size();
            SVR_CUDA(c, cudaMemcpyAsync(d_items, h_items.data(), na * sizeof(EvalItem), cudaMemcpyHostToDevice, c->stream));
            SVR_CUDA(c, cudaMemsetAsync(d_sums, 0, (size_t)na * 6 * sizeof(unsigned long long), c->stream));
            rreg_eval_kernel<<<(rows + 127) / 128, 128, 0, c->stream>>>(d_items, na, rows, d_sums);
            SVR_KERNEL_CHECK(c);
            SVR_CUDA(c, cudaMemcpyAsync(h_sums.data(), d_sums, (size_t)na * 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
            SVR_CUDA(c, cudaStreamSynchronize(c->stream));
            total_evals += na;
            std::vector<int> next;
            for (int a = 0; a < na; ++a) {
                const int i = active[a];
                const long long* s = (const long long*)&h_sums[(size_t)a * 6];
                const double n = (double)s[0], x = (double)s[1], y = (double)s[2], x2 = (double)s[3], y2 = (double)s[4], xy = (double)s[5];
                // irtkCrossCorrelationSimilarityMetric::Evaluate
                const double sim = n > 0 ? (xy - (x * y) / n) / (sqrt(x2 - x * x / n) * sqrt(y2 - y * y / n)) : 0.0;
                if (level_only >= 0) { if (similarity) similarity[i] = sim; continue; }
                opt[i].feed(sim, params[i].steps[level], params[i].iterations[level], params[i].epsilon);
                if (opt[i].phase != ItemOpt::LEVEL_DONE) next.push_back(i);
                else if (similarity) similarity[i] = opt[i].new_sim;
            }
            active.swap(next);
        }
    }
    if (level_only < 0) for (int i = 0; i < n_items; ++i) for (int q = 0; q < 6; ++q) dofs[6 * i + q] = opt[i].dof[q];
    if (evaluations) *evaluations = total_evals;
    return 0;
}

}  // extern "C"
