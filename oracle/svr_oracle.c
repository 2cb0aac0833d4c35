/*
 * svr_oracle.c -- CPU restatement of the SVR hot path of bkainz/fetalReconstruction
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (fetalreconstruction_b200/,
 * include/) may include, link or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
 * or as the timed CPU arm.
 *
 * PARITY STATUS: pinned against the reference's own CUDA path.  The reference ships no golden
 * vectors, tests or expected outputs for this path; its CUDA sources do
 * compile for sm_100a behind the shim headers of oracle/ref_shim/ (oracle/Makefile, `make ref`),
 * and their outputs on a B200 (tests/golden/ref_{steps,svr}_small.npz, oracle/ref_runner.py) are what
 * tests/test_ref_golden.py holds this line-by-line restatement to, stage by stage; the analytic
 * known-answer tests of tests/test_oracle_known_answers.py pin the mathematics.
 *
 * Every function cites the reference file:line it follows; paths are relative to
 * /root/reference/source/reconstructionGPU2/ ("cuda2.cu" = reconstruction_cuda2.cu,
 * "GPU.cc" = irtkReconstructionGPU.cc).
 *
 * Arithmetic: float, in the operation order of the reference kernels (build with
 * -ffp-contract=off).  Where the reference accumulates with float atomicAdd in a
 * non-deterministic order (the scatter into the volume), the oracle accumulates in
 * double and rounds once: that is the value the float atomics approximate.
 *
 * Documented deviations from the reference (same list as DESIGN.md):
 *   D1 (Q5)  x/y bounds are checked; the reference lets grid-overhang threads alias.
 *   D2 (Q6)  all slices are processed; the reference drops the last slice(s).
 *   D3 (Q9)  the regulariser reads a frozen copy of the post-gradient-step volume (the
 *            reference updates in place while neighbours still read; its CPU twin uses a copy).
 *   D4 (Q10) per-slice "voxel_num" counts (the reference returns one count per device).
 *   D5       sinc(0) = 1; the reference evaluates sin(0)/0 = NaN and silently drops the pixel.
 * Reproduced quirks: Q1-Q4 of the tap loop (epsilon skip against the last accepted tap,
 * first tap of a row never skipped, sume ignores the mask, negative coordinates saturate
 * to index 0), stale v_PSF_sums (only written when sume > 0.5), simulated slices only
 * written when weight > 0, min/max of the M-step seeded with 0.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_STEP 0.0001f          /* __step, include/reconstruction_cuda2.cuh:54 */
#define ORC_PSF_EPSILON 0.00001   /* PSF_EPSILON (double literal), reconstruction_cuda2.cuh:72 */
#define ORC_PSF_SUPPORT 16        /* MAX_PSF_SUPPORT, reconstruction_cuda2.cuh:74 */

typedef struct { float m[16]; } orc_mat4;   /* row-major; Matrix4 = float4 data[4], recon_volumeHelper.cuh:41 */
typedef struct { float x, y, z; } orc_f3;

/* ---- Matrix4 algebra: recon_volumeHelper.cuh:106-145 ------------------------------- */
static inline orc_f3 mat_mul_pt(const orc_mat4 *M, orc_f3 v)
{   /* operator*(Matrix4, float3), recon_volumeHelper.cuh:106-117 */
    orc_f3 r;
    r.x = M->m[0] * v.x + M->m[1] * v.y + M->m[2] * v.z + M->m[3];
    r.y = M->m[4] * v.x + M->m[5] * v.y + M->m[6] * v.z + M->m[7];
    r.z = M->m[8] * v.x + M->m[9] * v.y + M->m[10] * v.z + M->m[11];
    return r;
}

static inline orc_mat4 mat_mul(const orc_mat4 *A, const orc_mat4 *B)
{   /* operator*(Matrix4, Matrix4), recon_volumeHelper.cuh:134-145 */
    orc_mat4 t;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            t.m[4 * i + j] = A->m[4 * i + 0] * B->m[0 + j] + A->m[4 * i + 1] * B->m[4 + j] +
                             A->m[4 * i + 2] * B->m[8 + j] + A->m[4 * i + 3] * B->m[12 + j];
    return t;
}

/* CUDA float -> unsigned int conversion saturates: negatives and NaN -> 0 (quirk Q4). */
static inline uint32_t f2u_sat(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

/* ---- PSF: cuda2.cu:112-144 (calcPSF, USE_SINC_PSF=1) -------------------------------- */
static inline float orc_calc_psf(orc_f3 sPos, orc_f3 dim)
{
    const float sigmaz = dim.z / 2.3548f;
    sPos.x = sPos.x * dim.x / 2.3548f;
    sPos.y = sPos.y * dim.y / 2.3548f;
    float x = sqrtf(sPos.x * sPos.x + sPos.y * sPos.y);
    float R = 3.14159265359f * x;
    float si = (R == 0.0f) ? 1.0f : sinf(R) / R;       /* deviation D5 */
    return si * si * expf((-sPos.z * sPos.z) / (2.0f * sigmaz * sigmaz));
}

float orc_psf_value(float px, float py, float pz, float dx, float dy, float dz)
{
    orc_f3 p = { px, py, pz }, d = { dx, dy, dz };
    return orc_calc_psf(p, d);
}

/* getPSFParamsPrecomp, cuda2.cu:164-174.  psf_c = d_PSFI2W * ((PSFsize-1)*0.5), passed in. */
static inline float orc_psf_params(orc_f3 *ofsPos, orc_f3 c, int ox, int oy, int oz, const orc_mat4 *comb,
                                   orc_f3 slicePos, orc_f3 sliceDim, orc_f3 psf_c)
{
    ofsPos->x = (float)ox + c.x;
    ofsPos->y = (float)oy + c.y;
    ofsPos->z = (float)oz + c.z;
    orc_f3 p2 = mat_mul_pt(comb, *ofsPos);
    orc_f3 d;
    d.x = (p2.x - slicePos.x) * sliceDim.x;
    d.y = (p2.y - slicePos.y) * sliceDim.y;
    d.z = (p2.z - slicePos.z) * sliceDim.z;
    d.x -= psf_c.x; d.y -= psf_c.y; d.z -= psf_c.z;
    return orc_calc_psf(d, sliceDim);
}

/* Geometry handed to every PSF pass (what SetSliceMatrices / setSliceDims / generatePSFVolume
 * upload in the reference: cuda2.cu:741-750, 815-833, 870-907). */
typedef struct {
    int S, Nx, Ny;                 /* padded slice cube */
    int vx, vy, vz;                /* volume size */
    const orc_mat4 *I2W, *W2I;     /* per slice image<->world           */
    const orc_mat4 *T, *Tinv;      /* per slice slice->volume transform  */
    const orc_f3 *dims;            /* per slice voxel size (dx,dy,thickness) */
    orc_mat4 RI2W, RW2I;           /* volume image<->world */
    orc_f3 psf_c;                  /* PSF centre offset (approximately 0) */
} orc_geom;

static inline void pixel_setup(const orc_geom *g, int k, int x, int y, orc_mat4 *comb, orc_f3 *c, orc_f3 *slicePos)
{   /* cuda2.cu:222-226 (identical in K1/K2/K3) */
    orc_mat4 t = mat_mul(&g->W2I[k], &g->Tinv[k]);
    *comb = mat_mul(&t, &g->RI2W);
    slicePos->x = (float)x; slicePos->y = (float)y; slicePos->z = 0.0f;
    orc_f3 p = mat_mul_pt(&g->RW2I, mat_mul_pt(&g->T[k], mat_mul_pt(&g->I2W[k], *slicePos)));
    c->x = roundf(p.x); c->y = roundf(p.y); c->z = roundf(p.z);
}

/* ---- K1: gaussianReconstructionKernel3D_tex, cuda2.cu:176-295 ------------------------
 * recon / volweights are zeroed first (cuda2.cu:2410-2411); psf_sums persists between calls
 * (allocated zeroed at cuda2.cu:1567 and only written when sume > 0.5); voxel_count zeroed
 * (cuda2.cu:2409).  Then equalizeVol (cuda2.cu:2312-2327).  voxel_num[k] = number of pixels
 * of slice k with at least one masked in-volume tap (deviation D4).                      */
void orc_gaussian_reconstruction(const orc_geom *g, const float *slices, const float *scales, const float *mask,
                                 float *recon, float *volweights, float *psf_sums, int *voxel_count, int *voxel_num,
                                 int equalize)
{
    const size_t V = (size_t)g->vx * g->vy * g->vz;
    const size_t P = (size_t)g->Nx * g->Ny;
    double *acc = (double *)calloc(2 * V, sizeof(double));
    double *accw = acc + V;
    memset(voxel_count, 0, sizeof(int) * P * g->S);

    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    (void)nthreads;
    for (int k = 0; k < g->S; ++k) {
        int count = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : count)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float s = slices[idx];
            if (s == -1.0f) continue;
            const orc_f3 sliceDim = g->dims[k];
            s = s * scales[k];
            orc_mat4 comb; orc_f3 c, slicePos;
            pixel_setup(g, k, x, y, &comb, &c, &slicePos);
            const int dim = ORC_PSF_SUPPORT, centre = (ORC_PSF_SUPPORT - 1) / 2;

            float sume = 0;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        orc_f3 ofs;
                        float psfval = orc_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, slicePos, sliceDim, g->psf_c);
                        if ((double)fabsf(oldPSF - psfval) < ORC_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(ofs.x), ay = f2u_sat(ofs.y), az = f2u_sat(ofs.z);
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) sume += psfval;
                    }
                }
            if (sume > 0.5f) psf_sums[idx] = sume; else continue;

            int addvox = 0;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        orc_f3 ofs;
                        float psfval = orc_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, slicePos, sliceDim, g->psf_c);
                        if ((double)fabsf(oldPSF - psfval) < ORC_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(roundf(ofs.x)), ay = f2u_sat(roundf(ofs.y)), az = f2u_sat(roundf(ofs.z));
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                float a = psfval, b = psfval * s;
#pragma omp atomic
                                accw[v] += (double)a;
#pragma omp atomic
                                acc[v] += (double)b;
                                addvox = 1;
                            }
                        }
                    }
                }
            if (addvox) { voxel_count[idx] += 1; count += 1; }
        }
        if (voxel_num) voxel_num[k] = count;
    }
    for (size_t v = 0; v < V; ++v) {
        float a = (float)acc[v], b = (float)accw[v];
        volweights[v] = b;
        recon[v] = equalize ? ((b != 0) ? a / b : a) : a;     /* equalizeVol, cuda2.cu:2312-2327 */
    }
    free(acc);
}

/* ---- K2: simulateSlicesKernel3D_tex, cuda2.cu:298-404; slice_inside cuda2.cu:2742-2752 */
void orc_simulate_slices(const orc_geom *g, const float *slices, const float *psf_sums, const float *recon,
                         const float *mask, float *simslices, float *simweights, signed char *siminside,
                         unsigned char *slice_inside)
{
    const size_t P = (size_t)g->Nx * g->Ny;
    for (int k = 0; k < g->S; ++k) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float s = slices[idx];
            if (s == -1.0f) continue;
            float sume = psf_sums[idx];
            if (sume == 0.0f) continue;
            float sim = 0, weight = 0; int inside = 0;
            const orc_f3 sliceDim = g->dims[k];
            orc_mat4 comb; orc_f3 c, slicePos;
            pixel_setup(g, k, x, y, &comb, &c, &slicePos);
            const int dim = ORC_PSF_SUPPORT, centre = (ORC_PSF_SUPPORT - 1) / 2;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        orc_f3 ofs;
                        float psfval = orc_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, slicePos, sliceDim, g->psf_c);
                        if ((double)fabsf(oldPSF - psfval) < ORC_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(roundf(ofs.x)), ay = f2u_sat(roundf(ofs.y)), az = f2u_sat(roundf(ofs.z));
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                sim += psfval * recon[v];
                                weight += psfval;
                                inside = 1;
                            }
                        }
                    }
                }
            if (weight > 0) {
                simslices[idx] = sim / weight;
                simweights[idx] = weight;
                siminside[idx] = (signed char)inside;
            }
        }
        if (slice_inside) {
            int n = 0;
            for (size_t i = 0; i < P; ++i) n += (siminside[(size_t)k * P + i] == 1);
            slice_inside[k] = n > 0;
        }
    }
}

/* ---- K3: SuperresolutionKernel3D_tex, cuda2.cu:408-522 (addon/cmap zeroed, cuda2.cu:2202-2203) */
void orc_superresolution_backproject(const orc_geom *g, const float *slices, const float *weights,
                                     const float *simslices, const float *slice_weights, const float *scales,
                                     const float *mask, const float *psf_sums, float *addon, float *cmap)
{
    const size_t V = (size_t)g->vx * g->vy * g->vz;
    const size_t P = (size_t)g->Nx * g->Ny;
    double *acc = (double *)calloc(2 * V, sizeof(double));
    double *accc = acc + V;
    for (int k = 0; k < g->S; ++k) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float s = slices[idx];
            if (s == -1.0f) continue;
            float sume = psf_sums[idx];
            if (sume == 0.0f) continue;
            float w = weights[idx], ss = simslices[idx];
            float slice_weight = slice_weights[k], scale = scales[k];
            float sliceVal = s * scale;
            if (ss > 0.0f) sliceVal = sliceVal - ss; else sliceVal = 0.0f;
            const orc_f3 sliceDim = g->dims[k];
            orc_mat4 comb; orc_f3 c, slicePos;
            pixel_setup(g, k, x, y, &comb, &c, &slicePos);
            const int dim = ORC_PSF_SUPPORT, centre = (ORC_PSF_SUPPORT - 1) / 2;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        orc_f3 ofs;
                        float psfval = orc_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, slicePos, sliceDim, g->psf_c);
                        if ((double)fabsf(oldPSF - psfval) < ORC_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(ofs.x), ay = f2u_sat(ofs.y), az = f2u_sat(ofs.z);
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                float a = psfval * w * slice_weight * sliceVal;
                                float b = psfval * w * slice_weight;
#pragma omp atomic
                                acc[v] += (double)a;
#pragma omp atomic
                                accc[v] += (double)b;
                            }
                        }
                    }
                }
        }
    }
    for (size_t v = 0; v < V; ++v) { addon[v] = (float)acc[v]; cmap[v] = (float)accc[v]; }
    free(acc);
}

/* ---- K4 + K5: AdaptiveRegularizationPrep cuda2.cu:1944-1969, AdaptiveRegularizationKernel
 * cuda2.cu:2046-2117, constants cuda2.cu:666-695.  `original` = volume before the gradient
 * step (cuda2.cu:2138-2141).  In place on recon/addon/cmap.  Deviation D3: neighbours are
 * read from a frozen copy of the post-K4 volume.                                           */
static const int orc_dirs[13][3] = {
    { 1, 0, -1 }, { 0, 1, -1 }, { 1, 1, -1 }, { 1, -1, -1 }, { 1, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 },
    { 1, -1, 0 }, { 1, 0, 1 }, { 0, 1, 1 }, { 1, 1, 1 }, { 1, -1, 1 }, { 0, 0, 1 } };

void orc_regularize(int vx, int vy, int vz, float *recon, float *addon, float *cmap, int adaptive, float alpha,
                    float min_intensity, float max_intensity, float delta, float lambda)
{
    const size_t V = (size_t)vx * vy * vz;
    float *original = (float *)malloc(V * sizeof(float));
    float *frozen = (float *)malloc(V * sizeof(float));
    memcpy(original, recon, V * sizeof(float));
    float factor[13];
    for (int i = 0; i < 13; i++) {
        float f = 0;
        for (int j = 0; j < 3; j++) f += fabsf((float)orc_dirs[i][j]);
        factor[i] = 1.0f / f;
    }
    /* K4 */
    for (size_t v = 0; v < V; ++v) {
        if (!adaptive) {
            if (cmap[v] != 0) { addon[v] = addon[v] / cmap[v]; cmap[v] = 1.0f; }
        }
        float r = recon[v] + addon[v] * alpha;
        if ((double)r < (double)min_intensity * 0.9) r = (float)((double)min_intensity * 0.9);
        if ((double)r > (double)max_intensity * 1.1) r = (float)((double)max_intensity * 1.1);
        recon[v] = r;
    }
    memcpy(frozen, recon, V * sizeof(float));
    /* K5 */
#pragma omp parallel for schedule(static)
    for (int z = 0; z < vz; ++z)
        for (int y = 0; y < vy; ++y)
            for (int x = 0; x < vx; ++x) {
                const size_t p = x + (size_t)y * vx + (size_t)z * vx * vy;
                float val = 0, valW = 0, sum = 0;
                for (int i = 0; i < 13; i++) {
                    int x2 = x + orc_dirs[i][0], y2 = y + orc_dirs[i][1], z2 = z + orc_dirs[i][2];
                    int in2 = x2 >= 0 && x2 < vx && y2 >= 0 && y2 < vy && z2 >= 0 && z2 < vz;
                    size_t p2 = in2 ? (x2 + (size_t)y2 * vx + (size_t)z2 * vx * vy) : 0;
                    if (in2) {
                        float bi = 0.0f;
                        if (!(cmap[p] <= 0 || cmap[p2] <= 0)) {   /* AdaptiveRegularization1, cuda2.cu:2046-2057 */
                            float diff = (original[p2] - original[p]) * sqrtf(factor[i]) / delta;
                            bi = factor[i] / sqrtf(1.0f + diff * diff);
                        }
                        val += bi * frozen[p2] * cmap[p2];
                        valW += bi * cmap[p2];
                        sum += bi;
                    }
                    int x3 = x - orc_dirs[i][0], y3 = y - orc_dirs[i][1], z3 = z - orc_dirs[i][2];
                    int in3 = x3 >= 0 && x3 < vx && y3 >= 0 && y3 < vy && z3 >= 0 && z3 < vz;
                    if (in3 && in2) {                     /* cuda2.cu:2089-2093: both must be inside */
                        size_t p3 = x3 + (size_t)y3 * vx + (size_t)z3 * vx * vy;
                        float bi = 0.0f;
                        if (!(cmap[p3] <= 0 || cmap[p2] <= 0)) {   /* called with (pos3, pos2), cuda2.cu:2095 */
                            float diff = (original[p2] - original[p3]) * sqrtf(factor[i]) / delta;
                            bi = factor[i] / sqrtf(1.0f + diff * diff);
                        }
                        val += bi * frozen[p3] * cmap[p3];
                        valW += bi * cmap[p3];
                        sum += bi;
                    }
                }
                val -= sum * frozen[p] * cmap[p];
                valW -= sum * cmap[p];
                val = frozen[p] * cmap[p] + alpha * lambda / (delta * delta) * val;
                valW = cmap[p] + alpha * lambda / (delta * delta) * valW;
                recon[p] = (valW > 0.0f) ? val / valW : 0.0f;
            }
    free(original);
    free(frozen);
}

/* ---- K12: InitializeEMValuesKernel, cuda2.cu:3241-3267 ------------------------------- */
void orc_initialize_em_values(size_t n, const float *slices, float *weights)
{
    for (size_t i = 0; i < n; ++i) weights[i] = (slices[i] != -1) ? 1.0f : 0.0f;
}

/* ---- K11: InitializeRobustStatistics, cuda2.cu:2243-2308 ----------------------------- */
float orc_initialize_robust_statistics(size_t n, const float *slices, const signed char *siminside,
                                       const float *simslices, const float *simweights, double *out_sum,
                                       double *out_num)
{
    double sa = 0, sb = 0;
    for (size_t i = 0; i < n; ++i) {
        if (slices[i] != -1 && siminside[i] == 1 && simweights[i] > 0.99) {
            float sval = slices[i] - simslices[i];
            sa += (double)(sval * sval);
            sb += 1.0;
        }
    }
    if (out_sum) *out_sum = sa;
    if (out_num) *out_num = sb;
    return (float)sa / (float)sb;
}

/* G_ / M_: cuda2.cu:62-70 */
static inline float orc_G(float x, float s) { return ORC_STEP * expf(-x * x / (2.0f * s)) / (sqrtf(6.28f * s)); }
static inline float orc_M(float m) { return m * ORC_STEP; }

/* ---- K7 + K8: EStepKernel3D_tex cuda2.cu:2766-2813; slice potentials cuda2.cu:2816-2911 */
void orc_estep(int S, int Nx, int Ny, const float *slices, const float *simslices, const float *simweights,
               const float *scales, float m_, float sigma_, float mix_, float *weights, float *slice_potential)
{
    const size_t P = (size_t)Nx * Ny;
    memset(weights, 0, sizeof(float) * P * S);              /* cuda2.cu:2881 */
    for (int k = 0; k < S; ++k) {
        double sum = 0, num = 0;
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = slices[idx], sw = simweights[idx];
            if (!((s == -1) || sw <= 0)) {
                float sliceVal = s * scales[k];
                sliceVal -= simslices[idx];
                float g = orc_G(sliceVal, sigma_);
                float m = orc_M(m_);
                float weight = (g * mix_) / (g * mix_ + m * (1.0f - mix_));
                weights[idx] = weight;
            }
            if (sw > 0.99) {                                 /* transformSlicePotential */
                float v = (float)((1.0 - weights[idx]) * (1.0 - weights[idx]));
                sum += (double)v; num += 1.0;
            }
        }
        slice_potential[k] = (num > 0) ? sqrtf((float)sum / (float)num) : -1.0f;
    }
}

/* ---- K9: MStep, cuda2.cu:2966-3112.  out5 = {sum e^2 w, sum w, n, min e, max e} with the
 * min/max seeded with 0 (cuda2.cu:3103).  mstep_scales = the HOST h_scales vector (cuda2.cu:3093). */
void orc_mstep_sums(int S, int Nx, int Ny, const float *slices, const float *weights, const float *simslices,
                    const float *simweights, const float *mstep_scales, double *out5)
{
    const size_t P = (size_t)Nx * Ny;
    double sigma = 0, mix = 0, num = 0; float mn = 0.0f, mx = 0.0f;
    for (int k = 0; k < S; ++k)
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            if (slices[idx] != -1.0f && simweights[idx] > 0.99f) {
                float e = (slices[idx] * mstep_scales[k]) - simslices[idx];
                sigma += (double)(e * e * weights[idx]);
                mix += (double)weights[idx];
                num += 1.0;
                if (e < mn) mn = e;
                if (e > mx) mx = e;
            }
        }
    out5[0] = sigma; out5[1] = mix; out5[2] = num; out5[3] = mn; out5[4] = mx;
}

/* Host part of Reconstruction::MStep, cuda2.cu:3014-3072 */
void orc_mstep_finish(const double *sums5, int iter, float step, float *sigma_, float *mix_, float *m_)
{
    float sigma = (float)sums5[0], mix = (float)sums5[1], num = (float)sums5[2];
    float min_ = FLT_MAX, max_ = FLT_MIN;
    min_ = fminf(min_, (float)sums5[3]);
    max_ = fmaxf(max_, (float)sums5[4]);
    if (mix > 0) *sigma_ = sigma / mix;
    if (*sigma_ < step * step / 6.28f) *sigma_ = step * step / 6.28f;
    if (iter > 1) *mix_ = mix / num;
    *m_ = 1.0f / (max_ - min_);
}

/* ---- K10: CalculateScaleVector, cuda2.cu:3142-3239 ----------------------------------- */
void orc_calculate_scale_vector(int S, int Nx, int Ny, const float *slices, const float *weights,
                                const float *simslices, const float *simweights, float *scale_vec)
{
    const size_t P = (size_t)Nx * Ny;
    for (int k = 0; k < S; ++k) {
        double num = 0, den = 0;
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = slices[idx];
            if ((s == -1.0f) || simweights[idx] <= 0.99f) continue;
            num += (double)(weights[idx] * s * simslices[idx]);
            den += (double)(weights[idx] * s * s);
        }
        scale_vec[k] = ((float)den != 0.0f) ? (float)num / (float)den : 1.0f;
    }
}

/* ---- K15 maskVolumeKernel cuda2.cu:3313-3326 ----------------------------------------- */
void orc_mask_volume(size_t V, float *recon, const float *mask)
{
    for (size_t v = 0; v < V; ++v) if (mask[v] == 0) recon[v] = -1;
}

/* ---- K16 ScaleVolume cuda2.cu:3386-3470 ---------------------------------------------- */
float orc_scale_volume(int S, int Nx, int Ny, size_t V, const float *slices, const float *weights,
                       const float *simslices, const float *simweights, const float *slice_weights, float *recon)
{
    const size_t P = (size_t)Nx * Ny;
    double num = 0, den = 0;
    for (int k = 0; k < S; ++k)
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = slices[idx];
            if (s == -1) continue;
            if (simweights[idx] <= 0.99) continue;
            float ss = simslices[idx], w = weights[idx], sw = slice_weights[k];
            num += (double)(w * sw * s * ss);
            den += (double)(w * sw * ss * ss);
        }
    float scale = (float)(num / den);
    if (recon) for (size_t v = 0; v < V; ++v) if (recon[v] > 0) recon[v] = recon[v] * scale;
    return scale;
}

/* ---- K17 RestoreSliceIntensitiesKernel cuda2.cu:3349-3367 ---------------------------- */
void orc_restore_slice_intensities(int S, int Nx, int Ny, float *slices, const float *stack_factors, const int *stack_index)
{
    const size_t P = (size_t)Nx * Ny;
    for (int k = 0; k < S; ++k) {
        float f = stack_factors[stack_index[k]];
        for (size_t i = 0; i < P; ++i) {
            float s = slices[(size_t)k * P + i];
            if (s > 0) slices[(size_t)k * P + i] = s / f;
        }
    }
}

/* ---- a5: slice-level EM on the host, GPU.cc:3184-3440 (irtkReconstruction::EStepGPU) ---
 * G(x,s) GPU.cc include irtkReconstructionGPU.h:529-532 (double).  state = {sigma_s, mix_s,
 * mean_s, mean_s2, sigma_s2} in/out (floats as in the reference members).               */
static double orc_Gd(double x, double s, double step) { return step * exp(-x * x / (2 * s)) / (sqrt(6.28 * s)); }

void orc_host_slice_em(int S, float *slice_potential, const float *scale, float *slice_weight,
                       const int *force_excluded, int n_force, const int *small_slices, int n_small,
                       double step, float *state5)
{
    float sigma_s = state5[0], mix_s = state5[1], mean_s = state5[2], mean_s2 = state5[3], sigma_s2 = state5[4];
    for (int i = 0; i < n_force; i++) slice_potential[force_excluded[i]] = -1;
    for (int i = 0; i < n_small; i++) slice_potential[small_slices[i]] = -1;
    for (int i = 0; i < S; i++) if ((scale[i] < 0.2) || (scale[i] > 5)) slice_potential[i] = -1;

    double sum = 0, den = 0, sum2 = 0, den2 = 0, maxs = 0, mins = 1;
    for (int i = 0; i < S; i++)
        if (slice_potential[i] >= 0) {
            sum += slice_potential[i] * slice_weight[i];
            den += slice_weight[i];
            sum2 += slice_potential[i] * (1.0 - slice_weight[i]);
            den2 += (1.0 - slice_weight[i]);
            if (slice_potential[i] > maxs) maxs = slice_potential[i];
            if (slice_potential[i] < mins) mins = slice_potential[i];
        }
    mean_s = (den > 0) ? (float)(sum / den) : (float)mins;
    mean_s2 = (den2 > 0) ? (float)(sum2 / den2) : (float)((maxs + mean_s) / 2.0);

    sum = 0; den = 0; sum2 = 0; den2 = 0;
    for (int i = 0; i < S; i++)
        if (slice_potential[i] >= 0) {
            sum += (slice_potential[i] - mean_s) * (slice_potential[i] - mean_s) * slice_weight[i];
            den += slice_weight[i];
            sum2 += (slice_potential[i] - mean_s2) * (slice_potential[i] - mean_s2) * (1 - slice_weight[i]);
            den2 += (1 - slice_weight[i]);
        }
    if ((sum > 0) && (den > 0)) {
        sigma_s = (float)(sum / den);
        if (sigma_s < step * step / 6.28) sigma_s = (float)(step * step / 6.28);
    } else {
        sigma_s = 0.025f;
    }
    if ((sum2 > 0) && (den2 > 0)) {
        sigma_s2 = (float)(sum2 / den2);
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    } else {
        sigma_s2 = (mean_s2 - mean_s) * (mean_s2 - mean_s) / 4;
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    }

    for (int i = 0; i < S; i++) {
        if (slice_potential[i] == -1) { slice_weight[i] = 0; continue; }
        if ((den <= 0) || (mean_s2 <= mean_s)) { slice_weight[i] = 1; continue; }
        double gs1, gs2;
        if (slice_potential[i] < mean_s2) gs1 = orc_Gd(slice_potential[i] - mean_s, sigma_s, step); else gs1 = 0;
        if (slice_potential[i] > mean_s) gs2 = orc_Gd(slice_potential[i] - mean_s2, sigma_s2, step); else gs2 = 0;
        double likelihood = gs1 * mix_s + gs2 * (1 - mix_s);
        if (likelihood > 0)
            slice_weight[i] = (float)(gs1 * mix_s / likelihood);
        else {
            if (slice_potential[i] <= mean_s) slice_weight[i] = 1;
            if (slice_potential[i] >= mean_s2) slice_weight[i] = 0;
            if ((slice_potential[i] < mean_s2) && (slice_potential[i] > mean_s)) slice_weight[i] = 1;
        }
    }
    sum = 0; int num = 0;
    for (int i = 0; i < S; i++)
        if (slice_potential[i] >= 0) { sum += slice_weight[i]; num++; }
    mix_s = (num > 0) ? (float)(sum / num) : 0.9f;

    state5[0] = sigma_s; state5[1] = mix_s; state5[2] = mean_s; state5[3] = mean_s2; state5[4] = sigma_s2;
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
