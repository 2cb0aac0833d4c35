/*
 * pvr_oracle.c -- CPU restatement of the PVR (patch-to-volume) twins of the hot path of
 * bkainz/fetalReconstruction: patchBasedPSFReconstruction_gpu (P1), patchBasedSimulatePatches_gpu (P2),
 * patchBasedSuperresolution_gpu::run / regularize (P3, P4), ReconVolume::equalize (P5) and
 * patchBasedRobustStatistics_gpu (P6, P7), with the PVR constants of include/reconConfig.cuh and
 * include/pointSpreadFunction.cuh.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as svr_oracle.c).  PARITY STATUS: pinned against the reference's own PVR
 * CUDA path: its unmodified patchBased*_gpu.cu / reconVolume.cu compile for sm_100a behind oracle/ref_shim/
 * (libref_pvr.so), and every stage of one PVR iteration they produce on a B200 (tests/golden/ref_pvr_small.npz,
 * oracle/ref_runner_pvr.py) is what tests/test_ref_golden.py holds this file to (final volume 1.2e-5 relative RMS).
 * Not covered by the reference build: patch enumeration / extraction (IRTK host code), restated from the sources.
 *
 * Reference lines followed (paths relative to source/reconstructionGPU2/):
 *   PointSpreadFunction::sinc_pi / calcPSF / getPSFParamsPrecomp   include/pointSpreadFunction.cuh:45-116
 *   patchBasedPatchInitKernel (P0)                                  initPatchBasedRecon_gpu.cu:44-86
 *   patchBasedPSFReconstructionKernel (P1)                          patchBasedPSFReconstruction_gpu.cu:41-144
 *   patchBasedSimulatePatchesKernel (P2)                            patchBasedSimulatePatches_gpu.cu:41-125
 *   ReconVolume::getReconValueFromTexture / updateReconTex          reconVolume.cu:102-187
 *   patchBasedSuperresolution_gpuKernel (P3)                        patchBasedSuperresolution_gpu.cu:34-110
 *   AdaptiveRegularizationPrepKernel / AdaptiveRegularizationKernel patchBasedSuperresolution_gpu.cu:152-287
 *   equalizeVol (P5)                                                reconVolume.cu:59-100
 *   InitializeEMValuesKernel, EStepKernel, EStep host part, MStep, Scale, InitializeRobustStatistics
 *                                                                   patchBasedRobustStatistics_gpu.cu:41-868
 *
 * Data model: the reference keeps one PatchBasedVolume per stack (a pbb.x x pbb.y x nPatches grid).  The
 * oracle (and the CUDA library) concatenate the grids of all stacks into one float[nPatches][pby][pbx] cube
 * with per-patch voxel sizes (= the voxel size of the patch's stack, patchBasedVolume.cuh:111) and per-patch
 * matrices; the geometry struct is svr_oracle.c's orc_geom (S = number of patches).
 *
 * Documented deviations (as for SVR): x/y bounds are checked (P1/P3 do not, patchBasedPSFReconstruction_gpu.cu:51);
 * the regulariser reads a frozen copy of the post-step volume (the reference updates in place,
 * patchBasedSuperresolution_gpu.cu:222-262); scatter sums are accumulated in double and rounded once; the
 * texture read of P2 uses exact 1/8 weights (the hardware's 1.8 fixed-point weights are exactly 0.5 here).
 * Reproduced: epsilon skip with a FLOAT epsilon, mask-blind sume, sume accepted when > 1e-5, saturating
 * float->unsigned conversion, stale PSF sums, pixels outside the mask kept as 0 (not -1) by P0 and therefore
 * projected by P1/P2, E-step gated on the previous weight, __step = 1e-5 in G_/M_ but 1e-4 in the sigma floors,
 * the un-offset half-voxel texture read of P2 (an 8-voxel average), and -- in pvr_host_patch_em -- the
 * patch_potential[j] indexing without the stack offset (patchBasedRobustStatistics_gpu.cu:268,272).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PVR_STEP 0.00001f          /* __step, include/reconConfig.cuh:120 */
#define PVR_PSF_EPSILON 0.00001f   /* PSF_EPSILON (float literal), reconConfig.cuh:138 */
#define PVR_PSF_SUPPORT 12         /* MAX_PSF_SUPPORT, reconConfig.cuh:140 */

typedef struct { float m[16]; } pvr_mat4;
typedef struct { float x, y, z; } pvr_f3;
typedef struct {                     /* identical layout to orc_geom (svr_oracle.c) */
    int S, Nx, Ny;
    int vx, vy, vz;
    const pvr_mat4 *I2W, *W2I;
    const pvr_mat4 *T, *Tinv;
    const pvr_f3 *dims;
    pvr_mat4 RI2W, RW2I;
    pvr_f3 psf_c;
} pvr_geom;

static inline pvr_f3 mul_pt(const pvr_mat4 *M, pvr_f3 v)
{
    pvr_f3 r;
    r.x = M->m[0] * v.x + M->m[1] * v.y + M->m[2] * v.z + M->m[3];
    r.y = M->m[4] * v.x + M->m[5] * v.y + M->m[6] * v.z + M->m[7];
    r.z = M->m[8] * v.x + M->m[9] * v.y + M->m[10] * v.z + M->m[11];
    return r;
}
static inline pvr_mat4 mul(const pvr_mat4 *A, const pvr_mat4 *B)
{
    pvr_mat4 t;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            t.m[4 * i + j] = A->m[4 * i + 0] * B->m[0 + j] + A->m[4 * i + 1] * B->m[4 + j] +
                             A->m[4 * i + 2] * B->m[8 + j] + A->m[4 * i + 3] * B->m[12 + j];
    return t;
}
static inline uint32_t f2u_sat(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)f;
}

/* sinc_pi, pointSpreadFunction.cuh:45-72 (T = float) */
static inline float pvr_sinc_pi(float x)
{
    const float taylor_0_bound = FLT_EPSILON;
    const float taylor_2_bound = sqrtf(taylor_0_bound);
    const float taylor_n_bound = sqrtf(taylor_2_bound);
    if (fabsf(x) >= taylor_n_bound) return sinf(x) / x;
    float result = 1.0f;
    if (fabsf(x) >= taylor_0_bound) {
        float x2 = x * x;
        result -= x2 / 6.0f;
        if (fabsf(x) >= taylor_2_bound) result += (x2 * x2) / 120.0f;
    }
    return result;
}

/* calcPSF, pointSpreadFunction.cuh:75-107 (USE_SINC_PSF) */
static inline float pvr_calc_psf(pvr_f3 sPos, pvr_f3 dim)
{
    const float sigmaz = dim.z;
    sPos.x = sPos.x * dim.x / 2.3548f;
    sPos.y = sPos.y * dim.y / 2.3548f;
    float x = sqrtf(sPos.x * sPos.x + sPos.y * sPos.y);
    float R = 3.14159265359f * x;
    float si = pvr_sinc_pi(R);
    return si * si * expf((-sPos.z * sPos.z) / (2.0f * sigmaz * sigmaz));
}

float pvr_psf_value(float px, float py, float pz, float dx, float dy, float dz)
{
    pvr_f3 p = { px, py, pz }, d = { dx, dy, dz };
    return pvr_calc_psf(p, d);
}

/* getPSFParamsPrecomp, pointSpreadFunction.cuh:109-116 */
static inline float pvr_psf_params(pvr_f3 *ofsPos, pvr_f3 c, int ox, int oy, int oz, const pvr_mat4 *comb,
                                   pvr_f3 patchPos, pvr_f3 patchDim, pvr_f3 psf_c)
{
    ofsPos->x = (float)ox + c.x; ofsPos->y = (float)oy + c.y; ofsPos->z = (float)oz + c.z;
    pvr_f3 p2 = mul_pt(comb, *ofsPos);
    pvr_f3 d;
    d.x = (p2.x - patchPos.x) * patchDim.x;
    d.y = (p2.y - patchPos.y) * patchDim.y;
    d.z = (p2.z - patchPos.z) * patchDim.z / 2.5f;
    d.x -= psf_c.x; d.y -= psf_c.y; d.z -= psf_c.z;
    return pvr_calc_psf(d, patchDim);
}

static inline void pixel_setup(const pvr_geom *g, int k, int x, int y, pvr_mat4 *comb, pvr_f3 *c, pvr_f3 *pos)
{   /* patchBasedPSFReconstruction_gpu.cu:78-81: W2I * (InvTransformation * reconstructedI2W) */
    pvr_mat4 t = mul(&g->Tinv[k], &g->RI2W);
    *comb = mul(&g->W2I[k], &t);
    pos->x = (float)x; pos->y = (float)y; pos->z = 0.0f;
    pvr_f3 p = mul_pt(&g->RW2I, mul_pt(&g->T[k], mul_pt(&g->I2W[k], *pos)));
    c->x = roundf(p.x); c->y = roundf(p.y); c->z = roundf(p.z);
}

/* ---- P0: patchBasedPatchInitKernel, initPatchBasedRecon_gpu.cu:44-86, for patches [p0, p0+n) of one stack.
 * patches[] must be zero-filled beforehand (PatchBasedVolume::reset, patchBasedVolume.cuh:196-210). */
void pvr_patch_init(const pvr_geom *g, int p0, int n, const float *stack, int sx, int sy, int sz,
                    const float *stackW2I16, const signed char *mask, const char *spx /* [S][4096] or NULL */,
                    float *patches)
{
    const pvr_mat4 *sw2i = (const pvr_mat4 *)stackW2I16;
    const size_t P = (size_t)g->Nx * g->Ny;
    for (int k = p0; k < p0 + n; ++k)
        for (int y = 0; y < g->Ny; ++y)
            for (int x = 0; x < g->Nx; ++x) {
                pvr_f3 pos = { (float)x, (float)y, 0.0f };
                /* getValueFromPatchCoords, patchBasedVolume.cuh:233-249: stackW2I * p.I2W * scoord */
                pvr_mat4 m = mul(sw2i, &g->I2W[k]);
                pvr_f3 sc = mul_pt(&m, pos);
                float s = 0.0f;
                if (sc.x >= 0 && sc.x < sx && sc.y >= 0 && sc.y < sy && sc.z >= 0 && sc.z < sz) {
                    unsigned int idx = (unsigned int)(sc.x + sc.y * sx + sc.z * sx * sy);
                    s = stack[idx];
                }
                if (s == -1.0f) continue;
                pvr_mat4 ti = mul(&g->T[k], &g->I2W[k]);       /* patch.Transformation*patch.I2W*patchPos */
                pvr_f3 w = mul_pt(&ti, pos);
                pvr_f3 vp = mul_pt(&g->RW2I, w);
                uint32_t ax = f2u_sat(vp.x), ay = f2u_sat(vp.y), az = f2u_sat(vp.z);
                int masked = 0;                                 /* ReconVolume::isMasked, reconVolume.cuh:241-254 */
                if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                    signed char mv = mask[ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy];
                    masked = !(mv == -1 || mv == 0);
                }
                const size_t idx = (size_t)k * P + (size_t)y * g->Nx + x;
                if (spx) {
                    int on = spx[(size_t)k * 4096 + x + 64 * y] == '1';
                    if (masked && on) patches[idx] = s;
                    else if (masked && !on) patches[idx] = -1.0f;
                } else if (masked) patches[idx] = s;
            }
}

/* ---- P1: patchBasedPSFReconstructionKernel.  recon / volweights are ACCUMULATED INTO (the caller resets them:
 * ReconVolume::reset, irtkPatchBasedReconstruction.cpp:492); psf_sums persists between calls. */
void pvr_psf_reconstruction(const pvr_geom *g, const float *patches, const float *scales, const signed char *mask,
                            const char *spx, float *recon, float *volweights, float *psf_sums)
{
    const size_t V = (size_t)g->vx * g->vy * g->vz;
    const size_t P = (size_t)g->Nx * g->Ny;
    double *acc = (double *)calloc(2 * V, sizeof(double));
    double *accw = acc + V;
    const int dim = PVR_PSF_SUPPORT, centre = (PVR_PSF_SUPPORT - 1) / 2;
    for (int k = 0; k < g->S; ++k) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float s = patches[idx];
            if (s == -1.0f) continue;
            s = s * scales[k];
            const pvr_f3 patchDim = g->dims[k];
            pvr_mat4 comb; pvr_f3 c, pos;
            pixel_setup(g, k, x, y, &comb, &c, &pos);
            const int spx_on = spx ? (spx[(size_t)k * 4096 + x + 64 * y] == '1') : 1;
            float sume = 0;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        pvr_f3 ofs;
                        float psfval = pvr_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, pos, patchDim, g->psf_c);
                        if (fabsf(oldPSF - psfval) < PVR_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(ofs.x), ay = f2u_sat(ofs.y), az = f2u_sat(ofs.z);
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz && spx_on) sume += psfval;
                    }
                }
            if ((sume > PVR_PSF_EPSILON) || isnan(sume)) psf_sums[idx] = sume; else continue;
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        pvr_f3 ofs;
                        float psfval = pvr_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, pos, patchDim, g->psf_c);
                        if (fabsf(oldPSF - psfval) < PVR_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(roundf(ofs.x)), ay = f2u_sat(roundf(ofs.y)), az = f2u_sat(roundf(ofs.z));
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                float a = psfval, b = s * psfval;
#pragma omp atomic
                                accw[v] += (double)a;
#pragma omp atomic
                                acc[v] += (double)b;
                            }
                        }
                    }
                }
        }
    }
    for (size_t v = 0; v < V; ++v) {
        recon[v] = (float)((double)recon[v] + acc[v]);
        volweights[v] = (float)((double)volweights[v] + accw[v]);
    }
    free(acc);
}

/* ---- P5: equalizeVol, reconVolume.cu:59-75 */
void pvr_equalize(size_t V, float *recon, const float *volweights)
{
    for (size_t v = 0; v < V; ++v) {
        float a = recon[v], b = volweights[v];
        recon[v] = (b != 0) ? a / b : a;
    }
}

/* ---- tex3D(reconTex_, pos / size) with linear filtering, normalised coordinates and border mode read at an
 * integer voxel position (reconVolume.cu:169-187): texel coordinate pos - 0.5, i.e. the average of the 8 voxels
 * pos + {-1,0}^3 with out-of-volume voxels reading 0. */
void pvr_texture_volume(int vx, int vy, int vz, const float *recon, float *tex)
{
#pragma omp parallel for schedule(static)
    for (int z = 0; z < vz; ++z)
        for (int y = 0; y < vy; ++y)
            for (int x = 0; x < vx; ++x) {
                float s = 0.0f;
                for (int dz = -1; dz <= 0; ++dz)
                    for (int dy = -1; dy <= 0; ++dy)
                        for (int dx = -1; dx <= 0; ++dx) {
                            int xx = x + dx, yy = y + dy, zz = z + dz;
                            if (xx >= 0 && yy >= 0 && zz >= 0) s += 0.125f * recon[xx + (size_t)yy * vx + (size_t)zz * vx * vy];
                        }
                tex[x + (size_t)y * vx + (size_t)z * vx * vy] = s;
            }
}

/* ---- P2: patchBasedSimulatePatchesKernel */
void pvr_simulate_patches(const pvr_geom *g, const float *patches, const float *psf_sums, const float *recon,
                          const signed char *mask, float *simpatches, float *simweights, signed char *siminside)
{
    const size_t V = (size_t)g->vx * g->vy * g->vz;
    const size_t P = (size_t)g->Nx * g->Ny;
    float *tex = (float *)malloc(sizeof(float) * V);
    pvr_texture_volume(g->vx, g->vy, g->vz, recon, tex);
    const int dim = PVR_PSF_SUPPORT, centre = (PVR_PSF_SUPPORT - 1) / 2;
    for (int k = 0; k < g->S; ++k) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float s = patches[idx];
            if (s == -1.0f) continue;
            float sume = psf_sums[idx];
            if (sume == 0.0f) continue;
            float sim = 0, weight = 0; int inside = 0;
            const pvr_f3 patchDim = g->dims[k];
            pvr_mat4 comb; pvr_f3 c, pos;
            pixel_setup(g, k, x, y, &comb, &c, &pos);
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        pvr_f3 ofs;
                        float psfval = pvr_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, pos, patchDim, g->psf_c);
                        if (fabsf(oldPSF - psfval) < PVR_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(roundf(ofs.x)), ay = f2u_sat(roundf(ofs.y)), az = f2u_sat(roundf(ofs.z));
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                sim += psfval * tex[v];
                                weight += psfval;
                                inside = 1;
                            }
                        }
                    }
                }
            if (weight > 0) {
                simpatches[idx] = sim / weight;
                simweights[idx] = weight;
                siminside[idx] = (signed char)inside;
            }
        }
    }
    free(tex);
}

/* ---- P3: patchBasedSuperresolution_gpuKernel; addon / cmap are ACCUMULATED INTO (resetAddonCmap by the caller) */
void pvr_superresolution(const pvr_geom *g, const float *patches, const float *weights, const float *simpatches,
                         const float *patch_weights, const float *scales, const signed char *mask,
                         const float *psf_sums, float *addon, float *cmap)
{
    const size_t V = (size_t)g->vx * g->vy * g->vz;
    const size_t P = (size_t)g->Nx * g->Ny;
    double *acc = (double *)calloc(2 * V, sizeof(double));
    double *accc = acc + V;
    const int dim = PVR_PSF_SUPPORT, centre = (PVR_PSF_SUPPORT - 1) / 2;
    for (int k = 0; k < g->S; ++k) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int pix = 0; pix < (int)P; ++pix) {
            const int x = pix % g->Nx, y = pix / g->Nx;
            const size_t idx = (size_t)k * P + pix;
            float patchVal = patches[idx];
            if (patchVal == -1.0f) continue;
            patchVal = patchVal * scales[k];
            float sume = psf_sums[idx];
            if (sume == 0.0f) continue;
            float w = weights[idx], ss = simpatches[idx], patch_weight = patch_weights[k];
            if (ss > 0.0f) patchVal = patchVal - ss; else patchVal = 0.0f;
            const pvr_f3 patchDim = g->dims[k];
            pvr_mat4 comb; pvr_f3 c, pos;
            pixel_setup(g, k, x, y, &comb, &c, &pos);
            for (int z = 0; z < dim; z++)
                for (int yy = 0; yy < dim; yy++) {
                    float oldPSF = FLT_MAX;
                    for (int xx = 0; xx < dim; xx++) {
                        pvr_f3 ofs;
                        float psfval = pvr_psf_params(&ofs, c, xx - centre, yy - centre, z - centre, &comb, pos, patchDim, g->psf_c);
                        if (fabsf(oldPSF - psfval) < PVR_PSF_EPSILON) continue;
                        oldPSF = psfval;
                        uint32_t ax = f2u_sat(roundf(ofs.x)), ay = f2u_sat(roundf(ofs.y)), az = f2u_sat(roundf(ofs.z));
                        if (ax < (uint32_t)g->vx && ay < (uint32_t)g->vy && az < (uint32_t)g->vz) {
                            size_t v = ax + (size_t)ay * g->vx + (size_t)az * g->vx * g->vy;
                            if (mask[v] != 0) {
                                psfval /= sume;
                                float a = psfval * w * patch_weight * patchVal;
                                float b = psfval * w * patch_weight;
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
    for (size_t v = 0; v < V; ++v) {
        addon[v] = (float)((double)addon[v] + acc[v]);
        cmap[v] = (float)((double)cmap[v] + accc[v]);
    }
    free(acc);
}

/* ---- P4: regularize, patchBasedSuperresolution_gpu.cu:152-287 (original = volume before the step, :269-271) */
static const int pvr_dirs[13][3] = {
    { 1, 0, -1 }, { 0, 1, -1 }, { 1, 1, -1 }, { 1, -1, -1 }, { 1, 0, 0 }, { 0, 1, 0 }, { 1, 1, 0 },
    { 1, -1, 0 }, { 1, 0, 1 }, { 0, 1, 1 }, { 1, 1, 1 }, { 1, -1, 1 }, { 0, 0, 1 } };

void pvr_regularize(int vx, int vy, int vz, float *recon, float *addon, float *cmap, int adaptive, float alpha,
                    float min_intensity, float max_intensity, float delta, float lambda)
{
    const size_t V = (size_t)vx * vy * vz;
    float *original = (float *)malloc(V * sizeof(float));
    float *frozen = (float *)malloc(V * sizeof(float));
    memcpy(original, recon, V * sizeof(float));
    float factor[13];
    for (int i = 0; i < 13; i++) {
        float f = 0;
        for (int j = 0; j < 3; j++) f += fabsf((float)pvr_dirs[i][j]);
        factor[i] = 1.0f / f;
    }
    for (size_t v = 0; v < V; ++v) {            /* AdaptiveRegularizationPrepKernel */
        float a = addon[v], c = cmap[v], r = recon[v];
        if (!adaptive) { if (c != 0) { a = a / c; c = 1.0f; } }
        r = r + a * alpha;
        if (r < min_intensity * 0.9f) r = min_intensity * 0.9f;
        if (r > max_intensity * 1.1f) r = max_intensity * 1.1f;
        recon[v] = r; addon[v] = a; cmap[v] = c;
    }
    memcpy(frozen, recon, V * sizeof(float));
#pragma omp parallel for schedule(static)
    for (int z = 0; z < vz; ++z)
        for (int y = 0; y < vy; ++y)
            for (int x = 0; x < vx; ++x) {
                const size_t p = x + (size_t)y * vx + (size_t)z * vx * vy;
                float val = 0, valW = 0, sum = 0;
                for (int i = 0; i < 13; i++) {
                    int x2 = x + pvr_dirs[i][0], y2 = y + pvr_dirs[i][1], z2 = z + pvr_dirs[i][2];
                    int in2 = x2 >= 0 && x2 < vx && y2 >= 0 && y2 < vy && z2 >= 0 && z2 < vz;
                    size_t p2 = in2 ? (x2 + (size_t)y2 * vx + (size_t)z2 * vx * vy) : 0;
                    if (in2) {
                        float bi = 0.0f;
                        if (!(cmap[p] <= 0 || cmap[p2] <= 0)) {
                            float diff = (original[p2] - original[p]) * sqrtf(factor[i]) / delta;
                            bi = (float)(factor[i] / sqrt(1.0 + diff * diff));
                        }
                        val += bi * frozen[p2] * cmap[p2];
                        valW += bi * cmap[p2];
                        sum += bi;
                    }
                    int x3 = x - pvr_dirs[i][0], y3 = y - pvr_dirs[i][1], z3 = z - pvr_dirs[i][2];
                    int in3 = x3 >= 0 && x3 < vx && y3 >= 0 && y3 < vy && z3 >= 0 && z3 < vz;
                    if (in3 && in2) {
                        size_t p3 = x3 + (size_t)y3 * vx + (size_t)z3 * vx * vy;
                        float bi = 0.0f;
                        if (!(cmap[p3] <= 0 || cmap[p2] <= 0)) {
                            float diff = (original[p2] - original[p3]) * sqrtf(factor[i]) / delta;
                            bi = (float)(factor[i] / sqrt(1.0 + diff * diff));
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
                recon[p] = (valW > 0.0) ? val / valW : 0.0f;
            }
    free(original);
    free(frozen);
}

/* ---- P6: InitializeEMValuesKernel, patchBasedRobustStatistics_gpu.cu:57-78 */
void pvr_initialize_em_values(size_t n, const float *patches, float *weights)
{
    for (size_t i = 0; i < n; ++i) weights[i] = (patches[i] != -1 && patches[i] != 0) ? 1.0f : 0.0f;
}

static inline float pvr_G(float x, float s) { return PVR_STEP * expf(-x * x / (2.0f * s)) / (sqrtf(6.28f * s)); }
static inline float pvr_M(float m) { return m * PVR_STEP; }

/* ---- P6 + P7: EStepKernel (:110-150) and the per-patch potentials (:254-275); weights updated in place */
void pvr_estep(int S, int Nx, int Ny, const float *patches, const float *simpatches, const float *simweights,
               const float *scales, float m_, float sigma_, float mix_, float *weights, float *patch_potential)
{
    const size_t P = (size_t)Nx * Ny;
    for (int k = 0; k < S; ++k) {
        double sum = 0, num = 0;
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = patches[idx], sw = weights[idx];
            if (!((s == -1) || sw <= 0)) {
                float patchVal = s * scales[k];
                patchVal -= simpatches[idx];
                float g = pvr_G(patchVal, sigma_);
                float m = pvr_M(m_);
                float weight = (float)((g * mix_) / (g * mix_ + m * (1.0 - mix_)));
                weights[idx] = weight;
            }
            if (simweights[idx] > 0.99) {
                double d = 1.0 - weights[idx];
                sum += (float)(d * d);
                num += 1.0;
            }
        }
        patch_potential[k] = (num > 0) ? sqrtf((float)sum / (float)num) : -1.0f;
    }
}

/* ---- EStep host part, patchBasedRobustStatistics_gpu.cu:277-520, LITERAL including the patch_potential[j]
 * indexing without the stack offset (:268,272): potentials_per_patch is the correctly indexed vector of all
 * stacks; it is replayed through the reference's loop.  state5 = {sigma_s, mix_s, mean_s, mean_s2, sigma_s2}. */
void pvr_host_patch_em(int n_stacks, const int *patches_per_stack, const float *potentials_per_patch,
                       const float *scale, float *patch_weight, float step, float *state5, float *potential_used)
{
    int numPatches = 0;
    for (int i = 0; i < n_stacks; ++i) numPatches += patches_per_stack[i];
    float *pot = (float *)calloc((size_t)(numPatches > 0 ? numPatches : 1), sizeof(float));
    int ofs = 0;
    for (int i = 0; i < n_stacks; ++i) {
        for (int j = 0; j < patches_per_stack[i]; ++j) pot[j] = potentials_per_patch[ofs + j];
        ofs += patches_per_stack[i];
    }
    float sigma_s = state5[0], mix_s = state5[1], mean_s, mean_s2, sigma_s2;
    for (int i = 0; i < numPatches; ++i)
        if ((scale[i] < 0.2) || (scale[i] > 5)) pot[i] = -1;
    double sum = 0, den = 0, sum2 = 0, den2 = 0, maxs = 0, mins = 1;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) {
            sum += pot[i] * patch_weight[i];
            den += patch_weight[i];
            sum2 += pot[i] * (1.0 - patch_weight[i]);
            den2 += (1.0 - patch_weight[i]);
            if (pot[i] > maxs) maxs = pot[i];
            if (pot[i] < mins) mins = pot[i];
        }
    mean_s = (den > 0) ? (float)(sum / den) : (float)mins;
    mean_s2 = (den2 > 0) ? (float)(sum2 / den2) : (float)((maxs + mean_s) / 2.0);
    sum = den = sum2 = den2 = 0;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) {
            sum += (pot[i] - mean_s) * (pot[i] - mean_s) * patch_weight[i];
            den += patch_weight[i];
            sum2 += (pot[i] - mean_s2) * (pot[i] - mean_s2) * (1 - patch_weight[i]);
            den2 += (1 - patch_weight[i]);
        }
    if ((sum > 0) && (den > 0)) {
        sigma_s = (float)(sum / den);
        if (sigma_s < step * step / 6.28) sigma_s = (float)(step * step / 6.28);
    } else sigma_s = 0.025f;
    if ((sum2 > 0) && (den2 > 0)) {
        sigma_s2 = (float)(sum2 / den2);
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    } else {
        sigma_s2 = (mean_s2 - mean_s) * (mean_s2 - mean_s) / 4;
        if (sigma_s2 < step * step / 6.28) sigma_s2 = (float)(step * step / 6.28);
    }
    for (int i = 0; i < numPatches; ++i) {
        if (pot[i] == -1) { patch_weight[i] = 0; continue; }
        if ((den <= 0) || (mean_s2 <= mean_s)) { patch_weight[i] = 1; continue; }
        double gs1 = (pot[i] < mean_s2) ? pvr_G(pot[i] - mean_s, sigma_s) : 0;
        double gs2 = (pot[i] > mean_s) ? pvr_G(pot[i] - mean_s2, sigma_s2) : 0;
        double likelihood = gs1 * mix_s + gs2 * (1 - mix_s);
        if (likelihood > 0) patch_weight[i] = (float)(gs1 * mix_s / likelihood);
        else {
            if (pot[i] <= mean_s) patch_weight[i] = 1;
            if (pot[i] >= mean_s2) patch_weight[i] = 0;
            if ((pot[i] < mean_s2) && (pot[i] > mean_s)) patch_weight[i] = 1;
        }
    }
    sum = 0; int num = 0;
    for (int i = 0; i < numPatches; ++i)
        if (pot[i] >= 0) { sum += patch_weight[i]; num++; }
    mix_s = (num > 0) ? (float)(sum / num) : 0.9f;
    state5[0] = sigma_s; state5[1] = mix_s; state5[2] = mean_s; state5[3] = mean_s2; state5[4] = sigma_s2;
    if (potential_used) memcpy(potential_used, pot, sizeof(float) * numPatches);
    free(pot);
}

/* ---- MStep sums, patchBasedRobustStatistics_gpu.cu:524-624: {sum e^2 w, sum w, n, min e, max e}, reduce init 0 */
void pvr_mstep_sums(int S, int Nx, int Ny, const float *patches, const float *weights, const float *simpatches,
                    const float *simweights, const float *scales, double *out5)
{
    const size_t P = (size_t)Nx * Ny;
    double sigma = 0, mix = 0, num = 0; float mn = 0.0f, mx = 0.0f;
    for (int k = 0; k < S; ++k)
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = patches[idx], sw = simweights[idx];
            if (s != -1.0f && sw > 0.99f) {
                float e = (s * scales[k]) - simpatches[idx];
                sigma += (float)(e * e * weights[idx]);
                mix += weights[idx];
                num += 1.0;
                if (e < mn) mn = e;
                if (e > mx) mx = e;
            }
        }
    out5[0] = sigma; out5[1] = mix; out5[2] = num; out5[3] = mn; out5[4] = mx;
}

/* host arithmetic of MStep, patchBasedRobustStatistics_gpu.cu:574-640 (one stack's worth of sums; the per-stack
 * accumulation of the reference is a plain sum / min / max and is done by the caller) */
void pvr_mstep_finish(const double *sums5, int iter, float step, float *sigma_, float *mix_, float *m_)
{
    float sigma = (float)sums5[0], mix = (float)sums5[1], num = (float)sums5[2];
    float min_ = FLT_MAX, max_ = FLT_MIN;
    if ((float)sums5[3] < min_) min_ = (float)sums5[3];
    if ((float)sums5[4] > max_) max_ = (float)sums5[4];
    if (mix > 0) *sigma_ = sigma / mix;
    if (*sigma_ < step * step / 6.28f) *sigma_ = step * step / 6.28f;
    if (iter > 1) *mix_ = mix / num;
    *m_ = 1.0f / (max_ - min_);
}

/* ---- Scale, patchBasedRobustStatistics_gpu.cu:642-744 */
void pvr_scale(int S, int Nx, int Ny, const float *patches, const float *weights, const float *simpatches,
               const float *simweights, float *scale_vec)
{
    const size_t P = (size_t)Nx * Ny;
    for (int k = 0; k < S; ++k) {
        double num = 0, den = 0;
        for (size_t i = 0; i < P; ++i) {
            const size_t idx = (size_t)k * P + i;
            float s = patches[idx], sw = simweights[idx];
            if ((s == -1.0f) || sw <= 0.99f) continue;
            num += (float)(weights[idx] * s * simpatches[idx]);
            den += (float)(weights[idx] * s * s);
        }
        scale_vec[k] = ((float)den != 0.0f) ? (float)num / (float)den : 1.0f;
    }
}

/* ---- InitializeRobustStatistics, patchBasedRobustStatistics_gpu.cu:746-851: returns sigma = sa / sb */
float pvr_initialize_robust_statistics(size_t n, const float *patches, const signed char *siminside,
                                       const float *simpatches, const float *simweights, double *sa_out, double *sb_out)
{
    double sa = 0, sb = 0;
    for (size_t i = 0; i < n; ++i)
        if (patches[i] != -1 && siminside[i] == 1 && simweights[i] > 0.99) {
            float d = patches[i] - simpatches[i];
            sa += (float)(d * d);
            sb += 1.0;
        }
    if (sa_out) *sa_out = sa;
    if (sb_out) *sb_out = sb;
    return (float)sa / (float)sb;
}
