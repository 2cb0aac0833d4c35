/*
 * reg_oracle.c -- CPU restatement of the GPU slice-to-volume registration (--useGPUReg) of
 * bkainz/fetalReconstruction: Reconstruction::registerSlicesToVolume and everything below it.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as svr_oracle.c).  PARITY STATUS: pinned against the reference's own
 * CUDA path (tests/golden/ref_reg_small.npz, ref_regtrace.npz; tests/test_ref_golden.py) up to the texture
 * unit's filter arithmetic, which a CPU cannot reproduce (deviation D6 below: similarities within 2e-3; the
 * CUDA path samples through the texture unit and agrees to 1.5e-6), and by the known-answer tests in
 * tests/test_reg_oracle.py.
 *
 * Reference lines followed (paths relative to source/reconstructionGPU2/, cuda2.cu = reconstruction_cuda2.cu):
 *   generateGaussianKernel / GaussX/YKernel / FilterGaussStack    GPUGauss/gaussfilter.cu:56-277
 *   genenerateRegistrationSlices (R1)                              cuda2.cu:3504-3529
 *   prepareSliceToVolumeReg (levels, steps, blurring)              cuda2.cu:3884-3900
 *   registerMultipleSlicesToVolume (driver loop)                   cuda2.cu:4001-4141
 *   evaluateCostsMultipleSlices                                    cuda2.cu:4150-4221
 *   adjustSamplingMatrixForCentralDifferences ... checkImprovement cuda2.cu:4224-4457
 *   averageIf, computeNCCAndReduce, addNccValues, writeSimilarities cuda2.cu:4459-4575
 *
 * The restatement is LITERAL, including behaviours that look unintended (they decide the numbers the
 * reference produces, so the drop-in reproduces them; DESIGN.md section 6 lists them):
 *   G1  the second memset of evaluateCosts clears temp_float[2*slices, 5*slices) -- offsets in units of
 *       ALL device slices -- while the per-slice accumulators live at offsets in units of ACTIVE slices
 *       ([2a,3a) NCC accumulator, [3a,6a) interleaved triplets).  Which accumulators are cleared between
 *       the three in-slice offsets therefore depends on (a, slices, position in the active list).
 *   G2  the sum/count of the sampled slice ([a,2a)) are cleared once per evaluation, not per in-slice
 *       offset: the mean used for offset o is the mean over offsets -1..o.
 *   G3  after the central-difference loop the working matrices still hold the last probe (rz - step);
 *       the line search starts from there.
 *   G4  the texture read has no +0.5 texel-centre offset: volumePos = i samples half-way between voxels
 *       i-1 and i (border texels are 0).
 * The texture unit's 1.8 fixed-point interpolation weights are modelled (round to nearest 1/256); what remains
 * against the hardware filter of a B200 is ~3e-4 relative on a sample (measured against the reference's own
 * sampled slices, tools/ref_reg_debug.py) -- the CUDA path samples through the texture unit itself and has no
 * such residue (deviation D6 applies to this CPU oracle only).
 * Not reproduced: the x32 factor of averageIf/computeNCCAndReduce (32 threads per block pass the
 * `threadIdx.x == 0` test: a power-of-two factor that cancels exactly in every ratio), fast-math
 * sin/cos/atan in the parameter kernels, and the nondeterministic order of the compacted active list
 * for more than 512 slices (a stable order is used).
 * Sums the reference forms with float tree reductions + atomics are accumulated in double and rounded
 * once to the float the reference keeps in memory.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float m[16]; } reg_mat4;

static inline void reg_mul_pt(const reg_mat4 *M, float x, float y, float z, float *ox, float *oy, float *oz)
{   /* operator*(Matrix4, float3), recon_volumeHelper.cuh:106-117 */
    *ox = M->m[0] * x + M->m[1] * y + M->m[2] * z + M->m[3];
    *oy = M->m[4] * x + M->m[5] * y + M->m[6] * z + M->m[7];
    *oz = M->m[8] * x + M->m[9] * y + M->m[10] * z + M->m[11];
}

/* ---- gaussfilter.cu:56-88 and :189-192.  Returns klength; half[] receives kernel[mid .. klength-1]. */
int reg_gauss_kernel(float sigma, float *half /* >= 32 floats */)
{
    int klength = (int)(sigma * 5);
    if (klength > 63) klength = 63;           /* MAX_LENGTH_SK = BLOCK_SIZE_SK_1-1 */
    if (klength < 7) klength = 7;
    klength -= 1 - klength % 2;
    float kernel[64];
    float sum = 0;
    int mid = (int)floorf(klength / 2.0f);
    for (int i = 0; i < klength; i++) {
        kernel[i] = (float)exp(-(float)abs(i - mid) * (float)abs(i - mid) / (2 * sigma * sigma));
        sum += kernel[i];
    }
    for (int i = 0; i < klength; i++) kernel[i] /= sum;
    for (int i = 0; i < (klength + 1) / 2; i++) half[i] = kernel[klength / 2 + i];
    return klength;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- FilterGaussStack, gaussfilter.cu:92-173,189-277: X pass into temp, Y pass into out (in place ok).
 * Padding: a centre of exactly -1 is passed through; neighbours are clamped to >= 0; reads outside the
 * image clamp to the edge (cudaBoundaryModeClamp). */
void reg_filter_gauss_stack(float *data, int W, int H, int n, float sigma)
{
    float half[32];
    int klength = reg_gauss_kernel(sigma, half);
    int K = (klength + 1) / 2;
    float *tmp = (float *)malloc(sizeof(float) * (size_t)W * H);
#pragma omp parallel for schedule(static)
    for (int z = 0; z < n; ++z) {
        /* tmp is shared by the serial loop only: allocate per iteration when threaded */
        float *t = (float *)malloc(sizeof(float) * (size_t)W * H);
        float *img = data + (size_t)z * W * H;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float v = img[y * W + x];
                if (v != -1) {
                    v = v * half[0];
                    for (int i = 1; i < K; ++i)
                        v = v + half[i] * (fmaxf(0.0f, img[y * W + clampi(x + i, 0, W - 1)]) +
                                           fmaxf(0.0f, img[y * W + clampi(x - i, 0, W - 1)]));
                }
                t[y * W + x] = v;
            }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float v = t[y * W + x];
                if (v != -1) {
                    v = v * half[0];
                    for (int i = 1; i < K; ++i)
                        v = v + half[i] * (fmaxf(0.0f, t[clampi(y + i, 0, H - 1) * W + x]) +
                                           fmaxf(0.0f, t[clampi(y - i, 0, H - 1) * W + x]));
                }
                img[y * W + x] = v;
            }
        free(t);
    }
    free(tmp);
}

/* ---- tex3D(reconstructedTex_, x/sx, y/sy, z/sz): linear filter, normalised coordinates, border mode
 * (cuda2.cu:3909-3959).  Texel space coordinate = u*N - 0.5 (quirk G4); texels outside the array read 0.
 * Weights rounded to the hardware's 1.8 fixed point; the filter arithmetic itself is plain float (deviation D6). */
static inline float reg_fetch(const float *vol, int vx, int vy, int vz, int x, int y, int z)
{
    if (x < 0 || y < 0 || z < 0 || x >= vx || y >= vy || z >= vz) return 0.0f;
    return vol[(size_t)z * vx * vy + (size_t)y * vx + x];
}

float reg_tex3d(const float *vol, int vx, int vy, int vz, float px, float py, float pz)
{
    /* px = volumePos.x / reconsize.x * N = volumePos.x up to one rounding of the divide/multiply pair; the
       unnormalisation is exact enough that we restate it as volumePos - 0.5 */
    float fx = px - 0.5f, fy = py - 0.5f, fz = pz - 0.5f;
    float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    float ax = fx - flx, ay = fy - fly, az = fz - flz;
    /* the texture unit keeps the interpolation weights in 1.8 fixed point */
    ax = floorf(ax * 256.0f + 0.5f) * (1.0f / 256.0f);
    ay = floorf(ay * 256.0f + 0.5f) * (1.0f / 256.0f);
    az = floorf(az * 256.0f + 0.5f) * (1.0f / 256.0f);
    /* guard the float -> int conversion for far-away positions */
    if (!(flx > -2.0f && flx < (float)vx + 1.0f && fly > -2.0f && fly < (float)vy + 1.0f && flz > -2.0f &&
          flz < (float)vz + 1.0f))
        return 0.0f;
    int x0 = (int)flx, y0 = (int)fly, z0 = (int)flz;
    float c000 = reg_fetch(vol, vx, vy, vz, x0, y0, z0), c100 = reg_fetch(vol, vx, vy, vz, x0 + 1, y0, z0);
    float c010 = reg_fetch(vol, vx, vy, vz, x0, y0 + 1, z0), c110 = reg_fetch(vol, vx, vy, vz, x0 + 1, y0 + 1, z0);
    float c001 = reg_fetch(vol, vx, vy, vz, x0, y0, z0 + 1), c101 = reg_fetch(vol, vx, vy, vz, x0 + 1, y0, z0 + 1);
    float c011 = reg_fetch(vol, vx, vy, vz, x0, y0 + 1, z0 + 1), c111 = reg_fetch(vol, vx, vy, vz, x0 + 1, y0 + 1, z0 + 1);
    float c00 = c000 + ax * (c100 - c000), c10 = c010 + ax * (c110 - c010);
    float c01 = c001 + ax * (c101 - c001), c11 = c011 + ax * (c111 - c011);
    float c0 = c00 + ay * (c10 - c00), c1 = c01 + ay * (c11 - c01);
    return c0 + az * (c1 - c0);
}

/* ---- R1 genenerateRegistrationSlices, cuda2.cu:3504-3529, for one slice into out[H][W] */
void reg_generate_slice(const float *vol, int vx, int vy, int vz, const reg_mat4 *reconW2I, const reg_mat4 *T,
                        const reg_mat4 *ofs, int W, int H, int insliceofs, float *out)
{
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            float wx, wy, wz, tx, ty, tz, px, py, pz;
            reg_mul_pt(ofs, (float)x, (float)y, (float)(insliceofs * 2), &wx, &wy, &wz);
            reg_mul_pt(T, wx, wy, wz, &tx, &ty, &tz);
            reg_mul_pt(reconW2I, tx, ty, tz, &px, &py, &pz);
            float val = reg_tex3d(vol, vx, vy, vz, px, py, pz);
            if (val < 0) val = -1.0f;
            out[y * W + x] = val;
        }
}

/* ---- state of one registration run (the dev_* buffers of cuda2.cu:4892-5018) */
typedef struct {
    int W, H, S;
    int vx, vy, vz;
    const float *vol;
    reg_mat4 reconW2I;
    const float *resampled;      /* dev_v_slices_resampled  [S][H][W] */
    float *blurred;              /* dev_v_slices_resampled_float      */
    float *reg;                  /* dev_regSlices                     */
    const reg_mat4 *ofs;         /* dev_d_slicesOfs                   */
    reg_mat4 *M, *Morig;         /* dev_recon_matrices(_orig)         */
    float *sim;                  /* dev_recon_similarities [5][S]     */
    float *grad;                 /* dev_recon_gradient     [7][S]     */
    int *active, *active2, *active_prev;
    float *tf;                   /* dev_temp_float [6*S] */
    int *ti;                     /* dev_temp_int   [2*S] */
    long long evals;             /* number of (slice, in-slice offset) cost evaluations */
} reg_state;

/* averageIf, cuda2.cu:4459-4496 */
static void reg_average_if(const float *img, int W, int H, float *sum, int *count)
{
    double s = 0; int c = 0;
    for (int i = 0; i < W * H; ++i) {
        float v = img[i];
        if (v > -1.0f) { ++c; s += v; }
    }
    *sum += (float)s;
    *count += c;
}

/* evaluateCostsMultipleSlices, cuda2.cu:4150-4221 (literal buffer layout, quirks G1/G2) */
static void reg_evaluate_costs(reg_state *st, int a, int level, float blur, int writeoffset, int writestep, int writenum)
{
    if (a == 0) return;
    const int S = st->S, W = st->W, H = st->H;
    const size_t P = (size_t)W * H;
    float *tf = st->tf; int *ti = st->ti;
    memset(tf, 0, sizeof(float) * 6 * a);
    memset(ti, 0, sizeof(int) * 2 * a);
    for (int t = 0; t < a; ++t) reg_average_if(st->blurred + P * st->active[t], W, H, &tf[t], &ti[t]);

    for (int insofs = -1; insofs <= 1; insofs++) {
#pragma omp parallel for schedule(dynamic)
        for (int t = 0; t < a; ++t) {
            int j = st->active[t];
            reg_generate_slice(st->vol, st->vx, st->vy, st->vz, &st->reconW2I, &st->M[j], &st->ofs[j], W, H, insofs,
                               st->reg + P * t);
        }
        reg_filter_gauss_stack(st->reg, W, H, a, blur);
        for (int t = 0; t < a; ++t) reg_average_if(st->reg + P * t, W, H, &tf[a + t], &ti[a + t]);
        memset(tf + 2 * S, 0, sizeof(float) * 3 * S);              /* G1: offsets in units of `slices` */
        /* computeNCCAndReduce(..., sums = tf, counts = ti, results = tf + 3a, slices = a, level + 1) */
#pragma omp parallel for schedule(dynamic)
        for (int t = 0; t < a; ++t) {
            const float *A = st->blurred + P * st->active[t];
            const float *B = st->reg + P * t;
            float avg_a = tf[t], avg_b = tf[a + t];
            if (avg_a != 0) avg_a /= ti[t];
            if (avg_b != 0) avg_b /= ti[a + t];
            double vxy = 0, vxx = 0, vyy = 0;
            for (int lin = 0; lin < W * H; ++lin) {
                float va = A[lin], vb = B[lin];
                if (va >= 0.0f && vb >= 0.0f && lin % (level + 1) == 0) {
                    float s = va - avg_a, stex = vb - avg_b;
                    vxy += s * stex; vxx += s * s; vyy += stex * stex;
                }
            }
            float *res = tf + 3 * a + 3 * t;
            res[0] += (float)vxy; res[1] += (float)vxx; res[2] += (float)vyy;
        }
        /* addNccValues(prevData = tf + 3a, result = tf + 2a, a), cuda2.cu:4552-4563 */
        for (int t = 0; t < a; ++t) {
            const float *prev = tf + 3 * a + 3 * t;
            float norm = prev[1] * prev[2];
            float res = 0;
            if (norm > 0) res = prev[0] / sqrtf(norm);
            tf[2 * a + t] += res;
        }
        st->evals += a;
    }
    /* writeSimilarities, cuda2.cu:4565-4575 */
    for (int t = 0; t < a; ++t) {
        float res = tf[2 * a + t];
        int slice = st->active[t];
        for (int i = 0; i < writenum; ++i) st->sim[(size_t)writeoffset * S + (size_t)S * writestep * i + slice] = res;
    }
}

/* Euler decomposition shared by adjustSamplingMatrixForCentralDifferences and gradientStep
 * (cuda2.cu:4245-4290, 4343-4390) */
static void reg_rot_params(const reg_mat4 *in, float p_rot[3])
{
    const float TOL = 0.000001f;
    float tmp = asinf(-1.0f * in->m[2]);
    if (fabsf(cosf(tmp)) > TOL) {
        p_rot[0] = atan2f(in->m[6], in->m[10]);
        p_rot[1] = tmp;
        p_rot[2] = atan2f(in->m[1], in->m[0]);
    } else {
        p_rot[0] = atan2f(-in->m[2] * in->m[4], -in->m[2] * in->m[8]);
        p_rot[1] = tmp;
        p_rot[2] = 0;
    }
}

static void reg_set_rotation(reg_mat4 *out, const float p_rot[3])
{
    float cosrx = cosf(p_rot[0]), cosry = cosf(p_rot[1]), cosrz = cosf(p_rot[2]);
    float sinrx = sinf(p_rot[0]), sinry = sinf(p_rot[1]), sinrz = sinf(p_rot[2]);
    out->m[0] = cosry * cosrz; out->m[1] = cosry * sinrz; out->m[2] = -sinry;
    out->m[4] = (sinrx * sinry * cosrz - cosrx * sinrz);
    out->m[5] = (sinrx * sinry * sinrz + cosrx * cosrz);
    out->m[6] = sinrx * cosry;
    out->m[8] = (cosrx * sinry * cosrz + sinrx * sinrz);
    out->m[9] = (cosrx * sinry * sinrz - sinrx * cosrz);
    out->m[10] = cosrx * cosry;
}

void reg_adjust_matrix(const reg_mat4 *in, reg_mat4 *out, int part, float step)
{   /* cuda2.cu:4231-4293 */
    const float pi = 3.14159265358979323846f;
    *out = *in;
    if (part < 3) {
        out->m[4 * part + 3] = in->m[4 * part + 3] + step;
    } else {
        float p_rot[3];
        reg_rot_params(in, p_rot);
        p_rot[part - 3] += step * pi / 180.0f;
        reg_set_rotation(out, p_rot);
    }
}

void reg_gradient_step(reg_mat4 *matrix, const float g[6], float step)
{   /* cuda2.cu:4335-4391 */
    const float pi = 3.14159265358979323846f;
    for (int p = 0; p < 3; ++p) matrix->m[4 * p + 3] = matrix->m[4 * p + 3] + step * g[p];
    float p_rot[3];
    reg_rot_params(matrix, p_rot);
    for (int p = 0; p < 3; ++p) p_rot[p] += g[p + 3] * step * pi / 180.0f;
    reg_set_rotation(matrix, p_rot);
}

/* checkImprovement, cuda2.cu:4393-4457 (stable compaction) */
static int reg_check_improvement(int *newActive, int a, int S, const int *active, const float *sim, int cursim, int prev, float eps)
{
    int n = 0;
    for (int t = 0; t < a; ++t) {
        int slice = active[t];
        if (sim[(size_t)cursim * S + slice] > sim[(size_t)prev * S + slice] + eps) newActive[n++] = slice;
    }
    return n;
}

/* One cost evaluation of all S slices for given transforms (test hook; = evaluateCostsMultipleSlices with all
 * slices active, after blurring the slices for `level`).  similarity[S]. */
void reg_evaluate(int W, int H, int S, const float *resampled, const float *ofs16, const float *vol, int vx, int vy,
                  int vz, float voxel, const float *reconW2I16, const float *transforms16, int level, float *similarity)
{
    reg_state st;
    memset(&st, 0, sizeof(st));
    const size_t P = (size_t)W * H;
    st.W = W; st.H = H; st.S = S; st.vx = vx; st.vy = vy; st.vz = vz; st.vol = vol;
    memcpy(&st.reconW2I, reconW2I16, sizeof(reg_mat4));
    st.resampled = resampled;
    st.blurred = (float *)malloc(sizeof(float) * P * S);
    st.reg = (float *)malloc(sizeof(float) * P * S);
    st.ofs = (const reg_mat4 *)ofs16;
    st.M = (reg_mat4 *)malloc(sizeof(reg_mat4) * S);
    memcpy(st.M, transforms16, sizeof(reg_mat4) * S);
    st.sim = (float *)calloc((size_t)5 * S, sizeof(float));
    st.active = (int *)malloc(sizeof(int) * S);
    st.tf = (float *)calloc((size_t)6 * S, sizeof(float));
    st.ti = (int *)calloc((size_t)2 * S, sizeof(int));
    for (int i = 0; i < S; ++i) st.active[i] = i;
    float blur = voxel / 2.0f;
    for (int i = 0; i < level; ++i) blur *= 2;
    memcpy(st.blurred, resampled, sizeof(float) * P * S);
    reg_filter_gauss_stack(st.blurred, W, H, S, blur);
    reg_evaluate_costs(&st, S, level, blur, 0, 1, 1);
    memcpy(similarity, st.sim, sizeof(float) * S);
    free(st.blurred); free(st.reg); free(st.M); free(st.sim); free(st.active); free(st.tf); free(st.ti);
}

/* registerMultipleSlicesToVolume, cuda2.cu:4001-4141, with the parameters of prepareSliceToVolumeReg
 * (cuda2.cu:3884-3900): 2 levels, 4 steps, 20 iterations, epsilon 1e-4, blurring voxel/2 * 2^level,
 * step 0.1 * 2^level.  transforms16 [S][16] in/out.  Returns the number of (slice, offset) evaluations. */
long long reg_register_slices(int W, int H, int S, const float *resampled, const float *ofs16, const float *vol,
                              int vx, int vy, int vz, float voxel, const float *reconW2I16, float *transforms16,
                              int n_levels, int n_steps, int n_iterations)
{
    reg_state st;
    memset(&st, 0, sizeof(st));
    const size_t P = (size_t)W * H;
    st.W = W; st.H = H; st.S = S; st.vx = vx; st.vy = vy; st.vz = vz; st.vol = vol;
    memcpy(&st.reconW2I, reconW2I16, sizeof(reg_mat4));
    st.resampled = resampled;
    st.blurred = (float *)malloc(sizeof(float) * P * S);
    st.reg = (float *)malloc(sizeof(float) * P * S);
    st.ofs = (const reg_mat4 *)ofs16;
    st.M = (reg_mat4 *)malloc(sizeof(reg_mat4) * S);
    st.Morig = (reg_mat4 *)malloc(sizeof(reg_mat4) * S);
    st.sim = (float *)calloc((size_t)5 * S, sizeof(float));
    st.grad = (float *)calloc((size_t)7 * S, sizeof(float));
    st.active = (int *)malloc(sizeof(int) * S);
    st.active2 = (int *)malloc(sizeof(int) * S);
    st.active_prev = (int *)malloc(sizeof(int) * S);
    st.tf = (float *)calloc((size_t)6 * S, sizeof(float));
    st.ti = (int *)calloc((size_t)2 * S, sizeof(int));
    memcpy(st.M, transforms16, sizeof(reg_mat4) * S);
    memcpy(st.Morig, transforms16, sizeof(reg_mat4) * S);

    const float Epsilon = 0.0001f;
    float Blurring[8], LengthOfSteps[8];
    Blurring[0] = voxel / 2.0f;
    for (int i = 0; i < n_levels; i++) LengthOfSteps[i] = (float)(0.1 * pow(2.0f, i));
    for (int i = 1; i < n_levels; i++) Blurring[i] = Blurring[i - 1] * 2;

    for (int level = n_levels - 1; level >= 0; --level) {
        float blur = Blurring[level];
        float StepSize = LengthOfSteps[level];
        memcpy(st.blurred, resampled, sizeof(float) * P * S);
        reg_filter_gauss_stack(st.blurred, W, H, S, blur);
        for (int s = 0; s < n_steps; s++) {
            for (int i = 0; i < S; ++i) st.active[i] = i;
            int a = S;
            for (int iter = 0; iter < n_iterations; iter++) {
                reg_evaluate_costs(&st, a, level, blur, 0, 1, 3);
                for (int p = 0; p < 6; ++p) {
                    for (int t = 0; t < a; ++t) reg_adjust_matrix(&st.Morig[st.active[t]], &st.M[st.active[t]], p, StepSize);
                    reg_evaluate_costs(&st, a, level, blur, 3, 0, 1);
                    for (int t = 0; t < a; ++t) reg_adjust_matrix(&st.Morig[st.active[t]], &st.M[st.active[t]], p, -StepSize);
                    reg_evaluate_costs(&st, a, level, blur, 4, 0, 1);
                    /* computeGradientCentralDiff(similarities + 3*slices, ...), cuda2.cu:4295-4308 */
                    for (int t = 0; t < a; ++t) {
                        int slice = st.active[t];
                        float dx = st.sim[(size_t)3 * S + slice] - st.sim[(size_t)4 * S + slice];
                        st.grad[(size_t)p * S + slice] = dx;
                        if (p == 0) st.grad[(size_t)6 * S + slice] = dx * dx;
                        else st.grad[(size_t)6 * S + slice] += dx * dx;
                    }
                }
                for (int t = 0; t < a; ++t) {                       /* normalizeGradient, cuda2.cu:4309-4323 */
                    int slice = st.active[t];
                    float norm = st.grad[(size_t)6 * S + slice];
                    if (norm > 0) norm = 1.0f / sqrtf(norm);
                    for (int j = 0; j < 6; ++j) st.grad[(size_t)j * S + slice] *= norm;
                }
                int prevActive = a;
                memcpy(st.active_prev, st.active, sizeof(int) * a);
                do {
                    for (int t = 0; t < a; ++t) st.sim[(size_t)2 * S + st.active[t]] = st.sim[st.active[t]];
                    for (int t = 0; t < a; ++t) {                   /* G3: M still holds the rz - step probe */
                        int slice = st.active[t];
                        float g[6];
                        for (int j = 0; j < 6; ++j) g[j] = st.grad[(size_t)j * S + slice];
                        reg_gradient_step(&st.M[slice], g, StepSize);
                    }
                    reg_evaluate_costs(&st, a, level, blur, 0, 1, 1);
                    a = reg_check_improvement(st.active2, a, S, st.active, st.sim, 0, 2, Epsilon);
                    int *tmp = st.active; st.active = st.active2; st.active2 = tmp;
                } while (a > 0);
                for (int t = 0; t < prevActive; ++t) {
                    int slice = st.active_prev[t];
                    float g[6];
                    for (int j = 0; j < 6; ++j) g[j] = st.grad[(size_t)j * S + slice];
                    reg_gradient_step(&st.M[slice], g, -StepSize);
                }
                memcpy(st.Morig, st.M, sizeof(reg_mat4) * S);
                a = reg_check_improvement(st.active, prevActive, S, st.active_prev, st.sim, 2, 1, Epsilon);
                if (a == 0) break;
            }
            StepSize /= 2.0f;
        }
    }
    memcpy(transforms16, st.M, sizeof(reg_mat4) * S);
    long long evals = st.evals;
    free(st.blurred); free(st.reg); free(st.M); free(st.Morig); free(st.sim); free(st.grad);
    free(st.active); free(st.active2); free(st.active_prev); free(st.tf); free(st.ti);
    return evals;
}
