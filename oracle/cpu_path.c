/*
 * cpu_path.c -- restatement of the reference's CPU (--useCPU) slice <-> volume projections: the sparse slice-to-volume
 * matrix of irtkReconstruction::CoeffInit and the functors that apply it.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (same rules as svr_oracle.c): used by bench.py's cpu_baseline / --impl reference legs
 * as "the reference's IRTK/TBB CPU path timed on the host cores" (SURVEY.md section 8d), which cannot be built here
 * (Boost/TBB/GSL absent).  Not pinned against the reference (nothing of that path runs here): a timing baseline whose
 * numerical sanity is checked by tests/test_cpu_path.py (rows of the matrix sum to 1, a unit volume simulates to 1, the
 * phantom is recovered).
 *
 * Reference lines followed (source/reconstructionGPU2/irtkReconstructionGPU.cc):
 *   ParallelCoeffInit::operator()      :2305-2612   Gaussian PSF (sigma = 1.2 dx / 2.3548 in plane, dz / 2.3548 through plane)
 *                                                   on a grid of voxel size res / quality_factor spanning 2 voxels of the
 *                                                   slice, each PSF point trilinearly distributed onto the volume grid
 *                                                   (normalised over the in-volume neighbours, dropped when no neighbour
 *                                                   lies in the mask), collected in a (dim^3) transformed-PSF scratch
 *                                                   around the rounded centre voxel, non-zero entries stored per pixel
 *   CoeffInit (volume weights)         :2627-2665
 *   GaussianReconstruction (CPU)       :2765-2860   serial, as in the reference
 *   ParallelSimulateSlices             :1090-1144   parallel over slices
 *   ParallelSuperresolution            :3940-4022   parallel_reduce over slices with per-thread addon / confidence map
 * Parallelised with OpenMP where the reference uses tbb::parallel_for / parallel_reduce.  Arithmetic in double, as IRTK.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int S, Nx, Ny, vx, vy, vz;
    /* per slice: offsets into (vox, val) per pixel, CSR style */
    long **start;      /* [S][Nx*Ny + 1] */
    int **vox;         /* [S][nnz_s] linear voxel index */
    float **val;       /* [S][nnz_s] POINT3D::value is a float */
} cpu_coeffs;

static void mat34_from16(const float *m, double *o) { for (int i = 0; i < 12; ++i) o[i] = m[i]; }
static void mat34_mul(const double *a, const double *b, double *c)
{   /* c = a * b, both 3x4 affine (implicit last row 0 0 0 1) */
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c[4 * i + j] = a[4 * i] * b[j] + a[4 * i + 1] * b[4 + j] + a[4 * i + 2] * b[8 + j];
        c[4 * i + 3] = a[4 * i] * b[3] + a[4 * i + 1] * b[7] + a[4 * i + 2] * b[11] + a[4 * i + 3];
    }
}
static inline void apply34(const double *m, double x, double y, double z, double *ox, double *oy, double *oz)
{
    *ox = m[0] * x + m[1] * y + m[2] * z + m[3];
    *oy = m[4] * x + m[5] * y + m[6] * z + m[7];
    *oz = m[8] * x + m[9] * y + m[10] * z + m[11];
}

void cpu_coeff_free(cpu_coeffs *c)
{
    if (!c) return;
    for (int s = 0; s < c->S; ++s) { free(c->start[s]); free(c->vox[s]); free(c->val[s]); }
    free(c->start); free(c->vox); free(c->val); free(c);
}

/* ParallelCoeffInit.  slices [S][Ny][Nx] (-1 = padding), I2W / T [S][16] row-major, dims [S][3] (dx, dy, thickness),
 * RW2I[16] volume world->image, mask [vz][vy][vx] float, res = volume voxel size, quality = _quality_factor.
 * slice_inside[S] (out): the slice has a PSF point with a neighbour in the mask. */
cpu_coeffs *cpu_coeff_init(int S, int Nx, int Ny, const float *slices, const float *I2W, const float *T, const float *dims,
                           const float *RW2I, int vx, int vy, int vz, const float *mask, double res, double quality,
                           unsigned char *slice_inside)
{
    cpu_coeffs *c = (cpu_coeffs *)calloc(1, sizeof(cpu_coeffs));
    c->S = S; c->Nx = Nx; c->Ny = Ny; c->vx = vx; c->vy = vy; c->vz = vz;
    c->start = (long **)calloc(S > 0 ? S : 1, sizeof(long *));
    c->vox = (int **)calloc(S > 0 ? S : 1, sizeof(int *));
    c->val = (float **)calloc(S > 0 ? S : 1, sizeof(float *));
    double rw[12];
    mat34_from16(RW2I, rw);
    const int P = Nx * Ny;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < S; ++s) {
        const float *slice = slices + (size_t)s * P;
        const double dx = dims[3 * s], dy = dims[3 * s + 1], dz = dims[3 * s + 2];
        const double sigmax = 1.2 * dx / 2.3548, sigmay = 1.2 * dy / 2.3548, sigmaz = dz / 2.3548;
        const double size = res / quality;
        const int xDim = (int)round(2 * dx / size), yDim = (int)round(2 * dy / size), zDim = (int)round(2 * dz / size);
        const int np = xDim * yDim * zDim;
        double *PSF = (double *)malloc(sizeof(double) * (np > 0 ? np : 1));
        double sum = 0;
        for (int i = 0; i < xDim; ++i) for (int j = 0; j < yDim; ++j) for (int k = 0; k < zDim; ++k) {
            const double x = (i - 0.5 * (xDim - 1)) * size, y = (j - 0.5 * (yDim - 1)) * size, z = (k - 0.5 * (zDim - 1)) * size;
            const double v = exp(-x * x / (2 * sigmax * sigmax) - y * y / (2 * sigmay * sigmay) - z * z / (2 * sigmaz * sigmaz));
            PSF[(i * yDim + j) * zDim + k] = v;
            sum += v;
        }
        for (int i = 0; i < np; ++i) PSF[i] /= sum;
        const int dim = (int)(floor(ceil(sqrt((double)(xDim * xDim + yDim * yDim + zDim * zDim)) * size / res) / 2)) * 2 + 1 + 2;
        const int centre = (dim - 1) / 2;
        double *tPSF = (double *)malloc(sizeof(double) * dim * dim * dim);
        double a[12], b[12], m[12];                 /* m = RW2I * T * I2W: slice image -> volume image */
        mat34_from16(I2W + 16 * s, a);
        mat34_from16(T + 16 * s, b);
        double tb[12];
        mat34_mul(b, a, tb);
        mat34_mul(rw, tb, m);
        long *start = (long *)malloc(sizeof(long) * (P + 1));
        long cap = (long)P * 32 + 64, nnz = 0;
        int *vox = (int *)malloc(sizeof(int) * cap);
        float *val = (float *)malloc(sizeof(float) * cap);
        int inside_slice = 0;
        for (int j = 0; j < Ny; ++j) for (int i = 0; i < Nx; ++i) {
            const int pix = j * Nx + i;
            start[pix] = nnz;
            if (slice[pix] == -1.0f) continue;
            double x, y, z;
            apply34(m, i, j, 0, &x, &y, &z);
            const int tx = (int)round(x), ty = (int)round(y), tz = (int)round(z);
            memset(tPSF, 0, sizeof(double) * dim * dim * dim);
            for (int ii = 0; ii < xDim; ++ii) for (int jj = 0; jj < yDim; ++jj) for (int kk = 0; kk < zDim; ++kk) {
                double px = (ii - 0.5 * (xDim - 1)) * size / dx + i, py = (jj - 0.5 * (yDim - 1)) * size / dy + j,
                       pz = (kk - 0.5 * (zDim - 1)) * size / dz;
                apply34(m, px, py, pz, &x, &y, &z);
                const int nx = (int)floor(x), ny = (int)floor(y), nz = (int)floor(z);
                double wsum = 0;
                int inside = 0;
                for (int l = nx; l <= nx + 1; ++l) if (l >= 0 && l < vx)
                    for (int mm = ny; mm <= ny + 1; ++mm) if (mm >= 0 && mm < vy)
                        for (int n = nz; n <= nz + 1; ++n) if (n >= 0 && n < vz) {
                            wsum += (1 - fabs(l - x)) * (1 - fabs(mm - y)) * (1 - fabs(n - z));
                            if (mask[((size_t)n * vy + mm) * vx + l] == 1.0f) { inside = 1; inside_slice = 1; }
                        }
                if (wsum <= 0 || !inside) continue;
                const double pv = PSF[(ii * yDim + jj) * zDim + kk];
                for (int l = nx; l <= nx + 1; ++l) if (l >= 0 && l < vx)
                    for (int mm = ny; mm <= ny + 1; ++mm) if (mm >= 0 && mm < vy)
                        for (int n = nz; n <= nz + 1; ++n) if (n >= 0 && n < vz) {
                            const double w = (1 - fabs(l - x)) * (1 - fabs(mm - y)) * (1 - fabs(n - z));
                            const int aa = l - tx + centre, bb = mm - ty + centre, cc = n - tz + centre;
                            if (aa < 0 || aa >= dim || bb < 0 || bb >= dim || cc < 0 || cc >= dim) continue;   /* the reference exits here */
                            tPSF[(aa * dim + bb) * dim + cc] += pv * w / wsum;
                        }
            }
            for (int ii = 0; ii < dim; ++ii) for (int jj = 0; jj < dim; ++jj) for (int kk = 0; kk < dim; ++kk) {
                const double v = tPSF[(ii * dim + jj) * dim + kk];
                if (v > 0) {
                    if (nnz == cap) { cap *= 2; vox = (int *)realloc(vox, sizeof(int) * cap); val = (float *)realloc(val, sizeof(float) * cap); }
                    const int X = ii + tx - centre, Y = jj + ty - centre, Z = kk + tz - centre;
                    vox[nnz] = (Z * vy + Y) * vx + X;
                    val[nnz] = (float)v;
                    ++nnz;
                }
            }
        }
        start[P] = nnz;
        c->start[s] = start; c->vox[s] = vox; c->val[s] = val;
        if (slice_inside) slice_inside[s] = (unsigned char)inside_slice;
        free(PSF); free(tPSF);
    }
    return c;
}

long cpu_coeff_nnz(const cpu_coeffs *c)
{
    long n = 0;
    for (int s = 0; s < c->S; ++s) n += c->start[s][c->Nx * c->Ny];
    return n;
}

/* CoeffInit :2627-2640 */
void cpu_volume_weights(const cpu_coeffs *c, double *volw)
{
    const size_t V = (size_t)c->vx * c->vy * c->vz;
    const int P = c->Nx * c->Ny;
    memset(volw, 0, sizeof(double) * V);
    for (int s = 0; s < c->S; ++s) {
        const long n = c->start[s][P];
        for (long k = 0; k < n; ++k) volw[c->vox[s][k]] += c->val[s][k];
    }
}

/* GaussianReconstruction :2765-2830 (serial in the reference).  recon [V] out (already divided by the volume weights). */
void cpu_gaussian_reconstruction(const cpu_coeffs *c, const float *slices, const float *scales, const double *volw, float *recon,
                                 int *voxel_num)
{
    const size_t V = (size_t)c->vx * c->vy * c->vz;
    const int P = c->Nx * c->Ny;
    double *acc = (double *)calloc(V, sizeof(double));
    for (int s = 0; s < c->S; ++s) {
        int cnt = 0;
        const long *st = c->start[s];
        for (int pix = 0; pix < P; ++pix) {
            const float v = slices[(size_t)s * P + pix];
            if (v == -1.0f) continue;
            const double sv = (double)v * scales[s];
            if (st[pix + 1] > st[pix]) cnt++;
            for (long k = st[pix]; k < st[pix + 1]; ++k) acc[c->vox[s][k]] += c->val[s][k] * sv;
        }
        if (voxel_num) voxel_num[s] = cnt;
    }
    for (size_t v = 0; v < V; ++v) recon[v] = volw[v] != 0 ? (float)(acc[v] / volw[v]) : 0.0f;
    free(acc);
}

/* ParallelSimulateSlices :1090-1144 */
void cpu_simulate_slices(const cpu_coeffs *c, const float *slices, const float *recon, const float *mask, float *sim, float *simw,
                         signed char *siminside, int *slice_inside)
{
    const int P = c->Nx * c->Ny;
#pragma omp parallel for schedule(dynamic)
    for (int s = 0; s < c->S; ++s) {
        const long *st = c->start[s];
        int any = 0;
        for (int pix = 0; pix < P; ++pix) {
            const size_t idx = (size_t)s * P + pix;
            sim[idx] = 0; simw[idx] = 0; siminside[idx] = 0;
            if (slices[idx] == -1.0f) continue;
            double acc = 0, w = 0;
            for (long k = st[pix]; k < st[pix + 1]; ++k) {
                const int v = c->vox[s][k];
                acc += c->val[s][k] * (double)recon[v];
                w += c->val[s][k];
                if (mask[v] == 1.0f) { siminside[idx] = 1; any = 1; }
            }
            if (w > 0) { sim[idx] = (float)(acc / w); simw[idx] = (float)w; }
        }
        if (slice_inside) slice_inside[s] = any;
    }
}

/* ParallelSuperresolution :3940-4022: addon / confidence map, per-thread volumes joined at the end (parallel_reduce) */
void cpu_superresolution(const cpu_coeffs *c, const float *slices, const float *weights, const float *sim, const float *slice_weights,
                         const float *scales, float *addon, float *cmap)
{
    const size_t V = (size_t)c->vx * c->vy * c->vz;
    const int P = c->Nx * c->Ny;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    if (nt > c->S) nt = c->S > 0 ? c->S : 1;
    double *A = (double *)calloc(V * nt, sizeof(double)), *C = (double *)calloc(V * nt, sizeof(double));
#pragma omp parallel for schedule(dynamic) num_threads(nt)
    for (int s = 0; s < c->S; ++s) {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        double *a = A + V * t, *cm = C + V * t;
        const long *st = c->start[s];
        for (int pix = 0; pix < P; ++pix) {
            const size_t idx = (size_t)s * P + pix;
            if (slices[idx] == -1.0f) continue;
            double v = (double)slices[idx] * scales[s];
            v = sim[idx] > 0 ? v - sim[idx] : 0;
            const double ww = (double)weights[idx] * slice_weights[s];
            for (long k = st[pix]; k < st[pix + 1]; ++k) {
                a[c->vox[s][k]] += c->val[s][k] * v * ww;
                cm[c->vox[s][k]] += c->val[s][k] * ww;
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long v = 0; v < (long)V; ++v) {
        double sa = 0, sc = 0;
        for (int t = 0; t < nt; ++t) { sa += A[V * t + v]; sc += C[V * t + v]; }
        addon[v] = (float)sa; cmap[v] = (float)sc;
    }
    free(A); free(C);
}
