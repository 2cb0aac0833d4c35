// gsl_shim.h -- TEST INFRASTRUCTURE.  A minimal stand-in for the handful of GSL entry points the reference's vendored IRTK calls
// (geometry++/src/irtkMatrix.cc, irtkVector.cc, packages/registration/src/irtkOptimizer.cc, ...): dense vectors / matrices, LU
// decomposition with partial pivoting (solve, invert, determinant), SVD (one-sided Jacobi) and the symmetric eigenproblem (Jacobi).
// GSL itself is not in this image (un-vendored dependency, version unpinned); this header contains no GSL
// code, only textbook algorithms behind GSL's function names, so that the reference's own IRTK sources compile unmodified into
// oracle/_ref/libref_irtk.so (oracle/Makefile: make ref_irtk).  Results agree with GSL to rounding, not bit for bit.
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

extern "C++" {
struct gsl_vector { size_t size; size_t stride; double* data; };
struct gsl_matrix { size_t size1, size2, tda; double* data; };
struct gsl_permutation { size_t size; size_t* data; };

inline gsl_vector* gsl_vector_alloc(size_t n) { gsl_vector* v = new gsl_vector; v->size = n; v->stride = 1; v->data = (double*)calloc(n ? n : 1, sizeof(double)); return v; }
inline gsl_vector* gsl_vector_calloc(size_t n) { return gsl_vector_alloc(n); }
inline void gsl_vector_free(gsl_vector* v) { if (v) { free(v->data); delete v; } }
inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i * v->stride]; }
inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i * v->stride] = x; }
inline void gsl_vector_set_zero(gsl_vector* v) { for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] = 0; }
inline gsl_matrix* gsl_matrix_alloc(size_t r, size_t c) { gsl_matrix* m = new gsl_matrix; m->size1 = r; m->size2 = c; m->tda = c; m->data = (double*)calloc(r * c ? r * c : 1, sizeof(double)); return m; }
inline gsl_matrix* gsl_matrix_calloc(size_t r, size_t c) { return gsl_matrix_alloc(r, c); }
inline void gsl_matrix_free(gsl_matrix* m) { if (m) { free(m->data); delete m; } }
inline double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j) { return m->data[i * m->tda + j]; }
inline void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x) { m->data[i * m->tda + j] = x; }
inline void gsl_matrix_set_zero(gsl_matrix* m) { memset(m->data, 0, sizeof(double) * m->size1 * m->tda); }
inline void gsl_matrix_set_identity(gsl_matrix* m) { gsl_matrix_set_zero(m); for (size_t i = 0; i < std::min(m->size1, m->size2); ++i) gsl_matrix_set(m, i, i, 1.0); }
inline int gsl_matrix_memcpy(gsl_matrix* d, const gsl_matrix* s) { for (size_t i = 0; i < s->size1; ++i) for (size_t j = 0; j < s->size2; ++j) gsl_matrix_set(d, i, j, gsl_matrix_get(s, i, j)); return 0; }
inline gsl_permutation* gsl_permutation_alloc(size_t n) { gsl_permutation* p = new gsl_permutation; p->size = n; p->data = (size_t*)calloc(n ? n : 1, sizeof(size_t)); for (size_t i = 0; i < n; ++i) p->data[i] = i; return p; }
inline void gsl_permutation_free(gsl_permutation* p) { if (p) { free(p->data); delete p; } }

// A = P^T L U in place (L unit lower, U upper), partial pivoting; signum = sign of the permutation
inline int gsl_linalg_LU_decomp(gsl_matrix* A, gsl_permutation* p, int* signum)
{
    const size_t n = A->size1;
    *signum = 1;
    for (size_t i = 0; i < n; ++i) p->data[i] = i;
    for (size_t j = 0; j + 1 < n || j < n; ++j) {
        double big = std::fabs(gsl_matrix_get(A, j, j)); size_t piv = j;
        for (size_t i = j + 1; i < n; ++i) { const double a = std::fabs(gsl_matrix_get(A, i, j)); if (a > big) { big = a; piv = i; } }
        if (piv != j) {
            for (size_t k = 0; k < n; ++k) std::swap(A->data[j * A->tda + k], A->data[piv * A->tda + k]);
            std::swap(p->data[j], p->data[piv]);
            *signum = -*signum;
        }
        const double ajj = gsl_matrix_get(A, j, j);
        if (ajj != 0.0)
            for (size_t i = j + 1; i < n; ++i) {
                const double f = gsl_matrix_get(A, i, j) / ajj;
                gsl_matrix_set(A, i, j, f);
                for (size_t k = j + 1; k < n; ++k) gsl_matrix_set(A, i, k, gsl_matrix_get(A, i, k) - f * gsl_matrix_get(A, j, k));
            }
        if (j + 1 >= n) break;
    }
    return 0;
}
inline int gsl_linalg_LU_solve(const gsl_matrix* LU, const gsl_permutation* p, const gsl_vector* b, gsl_vector* x)
{
    const size_t n = LU->size1;
    for (size_t i = 0; i < n; ++i) gsl_vector_set(x, i, gsl_vector_get(b, p->data[i]));
    for (size_t i = 0; i < n; ++i) { double s = gsl_vector_get(x, i); for (size_t k = 0; k < i; ++k) s -= gsl_matrix_get(LU, i, k) * gsl_vector_get(x, k); gsl_vector_set(x, i, s); }
    for (size_t ii = n; ii-- > 0;) { double s = gsl_vector_get(x, ii); for (size_t k = ii + 1; k < n; ++k) s -= gsl_matrix_get(LU, ii, k) * gsl_vector_get(x, k); gsl_vector_set(x, ii, s / gsl_matrix_get(LU, ii, ii)); }
    return 0;
}
inline int gsl_linalg_LU_invert(const gsl_matrix* LU, const gsl_permutation* p, gsl_matrix* inv)
{
    const size_t n = LU->size1;
    gsl_vector* b = gsl_vector_alloc(n); gsl_vector* x = gsl_vector_alloc(n);
    for (size_t c = 0; c < n; ++c) {
        gsl_vector_set_zero(b); gsl_vector_set(b, c, 1.0);
        gsl_linalg_LU_solve(LU, p, b, x);
        for (size_t r = 0; r < n; ++r) gsl_matrix_set(inv, r, c, gsl_vector_get(x, r));
    }
    gsl_vector_free(b); gsl_vector_free(x);
    return 0;
}
inline double gsl_linalg_LU_det(gsl_matrix* LU, int signum)
{
    double d = signum;
    for (size_t i = 0; i < LU->size1; ++i) d *= gsl_matrix_get(LU, i, i);
    return d;
}

// Symmetric eigenproblem, cyclic Jacobi: eigenvalues -> eval, eigenvectors -> columns of evec
struct gsl_eigen_symmv_workspace { size_t n; };
inline gsl_eigen_symmv_workspace* gsl_eigen_symmv_alloc(size_t n) { gsl_eigen_symmv_workspace* w = new gsl_eigen_symmv_workspace; w->n = n; return w; }
inline void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace* w) { delete w; }
inline int gsl_eigen_symmv(gsl_matrix* A, gsl_vector* eval, gsl_matrix* evec, gsl_eigen_symmv_workspace*)
{
    const size_t n = A->size1;
    gsl_matrix_set_identity(evec);
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0;
        for (size_t i = 0; i < n; ++i) for (size_t j = i + 1; j < n; ++j) off += gsl_matrix_get(A, i, j) * gsl_matrix_get(A, i, j);
        if (off < 1e-300) break;
        for (size_t p = 0; p < n; ++p) for (size_t q = p + 1; q < n; ++q) {
            const double apq = gsl_matrix_get(A, p, q);
            if (apq == 0.0) continue;
            const double theta = (gsl_matrix_get(A, q, q) - gsl_matrix_get(A, p, p)) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
            for (size_t k = 0; k < n; ++k) { const double akp = gsl_matrix_get(A, k, p), akq = gsl_matrix_get(A, k, q); gsl_matrix_set(A, k, p, c * akp - s * akq); gsl_matrix_set(A, k, q, s * akp + c * akq); }
            for (size_t k = 0; k < n; ++k) { const double apk = gsl_matrix_get(A, p, k), aqk = gsl_matrix_get(A, q, k); gsl_matrix_set(A, p, k, c * apk - s * aqk); gsl_matrix_set(A, q, k, s * apk + c * aqk); }
            for (size_t k = 0; k < n; ++k) { const double vkp = gsl_matrix_get(evec, k, p), vkq = gsl_matrix_get(evec, k, q); gsl_matrix_set(evec, k, p, c * vkp - s * vkq); gsl_matrix_set(evec, k, q, s * vkp + c * vkq); }
        }
    }
    for (size_t i = 0; i < n; ++i) gsl_vector_set(eval, i, gsl_matrix_get(A, i, i));
    return 0;
}

// SVD A (m x n, m >= n) = U S V^T: A is replaced by U, singular values -> S (descending), V -> V.  One-sided Jacobi.
inline int gsl_linalg_SV_decomp(gsl_matrix* A, gsl_matrix* V, gsl_vector* S, gsl_vector*)
{
    const size_t m = A->size1, n = A->size2;
    gsl_matrix_set_identity(V);
    for (int sweep = 0; sweep < 100; ++sweep) {
        bool rotated = false;
        for (size_t p = 0; p < n; ++p) for (size_t q = p + 1; q < n; ++q) {
            double a = 0, b = 0, c = 0;
            for (size_t k = 0; k < m; ++k) { const double x = gsl_matrix_get(A, k, p), y = gsl_matrix_get(A, k, q); a += x * x; b += y * y; c += x * y; }
            if (std::fabs(c) <= 1e-15 * std::sqrt(a * b) || c == 0.0) continue;
            rotated = true;
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
            for (size_t k = 0; k < m; ++k) { const double x = gsl_matrix_get(A, k, p), y = gsl_matrix_get(A, k, q); gsl_matrix_set(A, k, p, cs * x - sn * y); gsl_matrix_set(A, k, q, sn * x + cs * y); }
            for (size_t k = 0; k < n; ++k) { const double x = gsl_matrix_get(V, k, p), y = gsl_matrix_get(V, k, q); gsl_matrix_set(V, k, p, cs * x - sn * y); gsl_matrix_set(V, k, q, sn * x + cs * y); }
        }
        if (!rotated) break;
    }
    std::vector<double> sv(n);
    for (size_t j = 0; j < n; ++j) { double s = 0; for (size_t k = 0; k < m; ++k) s += gsl_matrix_get(A, k, j) * gsl_matrix_get(A, k, j); sv[j] = std::sqrt(s); }
    std::vector<size_t> order(n);
    for (size_t j = 0; j < n; ++j) order[j] = j;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return sv[x] > sv[y]; });
    std::vector<double> Ac(m * n), Vc(n * n);
    for (size_t j = 0; j < n; ++j) {
        const size_t o = order[j];
        for (size_t k = 0; k < m; ++k) Ac[k * n + j] = sv[o] > 0 ? gsl_matrix_get(A, k, o) / sv[o] : 0.0;
        for (size_t k = 0; k < n; ++k) Vc[k * n + j] = gsl_matrix_get(V, k, o);
        gsl_vector_set(S, j, sv[o]);
    }
    for (size_t k = 0; k < m; ++k) for (size_t j = 0; j < n; ++j) gsl_matrix_set(A, k, j, Ac[k * n + j]);
    for (size_t k = 0; k < n; ++k) for (size_t j = 0; j < n; ++j) gsl_matrix_set(V, k, j, Vc[k * n + j]);
    return 0;
}
inline int gsl_linalg_SV_solve(const gsl_matrix* U, const gsl_matrix* V, const gsl_vector* S, const gsl_vector* b, gsl_vector* x)
{
    const size_t m = U->size1, n = U->size2;
    std::vector<double> w(n);
    for (size_t j = 0; j < n; ++j) { double s = 0; for (size_t k = 0; k < m; ++k) s += gsl_matrix_get(U, k, j) * gsl_vector_get(b, k); const double sj = gsl_vector_get(S, j); w[j] = sj != 0 ? s / sj : 0.0; }
    for (size_t i = 0; i < n; ++i) { double s = 0; for (size_t j = 0; j < n; ++j) s += gsl_matrix_get(V, i, j) * w[j]; gsl_vector_set(x, i, s); }
    return 0;
}
}
