// smallmat.cuh — register-resident small dense kernels for the thread-per-instance path.
// Every loop has compile-time bounds and is fully unrolled, so all arrays live in registers.
// These are the in-kernel replacements of the LAPACK/BLAS calls the reference makes on tiny blocks
// (potrf!/potrs!/trsv!/trsm!/mul!).
#pragma once
#include "common.cuh"

#define SM_UNROLL _Pragma("unroll")

// Upper Cholesky of a packed symmetric k x k matrix, in place (A = U'U; LAPACK.potrf!('U')).
// dinv[j] = 1/U(j,j) is kept so that the triangular solves multiply instead of divide.
// Returns 0 or the 1-based index of the first non-positive pivot (potrf info semantics).
template <int k>
__device__ __forceinline__ int chol_packed(double *a, double *dinv) {
    int info = 0;
    SM_UNROLL
    for (int j = 0; j < k; ++j) {
        double s = a[tri_idx(j, j)];
        SM_UNROLL
        for (int l = 0; l < j; ++l) s = fma(-a[tri_idx(l, j)], a[tri_idx(l, j)], s);
        if (!(s > 0.0) && info == 0) info = j + 1;
        const double dj = sqrt(s);
        const double ij = 1.0 / dj;
        a[tri_idx(j, j)] = dj;
        dinv[j] = ij;
        SM_UNROLL
        for (int i = j + 1; i < k; ++i) {
            double t = a[tri_idx(j, i)];
            SM_UNROLL
            for (int l = 0; l < j; ++l) t = fma(-a[tri_idx(l, j)], a[tri_idx(l, i)], t);
            a[tri_idx(j, i)] = t * ij;
        }
    }
    return info;
}

// x <- U^-T x  (BLAS.trsv!('U','T')), U packed upper, stride-s vector
template <int k, int s = 1>
__device__ __forceinline__ void solve_ut(const double *u, const double *dinv, double *x) {
    SM_UNROLL
    for (int i = 0; i < k; ++i) {
        double t = x[i * s];
        SM_UNROLL
        for (int l = 0; l < i; ++l) t = fma(-u[tri_idx(l, i)], x[l * s], t);
        x[i * s] = t * dinv[i];
    }
}

// x <- U^-1 x  (BLAS.trsv!('U','N'))
template <int k, int s = 1>
__device__ __forceinline__ void solve_un(const double *u, const double *dinv, double *x) {
    SM_UNROLL
    for (int i = k - 1; i >= 0; --i) {
        double t = x[i * s];
        SM_UNROLL
        for (int l = i + 1; l < k; ++l) t = fma(-u[tri_idx(i, l)], x[l * s], t);
        x[i * s] = t * dinv[i];
    }
}

// x <- (U'U)^-1 x  (LAPACK.potrs!)
template <int k, int s = 1>
__device__ __forceinline__ void solve_chol(const double *u, const double *dinv, double *x) {
    solve_ut<k, s>(u, dinv, x);
    solve_un<k, s>(u, dinv, x);
}

// streaming (read-once) and default loads of one row of a packed tile
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ double ld_keep(const double *p) { return __ldg(p); }
