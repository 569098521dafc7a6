// coop_prims.cuh — cooperative dense primitives: G threads of a group (a warp, or a whole CTA) work on one small
// matrix in shared or global memory.  Stand-ins for the BLAS / LAPACK calls the reference makes (gemm, potrf, trsm).
// Used by the general KKT kernel (kkt_coop.cu) and the condensed least-squares path (lsq.cu).
#pragma once
#include "riccati_kernels.cuh"  // group_sync

// ------------------------------------------------------------------ cooperative primitives ----
// All matrices column-major with explicit leading dimension; every primitive ends with a group sync.
template <int G>
__device__ __forceinline__ void co_gemm(int ta, int tb, int M, int Nn, int K, double alpha,
                                        const double *A, int lda, const double *B, int ldb, double beta,
                                        double *C, int ldc, int t) {
    for (int e = t; e < M * Nn; e += G) {
        const int i = e % M, j = e / M;
        double s = 0.0;
        for (int l = 0; l < K; ++l) {
            const double a = ta ? A[l + i * lda] : A[i + l * lda];
            const double b = tb ? B[j + l * ldb] : B[l + j * ldb];
            s = fma(a, b, s);
        }
        C[i + j * ldc] = (beta == 0.0) ? alpha * s : fma(alpha, s, beta * C[i + j * ldc]);
    }
    group_sync<G>();
}

// in-place upper Cholesky (right-looking); strict lower untouched.  info: 1-based bad pivot or 0.
template <int G>
__device__ __forceinline__ int co_chol(double *A, int k, int lda, int t) {
    int info = 0;
    for (int j = 0; j < k; ++j) {
        const double djj = A[j + j * lda];
        if (!(djj > 0.0) && info == 0) info = j + 1;
        const double d = sqrt(djj);
        group_sync<G>();
        for (int i = j + t; i < k; i += G) A[j + i * lda] = (i == j) ? d : A[j + i * lda] / d;
        group_sync<G>();
        const int rem = k - j - 1;
        for (int e = t; e < rem * rem; e += G) {
            const int a = j + 1 + e % rem, b = j + 1 + e / rem;
            if (a <= b) A[a + b * lda] = fma(-A[j + a * lda], A[j + b * lda], A[a + b * lda]);
        }
        group_sync<G>();
    }
    return info;
}

// B <- U^-T B  (one thread per right-hand-side column)
template <int G>
__device__ __forceinline__ void co_trsm_ut(const double *U, int k, int ldu, double *B, int nrhs, int ldb,
                                           int t) {
    for (int c = t; c < nrhs; c += G) {
        double *x = B + (size_t)c * ldb;
        for (int i = 0; i < k; ++i) {
            double s = x[i];
            for (int l = 0; l < i; ++l) s = fma(-U[l + i * ldu], x[l], s);
            x[i] = s / U[i + i * ldu];
        }
    }
    group_sync<G>();
}

// B <- U^-1 B
template <int G>
__device__ __forceinline__ void co_trsm_un(const double *U, int k, int ldu, double *B, int nrhs, int ldb,
                                           int t) {
    for (int c = t; c < nrhs; c += G) {
        double *x = B + (size_t)c * ldb;
        for (int i = k - 1; i >= 0; --i) {
            double s = x[i];
            for (int l = i + 1; l < k; ++l) s = fma(-U[i + l * ldu], x[l], s);
            x[i] = s / U[i + i * ldu];
        }
    }
    group_sync<G>();
}

