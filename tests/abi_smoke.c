/*
 * abi_smoke.c — the C ABI of liblqrb200.so driven from a plain C99 caller (no Python, no torch): what a `ccall`
 * from Julia does, minus Julia.  Built and run by tests/test_gpu_abi.py; tests/test_host_cpu.py compiles it as the
 * "header is valid C" check.
 *
 *   abi_smoke <problem.bin> <result.bin>
 *
 * problem.bin (written by the test from the reference's cartpole fixture, test/problems.jl:58-88):
 *   int32 n, m, N, batch, hess_mode, sC, sc, pad;  int32 p[N];  then doubles
 *   Q[n,n,N,b] R[m,m,N-1,b] q[n,N,b] r[m,N-1,b] A[n,n,N-1,b] B[n,m,N-1,b] d[n,N-1,b] C[sC,b] c[sc,b]
 * result.bin: doubles dz[NN,b] mult[P,b] res[NN,b], then Z_lti[NN,b] of the LTI Riccati leg.
 * Checks return codes, info[], the LAPACK-style argument errors and, for the Riccati leg, the closed loop
 * x_{k+1} = A x_k + B u_k in C.  Prints ABI_SMOKE_OK on success.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lqrb200.h"

#define CHECK(cond, ...)                      \
    do {                                      \
        if (!(cond)) {                        \
            fprintf(stderr, "abi_smoke: ");   \
            fprintf(stderr, __VA_ARGS__);     \
            fprintf(stderr, "\n");            \
            return 1;                         \
        }                                     \
    } while (0)

static double *read_doubles(FILE *f, size_t count) {
    double *p = (double *)malloc((count ? count : 1) * sizeof(double));
    if (p && count && fread(p, sizeof(double), count, f) != count) {
        free(p);
        return NULL;
    }
    return p;
}

int main(int argc, char **argv) {
    CHECK(argc == 3, "usage: abi_smoke <problem.bin> <result.bin>");
    FILE *f = fopen(argv[1], "rb");
    CHECK(f, "cannot open %s", argv[1]);
    int32_t hd[8];
    CHECK(fread(hd, sizeof(int32_t), 8, f) == 8, "short header");
    const int32_t n = hd[0], m = hd[1], N = hd[2], hess = hd[4];
    const int64_t b = hd[3], sC = hd[5], sc = hd[6];
    int32_t *p = (int32_t *)malloc((size_t)N * sizeof(int32_t));
    CHECK(p && fread(p, sizeof(int32_t), (size_t)N, f) == (size_t)N, "short p");
    const size_t K1 = (size_t)(N - 1);
    double *Q = read_doubles(f, (size_t)n * n * N * b), *R = read_doubles(f, (size_t)m * m * K1 * b);
    double *q = read_doubles(f, (size_t)n * N * b), *r = read_doubles(f, (size_t)m * K1 * b);
    double *A = read_doubles(f, (size_t)n * n * K1 * b), *B = read_doubles(f, (size_t)n * m * K1 * b);
    double *d = read_doubles(f, (size_t)n * K1 * b), *C = read_doubles(f, (size_t)sC * b), *c = read_doubles(f, (size_t)sc * b);
    fclose(f);
    CHECK(Q && R && q && r && A && B && d && C && c, "short problem file");

    CHECK(lqrb_version() == LQRB_VERSION, "header / library version mismatch: %d vs %d", lqrb_version(), LQRB_VERSION);
    int32_t ndev = 0;
    CHECK(lqrb_device_count(&ndev) == 0 && ndev >= 1, "no CUDA device (the library has no CPU fallback)");
    lqrb_handle_t h = NULL;
    int32_t rc = lqrb_create(&h, 0);
    CHECK(rc == 0 && h, "lqrb_create -> %d", rc);

    const int64_t NN = lqrb_num_vars(n, m, N), P = lqrb_num_cons(n, N, p);
    CHECK(NN == (int64_t)N * n + (int64_t)(N - 1) * m, "lqrb_num_vars");
    double *dz = (double *)calloc((size_t)(NN * b), 8), *mult = (double *)calloc((size_t)(P * b), 8);
    double *res = (double *)calloc((size_t)(NN * b), 8);
    int32_t *info = (int32_t *)calloc((size_t)b, 4);

    /* _solve!(::CholeskySolver), src/cholesky_solver.jl:166-182 */
    rc = lqrb_kkt_solve_f64(h, n, m, N, b, p, hess, 0, Q, R, NULL, q, r, A, B, d, NULL, C, c, dz, mult, res, info);
    CHECK(rc == 0, "lqrb_kkt_solve_f64 -> %d: %s", rc, lqrb_last_error_string(h));
    for (int64_t i = 0; i < b; ++i) CHECK(info[i] == 0, "info[%lld] = %d", (long long)i, info[i]);
    printf("kkt kernel: %s, launches so far %lld\n", lqrb_last_kernel_name(h), (long long)lqrb_launch_count(h));

    /* the same through factor once / solve with the kept factor */
    double *dz2 = (double *)calloc((size_t)(NN * b), 8), *mult2 = (double *)calloc((size_t)(P * b), 8);
    rc = lqrb_kkt_factor_f64(h, n, m, N, b, p, hess, 0, Q, R, NULL, A, B, NULL, C, info);
    CHECK(rc == 0, "lqrb_kkt_factor_f64 -> %d: %s", rc, lqrb_last_error_string(h));
    rc = lqrb_kkt_solve_factored_f64(h, n, m, N, b, p, hess, 0, 0, q, r, d, c, dz2, mult2, NULL, info);
    CHECK(rc == 0, "lqrb_kkt_solve_factored_f64 -> %d: %s", rc, lqrb_last_error_string(h));
    double worst = 0.0, scale = 0.0;
    for (int64_t i = 0; i < NN * b; ++i) {
        worst = fmax(worst, fabs(dz[i] - dz2[i]));
        scale = fmax(scale, fabs(dz[i]));
    }
    CHECK(worst <= 1e-9 * fmax(1.0, scale), "factored solve differs from the fused solve by %g", worst);

    /* argument errors: LAPACK-style negative index, nothing thrown across the ABI */
    CHECK(lqrb_kkt_solve_f64(h, 0, m, N, b, p, hess, 0, Q, R, NULL, q, r, A, B, d, NULL, C, c, dz, mult, res, info) == -2, "n = 0 accepted");
    CHECK(lqrb_kkt_solve_f64(h, n, m, N, b, NULL, hess, 0, Q, R, NULL, q, r, A, B, d, NULL, C, c, dz, mult, res, info) == -6, "p = NULL accepted");
    CHECK(strlen(lqrb_last_error_string(h)) > 0, "empty error string");

    /* DPSolver solve! (src/dynamic_programming.jl:54-72) on the LTI problem made of knot 1 of instance 1 */
    double *x0 = (double *)calloc((size_t)n, 8), *Z = (double *)calloc((size_t)NN, 8);
    double *K = (double *)calloc((size_t)m * n * K1, 8), *kff = (double *)calloc((size_t)m * K1, 8);
    for (int i = 0; i < n; ++i) x0[i] = 0.1 * (i + 1);
    rc = lqrb_riccati_f64(h, n, m, N, 1, LQRB_FLAG_LTI | LQRB_FLAG_NO_AFFINE, A, B, Q, R, NULL, NULL, Q + (size_t)n * n * (N - 1),
                          NULL, x0, Z, K, kff, info);
    CHECK(rc == 0 && info[0] == 0, "lqrb_riccati_f64 -> %d, info %d: %s", rc, info[0], lqrb_last_error_string(h));
    printf("riccati kernel: %s\n", lqrb_last_kernel_name(h));
    double dyn = 0.0;
    for (int k = 0; k < N - 1; ++k) {
        const double *x = Z + (size_t)k * (n + m), *u = x + n, *xn = Z + (size_t)(k + 1) * (n + m);
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j < n; ++j) s += A[i + j * n] * x[j];
            for (int j = 0; j < m; ++j) s += B[i + j * n] * u[j];
            dyn = fmax(dyn, fabs(s - xn[i]));
        }
        for (int j = 0; j < m; ++j) { /* u_k = -K_k x_k (no affine terms) */
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += K[(size_t)k * m * n + j + i * m] * x[i];
            dyn = fmax(dyn, fabs(u[j] + s));
        }
    }
    for (int i = 0; i < n; ++i) dyn = fmax(dyn, fabs(Z[i] - x0[i]));
    CHECK(dyn <= 1e-12, "closed loop / rollout residual %g", dyn);

    /* rollout! (src/least_squares.jl:195-202) with the controls just computed reproduces the states */
    double *U = (double *)calloc((size_t)m * K1, 8), *X = (double *)calloc((size_t)n * N, 8);
    for (int k = 0; k < N - 1; ++k)
        for (int j = 0; j < m; ++j) U[(size_t)k * m + j] = Z[(size_t)k * (n + m) + n + j];
    rc = lqrb_rollout_f64(h, n, m, N, 1, LQRB_FLAG_LTI, A, B, x0, U, X);
    CHECK(rc == 0, "lqrb_rollout_f64 -> %d", rc);
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < n; ++i) CHECK(fabs(X[(size_t)k * n + i] - Z[(size_t)k * (n + m) + i]) <= 1e-12, "rollout differs at knot %d", k);

    FILE *o = fopen(argv[2], "wb");
    CHECK(o, "cannot open %s", argv[2]);
    fwrite(dz, 8, (size_t)(NN * b), o);
    fwrite(mult, 8, (size_t)(P * b), o);
    fwrite(res, 8, (size_t)(NN * b), o);
    fwrite(Z, 8, (size_t)NN, o);
    fclose(o);
    CHECK(lqrb_synchronize(h) == 0, "lqrb_synchronize");
    CHECK(lqrb_destroy(h) == 0, "lqrb_destroy");
    printf("ABI_SMOKE_OK\n");
    return 0;
}
