/*
 * lqr_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the LQR.jl hot path, one function per reference
 * routine, citing the reference file:line each follows.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (liblqrb200.so) never links or calls it.
 *
 * PARITY PIN STATUS: the reference (bjack205/LQR.jl) stores no golden vectors
 * and Julia is not installed, so this restatement cannot be checked against
 * reference *outputs* ("parity unpinned" in that sense).  It is pinned instead
 * by the reference's own test identities (test/cholesky_solve.jl:18-44,
 * test/constraint_blocks.jl:70-133, test/block_cholesky.jl:24-67): see
 * tests/test_oracle.py, which evaluates them against an independent dense
 * KKT solve with extended-precision refinement (oracle/dense_kkt.py).
 *
 * Data layout here is the reference's own: column-major small matrices,
 * instance-major (one problem instance contiguous), 0-based knot index k.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LQRO_HESS_DENSE 0
#define LQRO_HESS_BLOCKDIAG 1
#define LQRO_HESS_DIAG 2
#define LQRO_FLAG_SOC 1 /* Ginv=false: H=I, g=0 (src/cholesky_solver.jl:254-273) */
#define LQRO_FLAG_LTI 2 /* Riccati: A,B,Q,R,q,r have no knot axis (src/lqr_problem.jl:1-11) */

/* ------------------------------------------------------------------------ */
/* small dense helpers (column-major).  Stand-ins for the LAPACK/BLAS calls  */
/* the reference makes.                                        */
/* ------------------------------------------------------------------------ */

/* LAPACK.potrf!('U', A): A = U'U, U in the upper triangle; strict lower is
 * left untouched.  Returns 0 or the 1-based index of the first bad pivot
 * (src/cholesky_solve.jl:1-3 returns this info; callers ignore it). */
static int potrf_u(int n, double *a, int lda) {
    for (int j = 0; j < n; ++j) {
        double s = a[j + j * lda];
        for (int k = 0; k < j; ++k) s -= a[k + j * lda] * a[k + j * lda];
        if (!(s > 0.0)) return j + 1;
        s = sqrt(s);
        a[j + j * lda] = s;
        for (int i = j + 1; i < n; ++i) {
            double t = a[j + i * lda];
            for (int k = 0; k < j; ++k) t -= a[k + j * lda] * a[k + i * lda];
            a[j + i * lda] = t / s;
        }
    }
    return 0;
}

/* BLAS.trsm!('L','U','T','N'): B <- U^-T B  (n x nrhs)  (src/cholesky_solve.jl:42-45) */
static void trsm_ut(int n, int nrhs, const double *u, int ldu, double *b, int ldb) {
    for (int c = 0; c < nrhs; ++c)
        for (int i = 0; i < n; ++i) {
            double t = b[i + c * ldb];
            for (int k = 0; k < i; ++k) t -= u[k + i * ldu] * b[k + c * ldb];
            b[i + c * ldb] = t / u[i + i * ldu];
        }
}

/* BLAS.trsm!('L','U','N','N'): B <- U^-1 B */
static void trsm_un(int n, int nrhs, const double *u, int ldu, double *b, int ldb) {
    for (int c = 0; c < nrhs; ++c)
        for (int i = n - 1; i >= 0; --i) {
            double t = b[i + c * ldb];
            for (int k = i + 1; k < n; ++k) t -= u[i + k * ldu] * b[k + c * ldb];
            b[i + c * ldb] = t / u[i + i * ldu];
        }
}

/* LAPACK.potrs!('U'): B <- (U'U)^-1 B */
static void potrs_u(int n, int nrhs, const double *u, int ldu, double *b, int ldb) {
    trsm_ut(n, nrhs, u, ldu, b, ldb);
    trsm_un(n, nrhs, u, ldu, b, ldb);
}

/* C(mxn) = alpha*op(A)*op(B) + beta*C ; ta/tb: 0 = N, 1 = T */
static void gemm(int ta, int tb, int m, int n, int k, double alpha, const double *a, int lda,
                 const double *b, int ldb, double beta, double *c, int ldc) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) {
                double av = ta ? a[l + i * lda] : a[i + l * lda];
                double bv = tb ? b[j + l * ldb] : b[l + j * ldb];
                s += av * bv;
            }
            c[i + j * ldc] = alpha * s + (beta == 0.0 ? 0.0 : beta * c[i + j * ldc]);
        }
}

/* ------------------------------------------------------------------------ */
/* Riccati: src/dynamic_programming.jl:28-72 (LTI, no affine terms there),   */
/* generalised to per-knot A_k,B_k,Q_k,R_k and affine q_k,r_k,qf as SURVEY   */
/* Appendix A states.  K is m x n per knot, kff is m per knot.               */
/* ------------------------------------------------------------------------ */
int lqro_riccati(int n, int m, int N, int flags, const double *A, const double *B, const double *Q,
                 const double *R, const double *q, const double *r, const double *Qf,
                 const double *qf, const double *x0, double *X, double *U, double *K,
                 double *kff, double *work /* >= 4n^2+3nm+m^2+2m+2n */) {
    const int lti = (flags & LQRO_FLAG_LTI) != 0;
    double *P = work, *P_ = P + n * n, *PA = P_ + n * n, *PB = PA + n * n, *APB = PB + n * m,
           *E = APB + n * m, *p = E + m * m, *p_ = p + n, *rr = p_ + n, *tmp = rr + m;
    int info = 0;
    /* Terminal ctg: solver.P .= prob.Qf  (:58) */
    memcpy(P, Qf, sizeof(double) * n * n);
    for (int i = 0; i < n; ++i) p[i] = qf ? qf[i] : 0.0;
    for (int k = N - 2; k >= 0; --k) { /* :61-64 */
        const int kk = lti ? 0 : k;
        const double *Ak = A + (size_t)kk * n * n, *Bk = B + (size_t)kk * n * m;
        const double *Qk = Q + (size_t)kk * n * n, *Rk = R + (size_t)kk * m * m;
        const double *qk = q ? q + (size_t)kk * n : NULL, *rk = r ? r + (size_t)kk * m : NULL;
        double *Kk = K + (size_t)k * m * n, *kk_ff = kff + (size_t)k * m;
        /* compute_gain! :37-43 */
        gemm(0, 0, n, m, n, 1.0, P, n, Bk, n, 0.0, PB, n);   /* PB = P*B       */
        memcpy(E, Rk, sizeof(double) * m * m);
        gemm(1, 0, m, m, n, 1.0, Bk, n, PB, n, 1.0, E, m);   /* E = R + B'PB   */
        gemm(0, 0, n, n, n, 1.0, P, n, Ak, n, 0.0, PA, n);   /* PA = P*A       */
        gemm(1, 0, m, n, n, 1.0, Bk, n, PA, n, 0.0, Kk, m);  /* K = B'PA       */
        for (int i = 0; i < m; ++i) {                        /* rr = r + B'p   */
            double s = rk ? rk[i] : 0.0;
            for (int l = 0; l < n; ++l) s += Bk[l + i * n] * p[l];
            rr[i] = s;
            kk_ff[i] = s;
        }
        int st = potrf_u(m, E, m); /* chol_solve! :28-31 */
        if (st && !info) info = (k + 1) * 1000 + st;
        potrs_u(m, n, E, m, Kk, m);
        potrs_u(m, 1, E, m, kk_ff, m);
        /* compute_ctg! :48-52 */
        gemm(1, 0, n, m, n, 1.0, Ak, n, PB, n, 0.0, APB, n); /* APB = A'PB     */
        memcpy(P_, Qk, sizeof(double) * n * n);
        gemm(1, 0, n, n, n, 1.0, Ak, n, PA, n, 1.0, P_, n);  /* Q + A'PA       */
        gemm(0, 0, n, n, m, -1.0, APB, n, Kk, m, 1.0, P_, n); /* - APB*K       */
        for (int i = 0; i < n; ++i) {                        /* p_ = q + A'p - K'rr */
            double s = qk ? qk[i] : 0.0;
            for (int l = 0; l < n; ++l) s += Ak[l + i * n] * p[l];
            for (int l = 0; l < m; ++l) s -= Kk[l + i * m] * rr[l];
            p_[i] = s;
        }
        memcpy(P, P_, sizeof(double) * n * n); /* solver.P .= solver.P_ (:63) */
        memcpy(p, p_, sizeof(double) * n);
    }
    /* forward rollout :66-70 (same recurrence as src/least_squares.jl:195-202) */
    memcpy(X, x0, sizeof(double) * n);
    for (int k = 0; k < N - 1; ++k) {
        const int kk = lti ? 0 : k;
        const double *Ak = A + (size_t)kk * n * n, *Bk = B + (size_t)kk * n * m;
        const double *Kk = K + (size_t)k * m * n, *kf = kff + (size_t)k * m;
        const double *xk = X + (size_t)k * n;
        double *uk = U + (size_t)k * m, *xn = X + (size_t)(k + 1) * n;
        for (int i = 0; i < m; ++i) {
            double s = -kf[i];
            for (int l = 0; l < n; ++l) s -= Kk[i + l * m] * xk[l];
            uk[i] = s;
        }
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int l = 0; l < n; ++l) s += Ak[i + l * n] * xk[l];
            for (int l = 0; l < m; ++l) s += Bk[i + l * n] * uk[l];
            tmp[i] = s;
        }
        memcpy(xn, tmp, sizeof(double) * n);
    }
    return info;
}

size_t lqro_riccati_work_doubles(int n, int m) {
    return (size_t)4 * n * n + 3 * n * m + m * m + 2 * m + 3 * n;
}

/* ------------------------------------------------------------------------ */
/* BlockCholesky: src/block_cholesky.jl:55-101.  M is (n+m)^2 column-major.  */
/* mode DENSE: potrf on the whole [A C';C B]; BLOCKDIAG: potrf on A and B    */
/* separately with zero coupling; DIAG: store reciprocals (:82-91).          */
/* ------------------------------------------------------------------------ */
int lqro_block_cholesky(int n, int m, int mode, const double *A, const double *B,
                        const double *C /* m x n or NULL */, double *M) {
    const int w = n + m;
    memset(M, 0, sizeof(double) * w * w);
    if (mode == LQRO_HESS_DIAG) { /* :82-91 */
        for (int i = 0; i < n; ++i) M[i + i * w] = 1.0 / A[i + i * n];
        for (int i = 0; i < m; ++i) M[(n + i) + (n + i) * w] = 1.0 / B[i + i * m];
        return 0;
    }
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) M[i + j * w] = A[i + j * n];
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) M[(n + i) + (n + j) * w] = B[i + j * m];
    if (mode == LQRO_HESS_BLOCKDIAG) { /* :69-77 */
        int st = potrf_u(n, M, w);
        if (st) return st;
        st = potrf_u(m, M + n + n * w, w);
        return st ? n + st : 0;
    }
    if (C) /* :55-66: chol.C .= C; transpose!(chol.Ct, C) */
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < m; ++i) {
                M[(n + i) + j * w] = C[i + j * m];
                M[j + (n + i) * w] = C[i + j * m];
            }
    return potrf_u(w, M, w);
}

/* ldiv!(chol, b): src/block_cholesky.jl:93-96 */
void lqro_block_ldiv(int n, int m, int mode, const double *M, int nrhs, double *b, int ldb) {
    const int w = n + m;
    if (mode == LQRO_HESS_DIAG) {
        for (int c = 0; c < nrhs; ++c)
            for (int i = 0; i < w; ++i) b[i + c * ldb] *= M[i + i * w];
        return;
    }
    /* block-diag storage has exact zeros in the coupling, so the whole-matrix
     * triangular solves are what the reference does too (:50, :93). */
    potrs_u(w, nrhs, M, w, b, ldb);
}

/* ------------------------------------------------------------------------ */
/* Constrained KKT solve: src/cholesky_solver.jl:166-236 and callees.        */
/* ------------------------------------------------------------------------ */
typedef struct {
    int n, m, N, hess_mode, flags;
    const int32_t *p;   /* stage-constraint rows per knot [N] (src/conblocks.jl:74-96) */
    const double *Q;    /* n*n*N    */
    const double *R;    /* m*m*(N-1) */
    const double *Hux;  /* m*n*(N-1) or NULL (dense mode only) */
    const double *q;    /* n*N */
    const double *r;    /* m*(N-1) */
    const double *A;    /* n*n*(N-1) : D1_k = [A_k B_k] */
    const double *B;    /* n*m*(N-1) */
    const double *d;    /* n*(N-1)  dynamics constraint value */
    const double *D2;   /* NULL => [-I 0] ; else sum_k n*w_{k+1} */
    const double *C;    /* concatenated C_k (p_k x w_k col-major) */
    const double *c;    /* concatenated c_k */
} lqro_kkt_problem;

static int knot_w(const lqro_kkt_problem *pr, int k) { return pr->n + (k < pr->N - 1 ? pr->m : 0); }

size_t lqro_kkt_num_vars(int n, int m, int N) { return (size_t)N * n + (size_t)(N - 1) * m; }
size_t lqro_kkt_num_cons(int n, int N, const int32_t *p) {
    size_t P = (size_t)(N - 1) * n;
    for (int k = 0; k < N; ++k) P += p[k];
    return P;
}

/* Per-knot storage of one block row of S (and of its factor), the analogue of
 * BlockUpperTriangular3 (src/jacobian_blocks.jl:97-127).  A_k aliases C_{k-1}
 * (:165-167): we keep one C array per knot and read it as the next knot's A. */
typedef struct {
    int p1, ps, p2;
    double *B, *C, *D, *E, *F, *c, *d, *mu, *lam;
} tri3;

/* Solve the whole instance.  Optional dense outputs for the test identities:
 *   S_out (P x P, upper filled), h_out (P), U_out (P x P upper) — the analogues
 *   of get_shur_factors / get_cholesky (src/cholesky_solver.jl:333-363).       */
int lqro_kkt_solve(const lqro_kkt_problem *pr, double *dz, double *mult, double *res_out,
                   double *S_out, double *h_out, double *U_out) {
    const int n = pr->n, m = pr->m, N = pr->N, wmax = n + m;
    const int soc = (pr->flags & LQRO_FLAG_SOC) != 0;
    int pmax = 0;
    for (int k = 0; k < N; ++k)
        if (pr->p[k] > pmax) pmax = pr->p[k];
    const int rmax = 2 * n + pmax;
    const size_t P = lqro_kkt_num_cons(n, N, pr->p);
    int info = 0;

    /* per-knot blocks */
    tri3 *T = (tri3 *)calloc(N, sizeof(tri3));
    double **Yk = (double **)calloc(N, sizeof(double *));   /* Y_k  (rho x w)      */
    double **Mk = (double **)calloc(N, sizeof(double *));   /* chol(H_k) storage   */
    double **rk = (double **)calloc(N, sizeof(double *));   /* r = Y H^-1 g        */
    double *JYt = (double *)malloc(sizeof(double) * wmax * rmax);
    double *YYt = (double *)malloc(sizeof(double) * rmax * rmax);
    double *g = (double *)malloc(sizeof(double) * wmax);
    double *tmpv = (double *)malloc(sizeof(double) * (rmax + wmax));

    size_t coff = 0, Coff = 0, D2off = 0;
    for (int k = 0; k < N; ++k) {
        const int w = knot_w(pr, k), p1 = k > 0 ? n : 0, ps = pr->p[k], p2 = k < N - 1 ? n : 0;
        const int rho = p1 + ps + p2, mk = w - n;
        tri3 *t = &T[k];
        t->p1 = p1; t->ps = ps; t->p2 = p2;
        t->B = (double *)calloc((size_t)ps * ps + 1, sizeof(double));
        t->C = (double *)calloc((size_t)p2 * p2 + 1, sizeof(double));
        t->D = (double *)calloc((size_t)p1 * ps + 1, sizeof(double));
        t->E = (double *)calloc((size_t)ps * p2 + 1, sizeof(double));
        t->F = (double *)calloc((size_t)p1 * p2 + 1, sizeof(double));
        t->c = (double *)calloc(ps + 1, sizeof(double));
        t->d = (double *)calloc(p2 + 1, sizeof(double));
        t->mu = (double *)calloc(ps + 1, sizeof(double));
        t->lam = (double *)calloc(p2 + 1, sizeof(double));

        /* ---- ConstraintBlock Y=[D2;C;D1], y=[c;d]  (src/conblocks.jl:36-72) ---- */
        double *Y = Yk[k] = (double *)calloc((size_t)rho * w + 1, sizeof(double));
        if (p1) {
            if (pr->D2) {
                for (int j = 0; j < w; ++j)
                    for (int i = 0; i < n; ++i) Y[i + j * rho] = pr->D2[D2off + i + (size_t)j * n];
                D2off += (size_t)n * w;
            } else {
                for (int i = 0; i < n; ++i) Y[i + i * rho] = -1.0; /* test/cartpole.jl:34-42 */
            }
        }
        for (int j = 0; j < w; ++j)
            for (int i = 0; i < ps; ++i) Y[(p1 + i) + j * rho] = pr->C[Coff + i + (size_t)j * ps];
        if (p2) {
            const double *Ak = pr->A + (size_t)k * n * n, *Bk = pr->B + (size_t)k * n * m;
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) Y[(p1 + ps + i) + j * rho] = Ak[i + j * n];
            for (int j = 0; j < m; ++j)
                for (int i = 0; i < n; ++i) Y[(p1 + ps + i) + (n + j) * rho] = Bk[i + j * n];
        }

        /* ---- InvertedQuadratic: chol(H_k), g_k  (src/block_cholesky.jl:107-153) ---- */
        double *M = Mk[k] = (double *)calloc((size_t)w * w + 1, sizeof(double));
        if (!soc) {
            int st = lqro_block_cholesky(n, mk, pr->hess_mode, pr->Q + (size_t)k * n * n,
                                         mk ? pr->R + (size_t)k * m * m : NULL,
                                         (mk && pr->Hux && pr->hess_mode == LQRO_HESS_DENSE)
                                             ? pr->Hux + (size_t)k * m * n : NULL, M);
            if (st && !info) info = (k + 1) * 1000 + st;
            for (int i = 0; i < n; ++i) g[i] = pr->q[(size_t)k * n + i];
            for (int i = 0; i < mk; ++i) g[n + i] = pr->r[(size_t)k * m + i];
        }

        /* ---- shur!  (src/jacobian_blocks.jl:231-242) ---- */
        for (int j = 0; j < rho; ++j)
            for (int i = 0; i < w; ++i) JYt[i + j * w] = Y[j + i * rho]; /* transpose!(JYt, Y) */
        double *rv = rk[k] = (double *)calloc(rho + 1, sizeof(double));
        if (!soc) {
            lqro_block_ldiv(n, mk, pr->hess_mode, M, rho, JYt, w);      /* ldiv!(Jinv.chol, JYt) */
            gemm(1, 0, rho, 1, w, 1.0, JYt, w, g, w, 0.0, rv, rho);     /* r = YJ*g */
        }
        gemm(0, 0, rho, rho, w, 1.0, Y, rho, JYt, w, 0.0, YYt, rho);    /* YYt = Y*JYt */

        /* ---- copy_shur!  (src/jacobian_blocks.jl:271-286, upper variant) ---- */
        if (p1) { /* res.A .+= YYt[ip1,ip1], A aliases the previous C */
            double *Cp = T[k - 1].C;
            for (int j = 0; j < p1; ++j)
                for (int i = 0; i < p1; ++i) Cp[i + j * p1] += YYt[i + j * rho];
        }
        for (int j = 0; j < ps; ++j)
            for (int i = 0; i < ps; ++i) t->B[i + j * ps] = YYt[(p1 + i) + (p1 + j) * rho];
        for (int j = 0; j < p2; ++j)
            for (int i = 0; i < p2; ++i) t->C[i + j * p2] = YYt[(p1 + ps + i) + (p1 + ps + j) * rho];
        for (int j = 0; j < ps; ++j)
            for (int i = 0; i < p1; ++i) t->D[i + j * p1] = YYt[i + (p1 + j) * rho];
        for (int j = 0; j < p2; ++j)
            for (int i = 0; i < ps; ++i) t->E[i + j * ps] = YYt[(p1 + i) + (p1 + ps + j) * rho];
        for (int j = 0; j < p2; ++j)
            for (int i = 0; i < p1; ++i) t->F[i + j * p1] = YYt[i + (p1 + ps + j) * rho];
        for (int i = 0; i < ps; ++i) t->c[i] = rv[p1 + i] - pr->c[coff + i]; /* c = r_[2] - c */
        for (int i = 0; i < p2; ++i) t->d[i] = rv[p1 + ps + i] - pr->d[(size_t)k * n + i];
        if (p1) /* copy_shur!(F[k-1], blocks[k-1], blocks[k]): d += next.r_[1]  (:249-252) */
            for (int i = 0; i < p1; ++i) T[k - 1].d[i] += rv[i];
        coff += ps;
        Coff += (size_t)ps * w;
    }

    /* optional dense S,h  (copy_shur_factors!, src/jacobian_blocks.jl:173-211) */
    if (S_out) memset(S_out, 0, sizeof(double) * P * P);
    if (U_out) memset(U_out, 0, sizeof(double) * P * P);
#define SCATTER(dst)                                                                          \
    do {                                                                                      \
        size_t off = 0;                                                                       \
        for (int k = 0; k < N; ++k) {                                                         \
            tri3 *t = &T[k];                                                                  \
            size_t i1 = off, is = off + t->p1, i2 = off + t->p1 + t->ps;                      \
            if (t->p1) {                                                                      \
                double *Ap = T[k - 1].C;                                                      \
                for (int j = 0; j < t->p1; ++j)                                               \
                    for (int i = 0; i <= j; ++i) dst[(i1 + i) + (i1 + j) * P] = Ap[i + j * t->p1]; \
            }                                                                                 \
            for (int j = 0; j < t->ps; ++j)                                                   \
                for (int i = 0; i <= j; ++i) dst[(is + i) + (is + j) * P] = t->B[i + j * t->ps]; \
            for (int j = 0; j < t->p2; ++j)                                                   \
                for (int i = 0; i <= j; ++i) dst[(i2 + i) + (i2 + j) * P] = t->C[i + j * t->p2]; \
            for (int j = 0; j < t->ps; ++j)                                                   \
                for (int i = 0; i < t->p1; ++i) dst[(i1 + i) + (is + j) * P] = t->D[i + j * t->p1]; \
            for (int j = 0; j < t->p2; ++j)                                                   \
                for (int i = 0; i < t->ps; ++i) dst[(is + i) + (i2 + j) * P] = t->E[i + j * t->ps]; \
            for (int j = 0; j < t->p2; ++j)                                                   \
                for (int i = 0; i < t->p1; ++i) dst[(i1 + i) + (i2 + j) * P] = t->F[i + j * t->p1]; \
            off += t->p1 + t->ps;                                                             \
        }                                                                                     \
    } while (0)
    if (S_out) SCATTER(S_out);
    if (h_out) {
        size_t off = 0;
        for (int k = 0; k < N; ++k) {
            tri3 *t = &T[k];
            for (int i = 0; i < t->ps; ++i) h_out[off + t->p1 + i] = t->c[i];
            for (int i = 0; i < t->p2; ++i) h_out[off + t->p1 + t->ps + i] = t->d[i];
            off += t->p1 + t->ps;
        }
    }

    /* ---- cholesky!(chol, shur): src/cholesky_solve.jl:28-33,47-67 (in place, :69-91) ---- */
    for (int k = 0; k < N; ++k) {
        tri3 *t = &T[k];
        const int p1 = t->p1, ps = t->ps, p2 = t->p2;
        const double *Afac = p1 ? T[k - 1].C : NULL; /* U.A (already factored) */
        if (p1 && ps) trsm_ut(p1, ps, Afac, p1, t->D, p1);             /* D <- A^-T D     */
        if (ps) {
            if (p1) gemm(1, 0, ps, ps, p1, -1.0, t->D, p1, t->D, p1, 1.0, t->B, ps); /* B -= D'D */
            int st = potrf_u(ps, t->B, ps);
            if (st && !info) info = (k + 1) * 1000 + 100 + st;
        }
        if (p1 && p2) trsm_ut(p1, p2, Afac, p1, t->F, p1);             /* F <- A^-T F     */
        if (ps && p2) {
            if (p1) gemm(1, 0, ps, p2, p1, -1.0, t->D, p1, t->F, p1, 1.0, t->E, ps); /* E -= D'F */
            trsm_ut(ps, p2, t->B, ps, t->E, ps);                       /* E <- B^-T E     */
        }
        if (p2) {
            if (p1) gemm(1, 0, p2, p2, p1, -1.0, t->F, p1, t->F, p1, 1.0, t->C, p2); /* C -= F'F */
            if (ps) gemm(1, 0, p2, p2, ps, -1.0, t->E, ps, t->E, ps, 1.0, t->C, p2); /* C -= E'E */
            int st = potrf_u(p2, t->C, p2);
            if (st && !info) info = (k + 1) * 1000 + 200 + st;
        }
    }
    if (U_out) SCATTER(U_out);
#undef SCATTER

    /* ---- forward_substitution!: src/cholesky_solve.jl:93-117 ---- */
    for (int k = 0; k < N; ++k) {
        tri3 *t = &T[k];
        const int p1 = t->p1, ps = t->ps, p2 = t->p2;
        const double *lprev = p1 ? T[k - 1].lam : NULL;
        for (int i = 0; i < ps; ++i) {
            double s = t->c[i];
            for (int l = 0; l < p1; ++l) s -= t->D[l + i * p1] * lprev[l]; /* c - D'λ_prev */
            t->mu[i] = s;
        }
        if (ps) trsm_ut(ps, 1, t->B, ps, t->mu, ps);
        for (int i = 0; i < p2; ++i) {
            double s = t->d[i];
            for (int l = 0; l < p1; ++l) s -= t->F[l + i * p1] * lprev[l];
            for (int l = 0; l < ps; ++l) s -= t->E[l + i * ps] * t->mu[l];
            t->lam[i] = s;
        }
        if (p2) trsm_ut(p2, 1, t->C, p2, t->lam, p2);
    }
    /* ---- backward_substitution!: src/cholesky_solve.jl:119-143 (negates) ---- */
    for (int k = N - 1; k >= 0; --k) {
        tri3 *t = &T[k];
        const int ps = t->ps, p2 = t->p2;
        if (k < N - 1) {
            tri3 *nx = &T[k + 1]; /* "Lprev" in the reference = the later knot */
            for (int i = 0; i < p2; ++i) {
                double s = t->lam[i];
                for (int l = 0; l < nx->ps; ++l) s += nx->D[i + l * nx->p1] * nx->mu[l];
                for (int l = 0; l < nx->p2; ++l) s += nx->F[i + l * nx->p1] * nx->lam[l];
                tmpv[i] = s;
            }
            memcpy(t->lam, tmpv, sizeof(double) * p2);
        }
        if (p2) trsm_un(p2, 1, t->C, p2, t->lam, p2);
        for (int i = 0; i < ps; ++i) {
            double s = t->mu[i];
            for (int l = 0; l < p2; ++l) s -= t->E[i + l * ps] * t->lam[l];
            t->mu[i] = s;
        }
        if (ps) trsm_un(ps, 1, t->B, ps, t->mu, ps);
        for (int i = 0; i < p2; ++i) t->lam[i] = -t->lam[i];
        for (int i = 0; i < ps; ++i) t->mu[i] = -t->mu[i];
    }
    /* NOTE on the reference's sign bookkeeping: backward_substitution!(L, Lprev)
     * adds Lprev.D*Lprev.μ + Lprev.F*Lprev.λ where the successors were already
     * negated (:127,136-137), which is the same as subtracting the un-negated
     * values — i.e. plain back-substitution followed by a global negation. */

    /* ---- multipliers in flat order [μ1;λ1;μ2;λ2;…;μN] (src/jacobian_blocks.jl:181-195) ---- */
    {
        size_t off = 0;
        for (int k = 0; k < N; ++k) {
            tri3 *t = &T[k];
            for (int i = 0; i < t->ps; ++i) mult[off + i] = t->mu[i];
            for (int i = 0; i < t->p2; ++i) mult[off + t->ps + i] = t->lam[i];
            off += t->ps + t->p2;
        }
    }
    /* ---- calculate_primals!: src/cholesky_solver.jl:185-236 ---- */
    {
        size_t zoff = 0;
        for (int k = 0; k < N; ++k) {
            tri3 *t = &T[k];
            const int w = knot_w(pr, k), p1 = t->p1, ps = t->ps, p2 = t->p2, rho = p1 + ps + p2;
            const int mk = w - n;
            const double *Y = Yk[k];
            double *z = tmpv;
            for (int j = 0; j < w; ++j) {
                double s = 0.0;
                for (int i = 0; i < p2; ++i) s += Y[(p1 + ps + i) + j * rho] * t->lam[i];      /* D1'λ_k */
                for (int i = 0; i < ps; ++i) s += Y[(p1 + i) + j * rho] * t->mu[i];            /* C'μ_k  */
                for (int i = 0; i < p1; ++i) s += Y[i + j * rho] * T[k - 1].lam[i];            /* D2'λ_{k-1} */
                if (!soc) s += (j < n) ? pr->q[(size_t)k * n + j] : pr->r[(size_t)k * m + (j - n)];
                z[j] = s;
            }
            if (res_out) memcpy(res_out + zoff, z, sizeof(double) * w);
            if (!soc) lqro_block_ldiv(n, mk, pr->hess_mode, Mk[k], 1, z, w); /* calc_primals! :195-199 */
            for (int j = 0; j < w; ++j) dz[zoff + j] = -z[j];
            zoff += w;
        }
    }

    for (int k = 0; k < N; ++k) {
        free(T[k].B); free(T[k].C); free(T[k].D); free(T[k].E); free(T[k].F);
        free(T[k].c); free(T[k].d); free(T[k].mu); free(T[k].lam);
        free(Yk[k]); free(Mk[k]); free(rk[k]);
    }
    free(T); free(Yk); free(Mk); free(rk); free(JYt); free(YYt); free(g); free(tmpv);
    return info;
}

/* flat-argument wrapper for ctypes */
int lqro_kkt_solve_flat(int n, int m, int N, int hess_mode, int flags, const int32_t *p,
                        const double *Q, const double *R, const double *Hux, const double *q,
                        const double *r, const double *A, const double *B, const double *d,
                        const double *D2, const double *C, const double *c, double *dz,
                        double *mult, double *res_out, double *S_out, double *h_out,
                        double *U_out) {
    lqro_kkt_problem pr = {n, m, N, hess_mode, flags, p, Q, R, Hux, q, r, A, B, d, D2, C, c};
    return lqro_kkt_solve(&pr, dz, mult, res_out, S_out, h_out, U_out);
}

/* ------------------------------------------------------------------------ */
/* Batched drivers (instance-major, one instance per loop iteration, OpenMP  */
/* over instances) — the CPU baseline BASELINE.md §4 describes.              */
/* ------------------------------------------------------------------------ */
int lqro_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int lqro_riccati_batch(int n, int m, int N, int flags, long batch, const double *A,
                       const double *B, const double *Q, const double *R, const double *q,
                       const double *r, const double *Qf, const double *qf, const double *x0,
                       double *X, double *U, double *K, double *kff, int32_t *info,
                       int nthreads) {
    const int lti = (flags & LQRO_FLAG_LTI) != 0;
    const size_t kn = lti ? 1 : (size_t)(N - 1);
    const size_t sA = kn * n * n, sB = kn * n * m, sQ = kn * n * n, sR = kn * m * m, sq = kn * n,
                 sr = kn * m;
    const size_t wd = lqro_riccati_work_doubles(n, m);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        double *work = (double *)malloc(sizeof(double) * wd);
#pragma omp for schedule(static)
        for (long i = 0; i < batch; ++i) {
            int st = lqro_riccati(n, m, N, flags, A + i * sA, B + i * sB, Q + i * sQ, R + i * sR,
                                  q ? q + i * sq : NULL, r ? r + i * sr : NULL,
                                  Qf + (size_t)i * n * n, qf ? qf + (size_t)i * n : NULL,
                                  x0 + (size_t)i * n, X + (size_t)i * n * N,
                                  U + (size_t)i * m * (N - 1), K + (size_t)i * m * n * (N - 1),
                                  kff + (size_t)i * m * (N - 1), work);
            if (info) info[i] = st;
        }
        free(work);
    }
    return 0;
}

int lqro_kkt_batch(int n, int m, int N, int hess_mode, int flags, const int32_t *p, long batch,
                   const double *Q, const double *R, const double *Hux, const double *q,
                   const double *r, const double *A, const double *B, const double *d,
                   const double *D2, const double *C, const double *c, double *dz, double *mult,
                   double *res_out, int32_t *info, int nthreads) {
    size_t sC = 0, sc = 0, sD2 = 0;
    for (int k = 0; k < N; ++k) {
        const int w = n + (k < N - 1 ? m : 0);
        sC += (size_t)p[k] * w;
        sc += p[k];
        if (k > 0) sD2 += (size_t)n * w;
    }
    const size_t NN = lqro_kkt_num_vars(n, m, N), P = lqro_kkt_num_cons(n, N, p);
    const size_t K1 = (size_t)(N - 1);
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
    for (long i = 0; i < batch; ++i) {
        lqro_kkt_problem pr = {n, m, N, hess_mode, flags, p,
                               Q + i * (size_t)n * n * N,
                               R + i * (size_t)m * m * K1,
                               Hux ? Hux + i * (size_t)m * n * K1 : NULL,
                               q + i * (size_t)n * N,
                               r + i * (size_t)m * K1,
                               A + i * (size_t)n * n * K1,
                               B + i * (size_t)n * m * K1,
                               d + i * (size_t)n * K1,
                               D2 ? D2 + i * sD2 : NULL,
                               C + i * sC,
                               c + i * sc};
        int st = lqro_kkt_solve(&pr, dz + i * NN, mult + i * P, res_out ? res_out + i * NN : NULL,
                                NULL, NULL, NULL);
        if (info) info[i] = st;
    }
    return 0;
}
