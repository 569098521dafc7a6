// riccati_kernels.cuh — batched Riccati backward pass + forward rollout.
//
// Replaces solve!(sol, ::DPSolver, prob) : src/dynamic_programming.jl:54-72
//   compute_gain! :37-43  (PB=P*B; E=R+B'PB; PA=P*A; K=B'PA; potrf/potrs)
//   compute_ctg!  :48-52  (P_ = Q + A'PA - (A'PB) K)
//   rollout       :66-70  (u = -K x ; x+ = A x + B u)      [= src/least_squares.jl:195-202]
// generalised to per-knot (LTV) data and affine cost terms as SURVEY Appendix A states:
//   kff = E^-1 (r + B'p),  p_ = q + A'p - K'(r + B'p),  u = -K x - kff.
//
// Two kernel families:
//   riccati_tpi_kernel<n,m>   one THREAD per instance, everything in registers, tile width 32.
//                             (n=4: the per-knot work is ~250 DFMA on 36 doubles, so a warp per
//                             instance would idle most lanes and drown in shuffles; with a thread per
//                             instance each warp reads one 256-byte row of its tile per load.)
//   riccati_coop_kernel<G>    G threads per instance (32 = warp, 256 = CTA), runtime n,m, matrices
//                             in shared memory, tile width 1.  Any size; used for n+m > 9.
#pragma once
#include "smallmat.cuh"

// ------------------------------------------------------------------ thread-per-instance -------
template <int n, int m>
struct RiccatiRows {
    static constexpr int oA = 0, oB = n * n, oQ = oB + n * m, oR = oQ + tri(n), oq = oR + tri(m),
                         orr = oq + n, F = lqrb_riccati_knot_rows(n, m);
    static constexpr int TR = tri(n) + 2 * n;  // Qf | qf | x0
    static constexpr int GR = m * n + m;       // K | kff
    static constexpr int W = n + m;
};

// One backward step in registers.  P (packed sym), p updated in place; K (m x n col-major), kff out.
template <int n, int m>
__device__ __forceinline__ int riccati_step(const double *A, const double *B, const double *Q,
                                            const double *R, const double *q, const double *r,
                                            double *P, double *p, double *K, double *kff) {
    double PB[n * m], PA[n * n], E[tri(m)], Einv[m], Kraw[m * n], rr[m];
    // PB = P*B, PA = P*A     (compute_gain! :38,40)
    SM_UNROLL
    for (int j = 0; j < m; ++j)
        SM_UNROLL
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(P[sym_idx(i, l)], B[l + j * n], s);
            PB[i + j * n] = s;
        }
    SM_UNROLL
    for (int j = 0; j < n; ++j)
        SM_UNROLL
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(P[sym_idx(i, l)], A[l + j * n], s);
            PA[i + j * n] = s;
        }
    // E = R + B'PB (symmetric, packed)   (:39)
    SM_UNROLL
    for (int j = 0; j < m; ++j)
        SM_UNROLL
        for (int i = 0; i <= j; ++i) {
            double s = R[tri_idx(i, j)];
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(B[l + i * n], PB[l + j * n], s);
            E[tri_idx(i, j)] = s;
        }
    // Kraw = B'PA  (= (A'PB)' since P is symmetric: the reference's APB, :50)   (:41)
    SM_UNROLL
    for (int j = 0; j < n; ++j)
        SM_UNROLL
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(B[l + i * n], PA[l + j * n], s);
            Kraw[i + j * m] = s;
            K[i + j * m] = s;
        }
    SM_UNROLL
    for (int i = 0; i < m; ++i) {
        double s = r[i];
        SM_UNROLL
        for (int l = 0; l < n; ++l) s = fma(B[l + i * n], p[l], s);
        rr[i] = s;
        kff[i] = s;
    }
    // chol_solve!(E, K)  (:28-31)
    const int st = chol_packed<m>(E, Einv);
    SM_UNROLL
    for (int j = 0; j < n; ++j) solve_chol<m>(E, Einv, K + j * m);
    solve_chol<m>(E, Einv, kff);
    // p_ = q + A'p - K'rr
    double pn[n];
    SM_UNROLL
    for (int i = 0; i < n; ++i) {
        double s = q[i];
        SM_UNROLL
        for (int l = 0; l < n; ++l) s = fma(A[l + i * n], p[l], s);
        SM_UNROLL
        for (int l = 0; l < m; ++l) s = fma(-K[l + i * m], rr[l], s);
        pn[i] = s;
    }
    // P_ = Q + A'PA - APB*K, upper triangle only (the result is symmetric)   (compute_ctg! :50-51)
    SM_UNROLL
    for (int j = 0; j < n; ++j)
        SM_UNROLL
        for (int i = 0; i <= j; ++i) {
            double s = Q[tri_idx(i, j)];
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(A[l + i * n], PA[l + j * n], s);
            SM_UNROLL
            for (int l = 0; l < m; ++l) s = fma(-Kraw[l + i * m], K[l + j * m], s);
            P[tri_idx(i, j)] = s;
        }
    SM_UNROLL
    for (int i = 0; i < n; ++i) p[i] = pn[i];
    return st;
}

template <int n, int m, bool LTI, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
    riccati_tpi_kernel(const double *__restrict__ knots, const double *__restrict__ term,
                       double *__restrict__ Z, double *__restrict__ gains,
                       int32_t *__restrict__ info, int N, int64_t batch) {
    using L = RiccatiRows<n, m>;
    const int64_t inst = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    if (inst >= batch) return;
    const int64_t tile = inst >> 5;
    const int lane = (int)(inst & 31);
    const int Kn = LTI ? 1 : N - 1;
    const double *kb = knots + tile * Kn * L::F * 32 + lane;
    const double *tb = term + tile * L::TR * 32 + lane;
    double *zb = Z + tile * ((int64_t)N * n + (int64_t)(N - 1) * m) * 32 + lane;
    double *gb = gains + tile * (int64_t)(N - 1) * L::GR * 32 + lane;

    double P[tri(n)], p[n];
    SM_UNROLL
    for (int e = 0; e < tri(n); ++e) P[e] = ld_stream(tb + e * 32);
    SM_UNROLL
    for (int e = 0; e < n; ++e) p[e] = ld_stream(tb + (tri(n) + e) * 32);

    double A[n * n], B[n * m];
    int st_all = 0;
    // ---------------- backward pass: k = N-2 .. 0   (src/dynamic_programming.jl:61-64)
    for (int k = N - 2; k >= 0; --k) {
        const double *kp = kb + (int64_t)(LTI ? 0 : k) * L::F * 32;
        double Q[tri(n)], R[tri(m)], q[n], r[m];
        SM_UNROLL
        for (int e = 0; e < n * n; ++e) A[e] = ld_keep(kp + (L::oA + e) * 32);
        SM_UNROLL
        for (int e = 0; e < n * m; ++e) B[e] = ld_keep(kp + (L::oB + e) * 32);
        SM_UNROLL
        for (int e = 0; e < tri(n); ++e) Q[e] = ld_stream(kp + (L::oQ + e) * 32);
        SM_UNROLL
        for (int e = 0; e < tri(m); ++e) R[e] = ld_stream(kp + (L::oR + e) * 32);
        SM_UNROLL
        for (int e = 0; e < n; ++e) q[e] = ld_stream(kp + (L::oq + e) * 32);
        SM_UNROLL
        for (int e = 0; e < m; ++e) r[e] = ld_stream(kp + (L::orr + e) * 32);
        double K[m * n], kff[m];
        const int st = riccati_step<n, m>(A, B, Q, R, q, r, P, p, K, kff);
        if (st != 0 && st_all == 0) st_all = (k + 1) * 1000 + st;
        double *gk = gb + (int64_t)k * L::GR * 32;
        SM_UNROLL
        for (int e = 0; e < m * n; ++e) gk[e * 32] = K[e];
        SM_UNROLL
        for (int e = 0; e < m; ++e) gk[(m * n + e) * 32] = kff[e];
    }
    if (info) info[inst] = st_all;

    // ---------------- forward rollout   (src/dynamic_programming.jl:66-70)
    double x[n];
    SM_UNROLL
    for (int e = 0; e < n; ++e) x[e] = ld_stream(tb + (tri(n) + n + e) * 32);
    for (int k = 0; k < N - 1; ++k) {
        const double *kp = kb + (int64_t)(LTI ? 0 : k) * L::F * 32;
        const double *gk = gb + (int64_t)k * L::GR * 32;
        double K[m * n], kff[m], u[m], xn[n];
        SM_UNROLL
        for (int e = 0; e < n * n; ++e) A[e] = ld_stream(kp + (L::oA + e) * 32);
        SM_UNROLL
        for (int e = 0; e < n * m; ++e) B[e] = ld_stream(kp + (L::oB + e) * 32);
        SM_UNROLL
        for (int e = 0; e < m * n; ++e) K[e] = gk[e * 32];
        SM_UNROLL
        for (int e = 0; e < m; ++e) kff[e] = gk[(m * n + e) * 32];
        double *zk = zb + (int64_t)k * L::W * 32;
        SM_UNROLL
        for (int i = 0; i < n; ++i) __stcs(zk + i * 32, x[i]);
        SM_UNROLL
        for (int i = 0; i < m; ++i) {
            double s = -kff[i];
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(-K[i + l * m], x[l], s);
            u[i] = s;
            __stcs(zk + (n + i) * 32, s);
        }
        SM_UNROLL
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n; ++l) s = fma(A[i + l * n], x[l], s);
            SM_UNROLL
            for (int l = 0; l < m; ++l) s = fma(B[i + l * n], u[l], s);
            xn[i] = s;
        }
        SM_UNROLL
        for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    double *zN = zb + (int64_t)(N - 1) * L::W * 32;
    SM_UNROLL
    for (int i = 0; i < n; ++i) __stcs(zN + i * 32, x[i]);
}

// ------------------------------------------------------------------ cooperative (any size) ----
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G <= 32)
        __syncwarp();
    else
        __syncthreads();
}

// shared memory doubles one instance needs
__host__ __device__ inline size_t riccati_coop_smem_doubles(int n, int m) {
    return (size_t)4 * n * n + 4 * n * m + m * m + 4 * n + 3 * m + 8;
}

// G threads per instance, (THREADS/G) instances per CTA; packed layout tile width 1.
template <int G, int THREADS>
__global__ void __launch_bounds__(THREADS)
    riccati_coop_kernel(const double *__restrict__ knots, const double *__restrict__ term,
                        double *__restrict__ Z, double *__restrict__ gains,
                        int32_t *__restrict__ info, int n, int m, int N, int lti, int64_t batch) {
    extern __shared__ double smem[];
    constexpr int IPC = THREADS / G;
    const int g = threadIdx.x / G, t = threadIdx.x % G;
    const int64_t inst_raw = (int64_t)blockIdx.x * IPC + g;
    const bool active = inst_raw < batch;
    const int64_t inst = active ? inst_raw : batch - 1;  // idle groups shadow a valid instance
    const int nn = n * n, nm = n * m, tn = tri(n), tm = tri(m);
    const int F = lqrb_riccati_knot_rows(n, m), TR = tn + 2 * n, GR = m * n + m, W = n + m;
    const int Kn = lti ? 1 : N - 1;
    double *s = smem + (size_t)g * riccati_coop_smem_doubles(n, m);
    double *P = s, *Pn = P + nn, *PA = Pn + nn, *A = PA + nn, *B = A + nn, *PB = B + nm,
           *Kraw = PB + nm, *K = Kraw + nm, *E = K + nm, *p = E + m * m, *pn = p + n, *x = pn + n,
           *xn = x + n, *rr = xn + n, *kff = rr + m, *u = kff + m;
    const double *rec = knots + inst * (int64_t)Kn * F;
    const double *tb = term + inst * TR;
    double *zb = Z + inst * ((int64_t)N * n + (int64_t)(N - 1) * m);
    double *gb = gains + inst * (int64_t)(N - 1) * GR;
    int st_all = 0;

    for (int e = t; e < nn; e += G) {
        const int i = e % n, j = e / n;
        P[e] = tb[sym_idx(i, j)];
    }
    for (int e = t; e < n; e += G) p[e] = tb[tn + e];
    group_sync<G>();

    for (int k = N - 2; k >= 0; --k) {
        const double *kp = rec + (int64_t)(lti ? 0 : k) * F;
        const double *Qp = kp + nn + nm, *Rp = Qp + tn, *qp = Rp + tm, *rp = qp + n;
        for (int e = t; e < nn; e += G) A[e] = kp[e];
        for (int e = t; e < nm; e += G) B[e] = kp[nn + e];
        group_sync<G>();
        for (int e = t; e < nn + nm; e += G) {  // PA = P*A ; PB = P*B
            const bool isA = e < nn;
            const int ee = isA ? e : e - nn;
            const int i = ee % n, j = ee / n;
            const double *col = (isA ? A : B) + j * n;
            double acc = 0.0;
            for (int l = 0; l < n; ++l) acc = fma(P[i + l * n], col[l], acc);
            (isA ? PA : PB)[ee] = acc;
        }
        group_sync<G>();
        for (int e = t; e < m * m + nm + m; e += G) {  // E = R + B'PB ; Kraw = B'PA ; rr = r + B'p
            if (e < m * m) {
                const int i = e % m, j = e / m;
                double acc = Rp[sym_idx(i, j)];
                for (int l = 0; l < n; ++l) acc = fma(B[l + i * n], PB[l + j * n], acc);
                E[e] = acc;
            } else if (e < m * m + nm) {
                const int ee = e - m * m, i = ee % m, j = ee / m;
                double acc = 0.0;
                for (int l = 0; l < n; ++l) acc = fma(B[l + i * n], PA[l + j * n], acc);
                Kraw[ee] = acc;
                K[ee] = acc;
            } else {
                const int i = e - m * m - nm;
                double acc = rp[i];
                for (int l = 0; l < n; ++l) acc = fma(B[l + i * n], p[l], acc);
                rr[i] = acc;
                kff[i] = acc;
            }
        }
        group_sync<G>();
        // right-looking upper Cholesky of E (m x m), cooperative
        for (int j = 0; j < m; ++j) {
            const double djj = E[j + j * m];
            if (!(djj > 0.0) && st_all == 0) st_all = (k + 1) * 1000 + j + 1;
            const double d = sqrt(djj);
            group_sync<G>();
            for (int i = j + t; i < m; i += G) E[j + i * m] = (i == j) ? d : E[j + i * m] / d;
            group_sync<G>();
            const int rem = m - j - 1;
            for (int e = t; e < rem * rem; e += G) {
                const int a = j + 1 + e % rem, b = j + 1 + e / rem;
                if (a <= b) E[a + b * m] = fma(-E[j + a * m], E[j + b * m], E[a + b * m]);
            }
            group_sync<G>();
        }
        // potrs: each thread solves one column of [K | kff]
        for (int c = t; c < n + 1; c += G) {
            double *col = (c < n) ? K + c * m : kff;
            for (int i = 0; i < m; ++i) {
                double acc = col[i];
                for (int l = 0; l < i; ++l) acc = fma(-E[l + i * m], col[l], acc);
                col[i] = acc / E[i + i * m];
            }
            for (int i = m - 1; i >= 0; --i) {
                double acc = col[i];
                for (int l = i + 1; l < m; ++l) acc = fma(-E[i + l * m], col[l], acc);
                col[i] = acc / E[i + i * m];
            }
        }
        group_sync<G>();
        for (int e = t; e < nn + n; e += G) {  // P_ = Q + A'PA - Kraw'K ; p_ = q + A'p - K'rr
            if (e < nn) {
                const int i = e % n, j = e / n;
                double acc = Qp[sym_idx(i, j)];
                for (int l = 0; l < n; ++l) acc = fma(A[l + i * n], PA[l + j * n], acc);
                for (int l = 0; l < m; ++l) acc = fma(-Kraw[l + i * m], K[l + j * m], acc);
                Pn[e] = acc;
            } else {
                const int i = e - nn;
                double acc = qp[i];
                for (int l = 0; l < n; ++l) acc = fma(A[l + i * n], p[l], acc);
                for (int l = 0; l < m; ++l) acc = fma(-K[l + i * m], rr[l], acc);
                pn[i] = acc;
            }
        }
        if (active) {
            double *gk = gb + (int64_t)k * GR;
            for (int e = t; e < nm; e += G) gk[e] = K[e];
            for (int e = t; e < m; e += G) gk[nm + e] = kff[e];
        }
        group_sync<G>();
        // symmetrise while copying back (upper triangle is the reference value)
        for (int e = t; e < nn; e += G) {
            const int i = e % n, j = e / n;
            P[e] = (i <= j) ? Pn[e] : Pn[j + i * n];
        }
        for (int e = t; e < n; e += G) p[e] = pn[e];
        group_sync<G>();
    }
    if (active && info && t == 0) info[inst] = st_all;

    for (int e = t; e < n; e += G) x[e] = tb[tn + n + e];
    group_sync<G>();
    for (int k = 0; k < N - 1; ++k) {
        const double *kp = rec + (int64_t)(lti ? 0 : k) * F;
        const double *Ag = kp, *Bg = kp + nn;
        const double *gk = gb + (int64_t)k * GR;
        double *zk = zb + (int64_t)k * W;
        for (int i = t; i < m; i += G) {
            double acc = -gk[nm + i];
            for (int l = 0; l < n; ++l) acc = fma(-gk[i + l * m], x[l], acc);
            u[i] = acc;
        }
        group_sync<G>();
        for (int i = t; i < n; i += G) {
            double acc = 0.0;
            for (int l = 0; l < n; ++l) acc = fma(Ag[i + l * n], x[l], acc);
            for (int l = 0; l < m; ++l) acc = fma(Bg[i + l * n], u[l], acc);
            xn[i] = acc;
        }
        if (active) {
            for (int i = t; i < n; i += G) zk[i] = x[i];
            for (int i = t; i < m; i += G) zk[n + i] = u[i];
        }
        group_sync<G>();
        for (int i = t; i < n; i += G) x[i] = xn[i];
        group_sync<G>();
    }
    if (active) {
        double *zN = zb + (int64_t)(N - 1) * W;
        for (int i = t; i < n; i += G) zN[i] = x[i];
    }
}
