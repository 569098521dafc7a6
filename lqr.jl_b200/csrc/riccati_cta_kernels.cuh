// riccati_cta_kernels.cuh — CTA-per-instance Riccati recursion for the large-state class (n = 32, 64;
// m = 8, 16) where the per-knot updates are real dense contractions: everything of order n^3 runs on the
// FP64 tensor cores (mma.sync m8n8k4 DMMA).
//
// Replaces solve!(sol, ::DPSolver, prob) : src/dynamic_programming.jl:54-72
//   compute_gain! :37-43  PB, PA, E = R + B'PB, K = E^-1 B'PA      compute_ctg! :48-52  P_ = Q + A'PA - A'PB K
// in the same symmetric form as riccati_dmma_kernels.cuh:  F = [A B],  T = F'P,  M = T F + blkdiag(Q,R),
//   K = Muu^-1 Mux,  P_ = Mxx - Mxu K,   plus the affine terms  g^ = [q;r] + F'p, kff = Muu^-1 g^u,
//   p_ = g^x - Mxu kff.
// Work split: warp w of the n/8 warps owns state rows 8w..8w+7.
//   * T row tile w (8 x n) = F[:,rows w]' P stays in its accumulators and is fed straight back as the
//     A operand of M = T F (a C fragment is an A fragment under a permuted contraction index).
//   * M is symmetric: warp w forms only the tiles (w, w+j), j = 0..NT/2, plus its 8 x m slice of Mxu; the
//     new P tile is written to shared memory together with its mirror image, which also keeps P bitwise
//     symmetric (an antisymmetric rounding residue would be amplified by the open-loop |A|^2 per knot).
//   * Muu = B'PB is reduced over the warps (each contributes the contraction over its 8 states), warp 0
//     inverts it (Gauss-Jordan, one column per lane) while warp 1 already streams the next knot in, then
//     K's slice for tile w is Muu^-1 times the warp's own Mxu accumulators read as a (transposed) B fragment.
// Shared memory per CTA (n=64: 102 KB -> 2 CTAs per SM, so one CTA's serial phases hide under the other's
// DMMAs): P (ld n+8), K' (n x (m+8)), F = [A B] column-wise with ld n+8 (every fragment is one
// conflict-free 16-byte load; filled by one 512-byte cp.async.bulk per column), R|q|r, and small vectors.  Q is read once per
// knot straight from global memory into the M accumulators (only the owned tiles).
#pragma once
#include "riccati_dmma_kernels.cuh"

namespace rcta {
using rdmma::bulk_g2s;
using rdmma::mbar_expect_tx;
using rdmma::mbar_init;
using rdmma::mbar_wait;
using rdmma::mma884;

template <int n, int m>
struct Cfg {
    static_assert(n % 8 == 0 && m % 8 == 0 && m <= 16 && m <= n && n >= 16 && n <= 64, "tile map: n = 16..64 (multiples of 8), m = 8, 16");
    static constexpr int NT = n / 8, UT = m / 8, WARPS = NT, THREADS = WARPS * 32, w = n + m;
    static constexpr int JT = NT / 2 + 1;  // owned column tiles of M per warp (see own_ct)
    // every fragment is read as one 16-byte load at [row g][8*blk + 2q]: leading dimensions = 8 (mod 16)
    // doubles make each quarter-warp hit 8 distinct 16-byte bank groups
    static constexpr int LP = n + 8, LF = n + 8, LK = m + 8, LM = m + 8;
    static constexpr int oQ = n * w, oR = oQ + tri(n), oq = oR + tri(m), orr = oq + n, F = orr + m;
    static constexpr int HS = tri(m) + n + m;  // R | q | r  (one bulk copy)
    static constexpr int TR = tri(n) + 2 * n, GR = m * n + m;
    static_assert(F % 2 == 0 && oR % 2 == 0 && HS % 2 == 0, "bulk copies need 16-byte pieces");
    // shared memory map (doubles)
    // the WARPS partial-sum slots of Muu alias K' | z | red | pad: K' is written only after the slots are consumed,
    // z and red belong to the forward pass
    static constexpr int sP = 0, sK = sP + n * LP, sZ = sK + n * LK, sRed = sZ + 2 * w,
                         sPadEnd = sK + (WARPS * m * m > n * LK + 2 * w + WARPS * m + 4 * n ? WARPS * m * m : n * LK + 2 * w + WARPS * m + 4 * n),
                         sF = sPadEnd, sH = sF + w * LF, sPv = sH + HS, sG = sPv + n, sKff = sG + w, sMi = sKff + m,
                         sCol = sMi + m * LM, sBar = sCol + 2 * m, TOTAL = sBar + 2, sSlot = sK;
    static_assert(sZ - sP >= w * LF, "forward pass double-buffers [A B] in the P|K' region");
    static constexpr size_t SMEM = (size_t)TOTAL * 8;
};

// Which tiles of the symmetric M does warp wp form?  Tile (wp, (wp + j) % NT) for j < NT/2, plus the
// antipodal tile j = NT/2 for the lower half of the warps: every unordered pair exactly once, and the two
// warps that share an SM sub-partition (wp, wp + NT/2) carry NT/2 + NT/2 + 1 tiles together.
template <int NT>
__device__ __forceinline__ bool owns_j(int wp, int j) {
    if (NT % 2) return j <= NT / 2;  // odd NT: (NT-1)/2 ring tiles each, no antipodal tile
    return j < NT / 2 || (j == NT / 2 && wp < NT / 2);
}
template <int NT>
__device__ __forceinline__ int own_ct(int wp, int j) {
    return (wp + j) % NT;
}

template <int n, int m>
__global__ void __launch_bounds__(Cfg<n, m>::THREADS, 2)
    riccati_cta_kernel(const double *__restrict__ knots, const double *__restrict__ term,
                       double *__restrict__ Z, double *__restrict__ gains, int32_t *__restrict__ info,
                       int N, int lti, int64_t batch) {
    using C = Cfg<n, m>;
    constexpr int NT = C::NT, UT = C::UT, JT = C::JT, WARPS = C::WARPS, THREADS = C::THREADS, w = C::w;
    constexpr int LP = C::LP, LF = C::LF, LK = C::LK, LM = C::LM, F = C::F, GR = C::GR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *Ps = sm + C::sP, *Kt = sm + C::sK, *Fs = sm + C::sF, *Hs = sm + C::sH, *pv = sm + C::sPv,
           *gh = sm + C::sG, *kffs = sm + C::sKff, *Mi = sm + C::sMi, *colb = sm + C::sCol,
           *slot = sm + C::sSlot, *zs = sm + C::sZ, *red = sm + C::sRed;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + C::sBar);
    const double *Rs = Hs, *qs = Hs + tri(m), *rs = qs + n;

    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int64_t inst = blockIdx.x;
    const int Kn = lti ? 1 : N - 1;
    const int steps = N - 1;
    const double *rec_g = knots + inst * (int64_t)Kn * F;
    const double *tb = term + inst * C::TR;
    double *zb = Z + inst * ((int64_t)N * n + (int64_t)(N - 1) * m);
    double *gb = gains + inst * (int64_t)(N - 1) * GR;

    // ---------------- terminal cost-to-go: P = Qf (full storage), p = qf
    for (int e = tid; e < n * n; e += THREADS) {
        const int i = e % n, j = e / n;
        Ps[i * LP + j] = tb[sym_idx(i, j)];
    }
    if (tid < n) pv[tid] = tb[tri(n) + tid];
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one warp streams a knot into shared memory: [A B] column by column (padded ld), then R|q|r
    auto issue_knot = [&](int k, double *Fdst, bool with_cost, uint64_t *b) {
        const double *src = rec_g + (int64_t)(lti ? 0 : k) * F;
        if (lane == 0) mbar_expect_tx(b, (uint32_t)((w * n + (with_cost ? C::HS : 0)) * 8));
        __syncwarp();
        for (int c = lane; c < w; c += 32) bulk_g2s(Fdst + c * LF, src + c * n, n * 8, b);
        if (with_cost && lane == 0) bulk_g2s(Hs, src + C::oR, C::HS * 8, b);
    };
    if (wp == 0) issue_knot(steps - 1, Fs, true, bar);

    int st_all = 0;
    uint32_t ph0 = 0, ph1 = 0;  // phase parities of the two mbarriers (every thread waits on every use)
    const int xr = 8 * wp + g;  // the state row this lane's accumulator rows belong to
    // ---------------- backward pass: k = N-2 .. 0   (src/dynamic_programming.jl:61-64)
    for (int it = 0; it < steps; ++it) {
        const int k = steps - 1 - it;
        const double *recg = rec_g + (int64_t)(lti ? 0 : k) * F;
        // M accumulators start at Q (owned tiles, read once from global) and 0 (control columns)
        double M[JT][2], Mu[UT][2];
        SM_UNROLL
        for (int j = 0; j < JT; ++j) {
            const int ct = own_ct<NT>(wp, j);
            SM_UNROLL
            for (int e = 0; e < 2; ++e) {
                const int c = 8 * ct + 2 * q + e;
                M[j][e] = owns_j<NT>(wp, j) ? __ldg(recg + C::oQ + (xr <= c ? c * (c + 1) / 2 + xr : xr * (xr + 1) / 2 + c)) : 0.0;
            }
        }
        SM_UNROLL
        for (int ut = 0; ut < UT; ++ut) Mu[ut][0] = Mu[ut][1] = 0.0;
        // pull the next knot's record into L2 so that its bulk copies (issued after the GEMMs) are short
        if (tid == 0 && !lti && it + 1 < steps)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(recg - F), "r"(F * 8) : "memory");

        mbar_wait(bar, ph0);
        ph0 ^= 1;

        // ---- phase 1: this warp's share of Muu = R + B'PB (contraction over its own 8 states)
        {
            double Tu[UT][2];
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) Tu[ut][0] = Tu[ut][1] = 0.0;
            double gua[UT][2];
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) gua[ut][0] = gua[ut][1] = 0.0;
            const double *fu = Fs + (n + g) * LF + 2 * q;       // B[8cp+2q+e][u = 8ut+g]
            const double *pw = Ps + xr * LP + 2 * q;            // P[8cp+2q+e][8wp+g] (P symmetric)
            SM_UNROLL
            for (int cp = 0; cp < NT; ++cp) {
                const double2 bw = *reinterpret_cast<const double2 *>(pw + 8 * cp);
                const double2 pk = *reinterpret_cast<const double2 *>(pv + 8 * cp + 2 * q);
                SM_UNROLL
                for (int ut = 0; ut < UT; ++ut) {
                    const double2 au = *reinterpret_cast<const double2 *>(fu + 8 * ut * LF + 8 * cp);
                    mma884(Tu[ut][0], Tu[ut][1], au.x, bw.x);
                    mma884(Tu[ut][0], Tu[ut][1], au.y, bw.y);
                    if (wp == 0) {  // B'p is the same in every warp: only warp 0 publishes it
                        gua[ut][0] = fma(au.x, pk.x, gua[ut][0]);
                        gua[ut][1] = fma(au.y, pk.y, gua[ut][1]);
                    }
                }
            }
            if (wp == 0) {  // g^u = r + B'p
                SM_UNROLL
                for (int ut = 0; ut < UT; ++ut) {
                    double s = gua[ut][0] + gua[ut][1];
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (q == 0) gh[n + 8 * ut + g] = s + rs[8 * ut + g];
                }
            }
            double *sl = slot + wp * m * m;
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut)
                SM_UNROLL
                for (int vt = 0; vt < UT; ++vt) {
                    double mp[2];
                    SM_UNROLL
                    for (int e = 0; e < 2; ++e) {
                        const int a = 8 * ut + g, b = 8 * vt + 2 * q + e;
                        mp[e] = wp == 0 ? Rs[a <= b ? b * (b + 1) / 2 + a : a * (a + 1) / 2 + b] : 0.0;
                    }
                    const double2 b = *reinterpret_cast<const double2 *>(Fs + (n + 8 * vt + g) * LF + 8 * wp + 2 * q);
                    mma884(mp[0], mp[1], Tu[ut][0], b.x);
                    mma884(mp[0], mp[1], Tu[ut][1], b.y);
                    *reinterpret_cast<double2 *>(sl + (8 * ut + g) * m + 8 * vt + 2 * q) = make_double2(mp[0], mp[1]);
                }
        }
        // ---- phase 2: T row tile = F[:, rows]' P  (compute_gain! :38,40); g^x = q + A'p
        double T[NT][2];
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct) T[ct][0] = T[ct][1] = 0.0;
        {
            double ga0 = 0.0, ga1 = 0.0;
            const double *fa = Fs + xr * LF + 2 * q;  // F[8cp+2q+e][x row]
            const double *pb = Ps + g * LP + 2 * q;   // P[8cp+2q+e][8ct+g], read through the mirror image
            SM_UNROLL
            for (int cp = 0; cp < NT; ++cp) {
                const double2 a = *reinterpret_cast<const double2 *>(fa + 8 * cp);
                const double2 pk = *reinterpret_cast<const double2 *>(pv + 8 * cp + 2 * q);
                ga0 = fma(a.x, pk.x, ga0);
                ga1 = fma(a.y, pk.y, ga1);
                SM_UNROLL
                for (int ct = 0; ct < NT; ++ct) {
                    const double2 b = *reinterpret_cast<const double2 *>(pb + 8 * ct * LP + 8 * cp);
                    mma884(T[ct][0], T[ct][1], a.x, b.x);
                    mma884(T[ct][0], T[ct][1], a.y, b.y);
                }
            }
            double gacc = ga0 + ga1;
            gacc += __shfl_xor_sync(0xffffffffu, gacc, 1);
            gacc += __shfl_xor_sync(0xffffffffu, gacc, 2);
            if (q == 0) gh[xr] = gacc + qs[xr];
        }
        // ---- phase 3: M += T F : owned state tiles and the control columns  (E :39, K :41, A'PA :50)
        SM_UNROLL
        for (int cp = 0; cp < NT; ++cp) {
            const double ta0 = T[cp][0], ta1 = T[cp][1];
            const double *fb = Fs + g * LF + 8 * cp + 2 * q;  // F[8cp+2q+e][z = 8ct+g]
            SM_UNROLL
            for (int j = 0; j < JT; ++j) {
                if (owns_j<NT>(wp, j)) {
                    const int ct = own_ct<NT>(wp, j);
                    const double2 b = *reinterpret_cast<const double2 *>(fb + 8 * ct * LF);
                    mma884(M[j][0], M[j][1], ta0, b.x);
                    mma884(M[j][0], M[j][1], ta1, b.y);
                }
            }
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) {
                const double2 b = *reinterpret_cast<const double2 *>(fb + (n + 8 * ut) * LF);
                mma884(Mu[ut][0], Mu[ut][1], ta0, b.x);
                mma884(Mu[ut][0], Mu[ut][1], ta1, b.y);
            }
        }
        __syncthreads();  // every warp is done with F and R|q|r; the Muu partial sums are in the slots

        if (wp == 1 && it + 1 < steps) issue_knot(k - 1, Fs, true, bar);  // overlaps the phases below
        if (wp == 0) {
            // ---- Muu^-1 by Gauss-Jordan (chol_solve! :28-31 applied by multiplication).  Two lanes per
            // column (lane = 16*half + column, each holds m/2 rows) keep the number of FP64 instructions low:
            // they queue behind the other CTA's DMMAs.  Pivots are the squared Cholesky pivots (potrf's sign test).
            constexpr int HR = m / 2;
            const int hh = lane >> 4, j = (lane & 15) < m ? (lane & 15) : 0;
            double a[HR];
            SM_UNROLL
            for (int i = 0; i < HR; ++i) {
                const double *sp = slot + (HR * hh + i) * m + j;
                double s0 = sp[0], s1 = sp[m * m];  // WARPS >= 2
                SM_UNROLL
                for (int sl = 2; sl < WARPS; ++sl) {
                    if (sl & 1) s1 += sp[sl * m * m];
                    else s0 += sp[sl * m * m];
                }
                a[i] = s0 + s1;
            }
            int bad = 0;
            SM_UNROLL
            for (int kk = 0; kk < m; ++kk) {
                double *cb = colb + (kk & 1) * m;  // double-buffered: one warp barrier per pivot
                // the pivot itself comes by shuffle, so that its reciprocal (the longest link of the chain) runs
                // while the pivot column makes its round trip through shared memory
                const double ckk = __shfl_sync(0xffffffffu, a[kk % HR], ((kk / HR) << 4) | kk);
                if (!(ckk > 0.0) && bad == 0) bad = kk + 1;
                const double p = rdmma::fast_rcp3(ckk);
                if ((lane & 15) == kk) {
                    SM_UNROLL
                    for (int i = 0; i < HR; i += 2) *reinterpret_cast<double2 *>(cb + HR * hh + i) = make_double2(a[i], a[i + 1]);
                }
                // row kk of this lane's column lives in the half kk / HR
                const double akk = __shfl_sync(0xffffffffu, a[kk % HR], ((kk / HR) << 4) | (lane & 15));
                __syncwarp();
                double c[HR];
                SM_UNROLL
                for (int i = 0; i < HR; i += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(cb + HR * hh + i);
                    c[i] = v.x;
                    c[i + 1] = v.y;
                }
                const bool piv = (lane & 15) == kk;
                const double f = piv ? -p : akk * p;
                SM_UNROLL
                for (int i = 0; i < HR; ++i) {
                    if (i == kk % HR) {
                        const double upd = piv ? c[i] * f : fma(-c[i], f, a[i]);
                        a[i] = hh == kk / HR ? (piv ? p : f) : upd;
                    } else {
                        a[i] = piv ? c[i] * f : fma(-c[i], f, a[i]);
                    }
                }
            }
            if (bad != 0 && st_all == 0) st_all = (k + 1) * 1000 + bad;
            {
                double kf0 = 0.0, kf1 = 0.0;  // two chains: this warp is the critical path of the CTA
                SM_UNROLL
                for (int i = 0; i < HR; i += 2) {
                    if ((lane & 15) < m) {
                        Mi[(HR * hh + i) * LM + j] = a[i];
                        Mi[(HR * hh + i + 1) * LM + j] = a[i + 1];
                    }
                    kf0 = fma(a[i], gh[n + HR * hh + i], kf0);  // Muu^-1 symmetric: column j dotted with g^u
                    kf1 = fma(a[i + 1], gh[n + HR * hh + i + 1], kf1);
                }
                double kf = kf0 + kf1;
                kf += __shfl_xor_sync(0xffffffffu, kf, 16);
                if (lane < m) {
                    kffs[lane] = kf;
                    gb[(int64_t)k * GR + m * n + lane] = kf;
                }
            }
        }

        __syncthreads();  // Muu^-1 and kff are published

        // ---- K[:, own tile] = Muu^-1 Mux[:, own tile]; Mux is this warp's Mxu accumulators transposed
        SM_UNROLL
        for (int rt = 0; rt < UT; ++rt) {
            double acc[2] = {0.0, 0.0};
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) {
                const double2 av = *reinterpret_cast<const double2 *>(Mi + (8 * rt + g) * LM + 8 * ut + 2 * q);
                mma884(acc[0], acc[1], av.x, Mu[ut][0]);
                mma884(acc[0], acc[1], av.y, Mu[ut][1]);
            }
            // acc = K[u = 8rt+g][x = 8wp+2q+e]
            SM_UNROLL
            for (int e = 0; e < 2; ++e) {
                const int x = 8 * wp + 2 * q + e, u = 8 * rt + g;
                Kt[x * LK + u] = acc[e];
                gb[(int64_t)k * GR + u + m * x] = acc[e];
            }
        }
        __syncthreads();

        // ---- P_ = Mxx - Mxu K (owned tiles), p_ = g^x - Mxu kff   (compute_ctg! :50-51)
        SM_UNROLL
        for (int ut = 0; ut < UT; ++ut) {
            const double na0 = -Mu[ut][0], na1 = -Mu[ut][1];
            SM_UNROLL
            for (int j = 0; j < JT; ++j) {
                if (owns_j<NT>(wp, j)) {
                    const int ct = own_ct<NT>(wp, j);
                    const double2 b = *reinterpret_cast<const double2 *>(Kt + (8 * ct + g) * LK + 8 * ut + 2 * q);
                    mma884(M[j][0], M[j][1], na0, b.x);
                    mma884(M[j][0], M[j][1], na1, b.y);
                }
            }
        }
        {
            double s = 0.0;
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) {
                const double2 kf = *reinterpret_cast<const double2 *>(kffs + 8 * ut + 2 * q);
                s = fma(Mu[ut][0], kf.x, s);
                s = fma(Mu[ut][1], kf.y, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) pv[xr] = gh[xr] - s;
        }
        SM_UNROLL
        for (int j = 0; j < JT; ++j) {
            if (owns_j<NT>(wp, j)) {
                const int ct = own_ct<NT>(wp, j);
                SM_UNROLL
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * ct + 2 * q + e;
                    if (j > 0 || xr <= c) {  // diagonal tile: the upper triangle is the value, mirrored
                        Ps[xr * LP + c] = M[j][e];
                        Ps[c * LP + xr] = M[j][e];
                    }
                }
            }
        }
        __syncthreads();
    }
    if (info && tid == 0) info[inst] = st_all;

    // ---------------- forward rollout   (src/dynamic_programming.jl:66-70)
    // [A B] of knot `it` alternates between the F region and the (now free) P|K' region
    // The rollout reads [A B] row by row (thread = row), so the padded leading dimension of the DMMA fragments is
    // not needed here: ONE 8 n w-byte bulk copy per knot instead of w copies of one column each (the copy engine's
    // rate for 512-byte pieces, not HBM, was what a forward step waited for).
    double *Fb[2] = {Fs, Ps};
    auto issue_fwd = [&](int k, double *Fdst, uint64_t *b) {
        if (lane == 0) {
            mbar_expect_tx(b, (uint32_t)(w * n * 8));
            bulk_g2s(Fdst, rec_g + (int64_t)(lti ? 0 : k) * F, w * n * 8, b);
        }
    };
    if (wp == 0) {
        issue_fwd(0, Fb[0], bar);
        if (steps > 1) issue_fwd(1, Fb[1], bar + 1);
    }
    if (tid < n) zs[tid] = tb[tri(n) + n + tid];
    __syncthreads();
    constexpr int UP = THREADS / m, XP = THREADS / n;  // partial sums per control / per state
    constexpr int UX = n / UP, XZ = w / XP;
    static_assert(UP * m == THREADS && XP * n == THREADS && UX * UP == n && XZ * XP == w, "forward pass split");
    const int ut_ = tid % m, up_ = tid / m, xi_ = tid % n, xp_ = tid / n;
    double gk[UX];
    SM_UNROLL
    for (int i = 0; i < UX; ++i) gk[i] = gb[ut_ + m * (UX * up_ + i)];
    double kf = tid < m ? gb[m * n + tid] : 0.0;
    for (int it = 0; it < steps; ++it) {
        double *zc = zs + (it & 1) * w, *zn = zs + ((it + 1) & 1) * w;
        // u = -K x - kff : UP partial sums per control
        double acc = 0.0;
        SM_UNROLL
        for (int i = 0; i < UX; ++i) acc = fma(gk[i], zc[UX * up_ + i], acc);
        const double kfc = kf;
        if (it + 1 < steps) {
            SM_UNROLL
            for (int i = 0; i < UX; ++i) gk[i] = gb[(int64_t)(it + 1) * GR + ut_ + m * (UX * up_ + i)];
            if (tid < m) kf = gb[(int64_t)(it + 1) * GR + m * n + tid];
        }
        if (m < 32) {
            SM_UNROLL
            for (int o = m; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (lane < m) red[wp * m + lane] = acc;
        __syncthreads();
        double *zk = zb + (int64_t)it * w;
        if (tid < m) {
            double s = kfc;
            SM_UNROLL
            for (int ww = 0; ww < WARPS; ++ww) s += red[ww * m + tid];
            zc[n + tid] = -s;
            __stcs(zk + n + tid, -s);
        }
        if (tid < n) __stcs(zk + tid, zc[tid]);
        // this knot's [A B]: forward knot j completes on bar[j & 1]
        if (it & 1) {
            mbar_wait(bar + 1, ph1);
            ph1 ^= 1;
        } else {
            mbar_wait(bar, ph0);
            ph0 ^= 1;
        }
        __syncthreads();
        // x+ = A x + B u : XP partial sums per state
        const double *Fc = Fb[it & 1];
        double xa = 0.0;
        SM_UNROLL
        for (int zz = 0; zz < XZ; ++zz) {
            const int zi = XZ * xp_ + zz;
            xa = fma(Fc[zi * n + xi_], zc[zi], xa);
        }
        red[WARPS * m + xp_ * n + xi_] = xa;
        __syncthreads();
        if (tid < n) {
            double s = 0.0;
            SM_UNROLL
            for (int pp = 0; pp < XP; ++pp) s += red[WARPS * m + pp * n + tid];
            zn[tid] = s;
        }
        if (wp == 0 && it + 2 < steps) issue_fwd(it + 2, Fb[it & 1], bar + (it & 1));
        __syncthreads();
    }
    if (tid < n) __stcs(zb + (int64_t)steps * w + tid, zs[(steps & 1) * w + tid]);
}

}  // namespace rcta
