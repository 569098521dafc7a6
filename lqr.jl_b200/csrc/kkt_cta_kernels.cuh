// kkt_cta_kernels.cuh — constrained KKT solve for the large-state class (n = 64, m = 16): one CTA per
// instance, everything of order n^3 on the FP64 tensor cores (mma.sync m8n8k4 DMMA).  "Dense-Schur variant"
// of BASELINE.json config 5.
//
// Replaces _solve!(::CholeskySolver) : src/cholesky_solver.jl:166-182 for the stage pattern p = [n,0,..,0,n]
// (initial condition + dynamics + goal constraint, D2 = [-I 0], block-diagonal cost Hessian) with the same
// block-LDL' restatement as kkt_hw_kernels.cuh (see there for the algebra and the reference line numbers):
//   pre-pass (kkt_cta_prep_kernel, one CTA per knot, fully parallel — the reference's calculate_shur_factors!,
//     src/jacobian_blocks.jl:220-286):   Qi = Q_k^-1,  Ri = R_k^-1,  T = A Qi (= -F'),  G = A Qi A' + B Ri B',
//     hg = Hi g,  rho = D1 hg - d           (first knot: B0 = C Hi C', -E0' and -y0 in the same slots)
//   main kernel (kkt_cta_kernel, one CTA per instance, sequential in k — cholesky!(chol, shur) and the two
//     substitutions, src/cholesky_solve.jl:28-143):   Sigma = Cp + Qi,  Si = Sigma^-1,  Z = T Si (= -U'),
//     v = Si y,  Cp <- G - Z T',  dp <- rho + T v;   backward:  x_{k-1} = v_k + Z_k' x_k,  Lambda = -x,
//     res_k, dz_k = -Hi res_k   (calculate_primals!, src/cholesky_solver.jl:185-236).
// Both kernels invert n x n SPD blocks with an in-place block Gauss-Jordan on 8 x 8 tiles: warp w keeps row
// tile w of the matrix in DMMA accumulator registers for all n/8 block steps; per step the pivot column panel
// goes through shared memory, the owning warp inverts the 8 x 8 pivot (one lane per column), and the rank-8
// update of every other tile is two DMMAs (A operand = the warp's own accumulators, a C fragment read as an
// A fragment under the permuted contraction index; B operand = a panel tile read with one 16-byte load).
#pragma once
#include "kkt_hw_kernels.cuh"

namespace kcta {
using rdmma::bulk_g2s;
using rdmma::mbar_expect_tx;
using rdmma::mbar_init;
using rdmma::mbar_wait;
using rdmma::mma884;

template <int n, int m, int HESS = LQRB_HESS_BLOCKDIAG>
struct Lay {
    static_assert(n % 8 == 0 && m % 8 == 0 && n >= 16 && n <= 64 && m <= 16 && m <= n, "n = 16..64 (multiples of 8), m = 8, 16");
    static_assert(HESS == LQRB_HESS_BLOCKDIAG || HESS == LQRB_HESS_DIAG, "block-diagonal or diagonal cost Hessian");
    static constexpr int NT = n / 8, UT = m / 8, w = n + m, WARPS = NT, THREADS = WARPS * 32;
    static constexpr int HQ = HESS == LQRB_HESS_DIAG ? n : tri(n), HR = HESS == LQRB_HESS_DIAG ? m : tri(m);
    // packed knot records (tile width 1), identical to kkt_hw_kernels.cuh
    static constexpr int oQ = 0, oR = HQ, og = oR + HR, oD1 = og + w, od = oD1 + n * w, CORE = od + n;
    static constexpr int oC0 = CORE, FIRST = CORE + n * w + n, MID = CORE;
    static constexpr int oCl = HQ + n, LAST = oCl + n * n + n;
    // ps = stage-constraint rows of every interior knot (C (ps x w, column-major) | c (ps) follow d in its record),
    // 0 <= ps <= PSMAX; with an odd ps knot records can start on odd doubles (see the alignment notes in the kernels)
    static constexpr int PSMAX = 4;
    __host__ __device__ static constexpr int mid(int ps) { return MID + ps * (w + 1); }
    __host__ __device__ static constexpr int64_t data_rows(int N, int ps = 0) { return FIRST + (int64_t)(N - 2) * mid(ps) + LAST; }
    __host__ __device__ static constexpr int64_t knot_off(int k, int ps = 0) { return k == 0 ? 0 : FIRST + (int64_t)(k - 1) * mid(ps); }
    __host__ __device__ static constexpr int64_t mult_rows(int N, int ps = 0) { return 2 * n + (int64_t)(N - 1) * n + (int64_t)(N - 2) * ps; }
    __host__ __device__ static constexpr int64_t z_rows(int N) { return (int64_t)N * n + (int64_t)(N - 1) * m; }
    // pre-pass output slot of one knot (doubles): Qi | T | G | hg (w) | rho (n) | Ri (m*m)
    static constexpr int hQi = 0, hT = n * n, hG = 2 * n * n, hHg = 3 * n * n, hRho = hHg + w, hRi = hRho + n,
                         HS = hRi + m * m;
    static_assert(HS % 2 == 0 && hHg % 2 == 0, "16-byte pieces");
    // with stage rows the slot continues: [D_j (n) | E_j (n)] for j < 4 | B (4 x 4) | ct (4)
    //   D_j = -(Hi C_j')_x,  E_j = D1 Hi C_j',  B = C Hi C',  ct = C hg - c      (shur! for the stage rows)
    static constexpr int sB = 8 * n, sCt = sB + 16, HSS = sCt + 4;
    __host__ __device__ static constexpr int hs(int ps) { return HS + (ps > 0 ? HSS : 0); }
    // per instance: N slots + the true Qi of the first knot (its slot holds B0)
    __host__ __device__ static constexpr int64_t prep_rows(int N, int ps = 0) { return (int64_t)N * hs(ps) + n * n; }
    static constexpr int REC = n * n + n;  // Z (row-major) | v
    // with stage rows the record continues: [sd_j (n) | E'_j (n)] for j < 4 | Bi (4 x 4) | c' (4)
    __host__ __device__ static constexpr int rec(int ps) { return REC + (ps > 0 ? HSS : 0); }
    // shared memory of the pre-pass (doubles)
    static constexpr int LA = n + 4;  // k-major operand loads: leading dimension = 4 (mod 16)
    static constexpr int pA = 0, pQ = pA + w * LA, pPan = pQ + n * LA, pPi = pPan + 2 * NT * 64, pCol = pPi + 128,
                         pRi = pCol + 2 * m, pV = pRi + m * (m + 4), PREP_TOTAL = pV + 4 * w + 64,
                         pSt = PREP_TOTAL, PREP_TOTAL_ST = pSt + 8 * w;  // stage rows: C_j and Hi C_j' (4 x w each)
    // shared memory of the main kernel (doubles)
    static constexpr int LB = n + 8;  // paired (16-byte) operand loads: leading dimension = 8 (mod 16)
    static constexpr int mT = 0, mS = mT + n * LB, mPan = mS + n * LB, mPi = mPan + 2 * NT * 64, mCol = mPi + 128,
                         mV = mCol + 16, mBar = mV + 8 * n + 4 * w, MAIN_TOTAL = mBar + 2,
                         mSt = MAIN_TOTAL, MAIN_TOTAL_ST = mSt + 16 * n + 48;  // stage rows: D, sd, E', W (4 x n each) | B, Bi | ct, c', bc, xi
};

// ------------------------------------------------------------------ 8 x 8 pivot block ------------------
// Inverse of an SPD 8 x 8 tile by one warp, Gauss-Jordan without pivoting.  Lane (h = lane >> 3, c = lane & 7)
// holds rows 2h, 2h+1 of column c; the pivot column, the pivot and the lane's own row-kk entry come by
// shuffle, so a pivot step costs 7 FP64 instructions (they queue behind the other CTA's DMMAs) and no
// shared-memory round trip.  Returns the 1-based index of the first non-positive pivot or 0.
// lo / hi: running min / max of the high words of the (positive) pivots — integer compares order positive doubles, so
// (hi - lo) >> 20 is log2 of the pivot ratio, the conditioning estimate of the block (integer pipe, off the FP64 chain).
__device__ __forceinline__ int gj8_warp(double &a0, double &a1, int lane, int &lo, int &hi) {
    const int h = lane >> 3, c = lane & 7;
    int bad = 0;
    SM_UNROLL
    for (int kk = 0; kk < 8; ++kk) {
        const double mine = (kk & 1) ? a1 : a0;  // row kk lives in register kk & 1 of the lanes with h = kk / 2
        const double ck0 = __shfl_sync(0xffffffffu, a0, (h << 3) | kk);
        const double ck1 = __shfl_sync(0xffffffffu, a1, (h << 3) | kk);
        const double piv = __shfl_sync(0xffffffffu, mine, ((kk >> 1) << 3) | kk);
        const double akk = __shfl_sync(0xffffffffu, mine, ((kk >> 1) << 3) | c);
        if (!(piv > 0.0) && bad == 0) bad = kk + 1;
        lo = min(lo, __double2hiint(piv));
        hi = max(hi, __double2hiint(piv));
        const double p = rdmma::fast_rcp3(piv);
        const bool pc = c == kk;
        const double f = pc ? -p : akk * p;
        const double u0 = pc ? ck0 * f : fma(-ck0, f, a0);
        const double u1 = pc ? ck1 * f : fma(-ck1, f, a1);
        const bool prow = h == (kk >> 1);
        a0 = (prow && !(kk & 1)) ? (pc ? p : f) : u0;
        a1 = (prow && (kk & 1)) ? (pc ? p : f) : u1;
    }
    return bad;
}

// log2 of the pivot ratio from the min / max pivot high words (0 when a pivot was not positive: info says so)
__device__ __forceinline__ int pivot_bits(int lo, int hi) { return (lo > 0 && hi >= lo) ? (hi - lo) >> 20 : 0; }

// ------------------------------------------------------------------ block Gauss-Jordan ----------------
// In place: S[ct][e] = tile (wp, ct) of an SPD n x n matrix in C-fragment layout -> the same tiles of its
// inverse.  pan: 2 x NT*64 doubles (8 x 8 tiles, row-major, double-buffered column panels), pis: 2 x 64 doubles
// (double-buffered pivot inverses).  Returns the 1-based index of the first non-positive pivot (potrf
// semantics) or 0; identical in every thread.
//
// One CTA barrier per block step: the warp that owns the NEXT pivot updates that tile first, inverts it with
// shuffles (gj8_warp) while the other warps are still in their rank-8 updates, and every warp publishes its
// tile of the next column panel before the barrier (look-ahead).  In-place Gauss-Jordan leaves
// A_kj = A_jk' for the columns still to be eliminated (j > kb) and A_kj = -A_jk' for the ones already done
// (A_ik <- -A_ik Pi but A_kj <- +Pi A_kj), so the pivot row is taken from the column panel with that sign.
// flag[1], flag[2]: min / max pivot high word over the whole matrix (see gj8_warp), set by the caller to INT_MAX / 0.
template <int NT>
__device__ __forceinline__ int block_gj_inverse(double (&S)[NT][2], double *pan, double *pis, double *colb,
                                                int *flag, int wp, int lane) {
    (void)colb;
    const int g = lane >> 2, q = lane & 3;
    const int hh = lane >> 3, cc = lane & 7;
    // invert the tile this warp has just written to p64 (row-major) and leave the inverse in po
    auto pivot = [&](const double *p64, double *po, int kb) {
        __syncwarp();
        double a0 = p64[(2 * hh) * 8 + cc], a1 = p64[(2 * hh + 1) * 8 + cc];
        int lo = 0x7fffffff, hi = 0;
        const int bad = gj8_warp(a0, a1, lane, lo, hi);
        po[(2 * hh) * 8 + cc] = a0;
        po[(2 * hh + 1) * 8 + cc] = a1;
        if (lane == 0) {
            if (bad != 0 && *flag == 0) *flag = 8 * kb + bad;
            atomicMin(flag + 1, lo);
            atomicMax(flag + 2, hi);
        }
    };
    // prologue: column panel 0 and its pivot
    *reinterpret_cast<double2 *>(pan + wp * 64 + g * 8 + 2 * q) = make_double2(S[0][0], S[0][1]);
    if (wp == 0) pivot(pan, pis, 0);
    __syncthreads();
    SM_UNROLL
    for (int kb = 0; kb < NT; ++kb) {
        const double *pc = pan + (kb & 1) * NT * 64;
        double *pn = pan + ((kb + 1) & 1) * NT * 64;
        const double2 pv = *reinterpret_cast<const double2 *>(pis + (kb & 1) * 64 + g * 8 + 2 * q);  // Pi[g][2q..2q+1]
        if (wp != kb) {
            // T = A_wk Pi ; A_wj -= T A_kj ; A_wk = -T
            double t0 = 0.0, t1 = 0.0;
            mma884(t0, t1, S[kb][0], pv.x);
            mma884(t0, t1, S[kb][1], pv.y);
            const double n0 = -t0, n1 = -t1;
            auto update = [&](int j) {
                const double2 b = *reinterpret_cast<const double2 *>(pc + j * 64 + g * 8 + 2 * q);
                mma884(S[j][0], S[j][1], j < kb ? t0 : n0, b.x);
                mma884(S[j][0], S[j][1], j < kb ? t1 : n1, b.y);
            };
            if (kb + 1 < NT) {
                update(kb + 1);  // the next panel tile first
                *reinterpret_cast<double2 *>(pn + wp * 64 + g * 8 + 2 * q) = make_double2(S[kb + 1][0], S[kb + 1][1]);
                if (wp == kb + 1) pivot(pn + wp * 64, pis + ((kb + 1) & 1) * 64, kb + 1);
            }
            // The warp that owns the next pivot row stops here: in the next step its whole strip is rebuilt from
            // the column panel (A_kj = +-Pi A_jk'), so its other tiles are dead — and it is the critical path.
            if (wp != kb + 1 || kb + 1 >= NT) {
                SM_UNROLL
                for (int j = 0; j < NT; ++j)
                    if (j != kb && j != kb + 1) update(j);
                S[kb][0] = n0;
                S[kb][1] = n1;
            }
        } else {
            // the pivot row: A_kj = Pi A_kj, A_kk = Pi
            SM_UNROLL
            for (int j = 0; j < NT; ++j) {
                if (j != kb) {
                    const double2 b = *reinterpret_cast<const double2 *>(pc + j * 64 + g * 8 + 2 * q);
                    double s0 = 0.0, s1 = 0.0;
                    mma884(s0, s1, pv.x, b.x);
                    mma884(s0, s1, pv.y, b.y);
                    S[j][0] = j < kb ? -s0 : s0;
                    S[j][1] = j < kb ? -s1 : s1;
                    if (j == kb + 1)
                        *reinterpret_cast<double2 *>(pn + wp * 64 + g * 8 + 2 * q) = make_double2(S[j][0], S[j][1]);
                }
            }
            S[kb][0] = pv.x;
            S[kb][1] = pv.y;
        }
        __syncthreads();
    }
    return *flag;
}

// ------------------------------------------------------------------ R^-1 for every knot ---------------
// One half-warp per knot with controls, lane j holds column j (Gauss-Jordan through shared memory, m serial
// pivots).  Inside the pre-pass this chain kept seven of the eight warps of the CTA at a barrier for a fifth of
// its time; here it is a fully parallel launch and the pre-pass reads the m x m result from its slot.
template <int n, int m, int HESS>
__global__ void __launch_bounds__(128)
    kkt_cta_ri_kernel(const double *__restrict__ data, double *__restrict__ prep, int32_t *__restrict__ hinfo, int N,
                      int64_t batch, int soc, int ps) {
    using L = Lay<n, m, HESS>;
    static_assert(m <= 16, "one half-warp per R");
    __shared__ __align__(16) double colb_all[8][2 * m];
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31, hh = lane >> 4, hl = lane & 15;
    const int64_t total = batch * (N - 1);
    const int64_t pair = ((int64_t)blockIdx.x * 4 + wp) * 2;
    if (pair >= total) return;  // whole warp leaves
    const bool active = pair + hh < total;
    const int64_t idx = active ? pair + hh : total - 1;
    const int64_t inst = idx / (N - 1);
    const int k = (int)(idx % (N - 1));
    const double *kp = data + inst * L::data_rows(N, ps) + L::knot_off(k, ps);
    double *out = prep + inst * L::prep_rows(N, ps) + (int64_t)k * L::hs(ps) + L::hRi;
    double a[m];
    const int j = hl < m ? hl : 0;
    SM_UNROLL
    for (int i = 0; i < m; ++i) {
        if (soc) a[i] = i == j ? 1.0 : 0.0;  // second_order_correction!: H = I
        else if (HESS == LQRB_HESS_DIAG) a[i] = i == j ? kp[L::oR + j] : 0.0;
        else a[i] = kp[L::oR + (i <= j ? j * (j + 1) / 2 + i : i * (i + 1) / 2 + j)];
    }
    const int bad = khw::gj_inverse<m>(a, colb_all[2 * wp + hh], hl < m ? hl : 31);
    if (active && hl < m) {
        SM_UNROLL
        for (int i = 0; i < m; ++i) out[i * m + hl] = a[i];
    }
    if (active && bad != 0 && hl == 0) atomicMin(hinfo + inst, (k + 1) * 1000 + n + bad);
}

// ------------------------------------------------------------------ pre-pass --------------------------
// grid = batch * N CTAs of THREADS threads.  knot 0: generic scalar code for the C_1 blocks (once per instance).
// ST = false: no stage rows (ps is ignored; the instantiation config 5b-K runs); true: ps <= 4 rows on every interior knot
template <int n, int m, int HESS, bool ST>
__global__ void __launch_bounds__(Lay<n, m, HESS>::THREADS, 2)
    kkt_cta_prep_kernel(const double *__restrict__ data, double *__restrict__ prep, int32_t *__restrict__ hinfo,
                        int32_t *__restrict__ cinfo, int N, int64_t batch, int soc, int ps_arg) {
    using L = Lay<n, m, HESS>;
    const int ps = ST ? ps_arg : 0;
    constexpr int NT = L::NT, UT = L::UT, w = L::w, LA = L::LA, THREADS = L::THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *As = sm + L::pA, *Qs = sm + L::pQ, *pan = sm + L::pPan, *pis = sm + L::pPi, *colb = sm + L::pCol,
           *Ris = sm + L::pRi, *vq = sm + L::pV, *vhg = vq + w, *vd = vhg + w, *vtmp = vd + w;
    __shared__ int flag3[3];
    int &flag = flag3[0];
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int64_t inst = blockIdx.x / N;
    const int k = (int)(blockIdx.x % N);
    const bool first = k == 0, last = k == N - 1;
    const int mk = last ? 0 : m, wk = n + mk;
    const double *kp = data + inst * L::data_rows(N, ps) + L::knot_off(k, ps);
    double *out = prep + inst * L::prep_rows(N, ps) + (int64_t)k * L::hs(ps);
    constexpr int LR = m + 4;
    __shared__ __align__(8) uint64_t xbar;
    if (tid == 0) {
        flag = 0;
        flag3[1] = 0x7fffffff;
        flag3[2] = 0;
        mbar_init(&xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- stage X = A_k (or C_N), B_k column-major with padded leading dimension: one bulk copy per column,
    // in flight while Q is inverted (plain loads here put the whole HBM latency in front of the first barrier);
    // g, d wait in registers
    const double *Xg = last ? kp + L::oCl : kp + L::oD1;
    // (an odd number of stage rows puts every other knot record on an odd double: bulk copies need 16-byte aligned
    // sources, so those knots use plain loads, published by the barriers below)
    const bool xal = !ST || (reinterpret_cast<uintptr_t>(Xg) & 15) == 0;
    if (xal) {
        if (wp == 0) {
            if (lane == 0) mbar_expect_tx(&xbar, (uint32_t)(n * wk * 8));
            __syncwarp();
            for (int j = lane; j < wk; j += 32) bulk_g2s(As + j * LA, Xg + (int64_t)j * n, n * 8, &xbar);
        }
    } else {
        for (int e = tid; e < n * wk; e += THREADS) As[(e / n) * LA + (e % n)] = Xg[e];
    }
    static_assert(THREADS >= w, "one thread per entry of g");
    const double gq_reg = (tid < wk && !soc) ? kp[(last ? L::HQ : L::og) + tid] : 0.0;  // SOC: g = 0
    const double vd_reg = tid < n ? (last ? kp[L::oCl + n * n + tid] : kp[L::od + tid]) : 0.0;
    constexpr int RI_PER = (m * m + THREADS - 1) / THREADS;  // Ri (kkt_cta_ri_kernel) also waits in registers
    double ri_reg[RI_PER];
    SM_UNROLL
    for (int e = 0; e < RI_PER; ++e) ri_reg[e] = (!last && tid + e * THREADS < m * m) ? out[L::hRi + tid + e * THREADS] : 0.0;
    // ---- Q strip (C fragments) -> Qi
    double S[NT][2];
    SM_UNROLL
    for (int ct = 0; ct < NT; ++ct)
        SM_UNROLL
        for (int e = 0; e < 2; ++e) {
            const int r = 8 * wp + g, c = 8 * ct + 2 * q + e;
            if (soc) S[ct][e] = r == c ? 1.0 : 0.0;  // second_order_correction!: H = I
            else if (HESS == LQRB_HESS_DIAG) S[ct][e] = r == c ? kp[r] : 0.0;
            else S[ct][e] = kp[r <= c ? c * (c + 1) / 2 + r : r * (r + 1) / 2 + c];
        }
    __syncthreads();
    int bad = block_gj_inverse<NT>(S, pan, pis, colb, &flag, wp, lane);
    if (tid < wk) vq[tid] = gq_reg;
    if (tid < n) vd[tid] = vd_reg;
    if (xal) mbar_wait(&xbar, 0);  // X has landed (every thread observes the barrier: the bulk writes are then visible)
    // Qi -> shared (operand) and global (slot, or the extra block for the first knot)
    {
        double *qo = first ? prep + inst * L::prep_rows(N, ps) + (int64_t)N * L::hs(ps) : out + L::hQi;
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct) {
            const int r = 8 * wp + g, c = 8 * ct + 2 * q;
            *reinterpret_cast<double2 *>(Qs + r * LA + c) = make_double2(S[ct][0], S[ct][1]);
            *reinterpret_cast<double2 *>(qo + r * n + c) = make_double2(S[ct][0], S[ct][1]);
        }
    }
    // ---- Ri from its slot (kkt_cta_ri_kernel)
    SM_UNROLL
    for (int e = 0; e < RI_PER; ++e) {
        const int t = tid + e * THREADS;
        if (!last && t < m * m) Ris[(t / m) * LR + (t % m)] = ri_reg[e];
    }
    __syncthreads();
    bad = flag;
    if (bad != 0 && tid == 0) atomicMin(hinfo + inst, (k + 1) * 1000 + bad);
    if (tid == 0) atomicMax(cinfo + inst, pivot_bits(flag3[1], flag3[2]));  // conditioning estimate of Q_k
    // ---- hg = Hi g
    if (tid < n) {
        double s0 = 0.0, s1 = 0.0;
        SM_UNROLL
        for (int l = 0; l < n; l += 2) {
            s0 = fma(Qs[l * LA + tid], vq[l], s0);
            s1 = fma(Qs[(l + 1) * LA + tid], vq[l + 1], s1);
        }
        vhg[tid] = s0 + s1;
    } else if (tid < wk) {
        double s = 0.0;
        for (int l = 0; l < m; ++l) s = fma(Ris[l * LR + tid - n], vq[n + l], s);
        vhg[tid] = s;
    }
    __syncthreads();
    if (!first) {
        for (int e = tid; e < wk; e += THREADS) out[L::hHg + e] = vhg[e];
    }
    // ---- rho = X hg_x + B hg_u - d
    if (tid < n) {
        double s0 = -vd[tid], s1 = 0.0;
        SM_UNROLL
        for (int l = 0; l < n; l += 2) {
            s0 = fma(As[l * LA + tid], vhg[l], s0);
            s1 = fma(As[(l + 1) * LA + tid], vhg[l + 1], s1);
        }
        for (int l = 0; l < mk; ++l) s0 = fma(As[(n + l) * LA + tid], vhg[n + l], s0);
        out[L::hRho + tid] = s0 + s1;
    }
    // ---- stage rows of an interior knot (shur! for the ps rows, src/jacobian_blocks.jl:231-286), as vectors:
    //   tc_j = Hi C_j',  D_j = -(tc_j)_x,  E_j = [A B] tc_j,  B = C tc,  ct = C hg - c
    if (ST && ps > 0 && !first && !last) {
        double *vc = sm + L::pSt, *tcs = vc + 4 * w;
        const double *Cg = kp + L::CORE;  // ps x w column-major, then c (ps)
        for (int e = tid; e < ps * w; e += THREADS) vc[(e % ps) * w + e / ps] = Cg[e];
        __syncthreads();
        for (int j = 0; j < ps; ++j) {
            if (tid < n) {
                double s0 = 0.0, s1 = 0.0;
                SM_UNROLL
                for (int l = 0; l < n; l += 2) {
                    s0 = fma(Qs[l * LA + tid], vc[j * w + l], s0);
                    s1 = fma(Qs[(l + 1) * LA + tid], vc[j * w + l + 1], s1);
                }
                tcs[j * w + tid] = s0 + s1;
            } else if (tid < w) {
                double s0 = 0.0;
                for (int l = 0; l < m; ++l) s0 = fma(Ris[l * LR + tid - n], vc[j * w + n + l], s0);
                tcs[j * w + tid] = s0;
            }
        }
        __syncthreads();
        double *so = out + L::HS;
        if (tid < n) {
            for (int j = 0; j < ps; ++j) {
                double s0 = 0.0, s1 = 0.0;
                SM_UNROLL
                for (int l = 0; l < n; l += 2) {
                    s0 = fma(As[l * LA + tid], tcs[j * w + l], s0);
                    s1 = fma(As[(l + 1) * LA + tid], tcs[j * w + l + 1], s1);
                }
                for (int l = 0; l < m; ++l) s0 = fma(As[(n + l) * LA + tid], tcs[j * w + n + l], s0);
                so[j * 2 * n + tid] = -tcs[j * w + tid];
                so[j * 2 * n + n + tid] = s0 + s1;
            }
        }
        if (tid < ps * ps) {
            const int j = tid / ps, jp = tid % ps;
            double s0 = 0.0;
            for (int l = 0; l < w; ++l) s0 = fma(vc[j * w + l], tcs[jp * w + l], s0);
            so[L::sB + 4 * j + jp] = s0;
        } else if (tid >= 32 && tid < 32 + ps) {
            const int j = tid - 32;
            double s0 = -Cg[ps * w + j];
            for (int l = 0; l < w; ++l) s0 = fma(vc[j * w + l], vhg[l], s0);
            so[L::sCt + j] = s0;
        }
    }
    // ---- T strip = X[rows] Qi   (A operand: X[8wp+g][4s+q], B operand: Qi[4s+q][8ct+g], k-major loads)
    double T[NT][2];
    SM_UNROLL
    for (int ct = 0; ct < NT; ++ct) T[ct][0] = T[ct][1] = 0.0;
    {
        const double *xa = As + q * LA + 8 * wp + g;
        const double *qb = Qs + q * LA + g;
#pragma unroll 4
        for (int s = 0; s < n / 4; ++s) {
            const double a = xa[4 * s * LA];
            SM_UNROLL
            for (int ct = 0; ct < NT; ++ct) mma884(T[ct][0], T[ct][1], a, qb[4 * s * LA + 8 * ct]);
        }
    }
    // ---- G strip = T X' + (B Ri) B'   (A operand: the T accumulators; B operand: X[8ct+g][8cp+2q+e])
    double G[NT][2];
    SM_UNROLL
    for (int ct = 0; ct < NT; ++ct) G[ct][0] = G[ct][1] = 0.0;
    SM_UNROLL
    for (int cp = 0; cp < NT; ++cp) {
        const double *xb = As + (8 * cp + 2 * q) * LA + g;
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct) {
            mma884(G[ct][0], G[ct][1], T[cp][0], xb[8 * ct]);
            mma884(G[ct][0], G[ct][1], T[cp][1], xb[LA + 8 * ct]);
        }
    }
    if (!last) {
        double V[UT][2];
        SM_UNROLL
        for (int ut = 0; ut < UT; ++ut) V[ut][0] = V[ut][1] = 0.0;
        SM_UNROLL
        for (int s = 0; s < m / 4; ++s) {
            const double a = As[(n + 4 * s + q) * LA + 8 * wp + g];
            SM_UNROLL
            for (int ut = 0; ut < UT; ++ut) mma884(V[ut][0], V[ut][1], a, Ris[(4 * s + q) * LR + 8 * ut + g]);
        }
        SM_UNROLL
        for (int ut = 0; ut < UT; ++ut) {
            const double *bb = As + (n + 8 * ut + 2 * q) * LA + g;
            SM_UNROLL
            for (int ct = 0; ct < NT; ++ct) {
                mma884(G[ct][0], G[ct][1], V[ut][0], bb[8 * ct]);
                mma884(G[ct][0], G[ct][1], V[ut][1], bb[LA + 8 * ct]);
            }
        }
    }
    SM_UNROLL
    for (int ct = 0; ct < NT; ++ct) {
        const int r = 8 * wp + g, c = 8 * ct + 2 * q;
        *reinterpret_cast<double2 *>(out + L::hG + r * n + c) = make_double2(G[ct][0], G[ct][1]);
        if (!first) *reinterpret_cast<double2 *>(out + L::hT + r * n + c) = make_double2(T[ct][0], T[ct][1]);
    }
    if (!first) return;

    // ---- first knot: slots Qi := B0 = C Hi C', T := -E0' = -D1 Hi C', hg := -(C hg - c)   (generic code)
    // W0 = Hi C' (w x n) in shared memory over Qs|pan.. is too large; stream it: for each row i of C
    const double *C0 = kp + L::oC0;  // n x w column-major
    double *qo = out + L::hQi, *to = out + L::hT;
    __syncthreads();
    // vtmp reuse: per-CTA row buffer of Hi C[i,:]'  (w doubles)
    for (int i = 0; i < n; ++i) {
        // hc = Hi C[i,:]'
        if (tid < n) {
            double s = 0.0;
            for (int l = 0; l < n; ++l) s = fma(Qs[l * LA + tid], C0[i + n * l], s);
            vtmp[tid] = s;
        } else if (tid < w) {
            double s = 0.0;
            for (int l = 0; l < m; ++l) s = fma(Ris[l * LR + tid - n], C0[i + n * (n + l)], s);
            vtmp[tid] = s;
        }
        __syncthreads();
        // B0[:, i] = C hc ; (-E0')[:, i] ... E0[i][b] = sum_j hc[j] D1[b][j]  ->  T slot row b, column i = -E0[i][b]
        if (tid < n) {
            double s = 0.0;
            for (int j = 0; j < w; ++j) s = fma(C0[tid + n * j], vtmp[j], s);
            qo[tid * n + i] = s;  // B0[tid][i] (symmetric)
        } else if (tid < 2 * n) {
            const int b = tid - n;
            double s = 0.0;
            for (int j = 0; j < w; ++j) s = fma(As[j * LA + b], vtmp[j], s);
            to[b * n + i] = -s;
        }
        __syncthreads();
    }
    if (tid < n) {
        double s = -C0[n * w + tid];
        for (int j = 0; j < w; ++j) s = fma(C0[tid + n * j], vhg[j], s);
        out[L::hHg + tid] = -s;  // y0 = dp - slot with dp = 0
    } else if (tid < w) {
        out[L::hHg + tid] = vhg[tid];
    }
}

// ------------------------------------------------------------------ main kernel -----------------------
template <int n, int m, int HESS, bool ST>
__global__ void __launch_bounds__(Lay<n, m, HESS>::THREADS, 2)
    kkt_cta_kernel(const double *__restrict__ data, const double *__restrict__ prep, const int32_t *__restrict__ hinfo,
                   double *__restrict__ recs, double *__restrict__ dz, double *__restrict__ mult,
                   double *__restrict__ res, int32_t *__restrict__ info, int32_t *__restrict__ cinfo, int N,
                   int64_t batch, int soc, int ps_arg, int free_final) {
    using L = Lay<n, m, HESS>;
    constexpr int NT = L::NT, w = L::w, LB = L::LB, THREADS = L::THREADS;
    static_assert(THREADS == 4 * n, "four partial sums per row in the mat-vecs");
    const int ps = ST ? ps_arg : 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    double *Ts = sm + L::mT, *Ss = sm + L::mS, *pan = sm + L::mPan, *pis = sm + L::mPi, *colb = sm + L::mCol,
           *ys = sm + L::mV, *vs = ys + n, *dps = vs + n, *red = dps + n /* 4n */, *xs = red + 4 * n, *rsv = xs + w,
           *xps = rsv + w;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + L::mBar);
    // stage rows (allocated only when ps > 0): D_j, sd_j, E'_j, W_j (4 x n each) | B (16) | Bi (16) | ct, c', bc, xi (4 each)
    double *Dv = sm + L::mSt, *sdv = Dv + 4 * n, *Ev = sdv + 4 * n, *Wv = Ev + 4 * n, *Bm = Wv + 4 * n, *Bi = Bm + 16,
           *ctv = Bi + 16, *cpv = ctv + 4, *bcv = cpv + 4, *xiv = bcv + 4;
    __shared__ int flag3[3];
    int &flag = flag3[0];
    const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int64_t inst = blockIdx.x;
    const int hs = L::hs(ps), recw = L::rec(ps);
    const double *db = data + inst * L::data_rows(N, ps);
    const double *pb = prep + inst * L::prep_rows(N, ps);
    double *rb = recs + inst * (int64_t)N * recw;
    double *zb = dz + inst * L::z_rows(N);
    double *mb = mult + inst * L::mult_rows(N, ps);
    double *resb = res ? res + inst * L::z_rows(N) : nullptr;
    const int r0 = 8 * wp + g;  // the row of this lane's accumulator entries

    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        flag = 0;
    }
    int spread = 0;  // thread 0: largest log2 pivot ratio of any Sigma_k (conditioning estimate, see gj8_warp)
    if (tid < n) dps[tid] = 0.0;
    __syncthreads();
    uint32_t ph = 0;
    // T_k (n x n row-major in the slot) -> Ts with padded rows, one 512-byte bulk copy per row
    auto issue_T = [&](int k) {
        if (wp == 0) {
            const double *src = pb + (int64_t)k * hs + L::hT;
            if (lane == 0) mbar_expect_tx(bar, (uint32_t)(n * n * 8));
            __syncwarp();
            for (int r = lane; r < n; r += 32) bulk_g2s(Ts + r * LB, src + r * n, n * 8, bar);
        }
    };
    issue_T(0);

    int st_all = 0;
    double Cp[NT][2];
    SM_UNROLL
    for (int ct = 0; ct < NT; ++ct) Cp[ct][0] = Cp[ct][1] = 0.0;

    // one elimination step: (Cp, dps) of the previous row + slot k  ->  record k, new (Cp, dps)
    for (int k = 0; k < N; ++k) {
        const double *slot = pb + (int64_t)k * hs;
        const int psk = (ST && k > 0 && k < N - 1) ? ps : 0;  // stage rows of this knot
        // Sigma strip = Cp + Qi (slot), G strip (prefetched)
        double S[NT][2], G[NT][2];
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct) {
            const double2 qv = *reinterpret_cast<const double2 *>(slot + L::hQi + r0 * n + 8 * ct + 2 * q);
            const double2 gv = *reinterpret_cast<const double2 *>(slot + L::hG + r0 * n + 8 * ct + 2 * q);
            S[ct][0] = Cp[ct][0] + qv.x;
            S[ct][1] = Cp[ct][1] + qv.y;
            G[ct][0] = gv.x;
            G[ct][1] = gv.y;
        }
        if (tid < n) ys[tid] = dps[tid] - slot[L::hHg + tid];  // y = dp - hg_x  (d += next.r_[1])
        const double rho = tid < n ? slot[L::hRho + tid] : 0.0;
        if (tid == 0) {
            flag = 0;
            flag3[1] = 0x7fffffff;
            flag3[2] = 0;
        }
        __syncthreads();
        {
            const int bad = block_gj_inverse<NT>(S, pan, pis, colb, &flag, wp, lane);
            if (bad != 0 && st_all == 0) st_all = k == 0 ? 1000 + 100 + bad : k * 1000 + 200 + bad;
            if (tid == 0) spread = max(spread, pivot_bits(flag3[1], flag3[2]));
        }
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct)
            *reinterpret_cast<double2 *>(Ss + r0 * LB + 8 * ct + 2 * q) = make_double2(S[ct][0], S[ct][1]);
        __syncthreads();
        // v = Si y : 4 partial sums per row (Si symmetric: read down the column)
        {
            const int i = tid % n, part = tid / n;  // THREADS = 4 n
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n / 4; ++l) s = fma(Ss[(part * (n / 4) + l) * LB + i], ys[part * (n / 4) + l], s);
            red[part * n + i] = s;
        }
        mbar_wait(bar, ph);  // T_k has landed
        ph ^= 1;
        __syncthreads();
        if (tid < n) {
            const double v = (red[tid] + red[n + tid]) + (red[2 * n + tid] + red[3 * n + tid]);
            vs[tid] = v;
            rb[(int64_t)k * recw + n * n + tid] = v;
        }
        // Z strip = T[rows] Si   (A operand: T[8wp+g][8cp+2q+e], B operand: Si[8cp+2q+e][8ct+g] = Si[8ct+g][..])
        double Z[NT][2];
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct) Z[ct][0] = Z[ct][1] = 0.0;
        {
            const double *ta = Ts + r0 * LB + 2 * q;
            const double *sb = Ss + g * LB + 2 * q;
            SM_UNROLL
            for (int cp = 0; cp < NT; ++cp) {
                const double2 a = *reinterpret_cast<const double2 *>(ta + 8 * cp);
                SM_UNROLL
                for (int ct = 0; ct < NT; ++ct) {
                    const double2 b = *reinterpret_cast<const double2 *>(sb + 8 * ct * LB + 8 * cp);
                    mma884(Z[ct][0], Z[ct][1], a.x, b.x);
                    mma884(Z[ct][0], Z[ct][1], a.y, b.y);
                }
            }
        }
        {
            double *rk = rb + (int64_t)k * recw;
            SM_UNROLL
            for (int ct = 0; ct < NT; ++ct)
                *reinterpret_cast<double2 *>(rk + r0 * n + 8 * ct + 2 * q) = make_double2(Z[ct][0], Z[ct][1]);
        }
        __syncthreads();  // vs is published
        // Cp' strip = G - Z T'   (A operand: the Z accumulators, B operand: T[8ct+g][8cp+2q+e])
        {
            const double *tb = Ts + g * LB + 2 * q;
            SM_UNROLL
            for (int cp = 0; cp < NT; ++cp) {
                const double na0 = -Z[cp][0], na1 = -Z[cp][1];
                SM_UNROLL
                for (int ct = 0; ct < NT; ++ct) {
                    const double2 b = *reinterpret_cast<const double2 *>(tb + 8 * ct * LB + 8 * cp);
                    mma884(G[ct][0], G[ct][1], na0, b.x);
                    mma884(G[ct][0], G[ct][1], na1, b.y);
                }
            }
        }
        // dp' = rho + T v : warp wp sums rows 8wp..8wp+7 (lanes over the columns, shuffle reduction)
        {
            constexpr int NC = (n + 31) / 32;  // column chunks of a row per lane
            double vv[NC];
            SM_UNROLL
            for (int c = 0; c < NC; ++c) vv[c] = lane + 32 * c < n ? vs[lane + 32 * c] : 0.0;
            double mine = 0.0;
            SM_UNROLL
            for (int rr = 0; rr < 8; ++rr) {
                const double *row = Ts + (8 * wp + rr) * LB;
                double s = 0.0;
                SM_UNROLL
                for (int c = 0; c < NC; ++c)
                    if (lane + 32 * c < n) s = fma(row[lane + 32 * c], vv[c], s);
                SM_UNROLL
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == rr) mine = s;
            }
            if (lane < 8) red[8 * wp + lane] = mine;
        }
        if (psk > 0) {
            // Stage rows of knot k as vectors beside the tile algebra (same elimination as kkt_wp_kernels.cuh).
            // Eliminate lam_{k-1}:  sd_j = Si D_j,  B' = B - D' sd,  E'_j = E_j + T sd_j,  c'_j = ct_j - sd_j' y;
            // eliminate mu_k:  Bi = B'^-1,  W = Bi E',  Cp -= sym(E' W'),  dp -= E' Bi c'   (pan is free here)
            const double *so = slot + L::HS;
            for (int e = tid; e < psk * 2 * n; e += THREADS) {
                const int j = e / (2 * n), r = e % (2 * n);
                if (r < n) Dv[j * n + r] = so[e];
                else Ev[j * n + r - n] = so[e];
            }
            if (tid < 16) Bm[tid] = so[L::sB + tid];
            else if (tid < 20) ctv[tid - 16] = so[L::sCt + tid - 16];
            __syncthreads();
            {
                const int i = tid % n, part = tid / n;
                double sj[L::PSMAX];
                SM_UNROLL
                for (int j = 0; j < L::PSMAX; ++j) sj[j] = 0.0;
                SM_UNROLL
                for (int l = 0; l < n / 4; ++l) {
                    const int r = part * (n / 4) + l;
                    const double sv = Ss[r * LB + i];
                    SM_UNROLL
                    for (int j = 0; j < L::PSMAX; ++j)
                        if (j < psk) sj[j] = fma(sv, Dv[j * n + r], sj[j]);
                }
                SM_UNROLL
                for (int j = 0; j < L::PSMAX; ++j)
                    if (j < psk) pan[j * 4 * n + part * n + i] = sj[j];
            }
            __syncthreads();
            if (tid < n)
                for (int j = 0; j < psk; ++j) {
                    const double *pj = pan + j * 4 * n;
                    sdv[j * n + tid] = (pj[tid] + pj[n + tid]) + (pj[2 * n + tid] + pj[3 * n + tid]);
                }
            __syncthreads();
            {
                constexpr int NC = (n + 31) / 32;
                for (int j = 0; j < psk; ++j) {
                    double vv[NC];
                    SM_UNROLL
                    for (int c = 0; c < NC; ++c) vv[c] = lane + 32 * c < n ? sdv[j * n + lane + 32 * c] : 0.0;
                    double mine = 0.0;
                    SM_UNROLL
                    for (int rr = 0; rr < 8; ++rr) {
                        const double *row = Ts + (8 * wp + rr) * LB;
                        double t = 0.0;
                        SM_UNROLL
                        for (int c = 0; c < NC; ++c)
                            if (lane + 32 * c < n) t = fma(row[lane + 32 * c], vv[c], t);
                        SM_UNROLL
                        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                        if (lane == rr) mine = t;
                    }
                    if (lane < 8) Ev[j * n + 8 * wp + lane] += mine;
                }
            }
            if (tid < psk * psk) {
                const int j = tid / psk, jp = tid % psk;
                double t = Bm[4 * j + jp];
                for (int l = 0; l < n; ++l) t = fma(-Dv[j * n + l], sdv[jp * n + l], t);
                Bm[4 * j + jp] = t;
            } else if (tid >= 32 && tid < 32 + psk) {
                const int j = tid - 32;
                double t = ctv[j];
                for (int l = 0; l < n; ++l) t = fma(-sdv[j * n + l], ys[l], t);
                cpv[j] = t;
            }
            __syncthreads();
            if (tid == 0) {  // Bi = B'^-1: Gauss-Jordan on at most 4 x 4, potrf sign test on the pivots
                double bm[4][4];
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) bm[a][b] = (a < psk && b < psk) ? Bm[4 * a + b] : (a == b ? 1.0 : 0.0);
                int badp = 0;
                for (int kk = 0; kk < 4; ++kk) {
                    if (kk >= psk) break;
                    const double piv = bm[kk][kk];
                    if (!(piv > 0.0) && badp == 0) badp = kk + 1;
                    const double pinv = 1.0 / piv;
                    for (int b = 0; b < 4; ++b) bm[kk][b] = b == kk ? pinv : bm[kk][b] * pinv;
                    for (int a = 0; a < 4; ++a) {
                        if (a == kk) continue;
                        const double f = bm[a][kk];
                        for (int b = 0; b < 4; ++b) bm[a][b] = b == kk ? -f * pinv : fma(-f, bm[kk][b], bm[a][b]);
                    }
                }
                for (int a = 0; a < 4; ++a)
                    for (int b = 0; b < 4; ++b) Bi[4 * a + b] = bm[a][b];
                xiv[0] = (double)badp;
            }
            __syncthreads();
            {
                const int badp = (int)xiv[0];
                if (badp != 0 && st_all == 0) st_all = (k + 1) * 1000 + 100 + badp;
            }
            if (tid < n) {
                for (int j = 0; j < psk; ++j) {
                    double t = 0.0;
                    for (int jp = 0; jp < psk; ++jp) t = fma(Bi[4 * j + jp], Ev[jp * n + tid], t);
                    Wv[j * n + tid] = t;
                }
            } else if (tid < n + psk) {
                const int j = tid - n;
                double t = 0.0;
                for (int jp = 0; jp < psk; ++jp) t = fma(Bi[4 * j + jp], cpv[jp], t);
                bcv[j] = t;
            }
            // record: [sd_j | E'_j] | Bi | c'
            double *rs = rb + (int64_t)k * recw + L::REC;
            for (int e = tid; e < psk * 2 * n; e += THREADS) {
                const int j = e / (2 * n), r = e % (2 * n);
                rs[e] = r < n ? sdv[j * n + r] : Ev[j * n + r - n];
            }
            if (tid < 16) rs[L::sB + tid] = Bi[tid];
            else if (tid < 20) rs[L::sCt + tid - 16] = cpv[tid - 16];
        }
        __syncthreads();  // every warp is done with Ts and Ss
        if (k + 1 < N) issue_T(k + 1);
        if (tid < n) {
            double t = rho + red[tid];
            for (int j = 0; j < psk; ++j) t = fma(-bcv[j], Ev[j * n + tid], t);
            dps[tid] = t;
        }
        // exact symmetrisation of Cp' through Ss
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct)
            *reinterpret_cast<double2 *>(Ss + r0 * LB + 8 * ct + 2 * q) = make_double2(G[ct][0], G[ct][1]);
        __syncthreads();
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct)
            SM_UNROLL
            for (int e = 0; e < 2; ++e) Cp[ct][e] = 0.5 * (G[ct][e] + Ss[(8 * ct + 2 * q + e) * LB + r0]);
        // Cp -= 1/2 (E'_j W_j' + W_j E'_j')   (products rounded separately: bitwise symmetric)
        for (int j = 0; j < psk; ++j) {
            const double *E = Ev + j * n, *W = Wv + j * n;
            const double er = E[r0], wr = W[r0];
            SM_UNROLL
            for (int ct = 0; ct < NT; ++ct)
                SM_UNROLL
                for (int e = 0; e < 2; ++e) {
                    const int c = 8 * ct + 2 * q + e;
                    const double p1 = __dmul_rn(er, W[c]), p2 = __dmul_rn(wr, E[c]);
                    Cp[ct][e] = __dadd_rn(Cp[ct][e], -0.5 * __dadd_rn(p1, p2));
                }
        }
        // no barrier here: Ss / the stage vectors are next written after the barriers of the following Gauss-Jordan
    }
    // ---- last block: mu_N' = Bl'^-1 y_mu   (Cp, dps hold Bl' and y_mu)
    if (free_final) {  // no goal rows (the last knot carried a zero block): mu_N = 0, nothing to invert
        if (tid == 0) atomicMax(cinfo + inst, spread);
        if (tid < n) {
            xs[tid] = 0.0;
            __stcs(mb + L::mult_rows(N, ps) - n + tid, 0.0);
        }
        if (info && tid == 0) {
            const int hcode = hinfo[inst];
            info[inst] = hcode != 0x7f7f7f7f ? hcode : st_all;
        }
        __syncthreads();
    } else {
        if (tid == 0) {
            flag = 0;
            flag3[1] = 0x7fffffff;
            flag3[2] = 0;
        }
        if (tid < n) ys[tid] = dps[tid];
        __syncthreads();
        const int bad = block_gj_inverse<NT>(Cp, pan, pis, colb, &flag, wp, lane);
        if (bad != 0 && st_all == 0) st_all = N * 1000 + 100 + bad;
        if (tid == 0) atomicMax(cinfo + inst, max(spread, pivot_bits(flag3[1], flag3[2])));
        SM_UNROLL
        for (int ct = 0; ct < NT; ++ct)
            *reinterpret_cast<double2 *>(Ss + r0 * LB + 8 * ct + 2 * q) = make_double2(Cp[ct][0], Cp[ct][1]);
        __syncthreads();
        if (tid < n) {
            double s0 = 0.0, s1 = 0.0;
            SM_UNROLL
            for (int l = 0; l < n; l += 2) {
                s0 = fma(Ss[l * LB + tid], ys[l], s0);
                s1 = fma(Ss[(l + 1) * LB + tid], ys[l + 1], s1);
            }
            xs[tid] = s0 + s1;
            __stcs(mb + L::mult_rows(N, ps) - n + tid, -(s0 + s1));  // mu_N
        }
        if (info && tid == 0) {
            const int hcode = hinfo[inst];
            info[inst] = hcode != 0x7f7f7f7f ? hcode : st_all;
        }
        __syncthreads();
    }

    // ---------------- backward sweep: x_{k-1} = v_k + Z_k' x_k,  Lambda = -x  (all operands from global / L2)
    for (int k = N - 1; k >= 0; --k) {
        const bool first = k == 0, last = k == N - 1;
        const int mk = last ? 0 : m, wk = n + mk;
        const double *rk = rb + (int64_t)k * recw;
        const double *slot = pb + (int64_t)k * hs;
        const double *kp = db + L::knot_off(k, ps);
        const int psk = (ST && k > 0 && k < N - 1) ? ps : 0;
        const double *D1g = last ? kp + L::oCl : kp + L::oD1;  // [A B] or C_N, column-major n x wk
        // the three operands of the NEXT step (record, [g | A B], Hi slot) go to L2 now: the mat-vecs below read
        // straight from global memory and every step is three dependent load phases
        if (k > 0 && tid == 0) {
            const double *rn = rb + (int64_t)(k - 1) * recw;
            const double *sn = pb + (int64_t)(k - 1) * hs;
            // (16-byte aligned start, whole 16-byte pieces: with an odd number of stage rows a knot can start on an odd double)
            const double *kn = reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(db + L::knot_off(k - 1, ps) + L::og) & ~(uintptr_t)15);
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rn), "r"(recw * 8) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kn), "r"((L::CORE - L::og) / 2 * 16) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sn + L::hQi), "r"(n * n * 8) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(sn + L::hRi), "r"(m * m * 8) : "memory");
        }
        // x_prev = v + Z' x : 4 partial sums per entry
        {
            const int i = tid % n, part = tid / n;  // THREADS = 4 n
            // all n/4 loads in flight at once (they come from L2: one latency per phase, not one per batch of 4)
            double zv[n / 4];
            SM_UNROLL
            for (int l = 0; l < n / 4; ++l) zv[l] = rk[(part * (n / 4) + l) * n + i];
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n / 4; ++l) s = fma(zv[l], xs[part * (n / 4) + l], s);
            red[part * n + i] = s;
        }
        const double *rs = rk + L::REC;  // stage part of the record: [sd_j | E'_j] | Bi | c'
        if (psk > 0) {
            // xi = Bi (c' - E' x_k)   (mu_k = -xi);  x_{k-1} -= sum_j xi_j sd_j
            constexpr int NC = (n + 31) / 32;
            for (int j = wp; j < psk; j += L::WARPS) {
                double t = 0.0;
                SM_UNROLL
                for (int c = 0; c < NC; ++c)
                    if (lane + 32 * c < n) t = fma(rs[j * 2 * n + n + lane + 32 * c], xs[lane + 32 * c], t);
                SM_UNROLL
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) cpv[j] = rs[L::sCt + j] - t;
            }
            __syncthreads();
            if (tid < psk) {
                double t = 0.0;
                for (int jp = 0; jp < psk; ++jp) t = fma(rs[L::sB + 4 * tid + jp], cpv[jp], t);
                xiv[tid] = t;
            }
        }
        __syncthreads();
        if (tid < n) {
            double xp = rk[n * n + tid] + (red[tid] + red[n + tid]) + (red[2 * n + tid] + red[3 * n + tid]);
            for (int j = 0; j < psk; ++j) xp = fma(-xiv[j], rs[j * 2 * n + tid], xp);
            xps[tid] = xp;
        }
        __syncthreads();
        // res_j = g_j - sum_i D1[i][j] x_i (+ x_prev_j) (- sum_i C_1[i][j] mu1'_i at the first knot)
        {
            constexpr int NC = (n + 31) / 32, JW = (w + L::WARPS - 1) / L::WARPS;  // columns per warp
            double dv[JW][NC], gj[JW];
            SM_UNROLL
            for (int jj = 0; jj < JW; ++jj) {  // every load of the phase first
                const int j = wp + jj * L::WARPS;
                SM_UNROLL
                for (int c = 0; c < NC; ++c) dv[jj][c] = (j < wk && lane + 32 * c < n) ? D1g[lane + 32 * c + n * j] : 0.0;
                gj[jj] = (j < wk && lane == 0 && !soc) ? kp[(last ? L::HQ : L::og) + j] : 0.0;
            }
            SM_UNROLL
            for (int jj = 0; jj < JW; ++jj) {
                const int j = wp + jj * L::WARPS;
                if (j >= wk) break;  // warp-uniform
                double s = 0.0;
                SM_UNROLL
                for (int c = 0; c < NC; ++c)
                    if (lane + 32 * c < n) s = fma(dv[jj][c], xs[lane + 32 * c], s);
                if (first) {
                    const double *C0 = kp + L::oC0;
                    SM_UNROLL
                    for (int c = 0; c < NC; ++c)
                        if (lane + 32 * c < n) s = fma(C0[lane + 32 * c + n * j], xps[lane + 32 * c], s);
                }
                SM_UNROLL
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) {
                    double r = gj[jj] - s;
                    if (!first && j < n) r += xps[j];
                    for (int jp = 0; jp < psk; ++jp) r = fma(-xiv[jp], kp[L::CORE + jp + psk * j], r);  // C' mu_k, mu = -xi
                    rsv[j] = r;
                }
            }
        }
        __syncthreads();
        // dz = -Hi res
        {
            const double *Qi = first ? pb + (int64_t)N * hs : slot + L::hQi;
            const int i = tid % n, part = tid / n;  // THREADS = 4 n
            double qv[n / 4];
            SM_UNROLL
            for (int l = 0; l < n / 4; ++l) qv[l] = Qi[(part * (n / 4) + l) * n + i];
            double s = 0.0;
            SM_UNROLL
            for (int l = 0; l < n / 4; ++l) s = fma(qv[l], rsv[part * (n / 4) + l], s);
            red[part * n + i] = s;
        }
        __syncthreads();
        if (tid < n) {
            const double z = -((red[tid] + red[n + tid]) + (red[2 * n + tid] + red[3 * n + tid]));
            __stcs(zb + (int64_t)k * w + tid, z);
            if (resb) __stcs(resb + (int64_t)k * w + tid, rsv[tid]);
            // multipliers: [mu_1 (n); lam_1 (n); mu_2 (ps); lam_2; ...; lam_{N-1}; mu_N]: this knot writes lam_{k-1} (or mu_1)
            // and its own stage multipliers mu_k
            const int64_t lo_ = first ? 0 : (int64_t)n + (int64_t)(k - 1) * (n + ps);
            __stcs(mb + lo_ + tid, -xps[tid]);
            if (tid < psk) __stcs(mb + lo_ + n + tid, -xiv[tid]);
        } else if (tid < wk) {
            const double *Ri = slot + L::hRi;
            double s = 0.0;
            for (int l = 0; l < m; ++l) s = fma(Ri[l * m + tid - n], rsv[n + l], s);
            __stcs(zb + (int64_t)k * w + tid, -s);
            if (resb) __stcs(resb + (int64_t)k * w + tid, rsv[tid]);
        }
        __syncthreads();
        if (tid < n) xs[tid] = xps[tid];
        __syncthreads();
    }
}

}  // namespace kcta
