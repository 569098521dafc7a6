// kkt_tpi_part.cuh — body of one translation unit of the thread-per-instance KKT kernels (kkt_kernels.cuh).  The size
// list (kkt_dispatch.cuh) is cut into parts that compile in parallel: kkt_tpi.cu, kkt_tpi_b.cu, kkt_tpi_c.cu define
// KKT_TPI_PART_SIZES / KKT_TPI_PART_NAME and include this file.
#include "kkt_dispatch.cuh"
#include "kkt_kernels.cuh"

template <int n, int m, int P1, int PM, int PN>
static int32_t launch_kkt_tpi(lqrb_context *h, const KktShape &s, int64_t batch, int flags,
                              const double *data, double *scratch, double *dz, double *mult,
                              double *res, int32_t *info, cudaStream_t st) {
    constexpr int THREADS = 64;
    const unsigned grid = (unsigned)((batch + THREADS - 1) / THREADS);
    const bool soc = (flags & LQRB_FLAG_SOC) != 0;
#define LAUNCH(HESS, SOC) \
    kkt_tpi_kernel<n, m, P1, PM, PN, HESS, SOC, THREADS><<<grid, THREADS, 0, st>>>(data, scratch, dz, mult, res, info, s.N, batch)
    if (soc) {
        // H and g are ignored: any HESS instantiation reads the same rows layout it was packed with
        if (s.hess == LQRB_HESS_DIAG) LAUNCH(LQRB_HESS_DIAG, true);
        else if (s.hess == LQRB_HESS_BLOCKDIAG) LAUNCH(LQRB_HESS_BLOCKDIAG, true);
        else LAUNCH(LQRB_HESS_DENSE, true);
    } else {
        if (s.hess == LQRB_HESS_DIAG) LAUNCH(LQRB_HESS_DIAG, false);
        else if (s.hess == LQRB_HESS_BLOCKDIAG) LAUNCH(LQRB_HESS_BLOCKDIAG, false);
        else LAUNCH(LQRB_HESS_DENSE, false);
    }
#undef LAUNCH
    char nm[96];
    snprintf(nm, sizeof nm, "kkt_tpi<%d,%d,p=%d/%d/%d,hess=%d%s>", n, m, P1, PM, PN, s.hess, soc ? ",soc" : "");
    h->kernel_name = nm;
    LQRB_LAUNCH_CHECK(h, "kkt_tpi_kernel");
    return 0;
}

int32_t KKT_TPI_PART_NAME(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                          double *dz, double *mult, double *res, int32_t *info, cudaStream_t st) {
#define X(N_, M_, A_, B_, C_)                                                            \
    if (s.n == N_ && s.m == M_ && s.P1 == A_ && (s.N == 2 || s.PM == B_) && s.PN == C_) \
        return launch_kkt_tpi<N_, M_, A_, B_, C_>(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    KKT_TPI_PART_SIZES(X)
#undef X
    return LQRB_NO_KERNEL;
}
