// sqp.cu — batched Dubins SQP driver (placeholder until the device linearisation lands).
#include "common.cuh"

extern "C" int32_t lqrb_sqp_dubins_f64(lqrb_handle_t h, int64_t batch, const lqrb_sqp_options_t *opts,
                                       const double *x0, const double *xf, double *Z, double *feas_p,
                                       double *feas_d, int32_t *iters_done, int64_t *kkt_solves) {
    if (!h) return -1;
    return lqrb_fail(h, -1, "lqrb_sqp_dubins_f64: not built yet");
}
