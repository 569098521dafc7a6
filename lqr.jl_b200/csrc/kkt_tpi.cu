// kkt_tpi.cu — thread-per-instance KKT kernels, part A of the size list (the reference's fixtures and BASELINE configs).
#define KKT_TPI_PART_SIZES KKT_TPI_SIZES_A
#define KKT_TPI_PART_NAME kkt_launch_tpi_a
#include "kkt_tpi_part.cuh"

int32_t kkt_launch_tpi_b(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                         double *dz, double *mult, double *res, int32_t *info, cudaStream_t st);
int32_t kkt_launch_tpi_c(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                         double *dz, double *mult, double *res, int32_t *info, cudaStream_t st);
int32_t kkt_launch_tpi_d(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                         double *dz, double *mult, double *res, int32_t *info, cudaStream_t st);
int32_t kkt_launch_tpi_e(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                         double *dz, double *mult, double *res, int32_t *info, cudaStream_t st);

int32_t kkt_launch_tpi(lqrb_context *h, const KktShape &s, int64_t batch, int flags, const double *data, double *scratch,
                       double *dz, double *mult, double *res, int32_t *info, cudaStream_t st) {
    int32_t rc = kkt_launch_tpi_a(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    if (rc == LQRB_NO_KERNEL) rc = kkt_launch_tpi_b(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    if (rc == LQRB_NO_KERNEL) rc = kkt_launch_tpi_c(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    if (rc == LQRB_NO_KERNEL) rc = kkt_launch_tpi_d(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    if (rc == LQRB_NO_KERNEL) rc = kkt_launch_tpi_e(h, s, batch, flags, data, scratch, dz, mult, res, info, st);
    return rc;
}
