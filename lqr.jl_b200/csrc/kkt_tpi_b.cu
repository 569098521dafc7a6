// kkt_tpi_b.cu — thread-per-instance KKT kernels, part B of the size list (kkt_dispatch.cuh).
#define KKT_TPI_PART_SIZES KKT_TPI_SIZES_B
#define KKT_TPI_PART_NAME kkt_launch_tpi_b
#include "kkt_tpi_part.cuh"
