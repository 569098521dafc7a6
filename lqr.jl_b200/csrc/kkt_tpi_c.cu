// kkt_tpi_c.cu — thread-per-instance KKT kernels, part C of the size list (kkt_dispatch.cuh).
#define KKT_TPI_PART_SIZES KKT_TPI_SIZES_C
#define KKT_TPI_PART_NAME kkt_launch_tpi_c
#include "kkt_tpi_part.cuh"
