// kkt_tpi_d.cu — thread-per-instance KKT kernels, part D of the size list (kkt_dispatch.cuh).
#define KKT_TPI_PART_SIZES KKT_TPI_SIZES_D
#define KKT_TPI_PART_NAME kkt_launch_tpi_d
#include "kkt_tpi_part.cuh"
