// kkt_tpi_e.cu — thread-per-instance KKT kernels, part E of the size list (kkt_dispatch.cuh).
#define KKT_TPI_PART_SIZES KKT_TPI_SIZES_E
#define KKT_TPI_PART_NAME kkt_launch_tpi_e
#include "kkt_tpi_part.cuh"
