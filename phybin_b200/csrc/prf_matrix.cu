// prf_matrix.cu -- the sparse matrix kernels: prf_rows.cuh (one warp per row) and the round-1 kernel.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_sparse.cuh"
#include "prf_rows.cuh"
