// prf_dense.cu -- the tcgen05 Gram variant, see prf_dense.cuh.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_dense.cuh"

namespace prf {
uint32_t dense_kpad(uint32_t n_shared) { return (n_shared + kDK - 1) / kDK * kDK; }
uint32_t dense_tpad(uint32_t T) { return (T + kDN - 1) / kDN * kDN; }
}  // namespace prf
