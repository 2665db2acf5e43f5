// prf_build.cu -- the build phase of hashRF (dedup + incidence), see prf_incidence.cuh.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_incidence.cuh"
