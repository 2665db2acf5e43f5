// prf_cluster.cuh -- first consumer of the matrix: flat clusters at edit distance <= D.
//
// The reference cuts the dendrogram of C.dendrogram at D (sliceDendro, PhyBin.hs:415-422).  For single
// linkage the resulting flat clusters are exactly the connected components of the graph
// { (a,b) : d(a,b) <= D }; for complete linkage and UPGMA every flat cluster lies inside one component,
// so the components split the O(T^2)..O(T^3) host clustering into independent small problems.  The
// graph is the fused threshold mask the matrix kernels already wrote: one pass over its words with a
// lock-free union-find (hook the larger root under the smaller with atomicCAS, path halving).
#pragma once

#include "prf_common.cuh"

namespace prf {

__device__ __forceinline__ uint32_t cc_find(uint32_t* parent, uint32_t x) {
  for (;;) {
    const uint32_t p = ((volatile uint32_t*)parent)[x];
    if (p == x) return x;
    const uint32_t gp = ((volatile uint32_t*)parent)[p];
    if (gp != p) atomicCAS(&parent[x], p, gp);  // path halving; losing the race is harmless
    x = p;
  }
}

__device__ __forceinline__ void cc_unite(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    const uint32_t hi = max(a, b), lo = min(a, b);
    if (atomicCAS(&parent[hi], hi, lo) == hi) return;  // hi was still a root: hooked under the smaller id
  }
}

__global__ void __launch_bounds__(256) cc_init_kernel(uint32_t* parent, uint32_t T) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) parent[t] = t;
}

// One thread per mask word: every set bit p is the pair (a, b) of packed position p_begin + 32 w + bit.
__global__ void __launch_bounds__(256) cc_hook_kernel(const uint32_t* __restrict__ mask, uint64_t n_words, uint64_t p_begin,
                                                      uint64_t p_end, uint32_t T, uint32_t* parent) {
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t bits = mask[w];
  if (!bits) return;
  const uint64_t p0 = p_begin + (w << 5);
  uint32_t a = row_of(p0, T);
  uint64_t rs = row_start(a, T), re = rs + (T - 1 - a);  // row a holds positions [rs, re)
  while (bits) {
    const uint32_t i = __ffs(bits) - 1;
    bits &= bits - 1;
    const uint64_t p = p0 + i;
    if (p >= p_end) break;
    while (p >= re) {  // the word straddles a row boundary
      ++a;
      rs = re;
      re = rs + (T - 1 - a);
    }
    cc_unite(parent, a, a + 1 + (uint32_t)(p - rs));
  }
}

// Edges (t, seed[i * T + t]): the components another shard (rank) found, merged into this union-find.
__global__ void __launch_bounds__(256) cc_seed_kernel(const uint32_t* __restrict__ seed, uint64_t n, uint32_t T, uint32_t* parent) {
  const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint32_t t = (uint32_t)(x % T), l = seed[x];
  if (l != t) cc_unite(parent, t, l);
}

// label[t] = smallest tree id of t's component; counts the roots.
__global__ void __launch_bounds__(256) cc_flatten_kernel(uint32_t* parent, uint32_t T, uint32_t* __restrict__ label,
                                                         uint32_t* n_roots) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t r = cc_find(parent, t);
  label[t] = r;
  if (r == t) atomicAdd(n_roots, 1u);
}

// Sub-matrix of one component: out[q], q = packed position of (i, j), i < j < m, in an m x m triangle, is
// d(ids[i], ids[j]) read from the full packed triangle `src` (which starts at packed position src_begin).
// ids must be ascending.
__global__ void __launch_bounds__(256) submatrix_gather_kernel(const uint16_t* __restrict__ src, uint64_t src_begin,
                                                               uint32_t T, const uint32_t* __restrict__ ids, uint32_t m,
                                                               uint64_t n_out, uint16_t* __restrict__ out) {
  const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_out) return;
  const uint32_t i = row_of(q, m);
  const uint32_t j = i + 1 + (uint32_t)(q - row_start(i, m));
  const uint32_t a = ids[i], b = ids[j];
  out[q] = src[row_start(a, T) + (b - a - 1) - src_begin];
}

}  // namespace prf
