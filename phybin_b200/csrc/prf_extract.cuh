// prf_extract.cuh -- allBips (RFDistance.hs:207-248) and the tolerant mode's raw clades on the device
// (SURVEY 8f-2): the host hands over the parsed trees as flat pre-order arrays and gets back, in device
// memory, exactly what phb_allbips / phb_raw_clades (phybin_host.cpp) produce -- per tree the SET of
// canonical bipartitions (or un-normalised interior clades), sorted by word-lexicographic key order.
//
// One warp per tree, everything in that warp's slice of shared memory:
//   1. clade bit sets: leaves set their bit, then one reverse pre-order sweep ORs every node into its
//      parent (labelBips, :207-221).  Lane w owns word w of every set, so the sweep needs no barrier.
//   2. normBip (:229-239) per node, lanes over nodes: keep the side with < size/2 members, flip to
//      {0..size-1} \ b above it, IntSet-min of the two on a tie; drop sets of < 2 members (:248), the
//      interior root, and leaves unless size < 4.
//   3. bitonic sort of the surviving node ids by key, adjacent-unique (a tree's bipartitions are a Set),
//      keys written behind the tree's node offset in a scratch array; a scan of the per-tree counts and a
//      copy pack them.
// Shared memory per warp = max_nodes * (8 W + 2) + 2 * pow2(max_nodes) bytes; the host picks the warps
// per CTA that fit and refuses shapes where one tree does not fit.
#pragma once

#include "prf_primitives.cuh"

namespace prf {

constexpr uint32_t kExtractMaxNodes = 32768;   // node ids are uint16 in shared memory, 0xffff = padding
constexpr size_t kExtractSmemBudget = 200 * 1024;

__host__ __device__ inline uint32_t pow2_ceil(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

__host__ inline size_t extract_warp_bytes(uint32_t max_nodes, uint32_t W) {
  size_t b = (size_t)max_nodes * W * 8 + (size_t)max_nodes * 2 + (size_t)pow2_ceil(max_nodes) * 2;
  return (b + 15) & ~(size_t)15;
}

template <int W>
__device__ __forceinline__ bool key_less(const uint64_t* __restrict__ bits, uint32_t x, uint32_t y) {
  if (x == 0xffffu) return false;  // padding sorts last
  if (y == 0xffffu) return true;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    const uint64_t a = bits[x * W + w], b = bits[y * W + w];
    if (a != b) return a < b;
  }
  return false;
}

// err[0] |= 1: a parent outside [tree start, node); |= 2: a leaf label >= 64 W; |= 4: num_leaves > 64 W.
template <int W, int kMode>
__global__ void __launch_bounds__(256) extract_kernel(const int32_t* __restrict__ label,
                                                      const uint32_t* __restrict__ parent,
                                                      const uint64_t* __restrict__ tree_begin,
                                                      const uint32_t* __restrict__ num_leaves, uint32_t T,
                                                      uint32_t warp_bytes, uint32_t max_nodes,
                                                      uint64_t* __restrict__ scratch, uint32_t* __restrict__ count,
                                                      uint64_t* __restrict__ leafsets, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= T) return;
  unsigned char* mine = smem_raw + (size_t)warp * warp_bytes;
  uint64_t* bits = reinterpret_cast<uint64_t*>(mine);
  uint16_t* par = reinterpret_cast<uint16_t*>(mine + (size_t)max_nodes * W * 8);
  uint16_t* perm = par + max_nodes;

  const uint64_t nb = tree_begin[t];
  const uint32_t m = (uint32_t)(tree_begin[t + 1] - nb);
  const uint32_t size = num_leaves[t];
  uint32_t bad = size > 64u * W ? 4u : 0u;

  // 1. leaf bits and tree-relative parents
  for (uint32_t i = lane; i < m * W; i += 32) bits[i] = 0;
  __syncwarp();
  for (uint32_t v = lane; v < m; v += 32) {
    const int32_t lab = label[nb + v];
    if (lab >= 0) {
      if ((uint32_t)lab < 64u * W) bits[v * W + ((uint32_t)lab >> 6)] = 1ull << (lab & 63);
      else bad |= 2u;
    }
    uint32_t p = 0;
    if (v) {
      const uint64_t pg = parent[nb + v];
      if (pg < nb || pg >= nb + v) bad |= 1u; else p = (uint32_t)(pg - nb);
    }
    par[v] = (uint16_t)p;
  }
  __syncwarp();
  if (lane < W) {  // lane w sweeps word w bottom-up: pre-order puts every child behind its parent
    for (uint32_t v = m; v-- > 1;) {
      const uint32_t p = par[v];
      bits[p * W + lane] |= bits[v * W + lane];
    }
  }
  __syncwarp();
  if (kMode == PRF_EXTRACT_CLADES && m && lane < W) leafsets[(uint64_t)t * W + lane] = bits[lane];

  // 2. normalise in place, collect the surviving node ids in node order
  uint32_t k = 0;
  for (uint32_t base = 0; base < m; base += 32) {
    const uint32_t v = base + lane;
    bool keep = false;
    if (v < m) {
      const bool is_leaf = label[nb + v] >= 0;
      if (kMode == PRF_EXTRACT_CLADES) {
        keep = v > 0 && !is_leaf;
      } else if (!(v == 0 && !is_leaf) && !(is_leaf && size >= 4)) {
        uint64_t b[W];
        uint32_t cnt = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) { b[w] = bits[v * W + w]; cnt += __popcll(b[w]); }
        const uint32_t half = size >> 1;
        if (cnt >= half) {
          uint64_t f[W];
          int lf = -1, lb = -1;
          uint32_t fc = 0;
#pragma unroll
          for (int w = 0; w < W; ++w) {
            const int64_t rem = (int64_t)size - 64ll * w;
            const uint64_t uni = rem >= 64 ? ~0ull : (rem <= 0 ? 0ull : ((1ull << rem) - 1));
            f[w] = uni & ~b[w];
            fc += __popcll(f[w]);
            if (lf < 0 && f[w]) lf = w * 64 + __ffsll((long long)f[w]) - 1;
            if (lb < 0 && b[w]) lb = w * 64 + __ffsll((long long)b[w]) - 1;
          }
          const bool flip = cnt > half || lf < 0 || (lb >= 0 && lf < lb);
          if (flip) {
#pragma unroll
            for (int w = 0; w < W; ++w) bits[v * W + w] = f[w];
            cnt = fc;
          }
        }
        keep = cnt > 1;
      }
    }
    const uint32_t vote = __ballot_sync(0xffffffffu, keep);
    if (keep) perm[k + __popc(vote & lanemask_lt())] = (uint16_t)v;
    k += __popc(vote);
  }
  const uint32_t np = pow2_ceil(k < 2 ? 2 : k);
  for (uint32_t i = k + lane; i < np; i += 32) perm[i] = 0xffffu;
  __syncwarp();

  // 3. bitonic sort of perm[0, np) by key
  for (uint32_t len = 2; len <= np; len <<= 1) {
    for (uint32_t stride = len >> 1; stride; stride >>= 1) {
      for (uint32_t i = lane; i < (np >> 1); i += 32) {
        const uint32_t lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1)), hi = lo + stride;
        const uint32_t x = perm[lo], y = perm[hi];
        const bool up = (lo & len) == 0;
        if (key_less<W>(bits, up ? y : x, up ? x : y)) { perm[lo] = (uint16_t)y; perm[hi] = (uint16_t)x; }
      }
      __syncwarp();
    }
  }

  // adjacent-unique, written behind the tree's node offset (a tree has at most m surviving nodes)
  uint32_t n_out = 0;
  uint64_t* dst = scratch + nb * W;
  for (uint32_t base = 0; base < k; base += 32) {
    const uint32_t i = base + lane;
    bool fresh = false;
    uint32_t x = 0;
    if (i < k) {
      x = perm[i];
      fresh = i == 0 || key_less<W>(bits, perm[i - 1], x);
    }
    const uint32_t vote = __ballot_sync(0xffffffffu, fresh);
    if (fresh) {
      uint64_t* o = dst + (uint64_t)(n_out + __popc(vote & lanemask_lt())) * W;
#pragma unroll
      for (int w = 0; w < W; ++w) o[w] = bits[x * W + w];
    }
    n_out += __popc(vote);
  }
  if (lane == 0) count[t] = n_out;
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad && lane == 0) atomicOr(err, bad);
}

// One warp per tree: keys from the scratch array to their packed place; off64[t] for t <= T.
__global__ void __launch_bounds__(256) extract_pack_kernel(const uint64_t* __restrict__ scratch,
                                                           const uint64_t* __restrict__ tree_begin,
                                                           const uint32_t* __restrict__ off32, uint32_t T, uint32_t W,
                                                           uint64_t* __restrict__ keys, uint64_t* __restrict__ off64) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t > T) return;
  const uint32_t o = off32[t];
  if (lane_id() == 0) off64[t] = o;
  if (t == T) return;
  const uint64_t n = (uint64_t)(off32[t + 1] - o) * W;
  const uint64_t* src = scratch + tree_begin[t] * W;
  uint64_t* dst = keys + (uint64_t)o * W;
  for (uint64_t i = lane_id(); i < n; i += 32) dst[i] = src[i];
}

template <int W>
int launch_extract_w(Ctx* ctx, int mode, const int32_t* label, const uint32_t* parent, const uint64_t* tree_begin,
                     const uint32_t* num_leaves, uint32_t T, uint32_t warps, uint32_t warp_bytes, uint32_t max_nodes,
                     uint64_t* scratch, uint32_t* count, uint64_t* leafsets, uint32_t* err) {
  const size_t smem = (size_t)warps * warp_bytes;
  const uint32_t grid = (T + warps - 1) / warps;
  if (mode == PRF_EXTRACT_CLADES) {
    PRF_CK(ctx, cudaFuncSetAttribute(extract_kernel<W, PRF_EXTRACT_CLADES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    extract_kernel<W, PRF_EXTRACT_CLADES><<<grid, warps * 32, smem, ctx->stream>>>(label, parent, tree_begin, num_leaves, T, warp_bytes,
                                                                                  max_nodes, scratch, count, leafsets, err);
  } else {
    PRF_CK(ctx, cudaFuncSetAttribute(extract_kernel<W, PRF_EXTRACT_BIPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    extract_kernel<W, PRF_EXTRACT_BIPS><<<grid, warps * 32, smem, ctx->stream>>>(label, parent, tree_begin, num_leaves, T, warp_bytes,
                                                                                max_nodes, scratch, count, leafsets, err);
  }
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
