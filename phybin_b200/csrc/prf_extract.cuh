// prf_extract.cuh -- allBips (RFDistance.hs:207-248) and the tolerant mode's raw clades on the device
// (SURVEY 8f-2): the host hands over the parsed trees as flat pre-order arrays and gets back, in device
// memory, exactly what phb_allbips / phb_raw_clades (phybin_host.cpp) produce -- per tree the SET of
// canonical bipartitions (or un-normalised interior clades), each set once, in the pre-order of the first
// node that yields it.
//
// One warp per tree, everything in that warp's slice of shared memory:
//   1. clade bit sets: every leaf sets its bit in its own set and (atomicOr) in its parent's; the non-root
//      interior nodes are then swept bottom-up -- reverse pre-order -- each ORed into its parent
//      (labelBips, :207-221).  Lane w owns word w of every set, so the sweep needs no barrier.
//   2. normBip (:229-239) per node, lanes over nodes: keep the side with < size/2 members, flip to
//      {0..size-1} \ b above it, IntSet-min of the two on a tie; drop sets of < 2 members (:248), the
//      interior root, and leaves unless size < 4.
//   3. a tree's bipartitions are a Set: the surviving nodes go into a per-warp open-addressing table keyed by
//      the whole bit set (exact compare); equal sets keep the smallest node id.  The representatives are
//      written, in node order, behind the tree's node offset in a scratch array; a scan of the per-tree counts
//      and a copy pack them.
// Only nodes whose set can survive own a slot of W words: the interior nodes (leaf singletons are dropped by
// the `> 1` filter unless the tree has fewer than 4 leaves, in which case every node gets a slot).
// Shared memory per warp = max_slots * (8 (W|1) + 4) + 2 * max_nodes + 2 * pow2(2 max_slots + 2) bytes; the host
// picks the CTA shape that keeps most warps resident and refuses shapes where one tree does not fit.
#pragma once

#include "prf_primitives.cuh"

namespace prf {

constexpr uint32_t kExtractMaxNodes = 16384;   // node ids and table slots are uint16 in shared memory
constexpr size_t kExtractSmemBudget = 200 * 1024;

__host__ __device__ inline uint32_t pow2_ceil(uint32_t x) {
  uint32_t p = 1;
  while (p < x) p <<= 1;
  return p;
}

// slots of a tree of m nodes and `size` leaves
__host__ __device__ inline uint32_t extract_slots(uint32_t m, uint32_t size) { return size < 4 ? m : m - (size < m ? size : m); }

// words between consecutive sets in shared memory: odd, so that lanes reading "their" set do not all hit the same banks
__host__ __device__ constexpr uint32_t extract_stride(uint32_t W) { return W | 1u; }
// table positions for a tree with n_slots sets: load factor <= 1/2
__host__ __device__ inline uint32_t extract_table(uint32_t n_slots) { return pow2_ceil(2 * n_slots + 2); }

__host__ inline size_t extract_warp_bytes(uint32_t max_slots, uint32_t max_nodes, uint32_t W) {
  size_t b = (size_t)max_slots * extract_stride(W) * 8 + (size_t)max_slots * 4 + (size_t)max_nodes * 2 + (size_t)extract_table(max_slots) * 2;
  return (b + 15) & ~(size_t)15;
}

// err[0] |= 1: a parent outside [tree start, node) or not an interior node; |= 2: a leaf label >= 64 W;
// |= 4: num_leaves > 64 W or not the tree's number of leaf nodes.
template <int W, int kMode>
__global__ void __launch_bounds__(256) extract_kernel(const int32_t* __restrict__ label,
                                                      const uint32_t* __restrict__ parent,
                                                      const uint64_t* __restrict__ tree_begin,
                                                      const uint32_t* __restrict__ num_leaves, uint32_t T,
                                                      uint32_t warp_bytes, uint32_t max_slots, uint32_t max_nodes,
                                                      uint64_t* __restrict__ scratch, uint32_t* __restrict__ count,
                                                      uint64_t* __restrict__ leafsets, uint32_t* __restrict__ err) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + warp;
  if (t >= T) return;
  unsigned char* mine = smem_raw + (size_t)warp * warp_bytes;
  constexpr uint32_t S = extract_stride(W);
  uint64_t* bits = reinterpret_cast<uint64_t*>(mine);                                   // [slot][S], W words used
  uint16_t* pslot = reinterpret_cast<uint16_t*>(mine + (size_t)max_slots * S * 8);      // [slot] parent's slot
  uint16_t* where = pslot + max_slots;     // [slot] 0xffff = dropped, else (after step 3) its table position
  uint16_t* slot_of = where + max_slots;   // [node] slot, 0xffff = none (a leaf of a tree with >= 4 leaves)
  unsigned short* tab = slot_of + max_nodes;  // open-addressing table of slots, 0xffff = free

  const uint64_t nb = tree_begin[t];
  const uint32_t m = (uint32_t)(tree_begin[t + 1] - nb);
  const uint32_t size = num_leaves[t];
  const bool all_slots = size < 4;
  const uint32_t n_slots = extract_slots(m, size);
  const uint32_t tmask = extract_table(n_slots) - 1;
  uint32_t bad = size > 64u * W ? 4u : 0u;

  // 1. slots in pre-order; leaf bits go straight into the parent's set (and the leaf's own, if it has one)
  for (uint32_t i = lane; i < n_slots * S; i += 32) bits[i] = 0;
  for (uint32_t i = lane; i <= tmask; i += 32) tab[i] = 0xffffu;
  uint32_t ns = 0;
#pragma unroll 4
  for (uint32_t base = 0; base < m; base += 32) {
    const uint32_t v = base + lane;
    const bool slotted = v < m && (all_slots || label[nb + v] < 0);
    const uint32_t vote = __ballot_sync(0xffffffffu, slotted);
    if (v < m) slot_of[v] = slotted ? (uint16_t)(ns + __popc(vote & lanemask_lt())) : (uint16_t)0xffffu;
    ns += __popc(vote);
  }
  if (ns != n_slots) bad |= 4u;  // num_leaves is not this tree's leaf count: the slot budget would be wrong
  __syncwarp();
  if (!(bad & 4u)) {
#pragma unroll 4
    for (uint32_t v = lane; v < m; v += 32) {
      const int32_t lab = label[nb + v];
      const uint32_t sv = slot_of[v];
      uint32_t sp = 0xffffu;
      if (v) {
        const uint64_t pg = parent[nb + v];
        if (pg >= nb && pg < nb + v) sp = slot_of[(uint32_t)(pg - nb)];
        if (sp == 0xffffu) bad |= 1u;
      }
      if (sv != 0xffffu) pslot[sv] = (uint16_t)sp;
      if (lab >= 0) {
        if ((uint32_t)lab < 64u * W) {
          const unsigned long long bit = 1ull << (lab & 63);
          if (sv != 0xffffu) bits[sv * S + ((uint32_t)lab >> 6)] = bit;
          if (sp != 0xffffu) atomicOr(reinterpret_cast<unsigned long long*>(&bits[sp * S + ((uint32_t)lab >> 6)]), bit);
        } else {
          bad |= 2u;
        }
      }
    }
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad) {  // malformed input: report, produce nothing for this tree
    if (lane == 0) { atomicOr(err, bad); count[t] = 0; }
    return;
  }
  __syncwarp();
  if (lane < W) {  // lane w sweeps word w bottom-up: slots are in pre-order, every child behind its parent
    uint32_t sv = n_slots;
    while (sv > 4) {  // parent slots four at a time, so that only the OR chain itself is serial
      const uint32_t p0 = pslot[sv - 1], p1 = pslot[sv - 2], p2 = pslot[sv - 3], p3 = pslot[sv - 4];
      bits[p0 * S + lane] |= bits[(sv - 1) * S + lane];
      bits[p1 * S + lane] |= bits[(sv - 2) * S + lane];
      bits[p2 * S + lane] |= bits[(sv - 3) * S + lane];
      bits[p3 * S + lane] |= bits[(sv - 4) * S + lane];
      sv -= 4;
    }
    while (sv-- > 1) {
      const uint32_t sp = pslot[sv];
      bits[sp * S + lane] |= bits[sv * S + lane];
    }
  }
  __syncwarp();
  if (kMode == PRF_EXTRACT_CLADES && m && lane < W) {
    // the root's set: its slot, or for a single-leaf tree of >= 4 "leaves" (impossible) nothing
    leafsets[(uint64_t)t * W + lane] = n_slots ? bits[lane] : 0ull;
  }

  // 2. normalise in place; where[slot] = 0 for a surviving set, 0xffff otherwise
  for (uint32_t v = lane; v < m; v += 32) {
    const uint32_t sv = slot_of[v];
    if (sv == 0xffffu) continue;
    bool keep = false;
    const bool is_leaf = label[nb + v] >= 0;
    if (kMode == PRF_EXTRACT_CLADES) {
      keep = v > 0 && !is_leaf;
    } else if (!(v == 0 && !is_leaf)) {
      uint64_t b[W];
      uint32_t cnt = 0;
#pragma unroll
      for (int w = 0; w < W; ++w) { b[w] = bits[sv * S + w]; cnt += __popcll(b[w]); }
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
          for (int w = 0; w < W; ++w) bits[sv * S + w] = f[w];
          cnt = fc;
        }
      }
      keep = cnt > 1;
    }
    where[sv] = keep ? 0u : 0xffffu;
  }
  __syncwarp();

  // 3. set semantics: equal keys share a table position that ends up holding their smallest slot
  for (uint32_t sv = lane; sv < n_slots; sv += 32) {
    if (where[sv] == 0xffffu) continue;
    uint64_t k[W];
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int w = 0; w < W; ++w) { k[w] = bits[sv * S + w]; h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1))); }
    uint32_t s = (uint32_t)h & tmask;
    for (;;) {
      const uint32_t prev = atomicCAS(&tab[s], (unsigned short)0xffffu, (unsigned short)sv);
      if (prev == 0xffffu) break;
      bool eq = true;
#pragma unroll
      for (int w = 0; w < W; ++w) eq &= bits[prev * S + w] == k[w];
      if (eq) {
        uint32_t cur = prev;
        while (sv < cur) {  // atomicMin on a half-word
          const uint32_t old = atomicCAS(&tab[s], (unsigned short)cur, (unsigned short)sv);
          if (old == cur) break;
          cur = old;
        }
        break;
      }
      s = (s + 1) & tmask;
    }
    where[sv] = (uint16_t)s;
  }
  __syncwarp();

  // representatives, in pre-order, behind the tree's node offset (a tree has at most m surviving nodes)
  uint32_t n_out = 0;
  uint64_t* dst = scratch + nb * W;
  for (uint32_t base = 0; base < n_slots; base += 32) {
    const uint32_t sv = base + lane;
    const bool rep = sv < n_slots && where[sv] != 0xffffu && tab[where[sv]] == sv;
    const uint32_t vote = __ballot_sync(0xffffffffu, rep);
    if (rep) {
      uint64_t* o = dst + (uint64_t)(n_out + __popc(vote & lanemask_lt())) * W;
#pragma unroll
      for (int w = 0; w < W; ++w) o[w] = bits[sv * S + w];
    }
    n_out += __popc(vote);
  }
  if (lane == 0) count[t] = n_out;
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
                     const uint32_t* num_leaves, uint32_t T, uint32_t warps, uint32_t warp_bytes, uint32_t max_slots, uint32_t max_nodes,
                     uint64_t* scratch, uint32_t* count, uint64_t* leafsets, uint32_t* err) {
  const size_t smem = (size_t)warps * warp_bytes;
  const uint32_t grid = (T + warps - 1) / warps;
  if (mode == PRF_EXTRACT_CLADES) {
    PRF_CK(ctx, cudaFuncSetAttribute(extract_kernel<W, PRF_EXTRACT_CLADES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    extract_kernel<W, PRF_EXTRACT_CLADES><<<grid, warps * 32, smem, ctx->stream>>>(label, parent, tree_begin, num_leaves, T, warp_bytes,
                                                                                  max_slots, max_nodes, scratch, count, leafsets, err);
  } else {
    PRF_CK(ctx, cudaFuncSetAttribute(extract_kernel<W, PRF_EXTRACT_BIPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    extract_kernel<W, PRF_EXTRACT_BIPS><<<grid, warps * 32, smem, ctx->stream>>>(label, parent, tree_begin, num_leaves, T, warp_bytes,
                                                                                max_slots, max_nodes, scratch, count, leafsets, err);
  }
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
