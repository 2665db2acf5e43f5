// prf_consensus.cuh -- the bipartitions of a strict consensus tree for EVERY cluster at once:
//     consensusTree _ (hd:tl) = bipsToTree num_taxa (foldl' S.intersection (allBips hd) (map allBips tl))
// (RFDistance.hs:396-400; called per cluster by outputClusters, PhyBin.hs:432-436).  The intersection runs on
// the device over the same canonical bit sets hashRF takes; bipsToTree (:410-445) is host work on the few
// surviving sets (phb_consensus_newick).
//
// A cluster is named by its smallest tree id (the labels prf_mask_components returns).  For a cluster c of s
// trees, a bipartition is in the consensus iff all s trees have it, i.e. iff it is one of the representative
// tree c's bipartitions and the other s-1 members have it too:
//   cons_table<insert>  (c, key) of every representative's entry -> open-addressing table (exact key compare)
//   cons_table<probe>   every entry of every other tree looks up (its cluster, key); a hit bumps the
//                       representative entry's counter
//   cons_flag           representative entries whose counter reached s-1 survive (all of them when s = 1)
//   scan + cons_gather   pack the survivors, in the representative's (sorted) order, one range per tree
// Every tree's entries must be distinct (prf_allbips / phb_allbips output is).
#pragma once

#include "prf_primitives.cuh"

namespace prf {

template <int W>
__device__ __forceinline__ uint64_t cons_hash(const uint64_t (&k)[W], uint32_t c) {
  uint64_t h = 0xD6E8FEB86659FD93ull ^ c;
#pragma unroll
  for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
  return h;
}

// One warp per tree.  kProbe = false: representatives of clusters of >= 2 trees insert (c << 32 | entry);
// kProbe = true: the other trees look their entries up and bump cnt[representative entry].
template <int W, bool kProbe>
__global__ void __launch_bounds__(256) cons_table_kernel(const uint64_t* __restrict__ bips,
                                                         const uint64_t* __restrict__ tree_off,
                                                         const uint32_t* __restrict__ labels,
                                                         const uint32_t* __restrict__ size, uint32_t T,
                                                         unsigned long long* slots, uint64_t cap_mask,
                                                         uint32_t* __restrict__ cnt) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint32_t c = labels[t];
  if (kProbe ? (c == t) : (c != t || size[c] < 2)) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  for (uint64_t e = e0 + lane_id(); e < e1; e += 32) {
    uint64_t k[W];
    load_key<W>(bips, e, k);
    uint64_t s = cons_hash<W>(k, c) & cap_mask;
    const unsigned long long mine = ((unsigned long long)c << 32) | (uint32_t)e;
    for (;;) {
      unsigned long long cur = slots[s];
      if (cur == kEmptySlot) {
        if (kProbe) break;  // the representative does not have this bipartition
        cur = atomicCAS(&slots[s], (unsigned long long)kEmptySlot, mine);
        if (cur == kEmptySlot) break;
      }
      if ((uint32_t)(cur >> 32) == c) {
        uint64_t r[W];
        load_key<W>(bips, cur & 0xffffffffull, r);
        bool eq = true;
#pragma unroll
        for (int w = 0; w < W; ++w) eq &= (r[w] == k[w]);
        if (eq) {
          if (kProbe) atomicAdd(&cnt[cur & 0xffffffffull], 1u);
          break;
        }
      }
      s = (s + 1) & cap_mask;
    }
  }
}

// cnt[e] becomes the survivor flag: 1 for a representative's entry that all other members hit.
__global__ void __launch_bounds__(256) cons_flag_kernel(const uint64_t* __restrict__ tree_off,
                                                        const uint32_t* __restrict__ labels,
                                                        const uint32_t* __restrict__ size, uint32_t T,
                                                        uint32_t* __restrict__ cnt) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const bool rep = labels[t] == t;
  const uint32_t need = rep ? size[t] - 1 : 0;
  for (uint64_t e = tree_off[t] + lane_id(); e < tree_off[t + 1]; e += 32) cnt[e] = (rep && cnt[e] == need) ? 1u : 0u;
}

// One warp per tree (t == T writes the closing offset): survivors to their packed place.
__global__ void __launch_bounds__(256) cons_gather_kernel(const uint64_t* __restrict__ bips,
                                                          const uint64_t* __restrict__ tree_off,
                                                          const uint32_t* __restrict__ flag,
                                                          const uint32_t* __restrict__ pos, uint32_t T, uint32_t W,
                                                          uint64_t* __restrict__ out, uint64_t* __restrict__ out_off) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t > T) return;
  const uint64_t e0 = tree_off[t];
  if (lane_id() == 0) out_off[t] = pos[e0];
  if (t == T) return;
  for (uint64_t e = e0 + lane_id(); e < tree_off[t + 1]; e += 32) {
    if (!flag[e]) continue;
    const uint64_t* src = bips + e * W;
    uint64_t* dst = out + (uint64_t)pos[e] * W;
    for (uint32_t w = 0; w < W; ++w) dst[w] = src[w];
  }
}

template <int W>
int launch_cons_table_w(Ctx* ctx, const uint64_t* bips, const uint64_t* tree_off, const uint32_t* labels,
                        const uint32_t* size, uint32_t T, unsigned long long* slots, uint64_t cap_mask, uint32_t* cnt) {
  const uint32_t grid = (T + 7) / 8;
  cons_table_kernel<W, false><<<grid, 256, 0, ctx->stream>>>(bips, tree_off, labels, size, T, slots, cap_mask, cnt);
  PRF_LAUNCHED(ctx);
  cons_table_kernel<W, true><<<grid, 256, 0, ctx->stream>>>(bips, tree_off, labels, size, T, slots, cap_mask, cnt);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
