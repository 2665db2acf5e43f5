// prf_shard.cuh -- multi-GPU build of the hashRF incidence (RFDistance.hs:289-297) without
// replicating the dedup on every rank.
//
// Bipartitions are partitioned by a hash of the key: rank r deduplicates only the keys of hash class
// r.  All occurrences of a key land in one class, so each rank's ordinary single-GPU build
// (prf_incidence.cuh, unchanged) over "every tree restricted to class r" yields a complete partial
// structure (|B_t|_r, inverted lists, per-tree suffix ranges), and
//     d(a,b) = sum_r ( |B_a|_r + |B_b|_r - 2 |B_a n B_b|_r ).
// The partial structures are small (~80 MB in total at 100k x 200); after an all-gather every rank
// MERGES them at the range level -- tree a's ranges are the union of its ranges in every class, the
// member arrays are concatenated -- and the matrix kernels run unchanged on the merged structure.
//
//   sender   part_owner_kernel   hash class of every local entry + per-(class, tree) counts
//            one stable radix pass by class, gather_keys_kernel -> send buffer grouped by class
//   host     all-to-all of keys and counts (NCCL), prefix sum -> per-class tree offsets
//   receiver prf_hashrf_load on the received keys (device memory)                  [existing code]
//   host     all-gather of the partial structures
//   every    merge_* kernels -> merged structure installed as the context's loaded problem
#pragma once

#include "prf_primitives.cuh"

namespace prf {

// One warp per local tree.  owner[e] = hash class of entry e; counts[c * T_local + t] = entries of
// tree t in class c.
template <int W>
__global__ void __launch_bounds__(256) part_owner_kernel(const uint64_t* __restrict__ bips,
                                                         const uint64_t* __restrict__ tree_off, uint32_t T_local,
                                                         uint32_t world, uint32_t* __restrict__ owner,
                                                         uint32_t* __restrict__ idx, uint32_t* __restrict__ counts) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T_local) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  uint32_t my_count = 0;  // lane c accumulates the count of class c (world <= 32)
  for (uint64_t eb = e0; eb < e1; eb += 32) {
    const uint64_t e = eb + lane_id();
    uint32_t c = 0xffffffffu;
    if (e < e1) {
      uint64_t k[W];
      load_key<W>(bips, e, k);
      uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
      for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
      c = (uint32_t)((mix64(h ^ 0xA5A5A5A5A5A5A5A5ull) >> 33) % world);  // independent of the dedup table's slot/tag bits
      owner[e] = c;
      idx[e] = (uint32_t)e;
    }
    for (uint32_t cls = 0; cls < world; ++cls) {
      const uint32_t n = __popc(__ballot_sync(0xffffffffu, c == cls));
      if (lane_id() == cls) my_count += n;
    }
  }
  if (lane_id() < world) counts[(size_t)lane_id() * T_local + t] = my_count;
}

// send[i] = bips[idx[i]] (W words each); one thread per word.
__global__ void __launch_bounds__(256) gather_keys_kernel(const uint64_t* __restrict__ bips,
                                                          const uint32_t* __restrict__ idx, uint64_t n_words, uint32_t W,
                                                          uint64_t* __restrict__ send) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_words) return;
  const uint64_t e = i / W, w = i - e * W;
  send[i] = bips[(uint64_t)idx[e] * W + w];
}

// first[c] = first position of class c in the sorted class array (first[world] = n).
__global__ void class_bounds_kernel(const uint32_t* __restrict__ sorted_owner, uint32_t n, uint32_t world,
                                    unsigned long long* __restrict__ first) {
  const uint32_t c = threadIdx.x;
  if (c > world) return;
  uint32_t lo = 0, hi = n;  // first index with owner >= c
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (sorted_owner[mid] < c) lo = mid + 1; else hi = mid;
  }
  first[c] = lo;
}

// ---- merge of the gathered partial structures -------------------------------------------------------

struct MergeParts {           // device arrays, one entry per rank (world <= 16)
  const uint16_t* nb[16];
  const uint32_t* roff[16];
  const uint2* ranges[16];
  const uint32_t* members[16];
  const uint32_t* lid[16];
  uint32_t member_base[16];   // position of rank r's members in the concatenated array
  uint32_t list_base[16];     // first dense list id of rank r
  uint32_t n_members[16];
  uint32_t world;
};

// nb_all[t] = sum_r nb_r[t]; cnt[t] = ranges of tree t over all ranks.
__global__ void __launch_bounds__(256) merge_count_kernel(const MergeParts P, uint32_t T, uint32_t* __restrict__ nb32,
                                                          uint32_t* __restrict__ cnt) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > T) return;
  uint32_t nb = 0, c = 0;
  if (t < T)
    for (uint32_t r = 0; r < P.world; ++r) {
      nb += P.nb[r][t];
      c += P.roff[r][t + 1] - P.roff[r][t];
    }
  if (t < T) nb32[t] = nb;
  cnt[t] = c;  // cnt[T] = 0: the scan's last element becomes the total
}

// One thread per (rank, tree): copies the tree's ranges of that rank behind those of the ranks before it.
__global__ void __launch_bounds__(256) merge_ranges_kernel(const MergeParts P, uint32_t T,
                                                           const uint32_t* __restrict__ roff_all,
                                                           uint2* __restrict__ ranges_all) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint64_t)P.world * T) return;
  const uint32_t r = (uint32_t)(i / T), t = (uint32_t)(i - (uint64_t)r * T);
  uint32_t dst = roff_all[t];
  for (uint32_t q = 0; q < r; ++q) dst += P.roff[q][t + 1] - P.roff[q][t];
  const uint32_t k0 = P.roff[r][t], k1 = P.roff[r][t + 1], base = P.member_base[r];
  for (uint32_t k = k0; k < k1; ++k) {
    const uint2 v = P.ranges[r][k];
    ranges_all[dst + (k - k0)] = make_uint2(v.x + base, v.y + base);
  }
}

// members_all: concatenation of the parts' padded list arrays.
__global__ void __launch_bounds__(256) merge_members_kernel(const MergeParts P, uint32_t r,
                                                            uint32_t* __restrict__ members_all) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= P.n_members[r]) return;
  members_all[P.member_base[r] + k] = P.members[r][k];
}

}  // namespace prf
