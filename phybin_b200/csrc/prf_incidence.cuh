// prf_incidence.cuh -- the `build` phase of hashRF (RFDistance.hs:289-297) on the device:
//   K1  global bipartition dedup: open-addressing hash table over the W-word bit sets with exact
//       full-key verification (a 64-bit hash only routes; equality is always decided on the key);
//   K2  trees x unique-bipartitions incidence in the two forms the matrix kernels want:
//       inverted lists (bipartition -> ascending tree ids) for bipartitions present in >= 2 trees
//       -- a bipartition seen in one tree can never be shared, it only counts in |B_t| -- and,
//       per tree, the list suffixes "members after me" that the sparse kernel walks.
// The reference's Map DenseLabelSet (IntSet of tree ids) (:289-297) is exactly this inverted index.
#pragma once

#include "prf_primitives.cuh"

namespace prf {

constexpr uint64_t kEmptySlot = ~0ull;

template <int W>
__device__ __forceinline__ void load_key(const uint64_t* __restrict__ bips, uint64_t e, uint64_t (&k)[W]) {
  const uint64_t* p = bips + e * W;
  if constexpr (W % 2 == 0) {
#pragma unroll
    for (int w = 0; w < W; w += 2) {
      ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p + w);
      k[w] = q.x;
      k[w + 1] = q.y;
    }
  } else {
#pragma unroll
    for (int w = 0; w < W; ++w) k[w] = p[w];
  }
}

// One warp per tree; lanes stride over the tree's entries.  Finds (or claims) the slot of every
// key, records it, and counts the slot's multiplicity.
//   slots[s]  = (tag << 32) | representative entry, kEmptySlot when free
//   cnt[s]    = number of entries that landed in s
template <int W>
__global__ void __launch_bounds__(256) dedup_insert_kernel(const uint64_t* __restrict__ bips,
                                                           const uint64_t* __restrict__ tree_off,
                                                           uint32_t T, unsigned long long* slots,
                                                           uint32_t* cnt, uint64_t cap_mask,
                                                           uint32_t* __restrict__ ent_slot,
                                                           uint32_t* __restrict__ ent_tree) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  for (uint64_t e = e0 + lane_id(); e < e1; e += 32) {
    uint64_t k[W];
    load_key<W>(bips, e, k);
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
    const uint64_t tag = h >> 32;
    const unsigned long long mine = (tag << 32) | (uint64_t)(uint32_t)e;
    uint64_t s = h & cap_mask;
    for (;;) {
      unsigned long long cur = slots[s];
      if (cur == kEmptySlot) {
        cur = atomicCAS(&slots[s], (unsigned long long)kEmptySlot, mine);
        if (cur == kEmptySlot) break;  // claimed: this entry is the representative
      }
      if ((cur >> 32) == tag) {  // same tag: decide on the full key
        uint64_t r[W];
        load_key<W>(bips, cur & 0xffffffffull, r);
        bool eq = true;
#pragma unroll
        for (int w = 0; w < W; ++w) eq &= (r[w] == k[w]);
        if (eq) break;
      }
      s = (s + 1) & cap_mask;
    }
    ent_slot[e] = (uint32_t)s;
    ent_tree[e] = t;
    atomicAdd(&cnt[s], 1u);
  }
}

struct SlotShared {  // 1 when the slot's bipartition occurs in >= 2 trees
  const uint32_t* cnt;
  __device__ uint32_t operator()(uint64_t i) const { return cnt[i] >= 2u; }
};
struct EntryShared {
  const uint32_t* cnt;
  const uint32_t* ent_slot;
  __device__ uint32_t operator()(uint64_t e) const { return cnt[ent_slot[e]] >= 2u; }
};

// Table statistics: number of occupied slots (U), sum m(m-1)/2 (shared-split hits).
__global__ void __launch_bounds__(256) slot_stats_kernel(const uint32_t* __restrict__ cnt, uint64_t cap,
                                                         unsigned long long* out /* [2] */) {
  unsigned long long u = 0, hits = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap;
       i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t m = cnt[i];
    u += m > 0;
    hits += m * (m - (m > 0)) / 2;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    u += __shfl_xor_sync(0xffffffffu, u, d);
    hits += __shfl_xor_sync(0xffffffffu, hits, d);
  }
  if (lane_id() == 0) {
    if (u) atomicAdd(&out[0], u);
    if (hits) atomicAdd(&out[1], hits);
  }
}

// For shared slots: list_cnt[sid] = multiplicity.
__global__ void __launch_bounds__(256) list_count_kernel(const uint32_t* __restrict__ cnt,
                                                         const uint32_t* __restrict__ sid_of_slot,
                                                         uint64_t cap, uint32_t* __restrict__ list_cnt) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap && cnt[i] >= 2u) list_cnt[sid_of_slot[i]] = cnt[i];
}

// Compacts the entries of shared bipartitions, keeping tree order: (sid, tree) pairs.
__global__ void __launch_bounds__(256) gather_shared_kernel(const uint32_t* __restrict__ cnt,
                                                            const uint32_t* __restrict__ ent_slot,
                                                            const uint32_t* __restrict__ ent_tree,
                                                            const uint32_t* __restrict__ ent_pos,
                                                            const uint32_t* __restrict__ sid_of_slot,
                                                            uint64_t nnz, uint32_t* __restrict__ sid,
                                                            uint32_t* __restrict__ tree) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  uint32_t s = ent_slot[e];
  if (cnt[s] >= 2u) {
    uint32_t k = ent_pos[e];
    sid[k] = sid_of_slot[s];
    tree[k] = ent_tree[e];
  }
}

// Flags sorted (sid, tree) pairs that repeat their predecessor: a bipartition listed twice in one
// tree.  A tree's bipartitions are a set, so those are dropped.
struct DupFlag {
  const uint32_t* sid;
  const uint32_t* tree;
  __device__ uint32_t operator()(uint64_t k) const {
    return k > 0 && sid[k] == sid[k - 1] && tree[k] == tree[k - 1];
  }
};
__global__ void __launch_bounds__(256) drop_dups_kernel(const uint32_t* __restrict__ sid,
                                                        const uint32_t* __restrict__ tree,
                                                        const uint32_t* __restrict__ dup_before,
                                                        uint32_t n, uint32_t* __restrict__ sid_out,
                                                        uint32_t* __restrict__ tree_out,
                                                        uint32_t* list_cnt, uint32_t* nb32) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  bool dup = k > 0 && sid[k] == sid[k - 1] && tree[k] == tree[k - 1];
  if (dup) {
    atomicSub(&list_cnt[sid[k]], 1u);
    atomicSub(&nb32[tree[k]], 1u);
  } else {
    uint32_t o = k - dup_before[k];
    sid_out[o] = sid[k];
    tree_out[o] = tree[k];
  }
}

__global__ void __launch_bounds__(256) tree_counts_kernel(const uint64_t* __restrict__ tree_off, uint32_t T,
                                                          uint32_t* __restrict__ nb32) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) nb32[t] = (uint32_t)(tree_off[t + 1] - tree_off[t]);
}

// nb32 -> uint16 + min/max.  mm[0] = min, mm[1] = max.
__global__ void __launch_bounds__(256) finish_nb_kernel(const uint32_t* __restrict__ nb32, uint32_t T,
                                                        uint16_t* __restrict__ nb16, uint32_t* mm) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = t < T ? nb32[t] : 0, lo = t < T ? v : 0xffffffffu, hi = v;
  if (t < T) nb16[t] = (uint16_t)(v > 0xffffu ? 0xffffu : v);
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
  }
  if (lane_id() == 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
  }
}

// Sorted position k holds tree a = members[k] inside list sid[k] = [list_off, list_off_next).
// Its suffix (k+1, end) lists the trees b > a sharing that bipartition.  Pass 1 counts the
// non-empty suffixes per tree, pass 2 stores them.
__global__ void __launch_bounds__(256) range_count_kernel(const uint32_t* __restrict__ sid,
                                                          const uint32_t* __restrict__ members,
                                                          const uint32_t* __restrict__ list_off,
                                                          uint32_t n, uint32_t* rcount) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (k + 1 < list_off[sid[k] + 1]) atomicAdd(&rcount[members[k]], 1u);
}
__global__ void __launch_bounds__(256) range_fill_kernel(const uint32_t* __restrict__ sid,
                                                         const uint32_t* __restrict__ members,
                                                         const uint32_t* __restrict__ list_off,
                                                         uint32_t n, const uint32_t* __restrict__ roff,
                                                         uint32_t* cursor, uint2* __restrict__ ranges) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t end = list_off[sid[k] + 1];
  if (k + 1 < end) {
    uint32_t a = members[k];
    uint32_t slot = roff[a] + atomicAdd(&cursor[a], 1u);
    ranges[slot] = make_uint2(k + 1, end);
  }
}

__global__ void fill_u64_kernel(unsigned long long* p, uint64_t n, unsigned long long v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

template <int W>
static void launch_dedup(Ctx* ctx, const uint64_t* bips, const uint64_t* off, uint32_t T,
                         unsigned long long* slots, uint32_t* cnt, uint64_t cap_mask,
                         uint32_t* ent_slot, uint32_t* ent_tree) {
  dedup_insert_kernel<W><<<(T + 7) / 8, 256, 0, ctx->stream>>>(bips, off, T, slots, cnt, cap_mask,
                                                                ent_slot, ent_tree);
}

static inline uint32_t grid_for(uint64_t n, uint32_t tpb = 256) { return (uint32_t)((n + tpb - 1) / tpb); }

static inline int bits_for(uint32_t n_values) {  // bits needed to represent 0..n_values-1
  int b = 0;
  while (b < 32 && (1ull << b) < n_values) ++b;
  return b;
}

// d_bips: nnz*W words, d_off: T+1 offsets, both already on the device.
static int hashrf_build(Ctx* ctx, const uint64_t* d_bips, const uint64_t* d_off, uint32_t T, uint32_t W,
                        uint64_t nnz) {
  cudaStream_t s = ctx->stream;
  HashRFState& H = ctx->h;
  H = HashRFState{};
  H.T = T;
  H.W = W;
  H.nnz = nnz;
  H.P = (uint64_t)T * (T ? T - 1 : 0) / 2;

  uint64_t cap = 1024;
  while (cap < 2 * nnz) cap <<= 1;

  DevBuf<unsigned long long> slots, acc;
  DevBuf<uint32_t> cnt, ent_slot, ent_tree, nb32, scalars;
  PRF_CK(ctx, slots.alloc(cap, s));
  PRF_CK(ctx, cnt.alloc(cap, s));
  PRF_CK(ctx, ent_slot.alloc(nnz, s));
  PRF_CK(ctx, ent_tree.alloc(nnz, s));
  PRF_CK(ctx, nb32.alloc(T, s));
  PRF_CK(ctx, acc.alloc(2, s));
  PRF_CK(ctx, scalars.alloc(8, s));
  PRF_CK(ctx, cudaMemsetAsync(cnt.p, 0, cnt.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(acc.p, 0, acc.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(scalars.p, 0, scalars.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(slots.p, 0xff, slots.bytes(), s));

  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  // ---- K1: dedup ----
  if (T) {
    switch (W) {
#define PRF_W_CASE(w) case w: launch_dedup<w>(ctx, d_bips, d_off, T, slots.p, cnt.p, cap - 1, ent_slot.p, ent_tree.p); break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7)
      PRF_W_CASE(8) PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13)
      PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
      default: return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
    }
    PRF_LAUNCHED(ctx);
    tree_counts_kernel<<<grid_for(T), 256, 0, s>>>(d_off, T, nb32.p);
    PRF_LAUNCHED(ctx);
  }
  slot_stats_kernel<<<std::min<uint32_t>(grid_for(cap), 148 * 8), 256, 0, s>>>(cnt.p, cap, acc.p);
  PRF_LAUNCHED(ctx);
  PRF_CK(ctx, cudaEventRecord(ctx->ev[2], s));

  // ---- K2: incidence ----
  // shared slots -> dense ids
  DevBuf<uint32_t> sid_of_slot, ent_pos;
  PRF_CK(ctx, sid_of_slot.alloc(cap, s));
  PRF_CK(ctx, ent_pos.alloc(nnz, s));
  int rc;
  if ((rc = exclusive_scan(ctx, SlotShared{cnt.p}, cap, sid_of_slot.p, scalars.p + 0))) return rc;
  if ((rc = exclusive_scan(ctx, EntryShared{cnt.p, ent_slot.p}, nnz, ent_pos.p, scalars.p + 1))) return rc;
  uint32_t hs[8];
  unsigned long long hacc[2];
  PRF_CK(ctx, cudaMemcpyAsync(hs, scalars.p, sizeof hs, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaMemcpyAsync(hacc, acc.p, sizeof hacc, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  uint32_t n_shared = hs[0], n_sh_nnz = hs[1];
  H.n_shared = n_shared;
  H.n_shared_nnz = n_sh_nnz;

  DevBuf<uint32_t> list_cnt, list_off;
  PRF_CK(ctx, list_cnt.alloc((size_t)n_shared + 1, s));
  PRF_CK(ctx, list_off.alloc((size_t)n_shared + 1, s));
  PRF_CK(ctx, cudaMemsetAsync(list_cnt.p, 0, list_cnt.bytes(), s));
  PRF_CK(ctx, H.sh_sid.alloc(n_sh_nnz, s));
  PRF_CK(ctx, H.sh_tree.alloc(n_sh_nnz, s));
  DevBuf<uint32_t> sid_sorted;
  PRF_CK(ctx, sid_sorted.alloc(n_sh_nnz, s));
  PRF_CK(ctx, H.members.alloc(n_sh_nnz, s));
  uint32_t n_sorted = n_sh_nnz;
  if (n_sh_nnz) {
    list_count_kernel<<<grid_for(cap), 256, 0, s>>>(cnt.p, sid_of_slot.p, cap, list_cnt.p);
    PRF_LAUNCHED(ctx);
    gather_shared_kernel<<<grid_for(nnz), 256, 0, s>>>(cnt.p, ent_slot.p, ent_tree.p, ent_pos.p,
                                                       sid_of_slot.p, nnz, H.sh_sid.p, H.sh_tree.p);
    PRF_LAUNCHED(ctx);
    PRF_CK(ctx, cudaMemcpyAsync(sid_sorted.p, H.sh_sid.p, (size_t)n_sh_nnz * 4, cudaMemcpyDeviceToDevice, s));
    PRF_CK(ctx, cudaMemcpyAsync(H.members.p, H.sh_tree.p, (size_t)n_sh_nnz * 4, cudaMemcpyDeviceToDevice, s));
    // stable sort by bipartition id: inside a list the trees stay ascending
    if ((rc = radix_sort_pairs(ctx, sid_sorted.p, H.members.p, n_sh_nnz, bits_for(n_shared)))) return rc;
    // a bipartition listed twice by one tree counts once
    DevBuf<uint32_t> dup_before;
    PRF_CK(ctx, dup_before.alloc(n_sh_nnz, s));
    if ((rc = exclusive_scan(ctx, DupFlag{sid_sorted.p, H.members.p}, n_sh_nnz, dup_before.p, scalars.p + 2)))
      return rc;
    uint32_t n_dups = 0;
    PRF_CK(ctx, cudaMemcpyAsync(&n_dups, scalars.p + 2, 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    if (n_dups) {
      DevBuf<uint32_t> sid2, mem2;
      PRF_CK(ctx, sid2.alloc(n_sh_nnz - n_dups, s));
      PRF_CK(ctx, mem2.alloc(n_sh_nnz - n_dups, s));
      drop_dups_kernel<<<grid_for(n_sh_nnz), 256, 0, s>>>(sid_sorted.p, H.members.p, dup_before.p, n_sh_nnz,
                                                          sid2.p, mem2.p, list_cnt.p, nb32.p);
      PRF_LAUNCHED(ctx);
      sid_sorted = std::move(sid2);
      H.members = std::move(mem2);
      n_sorted = n_sh_nnz - n_dups;
    }
  }
  if ((rc = exclusive_scan(ctx, LoadU32{list_cnt.p}, (uint64_t)n_shared + 1, list_off.p, nullptr))) return rc;

  // per-tree suffix ranges
  DevBuf<uint32_t> rcount, cursor;
  PRF_CK(ctx, rcount.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cursor.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.roff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cudaMemsetAsync(rcount.p, 0, rcount.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(cursor.p, 0, cursor.bytes(), s));
  if (n_sorted) {
    range_count_kernel<<<grid_for(n_sorted), 256, 0, s>>>(sid_sorted.p, H.members.p, list_off.p, n_sorted, rcount.p);
    PRF_LAUNCHED(ctx);
  }
  if ((rc = exclusive_scan(ctx, LoadU32{rcount.p}, (uint64_t)T + 1, H.roff.p, scalars.p + 3))) return rc;
  uint32_t n_ranges = 0;
  PRF_CK(ctx, cudaMemcpyAsync(&n_ranges, scalars.p + 3, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  PRF_CK(ctx, H.ranges.alloc(n_ranges, s));
  if (n_ranges) {
    range_fill_kernel<<<grid_for(n_sorted), 256, 0, s>>>(sid_sorted.p, H.members.p, list_off.p, n_sorted,
                                                         H.roff.p, cursor.p, H.ranges.p);
    PRF_LAUNCHED(ctx);
  }

  // |B_t| as uint16 + min/max
  PRF_CK(ctx, H.nb.alloc(T, s));
  uint32_t mm_init[2] = {0xffffffffu, 0u};
  PRF_CK(ctx, cudaMemcpyAsync(scalars.p + 4, mm_init, 8, cudaMemcpyHostToDevice, s));
  if (T) {
    finish_nb_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, T, H.nb.p, scalars.p + 4);
    PRF_LAUNCHED(ctx);
  }
  uint32_t mm[2];
  PRF_CK(ctx, cudaMemcpyAsync(mm, scalars.p + 4, 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (T == 0) mm[0] = mm[1] = 0;
  if (mm[1] > kMaxBipsPerTree)
    return ctx->fail(PRF_E_RANGE, "a tree has %u bipartitions; distances would overflow uint16 (max %u per tree)",
                     mm[1], kMaxBipsPerTree);
  H.uniform_nb = mm[0] == mm[1];
  H.nb_uniform = mm[0];

  prf_stats& st = H.stats;
  st = prf_stats{};
  st.n_trees = T;
  st.n_words = W;
  st.nnz = nnz;
  st.n_unique = hacc[0];
  st.n_shared = n_shared;
  st.n_shared_nnz = n_sh_nnz;
  st.n_hits = hacc[1];
  st.n_pairs = H.P;
  st.max_bips = mm[1];
  st.min_bips = mm[0];
  H.loaded = true;
  return PRF_OK;
}

}  // namespace prf
