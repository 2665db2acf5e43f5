// prf_incidence.cuh -- the `build` phase of hashRF (RFDistance.hs:289-297) on the device:
//   K1  global bipartition dedup: open-addressing hash table over the W-word bit sets with exact
//       full-key verification (a 64-bit hash only routes; equality is always decided on the key).
//       A bit per slot records "seen by a second entry": bipartitions present in one tree only
//       (~3/4 of them on random topologies) can never be shared, they only count in |B_t|.
//   K2  trees x unique-bipartitions incidence in the form the matrix kernels want: the entries of
//       shared bipartitions, stably radix-sorted by table slot, are the inverted lists
//       (bipartition -> ascending tree ids); per tree, the list suffixes "members after me" are
//       what the sparse kernel walks.
// The reference's Map DenseLabelSet (IntSet of tree ids) (:289-297) is exactly this inverted index.
#pragma once

#include "prf_primitives.cuh"

namespace prf {

// counters[0] = unique keys (claims), [1] = shared lists, [2] = duplicate (slot, tree) entries
// One warp per tree; lanes stride over the tree's entries.  Finds (or claims) the slot of every key.
//   slots[s]  = (tag << 32) | representative entry, kEmptySlot when free
//   shared[s/32] bit s%32 = some entry other than the representative landed in s
template <int W>
__global__ void __launch_bounds__(256) dedup_insert_kernel(const uint64_t* __restrict__ bips,
                                                           const uint64_t* __restrict__ tree_off,
                                                           uint32_t T, unsigned long long* slots,
                                                           uint32_t* shared_bits, uint64_t cap_mask,
                                                           uint32_t* __restrict__ ent_slot,
                                                           unsigned long long* counters) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  uint32_t claims = 0;
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
        if (cur == kEmptySlot) { ++claims; break; }  // claimed: this entry is the representative
      }
      if ((cur >> 32) == tag) {  // same tag: decide on the full key
        uint64_t r[W];
        load_key<W>(bips, cur & 0xffffffffull, r);
        bool eq = true;
#pragma unroll
        for (int w = 0; w < W; ++w) eq &= (r[w] == k[w]);
        if (eq) {
          const uint32_t bit = 1u << (s & 31);
          if (!(shared_bits[s >> 5] & bit)) atomicOr(&shared_bits[s >> 5], bit);
          break;
        }
      }
      s = (s + 1) & cap_mask;
    }
    ent_slot[e] = (uint32_t)s;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) claims += __shfl_xor_sync(0xffffffffu, claims, d);
  if (lane_id() == 0 && claims) atomicAdd(&counters[0], (unsigned long long)claims);
}

// Compact variant for nnz < 2^25 - 1 entries: 4-byte slots ( tag:7 | representative entry + 1 : 25, 0 = free )
// and a table of 1.25 * nnz slots addressed by multiply-shift instead of a power-of-two mask.  At the
// headline shape the table is 98 MB instead of 268 MB: it stays in the 126 MB L2 (the key stream is
// loaded with an evict-first policy), so the random probes and CAS are L2 hits instead of DRAM sectors.
template <int W>
__global__ void __launch_bounds__(256) dedup_insert32_kernel(const uint64_t* __restrict__ bips,
                                                             const uint64_t* __restrict__ tree_off, uint32_t T,
                                                             uint32_t* slots, uint32_t* shared_bits, uint32_t cap,
                                                             uint32_t* __restrict__ ent_slot,
                                                             unsigned long long* counters) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  uint32_t claims = 0;
  for (uint64_t e = e0 + lane_id(); e < e1; e += 32) {
    uint64_t k[W];
    {
      const uint64_t* p = bips + e * W;
#pragma unroll
      for (int w = 0; w < W; ++w)
        asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(k[w]) : "l"(p + w), "l"(policy));
    }
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
    const uint32_t tag = (uint32_t)(h >> 57);  // 7 bits, independent of the slot bits below
    const uint32_t mine = (tag << 25) | ((uint32_t)e + 1u);
    uint32_t s = (uint32_t)(((h & 0xffffffffull) * (uint64_t)cap) >> 32);
    for (;;) {
      uint32_t cur = slots[s];
      if (cur == 0u) {
        cur = atomicCAS(&slots[s], 0u, mine);
        if (cur == 0u) { ++claims; break; }  // claimed: this entry is the representative
      }
      if ((cur >> 25) == tag) {  // same tag: decide on the full key
        uint64_t r[W];
        load_key<W>(bips, (cur & 0x1ffffffu) - 1u, r);
        bool eq = true;
#pragma unroll
        for (int w = 0; w < W; ++w) eq &= (r[w] == k[w]);
        if (eq) {
          const uint32_t bit = 1u << (s & 31);
          if (!(shared_bits[s >> 5] & bit)) atomicOr(&shared_bits[s >> 5], bit);
          break;
        }
      }
      s = s + 1 == cap ? 0u : s + 1;
    }
    ent_slot[e] = s;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) claims += __shfl_xor_sync(0xffffffffu, claims, d);
  if (lane_id() == 0 && claims) atomicAdd(&counters[0], (unsigned long long)claims);
}

// Per tree: number of entries whose bipartition is shared; also |B_t| before duplicate removal.
__global__ void __launch_bounds__(256) count_shared_kernel(const uint64_t* __restrict__ tree_off, uint32_t T,
                                                           const uint32_t* __restrict__ ent_slot,
                                                           const uint32_t* __restrict__ shared_bits,
                                                           uint32_t* __restrict__ sc, uint32_t* __restrict__ nb32) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  uint32_t c = 0;
  for (uint64_t e = e0 + lane_id(); e < e1; e += 32) {
    const uint32_t s = ent_slot[e];
    c += (shared_bits[s >> 5] >> (s & 31)) & 1u;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane_id() == 0) {
    sc[t] = c;
    nb32[t] = (uint32_t)(e1 - e0);
  }
}

// Number of shared slots in one 32-slot word of the bitmap (scanned into per-word ranks).
struct PopcWord {
  const uint32_t* bits;
  __device__ uint32_t operator()(uint64_t w) const { return __popc(bits[w]); }
};

// Writes the (list id, tree) pairs of the shared entries in tree order (stable warp compaction).  The
// list id is the rank of the entry's slot among the shared slots (per-word rank + popcount below the
// slot's bit): dense ids in slot order, so two 10-bit radix passes sort them and the sorted key array
// IS the per-entry list id.
__global__ void __launch_bounds__(256) gather_shared_kernel(const uint64_t* __restrict__ tree_off, uint32_t T,
                                                            const uint32_t* __restrict__ ent_slot,
                                                            const uint32_t* __restrict__ shared_bits,
                                                            const uint32_t* __restrict__ wrank,
                                                            const uint32_t* __restrict__ soff,
                                                            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint64_t e0 = tree_off[t], e1 = tree_off[t + 1];
  uint32_t pos = soff[t];
  const uint32_t lt = lanemask_lt();
  for (uint64_t eb = e0; eb < e1; eb += 32) {
    const uint64_t e = eb + lane_id();
    uint32_t id = 0;
    bool sh = false;
    if (e < e1) {
      const uint32_t s = ent_slot[e];
      const uint32_t word = shared_bits[s >> 5];
      sh = (word >> (s & 31)) & 1u;
      id = wrank[s >> 5] + __popc(word & ((1u << (s & 31)) - 1u));
    }
    const uint32_t m = __ballot_sync(0xffffffffu, sh);
    if (sh) {
      const uint32_t o = pos + __popc(m & lt);
      keys[o] = id;
      vals[o] = t;
    }
    pos += __popc(m);
  }
}

// Sorted dense ids (every id 0..L-1 occurs): head_pos[l] = first position of list l, head_pos[L] = n.
__global__ void __launch_bounds__(256) id_heads_kernel(const uint32_t* __restrict__ lid, uint32_t n, uint32_t n_lists,
                                                       uint32_t* __restrict__ head_pos) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  if (k == 0 || lid[k] != lid[k - 1]) head_pos[lid[k]] = k;
  if (k == n - 1) head_pos[n_lists] = n;
}

// Sorted (slot, tree) pairs: a list starts where the slot changes.
struct HeadFlag {
  const uint32_t* slot;
  __device__ uint32_t operator()(uint64_t k) const { return k == 0 || slot[k] != slot[k - 1]; }
};
// A pair that repeats its predecessor is a bipartition listed twice by one tree.
struct DupFlag {
  const uint32_t* slot;
  const uint32_t* tree;
  __device__ uint32_t operator()(uint64_t k) const {
    return k > 0 && slot[k] == slot[k - 1] && tree[k] == tree[k - 1];
  }
};

// head_pos[l] = first sorted position of list l; head_pos[L] = n.  Also raises counters[2] when
// a duplicate pair is seen.
__global__ void __launch_bounds__(256) list_heads_kernel(const uint32_t* __restrict__ slot,
                                                         const uint32_t* __restrict__ tree,
                                                         const uint32_t* __restrict__ lid_ex, uint32_t n,
                                                         uint32_t* __restrict__ head_pos,
                                                         unsigned long long* counters, uint32_t* __restrict__ lid_out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const bool head = k == 0 || slot[k] != slot[k - 1];
  lid_out[k] = lid_ex[k] + (head ? 1 : 0) - 1;
  if (head) head_pos[lid_ex[k]] = k;
  if (k == n - 1) head_pos[lid_ex[k] + (head ? 1 : 0)] = n;
}

// Position k holds tree a = tree[k] inside list l = [head_pos[l], head_pos[l+1]).  Its suffix
// (k+1, end) lists the trees b > a sharing that bipartition.  Pass 1 counts the non-empty suffixes
// per tree (and accumulates sum m(m-1)/2 at list heads), pass 2 stores them.
//
// Lists of at most `short_max` members are not walked by the matrix kernel: their pairs are expanded into
// per-row "singles" (tree a gets the column b of every later member), counted in scount / stored by the
// fill pass.  Only the longer lists produce suffix ranges.
__global__ void __launch_bounds__(256) range_count_kernel(const uint32_t* __restrict__ lid,
                                                          const uint32_t* __restrict__ tree,
                                                          const uint32_t* __restrict__ head_pos, uint32_t n,
                                                          uint32_t short_max, uint32_t heavy_min, uint32_t* rcount,
                                                          uint32_t* hcount, uint32_t* scount,
                                                          unsigned long long* counters) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long hits = 0, longest = 0;
  if (k < n) {
    const uint32_t l = lid[k];
    const uint32_t end = head_pos[l + 1];
    if (k + 1 < end) {
      const uint32_t m = end - head_pos[l];
      if (m <= short_max) {
        atomicAdd(&scount[tree[k]], end - k - 1);
      } else {
        atomicAdd(&rcount[tree[k]], 1u);
        if (m >= heavy_min) atomicAdd(&hcount[tree[k]], 1u);
      }
    }
    if (k > 0 && lid[k - 1] == l && tree[k - 1] == tree[k]) atomicAdd(&counters[2], 1ull);  // listed twice by one tree (rare)
    if (k == head_pos[l]) {
      const unsigned long long m = end - k;
      hits = m * (m - 1) / 2;
      longest = m;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, d);
    longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, d));
  }
  // one pair of global atomics per block, not per warp: both targets are single hot addresses
  __shared__ unsigned long long s_hits[8], s_long[8];
  if (lane_id() == 0) {
    s_hits[threadIdx.x >> 5] = hits;
    s_long[threadIdx.x >> 5] = longest;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      hits += s_hits[w];
      longest = max(longest, s_long[w]);
    }
    if (hits) atomicAdd(&counters[3], hits);
    if (longest) atomicMax(&counters[4], longest);
  }
}
// Space of list l in the padded array of walked lists: 0 for short lists, else its length rounded up to 4.
struct PadLen {
  const uint32_t* head_pos;
  const uint32_t* n_lists;  // device scalar
  uint32_t short_max;
  __device__ uint32_t operator()(uint64_t l) const {
    if (l >= *n_lists) return 0;
    const uint32_t m = head_pos[l + 1] - head_pos[l];
    return m <= short_max ? 0u : (m + 3u) & ~3u;
  }
};
__global__ void __launch_bounds__(256) range_fill_kernel(const uint32_t* __restrict__ lid,
                                                         const uint32_t* __restrict__ tree,
                                                         const uint32_t* __restrict__ head_pos, uint32_t n,
                                                         uint32_t short_max, uint32_t heavy_min,
                                                         const uint32_t* __restrict__ roff,
                                                         const uint32_t* __restrict__ hcount, uint32_t* hcursor,
                                                         uint32_t* cursor, uint2* __restrict__ ranges,
                                                         const uint32_t* __restrict__ soff, uint32_t* scursor,
                                                         uint32_t* __restrict__ singles,
                                                         const uint32_t* __restrict__ phead,
                                                         uint32_t* __restrict__ lmembers) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t l = lid[k], end = head_pos[l + 1];
  const uint32_t m = end - head_pos[l];
  const uint32_t pk = phead[l] + (k - head_pos[l]), pend = phead[l] + m;  // my position / the list's end in lmembers
  if (m > short_max) {
    lmembers[pk] = tree[k];
    if (k + 1 == end)
      for (uint32_t j = pend; j < ((pend + 3u) & ~3u); ++j) lmembers[j] = 0xffffffffu;  // pad to a multiple of 4
  }
  if (k + 1 < end) {
    const uint32_t a = tree[k];
    if (m <= short_max) {
      uint32_t o = soff[a] + atomicAdd(&scursor[a], end - k - 1);
      for (uint32_t j = k + 1; j < end; ++j) singles[o++] = tree[j];
    } else if (m >= heavy_min) {  // a tree's heavy lists first, its light lists behind them
      ranges[roff[a] + atomicAdd(&hcursor[a], 1u)] = make_uint2(pk + 1, pend);
    } else {
      ranges[roff[a] + hcount[a] + atomicAdd(&cursor[a], 1u)] = make_uint2(pk + 1, pend);
    }
  }
}

// ---- majority flip -----------------------------------------------------------------------------------
// [u in B_a] xor [u in B_b] is unchanged when membership of bipartition u is complemented for EVERY
// tree, so a list held by more than half of the trees can be replaced by the (shorter) list of trees
// that lack it: d(a,b) = |D_a| + |D_b| - 2 |D_a n D_b| with D_t = B_t xor {majority bipartitions}.
// Similar trees (few differences from a common backbone) become a sparse problem again.
struct FlipLen {
  const uint32_t* head_pos;
  uint32_t T, n_lists;
  __device__ uint32_t operator()(uint64_t l) const {
    if (l >= n_lists) return 0;
    const uint32_t m = head_pos[l + 1] - head_pos[l];
    return 2ull * m > T ? T - m : m;
  }
};
// light lists are copied (one thread per entry); members of heavy lists are only counted per tree
__global__ void __launch_bounds__(256) flip_copy_kernel(const uint32_t* __restrict__ lid,
                                                        const uint32_t* __restrict__ tree,
                                                        const uint32_t* __restrict__ head_pos,
                                                        const uint32_t* __restrict__ head2, uint32_t n, uint32_t T,
                                                        uint32_t* __restrict__ tree2, uint32_t* __restrict__ lid2,
                                                        uint32_t* hcnt) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t l = lid[k], h = head_pos[l], m = head_pos[l + 1] - h;
  if (2ull * m > T) {
    atomicAdd(&hcnt[tree[k]], 1u);
  } else {
    const uint32_t o = head2[l] + (k - h);
    tree2[o] = tree[k];
    lid2[o] = l;
  }
}
// heavy lists: one thread per gap between consecutive members writes the trees of the gap
__global__ void __launch_bounds__(256) flip_complement_kernel(const uint32_t* __restrict__ lid,
                                                              const uint32_t* __restrict__ tree,
                                                              const uint32_t* __restrict__ head_pos,
                                                              const uint32_t* __restrict__ head2, uint32_t n, uint32_t T,
                                                              uint32_t* __restrict__ tree2, uint32_t* __restrict__ lid2,
                                                              unsigned long long* counters) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t l = lid[k], h = head_pos[l], e = head_pos[l + 1], m = e - h;
  if (2ull * m <= T) return;
  if (k == h) atomicAdd(&counters[5], 1ull);  // number of heavy lists
  const uint32_t i = k - h;                   // members before me in the list
  const uint32_t base = head2[l];
  // gap before member i: trees (previous member, this member)
  const uint32_t lo = i ? tree[k - 1] + 1 : 0u, hi = tree[k];
  for (uint32_t t = lo; t < hi; ++t) {
    tree2[base + t - i] = t;
    lid2[base + t - i] = l;
  }
  if (k + 1 == e)  // gap after the last member
    for (uint32_t t = tree[k] + 1; t < T; ++t) {
      tree2[base + t - (i + 1)] = t;
      lid2[base + t - (i + 1)] = l;
    }
}
__global__ void __launch_bounds__(256) flip_nb_kernel(uint32_t* nb32, const uint32_t* __restrict__ hcnt, uint32_t T,
                                                      const unsigned long long* counters) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) nb32[t] = nb32[t] + (uint32_t)counters[5] - 2u * hcnt[t];
}

// Slow path (a caller listed a bipartition twice in one tree): drop the repeats.
__global__ void __launch_bounds__(256) drop_dups_kernel(const uint32_t* __restrict__ slot,
                                                        const uint32_t* __restrict__ tree,
                                                        const uint32_t* __restrict__ dup_before, uint32_t n,
                                                        uint32_t* __restrict__ slot_out,
                                                        uint32_t* __restrict__ tree_out, uint32_t* nb32) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const bool dup = k > 0 && slot[k] == slot[k - 1] && tree[k] == tree[k - 1];
  if (dup) {
    atomicSub(&nb32[tree[k]], 1u);
  } else {
    const uint32_t o = k - dup_before[k];
    slot_out[o] = slot[k];
    tree_out[o] = tree[k];
  }
}

// nb_shift[k * stride + j] = nb[j + k] (0 past the end), k = 0..7
__global__ void __launch_bounds__(256) nb_shift_kernel(const uint16_t* __restrict__ nb, uint32_t T, uint32_t stride,
                                                       uint16_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 8 * stride) return;
  const uint32_t k = i / stride, j = i - k * stride;
  out[i] = j + k < T ? nb[j + k] : (uint16_t)0;
}

template <int W>
static void launch_dedup(Ctx* ctx, const uint64_t* bips, const uint64_t* off, uint32_t T,
                         unsigned long long* slots, uint32_t* shared_bits, uint64_t cap_mask,
                         uint32_t* ent_slot, unsigned long long* counters) {
  dedup_insert_kernel<W><<<(T + 7) / 8, 256, 0, ctx->stream>>>(bips, off, T, slots, shared_bits, cap_mask,
                                                                ent_slot, counters);
}


// The shifted copies of |B_t| the sparse kernel's producer bulk-loads (non-uniform |B_t| only).
int build_nb_shift(Ctx* ctx) {
  HashRFState& H = ctx->h;
  if (H.uniform_nb || H.T == 0) return PRF_OK;
  const uint32_t stride = (H.T + 16 + 7) / 8 * 8 + 8;
  PRF_CK(ctx, H.nb_shift.alloc((size_t)8 * stride, ctx->stream));
  nb_shift_kernel<<<(8 * stride + 255) / 256, 256, 0, ctx->stream>>>(H.nb.p, H.T, stride, H.nb_shift.p);
  PRF_LAUNCHED(ctx);
  H.nb_shift_stride = stride;
  return PRF_OK;
}

// d_bips: nnz*W words, d_off: T+1 offsets, both already on the device.
int hashrf_build(Ctx* ctx, const uint64_t* d_bips, const uint64_t* d_off, uint32_t T, uint32_t W,
                        uint64_t nnz) {
  cudaStream_t s = ctx->stream;
  HashRFState& H = ctx->h;
  H = HashRFState{};
  H.T = T;
  H.W = W;
  H.nnz = nnz;
  H.P = (uint64_t)T * (T ? T - 1 : 0) / 2;

  // compact 4-byte slots whenever the entry index fits 25 bits (see dedup_insert32_kernel)
  const bool compact = nnz < (1ull << 25) - 1 && !ctx->debug_wide_slots;
  uint64_t cap = 1024;
  if (compact) {
    static const double factor = [] { const char* e = getenv("PRF_DEDUP_FACTOR"); return e ? atof(e) : 1.25; }();
    cap = std::max<uint64_t>(1024, ((uint64_t)(nnz * factor) + 31) / 32 * 32);  // load factor <= 0.8, multiple of the bitmap word
  } else {
    while (cap < nnz + nnz / 2) cap <<= 1;  // load factor <= 2/3 even if every key is distinct
  }

  DevBuf<unsigned long long> slots, counters;
  DevBuf<uint32_t> slots32;
  DevBuf<uint32_t> shared_bits, ent_slot, nb32, sc, soff, scalars;
  if (compact) PRF_CK(ctx, slots32.alloc(cap, s));
  else PRF_CK(ctx, slots.alloc(cap, s));
  PRF_CK(ctx, shared_bits.alloc(cap / 32, s));
  PRF_CK(ctx, ent_slot.alloc(nnz, s));
  PRF_CK(ctx, nb32.alloc((size_t)T + 1, s));
  PRF_CK(ctx, sc.alloc((size_t)T + 1, s));
  PRF_CK(ctx, soff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, counters.alloc(6, s));
  PRF_CK(ctx, scalars.alloc(8, s));
  PRF_CK(ctx, cudaMemsetAsync(shared_bits.p, 0, shared_bits.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(counters.p, 0, counters.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(sc.p, 0, sc.bytes(), s));
  if (compact) PRF_CK(ctx, cudaMemsetAsync(slots32.p, 0, slots32.bytes(), s));
  else PRF_CK(ctx, cudaMemsetAsync(slots.p, 0xff, slots.bytes(), s));
  uint32_t mm_init[8] = {0, 0, 0, 0, 0xffffffffu, 0u, 0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(scalars.p, mm_init, sizeof mm_init, cudaMemcpyHostToDevice, s));

  trace_mark("build: temporaries allocated");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  // ---- K1: dedup ----
  if (T) {
    switch (W) {
#define PRF_W_CASE(w)                                                                                                  \
  case w:                                                                                                              \
    if (compact)                                                                                                       \
      dedup_insert32_kernel<w><<<(T + 7) / 8, 256, 0, s>>>(d_bips, d_off, T, slots32.p, shared_bits.p, (uint32_t)cap,   \
                                                           ent_slot.p, counters.p);                                    \
    else                                                                                                               \
      launch_dedup<w>(ctx, d_bips, d_off, T, slots.p, shared_bits.p, cap - 1, ent_slot.p, counters.p);                 \
    break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7)
      PRF_W_CASE(8) PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13)
      PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
      default: return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
    }
    PRF_LAUNCHED(ctx);
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[2], s));

  // ---- K2: incidence ----
  int rc;
  uint32_t n_sh_nnz = 0;
  if (T) {
    count_shared_kernel<<<(T + 7) / 8, 256, 0, s>>>(d_off, T, ent_slot.p, shared_bits.p, sc.p, nb32.p);
    PRF_LAUNCHED(ctx);
  }
  if ((rc = exclusive_scan(ctx, LoadU32{sc.p}, (uint64_t)T + 1, soff.p, scalars.p + 0))) return rc;
  // rank of every shared slot = dense list id (in slot order); scalars[1] = number of lists
  DevBuf<uint32_t> wrank;
  PRF_CK(ctx, wrank.alloc(cap / 32, s));
  if ((rc = exclusive_scan(ctx, PopcWord{shared_bits.p}, cap / 32, wrank.p, scalars.p + 1))) return rc;
  uint32_t h01[2] = {0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(h01, scalars.p + 0, 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  trace_mark("build: dedup + count synced");
  n_sh_nnz = h01[0];
  const uint32_t n_lists0 = h01[1];
  H.n_shared_nnz = n_sh_nnz;

  DevBuf<uint32_t> slot_sorted, lid_ex, head_pos, rcount, cursor, scount, scursor, hcursor;
  const uint32_t short_max = ctx->short_max;
  // lists expected to put more than ~48 members into a tile of the rows kernel are walked by the whole warp
  const uint32_t heavy_min = ctx->sparse_kernel == 1 ? 0u
                             : ctx->heavy_min ? ctx->heavy_min
                                              : std::max<uint32_t>(short_max + 1, (uint32_t)(48ull * T / 7680) + 1);
  PRF_CK(ctx, hcursor.alloc((size_t)T + 1, s));
  DevBuf<uint32_t> phead;
  PRF_CK(ctx, H.nheavy.alloc((size_t)T + 1, s));
  PRF_CK(ctx, scount.alloc((size_t)T + 1, s));
  PRF_CK(ctx, scursor.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.soff.alloc((size_t)T + 1, s));
  // upper bound: an entry of a short list is followed by at most short_max - 1 later members
  PRF_CK(ctx, H.singles.alloc((size_t)n_sh_nnz * (short_max > 1 ? (short_max - 1) / 2 + 1 : 0) + 1, s));
  // walked lists padded to multiples of 4 members: at most twice the entries (four times when one-member lists are walked)
  PRF_CK(ctx, phead.alloc((size_t)n_sh_nnz + 2, s));
  PRF_CK(ctx, H.lmembers.alloc((size_t)n_sh_nnz * (short_max ? 2 : 4) + 8, s));
  PRF_CK(ctx, slot_sorted.alloc(n_sh_nnz, s));
  PRF_CK(ctx, H.members.alloc((size_t)n_sh_nnz + 4, s));  // +4: the rows kernel reads members in aligned groups of 4
  PRF_CK(ctx, head_pos.alloc((size_t)n_sh_nnz + 1, s));
  PRF_CK(ctx, rcount.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cursor.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.roff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.ranges.alloc(n_sh_nnz, s));  // upper bound: at most one suffix per shared entry
  PRF_CK(ctx, H.nb.alloc((size_t)T + 16, s));  // slack: the matrix kernel reads |B| in aligned 32-bit pairs
  PRF_CK(ctx, cudaMemsetAsync(H.nb.p, 0, H.nb.bytes(), s));
  uint32_t n_sorted = n_sh_nnz;
  if (n_sh_nnz) {
    gather_shared_kernel<<<(T + 7) / 8, 256, 0, s>>>(d_off, T, ent_slot.p, shared_bits.p, wrank.p, soff.p, slot_sorted.p,
                                                     H.members.p);
    PRF_LAUNCHED(ctx);
    // stable sort by list id: inside a list the trees stay ascending
    if ((rc = radix_sort_pairs(ctx, slot_sorted.p, H.members.p, n_sh_nnz, bits_for(n_lists0)))) return rc;
  }
  bool ids_are_keys = true;  // until duplicates force the generic path, the sorted keys are the list ids
  unsigned long long hc[6] = {0, 0, 0, 0, 0, 0};
  uint32_t hs[8] = {0};
  uint32_t h_nranges = 0;
  for (int pass = 0; pass < 2; ++pass) {
    PRF_CK(ctx, cudaMemsetAsync(rcount.p, 0, rcount.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(cursor.p, 0, cursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(scount.p, 0, scount.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(scursor.p, 0, scursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(hcursor.p, 0, hcursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(H.nheavy.p, 0, H.nheavy.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(counters.p + 2, 0, 32, s));
    const uint32_t* lid_p = slot_sorted.p;
    if (n_sorted) {
      if (ids_are_keys) {
        id_heads_kernel<<<grid_for(n_sorted), 256, 0, s>>>(slot_sorted.p, n_sorted, n_lists0, head_pos.p);
        PRF_LAUNCHED(ctx);
      } else {
        if (!lid_ex.p) PRF_CK(ctx, lid_ex.alloc(n_sh_nnz, s));
        if (!H.lid.p) PRF_CK(ctx, H.lid.alloc(n_sh_nnz, s));
        if ((rc = exclusive_scan(ctx, HeadFlag{slot_sorted.p}, n_sorted, lid_ex.p, scalars.p + 1))) return rc;
        list_heads_kernel<<<grid_for(n_sorted), 256, 0, s>>>(slot_sorted.p, H.members.p, lid_ex.p, n_sorted,
                                                             head_pos.p, counters.p, H.lid.p);
        PRF_LAUNCHED(ctx);
        lid_p = H.lid.p;
      }
      range_count_kernel<<<grid_for(n_sorted), 256, 0, s>>>(lid_p, H.members.p, head_pos.p, n_sorted, short_max, heavy_min, rcount.p,
                                                            H.nheavy.p, scount.p, counters.p);
      PRF_LAUNCHED(ctx);
    }
    if ((rc = exclusive_scan(ctx, LoadU32{rcount.p}, (uint64_t)T + 1, H.roff.p, nullptr))) return rc;
    if ((rc = exclusive_scan(ctx, LoadU32{scount.p}, (uint64_t)T + 1, H.soff.p, nullptr))) return rc;
    if (n_sorted) {
      // scalars[1] = number of lists (ids_are_keys: written by the rank scan above; else by the head-flag scan)
      const uint64_t n_scan = (ids_are_keys ? (uint64_t)n_lists0 : (uint64_t)n_sorted) + 1;
      if ((rc = exclusive_scan(ctx, PadLen{head_pos.p, scalars.p + 1, short_max}, n_scan, phead.p, scalars.p + 7))) return rc;
      range_fill_kernel<<<grid_for(n_sorted), 256, 0, s>>>(lid_p, H.members.p, head_pos.p, n_sorted, short_max, heavy_min, H.roff.p,
                                                           H.nheavy.p, hcursor.p, cursor.p, H.ranges.p, H.soff.p, scursor.p,
                                                           H.singles.p, phead.p, H.lmembers.p);
      PRF_LAUNCHED(ctx);
    }
    if (T) {
      finish_nb_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, T, H.nb.p, scalars.p + 4, H.roff.p);
      PRF_LAUNCHED(ctx);
    }
    PRF_CK(ctx, cudaMemcpyAsync(hc, counters.p, sizeof hc, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaMemcpyAsync(hs, scalars.p, sizeof hs, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaMemcpyAsync(&h_nranges, H.roff.p + T, 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    trace_mark("build: lists synced");
    if (hc[2] == 0 || pass == 1) break;
    // a tree listed some bipartition more than once: its bipartitions are a set -- drop the repeats
    // from the sorted pairs, fix |B_t|, and rebuild the lists once more
    DevBuf<uint32_t> dup_before, slot2, mem2;
    PRF_CK(ctx, dup_before.alloc(n_sorted, s));
    PRF_CK(ctx, slot2.alloc(n_sorted, s));
    PRF_CK(ctx, mem2.alloc((size_t)n_sorted + 4, s));
    if ((rc = exclusive_scan(ctx, DupFlag{slot_sorted.p, H.members.p}, n_sorted, dup_before.p, nullptr))) return rc;
    drop_dups_kernel<<<grid_for(n_sorted), 256, 0, s>>>(slot_sorted.p, H.members.p, dup_before.p, n_sorted,
                                                        slot2.p, mem2.p, nb32.p);
    PRF_LAUNCHED(ctx);
    slot_sorted = std::move(slot2);
    H.members = std::move(mem2);
    ids_are_keys = false;
    n_sorted -= (uint32_t)hc[2];
    uint32_t mm_reset[3] = {0xffffffffu, 0u, 0u};
    PRF_CK(ctx, cudaMemcpyAsync(scalars.p + 4, mm_reset, 12, cudaMemcpyHostToDevice, s));
  }
  if (ids_are_keys) H.lid = std::move(slot_sorted);
  // ---- majority flip (only when some bipartition is held by more than half of the trees) ----
  uint32_t n_lists = n_sorted ? hs[1] : 0;
  if (n_sorted && 2ull * hc[4] > T && !ctx->debug_no_flip) {
    DevBuf<uint32_t> head2, mem2, lid2, hcnt;
    PRF_CK(ctx, head2.alloc((size_t)n_lists + 1, s));
    PRF_CK(ctx, hcnt.alloc((size_t)T + 1, s));
    PRF_CK(ctx, cudaMemsetAsync(hcnt.p, 0, hcnt.bytes(), s));
    if ((rc = exclusive_scan(ctx, FlipLen{head_pos.p, T, n_lists}, (uint64_t)n_lists + 1, head2.p, scalars.p + 2))) return rc;
    uint32_t n2 = 0;
    PRF_CK(ctx, cudaMemcpyAsync(&n2, scalars.p + 2, 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    PRF_CK(ctx, mem2.alloc((size_t)n2 + 4, s));
    PRF_CK(ctx, lid2.alloc(n2, s));
    flip_copy_kernel<<<grid_for(n_sorted), 256, 0, s>>>(H.lid.p, H.members.p, head_pos.p, head2.p, n_sorted, T, mem2.p,
                                                        lid2.p, hcnt.p);
    PRF_LAUNCHED(ctx);
    flip_complement_kernel<<<grid_for(n_sorted), 256, 0, s>>>(H.lid.p, H.members.p, head_pos.p, head2.p, n_sorted, T,
                                                              mem2.p, lid2.p, counters.p);
    PRF_LAUNCHED(ctx);
    flip_nb_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, hcnt.p, T, counters.p);
    PRF_LAUNCHED(ctx);
    H.members = std::move(mem2);
    H.lid = std::move(lid2);
    n_sorted = n2;
    // ranges, hits and |D_t| statistics over the flipped lists
    PRF_CK(ctx, H.ranges.alloc(n_sorted, s));
    PRF_CK(ctx, H.singles.alloc((size_t)n_sorted * (short_max > 1 ? (short_max - 1) / 2 + 1 : 0) + 1, s));
    PRF_CK(ctx, phead.alloc((size_t)n_lists + 2, s));
    PRF_CK(ctx, H.lmembers.alloc((size_t)n_sorted * (short_max ? 2 : 4) + 8, s));
    PRF_CK(ctx, cudaMemsetAsync(rcount.p, 0, rcount.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(cursor.p, 0, cursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(scount.p, 0, scount.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(scursor.p, 0, scursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(hcursor.p, 0, hcursor.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(H.nheavy.p, 0, H.nheavy.bytes(), s));
    PRF_CK(ctx, cudaMemsetAsync(counters.p + 3, 0, 16, s));
    const uint32_t mm_reset[3] = {0xffffffffu, 0u, 0u};
    PRF_CK(ctx, cudaMemcpyAsync(scalars.p + 4, mm_reset, 12, cudaMemcpyHostToDevice, s));
    if (n_sorted) {
      range_count_kernel<<<grid_for(n_sorted), 256, 0, s>>>(H.lid.p, H.members.p, head2.p, n_sorted, short_max, heavy_min, rcount.p,
                                                            H.nheavy.p, scount.p, counters.p);
      PRF_LAUNCHED(ctx);
    }
    if ((rc = exclusive_scan(ctx, LoadU32{rcount.p}, (uint64_t)T + 1, H.roff.p, nullptr))) return rc;
    if ((rc = exclusive_scan(ctx, LoadU32{scount.p}, (uint64_t)T + 1, H.soff.p, nullptr))) return rc;
    if (n_sorted) {
      if ((rc = exclusive_scan(ctx, PadLen{head2.p, scalars.p + 1, short_max}, (uint64_t)n_lists + 1, phead.p, scalars.p + 7))) return rc;
      range_fill_kernel<<<grid_for(n_sorted), 256, 0, s>>>(H.lid.p, H.members.p, head2.p, n_sorted, short_max, heavy_min, H.roff.p,
                                                           H.nheavy.p, hcursor.p, cursor.p, H.ranges.p, H.soff.p, scursor.p,
                                                           H.singles.p, phead.p, H.lmembers.p);
      PRF_LAUNCHED(ctx);
    }
    finish_nb_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, T, H.nb.p, scalars.p + 4, H.roff.p);
    PRF_LAUNCHED(ctx);
    unsigned long long hc2[6];
    PRF_CK(ctx, cudaMemcpyAsync(hc2, counters.p, sizeof hc2, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaMemcpyAsync(hs, scalars.p, sizeof hs, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaMemcpyAsync(&h_nranges, H.roff.p + T, 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    hc[3] = hc2[3];
    H.n_flipped = (uint32_t)hc2[5];
  }
  uint32_t mm[2] = {hs[4], hs[5]};
  if (T == 0) mm[0] = mm[1] = 0;
  if (mm[1] > kMaxBipsPerTree)
    return ctx->fail(PRF_E_RANGE, "a tree has %u bipartitions; distances would overflow uint16 (max %u per tree)",
                     mm[1], kMaxBipsPerTree);
  H.uniform_nb = mm[0] == mm[1];
  H.nb_uniform = mm[0];
  H.n_shared = n_lists;
  H.n_shared_nnz = n_sorted;  // after duplicate removal
  H.n_ranges = h_nranges;
  H.max_ranges = hs[6];
  H.n_lmembers = n_sorted ? hs[7] : 0;

  prf_stats& st = H.stats;
  st = prf_stats{};
  st.n_trees = T;
  st.n_words = W;
  st.nnz = nnz;
  st.n_unique = hc[0];
  st.n_shared = H.n_shared;
  st.n_shared_nnz = n_sorted;
  st.n_hits = hc[3];
  st.n_pairs = H.P;
  st.max_bips = mm[1];
  st.min_bips = mm[0];
  st.n_flipped = H.n_flipped;
  if ((rc = build_nb_shift(ctx))) return rc;
  H.loaded = true;
  return PRF_OK;
}

}  // namespace prf
