// prf_incidence.cuh -- the `build` phase of hashRF (RFDistance.hs:289-297) on the device:
//   K1  global bipartition dedup: open-addressing hash table over the W-word bit sets with exact
//       full-key verification (a 64-bit hash only routes; equality is always decided on the key).
//       A bit per slot records "seen by a second entry": bipartitions present in one tree only
//       (~3/4 of them on random topologies) can never be shared, they only count in |B_t|.
//   K2  trees x shared-bipartitions incidence in the form the matrix kernel wants (ListPart, prf_common.cuh),
//       WITHOUT a sort: the rank of a shared slot in the "seen twice" bitmap is the dense list id; list sizes
//       are counted with atomics; lists of more than short_max trees are counting-sorted into (list, column
//       block) buckets (any order inside a bucket), the others expanded into per-tree singles.
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
                                                           const uint64_t* __restrict__ tb,
                                                           const uint64_t* __restrict__ te,
                                                           uint32_t T, unsigned long long* slots,
                                                           uint32_t* shared_bits, uint64_t cap_mask,
                                                           uint32_t* __restrict__ ent_slot,
                                                           unsigned long long* counters) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const uint64_t e0 = tb[t], e1 = te[t];
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
// headline shape the table is 98 MB instead of 268 MB: it mostly stays in the 126 MB L2 (the key stream is
// loaded with an evict-first policy), so the random probes and CAS are L2 hits instead of DRAM sectors.
// One thread per entry over up to kMaxParts dense entry segments (one for a single-GPU build; one per source rank
// for a class build of a multi-GPU team, whose trees have too few entries each to fill a warp).
struct DedupSegs {
  uint64_t begin[kMaxParts];
  uint64_t start[kMaxParts + 1];  // entries of the segments before this one
  uint32_t n;
};
template <int W>
__global__ void __launch_bounds__(256, (W <= 6 ? 8 : W <= 10 ? 6 : 4)) dedup_insert32_kernel(const uint64_t* __restrict__ bips, const DedupSegs segs,
                                                             uint32_t* slots, uint32_t* shared_bits, uint32_t cap,
                                                             uint32_t* __restrict__ ent_slot,
                                                             unsigned long long* counters, int mode) {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  const uint64_t total = segs.start[segs.n];
  uint32_t claims = 0;
  // (the kernel is bound by the number of random accesses in flight, i.e. by resident threads: the launch bounds keep
  // W <= 6 at 32 registers = 2048 threads per SM, 0.55 vs 0.62 ms at 100k x 200.  Several keys in flight per thread --
  // first-slot loads issued back to back, keys read back on a tag match -- measured 0.98 ms: the per-thread chains of
  // the later probes serialise; prefetching the next key's slot into L2: 0.78 ms at 46 registers)
  // a warp's 32 keys are read with coalesced 16-byte loads and handed to their lanes through shared memory (a lane
  // loading its own 8 W bytes would touch one sector per lane and load: the L1 wavefronts were a third of the kernel)
  __shared__ __align__(16) uint64_t s_keys[8][32 * W + 2];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < total; i0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i0 + lane;
    uint32_t sg = 0;
    while (sg + 1 < segs.n && i >= segs.start[sg + 1]) ++sg;
    const uint64_t e = segs.begin[sg] + (i - segs.start[sg]);
    const bool valid = i < total;
    uint64_t k[W];
    const uint32_t sg0 = __shfl_sync(0xffffffffu, sg, 0);
    const uint32_t nv = __popc(__ballot_sync(0xffffffffu, valid));
    if (W % 2 == 0 && __all_sync(0xffffffffu, !valid || sg == sg0)) {  // one segment: the keys are contiguous
      const uint64_t ef = __shfl_sync(0xffffffffu, e, 0);
      const ulonglong2* src = reinterpret_cast<const ulonglong2*>(bips + ef * W);
      ulonglong2* dst = reinterpret_cast<ulonglong2*>(&s_keys[warp][0]);
#pragma unroll
      for (int c = 0; c < W / 2; ++c) {
        const uint32_t ch = c * 32 + lane;
        if (ch < nv * (W / 2)) {
          ulonglong2 q;
          asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(q.x), "=l"(q.y) : "l"(src + ch), "l"(policy));
          dst[ch] = q;
        }
      }
      __syncwarp();
#pragma unroll
      for (int w = 0; w < W; ++w) k[w] = s_keys[warp][lane * W + w];
      __syncwarp();
    } else if (valid) {
      const uint64_t* p = bips + e * W;
#pragma unroll
      for (int w = 0; w < W; ++w)
        asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(k[w]) : "l"(p + w), "l"(policy));
    }
    if (!valid) continue;
    uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
    for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
    const uint32_t tag = (uint32_t)(h >> 57);  // 7 bits, independent of the slot bits below
    const uint32_t mine = (tag << 25) | ((uint32_t)e + 1u);
    uint32_t s = (uint32_t)(((h & 0xffffffffull) * (uint64_t)cap) >> 32);
    for (;;) {
      uint32_t cur;
      if (mode & 1) {  // small table: see hashrf_build
        cur = atomicCAS(&slots[s], 0u, mine);
        if (cur == 0u) { ++claims; break; }
      } else {
        cur = slots[s];
        if (cur == 0u) {
          cur = atomicCAS(&slots[s], 0u, mine);
          if (cur == 0u) { ++claims; break; }  // claimed: this entry is the representative
        }
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

// Number of shared slots in one 32-slot word of the bitmap (scanned into per-word ranks).
struct PopcWord {
  const uint32_t* bits;
  __device__ uint32_t operator()(uint64_t w) const { return __popc(bits[w]); }
};

// ---- K2a: entries -> list ids, per-list sizes -------------------------------------------------------------
// One warp per tree.  ent[e]: in = table slot of entry e, out = dense list id (rank of the slot among the
// shared slots) or 0xffffffff when the bipartition is held by this tree only.  A tree's entries are a SET
// (S.Set DenseLabelSet): an exact per-warp hash set in shared memory drops a list id the tree already has.
// lcnt[l] = trees holding list l; nb32[t] = |B_t| (distinct entries).
template <int G>  // lanes per tree: 32, or 8 (four trees per warp) when the trees have few entries (class builds)
__global__ void __launch_bounds__(256) list_count_kernel(const uint64_t* __restrict__ tb, const uint64_t* __restrict__ te,
                                                         uint32_t T, uint32_t* __restrict__ ent,
                                                         const uint32_t* __restrict__ shared_bits,
                                                         const uint32_t* __restrict__ wrank, uint32_t set_slots,
                                                         uint32_t* lcnt, uint32_t* __restrict__ nb32) {
  extern __shared__ uint32_t s_sets[];
  constexpr uint32_t kPer = 32 / G;
  const uint32_t warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = lane_id(), sub = lane % G, grp = lane / G;
  uint32_t* set = s_sets + ((size_t)warp * kPer + grp) * set_slots;
  for (uint32_t t0 = (blockIdx.x * warps + warp) * kPer; t0 < T; t0 += gridDim.x * warps * kPer) {
    const uint32_t t = t0 + grp;
    const bool live = t < T;
    const uint64_t e0 = live ? tb[t] : 0, e1 = live ? te[t] : 0;
    for (uint32_t i = sub; i < set_slots; i += G) set[i] = 0xffffffffu;
    __syncwarp();
    uint32_t dups = 0;
    for (uint64_t e = e0 + sub; e < e1; e += G) {
      const uint32_t s = ent[e];
      const uint32_t word = shared_bits[s >> 5];
      uint32_t l = 0xffffffffu;
      if ((word >> (s & 31)) & 1u) {
        l = wrank[s >> 5] + __popc(word & ((1u << (s & 31)) - 1u));
        uint32_t h = (l * 0x9E3779B1u) & (set_slots - 1);
        for (;;) {
          const uint32_t old = atomicCAS(&set[h], 0xffffffffu, l);
          if (old == 0xffffffffu) break;
          if (old == l) {  // the tree listed this bipartition before
            l = 0xffffffffu;
            ++dups;
            break;
          }
          h = (h + 1) & (set_slots - 1);
        }
        if (l != 0xffffffffu) atomicAdd(&lcnt[l], 1u);
      }
      ent[e] = l;
    }
    __syncwarp();
#pragma unroll
    for (int d = G / 2; d; d >>= 1) dups += __shfl_xor_sync(0xffffffffu, dups, d);
    if (sub == 0 && live) nb32[t] = (uint32_t)(e1 - e0) - dups;
    __syncwarp();
  }
}

// ---- K2b: per-list classification -------------------------------------------------------------------------
struct ListCounts {  // device scalars written by the build (read back once)
  unsigned long long n_hits;      // sum over lists m (m - 1) / 2
  unsigned long long n_entries;   // sum over lists of >= 2 trees of m
  unsigned long long short_pairs; // sum over short lists m (m - 1) / 2  (= number of singles)
  unsigned long long long_members;  // sum over long lists m
  uint32_t n_lists;               // shared slots (rank scan total)
  uint32_t n_lists2;              // lists of >= 2 trees (a tree that lists a bipartition twice can leave a list of 1)
  uint32_t n_long;
  uint32_t short_members;         // sum over short lists of m
  uint32_t longest;
  uint32_t nb_min, nb_max, max_lists;
  uint32_t n_unique_lo, n_unique_hi;
  uint32_t n_lmem;                // padded size of lmem
  uint32_t n_heavy;               // lists held by more than half of the trees
};

struct IsLong {
  const uint32_t* lcnt;
  const ListCounts* c;
  uint32_t short_max;
  __device__ uint32_t operator()(uint64_t l) const { return l < c->n_lists && lcnt[l] > short_max ? 1u : 0u; }
};
struct ShortLen {
  const uint32_t* lcnt;
  const ListCounts* c;
  uint32_t short_max;
  __device__ uint32_t operator()(uint64_t l) const {
    if (l >= c->n_lists) return 0;
    const uint32_t m = lcnt[l];
    return m >= 2 && m <= short_max ? m : 0u;
  }
};
// With `flip`, a list held by more than half of the trees counts with the T - m trees that LACK it (majority flip).
__global__ void __launch_bounds__(256) list_stats_kernel(const uint32_t* __restrict__ lcnt, uint32_t short_max, uint32_t T,
                                                         int flip, ListCounts* c) {
  const uint32_t n = c->n_lists;
  unsigned long long hits = 0, ent = 0, sp = 0, lm = 0;
  uint32_t n2 = 0, longest = 0, heavy = 0;
  for (uint32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x) {
    const unsigned long long m0 = lcnt[l];
    if (m0 >= 2) {
      const bool hv = 2 * m0 > T && m0 > short_max;  // (short lists stay singles whatever their share of the trees)
      const unsigned long long m = hv && flip ? T - m0 : m0;
      ++n2;
      ent += m;
      hits += m ? m * (m - 1) / 2 : 0;
      if (m0 <= short_max) sp += m0 * (m0 - 1) / 2;
      else lm += m;
      longest = max(longest, (uint32_t)m0);
      if (hv) ++heavy;
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    hits += __shfl_xor_sync(0xffffffffu, hits, d);
    ent += __shfl_xor_sync(0xffffffffu, ent, d);
    sp += __shfl_xor_sync(0xffffffffu, sp, d);
    lm += __shfl_xor_sync(0xffffffffu, lm, d);
    n2 += __shfl_xor_sync(0xffffffffu, n2, d);
    heavy += __shfl_xor_sync(0xffffffffu, heavy, d);
    longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, d));
  }
  if (lane_id() == 0) {
    if (hits) atomicAdd(&c->n_hits, hits);
    if (ent) atomicAdd(&c->n_entries, ent);
    if (sp) atomicAdd(&c->short_pairs, sp);
    if (lm) atomicAdd(&c->long_members, lm);
    if (n2) atomicAdd(&c->n_lists2, n2);
    if (heavy) atomicAdd(&c->n_heavy, heavy);
    if (longest) atomicMax(&c->longest, longest);
  }
}

// ---- K2c: bucket sizes of the long lists, raw member arrays of the short lists ------------------------------
// One warp per tree over ent (list ids).  bcnt[ll * NB + (t >> block_log2)] counts the members of long list ll
// inside t's column block; a short list's members go to sraw[sbase[l] ..) in arrival order; rcnt[t] = long lists
// of tree t.
template <int G>
__global__ void __launch_bounds__(256) bucket_count_kernel(const uint64_t* __restrict__ tb, const uint64_t* __restrict__ te,
                                                           uint32_t T, const uint32_t* __restrict__ ent,
                                                           const uint32_t* __restrict__ lcnt, const uint32_t* __restrict__ llid,
                                                           const uint32_t* __restrict__ sbase, uint32_t short_max,
                                                           uint32_t block_log2, uint32_t NB, uint32_t* bcnt, uint32_t* sfill,
                                                           uint32_t* __restrict__ sraw, uint32_t* __restrict__ rcnt,
                                                           const uint32_t* __restrict__ fid, uint32_t* fbits, uint32_t fwords,
                                                           uint32_t* __restrict__ hcnt) {
  constexpr uint32_t kPer = 32 / G;
  const uint32_t warps = blockDim.x >> 5, lane = lane_id(), sub = lane % G, grp = lane / G;
  for (uint32_t t0 = (blockIdx.x * warps + (threadIdx.x >> 5)) * kPer; t0 < T; t0 += gridDim.x * warps * kPer) {
    const uint32_t t = t0 + grp;
    const bool live = t < T;
    const uint64_t e0 = live ? tb[t] : 0, e1 = live ? te[t] : 0;
    uint32_t nl = 0, nh = 0;
    for (uint64_t e = e0 + sub; e < e1; e += G) {
      const uint32_t l = ent[e];
      if (l == 0xffffffffu) continue;
      const uint32_t m = lcnt[l];
      if (fid && 2ull * m > T && m > short_max) {  // majority list: only its bitmap of holders is kept
        ++nh;
        atomicOr(&fbits[(size_t)fid[l] * fwords + (t >> 5)], 1u << (t & 31));
      } else if (m > short_max) {
        ++nl;
        atomicAdd(&bcnt[(size_t)llid[l] * NB + (t >> block_log2)], 1u);
      } else if (m >= 2) {
        sraw[sbase[l] + atomicAdd(&sfill[l], 1u)] = t;
      }
    }
    __syncwarp();
#pragma unroll
    for (int d = G / 2; d; d >>= 1) {
      nl += __shfl_xor_sync(0xffffffffu, nl, d);
      nh += __shfl_xor_sync(0xffffffffu, nh, d);
    }
    if (sub == 0 && live) {
      rcnt[t] = nl;
      if (hcnt) hcnt[t] = nh;
    }
  }
}

// ---- majority flip -------------------------------------------------------------------------------------------
// [u in B_a] xor [u in B_b] is unchanged when membership of bipartition u is complemented for EVERY tree, so a
// list held by more than half of the trees is replaced by the (shorter) list of trees that LACK it:
// d(a,b) = |D_a| + |D_b| - 2 |D_a n D_b| with D_t = B_t xor {majority bipartitions}.  Similar trees (few
// differences from a common backbone) become a sparse problem again.
struct IsHeavy {
  const uint32_t* lcnt;
  const ListCounts* c;
  uint32_t T, short_max;
  __device__ uint32_t operator()(uint64_t l) const {
    return l < c->n_lists && lcnt[l] > short_max && 2ull * lcnt[l] > T ? 1u : 0u;
  }
};
// fll[f] = long-list id of majority list f
__global__ void __launch_bounds__(256) flip_ids_kernel(const uint32_t* __restrict__ lcnt, const uint32_t* __restrict__ llid,
                                                       const uint32_t* __restrict__ fid, const ListCounts* c, uint32_t T,
                                                       uint32_t short_max, uint32_t* __restrict__ fll) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < c->n_lists && lcnt[l] > short_max && 2ull * lcnt[l] > T) fll[fid[l]] = llid[l];
}
// One warp per (majority list f, column block j): pass 0 counts the trees of the block lacking the bipartition
// into bcnt, pass 1 writes their column offsets into the bucket.
__global__ void __launch_bounds__(256) flip_buckets_kernel(const uint32_t* __restrict__ fbits, uint32_t fwords,
                                                           const uint32_t* __restrict__ fll, uint32_t n_heavy, uint32_t T,
                                                           uint32_t block_log2, uint32_t NB, int pass, uint32_t* bcnt,
                                                           const uint32_t* __restrict__ boff, uint16_t* __restrict__ lmem) {
  const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = lane_id();
  if (w >= n_heavy * NB) return;
  const uint32_t f = w / NB, j = w - f * NB;
  const uint32_t t0 = j << block_log2, t1 = min(T, t0 + (1u << block_log2));
  const uint32_t* bits = fbits + (size_t)f * fwords;
  const size_t b = (size_t)fll[f] * NB + j;
  uint32_t run = 0;
  for (uint32_t wb = t0 >> 5; wb < (t1 + 31) >> 5; wb += 32) {
    const uint32_t wi = wb + lane;
    uint32_t z = 0;  // trees of word wi lacking the bipartition
    if (wi < (t1 + 31) >> 5) {
      z = ~bits[wi];
      const uint32_t lo = wi << 5;
      if (lo + 32 > t1) z &= (1u << (t1 - lo)) - 1u;
    }
    const uint32_t n = __popc(z);
    uint32_t inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= (uint32_t)d) inc += x;
    }
    if (pass == 1) {
      uint32_t o = boff[b] + run + inc - n;
      while (z) {
        const uint32_t bit = __ffs(z) - 1;
        z &= z - 1;
        lmem[o++] = (uint16_t)(((wi << 5) + bit) & ((1u << block_log2) - 1u));
      }
    }
    run += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (pass == 0 && lane == 0) bcnt[b] = run;
}
__global__ void __launch_bounds__(256) nb16_kernel(const uint32_t* __restrict__ nb32, uint32_t T, uint16_t* __restrict__ nb16) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) nb16[t] = (uint16_t)min(nb32[t], 0xffffu);
}
// |D_t| and the number of lists of tree t with the majority lists flipped
__global__ void __launch_bounds__(256) flip_trees_kernel(uint32_t* nb32, uint32_t* rcnt, const uint32_t* __restrict__ hcnt,
                                                         uint32_t T, uint32_t n_heavy) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  nb32[t] = nb32[t] + n_heavy - 2u * hcnt[t];
  rcnt[t] += n_heavy - hcnt[t];
}
// One warp per tree: the majority lists it LACKS go behind its ordinary long lists in rlist.
__global__ void __launch_bounds__(256) flip_rlist_kernel(const uint32_t* __restrict__ fbits, uint32_t fwords,
                                                         const uint32_t* __restrict__ fll, uint32_t n_heavy, uint32_t T,
                                                         const uint32_t* __restrict__ roff, const uint32_t* __restrict__ hcnt,
                                                         uint32_t* __restrict__ rlist) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = lane_id(), lt = lanemask_lt();
  if (t >= T) return;
  uint32_t pos = roff[t + 1] - (n_heavy - hcnt[t]);
  for (uint32_t fb = 0; fb < n_heavy; fb += 32) {
    const uint32_t f = fb + lane;
    const bool lacks = f < n_heavy && !((fbits[(size_t)f * fwords + (t >> 5)] >> (t & 31)) & 1u);
    const uint32_t bal = __ballot_sync(0xffffffffu, lacks);
    if (lacks) rlist[pos + __popc(bal & lt)] = fll[f];
    pos += __popc(bal);
  }
}

struct Pad8 {  // bucket sizes rounded up to the 16-byte groups the matrix kernel loads
  const uint32_t* bcnt;
  uint64_t n;
  __device__ uint32_t operator()(uint64_t i) const { return i < n ? (bcnt[i] + 7u) & ~7u : 0u; }
};

// ---- K2d: fill the buckets and the trees' list arrays ------------------------------------------------------
// bcnt doubles as the fill cursor (counted down); lmem was preset to 0xffff (the padding value).
template <int G>
__global__ void __launch_bounds__(256) bucket_fill_kernel(const uint64_t* __restrict__ tb, const uint64_t* __restrict__ te,
                                                          uint32_t T, const uint32_t* __restrict__ ent,
                                                          const uint32_t* __restrict__ lcnt, const uint32_t* __restrict__ llid,
                                                          uint32_t short_max, uint32_t block_log2, uint32_t NB, uint32_t* bcnt,
                                                          const uint32_t* __restrict__ boff, uint16_t* __restrict__ lmem,
                                                          const uint32_t* __restrict__ roff, uint32_t* __restrict__ rlist,
                                                          int flip) {
  constexpr uint32_t kPer = 32 / G;
  const uint32_t warps = blockDim.x >> 5, lane = lane_id(), sub = lane % G, grp = lane / G;
  const uint32_t gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (G * grp)), lt = lanemask_lt() & gmask;
  for (uint32_t t0 = (blockIdx.x * warps + (threadIdx.x >> 5)) * kPer; t0 < T; t0 += gridDim.x * warps * kPer) {
    const uint32_t t = t0 + grp;
    const bool live = t < T;
    const uint64_t e0 = live ? tb[t] : 0, e1 = live ? te[t] : 0;
    uint32_t pos = live ? roff[t] : 0;
    uint32_t len = (uint32_t)(e1 - e0);  // the longest tree of the warp sets the (warp-uniform) trip count
#pragma unroll
    for (int d = 16; d; d >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, d));
    for (uint64_t i = 0; i < len; i += G) {
      const uint64_t e = e0 + i + sub;
      uint32_t ll = 0xffffffffu;
      if (e < e1) {
        const uint32_t l = ent[e];
        if (l != 0xffffffffu && lcnt[l] > short_max && !(flip && 2ull * lcnt[l] > T)) ll = llid[l];
      }
      const bool is = ll != 0xffffffffu;
      const uint32_t bal = __ballot_sync(0xffffffffu, is) & gmask;
      if (is) {
        const size_t b = (size_t)ll * NB + (t >> block_log2);
        lmem[boff[b] + (atomicSub(&bcnt[b], 1u) - 1u)] = (uint16_t)(t & ((1u << block_log2) - 1u));
        rlist[pos + __popc(bal & lt)] = ll;
      }
      pos += __popc(bal);
    }
  }
}

// ---- K2e: singles ------------------------------------------------------------------------------------------
// One thread per short list: tree a gets the column b of every member b > a.  Pass 0 counts per tree, pass 1 fills.
__global__ void __launch_bounds__(256) singles_kernel(const uint32_t* __restrict__ lcnt, const uint32_t* __restrict__ sbase,
                                                      const uint32_t* __restrict__ sraw, uint32_t short_max, const ListCounts* c,
                                                      int pass, uint32_t* scount, const uint32_t* __restrict__ soff,
                                                      uint32_t* scur, uint32_t* __restrict__ singles) {
  const uint32_t n = c->n_lists;
  for (uint32_t l = blockIdx.x * blockDim.x + threadIdx.x; l < n; l += gridDim.x * blockDim.x) {
    const uint32_t m = lcnt[l];
    if (m < 2 || m > short_max) continue;
    const uint32_t* mem = sraw + sbase[l];
    for (uint32_t i = 0; i < m; ++i) {
      const uint32_t a = mem[i];
      uint32_t k = 0;
      for (uint32_t j = 0; j < m; ++j) k += mem[j] > a ? 1u : 0u;
      if (!k) continue;
      if (pass == 0) {
        atomicAdd(&scount[a], k);
      } else {
        uint32_t o = soff[a] + atomicAdd(&scur[a], k);
        for (uint32_t j = 0; j < m; ++j)
          if (mem[j] > a) singles[o++] = mem[j];
      }
    }
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

// nb32 -> uint16, min / max |B_t| and the most long lists of one tree
__global__ void __launch_bounds__(256) finish_trees_kernel(const uint32_t* __restrict__ nb32, const uint32_t* __restrict__ rcnt,
                                                           uint32_t T, uint16_t* __restrict__ nb16, ListCounts* c) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = t < T ? nb32[t] : 0, lo = t < T ? v : 0xffffffffu, hi = v, ml = t < T ? rcnt[t] : 0;
  if (t < T) nb16[t] = (uint16_t)(v > 0xffffu ? 0xffffu : v);
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    ml = max(ml, __shfl_xor_sync(0xffffffffu, ml, d));
  }
  if (lane_id() == 0) {
    atomicMin(&c->nb_min, lo);
    atomicMax(&c->nb_max, hi);
    atomicMax(&c->max_lists, ml);
  }
}

template <int W>
static void launch_dedup(Ctx* ctx, const uint64_t* bips, const uint64_t* tb, const uint64_t* te, uint32_t T,
                         unsigned long long* slots, uint32_t* shared_bits, uint64_t cap_mask, uint32_t* ent_slot,
                         unsigned long long* counters) {
  dedup_insert_kernel<W><<<(T + 7) / 8, 256, 0, ctx->stream>>>(bips, tb, te, T, slots, shared_bits, cap_mask, ent_slot, counters);
}

// The shifted copies of |B_t| the sparse kernel's fill bulk-loads (non-uniform |B_t| only).
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

// d_bips: keys (W words per entry); tree t's entries are [tb[t], te[t]) (device arrays); `span` = one past the
// largest entry index; max_entries = most entries of one tree.
int hashrf_build(Ctx* ctx, const uint64_t* d_bips, const uint64_t* tb, const uint64_t* te, uint32_t T, uint32_t W,
                 uint64_t span, uint64_t nnz, uint32_t max_entries, const uint64_t* seg_begin, const uint64_t* seg_len,
                 uint32_t n_segs) {
  cudaStream_t s = ctx->stream;
  HashRFState& H = ctx->h;
  H = HashRFState{};
  H.T = T;
  H.W = W;
  H.nnz = nnz;
  H.P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  H.block_log2 = ctx->debug_small_tiles ? 8 : 13;
  const uint32_t NB = T ? ((T - 1) >> H.block_log2) + 1 : 1;
  H.NB = NB;
  const uint32_t short_max = ctx->short_max;
  if (max_entries > 4096)
    return ctx->fail(PRF_E_UNSUPPORTED, "a tree lists %u bipartition entries; supported <= 4096", max_entries);

  // compact 4-byte slots whenever the entry index fits 25 bits (see dedup_insert32_kernel)
  const bool compact = span < (1ull << 25) - 1 && !ctx->debug_wide_slots;
  uint64_t cap = 1024;
  // load factor <= 0.8 where the table has to fight for the L2 (above 8 M entries), <= 0.5 (shorter probe chains) below
  if (compact) cap = std::max<uint64_t>(1024, ((uint64_t)(nnz * (nnz > (8u << 20) ? 1.25 : 2.0)) + 31) / 32 * 32);
  else while (cap < nnz + nnz / 2) cap <<= 1;  // load factor <= 2/3 even if every key is distinct
  if (!compact && cap > (1ull << 32)) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu entries: the dedup table would exceed 2^32 slots", (unsigned long long)nnz);

  DevBuf<unsigned long long> slots, counters;
  DevBuf<uint32_t> slots32, shared_bits, wrank, nb32, rcnt, lcnt, llid, sbase, sfill;
  DevBuf<ListCounts> cnts;
  const size_t Lmax = (size_t)(nnz / 2) + 2;  // a shared slot holds at least two entries
  if (compact) PRF_CK(ctx, slots32.alloc(cap, s));
  else PRF_CK(ctx, slots.alloc(cap, s));
  PRF_CK(ctx, shared_bits.alloc(cap / 32, s));
  PRF_CK(ctx, wrank.alloc(cap / 32, s));
  PRF_CK(ctx, H.ent_lid.alloc(std::max<uint64_t>(span, 1), s));
  PRF_CK(ctx, H.ent_tb.alloc(std::max<uint32_t>(T, 1), s));
  PRF_CK(ctx, H.ent_te.alloc(std::max<uint32_t>(T, 1), s));
  PRF_CK(ctx, nb32.alloc((size_t)T + 1, s));
  PRF_CK(ctx, rcnt.alloc((size_t)T + 1, s));
  PRF_CK(ctx, lcnt.alloc(Lmax, s));
  PRF_CK(ctx, llid.alloc(Lmax, s));
  PRF_CK(ctx, sbase.alloc(Lmax, s));
  PRF_CK(ctx, sfill.alloc(Lmax, s));
  PRF_CK(ctx, counters.alloc(2, s));
  PRF_CK(ctx, cnts.alloc(1, s));
  PRF_CK(ctx, cudaMemsetAsync(shared_bits.p, 0, shared_bits.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(counters.p, 0, counters.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(lcnt.p, 0, lcnt.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(sfill.p, 0, sfill.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(rcnt.p, 0, rcnt.bytes(), s));
  if (compact) PRF_CK(ctx, cudaMemsetAsync(slots32.p, 0, slots32.bytes(), s));
  else PRF_CK(ctx, cudaMemsetAsync(slots.p, 0xff, slots.bytes(), s));
  ListCounts c0{};
  c0.nb_min = 0xffffffffu;
  PRF_CK(ctx, cudaMemcpyAsync(cnts.p, &c0, sizeof c0, cudaMemcpyHostToDevice, s));
  if (T) {
    PRF_CK(ctx, cudaMemcpyAsync(H.ent_tb.p, tb, (size_t)T * 8, cudaMemcpyDeviceToDevice, s));
    PRF_CK(ctx, cudaMemcpyAsync(H.ent_te.p, te, (size_t)T * 8, cudaMemcpyDeviceToDevice, s));
  }
  tb = H.ent_tb.p;  // the caller's arrays may go away after the load; the dense variant reads ent_lid later
  te = H.ent_te.p;

  trace_mark("build: temporaries allocated");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  // ---- K1: dedup ----
  DedupSegs segs{};  // the dense entry segments: the whole array, or one per source rank of a team
  if (n_segs == 0) {
    segs.n = 1;
    segs.begin[0] = 0;
    segs.start[1] = nnz;
  } else {
    segs.n = n_segs;
    for (uint32_t i = 0; i < n_segs; ++i) {
      segs.begin[i] = seg_begin[i];
      segs.start[i + 1] = segs.start[i] + seg_len[i];
    }
  }
  // a table that stays in L2 is probed with the compare-and-swap itself (one round trip for a new key: 0.047 vs 0.071 ms
  // at 10k x 100); a larger one with a plain load first (an atomic on a line that misses L2 is slower: 0.85 vs 0.62 ms
  // at 100k x 200)
  const int dedup_mode = compact && cap * 4 <= (32ull << 20) ? 1 : 0;
  const int dedup_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((segs.start[segs.n] + 255) / 256, (uint64_t)ctx->sm_count * 64));
  if (T) {
    switch (W) {
#define PRF_W_CASE(w)                                                                                                  \
  case w:                                                                                                              \
    if (compact)                                                                                                       \
      dedup_insert32_kernel<w><<<dedup_grid, 256, 0, s>>>(d_bips, segs, slots32.p, shared_bits.p, (uint32_t)cap,        \
                                                          H.ent_lid.p, counters.p, dedup_mode);                        \
    else                                                                                                               \
      launch_dedup<w>(ctx, d_bips, tb, te, T, slots.p, shared_bits.p, cap - 1, H.ent_lid.p, counters.p);               \
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
  trace_sync(s, "build:   dedup kernel");

  // ---- K2: lists ----
  int rc;
  // lanes per tree in the tree-major kernels: a class build of a multi-GPU team sees every tree with few entries
  const bool narrow = T && nnz / T < 48;
  const uint32_t tpw = narrow ? 4 : 1;  // trees per warp
  const int grid_t = std::max(1, std::min<int>((int)((T + 8 * tpw - 1) / (8 * tpw)), ctx->sm_count * 8));
  if ((rc = exclusive_scan(ctx, PopcWord{shared_bits.p}, cap / 32, wrank.p, &cnts.p->n_lists))) return rc;
  uint32_t set_slots = 64;
  while (set_slots < 2 * max_entries) set_slots <<= 1;
  if (T) {
    const int warps = (int)std::max<uint32_t>(1, std::min<uint32_t>(8, (48 * 1024) / (set_slots * 4 * tpw)));
    const int grid_l = std::max(1, std::min<int>((int)((T + warps * tpw - 1) / (warps * tpw)), ctx->sm_count * 8));
    const size_t smem_l = (size_t)warps * tpw * set_slots * 4;
    if (narrow) list_count_kernel<8><<<grid_l, 32 * warps, smem_l, s>>>(tb, te, T, H.ent_lid.p, shared_bits.p, wrank.p, set_slots, lcnt.p, nb32.p);
    else list_count_kernel<32><<<grid_l, 32 * warps, smem_l, s>>>(tb, te, T, H.ent_lid.p, shared_bits.p, wrank.p, set_slots, lcnt.p, nb32.p);
    PRF_LAUNCHED(ctx);
  }
  trace_sync(s, "build:   rank scan + list_count");
  if ((rc = exclusive_scan(ctx, IsLong{lcnt.p, cnts.p, short_max}, Lmax, llid.p, &cnts.p->n_long, &cnts.p->n_lists))) return rc;
  if ((rc = exclusive_scan(ctx, ShortLen{lcnt.p, cnts.p, short_max}, Lmax, sbase.p, &cnts.p->short_members, &cnts.p->n_lists))) return rc;
  const int flip_ok = ctx->debug_no_flip ? 0 : 1;
  list_stats_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(lcnt.p, short_max, T, flip_ok, cnts.p);
  PRF_LAUNCHED(ctx);
  ListCounts hc{};
  unsigned long long hcount[2] = {0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(&hc, cnts.p, sizeof hc, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaMemcpyAsync(hcount, counters.p, sizeof hcount, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));  // the one read-back inside the build: sizes of what follows
  trace_mark("build: lists counted");
  if (hc.long_members >= 0xfffffff0ull || hc.short_pairs >= 0xfffffff0ull || (uint64_t)hc.n_long * NB >= 0xfffffff0ull)
    return ctx->fail(PRF_E_UNSUPPORTED, "shared-bipartition structure does not fit 32-bit offsets");

  // majority flip (only when some list is held by more than half of the trees): bitmaps of the holders
  const uint32_t n_heavy = flip_ok && T ? hc.n_heavy : 0;
  const uint32_t fwords = (T + 31) / 32;
  DevBuf<uint32_t> fid, fll, fbits, hcnt;
  if (n_heavy) {
    PRF_CK(ctx, fid.alloc(Lmax, s));
    PRF_CK(ctx, fll.alloc(n_heavy, s));
    PRF_CK(ctx, fbits.alloc((size_t)n_heavy * fwords, s));
    PRF_CK(ctx, hcnt.alloc((size_t)T + 1, s));
    PRF_CK(ctx, cudaMemsetAsync(fbits.p, 0, fbits.bytes(), s));
    if ((rc = exclusive_scan(ctx, IsHeavy{lcnt.p, cnts.p, T, short_max}, Lmax, fid.p, nullptr, &cnts.p->n_lists))) return rc;
    flip_ids_kernel<<<grid_for(hc.n_lists), 256, 0, s>>>(lcnt.p, llid.p, fid.p, cnts.p, T, short_max, fll.p);
    PRF_LAUNCHED(ctx);
  }
  H.n_flipped = n_heavy;
  const uint64_t n_buckets = (uint64_t)hc.n_long * NB;
  const uint64_t lmem_cap = hc.long_members + 7 * n_buckets + 8;
  DevBuf<uint32_t> bcnt, sraw, scount, scur;
  PRF_CK(ctx, bcnt.alloc(n_buckets + 1, s));
  PRF_CK(ctx, sraw.alloc((size_t)hc.short_members + 1, s));
  PRF_CK(ctx, scount.alloc((size_t)T + 1, s));
  PRF_CK(ctx, scur.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.boff.alloc(n_buckets + 1, s));
  PRF_CK(ctx, H.lmem.alloc(lmem_cap, s));
  PRF_CK(ctx, H.roff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.rlist.alloc(hc.long_members + 1, s));
  PRF_CK(ctx, H.soff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, H.singles.alloc(hc.short_pairs + 1, s));
  PRF_CK(ctx, H.nb.alloc((size_t)T + 16, s));  // slack: the matrix kernel reads |B| in aligned 32-bit pairs
  PRF_CK(ctx, cudaMemsetAsync(H.nb.p, 0, H.nb.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(bcnt.p, 0, bcnt.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(scount.p, 0, scount.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(scur.p, 0, scur.bytes(), s));
  PRF_CK(ctx, cudaMemsetAsync(H.lmem.p, 0xff, H.lmem.bytes(), s));
  if (T) {
    if (narrow)
      bucket_count_kernel<8><<<grid_t, 256, 0, s>>>(tb, te, T, H.ent_lid.p, lcnt.p, llid.p, sbase.p, short_max, H.block_log2, NB,
                                                    bcnt.p, sfill.p, sraw.p, rcnt.p, fid.p, fbits.p, fwords, hcnt.p);
    else
      bucket_count_kernel<32><<<grid_t, 256, 0, s>>>(tb, te, T, H.ent_lid.p, lcnt.p, llid.p, sbase.p, short_max, H.block_log2, NB,
                                                     bcnt.p, sfill.p, sraw.p, rcnt.p, fid.p, fbits.p, fwords, hcnt.p);
    PRF_LAUNCHED(ctx);
    if (n_heavy) {
      flip_buckets_kernel<<<(n_heavy * NB + 7) / 8, 256, 0, s>>>(fbits.p, fwords, fll.p, n_heavy, T, H.block_log2, NB, 0, bcnt.p,
                                                                 nullptr, nullptr);
      PRF_LAUNCHED(ctx);
      PRF_CK(ctx, H.nb0.alloc((size_t)T + 16, s));
      PRF_CK(ctx, cudaMemsetAsync(H.nb0.p, 0, H.nb0.bytes(), s));
      nb16_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, T, H.nb0.p);
      PRF_LAUNCHED(ctx);
      flip_trees_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, rcnt.p, hcnt.p, T, n_heavy);
      PRF_LAUNCHED(ctx);
    }
  }
  trace_sync(s, "build:   memsets + bucket_count");
  if ((rc = exclusive_scan(ctx, Pad8{bcnt.p, n_buckets}, n_buckets + 1, H.boff.p, &cnts.p->n_lmem))) return rc;
  if ((rc = exclusive_scan(ctx, LoadU32{rcnt.p}, (uint64_t)T + 1, H.roff.p, nullptr))) return rc;
  if (T) {
    if (narrow)
      bucket_fill_kernel<8><<<grid_t, 256, 0, s>>>(tb, te, T, H.ent_lid.p, lcnt.p, llid.p, short_max, H.block_log2, NB, bcnt.p,
                                                   H.boff.p, H.lmem.p, H.roff.p, H.rlist.p, n_heavy ? 1 : 0);
    else
      bucket_fill_kernel<32><<<grid_t, 256, 0, s>>>(tb, te, T, H.ent_lid.p, lcnt.p, llid.p, short_max, H.block_log2, NB, bcnt.p,
                                                    H.boff.p, H.lmem.p, H.roff.p, H.rlist.p, n_heavy ? 1 : 0);
    PRF_LAUNCHED(ctx);
    if (n_heavy) {
      flip_buckets_kernel<<<(n_heavy * NB + 7) / 8, 256, 0, s>>>(fbits.p, fwords, fll.p, n_heavy, T, H.block_log2, NB, 1, bcnt.p,
                                                                 H.boff.p, H.lmem.p);
      PRF_LAUNCHED(ctx);
      flip_rlist_kernel<<<(T + 7) / 8, 256, 0, s>>>(fbits.p, fwords, fll.p, n_heavy, T, H.roff.p, hcnt.p, H.rlist.p);
      PRF_LAUNCHED(ctx);
    }
    if (hc.short_members) {
      singles_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(lcnt.p, sbase.p, sraw.p, short_max, cnts.p, 0, scount.p, nullptr, nullptr, nullptr);
      PRF_LAUNCHED(ctx);
    }
  }
  trace_sync(s, "build:   scans + bucket_fill + singles count");
  if ((rc = exclusive_scan(ctx, LoadU32{scount.p}, (uint64_t)T + 1, H.soff.p, nullptr))) return rc;
  if (T) {
    if (hc.short_members) {
      singles_kernel<<<ctx->sm_count * 4, 256, 0, s>>>(lcnt.p, sbase.p, sraw.p, short_max, cnts.p, 1, nullptr, H.soff.p, scur.p, H.singles.p);
      PRF_LAUNCHED(ctx);
    }
    finish_trees_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, rcnt.p, T, H.nb.p, cnts.p);
    PRF_LAUNCHED(ctx);
  }
  PRF_CK(ctx, cudaMemcpyAsync(&hc, cnts.p, sizeof hc, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  trace_mark("build: structure synced");
  if (T == 0) hc.nb_min = hc.nb_max = 0;
  if (hc.nb_max > kMaxBipsPerTree)
    return ctx->fail(PRF_E_RANGE, "a tree has %u bipartitions; distances would overflow uint16 (max %u per tree)", hc.nb_max,
                     kMaxBipsPerTree);
  H.uniform_nb = hc.nb_min == hc.nb_max;
  H.nb_uniform = hc.nb_min;
  H.n_shared = hc.n_lists2;
  H.n_list_ids = hc.n_lists;
  H.n_shared_nnz = hc.n_entries;
  H.n_long = hc.n_long;
  H.n_lmem = hc.n_lmem;
  H.rlist_n = hc.long_members;
  H.singles_n = hc.short_pairs;
  H.max_lists = hc.max_lists;
  H.n_parts = 1;
  H.part[0] = ListPart{H.roff.p, H.rlist.p, H.boff.p, H.lmem.p, H.soff.p, H.singles.p};

  prf_stats& st = H.stats;
  st = prf_stats{};
  st.n_trees = T;
  st.n_words = W;
  st.nnz = nnz;
  st.n_unique = hcount[0];
  st.n_shared = H.n_shared;
  st.n_shared_nnz = H.n_shared_nnz;
  st.n_hits = hc.n_hits;
  st.n_pairs = H.P;
  st.max_bips = hc.nb_max;
  st.min_bips = hc.nb_min;
  st.n_flipped = H.n_flipped;
  if ((rc = build_nb_shift(ctx))) return rc;
  H.loaded = true;
  return PRF_OK;
}

}  // namespace prf
