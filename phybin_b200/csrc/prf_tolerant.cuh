// prf_tolerant.cuh -- naiveDistMatrix (RFDistance.hs:168-198), the --tolerant path.
//
// Per pair (a, b) the reference prunes both trees to the shared taxa S (pruneTreeLeaves,
// PreProcessor.hs:21-32), recomputes allBips with size = |S| (RFDistance.hs:207-248) unless the
// tree already has exactly S (cached bips, :186-191), and counts the symmetric difference of the two
// bipartition sets (:193-194).  On bit sets that is:
//     raw_t = { c & S : c clade of a non-root node of t } minus {0} minus {S unless L_t == S}
//     F_t   = { normBip_k(r) : r in raw_t, |normBip_k(r)| > 1 }            (a SET: dedup needed)
//     d     = |F_a xor F_b|
// where normBip_k complements inside {0..k-1} (NOT inside S: invertDense, :127-131) and breaks
// |r| == k/2 ties with IntSet's lexicographic order (:239).  Distinct clades can collapse to one
// masked set and the two sides of one split normalise to the same set, so a masked-popcount Gram
// product is not equivalent; each pair needs an exact set computation.
//
// One warp per pair.  Lanes normalise the clades of both trees and insert them into a small
// per-warp open-addressing table in shared memory (hash routes, equality is decided on the full
// key, rebuilt on demand from the key's number so that the table is 4 bytes per slot and ~32 warps
// fit on an SM); each slot accumulates "seen in a" / "seen in b" flags; d = number of slots with
// exactly one.
#pragma once

#include "prf_common.cuh"

namespace prf {

constexpr uint32_t kTolMaxWords = 8;
constexpr uint32_t kTolMaxClades = 1000;  // per tree; slot word keeps an 11-bit key index
constexpr int kTolPairsPerWarp = 8;

template <int W>
struct BitSet {
  uint64_t w[W];
};

template <int W>
__device__ __forceinline__ uint32_t popc_set(const BitSet<W>& s) {
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < W; ++i) c += __popcll(s.w[i]);
  return c;
}

// normBip (RFDistance.hs:229-239) with universe R = {0..k-1}.  Returns the cardinality of the
// canonical side written to `out`.
template <int W>
__device__ __forceinline__ uint32_t norm_bip(const BitSet<W>& m, uint32_t pc, const BitSet<W>& R, uint32_t half,
                                             BitSet<W>& out) {
  if (pc < half) {
    out = m;
    return pc;
  }
  BitSet<W> fl;
#pragma unroll
  for (int i = 0; i < W; ++i) fl.w[i] = R.w[i] & ~m.w[i];  // invertDense k m
  bool take_fl = true;
  if (pc == half) {
    // min m fl on ascending element lists; m and fl are disjoint so their first elements decide,
    // and the empty list is smallest.
    bool decided = false;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      uint64_t u = m.w[i] | fl.w[i];
      if (!decided && u) {
        uint64_t low = u & (0 - u);
        take_fl = (fl.w[i] & low) != 0;
        decided = true;
      }
    }
    uint32_t fc = popc_set<W>(fl);
    if (fc == 0) take_fl = true;
  }
  out = take_fl ? fl : m;
  return popc_set<W>(out);
}

template <int W>
__device__ __forceinline__ uint32_t hash_set(const BitSet<W>& s) {
  // routes only (equality is decided on the full key): a multiply-add per 32-bit half word, then an avalanche
  uint32_t h = 0x9E3779B9u;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    h = h * 0x85EBCA6Bu + (uint32_t)s.w[i];
    h = h * 0xC2B2AE35u + (uint32_t)(s.w[i] >> 32);
  }
  h ^= h >> 16;
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

// Everything a lane needs to (re)build the canonical set of any clade of the pair.
template <int W>
struct PairCtx {
  const uint64_t* clades;
  uint64_t oa, ob;   // first clade of tree a / tree b
  uint32_t na, nb;   // clade counts
  BitSet<W> S, R;    // shared taxa; normBip's universe {0..k-1}
  uint32_t k, half;
  bool fullA, fullB; // the tree's taxon set equals S: the reference uses its cached, un-pruned bips
};

// Key number idx of the pair: idx < na -> clade idx of tree a; idx < na+nb -> clade idx-na of tree b;
// beyond -> the (idx-na-nb)-th taxon of S as a singleton (only reachable when k <= 3).
// Returns false when the reference would not keep this set (empty, the whole of S after pruning, or
// size <= 1 after normalisation).
template <int W>
__device__ __forceinline__ bool make_key(const PairCtx<W>& P, uint32_t idx, BitSet<W>& key) {
  BitSet<W> m;
  uint32_t pc;
  if (idx < P.na + P.nb) {
    const bool isb = idx >= P.na;
    const uint64_t* src = P.clades + (isb ? P.ob + (idx - P.na) : P.oa + idx) * W;
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = src[i] & P.S.w[i];
    pc = popc_set<W>(m);
    // drop the empty set, and the full set unless the cached (unpruned) bips are used
    if (pc == 0 || (pc == P.k && !(isb ? P.fullB : P.fullA))) return false;
  } else {
    uint32_t j = idx - P.na - P.nb, seen = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      m.w[i] = 0;
      uint64_t word = P.S.w[i];
      while (word) {
        const uint64_t low = word & (0 - word);
        if (seen == j) m.w[i] = low;
        ++seen;
        word ^= low;
      }
    }
    pc = 1;
  }
  return norm_bip<W>(m, pc, P.R, P.half, key) > 1;
}

// slot word: [31:30] flags (1 = in a, 2 = in b), [29:11] tag, [10:0] key number + 1; 0 = empty.
// The table holds no keys: on a tag match the other key is rebuilt from its number (its raw clade is
// an L1/L2 hit), so a warp needs only 4 bytes of shared memory per slot and many warps fit on an SM.
// Returns, packed, what the insertion changed: bit 0 / bit 1 = the slot gained flag 1 / 2 for the first time, bit 2 = the
// slot now has both flags and did not before.  Summed over a pair's keys: |F_a|, |F_b| and |F_a n F_b|.
template <int W>
__device__ __forceinline__ uint32_t table_insert(uint32_t* table, uint32_t tmask, const PairCtx<W>& P, uint32_t idx,
                                                 const BitSet<W>& key, uint32_t flag) {
  const uint32_t h = hash_set<W>(key);
  const uint32_t tag = (h >> 13) & 0x7ffffu;
  const uint32_t mine = (flag << 30) | (tag << 11) | (idx + 1);
  uint32_t s = h & tmask;
  for (;;) {
    uint32_t cur = *((volatile uint32_t*)&table[s]);
    if (cur == 0) {
      cur = atomicCAS(&table[s], 0u, mine);
      if (cur == 0) return flag | (flag == 3u ? 4u : 0u);
    }
    if (((cur >> 11) & 0x7ffffu) == tag) {
      BitSet<W> other;
      make_key<W>(P, (cur & 0x7ffu) - 1, other);  // it was inserted, so it is a kept key
      bool eq = true;
#pragma unroll
      for (int i = 0; i < W; ++i) eq &= (other.w[i] == key.w[i]);
      if (eq) {
        const uint32_t old = atomicOr(&table[s], flag << 30) >> 30;
        const uint32_t gained = flag & ~old;
        return gained | ((old | flag) == 3u && old != 3u ? 4u : 0u);
      }
    }
    s = (s + 1) & tmask;
  }
}

template <int W>
__global__ void __launch_bounds__(512) tolerant_kernel(uint32_t T, uint64_t p_begin, uint64_t p_end,
                                                       const uint64_t* __restrict__ clades,
                                                       const uint64_t* __restrict__ off,
                                                       const uint64_t* __restrict__ leafsets, uint32_t tsize,
                                                       int32_t editdist, uint16_t* __restrict__ out,
                                                       uint32_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warps = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * tsize;  // per-warp tables
  const uint32_t tmask = tsize - 1;
  for (uint32_t s = lane; s < tsize; s += 32) table[s] = 0;
  __syncwarp();

  const uint32_t per_cta = warps * kTolPairsPerWarp;
  const uint64_t n_blocks = (p_end - p_begin + per_cta - 1) / per_cta;
  for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const uint64_t pb = p_begin + blk * per_cta;
    uint32_t my_d = 0;
    for (int q = 0; q < kTolPairsPerWarp; ++q) {
      const uint32_t local = warp * kTolPairsPerWarp + q;
      const uint64_t p = pb + local;
      if (p >= p_end) break;  // (warp-uniform)
      const uint32_t a = row_of(p, T);
      const uint32_t b = a + 1 + (uint32_t)(p - row_start(a, T));
      PairCtx<W> P;
      P.clades = clades;
      P.fullA = P.fullB = true;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const uint64_t la = leafsets[(size_t)a * W + i], lb = leafsets[(size_t)b * W + i];
        P.S.w[i] = la & lb;
        P.fullA &= la == P.S.w[i];
        P.fullB &= lb == P.S.w[i];
      }
      P.k = popc_set<W>(P.S);
      uint32_t d = 0;
      if (P.k != 0) {  // inBoth empty -> 0 (RFDistance.hs:195-196)
        P.half = P.k >> 1;
#pragma unroll
        for (int i = 0; i < W; ++i) {
          const int64_t rem = (int64_t)P.k - 64 * i;
          P.R.w[i] = rem >= 64 ? ~0ull : (rem <= 0 ? 0ull : ((1ull << rem) - 1));
        }
        P.oa = off[a];
        P.na = (uint32_t)(off[a + 1] - P.oa);
        P.ob = off[b];
        P.nb = (uint32_t)(off[b + 1] - P.ob);
        // every key of the pair: the clades of a (flag 1), of b (flag 2) and, when k <= 3, the
        // singletons of S (both pruned trees have every leaf of S: flag 3; they still matter because
        // another clade may normalise to the same set)
        const uint32_t n_keys = P.na + P.nb + (P.k <= 3 ? P.k : 0u);
        for (uint32_t c0 = 0; c0 < n_keys; c0 += 32) {
          const uint32_t idx = c0 + lane;
          BitSet<W> key;
          const bool have = idx < n_keys && make_key<W>(P, idx, key);
          if (have) {
            const uint32_t g = table_insert<W>(table, tmask, P, idx, key, idx < P.na ? 1u : (idx < P.na + P.nb ? 2u : 3u));
            d += (g & 1u) + ((g >> 1) & 1u) - ((g >> 1) & 2u);  // |F_a| + |F_b| - 2 |F_a n F_b| = |F_a xor F_b|
          }
        }
        __syncwarp();
        // clear the table for the next pair
        for (uint32_t s = lane; s < tsize / 4; s += 32) reinterpret_cast<uint4*>(table)[s] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int sh = 16; sh; sh >>= 1) d += __shfl_xor_sync(0xffffffffu, d, sh);
        __syncwarp();
      }
      // lane q keeps the q-th result of this warp
      if (lane == (uint32_t)q) my_d = d;
    }
    // the warp's kTolPairsPerWarp consecutive results: one 16-byte store and one mask byte, no CTA-wide barrier (warps of
    // a CTA differ in how long their pairs take; waiting for the slowest one was ~10 % of the kernel)
    const uint64_t p0 = pb + (uint64_t)warp * kTolPairsPerWarp;  // (p0 - p_begin) is a multiple of 8
    if (p0 < p_end) {
      const uint32_t nv = (uint32_t)min((uint64_t)kTolPairsPerWarp, p_end - p0);
      uint32_t v[kTolPairsPerWarp];
#pragma unroll
      for (int q = 0; q < kTolPairsPerWarp; ++q) v[q] = __shfl_sync(0xffffffffu, my_d, q);
      if (lane == 0) {
        uint16_t* o = out + (p0 - p_begin);
        if (nv == kTolPairsPerWarp && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
          *reinterpret_cast<uint4*>(o) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        } else {
          for (uint32_t q = 0; q < nv; ++q) o[q] = (uint16_t)v[q];
        }
        if (editdist >= 0) {
          uint32_t bits = 0;
          for (uint32_t q = 0; q < nv; ++q) bits |= ((int32_t)v[q] <= editdist ? 1u : 0u) << q;
          reinterpret_cast<uint8_t*>(mask)[(p0 - p_begin) >> 3] = (uint8_t)bits;  // every mask byte has one writer
        }
      }
    }
  }
}

template <int W>
static int launch_tolerant_w(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                             uint32_t* d_mask) {
  const TolerantState& S = ctx->tol;
  const uint32_t key_cap = 2 * S.max_clades + 4;
  uint32_t tsize = 64;
  // load factor <= 1/3 even if no key repeats: a warp's 32 insertions finish with the longest of their probe chains, and
  // the registers (2 CTAs of 16 warps per SM) leave the shared memory for it
  while (tsize < 3 * key_cap) tsize <<= 1;
  if (tsize > 2048) tsize >>= 1;  // (very large trees: back to <= 2/3)
  uint32_t warps = 16;
  while (warps > 4 && (size_t)warps * tsize * 4 > 96 * 1024) warps -= 4;
  const size_t smem = (size_t)warps * tsize * 4;
  PRF_CK(ctx, cudaFuncSetAttribute(tolerant_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas = 0;
  PRF_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, tolerant_kernel<W>, (int)(warps * 32), smem));
  if (ctas < 1) ctas = 1;
  const uint32_t per_cta = warps * kTolPairsPerWarp;
  const uint64_t n_blocks = (p_end - p_begin + per_cta - 1) / per_cta;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(n_blocks, (uint64_t)ctx->sm_count * ctas * 8);
  // every mask byte is written by the warp that owns its 8 pairs; the bytes after p_end in the last word are not
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask + (p_end - p_begin) / 32, 0, ((p_end - p_begin) % 32) ? 4 : 0, ctx->stream));
  tolerant_kernel<W><<<grid, warps * 32, smem, ctx->stream>>>(S.T, p_begin, p_end, S.clades.p, S.off.p, S.leafsets.p,
                                                             tsize, editdist, d_out, d_mask);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

int launch_tolerant(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                           uint32_t* d_mask) {
  switch (ctx->tol.W) {
    case 1: return launch_tolerant_w<1>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 2: return launch_tolerant_w<2>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 3: return launch_tolerant_w<3>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 4: return launch_tolerant_w<4>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 5: return launch_tolerant_w<5>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 6: return launch_tolerant_w<6>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 7: return launch_tolerant_w<7>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 8: return launch_tolerant_w<8>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    default: return ctx->fail(PRF_E_UNSUPPORTED, "tolerant mode supports W in 1..%u", kTolMaxWords);
  }
}

}  // namespace prf
