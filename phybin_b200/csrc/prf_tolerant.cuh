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
//
// Most repeated keys of a pair come from pruning itself: a node whose siblings lose all their taxa has the
// same restricted set as its parent.  Those are removed BEFORE the table (a clade whose restricted set equals
// that of its parent clade is skipped: pure set semantics, the parent or the top of the chain stands for it),
// with the parent of every clade found once at load time (tol_parent_kernel); the equality path of the
// table (rebuild the other key, compare) then runs only for the rare real coincidences.
//
// Trees that REPEAT a leaf label (DUPS instantiation; prf_tolerant_load_dups): numLeaves counts such leaves
// twice (CoreTypes.hs:289-291), so normBip's size is per tree k_t = #surviving leaves, and "the node is the
// pruned tree's root" is decided by the label set of the leaves outside the node instead of by its own set.
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
// The EQ case (RFDistance.hs:239): min m fl on ascending element lists; m and fl are disjoint so their first
// elements decide, and the empty list is smallest.  Rare: kept out of line.
template <int W>
__device__ __noinline__ bool tie_takes_flipped(const BitSet<W> m, const BitSet<W> fl) {
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const uint64_t u = m.w[i] | fl.w[i];
    if (u) return (fl.w[i] & (u & (0 - u))) != 0;
  }
  return true;
}

template <int W>
__device__ __forceinline__ uint32_t norm_bip(const BitSet<W>& m, uint32_t pc, const BitSet<W>& R, uint32_t half,
                                             BitSet<W>& out) {
  out = m;
  if (pc < half) return pc;
  BitSet<W> fl;
#pragma unroll
  for (int i = 0; i < W; ++i) fl.w[i] = R.w[i] & ~m.w[i];  // invertDense k m
  const uint32_t fc = popc_set<W>(fl);
  if (pc == half && fc != 0 && !tie_takes_flipped<W>(m, fl)) return pc;
  out = fl;
  return fc;
}

// {0 .. k-1}
template <int W>
__device__ __forceinline__ BitSet<W> universe(uint32_t k) {
  BitSet<W> R;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int64_t rem = (int64_t)k - 64 * i;
    R.w[i] = rem >= 64 ? ~0ull : (rem <= 0 ? 0ull : ((1ull << rem) - 1));
  }
  return R;
}

template <int W>
__device__ __forceinline__ uint32_t hash_set(const BitSet<W>& s) {
  // routes only (equality is decided on the full key): two independent multiply-add chains over the 32-bit half
  // words, one avalanche round
  uint32_t h0 = 0x9E3779B9u, h1 = 0x85EBCA6Bu;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    h0 = h0 * 0xC2B2AE35u + (uint32_t)s.w[i];
    h1 = h1 * 0x27D4EB2Fu + (uint32_t)(s.w[i] >> 32);
  }
  uint32_t h = h0 ^ (h1 * 0x165667B1u);
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  h ^= h >> 16;
  return h;
}

struct TolParams {
  uint32_t T;
  uint64_t p_begin, p_end;
  const uint64_t* clades;    // [nnz * W]
  const uint64_t* outsets;   // [nnz * W]  DUPS only
  const uint16_t* par;       // [nnz] parent clade inside the tree, 0xffff = the root  (!DUPS only)
  const uint64_t* off;       // [T + 1]
  const uint64_t* leafsets;  // [T * W]
  const uint64_t* dup_off;   // [T + 1]    DUPS only
  const uint32_t* dup_label;
  const uint32_t* dup_extra;
  uint32_t tsize;
  int32_t editdist;
  uint16_t* out;
  uint32_t* mask;
};

// Everything a lane needs to (re)build the canonical set of any key of the pair.
template <int W, bool DUPS>
struct PairCtx {
  const uint64_t* clades;
  uint32_t oa, ob;   // first clade of tree a / tree b (nnz < 2^32 is checked at load: 32-bit index arithmetic)
  uint32_t na, nb;   // clade counts
  BitSet<W> S;       // shared taxa
  BitSet<W> R;       // normBip's universe {0..k-1} of side a (both sides when !DUPS)
  uint32_t ns;       // |S|
  uint32_t ka, kb;   // normBip's size per side = leaves of the pruned tree (= ns without repeated labels)
  bool fullA, fullB; // the tree's taxon set equals S: the reference uses its cached, un-pruned bips
};

// Key number idx of the pair: idx < na -> clade idx of tree a; idx < na+nb -> clade idx-na of tree b; beyond -> a
// singleton of S (only inserted when the side's k <= 3): its (idx-na-nb)-th taxon, for side a, or with DUPS the
// (idx-na-nb-ns)-th for side b.  Rebuilds the canonical set of a key that WAS inserted (no keep tests).
template <int W, bool DUPS>
__device__ __forceinline__ void rebuild_key(const PairCtx<W, DUPS>& P, uint32_t idx, BitSet<W>& key) {
  BitSet<W> m;
  uint32_t pc;
  bool isb;
  if (idx < P.na + P.nb) {
    isb = idx >= P.na;
    const uint64_t* src = P.clades + (size_t)(isb ? P.ob + (idx - P.na) : P.oa + idx) * W;
#pragma unroll
    for (int i = 0; i < W; ++i) m.w[i] = src[i] & P.S.w[i];
    pc = popc_set<W>(m);
  } else {
    uint32_t j = idx - P.na - P.nb, seen = 0;
    isb = DUPS && j >= P.ns;
    if (isb) j -= P.ns;
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
  const uint32_t k = isb ? P.kb : P.ka;
  if (DUPS) norm_bip<W>(m, pc, universe<W>(k), k >> 1, key);
  else norm_bip<W>(m, pc, P.R, k >> 1, key);
}

template <int W, bool DUPS>
__device__ __forceinline__ bool key_equals(const PairCtx<W, DUPS>& P, uint32_t idx, const BitSet<W>& key) {
  BitSet<W> other;
  rebuild_key<W, DUPS>(P, idx, other);
  bool eq = true;
#pragma unroll
  for (int i = 0; i < W; ++i) eq &= (other.w[i] == key.w[i]);
  return eq;
}

// slot word: [31:30] flags (1 = in a, 2 = in b), [29:11] tag, [10:0] key number + 1; 0 = empty.
// The table holds no keys: on a tag match the other key is rebuilt from its number (its raw clade is
// an L1/L2 hit), so a warp needs only 4 bytes of shared memory per slot and many warps fit on an SM.
// Returns, packed, what the insertion changed: bit 0 / bit 1 = the slot gained flag 1 / 2 for the first time, bit 2 = the
// slot now has both flags and did not before.  Summed over a pair's keys: |F_a|, |F_b| and |F_a n F_b|.
__device__ __forceinline__ uint32_t tol_lds(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t tol_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t tol_or(uint32_t addr, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.or.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(val) : "memory");
  return old;
}

// tbase: shared-window byte address of the warp's table; bmask = (tsize - 1) * 4.  The probe loop carries two
// registers (the slot's byte offset and the word to store); the asm barrier below stops ptxas from re-deriving them
// -- i.e. from re-hashing the key -- on every probe, which it otherwise does under the 64-register cap (ncu: the
// loop was 45 instructions long and 40 % of the kernel).
template <int W, bool DUPS>
__device__ __forceinline__ uint32_t table_insert(uint32_t tbase, uint32_t bmask, const PairCtx<W, DUPS>& P, uint32_t idx,
                                                 const BitSet<W>& key, uint32_t flag) {
  const uint32_t h = hash_set<W>(key);
  uint32_t mine = (flag << 30) | (((h >> 13) & 0x7ffffu) << 11) | (idx + 1);
  uint32_t s4 = (h << 2) & bmask;
  asm volatile("" : "+r"(mine), "+r"(s4));
  uint32_t old = 0;  // flags the key had before this insertion
  for (;;) {         // nothing but the probe in the loop: what the insertion changed is worked out once, after it
    uint32_t cur = tol_lds(tbase + s4);
    if (cur == 0) {
      cur = tol_cas(tbase + s4, 0u, mine);
      if (cur == 0) break;
    }
    if (((cur ^ mine) & 0x3ffff800u) == 0 && key_equals<W, DUPS>(P, (cur & 0x7ffu) - 1, key)) {
      old = tol_or(tbase + s4, flag << 30) >> 30;
      break;
    }
    s4 = (s4 + 4) & bmask;
  }
  return (flag & ~old) | ((old | flag) == 3u && old != 3u ? 4u : 0u);
}

// Parent of every clade inside its tree, found once per load: the smallest other clade of the tree that contains it
// (order (size, index), so equal clades chain upwards and never cycle); 0xffff when only the root contains it.
// One warp per tree, lanes over the clades; the inner loop reads the same clade in every lane (broadcast).
template <int W>
__global__ void __launch_bounds__(256) tol_parent_kernel(uint32_t T, const uint64_t* __restrict__ clades,
                                                         const uint64_t* __restrict__ off, uint16_t* __restrict__ par) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= T) return;
  const uint64_t o = off[t];
  const uint32_t n = (uint32_t)(off[t + 1] - o);
  for (uint32_t i = lane; i < n; i += 32) {
    BitSet<W> c;
#pragma unroll
    for (int w = 0; w < W; ++w) c.w[w] = clades[(o + i) * W + w];
    const uint32_t pci = popc_set<W>(c);
    uint32_t best = 0xffffu, best_pc = 0xffffffffu;
    for (uint32_t j = 0; j < n; ++j) {
      BitSet<W> x;
      bool sub = true;
#pragma unroll
      for (int w = 0; w < W; ++w) {
        x.w[w] = clades[(o + j) * W + w];
        sub &= (c.w[w] & ~x.w[w]) == 0;
      }
      const uint32_t pcj = popc_set<W>(x);
      const bool above = pcj > pci || (pcj == pci && j > i);
      if (sub && above && pcj < best_pc) { best = j; best_pc = pcj; }   // ascending j: the first of equal candidates
    }
    par[o + i] = (uint16_t)best;
  }
}

template <int W, bool DUPS>
__global__ void __launch_bounds__(512, (W <= 3 && !DUPS) ? 2 : 1) tolerant_kernel(const TolParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t warps = blockDim.x >> 5;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tsize = prm.tsize, T = prm.T;
  const uint64_t p_begin = prm.p_begin, p_end = prm.p_end;
  const uint64_t* __restrict__ clades = prm.clades;
  const uint64_t* __restrict__ leafsets = prm.leafsets;
  const uint64_t* __restrict__ off = prm.off;
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw) + (size_t)warp * tsize;  // per-warp tables
  uint32_t tbase = (uint32_t)__cvta_generic_to_shared(table), bmask = (tsize - 1) * 4;
  asm volatile("" : "+r"(tbase), "+r"(bmask));  // kept in registers (ptxas otherwise re-derives the address on every probe)
  for (uint32_t s = lane; s < tsize; s += 32) table[s] = 0;
  __syncwarp();

  const uint32_t per_cta = warps * kTolPairsPerWarp;
  const uint64_t n_blocks = (p_end - p_begin + per_cta - 1) / per_cta;
  for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const uint64_t p0 = p_begin + blk * per_cta + (uint64_t)warp * kTolPairsPerWarp;  // (p0 - p_begin) is a multiple of 8
    if (p0 >= p_end) continue;  // (warp-uniform; no CTA-wide barrier in this kernel)
    uint32_t my_d = 0;
    // the warp's pairs are consecutive positions of the packed triangle: (a, b), (a, b+1), ... wrapping to the next row
    uint32_t a = row_of(p0, T);
    uint32_t b = a + 1 + (uint32_t)(p0 - row_start(a, T));
    for (int q = 0; q < kTolPairsPerWarp; ++q, ++b) {
      if (p0 + q >= p_end) break;  // (warp-uniform)
      if (b == T) { ++a; b = a + 1; }
      PairCtx<W, DUPS> P;
      P.clades = clades;
      P.fullA = P.fullB = true;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const uint64_t la = leafsets[(size_t)a * W + i], lb = leafsets[(size_t)b * W + i];
        P.S.w[i] = la & lb;
        P.fullA &= la == P.S.w[i];
        P.fullB &= lb == P.S.w[i];
      }
      P.ns = popc_set<W>(P.S);
      uint32_t d = 0;
      if (P.ns != 0) {  // inBoth empty -> 0 (RFDistance.hs:195-196)
        P.ka = P.kb = P.ns;
        if (DUPS) {  // numLeaves of the pruned trees: repeated labels that survive count again (CoreTypes.hs:289-291)
          for (uint64_t x = prm.dup_off[a]; x < prm.dup_off[a + 1]; ++x) {
            const uint32_t l = prm.dup_label[x];
            if ((P.S.w[l >> 6] >> (l & 63)) & 1) P.ka += prm.dup_extra[x];
          }
          for (uint64_t x = prm.dup_off[b]; x < prm.dup_off[b + 1]; ++x) {
            const uint32_t l = prm.dup_label[x];
            if ((P.S.w[l >> 6] >> (l & 63)) & 1) P.kb += prm.dup_extra[x];
          }
        }
        P.R = universe<W>(P.ka);
        P.oa = (uint32_t)off[a];
        P.na = (uint32_t)off[a + 1] - P.oa;
        P.ob = (uint32_t)off[b];
        P.nb = (uint32_t)off[b + 1] - P.ob;
        const uint32_t n_cl = P.na + P.nb;
        for (uint32_t c0 = 0; c0 < n_cl; c0 += 32) {
          const uint32_t idx = c0 + lane;
          BitSet<W> key;
          bool have = false;
          bool isb = false;
          if (idx < n_cl) {
            isb = idx >= P.na;
            const uint32_t e = isb ? P.ob + (idx - P.na) : P.oa + idx;
            const bool full = isb ? P.fullB : P.fullA;
            BitSet<W> m;
#pragma unroll
            for (int i = 0; i < W; ++i) m.w[i] = clades[(size_t)e * W + i] & P.S.w[i];
            const uint32_t pc = popc_set<W>(m);
            bool keep = pc != 0;  // the node lost all its leaves
            if (DUPS) {
              // pruneTreeLeaves splices the node into the pruned tree's root when no leaf outside it survives
              // (PreProcessor.hs:28-31); the cached bips of an unpruned tree keep every node (RFDistance.hs:186-191)
              uint64_t outside = 0;
#pragma unroll
              for (int i = 0; i < W; ++i) outside |= prm.outsets[(size_t)e * W + i] & P.S.w[i];
              keep &= full || outside != 0;
            } else {
              const uint32_t pr = prm.par[e];
              if (pr == 0xffffu) {
                keep &= full || pc != P.ns;   // the whole of S: the pruned tree's root
              } else {
                const uint32_t pe = (isb ? P.ob : P.oa) + pr;
                bool same = true;
#pragma unroll
                for (int i = 0; i < W; ++i) same &= (clades[(size_t)pe * W + i] & P.S.w[i]) == m.w[i];
                keep &= !same;                // same set as its parent clade: the parent stands for both
              }
            }
            if (keep) {
              const uint32_t k = isb ? P.kb : P.ka;
              const uint32_t n = DUPS ? norm_bip<W>(m, pc, universe<W>(k), k >> 1, key) : norm_bip<W>(m, pc, P.R, k >> 1, key);
              have = n > 1;   // allBips' filter (RFDistance.hs:248)
            }
          }
          if (have) {
            const uint32_t g = table_insert<W, DUPS>(tbase, bmask, P, idx, key, isb ? 2u : 1u);
            d += (g & 1u) + ((g >> 1) & 1u) - ((g >> 1) & 2u);  // |F_a| + |F_b| - 2 |F_a n F_b| = |F_a xor F_b|
          }
        }
        // the leaves' singletons only survive the filter when k <= 3 (a singleton stays one for k >= 4); every taxon of S
        // is a leaf of both pruned trees; they still matter because a clade may normalise to the same set
        if (P.ka <= 3 || P.kb <= 3) {
          const uint32_t n_single = DUPS ? 2 * P.ns : P.ns;
          if (lane < n_single) {
            const bool sideb = DUPS && lane >= P.ns;
            const uint32_t k = sideb ? P.kb : P.ka;
            if (k <= 3) {
              BitSet<W> key;
              rebuild_key<W, DUPS>(P, n_cl + lane, key);
              if (popc_set<W>(key) > 1) {
                const uint32_t g = table_insert<W, DUPS>(tbase, bmask, P, n_cl + lane, key, DUPS ? (sideb ? 2u : 1u) : 3u);
                d += (g & 1u) + ((g >> 1) & 1u) - ((g >> 1) & 2u);
              }
            }
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
    {
      const uint32_t nv = (uint32_t)min((uint64_t)kTolPairsPerWarp, p_end - p0);
      uint32_t v[kTolPairsPerWarp];
#pragma unroll
      for (int q = 0; q < kTolPairsPerWarp; ++q) v[q] = __shfl_sync(0xffffffffu, my_d, q);
      if (lane == 0) {
        uint16_t* o = prm.out + (p0 - p_begin);
        if (nv == kTolPairsPerWarp && (reinterpret_cast<uintptr_t>(o) & 15u) == 0) {
          *reinterpret_cast<uint4*>(o) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
        } else {
          for (uint32_t q = 0; q < nv; ++q) o[q] = (uint16_t)v[q];
        }
        if (prm.editdist >= 0) {
          uint32_t bits = 0;
          for (uint32_t q = 0; q < nv; ++q) bits |= ((int32_t)v[q] <= prm.editdist ? 1u : 0u) << q;
          reinterpret_cast<uint8_t*>(prm.mask)[(p0 - p_begin) >> 3] = (uint8_t)bits;  // every mask byte has one writer
        }
      }
    }
  }
}

template <int W, bool DUPS>
static int launch_tolerant_wd(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  const TolerantState& S = ctx->tol;
  const uint32_t key_cap = 2 * S.max_clades + 8;
  uint32_t tsize = 64;
  // load factor <= 1/3 even if no key repeats: a warp's 32 insertions finish with the longest of their probe chains, and
  // the registers (2 CTAs of 16 warps per SM) leave the shared memory for it
  while (tsize < 3 * key_cap) tsize <<= 1;
  if (tsize > 2048) tsize >>= 1;  // (very large trees: back to <= 2/3)
  uint32_t warps = 16;
  while (warps > 4 && (size_t)warps * tsize * 4 > 96 * 1024) warps -= 4;
  const size_t smem = (size_t)warps * tsize * 4;
  PRF_CK(ctx, cudaFuncSetAttribute(tolerant_kernel<W, DUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int ctas = 0;
  PRF_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, tolerant_kernel<W, DUPS>, (int)(warps * 32), smem));
  if (ctas < 1) ctas = 1;
  const uint32_t per_cta = warps * kTolPairsPerWarp;
  const uint64_t n_blocks = (p_end - p_begin + per_cta - 1) / per_cta;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(n_blocks, (uint64_t)ctx->sm_count * ctas * 8);
  // every mask byte is written by the warp that owns its 8 pairs; the bytes after p_end in the last word are not
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask + (p_end - p_begin) / 32, 0, ((p_end - p_begin) % 32) ? 4 : 0, ctx->stream));
  TolParams prm;
  prm.T = S.T;
  prm.p_begin = p_begin;
  prm.p_end = p_end;
  prm.clades = S.clades.p;
  prm.outsets = S.outsets.p;
  prm.par = S.par.p;
  prm.off = S.off.p;
  prm.leafsets = S.leafsets.p;
  prm.dup_off = S.dup_off.p;
  prm.dup_label = S.dup_label.p;
  prm.dup_extra = S.dup_extra.p;
  prm.tsize = tsize;
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  tolerant_kernel<W, DUPS><<<grid, warps * 32, smem, ctx->stream>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

template <int W>
static int launch_tolerant_w(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  return ctx->tol.dups ? launch_tolerant_wd<W, true>(ctx, p_begin, p_end, editdist, d_out, d_mask)
                       : launch_tolerant_wd<W, false>(ctx, p_begin, p_end, editdist, d_out, d_mask);
}

template <int W>
static int launch_tol_parents_w(Ctx* ctx) {
  TolerantState& S = ctx->tol;
  if (!S.T) return PRF_OK;
  tol_parent_kernel<W><<<(S.T + 7) / 8, 256, 0, ctx->stream>>>(S.T, S.clades.p, S.off.p, S.par.p);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

int launch_tol_parents(Ctx* ctx) {
  switch (ctx->tol.W) {
    case 1: return launch_tol_parents_w<1>(ctx);
    case 2: return launch_tol_parents_w<2>(ctx);
    case 3: return launch_tol_parents_w<3>(ctx);
    case 4: return launch_tol_parents_w<4>(ctx);
    case 5: return launch_tol_parents_w<5>(ctx);
    case 6: return launch_tol_parents_w<6>(ctx);
    case 7: return launch_tol_parents_w<7>(ctx);
    case 8: return launch_tol_parents_w<8>(ctx);
    default: return ctx->fail(PRF_E_UNSUPPORTED, "tolerant mode supports W in 1..%u", kTolMaxWords);
  }
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
