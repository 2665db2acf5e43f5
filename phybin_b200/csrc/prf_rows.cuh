// prf_rows.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), sparse variant.
//
// The reference bumps d(i,j) once for every bipartition that exactly one of the two trees has
// (haveIt x dontHave, :327-338): d(a,b) = |B_a| + |B_b| - 2 |B_a n B_b|.  Random topologies share
// ~0.13 bipartitions per pair, so almost every entry is just |B_a| + |B_b| and the work is writing
// the matrix: 2 bytes per pair, HBM-write bound.
//
// ONE WARP OWNS A ROW.  There is no block-wide or inter-warp synchronisation anywhere in the kernel:
// a warp takes the next row a (dynamic counter, longest rows first) and produces it one COLUMN BLOCK
// (B = 8192 consecutive trees, the bucket size of the build's list structure, ListPart in
// prf_common.cuh) at a time in its private shared-memory buffer:
//   fill     d = |B_a| + |B_b| for the tile's columns (16-byte shared stores; rows whose |B_t| differ:
//            one bulk async copy of |B|[c_lo..] from one of 8 pre-shifted copies + a SIMD add of |B_a|)
//   singles  bipartitions held by only a few trees (lists of <= short_max members) were expanded by the
//            build into an explicit per-row list of columns; each is one shared-memory reduction
//   lists    for every long list l containing a, the build left the members inside this column block as one
//            bucket of 16-byte groups (8 x uint16 column offsets, padded).  The groups of all of the row's
//            lists form one sequence that is cut into 32 equal runs, one per lane ("merge path": a warp scan
//            of the bucket sizes, a binary search per lane): the work per lane is balanced whatever the list
//            lengths are, every load is a full 16-byte group, and a member costs ~8 instructions: range test
//            (also rejects padding, columns <= a and columns outside a shard), address, one shared-memory
//            reduction -2 on the entry's half-word (two lists may hit the same column)
//   mask     optional d <= editdist bits
//   store    the finished tile leaves as ONE bulk asynchronous copy shared -> global (TMA engine,
//            16-byte aligned interior; the < 8 entries before / after it are stored by a few lanes)
// Every output byte is written exactly once.  The structure may come in several PARTS (multi-GPU
// build: one part per hash class of bipartitions, prf_team.cuh): a row's lists are the union over the
// parts, d = sum over parts.
#pragma once

#include <algorithm>

#include "prf_common.cuh"
#include "prf_async.cuh"
#include "prf_internal.h"  // build_row_plan

namespace prf {

struct RowsParams {
  uint32_t T;
  uint64_t p_begin;
  const uint16_t* nb;
  const uint16_t* nb_shift;   // 8 shifted copies of nb (non-uniform |B_t|), see HashRFState
  uint32_t nb_shift_stride;
  int uniform_nb;
  uint32_t nb_uniform;
  uint32_t NB;                // column blocks
  uint32_t n_parts;
  ListPart part[kMaxParts];
  int32_t editdist;
  uint16_t* out;
  uint32_t* mask;
  const RowItem* items;
  uint32_t n_items;
  uint32_t* counter;          // next item (zeroed before the launch)
#ifdef PRF_CALIB
  int calib;                  // calibration builds only (scripts/rows_calib.py): 1 = fill + store only, 2 = everything but the
                              // reductions add 0, 3 = no group loads, 4 = no reduction instructions, 5 = bounds + scan only
#endif
};

#ifdef PRF_CALIB
__device__ int g_calib_no_red;
#endif

// L2 residency: the list structure (~45 MB at 100k x 200) is re-read by every row that uses it and fits the 126 MB L2,
// but the 10 GB stream of matrix tiles passing through L2 evicts it; every evicted sector comes back as a small random
// DRAM read in the middle of the write stream (0.39 GB per launch, and the kernel 0.7 ms slower than on an input whose
// structure is tiny).  So the structure is loaded with an evict-last policy and the tiles are stored evict-first.
__device__ __forceinline__ uint32_t ldg_keep(const uint32_t* p, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ldg_keep(const uint4* p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
}

// tile entry of column offset `off` (relative to the block) -= 2 when `on`; tb = byte address of the entry of the
// block's column 0 (mod 2^32).  One shared-memory reduction on the entry's 32-bit word, issued UNCONDITIONALLY (a
// lane that is off adds 0 to the block's first word): ptxas turns a predicated reduction into a branch around the
// address arithmetic, which costs more than the idle lanes of a reduction (the reductions are not the bottleneck,
// scripts/rows_calib.py).
__device__ __forceinline__ void tile_red2(uint32_t tb, uint32_t safe, uint32_t off, bool on) {
#ifdef PRF_CALIB
  if (g_calib_no_red == 2) return;
  if (g_calib_no_red) on = false;
#endif
  const uint32_t A = tb + 2u * off;
  const uint32_t val = (A & 2u) ? 0xfffe0000u : 0xfffffffeu;
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(on ? (A & ~3u) : safe), "r"(on ? val : 0u) : "memory");
}

// kLog2B: log2 of the column block (= tile) size; kCap: lists of a row handled per pass; kQ: 16-byte groups a lane
// loads ahead of processing them.
template <int kLog2B, int kWarps, int kCap, int kQ>
__global__ void __launch_bounds__(32 * kWarps, 1) rows_kernel(const RowsParams prm) {
  constexpr int kTile = 1 << kLog2B;
  constexpr int kStride = kTile + 8;  // entries per buffer: a tile keeps the 16-byte phase of its global address
  constexpr int kWarpBytes = kStride * 2 + kCap * 12 + 16;  // per warp: tile buffer, then trow / start / upre of the row's lists
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_fill[kWarps];  // completion of a warp's bulk load of |B| (non-uniform rows)

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw + (size_t)warp * kWarpBytes);
  uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
  uint32_t* s_trow = reinterpret_cast<uint32_t*>(smem_raw + (size_t)warp * kWarpBytes + kStride * 2);  // part << 28 | first bucket of the list
  uint32_t* s_start = s_trow + kCap;   // first lmem entry of the list's bucket in the current block
  uint32_t* s_upre = s_start + kCap;   // [kCap + 1] 16-byte groups of the lists before this one (exclusive prefix)
  const uint32_t T = prm.T, NB = prm.NB;

  if (lane == 0) {
    mbar_init(&s_fill[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t fill_phase = 0;

  uint64_t pol_keep, pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  const uint32_t vv = (2u * prm.nb_uniform) | ((2u * prm.nb_uniform) << 16);
  const uint4 q4 = make_uint4(vv, vv, vv, vv);

  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(prm.counter, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= prm.n_items) break;
    const uint4 iv = __ldg(reinterpret_cast<const uint4*>(prm.items + it));
    const uint32_t a = iv.x, cb = iv.z, ce = iv.w;

    // ---- the row's long lists and singles over all parts (lane p < n_parts looks at part p) ----
    uint32_t pk0 = 0, pR = 0, ps0 = 0, pS = 0;
    if (lane < prm.n_parts) {
      const ListPart& q = prm.part[lane];
      pk0 = ldg_keep(q.roff + a, pol_keep);
      pR = ldg_keep(q.roff + a + 1, pol_keep) - pk0;
      ps0 = ldg_keep(q.soff + a, pol_keep);
      pS = ldg_keep(q.soff + a + 1, pol_keep) - ps0;
    }
    uint32_t pRx = pR;  // inclusive scan over parts (n_parts <= 16)
#pragma unroll
    for (int d = 1; d < kMaxParts; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, pRx, d);
      if (lane >= (uint32_t)d) pRx += o;
    }
    const uint32_t R = __shfl_sync(0xffffffffu, pRx, kMaxParts - 1);
    pRx -= pR;  // exclusive: index of part p's first list among the row's lists
    uint32_t S = pS;
#pragma unroll
    for (int d = 16; d; d >>= 1) S += __shfl_xor_sync(0xffffffffu, S, d);
    // rows with <= 32 singles keep them in a register for all of the row's tiles (lane i: the i-th single over the
    // parts); loading them per tile left a global-load latency between the fill and the lists (8 % of the kernel's
    // stall samples at 100k x 200)
    uint32_t my_single = 0xffffffffu;
    const bool s_reg = S <= 32;
    if (S && s_reg) {
      uint32_t pSx = pS;
#pragma unroll
      for (int d = 1; d < kMaxParts; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, pSx, d);
        if (lane >= (uint32_t)d) pSx += o;
      }
      pSx -= pS;
      uint32_t p = 0;
      for (uint32_t q = 1; q < prm.n_parts; ++q)
        if (__shfl_sync(0xffffffffu, pSx, q) <= lane) p = q;
      const uint32_t s0 = __shfl_sync(0xffffffffu, ps0, p), x = __shfl_sync(0xffffffffu, pSx, p);
      if (lane < S) my_single = ldg_keep(prm.part[p].singles + s0 + (lane - x), pol_keep);
    }
    // lists [c0, c0 + Rc) of the row -> s_trow
    auto load_lists = [&](uint32_t c0, uint32_t Rc) {
      for (uint32_t kb = 0; kb < Rc; kb += 32) {
        const uint32_t k = c0 + kb + lane;
        uint32_t p = 0;
        for (uint32_t q = 1; q < prm.n_parts; ++q)  // warp-uniform loop, per-lane k
          if (__shfl_sync(0xffffffffu, pRx, q) <= k) p = q;
        const uint32_t k0 = __shfl_sync(0xffffffffu, pk0, p), x = __shfl_sync(0xffffffffu, pRx, p);
        if (kb + lane < Rc) s_trow[kb + lane] = (p << 28) | (ldg_keep(prm.part[p].rlist + k0 + (k - x), pol_keep) * NB);
      }
      __syncwarp();
    };
    if (R) load_lists(0, min(R, (uint32_t)kCap));
    // rows of <= 64 lists keep the bucket bounds of the NEXT block in registers (lists lane and lane + 32)
    const bool pf_ok = R > 0 && R <= 64 && R <= (uint32_t)kCap;
    uint32_t pf_s[2] = {0, 0}, pf_e[2] = {0, 0};
    auto prefetch_bounds = [&](uint32_t jn) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t kk = c * 32 + lane;
        pf_s[c] = pf_e[c] = 0;
        if (kk < R) {
          const uint32_t tr = s_trow[kk];
          const uint32_t* __restrict__ bo = prm.part[tr >> 28].boff + (tr & 0x0fffffffu) + jn;
          pf_s[c] = ldg_keep(bo, pol_keep);
          pf_e[c] = ldg_keep(bo + 1, pol_keep);
        }
      }
    };

    const uint64_t rs = row_start(a, T);
    const uint32_t na = prm.uniform_nb ? 0u : prm.nb[a];
    const uint32_t j_first = cb >> kLog2B, j_last = (ce - 1) >> kLog2B;
    if (pf_ok) prefetch_bounds(j_first);
    for (uint32_t j = j_first; j <= j_last; ++j) {
      const uint32_t c_lo = max(cb, j << kLog2B), c_hi = min(ce, (j + 1) << kLog2B), cnt = c_hi - c_lo;
      const uint64_t p_lo = rs + (c_lo - a - 1), P0 = p_lo & ~7ull;
      const uint32_t o = (uint32_t)(p_lo - P0);   // tile entry of column c_lo
      const int32_t base = (int32_t)o - (int32_t)c_lo;
      const uint32_t groups = (o + cnt + 7) >> 3;  // 16-byte groups of the buffer in use

      // ---- long lists, part 1: the 16-byte groups of the row's buckets in block j, dealt evenly to the lanes.
      // Bucket bounds come from registers loaded a tile ahead (rows of <= 64 lists); the first groups are requested
      // before the tile is filled, so their latency hides behind the fill.
      const uint32_t tb = smem_addr(tile) + 2u * (uint32_t)(base + (int32_t)(j << kLog2B));  // entry of the block's column 0 (mod 2^32)
      const uint32_t safe = smem_addr(tile) + 4u * lane;  // where an idle lane adds 0: one word per lane, no bank conflict
      const uint32_t lo16 = c_lo - (j << kLog2B), span16 = cnt;  // columns of the tile relative to the block
      uint32_t k = 0, u_next = 0, u1 = 0;  // my run of groups [u_next, u1) of the current pass; k = list of group u_next
      uint4 v[kQ];
      // bucket bounds of lists [0, Rc) in block j -> s_start, s_upre; my run of groups
      auto setup_pass = [&](uint32_t Rc, bool use_pf) {
        uint32_t run = 0;
#pragma unroll
        for (int c = 0; c < (kCap + 31) / 32; ++c) {
          const uint32_t kk = c * 32 + lane;
          if ((uint32_t)c * 32 >= Rc) break;
          uint32_t sb = 0, eb = 0;
          if (use_pf && c < 2) {
            sb = pf_s[c];
            eb = pf_e[c];
          } else if (kk < Rc) {
            const uint32_t tr = s_trow[kk];
            const uint32_t* __restrict__ bo = prm.part[tr >> 28].boff + (tr & 0x0fffffffu) + j;
            sb = ldg_keep(bo, pol_keep);
            eb = ldg_keep(bo + 1, pol_keep);
          }
          const uint32_t u = (eb - sb) >> 3;
          uint32_t inc = u;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += x;
          }
          if (kk < Rc) {
            s_start[kk] = sb;
            s_upre[kk] = run + inc - u;
          }
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_upre[Rc] = run;
        __syncwarp();
        const uint32_t q = (run + 31) >> 5;  // groups per lane
        u_next = min(run, lane * q);
        u1 = min(run, u_next + q);
        k = 0;
        if (u_next < u1) {  // the list holding my first group: last k with upre[k] <= u_next
          uint32_t lo = 0, hi = Rc;
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_upre[mid] <= u_next) lo = mid; else hi = mid;
          }
          k = lo;
        }
      };
      auto load_batch = [&]() {  // requests my next kQ groups
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          v[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          const uint32_t u = u_next + i;
#ifdef PRF_CALIB
          if (prm.calib == 3 && u < u1) {
            while (u >= s_upre[k + 1]) ++k;
            continue;
          }
#endif
          if (u < u1) {
            while (u >= s_upre[k + 1]) ++k;  // (empty buckets repeat the same prefix value)
            const ListPart& lp = prm.part[s_trow[k] >> 28];
            v[i] = ldg_keep(reinterpret_cast<const uint4*>(lp.lmem + s_start[k] + 8u * (u - s_upre[k])), pol_keep);
          }
        }
      };
      auto apply_batch = [&]() {
#pragma unroll
        for (int i = 0; i < kQ; ++i) {
          if (i > 0 && !__any_sync(0xffffffffu, u_next + i < u1)) break;  // warp-uniform
          const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t f0 = w[e] & 0xffffu, f1 = w[e] >> 16;
            tile_red2(tb, safe, f0, f0 - lo16 < span16);
            tile_red2(tb, safe, f1, f1 - lo16 < span16);
          }
        }
        u_next = min(u1, u_next + kQ);
      };
      const uint32_t R0 = min(R, (uint32_t)kCap);
#ifdef PRF_CALIB
      if (prm.calib == 1) u1 = 0; else
#endif
      if (R0) {
        setup_pass(R0, pf_ok);
#ifdef PRF_CALIB
        if (prm.calib == 5) u1 = u_next = 0;
#endif
        load_batch();
      }

      // the previous tile's bulk store has finished reading the buffer
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();

      // ---- fill: d = |B_a| + |B_b| ----
      uint4* t4 = reinterpret_cast<uint4*>(tile);
      if (prm.uniform_nb) {
        if (groups == kTile / 8 || groups == kTile / 8 + 1) {
#pragma unroll
          for (int i = 0; i < kTile / 256; ++i) t4[lane + 32 * i] = q4;
          if (lane == 0) t4[kTile / 8] = q4;
        } else {
          for (uint32_t g = lane; g < groups; g += 32) t4[g] = q4;
        }
      } else {
        const uint32_t na2 = na | (na << 16);
        if (prm.nb_shift != nullptr && c_lo >= o) {
          // the run nb[c_lo - o ..) lands on tile entry 0: pick the copy of nb whose shift makes the source
          // 16-byte aligned and let the bulk-copy engine bring it in, then add |B_a| two entries at a time.
          // (Entries of the first/last group outside [o, o + cnt) get values that are never stored.)
          const uint32_t d0 = c_lo - o, k = d0 & 7u;
          const uint16_t* src = prm.nb_shift + (size_t)k * prm.nb_shift_stride + (d0 - k);
          if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_fill[warp])), "r"(groups * 16u)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(tile)),
                         "l"(src), "r"(groups * 16u), "r"(smem_addr(&s_fill[warp]))
                         : "memory");
          }
          mbar_wait(&s_fill[warp], fill_phase);
          fill_phase ^= 1u;
#pragma unroll 4
          for (uint32_t g = lane; g < groups; g += 32) {
            uint4 v = t4[g];
            v.x = __vadd2(v.x, na2);
            v.y = __vadd2(v.y, na2);
            v.z = __vadd2(v.z, na2);
            v.w = __vadd2(v.w, na2);
            t4[g] = v;
          }
        } else {
          // the first columns of the matrix (c_lo < 8): entry by entry
          for (uint32_t i = lane; i < cnt; i += 32) tile[o + i] = (uint16_t)(na + prm.nb[c_lo + i]);
        }
      }
      __syncwarp();

      // ---- singles ----
      if (S && s_reg) {
        if (my_single - c_lo < cnt) {  // (0xffffffff - c_lo >= cnt: lanes without a single stay out)
          const uint32_t idx = (uint32_t)(base + (int32_t)my_single);
          atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
        }
      } else if (S) {
        for (uint32_t p = 0; p < prm.n_parts; ++p) {
          const uint32_t s0 = __shfl_sync(0xffffffffu, ps0, p), sn = __shfl_sync(0xffffffffu, pS, p);
          const uint32_t* __restrict__ sg = prm.part[p].singles + s0;
          for (uint32_t sb = 0; sb < sn; sb += 32) {
            const uint32_t i = sb + lane;
            const uint32_t b = i < sn ? ldg_keep(sg + i, pol_keep) : 0xffffffffu;
            if (b >= c_lo && b < c_hi) {
              const uint32_t idx = (uint32_t)(base + (int32_t)b);
              atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
            }
          }
        }
      }

      // ---- long lists, part 2 ----
      if (R0) {
        while (__any_sync(0xffffffffu, u_next < u1)) {
          apply_batch();
          if (__any_sync(0xffffffffu, u_next < u1)) load_batch();
        }
        __syncwarp();
      }
      for (uint32_t c0 = kCap; c0 < R; c0 += kCap) {  // rows with more lists than one pass holds
        const uint32_t Rc = min(R - c0, (uint32_t)kCap);
        load_lists(c0, Rc);
        setup_pass(Rc, false);
        while (__any_sync(0xffffffffu, u_next < u1)) {
          load_batch();
          apply_batch();
        }
        __syncwarp();
      }
      if (R > (uint32_t)kCap) load_lists(0, kCap);  // back to the first pass's lists for the next tile
      // bucket bounds of the next block, for rows of <= 64 lists (requesting them earlier, right after this block's
      // were consumed, measured slower: 2.68 vs 2.58 ms)
      if (pf_ok && j < j_last) prefetch_bounds(j + 1);
      __syncwarp();

      // ---- threshold mask: bit (p - p_begin) set iff d <= editdist ----
      if (prm.editdist >= 0) {
        // lane handles one 16-byte group (8 entries -> one mask byte) per step; four neighbouring lanes' bytes
        // make a mask word, stored by the lane on the word boundary.  Words wholly inside this tile's entries
        // are plain stores; words shared with a neighbouring tile are OR-ed into the pre-zeroed mask, so every
        // word is written either by one tile's stores or by atomics only.
        const uint32_t thr = (uint32_t)prm.editdist, thr2 = thr | (thr << 16);
        const uint64_t gb0 = (P0 - prm.p_begin) >> 3;   // mask byte of group 0
        const uint32_t ph = (uint32_t)gb0 & 3u;
        const uint32_t e_lo = o, e_hi = o + cnt;        // valid tile entries
        const uint32_t n_it = (groups + ph + 31) >> 5;
        for (uint32_t itn = 0; itn < n_it; ++itn) {
          const int32_t g = (int32_t)(itn * 32 + lane) - (int32_t)ph;
          uint32_t bits = 0;
          if (g >= 0 && (uint32_t)g < groups) {
            const uint4 v = t4[g];
            const uint32_t t = __vsetleu2(v.x, thr2) | (__vsetleu2(v.y, thr2) << 2) | (__vsetleu2(v.z, thr2) << 4) |
                               (__vsetleu2(v.w, thr2) << 6);
            bits = (t & 0x55u) | ((t >> 15) & 0xaau);
            const uint32_t e0 = 8u * (uint32_t)g;
            if (e0 < e_lo) bits &= 0xffu << (e_lo - e0);
            if (e0 + 8 > e_hi) bits &= e_hi > e0 ? (1u << (e_hi - e0)) - 1u : 0u;
          }
          uint32_t w = bits | (__shfl_down_sync(0xffffffffu, bits, 1) << 8);
          w |= __shfl_down_sync(0xffffffffu, w, 2) << 16;
          if ((lane & 3u) == 0) {
            const int32_t we0 = 8 * g;  // first tile entry of the word (may be negative)
            const uint64_t word = (gb0 + (int64_t)g) >> 2;
            if (we0 >= (int32_t)e_lo && we0 + 32 <= (int32_t)e_hi) prm.mask[word] = w;
            else if (w) atomicOr(&prm.mask[word], w);
          }
        }
      }

      // ---- store ----
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my shared writes -> visible to the bulk copy
      __syncwarp();
      {
        const uint32_t o_lo = o, o_hi = o + cnt;
        const uint32_t ia = (o_lo + 7) & ~7u, ib = o_hi & ~7u;
        uint16_t* gp = prm.out + (P0 - prm.p_begin);
        if (lane < 8) {
          const uint32_t i = o_lo + lane;
          if (i < min(ia, o_hi)) gp[i] = tile[i];
        } else if (lane < 16) {
          const uint32_t i = ib + (lane - 8);
          if (ib >= ia && i < o_hi) gp[i] = tile[i];
        }
        if (lane == 0) {
          if (ib > ia)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gp + ia),
                         "r"(smem_addr(tile + ia)), "r"((ib - ia) * 2u), "l"(pol_stream)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();  // lanes 0..15 have read their head/tail entries before the buffer is refilled
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// The cut of [p_begin, p_end) into whole-row work items (longest first), cached in the context.
static int prepare_row_plan(Ctx* ctx, uint64_t p_begin, uint64_t p_end) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  RowPlan& plan = ctx->plan;
  if (!plan.valid || plan.T != H.T || plan.p_begin != p_begin || plan.p_end != p_end || plan.tile != 0) {
    plan.valid = false;
    std::vector<RowItem> rows, groups;
    build_row_plan(H.T, p_begin, p_end, 0, rows, groups);  // tile 0: whole rows only, longest first
    PRF_CK(ctx, plan.items.alloc(rows.size(), s));
    PRF_CK(ctx, plan.counter.alloc(1, s));
    if (!rows.empty())
      PRF_CK(ctx, cudaMemcpyAsync(plan.items.p, rows.data(), rows.size() * sizeof(RowItem), cudaMemcpyHostToDevice, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));  // `rows` is pageable host memory about to go out of scope
    plan.n_items = (uint32_t)rows.size();
    plan.T = H.T;
    plan.p_begin = p_begin;
    plan.p_end = p_end;
    plan.tile = 0;
    plan.valid = true;
  }
  return PRF_OK;
}

static RowsParams make_rows_params(Ctx* ctx, uint64_t p_begin, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  const RowPlan& plan = ctx->plan;
  RowsParams prm{};
  prm.T = H.T;
  prm.p_begin = p_begin;
  prm.nb = H.nb.p;
  prm.nb_shift = H.nb_shift.p;
  prm.nb_shift_stride = H.nb_shift_stride;
  prm.uniform_nb = H.uniform_nb ? 1 : 0;
  prm.nb_uniform = H.nb_uniform;
  prm.NB = H.NB;
  prm.n_parts = H.n_parts;
  for (uint32_t p = 0; p < H.n_parts; ++p) prm.part[p] = H.part[p];
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  prm.items = plan.items.p;
  prm.n_items = plan.n_items;
  prm.counter = plan.counter.p;
#ifdef PRF_CALIB
  prm.calib = ctx->calib;
#endif
  return prm;
}

template <int kLog2B, int kWarps, int kCap, int kQ>
static int launch_rows_t(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  constexpr int kThreads = 32 * kWarps;
  constexpr size_t smem = (size_t)kWarps * (((1 << kLog2B) + 8) * 2 + kCap * 12 + 16);
  static_assert(smem + 256 <= 232448, "rows kernel: shared memory per CTA");
  if (H.block_log2 != (uint32_t)kLog2B) return ctx->fail(PRF_E_STATE, "list structure was built for another column block size");
  auto kern = rows_kernel<kLog2B, kWarps, kCap, kQ>;
  PRF_CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int rc = prepare_row_plan(ctx, p_begin, p_end);
  if (rc) return rc;
  RowPlan& plan = ctx->plan;
  if (plan.n_items == 0) return PRF_OK;
  const uint32_t grid = std::min<uint32_t>((plan.n_items + kWarps - 1) / kWarps, (uint32_t)ctx->sm_count);
  RowsParams prm = make_rows_params(ctx, p_begin, editdist, d_out, d_mask);
#ifdef PRF_CALIB
  {
    const int no_red = ctx->calib == 2 ? 1 : ctx->calib == 4 ? 2 : 0;
    PRF_CK(ctx, cudaMemcpyToSymbolAsync(g_calib_no_red, &no_red, sizeof(int), 0, cudaMemcpyHostToDevice, s));
  }
#endif
  PRF_CK(ctx, cudaMemsetAsync(plan.counter.p, 0, sizeof(uint32_t), s));
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  kern<<<grid, kThreads, smem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
