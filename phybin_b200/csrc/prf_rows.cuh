// prf_rows.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), sparse variant (round 2).
//
// The reference bumps d(i,j) once for every bipartition that exactly one of the two trees has
// (haveIt x dontHave, :327-338): d(a,b) = |B_a| + |B_b| - 2 |B_a n B_b|.  Random topologies share
// ~0.13 bipartitions per pair, so almost every entry is just |B_a| + |B_b| and the work is writing
// the matrix: 2 bytes per pair, HBM-write bound.
//
// ONE WARP OWNS A ROW.  There is no block-wide or inter-warp synchronisation anywhere in the kernel:
// a warp takes the next row a (dynamic counter, longest rows first) and produces it tile by tile
// (kTile consecutive columns) in its private shared-memory buffer:
//   fill     d = |B_a| + |B_b| for the tile's columns (16-byte shared stores; rows whose |B_t| differ:
//            one bulk async copy of |B|[c_lo..] from one of 8 pre-shifted copies + a SIMD add of |B_a|)
//   singles  bipartitions held by only a few trees (lists of <= kShortMax members) were expanded by the
//            build into an explicit per-row list of columns; each is one shared-memory atomic
//   walk     every long inverted list containing a keeps a cursor into "members after a" that only moves
//            forward from tile to tile (a per-warp array in shared memory; one copy of the loop body, so the
//            kernel stays inside the instruction cache).  Per list
//            and tile the warp loads a 64-member window (two coalesced 128-byte loads, issued a batch of
//            lists ahead of their use), and subtracts 2 at the columns inside the tile with plain 16-bit
//            shared loads/stores: the members of one list are distinct columns and the warp is the
//            buffer's only user, so no atomics are needed (measured: 4.8 updates/clk/SM for this pattern
//            at 7 warps against 1.8 with atomics, profiles/r02_ubench_atoms.log)
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

constexpr int kMaxParts = 16;

struct RowsPart {
  const uint32_t* roff;     // [T+1] per-tree offsets into ranges
  const uint2* ranges;      // (begin, end) into this part's members: the trees b > a of one long list;
                            // a tree's HEAVY lists (many members per tile: warp-per-list walk) come first
  const uint32_t* nheavy;   // [T] how many of the tree's ranges are heavy
  const uint32_t* soff;     // [T+1] per-tree offsets into singles
  const uint32_t* singles;  // columns b > a sharing a short-list bipartition with a (with multiplicity)
  uint32_t mbase;           // position of this part's members in `members`
};

struct RowsParams {
  uint32_t T;
  uint64_t p_begin;
  const uint16_t* nb;
  const uint16_t* nb_shift;   // 8 shifted copies of nb (non-uniform |B_t|), see HashRFState
  uint32_t nb_shift_stride;
  int uniform_nb;
  uint32_t nb_uniform;
  const uint32_t* members;    // all parts' inverted lists (one buffer)
  uint32_t n_parts;
  RowsPart part[kMaxParts];
  int32_t editdist;
  uint16_t* out;
  uint32_t* mask;
  const RowItem* items;
  uint32_t n_items;
  uint32_t* counter;          // next item (zeroed before the launch)
#ifdef PRF_CALIB
  int calib;                  // calibration builds only: 1 = fill + store only, 2 = member loads but no tile updates, 3 = no prefetch
#endif
  uint2* spill;               // per-warp cursor area for rows with more long lists than the register slots hold
  uint32_t spill_stride;
};

__device__ __forceinline__ void tile_sub2(uint16_t* tile, int32_t base, uint32_t b0, uint32_t b1, bool in0, bool in1) {
  uint16_t v0 = 0, v1 = 0;
  const uint32_t i0 = (uint32_t)(base + (int32_t)b0), i1 = (uint32_t)(base + (int32_t)b1);
  if (in0) v0 = tile[i0];
  if (in1) v1 = tile[i1];
  if (in0) tile[i0] = (uint16_t)(v0 - 2);
  if (in1) tile[i1] = (uint16_t)(v1 - 2);
  __syncwarp();  // the next list's loads may hit the same entries from other lanes
}

// Walks the row's long lists through one tile.  curs[k] = (cursor, end) of list k (shared memory, or the
// global spill area for rows with more lists than fit); consumes every member < c_hi.  The member windows
// of a batch of kB lists are loaded one batch ahead of their use.
template <int kB>
__device__ __forceinline__ void walk_lists(const uint32_t* __restrict__ members, uint2* curs, const uint32_t n,
                                           const uint32_t c_hi, const int32_t base, uint16_t* tile, const uint32_t lane) {
  struct Batch {
    uint32_t c[kB], m0[kB], m1[kB];
  };
  auto load = [&](uint32_t kb, Batch& b) {
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      const uint2 r = kb + u < n ? curs[kb + u] : make_uint2(0u, 0u);  // same address in every lane: broadcast
      const uint32_t i0 = r.x + lane;
      b.c[u] = r.x;
      b.m0[u] = i0 < r.y ? __ldg(members + i0) : 0xffffffffu;
      b.m1[u] = i0 + 32 < r.y ? __ldg(members + i0 + 32) : 0xffffffffu;
    }
  };
  auto process = [&](uint32_t kb, const Batch& b) {
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      const bool in0 = b.m0[u] < c_hi;
      const uint32_t bal0 = __ballot_sync(0xffffffffu, in0);
      if (bal0) {  // warp-uniform
        const bool in1 = b.m1[u] < c_hi;
        const uint32_t bal1 = __ballot_sync(0xffffffffu, in1);
        tile_sub2(tile, base, b.m0[u], b.m1[u], in0, in1);
        uint32_t c = b.c[u] + __popc(bal0) + __popc(bal1);
        if (bal1 == 0xffffffffu) {  // more than two windows of this list fall into the tile
          const uint32_t e = curs[kb + u].y;
          for (;;) {
            const uint32_t i0 = c + lane;
            const uint32_t x0 = i0 < e ? __ldg(members + i0) : 0xffffffffu;
            const uint32_t x1 = i0 + 32 < e ? __ldg(members + i0 + 32) : 0xffffffffu;
            const bool y0 = x0 < c_hi, y1 = x1 < c_hi;
            const uint32_t k = __popc(__ballot_sync(0xffffffffu, y0)) + __popc(__ballot_sync(0xffffffffu, y1));
            tile_sub2(tile, base, x0, x1, y0, y1);
            c += k;
            if (k != 64) break;
          }
        }
        if (lane == 0) curs[kb + u].x = c;
      }
    }
  };
  Batch A, B;
  load(0, A);
#pragma unroll 1
  for (uint32_t kb = 0; kb < n; kb += 2 * kB) {
    if (kb + kB < n) load(kb + kB, B);
    process(kb, A);
    if (kb + 2 * kB < n) load(kb + 2 * kB, A);
    if (kb + kB < n) process(kb + kB, B);
  }
}

// ---- LIGHT lists (few members per tile): ONE LANE PER LIST -------------------------------------------------
// A lane loads the next 4 * kU members of its own list with 16-byte loads from a 4-aligned index (all in flight
// at once) and applies those inside the tile with shared-memory reductions (two lanes' lists may hit the same
// column): ~8 instructions per member slot against ~60 per (list, tile) of the warp-per-list walk above.
// Lists are stored padded to a multiple of four members (pad = 0xffffffff, never below c_hi) and start at a
// multiple of four, so a 16-byte group never straddles two lists and the members before the cursor inside the
// first group belong to the same list: they are below c_lo and fail the range test.
struct LightWin {  // the window one lane holds of one list
  uint32_t x;      // cursor the window was loaded for (0 with y == 0: no list)
  uint32_t y;      // end of the list
};

template <int kU>
__device__ __forceinline__ void light_load(const uint32_t* __restrict__ members, const uint2 r, LightWin& w, uint4 (&buf)[kU]) {
  w.x = r.x;
  w.y = r.y;
  const uint32_t idx = r.x & ~3u;
#pragma unroll
  for (int j = 0; j < kU; ++j) {
    buf[j] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (idx + 4 * j < r.y) buf[j] = __ldg(reinterpret_cast<const uint4*>(members + idx + 4 * j));
  }
}

#ifdef PRF_CALIB
__device__ int g_calib_no_red;
#endif
__device__ __forceinline__ void tile_red2(uint32_t tb, uint32_t m, bool on) {
#ifdef PRF_CALIB
  if (g_calib_no_red) on = false;
#endif
  const uint32_t A = tb + 2u * m;  // byte address of column m's entry
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.u32 p, %2, 0;\n\t"
      "@p red.shared.add.u32 [%0], %1;\n\t"
      "}" ::"r"(A & ~3u),
      "r"(0xfffffffeu << ((A & 2u) << 3)), "r"((uint32_t)on)
      : "memory");
}

// Applies the window; returns the new cursor.  `tb` = byte address of column 0's entry (mod 2^32).  If the whole
// window was consumed and the list goes on (rare: more than 4 * kU - 3 members in one tile), continues with
// dependent loads.
template <int kU>
__device__ __forceinline__ uint32_t light_apply(const uint32_t* __restrict__ members, const LightWin w, const uint4 (&buf)[kU],
                                                const uint32_t c_lo, const uint32_t c_hi, const uint32_t tb) {
  const uint32_t span = c_hi - c_lo;
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kU; ++j) {
    const uint32_t mm[4] = {buf[j].x, buf[j].y, buf[j].z, buf[j].w};
    // warp-uniform early out: no lane's list has a member below c_hi from here on (lists ascend)
    if (j > 0 && !__any_sync(0xffffffffu, mm[0] < c_hi)) break;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool on = j == 0 ? mm[e] - c_lo < span : mm[e] < c_hi;
      cnt += on ? 1u : 0u;
      tile_red2(tb, mm[e], on);
    }
  }
  uint32_t cur = w.x + cnt;
  uint32_t idx = (w.x & ~3u) + 4 * kU;
  bool more = cur == idx && idx < w.y;  // consumed everything loaded and the list goes on
  while (__any_sync(0xffffffffu, more)) {
    uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    if (more) v = __ldg(reinterpret_cast<const uint4*>(members + idx));
    const uint32_t mm[4] = {v.x, v.y, v.z, v.w};
    uint32_t c = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool on = mm[e] < c_hi;
      c += on ? 1u : 0u;
      tile_red2(tb, mm[e], on);
    }
    cur += c;
    idx += 4;
    more = more && c == 4 && idx < w.y;
  }
  return cur;
}

// Generic light walk over lists [0, n) of `curs` (no prefetch): groups of 32 lists, one list per lane.
template <int kU>
__device__ __forceinline__ void walk_light(const uint32_t* __restrict__ members, uint2* curs, const uint32_t n,
                                           const uint32_t c_lo, const uint32_t c_hi, const uint32_t tb, const uint32_t lane) {
#pragma unroll 1
  for (uint32_t g0 = 0; g0 < n; g0 += 32) {
    const uint32_t k = g0 + lane;
    const uint2 r = k < n ? curs[k] : make_uint2(0u, 0u);
    LightWin w;
    uint4 buf[kU];
    light_load<kU>(members, r, w, buf);
    const uint32_t cur = light_apply<kU>(members, w, buf, c_lo, c_hi, tb);
    if (cur != r.x) curs[k].x = cur;
  }
}

// first index in [lo, hi) whose member is >= c
__device__ __forceinline__ uint32_t members_lower_bound(const uint32_t* __restrict__ members, uint32_t lo, uint32_t hi, uint32_t c) {
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(members + mid) < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int kTile, int kWarps, int kCurCap, int kB, int kU>
__global__ void __launch_bounds__(32 * kWarps, 1) rows_kernel(const RowsParams prm) {
  constexpr int kStride = kTile + 8;  // entries per buffer: a tile keeps the 16-byte phase of its global address
  constexpr int kWarpBytes = kStride * 2 + kCurCap * 8;  // per warp: the tile buffer, then the cursors of the row's lists
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_fill[kWarps];  // completion of a warp's bulk load of |B| (non-uniform rows)

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gwarp = blockIdx.x * kWarps + warp;
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw + (size_t)warp * kWarpBytes);
  uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
  uint2* s_cur = reinterpret_cast<uint2*>(smem_raw + (size_t)warp * kWarpBytes + kStride * 2);
  const uint32_t T = prm.T;
  const uint32_t* __restrict__ members = prm.members;

  if (lane == 0) {
    mbar_init(&s_fill[warp], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t fill_phase = 0;

  const uint32_t vv = (2u * prm.nb_uniform) | ((2u * prm.nb_uniform) << 16);
  const uint4 q4 = make_uint4(vv, vv, vv, vv);

  for (;;) {
    uint32_t it = 0;
    if (lane == 0) it = atomicAdd(prm.counter, 1u);
    it = __shfl_sync(0xffffffffu, it, 0);
    if (it >= prm.n_items) break;
    const uint4 iv = __ldg(reinterpret_cast<const uint4*>(prm.items + it));
    const uint32_t a = iv.x, cb = iv.z, ce = iv.w;

    // ---- the row's long lists: cursor l of group g lives in lane l ----
    // Cursor order of the row: the heavy lists of every part, then the light lists of every part.
    uint32_t R = 0, H = 0, S = 0;  // ranges / heavy ranges / singles of the row over all parts
    uint32_t pk0 = 0, pH = 0, pL = 0;  // lane p < n_parts: first range, heavy and light range counts of part p
    uint32_t ps0 = 0, pS = 0;
    if (lane < prm.n_parts) {
      const RowsPart& q = prm.part[lane];
      pk0 = __ldg(q.roff + a);
      pH = __ldg(q.nheavy + a);
      pL = __ldg(q.roff + a + 1) - pk0 - pH;
      ps0 = __ldg(q.soff + a);
      pS = __ldg(q.soff + a + 1) - ps0;
    }
    uint32_t pHx = pH, pLx = pL;  // inclusive scans over parts (n_parts <= 16)
#pragma unroll
    for (int d = 1; d < kMaxParts; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, pHx, d), o2 = __shfl_up_sync(0xffffffffu, pLx, d);
      if (lane >= (uint32_t)d) pHx += o, pLx += o2;
    }
    H = __shfl_sync(0xffffffffu, pHx, kMaxParts - 1);
    R = H + __shfl_sync(0xffffffffu, pLx, kMaxParts - 1);
    pHx -= pH;        // exclusive: cursor index of part p's first heavy list
    pLx += H - pL;    // exclusive + H: cursor index of part p's first light list
    {
      uint32_t s = pS;
#pragma unroll
      for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      S = s;
    }
    // cursor k of the row (0 <= k < R): heavy list (k < H) or light list of the last part p whose first index is <= k
    auto fetch_range = [&](uint32_t k) -> uint2 {
      const bool heavy = k < H;
      uint32_t p = 0;
      for (uint32_t q = 1; q < prm.n_parts; ++q)  // warp-uniform loop, per-lane k
        if ((heavy ? __shfl_sync(0xffffffffu, pHx, q) : __shfl_sync(0xffffffffu, pLx, q)) <= k) p = q;
      const uint32_t k0 = __shfl_sync(0xffffffffu, pk0, p), hx = __shfl_sync(0xffffffffu, pHx, p),
                     lx = __shfl_sync(0xffffffffu, pLx, p), ph = __shfl_sync(0xffffffffu, pH, p);
      uint2 r = make_uint2(0u, 0u);
      if (k < R) {
        r = __ldg(prm.part[p].ranges + k0 + (heavy ? k - hx : ph + (k - lx)));
        const uint32_t mb = prm.part[p].mbase;
        r.x += mb;
        r.y += mb;
        if (cb > a + 1) r.x = members_lower_bound(members, r.x, r.y, cb);  // shard starts inside the row
      }
      return r;
    };
    uint2* spill = prm.spill + (size_t)gwarp * prm.spill_stride;
    for (uint32_t kb = 0; kb < R; kb += 32) {
      const uint32_t k = kb + lane;
      const uint2 r = fetch_range(k);
      if (k < (uint32_t)kCurCap) s_cur[k] = r;
      else if (k < R) spill[k - kCurCap] = r;
    }
    __syncwarp();

    const uint32_t len = ce - cb, nseg = (len + kTile - 1) / kTile;
    const uint64_t rs = row_start(a, T);
    const uint32_t Hs = min(H, (uint32_t)kCurCap), Rs = min(R, (uint32_t)kCurCap);
    const uint32_t na = prm.uniform_nb ? 0u : prm.nb[a];
    for (uint32_t seg = 0; seg < nseg; ++seg) {
      const uint32_t c_lo = cb + seg * kTile, c_hi = min(ce, c_lo + (uint32_t)kTile), cnt = c_hi - c_lo;
      const uint64_t p_lo = rs + (c_lo - a - 1), P0 = p_lo & ~7ull;
      const uint32_t o = (uint32_t)(p_lo - P0);   // tile entry of column c_lo
      const int32_t base = (int32_t)o - (int32_t)c_lo;
      const uint32_t groups = (o + cnt + 7) >> 3;  // 16-byte groups of the buffer in use

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
      if (S) {
        for (uint32_t p = 0; p < prm.n_parts; ++p) {
          const uint32_t s0 = __shfl_sync(0xffffffffu, ps0, p), sn = __shfl_sync(0xffffffffu, pS, p);
          const uint32_t* __restrict__ sg = prm.part[p].singles + s0;
          for (uint32_t sb = 0; sb < sn; sb += 32) {
            const uint32_t i = sb + lane;
            const uint32_t b = i < sn ? __ldg(sg + i) : 0xffffffffu;
            if (b >= c_lo && b < c_hi) {
              const uint32_t idx = (uint32_t)(base + (int32_t)b);
              atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
            }
          }
        }
        __syncwarp();
      }

      // ---- walk the long lists ----
      // cursors [0, H) heavy, [H, R) light; cursors >= kCurCap live in the global spill area
      const uint32_t tb = smem_addr(tile) + 2u * (uint32_t)base;  // byte address of column 0's entry (mod 2^32)
#ifdef PRF_CALIB
      if (prm.calib != 1) {
#endif
      if (Hs) walk_lists<kB>(members, s_cur, Hs, c_hi, base, tile, lane);
      if (H > Hs) walk_lists<kB>(members, spill, H - kCurCap, c_hi, base, tile, lane);
      if (Rs > Hs) walk_light<kU>(members, s_cur + Hs, Rs - Hs, c_lo, c_hi, tb, lane);
      if (R > Rs) walk_light<kU>(members, spill + (max(H, (uint32_t)kCurCap) - kCurCap), R - max(H, (uint32_t)kCurCap), c_lo, c_hi, tb, lane);
      __syncwarp();
#ifdef PRF_CALIB
      }
#endif

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
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gp + ia),
                         "r"(smem_addr(tile + ia)), "r"((ib - ia) * 2u)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();  // lanes 0..15 have read their head/tail entries before the buffer is refilled
      }
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int kTile, int kWarps, int kCurCap, int kB, int kU>
static int launch_rows_t(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  constexpr int kThreads = 32 * kWarps;
  constexpr size_t smem = (size_t)kWarps * ((kTile + 8) * 2 + kCurCap * 8);
  static_assert(smem + 64 <= 232448, "rows kernel: shared memory per CTA");
  static_assert(kTile % 256 == 0 && kCurCap % (2 * kB) == 0, "rows kernel: tile / cursor array shape");
  auto kern = rows_kernel<kTile, kWarps, kCurCap, kB, kU>;
  PRF_CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
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
  if (plan.n_items == 0) return PRF_OK;
  const uint32_t grid = std::min<uint32_t>((plan.n_items + kWarps - 1) / kWarps, (uint32_t)ctx->sm_count);
  RowsParams prm{};
  prm.T = H.T;
  prm.p_begin = p_begin;
  prm.nb = H.nb.p;
  prm.nb_shift = H.nb_shift.p;
  prm.nb_shift_stride = H.nb_shift_stride;
  prm.uniform_nb = H.uniform_nb ? 1 : 0;
  prm.nb_uniform = H.nb_uniform;
  prm.members = H.lmembers.p;
  prm.n_parts = 1;
  prm.part[0].roff = H.roff.p;
  prm.part[0].ranges = H.ranges.p;
  prm.part[0].nheavy = H.nheavy.p;
  prm.part[0].soff = H.soff.p;
  prm.part[0].singles = H.singles.p;
  prm.part[0].mbase = 0;
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  prm.items = plan.items.p;
  prm.n_items = plan.n_items;
  prm.counter = plan.counter.p;
  if (H.max_ranges > (uint32_t)kCurCap) {
    const uint32_t stride = (H.max_ranges - kCurCap + 31) / 32 * 32;
    const size_t need = (size_t)grid * kWarps * stride;
    if (plan.cursor_ws.n < need) PRF_CK(ctx, plan.cursor_ws.alloc(need, s));
    prm.spill = plan.cursor_ws.p;
    prm.spill_stride = stride;
  }
#ifdef PRF_CALIB
  prm.calib = ctx->calib;
  {
    const int no_red = ctx->calib == 2;
    PRF_CK(ctx, cudaMemcpyToSymbolAsync(g_calib_no_red, &no_red, sizeof(int), 0, cudaMemcpyHostToDevice, s));
  }
#endif
  PRF_CK(ctx, cudaMemsetAsync(plan.counter.p, 0, sizeof(uint32_t), s));
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  kern<<<grid, kThreads, smem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

int launch_rows(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  // the tiny-tile build exists so that small test inputs exercise multi-tile rows and the cursor spill area
  if (ctx->debug_small_tiles) return launch_rows_t<256, 2, 8, 2, 1>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  switch (ctx->rows_cfg) {  // tuning instantiations (scripts/rows_probe.py); the default measured best at 100k x 200
    case 1: return launch_rows_t<7680, 14, 152, 4, 8>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 2: return launch_rows_t<6656, 16, 128, 4, 7>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 3: return launch_rows_t<8960, 12, 168, 4, 8>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    default: return launch_rows_t<5376, 20, 104, 4, 6>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  }
}

}  // namespace prf
