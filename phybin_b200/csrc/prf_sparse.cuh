// prf_sparse.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), sparse variant.
//
// The reference bumps d(i,j) once for every bipartition that exactly one of the two trees has
// (haveIt x dontHave, :327-338): d(a,b) = |B_a| + |B_b| - 2 |B_a n B_b|.  Random topologies share
// ~0.13 bipartitions per pair, so almost every entry is just |B_a| + |B_b| and the work is writing
// the matrix: 2 bytes per pair, HBM-write bound.
//
// Work is cut along ROWS of the packed upper triangle.  A work item is either one row a restricted
// to columns [cb, ce) (long rows, and the partial first/last row of a shard), or a group of whole
// short rows that fit one tile.  Persistent CTAs take items round-robin (longest first).  A row is
// produced tile by tile (kTile consecutive columns) through a ring of three shared-memory buffers.
//
// The CTA is warp-specialised and has no block-wide barrier in its main loop:
//   producer warp (warp 0)
//     waits for the ring slot's `empty` mbarrier, fills d = |B_a| + |B_b| for the tile's columns
//     (16-byte shared stores), publishes the row's suffix cursors and a tile descriptor, and arrives
//     on the tile's `full` mbarrier.  Work-item metadata is prefetched three items ahead (item ->
//     list offsets -> ranges, the ranges by cp.async straight into a ring of kRowBufs + 2 cursor
//     arrays), so no global latency sits on its path.  Rows whose |B_t| differ are filled by a bulk
//     copy of |B|[c_lo..] from one of 8 pre-shifted copies plus a SIMD add of |B_a|.
//   emitter warp (warp 1)
//     once every walker has arrived on the tile's `done` mbarrier: optional d <= editdist mask, then
//     the finished tile goes out as ONE bulk asynchronous copy (cp.async.bulk shared -> global, the
//     TMA engine; 16-byte aligned interior; the < 8 entries before and after it are stored by a few
//     lanes).  When the copy has finished reading shared memory the slot's `empty` barrier is signalled.
//   walker warps (the other warps)
//     tree a's shared bipartitions each own a suffix of an inverted list ("trees b > a that also have
//     it", ascending).  Every suffix keeps a cursor that only moves forward from tile to tile, so each
//     member is loaded once per row and there is no search.  A suffix always belongs to the same
//     walker, so cursors need no synchronisation: a walker keeps kSlots suffixes in registers (cursor,
//     end, and a 32-member look-ahead window that is reloaded as soon as the cursor moves, i.e. before
//     the next tile's `full` wait) and subtracts 2 at column b with a shared-memory reduction -- no
//     global atomics on the matrix.  Suffixes beyond the register slots use shared-memory cursors.
//     Walkers run ahead of each other by up to two tiles.
// Every output byte is written exactly once.
#pragma once

#include <algorithm>

#include "prf_common.cuh"
#include "prf_async.cuh"

namespace prf {

constexpr int kGroupRows = 256;   // most rows in a group item
constexpr int kWalkUnroll = 4;

struct RowParams {
  uint32_t T;
  uint64_t p_begin;
  const uint16_t* nb;
  const uint16_t* nb_shift;   // 8 shifted copies of nb (non-uniform |B_t|), see HashRFState
  uint32_t nb_shift_stride;
  int uniform_nb;
  uint32_t nb_uniform;
  const uint32_t* roff;
  const uint2* ranges;
  const uint32_t* members;
  int32_t editdist;
  uint16_t* out;
  uint32_t* mask;
  const RowItem* items;
  uint32_t n_items;
  uint2* cursor_ws;           // kRowBufs + 2 spill areas per CTA for rows with more suffixes than a shared cursor array
  uint32_t cursor_ws_stride;  // entries per area
  int debug_nowalk;           // calibration only: skip the walk (wrong results)
  uint32_t debug_sleep;       // ns of back-off between mbarrier polls (0 = spin)
  unsigned long long* timing; // PRF_TIMING builds: per-CTA cycle counters [grid][16]
};

enum : uint32_t { kTileRow = 0, kTileGroup = 1, kTileExit = 2 };

struct TileDesc {
  uint32_t kind;     // kTileRow / kTileGroup / kTileExit
  uint32_t a0;       // the row, or the group's first row
  uint32_t nr;       // rows in the group
  uint32_t k0;       // first range of the item
  uint32_t R;        // ranges of the item
  uint32_t c_hi;     // row tile: one past its last column
  int32_t base;      // row tile: entry index of column b is base + b
  uint32_t cur_sel;  // row tile: 0..kCurRing-1 = shared cursor array, kCurRing/kCurRing+1 = global spill area
  uint32_t flags;    // row tile: bit 0 = first tile of the row item, bit 1 = last
  uint64_t p_lo, p_hi;  // packed positions held by the tile; entry 0 is position p_lo & ~7
};

// What the service warp knows about an upcoming work item.
struct Staged {
  RowItem it;
  uint32_t k0, R;  // first range / number of ranges of the item's rows
  uint32_t arr;    // cursor array: < kCurRing shared, else global spill area (arr - kCurRing)
  uint32_t valid;
};

// One warp, one suffix: consume members < c_hi starting with the already loaded b0/b1 (members
// cur+lane and cur+32+lane, 0xffffffff past the end).  Returns the advanced cursor.
__device__ __forceinline__ uint32_t walk_suffix(const uint32_t* __restrict__ members, uint32_t cur, uint32_t end,
                                                uint32_t b0, uint32_t b1, uint32_t c_hi, int32_t base,
                                                uint32_t* tile32, uint32_t lane) {
  uint32_t b = b0;
  bool second = true;
  for (;;) {
    const bool in = b < c_hi;
    if (in) {
      const uint32_t idx = (uint32_t)(base + (int32_t)b);
      atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, in);
    cur += __popc(bal);
    if (bal != 0xffffffffu) break;
    if (second) {
      b = b1;
      second = false;
    } else {
      const uint32_t m = cur + lane;
      b = m < end ? __ldg(members + m) : 0xffffffffu;
    }
  }
  return cur;
}

template <int kTile, int kWalkers, int kCurCap, int kSlots, int kRowBufs, int kHelper>
__global__ void __launch_bounds__(32 * (kWalkers + 2 + kHelper), (kTile > 16384 ? 1 : 2)) sparse_rows_kernel(const RowParams prm) {
  constexpr int kStride = kTile + 8;  // entries per buffer: a tile keeps the 16-byte phase of its global address
  constexpr int kCurRing = kRowBufs + 2;  // cursor arrays in shared memory (see the safety argument at stage_prefetch)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint16_t* tiles = reinterpret_cast<uint16_t*>(smem_raw);
  uint2* s_cur = reinterpret_cast<uint2*>(smem_raw + (size_t)kRowBufs * kStride * 2);  // kCurRing x kCurCap
  __shared__ TileDesc s_desc[kRowBufs];
  __shared__ Staged s_st[4];
  __shared__ __align__(8) uint64_t s_fill;  // completion of the producer's bulk load of |B| (non-uniform rows)
  __shared__ uint32_t s_end[kSlots > 8 ? kWalkers * kSlots : 1];  // list ends of the walkers' slots (kSlots > 8)
  __shared__ __align__(8) uint64_t s_full[kRowBufs], s_done[kRowBufs], s_empty[kRowBufs];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t T = prm.T;
  const uint32_t* __restrict__ members = prm.members;

  if (tid == 0) {
    for (int b = 0; b < kRowBufs; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&s_done[b], kWalkers);
      mbar_init(&s_empty[b], 1 + kHelper);
    }
    mbar_init(&s_fill, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= 2 + kHelper) {
    // =========================== walkers ===========================
    const uint32_t w = warp - 2 - kHelper;
#ifdef PRF_TIMING
    long long t_wait = 0, t_work = 0, t0 = clock64(), t1;
#define TICK(acc) do { t1 = clock64(); acc += t1 - t0; t0 = t1; } while (0)
#else
#define TICK(acc) do {} while (0)
#endif
    // Register-resident suffixes of the current row: slot u is range w + u * kWalkers.  win[u] is a
    // look-ahead window, members[cx[u] + lane] (0xffffffff past the end), loaded as soon as the cursor
    // moves -- i.e. before the next tile's `full` wait -- so its latency hides behind the barrier.
    // With more than 8 slots the list ends move to shared memory (one broadcast load per window slide)
    // so that a slot costs two registers: rows with many shared bipartitions (500 taxa) stay off the
    // slower shared-memory cursor path.
    constexpr bool kEndInSmem = kSlots > 8;
    uint32_t cx[kSlots], ce[kEndInSmem ? 1 : kSlots], win[kSlots];
    uint32_t* my_end = &s_end[kEndInSmem ? w * kSlots : 0];
#pragma unroll
    for (int u = 0; u < kSlots; ++u) cx[u] = 0u, win[u] = 0xffffffffu;
    if constexpr (!kEndInSmem) {
#pragma unroll
      for (int u = 0; u < kSlots; ++u) ce[u] = 0u;
    }
    auto slot_end = [&](int u) -> uint32_t {
      if constexpr (kEndInSmem) return my_end[u];
      else return ce[u];
    };
    for (uint32_t q = 0;; ++q) {
      const uint32_t buf = q % kRowBufs;
      TICK(t_work);
      mbar_wait(&s_full[buf], (q / kRowBufs) & 1u, prm.debug_sleep);
      TICK(t_wait);
      const TileDesc d = s_desc[buf];
      if (d.kind == kTileExit) break;
      uint32_t* tile32 = reinterpret_cast<uint32_t*>(tiles + buf * kStride);
      if (d.kind == kTileRow) {
        uint2* cur = d.cur_sel < (uint32_t)kCurRing
                         ? s_cur + d.cur_sel * kCurCap
                         : prm.cursor_ws + ((size_t)blockIdx.x * kCurRing + (d.cur_sel - kCurRing)) * prm.cursor_ws_stride;
        const bool first = d.flags & 1u, last = d.flags & 2u;
        if (first) {
#pragma unroll
          for (int u = 0; u < kSlots; ++u) {
            const uint32_t k = w + u * kWalkers;
            const uint2 r = k < d.R ? cur[k] : make_uint2(0u, 0u);
            cx[u] = r.x;
            if constexpr (kEndInSmem) {
              if (lane == 0) my_end[u] = r.y;
            } else {
              ce[u] = r.y;
            }
          }
          if constexpr (kEndInSmem) __syncwarp();
#pragma unroll
          for (int u = 0; u < kSlots; ++u) {
            const uint32_t m = cx[u] + lane;
            win[u] = m < slot_end(u) ? __ldg(members + m) : 0xffffffffu;
          }
        }
        bool again;
        do {
          again = false;
#pragma unroll
          for (int u = 0; u < kSlots; ++u) {
            const bool in = win[u] < d.c_hi;
            const uint32_t bal = __ballot_sync(0xffffffffu, in);
            if (bal) {
              if (in) {
                const uint32_t idx = (uint32_t)(d.base + (int32_t)win[u]);
                atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
              }
              cx[u] += __popc(bal);
              const bool full = bal == 0xffffffffu;
              again |= full;
              if (full || !last) {  // slide the window (more of this tile, or the next tile of the row)
                const uint32_t m = cx[u] + lane;
                win[u] = m < slot_end(u) ? __ldg(members + m) : 0xffffffffu;
              } else {
                win[u] = 0xffffffffu;  // consumed; the row ends here, nothing to look ahead to
              }
            }
          }
        } while (again);
        // suffixes beyond the register slots (rows with more than kSlots * kWalkers shared bipartitions)
        for (uint32_t kb = w + kSlots * kWalkers; kb < d.R; kb += kWalkers * kWalkUnroll) {
          uint2 c[kWalkUnroll];
          uint32_t b0[kWalkUnroll], b1[kWalkUnroll];
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            const uint32_t k = kb + u * kWalkers;
            c[u] = k < d.R ? cur[k] : make_uint2(0u, 0u);
          }
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            const uint32_t m = c[u].x + lane;
            b0[u] = m < c[u].y ? __ldg(members + m) : 0xffffffffu;
            b1[u] = m + 32 < c[u].y ? __ldg(members + m + 32) : 0xffffffffu;
          }
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            if (c[u].x < c[u].y) {
              const uint32_t nc = walk_suffix(members, c[u].x, c[u].y, b0[u], b1[u], d.c_hi, d.base, tile32, lane);
              if (lane == 0 && nc != c[u].x) cur[kb + u * kWalkers].x = nc;
            }
          }
        }
      } else {
        const uint64_t P0 = d.p_lo & ~7ull;
        for (uint32_t kb = w; kb < d.R; kb += kWalkers * kWalkUnroll) {
          uint2 c[kWalkUnroll];
          uint32_t b0[kWalkUnroll], b1[kWalkUnroll];
          int32_t base[kWalkUnroll];
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            c[u] = make_uint2(0u, 0u);
            base[u] = 0;
            const uint32_t kk = kb + u * kWalkers;
            if (kk < d.R) {
              const uint32_t k = d.k0 + kk;
              c[u] = prm.ranges[k];
              uint32_t lo = 0, hi = d.nr - 1;  // largest i with roff[a0 + i] <= k
              while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (__ldg(prm.roff + d.a0 + mid) <= k) lo = mid; else hi = mid - 1;
              }
              const uint32_t a = d.a0 + lo;
              base[u] = (int32_t)(row_start(a, T) - P0) - (int32_t)(a + 1);
            }
          }
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            const uint32_t m = c[u].x + lane;
            b0[u] = m < c[u].y ? __ldg(members + m) : 0xffffffffu;
            b1[u] = m + 32 < c[u].y ? __ldg(members + m + 32) : 0xffffffffu;
          }
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u)
            if (c[u].x < c[u].y) walk_suffix(members, c[u].x, c[u].y, b0[u], b1[u], T, base[u], tile32, lane);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my shared writes -> visible to the bulk copy
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_done[buf]);
    }
#ifdef PRF_TIMING
    if (lane == 0 && prm.timing) {
      prm.timing[(size_t)blockIdx.x * 64 + 2 * warp] = t_wait;
      prm.timing[(size_t)blockIdx.x * 64 + 2 * warp + 1] = t_work;
    }
#endif
    return;
  }

  // =========================== producer (warp 0) and emitter (warp 1) ===========================
#ifdef PRF_TIMING
  long long t_wait = 0, t_work = 0, t_wait2 = 0, t0 = clock64(), t1;
#endif
  const uint32_t vv = (2u * prm.nb_uniform) | ((2u * prm.nb_uniform) << 16);
  const uint4 q4 = make_uint4(vv, vv, vv, vv);

  // d = |B_a| + |B_b| for packed positions [p_lo, p_hi) of row a starting at column c_lo; entry 0 = P0
  uint32_t fill_phase = 0;
  auto fill_row = [&](uint16_t* tile, uint32_t a, uint32_t c_lo, uint64_t p_lo, uint64_t p_hi, uint64_t P0,
                      bool bulk_ok = false) {
    if (prm.uniform_nb) {
      // the whole buffer, unconditionally: kTile/256 stores with immediate offsets
      uint4* t4 = reinterpret_cast<uint4*>(tile) + lane;
#pragma unroll
      for (int i = 0; i < kTile / 256; ++i) t4[32 * i] = q4;
      if (lane == 0) t4[kTile / 8] = q4;
    } else {
      const uint32_t na = prm.nb[a];
      const uint32_t na2 = na | (na << 16);
      const uint32_t cnt = (uint32_t)(p_hi - p_lo), o = (uint32_t)(p_lo - P0);
      uint4* t4 = reinterpret_cast<uint4*>(tile);
      if (bulk_ok && c_lo >= o) {
        // the run nb[c_lo - o ..) lands on tile entry 0: pick the copy of nb whose shift makes the source
        // 16-byte aligned and let the bulk-copy engine bring it in, then add |B_a| two entries at a time.
        // (Entries of the first/last group outside [o, o + cnt) get values that are never stored.)
        const uint32_t d0 = c_lo - o, k = d0 & 7u, groups = (o + cnt + 7) >> 3;
        const uint16_t* src = prm.nb_shift + (size_t)k * prm.nb_shift_stride + (d0 - k);
        if (lane == 0) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_fill)), "r"(groups * 16u)
                       : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_addr(tile)),
                       "l"(src), "r"(groups * 16u), "r"(smem_addr(&s_fill))
                       : "memory");
        }
        mbar_wait(&s_fill, fill_phase);
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
        return;
      }
      // generic path (group tiles, the first rows): one aligned 16-byte group of the tile per lane:
      // 5 aligned 32-bit loads of nb (its element offset is arbitrary), funnel-shifted into place.
      const uint32_t* nb32 = reinterpret_cast<const uint32_t*>(prm.nb);  // nb is allocated with 16 entries of slack
      const uint32_t g_lo = (o + 7) >> 3, g_hi = (o + cnt) >> 3;  // fully covered groups [g_lo, g_hi)
      // four groups per lane and round: 20 independent loads in flight (one warp fills the whole tile, so
      // a load-use round trip per group would make the producer the bottleneck)
      constexpr int kU = 4;
      for (uint32_t gb = g_lo + lane; gb < g_hi; gb += 32 * kU) {
        uint32_t wv[kU][5];
#pragma unroll
        for (int j = 0; j < kU; ++j) {
          const uint32_t g = gb + 32 * j;
          const uint32_t src = c_lo + 8 * g - o;  // nb index of the group's first entry
          const uint32_t* wp = nb32 + (src >> 1);
#pragma unroll
          for (int i = 0; i < 5; ++i) wv[j][i] = g < g_hi ? __ldg(wp + i) : 0u;
        }
#pragma unroll
        for (int j = 0; j < kU; ++j) {
          const uint32_t g = gb + 32 * j;
          if (g < g_hi) {
            const uint32_t sh = ((c_lo + 8 * g - o) & 1u) * 16u;
            uint4 v;
            v.x = __vadd2(__funnelshift_r(wv[j][0], wv[j][1], sh), na2);
            v.y = __vadd2(__funnelshift_r(wv[j][1], wv[j][2], sh), na2);
            v.z = __vadd2(__funnelshift_r(wv[j][2], wv[j][3], sh), na2);
            v.w = __vadd2(__funnelshift_r(wv[j][3], wv[j][4], sh), na2);
            t4[g] = v;
          }
        }
      }
      // the < 8 entries before the first / after the last full group (adjacent rows of a group tile own the rest)
      const uint32_t head_end = min(8 * g_lo, o + cnt), tail_beg = max(8 * g_hi, head_end);
      if (lane < 8 && o + lane < head_end) tile[o + lane] = (uint16_t)(na + prm.nb[c_lo + lane]);
      if (lane >= 8 && lane < 16 && tail_beg + (lane - 8) < o + cnt)
        tile[tail_beg + (lane - 8)] = (uint16_t)(na + prm.nb[c_lo + (tail_beg + (lane - 8) - o)]);
    }
  };

  // ---- work-item pipeline.  With item j current: item j+1's ranges are in flight (cp.async into its
  // cursor array), item j+2's list offsets and item j+3's RowItem are in flight in registers.  What
  // has landed lives in the shared ring s_st[item % 4] (the service warp is its only user), so the
  // pipeline costs six registers.  All values are warp-uniform; lane 0 writes shared memory.
  uint32_t row_seq = 0;    // row items that took a shared cursor array so far
  uint32_t spill_seq = 0;  // row items that took a spill area so far
  auto item_index = [&](uint32_t n) { return (uint64_t)blockIdx.x + (uint64_t)n * gridDim.x; };
  auto ld_item = [&](uint32_t n) {
    const uint64_t idx = item_index(n);
    return idx < prm.n_items ? __ldg(reinterpret_cast<const uint4*>(prm.items + idx)) : make_uint4(0u, 0u, 0u, 0u);
  };
  // Starts copying a row item's ranges into its cursor array.  Array choice: the s-th row item with
  // R <= kCurCap takes shared array s % kCurRing.  This runs when the item before-before it is being
  // finished, inside produce() of a tile t, which follows empty[t-kRowBufs] (hence done[t-kRowBufs]).
  // The array's previous user is kCurRing = kRowBufs + 2 row items back; the two row items in between
  // and it each own at least one tile, so its tiles are all <= t-kRowBufs: every walker is through with
  // it.  Spilled rows (R > kCurCap) rotate over kCurRing global areas, written synchronously in
  // produce() of the row's first tile (same argument with distance kCurRing >= kRowBufs).
  auto stage_prefetch = [&](Staged& st) {
    if (!st.valid || st.it.a0 != st.it.a1) return;
    if (st.R <= (uint32_t)kCurCap) {
      st.arr = row_seq++ % kCurRing;
      uint2* dst = s_cur + st.arr * kCurCap;
      for (uint32_t k = lane; k < st.R; k += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst + k)), "l"(prm.ranges + st.k0 + k)
                     : "memory");
    } else {
      st.arr = kCurRing + (spill_seq++ % kCurRing);
    }
  };
  auto make_staged = [&](uint32_t n, uint4 v) {
    Staged st;
    st.it = RowItem{v.x, v.y, v.z, v.w};
    st.k0 = st.R = st.arr = 0;
    st.valid = item_index(n) < prm.n_items;
    return st;
  };
  auto ld_offsets = [&](const Staged& st, uint32_t& k0, uint32_t& k1) {
    k0 = k1 = 0;
    if (st.valid) {
      k0 = __ldg(prm.roff + st.it.a0);
      k1 = prm.debug_nowalk ? k0 : __ldg(prm.roff + st.it.a1 + 1);
    }
  };

  uint32_t j = 0;   // current item (local sequence number)
  uint32_t seg = 0; // next tile of the current item
  bool exit_sent = false;
  uint4 pend_item;          // RowItem of item j+3
  uint32_t pend_k0, pend_k1;  // list offsets of item j+2
  {
    // prologue: items 0..2 complete enough, item 3 in flight
    Staged s0 = make_staged(0, ld_item(0)), s1 = make_staged(1, ld_item(1)), s2 = make_staged(2, ld_item(2));
    uint32_t a, b;
    ld_offsets(s0, a, b); s0.k0 = a; s0.R = b - a;
    ld_offsets(s1, a, b); s1.k0 = a; s1.R = b - a;
    stage_prefetch(s0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    stage_prefetch(s1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    ld_offsets(s2, pend_k0, pend_k1);
    pend_item = ld_item(3);
    if (lane == 0) { s_st[0] = s0; s_st[1] = s1; s_st[2] = s2; }
    __syncwarp();
  }

  auto advance_item = [&]() {
    // (a) item j+3's RowItem and (b) item j+2's offsets have landed
    Staged s3 = make_staged(j + 3, pend_item);
    Staged s2 = s_st[(j + 2) & 3];
    s2.k0 = pend_k0;
    s2.R = pend_k1 - pend_k0;
    // (c) ranges of item j+2, (d) offsets of item j+3, (e) RowItem of item j+4
    stage_prefetch(s2);
    asm volatile("cp.async.commit_group;" ::: "memory");  // one group per advance, possibly empty
    ld_offsets(s3, pend_k0, pend_k1);
    pend_item = ld_item(j + 4);
    __syncwarp();  // every lane has read s_st[(j+2)&3]
    if (lane == 0) { s_st[(j + 2) & 3] = s2; s_st[(j + 3) & 3] = s3; }
    __syncwarp();
    ++j;
    seg = 0;
  };

  // Produces the next tile of this CTA's sequence into ring slot `slot` (or the exit marker).
  auto produce = [&](uint32_t q) {
    if (exit_sent) return;
    const uint32_t slot = q % kRowBufs;
    TICK(t_work);
    if (q >= (uint32_t)kRowBufs) mbar_wait(&s_empty[slot], ((q / kRowBufs) & 1u) ^ 1u, prm.debug_sleep);  // tile q-3 has left the slot
    TICK(t_wait);
    TileDesc* d = &s_desc[slot];
    const Staged cs = s_st[j & 3];
    if (!cs.valid) {
      if (lane == 0) d->kind = kTileExit;
      exit_sent = true;
    } else {
      uint16_t* tile = tiles + slot * kStride;
      const RowItem it = cs.it;
      if (it.a0 == it.a1) {
        const uint32_t a = it.a0, cb = it.cb, ce = it.ce, R = cs.R, k0 = cs.k0;
        const uint32_t len = ce - cb, nseg = (len + kTile - 1) / kTile;
        if (seg == 0) {
          // the row's suffix cursors: all cp.async groups but the newest (item j+1's) have landed
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          __syncwarp();
          const bool spilled = cs.arr >= (uint32_t)kCurRing;
          uint2* cur = !spilled ? s_cur + cs.arr * kCurCap
                                : prm.cursor_ws + ((size_t)blockIdx.x * kCurRing + (cs.arr - kCurRing)) * prm.cursor_ws_stride;
          if (spilled || cb > a + 1) {
            for (uint32_t k = lane; k < R; k += 32) {
              uint2 r = spilled ? prm.ranges[k0 + k] : cur[k];
              if (cb > a + 1) {  // shard starts inside the row: first member >= cb
                uint32_t lo = r.x, hi = r.y;
                while (lo < hi) {
                  const uint32_t mid = (lo + hi) >> 1;
                  if (__ldg(members + mid) < cb) lo = mid + 1; else hi = mid;
                }
                r.x = lo;
              }
              cur[k] = r;
            }
          }
        }
        const uint32_t c_lo = cb + seg * kTile, c_hi = min(ce, c_lo + (uint32_t)kTile);
        const uint64_t p_lo = row_start(a, T) + (c_lo - a - 1), p_hi = p_lo + (c_hi - c_lo), P0 = p_lo & ~7ull;
        fill_row(tile, a, c_lo, p_lo, p_hi, P0, prm.nb_shift != nullptr);
        if (lane == 0) {
          d->kind = kTileRow;
          d->a0 = a;
          d->nr = 1;
          d->k0 = k0;
          d->R = R;
          d->c_hi = c_hi;
          d->base = (int32_t)(p_lo - P0) - (int32_t)c_lo;
          d->cur_sel = cs.arr;
          d->flags = (seg == 0 ? 1u : 0u) | (seg + 1 == nseg ? 2u : 0u);
          d->p_lo = p_lo;
          d->p_hi = p_hi;
        }
        if (++seg == nseg) advance_item();
      } else {
        const uint32_t a0 = it.a0, nr = it.a1 - it.a0 + 1;
        const uint64_t p_lo = row_start(a0, T), p_hi = row_start((uint64_t)it.a1 + 1, T), P0 = p_lo & ~7ull;
        if (prm.uniform_nb) {
          fill_row(tile, a0, a0 + 1, p_lo, p_hi, P0);
        } else {
          for (uint32_t r = 0; r < nr; ++r) {
            const uint32_t a = a0 + r;
            const uint64_t rs = row_start(a, T);
            fill_row(tile, a, a + 1, rs, rs + (T - 1 - a), P0);
          }
        }
        if (lane == 0) {
          d->kind = kTileGroup;
          d->a0 = a0;
          d->nr = nr;
          d->k0 = cs.k0;
          d->R = cs.R;
          d->c_hi = T;
          d->base = 0;
          d->cur_sel = 0;
          d->flags = 3u;
          d->p_lo = p_lo;
          d->p_hi = p_hi;
        }
        advance_item();
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_full[slot]);
  };

  // mask bits, unaligned ends and the bulk store of the finished tile in ring slot `slot`
  // part / nparts: the mask's interior groups are shared between the emitter (part 0, which also does the
  // partial groups and the stores) and, in the kHelper build, the helper warp (part 1, mask only)
  auto emit = [&](uint32_t slot, uint32_t part, uint32_t nparts) {
    const uint16_t* tile = tiles + slot * kStride;
    const uint64_t p_lo = s_desc[slot].p_lo, p_hi = s_desc[slot].p_hi, P0 = p_lo & ~7ull;
    if (prm.editdist >= 0) {
      // 8 entries (one aligned 16-byte group of the tile) per lane -> one mask byte: 1 shared load, 4 SIMD
      // compares and a handful of bit operations (the emitter warp is alone on this and has to keep up with
      // 14 walkers).  Bytes whose 32-bit mask word lies wholly inside the tile are plain stores; the few
      // groups in a word shared with a neighbouring tile are OR-ed into the pre-zeroed mask, so every word
      // is written either by one tile's stores or by atomics only.
      const uint32_t thr = (uint32_t)prm.editdist, thr2 = thr | (thr << 16);
      const uint32_t m_lo = (uint32_t)(p_lo - P0), m_hi = (uint32_t)(p_hi - P0);  // entries [m_lo, m_hi) of the tile
      const uint32_t g_lo = (m_lo + 7) >> 3, g_hi = m_hi >> 3;                     // fully covered groups [g_lo, g_hi)
      const uint64_t gb0 = (P0 - prm.p_begin) >> 3;                                // mask byte of group 0
      const uint32_t ph = (uint32_t)gb0 & 3u;
      const uint32_t gi_lo = g_lo + ((4u - ((ph + g_lo) & 3u)) & 3u);              // first group on a word boundary
      const uint32_t gi_hi = g_hi - ((ph + g_hi) & 3u);                            // interior = [gi_lo, gi_hi) if gi_lo < gi_hi
      uint8_t* mb = reinterpret_cast<uint8_t*>(prm.mask) + gb0;
      const uint4* t4 = reinterpret_cast<const uint4*>(tile);
      auto group_bits = [&](uint32_t g) {
        const uint4 v = t4[g];
        // per-halfword (entry <= thr) -> 1 at bits 0 and 16; interleave the four words, fold the high halves down
        const uint32_t t = __vsetleu2(v.x, thr2) | (__vsetleu2(v.y, thr2) << 2) | (__vsetleu2(v.z, thr2) << 4) |
                           (__vsetleu2(v.w, thr2) << 6);
        return (t & 0x55u) | ((t >> 15) & 0xaau);
      };
      auto or_byte = [&](uint32_t g, uint32_t bits) {
        const uint64_t byte = gb0 + g;
        if (bits) atomicOr(&prm.mask[byte >> 2], bits << ((byte & 3) * 8));
      };
#pragma unroll 2
      for (uint32_t g = g_lo + lane + 32 * part; g < g_hi; g += 32 * nparts) {
        const uint32_t bits = group_bits(g);
        if (g >= gi_lo && g < gi_hi) mb[g] = (uint8_t)bits;
        else or_byte(g, bits);
      }
      if (lane < 2 && part == 0) {
        // lane 0: the group holding the first entries (if it is partial); lane 1: the one holding the last
        const bool head = lane == 0;
        const bool have = head ? (m_lo & 7u) != 0 : ((m_hi & 7u) != 0 && (g_hi >= g_lo || (m_lo & 7u) == 0));
        if (have) {
          const uint32_t g = head ? m_lo >> 3 : g_hi;
          uint32_t valid = 0xffu;
          if ((m_lo >> 3) == g) valid &= 0xffu << (m_lo & 7u);
          if ((m_hi >> 3) == g) valid &= (1u << (m_hi & 7u)) - 1u;
          or_byte(g, group_bits(g) & valid);
        }
      }
    }
    if (part != 0) return;
    // tile-index space: entry i is packed position P0 + i
    const uint32_t o_lo = (uint32_t)(p_lo - P0), o_hi = (uint32_t)(p_hi - P0);
    const uint32_t ia = (o_lo + 7) & ~7u, ib = o_hi & ~7u;
    uint16_t* g = prm.out + (P0 - prm.p_begin);
    if (lane < 8) {
      const uint32_t i = o_lo + lane;
      if (i < min(ia, o_hi)) g[i] = tile[i];
    } else if (lane < 16) {
      const uint32_t i = ib + (lane - 8);
      if (ib >= ia && i < o_hi) g[i] = tile[i];
    }
    if (lane == 0 && ib > ia) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g + ia),
                   "r"(smem_addr(tile + ia)), "r"((ib - ia) * 2u)
                   : "memory");
    }
    if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // one group per tile, even if empty
  };

  if (warp == 0) {
    for (uint32_t q = 0; !exit_sent; ++q) produce(q);
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp == 1) {
    const uint32_t nparts = (kHelper && prm.editdist >= 0) ? 2u : 1u;
    for (uint32_t e = 0;; ++e) {
      const uint32_t slot = e % kRowBufs, par = (e / kRowBufs) & 1u;
      TICK(t_work);
      mbar_wait(&s_full[slot], par);  // descriptor published
      if (s_desc[slot].kind == kTileExit) break;
      mbar_wait(&s_done[slot], par, prm.debug_sleep);
      TICK(t_wait);
      emit(slot, 0, nparts);
      // hand the slot back as soon as the bulk copy has read it (the emitter has nothing else to do):
      // the producer can then refill it while the walkers are still busy with the next two tiles
      __syncwarp();  // this warp's own reads of the tile are done
      TICK(t_work);
      if (lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&s_empty[slot]);
      }
      __syncwarp();
      TICK(t_wait2);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  } else {
    // helper warp (kHelper builds): the other half of the threshold mask of every tile
    for (uint32_t e = 0;; ++e) {
      const uint32_t slot = e % kRowBufs, par = (e / kRowBufs) & 1u;
      mbar_wait(&s_full[slot], par);
      if (s_desc[slot].kind == kTileExit) break;
      mbar_wait(&s_done[slot], par, prm.debug_sleep);
      if (prm.editdist >= 0) emit(slot, 1, 2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[slot]);
    }
  }
  __syncwarp();
#ifdef PRF_TIMING
  if (lane == 0 && prm.timing) {
    prm.timing[(size_t)blockIdx.x * 64 + 2 * warp] = t_wait;
    prm.timing[(size_t)blockIdx.x * 64 + 2 * warp + 1] = t_work;
    if (warp == 1) prm.timing[(size_t)blockIdx.x * 64 + 63] = t_wait2;
  }
#endif
}

// Cuts packed positions [p_begin, p_end) into work items, longest first.
void build_row_plan(uint32_t T, uint64_t p_begin, uint64_t p_end, uint32_t tile, std::vector<RowItem>& rows,
                           std::vector<RowItem>& groups) {
  rows.clear();
  groups.clear();
  if (p_end <= p_begin || T < 2) return;
  const uint32_t a_first = row_of(p_begin, T), a_last = row_of(p_end - 1, T);
  uint32_t a = a_first;
  while (a <= a_last) {
    const uint64_t rs = row_start(a, T);
    const uint32_t cb = a + 1 + (uint32_t)(std::max(rs, p_begin) - rs);
    const uint32_t ce = a + 1 + (uint32_t)(std::min(rs + (T - 1 - a), p_end) - rs);
    const bool whole = cb == a + 1 && ce == T;
    if (whole) {
      // greedy: as many following whole rows as fit one tile
      uint32_t e = a;
      uint64_t tot = T - 1 - a;
      while (e + 1 <= a_last && e + 1 - a < (uint32_t)kGroupRows) {
        const uint64_t rs2 = row_start(e + 1, T);
        const bool whole2 = rs2 + (T - 2 - e) <= p_end;  // row e+1 ends inside the shard
        const uint64_t l2 = T - 2 - e;
        if (!whole2 || tot + l2 > tile) break;
        tot += l2;
        ++e;
      }
      if (e > a) {
        groups.push_back(RowItem{a, e, 0u, 0u});
        a = e + 1;
        continue;
      }
    }
    if (ce > cb) rows.push_back(RowItem{a, a, cb, ce});
    ++a;
  }
  std::stable_sort(rows.begin(), rows.end(),
                   [](const RowItem& x, const RowItem& y) { return x.ce - x.cb > y.ce - y.cb; });
}

template <int kTile, int kWalkers, int kCurCap, int kSlots, int kRowBufs, int kHelper>
static int launch_sparse_t(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                           uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  constexpr int kThreads = 32 * (kWalkers + 2 + kHelper);
  constexpr size_t smem = (size_t)kRowBufs * (kTile + 8) * 2 + (size_t)(kRowBufs + 2) * kCurCap * sizeof(uint2);
  auto kern = sparse_rows_kernel<kTile, kWalkers, kCurCap, kSlots, kRowBufs, kHelper>;
  PRF_CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // per instantiation
  RowPlan& plan = ctx->plan;
  if (!plan.valid || plan.T != H.T || plan.p_begin != p_begin || plan.p_end != p_end || plan.tile != kTile) {  // (occupancy is the same for every instantiation of a tile size)
    plan.valid = false;
    std::vector<RowItem> rows, groups;
    build_row_plan(H.T, p_begin, p_end, kTile, rows, groups);
    rows.insert(rows.end(), groups.begin(), groups.end());
    PRF_CK(ctx, plan.items.alloc(rows.size(), s));
    if (!rows.empty())
      PRF_CK(ctx, cudaMemcpyAsync(plan.items.p, rows.data(), rows.size() * sizeof(RowItem), cudaMemcpyHostToDevice, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));  // `rows` is pageable host memory about to go out of scope
    plan.n_items = (uint32_t)rows.size();
    plan.T = H.T;
    plan.p_begin = p_begin;
    plan.p_end = p_end;
    plan.tile = kTile;
    int nb = 0;
    PRF_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem));
    plan.ctas_per_sm = nb > 0 ? nb : 1;
    plan.valid = true;
  }
  if (plan.n_items == 0) return PRF_OK;
  const uint32_t grid = std::min<uint32_t>(plan.n_items, (uint32_t)(ctx->sm_count * plan.ctas_per_sm));
  RowParams prm{};
  prm.T = H.T;
  prm.p_begin = p_begin;
  prm.nb = H.nb.p;
  prm.nb_shift = H.nb_shift.p;
  prm.nb_shift_stride = H.nb_shift_stride;
  prm.uniform_nb = H.uniform_nb ? 1 : 0;
  prm.nb_uniform = H.nb_uniform;
  prm.roff = H.roff.p;
  prm.ranges = H.ranges.p;
  prm.members = H.lmembers.p;
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  prm.items = plan.items.p;
  prm.n_items = plan.n_items;
  if (H.max_ranges > (uint32_t)kCurCap) {
    const size_t need = (size_t)grid * (kRowBufs + 2) * H.max_ranges;
    if (plan.cursor_ws.n < need) PRF_CK(ctx, plan.cursor_ws.alloc(need, s));
    prm.cursor_ws = plan.cursor_ws.p;
    prm.cursor_ws_stride = H.max_ranges;
  }
  {
    const char* e = getenv("PRF_DEBUG_NOWALK");
    prm.debug_nowalk = e && *e == '1';
    const char* z = getenv("PRF_SPARSE_SLEEP");
    prm.debug_sleep = z ? (uint32_t)atoi(z) : 0u;
  }
#ifdef PRF_TIMING
  if (ctx->timing.n < (size_t)grid * 64) PRF_CK(ctx, ctx->timing.alloc((size_t)grid * 64, s));
  PRF_CK(ctx, cudaMemsetAsync(ctx->timing.p, 0, ctx->timing.bytes(), s));
  prm.timing = ctx->timing.p;
  ctx->timing_grid = grid;
#endif
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  kern<<<grid, kThreads, smem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

int launch_sparse(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                         uint32_t* d_mask) {
  // the tiny-tile build exists so that small test inputs exercise multi-tile rows and cursor spill
  // With a threshold mask the emitter would be the bottleneck (one warp, 16 K compares per tile): one of
  // the walker warps becomes a helper that builds half of the mask.
  const bool mask = d_mask != nullptr;
  if (ctx->debug_small_tiles)
    return mask ? launch_sparse_t<256, 2, 8, 8, 3, 1>(ctx, p_begin, p_end, editdist, d_out, d_mask)
                : launch_sparse_t<256, 2, 8, 8, 3, 0>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  static const int cfg = [] { const char* e = getenv("PRF_SPARSE_CFG"); return e ? atoi(e) : 0; }();  // 4: force 12 slots
  // trees with clearly more suffix ranges than 14 walkers x 8 register slots: 12 slots, list ends in shared memory
  const bool many = cfg == 4 || (cfg == 0 && ctx->h.max_ranges > 14 * 8 + 16);
  if (mask)
    return many ? launch_sparse_t<16384, 13, 320, 12, 3, 1>(ctx, p_begin, p_end, editdist, d_out, d_mask)
                : launch_sparse_t<16384, 13, 320, 8, 3, 1>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  return many ? launch_sparse_t<16384, 14, 320, 12, 3, 0>(ctx, p_begin, p_end, editdist, d_out, d_mask)
              : launch_sparse_t<16384, 14, 320, 8, 3, 0>(ctx, p_begin, p_end, editdist, d_out, d_mask);
}

}  // namespace prf
