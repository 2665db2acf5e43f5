// prf_sparse.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), sparse variant.
//
// The reference bumps d(i,j) once for every bipartition that exactly one of the two trees has
// (haveIt x dontHave, :327-338): d(a,b) = |B_a| + |B_b| - 2 |B_a n B_b|.  Random topologies share
// ~0.13 bipartitions per pair, so almost every entry is just |B_a| + |B_b| and the work is writing
// the matrix: 2 bytes per pair, HBM-write bound.
//
// Work is cut along ROWS of the packed upper triangle.  A work item is either one row a restricted
// to columns [cb, ce) (long rows, and the partial first/last row of a shard), or a group of whole
// short rows that fit one tile.  Persistent CTAs take items from an atomic counter (longest first).
// A row is produced tile by tile (kTile consecutive columns) through a ring of three shared-memory
// buffers:
//   fill   d = |B_a| + |B_b| for the tile's columns (16-byte shared stores);
//   walk   tree a's shared bipartitions each own a suffix of an inverted list ("trees b > a that also
//          have it", ascending).  Every suffix keeps a cursor in shared memory that only moves forward
//          from tile to tile, so each member is loaded exactly once per row and there is no search.
//          One warp takes kWalkUnroll suffixes at a time, issues their loads together, and subtracts
//          2 at column b with a shared-memory atomic -- no global atomics on the matrix;
//   store  the finished tile goes out as ONE bulk asynchronous copy (cp.async.bulk shared -> global,
//          the TMA engine; 16-byte aligned interior) while the CTA already walks the next tile.  The
//          < 8 entries before/after the aligned interior are stored by a few threads.
// Every output byte is written exactly once.  One __syncthreads per tile.
#pragma once

#include <algorithm>

#include "prf_common.cuh"

namespace prf {

constexpr int kRowTile = 16384;   // entries per tile buffer (32 KB)
constexpr int kRowThreads = 512;
constexpr int kRowBufs = 3;
constexpr int kRowCurCap = 1024;  // suffix cursors of one row kept in shared memory (more spill to global)
constexpr int kGroupRows = 256;   // most rows in a group item
constexpr int kWalkUnroll = 4;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

struct RowParams {
  uint32_t T;
  uint64_t p_begin;
  const uint16_t* nb;
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
  uint32_t* counter;
  uint2* cursor_ws;           // per-CTA spill area for rows with more than kCurCap suffixes
  uint32_t cursor_ws_stride;  // entries per CTA
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

template <int kTile, int kThreads, int kCurCap>
__global__ void __launch_bounds__(kThreads, 2) sparse_rows_kernel(const RowParams prm) {
  constexpr int kStride = kTile + 8;  // entries per buffer: a tile keeps the 16-byte phase of its global address
  constexpr uint32_t kWarps = kThreads / 32;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint16_t* tiles = reinterpret_cast<uint16_t*>(smem_raw);
  uint2* s_cur = reinterpret_cast<uint2*>(smem_raw + (size_t)kRowBufs * kStride * 2);
  __shared__ uint32_t s_roff[kGroupRows + 1];
  __shared__ uint32_t s_item[2];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t T = prm.T;
  const uint32_t* __restrict__ members = prm.members;
  const uint32_t vv = (2u * prm.nb_uniform) | ((2u * prm.nb_uniform) << 16);

  // d = |B_a| + |B_b| for packed positions [p_lo, p_hi) of row a starting at column c_lo
  auto fill_row = [&](uint16_t* tile, uint32_t a, uint32_t c_lo, uint64_t p_lo, uint64_t p_hi, uint64_t P0) {
    if (prm.uniform_nb) {
      const uint32_t cnt = (uint32_t)((p_hi - P0 + 7) >> 3);
      uint4* t4 = reinterpret_cast<uint4*>(tile);
      const uint4 q4 = make_uint4(vv, vv, vv, vv);
      for (uint32_t i = tid; i < cnt; i += kThreads) t4[i] = q4;
    } else {
      const uint32_t na = prm.nb[a];
      const uint32_t cnt = (uint32_t)(p_hi - p_lo), o = (uint32_t)(p_lo - P0);
      const uint16_t* nbb = prm.nb + c_lo;
      for (uint32_t i = tid; i < cnt; i += kThreads) tile[o + i] = (uint16_t)(na + nbb[i]);
    }
  };

  // mask bits, unaligned ends and the bulk store of a finished tile holding [p_lo, p_hi)
  auto emit = [&](const uint16_t* tile, uint64_t p_lo, uint64_t p_hi, uint64_t P0) {
    if (prm.editdist >= 0) {
      const uint64_t w_lo = (p_lo - prm.p_begin) >> 5, w_hi = (p_hi - 1 - prm.p_begin) >> 5;
      for (uint64_t w = w_lo + warp; w <= w_hi; w += kWarps) {
        const uint64_t pw = prm.p_begin + (w << 5);
        const uint64_t p = pw + lane;
        const bool hit = p >= p_lo && p < p_hi && (int32_t)tile[p - P0] <= prm.editdist;
        const uint32_t bits = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) {
          if (pw >= p_lo && pw + 32 <= p_hi) prm.mask[w] = bits;
          else if (bits) atomicOr(&prm.mask[w], bits);  // word shared with a neighbouring row (mask is pre-zeroed)
        }
      }
    }
    const uint64_t pa = (p_lo + 7) & ~7ull, pb = p_hi & ~7ull;
    uint16_t* outp = prm.out - prm.p_begin;  // indexed by absolute packed position
    if (pa >= p_hi) {
      if (tid < 8 && p_lo + tid < p_hi) outp[p_lo + tid] = tile[p_lo + tid - P0];
      return;
    }
    if (tid < 8 && p_lo + tid < pa) outp[p_lo + tid] = tile[p_lo + tid - P0];
    if (tid >= 32 && tid < 40 && pb + (tid - 32) < p_hi) outp[pb + (tid - 32)] = tile[pb + (tid - 32) - P0];
    if (tid == 0 && pb > pa) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(outp + pa),
                   "r"(smem_addr(tile + (pa - P0))), "r"((uint32_t)((pb - pa) * 2))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  };

  uint32_t q = 0;  // tiles produced so far by this CTA; tile q lives in buffer q % kRowBufs
  if (tid == 0) s_item[0] = blockIdx.x;
  __syncthreads();
  uint32_t par = 0;
  for (;;) {
    const uint32_t it = s_item[par];
    if (it >= prm.n_items) break;
    uint32_t next = 0;
    if (tid == 0) next = atomicAdd(prm.counter, 1u) + gridDim.x;  // latency hides behind this item
    const RowItem item = prm.items[it];

    if (item.a0 == item.a1) {
      // ---------------- one row, columns [cb, ce) ----------------
      const uint32_t a = item.a0, cb = item.cb, ce = item.ce;
      const uint32_t k0 = __ldg(prm.roff + a), R = __ldg(prm.roff + a + 1) - k0;
      const uint64_t p_row = row_start(a, T) + (cb - a - 1);
      uint2* cur = R <= (uint32_t)kCurCap ? s_cur : prm.cursor_ws + (size_t)blockIdx.x * prm.cursor_ws_stride;
      for (uint32_t k = tid; k < R; k += kThreads) {
        uint2 r = prm.ranges[k0 + k];
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
      const uint32_t len = ce - cb, nseg = (len + kTile - 1) / kTile;
      {
        const uint64_t p_hi = p_row + min(len, (uint32_t)kTile);
        fill_row(tiles + (q % kRowBufs) * kStride, a, cb, p_row, p_hi, p_row & ~7ull);
      }
      __syncthreads();
      for (uint32_t s = 0; s < nseg; ++s) {
        const uint32_t c_lo = cb + s * kTile, c_hi = min(ce, c_lo + (uint32_t)kTile);
        const uint64_t p_lo = p_row + (uint64_t)s * kTile, p_hi = p_lo + (c_hi - c_lo), P0 = p_lo & ~7ull;
        uint16_t* tile = tiles + (q % kRowBufs) * kStride;
        uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
        const int32_t base = (int32_t)(p_lo - P0) - (int32_t)c_lo;
        for (uint32_t kb = warp * kWalkUnroll; kb < R; kb += kWarps * kWalkUnroll) {
          uint2 c[kWalkUnroll];
          uint32_t b0[kWalkUnroll], b1[kWalkUnroll];
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) c[u] = kb + u < R ? cur[kb + u] : make_uint2(0u, 0u);
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            const uint32_t m = c[u].x + lane;
            b0[u] = m < c[u].y ? __ldg(members + m) : 0xffffffffu;
            b1[u] = m + 32 < c[u].y ? __ldg(members + m + 32) : 0xffffffffu;
          }
#pragma unroll
          for (int u = 0; u < kWalkUnroll; ++u) {
            if (c[u].x < c[u].y) {
              const uint32_t nc = walk_suffix(members, c[u].x, c[u].y, b0[u], b1[u], c_hi, base, tile32, lane);
              if (lane == 0 && nc != c[u].x) cur[kb + u].x = nc;
            }
          }
        }
        if (s + 1 < nseg) {
          const uint32_t n_lo = c_lo + kTile, n_hi = min(ce, n_lo + (uint32_t)kTile);
          const uint64_t np_lo = p_lo + kTile;
          fill_row(tiles + ((q + 1) % kRowBufs) * kStride, a, n_lo, np_lo, np_lo + (n_hi - n_lo), np_lo & ~7ull);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to the bulk copy
        if (tid == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // tile q-1 has left shared memory
          if (s + 1 == nseg) s_item[par ^ 1] = next;
        }
        __syncthreads();
        emit(tile, p_lo, p_hi, P0);
        ++q;
      }
    } else {
      // ---------------- a group of whole short rows a0..a1 in one tile ----------------
      const uint32_t a0 = item.a0, nr = item.a1 - item.a0 + 1;
      const uint64_t p_lo = row_start(a0, T), p_hi = row_start((uint64_t)item.a1 + 1, T), P0 = p_lo & ~7ull;
      uint16_t* tile = tiles + (q % kRowBufs) * kStride;
      uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
      for (uint32_t i = tid; i <= nr; i += kThreads) s_roff[i] = __ldg(prm.roff + a0 + i);
      if (prm.uniform_nb) {
        fill_row(tile, a0, a0 + 1, p_lo, p_hi, P0);
      } else {
        for (uint32_t r = 0; r < nr; ++r) {
          const uint32_t a = a0 + r;
          const uint64_t rs = row_start(a, T);
          fill_row(tile, a, a + 1, rs, rs + (T - 1 - a), P0);
        }
      }
      __syncthreads();
      const uint32_t k0 = s_roff[0], R = s_roff[nr] - k0;
      for (uint32_t kb = warp * kWalkUnroll; kb < R; kb += kWarps * kWalkUnroll) {
        uint2 c[kWalkUnroll];
        uint32_t b0[kWalkUnroll], b1[kWalkUnroll];
        int32_t base[kWalkUnroll];
#pragma unroll
        for (int u = 0; u < kWalkUnroll; ++u) {
          c[u] = make_uint2(0u, 0u);
          base[u] = 0;
          if (kb + u < R) {
            const uint32_t k = k0 + kb + u;
            c[u] = prm.ranges[k];
            uint32_t lo = 0, hi = nr - 1;  // largest i with s_roff[i] <= k
            while (lo < hi) {
              const uint32_t mid = (lo + hi + 1) >> 1;
              if (s_roff[mid] <= k) lo = mid; else hi = mid - 1;
            }
            const uint32_t a = a0 + lo;
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
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (tid == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        s_item[par ^ 1] = next;
      }
      __syncthreads();
      emit(tile, p_lo, p_hi, P0);
      ++q;
    }
    par ^= 1;
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Cuts packed positions [p_begin, p_end) into work items, longest first.
static void build_row_plan(uint32_t T, uint64_t p_begin, uint64_t p_end, uint32_t tile, std::vector<RowItem>& rows,
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

template <int kTile, int kThreads, int kCurCap>
static int launch_sparse_t(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                           uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  constexpr size_t smem = (size_t)kRowBufs * (kTile + 8) * 2 + (size_t)kCurCap * sizeof(uint2);
  auto kern = sparse_rows_kernel<kTile, kThreads, kCurCap>;
  RowPlan& plan = ctx->plan;
  if (!plan.valid || plan.T != H.T || plan.p_begin != p_begin || plan.p_end != p_end || plan.tile != kTile) {
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
    plan.ctas_per_sm = 0;
    if (!plan.counter.p) PRF_CK(ctx, plan.counter.alloc(1, s));
    if (plan.ctas_per_sm == 0) {
      PRF_CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int nb = 0;
      PRF_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kThreads, smem));
      plan.ctas_per_sm = nb > 0 ? nb : 1;
    }
    plan.valid = true;
  }
  if (plan.n_items == 0) return PRF_OK;
  const uint32_t grid = std::min<uint32_t>(plan.n_items, (uint32_t)(ctx->sm_count * plan.ctas_per_sm));
  RowParams prm{};
  prm.T = H.T;
  prm.p_begin = p_begin;
  prm.nb = H.nb.p;
  prm.uniform_nb = H.uniform_nb ? 1 : 0;
  prm.nb_uniform = H.nb_uniform;
  prm.roff = H.roff.p;
  prm.ranges = H.ranges.p;
  prm.members = H.members.p;
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  prm.items = plan.items.p;
  prm.n_items = plan.n_items;
  prm.counter = plan.counter.p;
  if (H.stats.max_bips > (uint32_t)kCurCap) {
    const size_t need = (size_t)grid * H.stats.max_bips;
    if (plan.cursor_ws.n < need) PRF_CK(ctx, plan.cursor_ws.alloc(need, s));
    prm.cursor_ws = plan.cursor_ws.p;
    prm.cursor_ws_stride = H.stats.max_bips;
  }
  PRF_CK(ctx, cudaMemsetAsync(plan.counter.p, 0, sizeof(uint32_t), s));
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  kern<<<grid, kThreads, smem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

static int launch_sparse(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                         uint32_t* d_mask) {
  // the tiny-tile build exists so that small test inputs exercise multi-tile rows and cursor spill
  if (ctx->debug_small_tiles) return launch_sparse_t<256, 128, 8>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  return launch_sparse_t<kRowTile, kRowThreads, kRowCurCap>(ctx, p_begin, p_end, editdist, d_out, d_mask);
}

}  // namespace prf
