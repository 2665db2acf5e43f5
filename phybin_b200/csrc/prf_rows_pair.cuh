// prf_rows_pair.cuh -- the matrix kernel of prf_rows.cuh with TWO warps per row.
//
// rows_kernel (one warp per row) is bound by the latency of each warp's own chain -- bucket bounds, scan, search, group
// loads, reductions, store -- at the 13 warps per SM that 16 KB tile buffers leave room for (issue slots 30 % busy).
// Here a PAIR of warps owns a row and its tile buffer, so 26 warps are resident with the same shared memory: the two
// warps fill half of the tile each, take half of the row's bucket groups each (the merge-path sequence is cut into 64
// runs instead of 32) and half of the singles / mask groups; warp 0 of the pair stores the tile.  They meet at three
// 64-thread named barriers per tile: buffer free (previous bulk store has read it), tile filled, tile finished.
// Everything a warp needs about the row (lists, bounds, prefix sums) it computes itself -- both warps write identical
// values into the pair's list arrays -- so no other synchronisation is needed.
#pragma once

#include "prf_rows.cuh"

namespace prf {

template <int kLog2B, int kPairs, int kCap, int kQ>
__global__ void __launch_bounds__(64 * kPairs, 1) rows_pair_kernel(const RowsParams prm) {
  constexpr int kTile = 1 << kLog2B;
  constexpr int kStride = kTile + 8;  // entries per buffer: a tile keeps the 16-byte phase of its global address
  constexpr int kPairBytes = kStride * 2 + kCap * 12 + 16;  // per pair: tile buffer, then trow / start / upre of the row's lists
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_fill[kPairs];  // completion of a pair's bulk load of |B| (non-uniform rows)
  __shared__ uint32_t s_item[kPairs];               // the pair's current work item

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1, h = warp & 1;
  const uint32_t L = h * 32 + lane;  // lane of the pair
  uint16_t* tile = reinterpret_cast<uint16_t*>(smem_raw + (size_t)pair * kPairBytes);
  uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);
  uint32_t* s_trow = reinterpret_cast<uint32_t*>(smem_raw + (size_t)pair * kPairBytes + kStride * 2);  // part << 28 | first bucket of the list
  uint32_t* s_start = s_trow + kCap;   // first lmem entry of the list's bucket in the current block
  uint32_t* s_upre = s_start + kCap;   // [kCap + 1] 16-byte groups of the lists before this one (exclusive prefix)
  const uint32_t T = prm.T, NB = prm.NB;
  auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); };

  if (L == 0) {
    mbar_init(&s_fill[pair], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pair_sync();
  uint32_t fill_phase = 0;

  uint64_t pol_keep, pol_stream;  // see ldg_keep in prf_rows.cuh
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  const uint32_t vv = (2u * prm.nb_uniform) | ((2u * prm.nb_uniform) << 16);
  const uint4 q4 = make_uint4(vv, vv, vv, vv);

  for (;;) {
    if (L == 0) s_item[pair] = atomicAdd(prm.counter, 1u);
    pair_sync();
    const uint32_t it = s_item[pair];  // (rewritten only after this row's tile barriers)
    if (it >= prm.n_items) break;
    const uint4 iv = __ldg(reinterpret_cast<const uint4*>(prm.items + it));
    const uint32_t a = iv.x, cb = iv.z, ce = iv.w;

    // ---- the row's long lists and singles over all parts (lane p < n_parts looks at part p; both warps alike) ----
    uint32_t pk0 = 0, pR = 0, ps0 = 0, pS = 0;
    if (lane < prm.n_parts) {
      const ListPart& q = prm.part[lane];
      pk0 = ldg_keep(q.roff + a, pol_keep);
      pR = ldg_keep(q.roff + a + 1, pol_keep) - pk0;
      ps0 = ldg_keep(q.soff + a, pol_keep);
      pS = ldg_keep(q.soff + a + 1, pol_keep) - ps0;
    }
    uint32_t pRx = pR, pSx = pS;  // inclusive scans over the parts (n_parts <= 16)
#pragma unroll
    for (int d = 1; d < kMaxParts; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, pRx, d), o2 = __shfl_up_sync(0xffffffffu, pSx, d);
      if (lane >= (uint32_t)d) {
        pRx += o;
        pSx += o2;
      }
    }
    const uint32_t R = __shfl_sync(0xffffffffu, pRx, kMaxParts - 1), S = __shfl_sync(0xffffffffu, pSx, kMaxParts - 1);
    pRx -= pR;  // exclusive: index of part p's first list among the row's lists
    pSx -= pS;
    // rows with <= 64 singles keep them in a register for all of the row's tiles (pair lane i: the i-th single)
    uint32_t my_single = 0xffffffffu;
    const bool s_reg = S <= 64;
    if (S && s_reg) {
      uint32_t p = 0;
      for (uint32_t q = 1; q < prm.n_parts; ++q)
        if (__shfl_sync(0xffffffffu, pSx, q) <= L) p = q;
      const uint32_t s0 = __shfl_sync(0xffffffffu, ps0, p), x = __shfl_sync(0xffffffffu, pSx, p);
      if (L < S) my_single = ldg_keep(prm.part[p].singles + s0 + (L - x), pol_keep);
    }
    // lists [c0, c0 + Rc) of the row -> s_trow (the pair's warps take alternate chunks of 32)
    auto load_lists = [&](uint32_t c0, uint32_t Rc) {
      for (uint32_t kb = 32 * h; kb < Rc; kb += 64) {
        const uint32_t k = c0 + kb + lane;
        uint32_t p = 0;
        for (uint32_t q = 1; q < prm.n_parts; ++q)  // warp-uniform loop, per-lane k
          if (__shfl_sync(0xffffffffu, pRx, q) <= k) p = q;
        const uint32_t k0 = __shfl_sync(0xffffffffu, pk0, p), x = __shfl_sync(0xffffffffu, pRx, p);
        if (kb + lane < Rc) s_trow[kb + lane] = (p << 28) | (ldg_keep(prm.part[p].rlist + k0 + (k - x), pol_keep) * NB);
      }
      pair_sync();
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

      const uint32_t tb = smem_addr(tile) + 2u * (uint32_t)(base + (int32_t)(j << kLog2B));  // entry of the block's column 0 (mod 2^32)
      const uint32_t safe = smem_addr(tile) + 4u * L;  // where an idle lane adds 0: one word per lane of the pair
      const uint32_t lo16 = c_lo - (j << kLog2B), span16 = cnt;  // columns of the tile relative to the block
      uint32_t k = 0, u_next = 0, u1 = 0;  // my run of groups [u_next, u1) of the current pass; k = list of group u_next
      uint4 v[kQ];
      // bucket bounds of lists [0, Rc) in block j -> s_start, s_upre (both warps compute and write the same values);
      // my run of groups: the sequence is cut into 64 runs
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
        const uint32_t q = (run + 63) >> 6;  // groups per lane of the pair
        u_next = min(run, L * q);
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
      if (R0) {
        setup_pass(R0, pf_ok);
        load_batch();
      }

      // ---- barrier 1: the previous tile's bulk store has finished reading the buffer (and the partner its mask) ----
      if (L == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      pair_sync();

      // ---- fill: d = |B_a| + |B_b|, half of the groups per warp ----
      uint4* t4 = reinterpret_cast<uint4*>(tile);
      if (prm.uniform_nb) {
        if (groups == kTile / 8 || groups == kTile / 8 + 1) {
#pragma unroll
          for (int i = 0; i < kTile / 512; ++i) t4[L + 64 * i] = q4;
          if (L == 0) t4[kTile / 8] = q4;
        } else {
          for (uint32_t g = L; g < groups; g += 64) t4[g] = q4;
        }
      } else {
        const uint32_t na2 = na | (na << 16);
        if (prm.nb_shift != nullptr && c_lo >= o) {
          // the run nb[c_lo - o ..) lands on tile entry 0: the copy of nb whose shift makes the source 16-byte aligned
          // comes in by one bulk copy, then |B_a| is added two entries at a time
          const uint32_t d0 = c_lo - o, ks = d0 & 7u;
          const uint16_t* src = prm.nb_shift + (size_t)ks * prm.nb_shift_stride + (d0 - ks);
          if (L == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&s_fill[pair])), "r"(groups * 16u)
                         : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_addr(tile)),
                         "l"(src), "r"(groups * 16u), "r"(smem_addr(&s_fill[pair]))
                         : "memory");
          }
          mbar_wait(&s_fill[pair], fill_phase);
          fill_phase ^= 1u;
#pragma unroll 4
          for (uint32_t g = L; g < groups; g += 64) {
            uint4 x = t4[g];
            x.x = __vadd2(x.x, na2);
            x.y = __vadd2(x.y, na2);
            x.z = __vadd2(x.z, na2);
            x.w = __vadd2(x.w, na2);
            t4[g] = x;
          }
        } else {
          // the first columns of the matrix (c_lo < 8): entry by entry
          for (uint32_t i = L; i < cnt; i += 64) tile[o + i] = (uint16_t)(na + prm.nb[c_lo + i]);
        }
      }
      // ---- barrier 2: the tile is filled (updates of either warp may land in either half) ----
      pair_sync();

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
          for (uint32_t sb = 0; sb < sn; sb += 64) {
            const uint32_t i = sb + L;
            const uint32_t b = i < sn ? ldg_keep(sg + i, pol_keep) : 0xffffffffu;
            if (b >= c_lo && b < c_hi) {
              const uint32_t idx = (uint32_t)(base + (int32_t)b);
              atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
            }
          }
        }
      }

      // ---- long lists ----
      if (R0) {
        while (__any_sync(0xffffffffu, u_next < u1)) {
          apply_batch();
          if (__any_sync(0xffffffffu, u_next < u1)) load_batch();
        }
      }
      for (uint32_t c0 = kCap; c0 < R; c0 += kCap) {  // rows with more lists than one pass holds
        const uint32_t Rc = min(R - c0, (uint32_t)kCap);
        pair_sync();  // the partner has finished with the previous pass's arrays
        load_lists(c0, Rc);
        setup_pass(Rc, false);
        while (__any_sync(0xffffffffu, u_next < u1)) {
          load_batch();
          apply_batch();
        }
      }
      if (R > (uint32_t)kCap) {  // back to the first pass's lists for the next tile
        pair_sync();
        load_lists(0, kCap);
      }
      // bucket bounds of the next block, for rows of <= 64 lists
      if (pf_ok && j < j_last) prefetch_bounds(j + 1);

      // ---- barrier 3: every update of the tile is done and visible to the bulk-copy engine ----
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      pair_sync();

      // ---- threshold mask: bit (p - p_begin) set iff d <= editdist (the warps take alternate spans of 32 groups) ----
      if (prm.editdist >= 0) {
        const uint32_t thr = (uint32_t)prm.editdist, thr2 = thr | (thr << 16);
        const uint64_t gb0 = (P0 - prm.p_begin) >> 3;   // mask byte of group 0
        const uint32_t ph = (uint32_t)gb0 & 3u;
        const uint32_t e_lo = o, e_hi = o + cnt;        // valid tile entries
        const uint32_t n_it = (groups + ph + 31) >> 5;
        for (uint32_t itn = h; itn < n_it; itn += 2) {
          const int32_t g = (int32_t)(itn * 32 + lane) - (int32_t)ph;
          uint32_t bits = 0;
          if (g >= 0 && (uint32_t)g < groups) {
            const uint4 x = t4[g];
            const uint32_t t = __vsetleu2(x.x, thr2) | (__vsetleu2(x.y, thr2) << 2) | (__vsetleu2(x.z, thr2) << 4) |
                               (__vsetleu2(x.w, thr2) << 6);
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

      // ---- store (warp 0 of the pair) ----
      if (h == 0) {
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
      }
    }
  }
  if (L == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int kLog2B, int kPairs, int kCap, int kQ>
static int launch_rows_pair_t(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  constexpr int kThreads = 64 * kPairs;
  constexpr size_t smem = (size_t)kPairs * (((1 << kLog2B) + 8) * 2 + kCap * 12 + 16);
  static_assert(smem + 512 <= 232448, "rows kernel: shared memory per CTA");
  if (H.block_log2 != (uint32_t)kLog2B) return ctx->fail(PRF_E_STATE, "list structure was built for another column block size");
  auto kern = rows_pair_kernel<kLog2B, kPairs, kCap, kQ>;
  PRF_CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int rc = prepare_row_plan(ctx, p_begin, p_end);
  if (rc) return rc;
  RowPlan& plan = ctx->plan;
  if (plan.n_items == 0) return PRF_OK;
  const uint32_t grid = std::min<uint32_t>((plan.n_items + kPairs - 1) / kPairs, (uint32_t)ctx->sm_count);
  RowsParams prm = make_rows_params(ctx, p_begin, editdist, d_out, d_mask);
  PRF_CK(ctx, cudaMemsetAsync(plan.counter.p, 0, sizeof(uint32_t), s));
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  kern<<<grid, kThreads, smem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
