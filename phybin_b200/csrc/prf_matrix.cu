// prf_matrix.cu -- the sparse matrix kernel (prf_rows.cuh) and the cut of a packed range into row work items.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_rows.cuh"
#include "prf_rows_pair.cuh"

namespace prf {

constexpr int kGroupRows = 256;  // most rows in a group item (tile > 0 only)

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

int launch_rows(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  // the tiny-block builds exist so that small test inputs exercise multi-tile rows and several passes over a row's lists
  if (ctx->debug_small_tiles)
    return ctx->rows_cfg >= 4 ? launch_rows_pair_t<8, 2, 8, 2>(ctx, p_begin, p_end, editdist, d_out, d_mask)
                              : launch_rows_t<8, 2, 8, 2>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  switch (ctx->rows_cfg) {  // tuning instantiations (scripts/rows_probe.py)
    case 1: return launch_rows_t<13, 13, 112, 6>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 2: return launch_rows_t<13, 12, 240, 6>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 3: return launch_rows_t<13, 13, 112, 4>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 4: return launch_rows_pair_t<13, 13, 112, 3>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 5: return launch_rows_pair_t<13, 12, 240, 3>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 6: return launch_rows_pair_t<13, 13, 112, 2>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    case 7: return launch_rows_pair_t<13, 13, 112, 4>(ctx, p_begin, p_end, editdist, d_out, d_mask);
    default: break;
  }
  // 13 warps per SM when every row's lists fit one pass of 112 (measured 2.53 vs 2.60 ms at 100k x 200), else 12 warps
  // x 240; 6 groups per lane in flight (4: 2.57 ms, 8: 2.54 ms)
  if (ctx->h.max_lists <= 112) return launch_rows_t<13, 13, 112, 6>(ctx, p_begin, p_end, editdist, d_out, d_mask);
  return launch_rows_pair_t<13, 12, 240, 3>(ctx, p_begin, p_end, editdist, d_out, d_mask);
}

}  // namespace prf
