// prf_agglo.cuh -- the agglomeration itself on the device: C.dendrogram linkage trees dist
// (PhyBin.hs:358) for one set of trees, as the full merge list (dendrogram.txt, PhyBin.hs:238-243) or cut at
// a threshold (sliceDendro, PhyBin.hs:415-422).
//
// Semantics are those of the host restatement phb_cluster_component (phybin_host.cpp, `agglomerate`), which
// follows hierarchical-clustering-0.4.6's distance-matrix algorithm: every step merges the pair of clusters
// at minimal distance, the first in (lower key, higher key) lexicographic order among ties; the merged
// cluster keeps the lower key; distances follow Lance-Williams (min / max / size-weighted mean in IEEE
// double, no fused multiply-add).  That order is inherently sequential (m - 1 dependent steps), so the
// kernel is ONE persistent cooperative launch whose step costs one grid barrier:
//
//   * the m x m matrix lives in HBM/L2 as a full symmetric array (uint16 for single / complete linkage,
//     whose values stay input values; double for UPGMA); row i belongs to CTA i mod G, which keeps the
//     row's nearest later neighbour (nn, nnd: first minimal column j > i) in shared memory;
//   * per step every CTA (1) reads the G published records and derives the same winner (a, b),
//     (2) updates D[x][a] / D[a][x] for its own rows x, kills column b with a sentinel, (3) repairs nn of
//     its own rows (a rescan of a row only reads that row, which only its owner ever writes -- except row
//     a, whose new nearest neighbour is reduced from the values the CTAs just computed and published with
//     the next record), (4) publishes its best row;
//   * rows are read with ld.global.cg (L2): no stale L1 lines of a row another SM has written.
#pragma once

#include <cooperative_groups.h>

#include "prf_common.cuh"

namespace prf {

constexpr uint32_t kAggloThreads = 512;
constexpr uint32_t kAggloNone = 0xffffffffu;

struct AggloRecord {   // what a CTA publishes per step
  double best_d;       // its best own row: nnd, row, nn   (row a of the step excluded)
  uint32_t best_i, best_j;
  double rowa_d;       // its share of row a's new nearest neighbour: min over its rows x > a of the new D[a][x]
  uint32_t rowa_j;
  uint32_t stamp;      // step + 1 (debugging aid)
};

struct AggloParams {
  void* D;             // m x ld
  uint32_t m, ld;
  int linkage;         // 0 single, 1 complete, 2 UPGMA
  double thresh;
  uint32_t* sizes;     // G copies of the cluster sizes (every CTA updates its own)
  AggloRecord* rec;    // 2 x G
  uint32_t* merge_a;   // m - 1
  uint32_t* merge_b;
  double* merge_d;
  uint32_t* n_merges;
};

template <class Tv> struct AggloVal;
template <> struct AggloVal<uint16_t> {
  static constexpr uint32_t kPerVec = 8;
  __device__ static double to_double(uint16_t v) { return v == 0xffffu ? __longlong_as_double(0x7ff0000000000000ll) : (double)v; }
  __device__ static uint16_t from_double(double d) { return (uint16_t)d; }
  __device__ static uint16_t dead() { return 0xffffu; }
};
template <> struct AggloVal<double> {
  static constexpr uint32_t kPerVec = 2;
  __device__ static double to_double(double v) { return v; }
  __device__ static double from_double(double d) { return d; }
  __device__ static double dead() { return __longlong_as_double(0x7ff0000000000000ll); }
};

__device__ __forceinline__ bool agglo_less(double d1, uint32_t i1, double d2, uint32_t i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

// min by (d, i) over the CTA, payload j; every thread returns the winner.  scratch: 3 x 32 words of shared memory.
__device__ __forceinline__ void agglo_block_min(double& d, uint32_t& i, uint32_t& j, double* s_d, uint32_t* s_i, uint32_t* s_j) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
#pragma unroll
  for (int sh = 16; sh; sh >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, sh);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i, sh), oj = __shfl_xor_sync(0xffffffffu, j, sh);
    if (agglo_less(od, oi, d, i)) { d = od; i = oi; j = oj; }
  }
  __syncthreads();  // the scratch of the previous call has been read
  if (lane == 0) { s_d[warp] = d; s_i[warp] = i; s_j[warp] = j; }
  __syncthreads();
  d = lane < n_warps ? s_d[lane] : __longlong_as_double(0x7ff0000000000000ll);
  i = lane < n_warps ? s_i[lane] : kAggloNone;
  j = lane < n_warps ? s_j[lane] : kAggloNone;
#pragma unroll
  for (int sh = 16; sh; sh >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, sh);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i, sh), oj = __shfl_xor_sync(0xffffffffu, j, sh);
    if (agglo_less(od, oi, d, i)) { d = od; i = oi; j = oj; }
  }
}

// Two reductions at once: (d1, i1, payload j1) and (d2, i2), each min by (d, i).  scratch: the arrays of agglo_block_min
// plus s_d2 / s_i2 (32 entries each).
__device__ __forceinline__ void agglo_block_min2(double& d1, uint32_t& i1, uint32_t& j1, double& d2, uint32_t& i2, double* s_d,
                                                 uint32_t* s_i, uint32_t* s_j, double* s_d2, uint32_t* s_i2) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
#pragma unroll
  for (int sh = 16; sh; sh >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d1, sh), pd = __shfl_xor_sync(0xffffffffu, d2, sh);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i1, sh), oj = __shfl_xor_sync(0xffffffffu, j1, sh);
    const uint32_t pi = __shfl_xor_sync(0xffffffffu, i2, sh);
    if (agglo_less(od, oi, d1, i1)) { d1 = od; i1 = oi; j1 = oj; }
    if (agglo_less(pd, pi, d2, i2)) { d2 = pd; i2 = pi; }
  }
  __syncthreads();  // the scratch of the previous call has been read
  if (lane == 0) { s_d[warp] = d1; s_i[warp] = i1; s_j[warp] = j1; s_d2[warp] = d2; s_i2[warp] = i2; }
  __syncthreads();
  d1 = lane < n_warps ? s_d[lane] : inf;
  i1 = lane < n_warps ? s_i[lane] : kAggloNone;
  j1 = lane < n_warps ? s_j[lane] : kAggloNone;
  d2 = lane < n_warps ? s_d2[lane] : inf;
  i2 = lane < n_warps ? s_i2[lane] : kAggloNone;
#pragma unroll
  for (int sh = 16; sh; sh >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d1, sh), pd = __shfl_xor_sync(0xffffffffu, d2, sh);
    const uint32_t oi = __shfl_xor_sync(0xffffffffu, i1, sh), oj = __shfl_xor_sync(0xffffffffu, j1, sh);
    const uint32_t pi = __shfl_xor_sync(0xffffffffu, i2, sh);
    if (agglo_less(od, oi, d1, i1)) { d1 = od; i1 = oi; j1 = oj; }
    if (agglo_less(pd, pi, d2, i2)) { d2 = pd; i2 = pi; }
  }
}

// First minimal entry of row x among the columns j > x (dead columns hold the sentinel = +inf): the whole CTA
// reads the row with 16-byte L2 loads; returns (value, column) in every thread, (+inf, NONE) when there is none.
template <class Tv>
__device__ __forceinline__ void agglo_rescan(const Tv* D, uint32_t ld, uint32_t m, uint32_t x, double& out_d, uint32_t& out_j,
                                             double* s_d, uint32_t* s_i, uint32_t* s_j) {
  constexpr uint32_t E = AggloVal<Tv>::kPerVec;
  const Tv* row = D + (size_t)x * ld;
  double bd = __longlong_as_double(0x7ff0000000000000ll);
  uint32_t bj = kAggloNone;
  for (uint32_t c = (x + 1) / E + threadIdx.x; c * E < m; c += blockDim.x) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(row + (size_t)c * E));
    Tv e[E];
    memcpy(e, &v, 16);
#pragma unroll
    for (uint32_t q = 0; q < E; ++q) {
      const uint32_t j = c * E + q;
      const double d = AggloVal<Tv>::to_double(e[q]);
      if (j > x && j < m && d < bd) { bd = d; bj = j; }  // ascending j inside a thread: the first minimum stays
    }
  }
  uint32_t dummy = 0;
  agglo_block_min(bd, bj, dummy, s_d, s_i, s_j);
  out_d = bd;
  out_j = bj;
}

template <class Tv>
__global__ void __launch_bounds__(kAggloThreads, 1) agglo_kernel(const AggloParams prm) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const uint32_t m = prm.m, ld = prm.ld;
  const uint32_t R = (m + G - 1) / G;                      // rows per CTA (row x = r * G + cta)
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double* s_nnd = reinterpret_cast<double*>(smem_raw);      // [R]
  double* s_d = s_nnd + R;                                  // [32] reduction scratch
  uint32_t* s_nn = reinterpret_cast<uint32_t*>(s_d + 32);   // [R]
  uint32_t* s_i = s_nn + R;                                 // [32]
  uint32_t* s_j = s_i + 32;                                 // [32]
  uint32_t* s_list = s_j + 32;                              // [R] rows to rescan
  uint32_t* s_act = s_list + R;                             // [R] row still a cluster
  uint32_t* s_i2 = s_act + R;                               // [32]
  double* s_d2 = reinterpret_cast<double*>(s_i2 + 32 + ((R & 1) ? 1 : 0));  // [32], 8-byte aligned: 4 R words + 32 words (+1)
  __shared__ uint32_t s_n_list;
  Tv* D = reinterpret_cast<Tv*>(prm.D);
  uint32_t* size = prm.sizes + (size_t)cta * m;

  for (uint32_t x = tid; x < m; x += blockDim.x) size[x] = 1;
  for (uint32_t r = 0; r < R; ++r) {  // initial nearest neighbours of the own rows
    const uint32_t x = r * G + cta;
    double d = inf;
    uint32_t j = kAggloNone;
    if (x < m) agglo_rescan<Tv>(D, ld, m, x, d, j, s_d, s_i, s_j);  // (x < m is uniform in the CTA)
    if (tid == 0) { s_nnd[r] = d; s_nn[r] = j; s_act[r] = x < m; }
  }
  __syncthreads();

  uint32_t a_prev = kAggloNone;   // row merged into in the previous step: its nn arrives with the records
  double pa_d = inf;              // this CTA's share of it
  uint32_t pa_j = kAggloNone;
  for (uint32_t step = 0;; ++step) {
    // ---- publish: best own row (without a_prev) and the share of a_prev's row
    {
      double d = inf;
      uint32_t i = kAggloNone, j = kAggloNone;
      for (uint32_t r = tid; r < R; r += blockDim.x) {
        const uint32_t x = r * G + cta;
        if (s_act[r] && x != a_prev && s_nn[r] != kAggloNone && agglo_less(s_nnd[r], x, d, i)) { d = s_nnd[r]; i = x; j = s_nn[r]; }
      }
      agglo_block_min2(d, i, j, pa_d, pa_j, s_d, s_i, s_j, s_d2, s_i2);
      if (tid == 0) {
        AggloRecord rc;
        rc.best_d = d; rc.best_i = i; rc.best_j = j;
        rc.rowa_d = pa_d; rc.rowa_j = pa_j; rc.stamp = step + 1;
        prm.rec[(size_t)(step & 1) * G + cta] = rc;
      }
    }
    grid.sync();
    // ---- every CTA derives the same winner from the G records
    double bd = inf, ad = inf;
    uint32_t bi = kAggloNone, bj = kAggloNone, aj = kAggloNone;
    for (uint32_t c = tid; c < G; c += blockDim.x) {
      const uint4* rp = reinterpret_cast<const uint4*>(prm.rec + (size_t)(step & 1) * G + c);
      const uint4 lo = __ldcg(rp), hi = __ldcg(rp + 1);
      const double d1 = __longlong_as_double(((long long)lo.y << 32) | lo.x), d2 = __longlong_as_double(((long long)hi.y << 32) | hi.x);
      const uint32_t i1 = lo.z, j1 = lo.w, j2 = hi.z;
      if (agglo_less(d1, i1, bd, bi)) { bd = d1; bi = i1; bj = j1; }
      if (agglo_less(d2, j2, ad, aj)) { ad = d2; aj = j2; }
    }
    agglo_block_min2(bd, bi, bj, ad, aj, s_d, s_i, s_j, s_d2, s_i2);
    if (a_prev != kAggloNone) {
      if (a_prev % G == cta && tid == 0) { s_nnd[a_prev / G] = ad; s_nn[a_prev / G] = aj; }
      if (aj != kAggloNone && agglo_less(ad, a_prev, bd, bi)) { bd = ad; bi = a_prev; bj = aj; }
    }
    if (bi == kAggloNone || bd > prm.thresh) {   // sliceDendro: dist > dstThresh splits (PhyBin.hs:418-419)
      if (cta == 0 && tid == 0) *prm.n_merges = step;
      break;
    }
    const uint32_t a = bi, b = bj;   // a < b
    if (cta == 0 && tid == 0) { prm.merge_a[step] = a; prm.merge_b[step] = b; prm.merge_d[step] = bd; }
    // cluster sizes only enter UPGMA's weights
    const double na = prm.linkage == 2 ? (double)size[a] : 1.0, nb = prm.linkage == 2 ? (double)size[b] : 1.0;
    if (tid == 0) s_n_list = 0;
    __syncthreads();   // sizes read, list counter cleared, a_prev's nn stored
    if (tid == 0 && prm.linkage == 2) size[a] += size[b];
    // ---- update the own rows
    pa_d = inf;
    pa_j = kAggloNone;
    for (uint32_t r = tid; r < R; r += blockDim.x) {
      const uint32_t x = r * G + cta;
      if (!s_act[r]) continue;
      if (x == b) {
        s_act[r] = 0; s_nnd[r] = inf; s_nn[r] = kAggloNone;
        __stcg(&D[(size_t)a * ld + b], AggloVal<Tv>::dead());
        continue;
      }
      if (x == a) continue;
      Tv* rowx = D + (size_t)x * ld;
      const double da = AggloVal<Tv>::to_double(__ldcg(rowx + a)), db = AggloVal<Tv>::to_double(__ldcg(rowx + b));
      double v;
      if (prm.linkage == 0) v = fmin(da, db);
      else if (prm.linkage == 1) v = fmax(da, db);
      else v = __ddiv_rn(__dadd_rn(__dmul_rn(na, da), __dmul_rn(nb, db)), __dadd_rn(na, nb));
      const Tv vv = AggloVal<Tv>::from_double(v);
      __stcg(rowx + a, vv);
      __stcg(D + (size_t)a * ld + x, vv);
      __stcg(rowx + b, AggloVal<Tv>::dead());
      if (x > a && agglo_less(v, x, pa_d, pa_j)) { pa_d = v; pa_j = x; }
      if (x < b) {
        if (s_nn[r] == a || s_nn[r] == b) {
          s_list[atomicAdd(&s_n_list, 1u)] = r;
        } else if (x < a && (v < s_nnd[r] || (v == s_nnd[r] && a < s_nn[r]))) {
          s_nnd[r] = v; s_nn[r] = a;
        }
      }
    }
    // (pa_d, pa_j: this thread's share of row a's nearest neighbour, reduced and published with the next record)
    __syncthreads();   // list complete, own-row stores ordered before the rescans of this CTA
    const uint32_t n_list = s_n_list;
    for (uint32_t q = 0; q < n_list; ++q) {
      const uint32_t r = s_list[q];
      double d;
      uint32_t j;
      agglo_rescan<Tv>(D, ld, m, r * G + cta, d, j, s_d, s_i, s_j);
      if (tid == 0) { s_nnd[r] = d; s_nn[r] = j; }
    }
    __syncthreads();
    a_prev = a;
  }
}

// D[i][j] = d(ids[i], ids[j]) (ids ascending, or NULL for the identity) from the packed triangle `src` that starts at
// packed position src_begin; the diagonal and the padding columns get the sentinel.
template <class Tv>
__global__ void __launch_bounds__(256) agglo_gather_kernel(const uint16_t* __restrict__ src, uint64_t src_begin, uint32_t T,
                                                           const uint32_t* __restrict__ ids, uint32_t m, uint32_t ld,
                                                           Tv* __restrict__ D) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ld) return;
  for (uint32_t i = blockIdx.y; i < m; i += gridDim.y) {
    Tv v = AggloVal<Tv>::dead();
    if (j < m && j != i) {
      const uint32_t ti = ids ? ids[i] : i, tj = ids ? ids[j] : j;
      const uint32_t lo = min(ti, tj), hi = max(ti, tj);
      v = (Tv)src[row_start(lo, T) + (hi - lo - 1) - src_begin];
    }
    D[(size_t)i * ld + j] = v;
  }
}

}  // namespace prf
