// prf_primitives.cuh -- device-wide exclusive scan and stable LSD radix sort of (key, value) pairs.
// Hand-written (no CUB/Thrust): single-pass scan with decoupled look-back and an 8-bit-digit radix sort (tile histogram, scan of digit-major histograms, stable
// warp-multisplit scatter).  Both are HBM-streaming kernels far off the critical path: at the
// headline shape they touch < 1 % of the bytes the matrix kernel writes.
#pragma once

#include "prf_common.cuh"

namespace prf {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix and writes the
// block total to `total` (same value in every thread).
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= (uint32_t)d) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  uint32_t wprefix = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    uint32_t s = warp_sums[w];
    if ((uint32_t)w < warp) wprefix += s;
    tot += s;
  }
  __syncthreads();  // warp_sums may be reused by the caller's next call
  total = tot;
  return wprefix + inc - v;
}

// F: struct with __device__ uint32_t operator()(uint64_t i) const

// ---- single-pass scan (decoupled look-back) -----------------------------------------------------------
// One kernel instead of three: a tile publishes its aggregate, then sums the published values of the tiles
// before it until it meets one that already knows its inclusive prefix.  Tile ids are handed out by an atomic
// counter, so a tile only ever waits for tiles that are already running.  status word = generation << 34 |
// flag << 32 | value (flag 1 = aggregate, 2 = inclusive prefix): the generation (one per scan call of the
// context) makes stale words read as "not ready", so the status array never needs clearing.
template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_lookback_kernel(F f, uint64_t n_max, const uint32_t* n_dev, uint32_t* out,
                                                                      uint32_t* total, volatile unsigned long long* status,
                                                                      uint32_t* tile_counter, uint32_t gen, uint32_t tiles_max) {
  __shared__ uint32_t s_tile, s_prefix;
  if (threadIdx.x == 0) {
    s_tile = atomicAdd(tile_counter, 1u);
    if (s_tile == tiles_max - 1) *tile_counter = 0u;  // every id has been handed out: ready for the next scan
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  // n_dev (optional): the number of elements, known only on the device (<= n_max); tiles beyond it leave at once
  const uint64_t n = n_dev ? min((uint64_t)*n_dev, n_max) : n_max;
  const uint32_t tiles = n ? (uint32_t)((n + kScanTile - 1) / kScanTile) : 1u;
  if (tile >= tiles) return;
  const uint64_t base = (uint64_t)tile * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + k;
    v[k] = i < n ? f(i) : 0;
    s += v[k];
  }
  uint32_t tot;
  uint32_t ex = block_exclusive_scan(s, tot);
  if (threadIdx.x == 0) {
    const unsigned long long g = (unsigned long long)gen << 34;
    uint32_t prefix = 0;
    if (tile == 0) {
      status[0] = g | (2ull << 32) | tot;
    } else {
      status[tile] = g | (1ull << 32) | tot;
      __threadfence();
      for (uint32_t p = tile; p-- > 0;) {
        unsigned long long w;
        do { w = status[p]; } while ((w >> 34) != gen || ((w >> 32) & 3ull) == 0);
        prefix += (uint32_t)w;
        if (((w >> 32) & 3ull) == 2) break;
      }
      status[tile] = g | (2ull << 32) | (uint32_t)(prefix + tot);
    }
    s_prefix = prefix;
    if (tile == tiles - 1 && total) *total = prefix + tot;
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const uint64_t i = base + k;
    if (i < n) out[i] = ex;
    ex += v[k];
  }
}

// out[i] = sum_{j<i} f(j) for i in [0, n); *d_total = sum of all (device pointer, may be null).
// out must not alias anything f reads unless f reads element i only (in-place is then safe:
// every tile is fully read into registers before it is written).
template <class F>
int exclusive_scan(Ctx* ctx, F f, uint64_t n, uint32_t* out, uint32_t* d_total, const uint32_t* n_dev = nullptr) {
  cudaStream_t s = ctx->stream;
  if (n == 0) {
    if (d_total) PRF_CK(ctx, cudaMemsetAsync(d_total, 0, sizeof(uint32_t), s));
    return PRF_OK;
  }
  const uint32_t tiles = (uint32_t)((n + kScanTile - 1) / kScanTile);
  if (ctx->scan_status.n < (size_t)tiles + 1) {
    PRF_CK(ctx, ctx->scan_status.alloc(std::max<size_t>((size_t)tiles + 1, 4096), s));
    PRF_CK(ctx, cudaMemsetAsync(ctx->scan_status.p, 0, ctx->scan_status.bytes(), s));
    ctx->scan_gen = 0;
  }
  if (!ctx->scan_counter.p) {
    PRF_CK(ctx, ctx->scan_counter.alloc(1, s));
    PRF_CK(ctx, cudaMemsetAsync(ctx->scan_counter.p, 0, sizeof(uint32_t), s));
  }
  if (++ctx->scan_gen >= (1u << 30)) {  // generation wrap: start over with a clean status array
    PRF_CK(ctx, cudaMemsetAsync(ctx->scan_status.p, 0, ctx->scan_status.bytes(), s));
    ctx->scan_gen = 1;
  }
  scan_lookback_kernel<F><<<tiles, kScanThreads, 0, s>>>(f, n, n_dev, out, d_total, ctx->scan_status.p, ctx->scan_counter.p,
                                                         ctx->scan_gen, tiles);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

// nb32 -> uint16 + min/max.  mm[0] = min, mm[1] = max.
// Also mm[2] = most suffix ranges of one tree (roff is the exclusive scan of the per-tree counts).
static __global__ void __launch_bounds__(256) finish_nb_kernel(const uint32_t* __restrict__ nb32, uint32_t T,
                                                        uint16_t* __restrict__ nb16, uint32_t* mm,
                                                        const uint32_t* __restrict__ roff) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = t < T ? nb32[t] : 0, lo = t < T ? v : 0xffffffffu, hi = v;
  uint32_t rmax = t < T ? roff[t + 1] - roff[t] : 0;
  if (t < T) nb16[t] = (uint16_t)(v > 0xffffu ? 0xffffu : v);
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, d));
  }
  if (lane_id() == 0) {
    atomicMin(&mm[0], lo);
    atomicMax(&mm[1], hi);
    atomicMax(&mm[2], rmax);
  }
}

struct LoadU32 {
  const uint32_t* p;
  __device__ uint32_t operator()(uint64_t i) const { return p[i]; }
};

// ---------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, value) uint32 pairs on key bits [0, nbits).
// ---------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;  // per thread
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096: 8 warps x 16 rounds x 32 lanes

template <int kBits>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t* __restrict__ keys,
                                                                  uint32_t n, int shift,
                                                                  uint32_t* __restrict__ hist,
                                                                  uint32_t tiles) {
  constexpr uint32_t kBins = 1u << kBits;
  __shared__ uint32_t h[kBins];
  for (uint32_t d = threadIdx.x; d < kBins; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    uint32_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kBins - 1)], 1u);
  }
  __syncthreads();
  for (uint32_t d = threadIdx.x; d < kBins; d += kSortThreads) hist[d * tiles + blockIdx.x] = h[d];  // digit-major
}

template <int kBits>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(
    const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t n,
    int shift, const uint32_t* __restrict__ hist_scanned, uint32_t tiles,
    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr uint32_t kBins = 1u << kBits;
  __shared__ uint32_t whist[kSortThreads / 32][kBins];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (kSortThreads / 32) * kBins; i += kSortThreads) (&whist[0][0])[i] = 0;
  __syncthreads();
  // warp w owns the contiguous sub-tile [base + w*512, base + (w+1)*512), 16 rounds of 32
  const uint32_t wbase = blockIdx.x * kSortTile + warp * (kSortItems * 32);
  uint32_t k[kSortItems], v[kSortItems];
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    uint32_t i = wbase + r * 32 + lane;
    bool valid = i < n;
    k[r] = valid ? keys_in[i] : 0;
    v[r] = valid ? vals_in[i] : 0;
    uint32_t act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      uint32_t d = (k[r] >> shift) & (kBins - 1);
      uint32_t peers = __match_any_sync(act, d);
      if ((peers & lt) == 0) whist[warp][d] += __popc(peers);  // lowest lane of the group
    }
    __syncwarp();
  }
  __syncthreads();
  // thread d: turn per-warp counts of digit d into global start offsets
  for (uint32_t d = threadIdx.x; d < kBins; d += kSortThreads) {
    uint32_t run = hist_scanned[d * tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) {
      uint32_t t = whist[w][d];
      whist[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    uint32_t i = wbase + r * 32 + lane;
    bool valid = i < n;
    uint32_t act = __ballot_sync(0xffffffffu, valid);
    uint32_t d = 0, peers = 0;
    if (valid) {
      d = (k[r] >> shift) & (kBins - 1);
      peers = __match_any_sync(act, d);
      uint32_t pos = whist[warp][d] + __popc(peers & lt);
      keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
    __syncwarp();  // everyone has read whist before the group leaders bump it
    if (valid && (peers & lt) == 0) whist[warp][d] += __popc(peers);
    __syncwarp();
  }
}

template <int kBits>
inline int radix_sort_passes(Ctx* ctx, uint32_t*& ki, uint32_t*& vi, uint32_t*& ko, uint32_t*& vo, uint32_t n, int nbits,
                             uint32_t* hist, uint32_t tiles) {
  cudaStream_t s = ctx->stream;
  for (int shift = 0; shift < nbits; shift += kBits) {
    sort_hist_kernel<kBits><<<tiles, kSortThreads, 0, s>>>(ki, n, shift, hist, tiles);
    PRF_LAUNCHED(ctx);
    int rc = exclusive_scan(ctx, LoadU32{hist}, (uint64_t)(1u << kBits) * tiles, hist, nullptr);
    if (rc) return rc;
    sort_scatter_kernel<kBits><<<tiles, kSortThreads, 0, s>>>(ki, vi, n, shift, hist, tiles, ko, vo);
    PRF_LAUNCHED(ctx);
    uint32_t* t;
    t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  return PRF_OK;
}

// Sorts in place (result ends in keys/vals); tmp buffers of the same size are allocated inside.
// 8-bit digits for short keys, 10-bit digits (1024 shared-memory bins per warp) when that saves a pass.
inline int radix_sort_pairs(Ctx* ctx, uint32_t* keys, uint32_t* vals, uint32_t n, int nbits) {
  if (n < 2 || nbits <= 0) return PRF_OK;
  cudaStream_t s = ctx->stream;
  uint32_t tiles = (n + kSortTile - 1) / kSortTile;
  const bool wide = (nbits + 9) / 10 < (nbits + 7) / 8;
  DevBuf<uint32_t> k2, v2, hist;
  PRF_CK(ctx, k2.alloc(n, s));
  PRF_CK(ctx, v2.alloc(n, s));
  PRF_CK(ctx, hist.alloc((size_t)(wide ? 1024 : 256) * tiles, s));
  uint32_t *ki = keys, *vi = vals, *ko = k2.p, *vo = v2.p;
  int rc = wide ? radix_sort_passes<10>(ctx, ki, vi, ko, vo, n, nbits, hist.p, tiles)
                : radix_sort_passes<8>(ctx, ki, vi, ko, vo, n, nbits, hist.p, tiles);
  if (rc) return rc;
  if (ki != keys) {
    PRF_CK(ctx, cudaMemcpyAsync(keys, ki, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    PRF_CK(ctx, cudaMemcpyAsync(vals, vi, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  }
  return PRF_OK;
}

}  // namespace prf
