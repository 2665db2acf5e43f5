// prf_sparse.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), sparse variant.
//
// The reference bumps d(i,j) once for every bipartition that exactly one of the two trees has
// (haveIt x dontHave, :327-338): d(a,b) = |B_a| + |B_b| - 2 |B_a n B_b|.  Random topologies share
// ~0.13 bipartitions per pair, so almost every entry is just |B_a| + |B_b| and the work is writing
// the matrix: 2 bytes per pair, HBM-write bound.
//
// One CTA produces one CHUNK of the packed upper triangle -- a contiguous run of kChunk uint16
// (rows of the packed layout follow each other without gaps, so a chunk is 1..many row segments):
//   1. fill the chunk in shared memory with |B_a| + |B_b|;
//   2. for every tree a that has a segment in the chunk, walk the suffixes of its shared
//      bipartitions' inverted lists ("trees b > a that also have it", ascending) and subtract 2 at
//      column b with a shared-memory atomic -- no global atomics, every output byte is written once.
//      Lists are entered at the chunk's first column: one thread per list estimates the position by
//      interpolation (tree ids are close to uniform) and gallops back until it is provably at or
//      before the lower bound, all lists in parallel, so the dependent-load chain is ~2 deep;
//   3. optionally ballot the d <= editdist mask, then hand the finished chunk to the TMA engine
//      (cp.async.bulk shared -> global): one 16-byte-aligned bulk store per chunk.
#pragma once

#include "prf_common.cuh"

namespace prf {

constexpr int kSparseThreads = 256;

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

struct StagedRange {
  uint32_t begin, end;  // walk members[begin, end)
  int32_t base;         // tile index of column b is base + b
  uint32_t b_lo;        // only columns in [b_lo, b_hi) belong to this chunk's segment of the row
  uint32_t b_hi;
};

template <int kChunk>
__global__ void __launch_bounds__(kSparseThreads) sparse_chunk_kernel(
    uint32_t T, uint64_t p_begin, uint64_t p_end, const uint16_t* __restrict__ nb,
    int uniform_nb, uint32_t nb_uniform, const uint32_t* __restrict__ roff,
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ members, int32_t editdist,
    uint16_t* __restrict__ out, uint32_t* __restrict__ mask) {
  extern __shared__ __align__(128) uint16_t tile[];  // kChunk entries
  __shared__ StagedRange staged[kSparseThreads];
  uint32_t* tile32 = reinterpret_cast<uint32_t*>(tile);

  const uint64_t p0 = p_begin + (uint64_t)blockIdx.x * kChunk;
  const uint32_t n = (uint32_t)min((uint64_t)kChunk, p_end - p0);
  const uint32_t a0 = row_of(p0, T), a1 = row_of(p0 + n - 1, T);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t k0 = __ldg(roff + a0), k1 = __ldg(roff + a1 + 1);
  const uint32_t b_first = a0 + 1 + (uint32_t)(p0 - row_start(a0, T));          // first column of row a0 in the chunk
  const uint32_t b_last_excl = a1 + 1 + (uint32_t)(p0 + n - row_start(a1, T));  // one past the last column of row a1

  for (uint32_t kb = k0; kb < k1 || kb == k0; kb += kSparseThreads) {
    // ---- 2a. stage up to 256 ranges: one thread each ----
    const uint32_t k = kb + threadIdx.x;
    if (k < k1) {
      // the row that owns range k: largest a in [a0, a1] with roff[a] <= k
      uint32_t lo = a0, hi = a1;
      while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(roff + mid) <= k) lo = mid; else hi = mid - 1;
      }
      const uint32_t a = lo;
      const uint2 r = ranges[k];
      const uint32_t b_lo = a == a0 ? b_first : a + 1;
      const uint32_t b_hi = a == a1 ? b_last_excl : T;
      uint32_t g = r.x;
      if (b_lo > a + 1) {
        // members of (a, T) are near-uniform: interpolate, then gallop back until members[g-1] < b_lo
        const uint32_t len = r.y - r.x;
        uint32_t est = r.x + (uint32_t)(((uint64_t)(b_lo - a - 1) * len) / (uint64_t)(T - a - 1));
        g = est > r.x + 16 ? est - 16 : r.x;
        uint32_t step = 32;
        while (g > r.x && __ldg(members + g - 1) >= b_lo) {
          g = g - r.x > step ? g - step : r.x;
          step <<= 1;
        }
      }
      StagedRange s;
      s.begin = g;
      s.end = r.y;
      s.base = (int32_t)((int64_t)row_start(a, T) - (int64_t)p0 - (int64_t)a - 1);
      s.b_lo = b_lo;
      s.b_hi = b_hi;
      staged[threadIdx.x] = s;
    }
    if (kb == k0) {
      // ---- 1. d = |B_a| + |B_b| (first pass only; overlaps the staging loads of other warps) ----
      if (uniform_nb) {
        const uint32_t v = 2u * nb_uniform;
        const uint32_t vv = v | (v << 16);
        uint4 q = make_uint4(vv, vv, vv, vv);
        uint4* t4 = reinterpret_cast<uint4*>(tile);
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < kChunk / 8; i += kSparseThreads) t4[i] = q;
      } else {
        for (uint32_t a = a0; a <= a1; ++a) {
          const uint64_t rs = row_start(a, T);
          const uint64_t lo = max(rs, p0), hi = min(rs + (T - 1 - a), p0 + n);
          const uint32_t na = nb[a];
          const uint32_t cnt = (uint32_t)(hi - lo);
          const uint32_t toff = (uint32_t)(lo - p0);
          const uint16_t* nbb = nb + (a + 1 + (lo - rs));  // column of position lo
          for (uint32_t i = threadIdx.x; i < cnt; i += kSparseThreads)
            tile[toff + i] = (uint16_t)(na + nbb[i]);
        }
      }
    }
    __syncthreads();
    // ---- 2b. subtract 2 per shared bipartition: one warp per staged range ----
    const uint32_t nstaged = min((uint32_t)kSparseThreads, k1 - kb);
    for (uint32_t i = warp; i < nstaged; i += kSparseThreads / 32) {
      const StagedRange s = staged[i];
      for (uint32_t m = s.begin + lane; m < s.end; m += 32) {
        const uint32_t b = __ldg(members + m);
        if (b >= s.b_hi) break;
        if (b >= s.b_lo) {
          const uint32_t idx = (uint32_t)(s.base + (int32_t)b);
          atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
        }
      }
    }
    if (kb + kSparseThreads < k1) __syncthreads();  // staged[] is reused by the next batch
  }
  // make the generic-proxy writes visible to the async proxy before the bulk store reads them
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // ---- 3. mask + store ----
  const uint64_t rel = p0 - p_begin;
  if (editdist >= 0) {
    const uint32_t groups = (n + 31) / 32;
    uint32_t* mrow = mask + rel / 32;
    for (uint32_t g = warp; g < groups; g += kSparseThreads / 32) {
      const uint32_t i = g * 32 + lane;
      const bool hit = i < n && (int32_t)tile[i] <= editdist;
      const uint32_t bits = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) mrow[g] = bits;
    }
  }
  uint16_t* dst = out + rel;
  if (n == (uint32_t)kChunk) {
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                   "r"(smem_addr(tile)), "r"((uint32_t)(kChunk * 2))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    for (uint32_t i = threadIdx.x; i < n; i += kSparseThreads) dst[i] = tile[i];
  }
}

template <int kChunk>
static int launch_sparse(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                         uint32_t* d_mask) {
  const HashRFState& H = ctx->h;
  PRF_CK(ctx, cudaFuncSetAttribute(sparse_chunk_kernel<kChunk>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kChunk * 2));
  uint64_t chunks = (p_end - p_begin + kChunk - 1) / kChunk;
  if (chunks == 0) return PRF_OK;
  if (chunks > 0x7fffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "range too large for one launch");
  sparse_chunk_kernel<kChunk><<<(uint32_t)chunks, kSparseThreads, kChunk * 2, ctx->stream>>>(
      H.T, p_begin, p_end, H.nb.p, H.uniform_nb ? 1 : 0, H.nb_uniform, H.roff.p, H.ranges.p, H.members.p,
      editdist, d_out, d_mask);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
