// prf_dense.cuh -- the `ingest` phase of hashRF (RFDistance.hs:300-341), dense variant.
//
// When the trees are similar (small split universe: e.g. one base tree plus a few NNI moves each)
// almost every pair shares almost every bipartition and the inverted-list join of prf_sparse.cuh
// degenerates to ~|B| hits per pair.  Then the shared-split count IS a dense contraction:
//     S = A * A^T,   A[t][u] = 1 iff tree t has shared bipartition u      (T x U_s, 0/1 as uint8)
//     d(a,b) = |B_a| + |B_b| - 2 S[a][b]
// computed on the 5th-generation tensor cores: tcgen05.mma kind::i8 (u8 x u8 -> s32), accumulators in
// TMEM, operands staged in shared memory by the bulk-copy (TMA) engine.
//
// Operand layout.  A is stored ONCE in global memory already in the tensor core's canonical K-major
// "no swizzle" shared-memory order: 8-row x 16-byte core matrices (128 contiguous bytes), ordered
//     [k-block of 128 columns][group of 8 trees][16-byte k-chunk (8)][tree in group (8)][16 bytes]
// so the operand tile of ANY run of consecutive tree groups for one k-block is one contiguous byte
// range: a single cp.async.bulk brings a 128-tree (16 KB) or 256-tree (32 KB) tile into shared
// memory, no tensor map and no swizzle needed, and both MMA operands (row block and column block)
// are tiles of the same array.  UMMA descriptors: LBO (next core matrix along K) = 128 B,
// SBO (next 8-tree group) = 1024 B; one instruction consumes K = 32 bytes = 2 core matrices.
//
// One CTA (256 threads) per SM, persistent over 128 x 256 output tiles of the upper triangle:
//   warp 0   one lane: bulk loads into a 3-stage ring (mbarrier complete_tx)
//   warp 1   one lane: tcgen05.mma issue, tcgen05.commit frees the stage / publishes the accumulator
//   warp 2   TMEM allocation (512 columns = two 128 x 256 s32 accumulators, double buffered)
//   warps 4-7 epilogue: tcgen05.ld the accumulator (thread = tree a, 32 columns at a time), fuse
//            d = |B_a| + |B_b| - 2S -> uint16 into a per-warp staging tile, release the accumulator,
//            then write every row's run of the packed upper triangle with 16-byte coalesced stores
//            (funnel-shifted to the run's global alignment), plus the optional d <= editdist mask.
#pragma once

#include <algorithm>

#include "prf_common.cuh"
#include "prf_async.cuh"
#include "prf_internal.h"

namespace prf {

constexpr int kDM = 128, kDN = 256, kDK = 128;  // output tile (trees x trees) and k-block (bytes = columns)
constexpr int kDStages = 3;
constexpr int kDStageBytes = (kDM + kDN) * kDK;          // 48 KB
constexpr int kDRowWords = kDN / 2 + 4;                   // staging row: 128 words + pad (528 B: conflict-free 16-byte stores)
constexpr int kDThreads = 256;
constexpr size_t kDSmem = (size_t)kDStages * kDStageBytes + 4 * 32 * kDRowWords * 4 + 1024;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// Operand tiles are re-read by every output tile of their row / column panel: they are loaded with an evict-last L2
// policy and the matrix is stored evict-first, so that the output stream does not push the operand out of L2.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void store_stream(uint4* p, const uint4& v) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
               : "memory");
}
// K-major, no swizzle: LBO = 128 B (>>4 = 8), SBO = 1024 B (>>4 = 64), descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)8 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

// Scatter the shared incidence into the tiled operand (see layout above).  One warp per tree over the build's
// per-entry list ids (HashRFState::ent_lid).
__global__ void __launch_bounds__(256) dense_scatter_kernel(const uint64_t* __restrict__ tb, const uint64_t* __restrict__ te,
                                                            uint32_t T, const uint32_t* __restrict__ ent_lid,
                                                            uint32_t row_groups, uint8_t* __restrict__ At) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  for (uint64_t e = tb[t] + lane_id(); e < te[t]; e += 32) {
    const uint32_t l = ent_lid[e];
    if (l == 0xffffffffu) continue;
    const size_t off = ((size_t)(l >> 7) * row_groups + (t >> 3)) * 1024 + ((l & 127u) >> 4) * 128 + (t & 7u) * 16 + (l & 15u);
    At[off] = 1;
  }
}

struct DenseParams {
  uint32_t T;
  uint64_t p_begin, p_end;
  const uint16_t* nb;
  int uniform_nb;
  uint32_t nb_uniform;
  const uint8_t* At;
  uint32_t row_groups;  // Tpad / 8
  uint32_t k_blocks;    // Kpad / 128
  int32_t editdist;
  uint16_t* out;
  uint32_t* mask;
  const uint2* tiles;   // (row block of 128, column block of 256)
  uint32_t n_tiles;
};

__global__ void __launch_bounds__(kDThreads, 1) dense_gram_kernel(const DenseParams prm) {
  extern __shared__ __align__(1024) unsigned char dsm[];
  unsigned char* stage_base = dsm;
  uint32_t* staging = reinterpret_cast<uint32_t*>(dsm + (size_t)kDStages * kDStageBytes);  // [4 warps][32 rows][kDRowWords]
  __shared__ __align__(8) uint64_t s_full[kDStages], s_empty[kDStages], s_tfull[2], s_tempty[2];
  __shared__ uint32_t s_tmem;
  __shared__ uint16_t s_nb[kDN];

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t T = prm.T;

  if (tid == 0) {
    for (int i = 0; i < kDStages; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_tfull[i], 1);
      mbar_init(&s_tempty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ===================== bulk-copy producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t t = blockIdx.x; t < prm.n_tiles; t += gridDim.x) {
        const uint2 tl = prm.tiles[t];
        for (uint32_t kb = 0; kb < prm.k_blocks; ++kb, ++it) {
          const uint32_t s = it % kDStages, ph = (it / kDStages) & 1u;
          mbar_wait(&s_empty[s], ph ^ 1u);
          mbar_expect_tx(&s_full[s], kDStageBytes);
          unsigned char* sa = stage_base + (size_t)s * kDStageBytes;
          const uint8_t* ga = prm.At + ((size_t)kb * prm.row_groups + (size_t)tl.x * (kDM / 8)) * 1024;
          const uint8_t* gb = prm.At + ((size_t)kb * prm.row_groups + (size_t)tl.y * (kDN / 8)) * 1024;
          bulk_load(sa, ga, kDM * kDK, &s_full[s]);
          bulk_load(sa + kDM * kDK, gb, kDN * kDK, &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // u8 x u8 -> s32, both operands K-major, M = 128, N = 256
      const uint32_t idesc = (2u << 4) | ((uint32_t)(kDN >> 3) << 17) | ((uint32_t)(kDM >> 4) << 24);
      uint32_t it = 0, tile_no = 0;
      for (uint32_t t = blockIdx.x; t < prm.n_tiles; t += gridDim.x, ++tile_no) {
        const uint32_t acc = tile_no & 1u, aph = (tile_no >> 1) & 1u;
        mbar_wait(&s_tempty[acc], aph ^ 1u);  // epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * kDN;
        for (uint32_t kb = 0; kb < prm.k_blocks; ++kb, ++it) {
          const uint32_t s = it % kDStages, ph = (it / kDStages) & 1u;
          mbar_wait(&s_full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_addr(stage_base + (size_t)s * kDStageBytes), sb = sa + kDM * kDK;
#pragma unroll
          for (int k = 0; k < kDK / 32; ++k)
            umma_i8(tmem_d, umma_desc(sa + k * 256), umma_desc(sb + k * 256), idesc, (kb | (uint32_t)k) != 0u);
          umma_commit(&s_empty[s]);  // stage free once these MMAs have read it
        }
        umma_commit(&s_tfull[acc]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const uint32_t q = warp - 4;  // TMEM lanes [32q, 32q + 32)
    uint32_t* my_stage = staging + (size_t)q * 32 * kDRowWords;
    uint32_t tile_no = 0;
    for (uint32_t t = blockIdx.x; t < prm.n_tiles; t += gridDim.x, ++tile_no) {
      const uint2 tl = prm.tiles[t];
      const uint32_t m0 = tl.x * kDM, n0 = tl.y * kDN;
      const uint32_t acc = tile_no & 1u, aph = (tile_no >> 1) & 1u;
      if (!prm.uniform_nb) {
        // |B_b| of the tile's 256 columns, shared by the four epilogue warps
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers are done
        const uint32_t e = tid - 128;
        s_nb[e] = n0 + e < T ? prm.nb[n0 + e] : (uint16_t)0;
        s_nb[e + 128] = n0 + e + 128 < T ? prm.nb[n0 + e + 128] : (uint16_t)0;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const uint32_t a = m0 + 32 * q + lane;
      const uint32_t na = prm.uniform_nb ? prm.nb_uniform : (a < T ? (uint32_t)prm.nb[a] : 0u);
      mbar_wait(&s_tfull[acc], aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + acc * kDN + ((32u * q) << 16);
#pragma unroll 1
      for (int c = 0; c < kDN / 32; ++c) {
        uint32_t v[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr + c * 32)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint32_t nb0, nb1;
          if (prm.uniform_nb) {
            nb0 = nb1 = prm.nb_uniform;
          } else {
            const uint32_t pr = reinterpret_cast<const uint32_t*>(s_nb)[c * 16 + i];
            nb0 = pr & 0xffffu;
            nb1 = pr >> 16;
          }
          const uint32_t d0 = (na + nb0 - 2u * v[2 * i]) & 0xffffu, d1 = (na + nb1 - 2u * v[2 * i + 1]) & 0xffffu;
          w[i] = d0 | (d1 << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(my_stage + (size_t)lane * kDRowWords + c * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      }
      // accumulator drained: hand it back to the MMA warp before the (long) global write-out
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_tempty[acc]);
      // ---- write-out: one row (tree a) at a time, the whole warp on its run of the packed triangle ----
      // lane r prepares row r's run once (64-bit index math); the row loop only shuffles it around
      uint64_t my_e_lo = 0;
      uint32_t my_cnt = 0, my_s0 = 0;
      {
        const uint32_t ar = m0 + 32 * q + lane;
        if (ar + 1 < T) {
          const uint32_t b_lo = max(n0, ar + 1), b_hi = min(n0 + (uint32_t)kDN, T);
          if (b_lo < b_hi) {
            const uint64_t e0 = row_start(ar, T) + (b_lo - ar - 1);
            const uint64_t e_lo = max(e0, prm.p_begin), e_hi = min(e0 + (b_hi - b_lo), prm.p_end);
            if (e_lo < e_hi) {
              my_e_lo = e_lo;
              my_cnt = (uint32_t)(e_hi - e_lo);
              my_s0 = (b_lo - n0) + (uint32_t)(e_lo - e0);  // staging entry of position e_lo
            }
          }
        }
      }
      const uint32_t rows_mask = __ballot_sync(0xffffffffu, my_cnt != 0);
      for (uint32_t rm = rows_mask; rm; rm &= rm - 1) {
        const uint32_t r = __ffs(rm) - 1;
        const uint64_t e_lo = __shfl_sync(0xffffffffu, my_e_lo, r);
        const uint32_t cnt = __shfl_sync(0xffffffffu, my_cnt, r), s0 = __shfl_sync(0xffffffffu, my_s0, r);
        const uint32_t* row = my_stage + (size_t)r * kDRowWords;
        const uint16_t* row16 = reinterpret_cast<const uint16_t*>(row);
        uint16_t* gout = prm.out + ((e_lo & ~7ull) - prm.p_begin);  // 16-byte aligned; entry i is position (e_lo & ~7) + i
        const uint32_t o_lo = (uint32_t)(e_lo & 7ull), o_hi = o_lo + cnt;
        for (uint32_t g = 8 * lane; g < o_hi; g += 256) {  // group of 8 entries at [g, g + 8)
          const int32_t idx = (int32_t)(s0 + g) - (int32_t)o_lo;  // staging entry of the group's first position
          if (g >= o_lo && g + 8 <= o_hi) {
            const uint32_t* wp = row + (idx >> 1);
            const uint32_t sh = ((uint32_t)idx & 1u) * 16u;
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
            uint4 o;
            o.x = __funnelshift_r(w0, w1, sh);
            o.y = __funnelshift_r(w1, w2, sh);
            o.z = __funnelshift_r(w2, w3, sh);
            o.w = __funnelshift_r(w3, w4, sh);
            store_stream(reinterpret_cast<uint4*>(gout + g), o);
            if (prm.editdist >= 0) {
              const uint32_t thr = (uint32_t)prm.editdist, thr2 = thr | (thr << 16);
              const uint32_t r0 = __vsetleu2(o.x, thr2), r1 = __vsetleu2(o.y, thr2), r2 = __vsetleu2(o.z, thr2),
                             r3 = __vsetleu2(o.w, thr2);
              const uint32_t bits = ((r0 | (r0 >> 15)) & 3u) | (((r1 | (r1 >> 15)) & 3u) << 2) |
                                    (((r2 | (r2 >> 15)) & 3u) << 4) | (((r3 | (r3 >> 15)) & 3u) << 6);
              // every mask write of this kernel is a 32-bit atomicOr into the pre-zeroed mask: a group's byte shares
              // its word with groups of other rows / warps / CTAs, and a plain byte store next to their word-sized
              // atomics would be a mixed-size race
              if (bits) {
                const uint64_t byte = ((e_lo & ~7ull) - prm.p_begin + g) >> 3;
                atomicOr(&prm.mask[byte >> 2], bits << ((byte & 3) * 8));
              }
            }
          } else {
            uint32_t bits = 0;
            for (uint32_t i = 0; i < 8; ++i) {
              const uint32_t e = g + i;
              if (e >= o_lo && e < o_hi) {
                const uint16_t dv = row16[idx + (int32_t)i];
                gout[e] = dv;
                if (prm.editdist >= 0 && (int32_t)dv <= prm.editdist) bits |= 1u << i;
              }
            }
            if (bits) {
              const uint64_t byte = ((e_lo & ~7ull) - prm.p_begin + g) >> 3;
              atomicOr(&prm.mask[byte >> 2], bits << ((byte & 3) * 8));  // group shared with another row/tile (mask pre-zeroed)
            }
          }
        }
      }
      __syncwarp();  // staging is rewritten by the next tile
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Builds (once per load) the tiled uint8 operand of the shared incidence.
static int dense_prepare(Ctx* ctx) {
  HashRFState& H = ctx->h;
  if (H.dense_ready) return PRF_OK;
  cudaStream_t s = ctx->stream;
  const uint32_t Tpad = (H.T + kDN - 1) / kDN * kDN;
  const uint32_t Kpad = std::max<uint32_t>(kDK, (H.n_list_ids + kDK - 1) / kDK * kDK);
  const size_t bytes = (size_t)Tpad * Kpad;
  size_t free_b = 0, total_b = 0;
  PRF_CK(ctx, cudaMemGetInfo(&free_b, &total_b));
  if (bytes > free_b / 2)
    return ctx->fail(PRF_E_NOMEM, "dense variant needs a %zu-byte operand (%u trees x %u shared bipartitions)", bytes, Tpad,
                     Kpad);
  PRF_CK(ctx, H.dense_At.alloc(bytes, s));
  PRF_CK(ctx, cudaMemsetAsync(H.dense_At.p, 0, bytes, s));
  if (!H.ent_lid.p) return ctx->fail(PRF_E_UNSUPPORTED, "the dense variant needs the incidence of a single-context build");
  if (H.T && H.n_list_ids) {
    dense_scatter_kernel<<<(H.T + 7) / 8, 256, 0, s>>>(H.ent_tb.p, H.ent_te.p, H.T, H.ent_lid.p, Tpad / 8, H.dense_At.p);
    PRF_LAUNCHED(ctx);
  }
  H.dense_Tpad = Tpad;
  H.dense_Kpad = Kpad;
  H.dense_ready = true;
  return PRF_OK;
}

int launch_dense(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask) {
  HashRFState& H = ctx->h;
  cudaStream_t s = ctx->stream;
  int rc = dense_prepare(ctx);
  if (rc) return rc;
  // tiles of the upper triangle that intersect the rows of [p_begin, p_end)
  const uint32_t a_first = row_of(p_begin, H.T), a_last = row_of(p_end - 1, H.T);
  std::vector<uint2> tiles;
  const uint32_t n_cb = H.dense_Tpad / kDN;
  // order: groups of 8 row blocks, column block by column block, so that the ~148 tiles in flight share
  // a few operand tiles (8 row-block operands, ~18 column-block operands: L2-resident working set)
  constexpr uint32_t kGroup = 8;
  for (uint32_t mg = a_first / kDM; mg <= a_last / kDM; mg += kGroup)
    for (uint32_t nj = 0; nj < n_cb; ++nj)
      for (uint32_t mi = mg; mi < mg + kGroup && mi <= a_last / kDM; ++mi)
        if ((uint64_t)nj * kDN + kDN - 1 > (uint64_t)mi * kDM) tiles.push_back(make_uint2(mi, nj));  // some b > some a
  if (tiles.empty()) return PRF_OK;
  PRF_CK(ctx, H.dense_tiles.alloc(tiles.size(), s));
  PRF_CK(ctx, cudaMemcpyAsync(H.dense_tiles.p, tiles.data(), tiles.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));  // `tiles` is pageable host memory
  PRF_CK(ctx, cudaFuncSetAttribute(dense_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDSmem));
  DenseParams prm{};
  prm.T = H.T;
  prm.p_begin = p_begin;
  prm.p_end = p_end;
  // the operand is the ORIGINAL incidence: with majority lists flipped for the sparse kernel, |B_t| before the flip
  prm.nb = H.n_flipped ? H.nb0.p : H.nb.p;
  prm.uniform_nb = H.n_flipped ? 0 : (H.uniform_nb ? 1 : 0);
  prm.nb_uniform = H.nb_uniform;
  prm.At = H.dense_At.p;
  prm.row_groups = H.dense_Tpad / 8;
  prm.k_blocks = H.dense_Kpad / kDK;
  prm.editdist = editdist;
  prm.out = d_out;
  prm.mask = d_mask;
  prm.tiles = H.dense_tiles.p;
  prm.n_tiles = (uint32_t)tiles.size();
  if (d_mask) PRF_CK(ctx, cudaMemsetAsync(d_mask, 0, ((p_end - p_begin + 31) / 32) * sizeof(uint32_t), s));
  const uint32_t grid = std::min<uint32_t>(prm.n_tiles, (uint32_t)ctx->sm_count);
  dense_gram_kernel<<<grid, kDThreads, kDSmem, s>>>(prm);
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace prf
