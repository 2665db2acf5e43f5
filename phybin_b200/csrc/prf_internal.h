// prf_internal.h -- what the translation units of libphybin_rf.so share: the context handle, small host
// helpers of the C ABI functions, and the launchers each kernel file exports.
//   prf_api.cu        lifetime, hashRF load / compute / fetch, multi-GPU build entry points, debug hooks
//   prf_build.cu      prf_incidence.cuh  (dedup + incidence)
//   prf_matrix.cu     prf_rows.cuh       (sparse matrix kernel)
//   prf_dense.cu      prf_dense.cuh      (tcgen05 Gram variant)
//   prf_tolerant.cu   prf_tolerant.cuh   + the prf_tolerant* entry points
//   prf_extract.cu    prf_extract.cuh    + prf_allbips*, prf_hashrf_trees, prf_tolerant_trees
//   prf_consumers.cu  prf_cluster.cuh, prf_consensus.cuh + their entry points
//   prf_team.cu       prf_team.cuh       multi-GPU build (prf_team_*, prf_multi_*)
#pragma once

#include <algorithm>
#include <cstring>
#include <vector>

#include "prf_common.cuh"

struct prf_ctx {
  prf::Ctx c;
};

namespace prf {

// kernel launchers (defined next to their kernels)
// tree t's entries are [tb[t], te[t]) of d_bips (W words each); span = one past the largest entry index
// seg_*: the dense entry segments holding all entries (n_segs = 0: the one segment [0, nnz))
int hashrf_build(Ctx* ctx, const uint64_t* d_bips, const uint64_t* tb, const uint64_t* te, uint32_t T, uint32_t W,
                 uint64_t span, uint64_t nnz, uint32_t max_entries, const uint64_t* seg_begin = nullptr,
                 const uint64_t* seg_len = nullptr, uint32_t n_segs = 0);
int build_nb_shift(Ctx* ctx);
void team_release_ctx(Ctx* ctx);  // prf_team.cu: frees a context's team arena (prf_destroy)
void build_row_plan(uint32_t T, uint64_t p_begin, uint64_t p_end, uint32_t tile, std::vector<RowItem>& rows,
                    std::vector<RowItem>& groups);
int launch_rows(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask);
int launch_dense(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask);
int launch_tol_parents(Ctx* ctx);  // parent clade of every clade of the loaded tolerant problem
int launch_tolerant(Ctx* ctx, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out, uint32_t* d_mask);
uint32_t dense_kpad(uint32_t n_shared);     // shared bipartitions padded to the dense variant's k-block
uint32_t dense_tpad(uint32_t T);            // trees padded to its column tile

inline int check_range(Ctx* ctx, uint64_t P, uint64_t p_begin, uint64_t p_end) {
  if (p_begin > p_end || p_end > P)
    return ctx->fail(PRF_E_INVALID, "packed range [%llu, %llu) outside [0, %llu]",
                     (unsigned long long)p_begin, (unsigned long long)p_end, (unsigned long long)P);
  if (p_begin % kShardAlign != 0 && p_begin != P)
    return ctx->fail(PRF_E_INVALID, "p_begin=%llu is not a multiple of prf_shard_align()=%llu",
                     (unsigned long long)p_begin, (unsigned long long)kShardAlign);
  return PRF_OK;
}

inline float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
  return ms;
}

// Points *d_out / *d_mask at context-owned buffers covering [p_begin, p_end) when the caller
// passed none.
inline int own_output(Ctx* ctx, uint64_t p_begin, uint64_t p_end, bool want_mask, uint16_t** d_out, uint32_t** d_mask) {
  uint64_t n = p_end - p_begin;
  if (!*d_out) {
    ctx->out_valid = false;
    if (ctx->out.n < n) PRF_CK(ctx, ctx->out.alloc(n, ctx->stream));
    *d_out = ctx->out.p;
    ctx->out_begin = p_begin;
    ctx->out_end = p_end;
    ctx->out_valid = true;
  }
  if (want_mask && !*d_mask) {
    ctx->mask_valid = false;
    uint64_t words = (n + 31) / 32;
    if (ctx->mask.n < words) PRF_CK(ctx, ctx->mask.alloc(words, ctx->stream));
    *d_mask = ctx->mask.p;
    ctx->mask_begin = p_begin;
    ctx->mask_end = p_end;
    ctx->mask_valid = true;
  }
  return PRF_OK;
}

inline int upload(Ctx* ctx, void* dst, const void* src, size_t bytes, int kind) {
  if (!bytes) return PRF_OK;
  PRF_CK(ctx, cudaMemcpyAsync(dst, src, bytes,
                              kind == PRF_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                              ctx->stream));
  return PRF_OK;
}


}  // namespace prf
