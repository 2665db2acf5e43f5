// prf_tolerant.cu -- --tolerant mode: kernel (prf_tolerant.cuh) and its C ABI entry points.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_tolerant.cuh"

using namespace prf;

extern "C" {

// ---------------------------------------------------------------------------------------------
// tolerant
// ---------------------------------------------------------------------------------------------

static int tolerant_load_impl(prf_ctx* h, const uint64_t* clades, const uint64_t* outsets, const uint64_t* tree_off,
                              const uint64_t* leafsets, const uint64_t* dup_off, const uint32_t* dup_label,
                              const uint32_t* dup_extra, bool dups, uint32_t T, uint32_t W, int kind, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!tree_off || (T && !leafsets)) return ctx->fail(PRF_E_INVALID, "tree_off / leafsets is NULL");
  if (dups && !dup_off) return ctx->fail(PRF_E_INVALID, "dup_off is NULL");
  if (W == 0 || W > kTolMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "tolerant mode supports W in 1..%u, got %u", kTolMaxWords, W);
  if (kind != PRF_MEM_HOST && kind != PRF_MEM_DEVICE) return ctx->fail(PRF_E_INVALID, "bad memory kind %d", kind);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint64_t launches0 = ctx->launches;
  TolerantState& S = ctx->tol;
  S = TolerantState{};
  ctx->out_valid = ctx->mask_valid = false;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));
  std::vector<uint64_t> hoff((size_t)T + 1);
  if (kind == PRF_MEM_HOST) {
    std::memcpy(hoff.data(), tree_off, ((size_t)T + 1) * 8);
  } else {
    PRF_CK(ctx, cudaMemcpyAsync(hoff.data(), tree_off, ((size_t)T + 1) * 8, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
  }
  if (hoff[0] != 0) return ctx->fail(PRF_E_INVALID, "tree_off[0] must be 0");
  uint32_t maxc = 0;
  for (uint32_t t = 0; t < T; ++t) {
    if (hoff[t + 1] < hoff[t]) return ctx->fail(PRF_E_INVALID, "tree_off is not ascending at tree %u", t);
    maxc = std::max<uint64_t>(maxc, hoff[t + 1] - hoff[t]);
  }
  uint64_t nnz = hoff[T];
  if (nnz && !clades) return ctx->fail(PRF_E_INVALID, "clades is NULL");
  if (nnz >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu clades; supported < 2^32 - 1", (unsigned long long)nnz);
  if (dups && nnz && !outsets) return ctx->fail(PRF_E_INVALID, "outsets is NULL");
  std::vector<uint64_t> hdoff;
  uint64_t n_dup = 0;
  if (dups) {  // (host arrays in both memory kinds: a handful of entries)
    hdoff.assign(dup_off, dup_off + T + 1);
    if (hdoff[0] != 0) return ctx->fail(PRF_E_INVALID, "dup_off[0] must be 0");
    for (uint32_t t = 0; t < T; ++t)
      if (hdoff[t + 1] < hdoff[t]) return ctx->fail(PRF_E_INVALID, "dup_off is not ascending at tree %u", t);
    n_dup = hdoff[T];
    if (n_dup && (!dup_label || !dup_extra)) return ctx->fail(PRF_E_INVALID, "dup_label / dup_extra is NULL");
    for (uint64_t x = 0; x < n_dup; ++x)
      if (dup_label[x] >= 64u * W) return ctx->fail(PRF_E_INVALID, "dup_label[%llu]=%u does not fit %u words", (unsigned long long)x, dup_label[x], W);
  }
  if (maxc > kTolMaxClades)
    return ctx->fail(PRF_E_UNSUPPORTED, "a tree has %u interior clades; tolerant mode supports <= %u", maxc, kTolMaxClades);
  PRF_CK(ctx, S.off.alloc((size_t)T + 1, s));
  PRF_CK(ctx, S.clades.alloc(nnz * W, s));
  PRF_CK(ctx, S.leafsets.alloc((size_t)T * W, s));
  int rc;
  if ((rc = upload(ctx, S.off.p, hoff.data(), ((size_t)T + 1) * 8, PRF_MEM_HOST))) return rc;
  if ((rc = upload(ctx, S.clades.p, clades, nnz * W * 8, kind))) return rc;
  if ((rc = upload(ctx, S.leafsets.p, leafsets, (size_t)T * W * 8, kind))) return rc;
  S.T = T;
  S.W = W;
  S.dups = dups;
  if (dups) {
    PRF_CK(ctx, S.outsets.alloc(nnz * W, s));
    PRF_CK(ctx, S.dup_off.alloc((size_t)T + 1, s));
    PRF_CK(ctx, S.dup_label.alloc(n_dup, s));
    PRF_CK(ctx, S.dup_extra.alloc(n_dup, s));
    if ((rc = upload(ctx, S.outsets.p, outsets, nnz * W * 8, kind))) return rc;
    if ((rc = upload(ctx, S.dup_off.p, hdoff.data(), ((size_t)T + 1) * 8, PRF_MEM_HOST))) return rc;
    if ((rc = upload(ctx, S.dup_label.p, dup_label, n_dup * 4, PRF_MEM_HOST))) return rc;
    if ((rc = upload(ctx, S.dup_extra.p, dup_extra, n_dup * 4, PRF_MEM_HOST))) return rc;
  } else {
    PRF_CK(ctx, S.par.alloc(nnz, s));
    if ((rc = launch_tol_parents(ctx))) return rc;
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  S.nnz = nnz;
  S.max_clades = maxc;
  S.P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  S.loaded = true;
  ctx->h.loaded = false;
  S.stats = prf_stats{};
  S.stats.n_trees = T;
  S.stats.n_words = W;
  S.stats.nnz = nnz;
  S.stats.n_pairs = S.P;
  S.stats.max_bips = maxc;
  S.stats.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
  S.stats.gpu_launches = (int32_t)(ctx->launches - launches0);
  if (stats) *stats = S.stats;
  return PRF_OK;
}

PRF_API int prf_tolerant_load(prf_ctx* h, const uint64_t* clades, const uint64_t* tree_off,
                              const uint64_t* leafsets, uint32_t T, uint32_t W, int kind, prf_stats* stats) {
  return tolerant_load_impl(h, clades, nullptr, tree_off, leafsets, nullptr, nullptr, nullptr, false, T, W, kind, stats);
}

PRF_API int prf_tolerant_load_dups(prf_ctx* h, const uint64_t* clades, const uint64_t* outsets, const uint64_t* tree_off,
                                   const uint64_t* leafsets, const uint64_t* dup_off, const uint32_t* dup_label,
                                   const uint32_t* dup_extra, uint32_t T, uint32_t W, int kind, prf_stats* stats) {
  return tolerant_load_impl(h, clades, outsets, tree_off, leafsets, dup_off, dup_label, dup_extra, true, T, W, kind, stats);
}

PRF_API int prf_tolerant_compute(prf_ctx* h, uint64_t p_begin, uint64_t p_end, int32_t editdist, uint16_t* d_out,
                                 uint32_t* d_mask, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  TolerantState& S = ctx->tol;
  if (!S.loaded) return ctx->fail(PRF_E_STATE, "prf_tolerant_compute before a successful prf_tolerant_load");
  int rc = check_range(ctx, S.P, p_begin, p_end);
  if (rc) return rc;
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  uint64_t launches0 = ctx->launches;
  if ((rc = own_output(ctx, p_begin, p_end, editdist >= 0, &d_out, &d_mask))) return rc;
  if (editdist < 0) d_mask = nullptr;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
  if (p_end > p_begin && (rc = launch_tolerant(ctx, p_begin, p_end, editdist, d_out, d_mask))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
  if (stats) {
    PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
    prf_stats st = S.stats;
    st.ms_matrix = ev_ms(ctx->ev[4], ctx->ev[5]);
    st.gpu_launches = (int32_t)(ctx->launches - launches0);
    *stats = st;
  }
  return PRF_OK;
}

PRF_API int prf_tolerant(prf_ctx* h, const uint64_t* clades, const uint64_t* tree_off, const uint64_t* leafsets,
                         uint32_t T, uint32_t W, int32_t editdist, uint16_t* out_upper, uint32_t* out_mask,
                         prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  prf_stats st{}, st2{};
  int rc = prf_tolerant_load(h, clades, tree_off, leafsets, T, W, PRF_MEM_HOST, &st);
  if (rc) return rc;
  uint64_t P = ctx->tol.P;
  if ((rc = prf_tolerant_compute(h, 0, P, editdist, nullptr, nullptr, &st2))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
  if (out_upper && (rc = prf_fetch(h, 0, P, out_upper))) return rc;
  if (out_mask && editdist >= 0 && (rc = prf_fetch_mask(h, 0, P, out_mask))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    *stats = st2;
    stats->ms_h2d = st.ms_h2d;
    stats->ms_d2h = ev_ms(ctx->ev[6], ctx->ev[7]);
    stats->ms_total = ev_ms(ctx->ev[0], ctx->ev[7]);
    stats->gpu_launches = st.gpu_launches + st2.gpu_launches;
  }
  return PRF_OK;
}

}  // extern "C"
