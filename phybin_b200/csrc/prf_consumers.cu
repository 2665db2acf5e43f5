// prf_consumers.cu -- what consumes the matrix: flat clusters from the threshold mask, sub-matrix gather,
// strict consensus per cluster (prf_cluster.cuh, prf_consensus.cuh).
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_cluster.cuh"
#include "prf_agglo.cuh"
#include "prf_consensus.cuh"

using namespace prf;

namespace {

template <class Tv>
int run_agglo(Ctx* ctx, const uint16_t* src, uint64_t src_begin, uint32_t T, const uint32_t* d_ids, uint32_t m, int linkage,
              double thresh, uint32_t* d_ma, uint32_t* d_mb, double* d_md, uint32_t* d_nm) {
  cudaStream_t s = ctx->stream;
  constexpr uint32_t E = AggloVal<Tv>::kPerVec;
  const uint32_t ld = (m + E - 1) / E * E;
  uint32_t G = std::min<uint32_t>((uint32_t)ctx->sm_count, (m + 15) / 16);
  if (ctx->debug_agglo_ctas) G = std::max<uint32_t>(1, std::min<uint32_t>(G, ctx->debug_agglo_ctas));
  const uint32_t R0 = (m + G - 1) / G;
  const size_t smem = (size_t)R0 * 20 + 512 + 4 + 128 + 256;
  PRF_CK(ctx, cudaFuncSetAttribute(agglo_kernel<Tv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PRF_CK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, agglo_kernel<Tv>, (int)kAggloThreads, smem));
  if (per_sm < 1) return ctx->fail(PRF_E_UNSUPPORTED, "agglomeration of %u trees does not fit a CTA's shared memory", m);
  DevBuf<Tv> D;
  DevBuf<uint32_t> sizes;
  DevBuf<AggloRecord> rec;
  if (D.alloc((size_t)m * ld, s) != cudaSuccess || sizes.alloc((size_t)G * m, s) != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PRF_E_NOMEM, "no device memory for a %u x %u matrix of %zu-byte distances", m, m, sizeof(Tv));
  }
  PRF_CK(ctx, rec.alloc((size_t)2 * G, s));
  PRF_CK(ctx, cudaMemsetAsync(rec.p, 0, (size_t)2 * G * sizeof(AggloRecord), s));  // stamp 0 = not yet published
  agglo_gather_kernel<Tv><<<dim3((ld + 255) / 256, std::min<uint32_t>(m, 65535u)), 256, 0, s>>>(src, src_begin, T, d_ids, m, ld, D.p);
  PRF_LAUNCHED(ctx);
  AggloParams prm;
  prm.D = D.p;
  prm.m = m;
  prm.ld = ld;
  prm.linkage = linkage;
  prm.thresh = thresh;
  prm.sizes = sizes.p;
  prm.rec = rec.p;
  prm.merge_a = d_ma;
  prm.merge_b = d_mb;
  prm.merge_d = d_md;
  prm.n_merges = d_nm;
  void* args[] = {&prm};
  PRF_CK(ctx, cudaLaunchCooperativeKernel((const void*)agglo_kernel<Tv>, dim3(G), dim3(kAggloThreads), args, smem, s));
  PRF_LAUNCHED(ctx);
  return PRF_OK;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------------------------
// flat clusters from the threshold mask
// ---------------------------------------------------------------------------------------------

// Union-find over the mask bits of [p_begin, p_end) plus the edges (t, seed[i][t]) of n_seed label arrays.
static int components_impl(Ctx* ctx, const uint32_t* d_mask, uint64_t p_begin, uint64_t p_end, uint32_t T, const uint32_t* seed,
                           uint32_t n_seed, uint32_t* labels, uint32_t* n_components) {
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<uint32_t> parent, label, roots, d_seed;
  PRF_CK(ctx, parent.alloc(T, s));
  PRF_CK(ctx, label.alloc(T, s));
  PRF_CK(ctx, roots.alloc(1, s));
  PRF_CK(ctx, cudaMemsetAsync(roots.p, 0, 4, s));
  uint32_t hroots = 0;
  if (T) {
    cc_init_kernel<<<grid_for(T), 256, 0, s>>>(parent.p, T);
    PRF_LAUNCHED(ctx);
    const uint64_t n_words = (p_end - p_begin + 31) / 32;
    if (n_words) {
      if (n_words > 0x7fffffffull * 256) return ctx->fail(PRF_E_UNSUPPORTED, "mask too large for one launch");
      cc_hook_kernel<<<(uint32_t)((n_words + 255) / 256), 256, 0, s>>>(d_mask, n_words, p_begin, p_end, T, parent.p);
      PRF_LAUNCHED(ctx);
    }
    if (n_seed) {  // another shard's components as edges (t, its label of t): the union is the graph of both shards
      const uint64_t n = (uint64_t)n_seed * T;
      PRF_CK(ctx, d_seed.alloc(n, s));
      PRF_CK(ctx, cudaMemcpyAsync(d_seed.p, seed, n * 4, cudaMemcpyHostToDevice, s));
      cc_seed_kernel<<<grid_for(n), 256, 0, s>>>(d_seed.p, n, T, parent.p);
      PRF_LAUNCHED(ctx);
    }
    cc_flatten_kernel<<<grid_for(T), 256, 0, s>>>(parent.p, T, label.p, roots.p);
    PRF_LAUNCHED(ctx);
    PRF_CK(ctx, cudaMemcpyAsync(labels, label.p, (size_t)T * 4, cudaMemcpyDeviceToHost, s));
  }
  PRF_CK(ctx, cudaMemcpyAsync(&hroots, roots.p, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (n_components) *n_components = hroots;
  return PRF_OK;
}

PRF_API int prf_mask_components(prf_ctx* h, const uint32_t* d_mask, uint32_t T, uint32_t* labels, uint32_t* n_components) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!labels && T) return ctx->fail(PRF_E_INVALID, "labels is NULL");
  const uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  if (!d_mask) {
    if (!ctx->mask_valid || ctx->mask_begin != 0 || ctx->mask_end != P)
      return ctx->fail(PRF_E_STATE, "no context-owned mask covering the whole triangle (compute with editdist >= 0 first)");
    d_mask = ctx->mask.p;
  }
  return components_impl(ctx, d_mask, 0, P, T, nullptr, 0, labels, n_components);
}

PRF_API int prf_mask_components_shard(prf_ctx* h, const uint32_t* d_mask, uint64_t p_begin, uint64_t p_end, uint32_t T,
                                      const uint32_t* seed, uint32_t n_seed, uint32_t* labels, uint32_t* n_components) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!labels && T) return ctx->fail(PRF_E_INVALID, "labels is NULL");
  if (n_seed && !seed) return ctx->fail(PRF_E_INVALID, "seed is NULL");
  const uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  if (!d_mask && p_end > p_begin) {
    if (!ctx->mask_valid || ctx->mask_begin != p_begin || ctx->mask_end != p_end)
      return ctx->fail(PRF_E_STATE, "the context-owned mask does not cover exactly [%llu, %llu)", (unsigned long long)p_begin,
                       (unsigned long long)p_end);
    d_mask = ctx->mask.p;
  }
  int rc = check_range(ctx, P, p_begin, p_end);
  if (rc) return rc;
  for (uint64_t x = 0; x < (uint64_t)n_seed * T; ++x)
    if (seed[x] >= T) return ctx->fail(PRF_E_INVALID, "seed label %u is not a tree id", seed[x]);
  return components_impl(ctx, d_mask, p_begin, p_end, T, seed, n_seed, labels, n_components);
}

PRF_API int prf_fetch_submatrix(prf_ctx* h, const uint16_t* d_matrix, uint32_t T, const uint32_t* ids, uint32_t m,
                                uint16_t* out) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  const uint16_t* src = d_matrix;
  uint64_t src_begin = 0;
  if (!src) {
    if (!ctx->out_valid) return ctx->fail(PRF_E_STATE, "no context-owned result to fetch");
    T = ctx->h.loaded ? ctx->h.T : (ctx->tol.loaded ? ctx->tol.T : 0);
    src = ctx->out.p;
    src_begin = ctx->out_begin;
  }
  if (m && !ids) return ctx->fail(PRF_E_INVALID, "ids is NULL");
  for (uint32_t i = 0; i < m; ++i)
    if (ids[i] >= T || (i && ids[i] <= ids[i - 1])) return ctx->fail(PRF_E_INVALID, "ids must be ascending tree ids below %u", T);
  const uint64_t n_out = (uint64_t)m * (m ? m - 1 : 0) / 2;
  if (!n_out) return PRF_OK;
  if (!out) return ctx->fail(PRF_E_INVALID, "out is NULL");
  if (!d_matrix) {
    const uint64_t lo = row_start(ids[0], T) + (ids[1] - ids[0] - 1), hi = row_start(ids[m - 2], T) + (ids[m - 1] - ids[m - 2] - 1);
    if (lo < ctx->out_begin || hi >= ctx->out_end) return ctx->fail(PRF_E_INVALID, "the computed range does not cover these pairs");
  }
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<uint32_t> d_ids;
  DevBuf<uint16_t> d_out;
  PRF_CK(ctx, d_ids.alloc(m, s));
  PRF_CK(ctx, d_out.alloc(n_out, s));
  PRF_CK(ctx, cudaMemcpyAsync(d_ids.p, ids, (size_t)m * 4, cudaMemcpyHostToDevice, s));
  submatrix_gather_kernel<<<grid_for(n_out), 256, 0, s>>>(src, src_begin, T, d_ids.p, m, n_out, d_out.p);
  PRF_LAUNCHED(ctx);
  PRF_CK(ctx, cudaMemcpyAsync(out, d_out.p, n_out * 2, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  return PRF_OK;
}

// ---------------------------------------------------------------------------------------------
// agglomerative clustering of one set of trees (the dendrogram)
// ---------------------------------------------------------------------------------------------

PRF_API int prf_agglomerate(prf_ctx* h, const uint16_t* d_matrix, uint32_t T, const uint32_t* ids, uint32_t m, int linkage,
                            double thresh, uint32_t* labels, uint32_t* merge_a, uint32_t* merge_b, double* merge_d,
                            prf_agglo_info* info) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  const uint16_t* src = d_matrix;
  uint64_t src_begin = 0;
  if (!src) {
    if (!ctx->out_valid) return ctx->fail(PRF_E_STATE, "no context-owned result to cluster");
    T = ctx->h.loaded ? ctx->h.T : (ctx->tol.loaded ? ctx->tol.T : 0);
    src = ctx->out.p;
    src_begin = ctx->out_begin;
  }
  if (!ids) m = T;
  if (linkage < 0 || linkage > 2) return ctx->fail(PRF_E_INVALID, "linkage must be 0 (single), 1 (complete) or 2 (UPGMA)");
  if (thresh != thresh) return ctx->fail(PRF_E_INVALID, "thresh is NaN");
  if (ids)
    for (uint32_t i = 0; i < m; ++i)
      if (ids[i] >= T || (i && ids[i] <= ids[i - 1])) return ctx->fail(PRF_E_INVALID, "ids must be ascending tree ids below %u", T);
  if (m >= 2 && !d_matrix) {
    const uint32_t i0 = ids ? ids[0] : 0, i1 = ids ? ids[1] : 1, y0 = ids ? ids[m - 2] : m - 2, y1 = ids ? ids[m - 1] : m - 1;
    const uint64_t lo = row_start(i0, T) + (i1 - i0 - 1), hi = row_start(y0, T) + (y1 - y0 - 1);
    if (lo < ctx->out_begin || hi >= ctx->out_end) return ctx->fail(PRF_E_INVALID, "the computed range does not cover these pairs");
  }
  if (info) *info = prf_agglo_info{};
  uint32_t n_merges = 0;
  std::vector<uint32_t> ha, hb;
  float ms = 0.f;
  const uint64_t launches0 = ctx->launches;
  if (m >= 2) {
    PRF_CK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int coop = 0;
    PRF_CK(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return ctx->fail(PRF_E_UNSUPPORTED, "device does not support cooperative launches");
    DevBuf<uint32_t> d_ids, d_ma, d_mb, d_nm;
    DevBuf<double> d_md;
    if (ids) {
      PRF_CK(ctx, d_ids.alloc(m, s));
      PRF_CK(ctx, cudaMemcpyAsync(d_ids.p, ids, (size_t)m * 4, cudaMemcpyHostToDevice, s));
    }
    PRF_CK(ctx, d_ma.alloc(m - 1, s));
    PRF_CK(ctx, d_mb.alloc(m - 1, s));
    PRF_CK(ctx, d_md.alloc(m - 1, s));
    PRF_CK(ctx, d_nm.alloc(1, s));
    PRF_CK(ctx, cudaMemsetAsync(d_nm.p, 0, 4, s));
    PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));
    // single / complete linkage only ever hold input values: the matrix stays uint16; UPGMA averages in double
    const int rc = linkage == 2 ? run_agglo<double>(ctx, src, src_begin, T, ids ? d_ids.p : nullptr, m, linkage, thresh, d_ma.p, d_mb.p, d_md.p, d_nm.p)
                                : run_agglo<uint16_t>(ctx, src, src_begin, T, ids ? d_ids.p : nullptr, m, linkage, thresh, d_ma.p, d_mb.p, d_md.p, d_nm.p);
    if (rc) return rc;
    PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
    PRF_CK(ctx, cudaMemcpyAsync(&n_merges, d_nm.p, 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    if (n_merges > m - 1) return ctx->fail(PRF_E_CUDA, "agglomeration returned %u merges for %u trees", n_merges, m);
    ha.resize(n_merges);
    hb.resize(n_merges);
    if (n_merges) {
      PRF_CK(ctx, cudaMemcpyAsync(ha.data(), d_ma.p, (size_t)n_merges * 4, cudaMemcpyDeviceToHost, s));
      PRF_CK(ctx, cudaMemcpyAsync(hb.data(), d_mb.p, (size_t)n_merges * 4, cudaMemcpyDeviceToHost, s));
      if (merge_d) PRF_CK(ctx, cudaMemcpyAsync(merge_d, d_md.p, (size_t)n_merges * 8, cudaMemcpyDeviceToHost, s));
      PRF_CK(ctx, cudaStreamSynchronize(s));
    }
    ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  }
  if (merge_a && n_merges) std::memcpy(merge_a, ha.data(), (size_t)n_merges * 4);
  if (merge_b && n_merges) std::memcpy(merge_b, hb.data(), (size_t)n_merges * 4);
  if (labels) {
    // a merge keeps the lower key, so a cluster's key is its smallest member
    std::vector<uint32_t> parent(m);
    for (uint32_t i = 0; i < m; ++i) parent[i] = i;
    for (uint32_t k = 0; k < n_merges; ++k) parent[hb[k]] = ha[k];
    for (uint32_t i = 0; i < m; ++i) labels[i] = parent[i] == i ? i : labels[parent[i]];  // parent[i] < i: already final
  }
  if (info) {
    info->n_merges = n_merges;
    info->n_clusters = m - n_merges;
    info->gpu_launches = (int32_t)(ctx->launches - launches0);
    info->ms_kernels = ms;
  }
  return PRF_OK;
}

// ---------------------------------------------------------------------------------------------
// strict consensus of every cluster
// ---------------------------------------------------------------------------------------------

PRF_API int prf_consensus(prf_ctx* h, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W, int kind,
                          const uint32_t* labels, prf_consensus_info* info) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!tree_off || (T && !labels)) return ctx->fail(PRF_E_INVALID, "tree_off / labels is NULL");
  if (W == 0 || W > kMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
  if (kind != PRF_MEM_HOST && kind != PRF_MEM_DEVICE) return ctx->fail(PRF_E_INVALID, "bad memory kind %d", kind);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const uint64_t launches0 = ctx->launches;
  ConsensusState& R = ctx->cons;
  R = ConsensusState{};

  std::vector<uint64_t> hoff((size_t)T + 1);
  if (kind == PRF_MEM_HOST) {
    std::memcpy(hoff.data(), tree_off, ((size_t)T + 1) * 8);
  } else {
    PRF_CK(ctx, cudaMemcpyAsync(hoff.data(), tree_off, ((size_t)T + 1) * 8, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
  }
  if (hoff[0] != 0) return ctx->fail(PRF_E_INVALID, "tree_off[0] must be 0");
  for (uint32_t t = 0; t < T; ++t)
    if (hoff[t + 1] < hoff[t]) return ctx->fail(PRF_E_INVALID, "tree_off is not ascending at tree %u", t);
  const uint64_t nnz = hoff[T];
  if (nnz >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu entries; supported < 2^32 - 1", (unsigned long long)nnz);
  if (nnz && !bips) return ctx->fail(PRF_E_INVALID, "bips is NULL");
  // cluster sizes; a label must be the smallest tree id of its cluster
  std::vector<uint32_t> size(T ? T : 1, 0);
  uint32_t n_clusters = 0;
  for (uint32_t t = 0; t < T; ++t) {
    const uint32_t c = labels[t];
    if (c > t || labels[c] != c)
      return ctx->fail(PRF_E_INVALID, "labels[%u]=%u is not the smallest tree id of a cluster", t, c);
    ++size[c];
    n_clusters += c == t;
  }
  uint64_t rep_entries = 0;
  for (uint32_t t = 0; t < T; ++t)
    if (labels[t] == t && size[t] >= 2) rep_entries += hoff[t + 1] - hoff[t];
  uint64_t cap = 1024;
  while (cap < 2 * rep_entries) cap <<= 1;

  DevBuf<uint64_t> d_off_own, d_bips_own;
  DevBuf<uint32_t> d_labels, d_size, cnt, pos;
  DevBuf<unsigned long long> slots;
  const uint64_t *d_off = tree_off, *d_bips = bips;
  int rc;
  if (kind == PRF_MEM_HOST) {
    PRF_CK(ctx, d_off_own.alloc((size_t)T + 1, s));
    PRF_CK(ctx, d_bips_own.alloc(nnz * W, s));
    if ((rc = upload(ctx, d_off_own.p, tree_off, ((size_t)T + 1) * 8, kind))) return rc;
    if ((rc = upload(ctx, d_bips_own.p, bips, nnz * W * 8, kind))) return rc;
    d_off = d_off_own.p;
    d_bips = d_bips_own.p;
  } else if (nnz && (reinterpret_cast<uintptr_t>(bips) & 15)) {
    return ctx->fail(PRF_E_INVALID, "device bips pointer must be 16-byte aligned");
  }
  PRF_CK(ctx, d_labels.alloc(T, s));
  PRF_CK(ctx, d_size.alloc(T, s));
  if ((rc = upload(ctx, d_labels.p, labels, (size_t)T * 4, PRF_MEM_HOST))) return rc;
  if ((rc = upload(ctx, d_size.p, size.data(), (size_t)T * 4, PRF_MEM_HOST))) return rc;
  PRF_CK(ctx, cnt.alloc(nnz + 1, s));
  PRF_CK(ctx, pos.alloc(nnz + 1, s));
  PRF_CK(ctx, slots.alloc(cap, s));
  PRF_CK(ctx, cudaMemsetAsync(cnt.p, 0, (nnz + 1) * 4, s));
  PRF_CK(ctx, cudaMemsetAsync(slots.p, 0xff, cap * 8, s));
  PRF_CK(ctx, R.off.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));
  if (T) {
    switch (W) {
#define PRF_W_CASE(w) \
  case w: rc = launch_cons_table_w<w>(ctx, d_bips, d_off, d_labels.p, d_size.p, T, slots.p, cap - 1, cnt.p); break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7) PRF_W_CASE(8)
      PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13) PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
      default: rc = PRF_E_UNSUPPORTED;
    }
    if (rc) return rc;
    cons_flag_kernel<<<(T + 7) / 8, 256, 0, s>>>(d_off, d_labels.p, d_size.p, T, cnt.p);
    PRF_LAUNCHED(ctx);
  }
  if ((rc = exclusive_scan(ctx, LoadU32{cnt.p}, nnz + 1, pos.p, nullptr))) return rc;
  uint32_t n_out = 0;
  PRF_CK(ctx, cudaMemcpyAsync(&n_out, pos.p + nnz, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  PRF_CK(ctx, R.keys.alloc((size_t)n_out * W, s));
  cons_gather_kernel<<<(T + 1 + 7) / 8, 256, 0, s>>>(d_bips, d_off, cnt.p, pos.p, T, W, R.keys.p, R.off.p);
  PRF_LAUNCHED(ctx);
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  R.valid = true;
  R.T = T;
  R.W = W;
  R.n_out = n_out;
  if (info) {
    info->n_out = n_out;
    info->n_clusters = n_clusters;
    info->gpu_launches = (int32_t)(ctx->launches - launches0);
    info->ms_kernels = ev_ms(ctx->ev[0], ctx->ev[1]);
    info->reserved = 0;
  }
  return PRF_OK;
}

PRF_API int prf_consensus_fetch(prf_ctx* h, uint64_t* keys, uint64_t* cons_off) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  const ConsensusState& R = ctx->cons;
  if (!R.valid) return ctx->fail(PRF_E_STATE, "no consensus result on this context (call prf_consensus first)");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (keys && R.n_out) PRF_CK(ctx, cudaMemcpyAsync(keys, R.keys.p, R.n_out * R.W * 8, cudaMemcpyDeviceToHost, s));
  if (cons_off) PRF_CK(ctx, cudaMemcpyAsync(cons_off, R.off.p, ((size_t)R.T + 1) * 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  return PRF_OK;
}

}  // extern "C"
