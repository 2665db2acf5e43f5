// prf_extract.cu -- allBips on the device (prf_extract.cuh) and the entry points that start from parsed trees.
#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_extract.cuh"

using namespace prf;

extern "C" {

// ---------------------------------------------------------------------------------------------
// bipartition extraction on the device
// ---------------------------------------------------------------------------------------------

PRF_API int prf_allbips(prf_ctx* h, const int32_t* node_label, const uint32_t* node_parent, const uint64_t* tree_begin,
                        const uint32_t* num_leaves, uint32_t T, uint32_t W, int kind, int mode, prf_extract_info* info) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!tree_begin || (T && !num_leaves)) return ctx->fail(PRF_E_INVALID, "tree_begin / num_leaves is NULL");
  if (W == 0 || W > kMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
  if (kind != PRF_MEM_HOST && kind != PRF_MEM_DEVICE) return ctx->fail(PRF_E_INVALID, "bad memory kind %d", kind);
  if (mode != PRF_EXTRACT_BIPS && mode != PRF_EXTRACT_CLADES) return ctx->fail(PRF_E_INVALID, "unknown extraction mode %d", mode);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const uint64_t launches0 = ctx->launches;
  ExtractState& E = ctx->ext;
  E = ExtractState{};
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));

  // tree offsets are needed on the host (largest tree -> shared memory per warp) and on the device
  std::vector<uint64_t> hbeg((size_t)T + 1);
  if (kind == PRF_MEM_HOST) {
    std::memcpy(hbeg.data(), tree_begin, ((size_t)T + 1) * 8);
  } else {
    PRF_CK(ctx, cudaMemcpyAsync(hbeg.data(), tree_begin, ((size_t)T + 1) * 8, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
  }
  if (hbeg[0] != 0) return ctx->fail(PRF_E_INVALID, "tree_begin[0] must be 0");
  std::vector<uint32_t> hleaves_own;
  const uint32_t* hleaves = num_leaves;  // needed on the host too: they decide how many nodes own a set
  if (kind == PRF_MEM_DEVICE && T) {
    hleaves_own.resize(T);
    PRF_CK(ctx, cudaMemcpyAsync(hleaves_own.data(), num_leaves, (size_t)T * 4, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    hleaves = hleaves_own.data();
  }
  uint32_t max_nodes = 1, max_slots = 1;
  for (uint32_t t = 0; t < T; ++t) {
    if (hbeg[t + 1] < hbeg[t]) return ctx->fail(PRF_E_INVALID, "tree_begin is not ascending at tree %u", t);
    const uint64_t m = hbeg[t + 1] - hbeg[t];
    if (m >= kExtractMaxNodes) return ctx->fail(PRF_E_UNSUPPORTED, "tree %u has %llu nodes; supported < %u", t, (unsigned long long)m, kExtractMaxNodes);
    max_nodes = std::max(max_nodes, (uint32_t)m);
    max_slots = std::max(max_slots, extract_slots((uint32_t)m, hleaves[t]));
  }
  const uint64_t n_nodes = hbeg[T];
  if (n_nodes >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "%llu nodes; supported < 2^32 - 1", (unsigned long long)n_nodes);
  if (n_nodes && (!node_label || !node_parent)) return ctx->fail(PRF_E_INVALID, "node_label / node_parent is NULL");
  const size_t warp_bytes = extract_warp_bytes(max_slots, max_nodes, W);
  if (warp_bytes > kExtractSmemBudget)
    return ctx->fail(PRF_E_UNSUPPORTED, "a tree with %u interior nodes x %u words needs %zu bytes of shared memory; the limit is %zu",
                     max_slots, W, warp_bytes, kExtractSmemBudget);
  // the kernel is latency-bound (a dependent sweep per tree): pick the CTA shape that keeps most warps resident
  uint32_t warps = 1, best = 0;
  for (uint32_t ctas = 1; ctas <= 4; ++ctas) {
    const size_t per_cta = (size_t)(220 * 1024) / ctas - 1024;
    const uint32_t w = (uint32_t)std::min<size_t>(8, std::min(per_cta, kExtractSmemBudget) / warp_bytes);
    if (w * ctas > best) { best = w * ctas; warps = w; }
  }

  DevBuf<int32_t> d_label_own;
  DevBuf<uint32_t> d_parent_own, d_leaves_own, count, off32, err;
  DevBuf<uint64_t> d_begin_own, scratch;
  const int32_t* d_label = node_label;
  const uint32_t *d_parent = node_parent, *d_leaves = num_leaves;
  const uint64_t* d_begin = tree_begin;
  int rc;
  if (kind == PRF_MEM_HOST) {
    PRF_CK(ctx, d_label_own.alloc(n_nodes, s));
    PRF_CK(ctx, d_parent_own.alloc(n_nodes, s));
    PRF_CK(ctx, d_leaves_own.alloc(T, s));
    PRF_CK(ctx, d_begin_own.alloc((size_t)T + 1, s));
    if ((rc = upload(ctx, d_label_own.p, node_label, n_nodes * 4, kind))) return rc;
    if ((rc = upload(ctx, d_parent_own.p, node_parent, n_nodes * 4, kind))) return rc;
    if ((rc = upload(ctx, d_leaves_own.p, num_leaves, (size_t)T * 4, kind))) return rc;
    if ((rc = upload(ctx, d_begin_own.p, tree_begin, ((size_t)T + 1) * 8, kind))) return rc;
    d_label = d_label_own.p; d_parent = d_parent_own.p; d_leaves = d_leaves_own.p; d_begin = d_begin_own.p;
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  PRF_CK(ctx, scratch.alloc(n_nodes * W, s));
  PRF_CK(ctx, count.alloc((size_t)T + 1, s));
  PRF_CK(ctx, off32.alloc((size_t)T + 1, s));
  PRF_CK(ctx, err.alloc(1, s));
  PRF_CK(ctx, cudaMemsetAsync(err.p, 0, 4, s));
  PRF_CK(ctx, cudaMemsetAsync(count.p + T, 0, 4, s));  // the scan's last element becomes the total
  PRF_CK(ctx, E.off.alloc((size_t)T + 1, s));
  if (mode == PRF_EXTRACT_CLADES) {
    PRF_CK(ctx, E.leafsets.alloc((size_t)T * W, s));
    PRF_CK(ctx, cudaMemsetAsync(E.leafsets.p, 0, E.leafsets.bytes(), s));
  }
  if (T) {
    switch (W) {
#define PRF_W_CASE(w)                                                                                                     \
  case w:                                                                                                                 \
    rc = launch_extract_w<w>(ctx, mode, d_label, d_parent, d_begin, d_leaves, T, warps, (uint32_t)warp_bytes, max_slots,   \
                             max_nodes, scratch.p, count.p, E.leafsets.p, err.p);                                         \
    break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7) PRF_W_CASE(8)
      PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13) PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
      default: rc = PRF_E_UNSUPPORTED;
    }
    if (rc) return rc;
  }
  if ((rc = exclusive_scan(ctx, LoadU32{count.p}, (uint64_t)T + 1, off32.p, nullptr))) return rc;
  uint32_t hnnz = 0, herr = 0;
  PRF_CK(ctx, cudaMemcpyAsync(&hnnz, off32.p + T, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaMemcpyAsync(&herr, err.p, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (herr & 1u) return ctx->fail(PRF_E_INVALID, "a node's parent is not an interior node that precedes it inside its own tree (nodes must be in pre-order)");
  if (herr & 2u) return ctx->fail(PRF_E_INVALID, "a leaf label does not fit %u words", W);
  if (herr & 4u) return ctx->fail(PRF_E_INVALID, "a tree's num_leaves is not its number of leaf nodes, or exceeds the %u labels %u words hold", 64 * W, W);
  PRF_CK(ctx, E.keys.alloc((size_t)hnnz * W, s));
  extract_pack_kernel<<<(T + 1 + 7) / 8, 256, 0, s>>>(scratch.p, d_begin, off32.p, T, W, E.keys.p, E.off.p);
  PRF_LAUNCHED(ctx);
  PRF_CK(ctx, cudaEventRecord(ctx->ev[2], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  E.valid = true;
  E.mode = mode;
  E.T = T;
  E.W = W;
  E.nnz = hnnz;
  if (info) {
    info->nnz = hnnz;
    info->max_nodes = max_nodes;
    info->gpu_launches = (int32_t)(ctx->launches - launches0);
    info->ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    info->ms_kernels = ev_ms(ctx->ev[1], ctx->ev[2]);
  }
  return PRF_OK;
}

PRF_API int prf_allbips_result(prf_ctx* h, const uint64_t** d_keys, const uint64_t** d_tree_off, const uint64_t** d_leafsets,
                               uint64_t* nnz) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  const ExtractState& E = ctx->ext;
  if (!E.valid) return ctx->fail(PRF_E_STATE, "no extraction result on this context (call prf_allbips first)");
  if (d_keys) *d_keys = E.keys.p;
  if (d_tree_off) *d_tree_off = E.off.p;
  if (d_leafsets) *d_leafsets = E.mode == PRF_EXTRACT_CLADES ? E.leafsets.p : nullptr;
  if (nnz) *nnz = E.nnz;
  return PRF_OK;
}

PRF_API int prf_allbips_fetch(prf_ctx* h, uint64_t* keys, uint64_t* tree_off, uint64_t* leafsets) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  const ExtractState& E = ctx->ext;
  if (!E.valid) return ctx->fail(PRF_E_STATE, "no extraction result on this context (call prf_allbips first)");
  if (leafsets && E.mode != PRF_EXTRACT_CLADES) return ctx->fail(PRF_E_STATE, "the last extraction was not in mode PRF_EXTRACT_CLADES");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (keys && E.nnz) PRF_CK(ctx, cudaMemcpyAsync(keys, E.keys.p, E.nnz * E.W * 8, cudaMemcpyDeviceToHost, s));
  if (tree_off) PRF_CK(ctx, cudaMemcpyAsync(tree_off, E.off.p, ((size_t)E.T + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (leafsets && E.T) PRF_CK(ctx, cudaMemcpyAsync(leafsets, E.leafsets.p, (size_t)E.T * E.W * 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  return PRF_OK;
}

static int trees_one_shot(prf_ctx* h, bool tolerant, const int32_t* node_label, const uint32_t* node_parent,
                          const uint64_t* tree_begin, const uint32_t* num_leaves, uint32_t T, uint32_t W, int32_t editdist,
                          uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats, prf_extract_info* extract) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  if (tolerant && node_label && tree_begin && W && W <= 64) {
    // the one-shot path extracts label SETS; a tree that repeats a leaf label needs prf_tolerant_load_dups' extra inputs
    std::vector<uint32_t> stamp((size_t)W * 64, 0xffffffffu);
    for (uint32_t t = 0; t < T; ++t)
      for (uint64_t v = tree_begin[t]; v < tree_begin[t + 1]; ++v) {
        const int32_t l = node_label[v];
        if (l < 0 || (uint32_t)l >= W * 64) continue;
        if (stamp[l] == t) {
          ctx->err.clear();
          return ctx->fail(PRF_E_UNSUPPORTED, "tree %u repeats leaf label %d: use prf_tolerant_load_dups (inputs from phb_raw_clades_ex)", t, l);
        }
        stamp[l] = t;
      }
  }
  prf_extract_info ei{};
  int rc = prf_allbips(h, node_label, node_parent, tree_begin, num_leaves, T, W, PRF_MEM_HOST,
                       tolerant ? PRF_EXTRACT_CLADES : PRF_EXTRACT_BIPS, &ei);
  if (rc) return rc;
  if (extract) *extract = ei;
  const ExtractState& E = ctx->ext;
  prf_stats st{}, st2{};
  rc = tolerant ? prf_tolerant_load(h, E.keys.p, E.off.p, E.leafsets.p, T, W, PRF_MEM_DEVICE, &st)
                : prf_hashrf_load(h, E.keys.p, E.off.p, T, W, PRF_MEM_DEVICE, &st);
  if (rc) return rc;
  const uint64_t P = tolerant ? ctx->tol.P : ctx->h.P;
  rc = tolerant ? prf_tolerant_compute(h, 0, P, editdist, nullptr, nullptr, &st2)
                : prf_hashrf_compute(h, 0, P, editdist, PRF_VARIANT_AUTO, nullptr, nullptr, &st2);
  if (rc) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
  if (out_upper && (rc = prf_fetch(h, 0, P, out_upper))) return rc;
  if (out_mask && editdist >= 0 && (rc = prf_fetch_mask(h, 0, P, out_mask))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (stats) {
    *stats = st2;
    stats->ms_h2d = ei.ms_h2d;  // the only host-to-device traffic: the tree arrays
    stats->ms_dedup = st.ms_dedup;
    stats->ms_incidence = st.ms_incidence;
    stats->ms_d2h = ev_ms(ctx->ev[6], ctx->ev[7]);
    stats->ms_total = ei.ms_h2d + ei.ms_kernels + st.ms_total + st2.ms_matrix + stats->ms_d2h;
    stats->gpu_launches = ei.gpu_launches + st.gpu_launches + st2.gpu_launches;
  }
  return PRF_OK;
}

PRF_API int prf_hashrf_trees(prf_ctx* h, const int32_t* node_label, const uint32_t* node_parent, const uint64_t* tree_begin,
                             const uint32_t* num_leaves, uint32_t T, uint32_t W, int32_t editdist, uint16_t* out_upper,
                             uint32_t* out_mask, prf_stats* stats, prf_extract_info* extract) {
  return trees_one_shot(h, false, node_label, node_parent, tree_begin, num_leaves, T, W, editdist, out_upper, out_mask, stats,
                        extract);
}

PRF_API int prf_tolerant_trees(prf_ctx* h, const int32_t* node_label, const uint32_t* node_parent, const uint64_t* tree_begin,
                               const uint32_t* num_leaves, uint32_t T, uint32_t W, int32_t editdist, uint16_t* out_upper,
                               uint32_t* out_mask, prf_stats* stats, prf_extract_info* extract) {
  return trees_one_shot(h, true, node_label, node_parent, tree_begin, num_leaves, T, W, editdist, out_upper, out_mask, stats,
                        extract);
}

}  // extern "C"
