// prf_api.cu -- the C ABI of include/phybin_rf.h.  One translation unit: the kernels live in the
// .cuh files next to this one.  Built for sm_100a only (see phybin_b200/build.py); there is no CPU
// path here and no dependency on the oracle.

#include <algorithm>
#include <cstring>
#include <new>

#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_shard.cuh"

using namespace prf;

static thread_local std::string g_create_error;


extern "C" {

PRF_API const char* prf_version(void) { return "phybin_rf 0.1 (sm_100a)"; }

PRF_API prf_ctx* prf_create(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                     " (phybin_rf has no CPU fallback)";
    cudaGetLastError();
    return nullptr;
  }
  if (device < 0 || device >= ndev) {
    g_create_error = "device index out of range";
    return nullptr;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return nullptr;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    return nullptr;
  }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                     "; this library is built for sm_100a (B200) only";
    return nullptr;
  }
  prf_ctx* h = new (std::nothrow) prf_ctx;
  if (!h) { g_create_error = "out of host memory"; return nullptr; }
  Ctx& c = h->c;
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete h;
    return nullptr;
  }
  for (auto& ev : c.ev) cudaEventCreate(&ev);
  // keep freed blocks in the pool: repeated load/compute cycles reuse them without cudaMalloc
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  g_create_error.clear();
  return h;
}

PRF_API void prf_destroy(prf_ctx* h) {
  if (!h) return;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  c.h = HashRFState{};
  c.tol = TolerantState{};
  c.ext = ExtractState{};
  c.cons = ConsensusState{};
  c.plan = RowPlan{};
  c.timing.release();
  c.out.release();
  c.mask.release();
  cudaStreamSynchronize(c.stream);
  for (auto& ev : c.ev) cudaEventDestroy(ev);
  cudaStreamDestroy(c.stream);
  delete h;
}

PRF_API const char* prf_last_error(const prf_ctx* h) { return h ? h->c.err.c_str() : g_create_error.c_str(); }
PRF_API void* prf_stream(prf_ctx* h) { return h ? (void*)h->c.stream : nullptr; }
PRF_API uint64_t prf_launch_count(const prf_ctx* h) { return h ? h->c.launches : 0; }
PRF_API uint64_t prf_shard_align(void) { return kShardAlign; }
PRF_API void prf_limits(uint32_t* max_words, uint32_t* max_bips) {
  if (max_words) *max_words = kMaxWords;
  if (max_bips) *max_bips = kMaxBipsPerTree;
}

PRF_API uint64_t prf_packed_index(uint32_t T, uint32_t a, uint32_t b) {
  if (!(a < b && b < T)) return ~0ull;
  return row_start(a, T) + (b - a - 1);
}

PRF_API int prf_shard_range(uint32_t T, uint32_t rank, uint32_t world, uint64_t* p_begin, uint64_t* p_end) {
  if (!world || rank >= world || !p_begin || !p_end) return PRF_E_INVALID;
  uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  uint64_t units = (P + kShardAlign - 1) / kShardAlign;  // aligned blocks, the last may be partial
  uint64_t b = units * rank / world, e = units * (rank + 1) / world;
  *p_begin = std::min(P, b * kShardAlign);
  *p_end = std::min(P, e * kShardAlign);
  return PRF_OK;
}

// ---------------------------------------------------------------------------------------------
// hashRF
// ---------------------------------------------------------------------------------------------

PRF_API int prf_hashrf_load(prf_ctx* h, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W,
                            int kind, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!tree_off) return ctx->fail(PRF_E_INVALID, "tree_off is NULL");
  if (W == 0 || W > kMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
  if (kind != PRF_MEM_HOST && kind != PRF_MEM_DEVICE) return ctx->fail(PRF_E_INVALID, "bad memory kind %d", kind);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint64_t launches0 = ctx->launches;
  ctx->h.loaded = false;
  ctx->tol.loaded = false;
  ctx->out_valid = ctx->mask_valid = false;

  trace_mark("hashrf_load: enter");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));
  // offsets first: nnz = tree_off[T]
  DevBuf<uint64_t> d_off_own, d_bips_own;
  const uint64_t* d_off = tree_off;
  uint64_t nnz = 0;
  if (kind == PRF_MEM_HOST) {
    if (tree_off[0] != 0) return ctx->fail(PRF_E_INVALID, "tree_off[0] must be 0");
    for (uint32_t t = 0; t < T; ++t)
      if (tree_off[t + 1] < tree_off[t]) return ctx->fail(PRF_E_INVALID, "tree_off is not ascending at tree %u", t);
    nnz = tree_off[T];
    PRF_CK(ctx, d_off_own.alloc((size_t)T + 1, s));
    int rc = upload(ctx, d_off_own.p, tree_off, ((size_t)T + 1) * 8, kind);
    if (rc) return rc;
    d_off = d_off_own.p;
  } else {
    PRF_CK(ctx, cudaMemcpyAsync(&nnz, tree_off + T, 8, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
  }
  if (nnz >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu entries; supported < 2^32 - 1", (unsigned long long)nnz);
  if (nnz && !bips) return ctx->fail(PRF_E_INVALID, "bips is NULL");
  const uint64_t* d_bips = bips;
  if (kind == PRF_MEM_HOST) {
    PRF_CK(ctx, d_bips_own.alloc(nnz * W, s));
    int rc = upload(ctx, d_bips_own.p, bips, nnz * W * 8, kind);
    if (rc) return rc;
    d_bips = d_bips_own.p;
  } else if (nnz && (reinterpret_cast<uintptr_t>(bips) & 15)) {
    return ctx->fail(PRF_E_INVALID, "device bips pointer must be 16-byte aligned");
  }
  trace_mark("hashrf_load: uploads queued");
  int rc = hashrf_build(ctx, d_bips, d_off, T, W, nnz);
  if (rc) return rc;
  trace_mark("hashrf_load: build done");
  prf_stats& st = ctx->h.stats;
  st.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
  st.ms_dedup = ev_ms(ctx->ev[1], ctx->ev[2]);
  st.ms_incidence = ev_ms(ctx->ev[2], ctx->ev[3]);
  st.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
  st.gpu_launches = (int32_t)(ctx->launches - launches0);
  if (stats) *stats = st;
  return PRF_OK;
}

PRF_API int prf_hashrf_compute(prf_ctx* h, uint64_t p_begin, uint64_t p_end, int32_t editdist, int variant,
                               uint16_t* d_out, uint32_t* d_mask, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  HashRFState& H = ctx->h;
  if (!H.loaded) return ctx->fail(PRF_E_STATE, "prf_hashrf_compute before a successful prf_hashrf_load");
  int rc = check_range(ctx, H.P, p_begin, p_end);
  if (rc) return rc;
  if (d_out && (reinterpret_cast<uintptr_t>(d_out) & 15))
    return ctx->fail(PRF_E_INVALID, "d_out must be 16-byte aligned");
  if (variant != PRF_VARIANT_AUTO && variant != PRF_VARIANT_SPARSE && variant != PRF_VARIANT_DENSE)
    return ctx->fail(PRF_E_INVALID, "unknown variant %d", variant);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  uint64_t launches0 = ctx->launches;
  trace_mark("hashrf_compute: enter");
  if ((rc = own_output(ctx, p_begin, p_end, editdist >= 0, &d_out, &d_mask))) return rc;
  if (editdist < 0) d_mask = nullptr;
  trace_mark("hashrf_compute: output ready");

  int use = variant;
  if (use == PRF_VARIANT_AUTO) {
    // Measured split density decides.  Cost models from this round's B200 runs, per pair
    // of trees: the sparse join pays ~1.9 ps per shared (pair, bipartition) incidence of the (majority-
    // flipped) lists plus ~0.9 ps when |B_t| is not uniform (per-column fill); the tensor-core Gram pays
    // 2 * Kpad int8 ops at the ~0.9 Pop/s the operand traffic allows.  Both write the same 2 B per pair.
    const double pairs = (double)(p_end - p_begin);
    const double hits = (double)H.stats.n_hits * (H.P ? pairs / (double)H.P : 0.0);
    const double kpad = (double)dense_kpad(H.n_shared);
    const double t_sparse = hits * 1.9e-12 + (H.uniform_nb ? 0.0 : pairs * 0.9e-12), t_dense = pairs * kpad * 2.0 / 0.9e15;
    const double operand = (double)dense_tpad(H.T) * kpad;
    use = (H.n_shared > 0 && H.lid.p && operand < 32e9 && t_dense < t_sparse) ? PRF_VARIANT_DENSE : PRF_VARIANT_SPARSE;
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
  if (p_end > p_begin) {
    if (use == PRF_VARIANT_DENSE && !H.lid.p && H.n_shared)
      return ctx->fail(PRF_E_UNSUPPORTED, "the dense variant needs the incidence of a single-context build");
    if (use == PRF_VARIANT_DENSE) rc = launch_dense(ctx, p_begin, p_end, editdist, d_out, d_mask);
    else if (ctx->sparse_kernel == 1) rc = launch_sparse(ctx, p_begin, p_end, editdist, d_out, d_mask);
    else rc = launch_rows(ctx, p_begin, p_end, editdist, d_out, d_mask);
    if (rc) return rc;
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
  H.stats.variant_used = use;
  trace_mark("hashrf_compute: launched");
  if (stats) {
    PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
    trace_mark("hashrf_compute: synced");
    prf_stats st = H.stats;
    st.ms_matrix = ev_ms(ctx->ev[4], ctx->ev[5]);
    st.gpu_launches = (int32_t)(ctx->launches - launches0);
    *stats = st;
  }
  return PRF_OK;
}

static int fetch_common(Ctx* ctx, uint64_t p_begin, uint64_t p_end, uint16_t* dst) {
  if (!ctx->out_valid) return ctx->fail(PRF_E_STATE, "no context-owned result to fetch");
  if (p_begin > p_end || p_begin < ctx->out_begin || p_end > ctx->out_end)
    return ctx->fail(PRF_E_INVALID, "fetch range outside the computed range");
  if (!dst) return ctx->fail(PRF_E_INVALID, "dst is NULL");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  if (p_end > p_begin)
    PRF_CK(ctx, cudaMemcpyAsync(dst, ctx->out.p + (p_begin - ctx->out_begin), (p_end - p_begin) * 2,
                                cudaMemcpyDeviceToHost, ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  return PRF_OK;
}

PRF_API int prf_fetch(prf_ctx* h, uint64_t p_begin, uint64_t p_end, uint16_t* dst) {
  if (!h) return PRF_E_INVALID;
  h->c.err.clear();
  return fetch_common(&h->c, p_begin, p_end, dst);
}

PRF_API int prf_fetch_mask(prf_ctx* h, uint64_t p_begin, uint64_t p_end, uint32_t* dst) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!ctx->mask_valid) return ctx->fail(PRF_E_STATE, "no context-owned mask to fetch");
  if (p_begin > p_end || p_begin < ctx->mask_begin || p_end > ctx->mask_end || (p_begin - ctx->mask_begin) % 32)
    return ctx->fail(PRF_E_INVALID, "mask fetch range must lie in the computed range and start on a 32-entry boundary");
  if (!dst) return ctx->fail(PRF_E_INVALID, "dst is NULL");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  uint64_t words = (p_end - p_begin + 31) / 32;
  if (words)
    PRF_CK(ctx, cudaMemcpyAsync(dst, ctx->mask.p + (p_begin - ctx->mask_begin) / 32, words * 4,
                                cudaMemcpyDeviceToHost, ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  return PRF_OK;
}

PRF_API int prf_fetch_rows(prf_ctx* h, uint32_t row0, uint32_t row1, uint16_t* dst) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  uint32_t T = ctx->h.loaded ? ctx->h.T : (ctx->tol.loaded ? ctx->tol.T : 0);
  if (row0 > row1 || row1 > T) return ctx->fail(PRF_E_INVALID, "row range [%u, %u) outside [0, %u]", row0, row1, T);
  uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  uint64_t b = row0 + 1 >= T ? P : row_start(row0, T);
  uint64_t e = row1 + 1 >= T ? P : row_start(row1, T);
  return fetch_common(ctx, b, e, dst);
}

PRF_API int prf_hashrf(prf_ctx* h, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W,
                       int32_t editdist, uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  prf_stats st{};
  int rc = prf_hashrf_load(h, bips, tree_off, T, W, PRF_MEM_HOST, &st);
  if (rc) return rc;
  uint64_t P = ctx->h.P;
  prf_stats st2{};
  if ((rc = prf_hashrf_compute(h, 0, P, editdist, PRF_VARIANT_AUTO, nullptr, nullptr, &st2))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
  if (out_upper && (rc = prf_fetch(h, 0, P, out_upper))) return rc;
  if (out_mask && editdist >= 0 && (rc = prf_fetch_mask(h, 0, P, out_mask))) return rc;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  trace_mark("hashrf: fetched");
  if (stats) {
    *stats = st2;
    stats->ms_h2d = st.ms_h2d;
    stats->ms_dedup = st.ms_dedup;
    stats->ms_incidence = st.ms_incidence;
    stats->ms_d2h = ev_ms(ctx->ev[6], ctx->ev[7]);
    stats->ms_total = ev_ms(ctx->ev[0], ctx->ev[7]);
    stats->gpu_launches = st.gpu_launches + st2.gpu_launches;
  }
  return PRF_OK;
}

// ---------------------------------------------------------------------------------------------
// hashRF: multi-GPU build
// ---------------------------------------------------------------------------------------------

PRF_API int prf_hashrf_partition(prf_ctx* h, const uint64_t* d_bips, const uint64_t* d_off, uint32_t T_local, uint32_t W,
                                 uint32_t world, uint64_t* d_send_keys, uint32_t* d_send_counts, uint64_t* h_first) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (!d_off || !d_send_counts || !h_first) return ctx->fail(PRF_E_INVALID, "null pointer");
  if (W == 0 || W > kMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
  if (world == 0 || world > 16) return ctx->fail(PRF_E_UNSUPPORTED, "world=%u; supported 1..16", world);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint64_t nnz = 0;
  PRF_CK(ctx, cudaMemcpyAsync(&nnz, d_off + T_local, 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (nnz >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu entries; supported < 2^32 - 1", (unsigned long long)nnz);
  if (nnz && (!d_bips || !d_send_keys)) return ctx->fail(PRF_E_INVALID, "null key buffer");
  DevBuf<uint32_t> owner, idx;
  DevBuf<unsigned long long> first;
  PRF_CK(ctx, owner.alloc(nnz, s));
  PRF_CK(ctx, idx.alloc(nnz, s));
  PRF_CK(ctx, first.alloc((size_t)world + 1, s));
  if (T_local) {
    switch (W) {
#define PRF_W_CASE(w) case w: part_owner_kernel<w><<<(T_local + 7) / 8, 256, 0, s>>>(d_bips, d_off, T_local, world, owner.p, idx.p, d_send_counts); break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7) PRF_W_CASE(8)
      PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13) PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
    }
    PRF_LAUNCHED(ctx);
  }
  int rc = radix_sort_pairs(ctx, owner.p, idx.p, (uint32_t)nnz, std::max(1, bits_for(world)));  // stable: tree order kept per class
  if (rc) return rc;
  class_bounds_kernel<<<1, 32, 0, s>>>(owner.p, (uint32_t)nnz, world, first.p);
  PRF_LAUNCHED(ctx);
  if (nnz) {
    gather_keys_kernel<<<grid_for(nnz * W), 256, 0, s>>>(d_bips, idx.p, nnz * W, W, d_send_keys);
    PRF_LAUNCHED(ctx);
  }
  unsigned long long hf[17];
  PRF_CK(ctx, cudaMemcpyAsync(hf, first.p, ((size_t)world + 1) * 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  for (uint32_t c = 0; c <= world; ++c) h_first[c] = hf[c];
  return PRF_OK;
}

PRF_API int prf_hashrf_export(prf_ctx* h, prf_part* out) {
  if (!h || !out) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  HashRFState& H = ctx->h;
  if (!H.loaded) return ctx->fail(PRF_E_STATE, "prf_hashrf_export before a successful prf_hashrf_load");
  out->n_trees = H.T;
  out->n_ranges = H.n_ranges;
  out->n_members = H.n_lmembers;  // the walked lists, each padded to a multiple of 4 members
  out->n_lists = H.n_shared;
  out->nnz = H.stats.nnz;
  out->n_unique = H.stats.n_unique;
  out->n_hits = H.stats.n_hits;
  out->nb = H.nb.p;
  out->roff = H.roff.p;
  out->ranges = H.ranges.p;
  out->members = H.lmembers.p;
  out->lid = nullptr;
  return PRF_OK;
}

PRF_API int prf_hashrf_merge(prf_ctx* h, uint32_t T, uint32_t W, uint32_t world, const prf_part* parts, prf_stats* stats) {
  if (!h || !parts) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (world == 0 || world > 16) return ctx->fail(PRF_E_UNSUPPORTED, "world=%u; supported 1..16", world);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  uint64_t launches0 = ctx->launches;
  MergeParts P{};
  P.world = world;
  uint64_t n_members = 0, n_lists = 0, n_ranges = 0, nnz = 0, n_unique = 0, n_hits = 0;
  for (uint32_t r = 0; r < world; ++r) {
    const prf_part& q = parts[r];
    if (q.n_trees != T) return ctx->fail(PRF_E_INVALID, "part %u covers %llu trees, expected %u", r, (unsigned long long)q.n_trees, T);
    P.nb[r] = q.nb;
    P.roff[r] = q.roff;
    P.ranges[r] = static_cast<const uint2*>(q.ranges);
    P.members[r] = q.members;
    P.lid[r] = q.lid;
    P.member_base[r] = (uint32_t)n_members;
    P.list_base[r] = (uint32_t)n_lists;
    P.n_members[r] = (uint32_t)q.n_members;
    n_members += q.n_members;
    n_lists += q.n_lists;
    n_ranges += q.n_ranges;
    nnz += q.nnz;
    n_unique += q.n_unique;
    n_hits += q.n_hits;
  }
  if (n_members >= 0xffffffffull) return ctx->fail(PRF_E_UNSUPPORTED, "merged member count does not fit 32 bits");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[2], s));
  // build into fresh buffers: a part may alias the context's current structure
  DevBuf<uint16_t> nb;
  DevBuf<uint32_t> roff, members, lid, nb32, cnt, mm, soff, singles;  // mm: min |B|, max |B|, max ranges per tree
  DevBuf<uint2> ranges;
  PRF_CK(ctx, soff.alloc((size_t)T + 1, s));  // the parts carry every list as ranges (built with short_max == 0)
  PRF_CK(ctx, cudaMemsetAsync(soff.p, 0, soff.bytes(), s));
  PRF_CK(ctx, singles.alloc(1, s));
  PRF_CK(ctx, nb.alloc((size_t)T + 16, s));
  PRF_CK(ctx, cudaMemsetAsync(nb.p, 0, nb.bytes(), s));
  PRF_CK(ctx, roff.alloc((size_t)T + 1, s));
  PRF_CK(ctx, members.alloc((size_t)n_members + 4, s));
  PRF_CK(ctx, ranges.alloc(n_ranges, s));
  PRF_CK(ctx, nb32.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cnt.alloc((size_t)T + 1, s));
  PRF_CK(ctx, mm.alloc(3, s));
  const uint32_t mm_init[3] = {0xffffffffu, 0u, 0u};
  PRF_CK(ctx, cudaMemcpyAsync(mm.p, mm_init, 12, cudaMemcpyHostToDevice, s));
  merge_count_kernel<<<grid_for((uint64_t)T + 1), 256, 0, s>>>(P, T, nb32.p, cnt.p);
  PRF_LAUNCHED(ctx);
  int rc = exclusive_scan(ctx, LoadU32{cnt.p}, (uint64_t)T + 1, roff.p, nullptr);
  if (rc) return rc;
  if (T) {
    merge_ranges_kernel<<<grid_for((uint64_t)world * T), 256, 0, s>>>(P, T, roff.p, ranges.p);
    PRF_LAUNCHED(ctx);
    finish_nb_kernel<<<grid_for(T), 256, 0, s>>>(nb32.p, T, nb.p, mm.p, roff.p);
    PRF_LAUNCHED(ctx);
  }
  for (uint32_t r = 0; r < world; ++r)
    if (P.n_members[r]) {
      merge_members_kernel<<<grid_for(P.n_members[r]), 256, 0, s>>>(P, r, members.p);
      PRF_LAUNCHED(ctx);
    }
  uint32_t hmm[3] = {0, 0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(hmm, mm.p, 12, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (T == 0) hmm[0] = hmm[1] = 0;
  if (hmm[1] > kMaxBipsPerTree)
    return ctx->fail(PRF_E_RANGE, "a tree has %u bipartitions; distances would overflow uint16 (max %u per tree)", hmm[1],
                     kMaxBipsPerTree);
  HashRFState& H = ctx->h;
  H = HashRFState{};  // frees the previous structure (stream-ordered, after the kernels above)
  H.T = T;
  H.W = W;
  H.nnz = nnz;
  H.P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  H.uniform_nb = hmm[0] == hmm[1];
  H.nb_uniform = hmm[0];
  H.nb = std::move(nb);
  H.roff = std::move(roff);
  H.ranges = std::move(ranges);
  H.lmembers = std::move(members);  // (the merged structure has no unpadded incidence: the dense variant is not available on it)
  H.n_lmembers = n_members;
  H.soff = std::move(soff);
  H.singles = std::move(singles);
  // the merged ranges of a tree interleave the parts: all of them take the warp-per-list walk
  PRF_CK(ctx, H.nheavy.alloc((size_t)T + 1, s));
  PRF_CK(ctx, cudaMemcpyAsync(H.nheavy.p, cnt.p, ((size_t)T + 1) * 4, cudaMemcpyDeviceToDevice, s));
  H.n_shared = (uint32_t)n_lists;
  H.n_shared_nnz = n_members;
  H.n_ranges = n_ranges;
  H.max_ranges = hmm[2];
  prf_stats& st = H.stats;
  st = prf_stats{};
  st.n_trees = T;
  st.n_words = W;
  st.nnz = nnz;
  st.n_unique = n_unique;
  st.n_shared = n_lists;
  st.n_shared_nnz = n_members;
  st.n_hits = n_hits;
  st.n_pairs = H.P;
  st.max_bips = hmm[1];
  st.min_bips = hmm[0];
  st.ms_incidence = ev_ms(ctx->ev[2], ctx->ev[3]);
  st.gpu_launches = (int32_t)(ctx->launches - launches0);
  if ((rc = build_nb_shift(ctx))) return rc;
  H.loaded = true;
  ctx->tol.loaded = false;
  ctx->out_valid = ctx->mask_valid = false;
  if (stats) *stats = st;
  return PRF_OK;
}

// ---------------------------------------------------------------------------------------------
// Debug entry points for tests of the primitives (not part of phybin_rf.h).
// ---------------------------------------------------------------------------------------------

// Host-only: the work items launch_sparse would cut [p_begin, p_end) into (tests of the plan builder;
// needs no device).  items: cap x 4 uint32 (a0, a1, cb, ce).  Returns the number of items or -1.
PRF_API long long prf_debug_row_plan(uint32_t T, uint64_t p_begin, uint64_t p_end, uint32_t tile, uint32_t* items,
                                     uint64_t cap) {
  std::vector<RowItem> rows, groups;
  build_row_plan(T, p_begin, p_end, tile, rows, groups);
  rows.insert(rows.end(), groups.begin(), groups.end());
  if (rows.size() > cap) return -1;
  for (size_t i = 0; i < rows.size(); ++i) {
    items[4 * i] = rows[i].a0;
    items[4 * i + 1] = rows[i].a1;
    items[4 * i + 2] = rows[i].cb;
    items[4 * i + 3] = rows[i].ce;
  }
  return (long long)rows.size();
}

PRF_API int prf_debug_wide_slots(prf_ctx* h, int on) {
  if (!h) return PRF_E_INVALID;
  h->c.debug_wide_slots = on != 0;
  return PRF_OK;
}

PRF_API int prf_debug_no_flip(prf_ctx* h, int on) {
  if (!h) return PRF_E_INVALID;
  h->c.debug_no_flip = on != 0;
  return PRF_OK;
}

PRF_API int prf_debug_short_max(prf_ctx* h, uint32_t short_max) {
  if (!h) return PRF_E_INVALID;
  h->c.short_max = short_max;
  return PRF_OK;
}

// Lists of at least this many members take the warp-per-list walk (0 = chosen from T).
PRF_API int prf_debug_heavy_min(prf_ctx* h, uint32_t heavy_min) {
  if (!h) return PRF_E_INVALID;
  h->c.heavy_min = heavy_min;
  return PRF_OK;
}

// which: 0 = rows kernel (one warp per row), 1 = round-1 CTA-per-row kernel (loads must then use short_max 0)
PRF_API int prf_debug_sparse_kernel(prf_ctx* h, int which, int cfg) {
  if (!h) return PRF_E_INVALID;
  h->c.sparse_kernel = which;
  h->c.rows_cfg = cfg;
  return PRF_OK;
}

#ifdef PRF_CALIB
// Calibration builds only (scripts/rows_calib.py; never in libphybin_rf.so): wrong results by design.
PRF_API int prf_debug_calib(prf_ctx* h, int mode) {
  if (!h) return PRF_E_INVALID;
  h->c.calib = mode;
  return PRF_OK;
}
#endif

PRF_API int prf_debug_small_tiles(prf_ctx* h, int on) {
  if (!h) return PRF_E_INVALID;
  h->c.debug_small_tiles = on != 0;
  return PRF_OK;
}

// PRF_TIMING builds: copies the per-CTA cycle counters of the last sparse launch ([grid][64]).
PRF_API int prf_debug_timing(prf_ctx* h, unsigned long long* dst, uint32_t cap, uint32_t* grid) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  if (grid) *grid = ctx->timing_grid;
  size_t n = std::min<size_t>((size_t)ctx->timing_grid * 64, cap);
  if (n && ctx->timing.p) {
    PRF_CK(ctx, cudaMemcpyAsync(dst, ctx->timing.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return PRF_OK;
}

PRF_API int prf_debug_scan(prf_ctx* h, const uint32_t* host_in, uint64_t n, uint32_t* host_out, uint32_t* total) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  cudaStream_t s = ctx->stream;
  DevBuf<uint32_t> in, out, tot;
  PRF_CK(ctx, in.alloc(n, s));
  PRF_CK(ctx, out.alloc(n, s));
  PRF_CK(ctx, tot.alloc(1, s));
  PRF_CK(ctx, cudaMemcpyAsync(in.p, host_in, n * 4, cudaMemcpyHostToDevice, s));
  int rc = exclusive_scan(ctx, LoadU32{in.p}, n, out.p, tot.p);
  if (rc) return rc;
  PRF_CK(ctx, cudaMemcpyAsync(host_out, out.p, n * 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaMemcpyAsync(total, tot.p, 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  return PRF_OK;
}

PRF_API int prf_debug_sort(prf_ctx* h, uint32_t* host_keys, uint32_t* host_vals, uint32_t n, int nbits) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  cudaStream_t s = ctx->stream;
  DevBuf<uint32_t> k, v;
  PRF_CK(ctx, k.alloc(n, s));
  PRF_CK(ctx, v.alloc(n, s));
  PRF_CK(ctx, cudaMemcpyAsync(k.p, host_keys, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  PRF_CK(ctx, cudaMemcpyAsync(v.p, host_vals, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  int rc = radix_sort_pairs(ctx, k.p, v.p, n, nbits);
  if (rc) return rc;
  PRF_CK(ctx, cudaMemcpyAsync(host_keys, k.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaMemcpyAsync(host_vals, v.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  return PRF_OK;
}

}  // extern "C"
