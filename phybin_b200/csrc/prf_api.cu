// prf_api.cu -- the C ABI of include/phybin_rf.h.  One translation unit: the kernels live in the
// .cuh files next to this one.  Built for sm_100a only (see phybin_b200/build.py); there is no CPU
// path here and no dependency on the oracle.

#include <algorithm>
#include <cstring>
#include <new>

#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"

using namespace prf;

static thread_local std::string g_create_error;

// ext[0] = tree_off[T]; ext[1] = most entries of one tree; ext[2] != 0 when the offsets do not start at 0 or descend
__global__ void __launch_bounds__(256) tree_extent_kernel(const uint64_t* __restrict__ off, uint32_t T, unsigned long long* ext) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long m = 0;
  if (t == T) {
    ext[0] = off[T];
    if (off[0] != 0) ext[2] = 1;
  } else if (t < T) {
    const uint64_t a = off[t], b = off[t + 1];
    if (b < a) ext[2] = 1;
    else m = b - a;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&ext[1], m);
}


namespace prf {
__global__ void __launch_bounds__(1024) wordop_peak_kernel(unsigned long long* sink, uint32_t iters, uint64_t seed) {
  uint64_t x[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed * (2 * i + 1) + threadIdx.x; acc[i] = 0; }
  uint64_t m = seed ^ blockIdx.x;
  for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __popcll(x[i] & m);
    m += 0x9E3779B97F4A7C15ull;
  }
  uint64_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += acc[i];
  if (t == 0x7fffffffffffffffull) *sink = t;  // keeps the loop alive
}
}  // namespace prf

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
  team_release_ctx(&c);
  c.h = HashRFState{};
  c.tol = TolerantState{};
  c.ext = ExtractState{};
  c.cons = ConsensusState{};
  c.plan = RowPlan{};
  c.timing.release();
  c.scan_status.release();
  c.scan_counter.release();
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
  // offsets first: nnz = tree_off[T], and the most entries of one tree (sizes the per-tree hash sets of the build)
  DevBuf<uint64_t> d_off_own, d_bips_own;
  const uint64_t* d_off = tree_off;
  uint64_t nnz = 0;
  uint32_t max_entries = 0;
  if (kind == PRF_MEM_HOST) {
    if (tree_off[0] != 0) return ctx->fail(PRF_E_INVALID, "tree_off[0] must be 0");
    for (uint32_t t = 0; t < T; ++t) {
      if (tree_off[t + 1] < tree_off[t]) return ctx->fail(PRF_E_INVALID, "tree_off is not ascending at tree %u", t);
      max_entries = (uint32_t)std::min<uint64_t>(0xffffffffull, std::max<uint64_t>(max_entries, tree_off[t + 1] - tree_off[t]));
    }
    nnz = tree_off[T];
    PRF_CK(ctx, d_off_own.alloc((size_t)T + 1, s));
    int rc = upload(ctx, d_off_own.p, tree_off, ((size_t)T + 1) * 8, kind);
    if (rc) return rc;
    d_off = d_off_own.p;
  } else {
    DevBuf<unsigned long long> ext;  // [0] = tree_off[T], [1] = max entries of a tree, [2] = offsets not ascending
    PRF_CK(ctx, ext.alloc(3, s));
    PRF_CK(ctx, cudaMemsetAsync(ext.p, 0, 24, s));
    tree_extent_kernel<<<grid_for((uint64_t)T + 1), 256, 0, s>>>(tree_off, T, ext.p);
    PRF_LAUNCHED(ctx);
    unsigned long long hext[3] = {0, 0, 0};
    PRF_CK(ctx, cudaMemcpyAsync(hext, ext.p, 24, cudaMemcpyDeviceToHost, s));
    PRF_CK(ctx, cudaStreamSynchronize(s));
    if (hext[2]) return ctx->fail(PRF_E_INVALID, "tree_off must start at 0 and ascend");
    nnz = hext[0];
    max_entries = (uint32_t)std::min<unsigned long long>(0xffffffffull, hext[1]);
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
  int rc = hashrf_build(ctx, d_bips, d_off, d_off + 1, T, W, nnz, nnz, max_entries);
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
    const double kpad = (double)dense_kpad(H.n_list_ids);
    const double t_sparse = hits * 1.9e-12 + (H.uniform_nb ? 0.0 : pairs * 0.9e-12), t_dense = pairs * kpad * 2.0 / 0.9e15;
    const double operand = (double)dense_tpad(H.T) * kpad;
    use = (H.n_shared > 0 && H.ent_lid.p && operand < 32e9 && t_dense < t_sparse) ? PRF_VARIANT_DENSE : PRF_VARIANT_SPARSE;
  }
  PRF_CK(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
  if (p_end > p_begin) {
    if (use == PRF_VARIANT_DENSE && !H.ent_lid.p && H.n_shared)
      return ctx->fail(PRF_E_UNSUPPORTED, "the dense variant needs the incidence of a single-context build");
    if (use == PRF_VARIANT_DENSE) rc = launch_dense(ctx, p_begin, p_end, editdist, d_out, d_mask);
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

PRF_API int prf_debug_agglo_ctas(prf_ctx* h, uint32_t n) {
  if (!h) return PRF_E_INVALID;
  h->c.debug_agglo_ctas = n;
  return PRF_OK;
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

// cfg: tuning instantiation of the rows kernel (prf_rows.cuh launch_rows)
PRF_API int prf_debug_rows_cfg(prf_ctx* h, int cfg) {
  if (!h) return PRF_E_INVALID;
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

// Measured integer "word-op" peak of the device (the tolerant mode's roofline denominator): one word-op =
// AND + popcount + accumulate of one 64-bit word, 8 independent chains per thread.  Returns word-ops per second.
PRF_API int prf_debug_wordop_peak(prf_ctx* h, double* wordops_per_s) {
  if (!h || !wordops_per_s) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  DevBuf<unsigned long long> sink;
  PRF_CK(ctx, sink.alloc(1, s));
  const uint32_t grid = (uint32_t)ctx->sm_count * 2, iters = 1 << 14;
  wordop_peak_kernel<<<grid, 1024, 0, s>>>(sink.p, iters, 0x0123456789abcdefull);  // warm-up
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));
  wordop_peak_kernel<<<grid, 1024, 0, s>>>(sink.p, iters, 0x0123456789abcdefull);
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  ctx->launches += 2;
  const double ms = ev_ms(ctx->ev[0], ctx->ev[1]);
  *wordops_per_s = ms > 0 ? (double)grid * 1024 * iters * 8 / (ms * 1e-3) : 0.0;
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
