// prf_team.cu -- multi-GPU hashRF build: the prf_team_* entry points (one context per GPU, in one process or one
// process per GPU) and prf_multi_* (single process, all GPUs of the box, host buffers in and out).  Kernels and
// the arena layout: prf_team.cuh.
#include <condition_variable>
#include <mutex>
#include <thread>

#include "prf_common.cuh"
#include "prf_internal.h"
#include "prf_primitives.cuh"
#include "prf_team.cuh"

using namespace prf;

namespace prf {

// Single-process teams: the ranks' host threads meet here, then every stream waits for every rank's event.
struct TeamShared {
  std::mutex mu;
  std::condition_variable cv;
  uint32_t world = 0, arrived = 0;
  uint64_t generation = 0;
  cudaEvent_t ev[kMaxParts] = {};
  int failed = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t g = generation;
    if (++arrived == world) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else if (!cv.wait_for(lk, std::chrono::seconds(60), [&] { return generation != g; })) {
      failed = 1;  // a rank never arrived (it failed before the barrier): do not wait for ever
    }
  }
};

void team_release_ctx(Ctx* ctx) {
  Team& tm = ctx->team;
  if (!tm.active) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (uint32_t p = 0; p < tm.world; ++p)
    if (tm.ipc_open[p] && tm.peer[p]) cudaIpcCloseMemHandle(tm.peer[p]);
  if (tm.arena) cudaFree(tm.arena);
  if (tm.ev) cudaEventDestroy(tm.ev);
  tm.cls_cnt.release();
  tm.cls_pos.release();
  tm.err.release();
  tm.cls.release();
  tm.peers_dev.release();
  cudaGetLastError();
  tm = Team{};
}

static int team_barrier(Ctx* ctx, const uint32_t* payload = nullptr) {
  Team& tm = ctx->team;
  cudaStream_t s = ctx->stream;
  ++tm.epoch;
  if (tm.shared) {  // single process: events + a host rendezvous of the ranks' threads
    PRF_CK(ctx, cudaEventRecord(tm.ev, s));
    if (payload) {  // same delivery as the flag kernel's, by a stream-ordered kernel before the event
      team_payload_kernel<<<1, 32, 0, s>>>(tm.peers_dev.p, tm.lay.off_flags, tm.rank, tm.world, payload);
      PRF_LAUNCHED(ctx);
      PRF_CK(ctx, cudaEventRecord(tm.ev, s));
    }
    tm.shared->wait();  // every rank has recorded its event
    for (uint32_t p = 0; p < tm.world; ++p)
      if (p != tm.rank) PRF_CK(ctx, cudaStreamWaitEvent(s, tm.shared->ev[p], 0));
    tm.shared->wait();  // nobody re-records its event before everybody has queued its waits
  } else {
    team_barrier_kernel<<<1, 32, 0, s>>>(tm.peers_dev.p, tm.lay.off_flags, tm.rank, tm.world, tm.epoch, payload, tm.err.p);
    PRF_LAUNCHED(ctx);
  }
  return PRF_OK;
}

}  // namespace prf

extern "C" {

PRF_API size_t prf_team_handle_bytes(void) { return kTeamHandleBytes; }

PRF_API int prf_team_init(prf_ctx* h, uint32_t rank, uint32_t world, uint32_t T, uint32_t W, uint64_t nnz_total,
                          uint64_t max_local_entries, uint32_t max_tree_entries, void* handle_out) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  if (world == 0 || world > (uint32_t)kMaxParts || rank >= world) return ctx->fail(PRF_E_INVALID, "rank %u of %u (at most %d ranks)", rank, world, kMaxParts);
  if (W == 0 || W > kMaxWords) return ctx->fail(PRF_E_UNSUPPORTED, "W=%u words per set; supported 1..%u", W, kMaxWords);
  if (nnz_total >= (1ull << 32) - 1) return ctx->fail(PRF_E_UNSUPPORTED, "nnz=%llu entries; supported < 2^32 - 1", (unsigned long long)nnz_total);
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  team_release_ctx(ctx);
  ctx->h.loaded = false;
  Team& tm = ctx->team;
  tm.rank = rank;
  tm.world = world;
  tm.max_entries = max_tree_entries;
  tm.nnz_total = nnz_total;
  const uint32_t log2B = ctx->debug_small_tiles ? 8 : 13;
  const uint32_t NB = T ? ((T - 1) >> log2B) + 1 : 1;
  tm.lay.compute(world, T, W, NB, nnz_total, std::max<uint64_t>(max_local_entries, 1), ctx->short_max);
  if ((uint64_t)world * tm.lay.cap_e >= (1ull << 32) - 1)
    return ctx->fail(PRF_E_UNSUPPORTED, "key regions of %u x %llu entries do not fit 32-bit entry indices", world, (unsigned long long)tm.lay.cap_e);
  cudaError_t e = cudaMalloc((void**)&tm.arena, tm.lay.bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return ctx->fail(PRF_E_NOMEM, "team arena of %zu bytes: %s", tm.lay.bytes, cudaGetErrorString(e));
  }
  PRF_CK(ctx, cudaMemsetAsync(tm.arena, 0, tm.lay.off_keys, ctx->stream));  // flags
  PRF_CK(ctx, tm.err.alloc(4, ctx->stream));
  PRF_CK(ctx, cudaMemsetAsync(tm.err.p, 0, 16, ctx->stream));
  PRF_CK(ctx, tm.peers_dev.alloc(kMaxParts, ctx->stream));
  PRF_CK(ctx, cudaEventCreateWithFlags(&tm.ev, cudaEventDisableTiming));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  tm.peer[rank] = tm.arena;
  tm.active = true;
  if (handle_out) {
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == kTeamHandleBytes, "cudaIpcMemHandle_t size");
    PRF_CK(ctx, cudaIpcGetMemHandle(&mh, tm.arena));
    std::memcpy(handle_out, &mh, sizeof mh);
  }
  if (world == 1) {  // a team of one needs no peers
    const uint64_t me = reinterpret_cast<uint64_t>(tm.arena);
    PRF_CK(ctx, cudaMemcpy(tm.peers_dev.p, &me, 8, cudaMemcpyHostToDevice));
    tm.connected = true;
  }
  return PRF_OK;
}

// Multi-process: `handles` = the world handles in rank order (every rank's prf_team_init output, exchanged by the caller).
PRF_API int prf_team_connect(prf_ctx* h, const void* handles) {
  if (!h || !handles) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  Team& tm = ctx->team;
  if (!tm.active) return ctx->fail(PRF_E_STATE, "prf_team_connect before prf_team_init");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  uint64_t addr[kMaxParts] = {};
  for (uint32_t p = 0; p < tm.world; ++p) {
    if (p != tm.rank && !tm.peer[p]) {
      cudaIpcMemHandle_t mh;
      std::memcpy(&mh, static_cast<const char*>(handles) + (size_t)p * kTeamHandleBytes, sizeof mh);
      void* ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx->fail(PRF_E_CUDA, "cudaIpcOpenMemHandle for rank %u: %s", p, cudaGetErrorString(e));
      }
      tm.peer[p] = static_cast<uint8_t*>(ptr);
      tm.ipc_open[p] = true;
    }
    addr[p] = reinterpret_cast<uint64_t>(tm.peer[p]);
  }
  PRF_CK(ctx, cudaMemcpy(tm.peers_dev.p, addr, sizeof addr, cudaMemcpyHostToDevice));
  tm.connected = true;
  return PRF_OK;
}

}  // extern "C"

namespace prf {

// Single process: the contexts of all ranks see each other's arenas directly.
static int team_connect_local(prf_ctx* const* hs, uint32_t world, TeamShared* shared) {
  for (uint32_t r = 0; r < world; ++r) {
    Ctx* ctx = &hs[r]->c;
    Team& tm = ctx->team;
    PRF_CK(ctx, cudaSetDevice(ctx->device));
    uint64_t addr[kMaxParts] = {};
    for (uint32_t p = 0; p < world; ++p) {
      Ctx* other = &hs[p]->c;
      if (other->device != ctx->device) {
        int can = 0;
        PRF_CK(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, other->device));
        if (!can) return ctx->fail(PRF_E_UNSUPPORTED, "device %d cannot access device %d", ctx->device, other->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return ctx->fail(PRF_E_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
      tm.peer[p] = other->team.arena;
      addr[p] = reinterpret_cast<uint64_t>(tm.peer[p]);
    }
    PRF_CK(ctx, cudaMemcpy(tm.peers_dev.p, addr, sizeof addr, cudaMemcpyHostToDevice));
    tm.shared = shared;
    shared->ev[r] = tm.ev;
    tm.connected = true;
  }
  shared->world = world;
  return PRF_OK;
}

// Collective: every rank calls it with the canonical bipartitions of ITS trees (device memory, laid out like
// prf_hashrf_load's input; rank r owns trees [T r / world, T (r + 1) / world)).
static int team_load(Ctx* ctx, const uint64_t* d_bips, const uint64_t* d_off, uint32_t Tl, uint64_t nnz_local, prf_stats* stats) {
  Team& tm = ctx->team;
  const TeamLayout& L = tm.lay;
  cudaStream_t s = ctx->stream;
  const uint32_t world = tm.world, rank = tm.rank, T = L.T, W = L.W;
  const uint32_t t_lo = (uint32_t)((uint64_t)T * rank / world), t_hi = (uint32_t)((uint64_t)T * (rank + 1) / world);
  if (Tl != t_hi - t_lo) return ctx->fail(PRF_E_INVALID, "rank %u of %u owns %u trees, got %u", rank, world, t_hi - t_lo, Tl);
  if (nnz_local > L.cap_e) return ctx->fail(PRF_E_INVALID, "%llu local entries exceed the %llu announced at prf_team_init", (unsigned long long)nnz_local, (unsigned long long)L.cap_e);
  const uint64_t launches0 = ctx->launches;
  ctx->h.loaded = false;
  ctx->tol.loaded = false;
  ctx->out_valid = ctx->mask_valid = false;
  PRF_CK(ctx, cudaEventRecord(ctx->ev[0], s));

  // ---- my keys to their owners: one kernel (classes, reserved ranges, stores into the peers' arenas) ----
  if (tm.cls_cnt.n < (size_t)kMaxParts) PRF_CK(ctx, tm.cls_cnt.alloc(kMaxParts, s));  // per-class cursors of my key regions
  PRF_CK(ctx, cudaMemsetAsync(tm.cls_cnt.p, 0, kMaxParts * 4, s));
  if (Tl) {
    switch (W) {
#define PRF_W_CASE(w) case w: team_partition_kernel<w><<<(Tl + 7) / 8, 256, 0, s>>>(d_bips, d_off, Tl, t_lo, world, rank, tm.peers_dev.p, L, tm.cls_cnt.p, tm.err.p); break;
      PRF_W_CASE(1) PRF_W_CASE(2) PRF_W_CASE(3) PRF_W_CASE(4) PRF_W_CASE(5) PRF_W_CASE(6) PRF_W_CASE(7) PRF_W_CASE(8)
      PRF_W_CASE(9) PRF_W_CASE(10) PRF_W_CASE(11) PRF_W_CASE(12) PRF_W_CASE(13) PRF_W_CASE(14) PRF_W_CASE(15) PRF_W_CASE(16)
#undef PRF_W_CASE
      default: return ctx->fail(PRF_E_UNSUPPORTED, "W=%u", W);
    }
    PRF_LAUNCHED(ctx);
  }
  trace_sync(s, "team:   partition kernel");
  int rc;
  if ((rc = team_barrier(ctx, tm.cls_cnt.p))) return rc;  // (+ how many entries every class got from me)
  trace_sync(s, "team:   barrier 1");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[1], s));
  trace_mark("team: keys pushed, barrier queued");

  // ---- the ordinary build on my class (keys, tb, te now sit in my arena) ----
  const uint64_t* keys = reinterpret_cast<const uint64_t*>(tm.arena + L.off_keys);
  const uint64_t* tb = reinterpret_cast<const uint64_t*>(tm.arena + L.off_tb);
  const uint64_t* te = reinterpret_cast<const uint64_t*>(tm.arena + L.off_te);
  // entries every source rank sent me: the dense segments of my key space, and the size of the dedup table
  unsigned long long seg_cnt[kMaxParts] = {};
  PRF_CK(ctx, cudaMemcpyAsync(seg_cnt, tm.arena + L.off_flags + 2048, (size_t)world * 8, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  uint64_t seg_begin[kMaxParts], seg_len[kMaxParts], class_entries = 0;
  for (uint32_t p = 0; p < world; ++p) {
    seg_begin[p] = (uint64_t)p * L.cap_e;
    seg_len[p] = seg_cnt[p];
    class_entries += seg_cnt[p];
  }
  trace_mark("team: class size known");
  if ((rc = hashrf_build(ctx, keys, tb, te, T, W, (uint64_t)world * L.cap_e, class_entries, tm.max_entries, seg_begin, seg_len, world))) {
    if (tm.shared) tm.shared->failed = 1;
    // keep the peers' barriers matched: fall through with an error part
  }
  HashRFState& H = ctx->h;
  const int build_rc = rc;
  const std::string build_err = ctx->err;

  // ---- my part into every peer's arena ----
  TeamPartHeader hd{};
  hd.err = build_rc ? 1u : 0u;
  if (!build_rc) {
    hd.n_unique = H.stats.n_unique;
    hd.n_hits = H.stats.n_hits;
    hd.n_entries = H.n_shared_nnz;
    hd.nnz = nnz_local;
    hd.n_lists2 = H.n_shared;
    hd.n_long = H.n_long;
    hd.n_flipped = H.n_flipped;
    hd.n_rlist = (uint32_t)H.rlist_n;
    hd.n_boff = (uint32_t)((uint64_t)H.n_long * L.NB + 1);
    hd.n_lmem = (uint32_t)H.n_lmem;
    hd.n_singles = (uint32_t)H.singles_n;
    hd.max_lists = H.max_lists;
    if (hd.n_rlist > L.cap_r || hd.n_boff > L.cap_b || hd.n_lmem > L.cap_l || hd.n_singles > L.cap_s) hd.err = 2;
  }
  DevBuf<TeamPartHeader> d_hd;
  PRF_CK(ctx, d_hd.alloc(1, s));
  PRF_CK(ctx, cudaMemcpyAsync(d_hd.p, &hd, sizeof hd, cudaMemcpyHostToDevice, s));
  TeamCopies cp{};
  auto add = [&](const void* src, size_t dst_off, size_t bytes) {
    if (bytes) cp.c[cp.n++] = TeamCopy{src, dst_off, bytes};
  };
  add(d_hd.p, 0, sizeof hd);
  if (!hd.err) {
    add(H.nb.p, L.p_nb, ((size_t)T + 16) * 2);
    add(H.roff.p, L.p_roff, ((size_t)T + 1) * 4);
    add(H.soff.p, L.p_soff, ((size_t)T + 1) * 4);
    add(H.rlist.p, L.p_rlist, (size_t)hd.n_rlist * 4);
    add(H.boff.p, L.p_boff, (size_t)hd.n_boff * 4);
    add(H.lmem.p, L.p_lmem, (size_t)hd.n_lmem * 2);
    add(H.singles.p, L.p_singles, (size_t)hd.n_singles * 4);
  }
  team_push_part_kernel<<<ctx->sm_count * 2, 256, 0, s>>>(cp, tm.peers_dev.p, L.off_parts + (size_t)rank * L.part_bytes, world);
  PRF_LAUNCHED(ctx);
  trace_sync(s, "team:   push part kernel");
  if ((rc = team_barrier(ctx))) return rc;
  trace_sync(s, "team:   barrier 2");
  PRF_CK(ctx, cudaEventRecord(ctx->ev[2], s));
  trace_mark("team: part pushed, barrier queued");

  // ---- every part is here: |B_t| = sum over the parts, install the parts as the loaded problem ----
  DevBuf<uint16_t> nb;
  DevBuf<uint32_t> mm;
  PRF_CK(ctx, nb.alloc((size_t)T + 16, s));
  PRF_CK(ctx, cudaMemsetAsync(nb.p, 0, nb.bytes(), s));
  PRF_CK(ctx, mm.alloc(4, s));
  const uint32_t mm0[4] = {0xffffffffu, 0u, 0u, 0u};
  PRF_CK(ctx, cudaMemcpyAsync(mm.p, mm0, 16, cudaMemcpyHostToDevice, s));
  std::vector<TeamPartHeader> hds(world);
  for (uint32_t p = 0; p < world; ++p)
    PRF_CK(ctx, cudaMemcpyAsync(&hds[p], tm.arena + L.off_parts + (size_t)p * L.part_bytes, sizeof(TeamPartHeader), cudaMemcpyDeviceToHost, s));
  uint32_t herr[4] = {0, 0, 0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(herr, tm.err.p, 16, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaStreamSynchronize(s));  // headers first: an error part has no arrays
  if (herr[0]) return ctx->fail(PRF_E_CUDA, herr[0] == 2 ? "team barrier timed out (a peer did not arrive)" : "team key exchange overflowed a region");
  if (build_rc) return ctx->fail(build_rc, "%s", build_err.c_str());
  for (uint32_t p = 0; p < world; ++p)
    if (hds[p].err)
      return ctx->fail(hds[p].err == 2 ? PRF_E_NOMEM : PRF_E_CUDA,
                       hds[p].err == 2 ? "hash class %u does not fit its team region (very uneven classes); use fewer ranks" : "rank %u failed to build its class", p);
  if (T) {
    team_finish_kernel<<<grid_for(T), 256, 0, s>>>(tm.arena, L, nb.p, mm.p);
    PRF_LAUNCHED(ctx);
  }
  uint32_t hmm[4] = {0, 0, 0, 0};
  PRF_CK(ctx, cudaMemcpyAsync(hmm, mm.p, 16, cudaMemcpyDeviceToHost, s));
  PRF_CK(ctx, cudaEventRecord(ctx->ev[3], s));
  PRF_CK(ctx, cudaStreamSynchronize(s));
  if (T == 0) hmm[0] = hmm[1] = 0;
  if (hmm[1] > kMaxBipsPerTree)
    return ctx->fail(PRF_E_RANGE, "a tree has %u bipartitions; distances would overflow uint16 (max %u per tree)", hmm[1], kMaxBipsPerTree);

  const uint32_t block_log2 = H.block_log2, NB = H.NB;
  H = HashRFState{};  // drops the class-local structure (stream-ordered, after the push kernel)
  H.T = T;
  H.W = W;
  H.nnz = tm.nnz_total;
  H.P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  H.block_log2 = block_log2;
  H.NB = NB;
  H.uniform_nb = hmm[0] == hmm[1];
  H.nb_uniform = hmm[0];
  H.nb = std::move(nb);
  H.max_lists = hmm[2];
  H.n_parts = world;
  prf_stats& st = H.stats;
  st = prf_stats{};
  for (uint32_t p = 0; p < world; ++p) {
    const uint8_t* reg = tm.arena + L.off_parts + (size_t)p * L.part_bytes;
    H.part[p] = ListPart{reinterpret_cast<const uint32_t*>(reg + L.p_roff), reinterpret_cast<const uint32_t*>(reg + L.p_rlist),
                         reinterpret_cast<const uint32_t*>(reg + L.p_boff), reinterpret_cast<const uint16_t*>(reg + L.p_lmem),
                         reinterpret_cast<const uint32_t*>(reg + L.p_soff), reinterpret_cast<const uint32_t*>(reg + L.p_singles)};
    st.n_unique += hds[p].n_unique;
    st.n_hits += hds[p].n_hits;
    st.n_shared += hds[p].n_lists2;
    st.n_shared_nnz += hds[p].n_entries;
    st.nnz += hds[p].nnz;
    st.n_flipped += hds[p].n_flipped;
    H.n_long += hds[p].n_long;
  }
  H.n_shared = (uint32_t)st.n_shared;
  H.n_shared_nnz = st.n_shared_nnz;
  H.n_flipped = st.n_flipped;
  st.n_trees = T;
  st.n_words = W;
  st.n_pairs = H.P;
  st.max_bips = hmm[1];
  st.min_bips = hmm[0];
  st.ms_h2d = 0.f;
  st.ms_dedup = ev_ms(ctx->ev[0], ctx->ev[1]);      // partition + key exchange
  st.ms_incidence = ev_ms(ctx->ev[1], ctx->ev[3]);  // class build + part exchange
  st.ms_total = ev_ms(ctx->ev[0], ctx->ev[3]);
  st.gpu_launches = (int32_t)(ctx->launches - launches0);
  if ((rc = build_nb_shift(ctx))) return rc;
  H.loaded = true;
  trace_mark("team: parts installed");
  if (stats) *stats = st;
  return PRF_OK;
}

}  // namespace prf

extern "C" {

PRF_API int prf_team_load(prf_ctx* h, const uint64_t* d_bips, const uint64_t* d_tree_off, uint32_t T_local, prf_stats* stats) {
  if (!h) return PRF_E_INVALID;
  Ctx* ctx = &h->c;
  ctx->err.clear();
  Team& tm = ctx->team;
  if (!tm.active || !tm.connected) return ctx->fail(PRF_E_STATE, "prf_team_load before prf_team_init / prf_team_connect");
  if (!d_tree_off) return ctx->fail(PRF_E_INVALID, "tree_off is NULL");
  PRF_CK(ctx, cudaSetDevice(ctx->device));
  uint64_t nnz_local = 0;
  PRF_CK(ctx, cudaMemcpyAsync(&nnz_local, d_tree_off + T_local, 8, cudaMemcpyDeviceToHost, ctx->stream));
  PRF_CK(ctx, cudaStreamSynchronize(ctx->stream));
  if (nnz_local && (!d_bips || (reinterpret_cast<uintptr_t>(d_bips) & 15))) return ctx->fail(PRF_E_INVALID, "device bips pointer must be non-null and 16-byte aligned");
  return team_load(ctx, d_bips, d_tree_off, T_local, nnz_local, stats);
}

PRF_API void prf_team_release(prf_ctx* h) {
  if (h) team_release_ctx(&h->c);
}

// ---------------------------------------------------------------------------------------------
// prf_multi: one process, every GPU of the box, host buffers in and out
// ---------------------------------------------------------------------------------------------

struct prf_multi {
  uint32_t world = 0;
  prf_ctx* ctx[kMaxParts] = {};
  TeamShared shared;
  std::string err;
  // team geometry the arenas were sized for
  uint32_t T = 0, W = 0;
  uint64_t nnz = 0, cap_e = 0;
  uint32_t max_entries = 0;
};

PRF_API prf_multi* prf_multi_create(int n_gpus, const int* devices) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return nullptr;
  }
  if (n_gpus <= 0) n_gpus = ndev;
  if (n_gpus > kMaxParts) return nullptr;
  prf_multi* m = new (std::nothrow) prf_multi;
  if (!m) return nullptr;
  m->world = (uint32_t)n_gpus;
  for (int r = 0; r < n_gpus; ++r) {
    m->ctx[r] = prf_create(devices ? devices[r] : r % ndev);
    if (!m->ctx[r]) {
      for (int q = 0; q < r; ++q) prf_destroy(m->ctx[q]);
      delete m;
      return nullptr;
    }
  }
  return m;
}

PRF_API void prf_multi_destroy(prf_multi* m) {
  if (!m) return;
  for (uint32_t r = 0; r < m->world; ++r) prf_destroy(m->ctx[r]);
  delete m;
}

PRF_API const char* prf_multi_last_error(const prf_multi* m) { return m ? m->err.c_str() : "no multi-GPU context"; }
PRF_API int prf_multi_gpus(const prf_multi* m) { return m ? (int)m->world : 0; }
PRF_API prf_ctx* prf_multi_ctx(prf_multi* m, int rank) { return m && rank >= 0 && (uint32_t)rank < m->world ? m->ctx[rank] : nullptr; }

// hashRF on all GPUs of the context: the trees are split into `world` contiguous shards, the build is
// hash-partitioned over the GPUs (prf_team.cuh), every GPU computes an aligned slice of the packed triangle and
// copies it into out_upper / out_mask over its own PCIe link.
PRF_API int prf_multi_hashrf(prf_multi* m, const uint64_t* bips, const uint64_t* tree_off, uint32_t T, uint32_t W, int32_t editdist,
                             uint16_t* out_upper, uint32_t* out_mask, prf_stats* stats) {
  if (!m) return PRF_E_INVALID;
  m->err.clear();
  if (!tree_off) { m->err = "tree_off is NULL"; return PRF_E_INVALID; }
  const uint32_t world = m->world;
  if (tree_off[0] != 0) { m->err = "tree_off[0] must be 0"; return PRF_E_INVALID; }
  uint64_t cap_e = 0;
  uint32_t max_entries = 0;
  for (uint32_t t = 0; t < T; ++t) {
    if (tree_off[t + 1] < tree_off[t]) { m->err = "tree_off is not ascending"; return PRF_E_INVALID; }
    max_entries = (uint32_t)std::min<uint64_t>(0xffffffffull, std::max<uint64_t>(max_entries, tree_off[t + 1] - tree_off[t]));
  }
  for (uint32_t r = 0; r < world; ++r) {
    const uint32_t lo = (uint32_t)((uint64_t)T * r / world), hi = (uint32_t)((uint64_t)T * (r + 1) / world);
    cap_e = std::max(cap_e, tree_off[hi] - tree_off[lo]);
  }
  const uint64_t nnz = tree_off[T];
  if (nnz && !bips) { m->err = "bips is NULL"; return PRF_E_INVALID; }
  // (re)size the arenas when the problem outgrows them
  if (T != m->T || W != m->W || nnz > m->nnz || cap_e > m->cap_e || max_entries > m->max_entries) {
    for (uint32_t r = 0; r < world; ++r) {
      int rc = prf_team_init(m->ctx[r], r, world, T, W, nnz, cap_e, max_entries, nullptr);
      if (rc) { m->err = prf_last_error(m->ctx[r]); return rc; }
    }
    prf_ctx* hs[kMaxParts];
    for (uint32_t r = 0; r < world; ++r) hs[r] = m->ctx[r];
    int rc = team_connect_local(hs, world, &m->shared);
    if (rc) { m->err = "team_connect_local failed"; return rc; }
    m->T = T; m->W = W; m->nnz = nnz; m->cap_e = cap_e; m->max_entries = max_entries;
  }
  const uint64_t P = (uint64_t)T * (T ? T - 1 : 0) / 2;
  int rcs[kMaxParts] = {};
  prf_stats sts[kMaxParts] = {};
  auto worker = [&](uint32_t r) {
    Ctx* ctx = &m->ctx[r]->c;
    int& rc = rcs[r];
    rc = PRF_OK;
    auto ck = [&](cudaError_t e) { if (e != cudaSuccess && rc == PRF_OK) { ctx->fail(PRF_E_CUDA, "%s", cudaGetErrorString(e)); rc = PRF_E_CUDA; } };
    ck(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const uint32_t lo = (uint32_t)((uint64_t)T * r / world), hi = (uint32_t)((uint64_t)T * (r + 1) / world), Tl = hi - lo;
    const uint64_t b0 = tree_off[lo], nl = tree_off[hi] - b0;
    DevBuf<uint64_t> d_keys, d_off;
    std::vector<uint64_t> off_l((size_t)Tl + 1);
    for (uint32_t t = 0; t <= Tl; ++t) off_l[t] = tree_off[lo + t] - b0;
    ck(d_keys.alloc(nl * W + 2, s));
    ck(d_off.alloc((size_t)Tl + 1, s));
    if (nl) ck(cudaMemcpyAsync(d_keys.p, bips + b0 * W, nl * W * 8, cudaMemcpyHostToDevice, s));
    ck(cudaMemcpyAsync(d_off.p, off_l.data(), ((size_t)Tl + 1) * 8, cudaMemcpyHostToDevice, s));
    ck(cudaStreamSynchronize(s));
    // the collective part runs even after a local failure so that the peers' barriers stay matched
    int lrc = team_load(ctx, d_keys.p, d_off.p, Tl, nl, &sts[r]);
    if (rc == PRF_OK) rc = lrc;
    if (rc != PRF_OK) return;
    uint64_t pb = 0, pe = 0;
    prf_shard_range(T, r, world, &pb, &pe);
    prf_stats st2{};
    rc = prf_hashrf_compute(m->ctx[r], pb, pe, editdist, PRF_VARIANT_SPARSE, nullptr, nullptr, &st2);
    if (rc) return;
    sts[r].ms_matrix = st2.ms_matrix;
    sts[r].variant_used = st2.variant_used;
    if (out_upper && pe > pb) rc = prf_fetch(m->ctx[r], pb, pe, out_upper + pb);
    if (!rc && out_mask && editdist >= 0 && pe > pb) rc = prf_fetch_mask(m->ctx[r], pb, pe, out_mask + pb / 32);
  };
  std::vector<std::thread> th;
  for (uint32_t r = 1; r < world; ++r) th.emplace_back(worker, r);
  worker(0);
  for (auto& t : th) t.join();
  for (uint32_t r = 0; r < world; ++r)
    if (rcs[r]) {
      m->err = "rank " + std::to_string(r) + ": " + prf_last_error(m->ctx[r]);
      return rcs[r];
    }
  if (stats) {
    *stats = sts[0];
    for (uint32_t r = 1; r < world; ++r) {
      stats->ms_matrix = std::max(stats->ms_matrix, sts[r].ms_matrix);
      stats->ms_total = std::max(stats->ms_total, sts[r].ms_total);
    }
    stats->n_pairs = P;
  }
  return PRF_OK;
}

PRF_API int prf_multi_components(prf_multi* m, uint32_t* labels, uint32_t* n_components) {
  if (!m) return PRF_E_INVALID;
  m->err.clear();
  const uint32_t world = m->world, T = m->T;
  if (!T || !labels) { m->err = "no problem loaded (prf_multi_hashrf with editdist >= 0 first) or labels is NULL"; return PRF_E_STATE; }
  std::vector<uint32_t> all((size_t)world * T);
  int rcs[kMaxParts] = {};
  auto worker = [&](uint32_t r) {
    uint64_t pb = 0, pe = 0;
    prf_shard_range(T, r, world, &pb, &pe);
    uint32_t n = 0;
    if (pe > pb) rcs[r] = prf_mask_components_shard(m->ctx[r], nullptr, pb, pe, T, nullptr, 0, all.data() + (size_t)r * T, &n);
    else for (uint32_t t = 0; t < T; ++t) all[(size_t)r * T + t] = t;
  };
  std::vector<std::thread> th;
  for (uint32_t r = 1; r < world; ++r) th.emplace_back(worker, r);
  worker(0);
  for (auto& t : th) t.join();
  for (uint32_t r = 0; r < world; ++r)
    if (rcs[r]) { m->err = "rank " + std::to_string(r) + ": " + prf_last_error(m->ctx[r]); return rcs[r]; }
  int rc = prf_mask_components_shard(m->ctx[0], nullptr, 0, 0, T, all.data(),
                                     world, labels, n_components);
  if (rc) m->err = std::string("merge: ") + prf_last_error(m->ctx[0]);
  return rc;
}

}  // extern "C"
