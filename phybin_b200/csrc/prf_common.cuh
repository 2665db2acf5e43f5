// prf_common.cuh -- context, error plumbing and small device helpers shared by every kernel file.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdlib>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "phybin_rf.h"

namespace prf {

constexpr uint32_t kMaxWords = 16;            // W <= 16 (1024 taxa); kernels are templated on W
constexpr uint32_t kMaxBipsPerTree = 32767;   // so that |B_a| + |B_b| fits uint16
constexpr uint64_t kShardAlign = 65536;       // multiple of every chunk size the sparse kernel uses

struct Ctx;

// Stream-ordered device buffer (cudaMallocAsync on the context's stream).
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; s = o.s; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count, cudaStream_t stream) {
    release();
    s = stream;
    n = count;
    return cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), stream);
  }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

constexpr int kMaxParts = 16;

// One hash class of the shared bipartitions in the form the sparse matrix kernel reads (prf_rows.cuh); a
// single-GPU build has one part holding everything, a multi-GPU build one part per rank (prf_team.cuh).
//   long lists  (more than short_max member trees): members bucketed by COLUMN BLOCK (B = 1 << block_log2 consecutive
//               tree ids); bucket (l, j) = lmem[boff[l * NB + j] .. boff[l * NB + j + 1]) holds `tree & (B - 1)` of the
//               list's members inside block j in any order, padded with 0xffff to a multiple of 8 entries
//   short lists expanded into per-tree "singles": for tree a the trees b > a it shares such a bipartition with
struct ListPart {
  const uint32_t* roff;     // [T + 1] offsets into rlist
  const uint32_t* rlist;    // per tree: ids of the long lists that contain it
  const uint32_t* boff;     // [n_long * NB + 1]
  const uint16_t* lmem;
  const uint32_t* soff;     // [T + 1] offsets into singles
  const uint32_t* singles;
};

// Device-resident state of a loaded hashRF problem (the incidence in the form the kernels use).
struct HashRFState {
  bool loaded = false;
  uint32_t T = 0, W = 0;
  uint64_t nnz = 0, P = 0;
  bool uniform_nb = false;       // every tree has the same |B|
  uint32_t nb_uniform = 0;
  DevBuf<uint16_t> nb;           // [T]   |B_t| (distinct bipartitions; |D_t| when majority lists were flipped)
  DevBuf<uint16_t> nb0;          // [T]   |B_t| before the flip (only when n_flipped > 0: the dense variant's operand is unflipped)
  DevBuf<uint16_t> nb_shift;     // 8 copies of nb shifted by 0..7 entries (only when |B_t| is not uniform): any run
                                 // nb[c..] is then a 16-byte aligned bulk copy away from a 16-byte aligned tile
  uint32_t nb_shift_stride = 0;
  uint32_t block_log2 = 13, NB = 0;  // column block size of the bucketed lists, number of blocks = ceil(T / B)
  uint32_t n_parts = 0;
  ListPart part[kMaxParts];
  uint32_t max_lists = 0;        // most long lists of one tree (summed over the parts)
  // the buffers behind part[0] of a single-context build
  DevBuf<uint32_t> roff, rlist, boff, soff, singles;
  DevBuf<uint16_t> lmem;
  uint32_t n_long = 0;           // long lists
  uint64_t n_lmem = 0;           // entries of lmem (with padding)
  uint64_t rlist_n = 0, singles_n = 0;  // entries of rlist / singles
  // tree-major incidence over ALL shared bipartitions (dense variant input; single-context builds only)
  DevBuf<uint32_t> ent_lid;      // [entry] dense id (0 .. n_shared-1) of the entry's list, 0xffffffff = bipartition of one tree only
  DevBuf<uint64_t> ent_tb, ent_te;  // [T] entry range of every tree
  uint32_t n_shared = 0;         // shared bipartitions (lists of >= 2 trees)
  uint32_t n_list_ids = 0;       // list ids in use (a tree listing a bipartition twice can leave an id with one tree)
  uint64_t n_shared_nnz = 0;     // their entries
  uint32_t n_flipped = 0;        // bipartitions held by > T/2 trees, stored as the list of trees lacking them
  // dense variant: tiled uint8 operand, built lazily by the first dense compute (prf_dense.cuh)
  bool dense_ready = false;
  uint32_t dense_Tpad = 0, dense_Kpad = 0;
  DevBuf<uint8_t> dense_At;
  DevBuf<uint2> dense_tiles;
  prf_stats stats{};
};

// Work item of the sparse matrix kernel (prf_sparse.cuh): one row a0 == a1 restricted to columns
// [cb, ce), or the whole rows a0..a1 (a1 > a0) packed into one tile.
struct RowItem {
  uint32_t a0, a1, cb, ce;
};

// Cached cut of a packed range into RowItems (depends only on T and the range, not on the trees).
struct RowPlan {
  bool valid = false;
  uint32_t T = 0, n_items = 0, tile = 0;
  uint64_t p_begin = 0, p_end = 0;
  int ctas_per_sm = 0;
  DevBuf<RowItem> items;
  DevBuf<uint32_t> counter;
  DevBuf<uint2> cursor_ws;
};

struct TolerantState {
  bool loaded = false;
  uint32_t T = 0, W = 0, max_clades = 0;
  uint64_t nnz = 0, P = 0;
  DevBuf<uint64_t> clades;    // [nnz*W]
  DevBuf<uint64_t> off;       // [T+1]
  DevBuf<uint64_t> leafsets;  // [T*W]
  DevBuf<uint16_t> par;       // [nnz]   parent clade of every clade inside its tree (prf_tolerant.cuh), 0xffff = root
  bool dups = false;          // some tree repeats a leaf label (prf_tolerant_load_dups): the arrays below are set
  DevBuf<uint64_t> outsets;   // [nnz*W] label set of the leaves outside each clade's node
  DevBuf<uint64_t> dup_off;   // [T+1]
  DevBuf<uint32_t> dup_label, dup_extra;
  prf_stats stats{};
};

// Result of the last device-side extraction (prf_extract.cuh).
struct ExtractState {
  bool valid = false;
  int mode = 0;
  uint32_t T = 0, W = 0;
  uint64_t nnz = 0;
  DevBuf<uint64_t> keys;      // [nnz * W]
  DevBuf<uint64_t> off;       // [T + 1]
  DevBuf<uint64_t> leafsets;  // [T * W]  (mode PRF_EXTRACT_CLADES only)
};

// Result of the last prf_consensus (prf_consensus.cuh).
struct ConsensusState {
  bool valid = false;
  uint32_t T = 0, W = 0;
  uint64_t n_out = 0;
  DevBuf<uint64_t> keys;  // [n_out * W]
  DevBuf<uint64_t> off;   // [T + 1]
};

constexpr size_t kTeamHandleBytes = 64;  // cudaIpcMemHandle_t

struct TeamPartHeader {  // what rank p tells everybody about its class (first bytes of part region p)
  unsigned long long n_unique, n_hits, n_entries, nnz;
  uint32_t n_lists2, n_long, n_flipped, n_rlist, n_boff, n_lmem, n_singles, max_lists;
  uint32_t err;          // != 0: the class did not fit its region
  uint32_t pad[3];
};

struct TeamLayout {
  uint32_t world = 0, T = 0, W = 0, NB = 0;
  uint64_t cap_e = 0;      // most entries one rank holds locally = capacity of a key region
  uint64_t cap_class = 0;  // most entries one hash class is sized for
  uint64_t cap_r = 0, cap_b = 0, cap_l = 0, cap_s = 0;  // rlist, boff, lmem (u16), singles entries per part
  size_t off_flags = 0, off_keys = 0, off_tb = 0, off_te = 0, off_parts = 0, part_bytes = 0, bytes = 0;
  size_t p_nb = 0, p_roff = 0, p_soff = 0, p_rlist = 0, p_boff = 0, p_lmem = 0, p_singles = 0;  // inside a part region
  static size_t up(size_t x) { return (x + 255) & ~(size_t)255; }
  void compute(uint32_t world_, uint32_t T_, uint32_t W_, uint32_t NB_, uint64_t nnz_total, uint64_t cap_e_, uint32_t short_max) {
    world = world_; T = T_; W = W_; NB = NB_; cap_e = cap_e_;
    cap_class = nnz_total <= (4u << 20) ? nnz_total : std::min<uint64_t>(nnz_total, 2 * nnz_total / world + 65536);
    cap_class = std::max<uint64_t>(cap_class, 16);
    cap_r = 2 * cap_class + 16;  // (majority lists stored as complements add at most as many again)
    cap_b = (cap_class / (short_max + 1) + 2) * NB + 16;
    cap_l = 2 * cap_class + 7 * std::min<uint64_t>(cap_b, 2 * cap_class) + 16;
    cap_s = cap_class * (short_max > 1 ? (short_max - 1) : 0) / 2 + 16;
    size_t o = 0;
    off_flags = o; o += 4096;
    off_keys = o; o += up((size_t)world * cap_e * W * 8);
    off_tb = o; o += up((size_t)T * 8 + 8);
    off_te = o; o += up((size_t)T * 8 + 8);
    off_parts = o;
    size_t q = up(sizeof(TeamPartHeader));
    p_nb = q; q += up(((size_t)T + 16) * 2);
    p_roff = q; q += up(((size_t)T + 1) * 4);
    p_soff = q; q += up(((size_t)T + 1) * 4);
    p_rlist = q; q += up(cap_r * 4);
    p_boff = q; q += up(cap_b * 4);
    p_lmem = q; q += up(cap_l * 2);
    p_singles = q; q += up(cap_s * 4);
    part_bytes = q;
    bytes = off_parts + (size_t)world * part_bytes;
  }
};

struct TeamShared;  // single-process teams: host barrier + events (prf_team.cu)

struct Team {
  bool active = false, connected = false;
  uint32_t rank = 0, world = 1, max_entries = 0;
  uint64_t nnz_total = 0;
  TeamLayout lay;
  uint8_t* arena = nullptr;           // mine (cudaMalloc)
  uint8_t* peer[kMaxParts] = {};      // every rank's arena as mapped here (peer[rank] == arena)
  bool ipc_open[kMaxParts] = {};
  uint64_t epoch = 0;                 // barriers passed
  TeamShared* shared = nullptr;       // non-null: single-process team (events instead of flag kernels)
  cudaEvent_t ev = nullptr;           // single-process: my arrival event
  DevBuf<uint32_t> cls_cnt, cls_pos, err;  // partition scratch
  DevBuf<uint8_t> cls;
  DevBuf<uint64_t> peers_dev;         // [world] peer arena addresses (device copy for the kernels)
};

struct Ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[8] = {};
  std::string err;
  uint64_t launches = 0;
  HashRFState h;
  TolerantState tol;
  ExtractState ext;
  ConsensusState cons;
  RowPlan plan;
  Team team;
  DevBuf<unsigned long long> scan_status;  // decoupled look-back scan: tile status words (prf_primitives.cuh)
  DevBuf<uint32_t> scan_counter;
  uint32_t scan_gen = 0;
  bool debug_small_tiles = false;
  uint32_t short_max = 8;             // lists of at most this many trees are expanded into per-row singles
  int calib = 0;                      // PRF_CALIB builds: calibration mode of the rows kernel (wrong results)
  int rows_cfg = 0;                   // tests / tuning: instantiation of the rows kernel
  bool debug_no_flip = false;         // tests: keep majority lists as they are
  bool debug_wide_slots = false;      // tests: force the 8-byte-slot dedup table
  uint32_t debug_agglo_ctas = 0;      // tuning: CTAs of the agglomeration kernel (0 = default)
  DevBuf<unsigned long long> timing;  // PRF_TIMING builds only
  uint32_t timing_grid = 0;
  // ctx-owned result of the last compute
  DevBuf<uint16_t> out;
  DevBuf<uint32_t> mask;
  uint64_t out_begin = 0, out_end = 0, mask_begin = 0, mask_end = 0;
  bool out_valid = false, mask_valid = false;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

// PRF_TRACE=1 in the environment: host wall-clock marks on stderr (where does a call spend its time).
inline bool trace_on() {
  static const bool on = [] { const char* e = getenv("PRF_TRACE"); return e && *e && *e != '0'; }();
  return on;
}
inline void trace_mark(const char* what) {
  if (!trace_on()) return;
  static std::chrono::steady_clock::time_point last = std::chrono::steady_clock::now();
  auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[prf] %-28s +%9.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
  last = now;
}

// PRF_TRACE=1: additionally waits for the stream, so that the mark's time is the GPU time of what was queued since
// the previous mark (serialises the phases: for diagnosis only).
inline void trace_sync(cudaStream_t s, const char* what) {
  if (!trace_on()) return;
  cudaStreamSynchronize(s);
  trace_mark(what);
}

#define PRF_CK(ctx, call)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return (ctx)->fail(PRF_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,                \
                         cudaGetErrorString(_e));                                              \
  } while (0)

// Counts a launch and checks it was accepted.
#define PRF_LAUNCHED(ctx)                                                                      \
  do {                                                                                         \
    ++(ctx)->launches;                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return (ctx)->fail(PRF_E_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,            \
                         cudaGetErrorString(_e));                                              \
  } while (0)

__host__ __device__ inline uint64_t row_start(uint64_t a, uint64_t T) {
  // first packed position of upper-triangle row a: a*T - a(a+1)/2 = a(2T - a - 1)/2
  return a * (2 * T - a - 1) / 2;
}

// Largest row a (0 <= a <= T-2) with row_start(a) <= p, for p < T(T-1)/2.
__host__ __device__ inline uint32_t row_of(uint64_t p, uint64_t T) {
  double tt = 2.0 * (double)T - 1.0;
  double disc = tt * tt - 8.0 * (double)p;
  double r = (tt - sqrt(disc > 0.0 ? disc : 0.0)) * 0.5;
  int64_t a = (int64_t)r;
  if (a < 0) a = 0;
  if (a > (int64_t)T - 2) a = (int64_t)T - 2;
  while (a > 0 && row_start((uint64_t)a, T) > p) --a;
  while (a + 1 <= (int64_t)T - 2 && row_start((uint64_t)a + 1, T) <= p) ++a;
  return (uint32_t)a;
}

inline uint32_t grid_for(uint64_t n, uint32_t tpb = 256) { return (uint32_t)((n + tpb - 1) / tpb); }
inline int bits_for(uint64_t n_values) {  // bits needed to represent 0..n_values-1
  int b = 0;
  while (b < 63 && (1ull << b) < n_values) ++b;
  return b;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr uint64_t kEmptySlot = ~0ull;

template <int W>
__device__ __forceinline__ void load_key(const uint64_t* __restrict__ bips, uint64_t e, uint64_t (&k)[W]) {
  const uint64_t* p = bips + e * W;
  if constexpr (W % 2 == 0) {
#pragma unroll
    for (int w = 0; w < W; w += 2) {
      ulonglong2 q = *reinterpret_cast<const ulonglong2*>(p + w);
      k[w] = q.x;
      k[w + 1] = q.y;
    }
  } else {
#pragma unroll
    for (int w = 0; w < W; ++w) k[w] = p[w];
  }
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace prf
