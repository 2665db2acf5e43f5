// prf_team.cuh -- the hashRF build (RFDistance.hs:289-297) over the GPUs of one box, without replicating the
// dedup on every GPU and without a host round trip per exchange.
//
// Bipartitions are partitioned by a hash of the key: rank r deduplicates only the keys of hash class r.  All
// occurrences of a key land in one class, so each rank's ordinary single-GPU build (prf_incidence.cuh) over
// "every tree restricted to class r" yields a complete ListPart, and
//     d(a,b) = sum_r ( |B_a|_r + |B_b|_r - 2 |B_a n B_b|_r ).
// The matrix kernel (prf_rows.cuh) reads the parts side by side, so nothing is merged: a row's lists are the
// union of its lists in every part.
//
// Every rank owns an ARENA (one cudaMalloc, mapped into every peer: cudaIpc handles between processes, peer
// access inside one process).  The exchanges are plain kernels that store straight into the peers' arenas over
// NVLink -- the partition kernel writes each key where its owner will read it, the part kernel copies a finished
// ListPart into every peer -- separated by a barrier: in the multi-process case a one-warp kernel that raises a
// flag in every peer's arena and spins (system-scope acquire, with a time-out) until every peer has raised its
// own; in the single-process case stream-ordered events.  No NCCL call, no host-side size exchange: every region
// has a capacity fixed at prf_team_init, and exceeding one is reported as an error, never truncated.
//
//   arena of rank r
//     flags[world]               barrier epochs, written by the peers
//     keys[world][cap_e][W]      region s: the keys of class r among the trees of rank s, tree order kept
//     tb[T], te[T]               entry range of every tree inside keys (written by the tree's owner)
//     part[world]                part p: header, |B_t|_p, roff, soff, rlist, boff, lmem, singles of class p
#pragma once

#include "prf_primitives.cuh"

namespace prf {

// hash class of a key: independent of the dedup table's slot / tag bits
template <int W>
__device__ __forceinline__ uint32_t key_class(const uint64_t* __restrict__ bips, uint64_t e, uint32_t world) {
  uint64_t k[W];
  load_key<W>(bips, e, k);
  uint64_t h = 0x9E3779B97F4A7C15ull;
#pragma unroll
  for (int w = 0; w < W; ++w) h = mix64(h ^ (k[w] + 0x632BE59BD9B4E019ull * (w + 1)));
  return (uint32_t)((mix64(h ^ 0xA5A5A5A5A5A5A5A5ull) >> 33) % world);
}

// One warp per local tree, one kernel: hash classes of the tree's keys, a contiguous range reserved for them in
// every class's (class, me) key region (atomic cursor per class: the trees of a region are in arrival order, every
// tree's entries contiguous), then every key stored straight into its owner's arena (peer memory over NVLink) and
// the tree's entry range into that class's tb / te.  The keys are read twice; the second read hits L1 / L2.
template <int W>
__global__ void __launch_bounds__(256) team_partition_kernel(const uint64_t* __restrict__ bips, const uint64_t* __restrict__ off,
                                                             uint32_t Tl, uint32_t t_lo, uint32_t world, uint32_t rank,
                                                             const uint64_t* __restrict__ peers, TeamLayout lay, uint32_t* cursor,
                                                             uint32_t* err) {
  const uint32_t t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = lane_id(), lt = lanemask_lt();
  if (t >= Tl) return;
  const uint64_t e0 = off[t], e1 = off[t + 1];
  uint32_t mine = 0;  // lane c: entries of class c
  for (uint64_t eb = e0; eb < e1; eb += 32) {
    const uint64_t e = eb + lane;
    const uint32_t c = e < e1 ? key_class<W>(bips, e, world) : 0xffffffffu;
    for (uint32_t k = 0; k < world; ++k) {
      const uint32_t n = __popc(__ballot_sync(0xffffffffu, c == k));
      if (lane == k) mine += n;
    }
  }
  uint64_t first = 0;
  if (lane < world) {
    const uint32_t at = atomicAdd(&cursor[lane], mine);
    if ((uint64_t)at + mine > lay.cap_e) *err = 1;  // (cannot happen: a region holds all of a rank's entries)
    first = (uint64_t)rank * lay.cap_e + at;
    reinterpret_cast<uint64_t*>(peers[lane] + lay.off_tb)[t_lo + t] = first;
    reinterpret_cast<uint64_t*>(peers[lane] + lay.off_te)[t_lo + t] = first + mine;
  }
  uint32_t run = 0;  // lane c: entries of class c written so far
  for (uint64_t eb = e0; eb < e1; eb += 32) {
    const uint64_t e = eb + lane;
    const uint32_t c = e < e1 ? key_class<W>(bips, e, world) : 0xffffffffu;
    for (uint32_t k = 0; k < world; ++k) {
      const uint32_t bal = __ballot_sync(0xffffffffu, c == k);
      if (!bal) continue;
      const uint64_t f = __shfl_sync(0xffffffffu, first, k);
      const uint32_t r = __shfl_sync(0xffffffffu, run, k);
      if (c == k) {  // (staging the runs in shared memory for whole-sector stores measured slower: 0.42 vs 0.33 ms at N = 4)
        uint64_t* out = reinterpret_cast<uint64_t*>(peers[k] + lay.off_keys) + (f + r + __popc(bal & lt)) * W;
        const uint64_t* in = bips + e * W;
        if ((W & 1u) == 0) {  // 16-byte stores over NVLink (key regions and the caller's keys are 16-byte aligned)
#pragma unroll
          for (int w = 0; w < W; w += 2) *reinterpret_cast<ulonglong2*>(out + w) = *reinterpret_cast<const ulonglong2*>(in + w);
        } else {
#pragma unroll
          for (int w = 0; w < W; ++w) out[w] = in[w];
        }
      }
      if (lane == k) run += __popc(bal);
    }
  }
  __threadfence_system();
}

// Multi-process barrier: tell every peer that I arrived (epoch), then wait until every peer has told me.
// `payload` (optional, [world] uint32 in my memory): payload[p] is delivered to rank p's arena (slot `rank` of its
// payload area) before the flag -- the key exchange uses it to tell every class how many entries I sent it.
__global__ void team_barrier_kernel(const uint64_t* __restrict__ peers, size_t off_flags, uint32_t rank, uint32_t world,
                                    unsigned long long epoch, const uint32_t* payload, uint32_t* err) {
  const uint32_t p = threadIdx.x;
  if (p >= world) return;
  if (payload) reinterpret_cast<unsigned long long*>(peers[p] + off_flags + 2048)[rank] = payload[p];
  __threadfence_system();
  unsigned long long* theirs = reinterpret_cast<unsigned long long*>(peers[p] + off_flags) + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank] + off_flags) + p;
  const long long t0 = clock64();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    if (v >= epoch) break;
    if (clock64() - t0 > 20000000000ll) {  // ~10 s: a peer is gone; fail instead of hanging the GPU
      *err = 2;
      break;
    }
    __nanosleep(200);
  }
}

// single-process teams: the payload delivery of team_barrier_kernel as a stream-ordered kernel of its own
__global__ void team_payload_kernel(const uint64_t* __restrict__ peers, size_t off_flags, uint32_t rank, uint32_t world,
                                    const uint32_t* payload) {
  const uint32_t p = threadIdx.x;
  if (p < world) reinterpret_cast<unsigned long long*>(peers[p] + off_flags + 2048)[rank] = payload[p];
}

struct TeamCopy {  // one array of my finished part
  const void* src;
  size_t dst_off;  // inside the part region
  size_t bytes;
};
struct TeamCopies {
  TeamCopy c[8];
  uint32_t n;
};
// Copies my part into part region `rank` of every peer (16-byte words; every source and offset is 16-byte aligned).
__global__ void __launch_bounds__(256) team_push_part_kernel(TeamCopies cp, const uint64_t* __restrict__ peers, size_t region_off,
                                                             uint32_t world) {
  for (uint32_t a = 0; a < cp.n; ++a) {
    const uint4* src = static_cast<const uint4*>(cp.c[a].src);
    const size_t n16 = (cp.c[a].bytes + 15) / 16;
    for (uint32_t p = 0; p < world; ++p) {
      uint4* dst = reinterpret_cast<uint4*>(peers[p] + region_off + cp.c[a].dst_off);
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
    }
  }
  __threadfence_system();
}

// |B_t| = sum over the parts; min / max; the most lists of one row.  out[0] = min, [1] = max, [2] = max lists
__global__ void __launch_bounds__(256) team_finish_kernel(const uint8_t* __restrict__ arena, TeamLayout lay, uint16_t* __restrict__ nb,
                                                          uint32_t* out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = 0, nl = 0;
  if (t < lay.T)
    for (uint32_t p = 0; p < lay.world; ++p) {
      const uint8_t* reg = arena + lay.off_parts + (size_t)p * lay.part_bytes;
      v += reinterpret_cast<const uint16_t*>(reg + lay.p_nb)[t];
      const uint32_t* roff = reinterpret_cast<const uint32_t*>(reg + lay.p_roff);
      nl += roff[t + 1] - roff[t];
    }
  uint32_t lo = t < lay.T ? v : 0xffffffffu, hi = v;
  if (t < lay.T) nb[t] = (uint16_t)min(v, 0xffffu);
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, d));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, d));
    nl = max(nl, __shfl_xor_sync(0xffffffffu, nl, d));
  }
  if (lane_id() == 0) {
    atomicMin(&out[0], lo);
    atomicMax(&out[1], hi);
    atomicMax(&out[2], nl);
  }
}

}  // namespace prf
