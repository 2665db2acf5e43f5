// Micro-benchmark: shared-memory update throughput per SM for random 16-bit entries in a 32 KB tile.
//   mode 0: atomicSub on the containing 32-bit word (ATOMS.ADD, no return)     -- what the round-1 sparse kernel does
//   mode 1: non-atomic 16-bit read-modify-write (LDS.U16 / STS.U16)
//   mode 2: atomicSub, only 16 of 32 lanes active
//   mode 3: loads only (LDS.U16), as a floor
//   mode 4: warp-private tile: every warp owns its own 16 KB / 32 KB tile, non-atomic 16-bit RMW + __syncwarp
//           (what a one-warp-per-tile walker would do), 20 of 32 lanes active
//   mode 5: as 4 but two independent RMWs in flight per step (two windows of one list)
//   mode 6: fill only: st.shared.v4 of a whole tile per "iteration" (bytes/cycle/SM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms atoms.cu ; run: ./atoms
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  extern __shared__ uint32_t tile32[];
  uint16_t* tile16 = (uint16_t*)tile32;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) tile32[i] = 0x40004000u;
  __syncthreads();
  uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    x = x * 1664525u + 1013904223u;
    uint32_t idx = (x >> 10) & 16383u;
    if (MODE == 0) atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u));
    if (MODE == 1) tile16[idx] = (uint16_t)(tile16[idx] - 2);
    if (MODE == 2) { if ((threadIdx.x & 1) == 0) atomicSub(&tile32[idx >> 1], 2u << ((idx & 1u) * 16u)); }
    if (MODE == 3) acc += tile16[idx];
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = tile32[seed & 8191] + acc;
}

// warp-private tiles of TILE entries (uint16)
template <int MODE, int TILE>
__global__ void kw(uint32_t* out, int iters, uint32_t seed) {
  extern __shared__ uint32_t tile32[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint16_t* t16 = (uint16_t*)tile32 + (size_t)warp * TILE;
  uint4* t4 = (uint4*)t16;
  for (int i = lane; i < TILE / 8; i += 32) t4[i] = make_uint4(0x40004000u, 0x40004000u, 0x40004000u, 0x40004000u);
  __syncwarp();
  uint32_t x = seed * 2654435761u + threadIdx.x * 40503u + blockIdx.x * 977u;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 4) {
      x = x * 1664525u + 1013904223u;
      // distinct per lane: lane-strided random base
      uint32_t idx = ((x >> 10) & (TILE / 32 - 1)) * 32 + ((lane * 7 + it) & 31);
      if (lane < 20) t16[idx] = (uint16_t)(t16[idx] - 2);
      __syncwarp();
    }
    if (MODE == 5) {
      x = x * 1664525u + 1013904223u;
      uint32_t i0 = ((x >> 10) & (TILE / 64 - 1)) * 32 + ((lane * 7 + it) & 31);
      uint32_t i1 = i0 + TILE / 2;
      uint16_t a = t16[i0], b = t16[i1];
      t16[i0] = (uint16_t)(a - 2);
      if (lane < 9) t16[i1] = (uint16_t)(b - 2);
      __syncwarp();
    }
    if (MODE == 6) {
      const uint4 q = make_uint4(it, it, it, it);
#pragma unroll
      for (int i = 0; i < TILE / 256; ++i) t4[lane + 32 * i] = q;
      __syncwarp();
    }
    if (MODE == 7) {  // atomics on a warp-private tile, 20 lanes
      x = x * 1664525u + 1013904223u;
      uint32_t idx = ((x >> 10) & (TILE / 32 - 1)) * 32 + ((lane * 7 + it) & 31);
      if (lane < 20) atomicSub(&((uint32_t*)t16)[idx >> 1], 2u << ((idx & 1u) * 16u));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = tile32[seed & 1023] + acc;
}

template <int MODE>
void run(const char* name, int threads, int blocks_per_sm) {
  int sms = 148;
  uint32_t* out; cudaMalloc(&out, 4096 * 4);
  int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<sms * blocks_per_sm, threads, 32768>>>(out, 100, 1);
  cudaEventRecord(a);
  k<MODE><<<sms * blocks_per_sm, threads, 32768>>>(out, iters, 2);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double lane_ops = (double)sms * blocks_per_sm * threads * iters * (MODE == 2 ? 0.5 : 1.0);
  printf("%-28s thr=%4d cta/sm=%d  %.3f ms  %.2f lane-updates/ns/GPU  %.2f per SM per cycle@1.9GHz  err=%s\n", name, threads,
         blocks_per_sm, ms, lane_ops / (ms * 1e6), lane_ops / (ms * 1e6) / 148 / 1.9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

template <int MODE, int TILE>
void runw(const char* name, int warps) {
  int sms = 148;
  uint32_t* out; cudaMalloc(&out, 4096 * 4);
  int iters = MODE == 6 ? 2000 : 20000;
  size_t smem = (size_t)warps * TILE * 2;
  cudaFuncSetAttribute(kw<MODE, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  kw<MODE, TILE><<<sms, warps * 32, smem>>>(out, 100, 1);
  cudaEventRecord(a);
  kw<MODE, TILE><<<sms, warps * 32, smem>>>(out, iters, 2);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double per = MODE == 4 || MODE == 7 ? 20.0 : (MODE == 5 ? 41.0 : (double)TILE * 2);  // lane updates (or bytes) per warp-iteration
  double ops = (double)sms * warps * iters * per;
  printf("%-34s tile=%5d warps/SM=%2d  %.3f ms  %.2f %s per SM per cycle@1.9GHz  err=%s\n", name, TILE, warps, ms,
         ops / (ms * 1e6) / 148 / 1.9, MODE == 6 ? "bytes" : "lane-updates", cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  for (int t : {256, 512, 1024}) {
    run<0>("atomicSub u16-in-u32", t, 2);
    run<1>("LDS.U16/STS.U16 rmw", t, 2);
    run<2>("atomicSub half lanes", t, 2);
    run<3>("LDS.U16 only", t, 2);
  }
  for (int w : {4, 7}) {
    runw<4, 16384>("warp-private rmw 20 lanes", w);
    runw<5, 16384>("warp-private 2x rmw 41 lanes", w);
    runw<7, 16384>("warp-private atomics 20 lanes", w);
    runw<6, 16384>("warp-private fill v4", w);
  }
  for (int w : {7, 14}) {
    runw<4, 8192>("warp-private rmw 20 lanes", w);
    runw<5, 8192>("warp-private 2x rmw 41 lanes", w);
    runw<7, 8192>("warp-private atomics 20 lanes", w);
    runw<6, 8192>("warp-private fill v4", w);
  }
  for (int w : {14, 28}) {
    runw<4, 4096>("warp-private rmw 20 lanes", w);
    runw<5, 4096>("warp-private 2x rmw 41 lanes", w);
    runw<7, 4096>("warp-private atomics 20 lanes", w);
  }
  return 0;
}
