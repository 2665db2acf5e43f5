// Micro-benchmark: shared-memory update throughput per SM for random 16-bit entries in a 32 KB tile.
//   mode 0: atomicSub on the containing 32-bit word (ATOMS.ADD, no return)     -- what the sparse kernel does
//   mode 1: non-atomic 16-bit read-modify-write (LDS.U16 / STS.U16)
//   mode 2: atomicSub, only 16 of 32 lanes active
//   mode 3: loads only (LDS.U16), as a floor
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

int main() {
  for (int t : {256, 512, 1024}) {
    run<0>("atomicSub u16-in-u32", t, 2);
    run<1>("LDS.U16/STS.U16 rmw", t, 2);
    run<2>("atomicSub half lanes", t, 2);
    run<3>("LDS.U16 only", t, 2);
  }
  return 0;
}
