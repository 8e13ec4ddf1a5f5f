// Same-address atomic throughput on B200 (what work cursors, queue reservations and barrier counters cost):
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/atomic_bench tools/atomic_bench.cu && /tmp/atomic_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_same(unsigned *c, int iters, int with_return, unsigned *sink) {
  unsigned acc = 0;
  if ((threadIdx.x & 31) == 0) {
    for (int i = 0; i < iters; i++) {
      if (with_return) acc += atomicAdd(c, 1u);
      else asm volatile("red.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
    }
  }
  if (acc == 0xdeadbeef) *sink = acc;
}
__global__ void k_spread(unsigned *c, int iters, unsigned *sink) {   // one counter (own 128-byte line) per CTA
  unsigned acc = 0;
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < iters; i++) acc += atomicAdd(c + 32 * blockIdx.x, 1u);
  if (acc == 0xdeadbeef) *sink = acc;
}
__global__ void k_poll(unsigned *c, int iters, unsigned *sink) {
  unsigned acc = 0;
  if ((threadIdx.x & 31) == 0)
    for (int i = 0; i < iters; i++) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory"); acc += v; }
  if (acc == 0xdeadbeef) *sink = acc;
}
int main() {
  unsigned *c, *sink;
  cudaMalloc(&c, 1 << 20); cudaMemset(c, 0, 1 << 20); cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int grid : {1, 16, 148, 296}) {
    for (int mode = 0; mode < 4; mode++) {
      const int iters = 64;
      float best = 1e9f;
      for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k_same<<<grid, 512>>>(c, iters, 1, sink);
        if (mode == 1) k_same<<<grid, 512>>>(c, iters, 0, sink);
        if (mode == 2) k_spread<<<grid, 512>>>(c, iters, sink);
        if (mode == 3) k_poll<<<grid, 512>>>(c, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
      }
      const double ops = (double)grid * 16 * iters;
      const char *names[] = {"atomicAdd with return, one address", "red (no return), one address", "atomicAdd, one line per CTA", "ld.acquire.gpu poll, one address"};
      printf("grid %3d x 16 warps  %-36s %8.1f us  %7.2f ns/op  (%.0f ops)\n", grid, names[mode], best * 1e3, best * 1e6 / ops, ops);
    }
  }
  return 0;
}
