// Event -> launch -> event on an idle stream: what a tiny frame pays outside its kernel body, as a function of the kernel's parameter
// bytes and of a final store to mapped host memory.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/launch_floor.bin
// tools/launch_floor.cu.  The cases are interleaved round-robin (clock ramps and drift hit all of them alike); median of 300 each.
#include <algorithm>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>
#include <cuda_runtime.h>
template <int N> struct P { int x[N]; };
template <int N> __global__ void __launch_bounds__(1024, 1) k_dev(const __grid_constant__ P<N> a, volatile int *out) {
  if (a.x[threadIdx.x % N] == 12345) out[0] = 1;
}
template <int N> __global__ void __launch_bounds__(1024, 1) k_host(const __grid_constant__ P<N> a, volatile int *host) {
  if (threadIdx.x < 20) host[threadIdx.x] = a.x[threadIdx.x % N];
}
__constant__ P<176> g_args;
__global__ void __launch_bounds__(1024, 1) k_sym(volatile int *out) { if (g_args.x[threadIdx.x % 176] == 12345) out[0] = 1; }
int main() {
  int *d; cudaMalloc(&d, 4096);
  int *h; cudaHostAlloc(&h, 4096, cudaHostAllocMapped); int *hd; cudaHostGetDevicePointer(&hd, h, 0);
  static P<4> p4{}; static P<16> p16{}; static P<32> p32{}; static P<64> p64{}; static P<96> p96{}; static P<128> p128{}; static P<176> p176{};
  std::vector<std::pair<std::string, std::function<void()>>> cases = {
      {"   16 B params, 1 x 1024", [&] { k_dev<4><<<1, 1024>>>(p4, d); }},
      {"   64 B params", [&] { k_dev<16><<<1, 1024>>>(p16, d); }},
      {"  128 B params", [&] { k_dev<32><<<1, 1024>>>(p32, d); }},
      {"  256 B params", [&] { k_dev<64><<<1, 1024>>>(p64, d); }},
      {"  384 B params", [&] { k_dev<96><<<1, 1024>>>(p96, d); }},
      {"  512 B params", [&] { k_dev<128><<<1, 1024>>>(p128, d); }},
      {"  704 B params", [&] { k_dev<176><<<1, 1024>>>(p176, d); }},
      {"  704 B params, 2 x 256", [&] { k_dev<176><<<2, 256>>>(p176, d); }},
      {"   16 B params + 80 B to mapped host memory", [&] { k_host<4><<<1, 1024>>>(p4, hd); }},
      {"  704 B params + 80 B to mapped host memory", [&] { k_host<176><<<1, 1024>>>(p176, hd); }},
      {"  704 B in a __constant__ symbol (8 B params)", [&] { k_sym<<<1, 1024>>>(d); }},
      {"  no kernel (event, event)", [&] {}},
  };
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<std::vector<float>> v(cases.size());
  for (int it = 0; it < 330; it++)
    for (size_t c = 0; c < cases.size(); c++) {
      cudaEventRecord(e0); cases[c].second(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 30) v[c].push_back(ms * 1e3f);
    }
  for (size_t c = 0; c < cases.size(); c++) {
    std::sort(v[c].begin(), v[c].end());
    printf("%-48s median %6.2f us   min %6.2f us\n", cases[c].first.c_str(), v[c][v[c].size() / 2], v[c][0]);
  }
  return 0;
}
