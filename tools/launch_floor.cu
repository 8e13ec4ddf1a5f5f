// Event -> launch -> event on an idle stream: what a tiny frame pays outside its kernel body.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// -o build/launch_floor tools/launch_floor.cu.  Prints the median microseconds of 200 launches per case.
#include <algorithm>
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
struct Small { int x[4]; };
struct Big { int x[176]; };   // 704 bytes, about sizeof(PixelArgs)
__global__ void k_small(Small a, int *out) { if (a.x[0] == 12345) out[0] = 1; }
__global__ void __launch_bounds__(1024, 1) k_big(const __grid_constant__ Big a, int *out) { if (a.x[threadIdx.x % 176] == 12345) out[0] = 1; }
__global__ void __launch_bounds__(1024, 1) k_big_host(const __grid_constant__ Big a, volatile int *host) { if (threadIdx.x < 20) host[threadIdx.x] = a.x[threadIdx.x]; }
__global__ void __launch_bounds__(1024, 1) k_big_dev(const __grid_constant__ Big a, volatile int *dev) { if (threadIdx.x < 20) dev[threadIdx.x] = a.x[threadIdx.x]; }
template <typename F> static float median_us(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<float> v;
  for (int i = 0; i < 220; i++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (i >= 20) v.push_back(ms * 1e3f);
  }
  std::sort(v.begin(), v.end());
  return v[v.size() / 2];
}
int main() {
  int *d; cudaMalloc(&d, 4096);
  int *h; cudaHostAlloc(&h, 4096, cudaHostAllocMapped); int *hd; cudaHostGetDevicePointer(&hd, h, 0);
  Small s{}; Big b{};
  printf("empty, 16-byte params, 1 x 32 threads        %6.2f us\n", median_us([&] { k_small<<<1, 32>>>(s, d); }));
  printf("empty, 16-byte params, 2 x 256 threads       %6.2f us\n", median_us([&] { k_small<<<2, 256>>>(s, d); }));
  printf("empty, 704-byte params, 1 x 1024 threads     %6.2f us\n", median_us([&] { k_big<<<1, 1024>>>(b, d); }));
  printf("empty, 704-byte params, 4 x 256 threads      %6.2f us\n", median_us([&] { k_big<<<4, 256>>>(b, d); }));
  printf("80 B to device memory, 1 x 1024              %6.2f us\n", median_us([&] { k_big_dev<<<1, 1024>>>(b, d); }));
  printf("80 B to mapped host memory, 1 x 1024         %6.2f us\n", median_us([&] { k_big_host<<<1, 1024>>>(b, hd); }));
  printf("no kernel (event, event)                     %6.2f us\n", median_us([&] {}));
  return 0;
}
