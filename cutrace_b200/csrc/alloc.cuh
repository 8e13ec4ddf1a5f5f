// alloc.cuh — stream-ordered device allocation for the render path.
// The reference allocates its three framebuffers with cudaMallocManaged and frees them inside every
// render() call (inc/kernel.hpp:99-101,116-118).  Here all device memory comes from the device's
// default memory pool (cudaMallocAsync) with the release threshold raised, so re-uploading a scene or
// resizing a frame reuses blocks instead of paying cudaMalloc/cudaFree (which synchronise the device).
#ifndef CUTRACE_B200_ALLOC_CUH
#define CUTRACE_B200_ALLOC_CUH
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctb {

inline void pool_keep_memory(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
}

template <typename T>
inline cudaError_t dmalloc(T **p, size_t bytes, cudaStream_t st) {
  *p = nullptr;
  if (bytes == 0) return cudaSuccess;
  return cudaMallocAsync(reinterpret_cast<void **>(p), bytes, st);
}

template <typename T>
inline void dfree(T *p, cudaStream_t st) {
  if (p) cudaFreeAsync(const_cast<typename std::remove_const<T>::type *>(p), st);
}

}  // namespace ctb
#endif
