// alloc.cuh — stream-ordered device allocation for the render path.
// The reference allocates its three framebuffers with cudaMallocManaged and frees them inside every
// render() call (inc/kernel.hpp:99-101,116-118).  Here all device memory comes from a memory pool this
// library OWNS (one cudaMemPool_t per device, created on first use, release threshold at its maximum), so
// re-uploading a scene or resizing a frame reuses blocks instead of paying cudaMalloc/cudaFree (which
// synchronise the device) — without touching the device's default pool, which belongs to the rest of the
// process (PyTorch's allocator, other cudaMallocAsync users).  cutrace_trim_memory() hands the cached
// blocks back to the driver.
#ifndef CUTRACE_B200_ALLOC_CUH
#define CUTRACE_B200_ALLOC_CUH
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <type_traits>

namespace ctb {

inline cudaMemPool_t *pool_slot(int device) {
  static cudaMemPool_t pools[64] = {};
  return (device >= 0 && device < 64) ? &pools[device] : nullptr;
}
inline std::mutex &pool_mutex() {
  static std::mutex m;
  return m;
}
// the library's pool of `device` (created on first use); nullptr -> fall back to the device's default pool
inline cudaMemPool_t pool_of(int device) {
  cudaMemPool_t *slot = pool_slot(device);
  if (!slot) return nullptr;
  std::lock_guard<std::mutex> lk(pool_mutex());
  if (!*slot) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    cudaMemPool_t p = nullptr;
    if (cudaMemPoolCreate(&p, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &threshold);
    *slot = p;
  }
  return *slot;
}
inline void pool_trim(int device) {
  cudaMemPool_t *slot = pool_slot(device);
  std::lock_guard<std::mutex> lk(pool_mutex());
  if (slot && *slot) cudaMemPoolTrimTo(*slot, 0);
}

template <typename T>
inline cudaError_t dmalloc(T **p, size_t bytes, cudaStream_t st) {
  *p = nullptr;
  if (bytes == 0) return cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool = pool_of(dev);
  if (pool) return cudaMallocFromPoolAsync(reinterpret_cast<void **>(p), bytes, pool, st);
  return cudaMallocAsync(reinterpret_cast<void **>(p), bytes, st);
}

template <typename T>
inline void dfree(T *p, cudaStream_t st) {
  if (p) cudaFreeAsync(const_cast<typename std::remove_const<T>::type *>(p), st);
}

}  // namespace ctb
#endif
