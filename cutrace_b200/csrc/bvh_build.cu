// bvh_build.cu — LBVH construction on the device.
//
//   1. prim_bounds_kernel   per-primitive AABB + centroid, scene bounds by ordered-int atomics
//   2. morton_kernel        63-bit Morton code of the centroid (21 bits / axis)
//   3. radix_sort_pairs     hand-written stable LSD radix sort, 8 bits / pass (CUB is used only by
//                           tests as a cross-check, never here)
//   4. gather_kernel        PrimRec + leaf boxes in sorted order
//   5. karras_kernel        Karras 2012 "Maximizing parallelism in the construction of BVHs":
//                           one thread per internal node finds its key range and split
//   6. refit_kernel         bottom-up boxes, second arrival at a node proceeds (atomic flags)
//   7. scan + emit_kernel   subtrees of <= leaf_size primitives collapse into leaves (their range is
//                           contiguous in sorted order); live nodes are compacted and written as
//                           64-byte two-child-box nodes
//   8. depth_kernel         deepest leaf, checked against the traversal stack
//
// What it replaces: the reference has no hierarchy — every ray tests every object and, after one
// AABB test per mesh, every triangle of the mesh (inc/ray_cast.hpp:37-52, inc/default_schema.hpp:125-144).
#include <chrono>
#include <cstdio>
#include <vector>
#include <type_traits>
#include "alloc.cuh"
#include "bvh.cuh"

namespace ctb {

#ifdef CTB_TIMING   // developer instrumentation, see api.cu
struct BuildTimer {
  std::chrono::high_resolution_clock::time_point t = std::chrono::high_resolution_clock::now();
  void lap(const char *what, cudaStream_t st) {
    cudaStreamSynchronize(st);
    auto n = std::chrono::high_resolution_clock::now();
    fprintf(stderr, "    [ctb-build] %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
#define BLAP(x) btimer.lap(x, st)
#else
struct BuildTimer {};
#define BLAP(x)
#endif

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
      rc = (e_ == cudaErrorMemoryAllocation) ? CUTRACE_ERR_OUT_OF_MEMORY : CUTRACE_ERR_CUDA; \
      goto done;                                                                         \
    }                                                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float ord2f_host(unsigned int u) {
  unsigned int v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  float f;
  memcpy(&f, &v, 4);
  return f;
}

struct Bounds6 { unsigned int lo[3], hi[3], clo[3], chi[3], bad_value, bad_index; };  // boxes as ordered ints + input errors

__global__ void init_bounds_kernel(Bounds6 *b) {
  if (threadIdx.x < 3) {
    b->lo[threadIdx.x] = 0xffffffffu; b->hi[threadIdx.x] = 0u;
    b->clo[threadIdx.x] = 0xffffffffu; b->chi[threadIdx.x] = 0u;
  }
  if (threadIdx.x == 0) { b->bad_value = 0u; b->bad_index = 0u; }
}

__global__ void prim_bounds_kernel(const float *__restrict__ p1, const float *__restrict__ p2,
                                   const float *__restrict__ p3, uint32_t n_tri,
                                   const float *__restrict__ sc, const float *__restrict__ sr, uint32_t n_sph,
                                   const uint32_t *__restrict__ tri_obj, const uint32_t *__restrict__ sph_obj, uint32_t n_objects,
                                   float4 *__restrict__ lo, float4 *__restrict__ hi, Bounds6 *bounds) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = n_tri + n_sph;
  float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
  bool valid = i < n;
  if (valid) {
    if (i < n_tri) {
      for (int c = 0; c < 3; c++) {
        float a = p1[3 * (size_t)i + c], b = p2[3 * (size_t)i + c], d = p3[3 * (size_t)i + c];
        l[c] = fminf(fminf(a, b), d);
        h[c] = fmaxf(fmaxf(a, b), d);
      }
    } else {
      uint32_t s = i - n_tri;
      float r = fabsf(sr[s]);
      r = r + r * 1e-6f;  // sphere roots are computed along the normalised direction; keep a margin
      for (int c = 0; c < 3; c++) { l[c] = sc[3 * (size_t)s + c] - r; h[c] = sc[3 * (size_t)s + c] + r; }
    }
    lo[i] = make_float4(l[0], l[1], l[2], 0.f);
    hi[i] = make_float4(h[0], h[1], h[2], 0.f);
    // input validation happens here, next to the data (a host pass over 10 M triangles costs 50 ms)
    if (!(isfinite(l[0]) && isfinite(l[1]) && isfinite(l[2]) && isfinite(h[0]) && isfinite(h[1]) && isfinite(h[2]))) {
      atomicAdd(&bounds->bad_value, 1u);
      valid = false;
    }
    if ((i < n_tri ? tri_obj[i] : sph_obj[i - n_tri]) >= n_objects) atomicAdd(&bounds->bad_index, 1u);
  }
  // warp reduce, then one set of atomics per BLOCK (3.8 M same-address atomics for 10 M primitives cost 2.4 ms)
  __shared__ float red[4][3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = 0; c < 3; c++) {
    float cl = valid ? 0.5f * (l[c] + h[c]) : INFINITY, ch = valid ? 0.5f * (l[c] + h[c]) : -INFINITY;
    float fl = l[c], fh = h[c];
    for (int o = 16; o > 0; o >>= 1) {
      fl = fminf(fl, __shfl_xor_sync(0xffffffffu, fl, o));
      fh = fmaxf(fh, __shfl_xor_sync(0xffffffffu, fh, o));
      cl = fminf(cl, __shfl_xor_sync(0xffffffffu, cl, o));
      ch = fmaxf(ch, __shfl_xor_sync(0xffffffffu, ch, o));
    }
    if (lane == 0) { red[0][c][warp] = fl; red[1][c][warp] = fh; red[2][c][warp] = cl; red[3][c][warp] = ch; }
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    const int q = threadIdx.x / 3, c = threadIdx.x % 3, nw = blockDim.x >> 5;
    float v = red[q][c][0];
    for (int w = 1; w < nw; w++) v = (q & 1) ? fmaxf(v, red[q][c][w]) : fminf(v, red[q][c][w]);
    if (isfinite(v)) {
      unsigned int *dst = q == 0 ? &bounds->lo[c] : q == 1 ? &bounds->hi[c] : q == 2 ? &bounds->clo[c] : &bounds->chi[c];
      if (q & 1) atomicMax(dst, f2ord(v)); else atomicMin(dst, f2ord(v));
    }
  }
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void morton_kernel(const float4 *__restrict__ lo, const float4 *__restrict__ hi, uint32_t n,
                              float3 clo, float3 cinv, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 l = lo[i], h = hi[i];
  float cx = (0.5f * (l.x + h.x) - clo.x) * cinv.x;
  float cy = (0.5f * (l.y + h.y) - clo.y) * cinv.y;
  float cz = (0.5f * (l.z + h.z) - clo.z) * cinv.z;
  const float S = 2097152.0f;  // 2^21
  unsigned long long x = (unsigned long long)fminf(fmaxf(cx * S, 0.f), S - 1.f);
  unsigned long long y = (unsigned long long)fminf(fmaxf(cy * S, 0.f), S - 1.f);
  unsigned long long z = (unsigned long long)fminf(fmaxf(cz * S, 0.f), S - 1.f);
  keys[i] = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
  vals[i] = i;
}

// ---------------------------------------------------------------------------------------------
// exclusive scan of uint32 (used by the radix sort and by node compaction)
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total, uint32_t *smem /*>=32*/) {
  uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
    uint32_t wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= (uint32_t)o) wi += t;
    }
    smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  uint32_t res = incl - v + smem[warp];
  if (total) *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void scan_tiles_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n,
                                  uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t sm[33];
  uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], sum = 0;
  for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (base + k < n) ? in[base + k] : 0; sum += v[k]; }
  uint32_t total;
  uint32_t off = block_exclusive_scan(sum, &total, sm);
  for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = off; off += v[k]; }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_sums_kernel(uint32_t *__restrict__ sums, uint32_t n_tiles, uint32_t *__restrict__ grand_total) {
  __shared__ uint32_t sm[33];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n_tiles ? sums[i] : 0, total;
    uint32_t off = block_exclusive_scan(v, &total, sm);
    if (i < n_tiles) sums[i] = off + carry;
    carry += total;
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry;
}

__global__ void scan_add_kernel(uint32_t *__restrict__ out, uint32_t n, const uint32_t *__restrict__ sums) {
  uint32_t i = blockIdx.x * SCAN_TILE + threadIdx.x;
  uint32_t add = sums[blockIdx.x];
  for (int k = 0; k < SCAN_ITEMS; k++, i += SCAN_THREADS)
    if (i < n) out[i] += add;
}

// out may alias in. tile_sums must hold ceil(n/SCAN_TILE) entries; d_total (optional) gets the sum.
static void exclusive_scan_u32(const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *tile_sums, uint32_t *d_total,
                               cudaStream_t st) {
  if (n == 0) { if (d_total) cudaMemsetAsync(d_total, 0, 4, st); return; }
  uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tiles_kernel<<<tiles, SCAN_THREADS, 0, st>>>(in, out, n, tile_sums);
  scan_sums_kernel<<<1, 1024, 0, st>>>(tile_sums, tiles, d_total);
  scan_add_kernel<<<tiles, SCAN_THREADS, 0, st>>>(out, n, tile_sums);
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort of (uint64 key, uint32 value), 8 bits per pass
// ---------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;    // 2048 keys per block
constexpr int RS_WARP_KEYS = 32 * RS_ROUNDS;       // 256 consecutive keys per warp

// block_hist[d * n_blocks + b] = number of keys with digit d in tile b
__global__ void rs_hist_kernel(const uint64_t *__restrict__ keys, uint32_t n, int shift, uint32_t *__restrict__ block_hist,
                               uint32_t n_blocks) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * RS_TILE;
  for (int r = 0; r < RS_ROUNDS; r++) {
    uint32_t i = base + r * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  block_hist[threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];
}

// block_hist now holds exclusive global offsets (digit-major). Order inside a tile: warp w owns keys
// [base + w*256, base + (w+1)*256) and walks them in rounds of 32 -> (warp, round, lane) = index order.
__global__ void rs_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                                  uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                                  int shift, const uint32_t *__restrict__ block_hist, uint32_t n_blocks) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * RS_TILE + warp * RS_WARP_KEYS;
  uint64_t key[RS_ROUNDS];
  uint32_t dig[RS_ROUNDS];
  for (int r = 0; r < RS_ROUNDS; r++) {
    uint32_t i = base + r * 32 + lane;
    bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0ull;
    dig[r] = ok ? ((uint32_t)(key[r] >> shift) & 255u) : (256u + lane);
    unsigned m = __match_any_sync(0xffffffffu, dig[r]);
    if (ok && (m & lt) == 0) cnt[warp][dig[r]] += __popc(m);
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: global tile offset, then exclusive over warps
    uint32_t d = threadIdx.x;
    uint32_t run = block_hist[d * n_blocks + blockIdx.x];
    for (int w = 0; w < RS_WARPS; w++) { uint32_t c = cnt[w][d]; cnt[w][d] = run; run += c; }
  }
  __syncthreads();
  for (int r = 0; r < RS_ROUNDS; r++) {
    uint32_t i = base + r * 32 + lane;
    bool ok = i < n;
    unsigned m = __match_any_sync(0xffffffffu, dig[r]);
    uint32_t pos = 0;
    if (ok) pos = cnt[warp][dig[r]] + __popc(m & lt);
    __syncwarp();
    if (ok && (m & lt) == 0) cnt[warp][dig[r]] += __popc(m);
    __syncwarp();
    if (ok) { keys_out[pos] = key[r]; vals_out[pos] = vals_in[i]; }
  }
}

// Up to SMALL_SORT_MAX pairs (the reference's own scenes: 1 .. 1005 primitives) are sorted by ONE block in shared memory
// instead of 8 passes x 5 launches: a bitonic network over (key, input position), which orders equal keys by position
// and therefore gives exactly the stable result of the LSD passes.  0.2 ms -> ~0.02 ms of the 0.4 ms LBVH build of bunny.json.
constexpr uint32_t SMALL_SORT_MAX = 2048;
__global__ void __launch_bounds__(1024) small_sort_kernel(uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t n) {
  __shared__ uint64_t sk[SMALL_SORT_MAX];
  __shared__ uint32_t sp[SMALL_SORT_MAX], sv[SMALL_SORT_MAX];
  uint32_t P = 2;
  while (P < n) P <<= 1;
  for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
    sk[i] = i < n ? keys[i] : ~0ull;     // padding sorts to the end (and behind real ~0 keys: larger position)
    sp[i] = i;
    sv[i] = i < n ? vals[i] : 0u;
  }
  __syncthreads();
  for (uint32_t k = 2; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        const uint32_t q = i ^ j;
        if (q > i) {
          const uint64_t a = sk[i], b = sk[q];
          const uint32_t pa = sp[i], pb = sp[q];
          const bool up = (i & k) == 0;                       // ascending run
          const bool a_gt_b = a > b || (a == b && pa > pb);
          if (a_gt_b == up) { sk[i] = b; sk[q] = a; sp[i] = pb; sp[q] = pa; }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { keys[i] = sk[i]; vals[i] = sv[sp[i]]; }
}

int radix_sort_pairs(uint64_t *d_keys, uint32_t *d_vals, uint32_t n, cudaStream_t st, std::string &err) {
  int rc = CUTRACE_OK;
  if (n < 2) return rc;
  if (n <= SMALL_SORT_MAX) {
    small_sort_kernel<<<1, 1024, 0, st>>>(d_keys, d_vals, n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { err = std::string("small_sort_kernel: ") + cudaGetErrorString(e); return CUTRACE_ERR_CUDA; }
    return rc;
  }
  uint64_t *k2 = nullptr;
  uint32_t *v2 = nullptr, *hist = nullptr, *tile_sums = nullptr;
  uint32_t n_blocks = (n + RS_TILE - 1) / RS_TILE;
  uint32_t table = 256u * n_blocks;
  uint64_t *kin = d_keys, *kout = nullptr;
  uint32_t *vin = d_vals, *vout = nullptr;
  CK(dmalloc(&k2, sizeof(uint64_t) * n, st));
  CK(dmalloc(&v2, sizeof(uint32_t) * n, st));
  CK(dmalloc(&hist, sizeof(uint32_t) * table, st));
  CK(dmalloc(&tile_sums, sizeof(uint32_t) * ((table + SCAN_TILE - 1) / SCAN_TILE + 1), st));
  kout = k2; vout = v2;
  for (int pass = 0; pass < 8; pass++) {
    int shift = pass * 8;
    rs_hist_kernel<<<n_blocks, RS_THREADS, 0, st>>>(kin, n, shift, hist, n_blocks);
    exclusive_scan_u32(hist, hist, table, tile_sums, nullptr, st);
    rs_scatter_kernel<<<n_blocks, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, n_blocks);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  CK(cudaGetLastError());
  // 8 passes = even number of swaps: the result is back in d_keys / d_vals
  CK(cudaStreamSynchronize(st));
done:
  dfree(k2, st); dfree(v2, st); dfree(hist, st); dfree(tile_sums, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// gather sorted primitive records
// ---------------------------------------------------------------------------------------------
__global__ void gather_kernel(const uint32_t *__restrict__ vals, uint32_t n, const float *__restrict__ p1,
                              const float *__restrict__ p2, const float *__restrict__ p3,
                              const uint32_t *__restrict__ tri_obj, uint32_t n_tri, const float *__restrict__ sc,
                              const float *__restrict__ sr, const uint32_t *__restrict__ sph_obj,
                              const float4 *__restrict__ lo, const float4 *__restrict__ hi, PrimRec *__restrict__ prims,
                              float4 *__restrict__ leaf_lo, float4 *__restrict__ leaf_hi) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t src = vals[k];
  PrimRec r;
  if (src < n_tri) {
    size_t o = 3 * (size_t)src;
    r.p1x = p1[o]; r.p1y = p1[o + 1]; r.p1z = p1[o + 2];
    r.p2x = p2[o]; r.p2y = p2[o + 1]; r.p2z = p2[o + 2];
    r.p3x = p3[o]; r.p3y = p3[o + 1]; r.p3z = p3[o + 2];
    r.obj = tri_obj[src]; r.idx = src; r.kind = CTB_PRIM_TRI;
  } else {
    uint32_t s = src - n_tri;
    size_t o = 3 * (size_t)s;
    r.p1x = sc[o]; r.p1y = sc[o + 1]; r.p1z = sc[o + 2];
    r.p2x = sr[s]; r.p2y = 0.f; r.p2z = 0.f;
    r.p3x = r.p3y = r.p3z = 0.f;
    r.obj = sph_obj[s]; r.idx = s; r.kind = CTB_PRIM_SPHERE;
  }
  prims[k] = r;
  leaf_lo[k] = lo[src];
  leaf_hi[k] = hi[src];
}

// ---------------------------------------------------------------------------------------------
// Karras hierarchy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t *__restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  uint64_t a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz((unsigned)i ^ (unsigned)j);
  return __clzll((long long)(a ^ b));
}

// children: >= 0 internal node index, < 0 leaf ~k
__global__ void karras_kernel(const uint64_t *__restrict__ keys, int n, int2 *__restrict__ children,
                              int2 *__restrict__ range, int *__restrict__ parent_node, int *__restrict__ parent_leaf) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + min(d, 0);
  int lo = min(i, j), hi = max(i, j);
  int left = (lo == gamma) ? ~gamma : gamma;
  int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  children[i] = make_int2(left, right);
  range[i] = make_int2(lo, hi);
  if (left >= 0) parent_node[left] = i; else parent_leaf[~left] = i;
  if (right >= 0) parent_node[right] = i; else parent_leaf[~right] = i;
  if (i == 0) parent_node[0] = -1;
}

__global__ void refit_kernel(int n, const int2 *__restrict__ children, const int *__restrict__ parent_node,
                             const int *__restrict__ parent_leaf, const float4 *leaf_lo, const float4 *leaf_hi,
                             float4 *node_lo, float4 *node_hi, unsigned int *flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int node = parent_leaf[k];
  while (node >= 0) {
    __threadfence();
    if (atomicAdd(&flags[node], 1u) == 0u) return;  // first arrival: the sibling subtree is not done yet
    __threadfence();
    int2 c = children[node];
    float4 l0 = c.x >= 0 ? __ldcg(&node_lo[c.x]) : __ldcg(&leaf_lo[~c.x]);
    float4 h0 = c.x >= 0 ? __ldcg(&node_hi[c.x]) : __ldcg(&leaf_hi[~c.x]);
    float4 l1 = c.y >= 0 ? __ldcg(&node_lo[c.y]) : __ldcg(&leaf_lo[~c.y]);
    float4 h1 = c.y >= 0 ? __ldcg(&node_hi[c.y]) : __ldcg(&leaf_hi[~c.y]);
    __stcg(&node_lo[node], make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f));
    __stcg(&node_hi[node], make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f));
    node = parent_node[node];
  }
}

__global__ void live_kernel(int n_internal, const int2 *__restrict__ range, uint32_t leaf_size, uint32_t *__restrict__ live) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  int2 r = range[i];
  live[i] = (uint32_t)(r.y - r.x + 1) > leaf_size ? 1u : 0u;
}

__global__ void emit_kernel(int n_internal, const int2 *__restrict__ children, const int2 *__restrict__ range,
                            const uint32_t *__restrict__ compact, uint32_t leaf_size, const float4 *__restrict__ leaf_lo,
                            const float4 *__restrict__ leaf_hi, const float4 *__restrict__ node_lo,
                            const float4 *__restrict__ node_hi, float eps, Node *__restrict__ nodes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  int2 r = range[i];
  if ((uint32_t)(r.y - r.x + 1) <= leaf_size) return;  // collapsed into a leaf of an ancestor
  int2 c = children[i];
  int ref[2];
  float4 lo[2], hi[2];
  int cc[2] = {c.x, c.y};
  for (int s = 0; s < 2; s++) {
    int ch = cc[s];
    if (ch < 0) {
      ref[s] = leaf_encode((uint32_t)~ch, 1u);
      lo[s] = leaf_lo[~ch]; hi[s] = leaf_hi[~ch];
    } else {
      int2 cr = range[ch];
      uint32_t cnt = (uint32_t)(cr.y - cr.x + 1);
      ref[s] = cnt <= leaf_size ? leaf_encode((uint32_t)cr.x, cnt) : (int)compact[ch];
      lo[s] = node_lo[ch]; hi[s] = node_hi[ch];
    }
  }
  Node nd;
  nd.n0xy = make_float4(lo[0].x - eps, hi[0].x + eps, lo[0].y - eps, hi[0].y + eps);
  nd.n1xy = make_float4(lo[1].x - eps, hi[1].x + eps, lo[1].y - eps, hi[1].y + eps);
  nd.nz = make_float4(lo[0].z - eps, hi[0].z + eps, lo[1].z - eps, hi[1].z + eps);
  nd.meta = make_int4(ref[0], ref[1], r.x, r.y);  // z,w = primitive range (debug / validation)
  nodes[compact[i]] = nd;
}

// ---- SAH treelets -------------------------------------------------------------------------------------------------------
// Karras' hierarchy splits a node where its Morton prefix ends — a spatial-median split on a fixed axis cycle.  Its lower
// levels decide most of a ray's cost (every ray that reaches a mesh walks them), and there a sweep-SAH tree is ~20 % cheaper
// (expected cost of bunny.json's 1000 triangles: 29.2 -> 23.3, profiles/r02_tuning.md).  So every maximal subtree of at most
// SAH_T primitives (a "treelet": the whole tree of the reference scenes, one mesh instance of the 10 M-triangle hall) gets
// its TOPOLOGY rebuilt by one CTA: full-sweep SAH over the three centroid orders, top down, level by level, one primitive
// per thread.  The treelet keeps its root index and its primitive range [lo, hi] in the sorted order; the primitives are
// permuted inside that range (leaves of the new tree, left to right) and the internal nodes lo+1 .. hi-1 (Karras numbering:
// the indices inside a subtree are its range minus the two endpoints, plus the root) are re-linked.  Everything downstream
// (gather, refit, leaf collapse, emit) works on the Karras arrays as before.
#ifndef CTB_SAH_TREELETS
#define CTB_SAH_TREELETS 1
#endif
#define SAH_T 1024
struct SahShared {
  float blo[3][SAH_T], bhi[3][SAH_T];   // primitive boxes by local id
  float key[SAH_T];                     // sort keys / scratch
  unsigned short ord[3][SAH_T];         // local ids per axis, partitioned by segment
  unsigned short tmp[SAH_T];
  unsigned short sa[SAH_T], sb[SAH_T];  // segment [sa, sb) a position belongs to
  short pg[SAH_T];                      // parent gap of that segment (-1: the treelet root)
  unsigned char side[SAH_T];            // which child of the parent the segment is
  unsigned char left[SAH_T];            // by local id: goes to the left child in this level
  float parea[SAH_T], sarea[SAH_T];     // prefix / suffix box areas in the current axis order
  float scan[6][32];                    // warp aggregates of a box scan
  int cnt[32];                          // ... of a count scan
  unsigned char side2[32];              // ... their "a head lies in the window" flags
  unsigned long long best[SAH_T];       // by segment start: min over (cost, axis, position)
  uint32_t vals_in[SAH_T];
  int open;                             // segments with more than one primitive
  int g_root;                           // gap (split position - 1) of the treelet root
};

__device__ __forceinline__ float box_area6(const float b[6]) {
  const float dx = fmaxf(b[3] - b[0], 0.f), dy = fmaxf(b[4] - b[1], 0.f), dz = fmaxf(b[5] - b[2], 0.f);
  return dx * dy + dy * dz + dz * dx;
}

// roots of the treelets: internal nodes of at most SAH_T (and more than leaf_size) primitives whose parent is bigger
__global__ void sah_roots_kernel(int n_internal, const int2 *__restrict__ range, const int *__restrict__ parent_node, uint32_t leaf_size,
                                 uint32_t *__restrict__ roots, uint32_t *__restrict__ n_roots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  const int2 r = range[i];
  const uint32_t cnt = (uint32_t)(r.y - r.x + 1);
  if (cnt > SAH_T || cnt <= leaf_size || cnt < 3) return;
  const int p = parent_node[i];
  if (p >= 0) { const int2 pr = range[p]; if ((uint32_t)(pr.y - pr.x + 1) <= SAH_T) return; }
  roots[atomicAdd(n_roots, 1u)] = (uint32_t)i;
}

// Block-wide segmented inclusive scans (1024 threads, one element each, index order): a warp scan with shuffles over
// (value, "a segment head lies in my window") pairs, the 32 warp aggregates scanned by warp 0, one fix-up — three barriers
// per scan.  The first version used Hillis-Steele steps through shared memory: twenty barriers per scan, six box scans and
// three count scans per level, ~0.6 ms per treelet — 25 ms on the 10,112 treelets of the hall.
struct SahBox { float v[6]; };
__device__ __forceinline__ void sah_box_merge(SahBox &into, const SahBox &o) {   // into = o (+) into
#pragma unroll
  for (int c = 0; c < 3; c++) { into.v[c] = fminf(into.v[c], o.v[c]); into.v[c + 3] = fmaxf(into.v[c + 3], o.v[c + 3]); }
}
__device__ __forceinline__ SahBox sah_seg_scan_box(SahBox x, bool head, float (*wagg)[32], unsigned char *wflag) {
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  bool f = head;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    SahBox nb;
#pragma unroll
    for (int c = 0; c < 6; c++) nb.v[c] = __shfl_up_sync(0xffffffffu, x.v[c], d);
    const bool nf = __shfl_up_sync(0xffffffffu, f, d);
    if (lane >= (unsigned)d && !f) { sah_box_merge(x, nb); f = nf; }
  }
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < 6; c++) wagg[c][w] = x.v[c];
    wflag[w] = f ? 1 : 0;
  }
  __syncthreads();
  if (w == 0) {   // scan of the warp aggregates
    SahBox a;
#pragma unroll
    for (int c = 0; c < 6; c++) a.v[c] = wagg[c][lane];
    bool af = wflag[lane] != 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      SahBox nb;
#pragma unroll
      for (int c = 0; c < 6; c++) nb.v[c] = __shfl_up_sync(0xffffffffu, a.v[c], d);
      const bool nf = __shfl_up_sync(0xffffffffu, af, d);
      if (lane >= (unsigned)d && !af) { sah_box_merge(a, nb); af = nf; }
    }
#pragma unroll
    for (int c = 0; c < 6; c++) wagg[c][lane] = a.v[c];
  }
  __syncthreads();
  if (w > 0 && !f) {   // no head between the start of my warp and me: the carry of the warps before applies
    SahBox cbox;
#pragma unroll
    for (int c = 0; c < 6; c++) cbox.v[c] = wagg[c][w - 1];
    sah_box_merge(x, cbox);
  }
  __syncthreads();
  return x;
}
__device__ __forceinline__ int sah_seg_scan_int(int x, bool head, int *wsum, unsigned char *wflag) {
  const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  bool f = head;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int nb = __shfl_up_sync(0xffffffffu, x, d);
    const bool nf = __shfl_up_sync(0xffffffffu, f, d);
    if (lane >= (unsigned)d && !f) { x += nb; f = nf; }
  }
  if (lane == 31) { wsum[w] = x; wflag[w] = f ? 1 : 0; }
  __syncthreads();
  if (w == 0) {
    int a = wsum[lane];
    bool af = wflag[lane] != 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int nb = __shfl_up_sync(0xffffffffu, a, d);
      const bool nf = __shfl_up_sync(0xffffffffu, af, d);
      if (lane >= (unsigned)d && !af) { a += nb; af = nf; }
    }
    wsum[lane] = a;
  }
  __syncthreads();
  if (w > 0 && !f) x += wsum[w - 1];
  __syncthreads();
  return x;
}

__global__ void __launch_bounds__(SAH_T, 1)
sah_treelet_kernel(const uint32_t *__restrict__ roots, const float4 *__restrict__ g_lo, const float4 *__restrict__ g_hi,
                   uint32_t *__restrict__ vals, int2 *__restrict__ children, int2 *__restrict__ range, int *__restrict__ parent_node,
                   int *__restrict__ parent_leaf) {
  extern __shared__ unsigned char sah_raw[];
  SahShared &S = *reinterpret_cast<SahShared *>(sah_raw);
  const int r = (int)roots[blockIdx.x];
  const int2 rr = range[r];
  const int lo = rr.x, m = rr.y - rr.x + 1;
  const int t = threadIdx.x;
  const bool on = t < m;
  // internal node index of a gap of this treelet (see the header comment); the root keeps r
  auto node_of_gap = [&](int g) -> int { return g == S.g_root ? r : lo + 1 + (g < S.g_root ? g : g - 1); };

  // ---- load ----
  if (on) {
    const uint32_t id = vals[lo + t];
    S.vals_in[t] = id;
    const float4 l = g_lo[id], h = g_hi[id];
    S.blo[0][t] = l.x; S.blo[1][t] = l.y; S.blo[2][t] = l.z; S.bhi[0][t] = h.x; S.bhi[1][t] = h.y; S.bhi[2][t] = h.z;
    S.sa[t] = 0; S.sb[t] = (unsigned short)m; S.pg[t] = -1; S.side[t] = 0;
  }
  if (t == 0) { S.open = 1; S.g_root = -1; }
  __syncthreads();
  // ---- three centroid orders: bitonic sort of (key, id) over the next power of two ----
  int P = 1;
  while (P < m) P <<= 1;
  for (int ax = 0; ax < 3; ax++) {
    S.key[t] = on ? S.blo[ax][t] + S.bhi[ax][t] : INFINITY;
    S.tmp[t] = (unsigned short)t;
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        const int x = t ^ j;
        if (t < P && x > t) {
          const bool up = (t & k) == 0;
          const float a = S.key[t], b = S.key[x];
          const unsigned short ia = S.tmp[t], ib = S.tmp[x];
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) { S.key[t] = b; S.key[x] = a; S.tmp[t] = ib; S.tmp[x] = ia; }
        }
        __syncthreads();
      }
    if (on) S.ord[ax][t] = S.tmp[t];
    __syncthreads();
  }

  // ---- top down, one level per iteration ----
  const int q = m - 1 - t;   // the position this thread handles in the suffix (reversed) scans
  for (int level = 0; level < SAH_T && S.open > 0; level++) {
    const int a = on ? S.sa[t] : 0, b = on ? S.sb[t] : 0;
    const bool live = on && b - a > 1;
    if (on && t == a) S.best[a] = ~0ull;
    __syncthreads();
    for (int ax = 0; ax < 3; ax++) {
      SahBox x;
      {   // prefix boxes in this axis' order
        const int id = on ? S.ord[ax][t] : 0;
#pragma unroll
        for (int c = 0; c < 3; c++) { x.v[c] = on ? S.blo[c][id] : INFINITY; x.v[c + 3] = on ? S.bhi[c][id] : -INFINITY; }
        x = sah_seg_scan_box(x, !on || t == a, S.scan, S.side2);
        if (on) S.parea[t] = box_area6(x.v);
      }
      {   // suffix boxes: the same scan over the reversed positions
        const bool qon = q >= 0;
        const int id = qon ? S.ord[ax][q] : 0;
#pragma unroll
        for (int c = 0; c < 3; c++) { x.v[c] = qon ? S.blo[c][id] : INFINITY; x.v[c + 3] = qon ? S.bhi[c][id] : -INFINITY; }
        x = sah_seg_scan_box(x, !qon || q == (int)S.sb[q] - 1, S.scan, S.side2);
        if (qon) S.sarea[q] = box_area6(x.v);
      }
      __syncthreads();
      // cost of splitting after position t (left = [a, t], right = [t+1, b))
      if (live && t < b - 1) {
        const float c = S.parea[t] * (float)(t - a + 1) + S.sarea[t + 1] * (float)(b - t - 1);
        const unsigned long long packed = ((unsigned long long)__float_as_uint(fmaxf(c, 0.f)) << 32) | ((unsigned long long)ax << 16) | (unsigned long long)t;
        atomicMin(&S.best[a], packed);
      }
      __syncthreads();
    }
    // ---- the split of my segment ----
    int ax_s = 0, nleft = 0;
    if (live) {
      const unsigned long long bst = S.best[a];
      ax_s = (int)((bst >> 16) & 3ull);
      nleft = (int)(bst & 0xffffull) - a + 1;
      S.left[S.ord[ax_s][t]] = (t - a) < nleft ? 1 : 0;
    }
    __syncthreads();
    // ---- stable partition of the three orders by the side flags ----
    for (int ax = 0; ax < 3; ax++) {
      const int f = live ? (int)S.left[S.ord[ax][t]] : 0;
      const int run = sah_seg_scan_int(f, !on || t == a, S.cnt, S.side2);
      if (live) {
        const int before = run - f;
        const int np = f ? a + before : a + nleft + (t - a - before);
        S.tmp[np] = S.ord[ax][t];
      }
      __syncthreads();
      if (live) S.ord[ax][t] = S.tmp[t];
      __syncthreads();
    }
    // ---- link the new node, hand the two halves down ----
    if (t == 0) S.open = 0;
    if (live && t == a && S.pg[t] < 0) S.g_root = a + nleft - 1;
    __syncthreads();
    if (live) {
      const int g = a + nleft - 1;                 // gap of this node
      const int me = node_of_gap(g);
      if (t == a) {
        range[me] = make_int2(lo + a, lo + b - 1);
        const int pgap = S.pg[t];
        if (pgap >= 0) {
          const int par = node_of_gap(pgap);
          if (S.side[t] == 0) children[par].x = me; else children[par].y = me;
          parent_node[me] = par;
        }
      }
    }
    __syncthreads();   // (every leader has read its parent gap before the segments are renamed)
    if (live) {
      const int g = a + nleft - 1;
      const int me = node_of_gap(g);
      const bool is_left = (t - a) < nleft;
      const int na = is_left ? a : a + nleft, nbd = is_left ? a + nleft : b;
      S.sa[t] = (unsigned short)na; S.sb[t] = (unsigned short)nbd; S.pg[t] = (short)g; S.side[t] = is_left ? 0 : 1;
      if (nbd - na == 1) {                         // a single primitive: a leaf of the tree
        if (is_left) children[me].x = ~(lo + t); else children[me].y = ~(lo + t);
        parent_leaf[lo + t] = me;
      } else if (t == na) {
        atomicAdd(&S.open, 1);
      }
    }
    __syncthreads();
  }
  // ---- the new left-to-right order of the primitives ----
  if (on) vals[lo + t] = S.vals_in[S.ord[0][t]];
}

// ---- 4-wide collapse (CTB_BVH4) ----------------------------------------------------------------------------------------
// A live internal node of the binary tree is KEPT when its depth is even; a kept node adopts the children of its live
// internal children (odd depth), so it ends up with 2..4 slots: leaves (single primitives or collapsed subtrees) and kept
// grandchildren.  Every ancestor of a live node is live, so the depth is the length of the parent chain.
__global__ void kept_kernel(int n_internal, const int2 *__restrict__ range, const int *__restrict__ parent_node, uint32_t leaf_size,
                            uint32_t *__restrict__ kept) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  int2 r = range[i];
  unsigned int d = 0;
  if ((uint32_t)(r.y - r.x + 1) > leaf_size)
    for (int p = parent_node[i]; p >= 0; p = parent_node[p]) d++;
  else
    d = 1;   // not live
  kept[i] = (d & 1u) ? 0u : 1u;
}

__global__ void emit4_kernel(int n_internal, const int2 *__restrict__ children, const int2 *__restrict__ range, const uint32_t *__restrict__ kept,
                             const uint32_t *__restrict__ idx4, uint32_t leaf_size, const float4 *__restrict__ leaf_lo,
                             const float4 *__restrict__ leaf_hi, const float4 *__restrict__ node_lo, const float4 *__restrict__ node_hi,
                             float eps, Node4 *__restrict__ nodes4) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal || !kept[i]) return;
  int ref[4] = {CTB_SENTINEL, CTB_SENTINEL, CTB_SENTINEL, CTB_SENTINEL};
  float lo[4][3], hi[4][3];
  for (int s = 0; s < 4; s++)
    for (int a = 0; a < 3; a++) { lo[s][a] = INFINITY; hi[s][a] = -INFINITY; }
  int n = 0;
  auto add = [&](int r_, float4 l, float4 h) {
    ref[n] = r_;
    lo[n][0] = l.x - eps; lo[n][1] = l.y - eps; lo[n][2] = l.z - eps;
    hi[n][0] = h.x + eps; hi[n][1] = h.y + eps; hi[n][2] = h.z + eps;
    n++;
  };
  // one child of a binary node as a slot: leaf, collapsed subtree, or (expand == false) a kept node
  auto visit = [&](int ch, bool may_expand, auto &&self) -> void {
    if (ch < 0) { add(leaf_encode((uint32_t)~ch, 1u), leaf_lo[~ch], leaf_hi[~ch]); return; }
    int2 cr = range[ch];
    uint32_t cnt = (uint32_t)(cr.y - cr.x + 1);
    if (cnt <= leaf_size) { add(leaf_encode((uint32_t)cr.x, cnt), node_lo[ch], node_hi[ch]); return; }
    if (may_expand) {   // live internal child at odd depth: adopt its two children
      int2 g = children[ch];
      self(g.x, false, self);
      self(g.y, false, self);
    } else {            // live internal grandchild at even depth: a kept node
      add((int)idx4[ch], node_lo[ch], node_hi[ch]);
    }
  };
  int2 c = children[i];
  visit(c.x, true, visit);
  visit(c.y, true, visit);
  Node4 nd;
  nd.lox = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]); nd.hix = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
  nd.loy = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]); nd.hiy = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
  nd.loz = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]); nd.hiz = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
  nd.ref = make_int4(ref[0], ref[1], ref[2], ref[3]);
  nd.pad = make_int4(range[i].x, range[i].y, n, 0);
  nodes4[idx4[i]] = nd;
}

__global__ void depth_kernel(int n_internal, const int2 *__restrict__ range, const int *__restrict__ parent_node,
                             uint32_t leaf_size, unsigned int *max_depth) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_internal) return;
  int2 r = range[i];
  if ((uint32_t)(r.y - r.x + 1) <= leaf_size) return;
  unsigned int d = 1;
  for (int p = parent_node[i]; p >= 0; p = parent_node[p]) d++;
  atomicMax(max_depth, d);
}

// ---------------------------------------------------------------------------------------------
// BFS relabelling of the top of the tree: the first `S` nodes in breadth-first order move to the front of the node
// array (indices 0..S-1, root stays 0) so that a kernel can stage "the top of the BVH" in shared memory with one
// contiguous copy; every other node keeps its relative (Morton) order behind them.
// ---------------------------------------------------------------------------------------------
__global__ void bfs_top_kernel(const Node *__restrict__ nodes, uint32_t n_nodes, uint32_t S, uint32_t *__restrict__ order,
                               uint32_t *__restrict__ n_top_out) {
  __shared__ uint32_t sm[33];
  __shared__ uint32_t head, tail;
  if (threadIdx.x == 0) { order[0] = 0; head = 0; tail = 1; }
  __syncthreads();
  while (head < tail && tail < S) {
    const uint32_t h0 = head, t0 = tail;
    uint32_t produced = 0;
    for (uint32_t base = h0; base < t0; base += blockDim.x) {   // deterministic order: parent order, child 0 before child 1
      const uint32_t i = base + threadIdx.x;
      int c0 = -1, c1 = -1;
      if (i < t0) { const int4 m = nodes[order[i]].meta; c0 = m.x; c1 = m.y; }
      const uint32_t cnt = (c0 >= 0 ? 1u : 0u) + (c1 >= 0 ? 1u : 0u);
      uint32_t total;
      uint32_t off = block_exclusive_scan(cnt, &total, sm) + t0 + produced;
      if (c0 >= 0) { if (off < S) order[off] = (uint32_t)c0; off++; }
      if (c1 >= 0) { if (off < S) order[off] = (uint32_t)c1; }
      produced += total;
    }
    __syncthreads();
    if (threadIdx.x == 0) { head = t0; tail = t0 + produced < S ? t0 + produced : S; }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_top_out = tail;
}

__global__ void top_flags_kernel(const uint32_t *__restrict__ order, uint32_t n_top, uint32_t *__restrict__ newidx,
                                 uint32_t *__restrict__ rest_flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_top) { newidx[order[i]] = i; rest_flag[order[i]] = 0u; }
}

__global__ void rest_index_kernel(const uint32_t *__restrict__ rest_flag, const uint32_t *__restrict__ rest_scan, uint32_t n_nodes,
                                  uint32_t n_top, uint32_t *__restrict__ newidx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes && rest_flag[i]) newidx[i] = n_top + rest_scan[i];
}

__global__ void permute_nodes_kernel(const Node *__restrict__ in, uint32_t n_nodes, const uint32_t *__restrict__ newidx,
                                     Node *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  Node nd = in[i];
  if (nd.meta.x >= 0) nd.meta.x = (int)newidx[nd.meta.x];
  if (nd.meta.y >= 0) nd.meta.y = (int)newidx[nd.meta.y];
  out[newidx[i]] = nd;
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int reorder_top(BvhResult &bvh, uint32_t S, cudaStream_t st, uint32_t *n_top_out, std::string &err) {
  int rc = CUTRACE_OK;
  *n_top_out = 0;
  if (bvh.root != 0 || bvh.n_nodes == 0 || S == 0) return rc;
  const uint32_t n = bvh.n_nodes;
  if (S > n) S = n;
  uint32_t *order = nullptr, *newidx = nullptr, *flag = nullptr, *scan = nullptr, *tile_sums = nullptr, *d_ntop = nullptr;
  Node *out = nullptr;
  uint32_t n_top = 0;
  const int T = 256;
  const uint32_t nb = (n + T - 1) / T;
  CK(dmalloc(&order, sizeof(uint32_t) * S, st));
  CK(dmalloc(&newidx, sizeof(uint32_t) * n, st));
  CK(dmalloc(&flag, sizeof(uint32_t) * n, st));
  CK(dmalloc(&scan, sizeof(uint32_t) * n, st));
  CK(dmalloc(&tile_sums, sizeof(uint32_t) * ((n + SCAN_TILE - 1) / SCAN_TILE + 1), st));
  CK(dmalloc(&d_ntop, sizeof(uint32_t), st));
  CK(dmalloc(&out, sizeof(Node) * n, st));
  bfs_top_kernel<<<1, 1024, 0, st>>>(bvh.nodes, n, S, order, d_ntop);
  CK(cudaMemcpyAsync(&n_top, d_ntop, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  fill_u32_kernel<<<nb, T, 0, st>>>(flag, n, 1u);
  top_flags_kernel<<<(n_top + T - 1) / T, T, 0, st>>>(order, n_top, newidx, flag);
  exclusive_scan_u32(flag, scan, n, tile_sums, nullptr, st);
  rest_index_kernel<<<nb, T, 0, st>>>(flag, scan, n, n_top, newidx);
  permute_nodes_kernel<<<nb, T, 0, st>>>(bvh.nodes, n, newidx, out);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(st));
  dfree(bvh.nodes, st);
  bvh.nodes = out;
  out = nullptr;
  *n_top_out = n_top;
done:
  dfree(order, st); dfree(newidx, st); dfree(flag, st); dfree(scan, st); dfree(tile_sums, st); dfree(d_ntop, st); dfree(out, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// validation: every primitive covered exactly once, child boxes contain their primitives / grandchildren
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prim_box(const PrimRec &p, float lo[3], float hi[3]) {
  if (p.kind == CTB_PRIM_TRI) {
    lo[0] = fminf(fminf(p.p1x, p.p2x), p.p3x); hi[0] = fmaxf(fmaxf(p.p1x, p.p2x), p.p3x);
    lo[1] = fminf(fminf(p.p1y, p.p2y), p.p3y); hi[1] = fmaxf(fmaxf(p.p1y, p.p2y), p.p3y);
    lo[2] = fminf(fminf(p.p1z, p.p2z), p.p3z); hi[2] = fmaxf(fmaxf(p.p1z, p.p2z), p.p3z);
  } else {
    float r = fabsf(p.p2x);
    lo[0] = p.p1x - r; hi[0] = p.p1x + r; lo[1] = p.p1y - r; hi[1] = p.p1y + r; lo[2] = p.p1z - r; hi[2] = p.p1z + r;
  }
}

__global__ void validate_kernel(const Node *__restrict__ nodes, uint32_t n_nodes, const PrimRec *__restrict__ prims,
                                uint32_t n_prims, unsigned int *cover, unsigned int *node_refs, unsigned int *errors) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  Node nd = nodes[i];
  int ch[2] = {nd.meta.x, nd.meta.y};
  float blo[2][3] = {{nd.n0xy.x, nd.n0xy.z, nd.nz.x}, {nd.n1xy.x, nd.n1xy.z, nd.nz.z}};
  float bhi[2][3] = {{nd.n0xy.y, nd.n0xy.w, nd.nz.y}, {nd.n1xy.y, nd.n1xy.w, nd.nz.w}};
  for (int s = 0; s < 2; s++) {
    if (ch[s] < 0) {
      uint32_t f = leaf_first(ch[s]), c = leaf_count(ch[s]);
      for (uint32_t k = f; k < f + c; k++) {
        if (k >= n_prims) { atomicAdd(errors, 1u); continue; }
        atomicAdd(&cover[k], 1u);
        float lo[3], hi[3];
        prim_box(prims[k], lo, hi);
        for (int a = 0; a < 3; a++)
          if (!(lo[a] >= blo[s][a] && hi[a] <= bhi[s][a])) atomicAdd(errors, 1u);
      }
    } else {
      if ((uint32_t)ch[s] >= n_nodes || ch[s] == 0) { atomicAdd(errors, 1u); continue; }
      atomicAdd(&node_refs[ch[s]], 1u);
      Node cn = nodes[ch[s]];
      float clo[3] = {fminf(cn.n0xy.x, cn.n1xy.x), fminf(cn.n0xy.z, cn.n1xy.z), fminf(cn.nz.x, cn.nz.z)};
      float chi[3] = {fmaxf(cn.n0xy.y, cn.n1xy.y), fmaxf(cn.n0xy.w, cn.n1xy.w), fmaxf(cn.nz.y, cn.nz.w)};
      for (int a = 0; a < 3; a++)
        if (!(clo[a] >= blo[s][a] - 1e-30f && chi[a] <= bhi[s][a] + 1e-30f)) atomicAdd(errors, 1u);
    }
  }
}

__global__ void validate_cover_kernel(const unsigned int *cover, uint32_t n_prims, const unsigned int *node_refs,
                                      uint32_t n_nodes, unsigned int *errors) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_prims && cover[i] != 1u) atomicAdd(errors + 1, 1u);
  if (i > 0 && i < n_nodes && node_refs[i] != 1u) atomicAdd(errors + 2, 1u);
}

int validate_bvh(const BvhResult &bvh, cudaStream_t st, std::string &err) {
  int rc = CUTRACE_OK;
  unsigned int *cover = nullptr, *refs = nullptr, *errors = nullptr;
  unsigned int h_err[3] = {0, 0, 0};
  if (bvh.n_prims == 0) return rc;
  CK(dmalloc(&cover, sizeof(unsigned int) * bvh.n_prims, st));
  CK(dmalloc(&refs, sizeof(unsigned int) * (bvh.n_nodes + 1), st));
  CK(dmalloc(&errors, sizeof(unsigned int) * 3, st));
  CK(cudaMemsetAsync(cover, 0, sizeof(unsigned int) * bvh.n_prims, st));
  CK(cudaMemsetAsync(refs, 0, sizeof(unsigned int) * (bvh.n_nodes + 1), st));
  CK(cudaMemsetAsync(errors, 0, sizeof(unsigned int) * 3, st));
  if (bvh.root >= 0 && bvh.root != CTB_SENTINEL) {
    validate_kernel<<<(bvh.n_nodes + 255) / 256, 256, 0, st>>>(bvh.nodes, bvh.n_nodes, bvh.prims, bvh.n_prims, cover, refs, errors);
    uint32_t m = bvh.n_prims > bvh.n_nodes ? bvh.n_prims : bvh.n_nodes;
    validate_cover_kernel<<<(m + 255) / 256, 256, 0, st>>>(cover, bvh.n_prims, refs, bvh.n_nodes, errors);
  } else if (bvh.root < 0) {
    if (leaf_first(bvh.root) != 0 || leaf_count(bvh.root) != bvh.n_prims) h_err[1] = 1;
  }
  CK(cudaGetLastError());
  if (bvh.root >= 0 && bvh.root != CTB_SENTINEL) CK(cudaMemcpyAsync(h_err, errors, sizeof h_err, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (h_err[0] || h_err[1] || h_err[2]) {
    char buf[200];
    snprintf(buf, sizeof buf, "BVH validation failed: %u box/containment errors, %u primitives not covered exactly once, %u nodes not referenced exactly once",
             h_err[0], h_err[1], h_err[2]);
    err = buf;
    rc = CUTRACE_ERR_INTERNAL;
  }
done:
  dfree(cover, st); dfree(refs, st); dfree(errors, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// per-object boxes for the reference's mesh pre-test (trace.cuh: mesh_gate)
// ---------------------------------------------------------------------------------------------
// acc: n_objects x 8 words = lo.xyz (ordered), hi.xyz (ordered), triangle count, unused
__global__ void object_bounds_init_kernel(unsigned int *acc, uint32_t n_objects) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_objects * 8u) return;
  const uint32_t w = i & 7u;
  acc[i] = w < 3u ? 0xffffffffu : 0u;
}
// min / max over the vertices of every triangle of an object — exactly cpu::schema::mesh::bounding_box
// (inc/default_schema.hpp:573-586; min and max are exact in float).  A warp whose 32 triangles belong to one object (the
// normal case: meshes are contiguous) reduces first and issues one set of atomics.
__global__ void object_bounds_kernel(const float *__restrict__ p1, const float *__restrict__ p2, const float *__restrict__ p3,
                                     const uint32_t *__restrict__ tri_obj, uint32_t n_tri, uint32_t n_objects, unsigned int *acc) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n_tri;
  uint32_t obj = valid ? tri_obj[i] : 0xffffffffu;
  if (obj >= n_objects) obj = 0xffffffffu;   // reported by prim_bounds_kernel
  float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (obj != 0xffffffffu) {
    for (int c = 0; c < 3; c++) {
      const float a = p1[3 * (size_t)i + c], b = p2[3 * (size_t)i + c], d = p3[3 * (size_t)i + c];
      l[c] = fminf(fminf(a, b), d);
      h[c] = fmaxf(fmaxf(a, b), d);
    }
  }
  const uint32_t obj0 = __shfl_sync(0xffffffffu, obj, 0);
  if (__all_sync(0xffffffffu, obj == obj0)) {
    if (obj0 == 0xffffffffu) return;
    for (int c = 0; c < 3; c++)
      for (int o = 16; o > 0; o >>= 1) {
        l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], o));
        h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], o));
      }
    if ((threadIdx.x & 31) == 0) {
      unsigned int *a = acc + 8 * (size_t)obj0;
      for (int c = 0; c < 3; c++) { atomicMin(a + c, f2ord(l[c])); atomicMax(a + 3 + c, f2ord(h[c])); }
      atomicAdd(a + 6, 32u);
    }
  } else if (obj != 0xffffffffu) {
    unsigned int *a = acc + 8 * (size_t)obj;
    for (int c = 0; c < 3; c++) { atomicMin(a + c, f2ord(l[c])); atomicMax(a + 3 + c, f2ord(h[c])); }
    atomicAdd(a + 6, 1u);
  }
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__global__ void object_bounds_finish_kernel(const unsigned int *__restrict__ acc, const uint32_t *__restrict__ obj_kind, uint32_t n_objects,
                                            ObjBound *out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_objects) return;
  const unsigned int *a = acc + 8 * (size_t)i;
  ObjBound b;
  const bool any = a[6] != 0u;
  for (int c = 0; c < 3; c++) { b.lo[c] = any ? ord2f(a[c]) : INFINITY; b.hi[c] = any ? ord2f(a[3 + c]) : -INFINITY; }
  // a mesh gets the pre-test, a loose triangle does not; without obj_kind an object with exactly one triangle is a loose triangle
  b.is_mesh = obj_kind ? (obj_kind[i] == CUTRACE_OBJ_MESH ? 1u : 0u) : (a[6] > 1u ? 1u : 0u);
  b.pad = 0u;
  out[i] = b;
}

int build_object_bounds(const float *d_p1, const float *d_p2, const float *d_p3, const uint32_t *d_tri_obj, uint32_t n_tri, uint32_t n_objects,
                        const uint32_t *d_obj_kind, ObjBound **out, cudaStream_t st, std::string &err) {
  int rc = CUTRACE_OK;
  unsigned int *acc = nullptr;
  *out = nullptr;
  if (n_objects == 0) return rc;
  CK(dmalloc(out, sizeof(ObjBound) * n_objects, st));
  CK(dmalloc(&acc, sizeof(unsigned int) * 8 * (size_t)n_objects, st));
  object_bounds_init_kernel<<<(n_objects * 8u + 255u) / 256u, 256, 0, st>>>(acc, n_objects);
  if (n_tri) object_bounds_kernel<<<(n_tri + 255u) / 256u, 256, 0, st>>>(d_p1, d_p2, d_p3, d_tri_obj, n_tri, n_objects, acc);
  object_bounds_finish_kernel<<<(n_objects + 255u) / 256u, 256, 0, st>>>(acc, d_obj_kind, n_objects, *out);
  CK(cudaGetLastError());
done:
  dfree(acc, st);
  if (rc) { dfree(*out, st); *out = nullptr; }
  return rc;
}

// ---------------------------------------------------------------------------------------------
// driver
// ---------------------------------------------------------------------------------------------
int build_bvh(const BvhInput &in, BvhResult &out, std::string &err) {
  int rc = CUTRACE_OK;
  cudaStream_t st = in.stream;
  const uint32_t n = in.n_tri + in.n_sph;
  uint32_t leaf_size = in.leaf_size ? in.leaf_size : 4;
  if (leaf_size > CTB_MAX_LEAF) leaf_size = CTB_MAX_LEAF;
  out = BvhResult();
  out.n_prims = n;
  if (n == 0) return rc;
  if (n >= (1u << 28)) { err = "too many primitives for the leaf encoding (max 2^28-1)"; return CUTRACE_ERR_INVALID_ARG; }

  float4 *lo = nullptr, *hi = nullptr, *leaf_lo = nullptr, *leaf_hi = nullptr, *node_lo = nullptr, *node_hi = nullptr;
  Bounds6 *d_bounds = nullptr;
  uint64_t *keys = nullptr;
  uint32_t *vals = nullptr, *live = nullptr, *tile_sums = nullptr, *d_total = nullptr;
  int2 *children = nullptr, *range = nullptr;
  int *parent_node = nullptr, *parent_leaf = nullptr;
  unsigned int *flags = nullptr, *d_depth = nullptr;
  Bounds6 hb;
  const int T = 256;
  const uint32_t nb = (n + T - 1) / T;
  const int ni = (int)n - 1;
  BuildTimer btimer; (void)btimer;

  CK(dmalloc(&lo, sizeof(float4) * n, st));
  CK(dmalloc(&hi, sizeof(float4) * n, st));
  CK(dmalloc(&d_bounds, sizeof(Bounds6), st));
  CK(dmalloc(&keys, sizeof(uint64_t) * n, st));
  CK(dmalloc(&vals, sizeof(uint32_t) * n, st));
  CK(dmalloc(&out.prims, sizeof(PrimRec) * n, st));
  CK(dmalloc(&leaf_lo, sizeof(float4) * n, st));
  CK(dmalloc(&leaf_hi, sizeof(float4) * n, st));

  BLAP("alloc 1");
  init_bounds_kernel<<<1, 32, 0, st>>>(d_bounds);
  prim_bounds_kernel<<<nb, T, 0, st>>>(in.d_p1, in.d_p2, in.d_p3, in.n_tri, in.d_sph_center, in.d_sph_radius, in.n_sph, in.d_tri_obj,
                                       in.d_sph_obj, in.n_objects, lo, hi, d_bounds);
  CK(cudaMemcpyAsync(&hb, d_bounds, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (hb.bad_value || hb.bad_index) {
    err = hb.bad_value ? "non-finite vertex / sphere data" : "primitive object index out of range";
    rc = CUTRACE_ERR_INVALID_ARG;
    goto done;
  }
  {
    float3 clo, cinv;
    float cl[3], ch[3], mag = 0.f;
    for (int c = 0; c < 3; c++) {
      out.lo[c] = ord2f_host(hb.lo[c]); out.hi[c] = ord2f_host(hb.hi[c]);
      cl[c] = ord2f_host(hb.clo[c]); ch[c] = ord2f_host(hb.chi[c]);
      mag = fmaxf(mag, fmaxf(fabsf(out.lo[c]), fabsf(out.hi[c])));
    }
    clo = make_float3(cl[0], cl[1], cl[2]);
    // ONE scale for the three axes (cubic Morton cells).  Normalising each axis by its own extent makes the
    // cells of a flat scene (the 106x106 instance grid is 106 x 0.85 x 106) extremely anisotropic: the top of
    // the tree then splits the thin axis over and over and every ray has to descend both halves
    // (profiles/r01_v2_synthetic10m.md: 9,000 thread-instructions per primary ray before this change).
    {
      float ext = fmaxf(fmaxf(ch[0] - cl[0], ch[1] - cl[1]), ch[2] - cl[2]);
      float inv = ext > 0.f ? 1.f / ext : 0.f;
      cinv = make_float3(inv, inv, inv);
    }
    BLAP("bounds");
    morton_kernel<<<nb, T, 0, st>>>(lo, hi, n, clo, cinv, keys, vals);
    rc = radix_sort_pairs(keys, vals, n, st, err);
    if (rc) goto done;
    BLAP("morton + sort");
    if (n <= leaf_size) {  // the whole scene is one leaf
      gather_kernel<<<nb, T, 0, st>>>(vals, n, in.d_p1, in.d_p2, in.d_p3, in.d_tri_obj, in.n_tri, in.d_sph_center,
                                      in.d_sph_radius, in.d_sph_obj, lo, hi, out.prims, leaf_lo, leaf_hi);
      CK(cudaGetLastError());
      out.root = leaf_encode(0u, n);
      out.n_nodes = 0;
      out.depth = 0;
      CK(cudaStreamSynchronize(st));
      goto done;
    }

    // box inflation: the reference's Cramer test accepts rays a few ulp outside the exact triangle,
    // so boxes get a margin proportional to the scene's coordinate magnitude (DESIGN.md "conservative culling")
    float eps = fmaxf(mag, 1e-30f) * 2e-6f;

    CK(dmalloc(&children, sizeof(int2) * ni, st));
    CK(dmalloc(&range, sizeof(int2) * ni, st));
    CK(dmalloc(&parent_node, sizeof(int) * ni, st));
    CK(dmalloc(&parent_leaf, sizeof(int) * n, st));
    CK(dmalloc(&node_lo, sizeof(float4) * ni, st));
    CK(dmalloc(&node_hi, sizeof(float4) * ni, st));
    CK(dmalloc(&flags, sizeof(unsigned int) * ni, st));
    CK(dmalloc(&live, sizeof(uint32_t) * ni, st));
    CK(dmalloc(&tile_sums, sizeof(uint32_t) * ((ni + SCAN_TILE - 1) / SCAN_TILE + 1), st));
    CK(dmalloc(&d_total, sizeof(uint32_t), st));
    CK(dmalloc(&d_depth, sizeof(unsigned int), st));
    CK(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * ni, st));
    CK(cudaMemsetAsync(d_depth, 0, sizeof(unsigned int), st));
    BLAP("gather + alloc 2");
    const uint32_t nbi = (ni + T - 1) / T;
    karras_kernel<<<nbi, T, 0, st>>>(keys, (int)n, children, range, parent_node, parent_leaf);
    if (CTB_SAH_TREELETS && in.sah_treelets && !getenv("CUTRACE_DEBUG_NO_SAH")) {
      // rebuild the topology of every subtree of at most SAH_T primitives with sweep SAH (`live` is free until live_kernel: treelet roots)
      uint32_t n_roots = 0;
      CK(cudaMemsetAsync(d_total, 0, 4, st));
      sah_roots_kernel<<<nbi, T, 0, st>>>(ni, range, parent_node, leaf_size, live, d_total);
      CK(cudaMemcpyAsync(&n_roots, d_total, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (n_roots) {
        CK(cudaFuncSetAttribute(sah_treelet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SahShared)));
        sah_treelet_kernel<<<n_roots, SAH_T, sizeof(SahShared), st>>>(live, lo, hi, vals, children, range, parent_node, parent_leaf);
        CK(cudaGetLastError());
      }
      BLAP("sah treelets");
    }
    // primitives in their final order (the treelets permute `vals` inside their ranges)
    gather_kernel<<<nb, T, 0, st>>>(vals, n, in.d_p1, in.d_p2, in.d_p3, in.d_tri_obj, in.n_tri, in.d_sph_center,
                                    in.d_sph_radius, in.d_sph_obj, lo, hi, out.prims, leaf_lo, leaf_hi);
    CK(cudaGetLastError());
    refit_kernel<<<nb, T, 0, st>>>((int)n, children, parent_node, parent_leaf, leaf_lo, leaf_hi, node_lo, node_hi, flags);
    live_kernel<<<nbi, T, 0, st>>>(ni, range, leaf_size, live);
    exclusive_scan_u32(live, live, (uint32_t)ni, tile_sums, d_total, st);
    uint32_t n_live = 0;
    CK(cudaMemcpyAsync(&n_live, d_total, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    BLAP("karras refit scan");
    if (n_live == 0) { err = "internal: LBVH has no live node"; rc = CUTRACE_ERR_INTERNAL; goto done; }
    out.n_nodes = n_live;
    CK(dmalloc(&out.nodes, sizeof(Node) * n_live, st));
    emit_kernel<<<nbi, T, 0, st>>>(ni, children, range, live, leaf_size, leaf_lo, leaf_hi, node_lo, node_hi, eps, out.nodes);
    if (CTB_BVH4) {   // `live` (the compact indices of the binary nodes) is not needed any more: reuse it for the kept flags / indices
      kept_kernel<<<nbi, T, 0, st>>>(ni, range, parent_node, leaf_size, live);
      exclusive_scan_u32(live, vals, (uint32_t)ni, tile_sums, d_total, st);   // vals (sort payload) is free by now: 4-wide indices
      uint32_t n4 = 0;
      CK(cudaMemcpyAsync(&n4, d_total, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (n4 == 0) { err = "internal: 4-wide collapse kept no node"; rc = CUTRACE_ERR_INTERNAL; goto done; }
      out.n_nodes4 = n4;
      CK(dmalloc(&out.nodes4, sizeof(Node4) * n4, st));
      emit4_kernel<<<nbi, T, 0, st>>>(ni, children, range, live, vals, leaf_size, leaf_lo, leaf_hi, node_lo, node_hi, eps, out.nodes4);
    }
    depth_kernel<<<nbi, T, 0, st>>>(ni, range, parent_node, leaf_size, d_depth);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&out.depth, d_depth, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    BLAP("emit depth");
    out.root = 0;
    if (out.depth > CTB_STACK - 2) {
      char buf[128];
      snprintf(buf, sizeof buf, "LBVH depth %u exceeds the traversal stack (%d)", out.depth, CTB_STACK - 2);
      err = buf; rc = CUTRACE_ERR_INTERNAL; goto done;
    }
  }
done:
  dfree(lo, st); dfree(hi, st); dfree(d_bounds, st); dfree(keys, st); dfree(vals, st); dfree(leaf_lo, st); dfree(leaf_hi, st);
  dfree(node_lo, st); dfree(node_hi, st); dfree(children, st); dfree(range, st); dfree(parent_node, st); dfree(parent_leaf, st);
  dfree(flags, st); dfree(live, st); dfree(tile_sums, st); dfree(d_total, st); dfree(d_depth, st);
  if (rc != CUTRACE_OK) {
    dfree(out.prims, st); dfree(out.nodes, st); dfree(out.nodes4, st); out.nodes4 = nullptr;
    out.prims = nullptr; out.nodes = nullptr;
  }
  return rc;
}

}  // namespace ctb
