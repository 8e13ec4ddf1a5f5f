// output.cu — framebuffer plumbing after the render: un-tiling of (gathered) tile-major rank buffers
// into row-major images, and the reference's float -> 8-bit output mappings (inc/images.hpp:26-88)
// done on the device so that 9 B/px instead of 28 B/px cross PCIe when only the JPEG inputs are wanted.
#include "render.cuh"

namespace ctb {

__global__ void untile_kernel(TileMap tm, uint32_t world, const float *__restrict__ g_depth, const float *__restrict__ g_normal,
                              const float *__restrict__ g_color, const uint32_t *__restrict__ g_id, uint64_t stride_px, int only_rank,
                              float *__restrict__ depth, float *__restrict__ normal, float *__restrict__ color,
                              uint32_t *__restrict__ hit_id) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= tm.width || y >= tm.height) return;
  const uint32_t slot = slot_of_tile(tm, (y / CUTRACE_TILE) * tm.tiles_x + x / CUTRACE_TILE);
  uint32_t rank = slot % world, lt = slot / world;
  if (only_rank >= 0) {   // a single rank's buffer: foreign tiles keep whatever the destination holds
    if ((int)rank != only_rank) return;
    rank = 0;
  }
  uint64_t src = (uint64_t)rank * stride_px + ((uint64_t)lt << (2 * CUTRACE_TILE_SHIFT)) + ((y % CUTRACE_TILE) << CUTRACE_TILE_SHIFT) + (x % CUTRACE_TILE);
  uint64_t dst = (uint64_t)y * tm.width + x;
  if (depth) depth[dst] = g_depth[src];
  if (hit_id) hit_id[dst] = g_id[src];
  if (normal) { normal[3 * dst] = g_normal[3 * src]; normal[3 * dst + 1] = g_normal[3 * src + 1]; normal[3 * dst + 2] = g_normal[3 * src + 2]; }
  if (color) { color[3 * dst] = g_color[3 * src]; color[3 * dst + 1] = g_color[3 * src + 1]; color[3 * dst + 2] = g_color[3 * src + 2]; }
}

void launch_untile(const TileMap &tm, uint32_t world, const float *g_depth, const float *g_normal, const float *g_color,
                   const uint32_t *g_id, uint64_t stride_px, int only_rank, float *depth, float *normal, float *color,
                   uint32_t *hit_id, cudaStream_t st) {
  dim3 block(32, 8), grid((tm.width + 31) / 32, (tm.height + 7) / 8);
  untile_kernel<<<grid, block, 0, st>>>(tm, world ? world : 1u, g_depth, g_normal, g_color, g_id, stride_px, only_rank, depth, normal, color, hit_id);
}

// inc/images.hpp:27-29 (depth), :49-55 (normal), :72-76 (colour)
__global__ void encode_bytes_kernel(const float *__restrict__ depth, const float *__restrict__ normal, const float *__restrict__ color,
                                    float m, uint64_t n, uint8_t *__restrict__ depth_rgb, uint8_t *__restrict__ normal_rgb,
                                    uint8_t *__restrict__ color_rgb) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (depth && depth_rgb) {
    float v = depth[i];
    // host arithmetic in the reference: explicit _rn intrinsics keep nvcc from contracting into FMAs
    uint8_t b = isfinite(v) ? (uint8_t)__fdiv_rn(__fmul_rn(255.0f, __fsub_rn(m, v)), m) : 0;
    depth_rgb[3 * i] = b; depth_rgb[3 * i + 1] = b; depth_rgb[3 * i + 2] = b;
  }
  if (normal && normal_rgb) {
    float x = normal[3 * i], y = normal[3 * i + 1], z = normal[3 * i + 2];
    float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    if ((double)nrm <= 1e-6) {
      normal_rgb[3 * i] = 0; normal_rgb[3 * i + 1] = 0; normal_rgb[3 * i + 2] = 0;
    } else {
      float inv = __fdiv_rn(1.0f, nrm);
      normal_rgb[3 * i] = (uint8_t)__fmul_rn(255.0f, __fadd_rn(0.5f, __fmul_rn(0.5f, __fmul_rn(inv, x))));
      normal_rgb[3 * i + 1] = (uint8_t)__fmul_rn(255.0f, __fadd_rn(0.5f, __fmul_rn(0.5f, __fmul_rn(inv, y))));
      normal_rgb[3 * i + 2] = (uint8_t)__fmul_rn(255.0f, __fadd_rn(0.5f, __fmul_rn(0.5f, __fmul_rn(inv, z))));
    }
  }
  if (color && color_rgb) {
    for (int c = 0; c < 3; c++) {
      float v = color[3 * i + c];
      float cl = v > 0.0f ? v : 0.0f;
      cl = cl < 1.0f ? cl : 1.0f;
      color_rgb[3 * i + c] = (uint8_t)(255 * cl);
    }
  }
}

void launch_encode_bytes(const float *depth, const float *normal, const float *color, float max_depth, uint64_t n_px,
                         uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb, cudaStream_t st) {
  if (!n_px) return;
  encode_bytes_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(depth, normal, color, max_depth, n_px, depth_rgb, normal_rgb, color_rgb);
}

__global__ void fill_sentinels_kernel(float *depth, float *normal, float *color, uint32_t *hit_id, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (depth) depth[i] = INFINITY;
  if (hit_id) hit_id[i] = CUTRACE_NO_HIT;
  if (normal) { normal[3 * i] = 0.f; normal[3 * i + 1] = 0.f; normal[3 * i + 2] = 0.f; }
  if (color) { color[3 * i] = 0.f; color[3 * i + 1] = 0.f; color[3 * i + 2] = 0.f; }
}

void launch_fill_sentinels(float *depth, float *normal, float *color, uint32_t *hit_id, uint64_t n_px, cudaStream_t st) {
  if (!n_px) return;
  fill_sentinels_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, st>>>(depth, normal, color, hit_id, n_px);
}

}  // namespace ctb
