// render.cuh — launch interface of the wavefront render kernels (render.cu) and output kernels (output.cu).
#ifndef CUTRACE_B200_RENDER_CUH
#define CUTRACE_B200_RENDER_CUH
#include "common.cuh"

namespace ctb {

#ifndef CTB_THREADS
#define CTB_THREADS 512
#endif
constexpr int TRACE_THREADS = CTB_THREADS;
constexpr int WORK_CHUNK_MAX = 512;   // most rays a warp claims per work-stealing atomic (guided: shrinks to 32 at the tail)
constexpr int SLOT_BLOCK = 256;   // queue slots a warp reserves per atomicAdd (>= 64: one warp-iteration emits <= 32 + 32 rays)
#ifndef CTB_SLOT_DIV
#define CTB_SLOT_DIV 8
#endif
#ifndef CTB_SLOT_MIN
#define CTB_SLOT_MIN 64
#endif
#define CTB_HOLE 0xffffffffu       // pix of a retired (unused) queue slot
#ifndef CTB_EXPORT_CTAS
#define CTB_EXPORT_CTAS 74   // CTAs of the G-buffer export kernel (half a CTA per SM: it must not crowd out the render kernels)
#endif
#ifndef CTB_SHADE_MIN_BLOCKS
#define CTB_SHADE_MIN_BLOCKS 2   // measured on B200: 2 CTAs/SM (64 regs) beat 1 CTA/SM for K = 1 (profiles/r01_tuning.md)
#endif
#ifndef CTB_MIN_BLOCKS
#define CTB_MIN_BLOCKS 2   // resident CTAs per SM the register allocator must allow (tuned on B200)
#endif

// Where a frame's results go.  row_major = 0: this ctx's tile-major local buffers (index = local pixel index; the
// NCCL-gather path un-tiles them later).  row_major = 1: a row-major full frame (index = y*width + x) — the ctx's own
// frame on one GPU, or rank 0's frame mapped through CUDA IPC on the other ranks, in which case the G-buffer and the
// final colour are stored straight into the peer GPU's HBM over NVLink by the kernels that produce them.
struct FrameTargets {
  float *depth;         // n px
  float *normal;        // n px * 3
  float *color;         // n px * 3
  uint32_t *hit_id;     // n px
  uint32_t row_major;
};

struct LaunchCfg {
  int mode;             // 0 global, 1 BVH staged in shared memory
  size_t smem_bytes;
  int grid_trace, grid_shade;   // persistent grid sizes (multiples of the SM count)
  int grid_frame;               // co-resident CTAs of the cooperative frame kernel (0: not available on this device)
  int grid_pixel;               // persistent grid of the pixel kernel
  int pixel_refill;             // pixel kernel: a lane takes its next pixel as soon as its path has ended (render.cu)
  int pixel_seg;                // pixel kernel: one work cursor per CTA instead of one for the grid (render.cu: claim_segment)
};
// pixel kernel: ONE CTA of 1024 threads per SM (64 registers).  Measured on bunny.json 4K (profiles/r02_tuning.md): 256 x 4 10.3 ms
// (71 KB of staged scene per CTA caps the SM at 3 CTAs), 256 x 3 9.44, 256 x 2 (128 registers) 10.7, 512 x 2 9.01; and on the
// final build 512 x 2 7.83, 768 x 1 (80 registers) 8.03, 640 x 1 (96 registers) 8.58, 1024 x 1 7.70 — the scene is staged once
// per SM, and more registers per thread do not pay for the warps they cost.
#ifndef CTB_PIXEL_THREADS
#define CTB_PIXEL_THREADS 1024
#endif
#ifndef CTB_PIXEL_MIN_BLOCKS
#define CTB_PIXEL_MIN_BLOCKS 1
#endif

#ifndef CTB_FILL_CHUNK_MAX
#define CTB_FILL_CHUNK_MAX 128   // shade records a warp claims at a time while it is only filling the tail of a trace phase
#endif

// Everything the persistent frame kernel needs (one kernel = all bounce levels of one pixel batch), passed by value.
struct FrameArgs {
  SceneView sv;
  TileMap tm;
  uint32_t bounces, levels;      // levels = bounces + 1 when some material spawns secondary rays, else 1
  uint32_t first_level;          // 0, or 1 when trace(0) already ran as its own kernel (cutrace_render_download: the G-buffer
                                 // copy to the host starts behind that kernel)
  uint32_t px_base, n_px;        // local pixel range of this batch
  RayRec *rays[2];               // ping-pong ray queues
  uint32_t ray_cap;
  ShadeRec *shade[16];           // one shade queue per level
  uint32_t shade_cap[16];
  FrameCounters *ctr;            // device counters: zero on entry, published to host_ctr and cleared again on exit
  FrameStats *host_stats;        // mapped pinned host memory (device pointer), may be NULL
  FrameTargets gbuf;             // where trace(0) stores the G-buffer
  FrameTargets out;              // where the finished frame lives (own HBM / peer GPU / pinned host memory)
  FrameTargets gsrc;             // != NULL: local tile-major G-buffer that has to travel to `out`
  FrameTargets acc;              // branching scenes: local colour accumulator (float atomics)
  float *level_color;            // non-branching scenes: levels x level_stride floats of per-level partial images
  uint64_t level_stride;
  uint32_t *nlev;
  const float *local_color;      // branching scenes: what the frame assembly copies
  uint32_t combine_levels;       // levels of the ordered sum; 0 = copy local_color
  int atomic_accumulate;
  int combine;                   // run the frame assembly inside the kernel
  int export_with_color;         // 1: the G-buffer travels with the colour at the end, 0: as filler work under the bounce levels
};

// the per-pixel kernel (render.cu: pixel_kernel)
struct PixelArgs {
  SceneView sv;
  TileMap tm;
  uint32_t bounces;
  uint32_t px_base, n_px;
  FrameCounters *ctr;        // zero on entry; the statistics are published to host_stats and cleared again on exit
  FrameStats *host_stats;    // mapped pinned host memory (device pointer), may be NULL
  FrameTargets out;          // where the frame lives (any layout FrameTargets describes)
  FrameTargets out2;         // optional second, row-major copy (any pointer may be NULL): the caller's pinned host images of
                             // cutrace_render_download — the pixels travel over PCIe while the kernel runs, there is no copy afterwards
};


cudaError_t launch_pixel(const LaunchCfg &cfg, const PixelArgs &args, cudaStream_t st);

// fills cfg for a scene on the current device; smem budget from the device attributes
cudaError_t plan_launch(const SceneView &sv, bool allow_smem, LaunchCfg *cfg);

// One bounce level of the wavefront. Level 0 generates primary rays for local pixel indices
// [px_base, px_base + n_px) itself; deeper levels read rays_in (count in ctr->n_rays[level]).
void launch_trace(const LaunchCfg &cfg, const SceneView &sv, const TileMap &tm, uint32_t level, uint32_t bounces,
                  uint32_t px_base, uint32_t n_px, const RayRec *rays_in, RayRec *rays_out, uint32_t ray_cap, ShadeRec *shade_out,
                  uint32_t shade_cap, FrameCounters *ctr, const FrameTargets &fb, uint32_t *nlev, uint32_t work_bound, cudaStream_t st);
// shadow rays + Phong for the shade records of one level; accumulates into fb.color
void launch_shade(const LaunchCfg &cfg, const SceneView &sv, uint32_t level, const ShadeRec *shade, uint32_t shade_cap, FrameCounters *ctr,
                  const FrameTargets &fb, bool atomic_accumulate, float *level_color, uint32_t px_base, uint32_t work_bound,
                  cudaStream_t st);
// all levels of one pixel batch in ONE cooperative persistent kernel (device-side level loop, see render.cu)
cudaError_t launch_frame(const LaunchCfg &cfg, const FrameArgs &args, uint32_t work_bound, cudaStream_t st);
// colour[px] = sum over levels l < nlev[px] of level_color[l][px - px_base], in level order; levels == 0: copy the
// locally accumulated colour (branching scenes) instead.  Writes into `out` in its layout; gsrc.depth != NULL: also
// forwards the tile-major G-buffer gsrc to `out` (peer frame).
void launch_combine(const TileMap &tm, const uint32_t *nlev, const float *level_color, uint64_t level_stride, uint32_t levels,
                    const float *local_color, uint32_t px_base, uint32_t n_px, const FrameTargets &out, const FrameTargets &gsrc,
                    cudaStream_t st);

void launch_export_gbuffer(const TileMap &tm, uint32_t px_base, uint32_t n_px, const FrameTargets &src, const FrameTargets &out,
                           cudaStream_t st);

// output.cu
void launch_untile(const TileMap &tm, uint32_t world, const float *g_depth, const float *g_normal, const float *g_color,
                   const uint32_t *g_id, uint64_t stride_px, int only_rank, float *depth, float *normal, float *color,
                   uint32_t *hit_id, cudaStream_t st);
void launch_encode_bytes(const float *depth, const float *normal, const float *color, float max_depth, uint64_t n_px,
                         uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb, cudaStream_t st);
void launch_fill_sentinels(float *depth, float *normal, float *color, uint32_t *hit_id, uint64_t n_px, cudaStream_t st);

}  // namespace ctb
#endif
