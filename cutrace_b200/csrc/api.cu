// api.cu — the C-ABI (include/cutrace.h): scene upload + LBVH build, frame render, download.
//
//   cutrace_upload_scene  replaces default_to_gpu / cpu_to_gpu::convert (inc/cpu_to_gpu.hpp:188-198):
//                         instead of cudaMallocManaged arrays of tagged unions with a nested
//                         allocation per mesh (inc/default_schema.hpp:592-594) the scene becomes
//                         flat 16-byte aligned records in device memory plus an LBVH.
//   cutrace_render        replaces gpu::render's launch + sync (inc/kernel.hpp:103-108).
//   cutrace_download      replaces its 3*h row cudaMemcpy calls and host max-depth scan
//                         (inc/kernel.hpp:110-125) with one device un-tile + one copy per image.
// There is NO CPU fallback anywhere in this library: without a CUDA device every entry point
// fails with CUTRACE_ERR_NO_DEVICE.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <type_traits>
#include "alloc.cuh"
#include "bvh.cuh"
#include "render.cuh"

using namespace ctb;

static thread_local std::string g_err = "";

#ifdef CTB_TIMING   // developer instrumentation (tools/build_variant.sh timing -DCTB_TIMING); never on in the product build
struct PhaseTimer {
  std::chrono::high_resolution_clock::time_point t = std::chrono::high_resolution_clock::now();
  void lap(const char *what) {
    auto n = std::chrono::high_resolution_clock::now();
    fprintf(stderr, "  [ctb-timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
    t = n;
  }
};
#define LAP(x) ptimer.lap(x)
#else
struct PhaseTimer {};
#define LAP(x)
#endif

static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      return fail(e_ == cudaErrorMemoryAllocation ? CUTRACE_ERR_OUT_OF_MEMORY                          \
                  : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? CUTRACE_ERR_NO_DEVICE \
                                                                                  : CUTRACE_ERR_CUDA, \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                                \
    }                                                                                                 \
  } while (0)

// cudaHostAlloc / cudaFreeHost cost 1-3 ms each (measured, tools/e2e_probe.py); the small pinned block every ctx
// needs for its frame counters is therefore recycled through a process-wide free list.
#include <mutex>
static std::mutex g_pinned_mu;
static std::vector<FrameStats *> g_pinned_free;
static FrameStats *pinned_counters_get() {
  {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (!g_pinned_free.empty()) { FrameStats *p = g_pinned_free.back(); g_pinned_free.pop_back(); return p; }
  }
  FrameStats *p = nullptr;
  if (cudaHostAlloc(&p, sizeof(FrameStats), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) return nullptr;
  return p;
}
static void pinned_counters_put(FrameStats *p) {
  if (!p) return;
  std::lock_guard<std::mutex> lk(g_pinned_mu);
  g_pinned_free.push_back(p);
}

// Streams and events of a ctx are recycled the same way: creating five streams and 73 events and destroying them again
// costs ~0.5 ms per ctx (measured with -DCTB_TIMING: 0.25 ms + 0.28 ms), 4 % of an upload-render-download-free cycle
// of bunny.json at 4K.  A set goes back to the free list of its device in cutrace_free, after all its streams are idle.
struct StreamSet {
  int device = -1;
  cudaStream_t main = nullptr;         // high priority: the trace chain (unused when the caller brings a stream)
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // low priority: shade kernels, G-buffer export
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_gbuf = nullptr;
  std::vector<cudaEvent_t> events;
};
#define CTB_N_EVENTS 72
static std::mutex g_streams_mu;
static std::vector<StreamSet *> g_streams_free;
static void stream_set_destroy(StreamSet *ss) {
  if (!ss) return;
  for (cudaEvent_t e : ss->events) cudaEventDestroy(e);
  for (int i = 0; i < 3; i++) if (ss->aux[i]) cudaStreamDestroy(ss->aux[i]);
  if (ss->copy) cudaStreamDestroy(ss->copy);
  if (ss->ev_gbuf) cudaEventDestroy(ss->ev_gbuf);
  if (ss->main) cudaStreamDestroy(ss->main);
  delete ss;
}
// the current device must be `device`
static cudaError_t stream_set_get(int device, StreamSet **out) {
  {
    std::lock_guard<std::mutex> lk(g_streams_mu);
    for (size_t i = 0; i < g_streams_free.size(); i++)
      if (g_streams_free[i]->device == device) {
        *out = g_streams_free[i];
        g_streams_free.erase(g_streams_free.begin() + (long)i);
        return cudaSuccess;
      }
  }
  StreamSet *ss = new StreamSet;
  ss->device = device;
  cudaError_t e;
  int pr_least = 0, pr_greatest = 0;
#define SS(call) do { e = (call); if (e != cudaSuccess) { stream_set_destroy(ss); return e; } } while (0)
  SS(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
  SS(cudaStreamCreateWithPriority(&ss->main, cudaStreamNonBlocking, pr_greatest));
  // shade kernels run on two lower-priority streams so that the trace chain (the critical path) gets SMs first
  for (int i = 0; i < 3; i++) SS(cudaStreamCreateWithPriority(&ss->aux[i], cudaStreamNonBlocking, pr_least));
  SS(cudaStreamCreateWithFlags(&ss->copy, cudaStreamNonBlocking));
  SS(cudaEventCreateWithFlags(&ss->ev_gbuf, cudaEventDisableTiming));
  for (int i = 0; i < CTB_N_EVENTS; i++) { cudaEvent_t ev; SS(cudaEventCreate(&ev)); ss->events.push_back(ev); }
#undef SS
  *out = ss;
  return cudaSuccess;
}
static void stream_set_put(StreamSet *ss) {
  if (!ss) return;
  cudaStreamSynchronize(ss->main);
  for (int i = 0; i < 3; i++) cudaStreamSynchronize(ss->aux[i]);
  cudaStreamSynchronize(ss->copy);
  {
    std::lock_guard<std::mutex> lk(g_streams_mu);
    if (g_streams_free.size() < 32) { g_streams_free.push_back(ss); return; }
  }
  stream_set_destroy(ss);
}

struct cutrace_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  StreamSet *streams = nullptr;
  cutrace_opts opts{};
  // scene
  BvhResult bvh;
  PlaneRec *planes = nullptr;
  MaterialRec *materials = nullptr;
  LightRec *lights = nullptr;
  uint32_t *obj_material = nullptr;
  ObjBound *obj_bounds = nullptr;    // per-object AABB + mesh flag (the reference's mesh pre-test, trace.cuh: mesh_gate)
  float *pl_tbl = nullptr, *pl_eps = nullptr;   // light-side table of the planes (trace.cuh: plane_side_prepass)
  SceneView sv{};
  int max_children = 0;
  // frame
  TileMap tm{};
  uint64_t n_local_px = 0;   // padded: n_local_tiles * CUTRACE_TILE_PIXELS
  uint64_t local_pixels = 0; // pixels of this ctx that are inside the image
  FrameTargets fb{};                 // tile-major local buffers (sharded ctx without a peer frame: NCCL-gather path)
  float *frame = nullptr;            // own row-major full frame: depth n | normal 3n | colour 3n | id n (one block)
  bool frame_is_ipc = false;         // allocated with cudaMalloc and exported through CUDA IPC
  void *peer_frame = nullptr;        // another ctx's frame (CUDA IPC import or same-process attach): stored into over NVLink
  bool peer_is_ipc = false;
  float *local_color = nullptr;      // tile-major colour accumulator (branching scenes: float atomics)
  RayRec *rays[2] = {nullptr, nullptr};
  ShadeRec *shade[16] = {};          // one shade queue per bounce level (shade(L) overlaps trace(L+1..))
  uint64_t shade_cap[16] = {};
  float *level_color = nullptr;      // levels x batch_px x 3: per-level partial images (non-branching scenes)
  uint32_t *nlev = nullptr;          // n_local_px: number of levels that contributed to a pixel
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // [0],[1]: shade kernels; [2]: G-buffer export to a peer frame
  cudaStream_t copy_stream = nullptr;   // D2H of the G-buffer underneath the bounce levels (cutrace_render_download)
  cudaEvent_t ev_gbuf = nullptr;        // "primary rays done": recorded inside the frame (external event when captured)
  bool want_gbuf_event = false;
  uint64_t batch_px = 0, cap = 0;
  uint32_t factor = 1;
  FrameCounters *d_ctr = nullptr;
  FrameStats *h_ctr = nullptr;      // pinned + mapped: the frame kernel publishes its statistics here
  FrameStats *h_ctr_dev = nullptr;  // device address of h_ctr
  float phase_ms[18] = {};          // frame kernel: when trace(p) was complete / the frame ended, relative to its start (last batch)
  uint32_t phase_count = 0;
  bool queues_ready = false;        // rays / shade queues, level images (alloc_queues)
  bool ctr_dirty = true;            // d_ctr may hold values of an earlier (multi-launch or failed) frame: clear before a frame kernel
  bool frame_kernel_failed = false; // a cooperative launch was refused: this ctx stays on the multi-launch path
  bool env_export_with_color = false;
  LaunchCfg cfg{};
  // download staging (row-major full frame), lazily allocated
  float *st_depth = nullptr;         // staging frame block (same layout as `frame`)
  uint64_t st_px = 0;
  uint8_t *st_bytes = nullptr;   // 3 images x n x 3 bytes
  uint64_t st_bytes_px = 0;
  uint32_t graph_launches = 0;
  bool env_no_graph = false, env_skip_export = false, env_local_color = false;   // developer toggles, read once at upload
  bool env_graph_first = false;
  uint32_t frames_rendered = 0;      // frames since the scene / camera / frame binding last changed
  cudaGraphExec_t graph = nullptr;   // the whole frame (all streams) captured once, replayed per cutrace_render
  bool graph_failed = false;
  cutrace_stats stats{};
  bool rendered = false;
  // host destinations of an in-flight cutrace_render_download (NULL otherwise)
  float *dl_depth = nullptr, *dl_normal = nullptr;
  uint32_t *dl_id = nullptr;
  FrameTargets dl_direct{};          // ... when they are pinned + mapped: device addresses the pixel kernel stores to directly
  bool dl_direct_on = false;
  std::vector<cudaEvent_t> events;
};

namespace {

struct DeviceGuard {
  int prev = -1;
  bool ok = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// the captured frame bakes in camera, resolution, queue and frame pointers: drop it whenever one of them changes
void drop_graph(cutrace_ctx *c) {
  if (c->graph) { cudaGraphExecDestroy(c->graph); c->graph = nullptr; }
  c->graph_failed = false;
  c->frames_rendered = 0;
}

// multiplier of the tile permutation (see TileMap): ~0.618 n, odd, coprime to n; and its inverse mod n
void tile_permutation(uint32_t n, uint32_t world, uint32_t *a_out, uint32_t *ainv_out) {
  *a_out = 1; *ainv_out = 1;
  if (world <= 1 || n < 3 || getenv("CUTRACE_DEBUG_NO_TILE_PERM")) return;
  auto gcd = [](uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; };
  uint64_t a = ((uint64_t)n * 618ull / 1000ull) | 1ull;
  while (a < n && gcd(a, n) != 1) a += 2;
  if (a >= n) return;
  // inverse by extended Euclid
  long long t = 0, nt = 1, r = n, nr = (long long)a;
  while (nr) { long long q = r / nr, tmp = t - q * nt; t = nt; nt = tmp; tmp = r - q * nr; r = nr; nr = tmp; }
  if (t < 0) t += n;
  *a_out = (uint32_t)a; *ainv_out = (uint32_t)t;
}

void free_frame(cutrace_ctx *c) {
  drop_graph(c);
  cudaStream_t st = c->stream;
  dfree(c->fb.depth, st); dfree(c->fb.normal, st); dfree(c->fb.color, st); dfree(c->fb.hit_id, st);
  if (c->frame_is_ipc) { if (st) cudaStreamSynchronize(st); cudaFree(c->frame); } else dfree(c->frame, st);
  if (c->peer_frame && c->peer_is_ipc) { if (st) cudaStreamSynchronize(st); cudaIpcCloseMemHandle(c->peer_frame); }
  dfree(c->local_color, st);
  c->frame = nullptr; c->frame_is_ipc = false; c->peer_frame = nullptr; c->local_color = nullptr;
  dfree(c->rays[0], st); dfree(c->rays[1], st);
  for (int i = 0; i < 16; i++) { dfree(c->shade[i], st); c->shade[i] = nullptr; c->shade_cap[i] = 0; }
  dfree(c->level_color, st); dfree(c->nlev, st);
  c->level_color = nullptr; c->nlev = nullptr;
  dfree(c->st_depth, st);
  dfree(c->st_bytes, st); c->st_bytes = nullptr; c->st_bytes_px = 0;
  c->fb = FrameTargets{};
  c->rays[0] = c->rays[1] = nullptr;
  c->queues_ready = false;
  c->st_depth = nullptr;
  c->st_px = 0;
  c->rendered = false;
}

int alloc_frame(cutrace_ctx *c, uint32_t width, uint32_t height) {
  if (width == 0 || height == 0) return fail(CUTRACE_ERR_INVALID_ARG, "camera width/height must be > 0");
  if ((uint64_t)width * height >= (1ull << 31)) return fail(CUTRACE_ERR_INVALID_ARG, "frame too large (>= 2^31 pixels)");
  free_frame(c);
  TileMap tm{};
  tm.width = width; tm.height = height;
  tm.tiles_x = (width + CUTRACE_TILE - 1) / CUTRACE_TILE;
  tm.tiles_y = (height + CUTRACE_TILE - 1) / CUTRACE_TILE;
  tm.world = c->opts.tile_world > 1 ? c->opts.tile_world : 1;
  tm.rank = tm.world > 1 ? c->opts.tile_rank : 0;
  uint32_t total_tiles = tm.tiles_x * tm.tiles_y;
  // every rank gets the same padded tile count so that gathered rank buffers have one stride
  tm.n_local_tiles = (total_tiles + tm.world - 1) / tm.world;
  tm.n_tiles = total_tiles;
  tile_permutation(total_tiles, tm.world, &tm.perm_a, &tm.perm_ainv);
  tm.curve = 1u;   // super-tile order (common.cuh: TileMap); CUTRACE_TILE_CURVE=0: round 1's multiplicative scatter (tuning comparisons)
  if (const char *e = getenv("CUTRACE_TILE_CURVE")) tm.curve = atoi(e) != 0 ? 1u : 0u;
  tm.wide_warps = 0u;   // set by cutrace_frame_attach for a frame in host memory   // a sharded ctx usually stores into a remote frame: 16 x 2 warps give 64 / 192-byte row segments
  c->tm = tm;
  c->n_local_px = (uint64_t)tm.n_local_tiles * CUTRACE_TILE_PIXELS;
  if (tm.world == 1) {
    c->local_pixels = (uint64_t)width * height;   // (the loop below costs 0.1 ms of every upload at 4K)
  } else {
    uint64_t px = 0;
    for (uint32_t lt = 0; lt < tm.n_local_tiles; lt++) {
      uint32_t tx, ty;
      if (!tile_of_slot(tm, lt * tm.world + tm.rank, tx, ty)) continue;
      uint32_t w = std::min(CUTRACE_TILE, tm.width - tx * CUTRACE_TILE), h = std::min(CUTRACE_TILE, tm.height - ty * CUTRACE_TILE);
      px += (uint64_t)w * h;
    }
    c->local_pixels = px;
  }
  c->sv.cam.w = width; c->sv.cam.h = height;

  uint32_t b = c->opts.bounces;
  const bool branching = c->max_children >= 2 && b > 0;
  const uint32_t levels = c->max_children > 0 ? b + 1 : 1;
  c->factor = branching ? (1u << b) : 1u;
  uint64_t fb_bytes = c->n_local_px * 36ull;
  // queue budget: 60 % of the free HBM.  cudaMemGetInfo costs 30-70 ms once gigabytes sit in the memory pool
  // (tools/build_probe.py), so it is only asked when the worst-case queues of the whole frame exceed 8 GB.
  double budget = 8.0 * 1073741824.0;
  if (const char *e = getenv("CUTRACE_QUEUE_BUDGET_MB")) budget = atof(e) * 1048576.0;
  else {
    double shade_all = 0;
    for (uint32_t L = 0; L < levels; L++) shade_all += (branching ? (double)(1u << L) : 1.0) * sizeof(ShadeRec);
    const double need_all = (double)c->n_local_px * ((c->max_children > 0 && b > 0 ? (double)c->factor * 2.0 * sizeof(RayRec) : 0.0) + shade_all +
                                                     (branching ? 0.0 : 12.0 * levels));
    if (need_all > budget) {
      size_t free_b = 0, total_b = 0;
      CU(cudaMemGetInfo(&free_b, &total_b));
      budget = (double)free_b * 0.6 - (double)fb_bytes;
    }
  }
  // bytes of queue memory per batch pixel: two ping-pong ray queues of the worst-case level, one shade queue per
  // level (level L holds at most 2^L hits per pixel when a material both reflects and transmits, else 1), and the
  // per-level partial colour images of the non-branching path
  double shade_per_px = 0;
  for (uint32_t L = 0; L < levels; L++) shade_per_px += (branching ? (double)(1u << L) : 1.0) * sizeof(ShadeRec);
  double per_px = (c->max_children > 0 && b > 0 ? (double)c->factor * 2.0 * sizeof(RayRec) : 0.0) + shade_per_px +
                  (branching ? 0.0 : 12.0 * levels);
  uint64_t batch = budget > per_px * 1024.0 ? (uint64_t)(budget / per_px) : 1024ull;
  batch = (batch / 1024ull) * 1024ull;
  if (batch < 1024) batch = 1024;
  if (batch > c->n_local_px) batch = c->n_local_px;
  while (batch * c->factor >= (1ull << 31) - (1ull << 25) && batch > 1024) batch = ((batch / 2) / 1024ull) * 1024ull;
  c->batch_px = batch;
  // every producer warp may leave one partly used slot block per queue behind
  // (reserve_slots in render.cu: the tail of a warp's last block of a level is the only thing it ever leaves unused)
  const uint64_t slack = (uint64_t)std::max(c->cfg.grid_trace, c->cfg.grid_frame) * (TRACE_THREADS / 32) * SLOT_BLOCK;
  c->cap = batch * c->factor + slack;

  cudaStream_t st = c->stream;
  if (c->tm.world > 1) {   // sharded: tile-major local buffers for the gather path (unused once a peer frame is imported)
    CU(dmalloc(&c->fb.depth, sizeof(float) * c->n_local_px, st));
    CU(dmalloc(&c->fb.normal, sizeof(float) * 3 * c->n_local_px, st));
    CU(dmalloc(&c->fb.color, sizeof(float) * 3 * c->n_local_px, st));
    CU(dmalloc(&c->fb.hit_id, sizeof(uint32_t) * c->n_local_px, st));
  } else {                 // one GPU: results are written row-major, no un-tile pass
    CU(dmalloc(&c->frame, 32ull * width * height, st));
  }
  for (uint32_t L = 0; L < levels; L++) c->shade_cap[L] = batch * (branching ? (1ull << L) : 1ull) + slack;
  c->queues_ready = false;   // the wavefront's queues are allocated by the first frame that needs them (the pixel kernel has none)
  return CUTRACE_OK;
}

// Scheduler of a ctx (see CUTRACE_FLAG_PIXEL_KERNEL in cutrace.h): the persistent per-pixel kernel unless the caller forces one of
// the wavefront schedulers; CUTRACE_FLAG_SERIALIZE (per-kernel timings) implies one launch per level and kind.
bool wants_pixel_kernel(const cutrace_ctx *c) {
  const uint32_t forced = c->opts.flags & (CUTRACE_FLAG_FRAME_KERNEL | CUTRACE_FLAG_LAUNCHES | CUTRACE_FLAG_PIXEL_KERNEL);
  if ((c->opts.flags & CUTRACE_FLAG_SERIALIZE) || !c->h_ctr_dev) return false;
  return forced ? (forced & CUTRACE_FLAG_PIXEL_KERNEL) != 0 : true;
}

// queue memory of the wavefront schedulers, sized by alloc_frame
int alloc_queues(cutrace_ctx *c) {
  if (c->queues_ready) return CUTRACE_OK;
  cudaStream_t st = c->stream;
  const uint32_t b = c->opts.bounces;
  const bool branching = c->max_children >= 2 && b > 0;
  const uint32_t levels = c->max_children > 0 ? b + 1 : 1;
  if (branching) CU(dmalloc(&c->local_color, sizeof(float) * 3 * c->n_local_px, st));
  if (c->max_children > 0 && b > 0) {
    CU(dmalloc(&c->rays[0], sizeof(RayRec) * c->cap, st));
    CU(dmalloc(&c->rays[1], sizeof(RayRec) * c->cap, st));
  }
  for (uint32_t L = 0; L < levels; L++) CU(dmalloc(&c->shade[L], sizeof(ShadeRec) * c->shade_cap[L], st));
  if (!branching) {
    CU(dmalloc(&c->level_color, sizeof(float) * 3 * c->batch_px * levels, st));
    CU(dmalloc(&c->nlev, sizeof(uint32_t) * c->n_local_px, st));
  }
  c->queues_ready = true;
  return CUTRACE_OK;
}

template <typename T>
int upload(T **dst, const std::vector<T> &src, cudaStream_t st, bool sync = true) {
  *dst = nullptr;
  if (src.empty()) return CUTRACE_OK;
  CU(dmalloc(dst, sizeof(T) * src.size(), st));
  CU(cudaMemcpyAsync(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, st));
  if (sync) CU(cudaStreamSynchronize(st));   // src may be a temporary host vector; sync = false: the caller keeps it alive until it has synchronised the stream
  return CUTRACE_OK;
}

bool all_finite(const float *p, uint64_t n) {
  for (uint64_t i = 0; i < n; i++)
    if (!std::isfinite(p[i])) return false;
  return true;
}

int validate_desc(const cutrace_scene_desc *s) {
  if (!s) return fail(CUTRACE_ERR_INVALID_ARG, "scene is NULL");
  if (s->abi_version != CUTRACE_ABI_VERSION) return fail(CUTRACE_ERR_INVALID_ARG, "cutrace_scene_desc.abi_version mismatch");
  if (s->width == 0 || s->height == 0) return fail(CUTRACE_ERR_INVALID_ARG, "camera width/height must be > 0");
  if (s->n_triangles && !(s->tri_p1 && s->tri_p2 && s->tri_p3 && s->tri_object)) return fail(CUTRACE_ERR_INVALID_ARG, "triangle arrays are NULL");
  if (s->n_spheres && !(s->sph_center && s->sph_radius && s->sph_object)) return fail(CUTRACE_ERR_INVALID_ARG, "sphere arrays are NULL");
  if (s->n_planes && !(s->pl_point && s->pl_normal && s->pl_object)) return fail(CUTRACE_ERR_INVALID_ARG, "plane arrays are NULL");
  if (s->n_objects && !s->obj_material) return fail(CUTRACE_ERR_INVALID_ARG, "obj_material is NULL");
  if (s->n_materials && !(s->mat_color && s->mat_specular && s->mat_reflect && s->mat_phong && s->mat_transparency))
    return fail(CUTRACE_ERR_INVALID_ARG, "material arrays are NULL");
  if (s->n_lights && !(s->light_kind && s->light_vec && s->light_color)) return fail(CUTRACE_ERR_INVALID_ARG, "light arrays are NULL");
  if (s->n_triangles + s->n_spheres >= (1ull << 28)) return fail(CUTRACE_ERR_INVALID_ARG, "too many primitives (max 2^28-1)");
  for (uint32_t i = 0; i < s->n_objects; i++)
    if (s->obj_material[i] >= s->n_materials) return fail(CUTRACE_ERR_INVALID_ARG, "object material index out of range");
  // triangle / sphere coordinates and their object indices are validated on the device (prim_bounds_kernel)
  for (uint64_t i = 0; i < s->n_planes; i++)
    if (s->pl_object[i] >= s->n_objects) return fail(CUTRACE_ERR_INVALID_ARG, "pl_object index out of range");
  for (uint32_t i = 0; i < s->n_lights; i++)
    if (s->light_kind[i] > CUTRACE_LIGHT_POINT) return fail(CUTRACE_ERR_INVALID_ARG, "unknown light kind");
  if (!all_finite(s->pl_point, 3 * s->n_planes) || !all_finite(s->pl_normal, 3 * s->n_planes))
    return fail(CUTRACE_ERR_INVALID_ARG, "non-finite plane data");
  return CUTRACE_OK;
}

// row-major frame block -> the four images
FrameTargets frame_views(float *block, uint64_t n) {
  FrameTargets t{};
  t.depth = block; t.normal = block + n; t.color = block + 4 * n; t.hit_id = reinterpret_cast<uint32_t *>(block + 7 * n);
  t.row_major = 1;
  return t;
}

// where this ctx's kernels store their results (see FrameTargets)
FrameTargets frame_targets(const cutrace_ctx *c) {
  const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
  if (c->peer_frame) return frame_views(static_cast<float *>(c->peer_frame), n);
  if (c->frame) return frame_views(c->frame, n);
  FrameTargets t = c->fb;
  t.row_major = 0;
  return t;
}

void set_cam(cutrace_ctx *c, const float pos[3], const float up[3], const float fwd[3], const float right[3], float ambient) {
  Camera &cam = c->sv.cam;
  cam.pos = mk3(pos[0], pos[1], pos[2]);
  cam.up = mk3(up[0], up[1], up[2]);
  cam.forward = mk3(fwd[0], fwd[1], fwd[2]);
  cam.right = mk3(right[0], right[1], right[2]);
  cam.ambient = ambient;
}

__global__ void phong_pow_debug_kernel(const float *x, const float *e, float *out_powf, float *out_fast, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out_powf[i] = powf(x[i], e[i]);
  out_fast[i] = phong_pow(x[i], e[i], phong_pow_floor(e[i]));
}

}  // namespace

extern "C" {

uint32_t cutrace_abi_version(void) { return CUTRACE_ABI_VERSION; }

uint32_t cutrace_tile_size(void) { return CUTRACE_TILE; }

uint32_t cutrace_debug_segment_length(uint32_t n_work, uint32_t k, uint32_t G) { return G ? seg_len(n_work, k, G) : 0u; }
uint32_t cutrace_debug_segment_work(uint32_t o, uint32_t k, uint32_t G) { return G ? seg_to_work(o, k, G) : 0u; }

static TileMap debug_tile_map(uint32_t width, uint32_t height, uint32_t world, uint32_t curve) {
  TileMap tm{};
  tm.width = width; tm.height = height;
  tm.tiles_x = (width + CUTRACE_TILE - 1) / CUTRACE_TILE;
  tm.tiles_y = (height + CUTRACE_TILE - 1) / CUTRACE_TILE;
  tm.world = world > 1 ? world : 1;
  tm.n_tiles = tm.tiles_x * tm.tiles_y;
  tm.n_local_tiles = (tm.n_tiles + tm.world - 1) / tm.world;
  tile_permutation(tm.n_tiles, tm.world, &tm.perm_a, &tm.perm_ainv);
  tm.curve = curve ? 1u : 0u;
  return tm;
}
int cutrace_debug_tile_of_slot(uint32_t width, uint32_t height, uint32_t tile_world, uint32_t curve, uint32_t slot, uint32_t *tx, uint32_t *ty) {
  const TileMap tm = debug_tile_map(width, height, tile_world, curve);
  uint32_t x = 0, y = 0;
  const bool ok = tile_of_slot(tm, slot, x, y);
  if (tx) *tx = x;
  if (ty) *ty = y;
  return ok ? 1 : 0;
}
uint32_t cutrace_debug_slot_of_tile(uint32_t width, uint32_t height, uint32_t tile_world, uint32_t curve, uint32_t tx, uint32_t ty) {
  const TileMap tm = debug_tile_map(width, height, tile_world, curve);
  return slot_of_tile(tm, ty * tm.tiles_x + tx);
}

const char *cutrace_last_error(void) { return g_err.c_str(); }

void cutrace_default_opts(cutrace_opts *o) {
  if (!o) return;
  memset(o, 0, sizeof *o);
  o->fudge = 1e-3f;   // main.cu:30 passes the double literal 1e-3 into a float parameter
  o->bounces = 5;
  o->device = -1;
  o->tile_world = 1;
  o->leaf_size = 3;   // measured with the SAH treelet build: 3 beats 4 by 1 % on bunny.json 4K and on the 10 M-triangle hall (tools/leaf_sweep.py)
}

void cutrace_free(cutrace_ctx *c) {
  if (!c) return;
  DeviceGuard g(c->device);
  PhaseTimer ptimer; (void)ptimer;
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_frame(c);
  LAP("free: frame + graph");
  dfree(c->bvh.nodes, c->stream); dfree(c->bvh.nodes4, c->stream); dfree(c->bvh.prims, c->stream);
  dfree(c->planes, c->stream); dfree(c->materials, c->stream); dfree(c->lights, c->stream); dfree(c->obj_material, c->stream);
  dfree(c->obj_bounds, c->stream); dfree(c->pl_tbl, c->stream); dfree(c->pl_eps, c->stream);
  dfree(c->d_ctr, c->stream);
  LAP("free: scene buffers");
  if (c->stream) cudaStreamSynchronize(c->stream);
  pinned_counters_put(c->h_ctr);
  LAP("free: sync");
  stream_set_put(c->streams);   // waits for every stream of the set, then recycles it
  LAP("free: events + streams");
  delete c;
}

int cutrace_upload_scene(const cutrace_scene_desc *s, const cutrace_opts *opts, cutrace_ctx **out) {
  if (!out) return fail(CUTRACE_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  int rc = validate_desc(s);
  if (rc) return rc;
  cutrace_opts o;
  if (opts) o = *opts; else cutrace_default_opts(&o);
  if (o.bounces > 15) return fail(CUTRACE_ERR_INVALID_ARG, "bounces must be <= 15");
  if (o.tile_world > 1 && o.tile_rank >= o.tile_world) return fail(CUTRACE_ERR_INVALID_ARG, "tile_rank >= tile_world");
  if (o.leaf_size == 0) o.leaf_size = 3;
  if (o.leaf_size > CTB_MAX_LEAF) return fail(CUTRACE_ERR_INVALID_ARG, "leaf_size must be <= 8");

  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(CUTRACE_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
  int dev = o.device;
  if (dev < 0) CU(cudaGetDevice(&dev));
  if (dev >= n_dev) return fail(CUTRACE_ERR_NO_DEVICE, "requested CUDA device ordinal does not exist");

  PhaseTimer ptimer; (void)ptimer;
  cutrace_ctx *c = new (std::nothrow) cutrace_ctx();
  if (!c) return fail(CUTRACE_ERR_OUT_OF_MEMORY, "host allocation failed");
  c->device = dev;
  c->opts = o;
  DeviceGuard guard(dev);
  if (!guard.ok) { delete c; return fail(CUTRACE_ERR_CUDA, "cudaSetDevice failed"); }
  c->env_no_graph = getenv("CUTRACE_NO_GRAPH") != nullptr;
  if (const char *e = getenv("CUTRACE_SCHEDULER")) {   // developer override of cutrace_opts.flags: "frame" | "launches"
    c->opts.flags &= ~(CUTRACE_FLAG_FRAME_KERNEL | CUTRACE_FLAG_LAUNCHES | CUTRACE_FLAG_PIXEL_KERNEL);
    if (!strcmp(e, "frame")) c->opts.flags |= CUTRACE_FLAG_FRAME_KERNEL;
    else if (!strcmp(e, "launches")) c->opts.flags |= CUTRACE_FLAG_LAUNCHES;
    else if (!strcmp(e, "pixel")) c->opts.flags |= CUTRACE_FLAG_PIXEL_KERNEL;
  }
  c->env_export_with_color = getenv("CUTRACE_EXPORT_WITH_COLOR") != nullptr;      // developer toggle: remote G-buffer stores at the end of the frame
  c->env_graph_first = getenv("CUTRACE_GRAPH_FIRST") != nullptr;
  c->env_skip_export = getenv("CUTRACE_DEBUG_SKIP_EXPORT") != nullptr;      // timing experiments of profiles/r01_tuning.md only:
  c->env_local_color = getenv("CUTRACE_DEBUG_LOCAL_COLOR") != nullptr;      // they leave the peer frame incomplete

#define UP(call) do { int rc_ = (call); if (rc_) { cutrace_free(c); return rc_; } } while (0)
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_); cutrace_free(c); \
    return fail(e_ == cudaErrorMemoryAllocation ? CUTRACE_ERR_OUT_OF_MEMORY : CUTRACE_ERR_CUDA, m_); } } while (0)

  CUF(stream_set_get(c->device, &c->streams));
  if (o.stream) c->stream = (cudaStream_t)o.stream;
  else { c->stream = c->streams->main; c->own_stream = true; }
  for (int i = 0; i < 3; i++) c->aux[i] = c->streams->aux[i];
  c->copy_stream = c->streams->copy;
  c->ev_gbuf = c->streams->ev_gbuf;
  c->events = c->streams->events;
  LAP("validate + streams + events");

  // ---- flat records ----
  std::vector<PlaneRec> planes(s->n_planes);
  for (uint64_t i = 0; i < s->n_planes; i++) {
    PlaneRec &p = planes[i];
    p.px = s->pl_point[3 * i]; p.py = s->pl_point[3 * i + 1]; p.pz = s->pl_point[3 * i + 2]; p.obj = s->pl_object[i];
    p.nx = s->pl_normal[3 * i]; p.ny = s->pl_normal[3 * i + 1]; p.nz = s->pl_normal[3 * i + 2]; p.pad = 0;
  }
  std::vector<MaterialRec> mats(s->n_materials);
  bool all_opaque = true, any_child = false, any_both = false;
  for (uint32_t i = 0; i < s->n_materials; i++) {
    MaterialRec &m = mats[i];
    m.r = s->mat_color[3 * i]; m.g = s->mat_color[3 * i + 1]; m.b = s->mat_color[3 * i + 2]; m.specular = s->mat_specular[i];
    m.reflect = s->mat_reflect[i]; m.phong = s->mat_phong[i]; m.transparency = s->mat_transparency[i]; m.pad = 0;
    bool r = (double)m.reflect >= 1e-6, t = (double)m.transparency >= 1e-6;   // inc/shading.hpp:130,141
    // intensity += 1 - trans saturates in one step only if 1 - trans >= 1 (inc/shading.hpp:36-39)
    if (!(1.0f - m.transparency >= 1.0f)) all_opaque = false;
    any_child = any_child || r || t;
    any_both = any_both || (r && t);
  }
  c->max_children = any_both ? 2 : (any_child ? 1 : 0);
  std::vector<LightRec> lights(s->n_lights);
  for (uint32_t i = 0; i < s->n_lights; i++) {
    LightRec &l = lights[i];
    l.vx = s->light_vec[3 * i]; l.vy = s->light_vec[3 * i + 1]; l.vz = s->light_vec[3 * i + 2]; l.kind = s->light_kind[i];
    l.r = s->light_color[3 * i]; l.g = s->light_color[3 * i + 1]; l.b = s->light_color[3 * i + 2]; l.pad = 0;
  }
  // side of every light relative to every plane, in double (see plane_side_prepass in trace.cuh); entries too close to the plane are 0
  std::vector<float> pl_tbl, pl_eps;
  if (CTB_PLANE_PREPASS) {
    pl_tbl.resize((size_t)s->n_planes * s->n_lights); pl_eps.resize(2 * s->n_planes);
    double scale = 1e-30;
    for (uint64_t i = 0; i < 3 * s->n_planes; i++) scale = std::max(scale, std::fabs((double)s->pl_point[i]));
    for (uint32_t i = 0; i < 3 * s->n_lights; i++) if (s->light_kind[i / 3] == CUTRACE_LIGHT_POINT) scale = std::max(scale, std::fabs((double)s->light_vec[i]));
    for (uint64_t p = 0; p < s->n_planes; p++) {
      const double n[3] = {s->pl_normal[3 * p], s->pl_normal[3 * p + 1], s->pl_normal[3 * p + 2]};
      const double nn = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      pl_eps[2 * p] = (float)(1e-4 * nn * scale);
      pl_eps[2 * p + 1] = (float)(1e-5 * nn);
      for (uint32_t l = 0; l < s->n_lights; l++) {
        const double v[3] = {s->light_vec[3 * l], s->light_vec[3 * l + 1], s->light_vec[3 * l + 2]};
        double t, eps;
        if (s->light_kind[l] == CUTRACE_LIGHT_POINT) {
          t = n[0] * (s->pl_point[3 * p] - v[0]) + n[1] * (s->pl_point[3 * p + 1] - v[1]) + n[2] * (s->pl_point[3 * p + 2] - v[2]);
          eps = 1e-4 * nn * scale;
        } else {                                        // sun: shadow direction d = -direction / |direction|, table = -(d.n)
          const double vn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
          t = vn > 0 ? (v[0] * n[0] + v[1] * n[1] + v[2] * n[2]) / vn : 0.0;
          eps = 1e-4 * nn;
        }
        pl_tbl[p * s->n_lights + l] = (std::isfinite(t) && std::fabs(t) > eps) ? (float)t : 0.f;
      }
    }
  }
  std::vector<uint32_t> omat(s->obj_material, s->obj_material + s->n_objects);
  // (no stream synchronisation per array — six round trips were 0.04 ms of every upload: the vectors live until this function
  // returns, and the stream is synchronised after the primitive copies below)
  UP(upload(&c->planes, planes, c->stream, false));
  UP(upload(&c->materials, mats, c->stream, false));
  UP(upload(&c->lights, lights, c->stream, false));
  UP(upload(&c->obj_material, omat, c->stream, false));
  UP(upload(&c->pl_tbl, pl_tbl, c->stream, false));
  UP(upload(&c->pl_eps, pl_eps, c->stream, false));
  LAP("small record uploads");

  // ---- primitives + LBVH ----
  float *d_p1 = nullptr, *d_p2 = nullptr, *d_p3 = nullptr, *d_sc = nullptr, *d_sr = nullptr;
  uint32_t *d_to = nullptr, *d_so = nullptr;
  auto free_tmp = [&]() { dfree(d_p1, c->stream); dfree(d_p2, c->stream); dfree(d_p3, c->stream); dfree(d_sc, c->stream); dfree(d_sr, c->stream); dfree(d_to, c->stream); dfree(d_so, c->stream); };
#define CUT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e_); free_tmp(); cutrace_free(c); \
    return fail(e_ == cudaErrorMemoryAllocation ? CUTRACE_ERR_OUT_OF_MEMORY : CUTRACE_ERR_CUDA, m_); } } while (0)
  const uint64_t nt = s->n_triangles, ns = s->n_spheres;
  if (nt) {
    CUT(dmalloc(&d_p1, sizeof(float) * 3 * nt, c->stream)); CUT(dmalloc(&d_p2, sizeof(float) * 3 * nt, c->stream));
    CUT(dmalloc(&d_p3, sizeof(float) * 3 * nt, c->stream)); CUT(dmalloc(&d_to, sizeof(uint32_t) * nt, c->stream));
    CUT(cudaMemcpyAsync(d_p1, s->tri_p1, sizeof(float) * 3 * nt, cudaMemcpyHostToDevice, c->stream));
    CUT(cudaMemcpyAsync(d_p2, s->tri_p2, sizeof(float) * 3 * nt, cudaMemcpyHostToDevice, c->stream));
    CUT(cudaMemcpyAsync(d_p3, s->tri_p3, sizeof(float) * 3 * nt, cudaMemcpyHostToDevice, c->stream));
    CUT(cudaMemcpyAsync(d_to, s->tri_object, sizeof(uint32_t) * nt, cudaMemcpyHostToDevice, c->stream));
  }
  if (ns) {
    CUT(dmalloc(&d_sc, sizeof(float) * 3 * ns, c->stream)); CUT(dmalloc(&d_sr, sizeof(float) * ns, c->stream)); CUT(dmalloc(&d_so, sizeof(uint32_t) * ns, c->stream));
    CUT(cudaMemcpyAsync(d_sc, s->sph_center, sizeof(float) * 3 * ns, cudaMemcpyHostToDevice, c->stream));
    CUT(cudaMemcpyAsync(d_sr, s->sph_radius, sizeof(float) * ns, cudaMemcpyHostToDevice, c->stream));
    CUT(cudaMemcpyAsync(d_so, s->sph_object, sizeof(uint32_t) * ns, cudaMemcpyHostToDevice, c->stream));
  }
  LAP("primitive H2D");
  {
    uint32_t *d_kind = nullptr;
    if (s->obj_kind && s->n_objects) {
      CUT(dmalloc(&d_kind, sizeof(uint32_t) * s->n_objects, c->stream));
      CUT(cudaMemcpyAsync(d_kind, s->obj_kind, sizeof(uint32_t) * s->n_objects, cudaMemcpyHostToDevice, c->stream));
    }
    std::string oerr;
    rc = build_object_bounds(d_p1, d_p2, d_p3, d_to, (uint32_t)nt, s->n_objects, d_kind, &c->obj_bounds, c->stream, oerr);
    dfree(d_kind, c->stream);
    if (rc) { free_tmp(); cutrace_free(c); return fail(rc, oerr); }
  }
  BvhInput bi;
  bi.d_p1 = d_p1; bi.d_p2 = d_p2; bi.d_p3 = d_p3; bi.d_tri_obj = d_to; bi.n_tri = (uint32_t)nt;
  bi.d_sph_center = d_sc; bi.d_sph_radius = d_sr; bi.d_sph_obj = d_so; bi.n_sph = (uint32_t)ns;
  bi.n_objects = s->n_objects;
  bi.leaf_size = o.leaf_size; bi.stream = c->stream;
  bi.sah_treelets = !(o.flags & CUTRACE_FLAG_FAST_BUILD);
  CUT(cudaEventRecord(c->events[0], c->stream));
  std::string berr;
  rc = build_bvh(bi, c->bvh, berr);
  if (rc) { free_tmp(); cutrace_free(c); return fail(rc, berr); }
  CUT(cudaEventRecord(c->events[1], c->stream));
  CUT(cudaStreamSynchronize(c->stream));
  CUT(cudaEventElapsedTime(&c->stats.build_ms, c->events[0], c->events[1]));
  free_tmp();
  LAP("build_bvh");
  if (o.flags & CUTRACE_FLAG_VALIDATE_BVH) {
    rc = validate_bvh(c->bvh, c->stream, berr);
    if (rc) { cutrace_free(c); return fail(rc, berr); }
  }

  SceneView &sv = c->sv;
  sv.nodes = c->bvh.nodes; sv.prims = c->bvh.prims; sv.planes = c->planes; sv.materials = c->materials;
  sv.lights = c->lights; sv.obj_material = c->obj_material; sv.obj_bounds = c->obj_bounds;
  sv.pl_tbl = c->pl_tbl; sv.pl_eps = c->pl_eps;
  sv.n_prims = c->bvh.n_prims; sv.n_nodes = c->bvh.n_nodes; sv.n_planes = (uint32_t)s->n_planes;
  sv.n_lights = s->n_lights; sv.n_materials = s->n_materials; sv.n_objects = s->n_objects;
  sv.root = c->bvh.root;
  if (CTB_BVH4 && c->bvh.nodes4) {   // the kernels of this build walk the 4-wide tree (node 0 is its root, too)
    sv.nodes = reinterpret_cast<const Node *>(c->bvh.nodes4);
    sv.n_nodes = c->bvh.n_nodes4;
  }
  sv.all_opaque = all_opaque ? 1u : 0u;
  // a handful of primitives is looped over directly: no node loads, no stack (sphere_plane.json 1080p 0.56 -> 0.51 ms, same image)
  sv.brute_force = ((o.flags & CUTRACE_FLAG_BRUTE_FORCE) || c->bvh.n_prims <= 8u) ? 1u : 0u;
  sv.fudge = o.fudge;
  sv.scene_mag = 0.f;
  for (int a = 0; a < 3; a++) sv.scene_mag = fmaxf(sv.scene_mag, fmaxf(fabsf(c->bvh.lo[a]), fabsf(c->bvh.hi[a])));
  set_cam(c, s->cam_pos, s->cam_up, s->cam_forward, s->cam_right, s->ambient);

  CUF(dmalloc(&c->d_ctr, sizeof(FrameCounters), c->stream));
  c->h_ctr = pinned_counters_get();
  if (!c->h_ctr) { cutrace_free(c); return fail(CUTRACE_ERR_OUT_OF_MEMORY, "cudaHostAlloc failed"); }
  {
    void *dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, c->h_ctr, 0) == cudaSuccess) c->h_ctr_dev = static_cast<FrameStats *>(dp);
    else cudaGetLastError();
  }
  LAP("ctr alloc");
  {
    // shared-memory plan: everything (nodes + primitives) if it fits next to the SM, else the top of the tree
    int smem_optin = 0;
    CUF(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t whole = (size_t)sv.n_nodes * CTB_NODE_BYTES + (size_t)sv.n_prims * sizeof(PrimRec);
    const bool allow = !(o.flags & CUTRACE_FLAG_NO_SMEM_TOP) && !sv.brute_force && !CTB_BVH4;
    sv.smem_nodes = 0;
    if (allow && whole + 1024 > (size_t)smem_optin && sv.root == 0 && sv.n_nodes > 0) {
      uint32_t want = CTB_SMEM_TOP_NODES, n_top = 0;
      if (const char *e = getenv("CUTRACE_SMEM_TOP_NODES")) want = (uint32_t)atoi(e);
      std::string terr;
      rc = reorder_top(c->bvh, want, c->stream, &n_top, terr);
      if (rc) { cutrace_free(c); return fail(rc, terr); }
      sv.nodes = c->bvh.nodes;
      sv.smem_nodes = n_top;
    }
  }
  CUF(plan_launch(sv, !(o.flags & CUTRACE_FLAG_NO_SMEM_TOP), &c->cfg));
  LAP("plan_launch");
  if (c->cfg.mode == 1) sv.smem_nodes = sv.n_nodes;
  sv.smem_prims = c->cfg.mode == 1 ? sv.n_prims : 0;
  UP(alloc_frame(c, s->width, s->height));
  LAP("alloc_frame");
  c->stats.bvh_nodes = c->bvh.n_nodes;
  c->stats.bvh_depth = c->bvh.depth;
  c->stats.smem_nodes = sv.smem_nodes;
  *out = c;
  return CUTRACE_OK;
#undef UP
#undef CUF
#undef CUT
}

int cutrace_set_camera(cutrace_ctx *c, const float pos[3], const float up[3], const float forward[3], const float right[3],
                       float ambient, uint32_t width, uint32_t height) {
  if (!c || !pos || !up || !forward || !right) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(c->device);
  CU(cudaStreamSynchronize(c->stream));
  drop_graph(c);
  set_cam(c, pos, up, forward, right, ambient);
  if (width != c->tm.width || height != c->tm.height) {
    int rc = alloc_frame(c, width, height);
    if (rc) return rc;
  }
  c->rendered = false;
  return CUTRACE_OK;
}

int cutrace_render(cutrace_ctx *c, cutrace_stats *stats) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  if (!g.ok) return fail(CUTRACE_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = c->stream;
  const uint32_t bounces = c->opts.bounces;
  const uint32_t levels = (c->max_children > 0) ? bounces + 1 : 1;
  cutrace_stats &S = c->stats;
  S.rays_primary = S.rays_reflect = S.rays_transmit = S.rays_shadow = S.shadow_casts = 0;
  S.kernel_launches = 0; S.max_depth = 0.f; S.trace_ms = S.shade_ms = 0.f; S.gather_ms = 0.f;
  S.local_pixels = 0;
  S.local_pixels = c->local_pixels;   // pixels of this ctx that are inside the image (counted once, in alloc_frame)
  S.rays_primary = c->local_pixels;
  PhaseTimer ptimer; (void)ptimer;
  cudaEvent_t ev_begin = c->events[0], ev_end = c->events[1];
  if (!wants_pixel_kernel(c)) { int rc_ = alloc_queues(c); if (rc_) return rc_; }
  const bool branching = c->max_children >= 2 && bounces > 0;
  const bool serialize = (c->opts.flags & CUTRACE_FLAG_SERIALIZE) != 0;
  const FrameTargets out = frame_targets(c);
  // a peer frame gets its G-buffer from a coalescing export kernel; trace(0) then writes the local tile-major buffers
  FrameTargets gbuf = out, gsrc{};
  if (c->peer_frame && c->fb.depth) { gbuf = c->fb; gbuf.row_major = 0; gsrc = gbuf; }
  FrameTargets acc{};   // shade kernels of a branching scene accumulate here
  acc.color = c->local_color;
  uint32_t launches = 0;

  // Everything one batch of pixels needs, in dependency order.
  //   trace(L) -> trace(L+1) (ray queue)   and   trace(L) -> shade(L) (shade queue L)
  // The trace chain runs on the ctx stream; shade(L) runs on one of two auxiliary streams behind an event, so the
  // persistent CTAs of later kernels fill the SMs that the tail of an earlier kernel leaves idle.
  bool capturing = false, early_event = false;
  auto enqueue = [&](uint64_t base, uint32_t n_px) -> cudaError_t {
    cudaError_t e;
#define EQ(call) do { e = (call); if (e != cudaSuccess) return e; } while (0)
    if (branching && base == 0) EQ(cudaMemsetAsync(c->local_color, 0, sizeof(float) * 3 * c->n_local_px, st));
    EQ(cudaMemsetAsync(c->d_ctr, 0, sizeof(FrameCounters), st));
    for (uint32_t L = 0; L < levels; L++) {
      uint64_t bound = (uint64_t)n_px * (branching ? (1ull << L) : 1ull) + (uint64_t)c->cfg.grid_trace * (TRACE_THREADS / 32) * SLOT_BLOCK;
      if (bound > c->shade_cap[L]) bound = c->shade_cap[L];
      RayRec *in = c->rays[L & 1], *outq = c->rays[(L + 1) & 1];
      cudaEvent_t e0 = c->events[2 + 3 * L], e1 = c->events[3 + 3 * L], e2 = c->events[4 + 3 * L];
      if (serialize) EQ(cudaEventRecord(e0, st));
      launch_trace(c->cfg, c->sv, c->tm, L, bounces, (uint32_t)base, n_px, in, outq, (uint32_t)c->cap, c->shade[L], (uint32_t)c->shade_cap[L], c->d_ctr, gbuf,
                   c->nlev, (uint32_t)bound, st);
      EQ(cudaEventRecord(e1, st));
      if (L == 0) EQ(cudaEventRecordWithFlags(c->ev_gbuf, st, capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
      cudaStream_t ss = serialize ? st : c->aux[L & 1];
      if (!serialize) EQ(cudaStreamWaitEvent(ss, e1, 0));
      float *lc = branching ? nullptr : c->level_color + (size_t)L * 3 * c->batch_px;
      launch_shade(c->cfg, c->sv, L, c->shade[L], (uint32_t)c->shade_cap[L], c->d_ctr, acc, branching, lc, (uint32_t)base, (uint32_t)bound, ss);
      EQ(cudaEventRecord(e2, ss));
      launches += 2;
      if (L == 0 && gsrc.depth && !c->env_skip_export) {   // peer frame: ship the G-buffer now, under the remaining levels
        cudaStream_t xs = serialize ? st : c->aux[2];
        if (!serialize) EQ(cudaStreamWaitEvent(xs, e1, 0));
        launch_export_gbuffer(c->tm, (uint32_t)base, n_px, gsrc, out, xs);
        EQ(cudaEventRecord(c->events[60], xs));
        launches += 1;
      }

    }
    if (!serialize) for (uint32_t L = 0; L < levels; L++) EQ(cudaStreamWaitEvent(st, c->events[4 + 3 * L], 0));
    if (!serialize && gsrc.depth && !c->env_skip_export) EQ(cudaStreamWaitEvent(st, c->events[60], 0));
    {
      FrameTargets cout = out;
      if (c->peer_frame && c->fb.color && c->env_local_color) { cout = c->fb; cout.row_major = 0; }   // timing experiment only
      launch_combine(c->tm, c->nlev, c->level_color, 3ull * c->batch_px, branching ? 0u : levels, c->local_color, (uint32_t)base, n_px, cout,
                     FrameTargets{}, st);
    }
    launches += 1;
    return cudaGetLastError();
#undef EQ
  };

  // Default path: ONE persistent cooperative kernel per pixel batch runs every bounce level (render.cu: frame_kernel).  It expects
  // cleared counters and leaves them cleared; its counters arrive in c->h_ctr (mapped pinned memory) without a copy.
  // With an early G-buffer download pending (cutrace_render_download) trace(0) runs as its own kernel first, so that the
  // copy engine can start behind it while the frame kernel works on the bounce levels.
  // which scheduler: see CUTRACE_FLAG_FRAME_KERNEL / CUTRACE_FLAG_PIXEL_KERNEL in cutrace.h
  const uint32_t forced = c->opts.flags & (CUTRACE_FLAG_FRAME_KERNEL | CUTRACE_FLAG_LAUNCHES | CUTRACE_FLAG_PIXEL_KERNEL);
  const bool use_pixel = wants_pixel_kernel(c);
  const bool want_frame = !use_pixel && (forced & CUTRACE_FLAG_FRAME_KERNEL) != 0;
  const bool use_frame = want_frame && !serialize && !c->frame_kernel_failed && c->cfg.grid_frame > 0 && c->h_ctr_dev;
  auto enqueue_frame = [&](uint64_t base, uint32_t n_px, bool split_primary) -> cudaError_t {
    cudaError_t e;
    if (c->ctr_dirty) {
      if ((e = cudaMemsetAsync(c->d_ctr, 0, sizeof(FrameCounters), st)) != cudaSuccess) return e;
      c->ctr_dirty = false;
    }
    if (branching && base == 0 && (e = cudaMemsetAsync(c->local_color, 0, sizeof(float) * 3 * c->n_local_px, st)) != cudaSuccess) return e;
    const uint64_t warps = (uint64_t)std::max(c->cfg.grid_trace, c->cfg.grid_frame) * (TRACE_THREADS / 32);
    FrameArgs fa{};
    fa.sv = c->sv; fa.tm = c->tm; fa.bounces = bounces; fa.levels = levels; fa.first_level = 0;
    fa.px_base = (uint32_t)base; fa.n_px = n_px;
    fa.rays[0] = c->rays[0]; fa.rays[1] = c->rays[1]; fa.ray_cap = (uint32_t)c->cap;
    uint64_t bound_max = n_px;
    for (uint32_t L = 0; L < levels; L++) {
      fa.shade[L] = c->shade[L]; fa.shade_cap[L] = (uint32_t)c->shade_cap[L];
      uint64_t bound = (uint64_t)n_px * (branching ? (1ull << L) : 1ull) + warps * SLOT_BLOCK;
      if (bound > c->shade_cap[L]) bound = c->shade_cap[L];
      bound_max = std::max(bound_max, bound);
    }
    fa.ctr = c->d_ctr; fa.host_stats = c->h_ctr_dev;
    fa.gbuf = gbuf; fa.out = out; fa.gsrc = gsrc; fa.acc = acc;
    if (c->env_skip_export) fa.gsrc = FrameTargets{};
    fa.level_color = branching ? nullptr : c->level_color; fa.level_stride = 3ull * c->batch_px; fa.nlev = c->nlev;
    fa.local_color = c->local_color; fa.combine_levels = branching ? 0u : levels; fa.atomic_accumulate = branching ? 1 : 0;
    fa.combine = 1; fa.export_with_color = c->env_export_with_color ? 1 : 0;
    if (c->peer_frame && c->fb.color && c->env_local_color) { fa.out = c->fb; fa.out.row_major = 0; }   // timing experiment only
    if (split_primary) {
      const uint64_t b0 = std::min<uint64_t>((uint64_t)n_px + warps * SLOT_BLOCK, c->shade_cap[0]);
      launch_trace(c->cfg, c->sv, c->tm, 0, bounces, (uint32_t)base, n_px, c->rays[0], c->rays[1], (uint32_t)c->cap, c->shade[0], (uint32_t)c->shade_cap[0],
                   c->d_ctr, gbuf, c->nlev, (uint32_t)b0, st);
      if ((e = cudaGetLastError()) != cudaSuccess) return e;
      if ((e = cudaEventRecord(c->ev_gbuf, st)) != cudaSuccess) return e;
      fa.first_level = 1;
      launches += 1;
    }
    c->ctr_dirty = true;   // until the kernel has run to its end
    if ((e = launch_frame(c->cfg, fa, (uint32_t)std::min<uint64_t>(bound_max, 0xffffffffull), st)) != cudaSuccess) return e;
    launches += 1;
    return cudaSuccess;
  };

  auto enqueue_pixel = [&](uint64_t base, uint32_t n_px) -> cudaError_t {
    cudaError_t e;
    if (c->ctr_dirty) {
      if ((e = cudaMemsetAsync(c->d_ctr, 0, sizeof(FrameCounters), st)) != cudaSuccess) return e;
      c->ctr_dirty = false;
    }
    PixelArgs pa{};
    pa.sv = c->sv; pa.tm = c->tm; pa.bounces = bounces; pa.px_base = (uint32_t)base; pa.n_px = n_px;
    static const bool dbg_no_host_stats = getenv("CUTRACE_DEBUG_NO_HOST_STATS") != nullptr;   // timing experiment: the statistics stay on the device
    pa.ctr = c->d_ctr; pa.host_stats = dbg_no_host_stats ? nullptr : c->h_ctr_dev; pa.out = out;
    if (c->dl_direct_on) { pa.out2 = c->dl_direct; pa.tm.wide_warps = 1u; }   // 16 x 2 warps: whole tile rows per store over PCIe
    static const int dbg_wide = [] { const char *e = getenv("CUTRACE_DEBUG_WIDE_WARPS"); return e ? atoi(e) : -1; }();   // tuning experiments
    if (dbg_wide >= 0) pa.tm.wide_warps = dbg_wide ? 1u : 0u;
    c->ctr_dirty = true;   // until the kernel has run to its end
    if ((e = launch_pixel(c->cfg, pa, st)) != cudaSuccess) return e;
    if (early_event) {   // cutrace_render_download: the G-buffer is complete when the kernel is
      if ((e = cudaEventRecord(c->ev_gbuf, st)) != cudaSuccess) return e;
    }
    launches += 1;
    return cudaSuccess;
  };

  // One batch (the normal case): the frame is a CUDA graph, captured from the code above and replayed afterwards — one
  // launch instead of ~40 API calls, which is what a 0.05 .. 2 ms frame is bound by.  The FIRST frame of a ctx is
  // enqueued directly: capture + instantiate cost more than the ~40 calls they replace, and a caller that renders one
  // frame per scene (the reference's main.cu does) never gets that back; the graph is built on the second frame.
  const bool single_batch = c->batch_px >= c->n_local_px;
  const bool use_graph = !use_frame && !use_pixel && single_batch && !serialize && !c->graph_failed && !c->env_no_graph &&
                         (c->frames_rendered > 0 || c->env_graph_first);
  if (use_graph && !c->graph) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      capturing = true;
      e = enqueue(0, (uint32_t)c->n_local_px);
      capturing = false;
      cudaError_t e2 = cudaStreamEndCapture(st, &g);
      if (e == cudaSuccess) e = e2;
      if (e == cudaSuccess) e = cudaGraphInstantiate(&c->graph, g, 0);
      if (g) cudaGraphDestroy(g);
    }
    if (e != cudaSuccess) { cudaGetLastError(); c->graph = nullptr; c->graph_failed = true; }
    c->graph_launches = launches;   // kernels per frame, remembered for the replays
    LAP("render: graph capture+instantiate");
  }
  float max_depth = 0.f;
  CU(cudaEventRecord(ev_begin, st));
  for (uint64_t base = 0; base < c->n_local_px; base += c->batch_px) {
    const uint32_t n_px = (uint32_t)std::min<uint64_t>(c->batch_px, c->n_local_px - base);
    const bool early_gbuf = single_batch && c->frame && !c->peer_frame && (c->dl_depth || c->dl_normal || c->dl_id);
    bool frame_done = false;
    uint32_t sched = 0;
    if (use_pixel) {
      launches = 0;
      early_event = early_gbuf;
      CU(enqueue_pixel(base, n_px));
      frame_done = true;
      sched = 2;
      S.kernel_launches += launches;
    } else if (use_frame) {
      launches = 0;
      cudaError_t fe = enqueue_frame(base, n_px, early_gbuf);
      if (fe == cudaSuccess) {
        frame_done = true;
        sched = 1;
        S.kernel_launches += launches;
      } else {   // e.g. cudaErrorCooperativeLaunchTooLarge under MPS limits: stay on the multi-launch path from now on
        cudaGetLastError();
        c->frame_kernel_failed = true;
        CU(cudaStreamSynchronize(st));
      }
    }
    if (!frame_done) {
      c->ctr_dirty = true;
      if (c->graph && use_graph) {
        CU(cudaGraphLaunch(c->graph, st));
        S.kernel_launches = c->graph_launches;
      } else {
        launches = 0;
        CU(enqueue(base, n_px));
        S.kernel_launches += launches;
      }
    }
    if (early_gbuf) {
      // G-buffer -> host underneath the remaining bounce levels (copy engine; the frame is already row-major)
      const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
      const FrameTargets v = frame_views(c->frame, n);
      CU(cudaStreamWaitEvent(c->copy_stream, c->ev_gbuf, 0));
      if (c->dl_depth) CU(cudaMemcpyAsync(c->dl_depth, v.depth, sizeof(float) * n, cudaMemcpyDeviceToHost, c->copy_stream));
      if (c->dl_id) CU(cudaMemcpyAsync(c->dl_id, v.hit_id, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, c->copy_stream));
      if (c->dl_normal) CU(cudaMemcpyAsync(c->dl_normal, v.normal, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->copy_stream));
      c->dl_depth = c->dl_normal = nullptr; c->dl_id = nullptr;   // consumed
    }
    if (!frame_done) CU(cudaMemcpyAsync(c->h_ctr, &c->d_ctr->st, sizeof(FrameStats), cudaMemcpyDeviceToHost, st));
    if (base + c->batch_px >= c->n_local_px) CU(cudaEventRecord(ev_end, st));
    LAP("render: enqueue");
    CU(cudaStreamSynchronize(st));
    LAP("render: wait for the frame");
    const FrameStats &h = *c->h_ctr;
    S.scheduler = sched;
    if (frame_done) {
      c->ctr_dirty = false;   // the frame kernel cleared the device counters on its way out
      c->phase_count = sched == 1 ? levels + 1 : 0;
      for (uint32_t p = 0; p <= levels && sched == 1; p++)
        c->phase_ms[p] = h.phase_ns[p] >= h.phase_ns[17] ? (float)((double)(h.phase_ns[p] - h.phase_ns[17]) * 1e-6) : 0.f;
    } else {
      c->phase_count = 0;
    }
    if (h.overflow) return fail(CUTRACE_ERR_INTERNAL, "internal: ray queue overflow (a queue reservation did not fit; the frame is incomplete)");
    S.rays_reflect += h.rays_reflect;
    S.rays_transmit += h.rays_transmit;
    S.shadow_casts += h.shadow_casts;
    S.rays_shadow += (uint64_t)h.shade_records * c->sv.n_lights;
    float md;
    memcpy(&md, &h.max_depth_bits, 4);
    if (md > max_depth) max_depth = md;
    if (serialize) {
      for (uint32_t L = 0; L < levels; L++) {
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, c->events[2 + 3 * L], c->events[3 + 3 * L]));
        CU(cudaEventElapsedTime(&b, c->events[3 + 3 * L], c->events[4 + 3 * L]));
        S.trace_ms += a; S.shade_ms += b;
      }
    }
  }
  CU(cudaEventElapsedTime(&S.render_ms, ev_begin, ev_end));
  S.max_depth = max_depth;
  c->rendered = true;
  c->frames_rendered++;
  if (stats) *stats = S;
  return CUTRACE_OK;
}

// device address of a pinned + mapped host pointer (cutrace_host_alloc / cutrace_host_register), NULL for anything else
static void *mapped_device_ptr(const void *p) {
  if (!p) return nullptr;
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int cutrace_render_download(cutrace_ctx *c, float *depth, float *normal, float *color, uint32_t *hit_id, float *max_depth,
                            cutrace_stats *stats) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  // Pixel kernel + pinned, mapped destinations (cutrace_host_alloc / cutrace_host_register): every pixel is stored to the
  // ctx's frame AND straight into the caller's images as it is finished — the download rides under the render, no copy follows.
  if (wants_pixel_kernel(c) && c->frame && !c->peer_frame && !getenv("CUTRACE_DEBUG_NO_DIRECT_DOWNLOAD")) {
    FrameTargets t{};
    t.depth = static_cast<float *>(mapped_device_ptr(depth)); t.normal = static_cast<float *>(mapped_device_ptr(normal));
    t.color = static_cast<float *>(mapped_device_ptr(color)); t.hit_id = static_cast<uint32_t *>(mapped_device_ptr(hit_id));
    t.row_major = 1;
    if ((!depth || t.depth) && (!normal || t.normal) && (!color || t.color) && (!hit_id || t.hit_id) && (depth || normal || color || hit_id)) {
      c->dl_direct = t; c->dl_direct_on = true;
      int rc = cutrace_render(c, stats);
      c->dl_direct_on = false;
      if (rc) return rc;
      if (max_depth) *max_depth = c->stats.max_depth;
      return CUTRACE_OK;
    }
  }
  const bool fused = c->frame && !c->peer_frame && c->batch_px >= c->n_local_px;
  if (fused) { c->dl_depth = depth; c->dl_normal = normal; c->dl_id = hit_id; }
  int rc = cutrace_render(c, stats);
  const bool early = fused && !c->dl_depth && !c->dl_normal && !c->dl_id;   // render consumed the request
  c->dl_depth = c->dl_normal = nullptr; c->dl_id = nullptr;
  if (rc) return rc;
  if (!early) return cutrace_download(c, depth, normal, color, hit_id, max_depth);
  const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
  PhaseTimer ptimer; (void)ptimer;
  if (color) CU(cudaMemcpyAsync(color, frame_views(c->frame, n).color, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  LAP("render_download: colour D2H");
  CU(cudaStreamSynchronize(c->copy_stream));
  LAP("render_download: G-buffer D2H tail");
  if (max_depth) *max_depth = c->stats.max_depth;
  return CUTRACE_OK;
}

int cutrace_get_phase_ms(cutrace_ctx *c, float *out, uint32_t capacity, uint32_t *n_out) {
  if (!c || (!out && capacity) || !n_out) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  *n_out = c->phase_count;
  for (uint32_t p = 0; p < c->phase_count && p < capacity; p++) out[p] = c->phase_ms[p];
  return CUTRACE_OK;
}

#ifdef CTB_PIXEL_STAMPS
int cutrace_debug_pixel_stamps(cutrace_ctx *c, unsigned long long *out) {   // tuning builds: the seven %globaltimer stamps of the last frame
  for (int i = 0; i < 7; i++) out[i] = c->h_ctr->phase_ns[i];
  return 0;
}
#endif

int cutrace_get_stats(cutrace_ctx *c, cutrace_stats *stats) {
  if (!c || !stats) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  *stats = c->stats;
  return CUTRACE_OK;
}

// Row-major device images of the last frame: the ctx's own frame, or (sharded ctx on the gather path) its tiles
// un-tiled into a staging frame whose foreign tiles read as misses.
static int full_frame_views(cutrace_ctx *c, FrameTargets *v) {
  cudaStream_t st = c->stream;
  const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
  if (c->peer_frame) return fail(CUTRACE_ERR_STATE, "this ctx renders into another ctx's frame (cutrace_frame_ipc_import); download from the exporting ctx");
  if (c->frame) { *v = frame_views(c->frame, n); return CUTRACE_OK; }
  if (c->st_px != n) {
    dfree(c->st_depth, st);
    c->st_depth = nullptr; c->st_px = 0;
    CU(dmalloc(&c->st_depth, 32ull * n, st));
    c->st_px = n;
  }
  *v = frame_views(c->st_depth, n);
  launch_fill_sentinels(v->depth, v->normal, v->color, v->hit_id, n, st);
  launch_untile(c->tm, c->tm.world, c->fb.depth, c->fb.normal, c->fb.color, c->fb.hit_id, 0, (int)c->tm.rank, v->depth, v->normal, v->color,
                v->hit_id, st);
  CU(cudaGetLastError());
  return CUTRACE_OK;
}

int cutrace_download(cutrace_ctx *c, float *depth, float *normal, float *color, uint32_t *hit_id, float *max_depth) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  if (!c->rendered) return fail(CUTRACE_ERR_STATE, "cutrace_download called before cutrace_render");
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream;
  const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
  FrameTargets v;
  int rc = full_frame_views(c, &v);
  if (rc) return rc;
  if (depth) CU(cudaMemcpyAsync(depth, v.depth, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  if (normal) CU(cudaMemcpyAsync(normal, v.normal, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
  if (color) CU(cudaMemcpyAsync(color, v.color, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
  if (hit_id) CU(cudaMemcpyAsync(hit_id, v.hit_id, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (max_depth) *max_depth = c->stats.max_depth;
  return CUTRACE_OK;
}

int cutrace_download_bytes(cutrace_ctx *c, uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb, float *max_depth) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  if (!c->rendered) return fail(CUTRACE_ERR_STATE, "cutrace_download_bytes called before cutrace_render");
  DeviceGuard g(c->device);
  cudaStream_t st = c->stream;
  const uint64_t n = (uint64_t)c->tm.width * c->tm.height;
  FrameTargets v;
  int rc = full_frame_views(c, &v);
  if (rc) return rc;
  if (c->st_bytes_px != n) {
    dfree(c->st_bytes, st); c->st_bytes = nullptr; c->st_bytes_px = 0;
    CU(dmalloc(&c->st_bytes, 9 * n, st));
    c->st_bytes_px = n;
  }
  uint8_t *d8 = c->st_bytes, *n8 = c->st_bytes + 3 * n, *c8 = c->st_bytes + 6 * n;
  launch_encode_bytes(depth_rgb ? v.depth : nullptr, normal_rgb ? v.normal : nullptr, color_rgb ? v.color : nullptr,
                      c->stats.max_depth, n, d8, n8, c8, st);
  CU(cudaGetLastError());
  if (depth_rgb) CU(cudaMemcpyAsync(depth_rgb, d8, 3 * n, cudaMemcpyDeviceToHost, st));
  if (normal_rgb) CU(cudaMemcpyAsync(normal_rgb, n8, 3 * n, cudaMemcpyDeviceToHost, st));
  if (color_rgb) CU(cudaMemcpyAsync(color_rgb, c8, 3 * n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (max_depth) *max_depth = c->stats.max_depth;
  return CUTRACE_OK;
}

int cutrace_frame_device(cutrace_ctx *c, float **depth, float **normal, float **color, uint32_t **hit_id) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  if (!c->frame) return fail(CUTRACE_ERR_STATE, "this ctx has no row-major frame of its own (sharded ctx without cutrace_frame_ipc_export)");
  FrameTargets v = frame_views(c->frame, (uint64_t)c->tm.width * c->tm.height);
  if (depth) *depth = v.depth;
  if (normal) *normal = v.normal;
  if (color) *color = v.color;
  if (hit_id) *hit_id = v.hit_id;
  return CUTRACE_OK;
}

int cutrace_frame_ipc_export(cutrace_ctx *c, void *handle) {
  if (!c || !handle) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) + 16 == CUTRACE_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(c->device);
  CU(cudaStreamSynchronize(c->stream));
  if (!c->frame_is_ipc) {   // pool memory cannot be exported: give the frame its own cudaMalloc block
    float *blk = nullptr;
    CU(cudaMalloc(&blk, 32ull * c->tm.width * c->tm.height));
    dfree(c->frame, c->stream);
    c->frame = blk;
    c->frame_is_ipc = true;
    c->rendered = false;
    drop_graph(c);
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->frame));
  memset(handle, 0, CUTRACE_IPC_HANDLE_BYTES);
  memcpy(handle, &h, sizeof h);
  const uint32_t dims[2] = {c->tm.width, c->tm.height};
  memcpy(static_cast<char *>(handle) + sizeof h, dims, sizeof dims);
  return CUTRACE_OK;
}

int cutrace_frame_ipc_import(cutrace_ctx *c, const void *handle) {
  if (!c || !handle) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard g(c->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  uint32_t dims[2];
  memcpy(dims, static_cast<const char *>(handle) + sizeof h, sizeof dims);
  if (dims[0] != c->tm.width || dims[1] != c->tm.height)
    return fail(CUTRACE_ERR_INVALID_ARG, "the exported frame is " + std::to_string(dims[0]) + "x" + std::to_string(dims[1]) + ", this ctx renders " +
                                             std::to_string(c->tm.width) + "x" + std::to_string(c->tm.height));
  CU(cudaStreamSynchronize(c->stream));
  if (c->peer_frame && c->peer_is_ipc) cudaIpcCloseMemHandle(c->peer_frame);
  c->peer_frame = nullptr;
  void *p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_frame = p;
  c->peer_is_ipc = true;
  c->rendered = false;
  drop_graph(c);
  return CUTRACE_OK;
}

int cutrace_enable_peer_access(int device, int peer_device) {
  if (device == peer_device) return CUTRACE_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(CUTRACE_ERR_NO_DEVICE, "cudaSetDevice failed");
  int can = 0;
  CU(cudaDeviceCanAccessPeer(&can, device, peer_device));
  if (!can) return fail(CUTRACE_ERR_CUDA, "devices cannot access each other's memory");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return CUTRACE_OK; }
  CU(e);
  return CUTRACE_OK;
}

int cutrace_set_frame_max_depth(cutrace_ctx *c, float max_depth) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  c->stats.max_depth = max_depth;
  return CUTRACE_OK;
}

int cutrace_frame_attach(cutrace_ctx *c, void *frame_block, uint32_t width, uint32_t height) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  void *dev_ptr = frame_block;
  bool host_frame = false;
  if (frame_block) {
    if (width != c->tm.width || height != c->tm.height)
      return fail(CUTRACE_ERR_INVALID_ARG, "the frame block is " + std::to_string(width) + "x" + std::to_string(height) + ", this ctx renders " +
                                               std::to_string(c->tm.width) + "x" + std::to_string(c->tm.height));
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, frame_block);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CUTRACE_ERR_INVALID_ARG, "frame_block is not memory CUDA knows (device, or registered / pinned host memory)"); }
    if (at.type == cudaMemoryTypeHost) {
      if (!at.devicePointer) return fail(CUTRACE_ERR_INVALID_ARG, "host frame_block is not mapped into the device address space");
      dev_ptr = at.devicePointer;
      host_frame = true;
    } else if (at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged) {
      return fail(CUTRACE_ERR_INVALID_ARG, "frame_block is unregistered host memory: register it with cutrace_host_register first");
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  if (c->peer_frame && c->peer_is_ipc) cudaIpcCloseMemHandle(c->peer_frame);
  c->peer_frame = dev_ptr;   // NULL detaches
  c->peer_is_ipc = false;
  // per-pixel stores over PCIe: 16 x 2 warps write whole 16-pixel tile rows (64 / 192-byte segments instead of 32 / 96)
  c->tm.wide_warps = (host_frame && !getenv("CUTRACE_DEBUG_NARROW_WARPS")) ? 1u : 0u;
  c->rendered = false;
  drop_graph(c);
  return CUTRACE_OK;
}

int cutrace_host_register(void *ptr, size_t bytes) {
  if (!ptr || !bytes) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  return CUTRACE_OK;
}

int cutrace_host_unregister(void *ptr) {
  if (!ptr) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  CU(cudaHostUnregister(ptr));
  return CUTRACE_OK;
}

int cutrace_device_buffers(cutrace_ctx *c, float **depth, float **normal, float **color, uint32_t **hit_id, uint64_t *n_local_px_padded) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  if (depth) *depth = c->fb.depth;
  if (normal) *normal = c->fb.normal;
  if (color) *color = c->fb.color;
  if (hit_id) *hit_id = c->fb.hit_id;
  if (n_local_px_padded) *n_local_px_padded = c->n_local_px;
  return CUTRACE_OK;
}

int cutrace_untile_device(cutrace_ctx *c, uint32_t world, const float *g_depth, const float *g_normal, const float *g_color,
                          const uint32_t *g_id, uint64_t stride_px, float *depth, float *normal, float *color, uint32_t *hit_id) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  if (world == 0) world = 1;
  if ((depth && !g_depth) || (normal && !g_normal) || (color && !g_color) || (hit_id && !g_id))
    return fail(CUTRACE_ERR_INVALID_ARG, "output requested without its gathered input");
  DeviceGuard g(c->device);
  cudaEvent_t e0 = c->events[70], e1 = c->events[71];
  CU(cudaEventRecord(e0, c->stream));
  launch_untile(c->tm, world, g_depth, g_normal, g_color, g_id, stride_px, -1, depth, normal, color, hit_id, c->stream);
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaEventElapsedTime(&c->stats.gather_ms, e0, e1));
  return CUTRACE_OK;
}

int cutrace_encode_bytes_device(cutrace_ctx *c, const float *depth, const float *normal, const float *color, float max_depth,
                                uint64_t n_px, uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  launch_encode_bytes(depth, normal, color, max_depth, n_px, depth_rgb, normal_rgb, color_rgb, c->stream);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  return CUTRACE_OK;
}

int cutrace_trim_memory(int device) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(CUTRACE_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0) CU(cudaGetDevice(&device));
  if (device >= n_dev) return fail(CUTRACE_ERR_NO_DEVICE, "requested CUDA device ordinal does not exist");
  DeviceGuard g(device);
  CU(cudaDeviceSynchronize());
  pool_trim(device);
  return CUTRACE_OK;
}

void *cutrace_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { g_err = "cudaHostAlloc failed"; return nullptr; }
  return p;
}

void cutrace_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int cutrace_debug_phong_pow(const float *x, const float *e, float *out_powf, float *out_fast, uint32_t n, int device) {
  if (n && (!x || !e || !out_powf || !out_fast)) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(CUTRACE_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0) CU(cudaGetDevice(&device));
  DeviceGuard g(device);
  if (n == 0) return CUTRACE_OK;
  float *d = nullptr;   // x | e | powf | fast
  CU(cudaMalloc(&d, sizeof(float) * 4 * (size_t)n));
  int rc = CUTRACE_OK;
  if (cudaMemcpy(d, x, sizeof(float) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d + n, e, sizeof(float) * n, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(CUTRACE_ERR_CUDA, "H2D copy failed");
  if (!rc) {
    phong_pow_debug_kernel<<<(n + 255) / 256, 256>>>(d, d + n, d + 2 * (size_t)n, d + 3 * (size_t)n, n);
    if (cudaDeviceSynchronize() != cudaSuccess) rc = fail(CUTRACE_ERR_CUDA, "phong_pow_debug_kernel failed");
  }
  if (!rc && (cudaMemcpy(out_powf, d + 2 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess ||
              cudaMemcpy(out_fast, d + 3 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost) != cudaSuccess)) rc = fail(CUTRACE_ERR_CUDA, "D2H copy failed");
  cudaFree(d);
  return rc;
}

int cutrace_debug_radix_sort(uint64_t *keys, uint32_t *values, uint32_t n, int device) {
  if (n && (!keys || !values)) return fail(CUTRACE_ERR_INVALID_ARG, "NULL argument");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(CUTRACE_ERR_NO_DEVICE, "no CUDA device");
  if (device < 0) CU(cudaGetDevice(&device));
  DeviceGuard g(device);
  if (n == 0) return CUTRACE_OK;
  uint64_t *dk = nullptr;
  uint32_t *dv = nullptr;
  CU(cudaMalloc(&dk, sizeof(uint64_t) * n));
  cudaError_t e = cudaMalloc(&dv, sizeof(uint32_t) * n);  // plain allocations: test hook, no ctx
  if (e != cudaSuccess) { cudaFree(dk); return fail(CUTRACE_ERR_OUT_OF_MEMORY, "cudaMalloc failed"); }
  std::string err;
  int rc = CUTRACE_OK;
  if (cudaMemcpy(dk, keys, sizeof(uint64_t) * n, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(dv, values, sizeof(uint32_t) * n, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(CUTRACE_ERR_CUDA, "H2D copy failed");
  if (!rc) { rc = radix_sort_pairs(dk, dv, n, nullptr, err); if (rc) fail(rc, err); }
  if (!rc && (cudaMemcpy(keys, dk, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess ||
              cudaMemcpy(values, dv, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess)) rc = fail(CUTRACE_ERR_CUDA, "D2H copy failed");
  cudaFree(dk); cudaFree(dv);
  return rc;
}

int cutrace_validate_bvh(cutrace_ctx *c) {
  if (!c) return fail(CUTRACE_ERR_INVALID_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  std::string err;
  int rc = validate_bvh(c->bvh, c->stream, err);
  return rc ? fail(rc, err) : CUTRACE_OK;
}

}  // extern "C"
