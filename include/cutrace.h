/*
 * cutrace.h — C-ABI of the B200-native render path for cutrace scenes.
 *
 * This is the drop-in boundary for the reference's render operator.  The reference has no FFI
 * layer; the path sits behind a header-only C++ template operator
 *
 *     cutrace::gpu::render<S, bounces, tpb>(scene, fudge, max, depth_map, color_map, normal_map,
 *                                           render_ms, total_ms)          (inc/kernel.hpp:86-130)
 *
 * whose input is produced by cutrace::cpu::schema::default_to_gpu() (inc/default_schema.hpp:935-937,
 * inc/cpu_to_gpu.hpp:188-198).  The three entry points below replace exactly those two calls plus
 * the row-wise cudaMemcpy read-back (inc/kernel.hpp:110-114):
 *
 *     default_to_gpu(scene)                     ->  cutrace_upload_scene()
 *     gpu::render<S,5,256>(scene, 1e-3, ...)     ->  cutrace_render()
 *     the per-row D2H copies + max-depth scan   ->  cutrace_download()
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  See INTEGRATION.md for
 * the binding a cutrace maintainer would add in main.cu.
 *
 * Conventions
 *   - every function returns CUTRACE_OK (0) or a negative cutrace_status; the library never
 *     prints, never calls exit(); cutrace_last_error() returns a thread-local message.
 *   - host pointers in cutrace_scene_desc are borrowed for the duration of the call only.
 *   - one ctx = one scene on one device; calls on a ctx must be externally serialised, distinct
 *     ctxs are independent.
 *   - images are row-major, row 0 = top, pixel (x,y) at y*width+x (inc/kernel.hpp:44-54).
 *   - miss sentinels follow the reference: depth +INF, normal {0,0,0}, colour {0,0,0}
 *     (inc/kernel.hpp:47-56, inc/shading.hpp:119,153); hit_id on a miss is CUTRACE_NO_HIT.
 */
#ifndef CUTRACE_B200_CUTRACE_H
#define CUTRACE_B200_CUTRACE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUTRACE_ABI_VERSION 1u
#define CUTRACE_NO_HIT 0xFFFFFFFFu

typedef enum cutrace_status {
  CUTRACE_OK = 0,
  CUTRACE_ERR_INVALID_ARG = -1,   /* NULL pointer, bad size, index out of range in the scene */
  CUTRACE_ERR_CUDA = -2,          /* a CUDA runtime call failed; message has the CUDA error   */
  CUTRACE_ERR_NO_DEVICE = -3,     /* no CUDA device / requested ordinal does not exist        */
  CUTRACE_ERR_OUT_OF_MEMORY = -4, /* device or host allocation failed                         */
  CUTRACE_ERR_STATE = -5,         /* e.g. download before any render                          */
  CUTRACE_ERR_INTERNAL = -6       /* BVH validation failed, queue overflow after retry, ...   */
} cutrace_status;

/* object kinds: variant order of default_gpu_object (inc/default_schema.hpp:920) */
#define CUTRACE_OBJ_TRIANGLE 0u
#define CUTRACE_OBJ_MESH 1u
#define CUTRACE_OBJ_PLANE 2u
#define CUTRACE_OBJ_SPHERE 3u

/* light kinds: variant order of default_gpu_light (inc/default_schema.hpp:921) */
#define CUTRACE_LIGHT_SUN 0u
#define CUTRACE_LIGHT_POINT 1u

/*
 * Flat structure-of-arrays scene.  Replaces gpu_scene_{gpu_array<gpu_variant<...>> x3, cam}
 * (inc/gpu_types.hpp:263-287, inc/cpu_to_gpu.hpp:69-199).
 *
 * object ids: `n_objects` is the length of the reference's scene.objects array; every primitive
 * carries the index of the object it belongs to (a mesh of 1000 triangles is ONE object, see
 * inc/ray_cast.hpp:45).  Triangles of one mesh must keep file order (tie-break rule,
 * inc/default_schema.hpp:133-141).
 */
typedef struct cutrace_scene_desc {
  uint32_t abi_version;      /* must be CUTRACE_ABI_VERSION */

  /* camera after look_at (inc/default_schema.hpp:370-374): unit forward/right/up */
  float cam_pos[3];
  float cam_up[3];
  float cam_forward[3];
  float cam_right[3];
  float ambient;             /* cam.ambient, inc/default_schema.hpp:357 */
  uint32_t width, height;

  /* triangles (loose triangles and mesh triangles alike), xyz interleaved, n_triangles*3 floats */
  uint64_t n_triangles;
  const float *tri_p1;
  const float *tri_p2;
  const float *tri_p3;
  const uint32_t *tri_object;    /* n_triangles */

  uint64_t n_spheres;
  const float *sph_center;       /* n_spheres*3 */
  const float *sph_radius;       /* n_spheres   */
  const uint32_t *sph_object;    /* n_spheres   */

  uint64_t n_planes;
  const float *pl_point;         /* n_planes*3 */
  const float *pl_normal;        /* n_planes*3, NOT normalised (reference keeps the JSON vector) */
  const uint32_t *pl_object;     /* n_planes   */

  uint32_t n_objects;
  const uint32_t *obj_material;  /* n_objects: material index of each object */
  const uint32_t *obj_kind;      /* optional (may be NULL), n_objects: CUTRACE_OBJ_*; used by the
                                    scene dump and by the oracle (a mesh gets the reference's AABB
                                    pre-test, a loose triangle does not). NULL = infer: an object
                                    with exactly one triangle is a loose triangle. */

  /* phong_material, inc/default_schema.hpp:319-343 */
  uint32_t n_materials;
  const float *mat_color;        /* n_materials*3 */
  const float *mat_specular;     /* n_materials */
  const float *mat_reflect;      /* n_materials */
  const float *mat_phong;        /* n_materials */
  const float *mat_transparency; /* n_materials */

  /* lights, inc/default_schema.hpp:267-311 */
  uint32_t n_lights;
  const uint32_t *light_kind;    /* CUTRACE_LIGHT_* */
  const float *light_vec;        /* n_lights*3: sun direction or point position */
  const float *light_color;      /* n_lights*3 */
} cutrace_scene_desc;

/* cutrace_opts.flags */
#define CUTRACE_FLAG_NO_SMEM_TOP 1u    /* do not stage the top of the BVH in shared memory */
#define CUTRACE_FLAG_VALIDATE_BVH 2u   /* run the device-side BVH validator after the build */
#define CUTRACE_FLAG_BRUTE_FORCE 4u    /* debug: skip the BVH, test every primitive per ray */
#define CUTRACE_FLAG_SERIALIZE 8u      /* run every kernel of a frame on one stream (per-kernel trace_ms / shade_ms are
                                          only measured in this mode); default: shade kernels overlap the trace chain */

/* Scheduler of a frame.  Default (none of the three flags): the persistent per-pixel kernel — ONE launch per frame, a thread
 * walks a pixel's whole path (primary ray, shadow rays, reflection / transmission chain) with an explicit stack, warps claim
 * pixels from a global cursor.  Measured on B200 (profiles/r02_tuning.md) it matches or beats the wavefront on all five
 * BASELINE configs (bunny.json 4K 9.0 against 9.8 ms; a 1/8 tile shard 1.36 against 1.60 ms) because nothing has to be
 * regrouped there and the queues cost 9 GB of DRAM traffic per frame.  The two wavefront schedulers stay selectable: */
#define CUTRACE_FLAG_FRAME_KERNEL 16u  /* wavefront, ONE cooperative launch: level loop and phase barriers on the device */
#define CUTRACE_FLAG_LAUNCHES 32u      /* wavefront, one launch per bounce level and kind, replayed as a CUDA graph */
#define CUTRACE_FLAG_PIXEL_KERNEL 64u  /* the per-pixel kernel (the default) */
/* Build quality.  Default: after the LBVH every subtree of at most 1024 primitives is rebuilt top-down with sweep SAH (one CTA
 * per treelet): frames get 10-13 % faster (bunny.json 4K 9.04 -> 8.17 ms, the 10 M-triangle hall 81.8 -> 70.9 ms) for
 * 1.7 ns/primitive of extra build time (bunny.json +0.23 ms, the hall +17 ms).  A caller that renders ONE frame per upload of
 * a scene with few pixels per primitive (the hall: 3 pixels per triangle) sets FAST_BUILD and keeps the LBVH topology. */
#define CUTRACE_FLAG_FAST_BUILD 128u

typedef struct cutrace_opts {
  float fudge;          /* min hit distance; the reference passes 1e-3 (main.cu:30)            */
  uint32_t bounces;     /* recursion budget; the reference passes 5 (main.cu:30); max 15       */
  int32_t device;       /* CUDA ordinal, -1 = current device                                   */
  uint32_t flags;
  /* screen-space sharding (multi-GPU): this ctx renders tiles t with t % tile_world == tile_rank.
   * tile_world = 0 or 1 means the whole frame. Tiles are CUTRACE_TILE x CUTRACE_TILE pixels in
   * row-major tile order. */
  uint32_t tile_rank, tile_world;
  void *stream;         /* cudaStream_t to launch on; NULL = a stream owned by the ctx         */
  uint32_t leaf_size;   /* max primitives per BVH leaf (1..8); 0 = default (3)                 */
  uint32_t reserved[7];
} cutrace_opts;

#define CUTRACE_TILE_SHIFT 4u                 /* log2 of the tile edge */
#define CUTRACE_TILE (1u << CUTRACE_TILE_SHIFT) /* 16 x 16 pixels: fine enough to balance 8 ranks (profiles/r01_tuning.md) */
#define CUTRACE_TILE_PIXELS (CUTRACE_TILE * CUTRACE_TILE)

typedef struct cutrace_stats {
  float build_ms;       /* LBVH build inside cutrace_upload_scene (device time)               */
  float render_ms;      /* device time of the last cutrace_render                              */
  float gather_ms;      /* device time of cutrace_gather_* in the last frame (0 if unused)     */
  float max_depth;      /* largest finite depth of the local pixels, 0 if none (kernel.hpp:120-125) */
  uint64_t rays_primary;
  uint64_t rays_reflect;
  uint64_t rays_transmit;
  uint64_t rays_shadow;     /* one per light per shaded hit                                    */
  uint64_t shadow_casts;    /* closest-hit casts issued by shadow marches (>= rays_shadow when
                               translucent surfaces are crossed, inc/shading.hpp:32)           */
  uint64_t local_pixels;    /* pixels rendered by this ctx                                     */
  uint32_t kernel_launches; /* kernels launched by the last cutrace_render                     */
  uint32_t bvh_nodes;       /* live internal nodes                                             */
  uint32_t bvh_depth;       /* deepest leaf                                                    */
  uint32_t smem_nodes;      /* BVH nodes staged in shared memory                               */
  float trace_ms;           /* device time in closest-hit kernels (sum over bounce levels)     */
  float shade_ms;           /* device time in shadow+phong kernels                             */
  uint32_t scheduler;       /* how the last frame ran: 0 = one launch per level and kind, 1 = the persistent frame kernel,
                               2 = the per-pixel kernel */
  uint32_t reserved[5];
} cutrace_stats;

typedef struct cutrace_ctx cutrace_ctx;

/* fills o with the reference's defaults: fudge 1e-3, bounces 5, device -1, whole frame */
void cutrace_default_opts(cutrace_opts *o);

/* Replaces default_to_gpu() (inc/default_schema.hpp:935, inc/cpu_to_gpu.hpp:188-198): copies the
 * scene to the device as flat SoA and builds the LBVH.  `opts` may be NULL (defaults). */
int cutrace_upload_scene(const cutrace_scene_desc *scene, const cutrace_opts *opts, cutrace_ctx **out);

/* Replaces the launch+sync in gpu::render (inc/kernel.hpp:103-108).  Blocks until the frame is
 * complete on the device.  `stats` may be NULL.  Re-runnable. */
int cutrace_render(cutrace_ctx *ctx, cutrace_stats *stats);

/* Replaces the D2H copies and the max-depth scan of gpu::render (inc/kernel.hpp:110-125).
 * Caller-owned host buffers of width*height (depth, hit_id) and width*height*3 (normal, colour)
 * elements; any of them may be NULL.  For a sharded ctx only the local tiles are written. */
int cutrace_download(cutrace_ctx *ctx, float *depth, float *normal, float *color, uint32_t *hit_id,
                     float *max_depth);

/* cutrace_render + cutrace_download in one call — the shape of the reference operator, which renders AND fills the
 * host images (inc/kernel.hpp:86-126).  The depth / normal / id images are final after the primary rays, so their
 * device->host copies run on a copy stream underneath the remaining bounce levels; only the colour copy is left for
 * the end.  Same arguments as cutrace_download (+ stats); host buffers should be pinned (cutrace_host_alloc) for the
 * copies to overlap.  Falls back to render-then-download for sharded ctxs. */
int cutrace_render_download(cutrace_ctx *ctx, float *depth, float *normal, float *color, uint32_t *hit_id,
                            float *max_depth, cutrace_stats *stats);

/* Output stage fused on the device (inc/images.hpp:26-88 + main.cu:34-36): the three 8-bit RGB images
 * the reference hands to stbi_write_jpg — depth (nearest = brightest, relative to max_depth), normal
 * (0.5 + 0.5 n), colour (clamped) — 9 bytes per pixel over PCIe instead of 28.  Each buffer is
 * width*height*3 bytes and may be NULL. */
int cutrace_download_bytes(cutrace_ctx *ctx, uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb,
                           float *max_depth);

void cutrace_free(cutrace_ctx *ctx);

/* thread-local, never NULL */
const char *cutrace_last_error(void);

/* ---- helpers around the three calls above ------------------------------------------------- */

/* new camera / resolution for an uploaded scene (no rebuild).  cam_* as in cutrace_scene_desc. */
int cutrace_set_camera(cutrace_ctx *ctx, const float pos[3], const float up[3], const float forward[3],
                       const float right[3], float ambient, uint32_t width, uint32_t height);

/* stats of the last upload/render */
int cutrace_get_stats(cutrace_ctx *ctx, cutrace_stats *stats);

/* Diagnostics of the last frame when it ran as one persistent kernel (the default): out[p] = milliseconds from the start
 * of the kernel until every ray of bounce level p was traced (p < levels), out[levels] = until the frame was assembled.
 * Taken from %globaltimer on the device.  *n_out = number of valid entries (0 when the frame ran as separate launches,
 * CUTRACE_FLAG_SERIALIZE).  The reference has no counterpart (one opaque kernel, inc/kernel.hpp:103-108). */
int cutrace_get_phase_ms(cutrace_ctx *ctx, float *out, uint32_t capacity, uint32_t *n_out);

/* Device-resident results of the last render, tile-major local layout: pixel j of local tile i is
 * at (i*CUTRACE_TILE_PIXELS + j).  Pointers stay valid until the next set_camera/free.
 * Used by the multi-GPU gather (NCCL over the caller's communicator, or peer copies). */
int cutrace_device_buffers(cutrace_ctx *ctx, float **depth, float **normal, float **color,
                           uint32_t **hit_id, uint64_t *n_local_px_padded);

/* Row-major device images of this ctx's own frame (one-GPU ctx, or the exporting ctx of a sharded render): what
 * cutrace_download copies.  Valid until the next set_camera/free. */
int cutrace_frame_device(cutrace_ctx *ctx, float **depth, float **normal, float **color, uint32_t **hit_id);

/* Multi-GPU without a gather: the ctx that owns the final frame (normally tile_rank 0) exports it through CUDA IPC;
 * the other ranks' ctxs import the handle and from then on their kernels store the G-buffer and the final colour of
 * their tiles straight into that frame over NVLink (peer stores fused into the producing kernels).  The caller only
 * has to synchronise the ranks (a barrier / the max-depth all-reduce) before downloading from the exporting ctx, and
 * once more before the next frame is rendered if it read the frame in between (the stores of frame N+1 must not
 * overtake a reader of frame N).
 * `handle` is CUTRACE_IPC_HANDLE_BYTES bytes, to be shipped to the other processes by the caller: the CUDA IPC handle
 * followed by the frame's width and height, which the importing ctx checks against its own (a mismatch would be an
 * out-of-bounds store into another GPU's memory).  Re-export after cutrace_set_camera changes the resolution. */
#define CUTRACE_IPC_HANDLE_BYTES 80
int cutrace_frame_ipc_export(cutrace_ctx *ctx, void *handle);
int cutrace_frame_ipc_import(cutrace_ctx *ctx, const void *handle);
/* Same idea for any frame block the caller can name by address: `frame_block` is 32*width*height bytes laid out as
 * depth n | normal 3n | colour 3n | hit id n (n = width*height), row-major — e.g.
 *   - the depth pointer cutrace_frame_device returned for the owning ctx of the same process (several ctxs / GPUs driven
 *     by one host, peer access enabled by the caller), or
 *   - PINNED HOST memory registered with cutrace_host_register (or from cutrace_host_alloc): every rank's kernels then
 *     store their tiles straight into the caller's host frame over their own PCIe link — the multi-GPU download without
 *     funnelling the frame through one GPU; a shared-memory mapping registered by every process works across processes.
 * width/height are the dimensions of the block and must equal the ctx's.  NULL detaches. */
int cutrace_frame_attach(cutrace_ctx *ctx, void *frame_block, uint32_t width, uint32_t height);
/* Device memory of every ctx comes from a memory pool this library owns (one per device; the device's default pool is not
 * touched), which keeps freed blocks for the next upload.  Hands the cached blocks of `device` (-1: current) back to the driver. */
int cutrace_trim_memory(int device);
/* page-locks + maps caller memory (e.g. a shared-memory frame) for cutrace_frame_attach / fast copies; undone by unregister */
int cutrace_host_register(void *ptr, size_t bytes);
int cutrace_host_unregister(void *ptr);
/* helpers for a one-process multi-GPU host: peer access from `device` to `peer_device` (idempotent), and the frame-wide
 * max depth (max over the ranks' cutrace_stats.max_depth, kernel.hpp:120-125) that cutrace_download_bytes of the ctx
 * owning the frame should use for the depth image. */
int cutrace_enable_peer_access(int device, int peer_device);
int cutrace_set_frame_max_depth(cutrace_ctx *ctx, float max_depth);

/* Un-tiles `world` gathered rank buffers (each laid out as cutrace_device_buffers describes, rank r
 * at gathered + r*stride elements of the respective type) into row-major full-frame DEVICE images
 * on the ctx's device.  Any of the pointers may be NULL. */
int cutrace_untile_device(cutrace_ctx *ctx, uint32_t world,
                          const float *g_depth, const float *g_normal, const float *g_color,
                          const uint32_t *g_id, uint64_t stride_px,
                          float *depth, float *normal, float *color, uint32_t *hit_id);

/* Output stage on the device (inc/images.hpp:26-88): maps full-frame row-major device images to
 * the three 8-bit RGB images the reference hands to stbi_write_jpg. Any output may be NULL. */
int cutrace_encode_bytes_device(cutrace_ctx *ctx, const float *depth, const float *normal,
                                const float *color, float max_depth, uint64_t n_px,
                                uint8_t *depth_rgb, uint8_t *normal_rgb, uint8_t *color_rgb);

/* pinned host memory for the download buffers (optional; any host memory works) */
void *cutrace_host_alloc(size_t bytes);
void cutrace_host_free(void *p);

/* device-side BVH self check (parent boxes contain children, every primitive reachable exactly
 * once); returns CUTRACE_OK or CUTRACE_ERR_INTERNAL with a message. */
int cutrace_validate_bvh(cutrace_ctx *ctx);

/* test hook: the LBVH builder's own stable LSD radix sort on host arrays of n (key, value) pairs,
 * sorted in place by key (device round trip inside). */
int cutrace_debug_radix_sort(uint64_t *keys, uint32_t *values, uint32_t n, int device);

/* test hook: out_powf[i] = powf(x[i], e[i]) and out_fast[i] = the specular power as the shading code computes it (powf skipped
 * below the underflow floor of e[i]), both on the device; the two arrays must be bit-identical. */
int cutrace_debug_phong_pow(const float *x, const float *e, float *out_powf, float *out_fast, uint32_t n, int device);

/* test hook (host only, no device needed): the screen tile (tx, ty) that tile slot `slot` of a width x height frame shows for
 * tile_world ranks, as the kernels compute it (curve: 1 = super-tile order, 0 = round 1's multiplicative scatter), and the inverse.
 * cutrace_debug_tile_of_slot returns 0 for a padding slot (slot >= number of tiles), 1 otherwise. */
int cutrace_debug_tile_of_slot(uint32_t width, uint32_t height, uint32_t tile_world, uint32_t curve, uint32_t slot, uint32_t *tx, uint32_t *ty);
uint32_t cutrace_debug_slot_of_tile(uint32_t width, uint32_t height, uint32_t tile_world, uint32_t curve, uint32_t tx, uint32_t ty);

/* test hook (host only): the pixel kernel's per-CTA work cursors.  cutrace_debug_segment_length = number of work items CTA k of a
 * grid of G CTAs owns in a frame of n_work items; cutrace_debug_segment_work = the frame work item behind offset o of CTA k's own
 * index space (o < that length). */
uint32_t cutrace_debug_segment_length(uint32_t n_work, uint32_t k, uint32_t G);
uint32_t cutrace_debug_segment_work(uint32_t o, uint32_t k, uint32_t G);

uint32_t cutrace_abi_version(void);
/* edge of the screen tiles the sharding works in (CUTRACE_TILE of the library that is loaded) */
uint32_t cutrace_tile_size(void);

#ifdef __cplusplus
}
#endif
#endif /* CUTRACE_B200_CUTRACE_H */
