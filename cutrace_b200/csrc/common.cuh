// common.cuh — shared device types of the B200 render path (flat SoA scene store, LBVH, queues).
//
// Replaces the reference's array-of-tagged-unions scene (inc/gpu_variant.hpp, inc/gpu_array.hpp,
// inc/gpu_types.hpp:263-287) with 16-byte aligned records that are read with 128-bit loads.
#ifndef CUTRACE_B200_COMMON_CUH
#define CUTRACE_B200_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/cutrace.h"

namespace ctb {

// ---- float3 math with PINNED roundings (inc/vector.hpp) ------------------------------------------------------------------
// The reference is compiled with nvcc's default -fmad=true: every a*b+c in its headers may or may not become one FMA, and
// which ones do is the compiler's choice per compilation context.  Round 1 wrote the same expression trees and relied on nvcc
// contracting them the same way; that held for the per-level kernels (depth bit-identical to the reference's kernel on all
// four reference scenes), but the SAME source inlined into the frame kernel or the pixel kernel was contracted differently:
// depths off by 1-2 ulp on ~0.002 % of the pixels, a silhouette pixel of triangle.json flipping (profiles/r02_parity.md).
// So the roundings are now spelled out with __fmul_rn / __fmaf_rn / __fadd_rn, which neither nvcc nor ptxas re-fuses.  The
// chosen forms are the ones read from the SASS of the kernels that were verified bit-identical against the reference
// (profiles/r02_parity.md lists the instruction sequences):
//   x*y + z*w      ->  fma(x, y, round(z*w))           (first product fused, second rounded)
//   dot(a, b)      ->  fma(a.z, b.z, fma(a.x, b.x, round(a.y*b.y)))
//   cross          ->  fma(u, v, -round(w*q)) per component
//   o + t*d        ->  fma(d, t, o)
//   det3 (Sarrus)  ->  round(round(a*e)*i), then one fma(round(product), factor, acc) per remaining term, in source order
// Host code (scene set-up) keeps the plain expressions.
struct vec3 { float x, y, z; };

__host__ __device__ __forceinline__ vec3 mk3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
#ifdef __CUDA_ARCH__
#define CTB_MUL(a, b) __fmul_rn((a), (b))
#define CTB_ADD(a, b) __fadd_rn((a), (b))
#define CTB_SUB(a, b) __fsub_rn((a), (b))
#define CTB_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define CTB_DIV(a, b) __fdiv_rn((a), (b))
#define CTB_SQRT(a) __fsqrt_rn((a))
#else
#define CTB_MUL(a, b) ((a) * (b))
#define CTB_ADD(a, b) ((a) + (b))
#define CTB_SUB(a, b) ((a) - (b))
#define CTB_FMA(a, b, c) fmaf((a), (b), (c))
#define CTB_DIV(a, b) ((a) / (b))
#define CTB_SQRT(a) sqrtf((a))
#endif
__host__ __device__ __forceinline__ vec3 vadd(vec3 a, vec3 b) { return mk3(CTB_ADD(a.x, b.x), CTB_ADD(a.y, b.y), CTB_ADD(a.z, b.z)); }   // :100
__host__ __device__ __forceinline__ vec3 vsub(vec3 a, vec3 b) { return mk3(CTB_SUB(a.x, b.x), CTB_SUB(a.y, b.y), CTB_SUB(a.z, b.z)); }   // :109
__host__ __device__ __forceinline__ vec3 vscale(vec3 a, float f) { return mk3(CTB_MUL(f, a.x), CTB_MUL(f, a.y), CTB_MUL(f, a.z)); }      // :118
__host__ __device__ __forceinline__ vec3 vmul(vec3 a, vec3 b) { return mk3(CTB_MUL(a.x, b.x), CTB_MUL(a.y, b.y), CTB_MUL(a.z, b.z)); }   // :136
// o + t * d  (e.g. *hit = r->start + *dist * r->dir, inc/default_schema.hpp:71)
__host__ __device__ __forceinline__ vec3 vmad(vec3 o, vec3 d, float t) { return mk3(CTB_FMA(d.x, t, o.x), CTB_FMA(d.y, t, o.y), CTB_FMA(d.z, t, o.z)); }
__host__ __device__ __forceinline__ float vdot(vec3 a, vec3 b) { return CTB_FMA(a.z, b.z, CTB_FMA(a.x, b.x, CTB_MUL(a.y, b.y))); }       // :127
__host__ __device__ __forceinline__ vec3 vcross(vec3 a, vec3 o) {                                                                       // :65-71
  return mk3(CTB_FMA(a.y, o.z, -CTB_MUL(a.z, o.y)), CTB_FMA(a.z, o.x, -CTB_MUL(a.x, o.z)), CTB_FMA(a.x, o.y, -CTB_MUL(a.y, o.x)));
}
// The specular term pow(max(0, n.h), e) (inc/shading.hpp:91): below x0 = 2^(-152 / e) the true power is under 2^-152, less than half
// the smallest denormal, and powf returns +0 (it is accurate to a few ulp and correctly rounded-to-zero there;
// tests/test_gpu_parity.py::test_phong_pow_floor checks the device function against this claim).  0.98 covers the 2-ulp MUFU.EX2
// behind exp2f: (0.98 x0)^e <= 0.98 * 2^-152.  Phong highlights are narrow (e = 200 / 500 in bunny.json: x0 = 0.59 / 0.81), so most
// warps skip the ~70-instruction powf — it was 4.9 % of the frame's instructions.  e < 1 or not finite: no floor, powf always runs.
__device__ __forceinline__ float phong_pow_floor(float e) { return (e >= 1.0f && e <= 3.0e38f) ? 0.98f * exp2f(-152.0f / e) : -1.0f; }
__device__ __forceinline__ float phong_pow(float x, float e, float floor_x) { return x < floor_x ? 0.0f : powf(x, e); }
__host__ __device__ __forceinline__ float vnorm(vec3 a) { return CTB_SQRT(vdot(a, a)); }                                                // :85-92
__host__ __device__ __forceinline__ vec3 vnormalized(vec3 a) { return vscale(a, CTB_DIV(1.0f, vnorm(a))); }                              // :77-79
__host__ __device__ __forceinline__ vec3 vreflect(vec3 incoming, vec3 normal) {                                                         // :204-206
  const float s = CTB_MUL(2.0f, vdot(normal, incoming));
  return mk3(CTB_FMA(-s, normal.x, incoming.x), CTB_FMA(-s, normal.y, incoming.y), CTB_FMA(-s, normal.z, incoming.z));
}
// matrix::determinant (Sarrus), inc/vector.hpp:218-224, columns c0 c1 c2:  a*e*i + b*f*g + c*d*h - c*e*g - a*f*h - b*d*i
__host__ __device__ __forceinline__ float det3(vec3 c0, vec3 c1, vec3 c2) {
  const float a = c0.x, b = c1.x, c = c2.x, d = c0.y, e = c1.y, f = c2.y, g = c0.z, h = c1.z, i = c2.z;
  float t = CTB_MUL(CTB_MUL(a, e), i);
  t = CTB_FMA(CTB_MUL(b, f), g, t);
  t = CTB_FMA(CTB_MUL(c, d), h, t);
  t = CTB_FMA(-CTB_MUL(c, e), g, t);
  t = CTB_FMA(-CTB_MUL(a, f), h, t);
  t = CTB_FMA(-CTB_MUL(b, d), i, t);
  return t;
}
// The same determinant as nvcc contracts it for the NUMERATOR of t in triangle::intersect (inc/default_schema.hpp:61,67 — the
// one Sarrus sum whose first product is not shared with another determinant): the second term is the rounded one and the
// first is fused onto it, fma(a*e, i, round((b*f)*g)); the remaining terms as in det3.
__host__ __device__ __forceinline__ float det3_t(vec3 c0, vec3 c1, vec3 c2) {
  const float a = c0.x, b = c1.x, c = c2.x, d = c0.y, e = c1.y, f = c2.y, g = c0.z, h = c1.z, i = c2.z;
  float t = CTB_MUL(CTB_MUL(b, f), g);
  t = CTB_FMA(CTB_MUL(a, e), i, t);
  t = CTB_FMA(CTB_MUL(c, d), h, t);
  t = CTB_FMA(-CTB_MUL(c, e), g, t);
  t = CTB_FMA(-CTB_MUL(a, f), h, t);
  t = CTB_FMA(-CTB_MUL(b, d), i, t);
  return t;
}

// Plane side pre-pass for shadow rays (trace.cuh: plane_side_prepass): one (point - X).n per plane and a sign comparison per
// light instead of a plane test per (light, plane).  Same images, but MEASURED SLOWER on B200 — bunny.json 4K 9.04 -> 9.94 ms
// (pixel kernel), 9.86 -> 10.79 (wavefront), hall 81.9 -> 84.1 (profiles/r02_tuning.md): the 20 dependent table loads per
// shaded hit cost more than the 20 short plane tests they replace.  Off; -DCTB_PLANE_PREPASS=1 builds it.
#ifndef CTB_PLANE_PREPASS
#define CTB_PLANE_PREPASS 0
#endif

// ---- scene records ------------------------------------------------------------------------------
// One 48-byte record per BVH primitive, stored in BVH-leaf (Morton) order: 3 x LDG.128 / LDS.128.
// Triangle: p1,p2,p3 = vertices exactly as uploaded (the reference's gpu::schema::triangle is also
// 48 B, inc/default_schema.hpp:26-30).  Sphere: p1 = centre, p2.x = radius.
// (obj, idx) implements the reference's tie-break: lowest object index wins an equal-t tie
// (inc/ray_cast.hpp:43), then lowest triangle index in file order (inc/default_schema.hpp:134).
struct __align__(16) PrimRec {
  float p1x, p1y, p1z; uint32_t obj;
  float p2x, p2y, p2z; uint32_t idx;
  float p3x, p3y, p3z; uint32_t kind;   // 0 triangle, 1 sphere
};
static_assert(sizeof(PrimRec) == 48, "PrimRec must be 48 bytes");
#define CTB_PRIM_TRI 0u
#define CTB_PRIM_SPHERE 1u

// 64-byte binary BVH node holding BOTH children's boxes (4 x 128-bit loads per visit).
//   n0xy = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1xy likewise for child 1
//   nz   = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   meta = (c0, c1, -, -)   child >= 0: node index; child < 0: leaf, ~child = (first << 3) | (count-1)
struct __align__(16) Node {
  float4 n0xy, n1xy, nz;
  int4 meta;
};
static_assert(sizeof(Node) == 64, "Node must be 64 bytes");

// -DCTB_BVH4=1: the kernels walk a 4-wide tree instead — every second level of the binary LBVH collapsed into its parent
// (bvh_build.cu: collapse4).  128-byte node, structure of arrays over the four children so that one float4 holds one box plane
// of all of them; an empty slot has an inverted box (never hit) and the reference CTB_SENTINEL.
struct __align__(16) Node4 {
  float4 lox, hix, loy, hiy, loz, hiz;
  int4 ref;     // child >= 0: Node4 index; child < 0: leaf code (as in Node); CTB_SENTINEL: empty slot
  int4 pad;
};
static_assert(sizeof(Node4) == 128, "Node4 must be 128 bytes");
#ifndef CTB_BVH4
#define CTB_BVH4 0
#endif
#define CTB_NODE_F4 (CTB_BVH4 ? 8u : 4u)          // float4 words per node of the tree the kernels walk
#define CTB_NODE_BYTES (16u * CTB_NODE_F4)
#define CTB_SENTINEL 0x7fffffff
#define CTB_STACK 64
#define CTB_MAX_LEAF 8
#ifndef CTB_SMEM_TOP_NODES
#define CTB_SMEM_TOP_NODES 0      // nodes of the BVH top (breadth-first) staged in shared memory for scenes that do not fit;
                                  // measured slower than L1 on B200 (profiles/r01_tuning.md), so off unless CUTRACE_SMEM_TOP_NODES is set
#endif

__host__ __device__ __forceinline__ int leaf_encode(uint32_t first, uint32_t count) { return ~(int)((first << 3) | (count - 1u)); }
__host__ __device__ __forceinline__ uint32_t leaf_first(int c) { return ((uint32_t)~c) >> 3; }
__host__ __device__ __forceinline__ uint32_t leaf_count(int c) { return (((uint32_t)~c) & 7u) + 1u; }

struct __align__(16) PlaneRec {   // plane::intersect, inc/default_schema.hpp:159-207
  float px, py, pz; uint32_t obj;
  float nx, ny, nz; uint32_t pad;
};
struct __align__(16) MaterialRec {  // phong_material, inc/default_schema.hpp:319-343
  float r, g, b, specular;
  float reflect, phong, transparency; uint32_t pad;
};
struct __align__(16) LightRec {     // sun / point_light, inc/default_schema.hpp:267-311
  float vx, vy, vz; uint32_t kind;
  float r, g, b; uint32_t pad;
};

struct Camera {   // cam, inc/default_schema.hpp:350-396 (post look_at)
  vec3 pos, up, forward, right;
  float ambient;
  uint32_t w, h;
};

// ---- wavefront queue records ---------------------------------------------------------------------
struct __align__(16) RayRec {       // 32 B: secondary ray (reflection / transmission)
  float ox, oy, oz; uint32_t pix;
  float dx, dy, dz; float weight;
};
struct __align__(16) ShadeRec {     // 48 B: one shaded hit = n_lights shadow rays + Phong
  float hx, hy, hz; uint32_t pix;   // hit point as the primitive reports it (shadow-ray origin)
  float nx, ny, nz; uint32_t mat;   // raw normal (normalised inside phong, inc/shading.hpp:83)
  float ix, iy, iz; float weight;   // incoming ray direction; path weight of this hit's Phong term
};
static_assert(sizeof(RayRec) == 32 && sizeof(ShadeRec) == 48, "queue record sizes");

// device-side counters of one frame
// Words that many warps hit with atomics at the same time each sit on their own 128-byte line: same-address atomics
// retire at ~0.67 ns each on B200 whatever the SM count, and atomics to ONE line from every warp of the grid (cursor,
// queue tails, barrier) had made a nearly empty bounce level cost 40 us (profiles/r02_tuning.md).
#define CTB_MAX_SEGS 320   // CTAs of the pixel kernel's persistent grid that can have a work cursor of their own (148 SMs x 1 or 2)
struct __align__(128) HotWord { unsigned int v; unsigned int pad_[31]; };
struct FrameStats {               // what the host reads after a frame
  unsigned long long rays_reflect, rays_transmit, shadow_casts, shade_records;
  unsigned int max_depth_bits;    // float bits of the largest finite primary depth
  unsigned int overflow;          // a queue reservation did not fit: the emission was dropped, the frame is reported as failed
  unsigned long long phase_ns[19];   // frame kernel: %globaltimer when phase p opened; [17] start, [levels] end of the frame
};
struct FrameCounters {
  HotWord n_rays[18];       // ray-queue slots reserved for level L (valid rays + retired holes)
  HotWord n_shade[18];      // shade-queue slots reserved by level L (valid records + retired holes)
  HotWord work_trace[18];   // work-stealing cursors
  HotWord work_shade[18];
  HotWord arrive[19];       // frame kernel: CTAs that have finished trace(p)   (arrive[levels]: all shading done)
  HotWord work_export;      // cursor of the G-buffer export (peer / host frame)
  HotWord finished;         // CTAs that have left the frame kernel: the last one publishes the stats and clears everything
  HotWord seg[CTB_MAX_SEGS];   // pixel kernel: one work cursor per CTA (render.cu: claim_segment)
  FrameStats st;
};

// Index arithmetic of the pixel kernel's per-CTA work cursors (render.cu: claim_segment; host-callable for tests/test_abi.py):
// the frame's work items in chunks of CTB_SEG_CHUNK, chunk j owned by CTA j % G; `o` is an offset in a CTA's own index space.
#define CTB_SEG_CHUNK 4096u
#define CTB_SEG_MIN_PX (1u << 20)   // frames below this keep the single cursor (see plan_launch)
__host__ __device__ __forceinline__ unsigned seg_len(unsigned n_work, unsigned k, unsigned G) {
  const unsigned nc = (n_work + CTB_SEG_CHUNK - 1u) / CTB_SEG_CHUNK;   // chunks of the frame (the last one may be partial)
  if (k >= nc) return 0u;
  const unsigned mine = (nc - 1u - k) / G + 1u;
  const unsigned len = mine * CTB_SEG_CHUNK;
  return ((nc - 1u) % G == k) ? len - (nc * CTB_SEG_CHUNK - n_work) : len;   // the owner of the last chunk
}
__host__ __device__ __forceinline__ unsigned seg_to_work(unsigned o, unsigned k, unsigned G) {
  return ((o / CTB_SEG_CHUNK) * G + k) * CTB_SEG_CHUNK + (o % CTB_SEG_CHUNK);
}

// Per-object data for the reference's mesh pre-test (inc/default_schema.hpp:99-114): the AABB cutrace computes on the host
// for every mesh (inc/default_schema.hpp:573-586) and whether the object is a mesh at all.
struct __align__(16) ObjBound {
  float lo[3]; uint32_t is_mesh;
  float hi[3]; uint32_t pad;
};

// everything a render kernel needs, passed by value (lives in the constant bank)
struct SceneView {
  const Node *nodes;
  const PrimRec *prims;
  const PlaneRec *planes;
  const MaterialRec *materials;
  const LightRec *lights;
  const uint32_t *obj_material;
  const ObjBound *obj_bounds;   // n_objects
  const float *pl_tbl;          // n_planes x n_lights: side of every light relative to every plane (trace.cuh: plane_side_prepass), or NULL
  const float *pl_eps;          // n_planes x 2: |sX| above eps[0] + eps[1] * |X|_1 is "clearly off the plane"
  uint32_t n_prims, n_nodes, n_planes, n_lights, n_materials, n_objects;
  int root;                 // node index, leaf code, or CTB_SENTINEL (empty BVH)
  uint32_t smem_nodes;      // nodes [0, smem_nodes) are staged in shared memory
  uint32_t smem_prims;      // prims [0, smem_prims) are staged in shared memory
  uint32_t all_opaque;      // no material has transparency >= 1e-6 (shadow rays can be any-hit)
  uint32_t brute_force;
  float fudge;
  float scene_mag;          // largest |coordinate| of the BVH primitives (box inflation is 2e-6 * scene_mag)
  Camera cam;
};

// Screen-space tiling.  The frame is cut into CUTRACE_TILE x CUTRACE_TILE tiles; tile "slots" 0..n_tiles-1 are dealt round-robin to the
// ranks (slot s belongs to rank s % world, local tile s / world).  Which screen tile a slot shows:
//   curve = 1 (default)  slots walk the frame super-tile by super-tile (CTB_SUPER_W x CTB_SUPER_H tiles, row-major inside and across):
//             tiles that are processed at the same time — by the warps of one CTA, and by one rank — lie next to each other on the
//             screen and walk the same part of the BVH.  The super-tile is 9 tiles wide so that inside it slot % world runs along
//             DIAGONALS for world = 2, 4, 8 ((iy * 9 + ix) % 8 = (ix + iy) % 8): every rank samples the whole image at tile
//             granularity (plain interleaving of rows gave a rank vertical stripes, and the 8-way shards of bunny.json differed by
//             20 % in cost, profiles/r01_tuning.md), and a rank's consecutive local tiles are at most a few tiles apart.
//   curve = 0            round 1's scatter: slot s shows screen tile (s * perm_a) % n_tiles, perm_a coprime to n_tiles, ~0.618 n
//             (one rank: perm_a = 1, natural order) — balanced, but a rank's consecutive tiles are 0.6 frames apart.
#define CTB_SUPER_W 9u
#define CTB_SUPER_H 8u
struct TileMap {
  uint32_t width, height;
  uint32_t tiles_x, tiles_y;
  uint32_t rank, world;       // this ctx owns slots s with s % world == rank
  uint32_t n_local_tiles;     // ceil(n_tiles / world): the same padded count on every rank
  uint32_t n_tiles, perm_a, perm_ainv;
  uint32_t wide_warps;        // 0: the 32 pixels of a warp form an 8 x 4 block (most coherent primary rays); 1: a 16 x 2 block — whole
                              // tile rows, so that per-pixel stores into a remote frame (peer GPU, pinned host memory) are 64 / 192-byte segments
  uint32_t curve;             // slot order, see above
};

// slot -> screen tile coordinates; false for the padding slots of the last local tile
__host__ __device__ __forceinline__ bool tile_of_slot(const TileMap &tm, uint32_t slot, uint32_t &tx, uint32_t &ty) {
  if (slot >= tm.n_tiles) return false;
  if (tm.curve) {
    const uint32_t row_tiles = tm.tiles_x * CTB_SUPER_H;                 // tiles of a full row of super-tiles
    const uint32_t sr = slot / row_tiles, rem = slot - sr * row_tiles;
    const uint32_t left_y = tm.tiles_y - sr * CTB_SUPER_H, h = left_y < CTB_SUPER_H ? left_y : CTB_SUPER_H;
    const uint32_t col_tiles = CTB_SUPER_W * h;                          // tiles of a full-width super-tile in this row
    const uint32_t sc = rem / col_tiles, rem2 = rem - sc * col_tiles;
    const uint32_t left_x = tm.tiles_x - sc * CTB_SUPER_W, w = left_x < CTB_SUPER_W ? left_x : CTB_SUPER_W;
    const uint32_t iy = rem2 / w, ix = rem2 - iy * w;
    tx = sc * CTB_SUPER_W + ix;
    ty = sr * CTB_SUPER_H + iy;
    return true;
  }
  const uint32_t t = (uint32_t)(((unsigned long long)slot * tm.perm_a) % tm.n_tiles);
  tx = t % tm.tiles_x;
  ty = t / tm.tiles_x;
  return true;
}
// screen tile index (ty * tiles_x + tx) -> slot
__host__ __device__ __forceinline__ uint32_t slot_of_tile(const TileMap &tm, uint32_t t) {
  if (tm.curve) {
    const uint32_t tx = t % tm.tiles_x, ty = t / tm.tiles_x;
    const uint32_t sr = ty / CTB_SUPER_H, sc = tx / CTB_SUPER_W;
    const uint32_t left_y = tm.tiles_y - sr * CTB_SUPER_H, h = left_y < CTB_SUPER_H ? left_y : CTB_SUPER_H;
    const uint32_t left_x = tm.tiles_x - sc * CTB_SUPER_W, w = left_x < CTB_SUPER_W ? left_x : CTB_SUPER_W;
    return sr * tm.tiles_x * CTB_SUPER_H + sc * CTB_SUPER_W * h + (ty - sr * CTB_SUPER_H) * w + (tx - sc * CTB_SUPER_W);
  }
  return (uint32_t)(((unsigned long long)t * tm.perm_ainv) % tm.n_tiles);
}
// local pixel index (tile-major, row-major inside the tile) -> pixel; false outside the image
__host__ __device__ __forceinline__ bool pixel_of_local(const TileMap &tm, uint32_t pix, uint32_t &x, uint32_t &y) {
  uint32_t tx, ty;
  if (!tile_of_slot(tm, (pix >> (2 * CUTRACE_TILE_SHIFT)) * tm.world + tm.rank, tx, ty)) return false;
  x = tx * CUTRACE_TILE + (pix & (CUTRACE_TILE - 1u));
  y = ty * CUTRACE_TILE + ((pix & (CUTRACE_TILE_PIXELS - 1u)) >> CUTRACE_TILE_SHIFT);
  return x < tm.width && y < tm.height;
}

}  // namespace ctb
#endif
