// trace.cuh — per-ray device functions: primitive tests and stack-based BVH traversal.
//
// The primitive tests restate the reference's arithmetic (same expression trees, so nvcc contracts
// them the way it contracts the reference) because parity on hit ids / depth is the first gate:
//   triangle  inc/default_schema.hpp:57-78   (Cramer's rule, two-sided, edge-inclusive)
//   sphere    inc/default_schema.hpp:226-251 (roots along the NORMALISED direction)
//   plane     inc/default_schema.hpp:189-201
// What is new is how few of them run: the reference tests every object and every triangle of a
// mesh whose AABB is hit (inc/ray_cast.hpp:37-52, inc/default_schema.hpp:133-141); here a ray walks
// a binary LBVH with an explicit stack and only tests the primitives of the leaves it reaches.
// The acceptance rule of ray_cast is kept: t > min_dist, strictly closer wins, equal-t ties go to the
// lowest object index and then the lowest triangle index in file order (inc/ray_cast.hpp:43,
// inc/default_schema.hpp:134).
//
// Code size is a first-class constraint here: the first version of these kernels was 150 KB of SASS
// and spent 14 of 15 stall cycles per issue in "no instruction" (profiles/r01_*).  Rare paths
// (spheres, IEEE-division slow paths, brute force, the translucent shadow march) are therefore
// kept out of line or in separate template instantiations, and leaf loops are not unrolled.
#ifndef CUTRACE_B200_TRACE_CUH
#define CUTRACE_B200_TRACE_CUH
#include "common.cuh"

#ifndef CTB_SPECULATIVE
#define CTB_SPECULATIVE 0
#endif
#ifndef CTB_PREFETCH_FAR
#define CTB_PREFETCH_FAR 1   // prefetch the pushed (far) child node into L1: -0.5..1 % on the 10 M-triangle hall
#endif

#ifndef CTB_BRANCHFREE_STACK
#define CTB_BRANCHFREE_STACK 1   // walks through L1 / L2 (MODE != 1): the node loops push speculatively (store always, move the pointer by "both children
                                 // hit") instead of branching three ways — hall 60.6 -> 60.0 ms; the staged scenes keep the branches (bunny.json 6.65 vs 6.69)
#endif
#ifndef CTB_PREFETCH_CHILDREN
#define CTB_PREFETCH_CHILDREN 0   // 1 / 2: when a node arrives, prefetch the records of BOTH its children into L1 (one / two 32-byte sectors each; leaves: their first primitive)
#endif

namespace ctb {

#define CTB_KIND_TRI 0
#define CTB_KIND_SPHERE 1
#define CTB_KIND_PLANE 2

struct Hit {
  float t;         // distance as ray_cast reports it (+INF: miss)
  uint32_t obj;    // object id (tie-break key 1)
  uint32_t idx;    // index in the primitive's input array (tie-break key 2)
  uint32_t ref;    // sorted primitive index, or plane index
  int kind;
};

__device__ __forceinline__ void hit_reset(Hit &h) {
  h.t = INFINITY; h.obj = 0xffffffffu; h.idx = 0xffffffffu; h.ref = 0; h.kind = -1;
}

__device__ __forceinline__ bool hit_better(const Hit &h, float t, uint32_t obj, uint32_t idx) {
  return t < h.t || (t == h.t && (obj < h.obj || (obj == h.obj && idx < h.idx)));
}

// triangle::intersect, inc/default_schema.hpp:57-69.  Returns the parametric distance via t.
// The three quotients are IEEE divisions of the reference's four Sarrus determinants, in its order.
// Before dividing, triangles whose barycentric numerators already have the wrong sign are rejected:
// fl(x/alpha) >= 0 can only hold for a negative real quotient if it underflows to -0, which the
// magnitude guard keeps on the exact path — so the accepted set and every accepted t are unchanged.
__device__ __forceinline__ bool tri_test(vec3 p1, vec3 p2, vec3 p3, vec3 o, vec3 dir, float min_t, float &t) {
  vec3 a = vsub(p2, p1), b = vsub(p2, p3), c = dir, d = vsub(p2, o);
  float alpha = det3(a, b, c);
  float nb = det3(d, b, c);
  float ng = det3(a, d, c);
  const float tiny = 1e-30f;
  bool neg_b = (__float_as_uint(nb) ^ __float_as_uint(alpha)) >> 31;   // quotient negative (or -0)
  bool neg_g = (__float_as_uint(ng) ^ __float_as_uint(alpha)) >> 31;
  float aa = CTB_MUL(tiny, fabsf(alpha));
  if ((neg_b && fabsf(nb) > aa) || (neg_g && fabsf(ng) > aa)) return false;
  float beta = CTB_DIV(nb, alpha);
  float gamma = CTB_DIV(ng, alpha);
  if (!(beta >= 0 && gamma >= 0 && CTB_ADD(beta, gamma) <= 1)) return false;
  float t0 = CTB_DIV(det3_t(a, b, d), alpha);
  t = t0;
  // `min_t <= t0` is the primitive's own test, `t0 > min_t` is ray_cast's `dist > min_dist`
  return isfinite(t0) && min_t <= t0 && t0 > min_t;
}

// sphere::intersect, inc/default_schema.hpp:226-243.  `d` is the NORMALISED direction (the reference normalises inside, :228); the
// brute-force loops over the <= 8 primitives of a scene like sphere_plane.json normalise once per ray instead of once per sphere.
// A negative (or NaN) discriminant leaves early: sqrt gives NaN, both roots are NaN, neither is finite — the reference returns
// false through the same values, only after two IEEE square roots (on their slow path: the argument is out of range) and two
// divisions.  Misses are the common case: in sphere_plane.json those four operations were a quarter of the frame's instructions.
__device__ __noinline__ bool sphere_test_n(float cx, float cy, float cz, float R, vec3 e, vec3 d, float min_t, float *tout) {
  vec3 c = mk3(cx, cy, cz);
  float dec = -vdot(d, vsub(e, c));
  // dec^2 - (d.d) * ((e-c).(e-c) - R^2), inc/default_schema.hpp:231
  float sub = CTB_FMA(dec, dec, -CTB_MUL(vdot(d, d), CTB_FMA(-R, R, vdot(vsub(e, c), vsub(e, c)))));
  if (!(sub >= 0.0f)) return false;
  float t0 = CTB_DIV(CTB_SUB(dec, CTB_SQRT(sub)), vdot(d, d)), t1 = CTB_DIV(CTB_ADD(dec, CTB_SQRT(sub)), vdot(d, d));
  bool t0v = isfinite(t0) && min_t <= t0, t1v = isfinite(t1) && min_t <= t1;
  if (!t0v && !t1v) return false;
  float t = (t0v && t1v) ? fminf(t0, t1) : (t0v ? t0 : t1);
  *tout = t;
  return t > min_t;
}
// The same test for the BVH walks, which meet a sphere rarely and keep nothing per ray: the direction is normalised here.  No early
// return on a negative discriminant in this copy: with it ptxas allocates the (sphere-free) bunny.json kernel differently —
// 8 more bytes of spills, 8.0 -> 8.2 ms per 4K frame — and the reference scenes have no BVH with spheres in it to pay that back.
__device__ __noinline__ bool sphere_test(float cx, float cy, float cz, float R, vec3 e, vec3 dir, float min_t, float *tout) {
  vec3 c = mk3(cx, cy, cz);
  vec3 d = vnormalized(dir);
  float dec = -vdot(d, vsub(e, c));
  float sub = CTB_FMA(dec, dec, -CTB_MUL(vdot(d, d), CTB_FMA(-R, R, vdot(vsub(e, c), vsub(e, c)))));
  float t0 = CTB_DIV(CTB_SUB(dec, CTB_SQRT(sub)), vdot(d, d)), t1 = CTB_DIV(CTB_ADD(dec, CTB_SQRT(sub)), vdot(d, d));
  bool t0v = isfinite(t0) && min_t <= t0, t1v = isfinite(t1) && min_t <= t1;
  if (!t0v && !t1v) return false;
  float t = (t0v && t1v) ? fminf(t0, t1) : (t0v ? t0 : t1);
  *tout = t;
  return t > min_t;
}

// plane::intersect, inc/default_schema.hpp:189-192
// A ray that starts ON the plane (a bounce or shadow ray leaving a wall) has a numerator of exactly 0 for axis-aligned planes: the
// quotient is +-0 or NaN, never accepted with min_t >= 0 — and a zero operand sends the IEEE division through its ~30-instruction
// slow path (2.3 % of the bunny.json frame, profiles/r02_tuning.md 8).  Left before dividing; same decisions.
__device__ __forceinline__ bool plane_test(vec3 point, vec3 n, vec3 o, vec3 dir, float min_t, float &t) {
  const float num = vdot(n, vsub(point, o));
  if (num == 0.0f && min_t >= 0.0f) { t = 0.0f; return false; }
  float t0 = CTB_DIV(num, vdot(dir, n));
  t = t0;
  return isfinite(t0) && min_t <= t0 && t0 > min_t;
}

// Slab test data of a ray.  Box distances are evaluated as fma(plane, inv, -o*inv): one FFMA per plane.  Its absolute
// error is <= 2^-24 (|o*inv| + |t|).  The |t| part is covered by the 4-ulp relative slack on t_far.  The |o*inv| part
// equals moving the box plane by |o| * 2^-24 in SPACE, whatever inv is — and the boxes are inflated by 2e-6 * (largest
// scene coordinate) at build time, which covers it for every ray that starts within 8x the scene's coordinate range
// (all secondary and shadow rays, any sane camera).  Only rays from further away get the explicit t-space bound
// `eabs` (which would otherwise let rays with a tiny direction component through every box).  Culling thus stays
// conservative: a box is never skipped when the reference's Cramer test could accept one of its triangles.  `inv`
// may be 1 ulp off (MUFU.RCP): that only perturbs the direction used for CULLING; primitive tests use exact o and d.
struct RayCtx {
  vec3 o, d;
  vec3 inv;   // ~1 / d, zeros replaced by +-1e-30
  vec3 oi;    // o * inv
  float eabs;
};

#ifndef CTB_RCP_APPROX
#define CTB_RCP_APPROX 1   // one MUFU.RCP (rcp.approx.ftz, 1 ulp) instead of __fdividef(1, x): bunny.json 4K 7.91 -> 7.83 ms; culling only, see RayCtx
#endif
__device__ __forceinline__ float safe_rcp(float x) {
  const float y = fabsf(x) < 1e-30f ? copysignf(1e-30f, x) : x;
#if CTB_RCP_APPROX
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
  return r;
#else
  return __fdividef(1.0f, y);
#endif
}

__device__ __forceinline__ void make_ray(RayCtx &r, vec3 o, vec3 d, float scene_mag) {
  r.o = o; r.d = d;
  r.inv = mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
  r.oi = mk3(o.x * r.inv.x, o.y * r.inv.y, o.z * r.inv.z);
  r.eabs = 0.f;
  if (fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)) > 8.0f * scene_mag) {   // far-away origin (distant camera)
    // an axis the ray does not move along (|d| < 1e-20) only decides by the SIGN of its huge slab distances
    const float ex = fabsf(d.x) < 1e-20f ? 0.f : fabsf(r.oi.x), ey = fabsf(d.y) < 1e-20f ? 0.f : fabsf(r.oi.y),
                ez = fabsf(d.z) < 1e-20f ? 0.f : fabsf(r.oi.z);
    r.eabs = 2.4e-7f * fmaxf(fmaxf(ex, ey), ez);
  }
}

// ---- the reference's per-mesh AABB pre-test -------------------------------------------------------------------------
// mesh::intersect (inc/default_schema.hpp:125-144) first runs bound_intersects (:99-114): a slab test of the ray against the
// AABB the host computed for the mesh (:573-586), tmin = 0, tmax = INF, device min/max (NaN-dropping), pass iff
// tmin <= tmax.  A ray that hits a triangle lying IN a face of that box (a flat mesh, an axis-aligned quad seen edge-on, a
// silhouette pixel exactly on the box edge) can fail this test by rounding although Cramer's rule accepts the triangle —
// the reference then misses the whole mesh.  The LBVH never looks at per-mesh boxes, so the test is replayed here for the
// only rays it can matter for: a triangle hit that is about to be ACCEPTED and whose hit point is not strictly inside its
// mesh's box.  `mesh_gate_exact` is the reference's arithmetic, operation for operation ((min - start) * (1 / dir) cannot be
// contracted into an FMA); the cheap inside-test in front of it is sound because a point inside the box by
// 1e-5 * (scene size + distance) keeps every slab interval open by 25x the rounding error of t1 / t2.
__device__ __noinline__ bool mesh_gate_exact(float lox, float loy, float loz, float hix, float hiy, float hiz, vec3 start, vec3 dir) {
  float tmin = 0.0f, tmax = INFINITY;
  const float r_inv[3] = {1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z};
  const float lo[3] = {lox, loy, loz}, hi[3] = {hix, hiy, hiz}, st[3] = {start.x, start.y, start.z};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float t1 = __fmul_rn(lo[d] - st[d], r_inv[d]);
    const float t2 = __fmul_rn(hi[d] - st[d], r_inv[d]);
    tmin = fminf(fmaxf(t1, tmin), fmaxf(t2, tmin));
    tmax = fmaxf(fminf(t1, tmax), fminf(t2, tmax));
  }
  return tmin <= tmax;
}
#ifndef CTB_MESH_GATE
#define CTB_MESH_GATE 1   // 0: tuning builds only (measures what the replayed pre-test costs)
#endif
__device__ __forceinline__ bool mesh_gate(const ObjBound *__restrict__ ob, uint32_t obj, vec3 o, vec3 d, float t, float scene_mag) {
  if (!CTB_MESH_GATE) return true;
  const float4 *bp = reinterpret_cast<const float4 *>(ob + obj);
  const float4 lo = __ldg(bp), hi = __ldg(bp + 1);
  if (__float_as_uint(lo.w) == 0u) return true;   // a loose triangle object: the reference has no pre-test for it
  const float px = fmaf(t, d.x, o.x), py = fmaf(t, d.y, o.y), pz = fmaf(t, d.z, o.z);
  const float m = 1e-5f * (scene_mag + fabsf(t) + fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
  if (px > lo.x + m && px < hi.x - m && py > lo.y + m && py < hi.y - m && pz > lo.z + m && pz < hi.z - m) return true;
  return mesh_gate_exact(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, o, d);
}

// MODE 0: nodes / primitives in global memory (LDG.128 through L1)
// MODE 1: the whole BVH and primitive store staged in shared memory (LDS.128) — the reference scenes
// MODE 2: the top `sv.smem_nodes` nodes (breadth-first) staged in shared memory, everything else global — big scenes
template <int MODE>
__device__ __forceinline__ float4 ld16(const float4 *p) {
  if (MODE == 1) return *p;
  return __ldg(p);
}

struct NodeData { float4 n0, n1, nz, mf; };

// The node loop's critical path is load node -> slab tests -> pick a child -> load ITS node: on the 10 M-triangle hall the first
// FFMA after the loads collects 20 % of all stall samples (profiles/r02_tuning.md 8).  The children's indices arrive with the node:
// their records can be on the way while the slab tests run.
template <int MODE>
__device__ __forceinline__ void prefetch_child(const float4 *__restrict__ nodes, const float4 *__restrict__ prims, int c) {
  if (MODE == 1 || !CTB_PREFETCH_CHILDREN) return;
  const char *p = c >= 0 ? reinterpret_cast<const char *>(nodes + 4 * (size_t)c) : reinterpret_cast<const char *>(prims + 3 * (size_t)leaf_first(c));
  if (c == CTB_SENTINEL) return;
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  if (CTB_PREFETCH_CHILDREN >= 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 32));
}

template <int MODE>
__device__ __forceinline__ NodeData load_node(const SceneView &sv, const float4 *__restrict__ nodes, int cur) {
  NodeData n;
  if (MODE == 2) {
    extern __shared__ float4 ctb_dyn_smem[];
    if ((uint32_t)cur < sv.smem_nodes) {
      const float4 *np = ctb_dyn_smem + 4 * cur;
      n.n0 = np[0]; n.n1 = np[1]; n.nz = np[2]; n.mf = np[3];
      return n;
    }
  }
  const float4 *np = nodes + 4 * (size_t)cur;
  n.n0 = ld16<MODE>(np); n.n1 = ld16<MODE>(np + 1); n.nz = ld16<MODE>(np + 2); n.mf = ld16<MODE>(np + 3);
  return n;
}

struct DirCache { vec3 dn; bool have; };   // the normalised direction of a ray, computed when the first sphere is met
template <int MODE>
__device__ __forceinline__ void test_prim(const SceneView &sv, const float4 *__restrict__ prims, uint32_t k, const RayCtx &r, float min_t, Hit &h,
                                          DirCache *dc = nullptr) {
  const float4 *pp = prims + 3 * (size_t)k;
  const float4 q0 = ld16<MODE>(pp), q1 = ld16<MODE>(pp + 1), q2 = ld16<MODE>(pp + 2);
  float t;
  bool ok;
  if (__float_as_uint(q2.w) == CTB_PRIM_TRI) ok = tri_test(mk3(q0.x, q0.y, q0.z), mk3(q1.x, q1.y, q1.z), mk3(q2.x, q2.y, q2.z), r.o, r.d, min_t, t);
  else if (dc) {
    if (!dc->have) { dc->dn = vnormalized(r.d); dc->have = true; }
    ok = sphere_test_n(q0.x, q0.y, q0.z, q1.x, r.o, dc->dn, min_t, &t);
  }
  else ok = sphere_test(q0.x, q0.y, q0.z, q1.x, r.o, r.d, min_t, &t);
  if (ok) {
    const uint32_t obj = __float_as_uint(q0.w), idx = __float_as_uint(q1.w);
    if (hit_better(h, t, obj, idx) && (__float_as_uint(q2.w) != CTB_PRIM_TRI || mesh_gate(sv.obj_bounds, obj, r.o, r.d, t, sv.scene_mag))) {
      h.t = t; h.obj = obj; h.idx = idx; h.ref = k; h.kind = (int)__float_as_uint(q2.w);
    }
  }
}

#if CTB_BVH4
// The same walk over the 4-wide tree (Node4): one iteration loads six float4 box planes + four references and tests four
// boxes; of the children that are hit the nearest (closest hit) or any one (any hit) is entered, the others are pushed.
#define CTB_NONE 0x7ffffffe
template <int MODE, bool ANY>
__device__ __forceinline__ bool traverse4(const SceneView &sv, const float4 *__restrict__ nodes, const float4 *__restrict__ prims,
                                          const RayCtx &r, float min_t, float max_t, Hit &h) {
  int stack[CTB_STACK];
  stack[0] = CTB_SENTINEL;
  int sp = 1;
  int cur = sv.root;
  const float slack = 1.0f + 4.0f * 1.1920929e-7f;
  while (cur != CTB_SENTINEL) {
#pragma unroll 1
    while ((unsigned)cur < (unsigned)CTB_SENTINEL) {   // internal node
      const float4 *np = nodes + 8 * (size_t)cur;
      const float4 lox = ld16<MODE>(np), hix = ld16<MODE>(np + 1), loy = ld16<MODE>(np + 2), hiy = ld16<MODE>(np + 3);
      const float4 loz = ld16<MODE>(np + 4), hiz = ld16<MODE>(np + 5), rf = ld16<MODE>(np + 6);
      const float limit = ANY ? fminf(max_t, h.t) : h.t;
      int nxt = CTB_NONE;
      float tnxt = INFINITY;
#define CTB_CHILD(C)                                                                                                                  \
      {                                                                                                                               \
        const float ax = fmaf(lox.C, r.inv.x, -r.oi.x), bx = fmaf(hix.C, r.inv.x, -r.oi.x);                                          \
        const float ay = fmaf(loy.C, r.inv.y, -r.oi.y), by = fmaf(hiy.C, r.inv.y, -r.oi.y);                                          \
        const float az = fmaf(loz.C, r.inv.z, -r.oi.z), bz = fmaf(hiz.C, r.inv.z, -r.oi.z);                                          \
        const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), min_t));                                     \
        const float tf = fmaf(fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), limit)), slack, r.eabs);                \
        const int rc = __float_as_int(rf.C);                                                                                          \
        if (tn <= tf && rc != CTB_SENTINEL) {   /* (an empty slot's inverted box passes the symmetric slab test) */                   \
          if (ANY || tn < tnxt) { if (nxt != CTB_NONE) stack[sp++] = nxt; nxt = rc; tnxt = tn; }                                      \
          else stack[sp++] = rc;                                                                                                      \
        }                                                                                                                             \
      }
      CTB_CHILD(x) CTB_CHILD(y) CTB_CHILD(z) CTB_CHILD(w)
#undef CTB_CHILD
      cur = nxt != CTB_NONE ? nxt : stack[--sp];
    }
    if (cur == CTB_SENTINEL) break;
    {   // leaf
      const uint32_t first = leaf_first(cur), count = leaf_count(cur);
#pragma unroll 1
      for (uint32_t k = first; k < first + count; k++) test_prim<MODE>(sv, prims, k, r, min_t, h);
      if (ANY && h.t < max_t) return true;
      cur = stack[--sp];
    }
  }
  return false;
}
#endif

// Closest hit over the BVH primitives (planes are handled by the caller).  `h` carries the best hit
// so far in and out.  ANY: stop at the first accepted hit with t < max_t (shadow rays in scenes
// without translucent materials) and return true.  while-while traversal: the inner loop walks
// internal nodes until it reaches a leaf, so the lanes of a warp spend most iterations in the same
// code; the stack bottom holds the sentinel, so popping needs no emptiness test.
template <int MODE, bool ANY, bool BRUTE>
__device__ __forceinline__ bool traverse(const SceneView &sv, const float4 *__restrict__ nodes, const float4 *__restrict__ prims,
                                         const RayCtx &r, float min_t, float max_t, Hit &h) {
  if (BRUTE) {
    DirCache dc;
    dc.have = false;
#pragma unroll 1
    for (uint32_t k = 0; k < sv.n_prims; k++) {
      test_prim<MODE>(sv, prims, k, r, min_t, h, &dc);
      if (ANY && h.t < max_t) return true;
    }
    return false;
  }
#if CTB_BVH4
  return traverse4<MODE, ANY>(sv, nodes, prims, r, min_t, max_t, h);
#else
  int stack[CTB_STACK];
  stack[0] = CTB_SENTINEL;
  int sp = 1;
  int cur = sv.root;
#if CTB_SPECULATIVE
  int pending = 0;   // a postponed leaf (leaf codes are negative, 0 = none)
#endif
  const float slack = 1.0f + 4.0f * 1.1920929e-7f;
  while (cur != CTB_SENTINEL) {
#pragma unroll 1
    while ((unsigned)cur < (unsigned)CTB_SENTINEL) {   // internal node
      const NodeData nd = load_node<MODE>(sv, nodes, cur);
      const float4 n0 = nd.n0, n1 = nd.n1, nz = nd.nz;
      const int c0 = __float_as_int(nd.mf.x), c1 = __float_as_int(nd.mf.y);
      prefetch_child<MODE>(nodes, prims, c0);
      prefetch_child<MODE>(nodes, prims, c1);
      const float limit = ANY ? fminf(max_t, h.t) : h.t;
      // slabs: one FFMA per box plane (error budget: see RayCtx)
      const float c0lox = fmaf(n0.x, r.inv.x, -r.oi.x), c0hix = fmaf(n0.y, r.inv.x, -r.oi.x);
      const float c0loy = fmaf(n0.z, r.inv.y, -r.oi.y), c0hiy = fmaf(n0.w, r.inv.y, -r.oi.y);
      const float c0loz = fmaf(nz.x, r.inv.z, -r.oi.z), c0hiz = fmaf(nz.y, r.inv.z, -r.oi.z);
      const float c1lox = fmaf(n1.x, r.inv.x, -r.oi.x), c1hix = fmaf(n1.y, r.inv.x, -r.oi.x);
      const float c1loy = fmaf(n1.z, r.inv.y, -r.oi.y), c1hiy = fmaf(n1.w, r.inv.y, -r.oi.y);
      const float c1loz = fmaf(nz.z, r.inv.z, -r.oi.z), c1hiz = fmaf(nz.w, r.inv.z, -r.oi.z);
      const float tn0 = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), min_t));
      const float tf0 = fmaf(fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), limit)), slack, r.eabs);
      const float tn1 = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), min_t));
      const float tf1 = fmaf(fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), limit)), slack, r.eabs);
      const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
      if (CTB_BRANCHFREE_STACK && MODE != 1) {
        const bool swap = h1 && (!h0 || tn1 < tn0);   // enter c1 first: it is the only child hit, or the nearer of two
        stack[sp] = swap ? c0 : c1;                   // speculative push of the other child (above the top: harmless when not kept)
        sp += (h0 && h1) ? 1 : 0;
        cur = swap ? c1 : c0;
        if (!(h0 || h1)) cur = stack[--sp];
      } else if (h0 && h1) {
        const bool swap = tn1 < tn0;
        const int far = swap ? c0 : c1;
        stack[sp++] = far;
        cur = swap ? c1 : c0;
#if CTB_PREFETCH_FAR
        if (MODE != 1 && far >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(nodes + 4 * (size_t)far));
#endif
      } else if (h0 || h1) {
        cur = h0 ? c0 : c1;
      } else {
        cur = stack[--sp];
      }
#if CTB_SPECULATIVE
      // speculative traversal: the first leaf a lane finds is postponed and the lane keeps walking nodes, so the lanes
      // of a warp leave the node loop together less often (they stay until they hold a second leaf or run dry)
      if (cur < 0 && pending == 0) { pending = cur; cur = stack[--sp]; }
#endif
    }
#if CTB_SPECULATIVE
    if (pending) {
      const uint32_t first = leaf_first(pending), count = leaf_count(pending);
#pragma unroll 1
      for (uint32_t k = first; k < first + count; k++) test_prim<MODE>(sv, prims, k, r, min_t, h);
      if (ANY && h.t < max_t) return true;
      pending = 0;
    }
#endif
    if (cur == CTB_SENTINEL) break;
    {   // leaf
      const uint32_t first = leaf_first(cur), count = leaf_count(cur);
#pragma unroll 1
      for (uint32_t k = first; k < first + count; k++) test_prim<MODE>(sv, prims, k, r, min_t, h);
      if (ANY && h.t < max_t) return true;
      cur = stack[--sp];
    }
  }
  return false;
#endif
}

// planes: unbounded, kept out of the BVH and always tested (the reference tests them like any other
// object in its linear walk).  Uniform addresses -> broadcast loads.
__device__ __forceinline__ void test_planes(const SceneView &sv, const RayCtx &r, float min_t, Hit &h) {
#pragma unroll 1
  for (uint32_t p = 0; p < sv.n_planes; p++) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + p);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    float t;
    if (plane_test(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), r.o, r.d, min_t, t)) {
      const uint32_t obj = __float_as_uint(a.w);
      if (hit_better(h, t, obj, 0u)) { h.t = t; h.obj = obj; h.idx = 0u; h.ref = p; h.kind = CTB_KIND_PLANE; }
    }
  }
}

// ray_cast (inc/ray_cast.hpp:29-55) over the flat store + LBVH
template <int MODE, bool BRUTE>
__device__ __forceinline__ void closest_hit(const SceneView &sv, const float4 *nodes, const float4 *prims, vec3 o, vec3 d,
                                            float min_t, Hit &h) {
  RayCtx r;
  make_ray(r, o, d, sv.scene_mag);
  hit_reset(h);
  test_planes(sv, r, min_t, h);
  traverse<MODE, false, BRUTE>(sv, nodes, prims, r, min_t, INFINITY, h);
}

// is there any surface with min_t < t < max_t ?  (shadow_intensity with only opaque materials,
// inc/shading.hpp:32-39: the first march step already saturates the intensity)
template <int MODE, bool BRUTE>
__device__ __forceinline__ bool any_hit(const SceneView &sv, const float4 *nodes, const float4 *prims, vec3 o, vec3 d,
                                        float min_t, float max_t) {
  RayCtx r;
  make_ray(r, o, d, sv.scene_mag);
#pragma unroll 1
  for (uint32_t p = 0; p < sv.n_planes; p++) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + p);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    float t;
    if (plane_test(mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), r.o, r.d, min_t, t) && t < max_t) return true;
  }
  Hit h;
  hit_reset(h);
  return traverse<MODE, true, BRUTE>(sv, nodes, prims, r, min_t, max_t, h);
}

// -------------------------------------------------------------------------------------------------
// Which lights can have a PLANE between them and the shaded point X?  plane::intersect (inc/default_schema.hpp:189-192) accepts
// t0 = ((point - X).n) / (d.n) in (1e-3, light_dist).  For a point light P:  d.n = ((P - X).n) / |P - X| = (sX - sP) / L with
// sX = (point - X).n and sP = (point - P).n, so t0 = sX L / (sX - sP): negative when sX and sP have the same sign and |sX| < |sP|,
// beyond the light by the factor 1 / (1 - sP/sX) when |sX| > |sP|, non-finite when they are equal — a plane that has X and P
// clearly on the same side never shadows.  For a sun the sign of d.n is a constant of (light, plane).  The host tabulates
// tbl[l][p] = sP (point light) or -(d.n) (sun), zeroed when it is within 1e-4 |n| x (scene scale) of zero; the test below needs
// sX once per plane (not per light) and per light only a sign comparison.  bit l of the result = "run the exact plane tests
// for light l" (set for every light when there are more than 32 of them or no table).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned plane_side_prepass(const SceneView &sv, vec3 x) {
  if (!sv.pl_tbl || sv.n_lights > 32u) return 0xffffffffu;
  unsigned maybe = 0;
#pragma unroll 1
  for (uint32_t p = 0; p < sv.n_planes; p++) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + p);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    const float sx = vdot(mk3(b.x, b.y, b.z), vsub(mk3(a.x, a.y, a.z), x));
    // "clearly off the plane": far above the rounding error of sx, which grows with |n| (|point| + |x|)
    const float2 e = __ldg(reinterpret_cast<const float2 *>(sv.pl_eps) + p);
    const bool clear = fabsf(sx) > fmaf(e.y, fabsf(x.x) + fabsf(x.y) + fabsf(x.z), e.x);
    const float *row = sv.pl_tbl + p * sv.n_lights;
#pragma unroll 1
    for (uint32_t l = 0; l < sv.n_lights; l++) {
      const float t = __ldg(row + l);
      // same side, and not so lopsided that t0 = L / (1 - sP/sX) could round below the light distance
      const bool same_side = clear && t != 0.f && !((__float_as_uint(sx) ^ __float_as_uint(t)) >> 31) && fabsf(sx) < 1e5f * fabsf(t);
      if (!same_side) maybe |= 1u << l;
    }
  }
  return maybe;
}

// -------------------------------------------------------------------------------------------------
// Shadow-ray packets.  All shadow rays of one shaded hit start at the same point (inc/shading.hpp:80),
// so up to K of them (one per light) walk the BVH together: one stack, one node fetch, the
// (plane - origin) subtractions and the ray-independent parts of Cramer's rule (a, b, d = p2 - o)
// shared.  A node is entered if ANY still-unoccluded ray of the packet hits its box; a ray leaves the
// packet as soon as it finds an occluder in (1e-3, light_dist).  Same accept/reject decisions per ray as
// any_hit(); returns the bit mask of occluded rays.  `act` = rays that take part.
// -------------------------------------------------------------------------------------------------
template <int MODE, int K, bool BRUTE>
__device__ __forceinline__ unsigned any_hit_packet(const SceneView &sv, const float4 *__restrict__ nodes,
                                                   const float4 *__restrict__ prims, vec3 o, const vec3 (&d)[K],
                                                   const float (&max_t)[K], unsigned act, bool test_planes_too = true) {
  const float min_t = (float)(0.0 + 1e-3);   // shadow_intensity: last_hit + 1e-3 with last_hit = 0
  unsigned occ = 0;
  // planes: (point - o).n is shared by the packet.  Skipped when the caller's side test (plane_side_prepass) has already
  // shown that no plane can lie between the origin and these lights.
#pragma unroll 1
  for (uint32_t p = 0; test_planes_too && p < sv.n_planes; p++) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + p);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    const vec3 n = mk3(b.x, b.y, b.z);
    const float num = vdot(n, vsub(mk3(a.x, a.y, a.z), o));
#pragma unroll
    for (int k = 0; k < K; k++) {
      const float den = vdot(d[k], n);
      // t0 > 1e-3 needs num and den of equal sign; then the exact IEEE quotient decides (plane::intersect)
      // (num == 0: the ray starts on the plane, t0 = +-0 or NaN < min_t — and a zero operand is the division's slow path, see plane_test)
      if (!((__float_as_uint(num) ^ __float_as_uint(den)) >> 31) && num != 0.0f && fabsf(num) < CTB_MUL(CTB_MUL(max_t[k], fabsf(den)), 1.0001f)) {
        const float t0 = CTB_DIV(num, den);
        if (isfinite(t0) && min_t <= t0 && t0 > min_t && t0 < max_t[k]) occ |= (act & (1u << k));
      }
    }
  }
  act &= ~occ;
  if (!act) return occ;

  if (BRUTE) {
    DirCache dc[K];
#pragma unroll
    for (int k = 0; k < K; k++) dc[k].have = false;
#pragma unroll 1
    for (uint32_t i = 0; i < sv.n_prims && act; i++) {
#pragma unroll
      for (int k = 0; k < K; k++) {
        if (act & (1u << k)) {
          RayCtx r;
          r.o = o; r.d = d[k];   // test_prim only reads origin and direction (the slab data is for the BVH walk)
          Hit h; hit_reset(h);
          test_prim<MODE>(sv, prims, i, r, min_t, h, &dc[k]);
          if (h.t < max_t[k]) { occ |= 1u << k; act &= ~(1u << k); }
        }
      }
    }
    return occ;
  }

#if CTB_BVH4
  {
    static_assert(K == 1, "the 4-wide walk has no packet form: build with SHADOW_PACKET=1");
    RayCtx r4;
    make_ray(r4, o, d[0], sv.scene_mag);
    Hit h4;
    hit_reset(h4);
    if (traverse4<MODE, true>(sv, nodes, prims, r4, min_t, max_t[0], h4)) occ |= act & 1u;
    return occ;
  }
#endif
  RayCtx rc[K];
#pragma unroll
  for (int k = 0; k < K; k++) make_ray(rc[k], o, d[k], sv.scene_mag);
  int stack[CTB_STACK];
  stack[0] = CTB_SENTINEL;
  int sp = 1;
  int cur = sv.root;
  const float slack = 1.0f + 4.0f * 1.1920929e-7f;
  while (cur != CTB_SENTINEL) {
#pragma unroll 1
    while ((unsigned)cur < (unsigned)CTB_SENTINEL) {
      const NodeData nd = load_node<MODE>(sv, nodes, cur);
      const float4 n0 = nd.n0, n1 = nd.n1, nz = nd.nz;
      const int c0 = __float_as_int(nd.mf.x), c1 = __float_as_int(nd.mf.y);
      prefetch_child<MODE>(nodes, prims, c0);
      prefetch_child<MODE>(nodes, prims, c1);
      bool h0 = false, h1 = false;
#pragma unroll
      for (int k = 0; k < K; k++) {
        const RayCtx &r = rc[k];
        const float lox = fmaf(n0.x, r.inv.x, -r.oi.x), hix = fmaf(n0.y, r.inv.x, -r.oi.x);
        const float loy = fmaf(n0.z, r.inv.y, -r.oi.y), hiy = fmaf(n0.w, r.inv.y, -r.oi.y);
        const float loz = fmaf(nz.x, r.inv.z, -r.oi.z), hiz = fmaf(nz.y, r.inv.z, -r.oi.z);
        const float tn = fmaxf(fmaxf(fminf(lox, hix), fminf(loy, hiy)), fmaxf(fminf(loz, hiz), min_t));
        const float tf = fmaf(fminf(fminf(fmaxf(lox, hix), fmaxf(loy, hiy)), fminf(fmaxf(loz, hiz), max_t[k])), slack, r.eabs);
        const float lox1 = fmaf(n1.x, r.inv.x, -r.oi.x), hix1 = fmaf(n1.y, r.inv.x, -r.oi.x);
        const float loy1 = fmaf(n1.z, r.inv.y, -r.oi.y), hiy1 = fmaf(n1.w, r.inv.y, -r.oi.y);
        const float loz1 = fmaf(nz.z, r.inv.z, -r.oi.z), hiz1 = fmaf(nz.w, r.inv.z, -r.oi.z);
        const float tn1 = fmaxf(fmaxf(fminf(lox1, hix1), fminf(loy1, hiy1)), fmaxf(fminf(loz1, hiz1), min_t));
        const float tf1 = fmaf(fminf(fminf(fmaxf(lox1, hix1), fmaxf(loy1, hiy1)), fminf(fmaxf(loz1, hiz1), max_t[k])), slack, r.eabs);
        const bool on = (act >> k) & 1u;
        h0 = h0 || (on && tn <= tf);
        h1 = h1 || (on && tn1 <= tf1);
      }
      if (CTB_BRANCHFREE_STACK && MODE != 1) {
        stack[sp] = c1;
        sp += (h0 && h1) ? 1 : 0;
        cur = h0 ? c0 : c1;
        if (!(h0 || h1)) cur = stack[--sp];
      } else if (h0 && h1) { stack[sp++] = c1; cur = c0; }
      else if (h0 || h1) cur = h0 ? c0 : c1;
      else cur = stack[--sp];
    }
    if (cur == CTB_SENTINEL) break;
    {
      const uint32_t first = leaf_first(cur), count = leaf_count(cur);
#pragma unroll 1
      for (uint32_t i = first; i < first + count; i++) {
        const float4 *pp = prims + 3 * (size_t)i;
        const float4 q0 = ld16<MODE>(pp), q1 = ld16<MODE>(pp + 1), q2 = ld16<MODE>(pp + 2);
        if (__float_as_uint(q2.w) == CTB_PRIM_TRI) {
          const vec3 p1 = mk3(q0.x, q0.y, q0.z), p2 = mk3(q1.x, q1.y, q1.z), p3 = mk3(q2.x, q2.y, q2.z);
          const vec3 a = vsub(p2, p1), b = vsub(p2, p3), dd = vsub(p2, o);
          unsigned cand = 0;
          float alpha[K], nb[K], ng[K];
#pragma unroll
          for (int k = 0; k < K; k++) {
            alpha[k] = det3(a, b, d[k]);
            nb[k] = det3(dd, b, d[k]);
            ng[k] = det3(a, dd, d[k]);
            const float aa = CTB_MUL(fabsf(alpha[k]), 1e-30f);
            const bool neg_b = (__float_as_uint(nb[k]) ^ __float_as_uint(alpha[k])) >> 31;
            const bool neg_g = (__float_as_uint(ng[k]) ^ __float_as_uint(alpha[k])) >> 31;
            if (!((neg_b && fabsf(nb[k]) > aa) || (neg_g && fabsf(ng[k]) > aa))) cand |= 1u << k;
          }
          cand &= act;
          if (cand) {   // exact path of triangle::intersect for the few rays that pass the sign test
            const float nt = det3_t(a, b, dd);
#pragma unroll
            for (int k = 0; k < K; k++) {
              if (cand & (1u << k)) {
                const float beta = CTB_DIV(nb[k], alpha[k]), gamma = CTB_DIV(ng[k], alpha[k]);
                if (beta >= 0 && gamma >= 0 && CTB_ADD(beta, gamma) <= 1) {
                  const float t0 = CTB_DIV(nt, alpha[k]);
                  if (isfinite(t0) && min_t <= t0 && t0 > min_t && t0 < max_t[k] &&
                      mesh_gate(sv.obj_bounds, __float_as_uint(q0.w), o, d[k], t0, sv.scene_mag)) { occ |= 1u << k; act &= ~(1u << k); }
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < K; k++) {
            float t;
            if ((act & (1u << k)) && sphere_test(q0.x, q0.y, q0.z, q1.x, o, d[k], min_t, &t) && t < max_t[k]) { occ |= 1u << k; act &= ~(1u << k); }
          }
        }
        if (!act) return occ;
      }
      cur = stack[--sp];
    }
  }
  return occ;
}

// -------------------------------------------------------------------------------------------------
// All shadow rays of one shaded hit in ONE traversal loop (big scenes: MODE 0 / 2).  A lane whose ray to light l has
// ended — occluder found, or stack empty — starts its ray to the next light at once instead of idling until the slowest lane
// of the warp has finished light l: the warp runs max over lanes of (sum over lights) node iterations instead of sum over
// lights of (max over lanes).  On the 10 M-triangle hall the shadow walk is 60 % of the frame at 11.8 of 32 lanes.  Directions
// and lengths come from per-thread arrays (local memory, indexed by the light) that the caller fills at full lane
// utilisation; the plane tests are the caller's too.  Per ray the accept / reject decisions are those of any_hit_packet<K = 1>.
// NEGATIVE RESULT (CTB_LIGHTS_IN_ONE_WALK, off): parity green, but the hall got 20 % SLOWER (70.7 -> 85.0 ms).  The lanes that
// are slow for one light are slow for all of them (the cost of a shadow ray is set by where its origin lies), so max-of-sums
// is hardly below sum-of-maxes, and lanes on different rays leave the node loop out of phase.
// `todo` = rays to trace (bit l = light l); returns the bit mask of the occluded ones.
// -------------------------------------------------------------------------------------------------
#define CTB_MULTI_LIGHTS 8
template <int MODE>
__device__ __forceinline__ unsigned any_hit_lights(const SceneView &sv, const float4 *__restrict__ nodes, const float4 *__restrict__ prims, vec3 o,
                                                   const float (&dx)[CTB_MULTI_LIGHTS], const float (&dy)[CTB_MULTI_LIGHTS],
                                                   const float (&dz)[CTB_MULTI_LIGHTS], const float (&len)[CTB_MULTI_LIGHTS], unsigned todo) {
  const float min_t = (float)(0.0 + 1e-3);
  const float slack = 1.0f + 4.0f * 1.1920929e-7f;
  unsigned occ = 0;
  int stack[CTB_STACK];
  int sp = 0, cur = CTB_SENTINEL;
  unsigned bit = 0;
  RayCtx r;
  vec3 d = mk3(0.f, 0.f, 1.f);
  float max_t = 0.f;
  r.o = o; r.d = d; r.inv = d; r.oi = d; r.eabs = 0.f;
  for (;;) {
    if (cur == CTB_SENTINEL) {   // this lane's ray has ended (or it has none yet): the next light
      if (!todo) break;
      const int l = __ffs((int)todo) - 1;
      bit = 1u << l;
      todo &= ~bit;
      d = mk3(dx[l], dy[l], dz[l]);
      max_t = len[l];
      make_ray(r, o, d, sv.scene_mag);
      stack[0] = CTB_SENTINEL;
      sp = 1;
      cur = sv.root;
    }
#pragma unroll 1
    while ((unsigned)cur < (unsigned)CTB_SENTINEL) {
      const NodeData nd = load_node<MODE>(sv, nodes, cur);
      const float4 n0 = nd.n0, n1 = nd.n1, nz = nd.nz;
      const int c0 = __float_as_int(nd.mf.x), c1 = __float_as_int(nd.mf.y);
      const float lox = fmaf(n0.x, r.inv.x, -r.oi.x), hix = fmaf(n0.y, r.inv.x, -r.oi.x);
      const float loy = fmaf(n0.z, r.inv.y, -r.oi.y), hiy = fmaf(n0.w, r.inv.y, -r.oi.y);
      const float loz = fmaf(nz.x, r.inv.z, -r.oi.z), hiz = fmaf(nz.y, r.inv.z, -r.oi.z);
      const float tn = fmaxf(fmaxf(fminf(lox, hix), fminf(loy, hiy)), fmaxf(fminf(loz, hiz), min_t));
      const float tf = fmaf(fminf(fminf(fmaxf(lox, hix), fmaxf(loy, hiy)), fminf(fmaxf(loz, hiz), max_t)), slack, r.eabs);
      const float lox1 = fmaf(n1.x, r.inv.x, -r.oi.x), hix1 = fmaf(n1.y, r.inv.x, -r.oi.x);
      const float loy1 = fmaf(n1.z, r.inv.y, -r.oi.y), hiy1 = fmaf(n1.w, r.inv.y, -r.oi.y);
      const float loz1 = fmaf(nz.z, r.inv.z, -r.oi.z), hiz1 = fmaf(nz.w, r.inv.z, -r.oi.z);
      const float tn1 = fmaxf(fmaxf(fminf(lox1, hix1), fminf(loy1, hiy1)), fmaxf(fminf(loz1, hiz1), min_t));
      const float tf1 = fmaf(fminf(fminf(fmaxf(lox1, hix1), fmaxf(loy1, hiy1)), fminf(fmaxf(loz1, hiz1), max_t)), slack, r.eabs);
      const bool h0 = tn <= tf, h1 = tn1 <= tf1;
      if (h0 && h1) { stack[sp++] = c1; cur = c0; }
      else if (h0 || h1) cur = h0 ? c0 : c1;
      else cur = stack[--sp];
    }
    if (cur == CTB_SENTINEL) continue;   // stack empty: not occluded
    {
      const uint32_t first = leaf_first(cur), count = leaf_count(cur);
      bool found = false;
#pragma unroll 1
      for (uint32_t i = first; i < first + count && !found; i++) {
        const float4 *pp = prims + 3 * (size_t)i;
        const float4 q0 = ld16<MODE>(pp), q1 = ld16<MODE>(pp + 1), q2 = ld16<MODE>(pp + 2);
        if (__float_as_uint(q2.w) == CTB_PRIM_TRI) {
          const vec3 p1 = mk3(q0.x, q0.y, q0.z), p2 = mk3(q1.x, q1.y, q1.z), p3 = mk3(q2.x, q2.y, q2.z);
          const vec3 a = vsub(p2, p1), b = vsub(p2, p3), dd = vsub(p2, o);
          const float alpha = det3(a, b, d), nb = det3(dd, b, d), ng = det3(a, dd, d);
          const float aa = CTB_MUL(fabsf(alpha), 1e-30f);
          const bool neg_b = (__float_as_uint(nb) ^ __float_as_uint(alpha)) >> 31;
          const bool neg_g = (__float_as_uint(ng) ^ __float_as_uint(alpha)) >> 31;
          if (!((neg_b && fabsf(nb) > aa) || (neg_g && fabsf(ng) > aa))) {   // exact path of triangle::intersect
            const float nt = det3_t(a, b, dd);
            const float beta = CTB_DIV(nb, alpha), gamma = CTB_DIV(ng, alpha);
            if (beta >= 0 && gamma >= 0 && CTB_ADD(beta, gamma) <= 1) {
              const float t0 = CTB_DIV(nt, alpha);
              if (isfinite(t0) && min_t <= t0 && t0 > min_t && t0 < max_t && mesh_gate(sv.obj_bounds, __float_as_uint(q0.w), o, d, t0, sv.scene_mag)) found = true;
            }
          }
        } else {
          float t;
          if (sphere_test(q0.x, q0.y, q0.z, q1.x, o, d, min_t, &t) && t < max_t) found = true;
        }
      }
      if (found) { occ |= bit; cur = CTB_SENTINEL; }
      else cur = stack[--sp];
    }
  }
  return occ;
}

// planes between a shadow ray's origin and its light (the plane loop of any_hit_packet for one ray)
__device__ __forceinline__ bool planes_occlude(const SceneView &sv, vec3 o, vec3 d, float max_t) {
  const float min_t = (float)(0.0 + 1e-3);
  bool occ = false;
#pragma unroll 1
  for (uint32_t p = 0; p < sv.n_planes; p++) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + p);
    const float4 a = __ldg(pp), b = __ldg(pp + 1);
    const vec3 n = mk3(b.x, b.y, b.z);
    const float num = vdot(n, vsub(mk3(a.x, a.y, a.z), o));
    const float den = vdot(d, n);
    if (!((__float_as_uint(num) ^ __float_as_uint(den)) >> 31) && num != 0.0f && fabsf(num) < CTB_MUL(CTB_MUL(max_t, fabsf(den)), 1.0001f)) {
      const float t0 = CTB_DIV(num, den);
      if (isfinite(t0) && min_t <= t0 && t0 > min_t && t0 < max_t) occ = true;
    }
  }
  return occ;
}

// surface point and raw normal of a hit, as the primitive's intersect() reports them
template <int MODE>
__device__ __forceinline__ void hit_surface(const SceneView &sv, const float4 *prims, const Hit &h, vec3 o, vec3 d,
                                            vec3 &point, vec3 &normal) {
  if (h.kind == CTB_KIND_PLANE) {
    const float4 *pp = reinterpret_cast<const float4 *>(sv.planes + h.ref);
    const float4 b = __ldg(pp + 1);
    point = vmad(o, d, h.t);                            // inc/default_schema.hpp:194
    normal = mk3(b.x, b.y, b.z);                        // :195 stored normal, not normalised
    return;
  }
  const float4 *pp = prims + 3 * (size_t)h.ref;
  const float4 q0 = ld16<MODE>(pp), q1 = ld16<MODE>(pp + 1), q2 = ld16<MODE>(pp + 2);
  if (h.kind == CTB_KIND_TRI) {
    vec3 p1 = mk3(q0.x, q0.y, q0.z), p2 = mk3(q1.x, q1.y, q1.z), p3 = mk3(q2.x, q2.y, q2.z);
    point = vmad(o, d, h.t);                                                      // :71
    normal = vscale(vnormalized(vcross(vsub(p2, p3), vsub(p1, p3))), -1.0f);      // :72
  } else {
    vec3 c = mk3(q0.x, q0.y, q0.z);
    point = vmad(o, vnormalized(d), h.t);                                         // :245
    normal = vnormalized(vsub(point, c));                                         // :246
  }
}

}  // namespace ctb
#endif
