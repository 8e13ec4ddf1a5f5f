// render.cu — the wavefront pipeline that replaces the reference's recursive one-thread-per-pixel
// kernel (inc/kernel.hpp:35-60 -> inc/shading.hpp:116-154 -> inc/ray_cast.hpp:29-55).
//
// Per bounce level L (0 = primary rays) two persistent, work-stealing kernels run:
//
//   trace_kernel  closest hit for every ray of level L (L = 0: rays are generated from the pixel
//                 index, cam::get_ray inc/default_schema.hpp:376-386).  Level 0 also writes the
//                 G-buffer (depth / raw normal / object id, inc/kernel.hpp:52-56) and reduces the
//                 largest finite depth (inc/kernel.hpp:120-125).  Every hit emits one ShadeRec;
//                 mirror / transparent materials push child rays into the level L+1 queue
//                 (inc/shading.hpp:130-149), compacted with warp ballot + popc and ONE atomicAdd
//                 per warp.
//   shade_kernel  per ShadeRec: all shadow rays (shadow_intensity, inc/shading.hpp:22-45) and the
//                 Phong sum (inc/shading.hpp:64-99), accumulated into the colour buffer with the
//                 path weight of the hit.
//
// The recursion  rgb = (1-t)*(phong + r*R) + t*T  (inc/shading.hpp:138,148) is unrolled into path
// weights: own Phong term w*(1-t), reflected child w*(1-t)*r, transmitted child w*t; at the last
// level (bounces exhausted) the blend is skipped exactly like `if constexpr(bounces != 0)`.
//
// Kernels are persistent: grid = SM count x resident CTAs; each warp claims WORK_CHUNK rays at a time
// from a global cursor (work stealing), so cheap and expensive rays balance without a tail.
#include "render.cuh"
#include "trace.cuh"

namespace ctb {

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Work stealing with guided chunk sizes: a warp claims `remaining / (2 * warps in flight)` items, at least one
// warp-iteration (32) and at most WORK_CHUNK_MAX.  Large claims while there is plenty of work keep the cursor
// atomics rare; 32-item claims at the end keep the tail short — at deep bounce levels of the 10 M-triangle scene a
// few grazing rays cost milliseconds each and fixed 128-ray claims left the GPU idle behind them
// (profiles/r01_tuning.md, "tile scaling").
__device__ __forceinline__ bool claim_work(unsigned *cursor, unsigned n_work, unsigned lane, unsigned &base, unsigned &end) {
  unsigned b = 0, chunk = 0;
  if (lane == 0) {
    const unsigned cur = *reinterpret_cast<volatile unsigned *>(cursor);
    const unsigned remaining = cur < n_work ? n_work - cur : 0u;
    const unsigned warps = gridDim.x * (blockDim.x >> 5);
    chunk = remaining / (2u * warps);
    chunk = chunk < 32u ? 32u : (chunk > (unsigned)WORK_CHUNK_MAX ? (unsigned)WORK_CHUNK_MAX : chunk);
    chunk &= ~31u;
    b = atomicAdd(cursor, chunk);
  }
  base = __shfl_sync(0xffffffffu, b, 0);
  chunk = __shfl_sync(0xffffffffu, chunk, 0);
  end = base + chunk < n_work ? base + chunk : n_work;
  return base < n_work;
}

// stage the BVH nodes and the primitive store in shared memory (MODE 1) — LDS.128 instead of
// divergent LDG.128 for the small reference scenes whose whole BVH fits next to the SM
template <int MODE>
__device__ __forceinline__ void stage_scene(const SceneView &sv, float4 *smem, const float4 *&nodes, const float4 *&prims) {
  if (MODE == 1) {
    const float4 *gn = reinterpret_cast<const float4 *>(sv.nodes);
    const float4 *gp = reinterpret_cast<const float4 *>(sv.prims);
    uint32_t nn = sv.n_nodes * 4u, np = sv.n_prims * 3u;
    for (uint32_t i = threadIdx.x; i < nn; i += blockDim.x) smem[i] = __ldg(gn + i);
    for (uint32_t i = threadIdx.x; i < np; i += blockDim.x) smem[nn + i] = __ldg(gp + i);
    __syncthreads();
    nodes = smem;
    prims = smem + nn;
  } else {
    if (MODE == 2) {   // top of the BVH (first smem_nodes nodes, breadth-first) -> shared memory
      const float4 *gn = reinterpret_cast<const float4 *>(sv.nodes);
      const uint32_t nn = sv.smem_nodes * 4u;
      for (uint32_t i = threadIdx.x; i < nn; i += blockDim.x) smem[i] = __ldg(gn + i);
      __syncthreads();
    }
    nodes = reinterpret_cast<const float4 *>(sv.nodes);
    prims = reinterpret_cast<const float4 *>(sv.prims);
  }
}

// local work index -> pixel.  Inside a tile, 32 consecutive indices form an 8x4 pixel block
// (coherent primary rays per warp); the framebuffer itself is row-major inside the tile.
__device__ __forceinline__ bool work_to_pixel(const TileMap &tm, uint32_t i, uint32_t &x, uint32_t &y, uint32_t &pix) {
  static_assert(CUTRACE_TILE_SHIFT >= 3, "a tile holds whole 8x4 warp blocks");
  constexpr uint32_t BX_SHIFT = CUTRACE_TILE_SHIFT - 3;   // log2(8x4 blocks per tile row)
  uint32_t lt = i >> (2 * CUTRACE_TILE_SHIFT), w = i & (CUTRACE_TILE_PIXELS - 1u);
  uint32_t b = w >> 5, l = w & 31u;
  uint32_t px = ((b & ((1u << BX_SHIFT) - 1u)) << 3) + (l & 7u), py = ((b >> BX_SHIFT) << 2) + (l >> 3);
  uint32_t tx, ty;
  if (!tile_of_slot(tm, lt * tm.world + tm.rank, tx, ty)) return false;
  x = tx * CUTRACE_TILE + px;
  y = ty * CUTRACE_TILE + py;
  pix = (lt << (2 * CUTRACE_TILE_SHIFT)) + (py << CUTRACE_TILE_SHIFT) + px;
  return x < tm.width && y < tm.height;
}

// cam::get_ray, inc/default_schema.hpp:376-386
__device__ __forceinline__ void camera_ray(const Camera &c, uint32_t x, uint32_t y, vec3 &o, vec3 &d) {
  float aspect = (float)c.w / (float)c.h;
  vec3 x_v = vscale(c.right, (((float)x / (float)c.w) - 0.5f) * aspect);
  vec3 y_v = vscale(c.up, 0.5f - ((float)y / (float)c.h));
  vec3 z_v = c.forward;
  o = c.pos;
  d = vnormalized(vadd(vadd(x_v, y_v), z_v));
}

template <int MODE, bool BRUTE>
__global__ void __launch_bounds__(TRACE_THREADS, CTB_MIN_BLOCKS)
trace_kernel(const SceneView sv, const TileMap tm, uint32_t level, uint32_t bounces, uint32_t px_base, uint32_t n_px,
             const RayRec *__restrict__ rays_in, RayRec *__restrict__ rays_out, ShadeRec *__restrict__ shade_out,
             FrameCounters *ctr, FrameTargets fb, uint32_t *__restrict__ nlev) {
  extern __shared__ float4 smem[];
  const float4 *nodes, *prims;
  stage_scene<MODE>(sv, smem, nodes, prims);
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = lanemask_lt();
  const uint32_t n_work = level == 0 ? n_px : ctr->n_rays[level];
  unsigned long long n_refl = 0, n_trans = 0, n_shaded = 0;
  float max_depth = 0.f;
  unsigned s_next = 0, s_end = 0, r_next = 0, r_end = 0;   // this warp's reserved slot blocks (warp-uniform)
  // block size: 1/8 of what a warp is expected to emit over the kernel, 32..SLOT_BLOCK — every warp leaves half a block
  // of holes behind on average, so small levels (deep bounces, 1/8 shards) fall back to one reservation per iteration
  unsigned slot_block = (n_work / (gridDim.x * (blockDim.x >> 5) * (unsigned)CTB_SLOT_DIV)) & ~31u;
  slot_block = slot_block < (unsigned)CTB_SLOT_MIN ? (unsigned)CTB_SLOT_MIN : (slot_block > (unsigned)SLOT_BLOCK ? (unsigned)SLOT_BLOCK : slot_block);

  for (;;) {
    unsigned base, end;
    if (!claim_work(&ctr->work_trace[level], n_work, lane, base, end)) break;
#pragma unroll 1
    for (unsigned off = 0; base + off < end; off += 32) {
      const uint32_t i = base + off + lane;
      bool active = i < n_work;
      vec3 o = mk3(0, 0, 0), d = mk3(0, 0, 1);
      float w = 1.0f;
      uint32_t pix = 0, gx = 0, gy = 0;
      if (level == 0) {
        active = active && work_to_pixel(tm, px_base + i, gx, gy, pix);
        if (active) camera_ray(sv.cam, gx, gy, o, d);
      } else if (active) {
        const float4 *rp = reinterpret_cast<const float4 *>(rays_in + i);
        float4 a = __ldcs(rp), b = __ldcs(rp + 1);
        o = mk3(a.x, a.y, a.z); pix = __float_as_uint(a.w);
        d = mk3(b.x, b.y, b.z); w = b.w;
        active = pix != CTB_HOLE;
      }
      Hit h;
      hit_reset(h);
      if (active) closest_hit<MODE, BRUTE>(sv, nodes, prims, o, d, sv.fudge, h);
      const bool hit = active && h.kind >= 0;
      vec3 point = mk3(0, 0, 0), nrm = mk3(0, 0, 0);
      uint32_t mat = 0;
      float reflect = 0.f, transp = 0.f;
      if (hit) {
        hit_surface<MODE>(sv, prims, h, o, d, point, nrm);
        mat = __ldg(sv.obj_material + h.obj);
        const float4 *mp = reinterpret_cast<const float4 *>(sv.materials + mat);
        float4 m1 = __ldg(mp + 1);
        reflect = m1.x; transp = m1.z;
      }
      if (level == 0 && active) {   // G-buffer, inc/kernel.hpp:52-56
        const size_t gi = fb.row_major ? (size_t)gy * tm.width + gx : (size_t)pix;   // possibly peer memory (NVLink store)
        fb.depth[gi] = h.t;
        fb.normal[3 * gi] = nrm.x; fb.normal[3 * gi + 1] = nrm.y; fb.normal[3 * gi + 2] = nrm.z;
        fb.hit_id[gi] = hit ? h.obj : CUTRACE_NO_HIT;
        if (hit && isfinite(h.t)) max_depth = fmaxf(max_depth, h.t);
        if (nlev && !hit) nlev[pix] = 0u;
      }
      // number of bounce levels that contributed to this pixel so far (levels run in order on one stream)
      if (nlev && hit) nlev[pix] = level + 1u;
      // inc/shading.hpp:126-149
      bool do_refl = false, do_trans = false;
      float w_own = w;
      if (hit && level < bounces) {
        do_refl = (double)reflect >= 1e-6;
        do_trans = (double)transp >= 1e-6;
        if (do_trans) w_own = w * (1.0f - transp);
      }
      // ---- output slots.  A warp reserves SLOT_BLOCK queue slots with ONE atomicAdd and fills them over the next
      // iterations (ballot + popc compaction inside the warp); the unused tail of a block is retired as holes
      // (pix = CTB_HOLE).  One same-address atomic WITH return per warp-iteration and queue had made the trace
      // kernels wait on the L2 atomic unit (ncu: long_scoreboard 11.5 warp-cycles per issue, 45 % issue slots).
      const unsigned m_hit = __ballot_sync(0xffffffffu, hit);
      if (m_hit) {
        const unsigned n_hit = __popc(m_hit);
        if (s_next + n_hit > s_end) {
          for (unsigned k = s_next + lane; k < s_end; k += 32)
            __stcs(reinterpret_cast<float4 *>(shade_out + k), make_float4(0.f, 0.f, 0.f, __uint_as_float(CTB_HOLE)));
          unsigned b = 0;
          const unsigned blk = slot_block > n_hit ? slot_block : n_hit;
          if (lane == 0) b = atomicAdd(&ctr->n_shade[level], blk);
          s_next = __shfl_sync(0xffffffffu, b, 0);
          s_end = s_next + blk;
        }
        if (hit) {
          float4 *sp = reinterpret_cast<float4 *>(shade_out + s_next + __popc(m_hit & lt_mask));
          __stcs(sp, make_float4(point.x, point.y, point.z, __uint_as_float(pix)));
          __stcs(sp + 1, make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(mat)));
          __stcs(sp + 2, make_float4(d.x, d.y, d.z, w_own));
          n_shaded++;
        }
        s_next += n_hit;
      }
      // ---- child rays ----
      const unsigned m_r = __ballot_sync(0xffffffffu, do_refl), m_t = __ballot_sync(0xffffffffu, do_trans);
      if (m_r | m_t) {
        const unsigned nr = __popc(m_r), nt = __popc(m_t);
        if (r_next + nr + nt > r_end) {
          for (unsigned k = r_next + lane; k < r_end; k += 32)
            __stcs(reinterpret_cast<float4 *>(rays_out + k), make_float4(0.f, 0.f, 0.f, __uint_as_float(CTB_HOLE)));
          unsigned b = 0;
          const unsigned blk = slot_block > nr + nt ? slot_block : nr + nt;
          if (lane == 0) b = atomicAdd(&ctr->n_rays[level + 1], blk);
          r_next = __shfl_sync(0xffffffffu, b, 0);
          r_end = r_next + blk;
        }
        const vec3 origin = vadd(o, vscale(d, h.t));   // incoming->start + distance * incoming->dir
        if (do_refl) {
          vec3 nd = vnormalized(d), nn = vnormalized(nrm);
          vec3 rd = vreflect(nd, nn);
          float4 *rp = reinterpret_cast<float4 *>(rays_out + r_next + __popc(m_r & lt_mask));
          __stcs(rp, make_float4(origin.x, origin.y, origin.z, __uint_as_float(pix)));
          __stcs(rp + 1, make_float4(rd.x, rd.y, rd.z, w_own * reflect));
          n_refl++;
        }
        if (do_trans) {
          float4 *rp = reinterpret_cast<float4 *>(rays_out + r_next + nr + __popc(m_t & lt_mask));
          __stcs(rp, make_float4(origin.x, origin.y, origin.z, __uint_as_float(pix)));
          __stcs(rp + 1, make_float4(d.x, d.y, d.z, w * transp));
          n_trans++;
        }
        r_next += nr + nt;
      }
    }
  }
  // retire the unused tails of this warp's last blocks
  for (unsigned k = s_next + lane; k < s_end; k += 32)
    __stcs(reinterpret_cast<float4 *>(shade_out + k), make_float4(0.f, 0.f, 0.f, __uint_as_float(CTB_HOLE)));
  for (unsigned k = r_next + lane; k < r_end; k += 32)
    __stcs(reinterpret_cast<float4 *>(rays_out + k), make_float4(0.f, 0.f, 0.f, __uint_as_float(CTB_HOLE)));
  // per-warp totals
  for (int s = 16; s > 0; s >>= 1) {
    n_refl += __shfl_xor_sync(0xffffffffu, n_refl, s);
    n_trans += __shfl_xor_sync(0xffffffffu, n_trans, s);
    n_shaded += __shfl_xor_sync(0xffffffffu, n_shaded, s);
    max_depth = fmaxf(max_depth, __shfl_xor_sync(0xffffffffu, max_depth, s));
  }
  if (lane == 0) {
    if (n_refl) atomicAdd(&ctr->rays_reflect, n_refl);
    if (n_shaded) atomicAdd(&ctr->shade_records, n_shaded);
    if (n_trans) atomicAdd(&ctr->rays_transmit, n_trans);
    if (level == 0 && max_depth > 0.f) atomicMax(&ctr->max_depth_bits, __float_as_uint(max_depth));
  }
}

// shadow_intensity, inc/shading.hpp:22-45
#ifndef SHADOW_PACKET
#define SHADOW_PACKET 1   // shadow rays of one hit that traverse together
#endif

template <int MODE, bool BRUTE, bool OPAQUE>
__device__ __forceinline__ float shadow_intensity(const SceneView &sv, const float4 *nodes, const float4 *prims, vec3 o, vec3 d,
                                                  float max_dist, unsigned long long &casts) {
  if (OPAQUE) {
    // every crossing adds 1 - 0 = 1 -> the first march step saturates: any surface in (1e-3, max_dist) shadows fully
    casts++;
    return any_hit<MODE, BRUTE>(sv, nodes, prims, o, d, (float)(0.0 + 1e-3), max_dist) ? 1.0f : 0.0f;
  }
  float intensity = 0.0f, last_hit = 0.0f;
  for (;;) {
    Hit h;
    casts++;
    closest_hit<MODE, BRUTE>(sv, nodes, prims, o, d, (float)((double)last_hit + 1e-3), h);
    if (!(h.kind >= 0 && h.t < max_dist)) break;
    const float4 *mp = reinterpret_cast<const float4 *>(sv.materials + __ldg(sv.obj_material + h.obj));
    float trans = __ldg(mp + 1).z;
    intensity += (1.0f - trans);
    if (intensity >= 1.0f) return 1.0f;
    last_hit = h.t;
  }
  return intensity;
}

template <int MODE, bool BRUTE, bool OPAQUE>
__global__ void __launch_bounds__(TRACE_THREADS, CTB_SHADE_MIN_BLOCKS)
shade_kernel(const SceneView sv, uint32_t level, const ShadeRec *__restrict__ shade, FrameCounters *ctr, FrameTargets fb,
             int atomic_accumulate, float *__restrict__ level_color, uint32_t px_base) {
  extern __shared__ float4 smem[];
  const float4 *nodes, *prims;
  stage_scene<MODE>(sv, smem, nodes, prims);
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t n_work = ctr->n_shade[level];
  unsigned long long casts = 0;

  for (;;) {
    unsigned base, end;
    if (!claim_work(&ctr->work_shade[level], n_work, lane, base, end)) break;
#pragma unroll 1
    for (unsigned off = 0; base + off < end; off += 32) {
      const uint32_t i = base + off + lane;
      if (i < n_work) {
        const float4 *sp = reinterpret_cast<const float4 *>(shade + i);
        const float4 s0 = __ldcs(sp), s1 = __ldcs(sp + 1), s2 = __ldcs(sp + 2);
        const vec3 hit = mk3(s0.x, s0.y, s0.z), normal = mk3(s1.x, s1.y, s1.z), in_dir = mk3(s2.x, s2.y, s2.z);
        const uint32_t pix = __float_as_uint(s0.w), mat = __float_as_uint(s1.w);
        const float weight = s2.w;
        if (pix == CTB_HOLE) continue;   // retired tail of a producer warp's slot block
        // phong, inc/shading.hpp:64-99
        const float4 *mp = reinterpret_cast<const float4 *>(sv.materials + mat);
        const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
        const vec3 diffuse = mk3(m0.x, m0.y, m0.z);
        const vec3 specular = vscale(diffuse, m0.w);        // *spec = specular * color
        const float phong_exp = m1.y;
        vec3 final = vscale(diffuse, sv.cam.ambient);
        const vec3 nn = vnormalized(normal);
        const vec3 in_n = vscale(vnormalized(in_dir), -1.0f);
        if (OPAQUE) {
          // every material is opaque: the shadow march saturates on its first step, so the K shadow rays
          // of this hit are any-hit queries and walk the BVH as one packet
          for (uint32_t l0 = 0; l0 < sv.n_lights; l0 += SHADOW_PACKET) {
            vec3 sd[SHADOW_PACKET];
            float md[SHADOW_PACKET];
            vec3 lcol[SHADOW_PACKET];
            unsigned valid = 0;
#pragma unroll
            for (int k = 0; k < SHADOW_PACKET; k++) {
              sd[k] = mk3(0.f, 0.f, 1.f); md[k] = 0.f; lcol[k] = mk3(0.f, 0.f, 0.f);
              if (l0 + k < sv.n_lights) {
                const float4 *lp = reinterpret_cast<const float4 *>(sv.lights + l0 + k);
                const float4 l0v = __ldg(lp), l1v = __ldg(lp + 1);
                vec3 direction;
                float distance;
                if (__float_as_uint(l0v.w) == CUTRACE_LIGHT_SUN) {     // inc/default_schema.hpp:280-283
                  direction = vscale(mk3(l0v.x, l0v.y, l0v.z), -1.0f);
                  distance = INFINITY;
                } else {                                               // inc/default_schema.hpp:305-308
                  vec3 P = mk3(l0v.x, l0v.y, l0v.z);
                  direction = vnormalized(vsub(P, hit));
                  distance = vnorm(vsub(P, hit));
                }
                sd[k] = vnormalized(direction);
                md[k] = distance * vnorm(direction);
                lcol[k] = mk3(l1v.x, l1v.y, l1v.z);
                valid |= 1u << k;
              }
            }
            casts += __popc(valid);
            const unsigned occ = any_hit_packet<MODE, SHADOW_PACKET, BRUTE>(sv, nodes, prims, hit, sd, md, valid);
#pragma unroll
            for (int k = 0; k < SHADOW_PACKET; k++) {
              if ((valid & ~occ) & (1u << k)) {        // shadow_fac = 0 < 1
                const vec3 nd = sd[k];
                float fd = fmaxf(0.0f, vdot(nn, nd));
                vec3 ld = vmul(diffuse, lcol[k]);
                vec3 hv = vnormalized(vadd(in_n, nd));
                float fs = powf(fmaxf(0.0f, vdot(nn, hv)), phong_exp);
                vec3 ls = vmul(specular, lcol[k]);
                vec3 term = vscale(vadd(vscale(ld, fd), vscale(ls, fs)), 1 - 0.0f);
                final.x += term.x; final.y += term.y; final.z += term.z;
              }
            }
          }
        } else {
          for (uint32_t l = 0; l < sv.n_lights; l++) {
            const float4 *lp = reinterpret_cast<const float4 *>(sv.lights + l);
            const float4 l0 = __ldg(lp), l1 = __ldg(lp + 1);
            vec3 direction;
            float distance;
            if (__float_as_uint(l0.w) == CUTRACE_LIGHT_SUN) {     // inc/default_schema.hpp:280-283
              direction = vscale(mk3(l0.x, l0.y, l0.z), -1.0f);
              distance = INFINITY;
            } else {                                              // inc/default_schema.hpp:305-308
              vec3 P = mk3(l0.x, l0.y, l0.z);
              direction = vnormalized(vsub(P, hit));
              distance = vnorm(vsub(P, hit));
            }
            const vec3 sdir = vnormalized(direction);
            const float light_dist = distance * vnorm(direction);
            const vec3 color = mk3(l1.x, l1.y, l1.z);
            const vec3 nd = sdir;
            const float shadow_fac = shadow_intensity<MODE, BRUTE, false>(sv, nodes, prims, hit, sdir, light_dist, casts);
            if (shadow_fac < 1.0f) {
              float fd = fmaxf(0.0f, vdot(nn, nd));
              vec3 ld = vmul(diffuse, color);
              vec3 hv = vnormalized(vadd(in_n, nd));
              float fs = powf(fmaxf(0.0f, vdot(nn, hv)), phong_exp);
              vec3 ls = vmul(specular, color);
              vec3 term = vscale(vadd(vscale(ld, fd), vscale(ls, fs)), 1 - shadow_fac);
              final.x += term.x; final.y += term.y; final.z += term.z;
            }
          }
        }
        if (level_color) {   // one hit per pixel and level: plain store into this level's partial image
          float *lp = level_color + 3 * (size_t)(pix - px_base);
          lp[0] = weight * final.x; lp[1] = weight * final.y; lp[2] = weight * final.z;
          continue;
        }
        float *cp = fb.color + 3 * (size_t)pix;
        if (atomic_accumulate) {
          atomicAdd(cp, weight * final.x); atomicAdd(cp + 1, weight * final.y); atomicAdd(cp + 2, weight * final.z);
        } else {
          cp[0] += weight * final.x; cp[1] += weight * final.y; cp[2] += weight * final.z;
        }
      }
    }
  }
  for (int s = 16; s > 0; s >>= 1) casts += __shfl_xor_sync(0xffffffffu, casts, s);
  if (lane == 0 && casts) atomicAdd(&ctr->shadow_casts, casts);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
// colour of a pixel = sum of its per-level partial images in level order (same order as a serial accumulation),
// stored where the frame lives: tile-major local buffer, own row-major frame, or a peer GPU's frame over NVLink
__global__ void combine_levels_kernel(const TileMap tm, const uint32_t *__restrict__ nlev, const float *__restrict__ level_color,
                                      uint64_t level_stride, uint32_t levels, const float *__restrict__ local_color,
                                      uint32_t px_base, uint32_t n_px, FrameTargets out, FrameTargets gsrc) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  const uint32_t pix = px_base + i;
  uint32_t x, y;
  if (!pixel_of_local(tm, pix, x, y)) return;
  float r = 0.f, g = 0.f, b = 0.f;
  if (levels == 0) {
    const float *p = local_color + 3 * (size_t)pix;
    r = p[0]; g = p[1]; b = p[2];
  } else {
    uint32_t n = nlev[pix];
    n = n < levels ? n : levels;
    for (uint32_t l = 0; l < n; l++) {
      const float *p = level_color + l * level_stride + 3 * (size_t)i;
      r += p[0]; g += p[1]; b += p[2];
    }
  }
  const size_t gi = out.row_major ? (size_t)y * tm.width + x : (size_t)pix;
  float *o = out.color + 3 * gi;
  o[0] = r; o[1] = g; o[2] = b;
  if (gsrc.depth) {
    // peer frame: the level-0 G-buffer was written to the local tile-major buffers (small 8x4-block stores are slow
    // over NVLink); it travels here together with the colour, 16-pixel tile rows (64..192 contiguous bytes) per half warp
    out.depth[gi] = gsrc.depth[pix];
    out.hit_id[gi] = gsrc.hit_id[pix];
    out.normal[3 * gi] = gsrc.normal[3 * (size_t)pix]; out.normal[3 * gi + 1] = gsrc.normal[3 * (size_t)pix + 1];
    out.normal[3 * gi + 2] = gsrc.normal[3 * (size_t)pix + 2];
  }
}

// forwards the level-0 G-buffer of this rank's tiles (tile-major, local) to a row-major frame in peer memory: one
// 16-pixel tile row per half warp -> 64 / 192-byte contiguous NVLink stores.  Runs on an auxiliary stream right after
// trace(0), i.e. the transfer overlaps the remaining bounce levels.
__global__ void export_gbuffer_kernel(const TileMap tm, uint32_t px_base, uint32_t n_px, FrameTargets src, FrameTargets out) {
  // a SMALL grid-stride grid: the kernel is bound by the NVLink stores (all ranks push into rank 0 at the same moment),
  // and every resident CTA of it takes SM slots away from the persistent trace/shade CTAs it is supposed to run under
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
    const uint32_t pix = px_base + i;
    uint32_t x, y;
    if (!pixel_of_local(tm, pix, x, y)) continue;
    const size_t gi = (size_t)y * tm.width + x;
    out.depth[gi] = src.depth[pix];
    out.hit_id[gi] = src.hit_id[pix];
    out.normal[3 * gi] = src.normal[3 * (size_t)pix]; out.normal[3 * gi + 1] = src.normal[3 * (size_t)pix + 1];
    out.normal[3 * gi + 2] = src.normal[3 * (size_t)pix + 2];
  }
}

void launch_export_gbuffer(const TileMap &tm, uint32_t px_base, uint32_t n_px, const FrameTargets &src, const FrameTargets &out,
                           cudaStream_t st) {
  if (!n_px) return;
  uint32_t grid = (n_px + 255) / 256;
  if (grid > CTB_EXPORT_CTAS) grid = CTB_EXPORT_CTAS;
  export_gbuffer_kernel<<<grid, 256, 0, st>>>(tm, px_base, n_px, src, out);
}

void launch_combine(const TileMap &tm, const uint32_t *nlev, const float *level_color, uint64_t level_stride, uint32_t levels,
                    const float *local_color, uint32_t px_base, uint32_t n_px, const FrameTargets &out, const FrameTargets &gsrc,
                    cudaStream_t st) {
  if (!n_px) return;
  combine_levels_kernel<<<(n_px + 255) / 256, 256, 0, st>>>(tm, nlev, level_color, level_stride, levels, local_color, px_base, n_px, out, gsrc);
}

typedef void (*trace_fn)(const SceneView, const TileMap, uint32_t, uint32_t, uint32_t, uint32_t, const RayRec *, RayRec *, ShadeRec *,
                         FrameCounters *, FrameTargets, uint32_t *);
typedef void (*shade_fn)(const SceneView, uint32_t, const ShadeRec *, FrameCounters *, FrameTargets, int, float *, uint32_t);

static trace_fn pick_trace(int mode, bool brute) {
  if (brute) return trace_kernel<0, true>;
  if (mode == 2) return trace_kernel<2, false>;
  return mode == 1 ? trace_kernel<1, false> : trace_kernel<0, false>;
}
static shade_fn pick_shade(int mode, bool brute, bool opaque) {
  if (brute) return opaque ? shade_kernel<0, true, true> : shade_kernel<0, true, false>;
  if (mode == 1) return opaque ? shade_kernel<1, false, true> : shade_kernel<1, false, false>;
  if (mode == 2) return opaque ? shade_kernel<2, false, true> : shade_kernel<2, false, false>;
  return opaque ? shade_kernel<0, false, true> : shade_kernel<0, false, false>;
}

cudaError_t plan_launch(const SceneView &sv, bool allow_smem, LaunchCfg *cfg) {
  int dev = 0, sms = 0, smem_optin = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  size_t need = (size_t)sv.n_nodes * sizeof(Node) + (size_t)sv.n_prims * sizeof(PrimRec);
  cfg->mode = 0;
  cfg->smem_bytes = 0;
  if (allow_smem && !sv.brute_force && sv.n_prims > 0 && need + 1024 <= (size_t)smem_optin) {
    cfg->mode = 1;
    cfg->smem_bytes = need;
  } else if (allow_smem && !sv.brute_force && sv.smem_nodes > 0) {
    cfg->mode = 2;   // api.cu moved the top sv.smem_nodes nodes (breadth-first) to the front of the node array
    cfg->smem_bytes = (size_t)sv.smem_nodes * sizeof(Node);
  }
  trace_fn tf = pick_trace(cfg->mode, sv.brute_force != 0);
  shade_fn sf = pick_shade(cfg->mode, sv.brute_force != 0, sv.all_opaque != 0);
  int occ_t = 1, occ_s = 1;
  if (cfg->mode != 0) {
    if ((e = cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes)) != cudaSuccess) return e;
  }
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, tf, TRACE_THREADS, cfg->smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, sf, TRACE_THREADS, cfg->smem_bytes)) != cudaSuccess) return e;
  if (occ_t < 1) occ_t = 1;
  if (occ_s < 1) occ_s = 1;
  cfg->grid_trace = sms * occ_t;
  cfg->grid_shade = sms * occ_s;
  return cudaSuccess;
}

static inline int clamp_grid(int persistent, uint32_t work_bound) {
  uint64_t need = ((uint64_t)work_bound + 32 * (TRACE_THREADS / 32) - 1) / (32 * (TRACE_THREADS / 32));
  if (need < 1) need = 1;
  return (int)(need < (uint64_t)persistent ? need : (uint64_t)persistent);
}

void launch_trace(const LaunchCfg &cfg, const SceneView &sv, const TileMap &tm, uint32_t level, uint32_t bounces,
                  uint32_t px_base, uint32_t n_px, const RayRec *rays_in, RayRec *rays_out, ShadeRec *shade_out,
                  FrameCounters *ctr, const FrameTargets &fb, uint32_t *nlev, uint32_t work_bound, cudaStream_t st) {
  int grid = clamp_grid(cfg.grid_trace, work_bound);
  pick_trace(cfg.mode, sv.brute_force != 0)<<<grid, TRACE_THREADS, cfg.smem_bytes, st>>>(sv, tm, level, bounces, px_base, n_px, rays_in,
                                                                                         rays_out, shade_out, ctr, fb, nlev);
}

void launch_shade(const LaunchCfg &cfg, const SceneView &sv, uint32_t level, const ShadeRec *shade, FrameCounters *ctr,
                  const FrameTargets &fb, bool atomic_accumulate, float *level_color, uint32_t px_base, uint32_t work_bound,
                  cudaStream_t st) {
  int grid = clamp_grid(cfg.grid_shade, work_bound);
  pick_shade(cfg.mode, sv.brute_force != 0, sv.all_opaque != 0)<<<grid, TRACE_THREADS, cfg.smem_bytes, st>>>(
      sv, level, shade, ctr, fb, atomic_accumulate ? 1 : 0, level_color, px_base);
}

}  // namespace ctb
