// render.cu — the wavefront pipeline that replaces the reference's recursive one-thread-per-pixel
// kernel (inc/kernel.hpp:35-60 -> inc/shading.hpp:116-154 -> inc/ray_cast.hpp:29-55).
//
// Per bounce level L (0 = primary rays) there are two kinds of work:
//
//   trace(L)   closest hit for every ray of level L (L = 0: rays are generated from the pixel index,
//              cam::get_ray inc/default_schema.hpp:376-386).  Level 0 also writes the G-buffer
//              (depth / raw normal / object id, inc/kernel.hpp:52-56) and reduces the largest finite
//              depth (inc/kernel.hpp:120-125).  Every hit emits one ShadeRec; mirror / transparent
//              materials push child rays into the level L+1 queue (inc/shading.hpp:130-149),
//              compacted with warp ballot + popc into slot blocks a warp reserves with ONE atomicAdd.
//   shade(L)   per ShadeRec: all shadow rays (shadow_intensity, inc/shading.hpp:22-45) and the Phong
//              sum (inc/shading.hpp:64-99), stored with the path weight of the hit.
//
// The recursion  rgb = (1-t)*(phong + r*R) + t*T  (inc/shading.hpp:138,148) is unrolled into path
// weights: own Phong term w*(1-t), reflected child w*(1-t)*r, transmitted child w*t; at the last
// level (bounces exhausted) the blend is skipped exactly like `if constexpr(bounces != 0)`.
//
// Three schedulers drive the same per-ray device functions:
//
//   pixel_kernel   (default) ONE persistent kernel per frame, one 1024-thread CTA per SM: a thread walks a pixel's whole path — the
//                  recursion of ray_color as a loop with an explicit stack — and warps claim 32..128 pixels at a time from a global
//                  cursor, or (scenes walked through L1 / L2, big frames) from their CTA's own cursor over round-robin 16-tile
//                  chunks with stealing (claim_segment).  No queues at all.
//   frame_kernel   the wavefront as ONE persistent cooperative kernel per frame: the level loop runs on the device.  Phase p
//                  = trace(p); a warp that runs out of trace(p) work arrives at the phase barrier and, instead of
//                  spinning, shades records of the levels < p until the barrier opens (all warps arrived), so the
//                  tail of every trace level is filled with shading.  After the last phase: ordered per-pixel sum of
//                  the level images, G-buffer / colour stores to wherever the frame lives (own HBM, a peer GPU over
//                  NVLink, pinned host memory), counters published to mapped host memory and cleared for the next
//                  frame.  The scene is staged into shared memory once per CTA and frame (cp.async.bulk + mbarrier).
//   trace_kernel / shade_kernel   the wavefront as one launch per level and kind (CUTRACE_FLAG_LAUNCHES; CUTRACE_FLAG_SERIALIZE:
//                  per-kernel timings; also the fallback when a cooperative launch is not possible).
//
// All kernels are persistent: each warp claims work from a global cursor (guided chunk sizes).
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include "render.cuh"
#include "trace.cuh"

namespace ctb {

#define CTB_FULL 0xffffffffu

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

#ifdef CTB_PHASE_DEBUG   // developer instrumentation (tools/build_variant.sh dbg -DCTB_PHASE_DEBUG): where a phase's time goes
// [0] first / [1] last warp out of trace(p) work, [2] last warp arrived (after flush), [3] first / [4] last warp that saw the phase open,
// [5][0] last CTA staged
__device__ unsigned long long g_dbg_ns[6][18];
#define DBG_MIN(k, p) do { if (lane == 0) atomicMin(&g_dbg_ns[k][p], globaltimer_ns()); } while (0)
#define DBG_MAX(k, p) do { if (lane == 0) atomicMax(&g_dbg_ns[k][p], globaltimer_ns()); } while (0)
#else
#define DBG_MIN(k, p)
#define DBG_MAX(k, p)
#endif

// Work stealing with guided chunk sizes: a warp claims `remaining / (2 * warps in flight)` items, at least one
// warp-iteration (32) and at most `max_chunk`.  Large claims while there is plenty of work keep the cursor
// atomics rare; 32-item claims at the end keep the tail short — at deep bounce levels of the 10 M-triangle scene a
// few grazing rays cost milliseconds each and fixed 128-ray claims left the GPU idle behind them
// (profiles/r01_tuning.md, "tile scaling").
// `first` (warp-uniform; trace phases, which every warp enters at the same moment): the first chunk of a warp is its
// STATIC slot gw * c0 — no atomic at all; the cursor then hands out what lies behind warps * c0.  A bounce level with fewer
// than 32 rays per warp costs no cursor atomics whatsoever (4,736 warps hitting one address at the start of every phase took
// 3 us by themselves, profiles/r02_tuning.md).
__device__ __forceinline__ unsigned guided_chunk(unsigned remaining, unsigned warps, unsigned max_chunk) {
  unsigned chunk = remaining / (2u * warps);
  chunk = chunk < 32u ? 32u : (chunk > max_chunk ? max_chunk : chunk);
  return chunk & ~31u;
}
__device__ __forceinline__ bool claim_work(unsigned *cursor, unsigned n_work, unsigned lane, unsigned max_chunk, bool static_first, bool first,
                                           unsigned &base, unsigned &end) {
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  unsigned dyn0 = 0;
  if (static_first) {
    const unsigned c0 = guided_chunk(n_work, warps, max_chunk);
    if (first) {
      const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
      base = gw * c0;
      end = base + c0 < n_work ? base + c0 : n_work;
      return base < n_work;
    }
    dyn0 = warps * c0;
    if (dyn0 >= n_work) return false;
  }
  unsigned b = 0, chunk = 0;
  if (lane == 0) {
    const unsigned cur = dyn0 + *reinterpret_cast<volatile unsigned *>(cursor);
    if (cur < n_work) {
      chunk = guided_chunk(n_work - cur, warps, max_chunk);
      b = dyn0 + atomicAdd(cursor, chunk);
    } else {
      b = n_work;   // exhausted: no atomic (the frame kernel polls exhausted cursors while it waits for a phase to open)
    }
  }
  base = __shfl_sync(CTB_FULL, b, 0);
  chunk = __shfl_sync(CTB_FULL, chunk, 0);
  end = base + chunk < n_work ? base + chunk : n_work;
  return base < n_work;
}

// The pixel kernel's lane refill (pixel_kernel<.., REFILL>) claims for the `asked` free lanes of a warp, not for a whole warp: the
// same static first slot and the same guided sizes while there is plenty of work, but never rounded to 32 and never less than
// `asked` — at the end of a frame a warp takes exactly the pixels its free lanes can start at once (a 32-pixel claim for one free
// lane would be walked one pixel at a time while other warps have nothing left).
__device__ __forceinline__ bool claim_pixels(unsigned *cursor, unsigned n_work, unsigned lane, unsigned max_chunk, bool first, unsigned asked,
                                             unsigned &base, unsigned &end) {
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  const unsigned c0 = guided_chunk(n_work, warps, max_chunk);
  if (first) {
    const unsigned gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    base = gw * c0;
    end = base + c0 < n_work ? base + c0 : n_work;
    return base < n_work;
  }
  const unsigned dyn0 = warps * c0;
  if (dyn0 >= n_work) return false;
  unsigned b = 0, chunk = 0;
  if (lane == 0) {
    const unsigned cur = dyn0 + *reinterpret_cast<volatile unsigned *>(cursor);
    if (cur < n_work) {
      chunk = (n_work - cur) / (2u * warps);
      chunk = chunk > max_chunk ? max_chunk : chunk;
      chunk = chunk < asked ? asked : chunk;
      b = dyn0 + atomicAdd(cursor, chunk);
    } else {
      b = n_work;
    }
  }
  base = __shfl_sync(CTB_FULL, b, 0);
  chunk = __shfl_sync(CTB_FULL, chunk, 0);
  end = base + chunk < n_work ? base + chunk : n_work;
  return base < n_work;
}

// ---- pixel kernel: one work cursor per CTA ----------------------------------------------------------------------------------
// With ONE cursor for the whole grid, consecutive claims go to whichever warps ask next: the 32 warps of an SM work on 32 unrelated
// places of the frame and share nothing in L1 but the top of the BVH (10 M-triangle hall: L1 hit rate 74 %, long-scoreboard stalls
// 4.3 warp-cycles per issue, profiles/r02_pixel_kernel_synthetic10m.md).  Here the frame's work items are cut into chunks of
// CTB_SEG_CHUNK (4096 pixels = 16 tiles = one 128-pixel claim for each of a CTA's 32 warps) that are dealt to the CTAs
// round-robin — CTA k owns chunks k, k + G, k + 2G, ..: every CTA samples the whole frame, so the CTAs finish within a few per
// cent of each other — and the warps of a CTA claim from THEIR cursor: at any moment an SM works on one or two 16-tile chunks.
// A warp whose CTA has run dry steals 32 pixels at a time from the CTA with the most work left (the warp reads all cursors, one
// per lane and round).  `o` below is an offset in a CTA's own index space [0, seg_len(k)).
// claims [base, end) in the index space of CTA `owner` (warp-uniform results); false: the frame has no unclaimed work left
__device__ __forceinline__ bool claim_segment(FrameCounters *ctr, unsigned n_work, unsigned lane, unsigned max_chunk, bool first, unsigned &owner,
                                              unsigned &base, unsigned &end) {
  const unsigned G = gridDim.x, wpc = blockDim.x >> 5, k = blockIdx.x;
  const unsigned len = seg_len(n_work, k, G);
  const unsigned c0 = guided_chunk(len, wpc, max_chunk);
  owner = k;
  if (first) {   // static first slot: no atomic (see claim_work)
    base = (threadIdx.x >> 5) * c0;
    end = base + c0 < len ? base + c0 : len;
    return base < len;
  }
  const unsigned dyn0 = wpc * c0;
  unsigned b = 0xffffffffu, chunk = 0;
  if (dyn0 < len) {
    if (lane == 0) {
      const unsigned cur = dyn0 + *reinterpret_cast<volatile unsigned *>(&ctr->seg[k].v);
      if (cur < len) {
        chunk = guided_chunk(len - cur, wpc, max_chunk);
        b = dyn0 + atomicAdd(&ctr->seg[k].v, chunk);
      }
    }
    b = __shfl_sync(CTB_FULL, b, 0);
    chunk = __shfl_sync(CTB_FULL, chunk, 0);
    if (b < len) { base = b; end = b + chunk < len ? b + chunk : len; return true; }
  }
  // own segment used up: steal from the CTA with the most work left
  for (;;) {
    unsigned best_rem = 0, best_v = 0;
    for (unsigned v = lane; v < G; v += 32) {
      const unsigned lv = seg_len(n_work, v, G);
      const unsigned d0 = wpc * guided_chunk(lv, wpc, max_chunk);
      const unsigned cur = d0 + *reinterpret_cast<volatile unsigned *>(&ctr->seg[v].v);
      const unsigned rem = cur < lv ? lv - cur : 0u;
      if (rem > best_rem) { best_rem = rem; best_v = v; }
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
      const unsigned r2 = __shfl_xor_sync(CTB_FULL, best_rem, sft), v2 = __shfl_xor_sync(CTB_FULL, best_v, sft);
      if (r2 > best_rem || (r2 == best_rem && v2 < best_v)) { best_rem = r2; best_v = v2; }
    }
    if (best_rem == 0) return false;
    const unsigned lv = seg_len(n_work, best_v, G);
    const unsigned d0 = wpc * guided_chunk(lv, wpc, max_chunk);
    unsigned sb = 0;
    if (lane == 0) sb = d0 + atomicAdd(&ctr->seg[best_v].v, 32u);
    sb = __shfl_sync(CTB_FULL, sb, 0);
    if (sb < lv) { owner = best_v; base = sb; end = sb + 32u < lv ? sb + 32u : lv; return true; }
  }
}

// ---- scene staging (MODE 1: whole BVH + primitive store, MODE 2: top of the BVH) ------------------------------------------
// One thread arms an mbarrier with the byte count and issues bulk asynchronous copies global -> shared (cp.async.bulk,
// SASS UBLKCP): the copy engine of the SM moves the 71 KB of bunny.json's BVH while no thread spends issue slots on
// LDG/STS pairs (round 1 staged with 4,444 LDG.128 + STS.128 per CTA and launch).  All threads then wait on the barrier.
#define CTB_BULK_CHUNK 32768u
__device__ __forceinline__ void bulk_stage(float4 *smem, const float4 *g0, uint32_t bytes0, const float4 *g1, uint32_t bytes1, uint64_t *bar) {
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes0 + bytes1) : "memory");
    const char *src[2] = {reinterpret_cast<const char *>(g0), reinterpret_cast<const char *>(g1)};
    const uint32_t len[2] = {bytes0, bytes1};
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
#pragma unroll 1
    for (int k = 0; k < 2; k++) {
#pragma unroll 1
      for (uint32_t off = 0; off < len[k]; off += CTB_BULK_CHUNK) {
        const uint32_t sz = len[k] - off < CTB_BULK_CHUNK ? len[k] - off : CTB_BULK_CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src[k] + off),
                     "r"(sz), "r"(bar_s)
                     : "memory");
        dst += sz;
      }
    }
  }
  // every thread waits for phase 0 of the barrier (all bytes landed)
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_s) : "memory");
  }
}

template <int MODE>
__device__ __forceinline__ void stage_scene(const SceneView &sv, float4 *smem, const float4 *&nodes, const float4 *&prims) {
  const float4 *gn = reinterpret_cast<const float4 *>(sv.nodes);
  const float4 *gp = reinterpret_cast<const float4 *>(sv.prims);
  if (MODE == 1) {
    const uint32_t nb = sv.n_nodes * CTB_NODE_BYTES, pb = sv.n_prims * (uint32_t)sizeof(PrimRec);
    uint64_t *bar = reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(smem) + nb + pb);   // 16-byte aligned: both sizes are multiples of 16
    bulk_stage(smem, gn, nb, gp, pb, bar);
    nodes = smem;
    prims = smem + sv.n_nodes * CTB_NODE_F4;
  } else {
    if (MODE == 2) {   // top of the BVH (first smem_nodes nodes, breadth-first) -> shared memory
      const uint32_t nb = sv.smem_nodes * (uint32_t)sizeof(Node);
      uint64_t *bar = reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(smem) + nb);
      bulk_stage(smem, gn, nb, gn, 0u, bar);
    }
    nodes = gn;
    prims = gp;
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
  if (tm.wide_warps) {
    static_assert(CUTRACE_TILE_SHIFT == 4, "16 x 2 warps assume 16-pixel tile rows");
    px = l & 15u; py = (b << 1) + (l >> 4);
  }
  uint32_t tx, ty;
  if (!tile_of_slot(tm, lt * tm.world + tm.rank, tx, ty)) return false;
  x = tx * CUTRACE_TILE + px;
  y = ty * CUTRACE_TILE + py;
  pix = (lt << (2 * CUTRACE_TILE_SHIFT)) + (py << CUTRACE_TILE_SHIFT) + px;
  return x < tm.width && y < tm.height;
}

// cam::get_ray, inc/default_schema.hpp:376-386 (roundings pinned, see common.cuh)
__device__ __forceinline__ void camera_ray(const Camera &c, uint32_t x, uint32_t y, vec3 &o, vec3 &d) {
  const float aspect = CTB_DIV((float)c.w, (float)c.h);
  const float sx = CTB_MUL(CTB_SUB(CTB_DIV((float)x, (float)c.w), 0.5f), aspect);   // ((x / w) - 0.5) * aspect
  const float sy = CTB_SUB(0.5f, CTB_DIV((float)y, (float)c.h));                   // 0.5 - y / h
  // x_v + y_v + z_v with x_v = sx * right, y_v = sy * up, z_v = forward
  const vec3 v = mk3(CTB_ADD(CTB_FMA(sx, c.right.x, CTB_MUL(sy, c.up.x)), c.forward.x), CTB_ADD(CTB_FMA(sx, c.right.y, CTB_MUL(sy, c.up.y)), c.forward.y),
                     CTB_ADD(CTB_FMA(sx, c.right.z, CTB_MUL(sy, c.up.z)), c.forward.z));
  o = c.pos;
  d = vnormalized(v);
}

// ---- queue slots --------------------------------------------------------------------------------------------------------
// A warp reserves a block of queue slots with ONE atomicAdd and fills it over the next iterations (ballot + popc
// compaction inside the warp).  One same-address atomic WITH return per warp-iteration and queue had made the trace
// kernels wait on the L2 atomic unit (ncu r01: long_scoreboard 11.5 warp-cycles per issue, 45 % issue slots).  An emission
// that does not fit into the rest of the current block is SPLIT: its first records fill the block to the last slot, the
// others open the next block — so the only unused slots a warp ever leaves are the tail of its last block of a level
// (retired as holes, pix = CTB_HOLE, by flush_block).  Hence   reserved(L) <= valid(L) + warps * SLOT_BLOCK,
// which is exactly the slack the queues are allocated with (api.cu: alloc_frame).  Every reservation is nevertheless
// checked against the capacity: on overflow the emission is dropped, the part of the failed reservation that lies inside
// the queue is retired as holes (consumers never read unwritten slots) and FrameCounters.overflow makes the host fail the
// frame with CUTRACE_ERR_INTERNAL.
struct SlotBlock { unsigned next, end; };
struct Emit { unsigned base0, rem, base1; bool fits; };

__device__ __forceinline__ void store_hole(void *rec) {
  __stcs(reinterpret_cast<float4 *>(rec), make_float4(0.f, 0.f, 0.f, __uint_as_float(CTB_HOLE)));
}

template <typename Rec>
__device__ __forceinline__ Emit reserve_slots(SlotBlock &b, unsigned n, unsigned lane, unsigned slot_block, unsigned *counter, unsigned cap,
                                              Rec *queue, FrameCounters *ctr) {
  Emit e;
  e.base0 = b.next; e.rem = b.end - b.next; e.base1 = 0; e.fits = true;
  if (n <= e.rem) { b.next += n; return e; }
  const unsigned need = n - e.rem;
  const unsigned blk = slot_block > need ? slot_block : need;
  unsigned nb = 0;
  if (lane == 0) nb = atomicAdd(counter, blk);
  nb = __shfl_sync(CTB_FULL, nb, 0);
  if (nb > cap || blk > cap - nb) {                       // does not fit: never write past the queue
    if (lane == 0) atomicExch(&ctr->st.overflow, 1u);
    for (unsigned k = nb + lane; k < cap; k += 32) store_hole(queue + k);
    e.fits = false;
    b.next = b.end;                                       // the old block is full (its `rem` slots are used by this emission)
    return e;
  }
  e.base1 = nb;
  b.next = nb + need; b.end = nb + blk;
  return e;
}
__device__ __forceinline__ bool emit_ok(const Emit &e, unsigned rank) { return rank < e.rem || e.fits; }
__device__ __forceinline__ unsigned emit_slot(const Emit &e, unsigned rank) { return rank < e.rem ? e.base0 + rank : e.base1 + (rank - e.rem); }

template <typename Rec>
__device__ __forceinline__ void flush_block(SlotBlock &b, Rec *queue, unsigned lane) {
  for (unsigned k = b.next + lane; k < b.end; k += 32) store_hole(queue + k);
  b.next = b.end = 0;
}

// what one bounce level reads and writes
struct LevelIO {
  const RayRec *rays_in;
  RayRec *rays_out;
  ShadeRec *shade_out;
  unsigned ray_cap, shade_cap;   // capacity of rays_out / shade_out in records
};

struct TraceAcc {   // per-lane tallies of one level, reduced and added to the frame counters by trace_flush
  unsigned n_refl, n_trans, n_shaded;
  float max_depth;
  SlotBlock sq, rq;   // this warp's current slot blocks in the shade / ray queue (warp-uniform)
};
__device__ __forceinline__ void trace_acc_reset(TraceAcc &a) {
  a.n_refl = a.n_trans = a.n_shaded = 0u; a.max_depth = 0.f;
  a.sq.next = a.sq.end = a.rq.next = a.rq.end = 0u;
}

// closest hit + G-buffer + queue emission for the work items [base, end) of level `level` (one warp)
template <int MODE, bool BRUTE>
__device__ __forceinline__ void trace_chunk(const SceneView &sv, const float4 *nodes, const float4 *prims, const TileMap &tm, uint32_t level,
                                            uint32_t bounces, uint32_t px_base, unsigned base, unsigned end, unsigned n_work, const LevelIO &io,
                                            FrameCounters *ctr, const FrameTargets &fb, uint32_t *__restrict__ nlev, unsigned slot_block,
                                            unsigned lane, unsigned lt_mask, TraceAcc &acc) {
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
      const float4 *rp = reinterpret_cast<const float4 *>(io.rays_in + i);
      float4 a = __ldcg(rp), b = __ldcg(rp + 1);   // L2 only: the queue was written by other SMs earlier in this kernel
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
      if (hit && isfinite(h.t)) acc.max_depth = fmaxf(acc.max_depth, h.t);
      if (nlev && !hit) nlev[pix] = 0u;
    }
    // number of bounce levels that contributed to this pixel so far (levels run in order)
    if (nlev && hit) nlev[pix] = level + 1u;
    // inc/shading.hpp:126-149
    bool do_refl = false, do_trans = false;
    float w_own = w;
    if (hit && level < bounces) {
      do_refl = (double)reflect >= 1e-6;
      do_trans = (double)transp >= 1e-6;
      if (do_trans) w_own = w * (1.0f - transp);
    }
    // ---- shade record ----
    const unsigned m_hit = __ballot_sync(CTB_FULL, hit);
    if (m_hit) {
      const Emit e = reserve_slots(acc.sq, __popc(m_hit), lane, slot_block, &ctr->n_shade[level].v, io.shade_cap, io.shade_out, ctr);
      const unsigned rank = __popc(m_hit & lt_mask);
      if (hit && emit_ok(e, rank)) {
        float4 *sp = reinterpret_cast<float4 *>(io.shade_out + emit_slot(e, rank));
        __stcs(sp, make_float4(point.x, point.y, point.z, __uint_as_float(pix)));
        __stcs(sp + 1, make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(mat)));
        __stcs(sp + 2, make_float4(d.x, d.y, d.z, w_own));
        acc.n_shaded++;
      }
    }
    // ---- child rays ----
    const unsigned m_r = __ballot_sync(CTB_FULL, do_refl), m_t = __ballot_sync(CTB_FULL, do_trans);
    if (m_r | m_t) {
      const unsigned nr = __popc(m_r), nt = __popc(m_t);
      const Emit e = reserve_slots(acc.rq, nr + nt, lane, slot_block, &ctr->n_rays[level + 1].v, io.ray_cap, io.rays_out, ctr);
      const vec3 origin = vmad(o, d, h.t);   // incoming->start + distance * incoming->dir
      if (do_refl) {
        const unsigned rank = __popc(m_r & lt_mask);
        if (emit_ok(e, rank)) {
          vec3 nd = vnormalized(d), nn = vnormalized(nrm);
          vec3 rd = vreflect(nd, nn);
          float4 *rp = reinterpret_cast<float4 *>(io.rays_out + emit_slot(e, rank));
          __stcs(rp, make_float4(origin.x, origin.y, origin.z, __uint_as_float(pix)));
          __stcs(rp + 1, make_float4(rd.x, rd.y, rd.z, w_own * reflect));
          acc.n_refl++;
        }
      }
      if (do_trans) {
        const unsigned rank = nr + __popc(m_t & lt_mask);
        if (emit_ok(e, rank)) {
          float4 *rp = reinterpret_cast<float4 *>(io.rays_out + emit_slot(e, rank));
          __stcs(rp, make_float4(origin.x, origin.y, origin.z, __uint_as_float(pix)));
          __stcs(rp + 1, make_float4(d.x, d.y, d.z, w * transp));
          acc.n_trans++;
        }
      }
    }
  }
}

// end of a level for this warp: retire the unused tails of its slot blocks
__device__ __forceinline__ void trace_flush(const LevelIO &io, unsigned lane, TraceAcc &acc) {
  flush_block(acc.sq, io.shade_out, lane);
  flush_block(acc.rq, io.rays_out, lane);
}
// adds a warp's tallies to the frame statistics (once per kernel and warp)
__device__ __forceinline__ void tallies_flush(FrameCounters *ctr, unsigned lane, TraceAcc &acc, unsigned casts) {
  unsigned long long c = casts;
  for (int s = 16; s > 0; s >>= 1) {
    acc.n_refl += __shfl_xor_sync(CTB_FULL, acc.n_refl, s);
    acc.n_trans += __shfl_xor_sync(CTB_FULL, acc.n_trans, s);
    acc.n_shaded += __shfl_xor_sync(CTB_FULL, acc.n_shaded, s);
    acc.max_depth = fmaxf(acc.max_depth, __shfl_xor_sync(CTB_FULL, acc.max_depth, s));
    c += __shfl_xor_sync(CTB_FULL, c, s);
  }
  if (lane == 0) {
    if (acc.n_refl) atomicAdd(&ctr->st.rays_reflect, (unsigned long long)acc.n_refl);
    if (acc.n_shaded) atomicAdd(&ctr->st.shade_records, (unsigned long long)acc.n_shaded);
    if (acc.n_trans) atomicAdd(&ctr->st.rays_transmit, (unsigned long long)acc.n_trans);
    if (acc.max_depth > 0.f) atomicMax(&ctr->st.max_depth_bits, __float_as_uint(acc.max_depth));
    if (c) atomicAdd(&ctr->st.shadow_casts, c);
  }
}

// shadow_intensity, inc/shading.hpp:22-45
#ifndef CTB_LIGHTS_IN_ONE_WALK
#define CTB_LIGHTS_IN_ONE_WALK 0   // 1: big scenes (MODE 0 / 2) walk all shadow rays of a hit in one traversal loop (any_hit_lights, trace.cuh);
                                   // measured on the 10 M-triangle hall: 70.7 -> 85.0 ms per frame (profiles/r02_tuning.md), off
#endif
#ifndef SHADOW_PACKET
#define SHADOW_PACKET 1   // shadow rays of one hit that traverse together
#endif

template <int MODE, bool BRUTE, bool OPAQUE>
__device__ __forceinline__ float shadow_intensity(const SceneView &sv, const float4 *nodes, const float4 *prims, vec3 o, vec3 d,
                                                  float max_dist, unsigned &casts) {
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

// final += (1 - shadow_fac) * (fd * ld + fs * ls), inc/shading.hpp:95.  Written with explicit roundings — fma(fd, ld, fs * ls),
// then fma(keep, ., final), the forms nvcc's contraction picks for the reference's expression — because the compiler's own
// choice depends on the surrounding code: the same source inlined into the frame kernel and into shade_kernel differed in the
// last bit of 0.4 % of the pixels (profiles/r02_tuning.md), and the schedulers are supposed to agree bit for bit.
__device__ __forceinline__ void phong_add(vec3 &final, vec3 ld, float fd, vec3 ls, float fs, float keep) {
  final.x = __fmaf_rn(keep, __fmaf_rn(fd, ld.x, __fmul_rn(fs, ls.x)), final.x);
  final.y = __fmaf_rn(keep, __fmaf_rn(fd, ld.y, __fmul_rn(fs, ls.y)), final.y);
  final.z = __fmaf_rn(keep, __fmaf_rn(fd, ld.z, __fmul_rn(fs, ls.z)), final.z);
}

// A light whose diffuse AND specular factors are exactly zero at this hit adds keep * (0 * ld + 0 * ls) = +0 to the sum whatever
// its shadow ray finds (inc/shading.hpp:95: `final` stays bit for bit what it was): the ray is not traced.  fd = max(0, n.l) is
// zero for every light behind the surface; fs = pow(max(0, n.h), phong) is then zero unless the half vector still leans towards
// the normal by more than the underflow floor of the exponent (phong_pow).  Roughly half of a closed mesh faces away from any
// given light.  The query still counts in shadow_casts (the statistics describe the reference's rays, not this path's shortcuts).
// All-opaque scenes only: a march through translucent surfaces is several casts in the reference's count, and mirror.json (lights
// behind many of its surfaces, low exponents: the test runs but rarely skips) got 19 % slower with it (profiles/r02_tuning.md 8).
#ifndef CTB_SKIP_DARK_LIGHTS
#define CTB_SKIP_DARK_LIGHTS 1
#endif
__device__ __forceinline__ bool light_is_dark(vec3 nn, vec3 in_n, vec3 nd, float phong_exp, float pow_floor) {
  if (!CTB_SKIP_DARK_LIGHTS) return false;
  if (fmaxf(0.0f, vdot(nn, nd)) != 0.0f) return false;
  const vec3 hv = vnormalized(vadd(in_n, nd));
  return phong_pow(fmaxf(0.0f, vdot(nn, hv)), phong_exp, pow_floor) == 0.0f;
}

// phong (inc/shading.hpp:64-99) of one shaded hit: all its shadow rays (shadow_intensity, :22-45) and the sum over the lights
template <int MODE, bool BRUTE, bool OPAQUE>
__device__ __forceinline__ vec3 phong_record(const SceneView &sv, const float4 *nodes, const float4 *prims, vec3 hit, vec3 normal, vec3 in_dir,
                                             uint32_t mat, unsigned &casts) {
  // phong, inc/shading.hpp:64-99
  const float4 *mp = reinterpret_cast<const float4 *>(sv.materials + mat);
  const float4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
  const vec3 diffuse = mk3(m0.x, m0.y, m0.z);
  const vec3 specular = vscale(diffuse, m0.w);        // *spec = specular * color
  const float phong_exp = m1.y, pow_floor = phong_pow_floor(phong_exp);
  vec3 final = vscale(diffuse, sv.cam.ambient);
  const vec3 nn = vnormalized(normal);
  const vec3 in_n = vscale(vnormalized(in_dir), -1.0f);
  if (OPAQUE && CTB_LIGHTS_IN_ONE_WALK && MODE != 1 && !BRUTE && SHADOW_PACKET == 1 && !CTB_BVH4 && sv.n_lights <= (uint32_t)CTB_MULTI_LIGHTS) {
    // big scenes: the shadow rays of this hit share one traversal loop, a lane moves on to its next light as soon as its
    // ray has ended (any_hit_lights, trace.cuh).  Ray set-up and the Phong sums run before / after it with all lanes.
    float dx[CTB_MULTI_LIGHTS], dy[CTB_MULTI_LIGHTS], dz[CTB_MULTI_LIGHTS], len[CTB_MULTI_LIGHTS];
    unsigned todo = 0;
#pragma unroll 1
    for (uint32_t l = 0; l < sv.n_lights; l++) {
      const float4 l0v = __ldg(reinterpret_cast<const float4 *>(sv.lights + l));
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
      const vec3 sdir = vnormalized(direction);
      const float md = distance * vnorm(direction);
      dx[l] = sdir.x; dy[l] = sdir.y; dz[l] = sdir.z; len[l] = md;
      if (!planes_occlude(sv, hit, sdir, md)) todo |= 1u << l;
    }
    casts += sv.n_lights;
    const unsigned lit = todo & ~any_hit_lights<MODE>(sv, nodes, prims, hit, dx, dy, dz, len, todo);
#pragma unroll 1
    for (uint32_t l = 0; l < sv.n_lights; l++) {
      if (lit & (1u << l)) {        // shadow_fac = 0 < 1
        const float4 l1v = __ldg(reinterpret_cast<const float4 *>(sv.lights + l) + 1);
        const vec3 lcol = mk3(l1v.x, l1v.y, l1v.z);
        const vec3 nd = mk3(dx[l], dy[l], dz[l]);
        float fd = fmaxf(0.0f, vdot(nn, nd));
        vec3 ld = vmul(diffuse, lcol);
        vec3 hv = vnormalized(vadd(in_n, nd));
        float fs = phong_pow(fmaxf(0.0f, vdot(nn, hv)), phong_exp, pow_floor);
        vec3 ls = vmul(specular, lcol);
        phong_add(final, ld, fd, ls, fs, 1.0f);
      }
    }
  } else if (OPAQUE) {
    // every material is opaque: the shadow march saturates on its first step, so the K shadow rays
    // of this hit are any-hit queries and walk the BVH as one packet
    const unsigned plane_maybe = CTB_PLANE_PREPASS ? plane_side_prepass(sv, hit) : 0xffffffffu;
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
      if (SHADOW_PACKET == 1 && valid && light_is_dark(nn, in_n, sd[0], phong_exp, pow_floor)) continue;
      const unsigned occ = any_hit_packet<MODE, SHADOW_PACKET, BRUTE>(sv, nodes, prims, hit, sd, md, valid,
                                                                      SHADOW_PACKET > 1 || l0 >= 32u || ((plane_maybe >> l0) & 1u));
#pragma unroll
      for (int k = 0; k < SHADOW_PACKET; k++) {
        if ((valid & ~occ) & (1u << k)) {        // shadow_fac = 0 < 1
          const vec3 nd = sd[k];
          float fd = fmaxf(0.0f, vdot(nn, nd));
          vec3 ld = vmul(diffuse, lcol[k]);
          vec3 hv = vnormalized(vadd(in_n, nd));
          float fs = phong_pow(fmaxf(0.0f, vdot(nn, hv)), phong_exp, pow_floor);
          vec3 ls = vmul(specular, lcol[k]);
          phong_add(final, ld, fd, ls, fs, 1.0f);
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
        float fs = phong_pow(fmaxf(0.0f, vdot(nn, hv)), phong_exp, pow_floor);
        vec3 ls = vmul(specular, color);
        phong_add(final, ld, fd, ls, fs, 1 - shadow_fac);
      }
    }
  }
  return final;
}

// shadow rays + Phong for the shade records [base, end) of one level (one warp)
template <int MODE, bool BRUTE, bool OPAQUE>
__device__ __forceinline__ void shade_chunk(const SceneView &sv, const float4 *nodes, const float4 *prims, unsigned base, unsigned end,
                                            unsigned n_work, const ShadeRec *__restrict__ shade, const FrameTargets &fb, int atomic_accumulate,
                                            float *__restrict__ level_color, uint32_t px_base, unsigned lane, unsigned &casts) {
#pragma unroll 1
  for (unsigned off = 0; base + off < end; off += 32) {
    const uint32_t i = base + off + lane;
    if (i < n_work) {
      const float4 *sp = reinterpret_cast<const float4 *>(shade + i);
      const float4 s0 = __ldcg(sp), s1 = __ldcg(sp + 1), s2 = __ldcg(sp + 2);
      const vec3 hit = mk3(s0.x, s0.y, s0.z), normal = mk3(s1.x, s1.y, s1.z), in_dir = mk3(s2.x, s2.y, s2.z);
      const uint32_t pix = __float_as_uint(s0.w), mat = __float_as_uint(s1.w);
      const float weight = s2.w;
      if (pix == CTB_HOLE) continue;   // retired tail of a producer warp's slot block
      const vec3 final = phong_record<MODE, BRUTE, OPAQUE>(sv, nodes, prims, hit, normal, in_dir, mat, casts);
      if (level_color) {   // one hit per pixel and level: plain store into this level's partial image
        float *lp = level_color + 3 * (size_t)(pix - px_base);
        lp[0] = weight * final.x; lp[1] = weight * final.y; lp[2] = weight * final.z;
        continue;
      }
      // a material reflects AND transmits: several hits of one pixel can sit in one level -> float atomics into the local accumulator
      float *cp = fb.color + 3 * (size_t)pix;
      atomicAdd(cp, weight * final.x); atomicAdd(cp + 1, weight * final.y); atomicAdd(cp + 2, weight * final.z);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// the two kinds of work as callable units
// -------------------------------------------------------------------------------------------------
// trace_chunk / shade_chunk are entered through these two functions from the per-level kernels AND from the frame kernel.
// nvcc chooses FMA contractions (and everything else) per compilation context: the same source inlined into two kernels gives
// hit points and colours that differ in the last bit (0.4 % of the pixels, profiles/r02_tuning.md).  Built with
// -DCTB_SHARED_BODIES=1 the two functions are __noinline__: one compiled body per kind of work, every scheduler executes the
// same SASS for a ray and all of them agree bit for bit by construction (checked: identical md5 of the colour image) — but the
// calling convention (SceneView through a generic pointer instead of constant-bank operands, registers saved around the call)
// costs 10-12 % of the frame (bunny.json 4K 9.83 -> 11.05 ms), so the product build inlines them and the schedulers are allowed
// to differ in the last bits (tests/parity.py: assert_same_frame(color_ulps=...)).
#ifndef CTB_SHARED_BODIES
#define CTB_SHARED_BODIES 0
#endif
#if CTB_SHARED_BODIES
#define CTB_CHUNK_LINKAGE __noinline__
#else
#define CTB_CHUNK_LINKAGE __forceinline__
#endif
struct TraceCall {
  const SceneView *sv;   // kernel parameter space (__grid_constant__)
  const TileMap *tm;
  uint32_t level, bounces, px_base;
  unsigned base, end, n_work, slot_block;
  LevelIO io;
  FrameCounters *ctr;
  FrameTargets fb;
  uint32_t *nlev;
};
struct ShadeCall {
  const SceneView *sv;
  unsigned base, end, n_work;
  const ShadeRec *shade;
  FrameTargets fb;
  int atomic_accumulate;
  float *level_color;
  uint32_t px_base;
};

// node / primitive pointers as the staged kernels see them (MODE 1: the shared-memory copy made by stage_scene)
template <int MODE>
__device__ __forceinline__ void scene_ptrs(const SceneView &sv, const float4 *&nodes, const float4 *&prims) {
  extern __shared__ float4 ctb_dyn_smem[];
  if (MODE == 1) {
    nodes = ctb_dyn_smem;
    prims = ctb_dyn_smem + sv.n_nodes * CTB_NODE_F4;
  } else {
    nodes = reinterpret_cast<const float4 *>(sv.nodes);
    prims = reinterpret_cast<const float4 *>(sv.prims);
  }
}

template <int MODE, bool BRUTE>
__device__ CTB_CHUNK_LINKAGE void trace_chunk_fn(const TraceCall *c, TraceAcc *acc_io) {
  const SceneView &sv = *c->sv;
  const float4 *nodes, *prims;
  scene_ptrs<MODE>(sv, nodes, prims);
  TraceAcc acc = *acc_io;
  trace_chunk<MODE, BRUTE>(sv, nodes, prims, *c->tm, c->level, c->bounces, c->px_base, c->base, c->end, c->n_work, c->io, c->ctr, c->fb, c->nlev,
                           c->slot_block, threadIdx.x & 31u, lanemask_lt(), acc);
  *acc_io = acc;
}

template <int MODE, bool BRUTE, bool OPAQUE>
__device__ CTB_CHUNK_LINKAGE unsigned shade_chunk_fn(const ShadeCall *c) {
  const SceneView &sv = *c->sv;
  const float4 *nodes, *prims;
  scene_ptrs<MODE>(sv, nodes, prims);
  unsigned casts = 0;
  shade_chunk<MODE, BRUTE, OPAQUE>(sv, nodes, prims, c->base, c->end, c->n_work, c->shade, c->fb, c->atomic_accumulate, c->level_color, c->px_base,
                                   threadIdx.x & 31u, casts);
  return casts;
}

template <int MODE, bool BRUTE>
__global__ void __launch_bounds__(TRACE_THREADS, CTB_MIN_BLOCKS)
trace_kernel(const __grid_constant__ SceneView sv, const __grid_constant__ TileMap tm, uint32_t level, uint32_t bounces, uint32_t px_base,
             uint32_t n_px, const LevelIO io, FrameCounters *ctr, FrameTargets fb, uint32_t *__restrict__ nlev) {
  extern __shared__ float4 smem[];
  const float4 *nodes, *prims;
  stage_scene<MODE>(sv, smem, nodes, prims);
  const unsigned lane = threadIdx.x & 31u;
  uint32_t n_work = level == 0 ? n_px : ctr->n_rays[level].v;
  if (level && n_work > io.ray_cap) n_work = io.ray_cap;   // (only after an overflow) rays_in has the capacity of rays_out
  // block size: 1/8 of what a warp is expected to emit over the kernel, CTB_SLOT_MIN..SLOT_BLOCK
  unsigned slot_block = (n_work / (gridDim.x * (blockDim.x >> 5) * (unsigned)CTB_SLOT_DIV)) & ~31u;
  slot_block = slot_block < (unsigned)CTB_SLOT_MIN ? (unsigned)CTB_SLOT_MIN : (slot_block > (unsigned)SLOT_BLOCK ? (unsigned)SLOT_BLOCK : slot_block);
  TraceAcc acc;
  trace_acc_reset(acc);
  TraceCall tc;
  tc.sv = &sv; tc.tm = &tm; tc.level = level; tc.bounces = bounces; tc.px_base = px_base; tc.n_work = n_work; tc.slot_block = slot_block;
  tc.io = io; tc.ctr = ctr; tc.fb = fb; tc.nlev = nlev;
  for (bool first = true;; first = false) {
    if (!claim_work(&ctr->work_trace[level].v, n_work, lane, (unsigned)WORK_CHUNK_MAX, true, first, tc.base, tc.end)) { if (first) continue; break; }
    trace_chunk_fn<MODE, BRUTE>(&tc, &acc);
  }
  trace_flush(io, lane, acc);
  tallies_flush(ctr, lane, acc, 0u);
}

template <int MODE, bool BRUTE, bool OPAQUE>
__global__ void __launch_bounds__(TRACE_THREADS, CTB_SHADE_MIN_BLOCKS)
shade_kernel(const __grid_constant__ SceneView sv, uint32_t level, const ShadeRec *__restrict__ shade, uint32_t shade_cap, FrameCounters *ctr,
             FrameTargets fb, int atomic_accumulate, float *__restrict__ level_color, uint32_t px_base) {
  extern __shared__ float4 smem[];
  const float4 *nodes, *prims;
  stage_scene<MODE>(sv, smem, nodes, prims);
  const unsigned lane = threadIdx.x & 31u;
  uint32_t n_work = ctr->n_shade[level].v;
  if (n_work > shade_cap) n_work = shade_cap;
  unsigned casts = 0;
  ShadeCall sc;
  sc.sv = &sv; sc.n_work = n_work; sc.shade = shade; sc.fb = fb; sc.atomic_accumulate = atomic_accumulate; sc.level_color = level_color;
  sc.px_base = px_base;
  for (bool first = true;; first = false) {
    if (!claim_work(&ctr->work_shade[level].v, n_work, lane, (unsigned)WORK_CHUNK_MAX, true, first, sc.base, sc.end)) { if (first) continue; break; }
    casts += shade_chunk_fn<MODE, BRUTE, OPAQUE>(&sc);
  }
  TraceAcc none;
  trace_acc_reset(none);
  tallies_flush(ctr, lane, none, casts);
}

// -------------------------------------------------------------------------------------------------
// frame assembly
// -------------------------------------------------------------------------------------------------
// colour of local pixel `pix` = sum of its per-level partial images in level order (same order as a serial accumulation),
// stored where the frame lives: tile-major local buffer, own row-major frame, a peer GPU's frame over NVLink, or pinned
// host memory.  gsrc.depth != NULL: the level-0 G-buffer (tile-major, local) travels with it.
__device__ __forceinline__ void combine_pixel(const TileMap &tm, uint32_t i, uint32_t px_base, const uint32_t *__restrict__ nlev,
                                              const float *__restrict__ level_color, uint64_t level_stride, uint32_t levels,
                                              const float *__restrict__ local_color, const FrameTargets &out, const FrameTargets &gsrc) {
  const uint32_t pix = px_base + i;
  uint32_t x, y;
  if (!pixel_of_local(tm, pix, x, y)) return;
  float r = 0.f, g = 0.f, b = 0.f;
  if (levels == 0) {
    const float *p = local_color + 3 * (size_t)pix;
    r = __ldcg(p); g = __ldcg(p + 1); b = __ldcg(p + 2);
  } else {
    uint32_t n = __ldcg(nlev + pix);
    n = n < levels ? n : levels;
    for (uint32_t l = 0; l < n; l++) {
      const float *p = level_color + l * level_stride + 3 * (size_t)i;
      r += __ldcg(p); g += __ldcg(p + 1); b += __ldcg(p + 2);
    }
  }
  const size_t gi = out.row_major ? (size_t)y * tm.width + x : (size_t)pix;
  float *o = out.color + 3 * gi;
  o[0] = r; o[1] = g; o[2] = b;
  if (gsrc.depth) {
    out.depth[gi] = __ldcg(gsrc.depth + pix);
    out.hit_id[gi] = __ldcg(gsrc.hit_id + pix);
    out.normal[3 * gi] = __ldcg(gsrc.normal + 3 * (size_t)pix); out.normal[3 * gi + 1] = __ldcg(gsrc.normal + 3 * (size_t)pix + 1);
    out.normal[3 * gi + 2] = __ldcg(gsrc.normal + 3 * (size_t)pix + 2);
  }
}

// forwards the level-0 G-buffer of one local pixel (tile-major, local) to a row-major frame elsewhere: one 16-pixel tile row
// per half warp -> 64 / 192-byte contiguous NVLink (or PCIe) stores
__device__ __forceinline__ void export_pixel(const TileMap &tm, uint32_t pix, const FrameTargets &src, const FrameTargets &out) {
  uint32_t x, y;
  if (!pixel_of_local(tm, pix, x, y)) return;
  const size_t gi = (size_t)y * tm.width + x;
  out.depth[gi] = __ldcg(src.depth + pix);
  out.hit_id[gi] = __ldcg(src.hit_id + pix);
  out.normal[3 * gi] = __ldcg(src.normal + 3 * (size_t)pix); out.normal[3 * gi + 1] = __ldcg(src.normal + 3 * (size_t)pix + 1);
  out.normal[3 * gi + 2] = __ldcg(src.normal + 3 * (size_t)pix + 2);
}

__global__ void combine_levels_kernel(const TileMap tm, const uint32_t *__restrict__ nlev, const float *__restrict__ level_color,
                                      uint64_t level_stride, uint32_t levels, const float *__restrict__ local_color,
                                      uint32_t px_base, uint32_t n_px, FrameTargets out, FrameTargets gsrc) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px) return;
  combine_pixel(tm, i, px_base, nlev, level_color, level_stride, levels, local_color, out, gsrc);
}

// Runs on an auxiliary stream right after trace(0) (multi-launch path), i.e. the transfer overlaps the remaining levels.
__global__ void export_gbuffer_kernel(const TileMap tm, uint32_t px_base, uint32_t n_px, FrameTargets src, FrameTargets out) {
  // a SMALL grid-stride grid: the kernel is bound by the NVLink stores (all ranks push into rank 0 at the same moment),
  // and every resident CTA of it takes SM slots away from the persistent trace/shade CTAs it is supposed to run under
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) export_pixel(tm, px_base + i, src, out);
}

// -------------------------------------------------------------------------------------------------
// the persistent frame kernel
// -------------------------------------------------------------------------------------------------
#define CTB_EXPORT_CHUNK 256u

// Grid barrier of the frame kernel, two-level.  Warps of a CTA count themselves in shared memory; the LAST warp of a CTA
// to arrive publishes the CTA to the global counter (one release-RMW per CTA and phase instead of one per warp) and then
// becomes the CTA's poller: it alone spins on the global counter (ld.acquire.gpu) and raises a shared-memory flag when
// every CTA has arrived.  The other warps keep doing filler work and only look at the shared flag between chunks.
//   fence protocol: writer warp: stores ... __threadfence(); smem arrive.   last warp: smem arrive; __threadfence();
//   red.release.gpu.  poller: ld.acquire.gpu; __threadfence(); smem flag.  reader: smem flag; __threadfence(); loads (.cg).
struct CtaSync {
  unsigned arrived;              // warps of this CTA that have arrived, monotonic over the phases
  volatile unsigned open;        // phases known to be open, monotonic
  unsigned long long refl, trans, shaded, casts;   // CTA-wide tallies, added to the frame statistics once per CTA
  unsigned max_depth_bits;
};

template <int MODE, bool BRUTE, bool OPAQUE>
__global__ void __launch_bounds__(TRACE_THREADS, CTB_MIN_BLOCKS) frame_kernel(const __grid_constant__ FrameArgs a) {
  extern __shared__ float4 smem[];
  __shared__ CtaSync cs;
  if (threadIdx.x == 0) { cs.arrived = 0u; cs.open = 0u; cs.refl = cs.trans = cs.shaded = cs.casts = 0ull; cs.max_depth_bits = 0u; }
  const float4 *nodes, *prims;
  stage_scene<MODE>(a.sv, smem, nodes, prims);
  if (MODE == 0) __syncthreads();   // (the staging paths synchronise the CTA themselves)
  const unsigned lane = threadIdx.x & 31u;
  const unsigned wpc = blockDim.x >> 5;
  FrameCounters *ctr = a.ctr;
  if (blockIdx.x == 0 && threadIdx.x == 0) ctr->st.phase_ns[17] = globaltimer_ns();   // diagnostics: start of the frame on the device
  DBG_MAX(5, 0);
  unsigned casts = 0;
  unsigned shade_lo = 0;                      // lowest shade level this warp still expects unclaimed records in
  bool export_left = a.gsrc.depth != nullptr && !a.export_with_color;   // G-buffer export to a remote frame pending
  unsigned k_phase = 0;                       // barriers passed so far
  TraceAcc acc;
  trace_acc_reset(acc);
  TraceCall tc;
  tc.sv = &a.sv; tc.tm = &a.tm; tc.bounces = a.bounces; tc.px_base = a.px_base; tc.ctr = ctr; tc.fb = a.gbuf; tc.nlev = a.nlev;
  ShadeCall sc;
  sc.sv = &a.sv; sc.fb = a.acc; sc.atomic_accumulate = a.atomic_accumulate; sc.px_base = a.px_base;

  // arrive at barrier number k_phase (global counter `gctr`); returns true for the warp that became the CTA's poller.
  // A grid of one CTA (tiny frames) never touches the global counter.
  auto cta_arrive = [&](unsigned *gctr) -> bool {
    __threadfence();
    __syncwarp();
    unsigned last = 0;
    if (lane == 0) {
      last = atomicAdd(&cs.arrived, 1u) == (k_phase + 1u) * wpc - 1u;
      if (last && gridDim.x > 1) red_release_add(gctr, 1u);   // release: orders the whole CTA's writes (fence cumulativity) before the count
    }
    return __shfl_sync(CTB_FULL, last, 0) != 0;
  };
  auto poll_until_open = [&](unsigned *gctr) {
    if (lane == 0) {
      if (gridDim.x > 1) while (ld_acquire_u32(gctr) < gridDim.x) __nanosleep(100);
      __threadfence_block();
      cs.open = k_phase + 1u;
    }
    __syncwarp();
  };
  for (uint32_t p = a.first_level; p <= a.levels; p++) {
    bool poller = false;
    // ---- trace(p): the critical path, every warp drains it first ----
    if (p < a.levels) {
      LevelIO io;
      io.rays_in = a.rays[p & 1]; io.rays_out = a.rays[(p + 1) & 1]; io.shade_out = a.shade[p];
      io.ray_cap = a.ray_cap; io.shade_cap = a.shade_cap[p];
      uint32_t n_work = a.n_px;
      if (p) { n_work = __ldcg(&ctr->n_rays[p].v); if (n_work > a.ray_cap) n_work = a.ray_cap; }
      unsigned slot_block = (n_work / (gridDim.x * wpc * (unsigned)CTB_SLOT_DIV)) & ~31u;
      slot_block = slot_block < (unsigned)CTB_SLOT_MIN ? (unsigned)CTB_SLOT_MIN : (slot_block > (unsigned)SLOT_BLOCK ? (unsigned)SLOT_BLOCK : slot_block);
      tc.level = p; tc.n_work = n_work; tc.slot_block = slot_block; tc.io = io;
      // a level with at most two warp-iterations per warp is dealt out statically (no cursor atomics)
      const bool tiny = n_work <= 64u * gridDim.x * wpc;
      bool first = tiny;
      for (;;) {
        unsigned base, end;
        if (!claim_work(&ctr->work_trace[p].v, n_work, lane, (unsigned)WORK_CHUNK_MAX, tiny, first, base, end)) {
          if (first) { first = false; continue; }
          break;
        }
        first = false;
        tc.base = base; tc.end = end;
        trace_chunk_fn<MODE, BRUTE>(&tc, &acc);
      }
      DBG_MIN(0, p); DBG_MAX(1, p);
      trace_flush(io, lane, acc);
      poller = cta_arrive(&ctr->arrive[p].v);   // everything this warp wrote for level p (queue records, holes, counters) is released
      DBG_MAX(2, p);
      if (poller) poll_until_open(&ctr->arrive[p].v);
    }
    // ---- while the phase is still closed (other warps are tracing): fill the time with work that is off the critical path ----
    for (;;) {
      if (p < a.levels && (poller || cs.open > k_phase)) break;
      if (export_left && p >= 1) {     // G-buffer of this rank's tiles -> remote frame, under the remaining levels
        unsigned b = 0;
        if (lane == 0) b = atomicAdd(&ctr->work_export.v, CTB_EXPORT_CHUNK);
        b = __shfl_sync(CTB_FULL, b, 0);
        if (b < a.n_px) {
          const unsigned e = b + CTB_EXPORT_CHUNK < a.n_px ? b + CTB_EXPORT_CHUNK : a.n_px;
          for (unsigned i = b + lane; i < e; i += 32) export_pixel(a.tm, a.px_base + i, a.gsrc, a.out);
          continue;
        }
        export_left = false;
      }
      if (shade_lo < p) {              // records of levels < p are complete
        uint32_t n_sw = 0;
        if (lane == 0) { n_sw = __ldcg(&ctr->n_shade[shade_lo].v); if (n_sw > a.shade_cap[shade_lo]) n_sw = a.shade_cap[shade_lo]; }
        n_sw = __shfl_sync(CTB_FULL, n_sw, 0);
        unsigned base, end;
        if (!claim_work(&ctr->work_shade[shade_lo].v, n_sw, lane, p < a.levels ? (unsigned)CTB_FILL_CHUNK_MAX : (unsigned)WORK_CHUNK_MAX, false, false,
                        base, end)) {
          shade_lo++;
          continue;
        }
        sc.base = base; sc.end = end; sc.n_work = n_sw; sc.shade = a.shade[shade_lo];
        sc.level_color = a.level_color ? a.level_color + (size_t)shade_lo * a.level_stride : nullptr;
        casts += shade_chunk_fn<MODE, BRUTE, OPAQUE>(&sc);
        continue;
      }
      if (p == a.levels) break;        // last phase: nothing left to claim
      __nanosleep(100);                // nothing to fill with: wait for the CTA's poller to raise the flag
    }
    if (p < a.levels) {
      __threadfence_block();           // reader side of the fence protocol (queue reads that follow are .cg: served by L2)
      k_phase++;
      DBG_MIN(3, p); DBG_MAX(4, p);
      if (lane == 0 && blockIdx.x == 0 && threadIdx.x == 0) ctr->st.phase_ns[p] = globaltimer_ns();
    }
  }
  // ---- all records are claimed; CTA-wide tallies, then wait until every warp of the grid has finished shading ----
  {
    unsigned long long c = casts;
    for (int s = 16; s > 0; s >>= 1) {
      acc.n_refl += __shfl_xor_sync(CTB_FULL, acc.n_refl, s);
      acc.n_trans += __shfl_xor_sync(CTB_FULL, acc.n_trans, s);
      acc.n_shaded += __shfl_xor_sync(CTB_FULL, acc.n_shaded, s);
      acc.max_depth = fmaxf(acc.max_depth, __shfl_xor_sync(CTB_FULL, acc.max_depth, s));
      c += __shfl_xor_sync(CTB_FULL, c, s);
    }
    if (lane == 0) {
      if (acc.n_refl) atomicAdd(&cs.refl, (unsigned long long)acc.n_refl);
      if (acc.n_trans) atomicAdd(&cs.trans, (unsigned long long)acc.n_trans);
      if (acc.n_shaded) atomicAdd(&cs.shaded, (unsigned long long)acc.n_shaded);
      if (c) atomicAdd(&cs.casts, c);
      if (acc.max_depth > 0.f) atomicMax(&cs.max_depth_bits, __float_as_uint(acc.max_depth));
    }
  }
  {
    // the CTA's last warp adds the CTA's tallies to the frame statistics BEFORE it publishes the CTA: whoever sees the
    // final barrier open also sees complete statistics
    __threadfence();
    __syncwarp();
    unsigned last = 0;
    if (lane == 0) {
      last = atomicAdd(&cs.arrived, 1u) == (k_phase + 1u) * wpc - 1u;
      if (last) {
        __threadfence_block();
        if (cs.refl) atomicAdd(&ctr->st.rays_reflect, cs.refl);
        if (cs.trans) atomicAdd(&ctr->st.rays_transmit, cs.trans);
        if (cs.shaded) atomicAdd(&ctr->st.shade_records, cs.shaded);
        if (cs.casts) atomicAdd(&ctr->st.shadow_casts, cs.casts);
        if (cs.max_depth_bits) atomicMax(&ctr->st.max_depth_bits, cs.max_depth_bits);
        if (gridDim.x > 1) red_release_add(&ctr->arrive[a.levels].v, 1u);
      }
    }
    last = __shfl_sync(CTB_FULL, last, 0);
    if (last) poll_until_open(&ctr->arrive[a.levels].v);
    else if (lane == 0) { while (cs.open <= k_phase) __nanosleep(100); }
    __syncwarp();
    __threadfence_block();
    k_phase++;
  }
  // ---- frame assembly: ordered sum of the level images (+ G-buffer) -> the frame ----
  if (a.combine) {
    const FrameTargets g = a.export_with_color ? a.gsrc : FrameTargets{};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_px; i += gridDim.x * blockDim.x)
      combine_pixel(a.tm, i, a.px_base, a.nlev, a.level_color, a.level_stride, a.combine_levels, a.local_color, a.out, g);
  }
  // ---- the last CTA out publishes the statistics to mapped host memory and clears all counters for the next frame ----
  __threadfence();
  __syncthreads();
  __shared__ unsigned s_last_cta;
  if (threadIdx.x == 0) s_last_cta = atomicAdd(&ctr->finished.v, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (s_last_cta) {
    __threadfence();
    if (threadIdx.x == 0) ctr->st.phase_ns[a.levels] = globaltimer_ns();
    __syncthreads();
    unsigned *src = reinterpret_cast<unsigned *>(ctr);
    constexpr unsigned NW = sizeof(FrameCounters) / 4, S0 = offsetof(FrameCounters, st) / 4, SN = sizeof(FrameStats) / 4;
    volatile unsigned *dst = reinterpret_cast<volatile unsigned *>(a.host_stats);
    if (dst && threadIdx.x < SN) dst[threadIdx.x] = __ldcg(src + S0 + threadIdx.x);
    __syncthreads();
    for (unsigned k = threadIdx.x; k < NW; k += blockDim.x) src[k] = 0u;
    __threadfence_system();
  }
}

// -------------------------------------------------------------------------------------------------
// the pixel kernel: one thread per pixel walks its whole path — for tiny scenes and tiny frames
// -------------------------------------------------------------------------------------------------
// The wavefront exists to keep 32 lanes busy on deep BVHs.  A scene of a handful of analytic primitives (sphere_plane.json:
// three spheres and a plane) or a frame of a few hundred pixels (triangle.json is 20 x 20) has nothing to regroup: queue
// records, a launch per level and the phase barriers are pure overhead there (the reference's own kernel, one launch, beat
// the round-1 wavefront on triangle.json: 11.5 us against 33 us).  This kernel keeps the reference's shape — a thread per
// pixel — without its cost: the recursion of ray_color (inc/shading.hpp:116-154) is a loop with an explicit stack of at
// most one deferred child per level, the primary ray is cast once (the reference casts it twice, inc/kernel.hpp:52 and
// inc/shading.hpp:123), path weights replace the nested blend, and the colour is summed in depth-first order — for
// non-branching scenes exactly the level order of the wavefront's ordered sum.  Counters arrive in mapped host memory and
// are cleared by the last block, like in the frame kernel: the frame is ONE launch, no memset, no copy.
#ifdef CTB_PIXEL_STAMPS   // tuning builds: %globaltimer at eight points of the kernel, taken by one thread in the middle of a 20 x 20 frame
#define CTB_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 200) s_stamp[i] = globaltimer_ns(); } while (0)
#else
#define CTB_STAMP(i) do { } while (0)
#endif

// One pixel's walk through the recursion of ray_color: the ray in flight, its path weight and level, the colour summed so far and the
// deferred children (at most one per level: the transmitted ray of a material that also reflects).
struct PendingRay { vec3 o, d; float w; uint32_t level; };
struct PixelPath {
  vec3 o, d;
  float w;
  uint32_t level;
  float r, g, b;
  size_t gi, g2;   // the pixel's index in a.out (its layout) and in a.out2 (row-major)
  int sp;
  PendingRay stack[16];
};

__device__ __forceinline__ void path_begin(const PixelArgs &a, uint32_t gx, uint32_t gy, uint32_t pix, PixelPath &p) {
  camera_ray(a.sv.cam, gx, gy, p.o, p.d);
  p.w = 1.0f;
  p.level = 0;
  p.r = p.g = p.b = 0.f;
  p.sp = 0;
  p.gi = a.out.row_major ? (size_t)gy * a.tm.width + gx : (size_t)pix;
  p.g2 = (size_t)gy * a.tm.width + gx;
}

// casts the path's current ray, shades its hit and moves on to the next ray of the pixel; false when the pixel is finished
template <int MODE, bool BRUTE, bool OPAQUE>
__device__ __forceinline__ bool path_step(const PixelArgs &a, const float4 *nodes, const float4 *prims, PixelPath &p, TraceAcc &acc, unsigned &casts) {
  const SceneView &sv = a.sv;
  Hit h;
  closest_hit<MODE, BRUTE>(sv, nodes, prims, p.o, p.d, sv.fudge, h);
  const bool hit = h.kind >= 0;
  vec3 point = mk3(0, 0, 0), nrm = mk3(0, 0, 0);
  if (hit) hit_surface<MODE>(sv, prims, h, p.o, p.d, point, nrm);
  if (p.level == 0) {   // G-buffer, inc/kernel.hpp:52-56
    const size_t gi = p.gi, g2 = p.g2;
    a.out.depth[gi] = h.t;
    a.out.normal[3 * gi] = nrm.x; a.out.normal[3 * gi + 1] = nrm.y; a.out.normal[3 * gi + 2] = nrm.z;
    a.out.hit_id[gi] = hit ? h.obj : CUTRACE_NO_HIT;
    if (a.out2.depth) a.out2.depth[g2] = h.t;
    if (a.out2.normal) { a.out2.normal[3 * g2] = nrm.x; a.out2.normal[3 * g2 + 1] = nrm.y; a.out2.normal[3 * g2 + 2] = nrm.z; }
    if (a.out2.hit_id) a.out2.hit_id[g2] = hit ? h.obj : CUTRACE_NO_HIT;
    if (hit && isfinite(h.t)) acc.max_depth = fmaxf(acc.max_depth, h.t);
  }
  bool next = false;
  if (hit) {
    const uint32_t mat = __ldg(sv.obj_material + h.obj);
    const float4 m1 = __ldg(reinterpret_cast<const float4 *>(sv.materials + mat) + 1);
    const float reflect = m1.x, transp = m1.z;
    // inc/shading.hpp:126-149
    const bool deeper = p.level < a.bounces;
    const bool do_refl = deeper && (double)reflect >= 1e-6, do_trans = deeper && (double)transp >= 1e-6;
    const float w_own = do_trans ? p.w * (1.0f - transp) : p.w;
    const vec3 final = phong_record<MODE, BRUTE, OPAQUE>(sv, nodes, prims, point, nrm, p.d, mat, casts);
    acc.n_shaded++;
    // product and sum rounded separately, like the wavefront's level image + ordered sum
    p.r = __fadd_rn(p.r, __fmul_rn(w_own, final.x)); p.g = __fadd_rn(p.g, __fmul_rn(w_own, final.y)); p.b = __fadd_rn(p.b, __fmul_rn(w_own, final.z));
    const vec3 origin = vmad(p.o, p.d, h.t);   // incoming->start + distance * incoming->dir
    if (do_trans) {
      acc.n_trans++;
      if (do_refl) { PendingRay &q = p.stack[p.sp++]; q.o = origin; q.d = p.d; q.w = p.w * transp; q.level = p.level + 1; }
      else { p.o = origin; p.w = p.w * transp; p.level++; next = true; }
    }
    if (do_refl) {
      acc.n_refl++;
      const vec3 nd = vnormalized(p.d), nn = vnormalized(nrm);
      p.d = vreflect(nd, nn);
      p.o = origin; p.w = w_own * reflect; p.level++; next = true;
    }
  }
  if (!next) {
    if (p.sp == 0) return false;
    const PendingRay &q = p.stack[--p.sp];
    p.o = q.o; p.d = q.d; p.w = q.w; p.level = q.level;
  }
  return true;
}

__device__ __forceinline__ void path_store_color(const PixelArgs &a, const PixelPath &p) {
  float *cp = a.out.color + 3 * p.gi;
  cp[0] = p.r; cp[1] = p.g; cp[2] = p.b;
  if (a.out2.color) { float *c2 = a.out2.color + 3 * p.g2; c2[0] = p.r; c2[1] = p.g; c2[2] = p.b; }
}

// REFILL (-DCTB_PIXEL_REFILL=1 builds + CUTRACE_PIXEL_REFILL=1): a LANE whose pixel is finished takes the next pixel of its warp's
// claim at once, instead of idling until the longest path of its 32 pixels has ended.  In the 10 M-triangle hall a mesh pixel is one
// level deep, a floor pixel two to six: 16 x 2 pixel groups mix them, and a warp runs at 11.6 of 32 lanes
// (profiles/r02_pixel_kernel_synthetic10m.md).  Every step of the loop is: hand unassigned work of the claim to the free lanes
// (ballot + popc ranks; a new claim from the global cursor when the current one is used up), then ONE path step for every lane that
// holds a pixel.  A pixel's arithmetic does not depend on the lane or the step it runs in: the frames are bit-identical (same md5
// on all five workloads, full frames and 1/8 shards).
// NEGATIVE RESULT (profiles/r02_tuning.md 7): the hall got 32 % SLOWER (70.1 -> 92.7 ms, 1/8 shard 9.45 -> 12.03 ms), bunny.json 4 %,
// mirror.json 16 %.  Lanes that sit at different bounce levels of different pixels share no nodes: what the warp gains in active
// lanes it loses in L1 sectors per node visit and in walks that leave the node loop out of phase — the same finding as
// any_hit_lights (trace.cuh).  Compiled out by default.
#ifndef CTB_PIXEL_REFILL
#define CTB_PIXEL_REFILL 0
#endif
template <int MODE, bool BRUTE, bool OPAQUE, bool REFILL, bool SEG>
__global__ void __launch_bounds__(CTB_PIXEL_THREADS, CTB_PIXEL_MIN_BLOCKS) pixel_kernel(const __grid_constant__ PixelArgs a) {
  extern __shared__ float4 smem[];
#ifdef CTB_PIXEL_STAMPS
  __shared__ unsigned long long s_stamp[10];
#endif
  CTB_STAMP(0);
  const SceneView &sv = a.sv;
  const float4 *nodes, *prims;
  stage_scene<MODE>(sv, smem, nodes, prims);
  const unsigned lane = threadIdx.x & 31u;
  CTB_STAMP(1);
  TraceAcc acc;
  trace_acc_reset(acc);
  unsigned casts = 0;
  // persistent CTAs: a warp claims 32 .. 128 pixels at a time (paths differ in length by an order of magnitude); frames
  // with fewer than ~16 claims per warp get the finest grain, or the last claim of a warp is a quarter of its work
  const unsigned warps_total = gridDim.x * (blockDim.x >> 5);
  unsigned max_chunk = (a.n_px / (warps_total * 16u)) & ~31u;
  max_chunk = max_chunk < 32u ? 32u : (max_chunk > 128u ? 128u : max_chunk);
  if (REFILL) {
    PixelPath p;
    bool have = false;                 // this lane holds a pixel
    unsigned nxt = 0, end = 0;         // warp-uniform: the unassigned work items of the current claim
    bool more = true, first = true;    // warp-uniform: the cursor may have work left / the static first claim is still to come
    for (;;) {
      unsigned need = __ballot_sync(CTB_FULL, !have);
      while (need) {
        if (nxt >= end) {
          if (!more) break;
          unsigned b, e;
          const bool got = claim_pixels(&a.ctr->work_trace[0].v, a.n_px, lane, max_chunk, first, __popc(need), b, e);
          if (!got && !first) { more = false; break; }
          first = false;
          if (!got) continue;
          nxt = b; end = e;
        }
        const unsigned avail = end - nxt, asked = __popc(need);
        const unsigned rank = __popc(need & lanemask_lt());
        if (!have && rank < avail) {
          const uint32_t i = nxt + rank;
          uint32_t gx = 0, gy = 0, pix = 0;
          if (i < a.n_px && work_to_pixel(a.tm, a.px_base + i, gx, gy, pix)) { path_begin(a, gx, gy, pix, p); have = true; }
        }
        nxt += asked < avail ? asked : avail;
        need = __ballot_sync(CTB_FULL, !have);   // out-of-image slots of the tile padding: ask again
      }
      if (!__any_sync(CTB_FULL, have)) break;
      if (have && !path_step<MODE, BRUTE, OPAQUE>(a, nodes, prims, p, acc, casts)) {
        path_store_color(a, p);
        have = false;
      }
    }
  } else {
    for (bool first = true;; first = false) {
      unsigned base, end, owner = 0;
      if (SEG) {
        if (!claim_segment(a.ctr, a.n_px, lane, max_chunk, first, owner, base, end)) { if (first) continue; break; }
      } else {
        if (!claim_work(&a.ctr->work_trace[0].v, a.n_px, lane, max_chunk, true, first, base, end)) { if (first) continue; break; }
      }
#pragma unroll 1
      for (unsigned off = 0; base + off < end; off += 32) {
        const uint32_t i = SEG ? seg_to_work(base + off, owner, gridDim.x) + lane : base + off + lane;
        uint32_t gx = 0, gy = 0, pix = 0;
        if (!(i < a.n_px && work_to_pixel(a.tm, a.px_base + i, gx, gy, pix))) continue;
        struct Pending { vec3 o, d; float w; uint32_t level; } stack[16];
        int sp = 0;
        vec3 o, d;
        camera_ray(sv.cam, gx, gy, o, d);
        float w = 1.0f;
        uint32_t level = 0;
        float r = 0.f, g = 0.f, b = 0.f;
        const size_t gi = a.out.row_major ? (size_t)gy * a.tm.width + gx : (size_t)pix;
        const size_t g2 = (size_t)gy * a.tm.width + gx;
        for (;;) {
          Hit h;
          closest_hit<MODE, BRUTE>(sv, nodes, prims, o, d, sv.fudge, h);
          if (level == 0) CTB_STAMP(2);
          const bool hit = h.kind >= 0;
          vec3 point = mk3(0, 0, 0), nrm = mk3(0, 0, 0);
          if (hit) hit_surface<MODE>(sv, prims, h, o, d, point, nrm);
          if (level == 0) {   // G-buffer, inc/kernel.hpp:52-56
            a.out.depth[gi] = h.t;
            a.out.normal[3 * gi] = nrm.x; a.out.normal[3 * gi + 1] = nrm.y; a.out.normal[3 * gi + 2] = nrm.z;
            a.out.hit_id[gi] = hit ? h.obj : CUTRACE_NO_HIT;
            if (a.out2.depth) a.out2.depth[g2] = h.t;
            if (a.out2.normal) { a.out2.normal[3 * g2] = nrm.x; a.out2.normal[3 * g2 + 1] = nrm.y; a.out2.normal[3 * g2 + 2] = nrm.z; }
            if (a.out2.hit_id) a.out2.hit_id[g2] = hit ? h.obj : CUTRACE_NO_HIT;
            if (hit && isfinite(h.t)) acc.max_depth = fmaxf(acc.max_depth, h.t);
          }
          bool next = false;
          if (hit) {
            const uint32_t mat = __ldg(sv.obj_material + h.obj);
            const float4 m1 = __ldg(reinterpret_cast<const float4 *>(sv.materials + mat) + 1);
            const float reflect = m1.x, transp = m1.z;
            // inc/shading.hpp:126-149
            const bool deeper = level < a.bounces;
            const bool do_refl = deeper && (double)reflect >= 1e-6, do_trans = deeper && (double)transp >= 1e-6;
            const float w_own = do_trans ? w * (1.0f - transp) : w;
            const vec3 final = phong_record<MODE, BRUTE, OPAQUE>(sv, nodes, prims, point, nrm, d, mat, casts);
            if (level == 0) CTB_STAMP(3);
            acc.n_shaded++;
            // product and sum rounded separately, like the wavefront's level image + ordered sum
            r = __fadd_rn(r, __fmul_rn(w_own, final.x)); g = __fadd_rn(g, __fmul_rn(w_own, final.y)); b = __fadd_rn(b, __fmul_rn(w_own, final.z));
            const vec3 origin = vmad(o, d, h.t);   // incoming->start + distance * incoming->dir
            if (do_trans) {
              acc.n_trans++;
              if (do_refl) { stack[sp].o = origin; stack[sp].d = d; stack[sp].w = w * transp; stack[sp].level = level + 1; sp++; }
              else { o = origin; w = w * transp; level++; next = true; }
            }
            if (do_refl) {
              acc.n_refl++;
              const vec3 nd = vnormalized(d), nn = vnormalized(nrm);
              d = vreflect(nd, nn);
              o = origin; w = w_own * reflect; level++; next = true;
            }
          }
          if (!next) {
            if (sp == 0) break;
            sp--;
            o = stack[sp].o; d = stack[sp].d; w = stack[sp].w; level = stack[sp].level;
          }
        }
        float *cp = a.out.color + 3 * gi;
        cp[0] = r; cp[1] = g; cp[2] = b;
        if (a.out2.color) { float *c2 = a.out2.color + 3 * g2; c2[0] = r; c2[1] = g; c2[2] = b; }
        CTB_STAMP(4);
      }
    }
  }
  CTB_STAMP(5);
  // ---- tallies: warp -> block -> frame statistics; the last block publishes them and clears the counters ----
  __shared__ unsigned long long s_t[4];
  __shared__ unsigned s_md, s_last;
  if (threadIdx.x == 0) { s_t[0] = s_t[1] = s_t[2] = s_t[3] = 0ull; s_md = 0u; }
  __syncthreads();
  {
    unsigned long long c = casts;
    for (int sft = 16; sft > 0; sft >>= 1) {
      acc.n_refl += __shfl_xor_sync(CTB_FULL, acc.n_refl, sft);
      acc.n_trans += __shfl_xor_sync(CTB_FULL, acc.n_trans, sft);
      acc.n_shaded += __shfl_xor_sync(CTB_FULL, acc.n_shaded, sft);
      acc.max_depth = fmaxf(acc.max_depth, __shfl_xor_sync(CTB_FULL, acc.max_depth, sft));
      c += __shfl_xor_sync(CTB_FULL, c, sft);
    }
    if (lane == 0) {
      if (acc.n_refl) atomicAdd(&s_t[0], (unsigned long long)acc.n_refl);
      if (acc.n_trans) atomicAdd(&s_t[1], (unsigned long long)acc.n_trans);
      if (acc.n_shaded) atomicAdd(&s_t[2], (unsigned long long)acc.n_shaded);
      if (c) atomicAdd(&s_t[3], c);
      if (acc.max_depth > 0.f) atomicMax(&s_md, __float_as_uint(acc.max_depth));
    }
  }
  __syncthreads();
  FrameCounters *ctr = a.ctr;
  if (gridDim.x == 1) {
    // a frame of at most 1024 tile-padded pixels (triangle.json: 20 x 20 = four 16 x 16 tiles): the block's tallies ARE the frame's.  The general path below
    // is a chain of dependent round trips — counters -> fence -> `finished` -> re-read -> host — about 1 us of an 11 us kernel.
    __shared__ FrameStats s_fs;
    constexpr unsigned SN1 = sizeof(FrameStats) / 4;
    if (threadIdx.x < SN1) reinterpret_cast<unsigned *>(&s_fs)[threadIdx.x] = 0u;
    __syncthreads();
    if (threadIdx.x == 0) {
      s_fs.rays_reflect = s_t[0]; s_fs.rays_transmit = s_t[1]; s_fs.shade_records = s_t[2]; s_fs.shadow_casts = s_t[3];
      s_fs.max_depth_bits = s_md;
      ctr->work_trace[0].v = 0u; ctr->seg[0].v = 0u;   // (only a debug grid override makes one block claim dynamically)
    }
    __syncthreads();
    volatile unsigned *dst1 = reinterpret_cast<volatile unsigned *>(a.host_stats);
    if (dst1 && threadIdx.x < SN1) dst1[threadIdx.x] = reinterpret_cast<const unsigned *>(&s_fs)[threadIdx.x];
#ifdef CTB_PIXEL_STAMPS
    CTB_STAMP(6);
    __syncthreads();
    if (threadIdx.x < 7 && a.host_stats) reinterpret_cast<volatile unsigned long long *>(a.host_stats->phase_ns)[threadIdx.x] = s_stamp[threadIdx.x];
#endif
    return;
  }
  if (threadIdx.x == 0) {
    if (s_t[0]) atomicAdd(&ctr->st.rays_reflect, s_t[0]);
    if (s_t[1]) atomicAdd(&ctr->st.rays_transmit, s_t[1]);
    if (s_t[2]) atomicAdd(&ctr->st.shade_records, s_t[2]);
    if (s_t[3]) atomicAdd(&ctr->st.shadow_casts, s_t[3]);
    if (s_md) atomicMax(&ctr->st.max_depth_bits, s_md);
    __threadfence();
    s_last = atomicAdd(&ctr->finished.v, 1u) == gridDim.x - 1u;
  }
  __syncthreads();
#ifdef CTB_PIXEL_STAMPS
  CTB_STAMP(6);
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 7 && a.host_stats) reinterpret_cast<volatile unsigned long long *>(a.host_stats->phase_ns)[threadIdx.x] = s_stamp[threadIdx.x];
#endif
  if (s_last) {
    __threadfence();
    unsigned *src = reinterpret_cast<unsigned *>(ctr);
#ifdef CTB_PIXEL_STAMPS
    constexpr unsigned S0 = offsetof(FrameCounters, st) / 4, SN = 10;   // (leaves phase_ns to the stamps)
#else
    constexpr unsigned S0 = offsetof(FrameCounters, st) / 4, SN = sizeof(FrameStats) / 4;
#endif
    volatile unsigned *dst = reinterpret_cast<volatile unsigned *>(a.host_stats);
    if (threadIdx.x < SN) {
      const unsigned v = __ldcg(src + S0 + threadIdx.x);
      if (dst) dst[threadIdx.x] = v;
      src[S0 + threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) { ctr->finished.v = 0u; ctr->work_trace[0].v = 0u; }
    if (SEG) for (unsigned k = threadIdx.x; k < gridDim.x; k += blockDim.x) ctr->seg[k].v = 0u;
    // (no system fence: the stores to the mapped host block are complete when the kernel is, and nobody reads them earlier)
  }
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
#ifdef CTB_PHASE_DEBUG
extern "C" int cutrace_debug_phase_dump(unsigned long long *out /* 6 x 18 */) {
  unsigned long long init[6][18];
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, g_dbg_ns, sizeof init) != cudaSuccess) return -1;
  for (int k = 0; k < 6; k++) for (int p = 0; p < 18; p++) init[k][p] = (k == 0 || k == 3) ? ~0ull : 0ull;
  return cudaMemcpyToSymbol(g_dbg_ns, init, sizeof init) == cudaSuccess ? 0 : -1;
}
#endif

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

typedef void (*trace_fn)(const SceneView, const TileMap, uint32_t, uint32_t, uint32_t, uint32_t, const LevelIO, FrameCounters *, FrameTargets,
                         uint32_t *);
typedef void (*shade_fn)(const SceneView, uint32_t, const ShadeRec *, uint32_t, FrameCounters *, FrameTargets, int, float *, uint32_t);
typedef void (*frame_fn)(const FrameArgs);

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
static frame_fn pick_frame(int mode, bool brute, bool opaque) {
  if (brute) return opaque ? frame_kernel<0, true, true> : frame_kernel<0, true, false>;
  if (mode == 1) return opaque ? frame_kernel<1, false, true> : frame_kernel<1, false, false>;
  if (mode == 2) return opaque ? frame_kernel<2, false, true> : frame_kernel<2, false, false>;
  return opaque ? frame_kernel<0, false, true> : frame_kernel<0, false, false>;
}

typedef void (*pixel_fn)(const PixelArgs);
template <bool REFILL, bool SEG>
static pixel_fn pick_pixel_r(int mode, bool brute, bool opaque) {
  if (brute) return opaque ? pixel_kernel<0, true, true, REFILL, SEG> : pixel_kernel<0, true, false, REFILL, SEG>;
  if (mode == 1) return opaque ? pixel_kernel<1, false, true, REFILL, SEG> : pixel_kernel<1, false, false, REFILL, SEG>;
  return opaque ? pixel_kernel<0, false, true, REFILL, SEG> : pixel_kernel<0, false, false, REFILL, SEG>;
}
static pixel_fn pick_pixel(int mode, bool brute, bool opaque, bool refill, bool seg) {
#if CTB_PIXEL_REFILL
  if (refill) return pick_pixel_r<true, false>(mode, brute, opaque);
#endif
  (void)refill;
  return seg ? pick_pixel_r<false, true>(mode, brute, opaque) : pick_pixel_r<false, false>(mode, brute, opaque);
}

cudaError_t plan_launch(const SceneView &sv, bool allow_smem, LaunchCfg *cfg) {
  int dev = 0, sms = 0, smem_optin = 0, coop = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return e;
  size_t need = (size_t)sv.n_nodes * CTB_NODE_BYTES + (size_t)sv.n_prims * sizeof(PrimRec);
  cfg->mode = 0;
  cfg->smem_bytes = 0;
  if (allow_smem && !sv.brute_force && sv.n_prims > 0 && need + 1024 <= (size_t)smem_optin) {
    cfg->mode = 1;
    cfg->smem_bytes = need + 16;   // + the staging mbarrier
  } else if (allow_smem && !sv.brute_force && sv.smem_nodes > 0) {
    cfg->mode = 2;   // api.cu moved the top sv.smem_nodes nodes (breadth-first) to the front of the node array
    cfg->smem_bytes = (size_t)sv.smem_nodes * sizeof(Node) + 16;
  }
  trace_fn tf = pick_trace(cfg->mode, sv.brute_force != 0);
  shade_fn sf = pick_shade(cfg->mode, sv.brute_force != 0, sv.all_opaque != 0);
  frame_fn ff = pick_frame(cfg->mode, sv.brute_force != 0, sv.all_opaque != 0);
  int occ_t = 1, occ_s = 1, occ_f = 1;
  if (cfg->mode != 0) {
    if ((e = cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(sf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(ff, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes)) != cudaSuccess) return e;
  }
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_t, tf, TRACE_THREADS, cfg->smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, sf, TRACE_THREADS, cfg->smem_bytes)) != cudaSuccess) return e;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, ff, TRACE_THREADS, cfg->smem_bytes)) != cudaSuccess) return e;
  if (occ_t < 1) occ_t = 1;
  if (occ_s < 1) occ_s = 1;
  cfg->grid_trace = sms * occ_t;
  cfg->grid_shade = sms * occ_s;
  cfg->grid_frame = occ_f >= 1 && coop ? sms * occ_f : 0;   // 0: no cooperative launch on this device -> multi-launch path
  {
    const int pmode = cfg->mode == 1 ? 1 : 0;
    cfg->pixel_refill = 0;   // lane refill (pixel_kernel<.., REFILL>): a tuning build's experiment, see render.cu
    if (CTB_PIXEL_REFILL) { if (const char *e = getenv("CUTRACE_PIXEL_REFILL")) cfg->pixel_refill = atoi(e) != 0; }
    // one work cursor per CTA (claim_segment) for scenes that are walked through L1 / L2, on frames of at least CTB_SEG_MIN_PX
    // pixels (launch_pixel): hall 8K 70.2 -> 68.7 ms, its 1/8 shard 9.36 -> 8.89; the staged scenes gain nothing (their BVH is in
    // shared memory) and small frames lose to the chunk granularity (mirror.json 1080p 0.26 -> 0.31 ms), profiles/r02_tuning.md 7.
    // CUTRACE_PIXEL_SEG=0|1 (developer override): never / whenever the grid allows.
    cfg->pixel_seg = pmode == 0 && !sv.brute_force ? 1 : 0;
    if (const char *e = getenv("CUTRACE_PIXEL_SEG")) cfg->pixel_seg = atoi(e) != 0 ? 2 : 0;
    pixel_fn pf = pick_pixel(pmode, sv.brute_force != 0, sv.all_opaque != 0, cfg->pixel_refill != 0, cfg->pixel_seg != 0);
    const size_t psmem = pmode == 1 ? cfg->smem_bytes : 0;
    int occ_p = 1;
    for (int sg = 0; sg < 2 && pmode == 1; sg++)   // launch_pixel may pick either cursor variant
      if ((e = cudaFuncSetAttribute(pick_pixel(pmode, sv.brute_force != 0, sv.all_opaque != 0, cfg->pixel_refill != 0, sg != 0),
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, pf, CTB_PIXEL_THREADS, psmem)) != cudaSuccess) return e;
    cfg->grid_pixel = sms * (occ_p < 1 ? 1 : occ_p);
  }
  return cudaSuccess;
}

static inline int clamp_grid(int persistent, uint32_t work_bound) {
  uint64_t need = ((uint64_t)work_bound + 32 * (TRACE_THREADS / 32) - 1) / (32 * (TRACE_THREADS / 32));
  if (need < 1) need = 1;
  return (int)(need < (uint64_t)persistent ? need : (uint64_t)persistent);
}

void launch_trace(const LaunchCfg &cfg, const SceneView &sv, const TileMap &tm, uint32_t level, uint32_t bounces,
                  uint32_t px_base, uint32_t n_px, const RayRec *rays_in, RayRec *rays_out, uint32_t ray_cap, ShadeRec *shade_out,
                  uint32_t shade_cap, FrameCounters *ctr, const FrameTargets &fb, uint32_t *nlev, uint32_t work_bound, cudaStream_t st) {
  int grid = clamp_grid(cfg.grid_trace, work_bound);
  LevelIO io;
  io.rays_in = rays_in; io.rays_out = rays_out; io.shade_out = shade_out; io.ray_cap = ray_cap; io.shade_cap = shade_cap;
  pick_trace(cfg.mode, sv.brute_force != 0)<<<grid, TRACE_THREADS, cfg.smem_bytes, st>>>(sv, tm, level, bounces, px_base, n_px, io, ctr, fb, nlev);
}

void launch_shade(const LaunchCfg &cfg, const SceneView &sv, uint32_t level, const ShadeRec *shade, uint32_t shade_cap, FrameCounters *ctr,
                  const FrameTargets &fb, bool atomic_accumulate, float *level_color, uint32_t px_base, uint32_t work_bound,
                  cudaStream_t st) {
  int grid = clamp_grid(cfg.grid_shade, work_bound);
  pick_shade(cfg.mode, sv.brute_force != 0, sv.all_opaque != 0)<<<grid, TRACE_THREADS, cfg.smem_bytes, st>>>(
      sv, level, shade, shade_cap, ctr, fb, atomic_accumulate ? 1 : 0, level_color, px_base);
}

cudaError_t launch_pixel(const LaunchCfg &cfg, const PixelArgs &args, cudaStream_t st) {
  if (!args.n_px) return cudaSuccess;
  // frames of a few thousand pixels read the scene through L1: staging it per CTA (mbarrier round trip) costs more than it saves
  const int mode = cfg.mode == 1 && args.n_px >= (1u << 16) ? 1 : 0;
  const size_t smem = mode == 1 ? cfg.smem_bytes : 0;
  // tuning experiments (read once per process: a 20 x 20 frame is a few microseconds, two walks over environ[] are not free)
  static const int dbg_threads = [] { const char *e = getenv("CUTRACE_DEBUG_PIXEL_THREADS"); return e ? atoi(e) : 0; }();
  static const int dbg_grid = [] { const char *e = getenv("CUTRACE_DEBUG_PIXEL_GRID"); return e ? atoi(e) : 0; }();
  int threads = CTB_PIXEL_THREADS;
  if (dbg_threads >= 32 && dbg_threads <= CTB_PIXEL_THREADS && dbg_threads % 32 == 0) threads = dbg_threads;
  uint64_t need = ((uint64_t)args.n_px + threads - 1) / threads;
  int grid = (int)(need < (uint64_t)cfg.grid_pixel ? need : (uint64_t)cfg.grid_pixel);
  if (dbg_grid > 0 && dbg_grid < grid) grid = dbg_grid;
  if (grid < 1) grid = 1;
  const bool seg = (cfg.pixel_seg == 2 || (cfg.pixel_seg == 1 && args.n_px >= CTB_SEG_MIN_PX)) && grid <= (int)CTB_MAX_SEGS;
  pick_pixel(mode, args.sv.brute_force != 0, args.sv.all_opaque != 0, cfg.pixel_refill != 0, seg)<<<grid, threads, smem, st>>>(args);
  return cudaGetLastError();
}

cudaError_t launch_frame(const LaunchCfg &cfg, const FrameArgs &args, uint32_t work_bound, cudaStream_t st) {
  if (cfg.grid_frame <= 0) return cudaErrorNotSupported;
  int grid = clamp_grid(cfg.grid_frame, work_bound);
  frame_fn f = pick_frame(cfg.mode, args.sv.brute_force != 0, args.sv.all_opaque != 0);
  void *params[] = {const_cast<FrameArgs *>(&args)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(f), dim3(grid), dim3(TRACE_THREADS), params, cfg.smem_bytes, st);
}

}  // namespace ctb
