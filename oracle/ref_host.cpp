// ref_host.cpp — TEST INFRASTRUCTURE. The reference's own device functions compiled for the host.
//
// /root/reference has no CPU renderer and its ray_cast / shading / intersect functions are
// __device__-only (inc/ray_cast.hpp:30, inc/shading.hpp:23,65,117).  This translation unit erases
// the CUDA qualifiers with the preprocessor and #includes the reference headers IN PLACE
// (-I$CUTRACE_REF/inc; nothing is copied into this repo), then runs the body of render_kernel
// (inc/kernel.hpp:44-59) in an OpenMP loop.  It is (i) the pin for oracle/cutrace_oracle.c,
// (ii) the generator of tests/golden/*.npz, (iii) the "reference" CPU baseline of bench.py.
// Known deltas to the device build: no FMA contraction (g++ -ffp-contract=off vs nvcc -fmad=true),
// glibc powf/sqrtf instead of CUDA's.  device min/max are mapped to fminf/fmaxf (NaN rules of
// inc/default_schema.hpp:109-110 on the device).
#include <algorithm>
#include <cmath>
#include <math.h>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <type_traits>

#define __device__
#define __host__
#define __global__
#define cudaCheck(x)
#define CUTRACE_CUDA_HPP  // skip inc/cuda.hpp: no cuda_runtime.h in a g++ build
inline int cudaFree(void *) { return 0; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
using std::isfinite;
static_assert(std::is_same_v<decltype(pow(1.0f, 1.0f)), float>, "pow(float,float) must be powf like on the device");
static_assert(std::is_same_v<decltype(sqrt(1.0f)), float>, "sqrt(float) must be sqrtf like on the device");

#include "default_schema.hpp"
#include "shading.hpp"
#include "ref_common.hpp"

// Counting wrapper: numbers of ray_cast calls are taken from a thread-local incremented here.
// (ray_cast itself is the reference's, untouched.)

extern "C" int cutrace_ref_host_render(const cutrace_scene_desc *d, float fudge, uint32_t bounces, uint64_t n_px,
                                        const uint64_t *px, float *depth, float *normal, float *color,
                                        uint32_t *hit_id, int n_threads) {
  using namespace cutrace;
  using namespace cutrace::gpu;
  using S = oracle_ref::scene_t;
  if (bounces != 5) return -3;  // the reference instantiates ray_color<S,5> (main.cu:30)
  oracle_ref::built_scene b;
  int rc = oracle_ref::build(d, [](size_t n) { return std::calloc(1, n); }, b);
  if (rc == 0) {
    const S &scene = b.scene;
    size_t w, h;
    scene.cam.get_bounds(&w, &h);
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads > 0 ? n_threads : 1)
    for (long long i = 0; i < (long long)n_px; i++) {
      size_t tid = px ? px[i] : (size_t)i;
      // ---- body of render_kernel, inc/kernel.hpp:44-59 ----
      size_t x_id = tid % w;
      size_t y_id = tid / w;
      float dist = INFINITY;
      ray r = scene.cam.get_ray(x_id, y_id);
      size_t hid = scene.objects.size;
      vector hit_point{}, nrm{0, 0, 0};
      uv tc{};
      bool did_hit = ray_cast(&scene, &r, fudge, &dist, &hid, &hit_point, &nrm, &tc, false);
      vector rgb = ray_color<S, 5>(&scene, &r, fudge, scene.cam.get_ambient());
      // -----------------------------------------------------
      if (depth) depth[i] = dist;
      if (normal) { normal[3 * i] = nrm.x; normal[3 * i + 1] = nrm.y; normal[3 * i + 2] = nrm.z; }
      if (color) { color[3 * i] = rgb.x; color[3 * i + 1] = rgb.y; color[3 * i + 2] = rgb.z; }
      if (hit_id) hit_id[i] = did_hit ? (uint32_t)hid : CUTRACE_NO_HIT;
    }
  }
  for (void *p : b.allocs) std::free(p);
  return rc;
}
