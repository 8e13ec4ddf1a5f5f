// ref_common.hpp — TEST INFRASTRUCTURE. Fills the REFERENCE's own GPU scene types
// (cutrace::cpu::schema::default_gpu_scene, /root/reference/inc/default_schema.hpp:926) from the
// flat cutrace_scene_desc, so that the reference's unmodified code renders exactly the inputs the
// new path gets.  Included by ref_host.cpp (g++) and ref_gpu.cu (nvcc) AFTER the reference headers.
#ifndef ORACLE_REF_COMMON_HPP
#define ORACLE_REF_COMMON_HPP
#include <cstdint>
#include <cstring>
#include <vector>
#include "../include/cutrace.h"

namespace oracle_ref {
namespace g = cutrace::gpu::schema;
using scene_t = cutrace::cpu::schema::default_gpu_scene;
using cutrace::vector;

struct built_scene {
  scene_t scene{};
  std::vector<void *> allocs;
};

inline vector v3(const float *p, uint64_t i) { return vector{p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }

// alloc(bytes) must return memory visible to whoever runs the reference code (malloc on the host,
// cudaMallocManaged on the device like inc/cpu_to_gpu.hpp:102).
template <typename Alloc>
int build(const cutrace_scene_desc *d, Alloc &&alloc, built_scene &out) {
  const uint32_t no = d->n_objects;
  auto *objects = static_cast<scene_t::object *>(alloc(sizeof(scene_t::object) * (no ? no : 1)));
  auto *lights = static_cast<scene_t::light *>(alloc(sizeof(scene_t::light) * (d->n_lights ? d->n_lights : 1)));
  auto *materials = static_cast<scene_t::material *>(alloc(sizeof(scene_t::material) * (d->n_materials ? d->n_materials : 1)));
  out.allocs = {objects, lights, materials};

  std::vector<uint64_t> count(no, 0), first(no, 0), fill(no, 0);
  std::vector<int> kind(no, -1);
  for (uint64_t k = 0; k < d->n_triangles; k++) { if (d->tri_object[k] >= no) return -2; count[d->tri_object[k]]++; }
  uint64_t acc = 0;
  for (uint32_t i = 0; i < no; i++) { first[i] = acc; acc += count[i]; }
  auto *tris = static_cast<g::triangle *>(alloc(sizeof(g::triangle) * (d->n_triangles ? d->n_triangles : 1)));
  out.allocs.push_back(tris);
  for (uint64_t k = 0; k < d->n_triangles; k++) {
    uint32_t o = d->tri_object[k];
    tris[first[o] + fill[o]++] = g::triangle{v3(d->tri_p1, k), v3(d->tri_p2, k), v3(d->tri_p3, k), d->obj_material[o]};
  }
  for (uint32_t i = 0; i < no; i++) {
    if (d->obj_material[i] >= d->n_materials) return -2;
    bool is_mesh = d->obj_kind ? d->obj_kind[i] == CUTRACE_OBJ_MESH : count[i] != 1;
    if (count[i] == 0 && !(d->obj_kind && d->obj_kind[i] == CUTRACE_OBJ_MESH)) continue;
    if (!is_mesh) {
      objects[i] = scene_t::object{tris[first[i]]};
    } else {
      // cpu mesh::bounding_box(), inc/default_schema.hpp:573-586
      cutrace::bound bb = cutrace::bound::incorrect();
      for (uint64_t k = first[i]; k < first[i] + count[i]; k++) {
        const auto &t = tris[k];
        bb.merge(cutrace::bound{
            {std::min(std::min(t.p1.x, t.p2.x), t.p3.x), std::min(std::min(t.p1.y, t.p2.y), t.p3.y), std::min(std::min(t.p1.z, t.p2.z), t.p3.z)},
            {std::max(std::max(t.p1.x, t.p2.x), t.p3.x), std::max(std::max(t.p1.y, t.p2.y), t.p3.y), std::max(std::max(t.p1.z, t.p2.z), t.p3.z)}});
      }
      objects[i] = scene_t::object{g::mesh{cutrace::gpu::gpu_array<g::triangle>{tris + first[i], count[i]}, d->obj_material[i], bb}};
    }
    kind[i] = 1;
  }
  for (uint64_t k = 0; k < d->n_spheres; k++) {
    uint32_t o = d->sph_object[k];
    if (o >= no) return -2;
    objects[o] = scene_t::object{g::sphere{v3(d->sph_center, k), d->sph_radius[k], d->obj_material[o]}};
    kind[o] = 3;
  }
  for (uint64_t k = 0; k < d->n_planes; k++) {
    uint32_t o = d->pl_object[k];
    if (o >= no) return -2;
    objects[o] = scene_t::object{g::plane{v3(d->pl_point, k), v3(d->pl_normal, k), d->obj_material[o]}};
    kind[o] = 2;
  }
  for (uint32_t i = 0; i < no; i++) if (kind[i] < 0) return -2;
  for (uint32_t l = 0; l < d->n_lights; l++) {
    if (d->light_kind[l] == CUTRACE_LIGHT_SUN) lights[l] = scene_t::light{g::sun{v3(d->light_vec, l), v3(d->light_color, l)}};
    else lights[l] = scene_t::light{g::point_light{v3(d->light_vec, l), v3(d->light_color, l)}};
  }
  for (uint32_t m = 0; m < d->n_materials; m++) {
    // field order inc/default_schema.hpp:320-324: color, specular, reflexivity, phong_exp, transparency
    materials[m] = scene_t::material{g::phong_material{v3(d->mat_color, m), d->mat_specular[m], d->mat_reflect[m], d->mat_phong[m], d->mat_transparency[m]}};
  }
  out.scene.objects = {objects, no};
  out.scene.lights = {lights, d->n_lights};
  out.scene.materials = {materials, d->n_materials};
  // camera: the desc carries the result of cam::look_at (inc/default_schema.hpp:370-374)
  g::cam cam{};
  cam.pos = v3(d->cam_pos, 0); cam.up = v3(d->cam_up, 0); cam.forward = v3(d->cam_forward, 0); cam.right = v3(d->cam_right, 0);
  cam.near = 0.1f; cam.far = 100.0f; cam.ambient = d->ambient; cam.w = d->width; cam.h = d->height;
  out.scene.cam = cam;
  return 0;
}
}  // namespace oracle_ref
#endif
