// scene_loader.hpp — host front-end: scene JSON + STL meshes -> flat cutrace_scene_desc arrays.
// Accepts exactly what the reference's default schema accepts (inc/default_schema.hpp:487-501,596-606,
// 633-645,672-684,719-729,754-764,805-821,876-897; top level inc/loader.hpp:679-745):
//   objects  : triangle{p1,p2,p3,material} | mesh{file,material} | plane{point,normal,material} |
//              sphere{center,radius,material}
//   lights   : sun{direction, color=[1,1,1]} | point{point, color=[1,1,1]}
//   materials: solid{color, specular=0.3, reflect=0, phong=32, transparency=0}
//   camera   : {eye, up, look, near_plane, far_plane, width, height, ambient} — all mandatory (MK_MANDATORY)
// Numbers are doubles cast to float / size_t (inc/json_helpers.hpp:91).  Meshes: the reference imports with
// Assimp and keeps faces in file order (inc/default_schema.hpp:516-545); here binary STL, ASCII STL and Wavefront OBJ
// (fan-triangulated) are read directly in face order.  With accept_aliases the stale spellings of schema.md are accepted as well.
#ifndef CUTRACE_B200_HOST_SCENE_LOADER_HPP
#define CUTRACE_B200_HOST_SCENE_LOADER_HPP
#include <string>
#include <vector>
#include "../../include/cutrace.h"

namespace cthost {

struct FlatScene {
  float cam_pos[3] = {0, 0, 0}, cam_up[3] = {0, 1, 0}, cam_forward[3] = {0, 0, 1}, cam_right[3] = {1, 0, 0};
  float ambient = 0.1f, near_plane = 0.1f, far_plane = 100.0f;
  uint32_t width = 1920, height = 1080;
  std::vector<float> tri_p1, tri_p2, tri_p3;
  std::vector<uint32_t> tri_object;
  std::vector<float> sph_center, sph_radius;
  std::vector<uint32_t> sph_object;
  std::vector<float> pl_point, pl_normal;
  std::vector<uint32_t> pl_object;
  std::vector<uint32_t> obj_material, obj_kind;
  std::vector<float> mat_color, mat_specular, mat_reflect, mat_phong, mat_transparency;
  std::vector<uint32_t> light_kind;
  std::vector<float> light_vec, light_color;

  cutrace_scene_desc desc() const;   // borrows this object's vectors
};

struct LoadOptions {
  std::string base_dir;        // mesh paths are resolved against this ("" = current directory, like the reference)
  bool accept_aliases = false; // also accept schema.md's stale spellings (model / position / points / untyped material)
};

// Returns true on success. On failure `errors` holds one message per problem, worded like the reference's
// stderr output ("Error while loading object #i: ...").
bool load_scene_file(const std::string &path, const LoadOptions &opt, FlatScene &out, std::vector<std::string> &errors);
bool load_scene_text(const std::string &json_text, const LoadOptions &opt, FlatScene &out, std::vector<std::string> &errors);
bool read_stl(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err);
bool read_obj(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err);
bool read_ply(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err);
bool read_off(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err);

// cam::look_at in float arithmetic (inc/default_schema.hpp:370-374)
void look_at(const float pos[3], const float up_in[3], const float look[3], float forward[3], float right[3], float up[3]);

extern const char *const kSchemaHelp;   // printed where the reference dumps its schema (main.cu:16-19)

}  // namespace cthost
#endif
