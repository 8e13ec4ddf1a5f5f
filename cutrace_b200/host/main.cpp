// main.cpp — `cutrace <scene file>`: the reference's driver (main.cu:8-47) on top of the C-ABI.
//   parse scene JSON -> flat scene -> cutrace_upload_scene -> scene dump -> cutrace_render ->
//   "Render time was X ms; kernel time with setup/teardown was Y ms." -> depth_map.jpg, normal_map.jpg, frame.jpg
// Exit codes as the reference: -1 usage, -2 scene rejected (with a schema help text where the reference dumps
// its schema, main.cu:16-19).  Extensions: --width/--height/--bounces/--fudge/--device/--out-dir/--aliases/
// --dump-raw/--host-bytes.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cutrace.h"
#include "jpeg.hpp"
#include "scene_loader.hpp"

static void dump_raw(const std::string &path, const void *p, size_t bytes) {
  if (FILE *f = fopen(path.c_str(), "wb")) { fwrite(p, 1, bytes, f); fclose(f); }
}

int main(int argc, const char **argv) {
  std::string scene_path, out_dir = ".";
  long width = 0, height = 0, bounces = 5, device = -1;
  double fudge = 1e-3;
  bool aliases = false, raw = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&](const char *name) -> const char * {
      if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", name); exit(-1); }
      return argv[++i];
    };
    if (a == "--width") width = atol(next("--width"));
    else if (a == "--height") height = atol(next("--height"));
    else if (a == "--bounces") bounces = atol(next("--bounces"));
    else if (a == "--fudge") fudge = atof(next("--fudge"));
    else if (a == "--device") device = atol(next("--device"));
    else if (a == "--out-dir") out_dir = next("--out-dir");
    else if (a == "--aliases") aliases = true;
    else if (a == "--dump-raw") raw = true;
    else if (scene_path.empty()) scene_path = a;
  }
  if (scene_path.empty()) {
    fprintf(stderr, "Usage: %s <scene file>\n", argv[0]);   // main.cu:10
    return -1;
  }

  cthost::FlatScene scene;
  cthost::LoadOptions lo;
  lo.accept_aliases = aliases;
  std::vector<std::string> errors;
  if (!cthost::load_scene_file(scene_path, lo, scene, errors)) {
    for (const auto &e : errors) fprintf(stderr, "%s\n", e.c_str());
    fputs(cthost::kSchemaHelp, stdout);
    return -2;
  }
  if (width > 0) scene.width = (uint32_t)width;
  if (height > 0) scene.height = (uint32_t)height;

  auto t_total0 = std::chrono::high_resolution_clock::now();
  cutrace_scene_desc d = scene.desc();
  cutrace_opts o;
  cutrace_default_opts(&o);
  o.fudge = (float)fudge; o.bounces = (uint32_t)bounces; o.device = (int32_t)device;
  cutrace_ctx *ctx = nullptr;
  if (cutrace_upload_scene(&d, &o, &ctx)) { fprintf(stderr, "[cutrace] %s\n", cutrace_last_error()); return -3; }

  // dump_scene_kernel, inc/kernel.hpp:152-165 (variant indices: objects 0 triangle,1 mesh,2 plane,3 sphere;
  // lights 0 sun,1 point; materials 0 solid)
  printf(" -> Have %-4llu objects:\n", (unsigned long long)scene.obj_kind.size());
  for (size_t i = 0; i < scene.obj_kind.size(); i++) printf("  -> Object   #%-4llu has type #%-2llu\n", (unsigned long long)i, (unsigned long long)scene.obj_kind[i]);
  printf(" -> Have %-4llu lights:\n", (unsigned long long)scene.light_kind.size());
  for (size_t i = 0; i < scene.light_kind.size(); i++) printf("  -> Light    #%-4llu has type #%-2llu\n", (unsigned long long)i, (unsigned long long)scene.light_kind[i]);
  printf(" -> Have %-4llu materials:\n", (unsigned long long)scene.mat_specular.size());
  for (size_t i = 0; i < scene.mat_specular.size(); i++) printf("  -> Material #%-4llu has type #%-2llu\n", (unsigned long long)i, 0ull);

  cutrace_stats st;
  if (cutrace_render(ctx, &st)) { fprintf(stderr, "[cutrace] %s\n", cutrace_last_error()); cutrace_free(ctx); return -3; }
  const size_t n = (size_t)scene.width * scene.height;
  std::vector<uint8_t> d8(3 * n), n8(3 * n), c8(3 * n);
  float max_d = 0.f;
  if (cutrace_download_bytes(ctx, d8.data(), n8.data(), c8.data(), &max_d)) { fprintf(stderr, "[cutrace] %s\n", cutrace_last_error()); cutrace_free(ctx); return -3; }
  auto t_total1 = std::chrono::high_resolution_clock::now();
  const double total_ms = std::chrono::duration<double, std::milli>(t_total1 - t_total0).count();
  // main.cu:32 (integer milliseconds in the reference; fractions are kept here because frames take < 1 ms .. tens of ms)
  printf("Render time was %.3f ms; kernel time with setup/teardown was %.3f ms.\n", (double)st.render_ms, total_ms);
  printf("[cutrace-b200] %llu rays (%llu primary, %llu reflect, %llu transmit, %llu shadow), LBVH %u nodes depth %u built in %.3f ms, %.1f Mrays/s\n",
         (unsigned long long)(st.rays_primary + st.rays_reflect + st.rays_transmit + st.rays_shadow), (unsigned long long)st.rays_primary,
         (unsigned long long)st.rays_reflect, (unsigned long long)st.rays_transmit, (unsigned long long)st.rays_shadow, st.bvh_nodes, st.bvh_depth,
         (double)st.build_ms, (double)(st.rays_primary + st.rays_reflect + st.rays_transmit + st.rays_shadow) / (st.render_ms * 1e3));

  if (raw) {
    std::vector<float> depth(n), normal(3 * n), color(3 * n);
    std::vector<uint32_t> ids(n);
    cutrace_download(ctx, depth.data(), normal.data(), color.data(), ids.data(), nullptr);
    dump_raw(out_dir + "/depth.f32", depth.data(), 4 * n); dump_raw(out_dir + "/normal.f32", normal.data(), 12 * n);
    dump_raw(out_dir + "/color.f32", color.data(), 12 * n); dump_raw(out_dir + "/hit_id.u32", ids.data(), 4 * n);
  }
  cutrace_free(ctx);
  bool ok = cthost::write_jpeg(out_dir + "/depth_map.jpg", (int)scene.width, (int)scene.height, d8.data(), 90);   // main.cu:34
  ok = cthost::write_jpeg(out_dir + "/normal_map.jpg", (int)scene.width, (int)scene.height, n8.data(), 90) && ok;  // main.cu:35
  ok = cthost::write_jpeg(out_dir + "/frame.jpg", (int)scene.width, (int)scene.height, c8.data(), 90) && ok;       // main.cu:36
  if (!ok) { fprintf(stderr, "[cutrace] could not write the output images to %s\n", out_dir.c_str()); return -4; }
  return 0;
}
