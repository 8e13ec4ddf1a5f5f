// main.cpp — `cutrace <scene file>`: the reference's driver (main.cu:8-47) on top of the C-ABI.
//   parse scene JSON -> flat scene -> cutrace_upload_scene -> scene dump -> cutrace_render ->
//   "Render time was X ms; kernel time with setup/teardown was Y ms." -> depth_map.jpg, normal_map.jpg, frame.jpg
// Exit codes as the reference: -1 usage, -2 scene rejected (with a schema help text where the reference dumps
// its schema, main.cu:16-19).  Extensions: --width/--height/--bounces/--fudge/--device/--out-dir/--aliases/
// --dump-raw, and --gpus N: one ctx per GPU in this process, screen tiles interleaved over the GPUs, every GPU's
// kernels storing its tiles straight into GPU 0's frame over NVLink (cutrace_frame_attach), one host thread per GPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cutrace.h"
#include "jpeg.hpp"
#include "scene_loader.hpp"

static void dump_raw(const std::string &path, const void *p, size_t bytes) {
  if (FILE *f = fopen(path.c_str(), "wb")) { fwrite(p, 1, bytes, f); fclose(f); }
}

int main(int argc, const char **argv) {
  std::string scene_path, out_dir = ".";
  long width = 0, height = 0, bounces = 5, device = -1, gpus = 1;
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
    else if (a == "--gpus") gpus = atol(next("--gpus"));
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
  if (width > 65535 || height > 65535) { fprintf(stderr, "--width / --height above 65535 cannot be written as baseline JPEG\n"); return -1; }
  if (width > 0) scene.width = (uint32_t)width;
  if (height > 0) scene.height = (uint32_t)height;

  auto t_total0 = std::chrono::high_resolution_clock::now();
  cutrace_scene_desc d = scene.desc();
  if (gpus < 1) gpus = 1;
  std::vector<cutrace_ctx *> ctxs((size_t)gpus, nullptr);
  auto fail_all = [&](const char *what) {
    fprintf(stderr, "[cutrace] %s: %s\n", what, cutrace_last_error());
    for (cutrace_ctx *c : ctxs) cutrace_free(c);
    return -3;
  };
  for (long r = 0; r < gpus; r++) {
    cutrace_opts o;
    cutrace_default_opts(&o);
    o.fudge = (float)fudge; o.bounces = (uint32_t)bounces;
    o.device = gpus > 1 ? (int32_t)r : (int32_t)device;
    o.tile_rank = (uint32_t)r; o.tile_world = (uint32_t)gpus;
    if (cutrace_upload_scene(&d, &o, &ctxs[(size_t)r])) return fail_all("upload");
  }
  cutrace_ctx *ctx = ctxs[0];
  if (gpus > 1) {   // GPU 0 owns the frame; the others store into it
    unsigned char handle[CUTRACE_IPC_HANDLE_BYTES];
    float *block = nullptr;
    if (cutrace_frame_ipc_export(ctx, handle) || cutrace_frame_device(ctx, &block, nullptr, nullptr, nullptr)) return fail_all("frame export");
    for (long r = 1; r < gpus; r++)
      if (cutrace_enable_peer_access((int)r, 0) || cutrace_frame_attach(ctxs[(size_t)r], block, scene.width, scene.height)) return fail_all("peer frame");
  }

  // dump_scene_kernel, inc/kernel.hpp:152-165 (variant indices: objects 0 triangle,1 mesh,2 plane,3 sphere;
  // lights 0 sun,1 point; materials 0 solid)
  printf(" -> Have %-4llu objects:\n", (unsigned long long)scene.obj_kind.size());
  for (size_t i = 0; i < scene.obj_kind.size(); i++) printf("  -> Object   #%-4llu has type #%-2llu\n", (unsigned long long)i, (unsigned long long)scene.obj_kind[i]);
  printf(" -> Have %-4llu lights:\n", (unsigned long long)scene.light_kind.size());
  for (size_t i = 0; i < scene.light_kind.size(); i++) printf("  -> Light    #%-4llu has type #%-2llu\n", (unsigned long long)i, (unsigned long long)scene.light_kind[i]);
  printf(" -> Have %-4llu materials:\n", (unsigned long long)scene.mat_specular.size());
  for (size_t i = 0; i < scene.mat_specular.size(); i++) printf("  -> Material #%-4llu has type #%-2llu\n", (unsigned long long)i, 0ull);

  std::vector<cutrace_stats> sts((size_t)gpus);
  std::vector<int> rcs((size_t)gpus, 0);
  std::vector<std::string> errs((size_t)gpus);
  {
    std::vector<std::thread> th;
    for (long r = 0; r < gpus; r++)
      th.emplace_back([&, r]() {
        rcs[(size_t)r] = cutrace_render(ctxs[(size_t)r], &sts[(size_t)r]);
        if (rcs[(size_t)r]) errs[(size_t)r] = cutrace_last_error();   // thread-local message
      });
    for (auto &t : th) t.join();   // the join is the barrier: every GPU's stores into the frame are complete
  }
  for (long r = 0; r < gpus; r++)
    if (rcs[(size_t)r]) { fprintf(stderr, "[cutrace] render on GPU %ld: %s\n", r, errs[(size_t)r].c_str()); for (cutrace_ctx *c : ctxs) cutrace_free(c); return -3; }
  cutrace_stats st = sts[0];
  for (long r = 1; r < gpus; r++) {
    const cutrace_stats &o = sts[(size_t)r];
    st.rays_primary += o.rays_primary; st.rays_reflect += o.rays_reflect; st.rays_transmit += o.rays_transmit;
    st.rays_shadow += o.rays_shadow; st.shadow_casts += o.shadow_casts;
    if (o.render_ms > st.render_ms) st.render_ms = o.render_ms;
    if (o.max_depth > st.max_depth) st.max_depth = o.max_depth;
  }
  cutrace_set_frame_max_depth(ctx, st.max_depth);
  const size_t n = (size_t)scene.width * scene.height;
  std::vector<uint8_t> d8(3 * n), n8(3 * n), c8(3 * n);
  float max_d = 0.f;
  if (cutrace_download_bytes(ctx, d8.data(), n8.data(), c8.data(), &max_d)) return fail_all("download");
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
  for (cutrace_ctx *c : ctxs) cutrace_free(c);
  bool ok = cthost::write_jpeg(out_dir + "/depth_map.jpg", (int)scene.width, (int)scene.height, d8.data(), 90);   // main.cu:34
  ok = cthost::write_jpeg(out_dir + "/normal_map.jpg", (int)scene.width, (int)scene.height, n8.data(), 90) && ok;  // main.cu:35
  ok = cthost::write_jpeg(out_dir + "/frame.jpg", (int)scene.width, (int)scene.height, c8.data(), 90) && ok;       // main.cu:36
  if (!ok) { fprintf(stderr, "[cutrace] could not write the output images to %s\n", out_dir.c_str()); return -4; }
  return 0;
}
