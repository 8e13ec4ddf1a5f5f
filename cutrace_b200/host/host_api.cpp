#include <cstring>
#include <string>
#include <vector>

#include "../../include/cutrace_host.h"
#include "jpeg.hpp"
#include "scene_loader.hpp"

struct cutrace_host_scene { cthost::FlatScene scene; };

extern "C" {

int cutrace_host_load_scene(const char *path, const char *base_dir, int accept_aliases, cutrace_host_scene **out,
                            cutrace_scene_desc *desc, char *errbuf, size_t errlen) {
  if (errbuf && errlen) errbuf[0] = 0;
  if (!path || !out || !desc) return -1;
  *out = nullptr;
  cthost::LoadOptions opt;
  opt.base_dir = base_dir ? base_dir : "";
  opt.accept_aliases = accept_aliases != 0;
  auto *h = new cutrace_host_scene();
  std::vector<std::string> errors;
  if (!cthost::load_scene_file(path, opt, h->scene, errors)) {
    std::string all;
    for (const auto &e : errors) all += e + "\n";
    if (errbuf && errlen) { strncpy(errbuf, all.c_str(), errlen - 1); errbuf[errlen - 1] = 0; }
    delete h;
    return -2;
  }
  *desc = h->scene.desc();
  *out = h;
  return 0;
}

void cutrace_host_free_scene(cutrace_host_scene *s) { delete s; }

int cutrace_host_write_jpeg(const char *path, int width, int height, const uint8_t *rgb, int quality) {
  if (!path || !rgb || width <= 0 || height <= 0 || width > 65535 || height > 65535) return -1;
  return cthost::write_jpeg(path, width, height, rgb, quality) ? 0 : -1;
}

void cutrace_host_look_at(const float pos[3], const float up_in[3], const float look[3], float forward[3], float right[3], float up[3]) {
  cthost::look_at(pos, up_in, look, forward, right, up);
}

}  // extern "C"
