/*
 * cutrace_host.h — C entry points of the host front-end (cutrace_b200/host, libcutrace_host.so; no CUDA).
 * Not part of the render-path ABI (include/cutrace.h); these wrap the callers either side of the path:
 *   - the scene JSON + STL front-end  (reference: default_schema::load_file, inc/loader.hpp:763-780,
 *     inc/default_schema.hpp:487-940)
 *   - the JPEG writer                 (reference: stbi_write_jpg(..., 3, data, 90), inc/images.hpp:39,64,86)
 * so that tests can drive the same C++ code the `cutrace` CLI uses.
 */
#ifndef CUTRACE_B200_CUTRACE_HOST_H
#define CUTRACE_B200_CUTRACE_HOST_H
#include "cutrace.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cutrace_host_scene cutrace_host_scene;

/* Parses `path` (mesh files resolved against base_dir, "" / NULL = current directory).  On success returns 0,
 * *out owns the arrays and *desc points into them.  On failure returns -2 (the reference's exit code for a
 * rejected scene, main.cu:16-19) and errbuf holds the messages, one per line. */
int cutrace_host_load_scene(const char *path, const char *base_dir, int accept_aliases, cutrace_host_scene **out,
                            cutrace_scene_desc *desc, char *errbuf, size_t errlen);
void cutrace_host_free_scene(cutrace_host_scene *s);

/* baseline JPEG, 3 components, 4:2:0 for quality <= 90 like stb_image_write; returns 0 on success */
int cutrace_host_write_jpeg(const char *path, int width, int height, const uint8_t *rgb, int quality);

/* cam::look_at in float arithmetic (inc/default_schema.hpp:370-374) */
void cutrace_host_look_at(const float pos[3], const float up_in[3], const float look[3], float forward[3], float right[3],
                          float up[3]);

#ifdef __cplusplus
}
#endif
#endif
