#include "scene_loader.hpp"

#include <cmath>
#include <cstdio>
#include <cctype>
#include <cstring>
#include <fstream>
#include <sstream>

#include "json.hpp"

namespace cthost {

const char *const kSchemaHelp =
    "Scene schema (as accepted by cutrace's default_schema.hpp):\n"
    "  { \"camera\":    { \"eye\":[x,y,z], \"up\":[x,y,z], \"look\":[x,y,z], \"near_plane\":n, \"far_plane\":n,\n"
    "                   \"width\":n, \"height\":n, \"ambient\":n },                  (all mandatory)\n"
    "    \"materials\": [ { \"type\":\"solid\", \"color\":[r,g,b], \"specular\":0.3, \"reflect\":0, \"phong\":32, \"transparency\":0 } ],\n"
    "    \"lights\":    [ { \"type\":\"sun\", \"direction\":[x,y,z], \"color\":[1,1,1] } | { \"type\":\"point\", \"point\":[x,y,z], \"color\":[1,1,1] } ],\n"
    "    \"objects\":   [ { \"type\":\"triangle\", \"p1\":[..], \"p2\":[..], \"p3\":[..], \"material\":i }\n"
    "                 | { \"type\":\"mesh\", \"file\":\"path.stl\", \"material\":i }\n"
    "                 | { \"type\":\"plane\", \"point\":[..], \"normal\":[..], \"material\":i }\n"
    "                 | { \"type\":\"sphere\", \"center\":[..], \"radius\":r, \"material\":i } ] }\n";

cutrace_scene_desc FlatScene::desc() const {
  cutrace_scene_desc d;
  memset(&d, 0, sizeof d);
  d.abi_version = CUTRACE_ABI_VERSION;
  for (int i = 0; i < 3; i++) { d.cam_pos[i] = cam_pos[i]; d.cam_up[i] = cam_up[i]; d.cam_forward[i] = cam_forward[i]; d.cam_right[i] = cam_right[i]; }
  d.ambient = ambient; d.width = width; d.height = height;
  auto ptr = [](const auto &v) { return v.empty() ? nullptr : v.data(); };
  d.n_triangles = tri_object.size(); d.tri_p1 = ptr(tri_p1); d.tri_p2 = ptr(tri_p2); d.tri_p3 = ptr(tri_p3); d.tri_object = ptr(tri_object);
  d.n_spheres = sph_object.size(); d.sph_center = ptr(sph_center); d.sph_radius = ptr(sph_radius); d.sph_object = ptr(sph_object);
  d.n_planes = pl_object.size(); d.pl_point = ptr(pl_point); d.pl_normal = ptr(pl_normal); d.pl_object = ptr(pl_object);
  d.n_objects = (uint32_t)obj_material.size(); d.obj_material = ptr(obj_material); d.obj_kind = ptr(obj_kind);
  d.n_materials = (uint32_t)mat_specular.size(); d.mat_color = ptr(mat_color); d.mat_specular = ptr(mat_specular);
  d.mat_reflect = ptr(mat_reflect); d.mat_phong = ptr(mat_phong); d.mat_transparency = ptr(mat_transparency);
  d.n_lights = (uint32_t)light_kind.size(); d.light_kind = ptr(light_kind); d.light_vec = ptr(light_vec); d.light_color = ptr(light_color);
  return d;
}

// ---- float3 helpers mirroring inc/vector.hpp (host build: no FMA contraction wanted) -----------------
static inline float norm3(const float v[3]) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static inline void normalized3(const float v[3], float o[3]) {
  float f = 1.0f / norm3(v);
  o[0] = f * v[0]; o[1] = f * v[1]; o[2] = f * v[2];
}
static inline void cross3(const float a[3], const float b[3], float o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

void look_at(const float pos[3], const float up_in[3], const float look[3], float forward[3], float right[3], float up[3]) {
  float d[3] = {look[0] - pos[0], look[1] - pos[1], look[2] - pos[2]}, t[3];
  normalized3(d, forward);
  cross3(forward, up_in, t); normalized3(t, right);
  cross3(right, forward, t); normalized3(t, up);
}

// ---- STL ---------------------------------------------------------------------------------------------

// a whole token must be one number ("0.5x" or "1e" is malformed, not 0.5 / 1): same rule as the Python mirror's float()
// grammar: [+-] ( digits [. digits] | . digits ) [ (e|E) [+-] digits ]  |  [+-] inf | infinity | nan   (no hex floats)
static bool parse_number(const std::string &tok, double &out) {
  size_t i = 0;
  const size_t n = tok.size();
  if (i < n && (tok[i] == '+' || tok[i] == '-')) i++;
  std::string rest;
  for (size_t k = i; k < n; k++) rest.push_back((char)tolower((unsigned char)tok[k]));
  if (rest != "inf" && rest != "infinity" && rest != "nan") {
    size_t digits = 0;
    while (i < n && isdigit((unsigned char)tok[i])) { i++; digits++; }
    if (i < n && tok[i] == '.') { i++; while (i < n && isdigit((unsigned char)tok[i])) { i++; digits++; } }
    if (!digits) return false;
    if (i < n && (tok[i] == 'e' || tok[i] == 'E')) {
      i++;
      if (i < n && (tok[i] == '+' || tok[i] == '-')) i++;
      size_t ed = 0;
      while (i < n && isdigit((unsigned char)tok[i])) { i++; ed++; }
      if (!ed) return false;
    }
    if (i != n) return false;
  }
  out = strtod(tok.c_str(), nullptr);
  return true;
}
bool read_stl(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open mesh file '" + path + "'"; return false; }
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  if (data.size() >= 84) {
    uint32_t n;
    memcpy(&n, data.data() + 80, 4);
    if (84ull + 50ull * n == data.size()) {
      for (uint32_t i = 0; i < n; i++) {
        float v[9];
        memcpy(v, data.data() + 84 + 50ull * i + 12, 36);   // skip the stored normal (the reference recomputes it)
        p1.insert(p1.end(), v, v + 3); p2.insert(p2.end(), v + 3, v + 6); p3.insert(p3.end(), v + 6, v + 9);
      }
      return true;
    }
  }
  std::istringstream in(data);
  std::string tok;
  std::vector<float> verts;
  while (in >> tok) {
    if (tok == "vertex") {
      double c[3];
      for (double &x : c) {
        std::string num;
        if (!(in >> num) || !parse_number(num, x)) { err = "bad vertex in ASCII STL '" + path + "'"; return false; }
      }
      verts.push_back((float)c[0]); verts.push_back((float)c[1]); verts.push_back((float)c[2]);
    }
  }
  if (verts.empty() || verts.size() % 9) { err = "cannot read STL file '" + path + "'"; return false; }
  for (size_t i = 0; i < verts.size(); i += 9) {
    p1.insert(p1.end(), &verts[i], &verts[i] + 3); p2.insert(p2.end(), &verts[i + 3], &verts[i + 3] + 3);
    p3.insert(p3.end(), &verts[i + 6], &verts[i + 6] + 3);
  }
  return true;
}

// ---- Wavefront OBJ (mesh import breadth; PLY and OFF follow below) --------------------------------
// The reference accepts whatever Assimp reads (inc/default_schema.hpp:516-545: aiProcess_Triangulate, faces in file
// order, positions only).  Here: `v x y z` and `f a b c ...` with a/b/c, a//c, a/b forms and negative (relative)
// indices; polygons are split as a fan from their first vertex; everything else (vt, vn, o, g, s, usemtl) is ignored.
bool read_obj(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err) {
  std::ifstream f(path);
  if (!f) { err = "cannot open mesh file '" + path + "'"; return false; }
  std::vector<float> v;
  std::string line;
  size_t faces = 0;
  while (std::getline(f, line)) {
    std::istringstream in(line);
    std::string tag;
    if (!(in >> tag)) continue;
    if (tag == "v") {
      double c[3];
      for (double &x : c) {
        std::string num;
        if (!(in >> num) || !parse_number(num, x)) { err = "bad vertex in OBJ '" + path + "'"; return false; }
      }
      v.push_back((float)c[0]); v.push_back((float)c[1]); v.push_back((float)c[2]);
    } else if (tag == "f") {
      std::vector<long> idx;
      std::string tok;
      while (in >> tok) {
        char *end = nullptr;
        long i = strtol(tok.c_str(), &end, 10);   // the part before the first '/'
        const bool whole = end == tok.c_str() + tok.size();   // not `*end == 0`: a token may hold an embedded NUL byte
        if (end == tok.c_str() || (!whole && *end != '/')) { err = "bad face index in OBJ '" + path + "'"; return false; }
        const long nv = (long)(v.size() / 3);
        if (i < 0) i = nv + i + 1;
        if (i < 1 || i > nv) { err = "face index out of range in OBJ '" + path + "'"; return false; }
        idx.push_back(i - 1);
      }
      for (size_t k = 1; k + 1 < idx.size(); k++) {
        p1.insert(p1.end(), &v[3 * idx[0]], &v[3 * idx[0]] + 3);
        p2.insert(p2.end(), &v[3 * idx[k]], &v[3 * idx[k]] + 3);
        p3.insert(p3.end(), &v[3 * idx[k + 1]], &v[3 * idx[k + 1]] + 3);
        faces++;
      }
    }
  }
  if (!faces) { err = "no faces in OBJ file '" + path + "'"; return false; }
  return true;
}

// fan triangulation of one polygon (indices into the flat xyz array `v`) in face order, like read_obj
static void emit_fan(const std::vector<float> &v, const std::vector<long> &idx, std::vector<float> &p1, std::vector<float> &p2,
                     std::vector<float> &p3, size_t &faces) {
  for (size_t k = 1; k + 1 < idx.size(); k++) {
    p1.insert(p1.end(), &v[3 * idx[0]], &v[3 * idx[0]] + 3);
    p2.insert(p2.end(), &v[3 * idx[k]], &v[3 * idx[k]] + 3);
    p3.insert(p3.end(), &v[3 * idx[k + 1]], &v[3 * idx[k + 1]] + 3);
    faces++;
  }
}

// ---- Stanford PLY (ASCII, binary little / big endian) ---------------------------------------------------------------------
// header: `element <name> <count>` blocks with `property <type> <name>` / `property list <count type> <item type> <name>` lines.
// Read: x, y, z of element `vertex` (any scalar type, converted to float) and the list property vertex_indices / vertex_index of
// element `face`; every other property and element is parsed only to be skipped.  Polygons are fan-triangulated in face order.
namespace {
struct PlyProp { bool list = false; int type = 0, count_type = 0; std::string name; };
struct PlyElem { std::string name; size_t count = 0; std::vector<PlyProp> props; };
// type codes: 1 int8 2 uint8 3 int16 4 uint16 5 int32 6 uint32 7 float32 8 float64
int ply_type(const std::string &t) {
  static const char *names[][2] = {{"char", "int8"}, {"uchar", "uint8"}, {"short", "int16"}, {"ushort", "uint16"},
                                   {"int", "int32"}, {"uint", "uint32"}, {"float", "float32"}, {"double", "float64"}};
  for (int i = 0; i < 8; i++) if (t == names[i][0] || t == names[i][1]) return i + 1;
  return 0;
}
const size_t ply_size[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
bool ply_binary_scalar(const std::string &d, size_t &at, int type, bool big, double &out) {
  const size_t n = ply_size[type];
  if (at + n > d.size()) return false;
  unsigned char b[8];
  for (size_t i = 0; i < n; i++) b[i] = (unsigned char)d[at + (big ? n - 1 - i : i)];   // to little endian
  at += n;
  switch (type) {
    case 1: { int8_t x; memcpy(&x, b, 1); out = x; break; }
    case 2: { uint8_t x; memcpy(&x, b, 1); out = x; break; }
    case 3: { int16_t x; memcpy(&x, b, 2); out = x; break; }
    case 4: { uint16_t x; memcpy(&x, b, 2); out = x; break; }
    case 5: { int32_t x; memcpy(&x, b, 4); out = x; break; }
    case 6: { uint32_t x; memcpy(&x, b, 4); out = x; break; }
    case 7: { float x; memcpy(&x, b, 4); out = x; break; }
    default: { double x; memcpy(&x, b, 8); out = x; break; }
  }
  return true;
}
}  // namespace

bool read_ply(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { err = "cannot open mesh file '" + path + "'"; return false; }
  const std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  // ---- header: lines up to end_header
  size_t at = 0;
  auto next_line = [&](std::string &line) {
    if (at >= data.size()) return false;
    size_t e = data.find('\n', at);
    if (e == std::string::npos) e = data.size();
    line = data.substr(at, e - at);
    if (!line.empty() && line.back() == '\r') line.pop_back();
    at = e + 1;
    return true;
  };
  std::string line;
  if (!next_line(line) || line != "ply") { err = "not a PLY file: '" + path + "'"; return false; }
  int format = -1;   // 0 ascii, 1 little endian, 2 big endian
  std::vector<PlyElem> elems;
  bool ended = false;
  while (next_line(line)) {
    std::istringstream in(line);
    std::string tag;
    if (!(in >> tag) || tag == "comment" || tag == "obj_info") continue;
    if (tag == "end_header") { ended = true; break; }
    if (tag == "format") {
      std::string kind;
      in >> kind;
      format = kind == "ascii" ? 0 : kind == "binary_little_endian" ? 1 : kind == "binary_big_endian" ? 2 : -1;
    } else if (tag == "element") {
      PlyElem e;
      std::string cnt;   // a whole token of digits (an optional sign like the Python mirror's int grammar), not "9abc"
      if (!(in >> e.name >> cnt)) { err = "bad element line in PLY '" + path + "'"; return false; }
      size_t k = (cnt[0] == '+' || cnt[0] == '-') ? 1 : 0;
      bool digits = k < cnt.size() && cnt.size() - k <= 18;
      for (size_t j = k; j < cnt.size(); j++) digits = digits && isdigit((unsigned char)cnt[j]);
      const long long n = digits ? strtoll(cnt.c_str(), nullptr, 10) : -1;
      if (n < 0) { err = "bad element line in PLY '" + path + "'"; return false; }
      e.count = (size_t)n;
      elems.push_back(e);
    } else if (tag == "property") {
      if (elems.empty()) { err = "property before any element in PLY '" + path + "'"; return false; }
      PlyProp pr;
      std::string t;
      in >> t;
      if (t == "list") {
        std::string ct, it;
        in >> ct >> it >> pr.name;
        pr.list = true; pr.count_type = ply_type(ct); pr.type = ply_type(it);
        if (!pr.count_type || pr.count_type >= 7) { err = "bad list count type in PLY '" + path + "'"; return false; }
      } else {
        pr.type = ply_type(t);
        in >> pr.name;
      }
      if (!pr.type || pr.name.empty()) { err = "bad property line in PLY '" + path + "'"; return false; }
      elems.back().props.push_back(pr);
    }
  }
  if (!ended || format < 0) { err = "bad PLY header in '" + path + "'"; return false; }
  // ---- body
  std::istringstream body;   // ASCII: one token stream
  if (format == 0) body.str(data.substr(at < data.size() ? at : data.size()));
  auto scalar = [&](int type, double &out) {
    if (format == 0) {
      std::string tok;
      return (bool)(body >> tok) && parse_number(tok, out);
    }
    return ply_binary_scalar(data, at, type, format == 2, out);
  };
  std::vector<float> v;
  size_t faces = 0;
  bool saw_vertex = false;
  for (const PlyElem &e : elems) {
    const bool is_vertex = e.name == "vertex", is_face = e.name == "face";
    int ix = -1, iy = -1, iz = -1, il = -1;
    for (size_t k = 0; k < e.props.size(); k++) {
      const PlyProp &pr = e.props[k];
      if (is_vertex && !pr.list) { if (pr.name == "x") ix = (int)k; else if (pr.name == "y") iy = (int)k; else if (pr.name == "z") iz = (int)k; }
      if (is_face && pr.list && (pr.name == "vertex_indices" || pr.name == "vertex_index")) il = (int)k;
    }
    if (is_vertex && (ix < 0 || iy < 0 || iz < 0)) { err = "PLY vertex element without x / y / z in '" + path + "'"; return false; }
    if (is_face && il < 0) { err = "PLY face element without vertex_indices in '" + path + "'"; return false; }
    if (is_face && !saw_vertex) { err = "PLY face element before the vertex element in '" + path + "'"; return false; }
    saw_vertex = saw_vertex || is_vertex;
    if (e.props.empty()) continue;   // rows without properties hold nothing (and a huge count must not spin here)
    for (size_t r = 0; r < e.count; r++) {
      double xyz[3] = {0, 0, 0};
      std::vector<long> idx;
      for (size_t k = 0; k < e.props.size(); k++) {
        const PlyProp &pr = e.props[k];
        double x;
        if (!pr.list) {
          if (!scalar(pr.type, x)) { err = "truncated or malformed PLY body in '" + path + "'"; return false; }
          if ((int)k == ix) xyz[0] = x; else if ((int)k == iy) xyz[1] = x; else if ((int)k == iz) xyz[2] = x;
          continue;
        }
        double cnt;
        if (!scalar(pr.count_type, cnt) || cnt < 0 || cnt > 1e6 || cnt != (double)(long)cnt) { err = "bad list length in PLY '" + path + "'"; return false; }
        for (long j = 0; j < (long)cnt; j++) {
          if (!scalar(pr.type, x)) { err = "truncated or malformed PLY body in '" + path + "'"; return false; }
          if ((int)k == il) {
            const long nv = (long)(v.size() / 3);
            if (x != (double)(long)x || x < 0 || (long)x >= nv) { err = "face index out of range in PLY '" + path + "'"; return false; }
            idx.push_back((long)x);
          }
        }
      }
      if (is_vertex) { v.push_back((float)xyz[0]); v.push_back((float)xyz[1]); v.push_back((float)xyz[2]); }
      if (is_face) emit_fan(v, idx, p1, p2, p3, faces);
    }
  }
  if (!faces) { err = "no faces in PLY file '" + path + "'"; return false; }
  return true;
}

// ---- Object File Format (OFF) ---------------------------------------------------------------------------------------------
// `OFF`, then `nv nf ne`, nv lines `x y z`, nf lines `n i0 .. i(n-1) [colour]`; `#` starts a comment; counts may share the OFF line.
bool read_off(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err) {
  std::ifstream f(path);
  if (!f) { err = "cannot open mesh file '" + path + "'"; return false; }
  std::vector<std::vector<std::string>> lines;   // the non-empty, comment-free lines as tokens
  std::string line;
  while (std::getline(f, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.resize(h);
    std::istringstream in(line);
    std::vector<std::string> toks;
    std::string t;
    while (in >> t) toks.push_back(t);
    if (!toks.empty()) lines.push_back(toks);
  }
  if (lines.empty() || lines[0][0] != "OFF") { err = "not an OFF file: '" + path + "'"; return false; }
  size_t row = 0;
  std::vector<std::string> head(lines[0].begin() + 1, lines[0].end());
  if (head.empty()) { if (lines.size() < 2) { err = "truncated OFF file '" + path + "'"; return false; } head = lines[1]; row = 2; } else row = 1;
  double nv = -1, nf = -1;
  if (head.size() < 2 || !parse_number(head[0], nv) || !parse_number(head[1], nf) || nv < 0 || nf < 0 || nv != (double)(long)nv || nf != (double)(long)nf) {
    err = "bad counts in OFF file '" + path + "'";
    return false;
  }
  if (lines.size() < row + (size_t)nv + (size_t)nf) { err = "truncated OFF file '" + path + "'"; return false; }
  std::vector<float> v;
  for (size_t i = 0; i < (size_t)nv; i++, row++) {
    double c[3];
    if (lines[row].size() < 3 || !parse_number(lines[row][0], c[0]) || !parse_number(lines[row][1], c[1]) || !parse_number(lines[row][2], c[2])) {
      err = "bad vertex in OFF '" + path + "'";
      return false;
    }
    v.push_back((float)c[0]); v.push_back((float)c[1]); v.push_back((float)c[2]);
  }
  size_t faces = 0;
  for (size_t i = 0; i < (size_t)nf; i++, row++) {
    double n;
    if (!parse_number(lines[row][0], n) || n < 0 || n != (double)(long)n || lines[row].size() < 1 + (size_t)n) { err = "bad face in OFF '" + path + "'"; return false; }
    std::vector<long> idx;
    for (size_t k = 0; k < (size_t)n; k++) {
      double x;
      if (!parse_number(lines[row][1 + k], x) || x != (double)(long)x || x < 0 || x >= nv) { err = "face index out of range in OFF '" + path + "'"; return false; }
      idx.push_back((long)x);
    }
    emit_fan(v, idx, p1, p2, p3, faces);
  }
  if (!faces) { err = "no faces in OFF file '" + path + "'"; return false; }
  return true;
}

static bool read_mesh(const std::string &path, std::vector<float> &p1, std::vector<float> &p2, std::vector<float> &p3, std::string &err) {
  const size_t dot = path.rfind('.');
  std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
  for (auto &ch : ext) ch = (char)tolower((unsigned char)ch);
  if (ext == "obj") return read_obj(path, p1, p2, p3, err);
  if (ext == "ply") return read_ply(path, p1, p2, p3, err);
  if (ext == "off") return read_off(path, p1, p2, p3, err);
  return read_stl(path, p1, p2, p3, err);
}

// ---- coercion (inc/json_helpers.hpp:88-126) -------------------------------------------------------------
static bool get_num(const JsonValue &o, const char *key, double &out, std::string &err, const double *def = nullptr) {
  const JsonValue *v = o.find(key);
  if (!v) {
    if (def) { out = *def; return true; }
    err = std::string("Cannot find key '") + key + "' in object.";
    return false;
  }
  if (!v->is_number()) { err = std::string("Expected a value of type number for '") + key + "'."; return false; }
  out = v->num;
  return true;
}
static bool get_vec(const JsonValue &o, const char *key, float out[3], std::string &err, const float *def = nullptr) {
  const JsonValue *v = o.find(key);
  if (!v) {
    if (def) { memcpy(out, def, 12); return true; }
    err = std::string("Cannot find key '") + key + "' in object.";
    return false;
  }
  if (!v->is_array() || v->arr.size() != 3 || !v->arr[0].is_number() || !v->arr[1].is_number() || !v->arr[2].is_number()) {
    err = std::string("Expected a 3-element array of numbers for '") + key + "'.";
    return false;
  }
  for (int i = 0; i < 3; i++) out[i] = (float)v->arr[i].num;
  return true;
}
static void push3(std::vector<float> &v, const float p[3]) { v.insert(v.end(), p, p + 3); }

bool load_scene_text(const std::string &text, const LoadOptions &opt, FlatScene &s, std::vector<std::string> &errors) {
  s = FlatScene();
  JsonValue root;
  std::string err;
  JsonParser parser(text);
  if (!parser.parse(root, err)) { errors.push_back("JSON parse error: " + err); return false; }
  if (!root.is_object()) { errors.push_back("Value is not a JSON object."); return false; }
  const size_t before = errors.size();

  // ---- materials first: object material indices are range-checked (the reference would read out of bounds)
  const JsonValue *mats = root.find("materials");
  if (!mats || !mats->is_array()) errors.push_back("Could not find 'materials' array.");
  else {
    for (size_t i = 0; i < mats->arr.size(); i++) {
      const JsonValue &m = mats->arr[i];
      auto bad = [&](const std::string &why) { errors.push_back("Error while loading material #" + std::to_string(i) + ": " + why); };
      if (!m.is_object()) { bad("Value is not a JSON object."); continue; }
      const JsonValue *ty = m.find("type");
      if (!(ty && ty->is_string() && ty->str == "solid") && !(opt.accept_aliases && !ty)) { bad("No matching material type (expected \"solid\")."); continue; }
      float col[3];
      double spec, refl, ph, tr;
      const double d_spec = 0.3, d_zero = 0.0, d_ph = 32.0;   // inc/default_schema.hpp:754-764 (0.3f, 0, 32, 0)
      if (!get_vec(m, "color", col, err) || !get_num(m, "specular", spec, err, &d_spec) || !get_num(m, "reflect", refl, err, &d_zero) ||
          !get_num(m, "phong", ph, err, &d_ph) || !get_num(m, "transparency", tr, err, &d_zero)) { bad(err); continue; }
      push3(s.mat_color, col);
      s.mat_specular.push_back(m.find("specular") ? (float)spec : 0.3f);
      s.mat_reflect.push_back((float)refl); s.mat_phong.push_back((float)ph); s.mat_transparency.push_back((float)tr);
    }
  }
  const size_t n_mat = s.mat_specular.size();

  const JsonValue *lights = root.find("lights");
  if (!lights || !lights->is_array()) errors.push_back("Could not find 'lights' array.");
  else {
    const float white[3] = {1, 1, 1};
    for (size_t i = 0; i < lights->arr.size(); i++) {
      const JsonValue &l = lights->arr[i];
      auto bad = [&](const std::string &why) { errors.push_back("Error while loading light #" + std::to_string(i) + ": " + why); };
      if (!l.is_object()) { bad("Value is not a JSON object."); continue; }
      const JsonValue *ty = l.find("type");
      float v[3], c[3];
      if (ty && ty->is_string() && ty->str == "sun") {
        if (!get_vec(l, "direction", v, err) || !get_vec(l, "color", c, err, white)) { bad(err); continue; }
        s.light_kind.push_back(CUTRACE_LIGHT_SUN);
      } else if (ty && ty->is_string() && ty->str == "point") {
        const char *key = (opt.accept_aliases && !l.find("point") && l.find("position")) ? "position" : "point";
        if (!get_vec(l, key, v, err) || !get_vec(l, "color", c, err, white)) { bad(err); continue; }
        s.light_kind.push_back(CUTRACE_LIGHT_POINT);
      } else { bad("No matching light type (expected \"sun\" or \"point\")."); continue; }
      push3(s.light_vec, v); push3(s.light_color, c);
    }
  }

  const JsonValue *objs = root.find("objects");
  if (!objs || !objs->is_array()) errors.push_back("Could not find 'objects' array.");
  else {
    for (size_t i = 0; i < objs->arr.size(); i++) {
      const JsonValue &o = objs->arr[i];
      auto bad = [&](const std::string &why) { errors.push_back("Error while loading object #" + std::to_string(i) + ": " + why); };
      if (!o.is_object()) { bad("Value is not a JSON object."); continue; }
      const JsonValue *ty = o.find("type");
      std::string type = (ty && ty->is_string()) ? ty->str : "";
      if (opt.accept_aliases && type == "model") type = "mesh";
      const uint32_t id = (uint32_t)s.obj_material.size();
      double mat_d;
      if (!get_num(o, "material", mat_d, err)) { bad(err); continue; }
      // range-check the double BEFORE the (size_t) cast of inc/json_helpers.hpp:88-93: picojson numbers like 1e300 make the cast undefined
      if (!(mat_d >= 0.0 && mat_d < (double)n_mat)) { bad("material index out of range"); continue; }
      const size_t mat = (size_t)mat_d;
      if (type == "triangle") {
        float a[3], b[3], c[3];
        const JsonValue *pts = o.find("points");
        if (opt.accept_aliases && !o.find("p1") && pts && pts->is_array() && pts->arr.size() == 3) {
          JsonValue tmp; tmp.kind = JsonValue::Object; tmp.obj["p1"] = pts->arr[0]; tmp.obj["p2"] = pts->arr[1]; tmp.obj["p3"] = pts->arr[2];
          if (!get_vec(tmp, "p1", a, err) || !get_vec(tmp, "p2", b, err) || !get_vec(tmp, "p3", c, err)) { bad(err); continue; }
        } else if (!get_vec(o, "p1", a, err) || !get_vec(o, "p2", b, err) || !get_vec(o, "p3", c, err)) { bad(err); continue; }
        push3(s.tri_p1, a); push3(s.tri_p2, b); push3(s.tri_p3, c); s.tri_object.push_back(id);
        s.obj_kind.push_back(CUTRACE_OBJ_TRIANGLE);
      } else if (type == "mesh") {
        const JsonValue *file = o.find("file");
        if (!file || !file->is_string()) { bad("Cannot find key 'file' in object."); continue; }
        std::string path = file->str;
        if (!opt.base_dir.empty() && !path.empty() && path[0] != '/') path = opt.base_dir + "/" + path;
        size_t n0 = s.tri_p1.size() / 3;
        if (!read_mesh(path, s.tri_p1, s.tri_p2, s.tri_p3, err)) { bad(err); continue; }
        s.tri_object.insert(s.tri_object.end(), s.tri_p1.size() / 3 - n0, id);
        s.obj_kind.push_back(CUTRACE_OBJ_MESH);
      } else if (type == "plane") {
        float p[3], n[3];
        if (!get_vec(o, "point", p, err) || !get_vec(o, "normal", n, err)) { bad(err); continue; }
        push3(s.pl_point, p); push3(s.pl_normal, n); s.pl_object.push_back(id);
        s.obj_kind.push_back(CUTRACE_OBJ_PLANE);
      } else if (type == "sphere") {
        float c[3];
        double r;
        if (!get_vec(o, "center", c, err) || !get_num(o, "radius", r, err)) { bad(err); continue; }
        push3(s.sph_center, c); s.sph_radius.push_back((float)r); s.sph_object.push_back(id);
        s.obj_kind.push_back(CUTRACE_OBJ_SPHERE);
      } else { bad("No matching object type (expected triangle, mesh, plane or sphere)."); continue; }
      s.obj_material.push_back((uint32_t)mat);
    }
  }

  const JsonValue *cam = root.find("camera");
  if (!cam || !cam->is_object()) errors.push_back("Could not find 'camera' object or it's invalid.");
  else {
    // defaults of inc/default_schema.hpp:835-842, only reachable with accept_aliases (the keys are mandatory)
    const float d_eye[3] = {0, 0, 0}, d_up[3] = {0, 1, 0}, d_look[3] = {0, 0, 1};
    const double d_near = 0.1f, d_far = 100.0f, d_w = 1920, d_h = 1080, d_amb = 0.1f;
    const bool al = opt.accept_aliases;
    float eye[3], up[3], look[3];
    double nearp, farp, w, h, amb;
    if (!get_vec(*cam, "eye", eye, err, al ? d_eye : nullptr) || !get_vec(*cam, "up", up, err, al ? d_up : nullptr) ||
        !get_vec(*cam, "look", look, err, al ? d_look : nullptr) || !get_num(*cam, "near_plane", nearp, err, al ? &d_near : nullptr) ||
        !get_num(*cam, "far_plane", farp, err, al ? &d_far : nullptr) || !get_num(*cam, "width", w, err, al ? &d_w : nullptr) ||
        !get_num(*cam, "height", h, err, al ? &d_h : nullptr) || !get_num(*cam, "ambient", amb, err, al ? &d_amb : nullptr)) {
      errors.push_back("Could not find 'camera' object or it's invalid: " + err);
    } else if (!(w >= 1 && h >= 1 && w < 65536.0 * 4 && h < 65536.0 * 4)) {
      errors.push_back("Could not find 'camera' object or it's invalid: width/height out of range");
    } else {
      memcpy(s.cam_pos, eye, 12);
      look_at(eye, up, look, s.cam_forward, s.cam_right, s.cam_up);
      s.near_plane = (float)nearp; s.far_plane = (float)farp; s.ambient = (float)amb;
      s.width = (uint32_t)(size_t)w; s.height = (uint32_t)(size_t)h;
    }
  }
  return errors.size() == before;
}

bool load_scene_file(const std::string &path, const LoadOptions &opt, FlatScene &out, std::vector<std::string> &errors) {
  std::ifstream f(path);
  if (!f) { errors.push_back("Error while loading file '" + path + "': cannot open"); return false; }
  std::stringstream ss;
  ss << f.rdbuf();
  return load_scene_text(ss.str(), opt, out, errors);
}

}  // namespace cthost
