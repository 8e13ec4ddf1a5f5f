// bvh.cuh — GPU LBVH build (Morton codes -> LSD radix sort -> Karras hierarchy -> refit -> emit).
// Replaces the reference's only culling structure, one AABB per mesh followed by a linear loop over
// every triangle (inc/default_schema.hpp:99-144), and the linear object walk of inc/ray_cast.hpp:37-52.
#ifndef CUTRACE_B200_BVH_CUH
#define CUTRACE_B200_BVH_CUH
#include <string>
#include "common.cuh"

namespace ctb {

struct BvhInput {
  const float *d_p1 = nullptr, *d_p2 = nullptr, *d_p3 = nullptr;  // device, n_tri*3 floats each
  const uint32_t *d_tri_obj = nullptr;
  uint32_t n_tri = 0;
  const float *d_sph_center = nullptr, *d_sph_radius = nullptr;
  const uint32_t *d_sph_obj = nullptr;
  uint32_t n_sph = 0;
  uint32_t n_objects = 0;     // object indices of the primitives are range-checked on the device
  uint32_t leaf_size = 3;
  bool sah_treelets = true;   // false: CUTRACE_FLAG_FAST_BUILD, the LBVH topology is kept
  cudaStream_t stream = nullptr;
};

struct BvhResult {
  Node4 *nodes4 = nullptr;    // CTB_BVH4 builds only: n_nodes4 nodes of the 4-wide tree; node 0 is the root when root >= 0
  uint32_t n_nodes4 = 0;
  Node *nodes = nullptr;      // n_nodes compacted live nodes; node 0 is the root when root >= 0
  PrimRec *prims = nullptr;   // n_prims records in leaf order
  uint32_t n_nodes = 0, n_prims = 0;
  int root = CTB_SENTINEL;
  uint32_t depth = 0;         // deepest leaf (number of internal nodes on the path)
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
};

// returns cutrace_status; on failure `err` has the message and nothing is left allocated
int build_bvh(const BvhInput &in, BvhResult &out, std::string &err);
// moves the first S nodes in breadth-first order to the front of bvh.nodes (the part kernels stage in shared memory)
int reorder_top(BvhResult &bvh, uint32_t S, cudaStream_t stream, uint32_t *n_top_out, std::string &err);
// device-side self check, see cutrace_validate_bvh()
int validate_bvh(const BvhResult &bvh, cudaStream_t stream, std::string &err);
// per-object AABB (min / max over the vertices of the object's triangles, like cutrace's host code) + "is a mesh" flag for the
// reference's mesh pre-test (trace.cuh: mesh_gate).  d_obj_kind may be NULL (an object with more than one triangle is a mesh).
int build_object_bounds(const float *d_p1, const float *d_p2, const float *d_p3, const uint32_t *d_tri_obj, uint32_t n_tri, uint32_t n_objects,
                        const uint32_t *d_obj_kind, ObjBound **out, cudaStream_t st, std::string &err);
// radix sort entry point (exposed for the sort unit test): sorts n (key,value) pairs ascending, stable.
// d_keys/d_vals are overwritten with the result; tmp buffers are allocated internally.
int radix_sort_pairs(uint64_t *d_keys, uint32_t *d_vals, uint32_t n, cudaStream_t stream, std::string &err);

}  // namespace ctb
#endif
