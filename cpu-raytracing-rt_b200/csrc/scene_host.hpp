// scene_host.hpp — host-side scene model: what the reference keeps in `Scene`
// (reference src/scene.rs:82-90) after Scene::new / gltf::build_scene, i.e. world-space
// primitives in post-build order plus the six hierarchies of src/bvh.rs.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "rtmath.cuh"

namespace rt {

struct Box6 {                       // src/aabb.rs:5-9
    double lo[3], hi[3];
};

struct HostNode {                   // src/bvh.rs:49-54
    Box6 box;
    uint32_t left, right;           // RT_NO_CHILD for leaves
    uint32_t first, count;          // primitive_indices
};

struct HostTri {                    // src/primitives/triangle.rs:6-17 + src/scene.rs:30-34
    d3 a, ba, ca, ng;
    double inv_area;
    d3 na, nb, nc;
    uint32_t material;
    uint32_t orig;                  // position in the list that was fed to BVH::new
    Box6 box;
};

struct HostShape {                  // src/scene.rs:21-27 Primitive<Box|Ellipsoid|Plane>
    int32_t type;
    uint32_t material;
    d3 param, position;
    quat rotation;
    uint32_t orig;
    Box6 box;
};

template <class P>
struct HostBvh {                    // src/bvh.rs:6-9
    std::vector<P> prims;
    std::vector<HostNode> nodes;
    uint32_t depth = 0;
};

// Structural check of a node array that did not come from build_bvh (a cache file, rt_scene_from_built):
// the reference's pre-order numbering (left child = node + 1, src/bvh.rs:104-108), every leaf range
// inside [0, n_prims), and EXACTLY ONE parent per non-root node — a node shared by two parents (a
// DAG) would make the real path length exceed the depth computed here, and the traversal stacks are
// sized from that depth.  Returns the depth (root = 1); throws std::runtime_error(prefix + reason).
uint32_t validate_hierarchy(const std::vector<HostNode> &nodes, uint64_t n_prims, const char *prefix);

struct HostScene {
    uint32_t width = 0, height = 0, samples = 0, ray_depth = 0;
    d3 bg_color{0, 0, 0};
    d3 cam_position{0, 0, 0}, cam_right{1, 0, 0}, cam_up{0, 1, 0}, cam_forward{0, 0, 1};
    double tan_half_fov_x = 1.0, tan_half_fov_y = 1.0;
    std::vector<rt_material> materials;
    std::vector<HostShape> planes;
    HostBvh<HostShape> boxes, ellipsoids, light_boxes, light_ellipsoids;
    HostBvh<HostTri> tris, light_tris;
};

// Full-sweep SAH build of src/bvh.rs:12-17, :75-144 over `n` primitive boxes.
// Outputs the pre-order node array and order[i] = input index of the primitive at slot i.
// Ties in the midpoint sort keep their current relative order ("BVH tie order").
void build_bvh(const Box6 *boxes, size_t n, std::vector<HostNode> &nodes, std::vector<uint32_t> &order,
               uint32_t &depth, int max_threads = 0);

// Scene::new + make_scenes (src/scene.rs:180-223) / build_scene (src/gltf/scene_builder.rs:9-22).
// Throws std::runtime_error on invalid input.
void build_scene(const rt_raw_scene &raw, HostScene &out);

void apply_view(const rt_raw_scene &raw, HostScene &s);     // image size, samples, depth, background, camera

// On-disk cache of the built hierarchies (scene_cache.cpp), keyed by a hash of the geometry.
uint64_t geometry_hash(const rt_raw_scene &raw);
std::string scene_cache_path(const std::string &dir, uint64_t hash);
void save_scene_cache(const std::string &path, uint64_t hash, const HostScene &s);
// false: absent or another mesh (hash or — when raw is given — the primitive / material counts differ)
bool load_scene_cache(const std::string &path, uint64_t hash, const rt_raw_scene *raw, HostScene &s);
bool build_scene_cached(const rt_raw_scene &raw, const std::string &cache_dir, HostScene &out);   // true: cache hit

// Loaders (throw std::runtime_error where the reference would panic).
struct RawSceneStorage {            // owns the arrays an rt_raw_scene points to
    rt_raw_scene raw{};
    std::vector<rt_material> materials;
    std::vector<rt_shape> shapes;
    std::vector<double> tri_pos, tri_nrm, tri_xform;
    std::vector<uint32_t> tri_mat;
    void link();
};
void parse_text_scene(std::istream &in, RawSceneStorage &out);                     // src/scene_parser.rs
void load_gltf_scene(const std::string &path, uint32_t w, uint32_t h, uint32_t spp,
                     RawSceneStorage &out);                                          // src/gltf/*
void save_ppm(const std::string &path, uint32_t w, uint32_t h, const double *rgb);  // src/ppm.rs

}  // namespace rt

struct rt_device_image;                 // flattened, page-locked copy of the scene (capi.cpp)
struct rt_scene {
    rt::HostScene s;
    mutable std::once_flag image_once;                  // guards the lazy build below (concurrent uploads from one thread per GPU)
    mutable std::shared_ptr<rt_device_image> image;     // built on the first rt_upload_scene
};
