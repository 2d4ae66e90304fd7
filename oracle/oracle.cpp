// oracle.cpp — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference render hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library, and only as the checker or the reported CPU baseline — the product
// (cpu-raytracing-rt_b200/) never links, imports or calls anything under oracle/.
//
// What it restates (reference = uncerso/cpu-raytracing-rt, all f64, src/types.rs:5), function by
// function with the reference's operation order and recursion structure:
//   aabb.rs, bvh.rs (build + both traversals), primitives/{triangle,box,ellipsoid,plane}.rs,
//   intersections.rs, intersection_probability.rs, ray_sampler.rs, raytrace.rs, camera.rs,
//   postprocessing.rs, scene.rs (Primitive::new, TrianglePrimitive::new, make_scenes),
//   gltf/scene_builder.rs (instantiate, flat_triangles, flat_light_triangles, cof), ppm.rs bytes.
//
// PARITY STATUS (also in DESIGN.md): the reference cannot be compiled here (no rustc/cargo, no
// crate sources).  This restatement is pinned by the reference's own 13 inline tests
// (tests/test_oracle_kat.py); beyond those it is "parity unpinned" against a real Rust build in
// three places: ulp-level operation order inside cgmath 0.18 (oracle_math.hpp), the tie order
// of `sort_unstable_by` in the BVH build (this file sorts ties in STABLE order, which equals
// Rust's for ranges <= 20 elements where both std algorithms use insertion sort), and the OS-seeded
// thread_rng (replaced by a counter-based stream with rand-0.8-equivalent distributions).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../include/rt_b200.h"
#include "oracle_math.hpp"

namespace orc {

// ------------------------------------------------------------------------------------------------
// aabb.rs
struct AABB {
    V3 min, max;
};
static inline F fmin_(F a, F b) { return a < b ? a : b; }   // aabb.rs:34-36
static inline F fmax_(F a, F b) { return b < a ? a : b; }   // aabb.rs:38-40
static inline AABB aabb_empty() { return {v3(INF, INF, INF), v3(-INF, -INF, -INF)}; }  // :16-21
static inline void extend(AABB &b, V3 v) {                   // aabb.rs:23-26
    b.min = v3(fmin_(b.min.x, v.x), fmin_(b.min.y, v.y), fmin_(b.min.z, v.z));
    b.max = v3(fmax_(b.max.x, v.x), fmax_(b.max.y, v.y), fmax_(b.max.z, v.z));
}
static inline void extend_aabb(AABB &b, const AABB &v) {     // aabb.rs:28-31
    b.min = v3(fmin_(b.min.x, v.min.x), fmin_(b.min.y, v.min.y), fmin_(b.min.z, v.min.z));
    b.max = v3(fmax_(b.max.x, v.max.x), fmax_(b.max.y, v.max.y), fmax_(b.max.z, v.max.z));
}
static inline F safe_min(F a, F b) {                         // aabb.rs:90-98
    if (!std::isfinite(a)) return b;
    if (!std::isfinite(b)) return a;
    return fmin_(a, b);
}
static inline F safe_max(F a, F b) {                         // aabb.rs:100-108
    if (!std::isfinite(a)) return b;
    if (!std::isfinite(b)) return a;
    return fmax_(a, b);
}
struct Ray {                                                 // ray.rs:4-13
    V3 origin, dir;
    V3 position_at(F t) const { return origin + dir * t; }
};
static inline bool aabb_inside(const AABB &b, V3 o) {        // aabb.rs:80-87
    for (int i = 0; i < 3; ++i)
        if (o[i] < b.min[i] || b.max[i] < o[i]) return false;
    return true;
}
// AABB::intersects, aabb.rs:51-78.  Returns false for None.
static inline bool aabb_intersects(const AABB &b, const Ray &ray, F &t_out) {
    for (int i = 0; i < 3; ++i)
        if (ray.dir[i] == 0.0 && (ray.origin[i] < b.min[i] || b.max[i] < ray.origin[i])) return false;
    if (aabb_inside(b, ray.origin)) { t_out = 0.0; return true; }
    V3 t_min = div_ew(b.min - ray.origin, ray.dir);
    V3 t_max = div_ew(b.max - ray.origin, ray.dir);
    V3 t1 = v3(safe_min(t_min.x, t_max.x), safe_min(t_min.y, t_max.y), safe_min(t_min.z, t_max.z));
    V3 t2 = v3(safe_max(t_min.x, t_max.x), safe_max(t_min.y, t_max.y), safe_max(t_min.z, t_max.z));
    F t_near = safe_max(safe_max(t1.x, t1.y), t1.z);
    F t_far = safe_min(safe_min(t2.x, t2.y), t2.z);
    if (t_near > t_far) return false;
    if (0.0 <= t_near) { t_out = t_near; return true; }
    if (0.0 <= t_far) { t_out = t_far; return true; }
    return false;
}

// ------------------------------------------------------------------------------------------------
// intersections.rs:10-29
struct Intersection {
    F t;
    V3 geometry_normal, shading_normal;
    bool inside;
};
static inline Intersection with_geometry_normals(F t, V3 n, bool inside) { return {t, n, n, inside}; }
static inline Intersection with_rotated_normal(const Intersection &i, Quat q) {  // :31-40
    return {i.t, normalize(rotate(q, i.geometry_normal)), normalize(rotate(q, i.shading_normal)), i.inside};
}

// ------------------------------------------------------------------------------------------------
// primitives
struct Material {
    int kind;
    F ior;
    V3 color, emission;
};
struct Tri {                                                 // triangle.rs:6-17 + scene.rs:30-34
    V3 a, ba, ca, geometry_normal;
    F inverted_area;
    V3 na, nb, nc;
    uint32_t material;
    AABB aabb;
    uint32_t orig;                                           // input index (for prim_order)
};
static inline void tri_props(V3 a, V3 b, V3 c, Tri &t) {     // triangle.rs:41-47
    t.a = a;
    t.ba = b - a;
    t.ca = c - a;
    V3 sized_normal = cross(t.ba, t.ca);
    F area = std::sqrt(dot(sized_normal, sized_normal)) / 2.0;
    t.geometry_normal = normalize(sized_normal);
    t.inverted_area = 1.0 / area;
}
// Triangle::intersection, triangle.rs:49-80
static inline bool tri_intersection(const Tri &tr, const Ray &ray, Intersection &out) {
    V3 c0 = tr.ba, c1 = tr.ca, c2 = -ray.dir;
    M3 m{{c0, c1, c2}};
    F d = det(m);
    if (std::fabs(d) < 1e-11) return false;
    V3 r0 = cross(c1, c2) / d, r1 = cross(c2, c0) / d, r2 = cross(c0, c1) / d;  // rows after transpose
    V3 w = ray.origin - tr.a;
    F u = dot(r0, w), v = dot(r1, w), t = dot(r2, w);
    if (u < 0.0 || v < 0.0 || 1.0 < u + v || t < 0.0) return false;
    V3 normal = tr.geometry_normal;
    V3 shading = normalize(tr.na + u * (tr.nb - tr.na) + v * (tr.nc - tr.na));   // :30-32
    bool inside = dot(ray.dir, normal) > 0.0;
    out = {t, inside ? -normal : normal, inside ? -shading : shading, inside};
    return true;
}

struct Shape {                                               // scene.rs:21-27 Primitive<T>
    int type;
    V3 param, position;
    Quat rotation;
    uint32_t material;
    AABB aabb;
    uint32_t orig;
};
static inline Ray model_space_ray(const Shape &s, const Ray &ray) {  // intersections.rs:93-99
    Quat rot = conjugate(s.rotation);
    return {rotate(rot, ray.origin - s.position), rotate(rot, ray.dir)};
}
// plane.rs:11-21
static inline bool plane_intersection(V3 normal, const Ray &ray, Intersection &out) {
    F nd = dot(normal, ray.dir);
    F t = -dot(normal, ray.origin) / nd;
    if (t < 0.0) return false;
    out = with_geometry_normals(t, (nd <= 0.0 ? 1.0 : -1.0) * normal, false);
    return true;
}
// box.rs:50-115
struct BoxPlane {
    F t, normal;
    int dim;
    V3 real_normal() const {
        if (dim == 0) return v3(normal, 0.0, 0.0);
        if (dim == 1) return v3(0.0, normal, 0.0);
        return v3(0.0, 0.0, normal);
    }
};
static inline BoxPlane bp_max(BoxPlane self, BoxPlane other) { return self.t < other.t ? other : self; }
static inline BoxPlane bp_min(BoxPlane self, BoxPlane other) { return self.t < other.t ? self : other; }
// returns 0 None, 1 One(far), 2 Two(near, far)
static inline int intersect_box_coef(V3 s, const Ray &ray, BoxPlane &near, BoxPlane &far) {
    bool have = false;
    BoxPlane max_t1{}, min_t2{};
    for (int i = 0; i < 3; ++i) {
        if (ray.dir[i] == 0.0 && s[i] < std::fabs(ray.origin[i])) return 0;
        if (ray.dir[i] == 0.0) continue;                     // box_planes_intersect -> None
        F t1 = (s[i] - ray.origin[i]) / ray.dir[i];
        F t2 = (-s[i] - ray.origin[i]) / ray.dir[i];
        F lo, hi, normal;
        if (t1 < t2) { lo = t1; hi = t2; normal = 1.0; } else { lo = t2; hi = t1; normal = -1.0; }
        if (have) {
            max_t1 = bp_max(BoxPlane{lo, normal, i}, max_t1);
            min_t2 = bp_min(BoxPlane{hi, normal, i}, min_t2);
        } else {
            max_t1 = BoxPlane{lo, normal, i};
            min_t2 = BoxPlane{hi, normal, i};
            have = true;
        }
    }
    if (!have) return 0;
    if (min_t2.t < max_t1.t) return 0;
    near = max_t1;
    far = min_t2;
    if (0.0 <= max_t1.t) return 2;
    if (0.0 <= min_t2.t) return 1;
    return 0;
}
// ellipsoid.rs:49-76: 0 None, 1 One(t2), 2 Two(t1,t2)
static inline int intersect_ellipsoid_coef(V3 r, const Ray &ray, F &t1, F &t2) {
    V3 origin = div_ew(ray.origin, r), dir = div_ew(ray.dir, r);
    F c = dot(origin, origin), b = dot(origin, dir), a = dot(dir, dir);
    F d = b * b - a * (c - 1.0);
    if (d < 0.0) return 0;
    F ds = std::sqrt(d);
    t1 = (-b + ds) / a;
    t2 = (-b - ds) / a;
    if (t2 < t1) std::swap(t1, t2);
    if (0.0 <= t1) return 2;
    if (0.0 <= t2) return 1;
    return 0;
}
static inline V3 ellipsoid_normal(V3 r, const Ray &ray, F t) {  // ellipsoid.rs:26,29
    return normalize(div_ew(div_ew(ray.position_at(t), r), r));
}
// Intersectable::all_intersections for Primitive<Box|Ellipsoid|Plane> (model space): count + up to 2
static inline int shape_all_intersections(const Shape &s, const Ray &world, Intersection out[2]) {
    Ray ray = model_space_ray(s, world);
    if (s.type == RT_SHAPE_BOX) {                            // box.rs:35-46
        BoxPlane n, f;
        int k = intersect_box_coef(s.param, ray, n, f);
        if (k == 1) { out[0] = with_geometry_normals(f.t, f.real_normal(), true); return 1; }
        if (k == 2) {
            out[0] = with_geometry_normals(n.t, n.real_normal(), false);
            out[1] = with_geometry_normals(f.t, f.real_normal(), true);
            return 2;
        }
        return 0;
    }
    if (s.type == RT_SHAPE_ELLIPSOID) {                      // ellipsoid.rs:34-46
        F t1, t2;
        int k = intersect_ellipsoid_coef(s.param, ray, t1, t2);
        if (k == 1) { out[0] = with_geometry_normals(t2, -ellipsoid_normal(s.param, ray, t2), true); return 1; }
        if (k == 2) {
            out[0] = with_geometry_normals(t1, ellipsoid_normal(s.param, ray, t1), false);
            out[1] = with_geometry_normals(t2, -ellipsoid_normal(s.param, ray, t2), true);
            return 2;
        }
        return 0;
    }
    return plane_intersection(s.param, ray, out[0]) ? 1 : 0;  // plane.rs:23-28
}
// Intersectable::intersection (closest) — box.rs:21-33, ellipsoid.rs:21-32, plane.rs:11-21
static inline bool shape_intersection(const Shape &s, const Ray &world, Intersection &out) {
    Ray ray = model_space_ray(s, world);
    if (s.type == RT_SHAPE_BOX) {
        BoxPlane n, f;
        int k = intersect_box_coef(s.param, ray, n, f);
        if (k == 1) { out = with_geometry_normals(f.t, f.real_normal(), true); return true; }
        if (k == 2) { out = with_geometry_normals(n.t, n.real_normal(), false); return true; }
        return false;
    }
    if (s.type == RT_SHAPE_ELLIPSOID) {
        F t1, t2;
        int k = intersect_ellipsoid_coef(s.param, ray, t1, t2);
        if (k == 1) { out = with_geometry_normals(t2, -ellipsoid_normal(s.param, ray, t2), true); return true; }
        if (k == 2) { out = with_geometry_normals(t1, ellipsoid_normal(s.param, ray, t1), false); return true; }
        return false;
    }
    return plane_intersection(s.param, ray, out);
}
static inline bool prim_intersection(const Tri &t, const Ray &r, Intersection &o) { return tri_intersection(t, r, o); }
static inline bool prim_intersection(const Shape &s, const Ray &r, Intersection &o) { return shape_intersection(s, r, o); }
static inline int prim_all_intersections(const Tri &t, const Ray &r, Intersection o[2]) {  // triangle.rs:82-87
    return tri_intersection(t, r, o[0]) ? 1 : 0;
}
static inline int prim_all_intersections(const Shape &s, const Ray &r, Intersection o[2]) {
    return shape_all_intersections(s, r, o);
}

// ------------------------------------------------------------------------------------------------
// bvh.rs
struct Counters {
    uint64_t segments = 0, light_pdf_rays = 0, paths = 0;
    uint64_t aabb_tests = 0, prim_tests = 0, light_aabb_tests = 0, light_prim_tests = 0;
    void add(const Counters &o) {
        segments += o.segments; light_pdf_rays += o.light_pdf_rays; paths += o.paths;
        aabb_tests += o.aabb_tests; prim_tests += o.prim_tests;
        light_aabb_tests += o.light_aabb_tests; light_prim_tests += o.light_prim_tests;
    }
};
static constexpr size_t NONE = ~(size_t)0;                   // usize::MAX, bvh.rs:148
struct Node {                                                // bvh.rs:49-54
    AABB aabb;
    size_t left_child, right_child;
    size_t first, count;                                     // primitive_indices: Range<usize>
};
template <class T>
struct BVH {
    std::vector<T> primitives;
    std::vector<Node> nodes;
};
static inline F aabb_score(const AABB &b) {                  // bvh.rs:115-118
    V3 s = b.max - b.min;
    return s.x * s.y + s.x * s.z + s.y * s.z;
}
// primitives.sort_unstable_by(midpoint_comparator(key)), bvh.rs:102,121,137-140.
// Ties: STABLE order (see header).  Sorted through (key, position) proxies, then permuted —
// the same permutation a stable sort of the structs themselves would produce.
template <class T>
static void sort_by_midpoint(T *p, size_t n, int axis, std::vector<std::pair<int64_t, uint32_t>> &proxy,
                             std::vector<T> &tmp) {
    proxy.resize(n);
    for (size_t i = 0; i < n; ++i) {
        F mid = (p[i].aabb.min[axis] + p[i].aabb.max[axis]) / 2.0;
        proxy[i] = {total_order_key(mid), (uint32_t)i};
    }
    std::stable_sort(proxy.begin(), proxy.end(),
                     [](const auto &a, const auto &b) { return a.first < b.first; });
    tmp.assign(p, p + n);
    for (size_t i = 0; i < n; ++i) p[i] = tmp[proxy[i].second];
}
template <class T>
struct Builder {
    std::vector<Node> &nodes;
    std::vector<AABB> forward, backward;                     // AABBSplitsBuilder, bvh.rs:224-256
    std::vector<std::pair<int64_t, uint32_t>> proxy;
    std::vector<T> tmp;
    explicit Builder(std::vector<Node> &n) : nodes(n) {}

    void make_splits(const T *e, size_t n) {                 // bvh.rs:238-255
        forward.clear();
        backward.clear();
        AABB b = aabb_empty();
        size_t cnt = n - 1;
        for (size_t i = 0; i < cnt; ++i) { extend_aabb(b, e[i].aabb); forward.push_back(b); }
        b = aabb_empty();
        for (size_t i = cnt; i >= 1; --i) { extend_aabb(b, e[i].aabb); backward.push_back(b); }
    }
    // subdivision_score, bvh.rs:120-135
    void subdivision_score(T *p, size_t n, int axis, size_t &best_first, F &best_score, int &best_type) {
        sort_by_midpoint(p, n, axis, proxy, tmp);
        make_splits(p, n);
        for (size_t i = 0; i + 1 < n; ++i) {
            size_t left_cnt = i + 1, right_cnt = n - left_cnt;
            F score = aabb_score(forward[i]) * (F)left_cnt + aabb_score(backward[backward.size() - i - 1]) * (F)right_cnt;
            if (score < best_score) { best_first = left_cnt; best_score = score; best_type = axis; }
        }
    }
    // build_nodes, bvh.rs:75-113
    size_t build(T *p, size_t n, size_t global_offset) {
        AABB aabb = aabb_empty();
        for (size_t i = 0; i < n; ++i) extend_aabb(aabb, p[i].aabb);     // compute_aabb :56-62
        if (n <= 4) {
            nodes.push_back(Node{aabb, NONE, NONE, global_offset, n});
            return nodes.size() - 1;
        }
        size_t best_first = n;
        F best_score = aabb_score(aabb) * (F)n;
        int best_type = -1;                                  // SameNode
        subdivision_score(p, n, 0, best_first, best_score, best_type);
        subdivision_score(p, n, 1, best_first, best_score, best_type);
        subdivision_score(p, n, 2, best_first, best_score, best_type);
        if (best_type < 0) {
            nodes.push_back(Node{aabb, NONE, NONE, global_offset, n});
            return nodes.size() - 1;
        }
        sort_by_midpoint(p, n, best_type, proxy, tmp);       // the re-sort, bvh.rs:102
        size_t node_index = nodes.size();
        nodes.push_back(Node{aabb, NONE, NONE, 0, 0});       // make_leaf(aabb, 0..0)
        size_t l = build(p, best_first, global_offset);
        size_t r = build(p + best_first, n - best_first, global_offset + best_first);
        nodes[node_index].left_child = l;
        nodes[node_index].right_child = r;
        return node_index;
    }
};
template <class T>
static BVH<T> bvh_new(std::vector<T> prims) {                // BVH::new, bvh.rs:12-17
    BVH<T> res;
    res.primitives = std::move(prims);
    Builder<T> b(res.nodes);
    b.build(res.primitives.data(), res.primitives.size(), 0);
    return res;
}

template <class T>
struct Best {                                                // Option<(Intersection, &T)>
    bool some = false;
    Intersection isect{};
    size_t prim = 0;
    F t_or_inf() const { return some ? isect.t : INF; }
};
template <class T>
static inline void update_best(const Intersection &i, size_t prim, Best<T> &best) {  // bvh.rs:213-222
    if (best.some) {
        if (i.t < best.isect.t) { best.isect = i; best.prim = prim; }
    } else {
        best.some = true; best.isect = i; best.prim = prim;
    }
}
// Node::intersection, bvh.rs:151-186
template <class T>
static void node_intersection(const BVH<T> &bvh, const Node &node, const Ray &ray, Best<T> &best, Counters &c) {
    for (size_t i = node.first; i < node.first + node.count; ++i) {
        Intersection isect;
        ++c.prim_tests;
        if (!prim_intersection(bvh.primitives[i], ray, isect)) continue;
        update_best(isect, i, best);
    }
    F lt = 0, rt = 0;
    bool lsome = false, rsome = false;
    if (node.left_child != NONE) { ++c.aabb_tests; lsome = aabb_intersects(bvh.nodes[node.left_child].aabb, ray, lt); }
    if (node.right_child != NONE) { ++c.aabb_tests; rsome = aabb_intersects(bvh.nodes[node.right_child].aabb, ray, rt); }
    F b = best.t_or_inf();
    F left = lsome ? (lt < b ? lt : b) : b;
    F right = rsome ? (rt < b ? rt : b) : b;
    if (left < b) {
        if (right < b) {
            if (left < right) {
                node_intersection(bvh, bvh.nodes[node.left_child], ray, best, c);
                F b2 = best.t_or_inf();
                if (right < b2) node_intersection(bvh, bvh.nodes[node.right_child], ray, best, c);
            } else {
                node_intersection(bvh, bvh.nodes[node.right_child], ray, best, c);
                F b2 = best.t_or_inf();
                if (left < b2) node_intersection(bvh, bvh.nodes[node.left_child], ray, best, c);
            }
        } else {
            node_intersection(bvh, bvh.nodes[node.left_child], ray, best, c);
        }
    } else if (right < b) {
        node_intersection(bvh, bvh.nodes[node.right_child], ray, best, c);
    }
}
template <class T>
static Best<T> bvh_intersection(const BVH<T> &bvh, const Ray &ray, Counters &c) {  // bvh.rs:27-36
    Best<T> best;
    if (!bvh.primitives.empty()) {
        F t;
        ++c.aabb_tests;
        if (aabb_intersects(bvh.nodes[0].aabb, ray, t)) node_intersection(bvh, bvh.nodes[0], ray, best, c);
    }
    return best;
}
// Node::intersections, bvh.rs:188-210
template <class T, class CB>
static void node_intersections(const BVH<T> &bvh, const Node &node, const Ray &ray, CB &cb, Counters &c) {
    for (size_t i = node.first; i < node.first + node.count; ++i) {
        Intersection hits[2];
        ++c.light_prim_tests;
        int k = prim_all_intersections(bvh.primitives[i], ray, hits);
        if (k >= 1) cb(hits[0], bvh.primitives[i]);
        if (k == 2) cb(hits[1], bvh.primitives[i]);
    }
    F t;
    if (node.left_child != NONE) {
        ++c.light_aabb_tests;
        if (aabb_intersects(bvh.nodes[node.left_child].aabb, ray, t)) node_intersections(bvh, bvh.nodes[node.left_child], ray, cb, c);
    }
    if (node.right_child != NONE) {
        ++c.light_aabb_tests;
        if (aabb_intersects(bvh.nodes[node.right_child].aabb, ray, t)) node_intersections(bvh, bvh.nodes[node.right_child], ray, cb, c);
    }
}
template <class T, class CB>
static void bvh_intersections(const BVH<T> &bvh, const Ray &ray, CB &cb, Counters &c) {  // bvh.rs:38-45
    if (!bvh.primitives.empty()) {
        F t;
        ++c.light_aabb_tests;
        if (aabb_intersects(bvh.nodes[0].aabb, ray, t)) node_intersections(bvh, bvh.nodes[0], ray, cb, c);
    }
}

// ------------------------------------------------------------------------------------------------
// scene.rs
struct ScenePrims {                                          // scene.rs:57-62
    std::vector<Shape> planes;
    BVH<Shape> ellipsoids, boxes;
    BVH<Tri> triangles;
};
struct LightPrims {                                          // scene.rs:65-79
    BVH<Shape> ellipsoids, boxes;
    BVH<Tri> triangles;
    size_t len() const { return ellipsoids.primitives.size() + boxes.primitives.size() + triangles.primitives.size(); }
    bool is_empty() const { return len() == 0; }
};
struct Scene {                                               // scene.rs:82-90
    ScenePrims primitives;
    LightPrims lights;
    uint32_t ray_depth, samples, width, height;
    V3 bg_color;
    V3 cam_position, cam_right, cam_up, cam_forward;
    int fov_kind;
    F fov;
    std::vector<Material> materials;
};
static inline bool is_light(const Material &m) { return !is_zero(m.emission); }  // scene.rs:225-227

static AABB rotated_aabb(const AABB &b, Quat r) {            // scene.rs:255-268
    AABB o = aabb_empty();
    extend(o, rotate(r, v3(b.min.x, b.min.y, b.min.z)));
    extend(o, rotate(r, v3(b.min.x, b.min.y, b.max.z)));
    extend(o, rotate(r, v3(b.min.x, b.max.y, b.min.z)));
    extend(o, rotate(r, v3(b.min.x, b.max.y, b.max.z)));
    extend(o, rotate(r, v3(b.max.x, b.min.y, b.min.z)));
    extend(o, rotate(r, v3(b.max.x, b.min.y, b.max.z)));
    extend(o, rotate(r, v3(b.max.x, b.max.y, b.min.z)));
    extend(o, rotate(r, v3(b.max.x, b.max.y, b.max.z)));
    return o;
}
static Shape make_shape(const rt_shape &s, uint32_t orig) {  // Primitive::new / new_without_aabb, scene.rs:108-137
    Shape o;
    o.type = s.type;
    o.param = v3(s.param[0], s.param[1], s.param[2]);
    o.position = v3(s.position[0], s.position[1], s.position[2]);
    o.rotation = quat_xyzw(s.rotation[0], s.rotation[1], s.rotation[2], s.rotation[3]);
    o.material = s.material;
    o.orig = orig;
    if (s.type == RT_SHAPE_PLANE) {
        o.aabb = aabb_empty();
    } else {
        AABB local = aabb_empty();                           // Box::new / Ellipsoid::new, box.rs:12-17
        extend(local, o.param);
        extend(local, -o.param);
        o.aabb = rotated_aabb(local, o.rotation);
        o.aabb.min = o.aabb.min + o.position;
        o.aabb.max = o.aabb.max + o.position;
    }
    return o;
}
static Tri make_triangle(const rt_raw_scene &raw, uint64_t i) {
    const double *p = raw.tri_pos + i * 9;
    V3 a = v3(p[0], p[1], p[2]), b = v3(p[3], p[4], p[5]), c = v3(p[6], p[7], p[8]);
    Tri t;
    tri_props(a, b, c, t);
    if (raw.tri_nrm) {
        const double *n = raw.tri_nrm + i * 9;
        t.na = v3(n[0], n[1], n[2]); t.nb = v3(n[3], n[4], n[5]); t.nc = v3(n[6], n[7], n[8]);
    } else {
        t.na = t.nb = t.nc = t.geometry_normal;              // new_with_geometry_normals, triangle.rs:25-28
    }
    t.material = raw.tri_mat ? raw.tri_mat[i] : 0;
    t.orig = (uint32_t)i;
    if (raw.tri_mode == RT_TRI_BAKED) {                      // instantiate, gltf/scene_builder.rs:42-55
        t.aabb = aabb_empty();
        extend(t.aabb, t.a);
        extend(t.aabb, t.a + t.ba);
        extend(t.aabb, t.a + t.ca);
        return t;
    }
    // TrianglePrimitive::new, scene.rs:139-165
    V3 position = v3(0, 0, 0);
    Quat rotation{1.0, v3(0, 0, 0)};
    if (raw.tri_xform) {
        const double *x = raw.tri_xform + i * 7;
        position = v3(x[0], x[1], x[2]);
        rotation = quat_xyzw(x[3], x[4], x[5], x[6]);
    }
    V3 wa = rotate(rotation, t.a) + position;
    V3 wb = rotate(rotation, t.ba + t.a) + position;
    V3 wc = rotate(rotation, t.ca + t.a) + position;
    V3 na = rotate(rotation, t.na), nb = rotate(rotation, t.nb), nc = rotate(rotation, t.nc);
    Tri w;
    tri_props(wa, wb, wc, w);
    w.na = na; w.nb = nb; w.nc = nc;
    w.material = t.material;
    w.orig = t.orig;
    w.aabb = aabb_empty();
    extend(w.aabb, wa);
    extend(w.aabb, wb);
    extend(w.aabb, wc);
    return w;
}
static Scene *scene_build(const rt_raw_scene &raw) {         // Scene::new + make_scenes / build_scene
    Scene *s = new Scene();
    s->ray_depth = raw.ray_depth; s->samples = raw.samples; s->width = raw.width; s->height = raw.height;
    s->bg_color = v3(raw.bg_color[0], raw.bg_color[1], raw.bg_color[2]);
    s->cam_position = v3(raw.cam_position[0], raw.cam_position[1], raw.cam_position[2]);
    s->cam_right = v3(raw.cam_right[0], raw.cam_right[1], raw.cam_right[2]);
    s->cam_up = v3(raw.cam_up[0], raw.cam_up[1], raw.cam_up[2]);
    s->cam_forward = v3(raw.cam_forward[0], raw.cam_forward[1], raw.cam_forward[2]);
    s->fov_kind = raw.fov_kind; s->fov = raw.fov;
    for (uint32_t i = 0; i < raw.n_materials; ++i) {
        const rt_material &m = raw.materials[i];
        s->materials.push_back({m.kind, m.ior, v3(m.color[0], m.color[1], m.color[2]), v3(m.emission[0], m.emission[1], m.emission[2])});
    }
    std::vector<Shape> planes, ellipsoids, boxes;
    std::vector<Tri> triangles;
    uint32_t nb = 0, ne = 0, np = 0;
    for (uint32_t i = 0; i < raw.n_shapes; ++i) {            // scene.rs:200-207
        const rt_shape &sh = raw.shapes[i];
        if (sh.type == RT_SHAPE_BOX) boxes.push_back(make_shape(sh, nb++));
        else if (sh.type == RT_SHAPE_ELLIPSOID) ellipsoids.push_back(make_shape(sh, ne++));
        else planes.push_back(make_shape(sh, np++));
    }
    triangles.reserve(raw.n_triangles);
    for (uint64_t i = 0; i < raw.n_triangles; ++i) triangles.push_back(make_triangle(raw, i));
    auto lights_of = [&](auto &v) {
        std::remove_reference_t<decltype(v)> out;
        for (auto &p : v) if (is_light(s->materials[p.material])) out.push_back(p);
        return out;
    };
    // glTF scenes use BVH::new_empty() for the shape hierarchies (gltf/scene_builder.rs:14-15):
    // no nodes at all, whereas BVH::new(vec![]) pushes one empty leaf (bvh.rs:77-80).
    const bool gltf = raw.tri_mode == RT_TRI_BAKED && raw.n_shapes == 0;
    if (!gltf) {
        s->lights.boxes = bvh_new(lights_of(boxes));         // scene.rs:209-213 (lights first)
        s->lights.ellipsoids = bvh_new(lights_of(ellipsoids));
    }
    s->lights.triangles = bvh_new(lights_of(triangles));
    s->primitives.planes = std::move(planes);                // scene.rs:215-220
    if (!gltf) {
        s->primitives.ellipsoids = bvh_new(std::move(ellipsoids));
        s->primitives.boxes = bvh_new(std::move(boxes));
    }
    s->primitives.triangles = bvh_new(std::move(triangles));
    return s;
}

// ------------------------------------------------------------------------------------------------
// intersections.rs:42-91
struct SceneHit {
    bool some = false;
    Intersection isect{};
    int kind = -1;
    size_t prim = 0;
    uint32_t material = 0;
    Quat rotation{1.0, {0, 0, 0}};
};
static inline void update_best_with(const Intersection &i, int kind, size_t prim, uint32_t material, Quat rot, SceneHit &res) {  // :76-85
    if (res.some) {
        if (!(i.t < res.isect.t)) return;
    }
    res.some = true; res.isect = i; res.kind = kind; res.prim = prim; res.material = material; res.rotation = rot;
}
static SceneHit intersect(const Ray &ray, const ScenePrims &scene, F max_dist, Counters &c) {
    SceneHit res;
    for (size_t i = 0; i < scene.planes.size(); ++i) {       // :45-49
        Intersection isect;
        ++c.prim_tests;
        if (!shape_intersection(scene.planes[i], ray, isect)) continue;
        update_best_with(isect, 0, i, scene.planes[i].material, scene.planes[i].rotation, res);
    }
    auto b = bvh_intersection(scene.boxes, ray, c);          // :51
    if (b.some) update_best_with(b.isect, 1, b.prim, scene.boxes.primitives[b.prim].material, scene.boxes.primitives[b.prim].rotation, res);
    auto e = bvh_intersection(scene.ellipsoids, ray, c);     // :52
    if (e.some) update_best_with(e.isect, 2, e.prim, scene.ellipsoids.primitives[e.prim].material, scene.ellipsoids.primitives[e.prim].rotation, res);
    auto t = bvh_intersection(scene.triangles, ray, c);      // :53, DONT_ROTATE :64
    if (t.some) update_best_with(t.isect, 3, t.prim, scene.triangles.primitives[t.prim].material, Quat{1.0, {0, 0, 0}}, res);
    if (res.some) {                                          // :55-61
        if (res.isect.t * magnitude(ray.dir) <= max_dist) {
            res.isect = with_rotated_normal(res.isect, res.rotation);
        } else {
            res.some = false; res.kind = -1;
        }
    }
    return res;
}

// intersection_probability.rs
static inline F prob_triangle(const Tri &t) { return t.inverted_area; }               // :9-13
static inline F prob_box(V3 s) { return 1.0 / sum(v3(s.y * s.z, s.x * s.z, s.x * s.y)) / 8.0; }  // :15-23
static inline F prob_ellipsoid(V3 r, const Intersection &i) {                          // :25-35
    V3 coef = mul_ew(v3(r.y * r.z, r.x * r.z, r.x * r.y), i.geometry_normal);
    return 1.0 / (4.0 * PI * std::sqrt(dot(coef, coef)));
}
static inline F to_direction_probability(const Intersection &i, const Ray &ray) {      // ray_sampler.rs:172-174
    return i.t * i.t / std::fabs(dot(ray.dir, i.geometry_normal));
}

// ------------------------------------------------------------------------------------------------
// ray_sampler.rs
static inline V3 vec3_standard_rng(Rng &rng) {               // :163-170
    F a0 = rng.gen_f64(), a1 = rng.gen_f64(), a2 = rng.gen_f64();
    return v3(a0 * 2.0 - 1.0, a1 * 2.0 - 1.0, a2 * 2.0 - 1.0);
}
static inline V3 uniform_on_sphere(Rng &rng) { return normalize(vec3_standard_rng(rng)); }  // :159-161
static inline V3 uniform_on_box(V3 sizes, Rng &rng) {        // :142-157
    F w4x = sizes.y * sizes.z, w4y = sizes.x * sizes.z, w4z = sizes.x * sizes.y;
    F choice = rng.range(0.0, w4x + w4y + w4z);
    F sign = (F)(rng.range_i32_0_1() * 2 - 1);
    F u1 = rng.range_inclusive(-1.0, 1.0);
    F u2 = rng.range_inclusive(-1.0, 1.0);
    V3 p;
    if (choice < w4x) p = v3(sign, u1, u2);
    else if (choice < w4x + w4y) p = v3(u1, sign, u2);
    else p = v3(u1, u2, sign);
    return mul_ew(p, sizes);
}
static inline V3 cosine_sample(V3 normal, Rng &rng) {        // :68-76
    V3 v = uniform_on_sphere(rng);
    V3 dir = v + normal;
    const F eps = EPSILON * 16.0;
    if (std::fabs(dir.x) <= eps && std::fabs(dir.y) <= eps && std::fabs(dir.z) <= eps) return normal;
    return normalize(dir);
}
static inline F cosine_pdf(V3 normal, V3 dir) {              // :78-83
    if (dot(normal, dir) <= 0.0) return 0.0;
    return dot(normal, dir) / PI;
}
static V3 light_sample(V3 pos, const LightPrims &lights, Rng &rng) {  // :101-130
    size_t index = (size_t)rng.range_usize(lights.len());
    const auto &boxes = lights.boxes.primitives;
    const auto &ellipsoids = lights.ellipsoids.primitives;
    const auto &triangles = lights.triangles.primitives;
    V3 world_pos;
    if (index < boxes.size()) {
        const Shape &l = boxes[index];
        V3 local = uniform_on_box(l.param, rng);
        world_pos = rotate(l.rotation, local) + l.position;
    } else if (index < boxes.size() + ellipsoids.size()) {
        const Shape &l = ellipsoids[index - boxes.size()];
        V3 local = mul_ew(uniform_on_sphere(rng), l.param);
        world_pos = rotate(l.rotation, local) + l.position;
    } else {
        const Tri &t = triangles[index - boxes.size() - ellipsoids.size()];
        F u = rng.range_inclusive(0.0, 1.0);
        F v = rng.range_inclusive(0.0, 1.0);
        if (u + v > 1.0) { u = 1.0 - u; v = 1.0 - v; }
        world_pos = t.ba * u + t.ca * v + t.a;
    }
    return normalize(world_pos - pos);
}
static F light_pdf(V3 pos, V3 dir, const LightPrims &lights, Counters &c) {  // :132-139
    Ray ray{pos + EPSILON * dir, dir};
    F impact = 0.0;
    ++c.light_pdf_rays;
    auto box_cb = [&](const Intersection &i, const Shape &p) {                // intersections.rs:88
        Intersection r = with_rotated_normal(i, p.rotation);
        impact += prob_box(p.param) * to_direction_probability(r, ray);
    };
    auto ell_cb = [&](const Intersection &i, const Shape &p) {                // :89
        Intersection r = with_rotated_normal(i, p.rotation);
        impact += prob_ellipsoid(p.param, r) * to_direction_probability(r, ray);
    };
    auto tri_cb = [&](const Intersection &i, const Tri &p) {                  // :90
        impact += prob_triangle(p) * to_direction_probability(i, ray);
    };
    bvh_intersections(lights.boxes, ray, box_cb, c);
    bvh_intersections(lights.ellipsoids, ray, ell_cb, c);
    bvh_intersections(lights.triangles, ray, tri_cb, c);
    return impact / (F)lights.len();
}

// ------------------------------------------------------------------------------------------------
// raytrace.rs
static inline F powi(F a, int b) {                           // compiler-rt __powidf2 (f64::powi)
    F r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}
static inline F reflection_power(F n1, F n2, const Ray &ray, const Intersection &i) {  // :62-65
    F r0 = powi((n1 - n2) / (n1 + n2), 2);
    return r0 + (1.0 - r0) * powi(1.0 + dot(ray.dir, i.shading_normal), 5);
}
static inline Ray reflected_ray(const Ray &ray, const Intersection &i) {               // :67-73
    V3 dir = ray.dir - 2.0 * i.shading_normal * dot(i.shading_normal, ray.dir);
    return {ray.position_at(i.t) + EPSILON * dir, dir};
}
static inline bool refracted_ray(const Ray &ray, const Intersection &i, F eta, Ray &out) {  // :75-88
    F cos_theta1 = -dot(i.shading_normal, ray.dir);
    F sin_theta2 = eta * std::sqrt(1.0 - cos_theta1 * cos_theta1);
    if (sin_theta2 > 1.0) return false;
    F cos_theta2 = std::sqrt(1.0 - sin_theta2 * sin_theta2);
    V3 dir = eta * ray.dir + (eta * cos_theta1 - cos_theta2) * i.shading_normal;
    out = {ray.position_at(i.t) + EPSILON * dir, dir};
    return true;
}
static V3 raytrace_impl(const Ray &ray, const Scene &scene, Rng &rng, uint32_t left, Counters &c) {  // :12-60
    if (left == 0) return v3(0, 0, 0);
    ++c.segments;
    SceneHit hit = intersect(ray, scene.primitives, INF, c);
    if (!hit.some) return scene.bg_color;
    const Intersection &isect = hit.isect;
    const Material &m = scene.materials[hit.material];
    V3 term;
    if (m.kind == RT_MAT_DIFFUSE) {                          // :16-34
        V3 pos = ray.position_at(isect.t);
        bool use_mix = !scene.lights.is_empty();
        V3 dir;
        if (use_mix) {                                       // Mix::sample :87-93
            if (rng.gen_bool(0.5)) dir = cosine_sample(isect.shading_normal, rng);
            else dir = light_sample(pos, scene.lights, rng);
        } else {
            dir = cosine_sample(isect.shading_normal, rng);
        }
        if (dot(dir, isect.shading_normal) <= 0.0) { term = v3(0, 0, 0); goto done; }
        {
            F pdf = use_mix ? (cosine_pdf(isect.shading_normal, dir) + light_pdf(pos, dir, scene.lights, c)) / 2.0   // :95-97
                            : cosine_pdf(isect.shading_normal, dir);
            if (pdf == 0.0) { term = v3(0, 0, 0); goto done; }
            V3 l = raytrace_impl(Ray{pos + EPSILON * dir, dir}, scene, rng, left - 1, c);
            term = dot(dir, isect.shading_normal) * mul_ew(m.color, l) / PI / pdf;   // :33
        }
    } else if (m.kind == RT_MAT_DIELECTRIC) {                // :36-54
        F n1 = 1.0, n2 = m.ior;
        if (isect.inside) std::swap(n1, n2);
        Ray refl = reflected_ray(ray, isect);
        Ray refr;
        if (!refracted_ray(ray, isect, n1 / n2, refr)) {
            term = raytrace_impl(refl, scene, rng, left - 1, c);
        } else {
            F rp = reflection_power(n1, n2, ray, isect);
            F p = rp < 0.0 ? 0.0 : (rp > 1.0 ? 1.0 : rp);    // f64::clamp(0,1) (NaN stays NaN)
            if (rng.gen_bool(p)) {
                term = raytrace_impl(refl, scene, rng, left - 1, c);
            } else {
                V3 col = raytrace_impl(refr, scene, rng, left - 1, c);
                term = isect.inside ? col : mul_ew(col, m.color);
            }
        }
    } else {                                                 // Metallic :56-58
        term = mul_ew(raytrace_impl(reflected_ray(ray, isect), scene, rng, left - 1, c), m.color);
    }
done:
    return m.emission + term;
}

// camera.rs
struct Camera {
    V3 position, right, up, forward;
    F tan_half_fov_x, tan_half_fov_y, width, height;
};
static Camera camera_new(const Scene &s) {                   // camera.rs:18-46
    Camera c;
    c.width = (F)s.width; c.height = (F)s.height;
    if (s.fov_kind == RT_FOV_Y) {
        c.tan_half_fov_y = std::tan(s.fov / 2.0);
        F aspect = c.height / c.width;
        c.tan_half_fov_x = c.tan_half_fov_y / aspect;
    } else {
        c.tan_half_fov_x = std::tan(s.fov / 2.0);
        F aspect = c.width / c.height;
        c.tan_half_fov_y = c.tan_half_fov_x / aspect;
    }
    c.position = s.cam_position; c.right = s.cam_right; c.up = s.cam_up; c.forward = s.cam_forward;
    return c;
}
static inline Ray camera_ray(const Camera &c, uint32_t px_x, uint32_t px_y, F jx, F jy) {  // fuzzy_ray :48-55
    F px = (F)px_x + jx, py = (F)px_y + jy;
    F x = (2.0 * px / c.width - 1.0) * c.tan_half_fov_x;
    F y = -(2.0 * py / c.height - 1.0) * c.tan_half_fov_y;
    V3 dir = x * c.right + y * c.up + 1.0 * c.forward;
    return {c.position, dir};
}

// postprocessing.rs
static inline F clamp01(F v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }   // num_traits::clamp
static inline F aces(F x) {                                  // :9-20
    const F a = 2.51, b = 0.03, c = 2.43, d = 0.59, e = 0.14;
    return clamp01(((a * x + b) * x) / (((c * x + d) * x) + e));
}
static inline F gamma_c(F x) { return std::pow(x, 1.0 / 2.2); }  // :5-7

}  // namespace orc

// =================================================================================================
// C API (ctypes) — names prefixed orc_ so nothing can be confused with the product's rt_ symbols.
using namespace orc;
#define ORC_API extern "C" __attribute__((visibility("default")))

ORC_API void *orc_scene_build(const rt_raw_scene *raw) { return scene_build(*raw); }
ORC_API void orc_scene_free(void *s) { delete (Scene *)s; }

template <class T>
static void bvh_counts(const BVH<T> &b, uint64_t *nodes, uint64_t *prims) { *nodes = b.nodes.size(); *prims = b.primitives.size(); }
template <class T>
static void bvh_export(const BVH<T> &b, double *bounds, uint32_t *left, uint32_t *right, uint32_t *first, uint32_t *count, uint32_t *order) {
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const Node &n = b.nodes[i];
        if (bounds) { double *o = bounds + i * 6; o[0] = n.aabb.min.x; o[1] = n.aabb.min.y; o[2] = n.aabb.min.z; o[3] = n.aabb.max.x; o[4] = n.aabb.max.y; o[5] = n.aabb.max.z; }
        if (left) left[i] = n.left_child == NONE ? RT_NO_CHILD : (uint32_t)n.left_child;
        if (right) right[i] = n.right_child == NONE ? RT_NO_CHILD : (uint32_t)n.right_child;
        if (first) first[i] = (uint32_t)n.first;
        if (count) count[i] = (uint32_t)n.count;
    }
    if (order) for (size_t i = 0; i < b.primitives.size(); ++i) order[i] = b.primitives[i].orig;
}
#define WITH_BVH(s, which, CALL)                                             \
    switch (which) {                                                         \
        case RT_BVH_BOXES: CALL((s)->primitives.boxes); break;               \
        case RT_BVH_ELLIPSOIDS: CALL((s)->primitives.ellipsoids); break;     \
        case RT_BVH_TRIANGLES: CALL((s)->primitives.triangles); break;       \
        case RT_BVH_LIGHT_BOXES: CALL((s)->lights.boxes); break;             \
        case RT_BVH_LIGHT_ELLIPSOIDS: CALL((s)->lights.ellipsoids); break;   \
        case RT_BVH_LIGHT_TRIANGLES: CALL((s)->lights.triangles); break;     \
        default: return 1;                                                   \
    }
ORC_API int orc_bvh_counts(const void *sv, int which, uint64_t *nodes, uint64_t *prims) {
    const Scene *s = (const Scene *)sv;
#define CALL(B) bvh_counts(B, nodes, prims)
    WITH_BVH(s, which, CALL)
#undef CALL
    return 0;
}
ORC_API int orc_bvh_get(const void *sv, int which, double *bounds, uint32_t *left, uint32_t *right, uint32_t *first, uint32_t *count, uint32_t *order) {
    const Scene *s = (const Scene *)sv;
#define CALL(B) bvh_export(B, bounds, left, right, first, count, order)
    WITH_BVH(s, which, CALL)
#undef CALL
    return 0;
}
ORC_API int orc_get_triangle(const void *sv, int which, uint64_t i, double *o, uint32_t *material) {
    const Scene *s = (const Scene *)sv;
    const BVH<Tri> &b = which == RT_BVH_TRIANGLES ? s->primitives.triangles : s->lights.triangles;
    if (i >= b.primitives.size()) return 1;
    const Tri &t = b.primitives[i];
    const V3 vs[7] = {t.a, t.ba, t.ca, t.geometry_normal, t.na, t.nb, t.nc};
    for (int k = 0; k < 7; ++k) { o[k * 3] = vs[k].x; o[k * 3 + 1] = vs[k].y; o[k * 3 + 2] = vs[k].z; }
    o[21] = t.inverted_area;
    o[22] = t.aabb.min.x; o[23] = t.aabb.min.y; o[24] = t.aabb.min.z;
    o[25] = t.aabb.max.x; o[26] = t.aabb.max.y; o[27] = t.aabb.max.z;
    if (material) *material = t.material;
    return 0;
}
ORC_API int orc_get_shape(const void *sv, int which, uint64_t i, rt_shape *out, double *aabb6) {
    const Scene *s = (const Scene *)sv;
    const std::vector<Shape> *v = nullptr;
    switch (which) {
        case -1: v = &s->primitives.planes; break;
        case RT_BVH_BOXES: v = &s->primitives.boxes.primitives; break;
        case RT_BVH_ELLIPSOIDS: v = &s->primitives.ellipsoids.primitives; break;
        case RT_BVH_LIGHT_BOXES: v = &s->lights.boxes.primitives; break;
        case RT_BVH_LIGHT_ELLIPSOIDS: v = &s->lights.ellipsoids.primitives; break;
        default: return 1;
    }
    if (i >= v->size()) return 1;
    const Shape &p = (*v)[i];
    out->type = p.type; out->material = p.material;
    out->param[0] = p.param.x; out->param[1] = p.param.y; out->param[2] = p.param.z;
    out->position[0] = p.position.x; out->position[1] = p.position.y; out->position[2] = p.position.z;
    out->rotation[0] = p.rotation.v.x; out->rotation[1] = p.rotation.v.y; out->rotation[2] = p.rotation.v.z; out->rotation[3] = p.rotation.s;
    if (aabb6) { aabb6[0] = p.aabb.min.x; aabb6[1] = p.aabb.min.y; aabb6[2] = p.aabb.min.z; aabb6[3] = p.aabb.max.x; aabb6[4] = p.aabb.max.y; aabb6[5] = p.aabb.max.z; }
    return 0;
}
ORC_API void orc_camera(const void *sv, double *tan_half_fov_xy) {
    Camera c = camera_new(*(const Scene *)sv);
    tan_half_fov_xy[0] = c.tan_half_fov_x;
    tan_half_fov_xy[1] = c.tan_half_fov_y;
}

static void fill_hit(const SceneHit &h, rt_hit &o) {
    if (!h.some) {
        o = rt_hit{};
        o.kind = -1;
        return;
    }
    o.t = h.isect.t;
    o.geometry_normal[0] = h.isect.geometry_normal.x; o.geometry_normal[1] = h.isect.geometry_normal.y; o.geometry_normal[2] = h.isect.geometry_normal.z;
    o.shading_normal[0] = h.isect.shading_normal.x; o.shading_normal[1] = h.isect.shading_normal.y; o.shading_normal[2] = h.isect.shading_normal.z;
    o.kind = h.kind; o.prim = (uint32_t)h.prim; o.inside = h.isect.inside ? 1 : 0; o.material = h.material;
}
template <class Fn>
static void parallel_for(uint64_t n, int nthreads, Fn fn) {   // rayon par_iter stand-in (main.rs:94)
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint64_t> next{0};
    const uint64_t chunk = 256;
    auto worker = [&](int tid) {
        for (;;) {
            uint64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            uint64_t e = std::min(n, b + chunk);
            fn(b, e, tid);
        }
    };
    if (nthreads == 1) { worker(0); return; }
    std::vector<std::thread> ts;
    for (int t = 0; t < nthreads; ++t) ts.emplace_back(worker, t);
    for (auto &t : ts) t.join();
}
static void export_counters(const Counters &c, rt_counters *out) {
    if (!out) return;
    *out = rt_counters{};
    out->segments = c.segments; out->light_pdf_rays = c.light_pdf_rays; out->paths = c.paths;
    out->aabb_tests = c.aabb_tests; out->prim_tests = c.prim_tests;
    out->light_aabb_tests = c.light_aabb_tests; out->light_prim_tests = c.light_prim_tests;
}

// intersect() for caller-supplied rays
ORC_API int orc_trace_rays(const void *sv, uint64_t n, const double *origins, const double *dirs, rt_hit *hits, int nthreads, rt_counters *counters) {
    const Scene *s = (const Scene *)sv;
    int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
    std::vector<Counters> cs(nt > 0 ? nt : 1);
    parallel_for(n, nt, [&](uint64_t b, uint64_t e, int tid) {
        for (uint64_t i = b; i < e; ++i) {
            Ray r{v3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]), v3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2])};
            ++cs[tid].segments;
            fill_hit(intersect(r, s->primitives, INF, cs[tid]), hits[i]);
        }
    });
    Counters tot;
    for (auto &c : cs) tot.add(c);
    export_counters(tot, counters);
    return 0;
}
// fuzzy_ray with fixed jitter + normalise + intersect, one ray per pixel
ORC_API int orc_trace_primary(const void *sv, double jx, double jy, rt_hit *hits, double *rays_out, int nthreads, rt_counters *counters) {
    const Scene *s = (const Scene *)sv;
    Camera cam = camera_new(*s);
    int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
    std::vector<Counters> cs(nt > 0 ? nt : 1);
    uint64_t n = (uint64_t)s->width * s->height;
    parallel_for(n, nt, [&](uint64_t b, uint64_t e, int tid) {
        for (uint64_t i = b; i < e; ++i) {
            Ray r = camera_ray(cam, (uint32_t)(i % s->width), (uint32_t)(i / s->width), jx, jy);
            r.dir = normalize(r.dir);                         // raytrace.rs:9
            if (rays_out) {
                double *o = rays_out + i * 6;
                o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z; o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z;
            }
            ++cs[tid].segments;
            if (hits) fill_hit(intersect(r, s->primitives, INF, cs[tid]), hits[i]);
        }
    });
    Counters tot;
    for (auto &c : cs) tot.add(c);
    export_counters(tot, counters);
    return 0;
}
ORC_API int orc_light_pdf(const void *sv, uint64_t n, const double *pos, const double *dir, double *pdf, rt_counters *counters) {
    const Scene *s = (const Scene *)sv;
    Counters c;
    for (uint64_t i = 0; i < n; ++i)
        pdf[i] = s->lights.is_empty() ? 0.0
                                      : light_pdf(v3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), v3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]), s->lights, c);
    export_counters(c, counters);
    return 0;
}
// Light::sample with the stream (seed, pixel=stream, sample=0): for sampler distribution tests
ORC_API int orc_light_sample(const void *sv, uint64_t seed, uint64_t n, const double *pos, double *dir_out) {
    const Scene *s = (const Scene *)sv;
    if (s->lights.is_empty()) return 1;
    for (uint64_t i = 0; i < n; ++i) {
        Rng rng(seed, (uint32_t)i, 0);
        V3 d = light_sample(v3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]), s->lights, rng);
        dir_out[i * 3] = d.x; dir_out[i * 3 + 1] = d.y; dir_out[i * 3 + 2] = d.z;
    }
    return 0;
}
ORC_API int orc_cosine_sample(uint64_t seed, uint64_t n, const double *normal, double *dir_out) {
    for (uint64_t i = 0; i < n; ++i) {
        Rng rng(seed, (uint32_t)i, 0);
        V3 d = cosine_sample(v3(normal[0], normal[1], normal[2]), rng);
        dir_out[i * 3] = d.x; dir_out[i * 3 + 1] = d.y; dir_out[i * 3 + 2] = d.z;
    }
    return 0;
}
// The pixel loop of generate_image (main.rs:94-103): accum[p] += sum over samples
// [sample_begin, sample_begin+sample_count) of raytrace(fuzzy_ray).  accum is W*H*3, NOT divided.
ORC_API int orc_render_accum(const void *sv, uint32_t sample_begin, uint32_t sample_count, uint64_t seed, int nthreads, double *accum, rt_counters *counters) {
    const Scene *s = (const Scene *)sv;
    Camera cam = camera_new(*s);
    int nt = nthreads <= 0 ? (int)std::thread::hardware_concurrency() : nthreads;
    if (nt < 1) nt = 1;
    std::vector<Counters> cs(nt);
    uint64_t n = (uint64_t)s->width * s->height;
    parallel_for(n, nt, [&](uint64_t b, uint64_t e, int tid) {
        for (uint64_t i = b; i < e; ++i) {
            V3 result = v3(accum[i * 3], accum[i * 3 + 1], accum[i * 3 + 2]);
            for (uint32_t k = 0; k < sample_count; ++k) {
                Rng rng(seed, (uint32_t)i, sample_begin + k);
                F jx = rng.range(0.0, 1.0), jy = rng.range(0.0, 1.0);     // camera.rs:49-50
                Ray r = camera_ray(cam, (uint32_t)(i % s->width), (uint32_t)(i / s->width), jx, jy);
                r.dir = normalize(r.dir);
                ++cs[tid].paths;
                result = result + raytrace_impl(r, *s, rng, s->ray_depth, cs[tid]);
            }
            accum[i * 3] = result.x; accum[i * 3 + 1] = result.y; accum[i * 3 + 2] = result.z;
        }
    });
    Counters tot;
    for (auto &c : cs) tot.add(c);
    export_counters(tot, counters);
    return 0;
}
// correct_gamma(aces_tonemap(sum / samples)), main.rs:104
ORC_API void orc_postprocess(uint64_t n_pixels, const double *accum, uint32_t total_samples, double *rgb) {
    for (uint64_t i = 0; i < n_pixels * 3; ++i) rgb[i] = gamma_c(aces(accum[i] / (F)total_samples));
}
// ppm.rs:13-19
ORC_API void orc_to_bytes(uint64_t n, const double *rgb, uint8_t *bytes) {
    for (uint64_t i = 0; i < n; ++i) {
        F v = rgb[i];
        v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
        F r = std::round(v * 255.0);
        bytes[i] = (uint8_t)(r != r ? 0 : r);
    }
}

// ---- known-answer entry points for the reference's 13 inline tests ------------------------------
ORC_API int orc_kat_aabb(const double *mn, const double *mx, const double *o, const double *d, double *t) {
    AABB b{v3(mn[0], mn[1], mn[2]), v3(mx[0], mx[1], mx[2])};
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2])};
    return aabb_intersects(b, r, *t) ? 1 : 0;
}
ORC_API int orc_kat_box(const double *sizes, const double *o, const double *d, double *t, double *normal, int *inside) {
    Shape s{};
    s.type = RT_SHAPE_BOX; s.param = v3(sizes[0], sizes[1], sizes[2]); s.position = v3(0, 0, 0); s.rotation = Quat{1.0, v3(0, 0, 0)};
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2])};
    BoxPlane n, f;                                           // Box::intersection on the raw ray (no Primitive wrapper)
    int k = intersect_box_coef(s.param, r, n, f);
    if (k == 0) return 0;
    const BoxPlane &h = k == 2 ? n : f;
    *t = h.t;
    V3 nn = h.real_normal();
    normal[0] = nn.x; normal[1] = nn.y; normal[2] = nn.z;
    *inside = k == 1;
    return 1;
}
ORC_API int orc_kat_triangle(const double *a, const double *b, const double *c, const double *o, const double *d, double *t_u_v) {
    Tri t;
    tri_props(v3(a[0], a[1], a[2]), v3(b[0], b[1], b[2]), v3(c[0], c[1], c[2]), t);
    t.na = t.nb = t.nc = t.geometry_normal;
    Ray r{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2])};
    Intersection i;
    if (!tri_intersection(t, r, i)) return 0;
    t_u_v[0] = i.t;
    return 1;
}
ORC_API void orc_kat_normalize(const double *v, double *out) {
    V3 n = normalize(v3(v[0], v[1], v[2]));
    out[0] = n.x; out[1] = n.y; out[2] = n.z;
}
// cof, gltf/scene_builder.rs:367-388.  m: column-major 9, out: column-major 9
ORC_API void orc_kat_cof(const double *m, double *out) {
    auto M = [&](int col, int row) { return m[col * 3 + row]; };
    auto others = [](int i, int &a, int &b) { if (i == 0) { a = 1; b = 2; } else if (i == 1) { a = 0; b = 2; } else { a = 0; b = 1; } };
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row) {
            int lc, rc, tr, br;
            others(col, lc, rc);
            others(row, tr, br);
            // Mat2::new(c0r0, c0r1, c1r0, c1r1).determinant() = c0r0*c1r1 - c1r0*c0r1
            F d = M(lc, tr) * M(rc, br) - M(rc, tr) * M(lc, br);
            out[col * 3 + row] = ((col + row) & 1) ? -d : d;
        }
}
// intersect_lights (intersections.rs:87-91) on the ray as given: number of callback invocations
ORC_API int orc_kat_intersect_lights(const void *sv, const double *o, const double *d) {
    const Scene *s = (const Scene *)sv;
    Ray ray{v3(o[0], o[1], o[2]), v3(d[0], d[1], d[2])};
    Counters c;
    int calls = 0;
    auto cb_s = [&](const Intersection &, const Shape &) { ++calls; };
    auto cb_t = [&](const Intersection &, const Tri &) { ++calls; };
    bvh_intersections(s->lights.boxes, ray, cb_s, c);
    bvh_intersections(s->lights.ellipsoids, ray, cb_s, c);
    bvh_intersections(s->lights.triangles, ray, cb_t, c);
    return calls;
}
// RNG probes (distribution tests + device RNG parity)
ORC_API void orc_rng_u64(uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n, uint64_t *out) {
    Rng rng(seed, pixel, sample);
    for (uint64_t i = 0; i < n; ++i) out[i] = rng.next_u64();
}
ORC_API void orc_rng_probe(uint64_t seed, uint32_t pixel, uint32_t sample, double *out8) {
    Rng rng(seed, pixel, sample);
    out8[0] = rng.range(0.0, 1.0);
    out8[1] = rng.gen_f64();
    out8[2] = rng.range_inclusive(-1.0, 1.0);
    out8[3] = rng.range_inclusive(0.0, 1.0);
    out8[4] = (double)rng.range_usize(7);
    out8[5] = (double)rng.range_i32_0_1();
    out8[6] = rng.gen_bool(0.25) ? 1.0 : 0.0;
    out8[7] = rng.range(0.0, 3.5);
}
ORC_API int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }
