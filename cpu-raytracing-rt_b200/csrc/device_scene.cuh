// device_scene.cuh — layout of the scene in HBM (see DESIGN.md "Data layout").
//
// EXACT mode reads the reference's own hierarchy (reference src/bvh.rs:49-54) flattened to
// 64-byte nodes: pre-order, left child = self + 1 (src/bvh.rs:104-108), so only `right` is
// stored.  Every record is a multiple of 16 bytes and is fetched with 16-byte vector loads.
// FAST mode reads a derived fp32 structure (4-wide nodes of 128 bytes, child boxes rounded
// outward) and confirms candidates against the same f64 triangle records.
#pragma once
#include <vector_types.h>

#include <cstdint>

#include "rtmath.cuh"

namespace rt {

struct __align__(16) DNode {            // 64 B = 4 x 16 B
    double lo[3], hi[3];
    uint32_t right;                     // RT_NO_CHILD (0xFFFFFFFF) for leaves
    uint32_t first, count, _pad;
};
static_assert(sizeof(DNode) == 64, "DNode must be 64 bytes");

struct __align__(16) DTriI {            // intersect record, 80 B = 5 x 16 B  (triangle.rs:7-9)
    double a[3], ba[3], ca[3];
    uint32_t material, _pad;
};
static_assert(sizeof(DTriI) == 80, "DTriI must be 80 bytes");

struct __align__(16) DTriS {            // shading record, 112 B = 7 x 16 B (triangle.rs:10-16)
    double ng[3], na[3], nb[3], nc[3];
    double inv_area;
    uint32_t material, _pad;
};
static_assert(sizeof(DTriS) == 112, "DTriS must be 112 bytes");

struct __align__(16) DShape {           // Primitive<Box|Ellipsoid|Plane>, 96 B (scene.rs:21-27)
    double param[3], position[3];
    double rot[4];                      // x y z w
    int32_t type;
    uint32_t material;
    uint32_t _pad[2];
};
static_assert(sizeof(DShape) == 96, "DShape must be 96 bytes");

struct __align__(16) DMaterial {        // Metadata, 64 B (scene.rs:13-18)
    double color[3], emission[3];
    double ior;
    int32_t kind, _pad;
};
static_assert(sizeof(DMaterial) == 64, "DMaterial must be 64 bytes");

// ---- FAST mode structures -------------------------------------------------------------------
// Inner node of the fp32 culling hierarchy, derived on the host by collapsing the reference's
// binary tree into 4-wide nodes: the boxes of up to FOUR children, rounded OUTWARD from the f64
// boxes, one plane per 16-byte vector (lo.x of the 4 children, lo.y, lo.z, hi.x, hi.y, hi.z), so
// a ray loads the NEAR planes of all four children with three 16-byte loads whose offsets depend
// only on the ray's direction signs, and the FAR planes with three more.  128 B = one cache line.
// child >= 0: FNode index; child < 0: leaf, ~child = (first << 3) | (count - 1) with
// 1 <= count <= 8 (longer reference leaves are split on the host); kFastEmpty: unused slot
// (its box is +inf / -inf: never hit).  Child order = the reference's left-to-right order.
struct __align__(16) FNode {            // 128 B
    float lo[3][4];                     // lo[axis][child]
    float hi[3][4];
    int32_t child[4];
    uint32_t _pad[4];
};
static_assert(sizeof(FNode) == 128, "FNode must be 128 bytes");

// fp32 pre-test record: anchor vertex in f64 (the subtraction origin - a is done in f64, then
// rounded once), edges in fp32.  48 B = 3 x 16 B.
struct __align__(16) FTri {
    double a[3];
    float ba[3], ca[3];
};
static_assert(sizeof(FTri) == 48, "FTri must be 48 bytes");

constexpr int32_t kFastEmpty = INT32_MIN;   // FBvh::root of a hierarchy without primitives
constexpr uint32_t kFastTopNodes = 85;      // 1 + 4 + 16 + 64: the levels staged in shared memory (10.6 KB)

struct DBvh {
    const DNode *nodes;
    const uint32_t *parent;             // parent[node], uploaded ONLY for hierarchies deeper than the traversal stack
                                        // (kExactStack): non-null selects the stackless walk of exact.cuh
    uint32_t n_nodes, n_prims;
};

struct FBvh {
    const FNode *nodes;
    const FTri *tris;
    int32_t root;                       // >= 0 inner node, < 0 leaf code, kFastEmpty: no primitives
    uint32_t depth;                     // max number of simultaneously deferred children
    uint32_t top;                       // nodes [0, top) = top levels, breadth-first (shared-memory staging)
    float bound;                        // >= max |coordinate| of the hierarchy's root box (error scale of hit_fragile)
};

struct DCamera {                        // camera.rs:6-15
    d3 position, right, up, forward;
    double tan_half_fov_x, tan_half_fov_y, width, height;
};

struct DScene {
    DBvh boxes, ellipsoids, tris, lboxes, lellipsoids, ltris;
    const DShape *planes, *box_prims, *ell_prims, *lbox_prims, *lell_prims;
    const DTriI *tri_i, *ltri_i;
    const DTriS *tri_s, *ltri_s;
    const DMaterial *materials;
    FBvh ftris, fltris;
    uint32_t n_planes, n_lights;        // n_lights = LightPrimitives::len()
    uint32_t width, height, ray_depth, _pad;
    DCamera cam;
    d3 bg;
};

// Rectangle of the image one render call covers (tile sharding, previews of a region).  Paths are
// keyed by the pixel's index in the WHOLE image, so a region renders exactly the values the full
// image has there.
struct DRegion {
    uint32_t x0, y0, w, h;
};

struct DCounters {                      // mirrors rt_counters (kernel_launches is host-side)
    unsigned long long segments, light_pdf_rays, paths, aabb_tests, prim_tests, light_aabb_tests, light_prim_tests;
    unsigned long long exact_fallbacks;     // FAST mode: rays handed to the exact traversal (failed proof, tie, list overflow)
};

}  // namespace rt
