// device_scene.cuh — layout of the scene in HBM (see DESIGN.md "Data layout").
//
// EXACT mode reads the reference's own hierarchy (reference src/bvh.rs:49-54) flattened to
// 64-byte nodes: pre-order, left child = self + 1 (src/bvh.rs:104-108), so only `right` is
// stored.  Every record is a multiple of 16 bytes and is fetched with 16-byte vector loads.
// FAST mode reads a derived fp32 structure (two child boxes per 64-byte node, rounded outward)
// and confirms candidates against the same f64 triangle records.
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
// Inner node of the fp32 culling hierarchy: the boxes of BOTH children (rounded outward from
// the f64 boxes and padded by the scene-scale rounding slack), so one 64-byte fetch decides both
// descents.  child < 0 encodes a leaf: ~child = index into the leaf table.
struct __align__(16) FNode {            // 64 B
    float lo0[3], hi0[3];               // left child box
    float lo1[3], hi1[3];               // right child box
    int32_t child0, child1;             // >= 0: FNode index; < 0: ~leaf (first | count<<?) see FLeaf
    uint32_t _pad[2];
};
static_assert(sizeof(FNode) == 64, "FNode must be 64 bytes");

struct __align__(16) FTri {             // fp32 pre-test record, 48 B = 3 x 16 B
    float a[3], ba[3], ca[3];
    float err;                          // absolute slack covering f64->f32 rounding of this record
    uint32_t _pad[2];
};
static_assert(sizeof(FTri) == 48, "FTri must be 48 bytes");

struct DBvh {
    const DNode *nodes;
    uint32_t n_nodes, n_prims;
};

struct FBvh {
    const FNode *nodes;                 // nullptr when the hierarchy has <= 1 leaf
    const uint2 *leaves;                // (first, count) per leaf
    const FTri *tris;
    int32_t root;                       // >= 0 inner node, < 0 ~leaf, INT32_MIN: empty
    uint32_t _pad;
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

struct DCounters {                      // mirrors rt_counters (kernel_launches is host-side)
    unsigned long long segments, light_pdf_rays, paths, aabb_tests, prim_tests, light_aabb_tests, light_prim_tests;
};

}  // namespace rt
