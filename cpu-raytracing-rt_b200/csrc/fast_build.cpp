// fast_build.cpp — host: FAST-mode culling structure.
//
// The reference hierarchy (reference src/bvh.rs:49-54, one f64 box per node, children fetched
// separately) is re-packed so that a parent carries BOTH child boxes in fp32.  Boxes are rounded
// outward, so every ray that touches an f64 box touches its fp32 image; topology, child order
// (left = first) and primitive order are unchanged, which keeps the all-hits visit order of
// Light::pdf (src/bvh.rs:188-210) identical.
#include "fast_build.hpp"

#include <cmath>
#include <limits>
#include <stdexcept>

namespace rt {
namespace {

float round_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -std::numeric_limits<float>::infinity());
    return f;
}
float round_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, std::numeric_limits<float>::infinity());
    return f;
}

struct Builder {
    const HostBvh<HostTri> &bvh;
    FastHierarchy &out;
    uint32_t max_depth = 0;

    Box6 range_box(uint32_t first, uint32_t count) const {
        const double inf = std::numeric_limits<double>::infinity();
        Box6 b{{inf, inf, inf}, {-inf, -inf, -inf}};
        for (uint32_t i = first; i < first + count; ++i)
            for (int k = 0; k < 3; ++k) {
                b.lo[k] = rmin(b.lo[k], bvh.prims[i].box.lo[k]);
                b.hi[k] = rmax(b.hi[k], bvh.prims[i].box.hi[k]);
            }
        return b;
    }
    static void put_box(float *lo, float *hi, const Box6 &b) {
        for (int k = 0; k < 3; ++k) { lo[k] = round_down(b.lo[k]); hi[k] = round_up(b.hi[k]); }
    }
    // a run of primitives as a child: a leaf code, or (count > 8) a synthesized split
    int32_t emit_range(uint32_t first, uint32_t count, uint32_t depth) {
        if (count == 0) throw std::runtime_error("empty leaf inside a non-empty hierarchy");
        if (count <= 8) {
            if (first >= (1u << 28)) throw std::runtime_error("too many triangles for the leaf encoding");
            return ~(int32_t)((first << 3) | (count - 1));
        }
        const uint32_t half = count / 2;
        const int32_t me = (int32_t)out.nodes.size();
        out.nodes.push_back(FNode{});
        max_depth = std::max(max_depth, depth + 1);
        const int32_t c0 = emit_range(first, half, depth + 1), c1 = emit_range(first + half, count - half, depth + 1);
        FNode &n = out.nodes[me];
        put_box(n.lo0, n.hi0, range_box(first, half));
        put_box(n.lo1, n.hi1, range_box(first + half, count - half));
        n.child0 = c0;
        n.child1 = c1;
        return me;
    }
    int32_t emit(uint32_t node, uint32_t depth) {
        const HostNode &h = bvh.nodes[node];
        if (h.left == RT_NO_CHILD) return emit_range(h.first, h.count, depth);
        const int32_t me = (int32_t)out.nodes.size();
        out.nodes.push_back(FNode{});
        max_depth = std::max(max_depth, depth + 1);
        const int32_t c0 = emit(h.left, depth + 1), c1 = emit(h.right, depth + 1);
        FNode &n = out.nodes[me];
        put_box(n.lo0, n.hi0, bvh.nodes[h.left].box);
        put_box(n.lo1, n.hi1, bvh.nodes[h.right].box);
        n.child0 = c0;
        n.child1 = c1;
        return me;
    }
};

}  // namespace

void build_fast_hierarchy(const HostBvh<HostTri> &bvh, FastHierarchy &out) {
    out = FastHierarchy{};
    if (bvh.prims.empty()) return;
    out.tris.resize(bvh.prims.size());
    for (size_t i = 0; i < bvh.prims.size(); ++i) {
        const HostTri &t = bvh.prims[i];
        FTri f{};
        f.a[0] = t.a.x; f.a[1] = t.a.y; f.a[2] = t.a.z;
        f.ba[0] = (float)t.ba.x; f.ba[1] = (float)t.ba.y; f.ba[2] = (float)t.ba.z;
        f.ca[0] = (float)t.ca.x; f.ca[1] = (float)t.ca.y; f.ca[2] = (float)t.ca.z;
        out.tris[i] = f;
    }
    Builder b{bvh, out};
    out.nodes.reserve(bvh.nodes.size() / 2 + 1);
    out.root = b.emit(0, 0);
    out.depth = b.max_depth;
}

}  // namespace rt
