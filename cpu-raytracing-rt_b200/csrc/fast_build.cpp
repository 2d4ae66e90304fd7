// fast_build.cpp — host: FAST-mode culling structure.
//
// The reference hierarchy (reference src/bvh.rs:49-54: binary, one f64 box per node, children
// fetched separately) is COLLAPSED into 4-wide nodes that carry the fp32 boxes of their children:
// starting from a node's two children, the child with the largest surface area is opened in place
// until four slots are filled.  Boxes are rounded outward, so every ray that touches an f64 box
// touches its fp32 image; the left-to-right order of children and the primitive order are kept,
// which keeps the all-hits visit order of Light::pdf (src/bvh.rs:188-210) identical.
#include "fast_build.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <stdexcept>

// Longest run of triangles kept as one leaf of the culling hierarchy (<= 8, the leaf encoding).
// The reference's leaves hold up to 4 triangles (or more after a SameNode decision); shorter
// leaves here mean the cheap 4-wide box test culls single triangles before the fp32 pre-test.
#ifndef RT_MAX_LEAF
#define RT_MAX_LEAF 2
#endif

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
    uint32_t max_stack = 0;

    struct Item {                       // a child slot being assembled: a binary node or a primitive run
        bool is_node;
        uint32_t node, first, count;
        Box6 box;
    };
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
    static double half_area(const Box6 &b) {
        const double sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
        return sx * sy + sx * sz + sy * sz;
    }
    Item node_item(uint32_t n) const {
        const HostNode &h = bvh.nodes[n];
        if (h.left == RT_NO_CHILD) return Item{false, 0, h.first, h.count, h.box};
        return Item{true, n, 0, 0, h.box};
    }
    static bool expandable(const Item &it) { return it.is_node || it.count > RT_MAX_LEAF; }
    // children of an item, in the reference's left-to-right order
    void split(const Item &it, Item &a, Item &b) const {
        if (it.is_node) {
            a = node_item(bvh.nodes[it.node].left);
            b = node_item(bvh.nodes[it.node].right);
        } else {                        // a reference leaf longer than 8 primitives: halve the run
            const uint32_t half = it.count / 2;
            a = Item{false, 0, it.first, half, range_box(it.first, half)};
            b = Item{false, 0, it.first + half, it.count - half, range_box(it.first + half, it.count - half)};
        }
    }
    // emits the 4-wide node for `it` (which must be expandable); returns its index
    int32_t emit(const Item &it, uint32_t stack_above) {
        std::vector<Item> kids(2);
        split(it, kids[0], kids[1]);
        while (kids.size() < 4) {       // open the largest expandable child, keeping the order
            int best = -1;
            double best_area = -1.0;
            for (size_t k = 0; k < kids.size(); ++k)
                if (expandable(kids[k]) && half_area(kids[k].box) > best_area) { best = (int)k; best_area = half_area(kids[k].box); }
            if (best < 0) break;
            Item a, b;
            split(kids[best], a, b);
            kids[best] = a;
            kids.insert(kids.begin() + best + 1, b);
        }
        const int32_t me = (int32_t)out.nodes.size();
        out.nodes.push_back(FNode{});
        // a visit can leave up to (children - 1) entries on the traversal stack
        const uint32_t stack_here = stack_above + (uint32_t)kids.size() - 1;
        max_stack = std::max(max_stack, stack_here);
        int32_t codes[4];
        for (size_t k = 0; k < kids.size(); ++k) {
            const Item &c = kids[k];
            if (expandable(c)) {
                codes[k] = emit(c, stack_here);
            } else {
                if (c.count == 0) throw std::runtime_error("empty leaf inside a non-empty hierarchy");
                if (c.first >= (1u << 28)) throw std::runtime_error("too many triangles for the leaf encoding");
                codes[k] = ~(int32_t)((c.first << 3) | (c.count - 1));
            }
        }
        FNode &n = out.nodes[me];
        const float inf = std::numeric_limits<float>::infinity();
        for (int k = 0; k < 4; ++k) {
            if (k < (int)kids.size()) {
                for (int a = 0; a < 3; ++a) { n.lo[a][k] = round_down(kids[k].box.lo[a]); n.hi[a][k] = round_up(kids[k].box.hi[a]); }
                n.child[k] = codes[k];
            } else {
                for (int a = 0; a < 3; ++a) { n.lo[a][k] = inf; n.hi[a][k] = -inf; }
                n.child[k] = kFastEmpty;
            }
        }
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
    out.nodes.reserve(bvh.nodes.size() / 3 + 1);
    const Builder::Item root = b.node_item(0);
    if (Builder::expandable(root)) {
        out.root = b.emit(root, 0);
    } else {
        out.root = ~(int32_t)((root.first << 3) | (root.count - 1));
    }
    out.depth = b.max_stack;
    // Put the top of the tree first, in breadth-first order, so that a kernel can stage nodes
    // [0, top) in shared memory with one contiguous copy.  The rest keeps its depth-first order.
    if (out.root >= 0) {
        const size_t n = out.nodes.size();
        const size_t want = std::min<size_t>(n, kFastTopNodes);
        std::vector<int32_t> bfs;
        bfs.reserve(want);
        bfs.push_back(out.root);
        for (size_t head = 0; head < bfs.size() && bfs.size() < want; ++head)
            for (int k = 0; k < 4 && bfs.size() < want; ++k) {
                const int32_t c = out.nodes[bfs[head]].child[k];
                if (c >= 0) bfs.push_back(c);
            }
        std::vector<int32_t> new_index(n, -1);
        for (size_t i = 0; i < bfs.size(); ++i) new_index[bfs[i]] = (int32_t)i;
        int32_t next = (int32_t)bfs.size();
        for (size_t i = 0; i < n; ++i)
            if (new_index[i] < 0) new_index[i] = next++;
        std::vector<FNode> moved(n);
        for (size_t i = 0; i < n; ++i) {
            FNode nd = out.nodes[i];
            for (int k = 0; k < 4; ++k)
                if (nd.child[k] >= 0) nd.child[k] = new_index[nd.child[k]];
            moved[new_index[i]] = nd;
        }
        out.nodes.swap(moved);
        out.root = new_index[out.root];
        out.top = (uint32_t)bfs.size();
    }
}

}  // namespace rt
