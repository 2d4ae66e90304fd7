// scene_build.cpp — host: primitive baking and the bit-exact SAH hierarchy build.
//
// The reference builds each BVH single-threaded by physically sorting ~290-byte primitive
// structs four times per node (reference src/bvh.rs:75-113, :120-135).  This builder produces
// the SAME node array and primitive order from 16-byte (key, index) items: per node three stable
// key sorts + one re-sort, a suffix-box sweep, and task-parallel subtrees that are stitched back
// into the reference's pre-order numbering.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <deque>
#include <exception>
#include <limits>
#include <stdexcept>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include "scene_host.hpp"

namespace rt {
namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

inline Box6 empty_box() {           // AABB::empty, src/aabb.rs:16-21
    return Box6{{kInf, kInf, kInf}, {-kInf, -kInf, -kInf}};
}
inline void grow(Box6 &b, const Box6 &o) {      // extend_aabb, src/aabb.rs:28-31
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = rmin(b.lo[k], o.lo[k]);
        b.hi[k] = rmax(b.hi[k], o.hi[k]);
    }
}
inline void grow(Box6 &b, d3 p) {               // extend, src/aabb.rs:23-26
    const double v[3] = {p.x, p.y, p.z};
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = rmin(b.lo[k], v[k]);
        b.hi[k] = rmax(b.hi[k], v[k]);
    }
}
inline double half_area(const Box6 &b) {        // aabb_score, src/bvh.rs:115-118
    const double sx = b.hi[0] - b.lo[0], sy = b.hi[1] - b.lo[1], sz = b.hi[2] - b.lo[2];
    return sx * sy + sx * sz + sy * sz;
}
inline int64_t order_key(double v) {            // f64::total_cmp as an integer key (src/bvh.rs:139)
    int64_t b;
    std::memcpy(&b, &v, 8);
    return b ^ (int64_t)(((uint64_t)(b >> 63)) >> 1);
}

struct Item {
    int64_t key;
    uint32_t idx;
    uint32_t pos;                   // position before the sort: (key, pos) makes an unstable sort stable
};

// Runs fn(t) for t in [0, team) on `team` threads (the caller is one of them).  Used by the few
// nodes at the top of the hierarchy, where one node is all the work there is: their sorts and
// sweeps are split over the threads that the subtree tasks further down will use afterwards.
// A worker thread that is ALWAYS joined — also when the spawning scope unwinds (bad_alloc in the
// other branch) — and whose own exception is re-thrown in the spawning thread by get(): nothing
// may end in std::terminate behind the C ABI (capi.cpp's guarded() only sees exceptions that
// arrive on the calling thread).
class Task {
    std::exception_ptr err_;
    std::thread th_;
public:
    template <class F>
    explicit Task(F &&f) : th_([this, fn = std::forward<F>(f)]() mutable {
        try { fn(); } catch (...) { err_ = std::current_exception(); }
    }) {}
    Task(const Task &) = delete;
    Task &operator=(const Task &) = delete;
    ~Task() { if (th_.joinable()) th_.join(); }
    void get() {
        if (th_.joinable()) th_.join();
        if (err_) std::rethrow_exception(err_);
    }
};
template <class F>
void run_team(int team, F &&fn) {
    if (team <= 1) { fn(0); return; }
    std::deque<Task> ts;                // a deque never moves its elements (Task hands `this` to its thread)
    for (int t = 1; t < team; ++t) ts.emplace_back([&fn, t]() { fn(t); });
    fn(0);
    for (auto &th : ts) th.get();
}
inline size_t chunk_begin(size_t n, int team, int t) { return n * (size_t)t / (size_t)team; }

// Stable sort of `ord[0..n)` by the midpoint of `axis`.  Small ranges: insertion sort; large
// ranges: LSD radix over the key bytes that actually differ; otherwise introsort on (key, position).
// LSD radix digits: 11 bits -> 6 passes over a 64-bit key instead of 8 (the 2048-entry histogram
// still sits in L1)
constexpr int kRadixBits = 11, kRadixPasses = 6;
constexpr uint64_t kRadixMask = (1ull << kRadixBits) - 1;

struct Sorter {
    std::vector<Item> a, b;
    // returns true when no two keys are equal: the order is then a function of the keys alone
    bool sort(const Box6 *boxes, uint32_t *ord, size_t n, int axis, int team = 1) {
        a.resize(n);
        if (team > 1) return sort_team(boxes, ord, n, axis, team);
        int64_t all_or = 0, all_and = -1;
        for (size_t i = 0; i < n; ++i) {
            const Box6 &bx = boxes[ord[i]];
            const int64_t k = order_key((bx.lo[axis] + bx.hi[axis]) / 2.0);     // src/bvh.rs:138
            a[i] = Item{k, ord[i], (uint32_t)i};
            all_or |= k;
            all_and &= k;
        }
        if (n <= 32) {
            for (size_t i = 1; i < n; ++i) {
                Item v = a[i];
                size_t j = i;
                while (j > 0 && v.key < a[j - 1].key) { a[j] = a[j - 1]; --j; }
                a[j] = v;
            }
        } else if (n >= 4096) {
            radix(n, (uint64_t)(all_or ^ all_and));
        } else {
            // introsort on (key, incoming position): the stable order without stable_sort's allocation
            std::sort(a.begin(), a.end(), [](const Item &x, const Item &y) { return x.key < y.key || (x.key == y.key && x.pos < y.pos); });
        }
        bool ties = false;
        ord[0] = a[0].idx;
        for (size_t i = 1; i < n; ++i) { ord[i] = a[i].idx; ties |= a[i].key == a[i - 1].key; }
        return !ties;
    }
    // The same LSD radix sort with every pass split over a team: per-thread histograms of
    // contiguous chunks, bucket offsets ordered (digit, thread), so the result is the stable
    // order — identical to the sequential one.
    bool sort_team(const Box6 *boxes, uint32_t *ord, size_t n, int axis, int team) {
        b.resize(n);
        std::vector<int64_t> ors(team, 0), ands(team, -1);
        run_team(team, [&](int t) {
            int64_t o = 0, an = -1;
            for (size_t i = chunk_begin(n, team, t); i < chunk_begin(n, team, t + 1); ++i) {
                const Box6 &bx = boxes[ord[i]];
                const int64_t k = order_key((bx.lo[axis] + bx.hi[axis]) / 2.0);
                a[i] = Item{k, ord[i], (uint32_t)i};
                o |= k; an &= k;
            }
            ors[t] = o; ands[t] = an;
        });
        int64_t all_or = 0, all_and = -1;
        for (int t = 0; t < team; ++t) { all_or |= ors[t]; all_and &= ands[t]; }
        const uint64_t differing = (uint64_t)(all_or ^ all_and);
        Item *src = a.data(), *dst = b.data();
        constexpr size_t kDigits = kRadixMask + 1;
        std::vector<size_t> hist((size_t)team * kDigits);
        for (int pass = 0; pass < kRadixPasses; ++pass) {
            const int sh = pass * kRadixBits;
            if (((differing >> sh) & kRadixMask) == 0) continue;
            const uint64_t flip = pass == kRadixPasses - 1 ? (1ull << (63 - sh)) : 0u;
            run_team(team, [&](int t) {
                size_t *h = hist.data() + (size_t)t * kDigits;
                for (size_t d = 0; d < kDigits; ++d) h[d] = 0;
                for (size_t i = chunk_begin(n, team, t); i < chunk_begin(n, team, t + 1); ++i) ++h[(((uint64_t)src[i].key >> sh) & kRadixMask) ^ flip];
            });
            size_t sum = 0;
            for (size_t d = 0; d < kDigits; ++d)
                for (int t = 0; t < team; ++t) { size_t &h = hist[(size_t)t * kDigits + d]; const size_t c = h; h = sum; sum += c; }
            run_team(team, [&](int t) {
                size_t *h = hist.data() + (size_t)t * kDigits;
                for (size_t i = chunk_begin(n, team, t); i < chunk_begin(n, team, t + 1); ++i) dst[h[(((uint64_t)src[i].key >> sh) & kRadixMask) ^ flip]++] = src[i];
            });
            std::swap(src, dst);
        }
        std::vector<char> tie(team, 0);
        run_team(team, [&](int t) {
            bool ties = false;
            for (size_t i = chunk_begin(n, team, t); i < chunk_begin(n, team, t + 1); ++i) {
                ord[i] = src[i].idx;
                ties |= i > 0 && src[i].key == src[i - 1].key;
            }
            tie[t] = ties;
        });
        for (int t = 0; t < team; ++t)
            if (tie[t]) return false;
        return true;
    }
    void radix(size_t n, uint64_t differing) {
        b.resize(n);
        Item *src = a.data(), *dst = b.data();
        for (int pass = 0; pass < kRadixPasses; ++pass) {
            const int sh = pass * kRadixBits;
            if (((differing >> sh) & kRadixMask) == 0) continue;        // digit identical in all keys
            size_t hist[kRadixMask + 1] = {0};
            // signed order: flip the sign bit, which is the top bit of the last (9-bit) digit
            const uint64_t flip = pass == kRadixPasses - 1 ? (1ull << (63 - sh)) : 0u;
            for (size_t i = 0; i < n; ++i) ++hist[(((uint64_t)src[i].key >> sh) & kRadixMask) ^ flip];
            size_t sum = 0;
            for (size_t d = 0; d <= kRadixMask; ++d) { size_t c = hist[d]; hist[d] = sum; sum += c; }
            for (size_t i = 0; i < n; ++i) dst[hist[(((uint64_t)src[i].key >> sh) & kRadixMask) ^ flip]++] = src[i];
            std::swap(src, dst);
        }
        if (src != a.data()) std::memcpy(a.data(), src, n * sizeof(Item));
    }
};

// nodes this large split their own sorts / sweeps over idle threads (RT_TEAM_MIN overrides: tests)
size_t team_min() {
    static const size_t v = []() {
        const char *e = std::getenv("RT_TEAM_MIN");
        const long long x = e ? std::atoll(e) : 0;
        return x >= 64 ? (size_t)x : (size_t)262144;
    }();
    return v;
}

struct BuildShared {
    const Box6 *boxes;
    std::atomic<int> spare_threads;
};

struct Worker {
    BuildShared &sh;
    Sorter sorter;
    std::vector<Box6> suffix;
    std::vector<uint32_t> saved[2];         // the X- and Y-sorted arrangements of the node being split
    explicit Worker(BuildShared &s) : sh(s) {}

    // Evaluates every split position of the current arrangement (src/bvh.rs:120-135).
    void sweep(const uint32_t *ord, size_t n, int axis, size_t &best_first, double &best_score, int &best_axis) {
        suffix.resize(n);
        Box6 acc = empty_box();
        for (size_t i = n - 1; i >= 1; --i) {       // AABBSplitsBuilder backward half, :249-253
            grow(acc, sh.boxes[ord[i]]);
            suffix[i] = acc;
        }
        acc = empty_box();
        for (size_t i = 0; i + 1 < n; ++i) {
            grow(acc, sh.boxes[ord[i]]);
            const size_t left_cnt = i + 1, right_cnt = n - left_cnt;
            const double score = half_area(acc) * (double)left_cnt + half_area(suffix[i + 1]) * (double)right_cnt;
            if (score < best_score) { best_first = left_cnt; best_score = score; best_axis = axis; }
        }
    }

    // The same sweep over a team: min / max are exactly associative, so per-chunk partial boxes
    // combined in order give bit-identical prefix and suffix boxes; every thread then scores its
    // chunk and the first minimum over the chunks in order is the sequential answer (strict <).
    void sweep_team(const uint32_t *ord, size_t n, int axis, size_t &best_first, double &best_score, int &best_axis, int team) {
        suffix.resize(n);
        std::vector<Box6> part(team), pre(team);
        run_team(team, [&](int t) {             // suffix boxes inside each chunk
            Box6 acc = empty_box();
            const size_t lo = std::max<size_t>(chunk_begin(n, team, t), 1), hi = chunk_begin(n, team, t + 1);
            for (size_t i = hi; i-- > lo;) { grow(acc, sh.boxes[ord[i]]); suffix[i] = acc; }
            part[t] = acc;
        });
        std::vector<Box6> after(team);          // union of all chunks to the right of chunk t
        Box6 acc = empty_box();
        for (int t = team - 1; t >= 0; --t) { after[t] = acc; grow(acc, part[t]); }
        run_team(team, [&](int t) {             // complete the suffix boxes, prefix box of each chunk
            const size_t lo = std::max<size_t>(chunk_begin(n, team, t), 1), hi = chunk_begin(n, team, t + 1);
            for (size_t i = lo; i < hi; ++i) grow(suffix[i], after[t]);
            Box6 p = empty_box();
            for (size_t i = chunk_begin(n, team, t); i < hi; ++i) grow(p, sh.boxes[ord[i]]);
            pre[t] = p;
        });
        std::vector<Box6> before(team);
        acc = empty_box();
        for (int t = 0; t < team; ++t) { before[t] = acc; grow(acc, pre[t]); }
        std::vector<size_t> bf(team, n);
        std::vector<double> bs(team, best_score);
        run_team(team, [&](int t) {
            Box6 left = before[t];
            size_t f = n;
            double sc = best_score;             // candidates must beat the incoming best, like the sequential loop
            for (size_t i = chunk_begin(n, team, t); i < chunk_begin(n, team, t + 1) && i + 1 < n; ++i) {
                grow(left, sh.boxes[ord[i]]);
                const size_t left_cnt = i + 1, right_cnt = n - left_cnt;
                const double score = half_area(left) * (double)left_cnt + half_area(suffix[i + 1]) * (double)right_cnt;
                if (score < sc) { f = left_cnt; sc = score; }
            }
            bf[t] = f; bs[t] = sc;
        });
        for (int t = 0; t < team; ++t)
            if (bf[t] != n && bs[t] < best_score) { best_first = bf[t]; best_score = bs[t]; best_axis = axis; }
    }

    // build_nodes, src/bvh.rs:75-113.  Nodes are appended to `out` in pre-order; child indices
    // are relative to out[0].
    uint32_t build(uint32_t *ord, size_t n, size_t global_offset, std::vector<HostNode> &out) {
        Box6 box = empty_box();
        for (size_t i = 0; i < n; ++i) grow(box, sh.boxes[ord[i]]);
        auto leaf = [&]() {
            out.push_back(HostNode{box, RT_NO_CHILD, RT_NO_CHILD, (uint32_t)global_offset, (uint32_t)n});
            return (uint32_t)(out.size() - 1);
        };
        if (n <= 4) return leaf();
        size_t best_first = n;
        double best_score = half_area(box) * (double)n;
        int best_axis = -1;
        // a node of the top levels borrows the idle threads for its own sorts and sweeps
        int team = 1;
        if (n >= team_min()) {
            const int want = (int)std::min<size_t>(n / (team_min() / 2), 64) - 1;
            int got = 0;
            while (got < want && sh.spare_threads.fetch_sub(1) > 0) ++got;
            if (got < want) sh.spare_threads.fetch_add(1);          // undo the failed attempt
            team = got + 1;
        }
        bool tie_free[3];
        for (int axis = 0; axis < 3; ++axis) {
            tie_free[axis] = sorter.sort(sh.boxes, ord, n, axis, team);
            // without equal keys the arrangement is a function of the keys alone: the re-sort by
            // the winning axis below would reproduce exactly this one, so keep a copy instead
            if (axis < 2 && tie_free[axis] && n > 32) saved[axis].assign(ord, ord + n);
            if (team > 1) sweep_team(ord, n, axis, best_first, best_score, best_axis, team);
            else sweep(ord, n, axis, best_first, best_score, best_axis);
        }
        if (best_axis >= 0 && best_axis != 2) {     // a stable re-sort by Z of a Z-sorted range is the identity
            if (tie_free[best_axis] && n > 32) std::memcpy(ord, saved[best_axis].data(), n * sizeof(uint32_t));
            else sorter.sort(sh.boxes, ord, n, best_axis, team);
        }
        if (team > 1) sh.spare_threads.fetch_add(team - 1);
        if (best_axis < 0) return leaf();
        const uint32_t me = (uint32_t)out.size();
        out.push_back(HostNode{box, RT_NO_CHILD, RT_NO_CHILD, 0, 0});
        const size_t nl = best_first, nr = n - best_first;
        if (n >= 16384 && sh.spare_threads.fetch_sub(1) > 0) {
            // right subtree on another thread into its own array, stitched in afterwards
            std::vector<HostNode> right_nodes;
            Task th([&]() {
                Worker w(sh);
                w.build(ord + nl, nr, global_offset + nl, right_nodes);
            });
            const uint32_t l = build(ord, nl, global_offset, out);
            th.get();
            sh.spare_threads.fetch_add(1);
            const uint32_t shift = (uint32_t)out.size();
            out.reserve(out.size() + right_nodes.size());
            for (HostNode nd : right_nodes) {
                if (nd.left != RT_NO_CHILD) nd.left += shift;
                if (nd.right != RT_NO_CHILD) nd.right += shift;
                out.push_back(nd);
            }
            out[me].left = l;
            out[me].right = shift;
        } else {
            if (n >= 16384) sh.spare_threads.fetch_add(1);
            const uint32_t l = build(ord, nl, global_offset, out);
            const uint32_t r = build(ord + nl, nr, global_offset + nl, out);
            out[me].left = l;
            out[me].right = r;
        }
        return me;
    }
};

}  // namespace

void build_bvh(const Box6 *boxes, size_t n, std::vector<HostNode> &nodes, std::vector<uint32_t> &order,
               uint32_t &depth, int max_threads) {
    if (n >= 0xFFFFFFF0ull) throw std::runtime_error("too many primitives for 32-bit indices");
    order.resize(n);
    for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
    nodes.clear();
    if (max_threads <= 0) max_threads = (int)std::thread::hardware_concurrency();
    if (max_threads < 1) max_threads = 1;
    BuildShared sh{boxes, {max_threads - 1}};
    Worker w(sh);
    w.build(order.data(), n, 0, nodes);
    // depth (root = 1); children always follow their parent in pre-order
    std::vector<uint32_t> d(nodes.size(), 0);
    depth = 0;
    if (!nodes.empty()) d[0] = 1;
    for (size_t i = 0; i < nodes.size(); ++i) {
        depth = std::max(depth, d[i]);
        if (nodes[i].left != RT_NO_CHILD) d[nodes[i].left] = d[i] + 1;
        if (nodes[i].right != RT_NO_CHILD) d[nodes[i].right] = d[i] + 1;
    }
}

// ------------------------------------------------------------------------------------------------
namespace {

inline d3 ld3(const double *p) { return mk3(p[0], p[1], p[2]); }

void triangle_props(d3 a, d3 b, d3 c, HostTri &t) {     // props, src/primitives/triangle.rs:41-47
    t.a = a;
    t.ba = sub(b, a);
    t.ca = sub(c, a);
    const d3 sized = cross3(t.ba, t.ca);
    const double area = sqrt(dot3(sized, sized)) / 2.0;
    t.ng = unit3(sized);
    t.inv_area = 1.0 / area;
}

Box6 rotated_box(const Box6 &b, quat r) {               // rotated_aabb, src/scene.rs:255-268
    Box6 o = empty_box();
    for (int ix = 0; ix < 2; ++ix)
        for (int iy = 0; iy < 2; ++iy)
            for (int iz = 0; iz < 2; ++iz)
                grow(o, qrot(r, mk3(ix ? b.hi[0] : b.lo[0], iy ? b.hi[1] : b.lo[1], iz ? b.hi[2] : b.lo[2])));
    return o;
}

HostShape bake_shape(const rt_shape &s, uint32_t orig) {    // Primitive::new, src/scene.rs:108-137
    HostShape o{};
    o.type = s.type;
    o.material = s.material;
    o.param = ld3(s.param);
    o.position = ld3(s.position);
    o.rotation = mkq_xyzw(s.rotation[0], s.rotation[1], s.rotation[2], s.rotation[3]);
    o.orig = orig;
    o.box = empty_box();
    if (s.type != RT_SHAPE_PLANE) {
        Box6 local = empty_box();                           // Box::new, src/primitives/box.rs:12-17
        grow(local, o.param);
        grow(local, neg(o.param));
        o.box = rotated_box(local, o.rotation);
        const double p[3] = {o.position.x, o.position.y, o.position.z};
        for (int k = 0; k < 3; ++k) { o.box.lo[k] += p[k]; o.box.hi[k] += p[k]; }
    }
    return o;
}

HostTri bake_triangle(const rt_raw_scene &raw, uint64_t i) {
    const double *p = raw.tri_pos + i * 9;
    HostTri t{};
    triangle_props(ld3(p), ld3(p + 3), ld3(p + 6), t);
    if (raw.tri_nrm) {
        const double *n = raw.tri_nrm + i * 9;
        t.na = ld3(n); t.nb = ld3(n + 3); t.nc = ld3(n + 6);
    } else {
        t.na = t.nb = t.nc = t.ng;                          // new_with_geometry_normals, triangle.rs:25-28
    }
    t.material = raw.tri_mat ? raw.tri_mat[i] : 0;
    t.orig = (uint32_t)i;
    if (raw.tri_mode == RT_TRI_BAKED) {                     // instantiate, gltf/scene_builder.rs:42-55
        t.box = empty_box();
        grow(t.box, t.a);
        grow(t.box, add(t.a, t.ba));
        grow(t.box, add(t.a, t.ca));
        return t;
    }
    d3 position = mk3(0, 0, 0);                             // TrianglePrimitive::new, src/scene.rs:139-165
    quat rotation = mkq_xyzw(0, 0, 0, 1);
    if (raw.tri_xform) {
        const double *x = raw.tri_xform + i * 7;
        position = ld3(x);
        rotation = mkq_xyzw(x[3], x[4], x[5], x[6]);
    }
    const d3 wa = add(qrot(rotation, t.a), position);
    const d3 wb = add(qrot(rotation, add(t.ba, t.a)), position);
    const d3 wc = add(qrot(rotation, add(t.ca, t.a)), position);
    HostTri w{};
    triangle_props(wa, wb, wc, w);
    w.na = qrot(rotation, t.na);
    w.nb = qrot(rotation, t.nb);
    w.nc = qrot(rotation, t.nc);
    w.material = t.material;
    w.orig = t.orig;
    w.box = empty_box();
    grow(w.box, wa);
    grow(w.box, wb);
    grow(w.box, wc);
    return w;
}

template <class P>
void make_bvh(std::vector<P> input, HostBvh<P> &out) {      // BVH::new, src/bvh.rs:12-17
    std::vector<Box6> boxes(input.size());
    for (size_t i = 0; i < input.size(); ++i) boxes[i] = input[i].box;
    std::vector<uint32_t> order;
    build_bvh(boxes.data(), input.size(), out.nodes, order, out.depth);
    out.prims.resize(input.size());
    for (size_t i = 0; i < input.size(); ++i) out.prims[i] = input[order[i]];
}

inline bool emits(const rt_material &m) {                   // is_light, src/scene.rs:225-227
    return !(m.emission[0] == 0.0 && m.emission[1] == 0.0 && m.emission[2] == 0.0);
}

}  // namespace

// Everything of the Scene that is not geometry: image size, sample count, depth, background and
// Camera::new (src/camera.rs:18-46).  Separate so that cached hierarchies (scene_cache.cpp) can be
// paired with the view of the current run.
void apply_view(const rt_raw_scene &raw, HostScene &s) {
    if (raw.width == 0 || raw.height == 0) throw std::runtime_error("zero image dimensions");
    if (raw.ray_depth > 255) throw std::runtime_error("ray_depth is a u8 in the reference");
    s.width = raw.width; s.height = raw.height; s.samples = raw.samples; s.ray_depth = raw.ray_depth;
    s.bg_color = ld3(raw.bg_color);
    s.cam_position = ld3(raw.cam_position); s.cam_right = ld3(raw.cam_right);
    s.cam_up = ld3(raw.cam_up); s.cam_forward = ld3(raw.cam_forward);
    const double fw = (double)raw.width, fh = (double)raw.height;      // Camera::new, src/camera.rs:18-46
    if (raw.fov_kind == RT_FOV_Y) {
        s.tan_half_fov_y = tan(raw.fov / 2.0);
        s.tan_half_fov_x = s.tan_half_fov_y / (fh / fw);
    } else {
        s.tan_half_fov_x = tan(raw.fov / 2.0);
        s.tan_half_fov_y = s.tan_half_fov_x / (fw / fh);
    }
}

void build_scene(const rt_raw_scene &raw, HostScene &s) {
    if ((raw.n_materials && !raw.materials) || (raw.n_shapes && !raw.shapes) || (raw.n_triangles && !raw.tri_pos))
        throw std::runtime_error("null array in rt_raw_scene");
    s = HostScene{};
    apply_view(raw, s);
    s.materials.assign(raw.materials, raw.materials + raw.n_materials);
    auto check_mat = [&](uint32_t m) { if (m >= raw.n_materials) throw std::runtime_error("material index out of range"); };

    std::vector<HostShape> boxes, ellipsoids;
    for (uint32_t i = 0; i < raw.n_shapes; ++i) {           // make_scenes, src/scene.rs:194-207
        const rt_shape &sh = raw.shapes[i];
        check_mat(sh.material);
        if (sh.type == RT_SHAPE_BOX) boxes.push_back(bake_shape(sh, (uint32_t)boxes.size()));
        else if (sh.type == RT_SHAPE_ELLIPSOID) ellipsoids.push_back(bake_shape(sh, (uint32_t)ellipsoids.size()));
        else if (sh.type == RT_SHAPE_PLANE) s.planes.push_back(bake_shape(sh, (uint32_t)s.planes.size()));
        else throw std::runtime_error("unknown shape type");
    }
    // RT_TIMING=1: phase times of the host build on stderr (development aid)
    const bool timing = std::getenv("RT_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rt_b200] %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    };
    std::vector<HostTri> tris(raw.n_triangles);
    {
        const uint64_t n = raw.n_triangles;
        int nt = (int)std::thread::hardware_concurrency();
        if (nt < 1 || n < 65536) nt = 1;
        std::deque<Task> ts;
        for (int t = 0; t < nt; ++t)
            ts.emplace_back([&, t]() {
                for (uint64_t i = n * t / nt; i < n * (t + 1) / nt; ++i) tris[i] = bake_triangle(raw, i);
            });
        for (auto &t : ts) t.get();
        if (raw.tri_mat)
            for (uint64_t i = 0; i < n; ++i) check_mat(raw.tri_mat[i]);
        else if (n) check_mat(0);
    }
    auto pick_lights = [&](const auto &v) {
        std::remove_cv_t<std::remove_reference_t<decltype(v)>> out;
        for (const auto &p : v)
            if (emits(s.materials[p.material])) out.push_back(p);
        return out;
    };
    lap("bake triangles");
    const bool gltf = raw.tri_mode == RT_TRI_BAKED && raw.n_shapes == 0;
    // lights first (src/scene.rs:209-213), then the scene sets (:215-220); glTF scenes use
    // BVH::new_empty for the shape hierarchies (gltf/scene_builder.rs:14-15): no nodes at all.
    if (!gltf) {
        make_bvh(pick_lights(boxes), s.light_boxes);
        make_bvh(pick_lights(ellipsoids), s.light_ellipsoids);
    }
    make_bvh(pick_lights(tris), s.light_tris);
    if (!gltf) {
        make_bvh(std::move(ellipsoids), s.ellipsoids);
        make_bvh(std::move(boxes), s.boxes);
    }
    lap("light + shape hierarchies");
    make_bvh(std::move(tris), s.tris);
    lap("triangle hierarchy");
}

}  // namespace rt
