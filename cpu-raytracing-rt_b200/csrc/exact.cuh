// exact.cuh — EXACT mode device functions: the reference's geometry queries operation for
// operation in f64, including its divisions and its depth-first visit order, so that hit ids,
// distances, normals and light PDFs are bit-identical to the CPU path.  Compiled with
// -fmad=false (the reference never contracts a*b+c).  Each function names what it replaces.
#pragma once
#include <cuda_runtime.h>

#include "device_scene.cuh"

namespace rt {

constexpr int kExactStack = 96;         // deeper hierarchies are walked without a stack (DBvh::parent, closest_exact_walk<DEEP>)
constexpr uint32_t kNoChild = 0xFFFFFFFFu;
#define RT_D __device__ __forceinline__

struct Cnt {                            // per-thread event counters, flushed once per kernel
    uint32_t seg = 0, lray = 0, paths = 0, aabb = 0, prim = 0, laabb = 0, lprim = 0;
    uint32_t xf = 0;                    // FAST mode: rays handed to the exact traversal
};

RT_D double dinf() { return __longlong_as_double(0x7FF0000000000000ll); }
RT_D d3 ld3(const double *p) { return mk3(p[0], p[1], p[2]); }

// 16-byte vector loads of a node: 3 x double2 (bounds) + 1 x uint4 (links)
RT_D void load_node_box(const DNode *n, double lo[3], double hi[3]) {
    const double2 *p = reinterpret_cast<const double2 *>(n);
    const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    lo[0] = a.x; lo[1] = a.y; lo[2] = b.x;
    hi[0] = b.y; hi[1] = c.x; hi[2] = c.y;
}
RT_D uint4 load_node_links(const DNode *n) { return __ldg(reinterpret_cast<const uint4 *>(n) + 3); }

// ---- AABB::intersects, reference src/aabb.rs:51-108 ----------------------------------------
RT_D double safe_min(double a, double b) {
    if (!isfinite(a)) return b;
    if (!isfinite(b)) return a;
    return rmin(a, b);
}
RT_D double safe_max(double a, double b) {
    if (!isfinite(a)) return b;
    if (!isfinite(b)) return a;
    return rmax(a, b);
}
RT_D bool aabb_exact(const double lo[3], const double hi[3], d3 o, d3 d, double &t) {
    if (d.x == 0.0 && (o.x < lo[0] || hi[0] < o.x)) return false;
    if (d.y == 0.0 && (o.y < lo[1] || hi[1] < o.y)) return false;
    if (d.z == 0.0 && (o.z < lo[2] || hi[2] < o.z)) return false;
    const bool outside = (o.x < lo[0] || hi[0] < o.x) || (o.y < lo[1] || hi[1] < o.y) || (o.z < lo[2] || hi[2] < o.z);
    if (!outside) { t = 0.0; return true; }
    const double ax = (lo[0] - o.x) / d.x, ay = (lo[1] - o.y) / d.y, az = (lo[2] - o.z) / d.z;
    const double bx = (hi[0] - o.x) / d.x, by = (hi[1] - o.y) / d.y, bz = (hi[2] - o.z) / d.z;
    const double t_near = safe_max(safe_max(safe_min(ax, bx), safe_min(ay, by)), safe_min(az, bz));
    const double t_far = safe_min(safe_min(safe_max(ax, bx), safe_max(ay, by)), safe_max(az, bz));
    if (t_near > t_far) return false;
    if (0.0 <= t_near) { t = t_near; return true; }
    if (0.0 <= t_far) { t = t_far; return true; }
    return false;
}
RT_D bool node_aabb_exact(const DNode *n, d3 o, d3 d, double &t) {
    double lo[3], hi[3];
    load_node_box(n, lo, hi);
    return aabb_exact(lo, hi, o, d, t);
}

// IEEE x / y with the exact-zero dividend answered directly: (+-0) / y = +-0 for finite non-zero
// y.  Axis-aligned triangles have exact zeros in their edge cross products, and the generic
// division sends a zero dividend down its ~40-instruction special-case path.  Same bits.
RT_D double divz1(double x, double y) {
    if (x == 0.0 && y == y && y != 0.0)
        return __longlong_as_double((__double_as_longlong(x) ^ __double_as_longlong(y)) & (long long)0x8000000000000000ull);
    return x / y;
}
RT_D d3 divz(d3 a, double s) { return mk3(divz1(a.x, s), divz1(a.y, s), divz1(a.z, s)); }

// ---- Triangle::intersection, reference src/primitives/triangle.rs:49-80 ----------------------
RT_D bool tri_exact(const DTriI *tp, d3 o, d3 d, double &t, double &u, double &v) {
    const double2 *p = reinterpret_cast<const double2 *>(tp);
    const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    const double q4 = __ldg(reinterpret_cast<const double *>(tp) + 8);
    const d3 a = mk3(q0.x, q0.y, q1.x), c0 = mk3(q1.y, q2.x, q2.y), c1 = mk3(q3.x, q3.y, q4), c2 = neg(d);
    const double det = c0.x * (c1.y * c2.z - c2.y * c1.z) - c1.x * (c0.y * c2.z - c2.y * c0.z) +
                       c2.x * (c0.y * c1.z - c1.y * c0.z);
    if (fabs(det) < kTriDetMin) return false;
    const d3 r0 = divz(cross3(c1, c2), det), r1 = divz(cross3(c2, c0), det), r2 = divz(cross3(c0, c1), det);
    const d3 w = sub(o, a);
    u = dot3(r0, w);
    v = dot3(r1, w);
    t = dot3(r2, w);
    return !(u < 0.0 || v < 0.0 || 1.0 < u + v || t < 0.0);
}
// normals of a confirmed triangle hit (shading_normal :30-32, inside/flip :73-78), then the
// renormalisation every hit gets in intersect() (src/intersections.rs:31-40 with DONT_ROTATE)
RT_D void tri_finish(const DTriS *sp, d3 d, double u, double v, d3 &ng, d3 &ns, bool &inside) {
    const double2 *p = reinterpret_cast<const double2 *>(sp);
    const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3), q4 = __ldg(p + 4), q5 = __ldg(p + 5);
    const d3 n = mk3(q0.x, q0.y, q1.x), na = mk3(q1.y, q2.x, q2.y), nb = mk3(q3.x, q3.y, q4.x), nc = mk3(q4.y, q5.x, q5.y);
    const d3 sh = unit3(add(add(na, scale_l(u, sub(nb, na))), scale_l(v, sub(nc, na))));
    inside = dot3(d, n) > 0.0;
    ng = inside ? neg(n) : n;
    ns = inside ? neg(sh) : sh;
}

// ---- Primitive<T>: model_space_ray, reference src/intersections.rs:93-104 --------------------
struct ShapeRec {
    d3 param, position;
    quat rot;
    int type;
    uint32_t material;
};
RT_D ShapeRec load_shape(const DShape *s) {
    const double2 *p = reinterpret_cast<const double2 *>(s);
    const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3), q4 = __ldg(p + 4);
    const uint4 m = __ldg(reinterpret_cast<const uint4 *>(s) + 5);
    ShapeRec r;
    r.param = mk3(q0.x, q0.y, q1.x);
    r.position = mk3(q1.y, q2.x, q2.y);
    r.rot = mkq_xyzw(q3.x, q3.y, q4.x, q4.y);
    r.type = (int)m.x;
    r.material = m.y;
    return r;
}
RT_D void model_ray(const ShapeRec &s, d3 o, d3 d, d3 &mo, d3 &md) {
    const quat c = qconj(s.rot);
    mo = qrot(c, sub(o, s.position));
    md = qrot(c, d);
}
// Plane::intersection, reference src/primitives/plane.rs:11-21 (model space)
RT_D bool plane_exact(d3 n, d3 o, d3 d, double &t, d3 &normal) {
    const double nd = dot3(n, d);
    t = -dot3(n, o) / nd;
    if (t < 0.0) return false;
    normal = scale_l(nd <= 0.0 ? 1.0 : -1.0, n);
    return true;
}
// intersect_box_coef, reference src/primitives/box.rs:75-115.  Returns 0 None, 1 One(far), 2 Two.
struct BoxFace {
    double t, sign;
    int dim;
};
RT_D d3 face_normal(BoxFace f) {                          // real_normal, box.rs:64-72
    return f.dim == 0 ? mk3(f.sign, 0.0, 0.0) : (f.dim == 1 ? mk3(0.0, f.sign, 0.0) : mk3(0.0, 0.0, f.sign));
}
RT_D int box_exact(d3 s, d3 o, d3 d, BoxFace &nearf, BoxFace &farf) {
    bool have = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double si = comp(s, i), oi = comp(o, i), di = comp(d, i);
        if (di == 0.0 && si < fabs(oi)) return 0;
        if (di == 0.0) continue;
        const double t1 = (si - oi) / di, t2 = (-si - oi) / di;
        BoxFace lo, hi;
        if (t1 < t2) { lo.t = t1; hi.t = t2; lo.sign = hi.sign = 1.0; } else { lo.t = t2; hi.t = t1; lo.sign = hi.sign = -1.0; }
        lo.dim = hi.dim = i;
        if (have) {
            nearf = (lo.t < nearf.t) ? nearf : lo;         // new.max(old): ties keep the NEW (later) axis
            farf = (hi.t < farf.t) ? hi : farf;            // new.min(old): ties keep the OLD axis
        } else {
            nearf = lo; farf = hi; have = true;
        }
    }
    if (!have) return 0;
    if (farf.t < nearf.t) return 0;
    if (0.0 <= nearf.t) return 2;
    if (0.0 <= farf.t) return 1;
    return 0;
}
// intersect_ellipsoid_coef, reference src/primitives/ellipsoid.rs:49-76. 0 None, 1 One(t2), 2 Two.
RT_D int ellipsoid_exact(d3 r, d3 o, d3 d, double &t1, double &t2) {
    const d3 so = div3(o, r), sd = div3(d, r);
    const double c = dot3(so, so), b = dot3(so, sd), a = dot3(sd, sd);
    const double disc = b * b - a * (c - 1.0);
    if (disc < 0.0) return 0;
    const double ds = sqrt(disc);
    t1 = (-b + ds) / a;
    t2 = (-b - ds) / a;
    if (t2 < t1) { const double x = t1; t1 = t2; t2 = x; }
    if (0.0 <= t1) return 2;
    if (0.0 <= t2) return 1;
    return 0;
}
RT_D d3 ellipsoid_normal(d3 r, d3 o, d3 d, double t) {    // ellipsoid.rs:26,29
    return unit3(div3(div3(add(o, scale(d, t)), r), r));
}
// Intersectable::intersection for a shape (closest only): t, model-space normal, inside
RT_D bool shape_closest(const ShapeRec &s, d3 o, d3 d, double &t, d3 &n, bool &inside) {
    d3 mo, md;
    model_ray(s, o, d, mo, md);
    if (s.type == 1) {                                    // RT_SHAPE_BOX, box.rs:21-33
        BoxFace nf, ff;
        const int k = box_exact(s.param, mo, md, nf, ff);
        if (k == 0) return false;
        const BoxFace f = k == 2 ? nf : ff;
        t = f.t; n = face_normal(f); inside = k == 1;
        return true;
    }
    if (s.type == 2) {                                    // RT_SHAPE_ELLIPSOID, ellipsoid.rs:21-32
        double t1, t2;
        const int k = ellipsoid_exact(s.param, mo, md, t1, t2);
        if (k == 0) return false;
        if (k == 2) { t = t1; n = ellipsoid_normal(s.param, mo, md, t1); inside = false; }
        else { t = t2; n = neg(ellipsoid_normal(s.param, mo, md, t2)); inside = true; }
        return true;
    }
    inside = false;
    return plane_exact(s.param, mo, md, t, n);
}

// ---- BVH::intersection / Node::intersection, reference src/bvh.rs:27-36, :151-186, :213-222 --
// KIND 3: triangles, otherwise shapes.  Replays the recursive near-first descent with an explicit
// stack: the deferred child is re-checked against the CURRENT best when popped (:170, :176).
struct Best {
    bool has;
    double t;
    uint32_t prim;
};
// DEEP = true is the same walk WITHOUT a stack, for hierarchies deeper than kExactStack (a sweep-SAH
// tree that peels one primitive per level, reference src/bvh.rs:93-111, has no depth bound): nothing
// is pushed, and when a subtree is finished the walk climbs parent links (DBvh::parent) and re-derives
// what the stack would have held.  That is possible because a deferred entry is fully determined by
// the ray: its distance is the sibling's box distance (the clamp at :161-163 never changes a value
// that passes `< best`), the sibling is visited second exactly when its distance is below the best
// hit AT THAT MOMENT (the pop re-check, :170 / :176 — and a sibling that was never pushed because it
// was not below the older, larger best fails this test too), and which child went first is `tl < tr`
// (:167, :173).  Two box tests per climbed level replace 12 bytes of stack; they are not counted
// (the counters report the reference's calls).
template <int KIND, bool DEEP>
RT_D Best closest_exact_walk(const DBvh bvh, const void *prims, d3 o, d3 d, Cnt &c) {
    Best best{false, 0.0, 0u};
    if (bvh.n_prims == 0) return best;
    double tt;
    ++c.aabb;
    if (!node_aabb_exact(bvh.nodes, o, d, tt)) return best;
    uint32_t stack_n[DEEP ? 1 : kExactStack];
    double stack_t[DEEP ? 1 : kExactStack];
    int sp = 0;
    uint32_t node = 0;
    for (;;) {
        const uint4 links = load_node_links(bvh.nodes + node);
        bool descended = false;
        if (links.x == kNoChild) {
            for (uint32_t i = links.y; i < links.y + links.z; ++i) {      // :152-156
                ++c.prim;
                double t;
                bool hit;
                if (KIND == 3) {
                    double u, v;
                    hit = tri_exact(reinterpret_cast<const DTriI *>(prims) + i, o, d, t, u, v);
                } else {
                    d3 n; bool in;
                    hit = shape_closest(load_shape(reinterpret_cast<const DShape *>(prims) + i), o, d, t, n, in);
                }
                if (hit && (!best.has || t < best.t)) { best.has = true; best.t = t; best.prim = i; }   // :213-222
            }
        } else {
            const uint32_t l = node + 1, r = links.x;
            double tl, tr;
            c.aabb += 2;
            const bool hl = node_aabb_exact(bvh.nodes + l, o, d, tl);    // :158-159
            const bool hr = node_aabb_exact(bvh.nodes + r, o, d, tr);
            const double b = best.has ? best.t : dinf();                  // :161-163
            const double left = hl ? (tl < b ? tl : b) : b;
            const double right = hr ? (tr < b ? tr : b) : b;
            if (left < b) {
                if (right < b) {
                    if (left < right) { if (!DEEP) { stack_n[sp] = r; stack_t[sp] = right; ++sp; } node = l; }
                    else { if (!DEEP) { stack_n[sp] = l; stack_t[sp] = left; ++sp; } node = r; }
                } else {
                    node = l;
                }
                descended = true;
            } else if (right < b) {
                node = r;
                descended = true;
            }
        }
        if (descended) continue;
        if (!DEEP) {
            for (;;) {
                if (sp == 0) return best;
                --sp;
                const double b = best.has ? best.t : dinf();
                if (stack_t[sp] < b) { node = stack_n[sp]; break; }
            }
        } else {
            for (;;) {                                                      // climb until a pending sibling is found
                if (node == 0) return best;
                const uint32_t par = __ldg(bvh.parent + node);
                const uint32_t l = par + 1, r = load_node_links(bvh.nodes + par).x;
                double tl, tr;
                const bool hl = node_aabb_exact(bvh.nodes + l, o, d, tl);
                const bool hr = node_aabb_exact(bvh.nodes + r, o, d, tr);
                const bool went_first = !(hl && hr) || node == (tl < tr ? l : r);
                const bool other_hit = node == l ? hr : hl;
                const double other_t = node == l ? tr : tl;
                const uint32_t other = node == l ? r : l;
                const double b = best.has ? best.t : dinf();
                if (went_first && other_hit && other_t < b) { node = other; break; }
                node = par;
            }
        }
    }
}
// DEEP is a COMPILE-TIME choice of the kernel (kernels.cu launches the ...Deep instantiations when some
// hierarchy of the scene is deeper than the stack): the ordinary kernels carry no trace of the stackless walk.
template <int KIND, bool DEEP = false>
RT_D Best closest_exact(const DBvh bvh, const void *prims, d3 o, d3 d, Cnt &c) {
    return closest_exact_walk<KIND, DEEP>(bvh, prims, o, d, c);
}

// ---- intersections::intersect, reference src/intersections.rs:42-85 --------------------------
struct SceneHit {
    int kind;               // -1 miss, 0 plane, 1 box, 2 ellipsoid, 3 triangle
    uint32_t prim, material;
    double t;
    d3 ng, ns;
    bool inside;
};
RT_D void finish_shape_hit(const DShape *sp, d3 o, d3 d, SceneHit &h) {
    const ShapeRec s = load_shape(sp);
    double t; d3 n; bool in;
    shape_closest(s, o, d, t, n, in);                     // same inputs, same arithmetic: same t
    const d3 w = unit3(qrot(s.rot, n));                   // with_rotated_normal, :31-40
    h.ng = w; h.ns = w; h.inside = in; h.material = s.material;
}
template <bool DEEP = false>
RT_D SceneHit intersect_exact(const DScene &sc, d3 o, d3 d, Cnt &c) {
    SceneHit h;
    h.kind = -1; h.prim = 0; h.material = 0; h.t = 0.0; h.inside = false;
    h.ng = h.ns = mk3(0.0, 0.0, 0.0);
    bool has = false;
    for (uint32_t i = 0; i < sc.n_planes; ++i) {          // :45-49
        ++c.prim;
        double t; d3 n; bool in;
        if (!shape_closest(load_shape(sc.planes + i), o, d, t, n, in)) continue;
        if (!has || t < h.t) { has = true; h.t = t; h.kind = 0; h.prim = i; }
    }
    const Best b = closest_exact<1, DEEP>(sc.boxes, sc.box_prims, o, d, c); // :51
    if (b.has && (!has || b.t < h.t)) { has = true; h.t = b.t; h.kind = 1; h.prim = b.prim; }
    const Best e = closest_exact<2, DEEP>(sc.ellipsoids, sc.ell_prims, o, d, c);  // :52
    if (e.has && (!has || e.t < h.t)) { has = true; h.t = e.t; h.kind = 2; h.prim = e.prim; }
    const Best t3 = closest_exact<3, DEEP>(sc.tris, sc.tri_i, o, d, c);     // :53
    if (t3.has && (!has || t3.t < h.t)) { has = true; h.t = t3.t; h.kind = 3; h.prim = t3.prim; }
    if (!has) return h;
    if (!(h.t * len3(d) <= dinf())) { h.kind = -1; return h; }             // :56 (false only for NaN)
    if (h.kind == 3) {
        double t, u, v;
        tri_exact(sc.tri_i + h.prim, o, d, t, u, v);
        d3 ng, ns;
        tri_finish(sc.tri_s + h.prim, d, u, v, ng, ns, h.inside);
        // DONT_ROTATE (:64) is still APPLIED: the identity rotation rewrites signed zeros and
        // the normalisation moves the last bits, so both are replayed rather than skipped.
        const quat ident = mkq_xyzw(0.0, 0.0, 0.0, 1.0);
        h.ng = unit3(qrot(ident, ng));
        h.ns = unit3(qrot(ident, ns));
        h.material = __ldg(&sc.tri_s[h.prim].material);         // same value as in the intersect record, on the line tri_finish just read
    } else {
        const DShape *sp = h.kind == 0 ? sc.planes + h.prim : (h.kind == 1 ? sc.box_prims + h.prim : sc.ell_prims + h.prim);
        finish_shape_hit(sp, o, d, h);
    }
    return h;
}

// ---- BVH::intersections / Node::intersections (all hits), reference src/bvh.rs:38-45, :188-210
// + intersect_lights (src/intersections.rs:87-91) + Light::pdf (src/ray_sampler.rs:132-139)
// + IntersectionProbability (src/intersection_probability.rs) + to_direction_probability.
// DEEP = true: no stack (see closest_exact_walk) — after a left subtree the walk climbs to the parent
// and enters the right child if the ray touches its box, after a right subtree it keeps climbing.
template <int KIND, bool DEEP>
RT_D void light_hits_exact_walk(const DBvh bvh, const void *prims, const DTriS *tri_s, d3 o, d3 d, double &impact, Cnt &c) {
    if (bvh.n_prims == 0) return;
    double tt;
    ++c.laabb;
    if (!node_aabb_exact(bvh.nodes, o, d, tt)) return;
    uint32_t stack_n[DEEP ? 1 : kExactStack];
    int sp = 0;
    stack_n[sp++] = 0;
    uint32_t node = 0;
    for (;;) {
        if (!DEEP) {
            if (sp == 0) return;
            node = stack_n[--sp];
        }
        bool descended = false;
        const uint4 links = load_node_links(bvh.nodes + node);
        if (links.x == kNoChild) {
            for (uint32_t i = links.y; i < links.y + links.z; ++i) {
                ++c.lprim;
                if (KIND == 3) {
                    double t, u, v;
                    if (!tri_exact(reinterpret_cast<const DTriI *>(prims) + i, o, d, t, u, v)) continue;
                    d3 ng, ns; bool in;
                    tri_finish(tri_s + i, d, u, v, ng, ns, in);
                    const double inv_area = __ldg(&tri_s[i].inv_area);
                    impact += inv_area * (t * t / fabs(dot3(d, ng)));
                } else {
                    const ShapeRec s = load_shape(reinterpret_cast<const DShape *>(prims) + i);
                    d3 mo, md;
                    model_ray(s, o, d, mo, md);
                    double t[2]; d3 n[2];
                    int k = 0;
                    if (KIND == 1) {                       // box.rs:35-46
                        BoxFace nf, ff;
                        const int r = box_exact(s.param, mo, md, nf, ff);
                        if (r == 2) { t[0] = nf.t; n[0] = face_normal(nf); t[1] = ff.t; n[1] = face_normal(ff); k = 2; }
                        else if (r == 1) { t[0] = ff.t; n[0] = face_normal(ff); k = 1; }
                    } else {                               // ellipsoid.rs:34-46
                        double t1, t2;
                        const int r = ellipsoid_exact(s.param, mo, md, t1, t2);
                        if (r == 2) { t[0] = t1; n[0] = ellipsoid_normal(s.param, mo, md, t1); t[1] = t2; n[1] = neg(ellipsoid_normal(s.param, mo, md, t2)); k = 2; }
                        else if (r == 1) { t[0] = t2; n[0] = neg(ellipsoid_normal(s.param, mo, md, t2)); k = 1; }
                    }
                    for (int j = 0; j < k; ++j) {
                        const d3 ng = unit3(qrot(s.rot, n[j]));
                        double prob;
                        if (KIND == 1) {
                            prob = 1.0 / ((s.param.y * s.param.z + s.param.x * s.param.z) + s.param.x * s.param.y) / 8.0;
                        } else {
                            const d3 coef = mul3(mk3(s.param.y * s.param.z, s.param.x * s.param.z, s.param.x * s.param.y), ng);
                            prob = 1.0 / (4.0 * kPi * sqrt(dot3(coef, coef)));
                        }
                        impact += prob * (t[j] * t[j] / fabs(dot3(d, ng)));
                    }
                }
            }
        } else {
            const uint32_t l = node + 1, r = links.x;
            c.laabb += 2;
            const bool hl = node_aabb_exact(bvh.nodes + l, o, d, tt);
            const bool hr = node_aabb_exact(bvh.nodes + r, o, d, tt);
            if (!DEEP) {
                if (hr) stack_n[sp++] = r;                 // left subtree completes before the right one
                if (hl) stack_n[sp++] = l;
            } else if (hl || hr) {
                node = hl ? l : r;
                descended = true;
            }
        }
        if (DEEP && !descended) {
            for (;;) {
                if (node == 0) return;
                const uint32_t par = __ldg(bvh.parent + node);
                const uint32_t r = load_node_links(bvh.nodes + par).x;
                const bool from_left = node == par + 1;
                node = par;
                if (from_left && node_aabb_exact(bvh.nodes + r, o, d, tt)) { node = r; break; }
            }
        }
    }
}
template <int KIND, bool DEEP = false>
RT_D void light_hits_exact(const DBvh bvh, const void *prims, const DTriS *tri_s, d3 o, d3 d, double &impact, Cnt &c) {
    light_hits_exact_walk<KIND, DEEP>(bvh, prims, tri_s, o, d, impact, c);
}
// o = pos + EPSILON * dir (ray_sampler.rs:133) is formed by the caller
template <bool DEEP = false>
RT_D double light_pdf_exact(const DScene &sc, d3 o, d3 dir, Cnt &c) {
    double impact = 0.0;
    ++c.lray;
    light_hits_exact<1, DEEP>(sc.lboxes, sc.lbox_prims, nullptr, o, dir, impact, c);
    light_hits_exact<2, DEEP>(sc.lellipsoids, sc.lell_prims, nullptr, o, dir, impact, c);
    light_hits_exact<3, DEEP>(sc.ltris, sc.ltri_i, sc.ltri_s, o, dir, impact, c);
    return impact / (double)sc.n_lights;
}

}  // namespace rt
