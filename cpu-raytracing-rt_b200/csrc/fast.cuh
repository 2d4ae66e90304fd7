// fast.cuh — FAST mode geometry queries: conservative fp32 culling, exact f64 confirmation.
//
// B200 has no RT cores; its fp32 pipes are twice as wide as its f64 pipes and an IEEE f64
// division costs ~20 instructions.  The reference spends 12 f64 divisions per node visit
// (reference src/aabb.rs:62-63) and 9 per triangle test (src/primitives/triangle.rs:56-58).
// FAST mode never divides while culling:
//   * boxes: a 4-wide hierarchy of fp32 boxes (collapsed from the reference's binary tree, rounded
//     outward on the host), slab-tested with a precomputed reciprocal direction; every rounding
//     error of the slab arithmetic is bounded and folded into per-ray constants, so any ray that
//     touches the f64 box passes the fp32 test;
//   * triangles: an fp32 Moller-Trumbore pre-test in determinant-scaled form whose rejection
//     thresholds carry the error bound of each triple product, so it can only reject triangles
//     the exact test also rejects;
//   * every surviving candidate runs tri_exact (the reference's own arithmetic) and only its
//     f64 (t, u, v) are ever used.  Results are therefore the EXACT-mode ones (ids, t, normals,
//     light PDFs), which tests/test_gpu_parity.py::test_fast_mode_equals_exact_mode asserts bitwise.
// Range: fp32 culling is used while the scene bound and the ray origin stay below kFastRange (1e15:
// far from every fp32 overflow — triple products of edges, __fdividef's 2^126 limit); beyond it the
// ray is answered by the exact traversal (fast_in_range).  Inside the range, a ray that starts very
// far from a small triangle (distance / triangle size ~ 1e9) is dominated by the f64 rounding of
// `origin - a`, in the reference's arithmetic as much as here: that error is folded into the pre-test's
// bounds (fast_wpad), so such rays turn into undecided candidates — confirmed exactly — never into
// wrong rejections.
#pragma once
#include "exact_limits.hpp"
#include "shade.cuh"

namespace rt {

constexpr int kFastStack = 64;              // checked against FBvh::depth at upload time
constexpr float kU = 5.9604645e-8f;         // 2^-24, fp32 unit roundoff
// Bound on the rounding error of an fp32 triple product w . (d x e) evaluated with FMAs from
// inputs that were each rounded once to fp32: every product term carries at most 4 roundings of
// relative size u on top of the 4u of the cross-product component it multiplies, so
// |computed - exact| <= 8u * sum_k |w_k| (|d_i e_j| + |d_j e_i|) <= 8u * |w|_1 |d|_1 |e|_1.
constexpr float kGamma = 5.4e-7f;           // 9u > 8u (1 + O(u))
// The pre-test forms w = (o + tref*d) - a in f64 and rounds it once to fp32; its error bounds are
// proportional to |w|_1 and assume that the f64 step is exact.  It is not: o - a is rounded (relative
// 2^-53 of |o - a|, the SAME rounded difference the reference's test starts from, triangle.rs:66),
// the reference's own matrix-vector product adds ~7 roundings of that size, and tref*d adds one.  All of
// it is bounded by 16 * 2^-53 * (|tref*d|_1 + |o - a|_1) <= 16 * 2^-53 * (2|o|_1 + 6*bound): the entry
// point o + tref*d lies between the origin and the leaf, and |a|_1 <= 3*bound.  Expressed as an
// addition to |w|_1 (divide by kGamma) it is one per-ray constant.
RT_D float fast_wpad(d3 o, float bound) {
    const float o1 = __double2float_ru(fabs(o.x) + fabs(o.y) + fabs(o.z));
    return 6.7e-9f * (o1 + 3.0f * bound);   // 2 * 16 * 2^-53 / kGamma = 6.6e-9
}
RT_D bool fast_in_range(d3 o, float bound) {
    // false also for a NaN / infinite origin
    return fabs(o.x) <= (double)kFastRange && fabs(o.y) <= (double)kFastRange && fabs(o.z) <= (double)kFastRange && bound <= kFastRange;
}

// Per-ray culling state.  For a box plane b the slab distance (b - o) / d is evaluated as
// fma(b, id, n) with id = 1/fl(d), n = -(fl(o) * id):  computed = true*(1+eta) + delta with
// |eta| <= 3u and |delta| <= 2.2u|n|.  The constants below fold that bound in once per ray:
//   NEAR planes use (idn, nn) = (id, n - 8u|n|) * (1 - 8u)  -> a true LOWER bound when >= 0,
//   FAR  planes use (idf, nf) = (id, n + 8u|n|) * (1 + 8u)  -> a true UPPER bound when >= 0
// (8u leaves room for the two extra roundings of the scaled constants), so the per-box work is
// 6 FMAs and 6 min/max with no pad arithmetic.  Which plane of a box is the near one depends only
// on the sign of d: `sx/sy/sz` select the 16-byte vector to load (lo or hi of the four children).
struct FRay {
    float inx, iny, inz, nnx, nny, nnz;     // near-plane constants
    float ifx, ify, ifz, nfx, nfy, nfz;     // far-plane constants
    float dx, dy, dz;                       // d rounded to fp32 (triangle pre-test)
    uint32_t sx, sy, sz;                    // 1 when the component of d is negative
};
RT_D float clamp_recip(float d) {
    const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
    return 1.0f / a;
}
RT_D FRay make_fray(d3 o, d3 d) {
    FRay r;
    r.dx = __double2float_rn(d.x); r.dy = __double2float_rn(d.y); r.dz = __double2float_rn(d.z);
    const float idx = clamp_recip(r.dx), idy = clamp_recip(r.dy), idz = clamp_recip(r.dz);
    const float nx = -(__double2float_rn(o.x) * idx), ny = -(__double2float_rn(o.y) * idy), nz = -(__double2float_rn(o.z) * idz);
    const float lo = 1.0f - 8.0f * kU, hi = 1.0f + 8.0f * kU, pad = 8.0f * kU;
    r.inx = idx * lo; r.iny = idy * lo; r.inz = idz * lo;
    r.ifx = idx * hi; r.ify = idy * hi; r.ifz = idz * hi;
    r.nnx = (nx - pad * fabsf(nx)) * lo; r.nny = (ny - pad * fabsf(ny)) * lo; r.nnz = (nz - pad * fabsf(nz)) * lo;
    r.nfx = (nx + pad * fabsf(nx)) * hi; r.nfy = (ny + pad * fabsf(ny)) * hi; r.nfz = (nz + pad * fabsf(nz)) * hi;
    r.sx = idx < 0.0f; r.sy = idy < 0.0f; r.sz = idz < 0.0f;
    return r;
}
// One visit of a 4-wide node: conservative slab tests of its four child boxes against [0, tmax].
// key[k] = conservative entry distance of child k (a true lower bound) when the f64 box may be
// hit, +inf otherwise.  7 x 16-byte loads: 3 near-plane vectors, 3 far-plane vectors, the links.
struct Visit4 {
    float key[4];
    int32_t child[4];
};
// GENERIC = true loads through a generic pointer (the node may live in shared memory).
template <bool GENERIC = false>
RT_D Visit4 visit4(const FNode *node, const FRay &r, float tmax) {
    const float4 *p = reinterpret_cast<const float4 *>(node);
    // vectors 0..2 = lo.x lo.y lo.z, 3..5 = hi.x hi.y hi.z of the four children
    float4 ax, ay, az, bx, by, bz;
    int4 lk;
    if (GENERIC) {
        ax = p[0 + 3 * r.sx]; ay = p[1 + 3 * r.sy]; az = p[2 + 3 * r.sz];
        bx = p[3 - 3 * r.sx]; by = p[4 - 3 * r.sy]; bz = p[5 - 3 * r.sz];
        lk = reinterpret_cast<const int4 *>(p)[6];
    } else {
        ax = __ldg(p + 0 + 3 * r.sx); ay = __ldg(p + 1 + 3 * r.sy); az = __ldg(p + 2 + 3 * r.sz);
        bx = __ldg(p + 3 - 3 * r.sx); by = __ldg(p + 4 - 3 * r.sy); bz = __ldg(p + 5 - 3 * r.sz);
        lk = __ldg(reinterpret_cast<const int4 *>(p) + 6);
    }
    const float inf = __int_as_float(0x7F800000);
    Visit4 v;
#define RT_SLAB(K, C)                                                                                          \
    {                                                                                                          \
        const float tn = fmaxf(fmaxf(fmaf(ax.C, r.inx, r.nnx), fmaf(ay.C, r.iny, r.nny)),                      \
                               fmaxf(fmaf(az.C, r.inz, r.nnz), 0.0f));                                         \
        const float tf = fminf(fminf(fmaf(bx.C, r.ifx, r.nfx), fmaf(by.C, r.ify, r.nfy)),                      \
                               fminf(fmaf(bz.C, r.ifz, r.nfz), tmax));                                         \
        v.key[K] = tn <= tf ? tn : inf;                                                                        \
    }
    RT_SLAB(0, x) RT_SLAB(1, y) RT_SLAB(2, z) RT_SLAB(3, w)
#undef RT_SLAB
    v.child[0] = lk.x; v.child[1] = lk.y; v.child[2] = lk.z; v.child[3] = lk.w;
    return v;
}
// ascending sort of the four (key, child) pairs: 5 compare-exchanges
RT_D void sort4(Visit4 &v) {
#define RT_CE(A, B)                                                                     \
    {                                                                                   \
        const bool sw = v.key[B] < v.key[A];                                            \
        const float ka = v.key[A], kb = v.key[B];                                       \
        const int32_t ca = v.child[A], cb = v.child[B];                                 \
        v.key[A] = sw ? kb : ka; v.key[B] = sw ? ka : kb;                               \
        v.child[A] = sw ? cb : ca; v.child[B] = sw ? ca : cb;                           \
    }
    RT_CE(0, 1) RT_CE(2, 3) RT_CE(0, 2) RT_CE(1, 3) RT_CE(1, 2)
#undef RT_CE
}

struct F3 {                 // the ray direction rounded to fp32 (all the pre-test needs of FRay)
    float dx, dy, dz;
};
RT_D F3 f3_of(const FRay &r) { F3 f; f.dx = r.dx; f.dy = r.dy; f.dz = r.dz; return f; }
struct FTriRec {
    d3 a;
    float bax, bay, baz, cax, cay, caz;
};
RT_D FTriRec load_ftri(const FTri *t) {
    const double2 q0 = __ldg(reinterpret_cast<const double2 *>(t));
    const float4 q1 = __ldg(reinterpret_cast<const float4 *>(t) + 1);
    const float4 q2 = __ldg(reinterpret_cast<const float4 *>(t) + 2);
    FTriRec r;
    r.a = mk3(q0.x, q0.y, __hiloint2double(__float_as_int(q1.y), __float_as_int(q1.x)));
    r.bax = q1.z; r.bay = q1.w; r.baz = q2.x;
    r.cax = q2.y; r.cay = q2.z; r.caz = q2.w;
    return r;
}
// fp32 pre-test in determinant-scaled Moller-Trumbore form (u = U/det, v = V/det, t = T/det).
// Returns 0: the exact test (triangle.rs:65-70) certainly rejects, or the hit lies beyond tmax;
//         1: undecided (within the error bound of an edge / of det = 0);
//         2: the exact test certainly accepts.
// For 1 and 2, [tlo, thi] bounds the distance the exact test would report.
//
// RAY-RELATIVE: the test is evaluated from the point o + tref*d (formed in f64), where tref is
// the conservative entry distance of the leaf that holds the triangle.  U and V do not change
// when the origin slides along the ray, but w = origin - a shrinks from the ray length to the
// leaf size, and the rounding error of the fp32 triple products — proportional to |w| — shrinks
// with it: distant grazing triangles no longer pile up as undecided candidates.
// The reference's own determinant (triangle.rs:50, same operation order as tri_exact): used when
// fp32 cannot tell the sign of det.  Rays lying IN the plane of many coplanar triangles (a path
// on an emissive quad that light-samples another quad of the same plane) would otherwise turn
// every one of them into an undecided candidate; the reference rejects them with |det| < 1e-11.
RT_D bool det_below_threshold(const DTriI *tp, d3 d) {
    const double2 *p = reinterpret_cast<const double2 *>(tp);
    const double2 q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    const double q4 = __ldg(reinterpret_cast<const double *>(tp) + 8);
    const d3 c0 = mk3(q1.y, q2.x, q2.y), c1 = mk3(q3.x, q3.y, q4), c2 = neg(d);
    const double det = c0.x * (c1.y * c2.z - c2.y * c1.z) - c1.x * (c0.y * c2.z - c2.y * c0.z) +
                       c2.x * (c0.y * c1.z - c1.y * c0.z);
    return fabs(det) < kTriDetMin;
}
RT_D int tri_classify(const FTri *tp, const DTriI *xp, const F3 r, d3 o, d3 d, float tref, float dl1g, float wpad, float tmax, float &tlo, float &thi) {
    const FTriRec t = load_ftri(tp);
    const double tr = (double)tref;
    const float wx = __double2float_rn(fma(tr, d.x, o.x - t.a.x)), wy = __double2float_rn(fma(tr, d.y, o.y - t.a.y)),
                wz = __double2float_rn(fma(tr, d.z, o.z - t.a.z));
    const float ppx = fmaf(r.dy, t.caz, -(r.dz * t.cay)), ppy = fmaf(r.dz, t.cax, -(r.dx * t.caz)), ppz = fmaf(r.dx, t.cay, -(r.dy * t.cax));
    const float det = fmaf(t.bax, ppx, fmaf(t.bay, ppy, t.baz * ppz));
    const float U = fmaf(wx, ppx, fmaf(wy, ppy, wz * ppz));
    const float qx = fmaf(wy, t.baz, -(wz * t.bay)), qy = fmaf(wz, t.bax, -(wx * t.baz)), qz = fmaf(wx, t.bay, -(wy * t.bax));
    const float V = fmaf(r.dx, qx, fmaf(r.dy, qy, r.dz * qz));
    const float Tn = fmaf(t.cax, qx, fmaf(t.cay, qy, t.caz * qz));
    const float w1 = fabsf(wx) + fabsf(wy) + fabsf(wz) + wpad;      // + the f64 rounding of the subtraction (fast_wpad)
    const float lb = fabsf(t.bax) + fabsf(t.bay) + fabsf(t.baz), lc = fabsf(t.cax) + fabsf(t.cay) + fabsf(t.caz);
    const float eu = dl1g * w1 * lc, ev = dl1g * w1 * lb, et = kGamma * w1 * lb * lc, ed = dl1g * lb * lc;
    const float ad = fabsf(det);
    const float inf = __int_as_float(0x7F800000);
    if (!(ad > ed)) {                                           // sign of det unknown (or NaN)
        if (det_below_threshold(xp, d)) return 0;               // the exact test returns None at :51
        tlo = 0.0f; thi = inf;
        return 1;
    }
    const float s = det < 0.0f ? -1.0f : 1.0f;
    const float us = s * U, vs = s * V, ts = s * Tn;            // ts / ad = distance measured from tref
    if (us < -eu || vs < -ev || us + vs > ad + eu + ev + ed || ts < -et) return 0;
    const float trem = __fadd_ru(tmax, -tref);                  // remaining range, rounded up
    if (ts > fmaf(trem, ad + ed, et)) return 0;
    tlo = __fadd_rd(tref, fmaxf(0.0f, __fdividef(ts - et, ad + ed) * 0.9999995f));
    const float den = ad - ed;
    thi = den > 1e-30f ? __fadd_ru(tref, __fdividef(ts + et, den) * 1.0000005f) : inf;
    if (!(thi == thi)) thi = inf;
    // the exact test also rejects |det| < 1e-11 (triangle.rs:51): "certain" needs det well above it
    const bool certain = us >= eu && vs >= ev && us + vs <= ad - (eu + ev + ed) && ts >= et && den > 1e-9f;
    return certain ? 2 : 1;
}
RT_D float up_f32(double t) { return __double2float_ru(t); }

// ---- closest hit over the triangle hierarchy ---------------------------------------------------
// The loop below is fp32 only.  Candidates that survive the pre-test are DEFERRED to a short
// list and confirmed with the reference's f64 arithmetic after the loop, when the whole warp
// has left it: an exact test is ~400 f64 instructions and would otherwise run with one or two
// lanes active while the other lanes of the warp wait.  "Certain" fp32 hits still tighten tmax
// (with the upper bound of their distance), so pruning works as if the hit were confirmed.
constexpr int kCand = 6;

// ---- when the conservative answer may not be the reference's ----------------------------------
// FAST culling is conservative with respect to GEOMETRY: it reaches every triangle whose exact
// test (triangle.rs:65-80) accepts the ray.  The reference is not: it reaches a triangle only if
// its own f64 box tests (aabb.rs:51-78) let it through every ancestor node, and those tests are
// rounded too.  A ray that clips the corner of an ancestor box — aimed at a mesh vertex, say —
// can be rejected there by one ulp although the triangle test would accept it, and a child whose
// computed entry distance is not `<` the current best (bvh.rs:160-176) is skipped although it holds a
// hit that is nearer by less than the rounding error.  For a ray drawn from a continuous
// distribution this never happens; for a ray aimed at a vertex it happens every time.
// FAST mode therefore PROVES, per accepted hit, that the reference reaches it, and hands the
// ray to the exact traversal when it cannot:
//  * hit_fragile: the box of the hit triangle itself lies inside every ancestor box, so if the ray
//    passes through it with room to spare — every cross-axis gap `exit_j - entry_i` and every exit
//    distance larger than the worst rounding error of such a test — all ancestors pass as well;
//  * tie_margin: two accepted hits closer together than that error (exact ties included) make the
//    reference's strict `<` order-dependent.
// Both are division-free; E_k = kRobust * (bound + |o_k|) over-estimates the error of one slab
// distance times |d_k| by three orders of magnitude.
constexpr double kRobust = 1e-13;
RT_D bool hit_fragile(const DTriI *tp, d3 o, d3 d, double t, double bound) {
    const double2 *p = reinterpret_cast<const double2 *>(tp);
    const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
    const double q4 = __ldg(reinterpret_cast<const double *>(tp) + 8);
    const double ax[3] = {q0.x, q0.y, q1.x}, bx[3] = {q0.x + q1.y, q0.y + q2.x, q1.x + q2.y}, cx[3] = {q0.x + q3.x, q0.y + q3.y, q1.x + q4};
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    double N[3], F[3], D[3], E[3];          // |d_k| * entry distance, |d_k| * exit distance, |d_k|, error of those
    bool frag = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double lo = fmin(ax[k], fmin(bx[k], cx[k])), hi = fmax(ax[k], fmax(bx[k], cx[k]));
        E[k] = kRobust * (bound + fabs(oo[k]));
        D[k] = fabs(dd[k]);
        if (dd[k] > 0.0) { N[k] = lo - oo[k]; F[k] = hi - oo[k]; }
        else if (dd[k] < 0.0) { N[k] = oo[k] - hi; F[k] = oo[k] - lo; }
        else {                              // aabb.rs:53-57: the origin has to be inside the slab, exactly
            N[k] = 0.0; F[k] = 0.0;
            frag = frag || !(oo[k] - lo > E[k] && hi - oo[k] > E[k]) || !(dd[k] == 0.0);     // (NaN direction: fragile)
        }
        if (D[k] > 0.0) {
            frag = frag || !(F[k] > E[k]);                          // the box is left robustly after the origin
            frag = frag || !(t * D[k] - N[k] > -E[k]);              // and the hit does not come before the box
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            if (i != j && D[i] > 0.0 && D[j] > 0.0) frag = frag || !(F[j] * D[i] - N[i] * D[j] > E[i] * D[j] + E[j] * D[i]);
    return frag;
}
// distance below which two hits of one ray may be ordered differently by the reference
RT_D double tie_margin(d3 o, d3 d, double t, double bound) {
    double m = 0.0;
    if (d.x != 0.0) m = fmax(m, (bound + fabs(o.x)) / fabs(d.x));
    if (d.y != 0.0) m = fmax(m, (bound + fabs(o.y)) / fabs(d.y));
    if (d.z != 0.0) m = fmax(m, (bound + fabs(o.z)) / fabs(d.z));
    return kRobust * (2.0 * m + fabs(t));
}

struct FBest {
    bool has, frag;         // frag: the reference may not reach this hit (hit_fragile)
    double t, u, v;
    double t2;              // smallest distance among the OTHER accepted candidates (lower bound for skipped ones)
    uint32_t prim;
};
RT_D void fbest_clear(FBest &b) {
    b.has = false; b.frag = false; b.t = 0.0; b.u = 0.0; b.v = 0.0; b.prim = 0;
    b.t2 = __longlong_as_double(0x7FF0000000000000ll);
}
RT_D void fbest_take(FBest &best, double t, double u, double v, uint32_t prim, bool frag) {
    // equal distances: lowest primitive index, just to have a rule — a tie is always handed to the
    // exact traversal afterwards (fbest_needs_exact), which applies the reference's visit order
    if (!best.has || t < best.t || (t == best.t && prim < best.prim)) {
        if (best.has) best.t2 = fmin(best.t2, best.t);
        best.has = true; best.frag = frag; best.t = t; best.u = u; best.v = v; best.prim = prim;
    } else {
        best.t2 = fmin(best.t2, t);
    }
}
RT_D bool fbest_needs_exact(const FBest &b, d3 o, d3 d, double bound) {
    if (!b.has) return false;
    if (b.frag) return true;
    return b.t2 < __longlong_as_double(0x7FF0000000000000ll) && !(b.t2 - b.t > tie_margin(o, d, b.t2, bound));
}
RT_D void confirm(const DTriI *tri_i, d3 o, d3 d, uint32_t prim, FBest &best, float &tmax, double bound) {
    double t, u, v;
    if (!tri_exact(tri_i + prim, o, d, t, u, v)) return;
    fbest_take(best, t, u, v, prim, hit_fragile(tri_i + prim, o, d, t, bound));
    tmax = fminf(tmax, up_f32(best.t));
}
// A ray with a NaN component can never produce a hit: every comparison of the reference's box
// and triangle tests is false, so it walks the WHOLE hierarchy, "hits" the first triangle with
// t = NaN and then fails `t * |dir| <= max_dist` (intersections.rs:56) -> None.  Such rays do occur
// (a dielectric hit at normal incidence can give 1 - cos^2 < 0 by one ulp, raytrace.rs:77; the
// reference itself panics on the NaN probability it then passes to gen_bool).  FAST mode answers
// "miss" directly instead of testing a million triangles.
RT_D bool ray_is_nan(d3 o, d3 d) { return !(o.x == o.x && o.y == o.y && o.z == o.z && d.x == d.x && d.y == d.y && d.z == d.z); }

RT_D void closest_fast(const FBvh f, const DTriI *tri_i, d3 o, d3 d, float tmax, FBest &best) {
    if (f.root == kFastEmpty || ray_is_nan(o, d)) return;
    const FRay r = make_fray(o, d);
    const float dl1g = kGamma * (fabsf(r.dx) + fabsf(r.dy) + fabsf(r.dz));
    const float wpad = fast_wpad(o, f.bound);
    int32_t stk_n[kFastStack];
    float stk_t[kFastStack];
    uint32_t cand_p[kCand];
    float cand_t[kCand];
    int nc = 0, sp = 0;
    int32_t cur = f.root;
    float cur_t = 0.0f;                     // conservative entry distance of `cur`
    const float inf = __int_as_float(0x7F800000);
    // "while-while": all lanes first walk inner nodes until each holds a leaf (or is done), then
    // all lanes that hold a leaf run the triangle pre-tests together.  kFastEmpty doubles as the
    // "nothing left" marker (it is negative and never a valid leaf code of a non-empty tree).
    while (cur != kFastEmpty) {
        while (cur >= 0) {
            Visit4 v = visit4(f.nodes + cur, r, tmax);
            sort4(v);                               // nearest first
            if (v.key[3] < inf) { stk_n[sp] = v.child[3]; stk_t[sp] = v.key[3]; ++sp; }
            if (v.key[2] < inf) { stk_n[sp] = v.child[2]; stk_t[sp] = v.key[2]; ++sp; }
            if (v.key[1] < inf) { stk_n[sp] = v.child[1]; stk_t[sp] = v.key[1]; ++sp; }
            if (v.key[0] < inf) {
                cur = v.child[0]; cur_t = v.key[0];
            } else {
                cur = kFastEmpty;
                while (sp > 0) {                    // pop, skipping children the bound now excludes
                    --sp;
                    if (stk_t[sp] <= tmax) { cur = stk_n[sp]; cur_t = stk_t[sp]; break; }
                }
            }
        }
        if (cur == kFastEmpty) break;
        {
            const uint32_t code = (uint32_t)~cur, first = code >> 3, count = (code & 7u) + 1u;
            for (uint32_t i = first; i < first + count; ++i) {
                float tlo, thi;
                const int cls = tri_classify(f.tris + i, tri_i + i, f3_of(r), o, d, cur_t, dl1g, wpad, tmax, tlo, thi);
                if (cls == 0) continue;
                if (cls == 2) tmax = fminf(tmax, thi);
                if (nc == kCand) {                  // rare: drop entries the new bound excludes
                    int k = 0;
                    for (int j = 0; j < kCand; ++j) {
                        if (cand_t[j] <= tmax) { cand_p[k] = cand_p[j]; cand_t[k] = cand_t[j]; ++k; }
                        else best.t2 = fmin(best.t2, (double)cand_t[j]);
                    }
                    nc = k;
                }
                if (nc == kCand) { confirm(tri_i, o, d, i, best, tmax, (double)f.bound); continue; }   // very rare
                cand_p[nc] = i; cand_t[nc] = tlo; ++nc;
            }
        }
        cur = kFastEmpty;
        while (sp > 0) {
            --sp;
            if (stk_t[sp] <= tmax) { cur = stk_n[sp]; cur_t = stk_t[sp]; break; }
        }
    }
    for (int j = 0; j < nc; ++j) {                  // warp-converged f64 confirmation
        if (cand_t[j] <= tmax) confirm(tri_i, o, d, cand_p[j], best, tmax, (double)f.bound);
        else best.t2 = fmin(best.t2, (double)cand_t[j]);           // skipped: farther than the best by an fp32 ulp at least
    }
}

// ---- intersections::intersect in FAST mode ------------------------------------------------------
// Split in three so the wavefront scheduler can run the pieces in different kernels:
//   shapes_closest  planes + box / ellipsoid hierarchies on the exact f64 path (a handful of
//                   analytic primitives, src/intersections.rs:45-52);
//   closest_fast    the triangle hierarchy, bounded by the shapes' best distance: a triangle only
//                   replaces an earlier category when strictly closer (src/intersections.rs:79);
//   finish_hit      the merge, the max_dist check (:56) and the rotated / renormalised normals.
struct ShapeBest {
    bool has;
    double t;
    int kind;
    uint32_t prim;
};
RT_D ShapeBest shapes_closest(const DScene &sc, d3 o, d3 d, Cnt &c) {
    ShapeBest sb;
    sb.has = false; sb.t = 0.0; sb.kind = -1; sb.prim = 0;
    for (uint32_t i = 0; i < sc.n_planes; ++i) {
        double t; d3 n; bool in;
        if (!shape_closest(load_shape(sc.planes + i), o, d, t, n, in)) continue;
        if (!sb.has || t < sb.t) { sb.has = true; sb.t = t; sb.kind = 0; sb.prim = i; }
    }
    if (sc.boxes.n_prims) {
        const Best b = closest_exact<1>(sc.boxes, sc.box_prims, o, d, c);
        if (b.has && (!sb.has || b.t < sb.t)) { sb.has = true; sb.t = b.t; sb.kind = 1; sb.prim = b.prim; }
    }
    if (sc.ellipsoids.n_prims) {
        const Best e = closest_exact<2>(sc.ellipsoids, sc.ell_prims, o, d, c);
        if (e.has && (!sb.has || e.t < sb.t)) { sb.has = true; sb.t = e.t; sb.kind = 2; sb.prim = e.prim; }
    }
    return sb;
}
// fp32 bound for the triangle search; skip = the shapes' best is NaN and can never be replaced
RT_D float shapes_tmax(const ShapeBest &sb, bool &skip) {
    skip = sb.has && sb.t != sb.t;
    return (sb.has && !skip) ? up_f32(sb.t) : __int_as_float(0x7F800000);
}
// SHAPES = false compiles the analytic-primitive code out (glTF scenes have none): the wavefront
// shade kernel is instruction-cache bound, every kilobyte of cold code in it costs.
template <bool SHAPES = true>
RT_D SceneHit finish_hit(const DScene &sc, d3 o, d3 d, const ShapeBest &sb, const FBest &fb) {
    SceneHit h;
    h.kind = sb.kind; h.prim = sb.prim; h.material = 0; h.t = sb.t; h.inside = false;
    h.ng = h.ns = mk3(0.0, 0.0, 0.0);
    bool has = sb.has;
    if (!has) h.kind = -1;
    if (fb.has && (!has || fb.t < h.t)) { has = true; h.t = fb.t; h.kind = 3; h.prim = fb.prim; }
    if (!has) return h;
    if (!(h.t * len3(d) <= dinf())) { h.kind = -1; return h; }
    if (h.kind == 3) {
        d3 ng, ns;
        tri_finish(sc.tri_s + h.prim, d, fb.u, fb.v, ng, ns, h.inside);
        const quat ident = mkq_xyzw(0.0, 0.0, 0.0, 1.0);
        h.ng = unit3(qrot(ident, ng));
        h.ns = unit3(qrot(ident, ns));
        h.material = __ldg(&sc.tri_s[h.prim].material);         // same value as in the intersect record, on the line tri_finish just read
    } else if (SHAPES) {
        const DShape *sp = h.kind == 0 ? sc.planes + h.prim : (h.kind == 1 ? sc.box_prims + h.prim : sc.ell_prims + h.prim);
        finish_shape_hit(sp, o, d, h);
    }
    return h;
}
// the triangle hit wins against the shapes (strict <, intersections.rs:79) but the proof that the
// reference finds the same one failed: its own traversal decides (rare, kept out of line)
RT_D bool triangle_hit_needs_exact(const DScene &sc, d3 o, d3 d, const ShapeBest &sb, const FBest &fb) {
    return fb.has && (!sb.has || fb.t < sb.t) && fbest_needs_exact(fb, o, d, (double)sc.ftris.bound);
}
static __device__ __noinline__ SceneHit intersect_exact_cold(const DScene &sc, d3 o, d3 d) {
    Cnt scratch;
    return intersect_exact(sc, o, d, scratch);
}
RT_D SceneHit intersect_fast(const DScene &sc, d3 o, d3 d, Cnt &c) {
    if (sc.ftris.root != kFastEmpty && !fast_in_range(o, sc.ftris.bound) && !ray_is_nan(o, d)) { ++c.xf; return intersect_exact_cold(sc, o, d); }
    const ShapeBest sb = shapes_closest(sc, o, d, c);
    FBest fb;
    fbest_clear(fb);
    bool skip;
    const float tmax0 = shapes_tmax(sb, skip);
    if (!skip) closest_fast(sc.ftris, sc.tri_i, o, d, tmax0, fb);
    if (triangle_hit_needs_exact(sc, o, d, sb, fb)) { ++c.xf; return intersect_exact_cold(sc, o, d); }
    return finish_hit(sc, o, d, sb, fb);
}

// ---- Light::pdf in FAST mode ---------------------------------------------------------------------
// All hits, no pruning, children visited left then right: the reference's DFS order
// (src/bvh.rs:188-210), so the floating-point SUM is formed in the same order.
// returns true when the hit is fragile (the reference's all-hits walk may not reach it)
RT_D bool light_confirm(const DTriI *tri_i, const DTriS *tri_s, d3 o, d3 d, uint32_t i, double &impact, double bound) {
    double t, u, v;
    if (!tri_exact(tri_i + i, o, d, t, u, v)) return false;
    const bool frag = hit_fragile(tri_i + i, o, d, t, bound);
    const double2 *p = reinterpret_cast<const double2 *>(tri_s + i);
    const double2 q0 = __ldg(p), q1 = __ldg(p + 1);
    const double inv_area = __ldg(&tri_s[i].inv_area);
    // |d . (+-n)| is the same for the flipped normal (triangle.rs:73-77)
    impact += inv_area * (t * t / fabs(dot3(d, mk3(q0.x, q0.y, q1.x))));
    return frag;
}
// returns true when some hit was fragile: the caller redoes the sum with light_hits_exact
RT_D bool light_hits_fast(const FBvh f, const DTriI *tri_i, const DTriS *tri_s, d3 o, d3 d, double &impact) {
    if (f.root == kFastEmpty) return false;
    const FRay r = make_fray(o, d);
    const float dl1g = kGamma * (fabsf(r.dx) + fabsf(r.dy) + fabsf(r.dz));
    const float wpad = fast_wpad(o, f.bound);
    const float inf = __int_as_float(0x7F800000);
    int32_t stk[kFastStack];
    float stk_t[kFastStack];
    uint32_t cand[kCand];
    int sp = 0, nc = 0;
    bool frag = false;
    stk[sp] = f.root; stk_t[sp] = 0.0f; ++sp;
    while (sp > 0) {
        --sp;
        const int32_t cur = stk[sp];
        const float cur_t = stk_t[sp];
        if (cur >= 0) {
            const Visit4 v = visit4(f.nodes + cur, r, inf);     // children pushed right to left: popped left first
#pragma unroll
            for (int k = 3; k >= 0; --k)
                if (v.key[k] < inf) { stk[sp] = v.child[k]; stk_t[sp] = v.key[k]; ++sp; }
        } else {
            const uint32_t code = (uint32_t)~cur, first = code >> 3, count = (code & 7u) + 1u;
            for (uint32_t i = first; i < first + count; ++i) {
                float tlo, thi;
                if (tri_classify(f.tris + i, tri_i + i, f3_of(r), o, d, cur_t, dl1g, wpad, inf, tlo, thi) == 0) continue;
                if (nc == kCand) {                  // flush in visit order, then keep collecting
                    for (int j = 0; j < kCand; ++j) frag = light_confirm(tri_i, tri_s, o, d, cand[j], impact, (double)f.bound) || frag;
                    nc = 0;
                }
                cand[nc++] = i;
            }
        }
    }
    for (int j = 0; j < nc; ++j) frag = light_confirm(tri_i, tri_s, o, d, cand[j], impact, (double)f.bound) || frag;   // same order as visited
    return frag;
}
// analytic lights (boxes, then ellipsoids: intersect_lights order, intersections.rs:88-89)
RT_D double light_shapes_impact(const DScene &sc, d3 o, d3 dir, Cnt &c) {
    double impact = 0.0;
    light_hits_exact<1>(sc.lboxes, sc.lbox_prims, nullptr, o, dir, impact, c);
    light_hits_exact<2>(sc.lellipsoids, sc.lell_prims, nullptr, o, dir, impact, c);
    return impact;
}
static __device__ __noinline__ void light_tris_exact_cold(const DScene &sc, d3 o, d3 dir, double &impact) {
    Cnt scratch;
    light_hits_exact<3>(sc.ltris, sc.ltri_i, sc.ltri_s, o, dir, impact, scratch);
}
// the emissive triangles' part of Light::pdf, added to `impact` (which holds the analytic lights' part)
RT_D void light_tris_impact_fast(const DScene &sc, d3 o, d3 dir, double &impact, Cnt &c) {
    const double base = impact;
    if ((sc.fltris.root != kFastEmpty && !fast_in_range(o, sc.fltris.bound)) || light_hits_fast(sc.fltris, sc.ltri_i, sc.ltri_s, o, dir, impact)) {
        ++c.xf;
        impact = base;
        light_tris_exact_cold(sc, o, dir, impact);
    }
}
RT_D double light_pdf_fast(const DScene &sc, d3 o, d3 dir, Cnt &c) {
    ++c.lray;
    double impact = light_shapes_impact(sc, o, dir, c);
    light_tris_impact_fast(sc, o, dir, impact, c);
    return impact / (double)sc.n_lights;
}

struct FastPolicy {
    RT_D static SceneHit intersect(const DScene &sc, d3 o, d3 d, Cnt &c) { return intersect_fast(sc, o, d, c); }
    RT_D static double light_pdf_ray(const DScene &sc, d3 o, d3 dir, Cnt &c) { return light_pdf_fast(sc, o, dir, c); }
};

}  // namespace rt
