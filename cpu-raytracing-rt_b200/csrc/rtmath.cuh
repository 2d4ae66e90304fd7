// rtmath.cuh — f64 vector / quaternion arithmetic of the render path, usable from host C++ and
// from device code.  The reference does all of this through cgmath ^0.18 (reference
// Cargo.toml:7, src/types.rs:5-11); every helper keeps that crate's published evaluation order
// because EXACT mode must round like the reference (compile device code with -fmad=false and
// host code with -ffp-contract=off; IEEE `/` and `sqrt` are correctly rounded on both sides).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rt {

constexpr double kPi = 3.14159265358979323846264338327950288;       // src/types.rs:13
constexpr double kEps = 2.220446049250313e-16 * 512.0;               // src/types.rs:14 (f64::EPSILON*512)
constexpr double kTriDetMin = 1e-11;                                 // src/primitives/triangle.rs:51

struct d3 {
    double x, y, z;
};
RT_HD d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD d3 add(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD d3 sub(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD d3 neg(d3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD d3 scale(d3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }          // Vector3 * S
RT_HD d3 scale_l(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }        // S * Vector3
RT_HD d3 divs(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }           // Vector3 / S
RT_HD d3 mul3(d3 a, d3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }         // mul_element_wise
RT_HD d3 div3(d3 a, d3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }         // div_element_wise
RT_HD double dot3(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }      // InnerSpace::dot
RT_HD d3 cross3(d3 a, d3 b) {
    return mk3((a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x));
}
RT_HD double len3(d3 a) { return sqrt(dot3(a, a)); }                               // magnitude
RT_HD d3 unit3(d3 a) { return scale(a, 1.0 / len3(a)); }                           // normalize
RT_HD double comp(d3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct quat {        // cgmath Quaternion {s, v}
    double s;
    d3 v;
};
RT_HD quat mkq_xyzw(double x, double y, double z, double w) { quat q; q.s = w; q.v = mk3(x, y, z); return q; }
RT_HD quat qconj(quat q) { quat r; r.s = q.s; r.v = neg(q.v); return r; }
RT_HD d3 qrot(quat q, d3 v) {      // Quaternion * Vector3 (Rotation::rotate_vector)
    d3 tmp = add(cross3(q.v, v), scale(v, q.s));
    return add(scale(cross3(q.v, tmp), 2.0), v);
}

// reference src/aabb.rs:34-40 (NOT fmin/fmax: the comparison direction decides NaN behaviour)
RT_HD double rmin(double a, double b) { return a < b ? a : b; }
RT_HD double rmax(double a, double b) { return b < a ? a : b; }

}  // namespace rt
