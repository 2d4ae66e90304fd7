// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// f64 vector / quaternion / matrix helpers restating the operation ORDER of cgmath ^0.18.0, the
// un-vendored crate the reference does its arithmetic with (reference Cargo.toml:7; its source is
// not under /root/reference and not in this container).  Orders restated from the published
// cgmath 0.18 sources:
//   dot        = (x*x' + y*y') + z*z'                 (mul_element_wise(..).sum(), sum = x+y+z)
//   cross      = (ay*bz - az*by, az*bx - ax*bz, ax*by - ay*bx)
//   magnitude  = sqrt(dot(v, v))
//   normalize  = v * (1.0 / magnitude)                (normalize_to(1): multiply by a reciprocal)
//   v / s      = component-wise IEEE division
//   q * v      = tmp = q.v x v + v*q.s ; (q.v x tmp)*2 + v
//   M3 * v     = (c0*v.x + c1*v.y) + c2*v.z            M4 * v = ((c0*x + c1*y) + c2*z) + c3*w
//   det(M3)    = m00(m11 m22 - m21 m12) - m10(m01 m22 - m21 m02) + m20(m01 m12 - m11 m02), m[c][r]
// Compile with -ffp-contract=off and without -ffast-math: rustc never contracts or reassociates.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

using F = double;                                       // reference src/types.rs:5
static constexpr F PI = 3.14159265358979323846264338327950288;          // src/types.rs:13
static constexpr F EPSILON = std::numeric_limits<F>::epsilon() * 512.0; // src/types.rs:14
static constexpr F INF = std::numeric_limits<F>::infinity();

struct V3 {
    F x, y, z;
    F operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 v3(F x, F y, F z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, F s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(F s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator/(V3 a, F s) { return {a.x / s, a.y / s, a.z / s}; }
static inline V3 mul_ew(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 div_ew(V3 a, V3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline F sum(V3 a) { return a.x + a.y + a.z; }
static inline F dot(V3 a, V3 b) { return sum(mul_ew(a, b)); }
static inline V3 cross(V3 a, V3 b) {
    return {(a.y * b.z) - (a.z * b.y), (a.z * b.x) - (a.x * b.z), (a.x * b.y) - (a.y * b.x)};
}
static inline F magnitude(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalize(V3 a) { return a * (1.0 / magnitude(a)); }
static inline bool is_zero(V3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }

struct Quat { F s; V3 v; };                             // cgmath Quaternion {s, v}
static inline Quat quat_xyzw(F x, F y, F z, F w) { return Quat{w, {x, y, z}}; }
static inline Quat conjugate(Quat q) { return Quat{q.s, -q.v}; }
static inline V3 rotate(Quat q, V3 v) {
    V3 tmp = cross(q.v, v) + (v * q.s);
    return (cross(q.v, tmp) * 2.0) + v;
}

struct M3 { V3 c[3]; };                                 // column-major like cgmath
static inline V3 mul(const M3 &m, V3 v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
static inline F det(const M3 &m) {
    return m.c[0].x * (m.c[1].y * m.c[2].z - m.c[2].y * m.c[1].z) -
           m.c[1].x * (m.c[0].y * m.c[2].z - m.c[2].y * m.c[0].z) +
           m.c[2].x * (m.c[0].y * m.c[1].z - m.c[1].y * m.c[0].z);
}

// f64::total_cmp key (used by the BVH midpoint comparator, src/bvh.rs:139)
static inline int64_t total_order_key(F v) {
    int64_t b;
    std::memcpy(&b, &v, 8);
    b ^= (int64_t)(((uint64_t)(b >> 63)) >> 1);
    return b;
}

// ---- counter-based RNG shared BY SPECIFICATION (not by code) with the CUDA kernels -----------
// The reference draws from rand 0.8 `thread_rng()` (OS-seeded ChaCha, src/main.rs:95): there is
// no reproducible stream to match, only distributions.  Both the oracle and the device use
// Philox4x32-10 keyed by `seed` with counter (pixel, global sample, block, 0); each block yields
// two u64 draws consumed in order.  The distribution transforms restate rand 0.8.5:
//   gen::<f64>()            (u64 >> 11) * 2^-53
//   gen_range(a..b)   f64   ((u64 >> 12 | 1.0 bits) - 1.0) * (b - a) + a, redraw while >= b
//   gen_range(a..=b)  f64   same with scale (b - a) / (1 - 2^-52), lowered until scale*max+a <= b
//   gen_range(0..n)   usize widening multiply, zone = (n << lzcnt(n)) - 1, redraw while lo > zone
//   gen_range(0..=1)  i32   the same on a u32 draw (range 2)
//   gen_bool(p)             p == 1 -> true without a draw, else u64 < (p * 2^64) as u64
struct Rng {
    uint32_t k0, k1, pixel, sample, block;
    uint64_t spare;
    bool have;
    Rng(uint64_t seed, uint32_t pixel_, uint32_t sample_)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), pixel(pixel_), sample(sample_), block(0),
          spare(0), have(false) {}

    static inline void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
            uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
            uint32_t n1 = (uint32_t)p1;
            uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
            uint32_t n3 = (uint32_t)p0;
            c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
    }
    uint64_t next_u64() {
        if (have) { have = false; return spare; }
        uint32_t c[4] = {pixel, sample, block++, 0u};
        philox(c, k0, k1);
        spare = ((uint64_t)c[3] << 32) | c[2];
        have = true;
        return ((uint64_t)c[1] << 32) | c[0];
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    F gen_f64() { return (F)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    static F bits_1_2(uint64_t u) {
        uint64_t b = (u >> 12) | 0x3FF0000000000000ull;
        F f;
        std::memcpy(&f, &b, 8);
        return f;
    }
    F range(F low, F high) {                        // gen_range(low..high)
        F scale = high - low;
        for (int guard = 0; guard < 64; ++guard) {
            F res = (bits_1_2(next_u64()) - 1.0) * scale + low;
            if (res < high) return res;
        }
        return low;
    }
    static F inclusive_scale(F low, F high) {       // UniformFloat::new_inclusive
        const F max_rand = bits_1_2(~0ull) - 1.0;
        F scale = (high - low) / max_rand;
        while (scale * max_rand + low > high) {
            uint64_t b;
            std::memcpy(&b, &scale, 8);
            b -= 1;
            std::memcpy(&scale, &b, 8);
        }
        return scale;
    }
    F range_inclusive(F low, F high) {              // gen_range(low..=high)
        F scale = inclusive_scale(low, high);
        return (bits_1_2(next_u64()) - 1.0) * scale + low;
    }
    uint64_t range_usize(uint64_t n) {              // gen_range(0..n), n >= 1
        uint64_t zone = (n << __builtin_clzll(n)) - 1;
        for (;;) {
            uint64_t v = next_u64();
            unsigned __int128 m = (unsigned __int128)v * n;
            if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
        }
    }
    int range_i32_0_1() {                           // gen_range(0..=1) on i32: u32 draw, range 2
        const uint32_t zone = (2u << 30) - 1;
        for (;;) {
            uint32_t v = next_u32();
            uint64_t m = (uint64_t)v * 2u;
            if ((uint32_t)m <= zone) return (int)(m >> 32);
        }
    }
    bool gen_bool(F p) {                            // Bernoulli::new(p).sample
        if (p == 1.0) return true;
        if (!(p >= 0.0 && p < 1.0)) p = 0.0;        // the reference panics here (NaN / out of range)
        F scaled = p * 18446744073709551616.0;
        uint64_t p_int = scaled >= 18446744073709551616.0 ? ~0ull : (uint64_t)scaled;
        return next_u64() < p_int;
    }
};

}  // namespace orc
