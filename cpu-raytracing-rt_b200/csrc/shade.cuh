// shade.cuh — the bounce loop: counter-based RNG with rand-0.8 distributions, the samplers of
// reference src/ray_sampler.rs, the material logic of reference src/raytrace.rs, the camera of
// reference src/camera.rs.  Geometry queries are injected (EXACT or FAST policy).
#pragma once
#include "exact.cuh"

namespace rt {

// ---- RNG -----------------------------------------------------------------------------------
// The reference uses rand 0.8 `thread_rng()` (src/main.rs:95): OS-seeded, not reproducible.
// Here every (pixel, global sample) owns a Philox4x32-10 stream keyed by the render seed, so a
// sample's path does not depend on which GPU, launch or thread computed it.  The transforms
// below reproduce rand 0.8.5's distributions for each call site.
// One Philox4x32-10 block of the stream (seed; pixel, sample): counter = (pixel, sample, block, 0).
// The stream of u64 draws is  #2k = (c1:c0) of block k,  #2k+1 = (c3:c2) of block k.
RT_D void philox_block(uint32_t k0, uint32_t k1, uint32_t pixel, uint32_t sample, uint32_t block,
                       unsigned long long &first, unsigned long long &second) {
    uint32_t c0 = pixel, c1 = sample, c2 = block, c3 = 0u, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ a; c1 = l1; c2 = h0 ^ c3 ^ b; c3 = l0;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    first = ((unsigned long long)c1 << 32) | c0;
    second = ((unsigned long long)c3 << 32) | c2;
}
// Stream position kept as (next block, spare half): one block per two draws, computed where the draw
// happens.  Used by the megakernel, whose lanes draw at different times anyway.
struct RngSeq {
    uint32_t k0, k1, pixel, sample, block;
    uint32_t s0, s1;        // spare u64 of the current block
    bool have;
    RT_D void init(unsigned long long seed, uint32_t px, uint32_t smp) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); pixel = px; sample = smp; block = 0; have = false; s0 = s1 = 0;
    }
    RT_D unsigned long long next_u64() {
        if (have) { have = false; return ((unsigned long long)s1 << 32) | s0; }
        unsigned long long first, second;
        philox_block(k0, k1, pixel, sample, block++, first, second);
        s0 = (uint32_t)second; s1 = (uint32_t)(second >> 32); have = true;
        return first;
    }
};
// The same stream addressed by the number of draws consumed (`pos`), with a WINDOW of draws computed
// ahead: the wavefront kernels fill the window once per bounce while the whole warp is converged
// (prefetch), instead of running the ten Philox rounds inside whichever divergent branch draws next
// (ncu, round 1: a fifth of the shade kernel's instructions were Philox rounds at 5 of 32 lanes).
// Draws beyond the window (rejection loops, the six draws of a box light) are computed in place.
// The per-slot state is `pos` alone (4 bytes instead of block + spare + flag).
struct RngWin {
    uint32_t k0, k1, pixel, sample, pos;
    uint32_t wbase, wn;                     // the window holds draws [wbase, wbase + wn), wbase even, wn <= 4
    unsigned long long w0, w1, w2, w3;
    RT_D void init(unsigned long long seed, uint32_t px, uint32_t smp, uint32_t at = 0u) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); pixel = px; sample = smp; pos = at; wbase = 0u; wn = 0u;
        w0 = w1 = w2 = w3 = 0ull;
    }
    RT_D void prefetch(int blocks) {        // 1 or 2 blocks starting at the one that holds draw `pos`
        wbase = pos & ~1u;
        philox_block(k0, k1, pixel, sample, wbase >> 1, w0, w1);
        wn = 2u;
        if (blocks > 1) { philox_block(k0, k1, pixel, sample, (wbase >> 1) + 1u, w2, w3); wn = 4u; }
        // the draws are MADE here, with the warp converged (the compiler must not sink the rounds into the
        // branches that consume them).  Four draws cover a cosine-sampled bounce; a light-sampled one takes four
        // or more — rand 0.8.5's gen_range(0..n) rejects half of its draws when n is a power of two (zone =
        // (n << lzcnt) - 1; 64 lights here) — and computes the rest in place.  A window of six (three blocks)
        // was measured: 1 497 against 1 504 Mseg/s, the kernel waits for memory, not for these instructions.
        asm volatile("" : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3));
    }
    RT_D unsigned long long next_u64() {
        const uint32_t i = pos - wbase;
        ++pos;
        if (i < wn) {
            const unsigned long long lo = (i & 1u) ? w1 : w0, hi = (i & 1u) ? w3 : w2;
            return (i & 2u) ? hi : lo;
        }
        unsigned long long first, second;
        philox_block(k0, k1, pixel, sample, (pos - 1u) >> 1, first, second);
        return ((pos - 1u) & 1u) ? second : first;
    }
};
// rand 0.8.5's distributions over either core
template <class CORE>
struct RngT : CORE {
    using CORE::next_u64;
    RT_D uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    RT_D double gen_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }      // Standard
    RT_D static double one_two(unsigned long long u) { return __longlong_as_double((long long)((u >> 12) | 0x3FF0000000000000ull)); }
    RT_D double range(double low, double high) {            // gen_range(low..high), UniformFloat::sample_single
        const double scale = high - low;
        for (int guard = 0; guard < 64; ++guard) {
            const double res = (one_two(next_u64()) - 1.0) * scale + low;
            if (res < high) return res;
        }
        return low;
    }
    RT_D double range_inclusive(double low, double high) {  // gen_range(low..=high), new_inclusive + sample
        const double max_rand = one_two(~0ull) - 1.0;
        double scale = (high - low) / max_rand;
        while (scale * max_rand + low > high) scale = __longlong_as_double(__double_as_longlong(scale) - 1);
        return (one_two(next_u64()) - 1.0) * scale + low;
    }
    RT_D unsigned long long range_usize(unsigned long long n) {     // gen_range(0..n), UniformInt<usize>
        const unsigned long long zone = (n << __clzll((long long)n)) - 1ull;
        for (;;) {
            const unsigned long long v = next_u64();
            if (v * n <= zone) return __umul64hi(v, n);
        }
    }
    RT_D int range_i32_0_1() {                              // gen_range(0..=1) on i32 (u32 draw, range 2)
        for (;;) {
            const uint32_t v = next_u32();
            if ((uint32_t)(v * 2u) <= 0x7FFFFFFFu) return (int)(v >> 31);
        }
    }
    RT_D bool gen_bool(double p) {                          // Bernoulli
        if (p == 1.0) return true;
        if (!(p >= 0.0 && p < 1.0)) p = 0.0;                // the reference panics here
        const double scaled = p * 18446744073709551616.0;
        const unsigned long long p_int = scaled >= 18446744073709551616.0 ? ~0ull : (unsigned long long)scaled;
        return next_u64() < p_int;
    }
};
using Rng = RngT<RngSeq>;
using RngW = RngT<RngWin>;

// ---- camera, reference src/camera.rs:48-55 + normalise (src/raytrace.rs:9) ---------------------
RT_D void camera_ray(const DCamera &cam, uint32_t px_x, uint32_t px_y, double jx, double jy, d3 &o, d3 &d) {
    const double px = (double)px_x + jx, py = (double)px_y + jy;
    const double x = (2.0 * px / cam.width - 1.0) * cam.tan_half_fov_x;
    const double y = -(2.0 * py / cam.height - 1.0) * cam.tan_half_fov_y;
    const d3 dir = add(add(scale_l(x, cam.right), scale_l(y, cam.up)), scale_l(1.0, cam.forward));
    o = cam.position;
    d = unit3(dir);
}

// ---- samplers, reference src/ray_sampler.rs ------------------------------------------------
template <class RNG>
RT_D d3 uniform_on_sphere(RNG &rng) {                       // :159-170 (cube-normalise: biased by design)
    const double a0 = rng.gen_f64(), a1 = rng.gen_f64(), a2 = rng.gen_f64();
    return unit3(mk3(a0 * 2.0 - 1.0, a1 * 2.0 - 1.0, a2 * 2.0 - 1.0));
}
template <class RNG>
RT_D d3 cosine_sample(d3 n, RNG &rng) {                     // :68-76
    const d3 dir = add(uniform_on_sphere(rng), n);
    const double eps = kEps * 16.0;
    if (fabs(dir.x) <= eps && fabs(dir.y) <= eps && fabs(dir.z) <= eps) return n;
    return unit3(dir);
}
RT_D double cosine_pdf(d3 n, d3 dir) {                      // :78-83
    const double c = dot3(n, dir);
    return c <= 0.0 ? 0.0 : c / kPi;
}
template <class RNG>
RT_D d3 uniform_on_box(d3 s, RNG &rng) {                    // :142-157
    const double w4x = s.y * s.z, w4y = s.x * s.z, w4z = s.x * s.y;
    const double choice = rng.range(0.0, w4x + w4y + w4z);
    const double sign = (double)(rng.range_i32_0_1() * 2 - 1);
    const double u1 = rng.range_inclusive(-1.0, 1.0);
    const double u2 = rng.range_inclusive(-1.0, 1.0);
    d3 p;
    if (choice < w4x) p = mk3(sign, u1, u2);
    else if (choice < w4x + w4y) p = mk3(u1, sign, u2);
    else p = mk3(u1, u2, sign);
    return mul3(p, s);
}
// SHAPES = false: the caller guarantees there are no box / ellipsoid lights (their code is dropped)
template <bool SHAPES = true, class RNG = Rng>
RT_D d3 light_sample(const DScene &sc, d3 pos, RNG &rng) {  // :101-130
    const uint32_t index = (uint32_t)rng.range_usize(sc.n_lights);
    const uint32_t nb = SHAPES ? sc.lboxes.n_prims : 0u, ne = SHAPES ? sc.lellipsoids.n_prims : 0u;
    d3 world;
    if (SHAPES && index < nb) {
        const ShapeRec l = load_shape(sc.lbox_prims + index);
        world = add(qrot(l.rot, uniform_on_box(l.param, rng)), l.position);
    } else if (SHAPES && index < nb + ne) {
        const ShapeRec l = load_shape(sc.lell_prims + (index - nb));
        world = add(qrot(l.rot, mul3(uniform_on_sphere(rng), l.param)), l.position);
    } else {
        const DTriI *t = sc.ltri_i + (index - nb - ne);
        double u = rng.range_inclusive(0.0, 1.0);
        double v = rng.range_inclusive(0.0, 1.0);
        if (u + v > 1.0) { u = 1.0 - u; v = 1.0 - v; }
        const double2 *p = reinterpret_cast<const double2 *>(t);
        const double2 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3);
        const double q4 = __ldg(reinterpret_cast<const double *>(t) + 8);
        const d3 a = mk3(q0.x, q0.y, q1.x), ba = mk3(q1.y, q2.x, q2.y), ca = mk3(q3.x, q3.y, q4);
        world = add(add(scale(ba, u), scale(ca, v)), a);
    }
    return unit3(sub(world, pos));
}

RT_D double powi(double a, int b) {                         // f64::powi (compiler-rt __powidf2)
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return r;
}

// One bounce of raytrace_impl (reference src/raytrace.rs:12-60) in throughput form: the
// reference returns e0 + w0*(e1 + w1*(...)) up its recursion; the loop carries T = w0*w1*... and
// adds T*e_k.  The bounce is split so that both schedulers run the SAME arithmetic:
//   shade_hit      emission + material + sampling of the next ray; for a diffuse hit in a scene
//                  with lights the weight needs Light::pdf of the NEW ray, which is left pending;
//   apply_pending  Mix::pdf (:95-97) and the throughput update (:33) once that pdf is known.
// The ray that Light::pdf shoots (pos + EPSILON*dir, dir — ray_sampler.rs:133) is exactly the
// next path segment (raytrace.rs:32), so one traversal serves both.
struct Pending {
    d3 v;               // dot * color / PI  (numerator of raytrace.rs:33)
    double cos_pdf;     // Cosine::pdf
};
enum { kPathEnd = 0, kPathContinue = 1, kPathPendingPdf = 2 };
template <bool SHAPES = true, class RNG = Rng>
RT_D int shade_hit(const DScene &sc, const SceneHit &h, d3 &o, d3 &d, d3 &T, d3 &L, RNG &rng, Pending &pd) {
    if (h.kind < 0) { L = add(L, mul3(T, sc.bg)); return kPathEnd; }         // :14
    const double2 *mp = reinterpret_cast<const double2 *>(sc.materials + h.material);
    const double2 m0 = __ldg(mp), m1 = __ldg(mp + 1), m2 = __ldg(mp + 2);
    const d3 color = mk3(m0.x, m0.y, m1.x), emission = mk3(m1.y, m2.x, m2.y);
    const double ior = __ldg(&sc.materials[h.material].ior);
    const int kind = __ldg(&sc.materials[h.material].kind);
    L = add(L, mul3(T, emission));                                          // :15
    const d3 pos = add(o, scale(d, h.t));                                   // position_at
    if (kind == 0) {                                                        // Diffuse :16-34
        const bool mix = sc.n_lights != 0;
        d3 dir;
        if (mix && !rng.gen_bool(0.5)) dir = light_sample<SHAPES>(sc, pos, rng);    // Mix::sample :87-93
        else dir = cosine_sample(h.ns, rng);
        const double cosv = dot3(dir, h.ns);
        if (cosv <= 0.0) return kPathEnd;                                   // :23-25
        pd.cos_pdf = cosine_pdf(h.ns, dir);
        pd.v = divs(scale_l(cosv, color), kPi);
        o = add(pos, scale_l(kEps, dir));                                   // :32 and ray_sampler.rs:133
        d = dir;
        if (mix) return kPathPendingPdf;
        if (pd.cos_pdf == 0.0) return kPathEnd;                             // :28-30
        T = mul3(T, divs(pd.v, pd.cos_pdf));                                // :33
        return kPathContinue;
    }
    const d3 refl = sub(d, scale(scale_l(2.0, h.ns), dot3(h.ns, d)));       // reflected_ray :67-73
    if (kind == 2) {                                                        // Dielectric :36-54
        double n1 = 1.0, n2 = ior;
        if (h.inside) { n1 = ior; n2 = 1.0; }
        const double eta = n1 / n2;
        const double cos1 = -dot3(h.ns, d);                                 // refracted_ray :75-88
        const double sin2 = eta * sqrt(1.0 - cos1 * cos1);
        if (!(sin2 > 1.0)) {
            const double r0 = powi((n1 - n2) / (n1 + n2), 2);               // reflection_power :62-65
            const double rp = r0 + (1.0 - r0) * powi(1.0 + dot3(d, h.ns), 5);
            const double p = rp < 0.0 ? 0.0 : (rp > 1.0 ? 1.0 : rp);
            if (!rng.gen_bool(p)) {                                         // :46
                const double cos2 = sqrt(1.0 - sin2 * sin2);
                const d3 dir = add(scale_l(eta, d), scale_l(eta * cos1 - cos2, h.ns));
                if (!h.inside) T = mul3(T, color);                          // :50
                o = add(pos, scale_l(kEps, dir));
                d = dir;
                return kPathContinue;
            }
        }
        o = add(pos, scale_l(kEps, refl));
        d = refl;
        return kPathContinue;
    }
    T = mul3(T, color);                                                     // Metallic :56-58
    o = add(pos, scale_l(kEps, refl));
    d = refl;
    return kPathContinue;
}
// light_pdf = Light::pdf of the ray (o, d) the path continues with.  false: the path ends (:28-30).
RT_D bool apply_pending(const Pending &pd, double light_pdf, d3 &T) {
    const double pdf = (pd.cos_pdf + light_pdf) / 2.0;                      // Mix::pdf :95-97
    if (pdf == 0.0) return false;
    T = mul3(T, divs(pd.v, pdf));                                           // :33
    return true;
}
template <class POLICY>
RT_D bool bounce(const DScene &sc, d3 &o, d3 &d, d3 &T, d3 &L, Rng &rng, Cnt &c) {
    ++c.seg;
    const SceneHit h = POLICY::intersect(sc, o, d, c);
    Pending pd;
    const int r = shade_hit(sc, h, o, d, T, L, rng, pd);
    if (r != kPathPendingPdf) return r == kPathContinue;
    return apply_pending(pd, POLICY::light_pdf_ray(sc, o, d, c), T);
}

template <bool DEEP>
struct ExactPolicyT {                   // DEEP: some hierarchy of the scene is deeper than the traversal stack (exact.cuh)
    RT_D static SceneHit intersect(const DScene &sc, d3 o, d3 d, Cnt &c) { return intersect_exact<DEEP>(sc, o, d, c); }
    RT_D static double light_pdf_ray(const DScene &sc, d3 o, d3 dir, Cnt &c) { return light_pdf_exact<DEEP>(sc, o, dir, c); }
};
using ExactPolicy = ExactPolicyT<false>;
using ExactDeepPolicy = ExactPolicyT<true>;

}  // namespace rt
