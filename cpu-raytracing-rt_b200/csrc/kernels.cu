// kernels.cu — sm_100a kernels of the render hot path and their launchers.
//   k_trace_rays / k_trace_primary : intersections::intersect          (src/intersections.rs:42)
//   k_light_pdf                    : Light::pdf                        (src/ray_sampler.rs:132)
//   k_render_mega                  : the pixel loop of generate_image  (src/main.rs:94-103)
//   k_postprocess                  : /spp, ACES, gamma                 (src/main.rs:104)
#include <cuda_runtime.h>

#include "../../include/rt_b200.h"
#include "fast.cuh"
#include "kernels.hpp"

namespace rt {

// ---- helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ void flush_counters(const Cnt &c, DCounters *out) {
    unsigned long long v[8] = {c.seg, c.lray, c.paths, c.aabb, c.prim, c.laabb, c.lprim, c.xf};
    unsigned long long *dst[8] = {&out->segments, &out->light_pdf_rays, &out->paths, &out->aabb_tests,
                                  &out->prim_tests, &out->light_aabb_tests, &out->light_prim_tests, &out->exact_fallbacks};
    __syncwarp();                       // every lane of every warp reaches this call
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        unsigned long long s = v[k];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xFFFFFFFFu, s, off);
        if (lane == 0 && s) atomicAdd(dst[k], s);
    }
}

__device__ __forceinline__ void store_hit(rt_hit *dst, const SceneHit &h) {
    rt_hit o;
    o.t = h.kind < 0 ? 0.0 : h.t;
    o.geometry_normal[0] = h.ng.x; o.geometry_normal[1] = h.ng.y; o.geometry_normal[2] = h.ng.z;
    o.shading_normal[0] = h.ns.x; o.shading_normal[1] = h.ns.y; o.shading_normal[2] = h.ns.z;
    o.kind = h.kind; o.prim = h.kind < 0 ? 0u : h.prim; o.inside = h.inside ? 1 : 0; o.material = h.kind < 0 ? 0u : h.material;
    *dst = o;
}

// A warp owns an 8x4 pixel tile: neighbouring lanes trace neighbouring camera rays, so the top of
// the hierarchy is fetched once per warp (warp-coherent node fetch).
__device__ __forceinline__ bool tile_pixel(uint32_t width, uint32_t height, uint32_t &x, uint32_t &y) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_x = (width + 7) >> 3;
    x = (warp % tiles_x) * 8 + (lane & 7);
    y = (warp / tiles_x) * 4 + (lane >> 3);
    return x < width && y < height;
}
static inline uint32_t tile_grid_threads(uint32_t width, uint32_t height) {
    return ((width + 7) / 8) * ((height + 3) / 4) * 32;
}

// ---- kernels -----------------------------------------------------------------------------------
template <class POLICY>
__global__ void __launch_bounds__(128) k_trace_rays(DScene sc, unsigned long long n, const double *__restrict__ origins,
                                                    const double *__restrict__ dirs, rt_hit *__restrict__ hits, DCounters *ctr) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Cnt c;
    if (i < n) {
        const d3 o = mk3(origins[i * 3], origins[i * 3 + 1], origins[i * 3 + 2]);
        const d3 d = mk3(dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]);
        ++c.seg;
        store_hit(hits + i, POLICY::intersect(sc, o, d, c));
    }
    flush_counters(c, ctr);
}

template <class POLICY>
__global__ void __launch_bounds__(128) k_trace_primary(DScene sc, double jx, double jy, rt_hit *__restrict__ hits,
                                                       double *__restrict__ rays_out, DCounters *ctr) {
    uint32_t x, y;
    Cnt c;
    if (tile_pixel(sc.width, sc.height, x, y)) {
        d3 o, d;
        camera_ray(sc.cam, x, y, jx, jy, o, d);
        const size_t p = (size_t)y * sc.width + x;
        if (rays_out) {
            double *r = rays_out + p * 6;
            r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
        }
        ++c.seg;
        const SceneHit h = POLICY::intersect(sc, o, d, c);
        if (hits) store_hit(hits + p, h);
    }
    flush_counters(c, ctr);
}

template <class POLICY>
__global__ void __launch_bounds__(128) k_light_pdf(DScene sc, unsigned long long n, const double *__restrict__ pos,
                                                   const double *__restrict__ dir, double *__restrict__ pdf, DCounters *ctr) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    Cnt c;
    if (i < n) {
        const d3 p = mk3(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2]);
        const d3 d = mk3(dir[i * 3], dir[i * 3 + 1], dir[i * 3 + 2]);
        pdf[i] = sc.n_lights ? POLICY::light_pdf_ray(sc, add(p, scale_l(kEps, d)), d, c) : 0.0;       // ray_sampler.rs:133
    }
    flush_counters(c, ctr);
}

// Megakernel: one persistent thread per pixel walks its sample range.  The loop is FLAT — one
// iteration = (start the next path if the lane has none) + one bounce — so a lane whose path ended
// regenerates immediately instead of idling until the longest path of its warp finishes; lanes of
// a warp are then always busy with a segment (paths end after ~2 segments on average but may take
// ray_depth).  Samples of a pixel are still summed in index order, so the sum is deterministic and
// independent of how sample ranges are split across launches or GPUs.
template <class POLICY>
__global__ void __launch_bounds__(128) k_render_mega(DScene sc, uint32_t sample_begin, uint32_t sample_count,
                                                     unsigned long long seed, double *__restrict__ accum, DCounters *ctr, DRegion rg) {
    uint32_t x, y;
    Cnt c;
    if (tile_pixel(rg.w, rg.h, x, y)) {
        x += rg.x0; y += rg.y0;
        const uint32_t p = y * sc.width + x;
        d3 sum = mk3(accum[(size_t)p * 3], accum[(size_t)p * 3 + 1], accum[(size_t)p * 3 + 2]);
        Rng rng;
        d3 o = mk3(0.0, 0.0, 0.0), d = o, T = o, L = o;
        uint32_t s = 0, left = 0;
        bool alive = false;
        for (;;) {
            if (!alive) {
                if (s == sample_count) break;
                rng.init(seed, p, sample_begin + s);
                const double jx = rng.range(0.0, 1.0), jy = rng.range(0.0, 1.0);   // camera.rs:49-50
                camera_ray(sc.cam, x, y, jx, jy, o, d);
                ++c.paths;
                T = mk3(1.0, 1.0, 1.0);
                L = mk3(0.0, 0.0, 0.0);
                left = sc.ray_depth;                                                // raytrace.rs:9
                alive = true;
            }
            bool more = false;
            if (left > 0) {                                                         // raytrace.rs:13
                more = bounce<POLICY>(sc, o, d, T, L, rng, c);
                --left;
            }
            if (!more || left == 0) {
                sum = add(sum, L);                                                  // main.rs:102
                alive = false;
                ++s;
            }
        }
        accum[(size_t)p * 3] = sum.x; accum[(size_t)p * 3 + 1] = sum.y; accum[(size_t)p * 3 + 2] = sum.z;
    }
    flush_counters(c, ctr);
}

// correct_gamma(aces_tonemap(sum / samples)), reference src/postprocessing.rs:5-38
__global__ void k_postprocess(unsigned long long n, const double *__restrict__ accum, double samples, double *__restrict__ rgb) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = accum[i] / samples;
    double v = ((2.51 * x + 0.03) * x) / (((2.43 * x + 0.59) * x) + 0.14);
    v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
    rgb[i] = pow(v, 1.0 / 2.2);
}

__global__ void k_rng_probe(unsigned long long seed, uint32_t pixel, uint32_t sample, unsigned long long n_u64,
                            unsigned long long *u64_out, double *probe8) {
    if (threadIdx.x || blockIdx.x) return;
    Rng rng;
    rng.init(seed, pixel, sample);
    for (unsigned long long i = 0; i < n_u64; ++i) u64_out[i] = rng.next_u64();
    rng.init(seed, pixel, sample);
    probe8[0] = rng.range(0.0, 1.0);
    probe8[1] = rng.gen_f64();
    probe8[2] = rng.range_inclusive(-1.0, 1.0);
    probe8[3] = rng.range_inclusive(0.0, 1.0);
    probe8[4] = (double)rng.range_usize(7);
    probe8[5] = (double)rng.range_i32_0_1();
    probe8[6] = rng.gen_bool(0.25) ? 1.0 : 0.0;
    probe8[7] = rng.range(0.0, 3.5);
}

// ---- launchers ----------------------------------------------------------------------------------
static inline unsigned blocks_for(unsigned long long n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
// a hierarchy deeper than the traversal stack was uploaded with parent links: the stackless instantiations serve (EXACT only:
// FAST requests on such a scene are turned into EXACT ones by the C ABI)
static inline bool deep(const DScene &sc) {
    return sc.boxes.parent || sc.ellipsoids.parent || sc.tris.parent || sc.lboxes.parent || sc.lellipsoids.parent || sc.ltris.parent;
}

cudaError_t launch_trace_rays(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *o, const double *d,
                              void *hits, DCounters *ctr) {
    if (n == 0) return cudaSuccess;
    if (mode == RT_MODE_FAST) k_trace_rays<FastPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, o, d, (rt_hit *)hits, ctr);
    else if (deep(sc)) k_trace_rays<ExactDeepPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, o, d, (rt_hit *)hits, ctr);
    else k_trace_rays<ExactPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, o, d, (rt_hit *)hits, ctr);
    return cudaGetLastError();
}
cudaError_t launch_trace_primary(cudaStream_t st, const DScene &sc, int mode, double jx, double jy, void *hits,
                                 double *rays, DCounters *ctr) {
    const unsigned n = tile_grid_threads(sc.width, sc.height);
    if (mode == RT_MODE_FAST) k_trace_primary<FastPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, jx, jy, (rt_hit *)hits, rays, ctr);
    else if (deep(sc)) k_trace_primary<ExactDeepPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, jx, jy, (rt_hit *)hits, rays, ctr);
    else k_trace_primary<ExactPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, jx, jy, (rt_hit *)hits, rays, ctr);
    return cudaGetLastError();
}
cudaError_t launch_light_pdf(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *pos, const double *dir,
                             double *pdf, DCounters *ctr) {
    if (n == 0) return cudaSuccess;
    if (mode == RT_MODE_FAST) k_light_pdf<FastPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, pos, dir, pdf, ctr);
    else if (deep(sc)) k_light_pdf<ExactDeepPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, pos, dir, pdf, ctr);
    else k_light_pdf<ExactPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, n, pos, dir, pdf, ctr);
    return cudaGetLastError();
}
cudaError_t launch_render_mega(cudaStream_t st, const DScene &sc, int mode, uint32_t sample_begin, uint32_t sample_count,
                               uint64_t seed, double *accum, DCounters *ctr, DRegion rg) {
    const unsigned n = tile_grid_threads(rg.w, rg.h);
    if (mode == RT_MODE_FAST) k_render_mega<FastPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, sample_begin, sample_count, seed, accum, ctr, rg);
    else if (deep(sc)) k_render_mega<ExactDeepPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, sample_begin, sample_count, seed, accum, ctr, rg);
    else k_render_mega<ExactPolicy><<<blocks_for(n, 128), 128, 0, st>>>(sc, sample_begin, sample_count, seed, accum, ctr, rg);
    return cudaGetLastError();
}
cudaError_t launch_postprocess(cudaStream_t st, uint64_t n, const double *accum, uint32_t samples, double *rgb) {
    if (n == 0) return cudaSuccess;
    k_postprocess<<<blocks_for(n, 256), 256, 0, st>>>(n, accum, (double)samples, rgb);
    return cudaGetLastError();
}
cudaError_t launch_rng_probe(cudaStream_t st, uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n, unsigned long long *u64_out,
                             double *probe8) {
    k_rng_probe<<<1, 32, 0, st>>>(seed, pixel, sample, n, u64_out, probe8);
    return cudaGetLastError();
}

}  // namespace rt
