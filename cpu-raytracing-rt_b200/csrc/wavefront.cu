// wavefront.cu — the wavefront scheduler of generate_image (reference src/main.rs:94-103).
//
// A pool of SLOTS, each holding one path in flight (ray, throughput, RNG stream position, pending
// diffuse weight), and a queue of ITEMS: item j = sample (j / n_px) of the pixel with ordinal
// (j % n_px).  A slot whose path has ended takes the next item, whichever pixel it belongs to, so
// all slots stay busy until the last sample of the call has been started.  An iteration:
//
//   k_wf_confirm  the exact f64 triangle test of the candidates the previous trace found, one
//                 (ray, candidate) pair per thread, folded per ray; proves that the reference
//                 reaches the same hit or hands the ray to
//   k_wf_exact    (forked branch) the reference's own traversal for those few rays;
//   k_wf_shade    one thread per slot, f64: finishes the pending Light::pdf weight, runs one bounce
//                 of raytrace_impl (src/raytrace.rs:12-60); a finished sample's radiance is parked
//                 in the plane of its sample index and the slot claims the next item; slots with a
//                 ray are appended to the trace list with a warp ballot + prefix-sum (stream
//                 compaction of live paths), fresh camera rays at the front;
//   k_wf_plan / k_wf_sum  complete sample planes are added to the accumulation buffer in sample
//                 order — main.rs:100-103's order, so the sum is the megakernel's bit for bit;
//   k_wf_trace    persistent warps, fp32 only.  Each lane walks the culling hierarchy for one ray
//                 (closest-hit search and, for a pending diffuse bounce, the all-hits search of
//                 the light hierarchy along the same ray) and writes candidate lists.  A lane that
//                 finishes its ray does not wait for the slowest ray of its warp: when too few
//                 lanes remain busy the warp refills the idle ones from the trace list.
//
// In a megakernel every lane waits for the slowest of its 32 rays at every bounce (ncu: 7 of 32
// lanes active in the traversal loop); here the traversal loop stays full and the ~400-instruction
// f64 confirmation runs with most lanes of a warp at once.  Arithmetic per path is identical to
// the megakernel in FAST mode, so both schedulers produce bitwise the same accumulation buffer.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "../../include/rt_b200.h"
#include "fast.cuh"
#include "kernels.hpp"

namespace rt {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kFlagAlive = 1u, kFlagRngHave = 2u, kFlagPendLight = 4u;
constexpr uint32_t kFlagExactPending = 16u;   // handed to k_wf_exact in this iteration: not shaded, not traced; cleared by the next confirm pass
constexpr uint32_t kOverflowClosest = 0x10000u, kOverflowLight = 0x20000u;
constexpr uint32_t kNoShape = 0xFFFFFFFFu;

RT_D d3 ld_soa(const double *base, uint32_t n, uint32_t i) { return mk3(base[i], base[n + i], base[2u * n + i]); }
RT_D void st_soa(double *base, uint32_t n, uint32_t i, d3 v) { base[i] = v.x; base[n + i] = v.y; base[2u * n + i] = v.z; }

__device__ __forceinline__ void flush_counters_wf(const Cnt &c, DCounters *out) {
    constexpr int kN = 4;
    unsigned long long v[kN] = {c.seg, c.lray, c.paths, c.xf};
    unsigned long long *dst[kN] = {&out->segments, &out->light_pdf_rays, &out->paths, &out->exact_fallbacks};
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < kN; ++k) {
        unsigned long long s = v[k];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(kFull, s, off);
        if (lane == 0 && s) atomicAdd(dst[k], s);
    }
}

// A slot IS a pixel of the whole image (state arrays, RNG key and accumulation index all use it);
// a render call restricted to a region simply starts with a live list that holds the region's
// pixels only, so the other kernels do not know about regions at all.
// A live list is filled from BOTH ends: slots whose ray is a fresh camera ray go to the front,
// slots in the middle of a path to the back.  Consumers walk front then back, so the warps of
// the trace kernel get runs of neighbouring primary rays (which visit the same nodes) instead of
// a mixture of primary and bounced rays, and the shade kernel's warps see similar hits.
// n_live[2 * list] = entries at the front, n_live[2 * list + 1] = entries at the back.
struct LiveList {
    const uint32_t *p;
    uint32_t nf, n, last;
    RT_D uint32_t at(uint32_t i) const { return p[i < nf ? i : last - (i - nf)]; }
};
RT_D LiveList live_list(const WfBuf &b, int list) {
    LiveList l;
    l.p = b.live[list];
    l.nf = b.n_live[2 * list];
    l.n = l.nf + b.n_live[2 * list + 1];
    l.last = b.n - 1u;
    return l;
}

// ordinal i of a region (0 .. w*h-1) -> pixel index in the whole image.  The region is walked in
// 8x4 pixel tiles, one tile per 32 consecutive ordinals (partial tiles at the right / bottom edge
// shrink, the map stays a bijection): the camera rays a warp starts together — a run in the trace
// list — cover a compact patch of the image.
RT_D uint32_t region_tile_pixel(const DRegion &rg, uint32_t width, uint32_t i) {
    const uint32_t ty = i / (4u * rg.w), r = i - ty * 4u * rg.w, ht = min(4u, rg.h - 4u * ty);
    const uint32_t tx = r / (8u * ht), l = r - tx * 8u * ht, wt = min(8u, rg.w - 8u * tx);
    return (rg.y0 + 4u * ty + l / wt) * width + rg.x0 + 8u * tx + l % wt;
}
__global__ void k_wf_init(WfBuf b, WfCall cl, uint32_t sample_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_px = cl.n_px;
    if (i < cl.n_slots) {
        b.flags[i] = 0;
        b.sample[i] = 0;
        b.live[0][i] = i;                               // the static list of all slots (k_wf_confirm)
    }
    if (i == 0) {
        *b.call = cl;                                   // what the kernels of this call read instead of arguments
        b.n_live[0] = cl.n_slots; b.n_live[1] = 0; b.n_live[2] = 0; b.n_live[3] = 0;
        b.cursor[0] = 0; b.cursor[1] = 0;
        const unsigned long long total = (unsigned long long)n_px * sample_count, first = (unsigned long long)(b.ring - cl.reserve) * n_px;
        b.item[0] = 0; b.item[1] = first < total ? first : total; b.item[2] = total;
        b.plane[0] = 0; b.plane[1] = 0xFFFFFFFFu; b.plane[2] = b.ring; b.plane[3] = 0;
    }
}

// ---- confirm -----------------------------------------------------------------------------------
// The exact f64 test of every deferred candidate (the reference's Triangle::intersection, ~400
// instructions with its nine divisions) as a kernel of its own.
// Phase 1, one (ray, candidate) PAIR per thread: a ray has 0..6 scene candidates and 0..6 light
// candidates, so a loop per ray runs with a third of the lanes; here a block gathers the pairs of
// its 128 rays with a prefix sum and deals them out evenly.
// Phase 2, one ray per thread: fold the pair results into what the shade kernel needs — the
// nearest confirmed triangle (slot entry 0 of cand / res_t / res_u / res_v, ncand = 0 or 1) and the
// complete Light::pdf impact (lres entry 0: the analytic lights' part + the terms in list order =
// the reference's DFS order, bvh.rs:188-210) — and decide whether the proof that the reference
// finds the same hits holds (hit_fragile, tie_margin).  If not the slot goes to k_wf_exact.
#ifndef WF_CONFIRM_MINB
#define WF_CONFIRM_MINB 8
#endif
// pair results a block keeps in shared memory between the two phases (0.9 pairs per ray on C3: 128
// rays need ~115); the rays of a block whose pairs do not fit are answered by k_wf_exact
constexpr uint32_t kConfirmPairs = 448;
__global__ void __launch_bounds__(128, WF_CONFIRM_MINB) k_wf_confirm(DScene sc, WfBuf b, int list, DCounters *ctr) {
    __shared__ uint32_t s_slot[128];
    __shared__ uint16_t s_pair[128 * 2 * kWfCand];
    __shared__ uint32_t s_warp[4];
    __shared__ double s_ray[6][128];
    __shared__ double s_t[kConfirmPairs], s_u[kConfirmPairs], s_v[kConfirmPairs];   // phase 1 -> phase 2
    const uint32_t tid = threadIdx.x, i = blockIdx.x * 128 + tid, lane = tid & 31, wid = tid >> 5;
    const LiveList ll = live_list(b, list);
    const uint32_t N = b.n;
    uint32_t slot = 0, nc = 0, nl = 0, flags = 0, nc_all = 0;
    bool mine = false;                                  // this thread's slot has results to fold
    if (i < ll.n) {
        slot = ll.at(i);
        flags = b.flags[slot];
        if (flags & kFlagExactPending) {
            // k_wf_exact answered this slot during the last iteration (the graph joined since):
            // its results are final, the shade kernel takes them in this iteration
            b.flags[slot] = flags & ~kFlagExactPending;
        } else if (flags & kFlagAlive) {                // not alive: nothing was traced for this slot
            mine = true;
            nc_all = b.ncand[slot];
            nc = (nc_all & kOverflowClosest) ? 0u : (nc_all & 0xFFu);       // overflowed lists: the shade kernel redoes the search
            nl = ((nc_all & kOverflowLight) || !(flags & kFlagPendLight)) ? 0u : ((nc_all >> 8) & 0xFFu);
            if (nc + nl) {
                const d3 o = ld_soa(b.ro, N, slot), d = ld_soa(b.rd, N, slot);
                s_ray[0][tid] = o.x; s_ray[1][tid] = o.y; s_ray[2][tid] = o.z;
                s_ray[3][tid] = d.x; s_ray[4][tid] = d.y; s_ray[5][tid] = d.z;
            }
        }
    }
    s_slot[tid] = slot;
    // exclusive prefix sum of the pair counts over the block
    uint32_t incl = nc + nl;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(kFull, incl, off);
        if ((int)lane >= off) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t base = incl - (nc + nl), total = 0;
#pragma unroll
    for (uint32_t w = 0; w < 4; ++w) {
        const uint32_t v = s_warp[w];
        if (w < wid) base += v;
        total += v;
    }
    for (uint32_t k = 0; k < nl; ++k) s_pair[base + k] = (uint16_t)(tid | (k << 7) | 0x400u);
    for (uint32_t k = 0; k < nc; ++k) s_pair[base + nl + k] = (uint16_t)(tid | (k << 7));
    __syncthreads();
    // ---- phase 1: pairs.  res_u / lres: -1 = no hit, -2 = a hit the reference may not reach
    for (uint32_t e = tid; e < min(total, kConfirmPairs); e += 128) {
        const uint32_t pr = s_pair[e], owner = pr & 127u, k = (pr >> 7) & 7u;
        const bool light = (pr & 0x400u) != 0;
        const size_t at = (size_t)k * N + s_slot[owner];
        const d3 o = mk3(s_ray[0][owner], s_ray[1][owner], s_ray[2][owner]);
        const d3 d = mk3(s_ray[3][owner], s_ray[4][owner], s_ray[5][owner]);
        const uint32_t prim = light ? b.lcand[at] : b.cand[at];
        double t, u, v;
        const DTriI *tp = (light ? sc.ltri_i : sc.tri_i) + prim;
        const bool hit = tri_exact(tp, o, d, t, u, v);
        const bool frag = hit && hit_fragile(tp, o, d, t, (double)(light ? sc.fltris.bound : sc.ftris.bound));
        if (light) {
            double term = frag ? -2.0 : -1.0;
            if (hit && !frag) {                         // the term of light_confirm (ray_sampler.rs:135-137)
                const double2 *p = reinterpret_cast<const double2 *>(sc.ltri_s + prim);
                const double2 q0 = __ldg(p), q1 = __ldg(p + 1);
                const double inv_area = __ldg(&sc.ltri_s[prim].inv_area);
                term = inv_area * (t * t / fabs(dot3(d, mk3(q0.x, q0.y, q1.x))));
            }
            s_t[e] = term;
        } else {
            s_u[e] = hit ? (frag ? -2.0 : u) : -1.0;
            s_t[e] = t; s_v[e] = v;
        }
    }
    __syncthreads();
    // ---- phase 2: one ray per thread
    if (mine) {
        // a candidate list that overflowed (rare: more than kWfCand survivors of the fp32 pre-test)
        // is not worth a second fp32 search: the exact traversal answers the ray
        bool need_exact = (nc_all & (kOverflowClosest | kOverflowLight)) != 0 || base + nl + nc > kConfirmPairs;
        if ((flags & kFlagPendLight) && !need_exact) {
            double impact = b.pimp[slot];
            for (uint32_t k = 0; k < nl; ++k) {
                const double term = s_t[base + k];
                if (term == -2.0) need_exact = true;
                else if (term != -1.0) impact += term;
            }
            b.lres[slot] = impact;
        }
        if (!need_exact) {
            FBest fb;
            fbest_clear(fb);
            for (uint32_t k = 0; k < nc; ++k) {
                const uint32_t e = base + nl + k;
                const double u = s_u[e];
                if (u != -1.0) fbest_take(fb, s_t[e], u, s_v[e], b.cand[(size_t)k * N + slot], u == -2.0);
            }
            if (fb.has) {
                const d3 o = mk3(s_ray[0][tid], s_ray[1][tid], s_ray[2][tid]), d = mk3(s_ray[3][tid], s_ray[4][tid], s_ray[5][tid]);
                need_exact = need_exact || fbest_needs_exact(fb, o, d, (double)sc.ftris.bound);
                b.cand[slot] = fb.prim;
                b.res_t[slot] = fb.t; b.res_u[slot] = fb.u; b.res_v[slot] = fb.v;
            }
            b.ncand[slot] = fb.has ? 1u : 0u;
        }
        if (need_exact) {
            b.flags[slot] = flags | kFlagExactPending;
            b.xlist[atomicAdd(b.n_x, 1u)] = slot;
            atomicAdd(&ctr->exact_fallbacks, 1ull);     // rare
        }
    }
}

// ---- shade -------------------------------------------------------------------------------------
#ifndef WF_SHADE_MINB
#define WF_SHADE_MINB 6        // glTF variant: 80 registers, no spills to speak of; measured 4 / 5 / 6 / 7 / 8 -> 1240 / 1270 / 1282 / 1276 / 1275 Mseg/s
#endif
template <bool SHAPES>
__global__ void __launch_bounds__(128, SHAPES ? 4 : WF_SHADE_MINB) k_wf_shade(DScene sc, WfBuf b, DCounters *ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const WfCall cl = *b.call;                          // uniform: one broadcast load
    const DRegion rg = cl.rg;
    const uint32_t sample_begin = cl.sample_begin;
    const unsigned long long seed = cl.seed;
    const uint32_t N = b.n, n_px = cl.n_px;
    const int lane = threadIdx.x & 31;
    uint32_t *__restrict__ next_live = b.live[1];
    Cnt c;
    bool keep = false, fresh = false, in_range = i < cl.n_slots, want = false, alive = false, defer = false, pend_light = false;
    const uint32_t slot = i;                            // slots are walked in index order; which PIXEL a slot works on changes
    uint32_t smp = 0, pix = 0, ord = 0, left = 0;
    Rng rng;
    rng.init(seed, 0u, 0u);
    d3 o = mk3(0.0, 0.0, 0.0), d = o, T = o, L = o;
    Pending pd;
    pd.v = o; pd.cos_pdf = 0.0;
    if (in_range) {
        const uint32_t flags = b.flags[slot];
        smp = b.sample[slot];
        left = (flags >> 8) & 0xFFu;
        defer = (flags & kFlagExactPending) != 0;
        alive = (flags & kFlagAlive) != 0 && !defer;
        bool ended = false;
        if (alive) {
            pix = b.pix[slot];
            rng.init(seed, pix, sample_begin + smp);
            rng.block = b.rng_block[slot]; rng.s0 = b.rng_s0[slot]; rng.s1 = b.rng_s1[slot];
            rng.have = (flags & kFlagRngHave) != 0;
            o = ld_soa(b.ro, N, slot); d = ld_soa(b.rd, N, slot);
            T = ld_soa(b.T, N, slot); L = ld_soa(b.L, N, slot);
            const uint32_t nc_all = b.ncand[slot];
            if (flags & kFlagPendLight) {               // finish the previous diffuse bounce: Light::pdf of THIS ray
                const double impact = b.lres[slot];     // folded by k_wf_confirm or answered by k_wf_exact
                pd.v = ld_soa(b.pv, N, slot);
                pd.cos_pdf = b.pcos[slot];
                if (!apply_pending(pd, impact / (double)sc.n_lights, T)) ended = true;
            }
            if (!ended) {
                ShapeBest sb;
                const uint32_t sid = b.sh_id[slot];
                sb.has = sid != kNoShape; sb.kind = sb.has ? (int)(sid >> 28) : -1; sb.prim = sid & 0x0FFFFFFFu; sb.t = b.sh_t[slot];
                FBest fb;
                fbest_clear(fb);
                bool skip;
                const float tmax = shapes_tmax(sb, skip);
                (void)tmax;
                if (!skip && (nc_all & 0xFFu)) {        // the nearest confirmed triangle, folded by k_wf_confirm
                    fb.has = true; fb.prim = b.cand[slot];
                    fb.t = b.res_t[slot]; fb.u = b.res_u[slot]; fb.v = b.res_v[slot];
                }
                ++c.seg;
                const SceneHit h = finish_hit<SHAPES>(sc, o, d, sb, fb);
                const int r = shade_hit<SHAPES>(sc, h, o, d, T, L, rng, pd);
                --left;
                if (r == kPathEnd) {
                    ended = true;
                } else if (left == 0) {
                    // the reference evaluates Light::pdf before recursing into a depth-0 call
                    // that returns at once (raytrace.rs:27-32, :13): count it, nothing to trace
                    if (r == kPathPendingPdf) ++c.lray;
                    ended = true;
                } else {
                    pend_light = r == kPathPendingPdf;
                }
            }
            if (ended) {
                // The finished sample's radiance goes to the plane of its sample index; k_wf_sum
                // adds the planes to the accumulation buffer in sample order (main.rs:102), whatever
                // slot computed them and whenever.
                const uint32_t od = b.ord[slot];
                double *sp = b.side + (size_t)(smp % b.ring) * 3u * n_px + od;
                sp[0] = L.x; sp[n_px] = L.y; sp[2u * (size_t)n_px] = L.z;
                alive = false;
            }
        }
        // A slot without a path (k_wf_exact's slots excepted) takes the next item of the queue.
        want = !alive && !defer;
    }
    // ---- claim: item j = sample (j / n_px) of the pixel with ordinal (j % n_px) in the region's
    // tile order.  One compare-and-swap per warp hands out a run of consecutive items — the lanes
    // of a warp that start a path together get neighbouring pixels — and never passes the limit
    // (the planes of the ring that are still being filled).
    __syncwarp();
    const unsigned mw = __ballot_sync(kFull, want);
    if (mw) {
        const int leader = __ffs(mw) - 1;
        unsigned long long first = 0;
        uint32_t got = 0, s0 = 0, o0 = 0;
        if (lane == leader) {
            // One atomicAdd per warp, no retry loop (a compare-and-swap on this one address from
            // 65 000 warps at once does not scale).  The limit is checked on a plain read first, so
            // all the warps of one launch together can overshoot it by at most one item per slot =
            // one plane (n_slots <= n_px): k_wf_plan keeps two planes of the ring in reserve.
            const unsigned long long limit = b.item[1], total = b.item[2];
            const uint32_t cnt = (uint32_t)__popc(mw);
            if (*(volatile unsigned long long *)b.item < limit) {
                first = atomicAdd(b.item, (unsigned long long)cnt);
                got = first < total ? (uint32_t)min((unsigned long long)cnt, total - first) : 0u;
            }
            s0 = (uint32_t)(first / n_px);
            o0 = (uint32_t)(first - (unsigned long long)s0 * n_px);
        }
        got = __shfl_sync(kFull, got, leader);
        s0 = __shfl_sync(kFull, s0, leader);
        o0 = __shfl_sync(kFull, o0, leader);
        const uint32_t rank = __popc(mw & ((1u << lane) - 1u));
        if (want && rank < got) {
            smp = s0;
            unsigned long long oo = (unsigned long long)o0 + rank;
            while (oo >= n_px) { oo -= n_px; ++smp; }
            ord = (uint32_t)oo;
            pix = region_tile_pixel(rg, sc.width, ord);
            rng.init(seed, pix, sample_begin + smp);
            const double jx = rng.range(0.0, 1.0), jy = rng.range(0.0, 1.0);       // camera.rs:49-50
            camera_ray(sc.cam, pix % sc.width, pix / sc.width, jx, jy, o, d);
            ++c.paths;
            T = mk3(1.0, 1.0, 1.0);
            L = mk3(0.0, 0.0, 0.0);
            left = sc.ray_depth;
            pend_light = false;
            fresh = true;
            alive = true;
        }
    }
    if (in_range) {
        if (defer) {
            // k_wf_confirm could not prove that the reference finds the same hits for this ray
            // (hit_fragile / a tie): k_wf_exact is answering both searches with the reference's own
            // traversal while this kernel runs.  Nothing is shaded, stored or traced for the slot
            // in this iteration; it stays in the list and is picked up in the next one.
            keep = true;
        } else if (alive) {                             // hand the ray to the trace kernel
            if (SHAPES) {
                const ShapeBest sb = shapes_closest(sc, o, d, c);
                b.sh_id[slot] = sb.has ? (((uint32_t)sb.kind << 28) | sb.prim) : kNoShape;
                b.sh_t[slot] = sb.t;
            } else {
                b.sh_id[slot] = kNoShape;
                b.sh_t[slot] = 0.0;
            }
            if (pend_light) {
                ++c.lray;
                b.pimp[slot] = SHAPES ? light_shapes_impact(sc, o, d, c) : 0.0;
                st_soa(b.pv, N, slot, pd.v);
                b.pcos[slot] = pd.cos_pdf;
            }
            st_soa(b.ro, N, slot, o); st_soa(b.rd, N, slot, d);
            st_soa(b.T, N, slot, T); st_soa(b.L, N, slot, L);
            b.rng_block[slot] = rng.block; b.rng_s0[slot] = rng.s0; b.rng_s1[slot] = rng.s1;
            if (fresh) { b.sample[slot] = smp; b.pix[slot] = pix; b.ord[slot] = ord; }
            b.flags[slot] = kFlagAlive | (rng.have ? kFlagRngHave : 0u) | (pend_light ? kFlagPendLight : 0u) | (left << 8);
            keep = true;
        } else {
            b.flags[slot] = 0;
        }
    }
    // the oldest sample still in flight: every plane before it is complete (k_wf_plan)
    {
        uint32_t m = (in_range && (alive || defer)) ? smp : 0xFFFFFFFFu;
        for (int off = 16; off > 0; off >>= 1) m = min(m, __shfl_xor_sync(kFull, m, off));
        if (lane == 0 && m != 0xFFFFFFFFu) atomicMin(&b.plane[1], m);
    }
    // stream compaction of the live slots: ballot + prefix sum inside the warp, one 64-bit atomic
    // per warp reserves room at both ends of the trace list (front: fresh camera rays, back: the rest)
    const unsigned mf = __ballot_sync(kFull, keep && fresh), mb = __ballot_sync(kFull, keep && !fresh);
    if (mf | mb) {
        unsigned long long base = 0;
        if (lane == 0)
            base = atomicAdd(reinterpret_cast<unsigned long long *>(b.n_live + 2),
                             (unsigned long long)__popc(mf) | ((unsigned long long)__popc(mb) << 32));
        base = __shfl_sync(kFull, base, 0);
        const unsigned below = (1u << lane) - 1u;
        if (keep) {
            if (fresh) next_live[(uint32_t)base + __popc(mf & below)] = slot;
            else next_live[N - 1u - ((uint32_t)(base >> 32) + __popc(mb & below))] = slot;
        }
    }
    flush_counters_wf(c, ctr);
}

// ---- sum -----------------------------------------------------------------------------------------
// Slots are not tied to pixels: a slot takes whatever item is next, so all slots stay busy until
// the last sample of the call has been started, however unevenly path lengths are spread over the
// image.  What keeps the result the megakernel's bit for bit is the ORDER of the additions: every
// finished sample is parked in the plane of its sample index, and complete planes are added to the
// accumulation buffer one after the other, sample 0 first (main.rs:100-103).
// k_wf_plan (one thread, after the shade kernel): planes [summed, upto) are complete — fully
// claimed and older than the oldest sample still in flight — and the claim limit moves on.
__global__ void k_wf_plan(WfBuf b) {
    const uint32_t n_px = b.call->n_px, reserve = b.call->reserve;
    const uint32_t summed = b.plane[3];                 // the sum kernel of the last iteration has run
    b.plane[0] = summed;
    const unsigned long long total = b.item[2], next = min(b.item[0], total);
    const unsigned long long claimed = next / n_px;     // planes whose items have all been handed out
    const uint32_t oldest = b.plane[1];
    uint32_t upto = (uint32_t)min(claimed, (unsigned long long)oldest);
    // batches of four planes amortise the read-modify-write of the accumulation buffer; at the end
    // of the call (nothing left to claim) everything that is complete goes
    if (upto < summed + b.batch && next < total) upto = summed;
    if (upto < summed) upto = summed;
    b.plane[3] = upto;
    // b.reserve planes of the ring stay free: the claims of one launch can pass the limit by one item per slot
    const unsigned long long limit = ((unsigned long long)upto + b.ring - reserve) * n_px;
    b.item[1] = limit < total ? limit : total;
    b.plane[1] = 0xFFFFFFFFu;
}
__global__ void __launch_bounds__(256) k_wf_sum(WfBuf b, uint32_t width, double *__restrict__ accum) {
    const uint32_t from = b.plane[0], upto = b.plane[3];
    if (upto <= from) return;
    const DRegion rg = b.call->rg;
    const uint32_t n_px = b.call->n_px;
    for (uint32_t od = blockIdx.x * blockDim.x + threadIdx.x; od < n_px; od += gridDim.x * blockDim.x) {
        const size_t p = (size_t)region_tile_pixel(rg, width, od) * 3;
        double ax = accum[p], ay = accum[p + 1], az = accum[p + 2];
        for (uint32_t smp = from; smp < upto; ++smp) {
            const double *sp = b.side + (size_t)(smp % b.ring) * 3u * n_px + od;
            ax += sp[0]; ay += sp[n_px]; az += sp[2u * (size_t)n_px];
        }
        accum[p] = ax; accum[p + 1] = ay; accum[p + 2] = az;
    }
}

// ---- exact ---------------------------------------------------------------------------------------
// The slots k_wf_confirm could not vouch for (normally a few dozen per iteration: paths leaving a
// flat light, rays through a mesh vertex): both searches of the ray redone with the reference's
// own traversal (exact.cuh) and written back in the form k_wf_confirm leaves its results in — the
// nearest triangle in entry 0, the complete light impact.  The slot's flag stays set for the whole
// iteration (the next confirm pass clears it), so whether this kernel is early or late changes nothing.  One such traversal is a long serial chain, so this kernel runs on a forked
// branch of the graph, next to the shade and trace kernels, which skip these slots meanwhile.
__global__ void __launch_bounds__(128) k_wf_exact(DScene sc, WfBuf b) {
    const uint32_t n = *b.n_x, N = b.n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = b.xlist[i];
        const uint32_t flags = b.flags[slot];
        const d3 o = ld_soa(b.ro, N, slot), d = ld_soa(b.rd, N, slot);
        Cnt scratch;
        if (flags & kFlagPendLight) {
            double impact = b.pimp[slot];               // the analytic lights' part (shade kernel), then the triangles
            light_hits_exact<3>(sc.ltris, sc.ltri_i, sc.ltri_s, o, d, impact, scratch);
            b.lres[slot] = impact;
        }
        const Best bt = closest_exact<3>(sc.tris, sc.tri_i, o, d, scratch);
        if (bt.has) {
            double t, u, v;
            tri_exact(sc.tri_i + bt.prim, o, d, t, u, v);
            b.cand[slot] = bt.prim;
            b.res_t[slot] = t; b.res_u[slot] = u; b.res_v[slot] = v;
        }
        b.ncand[slot] = bt.has ? 1u : 0u;               // no overflow: nothing left for the shade kernel to search
    }
}

// ---- trace -------------------------------------------------------------------------------------
// Warp-synchronous state machine.  A lane with a ray is in one of two states: at an INNER node
// (one step = fetch a 128-byte 4-wide node, four slab tests, pick the nearest child, push the others) or at a LEAF with
// triangles left (one step = one fp32 pre-test).  Every iteration the warp votes and runs the
// step kind more lanes are waiting for, so the 32 lanes always execute the same code and a
// finished lane is refilled from the live list as soon as enough lanes are idle — no lane waits
// for the slowest ray of its warp.  The closest-hit search of the scene hierarchy and the all-hits
// search of the light hierarchy use the same two steps with per-lane node / triangle pointers.
#ifndef WF_SMEM_TOP
#define WF_SMEM_TOP 0      // measured on B200 (C3): 947 Mseg/s with staging vs 1015 without — L1 hits cost the same as shared memory
#endif
#ifndef WF_COUNT
#define WF_COUNT 0           // 1: count fp32 node visits / pre-tests / candidates / overflows (diagnostics)
#endif
#ifndef WF_VOTE_I
#define WF_VOTE_I 1            // the INNER step runs when WF_VOTE_I * (#lanes at a node) >= WF_VOTE_T * (#lanes at a leaf)
#define WF_VOTE_T 1
#endif
#ifndef WF_REFILL
#define WF_REFILL 16
#endif
#ifndef WF_TRACE_BLOCKS
#define WF_TRACE_BLOCKS 8
#endif
constexpr int kRefillLanes = WF_REFILL;  // refill once this many lanes are idle

__global__ void __launch_bounds__(128, WF_TRACE_BLOCKS) k_wf_trace(DScene sc, WfBuf b, int list, DCounters *ctr) {
    const LiveList ll = live_list(b, list);
    const uint32_t n = ll.n, N = b.n;
    const int lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7F800000);
#if WF_SMEM_TOP
    // shared-memory staging of the top levels of the scene hierarchy (every ray starts there):
    // nodes [0, top) are its first levels in breadth-first order, one contiguous copy per block
    __shared__ float4 s_top[kFastTopNodes * 8];
    const uint32_t top_n = sc.ftris.root >= 0 ? (sc.ftris.top < kFastTopNodes ? sc.ftris.top : kFastTopNodes) : 0u;
    for (uint32_t k = threadIdx.x; k < top_n * 8; k += blockDim.x) s_top[k] = __ldg(reinterpret_cast<const float4 *>(sc.ftris.nodes) + k);
    __syncthreads();
#endif
    bool has = false, drained = false, closest = true, want_light = false;
    uint32_t nc = 0, nl = 0, over = 0, tri_i = 0, tri_end = 0;
    // per-lane values that are written at refill and read only when a candidate is stored or the ray
    // is published live in shared memory, not in registers (the kernel runs at 64 registers, 8 blocks per SM)
    __shared__ uint32_t s_slot[128];
#define slot s_slot[threadIdx.x]
#if WF_COUNT
    uint32_t n_inner = 0, n_tri = 0, n_cand = 0, n_over = 0;
#define WF_INC(X) ++X;
#else
#define WF_INC(X)
#endif
    FRay r = {};
    // the f64 origin and direction are only needed by the triangle pre-test: they live in shared
    // memory (48 B per thread) instead of 12 registers, which buys one more resident block per SM
    __shared__ double s_ray[6][128];
    double *const my_ray = &s_ray[0][threadIdx.x];
    float tmax = inf, cur_t = 0.0f;     // cur_t: conservative entry distance of `cur`
    __shared__ float s_wpad[128];                    // fast_wpad of the lane's ray (only the triangle pre-test reads it)

    int32_t cur = kFastEmpty;
    int sp = 0;
    uint2 stk[kFastStack];                  // (child code, entry distance bits): one 8-byte local access per push / pop
// decoded unconditionally (three integer ops, meaningful only when `cur` is a leaf code): no branch
#define WF_ENTER_LEAF()                                                                     \
    {                                                                                       \
        const uint32_t code_ = (uint32_t)~cur;                                              \
        tri_i = code_ >> 3;                                                                 \
        tri_end = tri_i + (code_ & 7u) + 1u;                                                \
    }
#define WF_POP()                                                                            \
    {                                                                                       \
        cur = kFastEmpty;                                                                   \
        while (sp > 0) {                                                                    \
            --sp;                                                                           \
            const uint2 e_ = stk[sp];                                                       \
            if (__uint_as_float(e_.y) <= tmax) { cur = (int32_t)e_.x; cur_t = __uint_as_float(e_.y); break; } \
        }                                                                                   \
        WF_ENTER_LEAF()                                                                     \
    }
    for (;;) {
        // ---- vote: which step do more lanes wait for?  (two ballots are the whole loop control
        // while every lane is busy)
        bool wi = has && cur >= 0;
        bool wt = has && cur < 0 && cur != kFastEmpty;
        unsigned mi = __ballot_sync(kFull, wi), mt = __ballot_sync(kFull, wt);
        if ((mi | mt) != kFull) {
            // some lanes are finished or idle
            if (has && cur == kFastEmpty) {
                if (closest && want_light && sc.fltris.root != kFastEmpty) {       // start the light search
                    closest = false;
                    tmax = inf; sp = 0;
                    cur = sc.fltris.root; cur_t = 0.0f;
                    WF_ENTER_LEAF()
                } else {                                                            // publish, go idle
                    b.ncand[slot] = nc | (nl << 8) | over;
                    has = false;
                }
            }
            // refill idle lanes from the live list: ballot + prefix sum, one atomic per warp
            const unsigned idle = __ballot_sync(kFull, !has);
            if (!drained && __popc(idle) >= kRefillLanes) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(b.cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(kFull, base, leader);
                // the highest index handed out in this refill tells whether the list ran out
                if (base + (uint32_t)__popc(idle) >= n) drained = true;
                if (!has) {
                    const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < n) {
                        const uint32_t slot_new = ll.at(idx);
                        slot = slot_new;
                        const d3 o = ld_soa(b.ro, N, slot_new), dd = ld_soa(b.rd, N, slot_new);
                        my_ray[0] = o.x; my_ray[128] = o.y; my_ray[256] = o.z;
                        my_ray[384] = dd.x; my_ray[512] = dd.y; my_ray[640] = dd.z;
                        r = make_fray(o, dd);
                        const uint32_t sid = b.sh_id[slot_new];
                        const double st = b.sh_t[slot_new];
                        bool skip = (sid != kNoShape && st != st) || ray_is_nan(o, dd);
                        tmax = (sid != kNoShape && !skip) ? up_f32(st) : inf;
                        const uint32_t fl = b.flags[slot_new];
                        want_light = (fl & kFlagPendLight) != 0;
                        nc = 0; nl = 0; over = 0; sp = 0;
                        const float range_bound = fmaxf(sc.ftris.bound, sc.fltris.bound);
                        s_wpad[threadIdx.x] = fast_wpad(o, range_bound);
                        if (!skip && !fast_in_range(o, range_bound)) {
                            // outside the range of the fp32 arithmetic: both searches go to k_wf_exact
                            // (the same route as a candidate list that overflowed)
                            over = kOverflowClosest | kOverflowLight;
                            want_light = false;
                            skip = true;
                        }
                        closest = true;
                        cur = skip ? kFastEmpty : sc.ftris.root; cur_t = 0.0f;
                        WF_ENTER_LEAF()
                        has = !(fl & kFlagExactPending);        // k_wf_exact is answering this ray: not traced
                    }
                }
            }
            if (drained && !__any_sync(kFull, has)) break;
            wi = has && cur >= 0;
            wt = has && cur < 0 && cur != kFastEmpty;
            mi = __ballot_sync(kFull, wi);
            mt = __ballot_sync(kFull, wt);
        }
        const int ni = __popc(mi), nt = __popc(mt);
        if (WF_VOTE_I * ni >= WF_VOTE_T * nt) {
            if (wi) {                                           // INNER step: one 4-wide node
                WF_INC(n_inner)
#if WF_SMEM_TOP
                const FNode *np = (closest && (uint32_t)cur < top_n) ? reinterpret_cast<const FNode *>(s_top) + cur : (closest ? sc.ftris.nodes : sc.fltris.nodes) + cur;
                Visit4 v = visit4<true>(np, r, tmax);
#else
                Visit4 v = visit4((closest ? sc.ftris.nodes : sc.fltris.nodes) + cur, r, tmax);
#endif
                // Nearest hit child becomes `cur`; the other hit children are deferred in child
                // order.  The child index rides in the two low mantissa bits of the (positive) entry
                // distance — cleared bits round it DOWN, so it stays a lower bound — and an integer
                // min picks the nearest.  The light search (all hits) uses the index alone as the
                // key: children left to right, the reference's visit order (bvh.rs:188-210), which
                // fixes the order of the pdf sum.
                const uint32_t none = 0xFFFFFFFFu;
                const uint32_t cmask = closest ? ~3u : 0u;          // branch-free: selects, no divergent paths
                const uint32_t b0 = (__float_as_uint(v.key[0]) & cmask) | 0u, b1 = (__float_as_uint(v.key[1]) & cmask) | 1u;
                const uint32_t b2 = (__float_as_uint(v.key[2]) & cmask) | 2u, b3 = (__float_as_uint(v.key[3]) & cmask) | 3u;
                const uint32_t k0 = v.key[0] < inf ? b0 : none, k1 = v.key[1] < inf ? b1 : none;
                const uint32_t k2 = v.key[2] < inf ? b2 : none, k3 = v.key[3] < inf ? b3 : none;
                const uint32_t kmin = min(min(k0, k1), min(k2, k3));
                if (kmin != none) {
                    // deferred children are pushed right to left, so they pop left to right
                    if (k3 != none && k3 != kmin) { stk[sp] = make_uint2((uint32_t)v.child[3], k3 & ~3u); ++sp; }
                    if (k2 != none && k2 != kmin) { stk[sp] = make_uint2((uint32_t)v.child[2], k2 & ~3u); ++sp; }
                    if (k1 != none && k1 != kmin) { stk[sp] = make_uint2((uint32_t)v.child[1], k1 & ~3u); ++sp; }
                    if (k0 != none && k0 != kmin) { stk[sp] = make_uint2((uint32_t)v.child[0], k0 & ~3u); ++sp; }
                    const uint32_t sel = kmin & 3u;
                    const int32_t c01 = (sel & 1u) ? v.child[1] : v.child[0], c23 = (sel & 1u) ? v.child[3] : v.child[2];
                    cur = (sel & 2u) ? c23 : c01;
                    cur_t = __uint_as_float(kmin & ~3u);
                    WF_ENTER_LEAF()
                } else {
                    WF_POP()
                }
            }
        } else if (wt) {                                        // TRIANGLE step: the whole (<= 2 triangle) leaf
            const FTri *tris = closest ? sc.ftris.tris : sc.fltris.tris;
            const DTriI *xtris = closest ? sc.tri_i : sc.ltri_i;
            const d3 o = mk3(my_ray[0], my_ray[128], my_ray[256]), dd = mk3(my_ray[384], my_ray[512], my_ray[640]);
            const float wpad = s_wpad[threadIdx.x];
            // the fp32 direction and its error scale are recomputed from the f64 ray in shared memory:
            // three conversions per leaf are cheaper than four registers held through the INNER steps
            F3 rd;
            rd.dx = __double2float_rn(dd.x); rd.dy = __double2float_rn(dd.y); rd.dz = __double2float_rn(dd.z);
            const float dl1g = kGamma * (fabsf(rd.dx) + fabsf(rd.dy) + fabsf(rd.dz));
            do {
                WF_INC(n_tri)
                float tlo, thi;
                const int cls = tri_classify(tris + tri_i, xtris + tri_i, rd, o, dd, cur_t, dl1g, wpad, tmax, tlo, thi);
                if (cls != 0) {
                    WF_INC(n_cand)
                    if (closest) {
                        if (cls == 2) tmax = fminf(tmax, thi);
                        if (nc < (uint32_t)kWfCand) {
                            b.cand[nc * N + slot] = tri_i;
                            ++nc;
                        } else {
                            over |= kOverflowClosest;           // rare: the shade kernel redoes this search
                            WF_INC(n_over)
                        }
                    } else {
                        if (nl < (uint32_t)kWfCand) { b.lcand[nl * N + slot] = tri_i; ++nl; }
                        else { over |= kOverflowLight; WF_INC(n_over) }
                    }
                }
            } while (++tri_i != tri_end);
            WF_POP()
        }
    }
#undef slot
#undef WF_ENTER_LEAF
#undef WF_POP
#undef WF_INC
#if WF_COUNT
    // diagnostics: fp32 node visits / pre-tests / candidates / overflows of this launch
    unsigned long long v[4] = {n_inner, n_tri, n_cand, n_over};
    unsigned long long *dst[4] = {&ctr->aabb_tests, &ctr->prim_tests, &ctr->light_aabb_tests, &ctr->light_prim_tests};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long s = v[k];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(kFull, s, off);
        if (lane == 0 && s) atomicAdd(dst[k], s);
    }
#else
    (void)ctr;
#endif
}

// ---- host loop -----------------------------------------------------------------------------------
size_t wf_bytes_per_slot() { return 15 * 8 + 3 * 8 + 7 * 4 + kWfCand * 8 + 4 * 8 + 3 * 4 + 2 * 4; }      // + kWfRing * 24 per PIXEL for the side planes
int wf_trace_blocks(int device) {
    int sms = 148, per_sm = 4;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_trace, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    return sms * per_sm;
}
cudaError_t wf_runtime_init(WfRuntime &rt) {
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&rt.side, cudaStreamNonBlocking)) != cudaSuccess) return e;
    cudaEvent_t *evs[4] = {&rt.ev_fork, &rt.ev_shade, &rt.ev_join, &rt.ev_poll};
    for (cudaEvent_t *ev : evs)
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    uint32_t *host = nullptr;
    if ((e = cudaHostAlloc((void **)&host, 4 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return e;
    rt.host_live = host;
    return cudaSuccess;
}
static void wf_drop_graph(WfRuntime &rt) {
    if (rt.exec) cudaGraphExecDestroy(rt.exec);
    if (rt.graph) cudaGraphDestroy(rt.graph);
    rt.exec = nullptr; rt.graph = nullptr; rt.tried_graph = false;
}
void wf_runtime_destroy(WfRuntime &rt) {
    wf_drop_graph(rt);
    if (rt.side) { cudaStreamSynchronize(rt.side); cudaStreamDestroy(rt.side); }
    cudaEvent_t evs[4] = {rt.ev_fork, rt.ev_shade, rt.ev_join, rt.ev_poll};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    if (rt.host_live) cudaFreeHost((void *)rt.host_live);
    rt = WfRuntime{};
}

constexpr int kGraphIters = 4;              // iterations per graph launch
constexpr int kLaunchesPerIter = 6;         // confirm, exact, shade, plan, sum, trace

cudaError_t launch_render_wavefront(cudaStream_t st, WfRuntime &rt, const DScene &sc, uint32_t sample_begin, uint32_t sample_count,
                                    uint64_t seed, double *accum, DCounters *ctr, const WfBuf &b, int trace_blocks,
                                    uint64_t *launches, DRegion rg) {
    const uint32_t n_px = rg.w * rg.h;
    const unsigned long long total_items = (unsigned long long)n_px * sample_count;
    // Slots are independent of pixels: all of the buffer's capacity is used, but never more than
    // (ring - 4) planes' worth — the claims of one launch can pass the limit by one item per slot,
    // and that many planes of the ring are kept free for it (reserve).
    unsigned long long want_slots = wf_slots_for(n_px) < b.n ? wf_slots_for(n_px) : b.n;
    if (rt.slots_override >= 1024 && rt.slots_override < want_slots) want_slots = rt.slots_override;
    const unsigned long long most = (unsigned long long)(b.ring - 4u) * n_px;
    if (want_slots > most) want_slots = most;
    if (want_slots > total_items) want_slots = total_items;
    if (want_slots < 1) want_slots = 1;
    WfCall cl{};
    cl.sample_begin = sample_begin;
    cl.n_slots = (uint32_t)want_slots;
    cl.reserve = (uint32_t)((want_slots + n_px - 1) / n_px) + 1u;
    cl.n_px = n_px;
    cl.seed = seed;
    cl.rg = rg;
    // The grids cover the buffer's CAPACITY (threads past n_slots leave at once), so the launch
    // configuration does not depend on the call either.
    const unsigned blocks = (b.n + 127) / 128;
    k_wf_init<<<(cl.n_slots + 127) / 128, 128, 0, st>>>(b, cl, sample_count);
    ++*launches;
    // a segment takes one iteration, two when it is deferred to k_wf_exact; an item can wait for
    // a free plane of the ring: generous bound, reaching it is reported as an error
    const uint64_t max_iters = 4 * ((uint64_t)sample_count * n_px / cl.n_slots + 1) * (sc.ray_depth ? sc.ray_depth : 1) + 64;
    // scenes without planes / boxes / ellipsoids (every glTF scene) run the kernel compiled without them
    const bool shapes = sc.n_planes || sc.boxes.n_prims || sc.ellipsoids.n_prims || sc.lboxes.n_prims || sc.lellipsoids.n_prims;
    const unsigned sum_blocks = (unsigned)std::min<uint64_t>(((uint64_t)b.n_px_cap + 255) / 256, 148 * 8);
    // One iteration: confirm -> shade -> trace on the main stream; on the forked stream k_wf_exact
    // (after confirm) and the one-thread k_wf_plan + k_wf_sum (after shade, next to the trace
    // kernel), joined before the next iteration.  Four iterations are captured ONCE per (scene,
    // buffers) into a CUDA graph and replayed by every call: the inner loop is launch-bound on small
    // images and at the end of every render.
    auto issue_iteration = [&]() {
        cudaMemsetAsync(&b.n_live[2], 0, 4 * sizeof(uint32_t), st);         // the trace list starts empty; trace fetch position; deferred-slot count
        k_wf_confirm<<<blocks, 128, 0, st>>>(sc, b, 0, ctr);
        cudaEventRecord(rt.ev_fork, st);
        cudaStreamWaitEvent(rt.side, rt.ev_fork, 0);
        k_wf_exact<<<64, 128, 0, rt.side>>>(sc, b);
        if (shapes) k_wf_shade<true><<<blocks, 128, 0, st>>>(sc, b, ctr);
        else k_wf_shade<false><<<blocks, 128, 0, st>>>(sc, b, ctr);
        cudaEventRecord(rt.ev_shade, st);
        cudaStreamWaitEvent(rt.side, rt.ev_shade, 0);
        k_wf_plan<<<1, 1, 0, rt.side>>>(b);
        k_wf_sum<<<sum_blocks, 256, 0, rt.side>>>(b, sc.width, accum);
        cudaEventRecord(rt.ev_join, rt.side);
        k_wf_trace<<<trace_blocks, 128, 0, st>>>(sc, b, 1, ctr);
        cudaStreamWaitEvent(st, rt.ev_join, 0);
    };
    const bool same_key = rt.tried_graph && std::memcmp(&rt.key_sc, &sc, sizeof sc) == 0 && std::memcmp(&rt.key_b, &b, sizeof b) == 0 &&
                          rt.key_accum == accum && rt.key_ctr == ctr && rt.key_blocks == trace_blocks;
    if (!same_key) {
        wf_drop_graph(rt);
        std::memcpy(&rt.key_sc, &sc, sizeof sc);
        std::memcpy(&rt.key_b, &b, sizeof b);
        rt.key_accum = accum; rt.key_ctr = ctr; rt.key_blocks = trace_blocks;
        rt.tried_graph = true;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            for (int k = 0; k < kGraphIters; ++k) issue_iteration();
            if (cudaStreamEndCapture(st, &rt.graph) != cudaSuccess) rt.graph = nullptr;
            if (rt.graph && cudaGraphInstantiate(&rt.exec, rt.graph, 0) != cudaSuccess) rt.exec = nullptr;
        }
        if (rt.exec) ++rt.graph_builds;
        else cudaGetLastError();                    // plain launches from now on (same kernels, same order)
    }
    const bool use_graph = rt.exec != nullptr;
    volatile uint32_t *host_live = rt.host_live;
    bool poll_pending = false, finished = false;
    cudaError_t result = cudaSuccess;
    for (uint64_t it = 0; it < max_iters; it += kGraphIters) {
        if (use_graph) {
            result = cudaGraphLaunch(rt.exec, st);
            if (result != cudaSuccess) break;
        } else {
            for (int k = 0; k < kGraphIters; ++k) issue_iteration();
        }
        *launches += kGraphIters * kLaunchesPerIter;
        // poll the state after the PREVIOUS group of four without stalling the stream: nothing in
        // flight and nothing left to claim means the call is done
        if (poll_pending) {
            result = cudaEventSynchronize(rt.ev_poll);
            if (result != cudaSuccess) break;
            const unsigned long long next = (unsigned long long)host_live[2] | ((unsigned long long)host_live[3] << 32);
            if (host_live[0] + host_live[1] == 0 && next >= total_items) { finished = true; break; }
        }
        cudaMemcpyAsync((void *)host_live, &b.n_live[2], 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync((void *)(host_live + 2), b.item, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(rt.ev_poll, st);
        poll_pending = true;
    }
    if (result == cudaSuccess) {
        // the planes completed by the last iterations (k_wf_plan sees no path in flight and nothing
        // left to claim, so everything that remains goes)
        for (int k = 0; k < 2; ++k) {
            k_wf_plan<<<1, 1, 0, st>>>(b);
            k_wf_sum<<<sum_blocks, 256, 0, st>>>(b, sc.width, accum);
        }
        *launches += 4;
    }
    if (result != cudaSuccess) return result;
    if (!finished) return cudaErrorLaunchTimeout;   // iteration bound reached with work left: never a silent partial image
    return cudaGetLastError();
}

}  // namespace rt
