// wavefront.cu — the wavefront scheduler of generate_image (reference src/main.rs:94-103).
//
// A pool of SLOTS, each holding one path in flight (ray, throughput, radiance, RNG stream position:
// records of whole sectors grouped by the kernels that touch them, kernels.hpp), and a queue of ITEMS: item j = sample (j / n_px)
// of the pixel with ordinal (j % n_px).  A slot whose path has ended takes the next item, whichever
// pixel it belongs to, so all slots stay busy until the last sample of the call has been started.
// One iteration — every kernel after the first walks a COMPACTED LIST, so its warps are full of
// lanes that all do the same thing:
//
//   k_wf_confirm    one (ray, candidate) pair per thread: the exact f64 triangle test of the
//                   candidates the previous trace found, folded per ray; proves that the reference
//                   reaches the same hit or hands the ray to k_wf_exact.  Then sorts every slot into
//                   the HIT list (its ray hit something), the REGEN list (its ray escaped, or the slot
//                   is empty) or neither (deferred to k_wf_exact);
//   k_wf_exact      (forked branch) the reference's own traversal for the few deferred rays;
//   k_wf_shade_hit  hit list, f64: finishes the pending Light::pdf weight and runs one bounce of
//                   raytrace_impl (src/raytrace.rs:12-60); the RNG draws of the bounce are computed
//                   ahead while the warp is converged.  Paths that go on are appended to the back of
//                   the trace list, paths that end to the regen list;
//   k_wf_regen      regen lists: adds the background to escaped paths, parks the finished sample's
//                   radiance in the plane of its sample index, takes the next item and starts its
//                   camera ray (front of the trace list, runs of neighbouring pixels).  Runs twice: for
//                   the slots k_wf_confirm found (on a forked branch, NEXT TO k_wf_shade_hit — both
//                   kernels wait for memory most of the time) and then for the paths the shade kernel
//                   ended.  Items are handed out by position in the list: no atomic operation;
//   k_wf_plan / k_wf_sum  complete sample planes are added to the accumulation buffer in sample
//                   order — main.rs:100-103's order, so the sum is the megakernel's bit for bit;
//   k_wf_trace      persistent warps, fp32 only.  Each lane walks the culling hierarchy for one ray
//                   (closest-hit search and, for a pending diffuse bounce, the all-hits search of
//                   the light hierarchy along the same ray) and writes candidate lists.  A lane that
//                   finishes its ray does not wait for the slowest ray of its warp: when too few
//                   lanes remain busy the warp refills the idle ones from the trace list.
//
// Round 1 ran confirm's successor as ONE kernel over all slots in static order: a warp held hits next
// to misses next to empty slots and executed each of those code paths with a third of its lanes
// (ncu: 12.4 of 32 lanes per instruction, a fifth of all instructions Philox rounds at 5 lanes).
// Arithmetic per path is identical to the megakernel in FAST mode, so both schedulers produce
// bitwise the same accumulation buffer, whatever slot or kernel a sample passed through.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

#include "../../include/rt_b200.h"
#include "fast.cuh"
#include "kernels.hpp"

namespace rt {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kFlagAlive = 1u;           // the slot holds a path whose ray is being / was traced
constexpr uint32_t kFlagEnded = 2u;           // k_wf_shade_hit ended the path: L is final, k_wf_regen parks it
constexpr uint32_t kFlagPendLight = 4u;       // the weight of the last diffuse bounce waits for Light::pdf of the ray in flight
constexpr uint32_t kFlagExactPending = 16u;   // handed to k_wf_exact in this iteration; cleared by the next confirm pass
constexpr uint32_t kOverflowClosest = 0x10000u, kOverflowLight = 0x20000u;
constexpr uint32_t kNoShape = 0xFFFFFFFFu;
enum { kCntFront = 0, kCntBack = 1, kCntCursor = 2, kCntExact = 3, kCntHit = 4, kCntRegen = 5 /* and 6: the two regen lists */, kCntWords = 8 };

// ---- slot records: 16-byte vector access; stores always cover whole 32-byte sectors ----------------
// WfHead as uint4[2]: (flags, sh_id, ncand, hit_prim) (sample, pix, ord, rng_pos)
RT_D uint4 ld_head(const WfHead *h) { return reinterpret_cast<const uint4 *>(h)[0]; }
RT_D uint4 ld_item(const WfHead *h) { return reinterpret_cast<const uint4 *>(h)[1]; }
RT_D void st_head_item(WfHead *h, uint4 head, uint4 item) {
    reinterpret_cast<uint4 *>(h)[0] = head;
    reinterpret_cast<uint4 *>(h)[1] = item;
}
// WfRay as double2[4]: (o.x o.y) (o.z d.x) (d.y d.z) (spare)
RT_D void ld_ray(const WfRay *s, d3 &o, d3 &d) {
    const double2 *p = reinterpret_cast<const double2 *>(s);
    const double2 q0 = p[0], q1 = p[1], q2 = p[2];
    o = mk3(q0.x, q0.y, q1.x);
    d = mk3(q1.y, q2.x, q2.y);
}
RT_D void st_ray(WfRay *s, d3 o, d3 d) {
    double2 *p = reinterpret_cast<double2 *>(s);
    p[0] = make_double2(o.x, o.y); p[1] = make_double2(o.z, d.x); p[2] = make_double2(d.y, d.z); p[3] = make_double2(0.0, 0.0);
}
// WfTL as double2[4]: (T.x T.y) (T.z L.x) (L.y L.z) (spare)
RT_D void ld_TL(const WfTL *s, d3 &T, d3 &L) {
    const double2 *p = reinterpret_cast<const double2 *>(s);
    const double2 q0 = p[0], q1 = p[1], q2 = p[2];
    T = mk3(q0.x, q0.y, q1.x);
    L = mk3(q1.y, q2.x, q2.y);
}
RT_D void st_TL(WfTL *s, d3 T, d3 L) {
    double2 *p = reinterpret_cast<double2 *>(s);
    p[0] = make_double2(T.x, T.y); p[1] = make_double2(T.z, L.x); p[2] = make_double2(L.y, L.z); p[3] = make_double2(0.0, 0.0);
}
// WfPend as double2[4]: (pv.x pv.y) (pv.z pcos) | (pimp lres) (sh_t spare)
RT_D void ld_pending(const WfPend *a, Pending &pd, double &lres) {
    const double2 *p = reinterpret_cast<const double2 *>(a);
    const double2 q0 = p[0], q1 = p[1];
    pd.v = mk3(q0.x, q0.y, q1.x);
    pd.cos_pdf = q1.y;
    lres = a->lres;
}
RT_D void st_pending(WfPend *a, const Pending &pd) {
    double2 *p = reinterpret_cast<double2 *>(a);
    p[0] = make_double2(pd.v.x, pd.v.y); p[1] = make_double2(pd.v.z, pd.cos_pdf);
}
RT_D void st_pend_tail(WfPend *a, double pimp, double lres, double sh_t) {
    double2 *p = reinterpret_cast<double2 *>(a);
    p[2] = make_double2(pimp, lres); p[3] = make_double2(sh_t, 0.0);
}
RT_D void st_res(WfRes *r, double t, double u, double v) {
    double2 *p = reinterpret_cast<double2 *>(r);
    p[0] = make_double2(t, u); p[1] = make_double2(v, 0.0);
}

// Per-block event counters and list appends: one global atomic per BLOCK and per counter.  (Per warp,
// the 65 000 warps of a kernel that does little else queue up on a handful of addresses.)
struct BlockShared {
    unsigned long long ctr[4];          // segments, light-pdf rays, paths, exact fallbacks
    uint32_t cnt[2][4], base[2];        // two lists: entries per warp, position of the block's run
    uint32_t oldest;
};
RT_D void block_shared_init(BlockShared &s) {
    if (threadIdx.x < 4) s.ctr[threadIdx.x] = 0ull;
    if (threadIdx.x == 4) s.oldest = 0xFFFFFFFFu;
    __syncthreads();
}
// warp-level sum into shared memory; block_shared_flush publishes after the last one
RT_D void block_count(BlockShared &s, const Cnt &c) {
    unsigned long long v[4] = {c.seg, c.lray, c.paths, c.xf};
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
        if (lane == 0 && x) atomicAdd(&s.ctr[k], x);
    }
}
RT_D void block_note_oldest(BlockShared &s, bool in_flight, uint32_t smp) {
    uint32_t m = in_flight ? smp : 0xFFFFFFFFu;
    for (int off = 16; off > 0; off >>= 1) m = min(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0 && m != 0xFFFFFFFFu) atomicMin(&s.oldest, m);
}
RT_D void block_shared_flush(BlockShared &s, DCounters *out, uint32_t *oldest) {
    __syncthreads();
    unsigned long long *dst[4] = {&out->segments, &out->light_pdf_rays, &out->paths, &out->exact_fallbacks};
    if (threadIdx.x < 4 && s.ctr[threadIdx.x]) atomicAdd(dst[threadIdx.x], s.ctr[threadIdx.x]);
    if (threadIdx.x == 4 && s.oldest != 0xFFFFFFFFu) atomicMin(oldest, s.oldest);      // the oldest sample still in flight (k_wf_plan)
}
// Stream compaction into two lists at once: ballots + prefix sums, one atomic per block and list.
// Called by every thread of the block; pos = where this thread's entry goes (meaningful when its
// predicate holds).  Entries of a block stay together, in thread order.
RT_D void block_append2(BlockShared &s, uint32_t *count_a, bool pa, uint32_t &pos_a, uint32_t *count_b, bool pb, uint32_t &pos_b) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned ma = __ballot_sync(0xFFFFFFFFu, pa), mb = __ballot_sync(0xFFFFFFFFu, pb);
    if (lane == 0) { s.cnt[0][wid] = __popc(ma); s.cnt[1][wid] = __popc(mb); }
    __syncthreads();
    if (threadIdx.x < 2) {
        const uint32_t tot = s.cnt[threadIdx.x][0] + s.cnt[threadIdx.x][1] + s.cnt[threadIdx.x][2] + s.cnt[threadIdx.x][3];
        s.base[threadIdx.x] = tot ? atomicAdd(threadIdx.x ? count_b : count_a, tot) : 0u;
    }
    __syncthreads();
    pos_a = s.base[0]; pos_b = s.base[1];
    for (int w = 0; w < wid; ++w) { pos_a += s.cnt[0][w]; pos_b += s.cnt[1][w]; }
    const unsigned below = (1u << lane) - 1u;
    pos_a += __popc(ma & below); pos_b += __popc(mb & below);
}

// The trace list is filled from BOTH ends: slots whose ray is a fresh camera ray go to the front,
// slots in the middle of a path to the back.  The trace kernel walks front then back, so its warps
// get runs of neighbouring primary rays (which visit the same nodes) instead of a mixture of primary
// and bounced rays.  counts[kCntFront] = entries at the front, counts[kCntBack] = entries at the back.
struct LiveList {
    const uint32_t *p;
    uint32_t nf, n, last;
    RT_D uint32_t at(uint32_t i) const { return p[i < nf ? i : last - (i - nf)]; }
};
RT_D LiveList live_list(const WfBuf &b) {
    LiveList l;
    l.p = b.live;
    l.nf = b.counts[kCntFront];
    l.n = l.nf + b.counts[kCntBack];
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
__global__ void k_wf_init(WfBuf b, WfCall cl, uint32_t sample_count, uint32_t first_stop) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_px = cl.n_px;
    if (i < cl.n_slots) st_head_item(b.head + i, make_uint4(0u, kNoShape, 0u, 0u), make_uint4(0u, 0u, 0u, 0u));      // an empty slot: the first confirm pass sends it to k_wf_regen
    if (i == 0) {
        *b.call = cl;                                   // what the kernels of this call read instead of arguments
        for (int k = 0; k < kCntWords; ++k) b.counts[k] = 0;
        const unsigned long long total = (unsigned long long)n_px * sample_count, first = (unsigned long long)(b.ring - cl.reserve) * n_px;
        b.item[0] = 0; b.item[1] = first < total ? first : total; b.item[2] = total; b.item[3] = 0;
        b.plane[0] = 0; b.plane[1] = 0xFFFFFFFFu; b.plane[2] = b.ring; b.plane[3] = 0; b.plane[4] = first_stop;
    }
}

// ---- confirm -----------------------------------------------------------------------------------
// The exact f64 test of every deferred candidate (the reference's Triangle::intersection, ~400
// instructions with its nine divisions) as a kernel of its own.
// Phase 1, one (ray, candidate) PAIR per thread: a ray has 0..6 scene candidates and 0..6 light
// candidates, so a loop per ray runs with a third of the lanes; here a block gathers the pairs of
// its 128 rays with a prefix sum and deals them out evenly.
// Phase 2, one ray per thread: fold the pair results into what the shade kernel needs — the
// nearest confirmed triangle (hit_prim, res_t / res_u / res_v, ncand = 0 or 1) and the complete
// Light::pdf impact (lres: the analytic lights' part + the terms in list order = the reference's DFS
// order, bvh.rs:188-210) — and decide whether the proof that the reference finds the same hits holds
// (hit_fragile, tie_margin).  If not the slot goes to k_wf_exact.
// Phase 3: every slot of the block is sorted into the hit list, the regen list or neither.
#ifndef WF_CONFIRM_MINB
#define WF_CONFIRM_MINB 8
#endif
// pair results a block keeps in shared memory between the two phases (0.9 pairs per ray on C3: 128
// rays need ~115); the rays of a block whose pairs do not fit are answered by k_wf_exact
constexpr uint32_t kConfirmPairs = 448;
__global__ void __launch_bounds__(128, WF_CONFIRM_MINB) k_wf_confirm(DScene sc, WfBuf b, DCounters *ctr) {
    __shared__ uint16_t s_pair[128 * 2 * kWfCand];
    __shared__ uint32_t s_warp[4];
    __shared__ double s_ray[6][128];
    __shared__ double s_t[kConfirmPairs], s_u[kConfirmPairs], s_v[kConfirmPairs];   // phase 1 -> phase 2
    __shared__ BlockShared s_blk;
    const uint32_t tid = threadIdx.x, slot0 = blockIdx.x * 128, i = slot0 + tid, lane = tid & 31, wid = tid >> 5;
    const uint32_t n_slots = b.call->n_slots;
    if (slot0 >= n_slots) return;                       // the grid covers the buffer's capacity
    uint32_t nc = 0, nl = 0, flags = 0, nc_all = 0, sh_id = kNoShape;
    uint4 item = make_uint4(0u, 0u, 0u, 0u);
    bool mine = false;                                  // this thread's slot has results to fold
    int cls = 0;                                        // 1: hit list, 2: regen list
    if (i < n_slots) {
        const uint4 h = ld_head(b.head + i);
        item = ld_item(b.head + i);                     // same sector: lets phase 2 write the record back whole
        flags = h.x; sh_id = h.y; nc_all = h.z;
        if (flags & kFlagExactPending) {
            // k_wf_exact answered this slot during the last iteration (the graph joined since):
            // its results are final, the shade kernel takes them in this iteration
            flags &= ~kFlagExactPending;
            b.head[i].flags = flags;
            cls = ((nc_all & 0xFFu) || sh_id != kNoShape) ? 1 : 2;
        } else if (flags & kFlagAlive) {
            mine = true;
            nc = (nc_all & kOverflowClosest) ? 0u : (nc_all & 0xFFu);       // overflowed lists: the exact traversal redoes the search
            nl = ((nc_all & kOverflowLight) || !(flags & kFlagPendLight)) ? 0u : ((nc_all >> 8) & 0xFFu);
            if (nc + nl) {
                d3 o, d;
                ld_ray(b.ray + i, o, d);
                s_ray[0][tid] = o.x; s_ray[1][tid] = o.y; s_ray[2][tid] = o.z;
                s_ray[3][tid] = d.x; s_ray[4][tid] = d.y; s_ray[5][tid] = d.z;
            }
        } else {
            cls = 2;                                    // no path: wants the next item
        }
    }
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
        const WfCand *oc = b.cand + slot0 + owner;
        const d3 o = mk3(s_ray[0][owner], s_ray[1][owner], s_ray[2][owner]);
        const d3 d = mk3(s_ray[3][owner], s_ray[4][owner], s_ray[5][owner]);
        const uint32_t prim = light ? oc->lcand[k] : oc->cand[k];
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
            const double2 *pt = reinterpret_cast<const double2 *>(b.pend + i);
            const double2 t0 = pt[2], t1 = pt[3];       // (pimp, lres) (sh_t, -): the sector is written back whole
            double impact = t0.x;
            for (uint32_t k = 0; k < nl; ++k) {
                const double term = s_t[base + k];
                if (term == -2.0) need_exact = true;
                else if (term != -1.0) impact += term;
            }
            st_pend_tail(b.pend + i, t0.x, impact, t1.x);
        }
        bool tri_hit = false;
        if (!need_exact) {
            FBest fb;
            fbest_clear(fb);
            for (uint32_t k = 0; k < nc; ++k) {
                const uint32_t e = base + nl + k;
                const double u = s_u[e];
                if (u != -1.0) fbest_take(fb, s_t[e], u, s_v[e], b.cand[i].cand[k], u == -2.0);
            }
            if (fb.has) {
                const d3 o = mk3(s_ray[0][tid], s_ray[1][tid], s_ray[2][tid]), d = mk3(s_ray[3][tid], s_ray[4][tid], s_ray[5][tid]);
                need_exact = fbest_needs_exact(fb, o, d, (double)sc.ftris.bound);
                st_res(b.res + i, fb.t, fb.u, fb.v);
            }
            if (!need_exact) {
                // a ray without candidates already says ncand = 0; otherwise the whole sector goes back (no partial write)
                if (nc_all != 0u) st_head_item(b.head + i, make_uint4(flags, sh_id, fb.has ? 1u : 0u, fb.prim), item);
                tri_hit = fb.has;
            }
        }
        if (need_exact) {
            // Nothing is shaded or traced for the slot in this iteration: k_wf_exact answers both searches
            // with the reference's own traversal meanwhile, the next confirm pass picks the slot up.
            b.head[i].flags = flags | kFlagExactPending;
            b.xlist[atomicAdd(&b.counts[kCntExact], 1u)] = i;
            atomicAdd(&ctr->exact_fallbacks, 1ull);     // rare
            atomicMin(&b.plane[1], item.x);             // its sample is still in flight
        } else {
            cls = (tri_hit || sh_id != kNoShape) ? 1 : 2;
        }
    }
    // ---- phase 3: the lists the next kernels walk
    uint32_t pos_hit, pos_regen;
    block_append2(s_blk, &b.counts[kCntHit], cls == 1, pos_hit, &b.counts[kCntRegen], cls == 2, pos_regen);
    if (cls == 1) b.hit_list[pos_hit] = i;
    if (cls == 2) b.regen_list[0][pos_regen] = i;
}

// ---- shade: the slots whose ray hit something ----------------------------------------------------
#ifndef WF_SHADE_MINB
#define WF_SHADE_MINB 5        // 96 registers; measured 4 / 5 / 6 blocks per SM -> 1504 / 1510 / 1495 Mseg/s
#endif
template <bool SHAPES>
__global__ void __launch_bounds__(128, SHAPES ? 4 : WF_SHADE_MINB) k_wf_shade_hit(DScene sc, WfBuf b, DCounters *ctr) {
    const uint32_t n = b.counts[kCntHit];
    if (blockIdx.x * 128u >= n) return;                 // the grid covers the buffer's capacity
    __shared__ BlockShared s_blk;
    block_shared_init(s_blk);
    Cnt c;
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    const bool on = i < n;
    // The lanes of a warp part ways twice — a pending weight or none, light sampling or cosine sampling —
    // and left to itself the compiler keeps the four groups apart until the end of the kernel (ncu: 12.5 of
    // 32 lanes with every lane holding a hit).  Warp barriers over the lanes that hold an entry bring them
    // back together after each of the two.
    const unsigned mon = __ballot_sync(kFull, on);
    bool go_on = false, ended = false;
    uint32_t slot = 0, smp = 0;
    if (on) {
        slot = b.hit_list[i];
        WfRay *const rp = b.ray + slot;
        WfTL *const tp = b.tl + slot;
        d3 o, d, T, L;
        ld_ray(rp, o, d);
        ld_TL(tp, T, L);
        const uint4 h = ld_head(b.head + slot), it = ld_item(b.head + slot);
        const uint32_t flags = h.x;
        smp = it.x;
        uint32_t left = (flags >> 8) & 0xFFu;
        // the draws of this bounce (a diffuse bounce takes four), computed while every lane of the warp is here
        RngW rng;
        rng.init(b.call->seed, it.y, b.call->sample_begin + smp, it.w);
        rng.prefetch(2);
        Pending pd;
        pd.v = mk3(0.0, 0.0, 0.0); pd.cos_pdf = 0.0;
        bool pend_light = false;
        if (flags & kFlagPendLight) {                   // finish the previous diffuse bounce: Light::pdf of THIS ray
            double impact;                              // folded by k_wf_confirm or answered by k_wf_exact
            ld_pending(b.pend + slot, pd, impact);
            if (!apply_pending(pd, impact / (double)sc.n_lights, T)) ended = true;
        }
        __syncwarp(mon);
        if (!ended) {
            ShapeBest sb;
            sb.has = SHAPES && h.y != kNoShape; sb.kind = sb.has ? (int)(h.y >> 28) : -1; sb.prim = h.y & 0x0FFFFFFFu;
            sb.t = sb.has ? b.pend[slot].sh_t : 0.0;
            FBest fb;
            fbest_clear(fb);
            bool skip;
            (void)shapes_tmax(sb, skip);
            if (!skip && (h.z & 0xFFu)) {               // the nearest confirmed triangle, folded by k_wf_confirm
                const double2 r = reinterpret_cast<const double2 *>(b.res + slot)[0];
                fb.has = true; fb.prim = h.w;
                fb.t = r.x; fb.u = r.y; fb.v = b.res[slot].v;
            }
            ++c.seg;
            const SceneHit hit = finish_hit<SHAPES>(sc, o, d, sb, fb);
            const int r = shade_hit<SHAPES>(sc, hit, o, d, T, L, rng, pd);
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
        __syncwarp(mon);
        if (ended) {
            // the second k_wf_regen pass parks the radiance in the plane of the sample and gives the slot its next item
            st_TL(tp, T, L);
            st_head_item(b.head + slot, make_uint4(kFlagEnded, kNoShape, 0u, 0u), it);
        } else {                                        // hand the ray to the trace kernel
            uint32_t sh_id = kNoShape;
            double sh_t = 0.0;
            if (SHAPES) {
                const ShapeBest nb = shapes_closest(sc, o, d, c);
                sh_id = nb.has ? (((uint32_t)nb.kind << 28) | nb.prim) : kNoShape;
                sh_t = nb.t;
            }
            if (pend_light) {
                ++c.lray;
                st_pending(b.pend + slot, pd);
            }
            if (pend_light || sh_id != kNoShape) st_pend_tail(b.pend + slot, (SHAPES && pend_light) ? light_shapes_impact(sc, o, d, c) : 0.0, 0.0, sh_t);
            st_ray(rp, o, d);
            st_TL(tp, T, L);
            st_head_item(b.head + slot, make_uint4(kFlagAlive | (pend_light ? kFlagPendLight : 0u) | (left << 8), sh_id, 0u, 0u),
                         make_uint4(it.x, it.y, it.z, rng.pos));
            go_on = true;
        }
    }
    // bounced rays go to the BACK of the trace list, ended paths to the second regen list
    uint32_t pos_back, pos_regen;
    block_append2(s_blk, &b.counts[kCntBack], go_on, pos_back, &b.counts[kCntRegen + 1], ended, pos_regen);
    if (go_on) b.live[b.n - 1u - pos_back] = slot;
    if (ended) b.regen_list[1][pos_regen] = slot;
    block_note_oldest(s_blk, go_on, smp);
    block_count(s_blk, c);
    block_shared_flush(s_blk, ctr, &b.plane[1]);
}

// ---- regen: escaped rays, ended paths, empty slots -> the next item -----------------------------
// Slots are not tied to pixels: a slot takes whatever item is next, so all slots stay busy until
// the last sample of the call has been started, however unevenly path lengths are spread over the
// image.  What keeps the result the megakernel's bit for bit is the ORDER of the additions: every
// finished sample is parked in the plane of its sample index, and complete planes are added to the
// accumulation buffer one after the other, sample 0 first (main.rs:100-103).
// Items are handed out BY POSITION: entry i of regen list 0 takes item next + i, entry i of list 1
// item next + |list 0| + i, as far as the claim limit (the planes of the ring still being filled)
// allows; the fresh camera ray goes to the same position of the trace list's front.  Every thread
// derives this from the list lengths — no atomic operation, and the camera rays of neighbouring
// threads are neighbouring pixels (8x4 tiles).  k_wf_plan moves `next` on afterwards.
struct RegenPlan {
    unsigned long long first;           // first item of this pass
    uint32_t grant, front;              // items this pass may take; trace-list position of its first camera ray
    uint32_t granted_so_far;            // grant of both passes up to and including this one
};
RT_D RegenPlan regen_plan(const WfBuf &b, int pass) {
    const unsigned long long next = b.item[0], limit = b.item[1];
    const unsigned long long room = limit > next ? limit - next : 0ull;
    const uint32_t n0 = b.counts[kCntRegen], g0 = (uint32_t)min((unsigned long long)n0, room);
    RegenPlan p;
    if (pass == 0) { p.first = next; p.grant = g0; p.front = 0u; p.granted_so_far = g0; return p; }
    const uint32_t n1 = b.counts[kCntRegen + 1], g1 = (uint32_t)min((unsigned long long)n1, room - g0);
    p.first = next + g0; p.grant = g1; p.front = g0; p.granted_so_far = g0 + g1;
    return p;
}
template <bool SHAPES>
__global__ void __launch_bounds__(128, SHAPES ? 4 : 8) k_wf_regen(DScene sc, WfBuf b, DCounters *ctr, int pass) {
    const uint32_t n = b.counts[kCntRegen + pass];
    const RegenPlan plan = regen_plan(b, pass);
    if (blockIdx.x == 0 && threadIdx.x == 0 && pass == 1) {
        // the bookkeeping of the whole iteration's claims (this block always runs, whatever the list lengths)
        b.item[3] = plan.granted_so_far;                // k_wf_plan adds it to `next`
        b.counts[kCntFront] = plan.granted_so_far;      // the trace list's front is complete
        if (plan.granted_so_far) {
            atomicAdd(&ctr->paths, (unsigned long long)plan.granted_so_far);
            atomicMin(&b.plane[1], (uint32_t)(b.item[0] / b.call->n_px));      // the oldest of the samples just started
        }
    }
    if (blockIdx.x * 128u >= n) return;
    __shared__ BlockShared s_blk;
    block_shared_init(s_blk);
    const int lane = threadIdx.x & 31;
    const WfCall cl = *b.call;                          // uniform: one broadcast load
    const uint32_t n_px = cl.n_px;
    Cnt c;
    const uint32_t i = blockIdx.x * 128u + threadIdx.x;
    const bool on = i < n;
    if (on) {
        const uint32_t slot = b.regen_list[pass][i];
        WfRay *const rp = b.ray + slot;
        WfTL *const tp = b.tl + slot;
        const uint32_t old_flags = b.head[slot].flags;
        if (old_flags & (kFlagAlive | kFlagEnded)) {
            const uint4 it = ld_item(b.head + slot);
            d3 T, L;
            ld_TL(tp, T, L);
            if (old_flags & kFlagAlive) {               // the ray escaped: raytrace_impl returns bg_color (raytrace.rs:14)
                bool counts_as_segment = true;
                if (old_flags & kFlagPendLight) {       // ... once the weight of the bounce that shot it is known
                    Pending pd;
                    double impact;
                    ld_pending(b.pend + slot, pd, impact);
                    counts_as_segment = apply_pending(pd, impact / (double)sc.n_lights, T);
                }
                if (counts_as_segment) {
                    ++c.seg;
                    L = add(L, mul3(T, sc.bg));
                }
            }
            // The finished sample's radiance goes to the plane of its sample index; k_wf_sum adds the
            // planes to the accumulation buffer in sample order (main.rs:102), whatever slot computed
            // them and whenever.
            double2 *side = reinterpret_cast<double2 *>(b.side + (size_t)(it.x % b.ring) * n_px + it.z);
            side[0] = make_double2(L.x, L.y); side[1] = make_double2(L.z, 0.0);
        }
        if (i < plan.grant) {
            // item j = sample (j / n_px) of the pixel with ordinal (j % n_px) in the region's tile order;
            // one 64-bit division per warp, the lanes count on from its first lane
            const unsigned long long j0 = plan.first + (i - (uint32_t)lane);
            uint32_t smp = (uint32_t)(j0 / n_px);
            unsigned long long oo = j0 - (unsigned long long)smp * n_px + (uint32_t)lane;
            while (oo >= n_px) { oo -= n_px; ++smp; }
            const uint32_t ord = (uint32_t)oo;
            const uint32_t pix = region_tile_pixel(cl.rg, sc.width, ord);
            RngW rng;
            rng.init(cl.seed, pix, cl.sample_begin + smp);
            rng.prefetch(1);
            const double jx = rng.range(0.0, 1.0), jy = rng.range(0.0, 1.0);       // camera.rs:49-50
            d3 o, d;
            camera_ray(sc.cam, pix % sc.width, pix / sc.width, jx, jy, o, d);
            uint32_t sh_id = kNoShape;
            if (SHAPES) {
                const ShapeBest nb = shapes_closest(sc, o, d, c);
                sh_id = nb.has ? (((uint32_t)nb.kind << 28) | nb.prim) : kNoShape;
                if (nb.has) st_pend_tail(b.pend + slot, 0.0, 0.0, nb.t);
            }
            st_ray(rp, o, d);
            st_TL(tp, mk3(1.0, 1.0, 1.0), mk3(0.0, 0.0, 0.0));
            st_head_item(b.head + slot, make_uint4(kFlagAlive | ((uint32_t)sc.ray_depth << 8), sh_id, 0u, 0u), make_uint4(smp, pix, ord, rng.pos));
            b.live[plan.front + i] = slot;              // fresh camera rays: the FRONT of the trace list, in item order
        } else if (old_flags != 0u) {
            b.head[slot].flags = 0u;                    // nothing to take right now: an empty slot, asks again in the next iteration
        }
    }
    block_count(s_blk, c);
    block_shared_flush(s_blk, ctr, &b.plane[1]);
}

// ---- sum -----------------------------------------------------------------------------------------
// k_wf_plan (one thread, after the regen passes): planes [summed, upto) are complete — fully
// claimed and older than the oldest sample still in flight — and the claim limit moves on.
__global__ void k_wf_plan(WfBuf b) {
    const uint32_t n_px = b.call->n_px, reserve = b.call->reserve;
    const uint32_t summed = b.plane[3];                 // the sum kernel of the last iteration has run
    b.plane[0] = summed;
    b.item[0] += b.item[3];                             // the items the regen passes of this iteration took
    b.item[3] = 0;
    const unsigned long long total = b.item[2], next = min(b.item[0], total);
    const unsigned long long claimed = next / n_px;     // planes whose items have all been handed out
    const uint32_t oldest = b.plane[1];
    uint32_t upto = (uint32_t)min(claimed, (unsigned long long)oldest);
    // a progressive render holds the sums at the end of the chunk the host is about to show (plane[4]) while
    // the paths of the next chunk are already in flight: the buffer then holds exactly that many samples
    const uint32_t stop = b.plane[4];
    const bool held = upto >= stop;
    if (held) upto = stop;
    // batches of four planes amortise the read-modify-write of the accumulation buffer; at the end
    // of the call (nothing left to claim) everything that is complete goes
    if (upto < summed + b.batch && next < total && !held) upto = summed;
    if (upto < summed) upto = summed;
    b.plane[3] = upto;
    // b.reserve planes of the ring stay free (headroom; the regen passes never take an item past the limit)
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
            const double2 *sp = reinterpret_cast<const double2 *>(b.side + (size_t)(smp % b.ring) * n_px + od);
            const double2 xy = sp[0];
            ax += xy.x; ay += xy.y; az += sp[1].x;
        }
        accum[p] = ax; accum[p + 1] = ay; accum[p + 2] = az;
    }
}

// ---- exact ---------------------------------------------------------------------------------------
// The slots k_wf_confirm could not vouch for (normally a few dozen per iteration: paths leaving a
// flat light, rays through a mesh vertex): both searches of the ray redone with the reference's
// own traversal (exact.cuh) and written back in the form k_wf_confirm leaves its results in — the
// nearest triangle, the complete light impact.  The slot's flag stays set for the whole iteration
// (the next confirm pass clears it) and the slot is in none of the lists, so whether this kernel is
// early or late changes nothing.  One such traversal is a long serial chain, so this kernel runs on a
// forked branch of the graph, next to the shade, regen and trace kernels.
__global__ void __launch_bounds__(128) k_wf_exact(DScene sc, WfBuf b) {
    const uint32_t n = b.counts[kCntExact];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t slot = b.xlist[i];
        d3 o, d;
        ld_ray(b.ray + slot, o, d);
        Cnt scratch;
        if (b.head[slot].flags & kFlagPendLight) {
            double impact = b.pend[slot].pimp;          // the analytic lights' part (shade kernel), then the triangles
            light_hits_exact<3>(sc.ltris, sc.ltri_i, sc.ltri_s, o, d, impact, scratch);
            b.pend[slot].lres = impact;
        }
        const Best bt = closest_exact<3>(sc.tris, sc.tri_i, o, d, scratch);
        if (bt.has) {
            double t, u, v;
            tri_exact(sc.tri_i + bt.prim, o, d, t, u, v);
            b.head[slot].hit_prim = bt.prim;
            st_res(b.res + slot, t, u, v);
        }
        b.head[slot].ncand = bt.has ? 1u : 0u;          // no overflow: nothing left to search
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
// Refills take 16 lanes at a time (measured on B200, Mseg/s on C3: 4 / 8 / 12 / 16 / 20 / 24 / 32 lanes
// -> 1325 / 1340 / 1366 / 1419 / 1426 / 1421 / 1331): rays that start at the root TOGETHER stay in
// step through the top of the hierarchy and fetch the same nodes; handing rays out a few at a time
// from a run claimed per warp loses that, however early the idle lanes get back to work.
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

// SMEM_TOP: the top levels of the scene hierarchy (nodes [0, top), breadth-first: every ray starts
// there) are staged in shared memory, one contiguous copy per block.  Measured on B200 (C3): 947
// Mseg/s with staging against 1015 without at the time — an L1 hit costs what a shared-memory read
// costs, and the per-lane choice between the two address spaces is extra work in the hottest loop —
// so the default is off; RT_WF_SMEM_TOP=1 (read at rt_create) selects this instantiation.
template <bool SMEM_TOP>
__global__ void __launch_bounds__(128, WF_TRACE_BLOCKS) k_wf_trace(DScene sc, WfBuf b, DCounters *ctr) {
    const LiveList ll = live_list(b);
    const uint32_t n = ll.n;
    const int lane = threadIdx.x & 31;
    const float inf = __int_as_float(0x7F800000);
    __shared__ float4 s_top[SMEM_TOP ? kFastTopNodes * 8 : 1];
    uint32_t top_n = 0;
    if (SMEM_TOP) {
        top_n = sc.ftris.root >= 0 ? (sc.ftris.top < kFastTopNodes ? sc.ftris.top : kFastTopNodes) : 0u;
        for (uint32_t k = threadIdx.x; k < top_n * 8; k += blockDim.x) s_top[k] = __ldg(reinterpret_cast<const float4 *>(sc.ftris.nodes) + k);
        __syncthreads();
    }
    bool has = false, drained = false, closest = true, want_light = false;
    uint32_t nc = 0, nl = 0, over = 0, tri_i = 0, tri_end = 0;
    // per-lane values that are written at refill and read only when a candidate is stored or the ray
    // is published live in shared memory, not in registers (the kernel runs at 64 registers, 8 blocks per SM)
    __shared__ uint32_t s_slot[128];
#define my_slot s_slot[threadIdx.x]
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
                    b.head[my_slot].ncand = nc | (nl << 8) | over;
                    has = false;
                }
            }
            // refill idle lanes from the live list: ballot + prefix sum, one atomic per warp
            const unsigned idle = __ballot_sync(kFull, !has);
            if (!drained && __popc(idle) >= kRefillLanes) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&b.counts[kCntCursor], (uint32_t)__popc(idle));
                base = __shfl_sync(kFull, base, leader);
                // the highest index handed out in this refill tells whether the list ran out
                if (base + (uint32_t)__popc(idle) >= n) drained = true;
                if (!has) {
                    const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < n) {
                        const uint32_t slot_new = ll.at(idx);
                        my_slot = slot_new;
                        d3 o, dd;
                        ld_ray(b.ray + slot_new, o, dd);
                        const uint4 h = ld_head(b.head + slot_new);
                        my_ray[0] = o.x; my_ray[128] = o.y; my_ray[256] = o.z;
                        my_ray[384] = dd.x; my_ray[512] = dd.y; my_ray[640] = dd.z;
                        r = make_fray(o, dd);
                        const uint32_t sid = h.y;
                        const double st = sid != kNoShape ? b.pend[slot_new].sh_t : 0.0;
                        bool skip = (sid != kNoShape && st != st) || ray_is_nan(o, dd);
                        tmax = (sid != kNoShape && !skip) ? up_f32(st) : inf;
                        want_light = (h.x & kFlagPendLight) != 0;
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
                        has = true;
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
                Visit4 v;
                if (SMEM_TOP) {
                    const FNode *np = (closest && (uint32_t)cur < top_n) ? reinterpret_cast<const FNode *>(s_top) + cur : (closest ? sc.ftris.nodes : sc.fltris.nodes) + cur;
                    v = visit4<true>(np, r, tmax);
                } else {
                    v = visit4((closest ? sc.ftris.nodes : sc.fltris.nodes) + cur, r, tmax);
                }
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
                            b.cand[my_slot].cand[nc] = tri_i;
                            ++nc;
                        } else {
                            over |= kOverflowClosest;           // rare: k_wf_exact redoes this search
                            WF_INC(n_over)
                        }
                    } else {
                        if (nl < (uint32_t)kWfCand) { b.cand[my_slot].lcand[nl] = tri_i; ++nl; }
                        else { over |= kOverflowLight; WF_INC(n_over) }
                    }
                }
            } while (++tri_i != tri_end);
            WF_POP()
        }
    }
#undef my_slot
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
size_t wf_bytes_per_slot() { return sizeof(WfHead) + sizeof(WfRay) + sizeof(WfTL) + sizeof(WfPend) + sizeof(WfRes) + sizeof(WfCand) + 5 * sizeof(uint32_t); }      // + ring * 32 per PIXEL for the side planes
int wf_trace_blocks(int device, bool smem_top) {
    int sms = 148, per_sm = 4;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const cudaError_t e = smem_top ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_trace<true>, 128, 0)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_trace<false>, 128, 0);
    if (e != cudaSuccess || per_sm < 1) per_sm = 4;
    return sms * per_sm;
}
cudaError_t wf_runtime_init(WfRuntime &rt, int device) {
    cudaError_t e;
    (void)device;
    if ((e = cudaStreamCreateWithFlags(&rt.side, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&rt.side2, cudaStreamNonBlocking)) != cudaSuccess) return e;
    cudaEvent_t *evs[5] = {&rt.ev_fork, &rt.ev_shade, &rt.ev_join, &rt.ev_poll, &rt.ev_regen};
    for (cudaEvent_t *ev : evs)
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    uint32_t *host = nullptr;
    if ((e = cudaHostAlloc((void **)&host, 16 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return e;
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
    if (rt.side2) { cudaStreamSynchronize(rt.side2); cudaStreamDestroy(rt.side2); }
    cudaEvent_t evs[5] = {rt.ev_fork, rt.ev_shade, rt.ev_join, rt.ev_poll, rt.ev_regen};
    for (cudaEvent_t ev : evs) if (ev) cudaEventDestroy(ev);
    if (rt.host_live) cudaFreeHost((void *)rt.host_live);
    rt = WfRuntime{};
}

constexpr int kGraphIters = 4;              // iterations per graph launch
constexpr int kLaunchesPerIter = 8;         // confirm, exact, shade_hit, regen x 2, plan, sum, trace

cudaError_t launch_render_wavefront(cudaStream_t st, WfRuntime &rt, const DScene &sc, uint32_t sample_begin, uint32_t sample_count,
                                    uint64_t seed, double *accum, DCounters *ctr, const WfBuf &b, int trace_blocks,
                                    uint64_t *launches, DRegion rg, const WfProgress *progress) {
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
    // The grids cover the buffer's CAPACITY (blocks past the slots in use, or past the end of a
    // list, leave at once), so the launch configuration does not depend on the call either.
    const unsigned blocks = (b.n + 127) / 128;
    // progressive: the sums stop at `stop` samples until the host has shown that image (WfProgress)
    const uint32_t chunk = progress && progress->chunk ? progress->chunk : 0u;
    uint32_t stop = chunk && chunk < sample_count ? chunk : 0xFFFFFFFFu;
    k_wf_init<<<(cl.n_slots + 127) / 128, 128, 0, st>>>(b, cl, sample_count, stop);
    ++*launches;
    // a segment takes one iteration, two when it is deferred to k_wf_exact; an item can wait for
    // a free plane of the ring: generous bound, reaching it is reported as an error
    const uint64_t max_iters = 4 * ((uint64_t)sample_count * n_px / cl.n_slots + 1) * (sc.ray_depth ? sc.ray_depth : 1) + 64;
    // scenes without planes / boxes / ellipsoids (every glTF scene) run the kernels compiled without them
    const bool shapes = sc.n_planes || sc.boxes.n_prims || sc.ellipsoids.n_prims || sc.lboxes.n_prims || sc.lellipsoids.n_prims;
    const unsigned sum_blocks = (unsigned)std::min<uint64_t>(((uint64_t)b.n_px_cap + 255) / 256, 148 * 8);
    // One iteration: confirm -> shade_hit -> regen (pass 1) -> trace on the main stream; forked after
    // confirm: k_wf_exact, and the first regen pass next to shade_hit; forked after the second regen pass:
    // the one-thread k_wf_plan + k_wf_sum, next to the trace kernel; all joined before the next iteration.  Four iterations are captured ONCE per (scene,
    // buffers) into a CUDA graph and replayed by every call: the inner loop is launch-bound on small
    // images and at the end of every render.
    auto issue_iteration = [&]() {
        cudaMemsetAsync(b.counts, 0, kCntWords * sizeof(uint32_t), st);     // the lists start empty; trace fetch position; deferred-slot count
        k_wf_confirm<<<blocks, 128, 0, st>>>(sc, b, ctr);
        cudaEventRecord(rt.ev_fork, st);
        cudaStreamWaitEvent(rt.side, rt.ev_fork, 0);
        cudaStreamWaitEvent(rt.side2, rt.ev_fork, 0);
        k_wf_exact<<<64, 128, 0, rt.side>>>(sc, b);
        if (shapes) {
            k_wf_regen<true><<<blocks, 128, 0, rt.side2>>>(sc, b, ctr, 0);
            k_wf_shade_hit<true><<<blocks, 128, 0, st>>>(sc, b, ctr);
        } else {
            k_wf_regen<false><<<blocks, 128, 0, rt.side2>>>(sc, b, ctr, 0);
            k_wf_shade_hit<false><<<blocks, 128, 0, st>>>(sc, b, ctr);
        }
        cudaEventRecord(rt.ev_regen, rt.side2);
        cudaStreamWaitEvent(st, rt.ev_regen, 0);
        if (shapes) k_wf_regen<true><<<blocks, 128, 0, st>>>(sc, b, ctr, 1);
        else k_wf_regen<false><<<blocks, 128, 0, st>>>(sc, b, ctr, 1);
        cudaEventRecord(rt.ev_shade, st);
        cudaStreamWaitEvent(rt.side, rt.ev_shade, 0);
        k_wf_plan<<<1, 1, 0, rt.side>>>(b);
        k_wf_sum<<<sum_blocks, 256, 0, rt.side>>>(b, sc.width, accum);
        cudaEventRecord(rt.ev_join, rt.side);
        if (rt.smem_top) k_wf_trace<true><<<trace_blocks, 128, 0, st>>>(sc, b, ctr);
        else k_wf_trace<false><<<trace_blocks, 128, 0, st>>>(sc, b, ctr);
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
    volatile uint32_t *host_stop = rt.host_live + 7;       // pinned staging word for the stop line
    uint32_t stopped_at = 0;                                // the callback asked to stop after this many samples
    bool poll_pending = false, finished = false;
    cudaError_t result = cudaSuccess;
    uint32_t showing = 0;                                   // a chunk image is on its way to the host (prepare done, show pending)
    for (uint64_t it = 0; it < max_iters; it += kGraphIters) {
        if (use_graph) {
            result = cudaGraphLaunch(rt.exec, st);
            if (result != cudaSuccess) break;
        } else {
            for (int k = 0; k < kGraphIters; ++k) issue_iteration();
        }
        *launches += kGraphIters * kLaunchesPerIter;
        if (showing) {
            // the image was enqueued before the group just launched: wait for it and hand it over while that group runs
            const int rc = progress->show(progress->user, showing);
            if (rc < 0) { result = cudaErrorUnknown; break; }
            if (rc > 0) { stopped_at = showing; finished = true; break; }
            showing = 0;
            stop = stop + chunk < sample_count ? stop + chunk : 0xFFFFFFFFu;
            *host_stop = stop;
            cudaMemcpyAsync(&b.plane[4], (const void *)host_stop, sizeof(uint32_t), cudaMemcpyHostToDevice, st);
        }
        // poll the state after the PREVIOUS group of four without stalling the stream: nothing
        // traced, nothing deferred and nothing left to claim means the call is done
        if (poll_pending) {
            result = cudaEventSynchronize(rt.ev_poll);
            if (result != cudaSuccess) break;
            const unsigned long long next = (unsigned long long)host_live[4] | ((unsigned long long)host_live[5] << 32);
            if (stop == 0xFFFFFFFFu && host_live[kCntFront] + host_live[kCntBack] + host_live[kCntExact] == 0 && next >= total_items) { finished = true; break; }
            if (stop != 0xFFFFFFFFu && !showing && host_live[6] >= stop) {
                // The accumulation buffer holds exactly `stop` samples of every pixel and will not change until
                // the stop line moves: tonemap + copy are enqueued now, the next group of iterations behind
                // them (the paths of the next chunk are in flight), and the image is handed over while it runs.
                if (progress->prepare(progress->user, stop) < 0) { result = cudaErrorUnknown; break; }
                showing = stop;
            }
        }
        cudaMemcpyAsync((void *)host_live, b.counts, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync((void *)(host_live + 4), b.item, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync((void *)(host_live + 6), &b.plane[3], sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        cudaEventRecord(rt.ev_poll, st);
        poll_pending = true;
    }
    if (progress) progress->samples_done = stopped_at ? stopped_at : sample_count;
    if (result == cudaSuccess && !stopped_at) {
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
