// kernels.hpp — host-callable launchers of the kernels in kernels.cu / wavefront.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_scene.cuh"

namespace rt {

cudaError_t launch_trace_rays(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *o, const double *d,
                              void *hits, DCounters *ctr);
cudaError_t launch_trace_primary(cudaStream_t st, const DScene &sc, int mode, double jx, double jy, void *hits,
                                 double *rays, DCounters *ctr);
cudaError_t launch_light_pdf(cudaStream_t st, const DScene &sc, int mode, uint64_t n, const double *pos, const double *dir,
                             double *pdf, DCounters *ctr);
cudaError_t launch_render_mega(cudaStream_t st, const DScene &sc, int mode, uint32_t sample_begin, uint32_t sample_count,
                               uint64_t seed, double *accum, DCounters *ctr, DRegion rg);
// ---- wavefront scheduler (wavefront.cu) ----------------------------------------------------------
constexpr int kWfCand = 6;              // deferred candidates kept per ray and per search
constexpr uint32_t kWfRing = 16;        // sample planes of finished radiance kept before they are added to the image, in order
// What differs from one render call to the next lives in DEVICE memory, so the kernel arguments —
// and with them the instantiated CUDA graph of the iteration — stay the same for as long as the
// scene and the buffers do: a call costs one k_wf_init launch, not a capture + instantiate.
struct WfCall {
    uint32_t sample_begin, n_slots, reserve, n_px;  // first global sample; slots in use (<= WfBuf::n); planes kept free for claim overshoot; pixels of the region
    unsigned long long seed;
    DRegion rg;
};
// Path state, grouped by WHO touches it.  The confirm kernel walks the slots in index order, the other
// kernels walk compacted lists (a gather): every group is a record of whole 32-byte sectors, and every
// writer writes whole sectors (a partial sector costs a read-modify-write in HBM: the kernels that walk
// lists run at half the HBM bandwidth and more, so those bytes are time).
struct __align__(16) WfHead {           // 32 B = one sector: every kernel reads it; shade_hit / regen rewrite it in full
    uint32_t flags;                     // alive / ended / pending Light::pdf / deferred to k_wf_exact | depth left << 8
    uint32_t sh_id;                     // nearest analytic shape (kind << 28 | index) or 0xFFFFFFFF
    uint32_t ncand;                     // trace -> confirm: #scene | #light << 8 | overflow bits; confirm -> shade: 0 / 1
    uint32_t hit_prim;                  // nearest confirmed triangle
    uint32_t sample, pix, ord, rng_pos; // the item in flight (sample of the call, pixel, ordinal in the region), RNG draws consumed
};
struct __align__(16) WfRay {            // 64 B: trace (gather), confirm, shade_hit, regen
    double o[3], d[3];                  // the ray to trace / just traced
    double _spare[2];
};
struct __align__(16) WfTL {             // 64 B: shade_hit, regen
    double T[3], L[3];                  // throughput, radiance so far
    double _spare[2];
};
struct __align__(16) WfPend {           // 64 B: slots with a pending diffuse bounce (and, text scenes, a shape hit)
    double pv[3], pcos;                 // sector 0: dot * color / PI, Cosine::pdf
    double pimp, lres;                  // sector 1: Light::pdf impact: the analytic lights' part, the complete sum
    double sh_t, _spare;                //           distance of sh_id
};
struct __align__(16) WfRes {            // 32 B: confirm / exact -> shade_hit
    double t, u, v, _spare;             // exact (t, u, v) of hit_prim
};
struct __align__(16) WfCand {           // 64 B: trace -> confirm
    uint32_t cand[kWfCand], lcand[kWfCand];     // triangles that survived the fp32 pre-test (scene / light hierarchy)
    uint32_t _spare[4];
};
static_assert(sizeof(WfHead) == 32 && sizeof(WfRay) == 64 && sizeof(WfTL) == 64 && sizeof(WfPend) == 64 && sizeof(WfRes) == 32 && sizeof(WfCand) == 64,
              "slot records are whole sectors");
struct WfBuf {
    uint32_t n;                         // slots allocated
    WfHead *head;
    WfRay *ray;
    WfTL *tl;
    WfPend *pend;
    WfRes *res;
    WfCand *cand;
    uint32_t *live;                     // n: the trace list, filled from both ends (front: fresh camera rays, back: bounced rays)
    uint32_t *hit_list;                 // n: slots whose ray hit something (k_wf_shade_hit)
    uint32_t *regen_list[2];            // n each: slots that need a new path — [0] found by k_wf_confirm (ray escaped, or no path),
                                        // [1] paths k_wf_shade_hit ended
    uint32_t *counts;                   // [0] trace list front, [1] back, [2] trace fetch cursor, [3] slots deferred to k_wf_exact,
                                        // [4] hit list, [5] regen list 0, [6] regen list 1 — cleared together at the start of an iteration
    uint32_t *xlist;                    // n: slots deferred to k_wf_exact in this iteration
    uint32_t batch;                     // planes added to the image per k_wf_sum pass (amortises the read-modify-write)
    uint32_t ring;                      // planes in the ring of finished samples
    uint32_t n_px_cap;                  // pixels the side planes were allocated for (width * height)
    WfCall *call;                       // device: the parameters of the render call in flight (written by k_wf_init)
    double4 *side;                      // ring * n_px: radiance of finished samples (x, y, z, -): plane (sample % ring), ordinal — one sector each
    unsigned long long *item;           // [0] next unclaimed item, [1] claim limit, [2] items of this call, [3] items taken in this iteration
    uint32_t *plane;                    // [0] planes added to accum, [1] oldest sample in flight, [2] ring, [3] planes complete, [4] stop line of a progressive render
};
size_t wf_bytes_per_slot();
// Slots a render of n_px pixels uses: twice the pixels for small images (fewer, fuller iterations: the
// loop is launch-bound there), one per pixel from 2 M pixels on.  Measured on B200 (C3 scene, Mseg/s):
// 256x256 360 / 567 / 550 at 1x / 2x / 4x; 512x512 1074 / 1073 / 984 at 2x / 3x / 4x; 1080p 1460 / 1475 /
// 1360 at 1x / 2x / 3x; 4K 1577 / 1554 at 1x / 1.5x.
inline uint64_t wf_slots_for(uint64_t n_px) {
    const uint64_t twice = 2 * n_px, most = n_px > (1u << 21) ? n_px : (1u << 21);
    return twice < most ? twice : most;
}
int wf_trace_blocks(int device, bool smem_top);
// Host state a context keeps between render calls: the forked stream, the events and the
// instantiated graph of four iterations, re-captured only when a kernel argument changes (another
// scene, other buffers).
struct WfRuntime {
    cudaStream_t side = nullptr, side2 = nullptr;       // forked branches of an iteration: k_wf_exact + plan / sum; k_wf_regen next to k_wf_shade_hit
    cudaEvent_t ev_fork = nullptr, ev_shade = nullptr, ev_join = nullptr, ev_poll = nullptr, ev_regen = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    bool tried_graph = false;           // capture failed once for this key: plain launches from then on
    DScene key_sc{};                    // the arguments the graph was captured with
    WfBuf key_b{};
    double *key_accum = nullptr;
    DCounters *key_ctr = nullptr;
    int key_blocks = 0;
    volatile uint32_t *host_live = nullptr;     // pinned, 6 words: what the host polls (counts[0..3], next item)
    uint64_t slots_override = 0;        // RT_WF_SLOTS (experiments), read once at rt_create
    bool smem_top = false;              // RT_WF_SMEM_TOP=1: the trace kernel stages the top levels of the hierarchy in shared memory
    uint64_t graph_builds = 0;          // diagnostics: how often the graph was (re)instantiated
};
cudaError_t wf_runtime_init(WfRuntime &rt, int device);
void wf_runtime_destroy(WfRuntime &rt);
// Progressive output without draining the pipeline between chunks: the ordered plane sums are held at
// the end of the chunk about to be shown while the paths of the next chunk are already in flight; `show`
// is called (on the host, between graph launches) when the accumulation buffer holds exactly `samples`
// samples of every pixel.  It returns 0 to go on, > 0 to stop the render there, < 0 on error.
// `prepare` runs first and only ENQUEUES (tonemap, copy to the host); the launcher then queues the next
// group of iterations before `show` waits for the image, so the device never runs dry.
struct WfProgress {
    uint32_t chunk = 0;                 // samples per chunk (0: no progressive output)
    int (*prepare)(void *user, uint32_t samples) = nullptr;
    int (*show)(void *user, uint32_t samples) = nullptr;
    void *user = nullptr;
    mutable uint32_t samples_done = 0;  // out: samples in the buffer when the call returned
};
cudaError_t launch_render_wavefront(cudaStream_t st, WfRuntime &rt, const DScene &sc, uint32_t sample_begin, uint32_t sample_count,
                                    uint64_t seed, double *accum, DCounters *ctr, const WfBuf &b, int trace_blocks,
                                    uint64_t *launches, DRegion rg, const WfProgress *progress = nullptr);

cudaError_t launch_postprocess(cudaStream_t st, uint64_t n, const double *accum, uint32_t samples, double *rgb);
cudaError_t launch_rng_probe(cudaStream_t st, uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n,
                             unsigned long long *u64_out, double *probe8);

}  // namespace rt
