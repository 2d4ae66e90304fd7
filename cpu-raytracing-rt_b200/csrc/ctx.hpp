// ctx.hpp — private: the device context behind rt_ctx, shared by capi.cpp (single device) and
// comm.cpp (NCCL reduce, device groups).  Nothing here is part of the C ABI.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/rt_b200.h"
#include "kernels.hpp"

struct rt_comm;                         // comm.cpp: the NCCL communicator of a context (rank, world, ncclComm_t)

struct rt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    char *arena = nullptr;              // one device allocation holding every scene record
    size_t arena_cap = 0, arena_used = 0;
    rt::DScene dsc{};
    bool has_scene = false;
    bool fast_ok = true;                // false: a hierarchy of the scene is too deep for the fp32 traversal -> EXACT answers FAST requests
    double *d_accum = nullptr, *d_rgb = nullptr;
    size_t n_px = 0;
    rt::DCounters *d_ctr = nullptr;
    uint64_t launches = 0;
    uint32_t samples = 0;
    size_t scene_bytes = 0;
    // wavefront scheduler state (allocated on first use, sized by the image)
    char *wf_arena = nullptr;
    rt::WfBuf wf{};
    rt::WfRuntime wf_rt;                // side stream, events, the instantiated iteration graph
    int wf_blocks = 0;
    // experiment knobs, read ONCE at rt_create (never on the render path)
    uint64_t knob_cap = 0;
    uint32_t knob_ring = rt::kWfRing, knob_batch = 4;
    cudaStream_t copy_stream = nullptr; // progressive output: the image of a chunk goes to the host while the next chunk renders
    cudaEvent_t ev_copy = nullptr;
    double *h_rgb = nullptr;            // page-locked staging image of that copy (a copy into pageable memory would not overlap anything)
    size_t h_rgb_px = 0;
    rt_comm *comm = nullptr;            // set by rt_comm_init / rt_group_create, torn down in rt_destroy
};

namespace rt {

extern thread_local std::string g_last_error;

inline int set_err(rt_ctx *ctx, int code, const std::string &msg) {
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}
#define CK(ctx, call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return rt::set_err(ctx, RT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
    } while (0)

void comm_release(rt_ctx *ctx);         // comm.cpp: destroys ctx->comm (called by rt_destroy)

}  // namespace rt
