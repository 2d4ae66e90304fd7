// comm.cpp — multi-GPU behind the C ABI (include/rt_b200.h, "multi-GPU").
//
// The reference's generate_image (reference src/main.rs:85-104) sums `samples` paths per pixel
// and tonemaps the mean.  Paths are independent, so GPUs split the SAMPLE RANGE: every device
// renders a disjoint block of global sample indices of every pixel into its own f64 accumulation
// buffer, one ncclReduce(sum) over NVLink brings the buffers to the root BEFORE the non-linear
// tonemap (main.rs:104), and only the root postprocesses.  Two ways in:
//   * one process per GPU (torchrun, MPI, ...): rt_comm_unique_id on rank 0, the launcher
//     broadcasts the 128 bytes, every rank calls rt_comm_init on its context, then rt_reduce_accum;
//   * one process, all GPUs (the CLI): rt_group_* — ncclCommInitAll, one host thread per device
//     while rendering, the same reduce.
// NCCL is resolved at run time (dlopen "libnccl.so.2", or the path in RT_NCCL_LIB): a single-GPU
// user needs no NCCL at all.  A process that will ALSO load PyTorch must point RT_NCCL_LIB at torch's
// bundled copy (the Python binding does): two libraries of one soname cannot coexist, and the first
// one loaded wins.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx.hpp"

using namespace rt;

static_assert(sizeof(ncclUniqueId) == RT_COMM_ID_BYTES, "rt_b200.h promises 128 bytes");

namespace {

struct Nccl {
    void *handle = nullptr;
    std::string why;                    // why it is unavailable
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

const Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, []() {
        const char *names[3] = {std::getenv("RT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *name : names) {
            if (!name || !*name) continue;
            n.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.handle) break;
            n.why = dlerror();
        }
        if (!n.handle) return;
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(n.handle, name);
            if (!p) { ok = false; n.why = std::string("missing symbol ") + name; }
            return p;
        };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.Reduce = (decltype(n.Reduce))sym("ncclReduce");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
        n.GetVersion = (decltype(n.GetVersion))sym("ncclGetVersion");
        if (!ok) { dlclose(n.handle); n.handle = nullptr; }
    });
    return n;
}

int need_nccl(rt_ctx *ctx) {
    if (!nccl().handle) return set_err(ctx, RT_ERR_COMM, "NCCL is not available (libnccl.so.2: " + nccl().why + ")");
    return RT_OK;
}
#define CKN(ctx, call)                                                                                  \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess)                                                                          \
            return set_err(ctx, RT_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r_));   \
    } while (0)

}  // namespace

struct rt_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

void rt::comm_release(rt_ctx *ctx) {
    if (!ctx || !ctx->comm) return;
    if (ctx->comm->comm && nccl().handle) nccl().CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
}

struct rt_group {
    std::vector<rt_ctx *> ctx;          // ctx[0] is the root: it receives the reduce and tonemaps
    std::string err;
};

namespace {

int group_err(rt_group *g, int code, const std::string &msg) {
    if (g) g->err = msg;
    g_last_error = msg;
    return code;
}
// fn(i) on one host thread per device (the render call blocks its thread while it polls the
// device); returns the first non-zero status, with that context's message as the group's
template <class F>
int for_each_device(rt_group *g, F &&fn) {
    const size_t n = g->ctx.size();
    std::vector<int> rc(n, RT_OK);
    if (n == 1) {
        rc[0] = fn(0);
    } else {
        std::vector<std::thread> ts;
        ts.reserve(n);
        try {
            for (size_t i = 0; i < n; ++i) ts.emplace_back([&, i]() { rc[i] = fn(i); });
        } catch (...) {
            for (auto &t : ts) t.join();
            return group_err(g, RT_ERR_LIMIT, "cannot start a host thread per device");
        }
        for (auto &t : ts) t.join();
    }
    for (size_t i = 0; i < n; ++i)
        if (rc[i] != RT_OK) return group_err(g, rc[i], "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err);
    return RT_OK;
}
// sum-reduce of the accumulation buffers onto ctx[0], every device on its own stream
int group_reduce(rt_group *g) {
    if (g->ctx.size() < 2) return RT_OK;
    const Nccl &n = nccl();
    ncclResult_t r = n.GroupStart();
    for (size_t i = 0; i < g->ctx.size() && r == ncclSuccess; ++i) {
        rt_ctx *c = g->ctx[i];
        r = n.Reduce(c->d_accum, c->d_accum, c->n_px * 3, ncclFloat64, ncclSum, 0, c->comm->comm, c->stream);
    }
    const ncclResult_t e = n.GroupEnd();
    if (r == ncclSuccess) r = e;
    if (r != ncclSuccess) return group_err(g, RT_ERR_COMM, std::string("ncclReduce: ") + n.GetErrorString(r));
    return RT_OK;
}

}  // namespace

extern "C" {

int rt_shard_sample_range(int rank, int world, uint32_t total_samples, uint32_t first_sample, uint32_t *begin, uint32_t *count) {
    if (!begin || !count || world < 1 || rank < 0 || rank >= world) return set_err(nullptr, RT_ERR_INVALID_ARG, "rank out of range");
    const uint32_t base = total_samples / (uint32_t)world, extra = total_samples % (uint32_t)world;
    *count = base + ((uint32_t)rank < extra ? 1u : 0u);
    *begin = first_sample + (uint32_t)rank * base + ((uint32_t)rank < extra ? (uint32_t)rank : extra);
    return RT_OK;
}

int rt_comm_unique_id(void *id128) {
    if (!id128) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    if (int r = need_nccl(nullptr)) return r;
    ncclUniqueId id;
    CKN(nullptr, nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return RT_OK;
}
int rt_comm_init(rt_ctx *ctx, const void *id128, int rank, int world) {
    if (!ctx || !id128) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    if (world < 1 || rank < 0 || rank >= world) return set_err(ctx, RT_ERR_INVALID_ARG, "rank out of range");
    if (ctx->comm) return set_err(ctx, RT_ERR_INVALID_ARG, "the context already belongs to a communicator");
    if (int r = need_nccl(ctx)) return r;
    CK(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    std::unique_ptr<rt_comm> c(new rt_comm());
    c->rank = rank; c->world = world;
    CKN(ctx, nccl().CommInitRank(&c->comm, world, id, rank));
    ctx->comm = c.release();
    return RT_OK;
}
int rt_comm_info(const rt_ctx *ctx, int *rank, int *world, int *nccl_version) {
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "null context");
    if (rank) *rank = ctx->comm ? ctx->comm->rank : 0;
    if (world) *world = ctx->comm ? ctx->comm->world : 1;
    if (nccl_version) {
        *nccl_version = 0;
        if (ctx->comm && nccl().handle) nccl().GetVersion(nccl_version);
    }
    return RT_OK;
}
int rt_reduce_accum(rt_ctx *ctx, int root) {
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "null context");
    if (!ctx->has_scene) return set_err(ctx, RT_ERR_NO_SCENE, "no scene uploaded");
    if (!ctx->comm) {                   // a context on its own: the sum over one buffer is the buffer
        if (root != 0) return set_err(ctx, RT_ERR_INVALID_ARG, "root out of range");
        return RT_OK;
    }
    if (root < 0 || root >= ctx->comm->world) return set_err(ctx, RT_ERR_INVALID_ARG, "root out of range");
    CK(ctx, cudaSetDevice(ctx->device));
    // in place; f64 like the buffer (the reference accumulates in f64, src/main.rs:100-103); enqueued on the
    // context's stream, i.e. after the render kernels and before a following rt_postprocess
    CKN(ctx, nccl().Reduce(ctx->d_accum, ctx->d_accum, ctx->n_px * 3, ncclFloat64, ncclSum, root, ctx->comm->comm, ctx->stream));
    return RT_OK;
}

// ---- one process, several devices -------------------------------------------------------------
int rt_group_create(int n_devices, const int *device_ids, rt_group **out) {
    if (!out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    int have = 0;
    if (int r = rt_device_count(&have)) return r;
    if (have == 0) return set_err(nullptr, RT_ERR_NO_DEVICE, "no CUDA device: the render path has no CPU fallback");
    std::vector<int> ids;
    if (n_devices <= 0) {
        for (int i = 0; i < have; ++i) ids.push_back(i);
    } else {
        if (!device_ids) return set_err(nullptr, RT_ERR_INVALID_ARG, "null device list");
        ids.assign(device_ids, device_ids + n_devices);
        for (size_t i = 0; i < ids.size(); ++i)
            for (size_t j = 0; j < i; ++j)
                if (ids[i] == ids[j]) return set_err(nullptr, RT_ERR_INVALID_ARG, "device listed twice");
    }
    std::unique_ptr<rt_group> g(new rt_group());
    auto fail = [&](int code) {
        for (rt_ctx *c : g->ctx) rt_destroy(c);
        return code;
    };
    for (int id : ids) {
        rt_ctx *c = nullptr;
        if (int r = rt_create(id, &c)) return fail(r);
        g->ctx.push_back(c);
    }
    if (ids.size() > 1) {
        if (int r = need_nccl(nullptr)) return fail(r);
        std::vector<ncclComm_t> comms(ids.size(), nullptr);
        const ncclResult_t r = nccl().CommInitAll(comms.data(), (int)ids.size(), ids.data());
        if (r != ncclSuccess) return fail(set_err(nullptr, RT_ERR_COMM, std::string("ncclCommInitAll: ") + nccl().GetErrorString(r)));
        for (size_t i = 0; i < ids.size(); ++i) {
            g->ctx[i]->comm = new rt_comm();
            g->ctx[i]->comm->comm = comms[i];
            g->ctx[i]->comm->rank = (int)i;
            g->ctx[i]->comm->world = (int)ids.size();
        }
    }
    *out = g.release();
    return RT_OK;
}
void rt_group_destroy(rt_group *g) {
    if (!g) return;
    for (rt_ctx *c : g->ctx) rt_destroy(c);
    delete g;
}
int rt_group_size(const rt_group *g, int *n) {
    if (!g || !n) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *n = (int)g->ctx.size();
    return RT_OK;
}
rt_ctx *rt_group_ctx(rt_group *g, int i) { return (g && i >= 0 && (size_t)i < g->ctx.size()) ? g->ctx[i] : nullptr; }
const char *rt_group_last_error(const rt_group *g) { return g ? g->err.c_str() : g_last_error.c_str(); }

int rt_group_upload_scene(rt_group *g, const rt_scene *scene) {
    if (!g || !scene) return group_err(g, RT_ERR_INVALID_ARG, "null argument");
    return for_each_device(g, [&](size_t i) { return rt_upload_scene(g->ctx[i], scene); });
}
int rt_group_render_accum(rt_group *g, const rt_render_params *params) {
    if (!g || !params) return group_err(g, RT_ERR_INVALID_ARG, "null argument");
    const int world = (int)g->ctx.size();
    return for_each_device(g, [&](size_t i) {
        rt_render_params p = *params;
        rt_shard_sample_range((int)i, world, params->sample_count, params->sample_begin, &p.sample_begin, &p.sample_count);
        // The reduce ADDS the buffers: what the image already holds stays on the root only, every
        // other device starts its share from zero.
        if (i != 0) p.clear = 1;
        return rt_render_accum(g->ctx[i], &p);
    });
}
int rt_group_reduce_accum(rt_group *g) {
    if (!g) return group_err(g, RT_ERR_INVALID_ARG, "null group");
    return group_reduce(g);
}
int rt_group_render(rt_group *g, uint64_t seed, int mode, int sched, double *rgb_out) {
    if (!g || g->ctx.empty()) return group_err(g, RT_ERR_INVALID_ARG, "null group");
    rt_ctx *root = g->ctx[0];
    if (!root->has_scene) return group_err(g, RT_ERR_NO_SCENE, "no scene uploaded");
    rt_render_params p{};
    p.sample_begin = 0; p.sample_count = root->samples;
    p.seed = seed; p.mode = mode; p.sched = sched; p.clear = 1;
    if (int r = rt_group_render_accum(g, &p)) return r;
    if (int r = group_reduce(g)) return r;
    if (int r = rt_postprocess(root, root->samples, rgb_out)) return group_err(g, r, root->err);
    if (!rgb_out) return RT_OK;
    // the other devices' reduce must not outlive the call either
    return for_each_device(g, [&](size_t i) { return i ? rt_sync(g->ctx[i]) : (int)RT_OK; });
}
int rt_group_render_progressive(rt_group *g, uint64_t seed, int mode, int sched, uint32_t chunk_samples, rt_progress_fn progress,
                                void *user, double *rgb_out, uint32_t *samples_done) {
    if (!g || g->ctx.empty()) return group_err(g, RT_ERR_INVALID_ARG, "null group");
    rt_ctx *root = g->ctx[0];
    if (!root->has_scene) return group_err(g, RT_ERR_NO_SCENE, "no scene uploaded");
    if (!rgb_out) return group_err(g, RT_ERR_INVALID_ARG, "null argument");
    if (chunk_samples == 0) return group_err(g, RT_ERR_INVALID_ARG, "chunk_samples must be positive");
    const uint32_t total = root->samples;
    uint32_t done = 0;
    if (samples_done) *samples_done = 0;
    do {
        rt_render_params p{};
        p.sample_begin = done;
        p.sample_count = total - done < chunk_samples ? total - done : chunk_samples;
        p.seed = seed; p.mode = mode; p.sched = sched; p.clear = done == 0 ? 1 : 0;
        if (int r = rt_group_render_accum(g, &p)) return r;
        if (int r = group_reduce(g)) return r;      // the root's buffer now holds every sample done so far
        done += p.sample_count;
        if (int r = rt_postprocess(root, done ? done : total, rgb_out)) return group_err(g, r, root->err);
        if (samples_done) *samples_done = done;
        if (progress && progress(user, done, total, rgb_out) != 0) break;
    } while (done < total);
    return for_each_device(g, [&](size_t i) { return i ? rt_sync(g->ctx[i]) : (int)RT_OK; });
}

}  // extern "C"
