// capi.cpp — the C ABI of include/rt_b200.h: scene objects, device context, upload, launches.
// No C++ exception and no CUDA error crosses the boundary: every entry point returns rt_status.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"
#include "exact_limits.hpp"
#include "fast_build.hpp"
#include "kernels.hpp"
#include "scene_host.hpp"

using namespace rt;

#include "ctx.hpp"

thread_local std::string rt::g_last_error;

namespace {

template <class F>
int guarded(rt_ctx *ctx, int code_on_throw, F &&f) {
    try {
        return f();
    } catch (const std::bad_alloc &) {
        return set_err(ctx, RT_ERR_LIMIT, "out of host memory");
    } catch (const std::exception &e) {
        return set_err(ctx, code_on_throw, e.what());
    } catch (...) {
        return set_err(ctx, code_on_throw, "unknown error");
    }
}

void free_wavefront(rt_ctx *ctx) {
    if (ctx->wf_arena) cudaFree(ctx->wf_arena);
    ctx->wf_arena = nullptr;
    ctx->wf = WfBuf{};
}
cudaError_t ensure_wavefront(rt_ctx *ctx) {
    // Slots are not tied to pixels (wavefront.cu): small images still get enough slots to fill the
    // GPU, large ones one slot per pixel.  RT_WF_CAP overrides the capacity (experiments).
    const uint32_t n_img = (uint32_t)ctx->n_px;
    uint64_t cap = wf_slots_for(n_img);
    if (ctx->knob_cap) cap = ctx->knob_cap;
    const uint32_t n = (uint32_t)cap;
    if (ctx->wf_arena && ctx->wf.n == n && ctx->wf.n_px_cap == n_img) return cudaSuccess;
    free_wavefront(ctx);
    const size_t per = wf_bytes_per_slot();
    const uint32_t ring = ctx->knob_ring, batch = ctx->knob_batch;
    const size_t total = per * n + (size_t)n_img * 32 * ring + 65536 + 64 * 256;
    cudaError_t e = cudaMalloc(&ctx->wf_arena, total);
    if (e != cudaSuccess) return e;
    char *p = ctx->wf_arena;
    auto take = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    WfBuf &b = ctx->wf;
    b.n = n;
    b.head = (WfHead *)take((size_t)n * sizeof(WfHead));
    b.ray = (WfRay *)take((size_t)n * sizeof(WfRay));
    b.tl = (WfTL *)take((size_t)n * sizeof(WfTL));
    b.pend = (WfPend *)take((size_t)n * sizeof(WfPend));
    b.res = (WfRes *)take((size_t)n * sizeof(WfRes));
    b.cand = (WfCand *)take((size_t)n * sizeof(WfCand));
    b.live = (uint32_t *)take((size_t)n * 4);
    b.hit_list = (uint32_t *)take((size_t)n * 4);
    b.regen_list[0] = (uint32_t *)take((size_t)n * 4);
    b.regen_list[1] = (uint32_t *)take((size_t)n * 4);
    b.xlist = (uint32_t *)take((size_t)n * 4);
    b.counts = (uint32_t *)take(32);
    b.ring = ring;
    b.batch = batch;
    b.side = (double4 *)take((size_t)n_img * 32 * ring);
    b.n_px_cap = n_img;
    b.item = (unsigned long long *)take(32);
    b.plane = (uint32_t *)take(32);
    b.call = (WfCall *)take(sizeof(WfCall));
    if ((size_t)(p - ctx->wf_arena) > total) return cudaErrorMemoryAllocation;
    if (!ctx->wf_rt.side && (e = wf_runtime_init(ctx->wf_rt, ctx->device)) != cudaSuccess) return e;
    ctx->wf_blocks = wf_trace_blocks(ctx->device, ctx->wf_rt.smem_top);
    return cudaSuccess;
}

void free_scene(rt_ctx *ctx, bool keep_arena) {
    if (!keep_arena) {
        free_wavefront(ctx);
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_cap = 0;
        if (ctx->d_accum) cudaFree(ctx->d_accum);
        if (ctx->d_rgb) cudaFree(ctx->d_rgb);
        ctx->d_accum = ctx->d_rgb = nullptr;
        ctx->n_px = 0;
    }
    ctx->arena_used = 0;
    ctx->has_scene = false;
}

inline size_t align_up(size_t n) { return (n + 255) & ~(size_t)255; }

template <class T>
cudaError_t upload(rt_ctx *ctx, const std::vector<T> &v, const T **out) {
    *out = nullptr;
    if (v.empty()) return cudaSuccess;
    const size_t bytes = v.size() * sizeof(T);
    if (ctx->arena_used + align_up(bytes) > ctx->arena_cap) return cudaErrorMemoryAllocation;
    char *p = ctx->arena + ctx->arena_used;
    ctx->arena_used += align_up(bytes);
    *out = (const T *)p;
    return cudaMemcpyAsync(p, v.data(), bytes, cudaMemcpyHostToDevice, ctx->stream);
}

std::vector<DNode> flatten_nodes(const std::vector<HostNode> &nodes) {
    std::vector<DNode> out(nodes.size());
    for (size_t i = 0; i < nodes.size(); ++i) {
        const HostNode &n = nodes[i];
        DNode d{};
        for (int k = 0; k < 3; ++k) { d.lo[k] = n.box.lo[k]; d.hi[k] = n.box.hi[k]; }
        if (n.left != RT_NO_CHILD && n.left != i + 1) throw std::runtime_error("BVH is not in pre-order");
        d.right = n.right;
        d.first = n.first;
        d.count = n.count;
        out[i] = d;
    }
    return out;
}
// parent[i] of every node, for hierarchies deeper than the explicit traversal stack (empty otherwise)
std::vector<uint32_t> parents_if_deep(const std::vector<HostNode> &nodes, uint32_t depth) {
    std::vector<uint32_t> out;
    if (depth <= (uint32_t)kExactStackHost) return out;
    out.assign(nodes.size(), 0u);
    for (size_t i = 0; i < nodes.size(); ++i)
        if (nodes[i].left != RT_NO_CHILD) { out[nodes[i].left] = (uint32_t)i; out[nodes[i].right] = (uint32_t)i; }
    return out;
}
std::vector<DShape> flatten_shapes(const std::vector<HostShape> &v) {
    std::vector<DShape> out(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
        const HostShape &s = v[i];
        DShape d{};
        d.param[0] = s.param.x; d.param[1] = s.param.y; d.param[2] = s.param.z;
        d.position[0] = s.position.x; d.position[1] = s.position.y; d.position[2] = s.position.z;
        d.rot[0] = s.rotation.v.x; d.rot[1] = s.rotation.v.y; d.rot[2] = s.rotation.v.z; d.rot[3] = s.rotation.s;
        d.type = s.type;
        d.material = s.material;
        out[i] = d;
    }
    return out;
}
void flatten_tris(const std::vector<HostTri> &v, std::vector<DTriI> &ti, std::vector<DTriS> &ts) {
    ti.resize(v.size());
    ts.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
        const HostTri &t = v[i];
        DTriI a{};
        a.a[0] = t.a.x; a.a[1] = t.a.y; a.a[2] = t.a.z;
        a.ba[0] = t.ba.x; a.ba[1] = t.ba.y; a.ba[2] = t.ba.z;
        a.ca[0] = t.ca.x; a.ca[1] = t.ca.y; a.ca[2] = t.ca.z;
        a.material = t.material;
        ti[i] = a;
        DTriS s{};
        s.ng[0] = t.ng.x; s.ng[1] = t.ng.y; s.ng[2] = t.ng.z;
        s.na[0] = t.na.x; s.na[1] = t.na.y; s.na[2] = t.na.z;
        s.nb[0] = t.nb.x; s.nb[1] = t.nb.y; s.nb[2] = t.nb.z;
        s.nc[0] = t.nc.x; s.nc[1] = t.nc.y; s.nc[2] = t.nc.z;
        s.inv_area = t.inv_area;
        s.material = t.material;
        ts[i] = s;
    }
}

// Everything rt_upload_scene copies, flattened once per rt_scene and kept page-locked so that
// repeated uploads (one per generate_image call) are plain DMA transfers.
}  // namespace
struct rt_device_image {
    // parent: filled only for a hierarchy deeper than the traversal stack (stackless walk, exact.cuh)
    struct ShapeSet { std::vector<DNode> nodes; std::vector<DShape> prims; std::vector<uint32_t> parent; };
    struct TriSet { std::vector<DNode> nodes; std::vector<DTriI> ti; std::vector<DTriS> ts; std::vector<uint32_t> parent; FastHierarchy fh; };
    bool fast_ok = true;                // false: some hierarchy is too deep for the fp32 traversal's stack -> EXACT answers FAST requests
    ShapeSet shapes[4];                 // boxes, ellipsoids, light boxes, light ellipsoids
    std::vector<DShape> planes;
    TriSet tris[2];                     // scene, lights
    std::vector<DMaterial> mats;
    std::vector<std::pair<void *, size_t>> pinned;
    size_t total_bytes = 0;
    template <class T>
    void pin(std::vector<T> &v) {
        total_bytes += align_up(v.size() * sizeof(T));
        if (v.empty()) return;
        if (cudaHostRegister(v.data(), v.size() * sizeof(T), cudaHostRegisterDefault) == cudaSuccess) pinned.emplace_back(v.data(), v.size() * sizeof(T));
        else cudaGetLastError();
    }
    ~rt_device_image() { for (auto &p : pinned) cudaHostUnregister(p.first); }
};
namespace {
std::shared_ptr<rt_device_image> make_image(const HostScene &s) {
    auto im = std::make_shared<rt_device_image>();
    const HostBvh<HostShape> *sh[4] = {&s.boxes, &s.ellipsoids, &s.light_boxes, &s.light_ellipsoids};
    for (int k = 0; k < 4; ++k) {
        im->shapes[k].nodes = flatten_nodes(sh[k]->nodes);
        im->shapes[k].prims = flatten_shapes(sh[k]->prims);
        im->shapes[k].parent = parents_if_deep(sh[k]->nodes, sh[k]->depth);
        if (!im->shapes[k].parent.empty()) im->fast_ok = false;
    }
    im->planes = flatten_shapes(s.planes);
    const HostBvh<HostTri> *tr[2] = {&s.tris, &s.light_tris};
    for (int k = 0; k < 2; ++k) {
        im->tris[k].nodes = flatten_nodes(tr[k]->nodes);
        flatten_tris(tr[k]->prims, im->tris[k].ti, im->tris[k].ts);
        im->tris[k].parent = parents_if_deep(tr[k]->nodes, tr[k]->depth);
        // The fp32 hierarchy keeps up to three deferred children per 4-wide level on a 64-entry stack: a
        // chain-like tree overflows it long before the exact walk needs its parent links.  Such a scene
        // is rendered by the exact traversal (same image bit for bit, rt_render_accum), so the fp32
        // structure is not built at all — its builder recurses once per level.
        if (!im->tris[k].parent.empty() || tr[k]->depth + 2 > (uint32_t)kFastStackHost) im->fast_ok = false;
    }
    for (int k = 0; k < 2 && im->fast_ok; ++k) {
        build_fast_hierarchy(*tr[k], im->tris[k].fh);
        if (im->tris[k].fh.depth + 2 > (uint32_t)kFastStackHost) im->fast_ok = false;
    }
    if (!im->fast_ok)
        for (int k = 0; k < 2; ++k) im->tris[k].fh = FastHierarchy{};
    im->mats.resize(s.materials.size());
    for (size_t i = 0; i < im->mats.size(); ++i) {
        DMaterial m{};
        for (int k = 0; k < 3; ++k) { m.color[k] = s.materials[i].color[k]; m.emission[k] = s.materials[i].emission[k]; }
        m.ior = s.materials[i].ior;
        m.kind = s.materials[i].kind;
        im->mats[i] = m;
    }
    int ndev = 0;
    const bool cuda_ok = cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0;
    if (!cuda_ok) cudaGetLastError();
    for (int k = 0; k < 4; ++k) { im->pin(im->shapes[k].nodes); im->pin(im->shapes[k].prims); im->pin(im->shapes[k].parent); }
    im->pin(im->planes);
    for (int k = 0; k < 2; ++k) { im->pin(im->tris[k].parent); im->pin(im->tris[k].nodes); im->pin(im->tris[k].ti); im->pin(im->tris[k].ts); im->pin(im->tris[k].fh.nodes); im->pin(im->tris[k].fh.tris); }
    im->pin(im->mats);
    return im;
}

int need_scene(rt_ctx *ctx) {
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "null context");
    if (!ctx->has_scene) return set_err(ctx, RT_ERR_NO_SCENE, "no scene uploaded");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return set_err(ctx, RT_ERR_CUDA, cudaGetErrorString(e));
    return RT_OK;
}
int check_mode(rt_ctx *ctx, int mode) {
    if (mode != RT_MODE_EXACT && mode != RT_MODE_FAST) return set_err(ctx, RT_ERR_INVALID_ARG, "unknown mode");
    return RT_OK;
}
// FAST gives the EXACT results bit for bit; where its fp32 structures cannot be used — a hierarchy
// deeper than the fp32 traversal's stack (the reference's sweep SAH has no depth bound,
// src/bvh.rs:93-111), coordinates beyond the fp32 range (fast.cuh, kFastRange) — the exact traversal
// answers the request instead of an error.
bool fast_usable(const rt_ctx *ctx) {
    return ctx->fast_ok && (ctx->dsc.tris.n_prims == 0 || ctx->dsc.ftris.bound <= kFastRange);
}

struct DevBuf {                         // scoped device allocation for the array-in / array-out calls
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
};

}  // namespace

template <class P>
static void export_bvh(const HostBvh<P> &b, double *bounds, uint32_t *left, uint32_t *right, uint32_t *first,
                       uint32_t *count, uint32_t *order) {
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        const HostNode &n = b.nodes[i];
        if (bounds) for (int k = 0; k < 3; ++k) { bounds[i * 6 + k] = n.box.lo[k]; bounds[i * 6 + 3 + k] = n.box.hi[k]; }
        if (left) left[i] = n.left;
        if (right) right[i] = n.right;
        if (first) first[i] = n.first;
        if (count) count[i] = n.count;
    }
    if (order) for (size_t i = 0; i < b.prims.size(); ++i) order[i] = b.prims[i].orig;
}
namespace {
template <class P>
void import_nodes(const rt_built_bvh &b, HostBvh<P> &out) {
    if (b.n_nodes && (!b.bounds || !b.left || !b.right || !b.first || !b.count)) throw std::runtime_error("null node array");
    if (b.n_nodes >= RT_NO_CHILD || b.n_prims >= RT_NO_CHILD) throw std::runtime_error("hierarchy too large for 32-bit indices");
    out.nodes.resize(b.n_nodes);
    for (uint64_t i = 0; i < b.n_nodes; ++i) {
        HostNode n{};
        for (int k = 0; k < 3; ++k) { n.box.lo[k] = b.bounds[i * 6 + k]; n.box.hi[k] = b.bounds[i * 6 + 3 + k]; }
        n.left = b.left[i]; n.right = b.right[i]; n.first = b.first[i]; n.count = b.count[i];
        out.nodes[i] = n;
    }
    // pre-order numbering, leaf ranges, one parent per node (a shared child would hide the real depth
    // from the stack-size check of rt_upload_scene)
    out.depth = validate_hierarchy(out.nodes, b.n_prims, "");
}
}  // namespace

// ---- host side -----------------------------------------------------------------------------------
extern "C" {

int rt_scene_build(const rt_raw_scene *raw, rt_scene **out) {
    if (!raw || !out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    return guarded(nullptr, RT_ERR_INVALID_ARG, [&]() {
        std::unique_ptr<rt_scene> s(new rt_scene());
        build_scene(*raw, s->s);
        *out = s.release();
        return (int)RT_OK;
    });
}
void rt_scene_free(rt_scene *scene) { delete scene; }

int rt_scene_load_text(const char *path, rt_scene **out) {
    if (!out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    return guarded(nullptr, RT_ERR_PARSE, [&]() {
        RawSceneStorage st;
        if (path) {
            std::ifstream f(path);
            if (!f) return set_err(nullptr, RT_ERR_IO, std::string("cannot open '") + path + "'");
            parse_text_scene(f, st);
        } else {
            parse_text_scene(std::cin, st);
        }
        std::unique_ptr<rt_scene> s(new rt_scene());
        build_scene(st.raw, s->s);
        *out = s.release();
        return (int)RT_OK;
    });
}
int rt_scene_load_gltf(const char *path, uint32_t width, uint32_t height, uint32_t samples, rt_scene **out) {
    if (!out || !path) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    return guarded(nullptr, RT_ERR_PARSE, [&]() {
        RawSceneStorage st;
        load_gltf_scene(path, width, height, samples, st);
        std::unique_ptr<rt_scene> s(new rt_scene());
        build_scene(st.raw, s->s);
        *out = s.release();
        return (int)RT_OK;
    });
}

int rt_scene_build_cached(const rt_raw_scene *raw, const char *cache_dir, rt_scene **out, int *cache_hit) {
    if (!raw || !out || !cache_dir) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cache_hit) *cache_hit = 0;
    return guarded(nullptr, RT_ERR_IO, [&]() {
        std::unique_ptr<rt_scene> s(new rt_scene());
        const bool hit = build_scene_cached(*raw, cache_dir, s->s);
        if (cache_hit) *cache_hit = hit ? 1 : 0;
        *out = s.release();
        return (int)RT_OK;
    });
}
int rt_scene_load_gltf_cached(const char *path, uint32_t width, uint32_t height, uint32_t samples, const char *cache_dir,
                              rt_scene **out, int *cache_hit) {
    if (!out || !path || !cache_dir) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (cache_hit) *cache_hit = 0;
    RawSceneStorage st;
    const int rc = guarded(nullptr, RT_ERR_PARSE, [&]() {
        load_gltf_scene(path, width, height, samples, st);
        return (int)RT_OK;
    });
    if (rc != RT_OK) return rc;
    return rt_scene_build_cached(&st.raw, cache_dir, out, cache_hit);
}

int rt_scene_from_built(const rt_built_scene *built, rt_scene **out) {
    if (!built || !out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    return guarded(nullptr, RT_ERR_INVALID_ARG, [&]() {
        const rt_built_scene &b = *built;
        if (b.width == 0 || b.height == 0) throw std::runtime_error("zero image dimensions");
        if (b.ray_depth > 255) throw std::runtime_error("ray_depth is a u8 in the reference");
        if ((b.n_materials && !b.materials) || (b.n_planes && !b.planes)) throw std::runtime_error("null array");
        std::unique_ptr<rt_scene> sc(new rt_scene());
        HostScene &s = sc->s;
        s.width = b.width; s.height = b.height; s.samples = b.samples; s.ray_depth = b.ray_depth;
        auto v3 = [](const double *p) { return mk3(p[0], p[1], p[2]); };
        s.bg_color = v3(b.bg_color);
        s.cam_position = v3(b.cam_position); s.cam_right = v3(b.cam_right); s.cam_up = v3(b.cam_up); s.cam_forward = v3(b.cam_forward);
        s.tan_half_fov_x = b.tan_half_fov_x; s.tan_half_fov_y = b.tan_half_fov_y;
        s.materials.assign(b.materials, b.materials + b.n_materials);
        auto shape_of = [&](const rt_shape &r, uint32_t orig, const double *box) {
            if (r.material >= b.n_materials) throw std::runtime_error("material index out of range");
            HostShape h{};
            h.type = r.type; h.material = r.material;
            h.param = v3(r.param); h.position = v3(r.position);
            h.rotation = mkq_xyzw(r.rotation[0], r.rotation[1], r.rotation[2], r.rotation[3]);
            h.orig = orig;
            if (box) for (int k = 0; k < 3; ++k) { h.box.lo[k] = box[k]; h.box.hi[k] = box[3 + k]; }
            return h;
        };
        for (uint32_t i = 0; i < b.n_planes; ++i) s.planes.push_back(shape_of(b.planes[i], i, nullptr));
        HostBvh<HostShape> *shape_sets[6] = {&s.boxes, &s.ellipsoids, nullptr, &s.light_boxes, &s.light_ellipsoids, nullptr};
        HostBvh<HostTri> *tri_sets[6] = {nullptr, nullptr, &s.tris, nullptr, nullptr, &s.light_tris};
        for (int w = 0; w < 6; ++w) {
            const rt_built_bvh &bb = b.bvh[w];
            if (shape_sets[w]) {
                if (bb.n_prims && !bb.shapes) throw std::runtime_error("null shape array");
                import_nodes(bb, *shape_sets[w]);
                for (uint64_t i = 0; i < bb.n_prims; ++i) shape_sets[w]->prims.push_back(shape_of(bb.shapes[i], (uint32_t)i, bb.shape_aabb ? bb.shape_aabb + i * 6 : nullptr));
            } else {
                if (bb.n_prims && (!bb.tri28 || !bb.tri_material)) throw std::runtime_error("null triangle array");
                import_nodes(bb, *tri_sets[w]);
                tri_sets[w]->prims.resize(bb.n_prims);
                for (uint64_t i = 0; i < bb.n_prims; ++i) {
                    const double *t = bb.tri28 + i * 28;
                    HostTri h{};
                    h.a = v3(t); h.ba = v3(t + 3); h.ca = v3(t + 6); h.ng = v3(t + 9);
                    h.na = v3(t + 12); h.nb = v3(t + 15); h.nc = v3(t + 18);
                    h.inv_area = t[21];
                    for (int k = 0; k < 3; ++k) { h.box.lo[k] = t[22 + k]; h.box.hi[k] = t[25 + k]; }
                    h.material = bb.tri_material[i];
                    if (h.material >= b.n_materials) throw std::runtime_error("material index out of range");
                    h.orig = (uint32_t)i;
                    tri_sets[w]->prims[i] = h;
                }
            }
        }
        *out = sc.release();
        return (int)RT_OK;
    });
}

int rt_scene_get_info(const rt_scene *scene, rt_scene_info *info) {
    if (!scene || !info) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    const HostScene &s = scene->s;
    *info = rt_scene_info{};
    info->width = s.width; info->height = s.height; info->samples = s.samples; info->ray_depth = s.ray_depth;
    const d3 *src[5] = {&s.bg_color, &s.cam_position, &s.cam_right, &s.cam_up, &s.cam_forward};
    double *dst[5] = {info->bg_color, info->cam_position, info->cam_right, info->cam_up, info->cam_forward};
    for (int i = 0; i < 5; ++i) { dst[i][0] = src[i]->x; dst[i][1] = src[i]->y; dst[i][2] = src[i]->z; }
    info->tan_half_fov_x = s.tan_half_fov_x; info->tan_half_fov_y = s.tan_half_fov_y;
    info->n_planes = (uint32_t)s.planes.size(); info->n_materials = (uint32_t)s.materials.size();
    auto put = [&](int which, size_t nodes, size_t prims, uint32_t depth) {
        info->bvh_nodes[which] = nodes; info->bvh_prims[which] = prims; info->bvh_depth[which] = depth;
    };
    put(RT_BVH_BOXES, s.boxes.nodes.size(), s.boxes.prims.size(), s.boxes.depth);
    put(RT_BVH_ELLIPSOIDS, s.ellipsoids.nodes.size(), s.ellipsoids.prims.size(), s.ellipsoids.depth);
    put(RT_BVH_TRIANGLES, s.tris.nodes.size(), s.tris.prims.size(), s.tris.depth);
    put(RT_BVH_LIGHT_BOXES, s.light_boxes.nodes.size(), s.light_boxes.prims.size(), s.light_boxes.depth);
    put(RT_BVH_LIGHT_ELLIPSOIDS, s.light_ellipsoids.nodes.size(), s.light_ellipsoids.prims.size(), s.light_ellipsoids.depth);
    put(RT_BVH_LIGHT_TRIANGLES, s.light_tris.nodes.size(), s.light_tris.prims.size(), s.light_tris.depth);
    return RT_OK;
}

int rt_scene_get_bvh(const rt_scene *scene, int which, double *bounds, uint32_t *left, uint32_t *right,
                     uint32_t *first, uint32_t *count, uint32_t *prim_order) {
    if (!scene) return set_err(nullptr, RT_ERR_INVALID_ARG, "null scene");
    const HostScene &s = scene->s;
    switch (which) {
        case RT_BVH_BOXES: export_bvh(s.boxes, bounds, left, right, first, count, prim_order); break;
        case RT_BVH_ELLIPSOIDS: export_bvh(s.ellipsoids, bounds, left, right, first, count, prim_order); break;
        case RT_BVH_TRIANGLES: export_bvh(s.tris, bounds, left, right, first, count, prim_order); break;
        case RT_BVH_LIGHT_BOXES: export_bvh(s.light_boxes, bounds, left, right, first, count, prim_order); break;
        case RT_BVH_LIGHT_ELLIPSOIDS: export_bvh(s.light_ellipsoids, bounds, left, right, first, count, prim_order); break;
        case RT_BVH_LIGHT_TRIANGLES: export_bvh(s.light_tris, bounds, left, right, first, count, prim_order); break;
        default: return set_err(nullptr, RT_ERR_INVALID_ARG, "unknown hierarchy");
    }
    return RT_OK;
}
int rt_scene_get_triangle(const rt_scene *scene, int which, uint64_t i, double *o, uint32_t *material) {
    if (!scene || !o) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    const HostBvh<HostTri> *b = which == RT_BVH_TRIANGLES ? &scene->s.tris : (which == RT_BVH_LIGHT_TRIANGLES ? &scene->s.light_tris : nullptr);
    if (!b || i >= b->prims.size()) return set_err(nullptr, RT_ERR_INVALID_ARG, "triangle index out of range");
    const HostTri &t = b->prims[i];
    const d3 vs[7] = {t.a, t.ba, t.ca, t.ng, t.na, t.nb, t.nc};
    for (int k = 0; k < 7; ++k) { o[k * 3] = vs[k].x; o[k * 3 + 1] = vs[k].y; o[k * 3 + 2] = vs[k].z; }
    o[21] = t.inv_area;
    for (int k = 0; k < 3; ++k) { o[22 + k] = t.box.lo[k]; o[25 + k] = t.box.hi[k]; }
    if (material) *material = t.material;
    return RT_OK;
}
int rt_scene_get_shape(const rt_scene *scene, int which, uint64_t i, rt_shape *out, double *aabb6) {
    if (!scene || !out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    const HostScene &s = scene->s;
    const std::vector<HostShape> *v = nullptr;
    switch (which) {
        case -1: v = &s.planes; break;
        case RT_BVH_BOXES: v = &s.boxes.prims; break;
        case RT_BVH_ELLIPSOIDS: v = &s.ellipsoids.prims; break;
        case RT_BVH_LIGHT_BOXES: v = &s.light_boxes.prims; break;
        case RT_BVH_LIGHT_ELLIPSOIDS: v = &s.light_ellipsoids.prims; break;
        default: return set_err(nullptr, RT_ERR_INVALID_ARG, "unknown shape set");
    }
    if (i >= v->size()) return set_err(nullptr, RT_ERR_INVALID_ARG, "shape index out of range");
    const HostShape &p = (*v)[i];
    *out = rt_shape{};
    out->type = p.type; out->material = p.material;
    out->param[0] = p.param.x; out->param[1] = p.param.y; out->param[2] = p.param.z;
    out->position[0] = p.position.x; out->position[1] = p.position.y; out->position[2] = p.position.z;
    out->rotation[0] = p.rotation.v.x; out->rotation[1] = p.rotation.v.y; out->rotation[2] = p.rotation.v.z; out->rotation[3] = p.rotation.s;
    if (aabb6) for (int k = 0; k < 3; ++k) { aabb6[k] = p.box.lo[k]; aabb6[3 + k] = p.box.hi[k]; }
    return RT_OK;
}
int rt_scene_get_material(const rt_scene *scene, uint32_t i, rt_material *out) {
    if (!scene || !out || i >= scene->s.materials.size()) return set_err(nullptr, RT_ERR_INVALID_ARG, "bad material index");
    *out = scene->s.materials[i];
    return RT_OK;
}
int rt_save_ppm(const char *path, uint32_t width, uint32_t height, const double *rgb) {
    if (!path || !rgb) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    return guarded(nullptr, RT_ERR_IO, [&]() { save_ppm(path, width, height, rgb); return (int)RT_OK; });
}

// ---- device side -----------------------------------------------------------------------------------
int rt_device_count(int *count) {
    if (!count) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return set_err(nullptr, RT_ERR_NO_DEVICE, cudaGetErrorString(e)); }
    return RT_OK;
}
int rt_create(int device_id, rt_ctx **out) {
    if (!out) return set_err(nullptr, RT_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return set_err(nullptr, RT_ERR_NO_DEVICE, std::string("no CUDA device: the render path has no CPU fallback (") +
                                                      (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
    }
    if (device_id < 0 || device_id >= n) return set_err(nullptr, RT_ERR_INVALID_ARG, "device id out of range");
    std::unique_ptr<rt_ctx> ctx(new rt_ctx());
    ctx->device = device_id;
    CK(nullptr, cudaSetDevice(device_id));
    CK(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(nullptr, cudaMalloc(&ctx->d_ctr, sizeof(DCounters)));
    CK(nullptr, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DCounters), ctx->stream));
    // experiment knobs of the wavefront scheduler: read here, once, not on the render path
    if (const char *e = std::getenv("RT_WF_CAP")) { const long long v = std::atoll(e); if (v >= 1024 && v <= (1ll << 28)) ctx->knob_cap = (uint64_t)v; }
    if (const char *e = std::getenv("RT_WF_RING")) { const int v = std::atoi(e); if (v >= 8 && v <= 256) ctx->knob_ring = (uint32_t)v; }
    if (const char *e = std::getenv("RT_WF_BATCH")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) ctx->knob_batch = (uint32_t)v; }
    // k_wf_plan only moves on when `batch` planes are complete, and at most ring - reserve (reserve <= 4)
    // planes are ever open: a larger batch would never be reached
    if (ctx->knob_batch > ctx->knob_ring - 4u) ctx->knob_batch = ctx->knob_ring - 4u;
    if (const char *e = std::getenv("RT_WF_SMEM_TOP")) ctx->wf_rt.smem_top = std::atoi(e) != 0;
    if (const char *e = std::getenv("RT_WF_SLOTS")) { const long long v = std::atoll(e); if (v >= 1024) ctx->wf_rt.slots_override = (uint64_t)v; }
    *out = ctx.release();
    return RT_OK;
}
void rt_destroy(rt_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_release(ctx);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
    if (ctx->h_rgb) cudaFreeHost(ctx->h_rgb);
    free_scene(ctx, false);
    if (ctx->d_ctr) cudaFree(ctx->d_ctr);
    wf_runtime_destroy(ctx->wf_rt);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char *rt_last_error(const rt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int rt_upload_scene(rt_ctx *ctx, const rt_scene *scene) {
    if (!ctx || !scene) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    const HostScene &s = scene->s;
    return guarded(ctx, RT_ERR_INVALID_ARG, [&]() -> int {
        CK(ctx, cudaSetDevice(ctx->device));
        // built once per scene, whichever context (thread) uploads first: one host thread per GPU may
        // upload the same rt_scene concurrently (rt_b200.h, "Threading")
        std::call_once(scene->image_once, [&]() { scene->image = make_image(s); });
        if (!scene->image) return set_err(ctx, RT_ERR_LIMIT, "the flattened scene image could not be built");
        const rt_device_image &im = *scene->image;
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        const size_t n_px = (size_t)s.width * s.height;
        const bool keep = ctx->arena_cap >= im.total_bytes && ctx->n_px == n_px;
        free_scene(ctx, keep);
        if (!keep) {
            CK(ctx, cudaMalloc(&ctx->arena, im.total_bytes ? im.total_bytes : 256));
            ctx->arena_cap = im.total_bytes;
            ctx->n_px = n_px;
            CK(ctx, cudaMalloc(&ctx->d_accum, n_px * 3 * sizeof(double)));
            CK(ctx, cudaMalloc(&ctx->d_rgb, n_px * 3 * sizeof(double)));
        }
        DScene d{};
        DBvh *sb[4] = {&d.boxes, &d.ellipsoids, &d.lboxes, &d.lellipsoids};
        const DShape **sp[4] = {&d.box_prims, &d.ell_prims, &d.lbox_prims, &d.lell_prims};
        for (int k = 0; k < 4; ++k) {
            CK(ctx, upload(ctx, im.shapes[k].nodes, &sb[k]->nodes));
            CK(ctx, upload(ctx, im.shapes[k].prims, sp[k]));
            CK(ctx, upload(ctx, im.shapes[k].parent, &sb[k]->parent));     // null unless the hierarchy is deeper than the stack
            sb[k]->n_nodes = (uint32_t)im.shapes[k].nodes.size();
            sb[k]->n_prims = (uint32_t)im.shapes[k].prims.size();
        }
        CK(ctx, upload(ctx, im.planes, &d.planes));
        d.n_planes = (uint32_t)im.planes.size();
        DBvh *tb[2] = {&d.tris, &d.ltris};
        const DTriI **ti[2] = {&d.tri_i, &d.ltri_i};
        const DTriS **ts[2] = {&d.tri_s, &d.ltri_s};
        FBvh *fb[2] = {&d.ftris, &d.fltris};
        for (int k = 0; k < 2; ++k) {
            CK(ctx, upload(ctx, im.tris[k].nodes, &tb[k]->nodes));
            CK(ctx, upload(ctx, im.tris[k].ti, ti[k]));
            CK(ctx, upload(ctx, im.tris[k].ts, ts[k]));
            CK(ctx, upload(ctx, im.tris[k].parent, &tb[k]->parent));
            tb[k]->n_nodes = (uint32_t)im.tris[k].nodes.size();
            tb[k]->n_prims = (uint32_t)im.tris[k].ti.size();
            CK(ctx, upload(ctx, im.tris[k].fh.nodes, &fb[k]->nodes));
            CK(ctx, upload(ctx, im.tris[k].fh.tris, &fb[k]->tris));
            fb[k]->root = im.tris[k].fh.root;
            fb[k]->depth = im.tris[k].fh.depth;
            fb[k]->top = im.tris[k].fh.top;
            double bound = 0.0;
            if (!im.tris[k].nodes.empty())
                for (int a = 0; a < 3; ++a) bound = std::max(bound, std::max(std::fabs(im.tris[k].nodes[0].lo[a]), std::fabs(im.tris[k].nodes[0].hi[a])));
            fb[k]->bound = std::nextafter((float)bound, INFINITY);
        }
        CK(ctx, upload(ctx, im.mats, &d.materials));
        d.n_lights = (uint32_t)(s.light_boxes.prims.size() + s.light_ellipsoids.prims.size() + s.light_tris.prims.size());
        d.width = s.width; d.height = s.height; d.ray_depth = s.ray_depth;
        d.cam.position = s.cam_position; d.cam.right = s.cam_right; d.cam.up = s.cam_up; d.cam.forward = s.cam_forward;
        d.cam.tan_half_fov_x = s.tan_half_fov_x; d.cam.tan_half_fov_y = s.tan_half_fov_y;
        d.cam.width = (double)s.width; d.cam.height = (double)s.height;
        d.bg = s.bg_color;
        CK(ctx, cudaMemsetAsync(ctx->d_accum, 0, n_px * 3 * sizeof(double), ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->dsc = d;
        ctx->fast_ok = im.fast_ok;
        ctx->samples = s.samples;
        ctx->scene_bytes = ctx->arena_used;
        ctx->has_scene = true;
        return RT_OK;
    });
}
int rt_scene_device_bytes(rt_ctx *ctx, uint64_t *bytes) {
    if (int r = need_scene(ctx)) return r;
    if (!bytes) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    *bytes = ctx->scene_bytes;
    return RT_OK;
}

int rt_trace_rays(rt_ctx *ctx, int mode, uint64_t n, const double *origins, const double *dirs, rt_hit *hits) {
    if (int r = need_scene(ctx)) return r;
    if (int r = check_mode(ctx, mode)) return r;
    if (n && (!origins || !dirs || !hits)) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    if (n == 0) return RT_OK;
    DevBuf o, d, h;
    CK(ctx, o.alloc(n * 24));
    CK(ctx, d.alloc(n * 24));
    CK(ctx, h.alloc(n * sizeof(rt_hit)));
    CK(ctx, cudaMemcpyAsync(o.p, origins, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d.p, dirs, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, launch_trace_rays(ctx->stream, ctx->dsc, (mode == RT_MODE_FAST && !ctx->fast_ok) ? (int)RT_MODE_EXACT : mode, n, (const double *)o.p, (const double *)d.p, h.p, ctx->d_ctr));
    ++ctx->launches;
    CK(ctx, cudaMemcpyAsync(hits, h.p, n * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_trace_primary(rt_ctx *ctx, int mode, double jx, double jy, rt_hit *hits, double *rays_out) {
    if (int r = need_scene(ctx)) return r;
    if (int r = check_mode(ctx, mode)) return r;
    DevBuf h, rb;
    if (hits) CK(ctx, h.alloc(ctx->n_px * sizeof(rt_hit)));
    if (rays_out) CK(ctx, rb.alloc(ctx->n_px * 48));
    CK(ctx, launch_trace_primary(ctx->stream, ctx->dsc, (mode == RT_MODE_FAST && !ctx->fast_ok) ? (int)RT_MODE_EXACT : mode, jx, jy, h.p, (double *)rb.p, ctx->d_ctr));
    ++ctx->launches;
    if (hits) CK(ctx, cudaMemcpyAsync(hits, h.p, ctx->n_px * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->stream));
    if (rays_out) CK(ctx, cudaMemcpyAsync(rays_out, rb.p, ctx->n_px * 48, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_light_pdf(rt_ctx *ctx, int mode, uint64_t n, const double *pos, const double *dir, double *pdf) {
    if (int r = need_scene(ctx)) return r;
    if (int r = check_mode(ctx, mode)) return r;
    if (n && (!pos || !dir || !pdf)) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    if (n == 0) return RT_OK;
    DevBuf p, d, o;
    CK(ctx, p.alloc(n * 24));
    CK(ctx, d.alloc(n * 24));
    CK(ctx, o.alloc(n * 8));
    CK(ctx, cudaMemcpyAsync(p.p, pos, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(d.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, launch_light_pdf(ctx->stream, ctx->dsc, (mode == RT_MODE_FAST && !ctx->fast_ok) ? (int)RT_MODE_EXACT : mode, n, (const double *)p.p, (const double *)d.p, (double *)o.p, ctx->d_ctr));
    ++ctx->launches;
    CK(ctx, cudaMemcpyAsync(pdf, o.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}

int rt_render_accum(rt_ctx *ctx, const rt_render_params *p) {
    if (int r = need_scene(ctx)) return r;
    if (!p) return set_err(ctx, RT_ERR_INVALID_ARG, "null params");
    rt_render_params resolved = *p;                     // RT_MODE_AUTO / RT_SCHED_AUTO -> what suits the scene
    const bool has_tris = ctx->dsc.tris.n_prims != 0;
    if (resolved.mode == RT_MODE_AUTO) resolved.mode = has_tris ? RT_MODE_FAST : RT_MODE_EXACT;
    // A scene whose coordinates leave the range of the fp32 culling arithmetic (fast.cuh, kFastRange),
    // or whose hierarchy is too deep for the fp32 traversal's stack, is rendered by the exact
    // traversal whatever was asked for: same image, bit for bit.
    if (resolved.mode == RT_MODE_FAST && !fast_usable(ctx)) {
        resolved.mode = RT_MODE_EXACT;
        resolved.sched = RT_SCHED_MEGAKERNEL;
    }
    if (resolved.sched == RT_SCHED_AUTO) resolved.sched = (resolved.mode == RT_MODE_FAST && has_tris) ? RT_SCHED_WAVEFRONT : RT_SCHED_MEGAKERNEL;
    p = &resolved;
    if (int r = check_mode(ctx, p->mode)) return r;
    if (p->sched != RT_SCHED_MEGAKERNEL && p->sched != RT_SCHED_WAVEFRONT) return set_err(ctx, RT_ERR_INVALID_ARG, "unknown scheduler");
    if ((uint64_t)p->sample_begin + p->sample_count > 0xFFFFFFFFull) return set_err(ctx, RT_ERR_INVALID_ARG, "sample range overflows u32");
    DRegion rg{0u, 0u, ctx->dsc.width, ctx->dsc.height};
    if (p->region[2] != 0 && p->region[3] != 0) {
        rg = DRegion{p->region[0], p->region[1], p->region[2], p->region[3]};
        if ((uint64_t)rg.x0 + rg.w > ctx->dsc.width || (uint64_t)rg.y0 + rg.h > ctx->dsc.height)
            return set_err(ctx, RT_ERR_INVALID_ARG, "render region reaches outside the image");
    }
    if (p->clear) CK(ctx, cudaMemsetAsync(ctx->d_accum, 0, ctx->n_px * 3 * sizeof(double), ctx->stream));
    if (p->sample_count == 0) return RT_OK;
    if (p->sched == RT_SCHED_WAVEFRONT && ctx->dsc.ray_depth > 0) {
        if (p->mode != RT_MODE_FAST) return set_err(ctx, RT_ERR_INVALID_ARG, "the wavefront scheduler runs FAST mode only (EXACT is the megakernel parity anchor)");
        CK(ctx, ensure_wavefront(ctx));
        CK(ctx, launch_render_wavefront(ctx->stream, ctx->wf_rt, ctx->dsc, p->sample_begin, p->sample_count, p->seed, ctx->d_accum,
                                        ctx->d_ctr, ctx->wf, ctx->wf_blocks, &ctx->launches, rg));
        return RT_OK;
    }
    CK(ctx, launch_render_mega(ctx->stream, ctx->dsc, p->mode, p->sample_begin, p->sample_count, p->seed, ctx->d_accum, ctx->d_ctr, rg));
    ++ctx->launches;
    return RT_OK;
}
int rt_accum_device_ptr(rt_ctx *ctx, void **dptr, uint64_t *n_doubles) {
    if (int r = need_scene(ctx)) return r;
    if (dptr) *dptr = ctx->d_accum;
    if (n_doubles) *n_doubles = ctx->n_px * 3;
    return RT_OK;
}
int rt_get_accum(rt_ctx *ctx, double *sum_rgb) {
    if (int r = need_scene(ctx)) return r;
    if (!sum_rgb) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    CK(ctx, cudaMemcpyAsync(sum_rgb, ctx->d_accum, ctx->n_px * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_set_accum(rt_ctx *ctx, const double *sum_rgb) {
    if (int r = need_scene(ctx)) return r;
    if (!sum_rgb) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    CK(ctx, cudaMemcpyAsync(ctx->d_accum, sum_rgb, ctx->n_px * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_postprocess(rt_ctx *ctx, uint32_t total_samples, double *rgb_out) {
    if (int r = need_scene(ctx)) return r;
    CK(ctx, launch_postprocess(ctx->stream, ctx->n_px * 3, ctx->d_accum, total_samples, ctx->d_rgb));
    ++ctx->launches;
    if (!rgb_out) return RT_OK;         // device-only (asynchronous): the image stays in HBM
    CK(ctx, cudaMemcpyAsync(rgb_out, ctx->d_rgb, ctx->n_px * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_render(rt_ctx *ctx, uint64_t seed, int mode, int sched, double *rgb_out) {
    if (int r = need_scene(ctx)) return r;
    rt_render_params p{};
    p.sample_begin = 0;
    p.sample_count = ctx->samples;
    p.seed = seed; p.mode = mode; p.sched = sched; p.clear = 1;
    if (int r = rt_render_accum(ctx, &p)) return r;
    return rt_postprocess(ctx, ctx->samples, rgb_out);
}
// what rt_render_accum would pick for this scene
static void resolve_mode(const rt_ctx *ctx, int &mode, int &sched) {
    const bool has_tris = ctx->dsc.tris.n_prims != 0;
    if (mode == RT_MODE_AUTO) mode = has_tris ? RT_MODE_FAST : RT_MODE_EXACT;
    if (mode == RT_MODE_FAST && !fast_usable(ctx)) { mode = RT_MODE_EXACT; sched = RT_SCHED_MEGAKERNEL; }
    if (sched == RT_SCHED_AUTO) sched = (mode == RT_MODE_FAST && has_tris) ? RT_SCHED_WAVEFRONT : RT_SCHED_MEGAKERNEL;
}
struct ProgressiveShow {                // WfProgress::show: tonemap the samples in the buffer, hand the image to the caller
    rt_ctx *ctx;
    rt_progress_fn progress;
    void *user;
    double *rgb_out;
    uint32_t total;
    uint32_t *samples_done;
    // tonemap on the render stream (behind the launches already queued), copy on a stream of its own: the
    // host later waits for the copy only, the render stream goes on with the next chunk
    static int prepare(void *self, uint32_t samples) {
        ProgressiveShow *p = (ProgressiveShow *)self;
        rt_ctx *ctx = p->ctx;
        if (!ctx->copy_stream) {
            if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return -1;
            if (cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming) != cudaSuccess) return -1;
        }
        if (rt_postprocess(ctx, samples, nullptr) != RT_OK) return -1;
        if (cudaEventRecord(ctx->ev_copy, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0) != cudaSuccess) return -1;
        if (ctx->h_rgb_px != ctx->n_px) {
            if (ctx->h_rgb) cudaFreeHost(ctx->h_rgb);
            ctx->h_rgb = nullptr; ctx->h_rgb_px = 0;
            if (cudaHostAlloc((void **)&ctx->h_rgb, ctx->n_px * 24, cudaHostAllocDefault) != cudaSuccess) return -1;
            ctx->h_rgb_px = ctx->n_px;
        }
        if (cudaMemcpyAsync(ctx->h_rgb, ctx->d_rgb, ctx->n_px * 24, cudaMemcpyDeviceToHost, ctx->copy_stream) != cudaSuccess) return -1;
        return 0;
    }
    static int show(void *self, uint32_t samples) {
        ProgressiveShow *p = (ProgressiveShow *)self;
        if (cudaStreamSynchronize(p->ctx->copy_stream) != cudaSuccess) return -1;
        std::memcpy(p->rgb_out, p->ctx->h_rgb, p->ctx->n_px * 24);
        if (p->samples_done) *p->samples_done = samples;
        return (p->progress && p->progress(p->user, samples, p->total, p->rgb_out) != 0) ? 1 : 0;
    }
};
int rt_render_progressive(rt_ctx *ctx, uint64_t seed, int mode, int sched, uint32_t chunk_samples, rt_progress_fn progress,
                          void *user, double *rgb_out, uint32_t *samples_done) {
    if (int r = need_scene(ctx)) return r;
    if (!rgb_out) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    if (chunk_samples == 0) return set_err(ctx, RT_ERR_INVALID_ARG, "chunk_samples must be positive");
    const uint32_t total = ctx->samples;
    if (samples_done) *samples_done = 0;
    int rmode = mode, rsched = sched;
    resolve_mode(ctx, rmode, rsched);
    if (rsched == RT_SCHED_WAVEFRONT && rmode == RT_MODE_FAST && ctx->dsc.ray_depth > 0 && total > 0) {
        // ONE render call: the wavefront scheduler holds its ordered sums at each chunk boundary while the next
        // chunk's paths are already in flight (WfProgress), so a chunk costs no pipeline drain and refill
        if (int r = check_mode(ctx, rmode)) return r;
        CK(ctx, cudaMemsetAsync(ctx->d_accum, 0, ctx->n_px * 3 * sizeof(double), ctx->stream));
        CK(ctx, ensure_wavefront(ctx));
        ProgressiveShow ps{ctx, progress, user, rgb_out, total, samples_done};
        WfProgress wp;
        wp.chunk = chunk_samples; wp.prepare = &ProgressiveShow::prepare; wp.show = &ProgressiveShow::show; wp.user = &ps;
        const DRegion rg{0u, 0u, ctx->dsc.width, ctx->dsc.height};
        CK(ctx, launch_render_wavefront(ctx->stream, ctx->wf_rt, ctx->dsc, 0u, total, seed, ctx->d_accum, ctx->d_ctr, ctx->wf,
                                        ctx->wf_blocks, &ctx->launches, rg, &wp));
        if (wp.samples_done < total) { CK(ctx, cudaStreamSynchronize(ctx->stream)); return RT_OK; }      // the callback stopped the render: rgb_out holds that image
        if (int r = rt_postprocess(ctx, total, rgb_out)) return r;
        if (samples_done) *samples_done = total;
        if (progress) progress(user, total, total, rgb_out);
        return RT_OK;
    }
    uint32_t done = 0;
    do {                                            // samples == 0 still clears and tonemaps once, like rt_render
        rt_render_params p{};
        p.sample_begin = done;
        p.sample_count = total - done < chunk_samples ? total - done : chunk_samples;
        p.seed = seed; p.mode = mode; p.sched = sched; p.clear = done == 0 ? 1 : 0;
        if (int r = rt_render_accum(ctx, &p)) return r;
        done += p.sample_count;
        // the last line of generate_image (src/main.rs:104) divides by the samples taken SO FAR
        if (int r = rt_postprocess(ctx, done ? done : total, rgb_out)) return r;
        if (samples_done) *samples_done = done;
        if (progress && progress(user, done, total, rgb_out) != 0) break;
    } while (done < total);
    return RT_OK;
}
int rt_debug_rng(rt_ctx *ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n, uint64_t *u64_out, double *probe8) {
    if (!ctx || !u64_out || !probe8) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    DevBuf u, pr;
    CK(ctx, u.alloc(n * 8));
    CK(ctx, pr.alloc(64));
    CK(ctx, launch_rng_probe(ctx->stream, seed, pixel, sample, n, (unsigned long long *)u.p, (double *)pr.p));
    ++ctx->launches;
    CK(ctx, cudaMemcpyAsync(u64_out, u.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(probe8, pr.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_sync(rt_ctx *ctx) {
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "null context");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return RT_OK;
}
int rt_stream(rt_ctx *ctx, void **stream) {
    if (!ctx || !stream) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    *stream = (void *)ctx->stream;
    return RT_OK;
}
int rt_get_counters(rt_ctx *ctx, rt_counters *out) {
    if (!ctx || !out) return set_err(ctx, RT_ERR_INVALID_ARG, "null argument");
    CK(ctx, cudaSetDevice(ctx->device));
    DCounters c{};
    CK(ctx, cudaMemcpyAsync(&c, ctx->d_ctr, sizeof c, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    out->segments = c.segments; out->light_pdf_rays = c.light_pdf_rays; out->paths = c.paths;
    out->aabb_tests = c.aabb_tests; out->prim_tests = c.prim_tests;
    out->light_aabb_tests = c.light_aabb_tests; out->light_prim_tests = c.light_prim_tests;
    out->kernel_launches = ctx->launches;
    out->exact_fallbacks = c.exact_fallbacks;
    return RT_OK;
}
int rt_reset_counters(rt_ctx *ctx) {
    if (!ctx) return set_err(nullptr, RT_ERR_INVALID_ARG, "null context");
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DCounters), ctx->stream));
    ctx->launches = 0;
    return RT_OK;
}

}  // extern "C"
