// scene_cache.cpp — on-disk cache of the built hierarchies, keyed by a hash of the geometry.
//
// the reference rebuilds its six BVHs on every run (src/bvh.rs:12-17 is
// called from src/scene.rs:209-220 and src/gltf/scene_builder.rs:12-17); for the 20 M-triangle
// configuration that is the longest step of the whole job.  The build is a pure function of the
// primitive list, so its OUTPUT (world-space primitives in post-build order + node arrays — the
// very bits the parity tests compare) is stored once and read back on the next run.  What is not
// geometry (image size, sample count, depth, background, camera) is applied from the caller's
// rt_raw_scene after the load, so a cached mesh can be rendered at any resolution / spp.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <unistd.h>
#include <stdexcept>

#include "scene_host.hpp"

namespace rt {
namespace {

// 4-lane multiply-rotate word hash (the lanes are independent so the loop pipelines); it only has
// to tell meshes apart, every count is ALSO stored in the header and compared.
struct Hasher {
    uint64_t h[4] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0x165667B19E3779F9ull, 0x27D4EB2F165667C5ull};
    static uint64_t mix(uint64_t a, uint64_t w) {
        a = (a ^ w) * 0xFF51AFD7ED558CCDull;
        return (a << 31) | (a >> 33);
    }
    void bytes(const void *p, size_t n) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        size_t i = 0;
        for (; i + 32 <= n; i += 32) {
            uint64_t w[4];
            memcpy(w, c + i, 32);
            for (int k = 0; k < 4; ++k) h[k] = mix(h[k], w[k]);
        }
        for (int k = 0; i < n; i += 8, ++k) {
            uint64_t w = 0;
            memcpy(&w, c + i, n - i < 8 ? n - i : 8);
            h[k & 3] = mix(h[k & 3], w ^ 0xA5A5A5A5A5A5A5A5ull);
        }
        h[0] = mix(h[0], (uint64_t)n);
    }
    template <class T>
    void pod(const T &v) { bytes(&v, sizeof v); }
    void array(const void *p, size_t n) {            // presence is part of the key (NULL has a meaning)
        pod<uint64_t>(p ? n : ~0ull);
        if (p) bytes(p, n);
    }
    uint64_t finish() const {
        uint64_t a = 0x2545F4914F6CDD1Dull;
        for (int k = 0; k < 4; ++k) a = mix(a, h[k]);
        a ^= a >> 32;
        return a * 0xD6E8FEB86659FD93ull;
    }
};

constexpr char kMagic[8] = {'R', 'T', 'B', '2', 'B', 'V', 'H', '1'};

struct Header {
    char magic[8];
    uint32_t version, sz_node, sz_tri, sz_shape, sz_material, n_materials;
    uint64_t hash;
    uint64_t n_planes;
    uint64_t n_nodes[6], n_prims[6];
};

struct File {
    FILE *f = nullptr;
    ~File() { if (f) fclose(f); }
};

void put(FILE *f, const void *p, size_t n) {
    if (n && fwrite(p, 1, n, f) != n) throw std::runtime_error(std::string("scene cache: write failed: ") + strerror(errno));
}
void get(FILE *f, void *p, size_t n) {
    if (n && fread(p, 1, n, f) != n) throw std::runtime_error("scene cache: file is truncated");
}

// The checks rt_scene_from_built applies to arrays handed over by a caller, applied to arrays
// read from disk: a damaged file must not become an out-of-range index on the device.
template <class P>
void check_hierarchy(HostBvh<P> &b, size_t n_materials) {
    b.depth = validate_hierarchy(b.nodes, b.prims.size(), "scene cache: ");
    for (const P &p : b.prims)
        if (p.material >= n_materials) throw std::runtime_error("scene cache: material index out of range");
}

}  // namespace

uint32_t validate_hierarchy(const std::vector<HostNode> &nodes, uint64_t n_prims, const char *prefix) {
    const size_t nn = nodes.size();
    const std::string pre(prefix);
    std::vector<uint32_t> depth(nn, 0);
    std::vector<uint8_t> parents(nn, 0);
    uint32_t deepest = 0;
    if (nn) depth[0] = 1;
    for (size_t i = 0; i < nn; ++i) {
        const HostNode &n = nodes[i];
        // children always have larger indices than their parent, so by the time node i is visited
        // every possible parent of it has been: a non-root node without one is unreachable
        if (i > 0 && parents[i] != 1) throw std::runtime_error(pre + (parents[i] ? "node with two parents" : "unreachable node"));
        const bool leaf = n.left == RT_NO_CHILD;
        if (leaf != (n.right == RT_NO_CHILD)) throw std::runtime_error(pre + "node with exactly one child");
        if (leaf) {
            if ((uint64_t)n.first + n.count > n_prims) throw std::runtime_error(pre + "leaf range outside the primitive array");
        } else {
            if (n.left != i + 1 || n.right <= n.left || n.right >= nn) throw std::runtime_error(pre + "nodes are not in the reference's pre-order numbering");
            for (const uint32_t c : {n.left, n.right}) {
                if (parents[c] < 2) ++parents[c];
                depth[c] = std::max(depth[c], depth[i] + 1);
            }
        }
        deepest = std::max(deepest, depth[i]);
    }
    return deepest;
}

uint64_t geometry_hash(const rt_raw_scene &raw) {
    Hasher h;
    h.pod<int32_t>(raw.tri_mode);
    h.pod<uint32_t>(raw.n_materials);
    for (uint32_t i = 0; i < raw.n_materials; ++i) {           // field by field: the struct has a pad word
        const rt_material &m = raw.materials[i];
        h.pod(m.kind); h.pod(m.ior); h.pod(m.color); h.pod(m.emission);
    }
    h.pod<uint32_t>(raw.n_shapes);
    for (uint32_t i = 0; i < raw.n_shapes; ++i) {
        const rt_shape &s = raw.shapes[i];
        h.pod(s.type); h.pod(s.material); h.pod(s.param); h.pod(s.position); h.pod(s.rotation);
    }
    const size_t n = raw.n_triangles;
    h.pod<uint64_t>(n);
    h.array(raw.tri_pos, n * 9 * sizeof(double));
    h.array(raw.tri_nrm, n * 9 * sizeof(double));
    h.array(raw.tri_mat, n * sizeof(uint32_t));
    h.array(raw.tri_xform, n * 7 * sizeof(double));
    return h.finish();
}

void save_scene_cache(const std::string &path, uint64_t hash, const HostScene &s) {
    Header hd{};
    memcpy(hd.magic, kMagic, 8);
    hd.version = 1;
    hd.sz_node = sizeof(HostNode); hd.sz_tri = sizeof(HostTri); hd.sz_shape = sizeof(HostShape); hd.sz_material = sizeof(rt_material);
    hd.n_materials = (uint32_t)s.materials.size();
    hd.hash = hash;
    hd.n_planes = s.planes.size();
    const HostBvh<HostShape> *sh[4] = {&s.boxes, &s.ellipsoids, &s.light_boxes, &s.light_ellipsoids};
    const int sh_id[4] = {RT_BVH_BOXES, RT_BVH_ELLIPSOIDS, RT_BVH_LIGHT_BOXES, RT_BVH_LIGHT_ELLIPSOIDS};
    const HostBvh<HostTri> *tr[2] = {&s.tris, &s.light_tris};
    const int tr_id[2] = {RT_BVH_TRIANGLES, RT_BVH_LIGHT_TRIANGLES};
    for (int k = 0; k < 4; ++k) { hd.n_nodes[sh_id[k]] = sh[k]->nodes.size(); hd.n_prims[sh_id[k]] = sh[k]->prims.size(); }
    for (int k = 0; k < 2; ++k) { hd.n_nodes[tr_id[k]] = tr[k]->nodes.size(); hd.n_prims[tr_id[k]] = tr[k]->prims.size(); }
    // written under a private name and renamed: a reader never sees a half-written file, and two
    // processes (one per GPU) racing to fill the cache both end with a complete one
    const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)hash ^ (unsigned long long)(uintptr_t)&hd);
    {
        File f;
        f.f = fopen(tmp.c_str(), "wb");
        if (!f.f) throw std::runtime_error("scene cache: cannot create " + tmp + ": " + strerror(errno));
        put(f.f, &hd, sizeof hd);
        put(f.f, s.materials.data(), s.materials.size() * sizeof(rt_material));
        put(f.f, s.planes.data(), s.planes.size() * sizeof(HostShape));
        for (int k = 0; k < 4; ++k) {
            put(f.f, sh[k]->nodes.data(), sh[k]->nodes.size() * sizeof(HostNode));
            put(f.f, sh[k]->prims.data(), sh[k]->prims.size() * sizeof(HostShape));
        }
        for (int k = 0; k < 2; ++k) {
            put(f.f, tr[k]->nodes.data(), tr[k]->nodes.size() * sizeof(HostNode));
            put(f.f, tr[k]->prims.data(), tr[k]->prims.size() * sizeof(HostTri));
        }
        // flushed to the disk before the rename publishes it: a crash must not leave a short file under the final name
        if (fflush(f.f) != 0 || fsync(fileno(f.f)) != 0) throw std::runtime_error(std::string("scene cache: write failed: ") + strerror(errno));
    }
    if (rename(tmp.c_str(), path.c_str()) != 0) {
        const std::string why = strerror(errno);
        remove(tmp.c_str());
        throw std::runtime_error("scene cache: cannot rename to " + path + ": " + why);
    }
}

bool load_scene_cache(const std::string &path, uint64_t hash, const rt_raw_scene *raw, HostScene &s) {
    File f;
    f.f = fopen(path.c_str(), "rb");
    if (!f.f) return false;                                  // not cached yet
    Header hd{};
    get(f.f, &hd, sizeof hd);
    if (memcmp(hd.magic, kMagic, 8) != 0 || hd.version != 1) throw std::runtime_error("scene cache: " + path + " is not a scene cache of this version");
    if (hd.sz_node != sizeof(HostNode) || hd.sz_tri != sizeof(HostTri) || hd.sz_shape != sizeof(HostShape) || hd.sz_material != sizeof(rt_material))
        throw std::runtime_error("scene cache: " + path + " was written with another record layout");
    if (hd.hash != hash) return false;                       // another mesh under this name
    if (raw) {
        // the 64-bit hash is not cryptographic: a file that agrees on the hash but not on the COUNTS of
        // the caller's scene is another mesh (a collision or a stale file) — a miss, the file is rewritten
        uint64_t by_type[3] = {0, 0, 0};
        for (uint32_t i = 0; i < raw->n_shapes; ++i)
            if (raw->shapes[i].type >= 0 && raw->shapes[i].type < 3) ++by_type[raw->shapes[i].type];
        if (hd.n_materials != raw->n_materials || hd.n_planes != by_type[RT_SHAPE_PLANE] || hd.n_prims[RT_BVH_BOXES] != by_type[RT_SHAPE_BOX] ||
            hd.n_prims[RT_BVH_ELLIPSOIDS] != by_type[RT_SHAPE_ELLIPSOID] || hd.n_prims[RT_BVH_TRIANGLES] != raw->n_triangles)
            return false;
    }
    for (int k = 0; k < 6; ++k)
        if (hd.n_nodes[k] >= RT_NO_CHILD || hd.n_prims[k] >= RT_NO_CHILD) throw std::runtime_error("scene cache: counts out of range");
    // the counts are bounded by the file size before anything is allocated from them
    if (fseek(f.f, 0, SEEK_END) != 0) throw std::runtime_error("scene cache: seek failed");
    const long long file_bytes = ftell(f.f);
    unsigned long long need = sizeof hd + (unsigned long long)hd.n_materials * sizeof(rt_material) + hd.n_planes * sizeof(HostShape);
    for (int k = 0; k < 6; ++k) {
        const bool tri = k == RT_BVH_TRIANGLES || k == RT_BVH_LIGHT_TRIANGLES;
        need += hd.n_nodes[k] * sizeof(HostNode) + hd.n_prims[k] * (tri ? sizeof(HostTri) : sizeof(HostShape));
    }
    if (file_bytes < 0 || (unsigned long long)file_bytes != need) throw std::runtime_error("scene cache: " + path + " has the wrong size for its header");
    if (fseek(f.f, (long)sizeof hd, SEEK_SET) != 0) throw std::runtime_error("scene cache: seek failed");

    HostScene g;
    g.materials.resize(hd.n_materials);
    get(f.f, g.materials.data(), g.materials.size() * sizeof(rt_material));
    g.planes.resize(hd.n_planes);
    get(f.f, g.planes.data(), g.planes.size() * sizeof(HostShape));
    HostBvh<HostShape> *sh[4] = {&g.boxes, &g.ellipsoids, &g.light_boxes, &g.light_ellipsoids};
    const int sh_id[4] = {RT_BVH_BOXES, RT_BVH_ELLIPSOIDS, RT_BVH_LIGHT_BOXES, RT_BVH_LIGHT_ELLIPSOIDS};
    HostBvh<HostTri> *tr[2] = {&g.tris, &g.light_tris};
    const int tr_id[2] = {RT_BVH_TRIANGLES, RT_BVH_LIGHT_TRIANGLES};
    for (int k = 0; k < 4; ++k) {
        sh[k]->nodes.resize(hd.n_nodes[sh_id[k]]);
        get(f.f, sh[k]->nodes.data(), sh[k]->nodes.size() * sizeof(HostNode));
        sh[k]->prims.resize(hd.n_prims[sh_id[k]]);
        get(f.f, sh[k]->prims.data(), sh[k]->prims.size() * sizeof(HostShape));
        check_hierarchy(*sh[k], g.materials.size());
    }
    for (int k = 0; k < 2; ++k) {
        tr[k]->nodes.resize(hd.n_nodes[tr_id[k]]);
        get(f.f, tr[k]->nodes.data(), tr[k]->nodes.size() * sizeof(HostNode));
        tr[k]->prims.resize(hd.n_prims[tr_id[k]]);
        get(f.f, tr[k]->prims.data(), tr[k]->prims.size() * sizeof(HostTri));
        check_hierarchy(*tr[k], g.materials.size());
    }
    for (const HostShape &p : g.planes)
        if (p.material >= g.materials.size()) throw std::runtime_error("scene cache: material index out of range");
    s = std::move(g);
    return true;
}

std::string scene_cache_path(const std::string &dir, uint64_t hash) {
    char name[40];
    snprintf(name, sizeof name, "%016llx.rtbvh", (unsigned long long)hash);
    if (dir.empty()) return name;
    return dir.back() == '/' ? dir + name : dir + "/" + name;
}

bool build_scene_cached(const rt_raw_scene &raw, const std::string &cache_dir, HostScene &out) {
    if ((raw.n_materials && !raw.materials) || (raw.n_shapes && !raw.shapes) || (raw.n_triangles && !raw.tri_pos))
        throw std::runtime_error("null array in rt_raw_scene");
    const uint64_t hash = geometry_hash(raw);
    const std::string path = scene_cache_path(cache_dir, hash);
    HostScene cached;
    if (load_scene_cache(path, hash, &raw, cached)) {
        apply_view(raw, cached);
        out = std::move(cached);
        return true;
    }
    build_scene(raw, out);
    save_scene_cache(path, hash, out);
    return false;
}

}  // namespace rt
