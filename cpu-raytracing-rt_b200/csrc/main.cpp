// main.cpp — the reference's command line (reference src/main.rs:30-83, run.sh):
//   cpu-raytracing-rt custom <out.ppm> < scene.txt
//   cpu-raytracing-rt glTF <in.gltf> <W> <H> <spp> <out.ppm>
// Scene loading and the BVH build run on the host; generate_image (src/main.rs:85) runs on the GPU.
// Extra, optional environment knobs (not in the reference): RT_SEED, RT_MODE=exact|fast (fast =
// wavefront scheduler, exact = megakernel parity anchor), RT_DEVICE=<id>, RT_DEVICES=<id,id,...>|all (the
// sample range is split over those GPUs and the accumulation buffers are sum-reduced over NCCL to the
// first one before the tonemap), RT_BVH_CACHE=<dir> (on-disk cache of the built hierarchies),
// RT_PREVIEW=1 (rewrite the output after every chunk of samples).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

static int die(const char *what, rt_ctx *ctx) {
    std::fprintf(stderr, "%s: %s\n", what, rt_last_error(ctx));
    return 101;     // a Rust panic exits with 101
}
static int die_group(const char *what, rt_group *g) {
    std::fprintf(stderr, "%s: %s\n", what, rt_group_last_error(g));
    return 101;
}

namespace {
struct Progress {
    const char *output;
    uint32_t width, height;
    bool preview;
    int last_percent;
};
int on_progress(void *user, uint32_t done, uint32_t total, const double *rgb) {
    Progress *pg = static_cast<Progress *>(user);
    const int percent = total ? (int)((uint64_t)done * 100 / total) : 100;
    if (percent != pg->last_percent) std::printf("%d%%\n", percent);
    pg->last_percent = percent;
    std::fflush(stdout);
    if (pg->preview && done < total) rt_save_ppm(pg->output, pg->width, pg->height, rgb);   // best effort
    return 0;
}
}  // namespace

int main(int argc, char **argv) {
    if (argc < 2) { std::puts("You must specify the input format"); return 0; }
    const std::string format = argv[1];
    if (format != "custom" && format != "glTF") { std::puts("Input format must be 'custom' or 'glTF'"); return 0; }
    const int out_idx = format == "custom" ? 2 : 6;
    if (argc <= out_idx) { std::puts("You must specify the output file"); return 0; }
    const char *output = argv[out_idx];

    std::puts("Started scene parsing");
    rt_scene *scene = nullptr;
    if (format == "glTF") {
        char *e1 = nullptr, *e2 = nullptr, *e3 = nullptr;
        const unsigned long w = std::strtoul(argv[3], &e1, 10), h = std::strtoul(argv[4], &e2, 10), spp = std::strtoul(argv[5], &e3, 10);
        if (*e1 || *e2 || *e3) { std::fprintf(stderr, "width, height and samples must be base-10 integers\n"); return 101; }
        // RT_BVH_CACHE=<dir>: keep the built hierarchies on disk, keyed by a hash of the mesh
        const char *cache = std::getenv("RT_BVH_CACHE");
        const int rc = cache && *cache ? rt_scene_load_gltf_cached(argv[2], (uint32_t)w, (uint32_t)h, (uint32_t)spp, cache, &scene, nullptr)
                                       : rt_scene_load_gltf(argv[2], (uint32_t)w, (uint32_t)h, (uint32_t)spp, &scene);
        if (rc) return die("glTF", nullptr);
    } else {
        if (rt_scene_load_text(nullptr, &scene)) return die("scene", nullptr);
    }
    rt_scene_info info;
    rt_scene_get_info(scene, &info);

    const char *seed_s = std::getenv("RT_SEED"), *mode_s = std::getenv("RT_MODE");
    const uint64_t seed = seed_s ? std::strtoull(seed_s, nullptr, 0) : 0xC0FFEEull;
    // default: FAST + wavefront for scenes with triangles, EXACT + megakernel for purely analytic ones
    const int mode = !mode_s ? RT_MODE_AUTO : (!std::strcmp(mode_s, "exact") ? RT_MODE_EXACT : RT_MODE_FAST);
    const int sched = mode == RT_MODE_AUTO ? RT_SCHED_AUTO : (mode == RT_MODE_FAST ? RT_SCHED_WAVEFRONT : RT_SCHED_MEGAKERNEL);
    std::vector<double> rgb((size_t)info.width * info.height * 3);
    // The reference prints "N%" lines while it renders (src/main.rs:105-110) and writes the image
    // once at the end.  Here the samples are rendered in 20 chunks: one line per chunk, and with
    // RT_PREVIEW=1 the output file is rewritten after every chunk with the image so far.
    const char *preview_s = std::getenv("RT_PREVIEW");
    Progress pg{output, info.width, info.height, preview_s && *preview_s && std::strcmp(preview_s, "0") != 0, 0};
    const uint32_t chunk = info.samples >= 20 ? (info.samples + 19) / 20 : 1;

    const char *devs = std::getenv("RT_DEVICES");
    rt_ctx *ctx = nullptr;
    rt_group *group = nullptr;
    if (devs && *devs) {
        // several GPUs of this box: one context per device inside the library, NCCL reduce to the first
        std::vector<int> ids;
        if (std::strcmp(devs, "all") != 0) {
            for (const char *p = devs; *p;) {
                char *end = nullptr;
                const long v = std::strtol(p, &end, 10);
                if (end == p) { std::fprintf(stderr, "RT_DEVICES must be 'all' or a comma-separated list of device ids\n"); return 101; }
                ids.push_back((int)v);
                p = *end == ',' ? end + 1 : end;
            }
        }
        if (rt_group_create((int)ids.size(), ids.empty() ? nullptr : ids.data(), &group)) return die("devices", nullptr);
        if (rt_group_upload_scene(group, scene)) return die_group("upload", group);
        std::puts("Started raytracing");
        // fewer, larger chunks than on one GPU: every chunk is split over the devices and ends in a reduce
        int n_dev = 1;
        rt_group_size(group, &n_dev);
        const uint32_t gchunk = chunk * (uint32_t)n_dev;
        if (rt_group_render_progressive(group, seed, mode, sched, gchunk, on_progress, &pg, rgb.data(), nullptr)) return die_group("render", group);
    } else {
        const char *dev = std::getenv("RT_DEVICE");
        if (rt_create(dev ? std::atoi(dev) : 0, &ctx)) return die("device", nullptr);
        if (rt_upload_scene(ctx, scene)) return die("upload", ctx);
        std::puts("Started raytracing");
        if (rt_render_progressive(ctx, seed, mode, sched, chunk, on_progress, &pg, rgb.data(), nullptr)) return die("render", ctx);
    }
    if (pg.last_percent < 100) std::puts("100%");
    if (rt_save_ppm(output, info.width, info.height, rgb.data())) return die("Cannot create output file", nullptr);
    rt_destroy(ctx);
    rt_group_destroy(group);
    rt_scene_free(scene);
    return 0;
}
