/* abi_smoke.c — the C ABI checked from plain C, against include/rt_b200.h ONLY (no Python, no C++):
 *
 *   abi_smoke --layout      prints sizeof / offsetof of every struct of the header, one line per
 *                           struct in the form of the `@layout` lines of rust/rt_b200-sys/src/lib.rs
 *                           (tests/test_abi.py asserts the two tables are equal, so the #[repr(C)]
 *                           mirrors cannot drift from the header although no rustc is at hand);
 *   abi_smoke --flow        the whole reference-facing flow a host performs in place of
 *                           generate_image (reference src/main.rs:73): rt_scene_build ->
 *                           rt_create -> rt_upload_scene -> rt_render -> rt_destroy, on a two-triangle
 *                           scene with a light; without a CUDA device it checks that rt_create fails
 *                           with RT_ERR_NO_DEVICE (there is no CPU fallback) and exits 0.
 * Built by tests/test_abi.py:  gcc -std=c11 -Wall -Wextra -Werror -Iinclude tests/abi/abi_smoke.c -L... -lrt_b200
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt_b200.h"

#define BEGIN(T) printf("@layout %s size=%zu", #T, sizeof(T))
#define F(T, f) printf(" %s=%zu", #f, offsetof(T, f))
#define END() printf("\n")

static void layout(void) {
    BEGIN(rt_material); F(rt_material, kind); F(rt_material, _pad); F(rt_material, ior); F(rt_material, color); F(rt_material, emission); END();
    BEGIN(rt_shape); F(rt_shape, type); F(rt_shape, material); F(rt_shape, param); F(rt_shape, position); F(rt_shape, rotation); END();
    BEGIN(rt_raw_scene); F(rt_raw_scene, width); F(rt_raw_scene, height); F(rt_raw_scene, samples); F(rt_raw_scene, ray_depth);
    F(rt_raw_scene, bg_color); F(rt_raw_scene, cam_position); F(rt_raw_scene, cam_right); F(rt_raw_scene, cam_up); F(rt_raw_scene, cam_forward);
    F(rt_raw_scene, fov_kind); F(rt_raw_scene, tri_mode); F(rt_raw_scene, fov); F(rt_raw_scene, n_materials); F(rt_raw_scene, n_shapes);
    F(rt_raw_scene, materials); F(rt_raw_scene, shapes); F(rt_raw_scene, n_triangles); F(rt_raw_scene, tri_pos); F(rt_raw_scene, tri_nrm);
    F(rt_raw_scene, tri_mat); F(rt_raw_scene, tri_xform); END();
    BEGIN(rt_scene_info); F(rt_scene_info, width); F(rt_scene_info, height); F(rt_scene_info, samples); F(rt_scene_info, ray_depth);
    F(rt_scene_info, bg_color); F(rt_scene_info, cam_position); F(rt_scene_info, cam_right); F(rt_scene_info, cam_up); F(rt_scene_info, cam_forward);
    F(rt_scene_info, tan_half_fov_x); F(rt_scene_info, tan_half_fov_y); F(rt_scene_info, n_planes); F(rt_scene_info, n_materials);
    F(rt_scene_info, bvh_nodes); F(rt_scene_info, bvh_prims); F(rt_scene_info, bvh_depth); END();
    BEGIN(rt_built_bvh); F(rt_built_bvh, n_nodes); F(rt_built_bvh, n_prims); F(rt_built_bvh, bounds); F(rt_built_bvh, left); F(rt_built_bvh, right);
    F(rt_built_bvh, first); F(rt_built_bvh, count); F(rt_built_bvh, shapes); F(rt_built_bvh, shape_aabb); F(rt_built_bvh, tri28);
    F(rt_built_bvh, tri_material); END();
    BEGIN(rt_built_scene); F(rt_built_scene, width); F(rt_built_scene, height); F(rt_built_scene, samples); F(rt_built_scene, ray_depth);
    F(rt_built_scene, bg_color); F(rt_built_scene, cam_position); F(rt_built_scene, cam_right); F(rt_built_scene, cam_up); F(rt_built_scene, cam_forward);
    F(rt_built_scene, tan_half_fov_x); F(rt_built_scene, tan_half_fov_y); F(rt_built_scene, n_materials); F(rt_built_scene, n_planes);
    F(rt_built_scene, materials); F(rt_built_scene, planes); F(rt_built_scene, bvh); END();
    BEGIN(rt_hit); F(rt_hit, t); F(rt_hit, geometry_normal); F(rt_hit, shading_normal); F(rt_hit, kind); F(rt_hit, prim); F(rt_hit, inside);
    F(rt_hit, material); END();
    BEGIN(rt_render_params); F(rt_render_params, sample_begin); F(rt_render_params, sample_count); F(rt_render_params, seed); F(rt_render_params, mode);
    F(rt_render_params, sched); F(rt_render_params, clear); F(rt_render_params, _pad); F(rt_render_params, region); END();
    BEGIN(rt_counters); F(rt_counters, segments); F(rt_counters, light_pdf_rays); F(rt_counters, paths); F(rt_counters, aabb_tests);
    F(rt_counters, prim_tests); F(rt_counters, light_aabb_tests); F(rt_counters, light_prim_tests); F(rt_counters, kernel_launches);
    F(rt_counters, exact_fallbacks); END();
    printf("@const RT_NO_CHILD=%u RT_COMM_ID_BYTES=%d RT_MODE_AUTO=%d RT_SCHED_AUTO=%d RT_ERR_COMM=%d RT_BVH_LIGHT_TRIANGLES=%d\n",
           RT_NO_CHILD, RT_COMM_ID_BYTES, RT_MODE_AUTO, RT_SCHED_AUTO, RT_ERR_COMM, RT_BVH_LIGHT_TRIANGLES);
}

#define CHECK(call)                                                                                      \
    do {                                                                                                 \
        int rc_ = (call);                                                                                \
        if (rc_ != RT_OK) {                                                                              \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rt_last_error(NULL));                          \
            return 1;                                                                                    \
        }                                                                                                \
    } while (0)

static int flow(void) {
    /* a diffuse floor quad under an emissive quad, camera looking along +z from above the floor */
    enum { W = 48, H = 32, SPP = 4 };
    rt_material mats[2];
    memset(mats, 0, sizeof mats);
    mats[0].kind = RT_MAT_DIFFUSE; mats[0].color[0] = 0.8; mats[0].color[1] = 0.7; mats[0].color[2] = 0.6;
    mats[1].kind = RT_MAT_DIFFUSE; mats[1].emission[0] = mats[1].emission[1] = mats[1].emission[2] = 12.0;
    const double pos[4 * 9] = {
        -4, -1, 0,   4, -1, 0,   4, -1, 9,      -4, -1, 0,   4, -1, 9,  -4, -1, 9,     /* floor, normal +y */
        -1, 2, 3,    1, 2, 5,    1, 2, 3,       -1, 2, 3,   -1, 2, 5,    1, 2, 5,      /* light, facing down */
    };
    const uint32_t tri_mat[4] = {0, 0, 1, 1};
    rt_raw_scene raw;
    memset(&raw, 0, sizeof raw);
    raw.width = W; raw.height = H; raw.samples = SPP; raw.ray_depth = 4;
    raw.cam_position[1] = 0.5; raw.cam_position[2] = -2.0;
    raw.cam_right[0] = 1.0; raw.cam_up[1] = 1.0; raw.cam_forward[2] = 1.0;
    raw.fov_kind = RT_FOV_X; raw.fov = 1.2; raw.tri_mode = RT_TRI_BAKED;
    raw.n_materials = 2; raw.materials = mats;
    raw.n_triangles = 4; raw.tri_pos = pos; raw.tri_mat = tri_mat;

    rt_scene *scene = NULL;
    CHECK(rt_scene_build(&raw, &scene));
    rt_scene_info info;
    CHECK(rt_scene_get_info(scene, &info));
    if (info.width != W || info.height != H || info.bvh_prims[RT_BVH_TRIANGLES] != 4 || info.bvh_prims[RT_BVH_LIGHT_TRIANGLES] != 2) {
        fprintf(stderr, "rt_scene_get_info: unexpected counts\n");
        return 1;
    }
    int n_dev = 0;
    const int rc_count = rt_device_count(&n_dev);
    rt_ctx *ctx = NULL;
    const int rc_create = rt_create(0, &ctx);
    if (rc_count != RT_OK || n_dev == 0) {
        /* no GPU: the render path must refuse, never fall back to the CPU */
        if (rc_create != RT_ERR_NO_DEVICE || ctx != NULL) {
            fprintf(stderr, "rt_create without a device returned %d (expected RT_ERR_NO_DEVICE)\n", rc_create);
            return 1;
        }
        rt_scene_free(scene);
        printf("flow ok (no CUDA device: rt_create -> RT_ERR_NO_DEVICE, as specified)\n");
        return 0;
    }
    if (rc_create != RT_OK) { fprintf(stderr, "rt_create: %s\n", rt_last_error(NULL)); return 1; }
    CHECK(rt_upload_scene(ctx, scene));
    double *rgb = (double *)malloc(sizeof(double) * W * H * 3), *rgb2 = (double *)malloc(sizeof(double) * W * H * 3);
    if (!rgb || !rgb2) return 1;
    CHECK(rt_render(ctx, 7u, RT_MODE_AUTO, RT_SCHED_AUTO, rgb));
    /* the same image from the parity anchor (EXACT arithmetic, megakernel): bit for bit */
    CHECK(rt_render(ctx, 7u, RT_MODE_EXACT, RT_SCHED_MEGAKERNEL, rgb2));
    double lit = 0.0;
    for (int i = 0; i < W * H * 3; ++i) {
        if (!(rgb[i] >= 0.0 && rgb[i] <= 1.0)) { fprintf(stderr, "pixel value %g outside [0, 1]\n", rgb[i]); return 1; }
        lit += rgb[i];
    }
    if (memcmp(rgb, rgb2, sizeof(double) * W * H * 3) != 0) { fprintf(stderr, "FAST/wavefront image differs from EXACT/megakernel\n"); return 1; }
    if (!(lit > 1.0)) { fprintf(stderr, "the image is black\n"); return 1; }
    rt_counters c;
    CHECK(rt_get_counters(ctx, &c));
    if (c.paths != 2ull * W * H * SPP || c.segments < c.paths) { fprintf(stderr, "counters: paths %llu segments %llu\n", (unsigned long long)c.paths, (unsigned long long)c.segments); return 1; }
    rt_destroy(ctx);
    rt_scene_free(scene);
    free(rgb); free(rgb2);
    printf("flow ok (%d device(s), %llu segments, mean value %.4f)\n", n_dev, (unsigned long long)c.segments, lit / (W * H * 3));
    return 0;
}

int main(int argc, char **argv) {
    if (argc == 2 && strcmp(argv[1], "--layout") == 0) { layout(); return 0; }
    if (argc == 2 && strcmp(argv[1], "--flow") == 0) return flow();
    fprintf(stderr, "usage: abi_smoke --layout | --flow\n");
    return 2;
}
