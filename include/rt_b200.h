/* rt_b200.h — C ABI of the B200-native path-tracing core.
 *
 * Drop-in boundary: the reference is one Rust binary whose render hot path is the single call
 *     let img = generate_image(&scene);                       (reference src/main.rs:73, :85-114)
 * Everything the reference does ABOVE that call (CLI, text / glTF loading, Scene::new, BVH::new)
 * is host work; everything BELOW it (camera rays, BVH traversal, primitive tests, the bounce loop,
 * light/MIS PDFs, tonemap + gamma) runs in hand-written sm_100a CUDA kernels behind this header.
 * A Rust `-sys` crate binds exactly these symbols.
 *
 * Conventions
 *   - every function returns int: 0 = RT_OK, otherwise an rt_status; rt_last_error() gives text.
 *   - all host pointers are caller-owned and never retained past the call that received them.
 *   - all arithmetic the reference does in f64 (src/types.rs:5) is exposed as double.
 *   - there is NO CPU fallback for rendering: rt_create fails when no CUDA device is present.
 *
 * Threading
 *   - calls on ONE rt_ctx (or one rt_group) are serialised by the caller; different contexts may
 *     be driven from different threads at the same time (one host thread per GPU).
 *   - an rt_scene is immutable once built: any number of contexts may upload the same scene
 *     concurrently (the flattened, page-locked image is built once, under std::call_once).
 *   - rt_last_error(NULL) is thread-local.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_API __attribute__((visibility("default")))

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID_ARG = 1,
    RT_ERR_NO_DEVICE = 2,   /* no CUDA device / driver: rendering is impossible, no fallback */
    RT_ERR_CUDA = 3,
    RT_ERR_IO = 4,
    RT_ERR_PARSE = 5,       /* scene text / glTF rejected (the reference would panic here)    */
    RT_ERR_NO_SCENE = 6,
    RT_ERR_LIMIT = 7,       /* out of memory, too many primitives for 32-bit indices, ...     */
    RT_ERR_COMM = 8         /* NCCL missing or a collective failed (multi-GPU calls only)     */
} rt_status;

/* ---- scene description (what the reference holds in parsed_scene::Scene / glTF Model) ------- */

/* reference src/scene.rs:7-18 (Material, Metadata) */
enum { RT_MAT_DIFFUSE = 0, RT_MAT_METALLIC = 1, RT_MAT_DIELECTRIC = 2 };
typedef struct rt_material {
    int32_t kind;           /* RT_MAT_*                          */
    int32_t _pad;
    double  ior;            /* Dielectric(ior); ignored otherwise */
    double  color[3];
    double  emission[3];    /* != 0 makes the primitive a light (src/scene.rs:225-227) */
} rt_material;

/* analytic primitives of the text format: reference src/scene.rs:21-27 Primitive<T>,
 * src/primitives/{plane,box,ellipsoid}.rs */
enum { RT_SHAPE_PLANE = 0, RT_SHAPE_BOX = 1, RT_SHAPE_ELLIPSOID = 2 };
typedef struct rt_shape {
    int32_t  type;          /* RT_SHAPE_*                                         */
    uint32_t material;      /* index into rt_raw_scene.materials                  */
    double   param[3];      /* plane normal | box half-sizes | ellipsoid radii    */
    double   position[3];
    double   rotation[4];   /* x y z w (text order, src/scene_parser.rs:51-57)    */
} rt_shape;

enum { RT_FOV_X = 0, RT_FOV_Y = 1 };        /* src/scene.rs:36-40 */
enum { RT_TRI_BAKED = 0,                    /* glTF: world-space a,b,c + smooth normals,
                                               src/gltf/scene_builder.rs:42-55, :342-356 */
       RT_TRI_MODEL = 1 };                  /* text: model-space a,b,c, geometry normals, per-
                                               triangle position+rotation, src/scene.rs:139-165 */

/* Pre-BVH scene: the input of Scene::new (src/scene.rs:180) / build_scene
 * (src/gltf/scene_builder.rs:9).  Triangle order is the order fed to BVH::new. */
typedef struct rt_raw_scene {
    uint32_t width, height;         /* Scene.dimensions                         */
    uint32_t samples;               /* Scene.samples                            */
    uint32_t ray_depth;             /* Scene.ray_depth (u8 in the reference)    */
    double   bg_color[3];
    double   cam_position[3], cam_right[3], cam_up[3], cam_forward[3];
    int32_t  fov_kind;              /* RT_FOV_X | RT_FOV_Y                      */
    int32_t  tri_mode;              /* RT_TRI_BAKED | RT_TRI_MODEL              */
    double   fov;                   /* radians                                  */
    uint32_t n_materials;
    uint32_t n_shapes;
    const rt_material *materials;
    const rt_shape    *shapes;      /* planes, boxes, ellipsoids in file order  */
    uint64_t n_triangles;
    const double   *tri_pos;        /* n*9: a,b,c                               */
    const double   *tri_nrm;        /* n*9: na,nb,nc; NULL -> geometry normals  */
    const uint32_t *tri_mat;        /* n material indices                       */
    const double   *tri_xform;      /* RT_TRI_MODEL only: n*7 = position[3], rotation xyzw[4];
                                       NULL -> identity                         */
} rt_raw_scene;

/* ---- host side: scene construction (reference L4 + L3, runs once per scene) ----------------- */

typedef struct rt_scene rt_scene;   /* built scene: world-space primitives + 6 BVHs */

/* Scene::new / make_scenes (src/scene.rs:180-223) and build_scene
 * (src/gltf/scene_builder.rs:9-22): bakes primitives, builds the 3 light BVHs then the 3 scene
 * BVHs with the reference's full-sweep SAH (src/bvh.rs:12-17, :75-144). */
RT_API int  rt_scene_build(const rt_raw_scene *raw, rt_scene **out);
RT_API void rt_scene_free(rt_scene *scene);

/* scene_parser::parse_scene (src/scene_parser.rs:5-43) + Scene::new.  path NULL -> stdin. */
RT_API int  rt_scene_load_text(const char *path, rt_scene **out);
/* gltf::parse + gltf::build_scene (src/gltf/parser.rs:189, src/gltf/scene_builder.rs:9),
 * buffers resolved against the gltf's directory prefix (src/main.rs:54-59, :77-83). */
RT_API int  rt_scene_load_gltf(const char *path, uint32_t width, uint32_t height, uint32_t samples,
                               rt_scene **out);

/* The same two calls with an on-disk cache of the BUILT scene (the
 * reference rebuilds src/bvh.rs:12-17 on every run).  The file <cache_dir>/<hash>.rtbvh is keyed by
 * a 64-bit hash of the geometry (materials, shapes, triangle arrays); image size, samples, depth,
 * background and camera always come from the arguments, so one cached mesh serves any view.
 * On a miss the scene is built and the file written (atomically: temporary name + rename).
 * *cache_hit (may be NULL) reports which happened.  A damaged file is an error (RT_ERR_IO), never
 * silently rebuilt; the loaded hierarchies pass the same checks as rt_scene_from_built. */
RT_API int  rt_scene_build_cached(const rt_raw_scene *raw, const char *cache_dir, rt_scene **out,
                                  int *cache_hit);
RT_API int  rt_scene_load_gltf_cached(const char *path, uint32_t width, uint32_t height,
                                      uint32_t samples, const char *cache_dir, rt_scene **out,
                                      int *cache_hit);

/* Which of the six hierarchies (src/scene.rs:57-69). Build order in the reference is lights
 * first (src/scene.rs:209-220); the numbering here is only a name. */
enum { RT_BVH_BOXES = 0, RT_BVH_ELLIPSOIDS = 1, RT_BVH_TRIANGLES = 2,
       RT_BVH_LIGHT_BOXES = 3, RT_BVH_LIGHT_ELLIPSOIDS = 4, RT_BVH_LIGHT_TRIANGLES = 5 };
#define RT_NO_CHILD 0xFFFFFFFFu     /* usize::MAX in src/bvh.rs:148 */

typedef struct rt_scene_info {
    uint32_t width, height, samples, ray_depth;
    double   bg_color[3];
    double   cam_position[3], cam_right[3], cam_up[3], cam_forward[3];
    double   tan_half_fov_x, tan_half_fov_y;        /* Camera::new, src/camera.rs:18-46 */
    uint32_t n_planes, n_materials;
    uint64_t bvh_nodes[6], bvh_prims[6];
    uint32_t bvh_depth[6];                          /* max node depth, root = 1        */
} rt_scene_info;
RT_API int rt_scene_get_info(const rt_scene *scene, rt_scene_info *info);

/* BVH in the reference's own layout (src/bvh.rs:49-54 Node): pre-order node array, per node
 * bounds[6] = min xyz, max xyz; left/right child (RT_NO_CHILD for leaves); first/count =
 * primitive_indices range.  prim_order[i] = index (in the order the primitives were fed to
 * BVH::new) of the primitive that ended up at slot i.  Any output pointer may be NULL. */
RT_API int rt_scene_get_bvh(const rt_scene *scene, int which, double *bounds, uint32_t *left,
                            uint32_t *right, uint32_t *first, uint32_t *count,
                            uint32_t *prim_order);
/* World-space triangle i of hierarchy `which` (2 or 5) in post-build order:
 * out[0..27] = a, ba, ca, geometry_normal, na, nb, nc (3 each), inverted_area, then the
 * primitive AABB min xyz, max xyz  (src/primitives/triangle.rs:6-17, src/scene.rs:30-34). */
RT_API int rt_scene_get_triangle(const rt_scene *scene, int which, uint64_t i, double *out28,
                                 uint32_t *material);
/* Shape i of hierarchy `which` (0,1,3,4) in post-build order, or plane i (which = -1). */
RT_API int rt_scene_get_shape(const rt_scene *scene, int which, uint64_t i, rt_shape *out,
                              double *aabb6);
RT_API int rt_scene_get_material(const rt_scene *scene, uint32_t i, rt_material *out);

/* A scene whose hierarchies were ALREADY built by the caller — the reference's own `Scene`
 * (src/scene.rs:82-90) handed over as flat arrays, for a host that keeps its BVH::new
 * (src/bvh.rs:12-17) and only wants the render path replaced.  Layouts are exactly what
 * rt_scene_get_bvh / rt_scene_get_triangle / rt_scene_get_shape return: node arrays in the
 * reference's pre-order numbering (left child = node + 1), primitives in post-build order.
 * Everything is copied; nothing is retained. */
typedef struct rt_built_bvh {
    uint64_t n_nodes, n_prims;
    const double   *bounds;         /* n_nodes * 6: min xyz, max xyz                       */
    const uint32_t *left, *right;   /* RT_NO_CHILD for leaves                              */
    const uint32_t *first, *count;  /* primitive_indices range                             */
    const rt_shape *shapes;         /* RT_BVH_(LIGHT_)BOXES / ELLIPSOIDS: n_prims shapes   */
    const double   *shape_aabb;     /* optional, n_prims * 6: world AABB of each shape     */
    const double   *tri28;          /* RT_BVH_(LIGHT_)TRIANGLES: n_prims * 28 doubles      */
    const uint32_t *tri_material;   /* n_prims                                             */
} rt_built_bvh;
typedef struct rt_built_scene {
    uint32_t width, height, samples, ray_depth;
    double   bg_color[3];
    double   cam_position[3], cam_right[3], cam_up[3], cam_forward[3];
    double   tan_half_fov_x, tan_half_fov_y;        /* Camera::new, src/camera.rs:18-46 */
    uint32_t n_materials, n_planes;
    const rt_material *materials;
    const rt_shape    *planes;
    rt_built_bvh bvh[6];            /* indexed by RT_BVH_*                                 */
} rt_built_scene;
RT_API int rt_scene_from_built(const rt_built_scene *built, rt_scene **out);

/* ppm::save_to_ppm (src/ppm.rs:4-19): binary P6 of post-gamma rgb (width*height*3 doubles). */
RT_API int rt_save_ppm(const char *path, uint32_t width, uint32_t height, const double *rgb);

/* ---- device side: the hot path --------------------------------------------------------------- */

typedef struct rt_ctx rt_ctx;       /* one CUDA device, its streams and all device buffers */

RT_API int  rt_device_count(int *count);
RT_API int  rt_create(int device_id, rt_ctx **out);
RT_API void rt_destroy(rt_ctx *ctx);
RT_API const char *rt_last_error(const rt_ctx *ctx);    /* ctx may be NULL: last global error */

/* Flatten Scene (src/scene.rs:82-90) + its six BVHs (src/bvh.rs:6-9, :49-54) into device
 * buffers; replaces any previous scene of the context. */
RT_API int rt_upload_scene(rt_ctx *ctx, const rt_scene *scene);
/* Bytes of scene records resident in HBM after the last upload (= H2D bytes of one upload). */
RT_API int rt_scene_device_bytes(rt_ctx *ctx, uint64_t *bytes);

/* Arithmetic modes.  EXACT follows the reference operation for operation in f64 (divisions
 * included) and its DFS visit order: it is the parity anchor.  FAST culls with conservative
 * fp32 boxes and an fp32 triangle pre-test, then confirms every candidate with the same exact
 * f64 test, so hit ids / t / normals are the EXACT ones ("fast mode"). */
enum { RT_MODE_EXACT = 0, RT_MODE_FAST = 1,
       RT_MODE_AUTO = 2 };      /* render calls only: FAST + wavefront when the scene has triangles, else EXACT +
                                   megakernel (analytic primitives gain nothing from fp32 culling; the images are
                                   bit for bit the same either way).  Pair it with RT_SCHED_AUTO. */
/* Schedulers for rt_render_accum: one persistent thread per pixel walking its samples
 * (megakernel), or queue-based wavefront with ballot/prefix-sum compaction of live paths. */
enum { RT_SCHED_MEGAKERNEL = 0, RT_SCHED_WAVEFRONT = 1, RT_SCHED_AUTO = 2 };

typedef struct rt_hit {             /* intersections::Intersection + which primitive */
    double   t;
    double   geometry_normal[3];
    double   shading_normal[3];
    int32_t  kind;                  /* -1 miss, 0 plane, 1 box, 2 ellipsoid, 3 triangle */
    uint32_t prim;                  /* index in post-build order of that hierarchy     */
    int32_t  inside;
    uint32_t material;
} rt_hit;

/* intersections::intersect (src/intersections.rs:42-62) for n caller-supplied rays
 * (origins/dirs n*3, used as given: not normalised). */
RT_API int rt_trace_rays(rt_ctx *ctx, int mode, uint64_t n, const double *origins,
                         const double *dirs, rt_hit *hits);
/* Camera::fuzzy_ray with a FIXED in-pixel offset (jx,jy) instead of the two random draws
 * (src/camera.rs:48-55), normalise (src/raytrace.rs:9), then intersect: one ray per pixel,
 * hits[y*W+x].  rays_out (W*H*6: origin, normalised dir) may be NULL. */
RT_API int rt_trace_primary(rt_ctx *ctx, int mode, double jx, double jy, rt_hit *hits,
                            double *rays_out);
/* Light::pdf (src/ray_sampler.rs:132-139): pdf[i] for shading point pos[i], direction dir[i]. */
RT_API int rt_light_pdf(rt_ctx *ctx, int mode, uint64_t n, const double *pos, const double *dir,
                        double *pdf);

typedef struct rt_render_params {
    uint32_t sample_begin;          /* global index of the first sample of this call       */
    uint32_t sample_count;
    uint64_t seed;                  /* counter-RNG key; streams are (pixel, global sample) */
    int32_t  mode;                  /* RT_MODE_*                                           */
    int32_t  sched;                 /* RT_SCHED_*                                          */
    int32_t  clear;                 /* 1: zero the accumulation buffer first               */
    int32_t  _pad;
    uint32_t region[4];             /* x0, y0, w, h of the pixels to render (tile sharding,
                                       region previews); w == 0 or h == 0: the whole image.
                                       Paths stay keyed by the pixel's index in the whole image:
                                       a region gets bit for bit the values the full render
                                       has there, the rest of the accumulation buffer is left
                                       untouched                                            */
} rt_render_params;

/* The pixel loop of generate_image without its last line: accum[p] += Σ_samples raytrace(...)
 * (src/main.rs:94-103).  Asynchronous on the context's stream. */
RT_API int rt_render_accum(rt_ctx *ctx, const rt_render_params *params);
/* Device address / element count of the f64 accumulation buffer (W*H*3), for the NCCL
 * sum-reduce across GPUs that has to happen BEFORE the non-linear tonemap. */
RT_API int rt_accum_device_ptr(rt_ctx *ctx, void **dptr, uint64_t *n_doubles);
RT_API int rt_get_accum(rt_ctx *ctx, double *sum_rgb);          /* D2H, W*H*3 */
RT_API int rt_set_accum(rt_ctx *ctx, const double *sum_rgb);    /* H2D, W*H*3 */
/* correct_gamma(aces_tonemap(sum / total_samples)) (src/main.rs:104, src/postprocessing.rs),
 * on the device; rgb_out = W*H*3 doubles on the host (the reference's Image.bytes), or NULL to
 * leave the image in HBM (asynchronous). */
RT_API int rt_postprocess(rt_ctx *ctx, uint32_t total_samples, double *rgb_out);
/* generate_image (src/main.rs:85-114) in one call: clear, accumulate scene.samples samples,
 * postprocess, copy to host. */
RT_API int rt_render(rt_ctx *ctx, uint64_t seed, int mode, int sched, double *rgb_out);

/* generate_image in chunks of `chunk_samples` samples per pixel, with the image so far after every
 * chunk (the reference only prints "N%" lines, src/main.rs:106-110, and
 * writes once at the end).  After each chunk the accumulation buffer is tonemapped for the samples
 * done so far into rgb_out and `progress(user, samples_done, samples_total, rgb_out)` is called;
 * a non-zero return stops the render there (rgb_out then holds the image of samples_done samples,
 * RT_OK is returned).  progress may be NULL.  Sample s of a pixel is the same path whatever the
 * chunking, and sums are taken in sample order: the final image is bit for bit rt_render's.
 * *samples_done (may be NULL) receives the number of samples in rgb_out. */
typedef int (*rt_progress_fn)(void *user, uint32_t samples_done, uint32_t samples_total, const double *rgb);
RT_API int rt_render_progressive(rt_ctx *ctx, uint64_t seed, int mode, int sched, uint32_t chunk_samples,
                                 rt_progress_fn progress, void *user, double *rgb_out, uint32_t *samples_done);
RT_API int rt_sync(rt_ctx *ctx);
/* CUDA stream handle (cudaStream_t) the context launches on, for event timing by the caller. */
RT_API int rt_stream(rt_ctx *ctx, void **stream);

typedef struct rt_counters {
    uint64_t segments;              /* closest-hit scene queries, src/raytrace.rs:14        */
    uint64_t light_pdf_rays;        /* Light::pdf rays, src/ray_sampler.rs:133              */
    uint64_t paths;                 /* camera rays started                                  */
    uint64_t aabb_tests;            /* AABB::intersects calls in closest-hit (EXACT only)   */
    uint64_t prim_tests;            /* primitive.intersection calls in closest-hit          */
    uint64_t light_aabb_tests;      /* same two for the all-hits light traversals           */
    uint64_t light_prim_tests;
    uint64_t kernel_launches;       /* kernels of this library launched since the last reset */
    uint64_t exact_fallbacks;       /* FAST mode: rays answered by the exact traversal because the
                                       proof that the reference finds the same hit failed (a hit
                                       on the edge of its own box, a tie, a full candidate list) */
} rt_counters;
RT_API int rt_get_counters(rt_ctx *ctx, rt_counters *out);      /* syncs the stream */
RT_API int rt_reset_counters(rt_ctx *ctx);

/* ---- multi-GPU (the reference has one rayon pool, src/main.rs:94) --------------
 * The image is partitioned by SAMPLE RANGES: device r of N renders the block
 * rt_shard_sample_range(r, N, samples, 0) of global sample indices of every pixel (paths are keyed
 * by pixel and GLOBAL sample, so the union is the one-GPU job), the f64 accumulation buffers are
 * sum-reduced to the root with ONE ncclReduce over NVLink before the non-linear tonemap
 * (src/main.rs:104), and only the root postprocesses.  NCCL is loaded at run time (libnccl.so.2);
 * without it these calls return RT_ERR_COMM and everything single-GPU still works. */

/* Contiguous block partition of [first_sample, first_sample + total_samples): sizes differ by at
 * most one, earlier ranks take the larger blocks. */
RT_API int rt_shard_sample_range(int rank, int world, uint32_t total_samples, uint32_t first_sample,
                                 uint32_t *begin, uint32_t *count);

/* One process per GPU (torchrun / MPI launchers): rank 0 creates the id, the launcher broadcasts
 * its RT_COMM_ID_BYTES bytes, every rank joins with its own context (ncclCommInitRank on the
 * context's device; collective: all ranks must call it). */
#define RT_COMM_ID_BYTES 128
RT_API int rt_comm_unique_id(void *id128);
RT_API int rt_comm_init(rt_ctx *ctx, const void *id128, int rank, int world);
RT_API int rt_comm_info(const rt_ctx *ctx, int *rank, int *world, int *nccl_version);   /* any may be NULL */
/* accum(root) = sum over ranks of accum (in place, f64), enqueued on the context's stream after the
 * render kernels; follow it with rt_postprocess on the root.  Collective.  A context without a
 * communicator is a world of one: a no-op. */
RT_API int rt_reduce_accum(rt_ctx *ctx, int root);

/* One process, several GPUs (what the CLI uses): n_devices <= 0 takes every visible device.
 * Device device_ids[0] is the root.  One host thread per device drives the render calls. */
typedef struct rt_group rt_group;
RT_API int  rt_group_create(int n_devices, const int *device_ids, rt_group **out);
RT_API void rt_group_destroy(rt_group *group);
RT_API int  rt_group_size(const rt_group *group, int *n_devices);
RT_API rt_ctx *rt_group_ctx(rt_group *group, int i);       /* borrowed; i = 0 is the root */
RT_API const char *rt_group_last_error(const rt_group *group);
RT_API int  rt_group_upload_scene(rt_group *group, const rt_scene *scene);     /* replicated on every device */
/* rt_render_accum with the sample range split over the devices (no reduce); params->clear applies
 * to the root, the other devices always start their share from zero. */
RT_API int  rt_group_render_accum(rt_group *group, const rt_render_params *params);
RT_API int  rt_group_reduce_accum(rt_group *group);         /* onto the root, asynchronous */
/* generate_image on all devices: split, render, reduce, tonemap on the root, copy to the host. */
RT_API int  rt_group_render(rt_group *group, uint64_t seed, int mode, int sched, double *rgb_out);
RT_API int  rt_group_render_progressive(rt_group *group, uint64_t seed, int mode, int sched,
                                        uint32_t chunk_samples, rt_progress_fn progress, void *user,
                                        double *rgb_out, uint32_t *samples_done);

/* Test hook: the first n u64 draws of the device RNG stream (seed, pixel, sample) and one value
 * of each rand-0.8 distribution used on the path, for parity with the
 * oracle's stream.  probe8 = 8 doubles. */
RT_API int rt_debug_rng(rt_ctx *ctx, uint64_t seed, uint32_t pixel, uint32_t sample, uint64_t n,
                        uint64_t *u64_out, double *probe8);


#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
