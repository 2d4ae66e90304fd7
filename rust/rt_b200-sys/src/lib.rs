//! Raw bindings of `include/rt_b200.h`, declaration by declaration.
//!
//! SOURCE ONLY: the build image has no Rust toolchain, so this crate has never been compiled here.
//! What keeps it honest meanwhile: every `#[repr(C)]` struct carries an `@layout` line (size and the
//! byte offset of every field); `tests/abi/abi_smoke.c --layout` prints the same table from the C
//! header with `sizeof` / `offsetof`, and `tests/test_abi.py` asserts that the two tables are equal
//! (and that every `extern "C"` name below is exported by `librt_b200.so`).  With a toolchain,
//! `cargo test` checks the same table against `std::mem::{size_of, offset_of}` (mod layout).
//! @const RT_NO_CHILD=4294967295 RT_COMM_ID_BYTES=128 RT_MODE_AUTO=2 RT_SCHED_AUTO=2 RT_ERR_COMM=8 RT_BVH_LIGHT_TRIANGLES=5
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const RT_OK: c_int = 0;
pub const RT_ERR_INVALID_ARG: c_int = 1;
pub const RT_ERR_NO_DEVICE: c_int = 2;
pub const RT_ERR_CUDA: c_int = 3;
pub const RT_ERR_IO: c_int = 4;
pub const RT_ERR_PARSE: c_int = 5;
pub const RT_ERR_NO_SCENE: c_int = 6;
pub const RT_ERR_LIMIT: c_int = 7;
pub const RT_ERR_COMM: c_int = 8;
pub const RT_MAT_DIFFUSE: i32 = 0;
pub const RT_MAT_METALLIC: i32 = 1;
pub const RT_MAT_DIELECTRIC: i32 = 2;
pub const RT_SHAPE_PLANE: i32 = 0;
pub const RT_SHAPE_BOX: i32 = 1;
pub const RT_SHAPE_ELLIPSOID: i32 = 2;
pub const RT_FOV_X: i32 = 0;
pub const RT_FOV_Y: i32 = 1;
pub const RT_TRI_BAKED: i32 = 0;
pub const RT_TRI_MODEL: i32 = 1;
pub const RT_BVH_BOXES: usize = 0;
pub const RT_BVH_ELLIPSOIDS: usize = 1;
pub const RT_BVH_TRIANGLES: usize = 2;
pub const RT_BVH_LIGHT_BOXES: usize = 3;
pub const RT_BVH_LIGHT_ELLIPSOIDS: usize = 4;
pub const RT_BVH_LIGHT_TRIANGLES: usize = 5;
/// `usize::MAX` in the reference's `Node` (src/bvh.rs:148)
pub const RT_NO_CHILD: u32 = 0xFFFF_FFFF;
pub const RT_MODE_EXACT: c_int = 0;
pub const RT_MODE_FAST: c_int = 1;
pub const RT_MODE_AUTO: c_int = 2;
pub const RT_SCHED_MEGAKERNEL: c_int = 0;
pub const RT_SCHED_WAVEFRONT: c_int = 1;
pub const RT_SCHED_AUTO: c_int = 2;
pub const RT_COMM_ID_BYTES: usize = 128;

/// @layout rt_material size=64 kind=0 _pad=4 ior=8 color=16 emission=40
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_material {
    pub kind: i32,
    pub _pad: i32,
    pub ior: f64,
    pub color: [f64; 3],
    pub emission: [f64; 3],
}

/// (`ty` is `type` in the header)
/// @layout rt_shape size=88 type=0 material=4 param=8 position=32 rotation=56
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_shape {
    pub ty: i32,
    pub material: u32,
    pub param: [f64; 3],
    pub position: [f64; 3],
    pub rotation: [f64; 4], // x y z w
}

/// @layout rt_raw_scene size=216 width=0 height=4 samples=8 ray_depth=12 bg_color=16 cam_position=40 cam_right=64 cam_up=88 cam_forward=112 fov_kind=136 tri_mode=140 fov=144 n_materials=152 n_shapes=156 materials=160 shapes=168 n_triangles=176 tri_pos=184 tri_nrm=192 tri_mat=200 tri_xform=208
#[repr(C)]
pub struct rt_raw_scene {
    pub width: u32,
    pub height: u32,
    pub samples: u32,
    pub ray_depth: u32,
    pub bg_color: [f64; 3],
    pub cam_position: [f64; 3],
    pub cam_right: [f64; 3],
    pub cam_up: [f64; 3],
    pub cam_forward: [f64; 3],
    pub fov_kind: i32,
    pub tri_mode: i32,
    pub fov: f64,
    pub n_materials: u32,
    pub n_shapes: u32,
    pub materials: *const rt_material,
    pub shapes: *const rt_shape,
    pub n_triangles: u64,
    pub tri_pos: *const f64,
    pub tri_nrm: *const f64,
    pub tri_mat: *const u32,
    pub tri_xform: *const f64,
}

/// @layout rt_scene_info size=280 width=0 height=4 samples=8 ray_depth=12 bg_color=16 cam_position=40 cam_right=64 cam_up=88 cam_forward=112 tan_half_fov_x=136 tan_half_fov_y=144 n_planes=152 n_materials=156 bvh_nodes=160 bvh_prims=208 bvh_depth=256
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_scene_info {
    pub width: u32,
    pub height: u32,
    pub samples: u32,
    pub ray_depth: u32,
    pub bg_color: [f64; 3],
    pub cam_position: [f64; 3],
    pub cam_right: [f64; 3],
    pub cam_up: [f64; 3],
    pub cam_forward: [f64; 3],
    pub tan_half_fov_x: f64,
    pub tan_half_fov_y: f64,
    pub n_planes: u32,
    pub n_materials: u32,
    pub bvh_nodes: [u64; 6],
    pub bvh_prims: [u64; 6],
    pub bvh_depth: [u32; 6],
}

/// @layout rt_built_bvh size=88 n_nodes=0 n_prims=8 bounds=16 left=24 right=32 first=40 count=48 shapes=56 shape_aabb=64 tri28=72 tri_material=80
#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_built_bvh {
    pub n_nodes: u64,
    pub n_prims: u64,
    pub bounds: *const f64,
    pub left: *const u32,
    pub right: *const u32,
    pub first: *const u32,
    pub count: *const u32,
    pub shapes: *const rt_shape,
    pub shape_aabb: *const f64,
    pub tri28: *const f64,
    pub tri_material: *const u32,
}

/// @layout rt_built_scene size=704 width=0 height=4 samples=8 ray_depth=12 bg_color=16 cam_position=40 cam_right=64 cam_up=88 cam_forward=112 tan_half_fov_x=136 tan_half_fov_y=144 n_materials=152 n_planes=156 materials=160 planes=168 bvh=176
#[repr(C)]
pub struct rt_built_scene {
    pub width: u32,
    pub height: u32,
    pub samples: u32,
    pub ray_depth: u32,
    pub bg_color: [f64; 3],
    pub cam_position: [f64; 3],
    pub cam_right: [f64; 3],
    pub cam_up: [f64; 3],
    pub cam_forward: [f64; 3],
    pub tan_half_fov_x: f64,
    pub tan_half_fov_y: f64,
    pub n_materials: u32,
    pub n_planes: u32,
    pub materials: *const rt_material,
    pub planes: *const rt_shape,
    pub bvh: [rt_built_bvh; 6],
}

/// @layout rt_hit size=72 t=0 geometry_normal=8 shading_normal=32 kind=56 prim=60 inside=64 material=68
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_hit {
    pub t: f64,
    pub geometry_normal: [f64; 3],
    pub shading_normal: [f64; 3],
    pub kind: i32,
    pub prim: u32,
    pub inside: i32,
    pub material: u32,
}

/// @layout rt_render_params size=48 sample_begin=0 sample_count=4 seed=8 mode=16 sched=20 clear=24 _pad=28 region=32
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_render_params {
    pub sample_begin: u32,
    pub sample_count: u32,
    pub seed: u64,
    pub mode: i32,
    pub sched: i32,
    pub clear: i32,
    pub _pad: i32,
    /// x0, y0, w, h of the pixels to render; w == 0 or h == 0: the whole image
    pub region: [u32; 4],
}

/// @layout rt_counters size=72 segments=0 light_pdf_rays=8 paths=16 aabb_tests=24 prim_tests=32 light_aabb_tests=40 light_prim_tests=48 kernel_launches=56 exact_fallbacks=64
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_counters {
    pub segments: u64,
    pub light_pdf_rays: u64,
    pub paths: u64,
    pub aabb_tests: u64,
    pub prim_tests: u64,
    pub light_aabb_tests: u64,
    pub light_prim_tests: u64,
    pub kernel_launches: u64,
    pub exact_fallbacks: u64,
}

/// Called by `rt_render_progressive` after every chunk of samples; non-zero return stops the render.
pub type rt_progress_fn =
    Option<unsafe extern "C" fn(user: *mut c_void, samples_done: u32, samples_total: u32, rgb: *const f64) -> c_int>;

pub enum rt_scene {}
pub enum rt_ctx {}
pub enum rt_group {}

extern "C" {
    // host side: scene construction
    pub fn rt_scene_build(raw: *const rt_raw_scene, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_free(scene: *mut rt_scene);
    pub fn rt_scene_load_text(path: *const c_char, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_load_gltf(path: *const c_char, width: u32, height: u32, samples: u32, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_build_cached(raw: *const rt_raw_scene, cache_dir: *const c_char, out: *mut *mut rt_scene, cache_hit: *mut c_int) -> c_int;
    pub fn rt_scene_load_gltf_cached(path: *const c_char, width: u32, height: u32, samples: u32, cache_dir: *const c_char,
                                     out: *mut *mut rt_scene, cache_hit: *mut c_int) -> c_int;
    pub fn rt_scene_get_info(scene: *const rt_scene, info: *mut rt_scene_info) -> c_int;
    pub fn rt_scene_get_bvh(scene: *const rt_scene, which: c_int, bounds: *mut f64, left: *mut u32, right: *mut u32, first: *mut u32,
                            count: *mut u32, prim_order: *mut u32) -> c_int;
    pub fn rt_scene_get_triangle(scene: *const rt_scene, which: c_int, i: u64, out28: *mut f64, material: *mut u32) -> c_int;
    pub fn rt_scene_get_shape(scene: *const rt_scene, which: c_int, i: u64, out: *mut rt_shape, aabb6: *mut f64) -> c_int;
    pub fn rt_scene_get_material(scene: *const rt_scene, i: u32, out: *mut rt_material) -> c_int;
    pub fn rt_scene_from_built(built: *const rt_built_scene, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_save_ppm(path: *const c_char, width: u32, height: u32, rgb: *const f64) -> c_int;
    // device side: the hot path
    pub fn rt_device_count(count: *mut c_int) -> c_int;
    pub fn rt_create(device_id: c_int, out: *mut *mut rt_ctx) -> c_int;
    pub fn rt_destroy(ctx: *mut rt_ctx);
    pub fn rt_last_error(ctx: *const rt_ctx) -> *const c_char;
    pub fn rt_upload_scene(ctx: *mut rt_ctx, scene: *const rt_scene) -> c_int;
    pub fn rt_scene_device_bytes(ctx: *mut rt_ctx, bytes: *mut u64) -> c_int;
    pub fn rt_trace_rays(ctx: *mut rt_ctx, mode: c_int, n: u64, origins: *const f64, dirs: *const f64, hits: *mut rt_hit) -> c_int;
    pub fn rt_trace_primary(ctx: *mut rt_ctx, mode: c_int, jx: f64, jy: f64, hits: *mut rt_hit, rays_out: *mut f64) -> c_int;
    pub fn rt_light_pdf(ctx: *mut rt_ctx, mode: c_int, n: u64, pos: *const f64, dir: *const f64, pdf: *mut f64) -> c_int;
    pub fn rt_render_accum(ctx: *mut rt_ctx, params: *const rt_render_params) -> c_int;
    pub fn rt_accum_device_ptr(ctx: *mut rt_ctx, dptr: *mut *mut c_void, n_doubles: *mut u64) -> c_int;
    pub fn rt_get_accum(ctx: *mut rt_ctx, sum_rgb: *mut f64) -> c_int;
    pub fn rt_set_accum(ctx: *mut rt_ctx, sum_rgb: *const f64) -> c_int;
    pub fn rt_postprocess(ctx: *mut rt_ctx, total_samples: u32, rgb_out: *mut f64) -> c_int;
    pub fn rt_render(ctx: *mut rt_ctx, seed: u64, mode: c_int, sched: c_int, rgb_out: *mut f64) -> c_int;
    pub fn rt_render_progressive(ctx: *mut rt_ctx, seed: u64, mode: c_int, sched: c_int, chunk_samples: u32,
                                 progress: rt_progress_fn, user: *mut c_void, rgb_out: *mut f64, samples_done: *mut u32) -> c_int;
    pub fn rt_sync(ctx: *mut rt_ctx) -> c_int;
    pub fn rt_stream(ctx: *mut rt_ctx, stream: *mut *mut c_void) -> c_int;
    pub fn rt_get_counters(ctx: *mut rt_ctx, out: *mut rt_counters) -> c_int;
    pub fn rt_reset_counters(ctx: *mut rt_ctx) -> c_int;
    // multi-GPU
    pub fn rt_shard_sample_range(rank: c_int, world: c_int, total_samples: u32, first_sample: u32, begin: *mut u32, count: *mut u32) -> c_int;
    pub fn rt_comm_unique_id(id128: *mut c_void) -> c_int;
    pub fn rt_comm_init(ctx: *mut rt_ctx, id128: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn rt_comm_info(ctx: *const rt_ctx, rank: *mut c_int, world: *mut c_int, nccl_version: *mut c_int) -> c_int;
    pub fn rt_reduce_accum(ctx: *mut rt_ctx, root: c_int) -> c_int;
    pub fn rt_group_create(n_devices: c_int, device_ids: *const c_int, out: *mut *mut rt_group) -> c_int;
    pub fn rt_group_destroy(group: *mut rt_group);
    pub fn rt_group_size(group: *const rt_group, n_devices: *mut c_int) -> c_int;
    pub fn rt_group_ctx(group: *mut rt_group, i: c_int) -> *mut rt_ctx;
    pub fn rt_group_last_error(group: *const rt_group) -> *const c_char;
    pub fn rt_group_upload_scene(group: *mut rt_group, scene: *const rt_scene) -> c_int;
    pub fn rt_group_render_accum(group: *mut rt_group, params: *const rt_render_params) -> c_int;
    pub fn rt_group_reduce_accum(group: *mut rt_group) -> c_int;
    pub fn rt_group_render(group: *mut rt_group, seed: u64, mode: c_int, sched: c_int, rgb_out: *mut f64) -> c_int;
    pub fn rt_group_render_progressive(group: *mut rt_group, seed: u64, mode: c_int, sched: c_int, chunk_samples: u32,
                                       progress: rt_progress_fn, user: *mut c_void, rgb_out: *mut f64, samples_done: *mut u32) -> c_int;
    // test hook
    pub fn rt_debug_rng(ctx: *mut rt_ctx, seed: u64, pixel: u32, sample: u32, n: u64, u64_out: *mut u64, probe8: *mut f64) -> c_int;
}

/// The `@layout` tables above, checked by the compiler (needs Rust >= 1.77 for `offset_of!`).
#[cfg(test)]
mod layout {
    use super::*;
    use std::mem::{offset_of, size_of};

    #[test]
    fn sizes_and_offsets_match_the_header() {
        assert_eq!(size_of::<rt_material>(), 64);
        assert_eq!(offset_of!(rt_material, kind), 0);
        assert_eq!(offset_of!(rt_material, _pad), 4);
        assert_eq!(offset_of!(rt_material, ior), 8);
        assert_eq!(offset_of!(rt_material, color), 16);
        assert_eq!(offset_of!(rt_material, emission), 40);
        assert_eq!(size_of::<rt_shape>(), 88);
        assert_eq!(offset_of!(rt_shape, ty), 0);
        assert_eq!(offset_of!(rt_shape, material), 4);
        assert_eq!(offset_of!(rt_shape, param), 8);
        assert_eq!(offset_of!(rt_shape, position), 32);
        assert_eq!(offset_of!(rt_shape, rotation), 56);
        assert_eq!(size_of::<rt_raw_scene>(), 216);
        assert_eq!(offset_of!(rt_raw_scene, width), 0);
        assert_eq!(offset_of!(rt_raw_scene, height), 4);
        assert_eq!(offset_of!(rt_raw_scene, samples), 8);
        assert_eq!(offset_of!(rt_raw_scene, ray_depth), 12);
        assert_eq!(offset_of!(rt_raw_scene, bg_color), 16);
        assert_eq!(offset_of!(rt_raw_scene, cam_position), 40);
        assert_eq!(offset_of!(rt_raw_scene, cam_right), 64);
        assert_eq!(offset_of!(rt_raw_scene, cam_up), 88);
        assert_eq!(offset_of!(rt_raw_scene, cam_forward), 112);
        assert_eq!(offset_of!(rt_raw_scene, fov_kind), 136);
        assert_eq!(offset_of!(rt_raw_scene, tri_mode), 140);
        assert_eq!(offset_of!(rt_raw_scene, fov), 144);
        assert_eq!(offset_of!(rt_raw_scene, n_materials), 152);
        assert_eq!(offset_of!(rt_raw_scene, n_shapes), 156);
        assert_eq!(offset_of!(rt_raw_scene, materials), 160);
        assert_eq!(offset_of!(rt_raw_scene, shapes), 168);
        assert_eq!(offset_of!(rt_raw_scene, n_triangles), 176);
        assert_eq!(offset_of!(rt_raw_scene, tri_pos), 184);
        assert_eq!(offset_of!(rt_raw_scene, tri_nrm), 192);
        assert_eq!(offset_of!(rt_raw_scene, tri_mat), 200);
        assert_eq!(offset_of!(rt_raw_scene, tri_xform), 208);
        assert_eq!(size_of::<rt_scene_info>(), 280);
        assert_eq!(offset_of!(rt_scene_info, width), 0);
        assert_eq!(offset_of!(rt_scene_info, height), 4);
        assert_eq!(offset_of!(rt_scene_info, samples), 8);
        assert_eq!(offset_of!(rt_scene_info, ray_depth), 12);
        assert_eq!(offset_of!(rt_scene_info, bg_color), 16);
        assert_eq!(offset_of!(rt_scene_info, cam_position), 40);
        assert_eq!(offset_of!(rt_scene_info, cam_right), 64);
        assert_eq!(offset_of!(rt_scene_info, cam_up), 88);
        assert_eq!(offset_of!(rt_scene_info, cam_forward), 112);
        assert_eq!(offset_of!(rt_scene_info, tan_half_fov_x), 136);
        assert_eq!(offset_of!(rt_scene_info, tan_half_fov_y), 144);
        assert_eq!(offset_of!(rt_scene_info, n_planes), 152);
        assert_eq!(offset_of!(rt_scene_info, n_materials), 156);
        assert_eq!(offset_of!(rt_scene_info, bvh_nodes), 160);
        assert_eq!(offset_of!(rt_scene_info, bvh_prims), 208);
        assert_eq!(offset_of!(rt_scene_info, bvh_depth), 256);
        assert_eq!(size_of::<rt_built_bvh>(), 88);
        assert_eq!(offset_of!(rt_built_bvh, n_nodes), 0);
        assert_eq!(offset_of!(rt_built_bvh, n_prims), 8);
        assert_eq!(offset_of!(rt_built_bvh, bounds), 16);
        assert_eq!(offset_of!(rt_built_bvh, left), 24);
        assert_eq!(offset_of!(rt_built_bvh, right), 32);
        assert_eq!(offset_of!(rt_built_bvh, first), 40);
        assert_eq!(offset_of!(rt_built_bvh, count), 48);
        assert_eq!(offset_of!(rt_built_bvh, shapes), 56);
        assert_eq!(offset_of!(rt_built_bvh, shape_aabb), 64);
        assert_eq!(offset_of!(rt_built_bvh, tri28), 72);
        assert_eq!(offset_of!(rt_built_bvh, tri_material), 80);
        assert_eq!(size_of::<rt_built_scene>(), 704);
        assert_eq!(offset_of!(rt_built_scene, width), 0);
        assert_eq!(offset_of!(rt_built_scene, height), 4);
        assert_eq!(offset_of!(rt_built_scene, samples), 8);
        assert_eq!(offset_of!(rt_built_scene, ray_depth), 12);
        assert_eq!(offset_of!(rt_built_scene, bg_color), 16);
        assert_eq!(offset_of!(rt_built_scene, cam_position), 40);
        assert_eq!(offset_of!(rt_built_scene, cam_right), 64);
        assert_eq!(offset_of!(rt_built_scene, cam_up), 88);
        assert_eq!(offset_of!(rt_built_scene, cam_forward), 112);
        assert_eq!(offset_of!(rt_built_scene, tan_half_fov_x), 136);
        assert_eq!(offset_of!(rt_built_scene, tan_half_fov_y), 144);
        assert_eq!(offset_of!(rt_built_scene, n_materials), 152);
        assert_eq!(offset_of!(rt_built_scene, n_planes), 156);
        assert_eq!(offset_of!(rt_built_scene, materials), 160);
        assert_eq!(offset_of!(rt_built_scene, planes), 168);
        assert_eq!(offset_of!(rt_built_scene, bvh), 176);
        assert_eq!(size_of::<rt_hit>(), 72);
        assert_eq!(offset_of!(rt_hit, t), 0);
        assert_eq!(offset_of!(rt_hit, geometry_normal), 8);
        assert_eq!(offset_of!(rt_hit, shading_normal), 32);
        assert_eq!(offset_of!(rt_hit, kind), 56);
        assert_eq!(offset_of!(rt_hit, prim), 60);
        assert_eq!(offset_of!(rt_hit, inside), 64);
        assert_eq!(offset_of!(rt_hit, material), 68);
        assert_eq!(size_of::<rt_render_params>(), 48);
        assert_eq!(offset_of!(rt_render_params, sample_begin), 0);
        assert_eq!(offset_of!(rt_render_params, sample_count), 4);
        assert_eq!(offset_of!(rt_render_params, seed), 8);
        assert_eq!(offset_of!(rt_render_params, mode), 16);
        assert_eq!(offset_of!(rt_render_params, sched), 20);
        assert_eq!(offset_of!(rt_render_params, clear), 24);
        assert_eq!(offset_of!(rt_render_params, _pad), 28);
        assert_eq!(offset_of!(rt_render_params, region), 32);
        assert_eq!(size_of::<rt_counters>(), 72);
        assert_eq!(offset_of!(rt_counters, segments), 0);
        assert_eq!(offset_of!(rt_counters, light_pdf_rays), 8);
        assert_eq!(offset_of!(rt_counters, paths), 16);
        assert_eq!(offset_of!(rt_counters, aabb_tests), 24);
        assert_eq!(offset_of!(rt_counters, prim_tests), 32);
        assert_eq!(offset_of!(rt_counters, light_aabb_tests), 40);
        assert_eq!(offset_of!(rt_counters, light_prim_tests), 48);
        assert_eq!(offset_of!(rt_counters, kernel_launches), 56);
        assert_eq!(offset_of!(rt_counters, exact_fallbacks), 64);
    }
}
