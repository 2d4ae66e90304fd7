//! Raw bindings of `include/rt_b200.h`.  SOURCE ONLY: the build image has no Rust toolchain, so
//! this crate has never been compiled; it mirrors the header declaration by declaration and the
//! struct layouts are the ones `tests/test_host_scene.py` checks through ctypes.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const RT_OK: c_int = 0;
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
pub const RT_MODE_EXACT: c_int = 0;
pub const RT_MODE_FAST: c_int = 1;
pub const RT_MODE_AUTO: c_int = 2;
pub const RT_SCHED_MEGAKERNEL: c_int = 0;
pub const RT_SCHED_WAVEFRONT: c_int = 1;
pub const RT_SCHED_AUTO: c_int = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_material {
    pub kind: i32,
    pub _pad: i32,
    pub ior: f64,
    pub color: [f64; 3],
    pub emission: [f64; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_shape {
    pub ty: i32,
    pub material: u32,
    pub param: [f64; 3],
    pub position: [f64; 3],
    pub rotation: [f64; 4], // x y z w
}

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

#[repr(C)]
#[derive(Clone, Copy)]
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

/// Called by `rt_render_progressive` after every chunk of samples; non-zero return stops the render.
pub type rt_progress_fn =
    Option<unsafe extern "C" fn(user: *mut c_void, samples_done: u32, samples_total: u32, rgb: *const f64) -> c_int>;

#[repr(C)]
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

pub enum rt_scene {}
pub enum rt_ctx {}

extern "C" {
    pub fn rt_scene_build(raw: *const rt_raw_scene, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_from_built(built: *const rt_built_scene, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_free(scene: *mut rt_scene);
    pub fn rt_scene_load_text(path: *const c_char, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_load_gltf(path: *const c_char, width: u32, height: u32, samples: u32, out: *mut *mut rt_scene) -> c_int;
    pub fn rt_save_ppm(path: *const c_char, width: u32, height: u32, rgb: *const f64) -> c_int;
    pub fn rt_device_count(count: *mut c_int) -> c_int;
    pub fn rt_create(device_id: c_int, out: *mut *mut rt_ctx) -> c_int;
    pub fn rt_destroy(ctx: *mut rt_ctx);
    pub fn rt_last_error(ctx: *const rt_ctx) -> *const c_char;
    pub fn rt_upload_scene(ctx: *mut rt_ctx, scene: *const rt_scene) -> c_int;
    pub fn rt_render_accum(ctx: *mut rt_ctx, params: *const rt_render_params) -> c_int;
    pub fn rt_accum_device_ptr(ctx: *mut rt_ctx, dptr: *mut *mut c_void, n_doubles: *mut u64) -> c_int;
    pub fn rt_get_accum(ctx: *mut rt_ctx, sum_rgb: *mut f64) -> c_int;
    pub fn rt_set_accum(ctx: *mut rt_ctx, sum_rgb: *const f64) -> c_int;
    pub fn rt_postprocess(ctx: *mut rt_ctx, total_samples: u32, rgb_out: *mut f64) -> c_int;
    pub fn rt_render(ctx: *mut rt_ctx, seed: u64, mode: c_int, sched: c_int, rgb_out: *mut f64) -> c_int;
    pub fn rt_render_progressive(ctx: *mut rt_ctx, seed: u64, mode: c_int, sched: c_int, chunk_samples: u32,
                                 progress: rt_progress_fn, user: *mut c_void, rgb_out: *mut f64, samples_done: *mut u32) -> c_int;
    pub fn rt_scene_build_cached(raw: *const rt_raw_scene, cache_dir: *const c_char, out: *mut *mut rt_scene, cache_hit: *mut c_int) -> c_int;
    pub fn rt_sync(ctx: *mut rt_ctx) -> c_int;
}
