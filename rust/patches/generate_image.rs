// Replacement for `generate_image` in the reference's src/main.rs (:85-114).  SOURCE ONLY (no
// Rust toolchain in the build image).  Everything else of main.rs stays: argument handling,
// scene_parser::parse_scene / gltf::build_scene, ppm::save_to_ppm.
//
// The flattening needs read access to what Scene::new received.  The smallest change to the
// reference is to keep the parsed primitives next to the built scene:
//     pub struct Scene { ..., pub source: SceneSource }
//     pub enum SceneSource { Text(Vec<parsed_scene::Primitive>), Gltf(Vec<(Vec<Triangle>, Metadata)>) }
// (src/scene.rs:180-191 and src/gltf/scene_builder.rs:9-22 already hold both values).
use rt_b200_sys::*;
use std::ffi::CStr;

fn check(code: i32, ctx: *const rt_ctx) {
    if code != RT_OK {
        let msg = unsafe { CStr::from_ptr(rt_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("rt_b200: {}", msg); // the reference panics on every error too
    }
}

fn generate_image(scene: &Scene) -> Image {
    println!("Started raytracing");
    let flat = Flattened::from(scene); // owns materials / shapes / tri_pos / tri_nrm / tri_mat / tri_xform
    let raw = rt_raw_scene {
        width: scene.dimensions.x as u32,
        height: scene.dimensions.y as u32,
        samples: scene.samples as u32,
        ray_depth: scene.ray_depth as u32,
        bg_color: scene.bg_color.into(),
        cam_position: scene.camera.position.into(),
        cam_right: scene.camera.right.into(),
        cam_up: scene.camera.up.into(),
        cam_forward: scene.camera.forward.into(),
        fov_kind: match scene.camera.fov { Fov::X(_) => RT_FOV_X, Fov::Y(_) => RT_FOV_Y },
        tri_mode: flat.tri_mode,
        fov: match scene.camera.fov { Fov::X(f) | Fov::Y(f) => f },
        n_materials: flat.materials.len() as u32,
        n_shapes: flat.shapes.len() as u32,
        materials: flat.materials.as_ptr(),
        shapes: flat.shapes.as_ptr(),
        n_triangles: (flat.tri_pos.len() / 9) as u64,
        tri_pos: flat.tri_pos.as_ptr(),
        tri_nrm: if flat.tri_nrm.is_empty() { std::ptr::null() } else { flat.tri_nrm.as_ptr() },
        tri_mat: flat.tri_mat.as_ptr(),
        tri_xform: if flat.tri_xform.is_empty() { std::ptr::null() } else { flat.tri_xform.as_ptr() },
    };
    let mut img = Image::new(scene.dimensions.x, scene.dimensions.y);
    unsafe {
        let mut s: *mut rt_scene = std::ptr::null_mut();
        let mut ctx: *mut rt_ctx = std::ptr::null_mut();
        check(rt_scene_build(&raw, &mut s), std::ptr::null());
        check(rt_create(0, &mut ctx), std::ptr::null());
        check(rt_upload_scene(ctx, s), ctx);
        // Image.bytes is Vec<Vector3<f64>>: three f64 per pixel, row-major — the layout rt_render writes
        check(rt_render(ctx, rand::random::<u64>(), RT_MODE_AUTO, RT_SCHED_AUTO, img.bytes.as_mut_ptr() as *mut f64), ctx);
        rt_destroy(ctx);
        rt_scene_free(s);
    }
    img
}
