// Replacement for `generate_image` in the reference's src/main.rs (:85-114).  SOURCE ONLY (no
// Rust toolchain in the build image, so this file has never been compiled; the layouts it relies
// on are checked against the C header by tests/abi/abi_smoke.c + tests/test_abi.py).
// Everything else of main.rs stays: argument handling, scene_parser::parse_scene /
// gltf::build_scene, Scene::new (with the reference's own BVH::new), ppm::save_to_ppm.
//
// The hand-over is rt_scene_from_built: the `Scene` the reference has ALREADY built
// (src/scene.rs:82-90) goes across as flat arrays — per hierarchy the node array of
// src/bvh.rs:49-54 and the primitives in post-build order (BVH::primitives(), src/bvh.rs:23).
// The device therefore traverses the hierarchy produced by the very `BVH::new` of the Rust
// build (whatever its `sort_unstable_by` tie order, SURVEY.md §7 hard part 3): nothing is rebuilt.
//
// Two one-line accessors are all the reference needs to gain (its fields are private to bvh.rs):
//
//     // src/bvh.rs, inside `impl<T: Intersectable + HasAABB> BVH<T>`
//     pub(crate) fn nodes(&self) -> &[Node] { &self.nodes }
//     // src/bvh.rs: `struct Node` and its four fields become `pub(crate)`
//
// `Flattened` below is the other route (rt_scene_build: the library re-runs the same SAH build on
// the un-built primitives); it is kept for hosts that prefer to drop their BVH::new entirely and
// needs the parsed primitives kept next to the scene (see `SceneSource`).
use rt_b200_sys::*;
use std::collections::HashMap;
use std::ffi::CStr;

use crate::aabb::{HasAABB, AABB};
use crate::bvh::BVH;
use crate::image::Image;
use crate::intersections::Intersectable;
use crate::primitives::{Box as BoxPrim, Ellipsoid, Plane, Triangle};
use crate::scene::{Fov, Material, Metadata, Primitive, Scene, TrianglePrimitive};
use crate::types::{Float, Quat, Vec3};

fn check(code: i32, ctx: *const rt_ctx) {
    if code != RT_OK {
        let msg = unsafe { CStr::from_ptr(rt_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("rt_b200: {}", msg); // the reference panics on every error too
    }
}

fn v3(v: &Vec3) -> [f64; 3] {
    [v.x, v.y, v.z]
}

/// Materials are per primitive in the reference (`Metadata`, src/scene.rs:13-18); the C ABI
/// indexes a table, so equal values are interned (bit patterns as the key: NaN-safe, -0.0 != 0.0).
#[derive(Default)]
struct Materials {
    table: Vec<rt_material>,
    index: HashMap<[u64; 8], u32>,
}

impl Materials {
    fn intern(&mut self, m: &Metadata) -> u32 {
        let (kind, ior) = match m.material {
            Material::Diffuse => (RT_MAT_DIFFUSE, 0.0),
            Material::Metallic => (RT_MAT_METALLIC, 0.0),
            Material::Dielectric(ior) => (RT_MAT_DIELECTRIC, ior),
        };
        let key = [
            kind as u64,
            ior.to_bits(),
            m.color.x.to_bits(),
            m.color.y.to_bits(),
            m.color.z.to_bits(),
            m.emission.x.to_bits(),
            m.emission.y.to_bits(),
            m.emission.z.to_bits(),
        ];
        if let Some(&i) = self.index.get(&key) {
            return i;
        }
        let i = self.table.len() as u32;
        self.table.push(rt_material { kind, _pad: 0, ior, color: v3(&m.color), emission: v3(&m.emission) });
        self.index.insert(key, i);
        i
    }
}

fn shape_of<T>(ty: i32, param: &Vec3, p: &Primitive<T>, mats: &mut Materials) -> rt_shape {
    rt_shape {
        ty,
        material: mats.intern(&p.metadata),
        param: v3(param),
        position: v3(&p.position),
        // text order x y z w (src/scene_parser.rs:51-57); cgmath stores (s, v)
        rotation: [p.rotation.v.x, p.rotation.v.y, p.rotation.v.z, p.rotation.s],
    }
}

// ---- route 1: the built Scene as it is (rt_scene_from_built) -----------------------------------

/// One hierarchy in the layout of rt_built_bvh; owns the arrays the raw pointers point into.
#[derive(Default)]
struct FlatBvh {
    bounds: Vec<f64>,
    left: Vec<u32>,
    right: Vec<u32>,
    first: Vec<u32>,
    count: Vec<u32>,
    shapes: Vec<rt_shape>,
    shape_aabb: Vec<f64>,
    tri28: Vec<f64>,
    tri_material: Vec<u32>,
    n_prims: u64,
}

fn child(i: usize) -> u32 {
    if i == usize::MAX { RT_NO_CHILD } else { i as u32 } // src/bvh.rs:148
}

impl FlatBvh {
    fn nodes_of<T: Intersectable + HasAABB>(&mut self, bvh: &BVH<T>) {
        for n in bvh.nodes() {
            self.bounds.extend_from_slice(&[n.aabb.min.x, n.aabb.min.y, n.aabb.min.z, n.aabb.max.x, n.aabb.max.y, n.aabb.max.z]);
            self.left.push(child(n.left_child));
            self.right.push(child(n.right_child));
            self.first.push(n.primitive_indices.start as u32);
            self.count.push((n.primitive_indices.end - n.primitive_indices.start) as u32);
        }
        self.n_prims = bvh.primitives().len() as u64;
    }

    fn aabb_of(&mut self, a: &AABB) {
        self.shape_aabb.extend_from_slice(&[a.min.x, a.min.y, a.min.z, a.max.x, a.max.y, a.max.z]);
    }

    fn boxes(bvh: &BVH<Primitive<BoxPrim>>, mats: &mut Materials) -> Self {
        let mut f = Self::default();
        f.nodes_of(bvh);
        for p in bvh.primitives() {
            f.shapes.push(shape_of(RT_SHAPE_BOX, &p.primitive.sizes, p, mats));
            f.aabb_of(p.aabb());
        }
        f
    }

    fn ellipsoids(bvh: &BVH<Primitive<Ellipsoid>>, mats: &mut Materials) -> Self {
        let mut f = Self::default();
        f.nodes_of(bvh);
        for p in bvh.primitives() {
            f.shapes.push(shape_of(RT_SHAPE_ELLIPSOID, &p.primitive.radiuses, p, mats));
            f.aabb_of(p.aabb());
        }
        f
    }

    /// 28 doubles per triangle, the order rt_scene_get_triangle documents: a, ba, ca,
    /// geometry_normal, na, nb, nc, inverted_area, AABB min, AABB max
    /// (src/primitives/triangle.rs:6-17, src/scene.rs:30-34).
    fn triangles(bvh: &BVH<TrianglePrimitive>, mats: &mut Materials) -> Self {
        let mut f = Self::default();
        f.nodes_of(bvh);
        for p in bvh.primitives() {
            let t: &Triangle = &p.primitive;
            for v in [&t.a, &t.ba, &t.ca, &t.geometry_normal, &t.na, &t.nb, &t.nc] {
                f.tri28.extend_from_slice(&v3(v));
            }
            f.tri28.push(t.inverted_area);
            let a = p.aabb();
            f.tri28.extend_from_slice(&[a.min.x, a.min.y, a.min.z, a.max.x, a.max.y, a.max.z]);
            f.tri_material.push(mats.intern(&p.metadata));
        }
        f
    }

    fn as_raw(&self) -> rt_built_bvh {
        fn ptr<T>(v: &Vec<T>) -> *const T {
            if v.is_empty() { std::ptr::null() } else { v.as_ptr() }
        }
        rt_built_bvh {
            n_nodes: self.left.len() as u64,
            n_prims: self.n_prims,
            bounds: ptr(&self.bounds),
            left: ptr(&self.left),
            right: ptr(&self.right),
            first: ptr(&self.first),
            count: ptr(&self.count),
            shapes: ptr(&self.shapes),
            shape_aabb: ptr(&self.shape_aabb),
            tri28: ptr(&self.tri28),
            tri_material: ptr(&self.tri_material),
        }
    }
}

/// Camera::new (src/camera.rs:18-46), the two tangents only.
fn tan_half_fov(fov: &Fov, width: usize, height: usize) -> (Float, Float) {
    let (w, h) = (width as Float, height as Float);
    match *fov {
        Fov::Y(yfov) => {
            let ty = (yfov / 2.0).tan();
            (ty / (h / w), ty)
        }
        Fov::X(xfov) => {
            let tx = (xfov / 2.0).tan();
            (tx, tx / (w / h))
        }
    }
}

/// The built scene in the layout of rt_built_scene.  Owns every array.
struct FlatScene {
    materials: Materials,
    planes: Vec<rt_shape>,
    bvh: [FlatBvh; 6], // indexed by RT_BVH_*
}

impl FlatScene {
    fn from(scene: &Scene) -> Self {
        let mut materials = Materials::default();
        let planes = scene.primitives.planes.iter().map(|p: &Primitive<Plane>| shape_of(RT_SHAPE_PLANE, &p.primitive.normal, p, &mut materials)).collect();
        let bvh = [
            FlatBvh::boxes(&scene.primitives.boxes, &mut materials),           // RT_BVH_BOXES
            FlatBvh::ellipsoids(&scene.primitives.ellipsoids, &mut materials), // RT_BVH_ELLIPSOIDS
            FlatBvh::triangles(&scene.primitives.triangles, &mut materials),   // RT_BVH_TRIANGLES
            FlatBvh::boxes(&scene.lights.boxes, &mut materials),               // RT_BVH_LIGHT_BOXES
            FlatBvh::ellipsoids(&scene.lights.ellipsoids, &mut materials),     // RT_BVH_LIGHT_ELLIPSOIDS
            FlatBvh::triangles(&scene.lights.triangles, &mut materials),       // RT_BVH_LIGHT_TRIANGLES
        ];
        Self { materials, planes, bvh }
    }

    fn as_raw(&self, scene: &Scene) -> rt_built_scene {
        let (tx, ty) = tan_half_fov(&scene.camera.fov, scene.dimensions.x, scene.dimensions.y);
        rt_built_scene {
            width: scene.dimensions.x as u32,
            height: scene.dimensions.y as u32,
            samples: scene.samples as u32,
            ray_depth: scene.ray_depth as u32,
            bg_color: v3(&scene.bg_color),
            cam_position: v3(&scene.camera.position),
            cam_right: v3(&scene.camera.right),
            cam_up: v3(&scene.camera.up),
            cam_forward: v3(&scene.camera.forward),
            tan_half_fov_x: tx,
            tan_half_fov_y: ty,
            n_materials: self.materials.table.len() as u32,
            n_planes: self.planes.len() as u32,
            materials: self.materials.table.as_ptr(),
            planes: if self.planes.is_empty() { std::ptr::null() } else { self.planes.as_ptr() },
            bvh: [
                self.bvh[0].as_raw(),
                self.bvh[1].as_raw(),
                self.bvh[2].as_raw(),
                self.bvh[3].as_raw(),
                self.bvh[4].as_raw(),
                self.bvh[5].as_raw(),
            ],
        }
    }
}

/// `RT_DEVICES=0,1,2` (or `all`) renders on several GPUs of the box: sample ranges split over the
/// devices, one in-library ncclReduce of the accumulation buffers, tonemap on the first device.
fn devices_from_env() -> Option<Vec<i32>> {
    let v = std::env::var("RT_DEVICES").ok()?;
    if v == "all" {
        return Some(vec![]); // n_devices <= 0: every visible device
    }
    Some(v.split(',').map(|s| s.trim().parse().expect("RT_DEVICES: comma-separated device indices")).collect())
}

fn generate_image(scene: &Scene) -> Image {
    println!("Started raytracing");
    let flat = FlatScene::from(scene);
    let built = flat.as_raw(scene);
    let mut img = Image::new(scene.dimensions.x, scene.dimensions.y);
    // Image.bytes is Vec<Vector3<f64>>: three f64 per pixel, row-major — the layout rt_render writes
    let rgb = img.bytes.as_mut_ptr() as *mut f64;
    let seed = rand::random::<u64>(); // thread_rng() in the reference (src/main.rs:95): OS-seeded too
    unsafe {
        let mut s: *mut rt_scene = std::ptr::null_mut();
        check(rt_scene_from_built(&built, &mut s), std::ptr::null());
        match devices_from_env() {
            Some(ids) => {
                let mut g: *mut rt_group = std::ptr::null_mut();
                check(rt_group_create(ids.len() as i32, if ids.is_empty() { std::ptr::null() } else { ids.as_ptr() }, &mut g), std::ptr::null());
                check(rt_group_upload_scene(g, s), rt_group_ctx(g, 0));
                check(rt_group_render(g, seed, RT_MODE_AUTO, RT_SCHED_AUTO, rgb), rt_group_ctx(g, 0));
                rt_group_destroy(g);
            }
            None => {
                let mut ctx: *mut rt_ctx = std::ptr::null_mut();
                check(rt_create(0, &mut ctx), std::ptr::null());
                check(rt_upload_scene(ctx, s), ctx);
                check(rt_render(ctx, seed, RT_MODE_AUTO, RT_SCHED_AUTO, rgb), ctx);
                rt_destroy(ctx);
            }
        }
        rt_scene_free(s);
    }
    img
}

// ---- route 2: the un-built primitives (rt_scene_build) -----------------------------------------
//
// What Scene::new / build_scene received, kept next to the scene:
//     pub struct Scene { ..., pub source: SceneSource }
// filled at src/scene.rs:180-191 (`SceneSource::Text(scene.primitives)` — make_scenes then takes a
// clone) and src/gltf/scene_builder.rs:12 (`SceneSource::Gltf(triangles.clone())`).
// CAVEAT, and the reason route 1 is the default: `Triangle` keeps (a, ba, ca), not the vertices it
// was made from (src/primitives/triangle.rs:41-47), and rt_raw_scene takes vertices.  b = a + ba is
// rounded once more and the library's own `b - a` need not give back the same `ba` bit for bit, so
// route 2 reproduces the reference up to one ulp of an edge vector; route 1 hands (a, ba, ca) over
// as they are (tri28) and is exact.
pub enum SceneSource {
    Text(Vec<crate::parsed_scene::Primitive>),
    Gltf(Vec<(Vec<Triangle>, Metadata)>), // ColoredTriangles, src/gltf/scene_builder.rs
}

/// The arrays of rt_raw_scene.  Triangle order is the order fed to BVH::new.
pub struct Flattened {
    pub tri_mode: i32,
    pub materials: Vec<rt_material>,
    pub shapes: Vec<rt_shape>,
    pub tri_pos: Vec<f64>,   // n*9: a, b, c
    pub tri_nrm: Vec<f64>,   // n*9: na, nb, nc (Gltf only; Text triangles use the geometry normal)
    pub tri_mat: Vec<u32>,
    pub tri_xform: Vec<f64>, // Text only, n*7: position, rotation x y z w
}

impl Flattened {
    pub fn from(source: &SceneSource) -> Self {
        let mut mats = Materials::default();
        let mut f = Flattened { tri_mode: RT_TRI_BAKED, materials: vec![], shapes: vec![], tri_pos: vec![], tri_nrm: vec![], tri_mat: vec![], tri_xform: vec![] };
        match source {
            // src/gltf/scene_builder.rs:24-30, :42-55: primitive order x triangle order; world-space
            // a / a+ba / a+ca are exactly the three points `instantiate` extends the AABB with
            SceneSource::Gltf(groups) => {
                for (triangles, metadata) in groups {
                    let m = mats.intern(metadata);
                    for t in triangles {
                        for p in [t.a, t.a + t.ba, t.a + t.ca] {
                            f.tri_pos.extend_from_slice(&v3(&p));
                        }
                        for n in [&t.na, &t.nb, &t.nc] {
                            f.tri_nrm.extend_from_slice(&v3(n));
                        }
                        f.tri_mat.push(m);
                    }
                }
            }
            // src/scene.rs:194-207: file order; planes / boxes / ellipsoids keep their own order
            // within their kind, which is all make_scenes preserves
            SceneSource::Text(prims) => {
                use crate::parsed_scene::PrimitiveType;
                f.tri_mode = RT_TRI_MODEL;
                for p in prims {
                    let props = &p.properties;
                    let metadata = Metadata::new(props);
                    let m = mats.intern(&metadata);
                    let position = props.position.unwrap_or(Vec3::new(0.0, 0.0, 0.0));
                    let rotation = props.rotation.unwrap_or(Quat::from_sv(1.0, Vec3::new(0.0, 0.0, 0.0)));
                    let shape = |ty: i32, param: &Vec3| rt_shape {
                        ty,
                        material: m,
                        param: v3(param),
                        position: v3(&position),
                        rotation: [rotation.v.x, rotation.v.y, rotation.v.z, rotation.s],
                    };
                    match p.prim_type.as_ref().unwrap() {
                        PrimitiveType::Plane(pl) => f.shapes.push(shape(RT_SHAPE_PLANE, &pl.normal)),
                        PrimitiveType::Box(b) => f.shapes.push(shape(RT_SHAPE_BOX, &b.sizes)),
                        PrimitiveType::Ellipsoid(e) => f.shapes.push(shape(RT_SHAPE_ELLIPSOID, &e.radiuses)),
                        // model space: TrianglePrimitive::new (src/scene.rs:139-165) rotates and
                        // translates inside the library, with the same operation order
                        PrimitiveType::Triangle(t) => {
                            for q in [t.a, t.ba + t.a, t.ca + t.a] {
                                f.tri_pos.extend_from_slice(&v3(&q));
                            }
                            f.tri_mat.push(m);
                            f.tri_xform.extend_from_slice(&v3(&position));
                            f.tri_xform.extend_from_slice(&[rotation.v.x, rotation.v.y, rotation.v.z, rotation.s]);
                        }
                    }
                }
            }
        }
        f.materials = mats.table;
        f
    }

    pub fn as_raw(&self, scene: &Scene) -> rt_raw_scene {
        fn ptr<T>(v: &Vec<T>) -> *const T {
            if v.is_empty() { std::ptr::null() } else { v.as_ptr() }
        }
        rt_raw_scene {
            width: scene.dimensions.x as u32,
            height: scene.dimensions.y as u32,
            samples: scene.samples as u32,
            ray_depth: scene.ray_depth as u32,
            bg_color: v3(&scene.bg_color),
            cam_position: v3(&scene.camera.position),
            cam_right: v3(&scene.camera.right),
            cam_up: v3(&scene.camera.up),
            cam_forward: v3(&scene.camera.forward),
            fov_kind: match scene.camera.fov { Fov::X(_) => RT_FOV_X, Fov::Y(_) => RT_FOV_Y },
            tri_mode: self.tri_mode,
            fov: match scene.camera.fov { Fov::X(f) | Fov::Y(f) => f },
            n_materials: self.materials.len() as u32,
            n_shapes: self.shapes.len() as u32,
            materials: ptr(&self.materials),
            shapes: ptr(&self.shapes),
            n_triangles: (self.tri_pos.len() / 9) as u64,
            tri_pos: ptr(&self.tri_pos),
            tri_nrm: ptr(&self.tri_nrm),
            tri_mat: ptr(&self.tri_mat),
            tri_xform: ptr(&self.tri_xform),
        }
    }
}

/// generate_image over route 2 (the library builds the hierarchies: rt_scene_build).
#[allow(dead_code)]
fn generate_image_rebuilt(scene: &Scene, source: &SceneSource) -> Image {
    println!("Started raytracing");
    let flat = Flattened::from(source);
    let raw = flat.as_raw(scene);
    let mut img = Image::new(scene.dimensions.x, scene.dimensions.y);
    unsafe {
        let mut s: *mut rt_scene = std::ptr::null_mut();
        let mut ctx: *mut rt_ctx = std::ptr::null_mut();
        check(rt_scene_build(&raw, &mut s), std::ptr::null());
        check(rt_create(0, &mut ctx), std::ptr::null());
        check(rt_upload_scene(ctx, s), ctx);
        check(rt_render(ctx, rand::random::<u64>(), RT_MODE_AUTO, RT_SCHED_AUTO, img.bytes.as_mut_ptr() as *mut f64), ctx);
        rt_destroy(ctx);
        rt_scene_free(s);
    }
    img
}
