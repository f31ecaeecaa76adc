//! `Camera::render_gpu`: the B200 twin of `Camera::render_parallel` (camera.rs:97-112).
//!
//! Flattens the `World` — trait-object shapes, materials, patterns, lights — into the plain arrays of
//! `include/rtgpu.h` (no virtual dispatch on the device) and hands them, with the camera's per-ray fields, to
//! `rtgpu_render`.  The returned `Canvas` holds the same `Color` values the CPU path computes; the library's f64 parity
//! mode follows the reference's arithmetic operation by operation.  There is no CPU fallback: without a usable
//! device the call panics with the library's error message.
//!
//! New file for `ray-tracer/src/composites/`; `gpu-rendering-mode.patch` declares the module, opens the few private
//! fields it reads (`pub(crate)`) and adds the CLI arm.

use crate::composites::{Camera, Canvas, Material, World};
use crate::dyn_partial_eq::DynPartialEq;
use crate::patterns::{CheckerPattern, ComplexPattern, GradientPattern, Pattern, RingPattern, StripePattern, TestPattern};
use crate::primitives::{Color, Transformation};
use crate::shapes::{Cone, Cube, Cylinder, Plane, Shape, Sphere, Triangle};
use rtgpu_sys as sys;

/// The arrays of `rtgpu_scene`, in `world.shapes` / `world.lights` order.
#[derive(Default)]
struct FlatScene {
    shape_type: Vec<u8>,
    shape_inv: Vec<f64>,
    shape_min: Vec<f64>,
    shape_max: Vec<f64>,
    shape_closed: Vec<u8>,
    shape_triangle: Vec<i32>,
    shape_material: Vec<u32>,
    shape_eq_class: Vec<u32>,
    tri_vertex_1: Vec<f64>,
    tri_edge_1: Vec<f64>,
    tri_edge_2: Vec<f64>,
    tri_normal: Vec<f64>,
    mat_color: Vec<f64>,
    mat_params: Vec<f64>,
    mat_casts_shadow: Vec<u8>,
    mat_pattern: Vec<i32>,
    pat_type: Vec<u8>,
    pat_color_a: Vec<f64>,
    pat_color_b: Vec<f64>,
    pat_inv: Vec<f64>,
    pat_child_a: Vec<i32>,
    pat_child_b: Vec<i32>,
    light_position: Vec<f64>,
    light_intensity: Vec<f64>,
    /// materials already emitted: de-duplicated by value (`Material: PartialEq`, material.rs:8)
    materials: Vec<Material>,
}

/// Rows 0..2 of a 4x4 matrix, row-major (matrix.rs:8); row 3 of an affine inverse is (0, 0, 0, 1).
fn push_rows_012(matrix: &Transformation, out: &mut Vec<f64>) {
    for row in 0..3 {
        for column in 0..4 {
            out.push(matrix[row][column]);
        }
    }
}

impl FlatScene {
    /// Emits `pattern` (children of a ComplexPattern first: the ABI requires child index < parent index) and returns
    /// its index.
    fn pattern(&mut self, pattern: &dyn Pattern) -> i32 {
        let any = DynPartialEq::as_any(pattern);
        let black = Color::BLACK;
        let (kind, color_a, color_b, child_a, child_b) = if let Some(p) = any.downcast_ref::<StripePattern>() {
            (sys::RTGPU_PATTERN_STRIPE, p.colors().0, p.colors().1, -1, -1)
        } else if let Some(p) = any.downcast_ref::<GradientPattern>() {
            (sys::RTGPU_PATTERN_GRADIENT, p.colors().0, p.colors().1, -1, -1)
        } else if let Some(p) = any.downcast_ref::<RingPattern>() {
            (sys::RTGPU_PATTERN_RING, p.colors().0, p.colors().1, -1, -1)
        } else if let Some(p) = any.downcast_ref::<CheckerPattern>() {
            (sys::RTGPU_PATTERN_CHECKER, p.colors().0, p.colors().1, -1, -1)
        } else if let Some(p) = any.downcast_ref::<ComplexPattern>() {
            let (a, b) = p.children();
            let index_a = self.pattern(a);
            let index_b = self.pattern(b);
            (sys::RTGPU_PATTERN_COMPLEX, black, black, index_a, index_b)
        } else if any.is::<TestPattern>() {
            (sys::RTGPU_PATTERN_TEST, black, black, -1, -1)
        } else {
            panic!("rtgpu: unsupported pattern type: {pattern}"); // loud, like scene_loader.rs on unknown kinds
        };
        self.pat_type.push(kind);
        self.pat_color_a.extend(color_a.channels());
        self.pat_color_b.extend(color_b.channels());
        push_rows_012(&pattern.transformation_inverse(), &mut self.pat_inv);
        self.pat_child_a.push(child_a);
        self.pat_child_b.push(child_b);
        return (self.pat_type.len() - 1) as i32;
    }

    fn material(&mut self, material: &Material) -> u32 {
        if let Some(index) = self.materials.iter().position(|known| known == material) {
            return index as u32;
        }
        let pattern = material.pattern.as_ref().map_or(-1, |p| self.pattern(p.as_ref()));
        self.mat_color.extend(material.color.channels());
        // order = RTGPU_MAT_PARAM_COUNT block of include/rtgpu.h
        self.mat_params.extend([
            material.ambient,
            material.diffuse,
            material.specular,
            material.shininess,
            material.reflectiveness,
            material.transparency,
            material.refractive_index,
        ]);
        self.mat_casts_shadow.push(u8::from(material.casts_shadow));
        self.mat_pattern.push(pattern);
        self.materials.push(material.clone());
        return (self.materials.len() - 1) as u32;
    }

    fn new(world: &World) -> Self {
        let mut flat = Self::default();
        for (index, shape) in world.shapes.iter().enumerate() {
            let shape: &dyn Shape = shape.as_ref();
            let any = DynPartialEq::as_any(shape);
            let (mut min, mut max, mut closed, mut triangle) = (0.0, 0.0, false, -1);
            let kind = if any.is::<Sphere>() {
                sys::RTGPU_SPHERE
            } else if any.is::<Plane>() {
                sys::RTGPU_PLANE
            } else if any.is::<Cube>() {
                sys::RTGPU_CUBE
            } else if let Some(c) = any.downcast_ref::<Cylinder>() {
                (min, max, closed) = (c.min, c.max, c.closed);
                sys::RTGPU_CYLINDER
            } else if let Some(c) = any.downcast_ref::<Cone>() {
                (min, max, closed) = (c.min, c.max, c.closed);
                sys::RTGPU_CONE
            } else if let Some(t) = any.downcast_ref::<Triangle>() {
                triangle = (flat.tri_vertex_1.len() / 3) as i32;
                flat.tri_vertex_1.extend([t.vertex_1.x, t.vertex_1.y, t.vertex_1.z]);
                flat.tri_edge_1.extend([t.edge_1.x, t.edge_1.y, t.edge_1.z]);
                flat.tri_edge_2.extend([t.edge_2.x, t.edge_2.y, t.edge_2.z]);
                flat.tri_normal.extend([t.normal.x, t.normal.y, t.normal.z]);
                sys::RTGPU_TRIANGLE
            } else {
                panic!("rtgpu: unsupported shape type: {shape:?}");
            };
            flat.shape_type.push(kind);
            push_rows_012(&shape.transformation_inverse(), &mut flat.shape_inv);
            flat.shape_min.push(min);
            flat.shape_max.push(max);
            flat.shape_closed.push(u8::from(closed));
            flat.shape_triangle.push(triangle);
            let material = flat.material(shape.material());
            flat.shape_material.push(material);
            // `dyn Shape == dyn Shape` is equality by value (shape.rs:34-38): the class is the lowest equal index.
            // It decides "same shape" in the refraction container walk (intersection.rs:38,47).
            let class = world.shapes[..index].iter().position(|earlier| earlier.as_ref() == shape).unwrap_or(index);
            flat.shape_eq_class.push(class as u32);
        }
        for light in &world.lights {
            flat.light_position.extend([light.position.x, light.position.y, light.position.z]);
            flat.light_intensity.extend(light.intensity.channels());
        }
        return flat;
    }

    /// The C view of the arrays; valid while `self` is alive and unchanged.
    fn as_c(&self) -> sys::rtgpu_scene {
        return sys::rtgpu_scene {
            abi_version: sys::RTGPU_ABI_VERSION,
            n_shapes: self.shape_type.len() as u32,
            shape_type: self.shape_type.as_ptr(),
            shape_inv: self.shape_inv.as_ptr(),
            shape_min: self.shape_min.as_ptr(),
            shape_max: self.shape_max.as_ptr(),
            shape_closed: self.shape_closed.as_ptr(),
            shape_triangle: self.shape_triangle.as_ptr(),
            shape_material: self.shape_material.as_ptr(),
            shape_eq_class: self.shape_eq_class.as_ptr(),
            n_triangles: (self.tri_vertex_1.len() / 3) as u32,
            tri_vertex_1: self.tri_vertex_1.as_ptr(),
            tri_edge_1: self.tri_edge_1.as_ptr(),
            tri_edge_2: self.tri_edge_2.as_ptr(),
            tri_normal: self.tri_normal.as_ptr(),
            n_materials: self.mat_pattern.len() as u32,
            mat_color: self.mat_color.as_ptr(),
            mat_params: self.mat_params.as_ptr(),
            mat_casts_shadow: self.mat_casts_shadow.as_ptr(),
            mat_pattern: self.mat_pattern.as_ptr(),
            n_patterns: self.pat_type.len() as u32,
            pat_type: self.pat_type.as_ptr(),
            pat_color_a: self.pat_color_a.as_ptr(),
            pat_color_b: self.pat_color_b.as_ptr(),
            pat_inv: self.pat_inv.as_ptr(),
            pat_child_a: self.pat_child_a.as_ptr(),
            pat_child_b: self.pat_child_b.as_ptr(),
            n_lights: (self.light_position.len() / 3) as u32,
            light_position: self.light_position.as_ptr(),
            light_intensity: self.light_intensity.as_ptr(),
        };
    }
}

impl Camera {
    /// GPU twin of `render_parallel`: the same `Canvas`, computed by `librtgpu.so` on `RTGPU_GPUS` devices
    /// (default 1; interleaved 16-row bands, no collective).  Panics if the library reports an error — the gpu
    /// rendering mode has no CPU fallback.
    pub fn render_gpu(&self, world: &World) -> Canvas {
        let flat = FlatScene::new(world);
        let scene = flat.as_c();
        let mut inv = [0.0_f64; 12];
        for row in 0..3 {
            for column in 0..4 {
                inv[row * 4 + column] = self.transformation_inverse[row][column];
            }
        }
        let camera = sys::rtgpu_camera {
            hsize: self.horizontal_size,
            vsize: self.vertical_size,
            half_width: self.half_width,
            half_height: self.half_height,
            pixel_size: self.pixel_size,
            inv,
            origin: [self.origin.x, self.origin.y, self.origin.z],
        };
        let n_gpus = std::env::var("RTGPU_GPUS").ok().and_then(|value| value.parse().ok()).unwrap_or(1);
        let opts = sys::rtgpu_opts {
            precision: sys::RTGPU_PRECISION_F64,
            max_depth: u32::from(World::MAX_REFLECTION_ITERATIONS),
            n_gpus,
            band_rows: 16,
            flags: sys::RTGPU_FLAG_NONE,
        };
        let pixel_count = (self.horizontal_size as usize) * (self.vertical_size as usize);
        let mut rgb = vec![0.0_f64; pixel_count * 3];
        // SAFETY: every pointer in `scene` borrows from `flat`, which outlives the call; `rgb` holds hsize * vsize * 3
        // doubles as rtgpu_render requires; out_rgb8 and stats may be NULL.
        let status = unsafe { sys::rtgpu_render(&scene, &camera, &opts, rgb.as_mut_ptr().cast(), core::ptr::null_mut(), core::ptr::null_mut()) };
        if status != sys::RTGPU_OK {
            // SAFETY: rtgpu_last_error returns a NUL-terminated string owned by the library (thread-local).
            let message = unsafe { core::ffi::CStr::from_ptr(sys::rtgpu_last_error()) }.to_string_lossy().into_owned();
            panic!("rtgpu_render failed ({status}): {message}");
        }
        let mut canvas = Canvas::new(self.horizontal_size, self.vertical_size);
        // Color is not repr(C) (color.rs:6-11): convert instead of transmuting
        for (pixel, channels) in canvas.pixels.iter_mut().zip(rgb.chunks_exact(3)) {
            *pixel = Color::new(channels[0], channels[1], channels[2]);
        }
        return canvas;
    }
}
