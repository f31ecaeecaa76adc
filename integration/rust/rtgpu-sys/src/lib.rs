//! Raw FFI declarations of `include/rtgpu.h` (ABI version 1).  Field order and types mirror the header one to
//! one; `tests/test_rust_binding.py` of the rtgpu repository parses this file and the header and compares them.
#![allow(non_camel_case_types)]
#![no_std]

use core::ffi::{c_char, c_int, c_void};

pub const RTGPU_ABI_VERSION: u32 = 1;

// rtgpu_status
pub const RTGPU_OK: c_int = 0;
pub const RTGPU_ERR_INVALID_ARGUMENT: c_int = -1;
pub const RTGPU_ERR_UNSUPPORTED: c_int = -2;
pub const RTGPU_ERR_NO_DEVICE: c_int = -3;
pub const RTGPU_ERR_CUDA: c_int = -4;
pub const RTGPU_ERR_OUT_OF_MEMORY: c_int = -5;

// rtgpu_shape_type
pub const RTGPU_SPHERE: u8 = 0;
pub const RTGPU_PLANE: u8 = 1;
pub const RTGPU_CUBE: u8 = 2;
pub const RTGPU_CYLINDER: u8 = 3;
pub const RTGPU_CONE: u8 = 4;
pub const RTGPU_TRIANGLE: u8 = 5;

// rtgpu_pattern_type
pub const RTGPU_PATTERN_STRIPE: u8 = 0;
pub const RTGPU_PATTERN_GRADIENT: u8 = 1;
pub const RTGPU_PATTERN_RING: u8 = 2;
pub const RTGPU_PATTERN_CHECKER: u8 = 3;
pub const RTGPU_PATTERN_COMPLEX: u8 = 4;
pub const RTGPU_PATTERN_TEST: u8 = 5;

/// Scalars per material in `mat_params`: ambient, diffuse, specular, shininess, reflectiveness, transparency, refractive_index
pub const RTGPU_MAT_PARAM_COUNT: usize = 7;

// rtgpu_precision
pub const RTGPU_PRECISION_F64: u32 = 0;
pub const RTGPU_PRECISION_F32: u32 = 1;

// rtgpu_opts.flags
pub const RTGPU_FLAG_NONE: u32 = 0;
pub const RTGPU_FLAG_WAVEFRONT: u32 = 1;
pub const RTGPU_FLAG_PERSISTENT: u32 = 2;

#[repr(C)]
pub struct rtgpu_scene {
    pub abi_version: u32,
    pub n_shapes: u32,
    pub shape_type: *const u8,
    pub shape_inv: *const f64,
    pub shape_min: *const f64,
    pub shape_max: *const f64,
    pub shape_closed: *const u8,
    pub shape_triangle: *const i32,
    pub shape_material: *const u32,
    pub shape_eq_class: *const u32,
    pub n_triangles: u32,
    pub tri_vertex_1: *const f64,
    pub tri_edge_1: *const f64,
    pub tri_edge_2: *const f64,
    pub tri_normal: *const f64,
    pub n_materials: u32,
    pub mat_color: *const f64,
    pub mat_params: *const f64,
    pub mat_casts_shadow: *const u8,
    pub mat_pattern: *const i32,
    pub n_patterns: u32,
    pub pat_type: *const u8,
    pub pat_color_a: *const f64,
    pub pat_color_b: *const f64,
    pub pat_inv: *const f64,
    pub pat_child_a: *const i32,
    pub pat_child_b: *const i32,
    pub n_lights: u32,
    pub light_position: *const f64,
    pub light_intensity: *const f64,
}

#[repr(C)]
pub struct rtgpu_camera {
    pub hsize: u32,
    pub vsize: u32,
    pub half_width: f64,
    pub half_height: f64,
    pub pixel_size: f64,
    pub inv: [f64; 12],
    pub origin: [f64; 3],
}

#[repr(C)]
pub struct rtgpu_rows {
    pub band_rows: u32,
    pub shard_index: u32,
    pub shard_count: u32,
}

#[repr(C)]
pub struct rtgpu_opts {
    pub precision: u32,
    pub max_depth: u32,
    pub n_gpus: i32,
    pub band_rows: u32,
    pub flags: u32,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct rtgpu_stats {
    pub rays_primary: u64,
    pub rays_shadow: u64,
    pub rays_reflect: u64,
    pub rays_refract: u64,
    pub hit_nodes: u64,
    pub pixels: u64,
    pub kernel_ms: f64,
    pub total_ms: f64,
}

/// A scene resident on one device (opaque).
#[repr(C)]
pub struct rtgpu_context {
    _private: [u8; 0],
}

unsafe extern "C" {
    pub fn rtgpu_abi_version() -> u32;
    pub fn rtgpu_last_error() -> *const c_char;
    pub fn rtgpu_device_count() -> c_int;
    pub fn rtgpu_rows_count(rows: *const rtgpu_rows, vsize: u32) -> u32;
    pub fn rtgpu_rows_list(rows: *const rtgpu_rows, vsize: u32, out_rows: *mut u32, capacity: u32) -> u32;
    pub fn rtgpu_render(
        scene: *const rtgpu_scene,
        camera: *const rtgpu_camera,
        opts: *const rtgpu_opts,
        out_rgb: *mut c_void, // f64 elements in RTGPU_PRECISION_F64, f32 in RTGPU_PRECISION_F32
        out_rgb8: *mut u8,
        stats: *mut rtgpu_stats,
    ) -> c_int;
    pub fn rtgpu_context_create(scene: *const rtgpu_scene, device: c_int, out_context: *mut *mut rtgpu_context) -> c_int;
    pub fn rtgpu_context_destroy(context: *mut rtgpu_context);
    pub fn rtgpu_context_render_device(
        context: *mut rtgpu_context,
        camera: *const rtgpu_camera,
        opts: *const rtgpu_opts,
        rows: *const rtgpu_rows,
        d_out_rgb: *mut c_void,
        d_out_rgb8: *mut u8,
        d_counters: *mut u64,
        cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn rtgpu_context_render(
        context: *mut rtgpu_context,
        camera: *const rtgpu_camera,
        opts: *const rtgpu_opts,
        rows: *const rtgpu_rows,
        out_rgb: *mut c_void,
        out_rgb8: *mut u8,
        stats: *mut rtgpu_stats,
    ) -> c_int;
    pub fn rtgpu_last_family() -> c_int;
    pub fn rtgpu_context_frame_records(context: *mut rtgpu_context, out: *mut u64) -> c_int;
    pub fn rtgpu_context_launch_count(context: *mut rtgpu_context) -> u64;
    pub fn rtgpu_host_alloc(bytes: usize) -> *mut c_void;
    pub fn rtgpu_host_free(ptr: *mut c_void);
    pub fn rtgpu_measure_fma_peak(device: c_int, precision: u32, out_tflops: *mut f64, out_ms: *mut f64) -> c_int;
    pub fn rtgpu_selftest_arith(
        device: c_int,
        a: *const f64,
        b: *const f64,
        n: usize,
        out_div_mismatches: *mut u64,
        out_sqrt_mismatches: *mut u64,
        out_div_fallbacks: *mut u64,
        out_sqrt_fallbacks: *mut u64,
    ) -> c_int;
    pub fn rtgpu_debug_probe(
        context: *mut rtgpu_context,
        camera: *const rtgpu_camera,
        kind: u32,
        input: *const f64,
        n_in: usize,
        out: *mut f64,
        n_out: usize,
    ) -> c_int;
    pub fn rtgpu_debug_color_at(
        context: *mut rtgpu_context,
        origin: *const f64,
        direction: *const f64,
        opts: *const rtgpu_opts,
        out_rgb: *mut f64,
    ) -> c_int;
}
