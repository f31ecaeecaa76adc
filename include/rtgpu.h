/*
 * rtgpu.h — C ABI of the B200-native camera render pass.
 *
 * This is the drop-in boundary for ONE hot path of przemo199/ray-tracer-challenge-rs:
 *     Camera::render / Camera::render_parallel     (ray-tracer/src/composites/camera.rs:79-112)
 * and everything those call per pixel (World::color_at, world.rs:89-95, and below).
 * The reference has no FFI of its own; the only seam is the `RenderingMode` match in
 * ray-tracer-cli/src/main.rs:18-21.  A new `RenderingMode::Gpu` arm calls
 * `Camera::render_gpu(&self, &World) -> Canvas`, which flattens the World (trait objects ->
 * the plain arrays below) and calls `rtgpu_render`.  INTEGRATION.md shows that Rust binding.
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer unless the name starts with `d_`.
 *   - the caller owns every array; the library copies what it needs and never keeps a caller
 *     pointer past the return of the call it was passed to.
 *   - every function returns RTGPU_OK (0) or a negative rtgpu_status; the message for the last
 *     failure on the calling thread is `rtgpu_last_error()`.
 *   - there is NO CPU fallback: without a usable CUDA device the render entry points fail with
 *     RTGPU_ERR_NO_DEVICE.
 *   - all arithmetic inputs are IEEE-754 binary64, exactly the values the reference holds.
 */
#ifndef RTGPU_H
#define RTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTGPU_ABI_VERSION 1u

typedef enum rtgpu_status {
    RTGPU_OK = 0,
    RTGPU_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, index out of range, bad enum value        */
    RTGPU_ERR_UNSUPPORTED = -2,      /* scene content the device path does not implement       */
    RTGPU_ERR_NO_DEVICE = -3,        /* no CUDA device / driver (gpu mode never falls back)    */
    RTGPU_ERR_CUDA = -4,             /* a CUDA runtime call failed; see rtgpu_last_error()     */
    RTGPU_ERR_OUT_OF_MEMORY = -5
} rtgpu_status;

/* Shape kinds = the six `impl Shape` types (ray-tracer/src/shapes/{sphere,plane,cube,cylinder,
 * cone,triangle}.rs).  The numeric values are part of the ABI. */
typedef enum rtgpu_shape_type {
    RTGPU_SPHERE = 0,
    RTGPU_PLANE = 1,
    RTGPU_CUBE = 2,
    RTGPU_CYLINDER = 3,
    RTGPU_CONE = 4,
    RTGPU_TRIANGLE = 5,
    RTGPU_SHAPE_TYPE_COUNT = 6
} rtgpu_shape_type;

/* Pattern kinds = `impl Pattern` types (ray-tracer/src/patterns/*.rs).  RTGPU_PATTERN_TEST is the
 * reference's crate-private TestPattern (patterns/pattern.rs:29-60: colour = the pattern-space
 * point), kept so the reference's own refraction test (world.rs:547-571) can be replayed. */
typedef enum rtgpu_pattern_type {
    RTGPU_PATTERN_STRIPE = 0,
    RTGPU_PATTERN_GRADIENT = 1,
    RTGPU_PATTERN_RING = 2,
    RTGPU_PATTERN_CHECKER = 3,
    RTGPU_PATTERN_COMPLEX = 4, /* patterns/complex_pattern.rs:24-33; children via pat_child_a/b */
    RTGPU_PATTERN_TEST = 5,
    RTGPU_PATTERN_TYPE_COUNT = 6
} rtgpu_pattern_type;

/* Per-material scalar block, in this order (composites/material.rs:9-20). */
enum {
    RTGPU_MAT_AMBIENT = 0,
    RTGPU_MAT_DIFFUSE = 1,
    RTGPU_MAT_SPECULAR = 2,
    RTGPU_MAT_SHININESS = 3,
    RTGPU_MAT_REFLECTIVENESS = 4,
    RTGPU_MAT_TRANSPARENCY = 5,
    RTGPU_MAT_REFRACTIVE_INDEX = 6,
    RTGPU_MAT_PARAM_COUNT = 7
};

/*
 * The flattened World (composites/world.rs:9-12): structure-of-arrays, one entry per element of
 * `world.shapes` IN THE WORLD'S ORDER (the order decides hit tie-breaks, intersections.rs:13-18
 * + the stable sort in world.rs:34).  3x4 matrices are rows 0..2 of the reference's row-major
 * `transformation_inverse` (primitives/matrix.rs:8); row 3 never enters the per-ray arithmetic
 * (matrix.rs:332-362 evaluates rows 0..2 only).
 */
typedef struct rtgpu_scene {
    uint32_t abi_version; /* = RTGPU_ABI_VERSION */

    /* shapes[S] */
    uint32_t n_shapes;
    const uint8_t *shape_type;      /* [S]     rtgpu_shape_type                                     */
    const double *shape_inv;        /* [S*12]  transformation_inverse rows 0..2 (shapes/shape.rs:31) */
    const double *shape_min;        /* [S]     cylinder/cone `min` (cylinder.rs:12, cone.rs:12); ignored otherwise */
    const double *shape_max;        /* [S]     cylinder/cone `max`                                   */
    const uint8_t *shape_closed;    /* [S]     cylinder/cone `closed`                                */
    const int32_t *shape_triangle;  /* [S]     index into tri_* for RTGPU_TRIANGLE, else -1; may be NULL when n_triangles == 0 */
    const uint32_t *shape_material; /* [S]     index into the material arrays                        */
    const uint32_t *shape_eq_class; /* [S]     lowest index of a shape that is `==` this one by value
                                               (shapes/shape.rs:34-38, dyn_partial_eq.rs:9-16); decides
                                               "same shape" in the refraction container walk
                                               (composites/intersection.rs:38,47)                    */

    /* triangles[T] (shapes/triangle.rs:9-18) */
    uint32_t n_triangles;
    const double *tri_vertex_1; /* [T*3] */
    const double *tri_edge_1;   /* [T*3] vertex_2 - vertex_1 */
    const double *tri_edge_2;   /* [T*3] vertex_3 - vertex_1 */
    const double *tri_normal;   /* [T*3] normalize(edge_2 x edge_1) (triangle.rs:24) */

    /* materials[M] (composites/material.rs:9-20) */
    uint32_t n_materials;
    const double *mat_color;         /* [M*3] */
    const double *mat_params;        /* [M*RTGPU_MAT_PARAM_COUNT] */
    const uint8_t *mat_casts_shadow; /* [M] */
    const int32_t *mat_pattern;      /* [M] index into the pattern arrays, -1 = none */

    /* patterns[Q] (patterns/*.rs) */
    uint32_t n_patterns;
    const uint8_t *pat_type;    /* [Q]    rtgpu_pattern_type */
    const double *pat_color_a;  /* [Q*3]  */
    const double *pat_color_b;  /* [Q*3]  */
    const double *pat_inv;      /* [Q*12] pattern transformation_inverse rows 0..2 */
    const int32_t *pat_child_a; /* [Q]    RTGPU_PATTERN_COMPLEX only, else -1 */
    const int32_t *pat_child_b; /* [Q]    */

    /* lights[L] (primitives/light.rs:6-9), in `world.lights` order (the fold order of world.rs:45-52) */
    uint32_t n_lights;
    const double *light_position;  /* [L*3] */
    const double *light_intensity; /* [L*3] */
} rtgpu_scene;

/* The Camera's per-ray fields (composites/camera.rs:10-19).  The host computes them exactly as
 * Camera::new / set_transformation do (camera.rs:25-49,114-127). */
typedef struct rtgpu_camera {
    uint32_t hsize;     /* horizontal_size */
    uint32_t vsize;     /* vertical_size   */
    double half_width;
    double half_height;
    double pixel_size;
    double inv[12];     /* transformation_inverse rows 0..2 */
    double origin[3];   /* transformation_inverse * Point::ORIGIN */
} rtgpu_camera;

typedef enum rtgpu_precision {
    RTGPU_PRECISION_F64 = 0, /* parity mode: the reference's operations, order and fma sites     */
    RTGPU_PRECISION_F32 = 1  /* fast mode: binary32 arithmetic, own epsilon, stated tolerance    */
} rtgpu_precision;

/* Which rows of the image a call renders.  Rows are grouped into bands of `band_rows` rows; band b
 * belongs to shard (b % shard_count).  A call renders the rows of shard `shard_index`, in
 * increasing row order.  {band_rows=0, shard_index=0, shard_count=1} = the whole frame. */
typedef struct rtgpu_rows {
    uint32_t band_rows;
    uint32_t shard_index;
    uint32_t shard_count;
} rtgpu_rows;

typedef struct rtgpu_opts {
    uint32_t precision;   /* rtgpu_precision                                                        */
    uint32_t max_depth;   /* World::MAX_REFLECTION_ITERATIONS (world.rs:15) = 6 for parity; 0..15     */
    int32_t n_gpus;       /* rtgpu_render only: devices 0..n_gpus-1 share the frame by row bands; <=0 -> 1 */
    uint32_t band_rows;   /* rtgpu_render only: band height for the multi-GPU split; 0 -> 16          */
    uint32_t flags;       /* RTGPU_FLAG_*                                                           */
} rtgpu_opts;

#define RTGPU_FLAG_NONE 0u
/* Kernel family (both produce the same pixels and counters, bit for bit).  Neither flag = auto: RTGPU_FAMILY=
 * persistent|wavefront|auto in the environment, else the library measures — for each (scene, frame shape) a
 * context times its first two frames of each family with CUDA events and renders the rest with the faster
 * one; scenes without reflective or transparent materials always take PERSISTENT.
 *   PERSISTENT : one launch; every lane walks one pixel's recursion tree with an explicit stack.
 *   WAVEFRONT  : one launch per recursion level over queues of rays (+ a bottom-up combine per level).
 *                rtgpu_context_render_device then blocks until the frame is complete (it has to verify that
 *                its ray / node buffers were large enough, and renders again with larger ones if not). */
#define RTGPU_FLAG_WAVEFRONT 1u
#define RTGPU_FLAG_PERSISTENT 2u

/* Work counters (integers, identical between devices, shardings and the CPU oracle) and timings. */
typedef struct rtgpu_stats {
    uint64_t rays_primary; /* one per pixel (camera.rs:108)                                          */
    uint64_t rays_shadow;  /* one per light per hit node (world.rs:45-52)                            */
    uint64_t rays_reflect; /* world.rs:120-126                                                       */
    uint64_t rays_refract; /* world.rs:136-154 (not counted when total internal reflection)          */
    uint64_t hit_nodes;    /* internal_color_at calls that found a hit (world.rs:79-84)               */
    uint64_t pixels;       /* pixels rendered by this call                                            */
    double kernel_ms;      /* device time of the render kernel(s), CUDA events; max over devices      */
    double total_ms;       /* host wall clock of the whole call                                       */
} rtgpu_stats;

typedef struct rtgpu_context rtgpu_context; /* a scene resident on ONE device */

/* -- library ------------------------------------------------------------------------------- */
uint32_t rtgpu_abi_version(void);
const char *rtgpu_last_error(void);
/* Number of usable CUDA devices; 0 when there is no driver or device (never negative). */
int rtgpu_device_count(void);

/* Number of rows / the row list `rows` selects out of `vsize` (host-side helper shared by every
 * caller that shards a frame).  `out_rows` may be NULL; otherwise it receives the row indices. */
uint32_t rtgpu_rows_count(const rtgpu_rows *rows, uint32_t vsize);
uint32_t rtgpu_rows_list(const rtgpu_rows *rows, uint32_t vsize, uint32_t *out_rows, uint32_t capacity);

/* -- one-shot render: replaces Camera::render / render_parallel (camera.rs:79-112) ---------- */
/*
 * Renders `camera.hsize * camera.vsize` pixels of `scene` and writes the linear colours the
 * reference's Canvas holds (composites/canvas.rs:13-17): out_rgb[(y*hsize + x)*3 + c], c = r,g,b.
 * out_rgb8, if not NULL, additionally receives the bytes Canvas::to_png_file would encode
 * (canvas.rs:117-123): clamp to [0,1], * 255, round half away from zero.  Either may be NULL, not
 * both.  Uses opts->n_gpus devices (row bands, no collective) and blocks until both host buffers
 * are complete.  `stats` may be NULL.
 * Element type of out_rgb: `double` in RTGPU_PRECISION_F64 (hsize*vsize*3 doubles), `float` in
 * RTGPU_PRECISION_F32 (hsize*vsize*3 floats) — hence `void *`; the same holds for rtgpu_context_render.
 */
int rtgpu_render(const rtgpu_scene *scene, const rtgpu_camera *camera, const rtgpu_opts *opts,
                 void *out_rgb, uint8_t *out_rgb8, rtgpu_stats *stats);

/* -- resident scene: what a caller rendering many frames / one rank of a multi-process job uses */
int rtgpu_context_create(const rtgpu_scene *scene, int device, rtgpu_context **out_context);
void rtgpu_context_destroy(rtgpu_context *context);

/*
 * Launches the render of the rows `rows` selects on `cuda_stream` (a cudaStream_t passed as
 * void*, NULL = the legacy default stream) and returns without synchronising.  d_out_rgb /
 * d_out_rgb8 are DEVICE pointers on the context's device, compact over the selected rows:
 * d_out_rgb[(k*hsize + x)*3 + c] for the k-th selected row.  d_counters is a DEVICE pointer to 6
 * uint64 (order of rtgpu_stats' first six fields) that the kernel adds into, or NULL.
 * In RTGPU_PRECISION_F32 d_out_rgb is `float*`-typed storage (still 3 values per pixel).
 * Asynchrony: only the PERSISTENT family returns without waiting.  A WAVEFRONT frame is complete on return, and
 * in automatic mode the calibration frames of a (scene, frame shape) wait for the previous frame's events; pass
 * RTGPU_FLAG_PERSISTENT to keep a pipeline of frames fully asynchronous.
 */
int rtgpu_context_render_device(rtgpu_context *context, const rtgpu_camera *camera,
                                const rtgpu_opts *opts, const rtgpu_rows *rows, void *d_out_rgb,
                                uint8_t *d_out_rgb8, uint64_t *d_counters, void *cuda_stream);

/* Host-buffer variant on the resident scene: H2D of the camera, kernel, D2H of the selected rows
 * into the FULL-FRAME host buffers at their row offsets; blocks until done. */
int rtgpu_context_render(rtgpu_context *context, const rtgpu_camera *camera, const rtgpu_opts *opts,
                         const rtgpu_rows *rows, void *out_rgb, uint8_t *out_rgb8,
                         rtgpu_stats *stats);

/* Kernel family the calling thread's most recent render ran: 0 = persistent, 1 = wavefront (what the automatic
 * choice settled on; benchmarks report it). */
int rtgpu_last_family(void);

/* Records the wavefront family moved through HBM for the context's most recent host-buffer frame (zeros for a
 * persistent-family frame): out[0] = hits queued over all levels (each written once and read once as a ray record),
 * out[1] = node records (each written once, read once by the combine pass), out[2] / out[3] = bytes per ray / node
 * record.  Benchmarks turn this into the frame's algorithmic queue traffic without a profiler. */
int rtgpu_context_frame_records(rtgpu_context *context, uint64_t out[4]);

/* Render kernels this library has launched for the context since it was created (level / bin / combine / commit /
 * status kernels of the wavefront family, one launch per frame of the persistent family).  Benchmarks report the
 * difference across their timed region as the number of kernels that ran in it. */
uint64_t rtgpu_context_launch_count(rtgpu_context *context);

/* -- pinned host memory ---------------------------------------------------------------------- */
/* Page-locked, device-mapped host memory.  When out_rgb / out_rgb8 of a host-buffer render live in such
 * memory (these functions, cudaHostAlloc, cudaHostRegister, torch pin_memory ...) the persistent kernel writes
 * the finished pixels straight into them over PCIe while the render is still running (no staging copy);
 * ordinary pageable buffers — and the wavefront family — go through a device staging buffer and a copy. */
void *rtgpu_host_alloc(size_t bytes);
void rtgpu_host_free(void *ptr);

/* -- measurement helpers ------------------------------------------------------------------- */
/* Dependent-free DFMA / FFMA chains on every SM: the measured FP64 / FP32 FMA-pipe peak that the
 * roofline of this path is quoted against (MEASURED_PEAKS.json has HBM and bf16 only).
 * Returns TFLOP/s (2 flop per fma) in *out_tflops; runs on `device`, on `cuda_stream`. */
int rtgpu_measure_fma_peak(int device, uint32_t precision, double *out_tflops, double *out_ms);

/* Self-test of the IEEE-exact fast division / square root the FP64 kernels use (csrc/rt_arith.cuh):
 * for every i < n evaluates a[i] / b[i] and sqrt(a[i]) with the restructured sequences and with
 * the native operators on `device`, and counts bitwise differences where the fast path declared
 * itself valid (must be 0) and how often it fell back to the native operator.  a, b: host arrays. */
int rtgpu_selftest_arith(int device, const double *a, const double *b, size_t n, uint64_t *out_div_mismatches,
                         uint64_t *out_sqrt_mismatches, uint64_t *out_div_fallbacks, uint64_t *out_sqrt_fallbacks);

/* -- known-answer probes -------------------------------------------------------------------- */
/* ONE device thread evaluates the device functions the render kernels are built from on the caller's inputs, so
 * that the reference's own unit tests can be replayed on the GPU value by value (tests/test_gpu_kat.py).  Test
 * infrastructure: not on the render path.  `in` / `out` are host arrays of doubles; integers (shape = index into
 * world.shapes, material, pattern, light, flags) travel as doubles.  Scene queries (IN_SHADOW, PREPARE) need a
 * scene below the BVH threshold.  f64 only. */
typedef enum rtgpu_probe_kind {
    RTGPU_PROBE_RAY_FOR_PIXEL = 0, /* Camera::ray_for_pixel, camera.rs:52-68.  in: px, py.  out: origin[3], direction[3]       */
    RTGPU_PROBE_INTERSECT = 1,     /* Ray::intersect, ray.rs:35-49 + shapes/*.rs local_intersect.  in: shape, origin[3],
                                      direction[3].  out: n, t[4] (push order)                                                 */
    RTGPU_PROBE_LOCAL_NORMAL = 2,  /* local_normal_at of the shape's type.  in: shape, object-space point[3].  out: normal[3]   */
    RTGPU_PROBE_NORMAL = 3,        /* Shape::normal_at, shapes/shape.rs:22-27.  in: shape, world point[3].  out: normal[3]      */
    RTGPU_PROBE_PATTERN = 4,       /* Pattern::color_at_shape, patterns/pattern.rs:10-14.  in: pattern, shape, point[3].
                                      out: colour[3]                                                                           */
    RTGPU_PROBE_LIGHTING = 5,      /* Material::lighting, composites/material.rs:53-114.  in: material, shape, light
                                      position[3], light intensity[3], point[3], eye[3], normal[3], in_shadow.  out: colour[3] */
    RTGPU_PROBE_IN_SHADOW = 6,     /* World::is_in_shadow, composites/world.rs:98-112.  in: light, point[3].  out: 0 / 1        */
    RTGPU_PROBE_PREPARE = 7,       /* Intersections::hit + Intersection::prepare_computations, composites/intersection.rs:21-75,
                                      + ComputedHit::schlicks_approximation, computed_hit.rs:50-68.  in: origin[3],
                                      direction[3], k (-1 = the hit), n, then n x (t, shape) = the sorted list.  out: found,
                                      distance, shape, inside, point[3], over[3], under[3], eye[3], normal[3], reflect[3],
                                      n1, n2, schlick  (25 values)                                                            */
    RTGPU_PROBE_QUANTISE = 8,      /* Canvas 8-bit quantisation, composites/canvas.rs:117-123.  in: v.  out: byte              */
    RTGPU_PROBE_COLLECT = 9        /* World::collect_intersections before its sort, world.rs:25-33.  in: origin[3],
                                      direction[3].  out: count, then (t, shape) pairs in push order                           */
} rtgpu_probe_kind;
int rtgpu_debug_probe(rtgpu_context *context, const rtgpu_camera *camera, uint32_t kind, const double *in, size_t n_in,
                      double *out, size_t n_out);
/* World::color_at (world.rs:89-95) of an arbitrary ray — origin, direction as given, not normalised — with
 * opts->max_depth remaining iterations, through the unmodified render kernels of the family opts->flags names
 * (default PERSISTENT): a 1 x 1 frame whose only ray is the caller's. */
int rtgpu_debug_color_at(rtgpu_context *context, const double origin[3], const double direction[3], const rtgpu_opts *opts,
                         double out_rgb[3]);

#ifdef __cplusplus
}
#endif
#endif /* RTGPU_H */
