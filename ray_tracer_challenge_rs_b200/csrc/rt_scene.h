// rt_scene.h — device-side scene layout shared by the host packer (rtgpu.cu) and the kernels.
//
// The reference keeps `Vec<Box<dyn Shape>>` (composites/world.rs:9-12).  On the device the world is
// ONE contiguous blob of `T` (double in parity mode, float in fast mode) plus one blob of int32
// metadata, with the shapes grouped by type so that the intersection loops contain no dispatch:
//
//   real blob : [ shape geometry S x 16 ][ triangle data NT x 12 ][ materials M x 12 ]
//               [ patterns Q x 18 ][ lights L x 6 ][ cull spheres S x 4 ][ BVH node boxes N x 12 ]
//   int  blob : [ shape meta S x 8 ][ material meta M x 2 ][ pattern meta Q x 4 ][ BVH children N x 2 ]
//
// Shape order: [ uniform part ][ BVH part ].
//   uniform part : shapes every lane tests in lockstep, grouped by type (`type_begin`), world order inside
//                  a type.  Small scenes: every shape (bounded ones behind a bounding-sphere pre-test).
//                  BVH scenes: only the UNBOUNDED shapes (planes, untruncated cylinders / cones).
//   BVH part     : the bounded shapes of a large scene, in BVH leaf order (`n_bvh_nodes > 0`).
//
// Both blobs are staged into shared memory by every CTA when they fit (always, for the shipped
// scenes: <= 3 KB), otherwise they are read through L1/L2 from global memory.
#pragma once
#include <stdint.h>

namespace rt {

// ---- shape geometry record: 16 reals (128 B in f64, 16-byte aligned rows) ---------------------
//  [0..11]  transformation_inverse rows 0..2 (row-major 3x4)         shapes/shape.rs:31
//  [12]     min   (cylinder / cone)                                   cylinder.rs:12, cone.rs:12
//  [13]     max
//  [14..15] unused (keeps records 16-byte aligned for 128-bit shared loads)
constexpr int SHAPE_REALS = 16;
constexpr int SHAPE_MIN = 12;
constexpr int SHAPE_MAX = 13;

// ---- shape meta record: 8 int32 (the first four are read as one int4) ---------------------------
//  [0] orig index in world.shapes (tie-breaks, intersections.rs:13-18 + world.rs:34)
//  [1] material index
//  [2] flags | (shape type << FLAG_TYPE_SHIFT)
//  [3] eq_class (lowest orig index of a value-equal shape; intersection.rs:38,47)
//  [4] triangle slot (index into the triangle records), -1 otherwise
constexpr int SHAPE_INTS = 8;
constexpr int FLAG_TYPE_SHIFT = 8;
constexpr int FLAG_CLOSED = 1;        // cylinder.rs:14 / cone.rs:14
constexpr int FLAG_CASTS_SHADOW = 2;  // material.casts_shadow of the shape's material (world.rs:108)
// The refraction-container walk treats value-equal shapes as one (intersection.rs:47).  A class of
// k identical shapes toggles k times per distinct distance: for even k it never stays in the
// container list, for odd k it behaves like ONE shape whose last push comes from its highest-index
// member.  The packer therefore marks exactly that member of every odd-sized class.
constexpr int FLAG_CONTAINER_REP = 4;

// ---- triangle record: 12 reals (shapes/triangle.rs:9-18) ----------------------------------------
//  vertex_1[3], edge_1[3], edge_2[3], normal[3]; indexed by (sorted position - first triangle)
constexpr int TRI_REALS = 12;

// ---- material record: 12 reals (composites/material.rs:9-20) -------------------------------------
//  color[3], ambient, diffuse, specular, shininess, reflectiveness, transparency, refractive_index, pad[2]
constexpr int MAT_REALS = 12;
constexpr int MAT_AMBIENT = 3, MAT_DIFFUSE = 4, MAT_SPECULAR = 5, MAT_SHININESS = 6, MAT_REFLECTIVENESS = 7,
              MAT_TRANSPARENCY = 8, MAT_REFRACTIVE_INDEX = 9;
//  material meta: [0] pattern index (-1 none), [1] casts_shadow
constexpr int MAT_INTS = 2;

// ---- pattern record: 18 reals: color_a[3], color_b[3], transformation_inverse rows 0..2 [12] ----
constexpr int PAT_REALS = 18;
//  pattern meta: [0] type (rtgpu_pattern_type), [1] child_a, [2] child_b, [3] pad
constexpr int PAT_INTS = 4;

// ---- light record: 6 reals: position[3], intensity[3] (primitives/light.rs:6-9) -----------------
constexpr int LIGHT_REALS = 6;

// ---- cull record: 4 reals: world-space bounding sphere centre[3], radius^2 (+inf = unbounded) ----
// Not part of the reference: a conservative pre-test.  A ray whose supporting line stays outside the
// (slightly inflated) sphere cannot produce an intersection in the reference's arithmetic either, so
// skipping the exact test changes no result (rt_kernel.cuh, trace_unified).
constexpr int CULL_REALS = 4;

// ---- BVH node (not in the reference; csrc/rt_bvh.h builds it): 12 reals = the two children's
// inflated world-space boxes (lo[3], hi[3] each); 2 ints = child references, >= 0 an inner node,
// < 0 the leaf shape at sorted position ~ref.  One shape per leaf.
constexpr int BVH_REALS = 12;
constexpr int BVH_INTS = 2;
// The traversal reads a single-precision copy of the node: 16 words = 64 bytes, 16-byte aligned, in the int blob:
//   [0..11] the two children's boxes as floats ROUNDED OUTWARDS from the double boxes (lo[3], hi[3] each),
//   [12..13] the child references, [14..15] unused.
// The box test is a conservative filter (rt_kernel.cuh box_hit32): single precision with rigorous margins costs a
// quarter of the instructions of the double test and halves the bytes per node visit.
constexpr int BVH32_WORDS = 16;
constexpr int BVH_MAX_DEPTH = 60;  // device traversal stack; the builder falls back to median splits to stay below

constexpr int NUM_SHAPE_TYPES = 6;  // order = rtgpu_shape_type: sphere, plane, cube, cylinder, cone, triangle

// What a kernel needs to find its way around the two blobs.  Passed by value as a kernel parameter.
struct SceneLayout {
    uint32_t n_shapes;
    uint32_t type_begin[NUM_SHAPE_TYPES + 1];  // sorted positions [type_begin[t], type_begin[t+1]) hold type t
    uint32_t n_materials, n_patterns, n_lights;
    uint32_t tri_off, mat_off, pat_off, light_off, cull_off, bvh_off;  // offsets into the real blob, in reals
    uint32_t mat_meta_off, pat_meta_off, bvh_meta_off;  // offsets into the int blob, in int32
    uint32_t bvh32_off;                             // single-precision BVH nodes (BVH32_WORDS each), int blob, multiple of 4
    float bvh_coord_max;                            // largest |coordinate| of any BVH box (sizes the origin margin of box_hit32)
    uint32_t cull32_off;                            // single-precision cull records (4 words each: centre, padded radius^2), int blob, multiple of 4
    float cull_coord_max;                           // largest |coordinate| of any finite cull-sphere centre, rounded up
    uint32_t n_bvh_nodes;                           // > 0: the bounded shapes are reached through a BVH
    int32_t bvh_root;                               // root reference (>= 0 node, < 0 single leaf)
    uint32_t n_reals, n_ints;                       // blob sizes
    uint32_t in_shared;                             // 1: CTAs stage both blobs in shared memory
    // shrink factor of the bounding-sphere pre-test (>> the rounding of the pre-test itself).  Lives here so that
    // the loop reads it as a constant-bank operand instead of rebuilding the literal every iteration.
    float cull_shrink32;
    double cull_shrink64;
};

// Camera (composites/camera.rs:10-19) + the row selection of one launch (include/rtgpu.h rtgpu_rows).
template <typename T>
struct CameraParams {
    T half_width, half_height, pixel_size;
    T inv[12];
    T origin[3];
    uint32_t hsize, vsize;
    // rows: with q = k / band_rows the compact row k of this launch is image row
    //   ((q / band_take) * shard_count + shard_index + q % band_take) * band_rows + k % band_rows
    // i.e. out of every `shard_count` consecutive bands the launch renders `band_take` of them, starting at band
    // `shard_index` (band_take = 1: the public rtgpu_rows selection; > 1: the unequal parts of a chunked host render)
    uint32_t n_rows;  // rows rendered by this launch
    uint32_t band_rows, shard_index, shard_count, band_take;
    uint32_t max_depth;  // World::MAX_REFLECTION_ITERATIONS (world.rs:15)
    uint32_t tile_stride;  // coprime to the number of 8x4 tiles: scattered tile order (rt_kernel.cuh)
    uint32_t out_full_frame;  // 1: outputs are full-frame buffers indexed by image row (zero-copy into the
                              //    caller's pinned host Canvas); 0: compact over the rows of this launch
    uint32_t probe_ray;       // 1 (rtgpu_debug_color_at): every pixel's ray is (origin, inv[0..2]) as given, not normalised
                              //    — World::color_at of an arbitrary ray through the unmodified kernels
};

// Work counters, in the order of the first six fields of rtgpu_stats.
constexpr int COUNTER_PRIMARY = 0, COUNTER_SHADOW = 1, COUNTER_REFLECT = 2, COUNTER_REFRACT = 3, COUNTER_HIT_NODES = 4,
              COUNTER_PIXELS = 5, NUM_COUNTERS = 6;

}  // namespace rt
