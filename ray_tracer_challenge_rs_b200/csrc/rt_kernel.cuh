// rt_kernel.cuh — the camera render pass as ONE persistent sm_100a kernel.
//
// Replaces, per pixel, Camera::render_parallel (composites/camera.rs:97-112) and everything under
// World::color_at (composites/world.rs:89-95).  Citations are paths under the reference's
// `ray-tracer/src/`.
//
// Design (B200-first, not a translation of the recursive, trait-object reference):
//   * every lane is a small state machine that owns one pixel at a time and always has exactly
//     ONE ray in flight: a radiance ray (nearest hit), a shadow ray (any hit) or a "container"
//     re-trace (refractive-index bookkeeping).  All three share ONE brute-force intersection loop
//     over the type-sorted shape table, so the dominant loop runs warp-converged no matter how far
//     the lanes' recursion trees have drifted apart;
//   * the reference's recursion (reflected_color / refracted_color, world.rs:114-157) becomes an
//     explicit post-order stack of <= max_depth frames per lane.  Children are fully evaluated and
//     only then scaled and added, exactly like the reference (`surface + reflected + refracted`),
//     so the f64 result is bit-comparable;
//   * the reference sorts every intersection list (world.rs:34) only to find the hit and to walk the
//     refraction containers; here the hit is a running (t, world-order) minimum and the container
//     walk is a second pass that needs no list (see ContainerAcc);
//   * lanes that finish a pixel refill from a warp-private chunk of pixel slots (8x4 tiles), so a
//     warp stays full until the frame runs out of pixels;
//   * arithmetic: this file must be compiled with -fmad=false.  `fma()` appears exactly where the
//     reference calls `mul_add`; everything else keeps the reference's operation order.
#pragma once

#include <cfloat>
#include <climits>
#include <cstdint>

#include "rt_arith.cuh"
#include "rt_scene.h"

namespace rt {

#ifndef RT_ACC_SPLIT
#define RT_ACC_SPLIT 1
#endif
#ifndef RT_BVH_WATCHDOG
#define RT_BVH_WATCHDOG 0
#endif
#ifndef RT_TMA_STAGE
#define RT_TMA_STAGE 1  // scene tables -> shared memory by cp.async.bulk + mbarrier (0: plain loads).  Must precede stage_scene.
#endif
#ifndef RT_F32_OFFSET_ULPS
#define RT_F32_OFFSET_ULPS 16  // f32 fast mode: over / under point offset in ulps of max(|point|, distance, 1); 8 .. 4096 swept in profiles/r2_notes.md
#endif
#ifndef RT_CANDIDATES
#define RT_CANDIDATES 0  // 1: per-lane candidate masks, exact tests candidate by candidate (cover -1 %, every other scene +4..8 %: profiles/r2_notes.md)
#endif
#ifndef RT_LEAN_LOOP
#define RT_LEAN_LOOP 1  // trace_unified: pointer-driven shape loop with a single branch per culled shape
#endif
#ifndef RT_CULL
#define RT_CULL 1  // bounding-sphere pre-test in the intersection loop (exact results either way)
#endif
#ifndef RT_BOUNDS_CHECK
#define RT_BOUNDS_CHECK 0  // 1: every table / queue / node / permutation index is range-checked on the device; a bad one traps
#endif
#if RT_BOUNDS_CHECK
#define RT_CHECK(cond) do { if (!(cond)) { printf("rtgpu bounds check failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define RT_CHECK(cond) do { } while (0)
#endif
#ifndef RT_CULL_F32
#define RT_CULL_F32 1  // f64 kernels: the pre-test runs in single precision with proven margins (trace_unified); exact results either way
#endif

#ifndef RT_STRICT_SIGNED_ZERO
// 1: keep the reference's `0.0 + ...` fold seeds and `+ m[r][3] * w` terms of Matrix*Point/Vector
//    (matrix.rs:332-362) literally.  `0.0 + p` differs from `p` only for p = -0.0, and `x + m*0.0`
//    differs from `x` only for x = -0.0, so the terms can only turn a -0.0 component into +0.0.
//    Nothing on the path can tell the two zeros apart (every division is guarded by an EPSILON or
//    `> 0.0` test, comparisons treat them as equal, `floor(-0.0) as i64` = 0): the default build
//    drops the 12 extra FP64 operations per ray-shape test.  DESIGN.md, "signed zeros".
#define RT_STRICT_SIGNED_ZERO 0
#endif

template <typename T>
struct Real;
template <>
struct Real<double> {
    static __device__ __forceinline__ double eps() { return 0.00000008; }  // consts.rs:2
    // computed_hit.rs:33-34: over / under point = point +- normal * EPSILON, a constant
    static __device__ __forceinline__ double offset(double, double, double, double) { return 0.00000008; }
    static __device__ __forceinline__ double max() { return DBL_MAX; }
    static __device__ __forceinline__ double cull_shrink() { return 1.0 - 1.0e-9; }  // >> f64 rounding of the pre-test
};
template <>
struct Real<float> {
    static __device__ __forceinline__ float eps() { return 0.00000008f; }
    // 8e-8 is below one f32 ulp at |x| >= 1: the fast mode needs its own offset, and a constant one is
    // either too small far from the origin or too large for thin shapes: it scales with the magnitude of the hit
    // point and of the hit distance (the rounding error of o + d * t is a few ulps of those), RT_F32_OFFSET_ULPS of them.
    static __device__ __forceinline__ float offset(float px, float py, float pz, float t) {
        const float scale = fmaxf(fmaxf(fmaxf(fabsf(px), fabsf(py)), fmaxf(fabsf(pz), fabsf(t))), 1.0f);
        return (float)RT_F32_OFFSET_ULPS * 1.1920929e-7f * scale;
    }
    static __device__ __forceinline__ float max() { return FLT_MAX; }
    static __device__ __forceinline__ float cull_shrink() { return 1.0f - 1.0e-3f; }
};

template <typename T>
struct V3 {
    T x, y, z;
};
template <typename T>
struct Ray {
    V3<T> o, d;
};

#define RT_DEV __device__ __forceinline__
// Rarely executed or multiply used helpers are kept out of line: the kernel's hot working set must
// stay inside the 32 KB L1.5 instruction cache (the first culled build spent 44 % of its stall
// samples on instruction fetch, profiles/r1_notes.md).
#ifndef RT_OUTLINE
#define RT_OUTLINE 1
#endif
#if RT_OUTLINE
#define RT_COLD __device__ __noinline__
#else
#define RT_COLD __device__ __forceinline__
#endif

template <typename T> RT_DEV V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> RT_DEV V3<T> ld3(const T* p) { return mk<T>(p[0], p[1], p[2]); }
template <typename T> RT_DEV V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> RT_DEV V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> RT_DEV V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> RT_DEV V3<T> hadamard(V3<T> a, V3<T> b) { return mk<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> RT_DEV V3<T> neg(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> RT_DEV T sq(T v) { return v * v; }  // utils.rs:27-32

// vector.rs:93-95
template <typename T> RT_DEV T dot(V3<T> a, V3<T> b) { return fma(a.z, b.z, fma(a.x, b.x, a.y * b.y)); }
// vector.rs:97-103
template <typename T> RT_DEV V3<T> cross(V3<T> a, V3<T> b) {
    return mk<T>(fma(a.y, b.z, -a.z * b.y), fma(a.z, b.x, -a.x * b.z), fma(a.x, b.y, -a.y * b.x));
}
template <typename T> RT_COLD T div_native(T a, T b) { return a / b; }
template <typename T> RT_COLD T sqrt_native(T a) { return sqrt(a); }

// vector.rs:84-91: magnitude = sqrt(x^2 + y^2 + z^2); normalized DIVIDES each component by it.
// One exact sqrt + one shared reciprocal for the three exact quotients (rt_arith.cuh).
template <typename T>
struct Normalized {
    V3<T> v;
    T magnitude;
};
template <typename T> RT_COLD Normalized<T> normalize_full(V3<T> a) {
    bool ok = true;
    T s = sq(a.x) + sq(a.y) + sq(a.z);
    T m = sqrt_fast(s, ok);
    Recip<T> r = recip(m, ok);
    V3<T> q = mk<T>(quot0(a.x, r, ok), quot0(a.y, r, ok), quot0(a.z, r, ok));
    if (!ok) {
        m = sqrt_native(s);
        q = mk<T>(div_native(a.x, m), div_native(a.y, m), div_native(a.z, m));
    }
    Normalized<T> out;
    out.v = q;
    out.magnitude = m;
    return out;
}
template <typename T> RT_DEV V3<T> normalized(V3<T> a) { return normalize_full(a).v; }
// vector.rs:105-107
template <typename T> RT_DEV V3<T> reflect(V3<T> v, V3<T> n) { return v - ((n * T(2)) * dot(v, n)); }

// matrix.rs:332-346: Matrix<4> * Point, rows 0..2
template <typename T> RT_DEV V3<T> mat_point(const T* m, V3<T> p) {
#if RT_STRICT_SIGNED_ZERO
    return mk<T>((((T(0) + m[0] * p.x) + m[1] * p.y) + m[2] * p.z) + m[3] * T(1),
                 (((T(0) + m[4] * p.x) + m[5] * p.y) + m[6] * p.z) + m[7] * T(1),
                 (((T(0) + m[8] * p.x) + m[9] * p.y) + m[10] * p.z) + m[11] * T(1));
#else
    return mk<T>(((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3], ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7],
                 ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11]);
#endif
}
// matrix.rs:348-362: Matrix<4> * Vector
template <typename T> RT_DEV V3<T> mat_vector(const T* m, V3<T> v) {
#if RT_STRICT_SIGNED_ZERO
    return mk<T>((((T(0) + m[0] * v.x) + m[1] * v.y) + m[2] * v.z) + m[3] * T(0),
                 (((T(0) + m[4] * v.x) + m[5] * v.y) + m[6] * v.z) + m[7] * T(0),
                 (((T(0) + m[8] * v.x) + m[9] * v.y) + m[10] * v.z) + m[11] * T(0));
#else
    return mk<T>((m[0] * v.x + m[1] * v.y) + m[2] * v.z, (m[4] * v.x + m[5] * v.y) + m[6] * v.z,
                 (m[8] * v.x + m[9] * v.y) + m[10] * v.z);
#endif
}
// shapes/shape.rs:25: transformation_inverse.transpose() * local_normal (row 3 of the inverse is 0,0,0,1)
template <typename T> RT_DEV V3<T> mat_transposed_vector(const T* m, V3<T> v) {
#if RT_STRICT_SIGNED_ZERO
    return mk<T>((((T(0) + m[0] * v.x) + m[4] * v.y) + m[8] * v.z) + T(0),
                 (((T(0) + m[1] * v.x) + m[5] * v.y) + m[9] * v.z) + T(0),
                 (((T(0) + m[2] * v.x) + m[6] * v.y) + m[10] * v.z) + T(0));
#else
    return mk<T>((m[0] * v.x + m[4] * v.y) + m[8] * v.z, (m[1] * v.x + m[5] * v.y) + m[9] * v.z,
                 (m[2] * v.x + m[6] * v.y) + m[10] * v.z);
#endif
}

// ---------------------------------------------------------------------------------------------
// The scene as the kernel sees it (shared memory when it fits, else global).
// SMEM = true: both blobs were staged at the start of dynamic shared memory (stage_scene): the accessors name
// that array directly, so every table read is an LDS with a uniform base.  Going through pointer members
// instead cost generic LD.E loads — the view's address escapes to the out-of-line helpers, so ptxas cannot
// tell which window the pointers refer to (measured: -5..14 % frame time on the persistent family).
// SMEM = false: the blobs are read from global memory through L1 / L2.
template <typename T, bool SMEM>
struct SceneView {
    const T* reals;
    const int* ints;
    SceneLayout L;
    RT_DEV const T* R() const {
        if constexpr (SMEM) {
            extern __shared__ __align__(16) unsigned char rt_scene_smem[];
            return reinterpret_cast<const T*>(rt_scene_smem);
        } else {
            return reals;
        }
    }
    RT_DEV const int* I() const {
        if constexpr (SMEM) {
            extern __shared__ __align__(16) unsigned char rt_scene_smem[];
            return reinterpret_cast<const int*>(rt_scene_smem + (((size_t)L.n_reals * sizeof(T) + 15) & ~size_t(15)));
        } else {
            return ints;
        }
    }
    // first byte of dynamic shared memory behind the staged tables (SMEM) or the start of it (tables in global memory)
    RT_DEV unsigned char* scratch() const {
        extern __shared__ __align__(16) unsigned char rt_scene_smem[];
        if constexpr (SMEM) {
            const size_t ints_at = ((size_t)L.n_reals * sizeof(T) + 15) & ~size_t(15);
            return rt_scene_smem + ((ints_at + (size_t)L.n_ints * sizeof(int) + 15) & ~size_t(15));
        } else {
            return rt_scene_smem;
        }
    }
    RT_DEV const T* shape(uint32_t pos) const { RT_CHECK(pos < L.n_shapes); return R() + (size_t)pos * SHAPE_REALS; }
    RT_DEV int4 shape_meta(uint32_t pos) const { RT_CHECK(pos < L.n_shapes); return reinterpret_cast<const int4*>(I())[(size_t)pos * (SHAPE_INTS / 4)]; }
    RT_DEV const T* triangle(uint32_t pos) const { return R() + L.tri_off + (size_t)I()[(size_t)pos * SHAPE_INTS + 4] * TRI_REALS; }
    RT_DEV const T* bvh_boxes(uint32_t node) const { RT_CHECK(node < L.n_bvh_nodes); return R() + L.bvh_off + (size_t)node * BVH_REALS; }
    RT_DEV const float4* bvh_node32(uint32_t node) const { RT_CHECK(node < L.n_bvh_nodes); return reinterpret_cast<const float4*>(I() + L.bvh32_off) + (size_t)node * (BVH32_WORDS / 4); }
    RT_DEV int2 bvh_children(uint32_t node) const { RT_CHECK(node < L.n_bvh_nodes); return reinterpret_cast<const int2*>(I() + L.bvh_meta_off)[node]; }
    RT_DEV const T* material(uint32_t m) const { RT_CHECK(m < L.n_materials); return R() + L.mat_off + (size_t)m * MAT_REALS; }
    RT_DEV int material_pattern(uint32_t m) const { RT_CHECK(m < L.n_materials); return I()[L.mat_meta_off + m * MAT_INTS]; }
    RT_DEV const T* pattern(uint32_t p) const { RT_CHECK(p < L.n_patterns); return R() + L.pat_off + (size_t)p * PAT_REALS; }
    RT_DEV const int* pattern_meta(uint32_t p) const { RT_CHECK(p < L.n_patterns); return I() + L.pat_meta_off + p * PAT_INTS; }
    RT_DEV const T* light(uint32_t l) const { RT_CHECK(l < L.n_lights); return R() + L.light_off + (size_t)l * LIGHT_REALS; }
    RT_DEV const T* cull(uint32_t pos) const { RT_CHECK(pos <= L.n_shapes); return R() + L.cull_off + (size_t)pos * CULL_REALS; }
    RT_DEV const float4* cull32(uint32_t pos) const { RT_CHECK(pos <= L.n_shapes); return reinterpret_cast<const float4*>(I() + L.cull32_off) + pos; }
    RT_DEV T cull_shrink() const { return sizeof(T) == 8 ? (T)L.cull_shrink64 : (T)L.cull_shrink32; }
};

// Copy both scene blobs to the start of dynamic shared memory with the TMA unit: two bulk asynchronous copies
// (cp.async.bulk, SASS UBLKCP) issued by one thread, completion counted in bytes on an mbarrier that every
// thread then waits on.  The packer pads both blobs to multiples of 16 bytes; cudaMalloc and the shared window
// are aligned.  Called by every thread of the CTA, once, before anything else touches shared memory.
template <typename T>
RT_DEV void stage_scene(const SceneLayout& layout, const T* __restrict__ g_reals, const int* __restrict__ g_ints) {
    extern __shared__ __align__(16) unsigned char rt_scene_smem[];
    T* s_reals = reinterpret_cast<T*>(rt_scene_smem);
    int* s_ints = reinterpret_cast<int*>(rt_scene_smem + (((size_t)layout.n_reals * sizeof(T) + 15) & ~size_t(15)));
#if RT_TMA_STAGE
    __shared__ __align__(8) unsigned long long stage_bar;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&stage_bar);
    const uint32_t bytes_reals = layout.n_reals * (uint32_t)sizeof(T), bytes_ints = layout.n_ints * (uint32_t)sizeof(int);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_reals + bytes_ints) : "memory");
        if (bytes_reals)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s_reals)),
                         "l"(g_reals), "r"(bytes_reals), "r"(bar)
                         : "memory");
        if (bytes_ints)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (uint32_t)__cvta_generic_to_shared(s_ints)),
                         "l"(g_ints), "r"(bytes_ints), "r"(bar)
                         : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }
#else
    for (uint32_t i = threadIdx.x; i < layout.n_reals; i += blockDim.x) s_reals[i] = g_reals[i];
    for (uint32_t i = threadIdx.x; i < layout.n_ints; i += blockDim.x) s_ints[i] = g_ints[i];
    __syncthreads();
#endif
}

// ---------------------------------------------------------------------------------------------
// What one trace accumulates.  Three query kinds share the intersection loop:
//   RADIANCE  : Intersections::hit (intersections.rs:13-18) = the first minimal distance >= 0 of the
//               stable-sorted list (world.rs:34), i.e. the minimum of (distance, world order).
//   SHADOW    : World::is_in_shadow (world.rs:98-112) = any casts_shadow shape with 0 <= t < light
//               distance — a nearest-hit query seeded with best_t = light distance.
//   CONTAINER : the refraction-container walk of Intersection::prepare_computations
//               (intersection.rs:33-62) without materialising the sorted list; see ContainerAcc.
enum : int { MODE_RADIANCE = 0, MODE_SHADOW = 1, MODE_CONTAINER = 2, MODE_IDLE = 3 };

// The walk toggles each shape class at every intersection that sorts before the hit, i.e. with
// distance < hit distance (the hit is the FIRST entry with its distance).  A class is in the
// container list iff it has an odd number of such intersections; its place in the list is the
// sort position of its last one, (max distance, world order).  n1 = refractive index of the last
// class in the list; n2 = the same after toggling the hit's own class (intersection.rs:39-58).
template <typename T>
struct ContainerAcc {
    T t_hit;
    int hit_class;
    bool hit_class_inside;  // the hit's class is in the list before the hit toggles it
    int all_pos, excl_pos;  // sorted position of the last class overall / excluding the hit's class; -1 = none
    T all_t, excl_t;
    int all_orig, excl_orig;
};

template <typename T>
struct TraceAcc {
    int mode;
    T dir_sq;       // |direction|^2 of the ray being traced (refracted rays are not unit, world.rs:150)
    T best_t;       // RADIANCE: +max seed; SHADOW: light distance seed
    int best_orig;  // world order of the best hit (tie-break), INT_MAX seed
    int best_pos;   // sorted position of the best hit, -1 = none
    ContainerAcc<T>* c;  // a separate object: its address escapes to the out-of-line consume_container, and the
                         // hot fields above must stay promotable to registers
#if !RT_ACC_SPLIT
    ContainerAcc<T> c_store;  // A/B: c points here, the whole accumulator escapes and lives in local memory
#endif
};

template <typename T>
RT_DEV ContainerAcc<T>* acc_store(TraceAcc<T>& a) {
#if RT_ACC_SPLIT
    return nullptr;
#else
    return &a.c_store;
#endif
}

// The container query's bookkeeping for one shape (rare: only hits on transparent materials ask for
// it), kept out of line so the three hot query loops stay small.
template <typename T>
RT_COLD void consume_container(ContainerAcc<T>& c, int n, T t0, T t1, T t2, T t3, int pos, int4 meta) {
    const T ts[4] = {t0, t1, t2, t3};
    int count = 0;
    T tmax = -Real<T>::max();
    bool any = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < n && ts[k] < c.t_hit) {
            ++count;
            if (!any || ts[k] > tmax) tmax = ts[k];
            any = true;
        }
    }
    if (count & 1) {
        if (meta.w == c.hit_class) {
            c.hit_class_inside = true;
        } else if (c.excl_pos < 0 || tmax > c.excl_t || (tmax == c.excl_t && meta.x > c.excl_orig)) {
            c.excl_pos = pos;
            c.excl_t = tmax;
            c.excl_orig = meta.x;
        }
        if (c.all_pos < 0 || tmax > c.all_t || (tmax == c.all_t && meta.x > c.all_orig)) {
            c.all_pos = pos;
            c.all_t = tmax;
            c.all_orig = meta.x;
        }
    }
}

// NMAX = the most intersections the shape type can push (sphere 2, plane 1, cube 2, cylinder / cone 4, triangle 1)
template <typename T, int NMAX>
RT_DEV void consume(TraceAcc<T>& a, int n, T t0, T t1, T t2, T t3, int pos, int4 meta) {
    if (a.mode != MODE_CONTAINER) {
        // intersections.rs:13-18 / intersection.rs:77-79.  NaN fails `t >= 0`.
        bool eligible = (a.mode == MODE_RADIANCE) || (meta.z & FLAG_CASTS_SHADOW);
        if (eligible) {
            const T ts[4] = {t0, t1, t2, t3};
#pragma unroll
            for (int k = 0; k < NMAX; ++k) {
                if (k < n) {
                    T t = ts[k];
                    if (t >= T(0) && (t < a.best_t || (t == a.best_t && meta.x < a.best_orig))) {
                        a.best_t = t;
                        a.best_orig = meta.x;
                        a.best_pos = pos;
                    }
                }
            }
        }
    } else if (meta.z & FLAG_CONTAINER_REP) {
        consume_container(*a.c, n, t0, t1, t2, t3, pos, meta);
    }
}

// utils.rs:47-57
template <typename T>
RT_DEV bool solve_quadratic(T a, T b, T c, T& s1, T& s2) {
    T discriminant = fma(T(4) * a, -c, sq(b));
    if (discriminant < T(0)) return false;
    T double_a = T(2) * a;
    bool ok = true;
    T root = sqrt_fast(discriminant, ok);
    Recip<T> r = recip(double_a, ok);
    s1 = quot(-b - root, r, ok);
    s2 = quot(-b + root, r, ok);
    if (!ok) {
        root = sqrt_native(discriminant);
        s1 = div_native(-b - root, double_a);
        s2 = div_native(-b + root, double_a);
    }
    return true;
}

// shapes/cube.rs:22-43, native operators (fallback of cube_axis_fast)
template <typename T>
RT_DEV void cube_check_axis_native(T origin, T direction, T& tmin, T& tmax) {
    T nmin = T(-1) - origin;
    T nmax = T(1) - origin;
    T dmin, dmax;
    if (fabs(direction) >= Real<T>::eps()) {
        dmin = nmin / direction;
        dmax = nmax / direction;
    } else {
        dmin = nmin * Real<T>::max();
        dmax = nmax * Real<T>::max();
    }
    if (dmin > dmax) {
        T t = dmin;
        dmin = dmax;
        dmax = t;
    }
    tmin = dmin;
    tmax = dmax;
}

// Everything by value: a reference parameter of an out-of-line function pins the caller's object in local memory
// (the object-space ray was stored there before every cube test although this fallback almost never runs).
template <typename T>
struct CubeSlabs {
    T xmin, xmax, ymin, ymax, zmin, zmax;
};

template <typename T>
RT_COLD CubeSlabs<T> cube_axes_native(V3<T> o, V3<T> d) {
    CubeSlabs<T> s;
    cube_check_axis_native(o.x, d.x, s.xmin, s.xmax);
    cube_check_axis_native(o.y, d.y, s.ymin, s.ymax);
    cube_check_axis_native(o.z, d.z, s.zmin, s.zmax);
    return s;
}

// shapes/cube.rs:22-43 without branches: both slab distances share one reciprocal; `ok` turns false
// when an operand leaves the exact fast range AND the quotients are actually used.
template <typename T>
RT_DEV void cube_axis_fast(T origin, T direction, T& tmin, T& tmax, bool& ok) {
    T nmin = T(-1) - origin;
    T nmax = T(1) - origin;
    const bool divide = fabs(direction) >= Real<T>::eps();
    bool ok_div = true;
    Recip<T> r = recip(direction, ok_div);
    T qmin = quot(nmin, r, ok_div);
    T qmax = quot(nmax, r, ok_div);
    T dmin = divide ? qmin : nmin * Real<T>::max();
    T dmax = divide ? qmax : nmax * Real<T>::max();
    const bool swap = dmin > dmax;
    tmin = swap ? dmax : dmin;
    tmax = swap ? dmin : dmax;
    ok = ok && (ok_div || !divide);
}

// cylinder.rs:34-39 / cone.rs:34-39 (radius 1 for the cylinder)
template <typename T>
RT_DEV bool check_cap(const Ray<T>& r, T distance, T radius_sq) {
    T x = fma(r.d.x, distance, r.o.x);
    T z = fma(r.d.z, distance, r.o.z);
    return (sq(x) + sq(z)) <= radius_sq;
}

// (min - o.y) / d.y and (max - o.y) / d.y of intersect_caps (cylinder.rs:47,52 / cone.rs:47,52)
template <typename T>
RT_DEV void cap_distances(T nlo, T nhi, T dy, T& dlo, T& dhi) {
    bool ok = true;
    Recip<T> r = recip(dy, ok);
    dlo = quot(nlo, r, ok);
    dhi = quot(nhi, r, ok);
    if (!ok) {
        dlo = div_native(nlo, dy);
        dhi = div_native(nhi, dy);
    }
}

// local_intersect of shape type TYPE on the object-space ray `r`; distances in push order.
template <typename T, int TYPE>
RT_DEV int local_intersect(const Ray<T>& r, const T* g, int flags, const T* tri, T& t0, T& t1, T& t2, T& t3) {
    int n = 0;
    T ts[4] = {T(0), T(0), T(0), T(0)};
    if (TYPE == 0 && sizeof(T) == 4) {
        // f32 fast mode: b^2 - 4ac cancels catastrophically in binary32 when the origin is far away in object units
        // (small or squashed spheres: shadow_puppets' 200 x 200 x 0.01 backdrop).  The discriminant is taken from the
        // ray's closest approach to the centre instead (Haines et al., "Precision improvements for ray / sphere
        // intersection"), and the roots from the stable pair q / a, c / q.  Not the reference's operations: this
        // mode is bound by its stated tolerance, not bit parity.
        const T a = dot(r.d, r.d);
        const T bh = -dot(r.o, r.d);                   // half of -b
        const T k = bh / a;
        const V3<T> l = r.o + r.d * k;                 // closest point of the line to the centre
        const T discr = T(1) - dot(l, l);              // x a = (b^2 - 4ac) / 4
        if (!(discr < T(0))) {
            const T c = dot(r.o, r.o) - T(1);
            const T root = sqrt(a * discr);
            const T q = bh + (bh < T(0) ? -root : root);
            T s1 = q != T(0) ? c / q : T(0), s2 = q / a;
            if (s1 > s2) {
                const T tmp = s1;
                s1 = s2;
                s2 = tmp;
            }
            ts[0] = s1;
            ts[1] = s2;
            n = 2;
        }
    } else if (TYPE == 0) {  // shapes/sphere.rs:41-53
        T a = dot(r.d, r.d);
        T b = T(2) * dot(r.d, r.o);
        T c = dot(r.o, r.o) - T(1);
        T s1, s2;
        if (solve_quadratic(a, b, c, s1, s2)) {
            ts[0] = s1;
            ts[1] = s2;
            n = 2;
        }
    } else if (TYPE == 1) {  // shapes/plane.rs:42-48
        if (!(fabs(r.d.y) < Real<T>::eps())) {
            ts[0] = div_exact(-r.o.y, r.d.y);
            n = 1;
        }
    } else if (TYPE == 2) {  // shapes/cube.rs:65-85
        T xmin, xmax, ymin, ymax, zmin, zmax;
        bool ok = true;
        cube_axis_fast(r.o.x, r.d.x, xmin, xmax, ok);
        cube_axis_fast(r.o.y, r.d.y, ymin, ymax, ok);
        cube_axis_fast(r.o.z, r.d.z, zmin, zmax, ok);
        if (!ok) {
            const CubeSlabs<T> s = cube_axes_native(r.o, r.d);
            xmin = s.xmin; xmax = s.xmax;
            ymin = s.ymin; ymax = s.ymax;
            zmin = s.zmin; zmax = s.zmax;
        }
        T dmin = fmax(fmax(fmax(-Real<T>::max(), xmin), ymin), zmin);
        T dmax = fmin(fmin(fmin(Real<T>::max(), xmax), ymax), zmax);
        if (dmin < dmax && dmax > T(0)) {
            ts[0] = dmin;
            ts[1] = dmax;
            n = 2;
        }
    } else if (TYPE == 3) {  // shapes/cylinder.rs:81-110 + 41-59
        T mn = g[SHAPE_MIN], mx = g[SHAPE_MAX];
        T a = sq(r.d.x) + sq(r.d.z);
        if (fabs(a) > T(0)) {
            T b = T(2) * fma(r.o.x, r.d.x, r.o.z * r.d.z);
            T c = sq(r.o.x) + sq(r.o.z) - T(1);
            T d1, d2;
            bool roots;
            if (sizeof(T) == 4) {
                // f32 fast mode: the same closest-approach discriminant as the sphere's, in the xz plane
                const T bh = -(r.o.x * r.d.x + r.o.z * r.d.z), k = bh / a;
                const T lx = fma(r.d.x, k, r.o.x), lz = fma(r.d.z, k, r.o.z);
                const T discr = T(1) - (lx * lx + lz * lz);
                roots = !(discr < T(0));
                if (roots) {
                    const T root = sqrt(a * discr);
                    const T q = bh + (bh < T(0) ? -root : root);
                    d1 = q != T(0) ? c / q : T(0);
                    d2 = q / a;
                }
            } else {
                roots = solve_quadratic(a, b, c, d1, d2);
            }
            if (roots) {
                if (d1 > d2) {
                    T t = d1;
                    d1 = d2;
                    d2 = t;
                }
                T y1 = fma(d1, r.d.y, r.o.y);
                if (mn < y1 && y1 < mx) ts[n++] = d1;
                T y2 = fma(d2, r.d.y, r.o.y);
                if (mn < y2 && y2 < mx) ts[n++] = d2;
            }
        }
        if ((flags & FLAG_CLOSED) && !(fabs(r.d.y) < Real<T>::eps())) {
            T dlo, dhi;
            cap_distances(mn - r.o.y, mx - r.o.y, r.d.y, dlo, dhi);
            if (check_cap(r, dlo, T(1))) ts[n++] = dlo;
            if (check_cap(r, dhi, T(1))) ts[n++] = dhi;
        }
    } else if (TYPE == 4) {  // shapes/cone.rs:81-112 + 41-59
        T mn = g[SHAPE_MIN], mx = g[SHAPE_MAX];
        T a = sq(r.d.x) - sq(r.d.y) + sq(r.d.z);
        T b = T(2) * fma(r.o.z, r.d.z, fma(r.o.x, r.d.x, -r.o.y * r.d.y));
        T c = sq(r.o.x) - sq(r.o.y) + sq(r.o.z);
        T d1, d2;
        if (fabs(a) < Real<T>::eps() && fabs(b) > Real<T>::eps()) {
            ts[n++] = div_exact(-c, T(2) * b);
        } else if (solve_quadratic(a, b, c, d1, d2)) {
            if (d1 > d2) {
                T t = d1;
                d1 = d2;
                d2 = t;
            }
            T y1 = fma(r.d.y, d1, r.o.y);
            if (mn < y1 && y1 < mx) ts[n++] = d1;
            T y2 = fma(r.d.y, d2, r.o.y);
            if (mn < y2 && y2 < mx) ts[n++] = d2;
        }
        if ((flags & FLAG_CLOSED) && !(fabs(r.d.y) < Real<T>::eps())) {
            T dlo, dhi;
            cap_distances(mn - r.o.y, mx - r.o.y, r.d.y, dlo, dhi);
            if (check_cap(r, dlo, sq(mn))) ts[n++] = dlo;
            if (check_cap(r, dhi, sq(mx))) ts[n++] = dhi;
        }
    } else {  // shapes/triangle.rs:39-56
        V3<T> v1 = ld3(tri), e1 = ld3(tri + 3), e2 = ld3(tri + 6);
        V3<T> dce2 = cross(r.d, e2);
        T det = dot(e1, dce2);
        if (!(fabs(det) < Real<T>::eps())) {
            V3<T> v1o = r.o - v1;
            // u, v and the distance are three exact quotients by the same determinant
            bool ok_r = true;
            Recip<T> rd = recip(det, ok_r);
            T un = dot(v1o, dce2);
            bool ok_u = ok_r;
            T u = quot(un, rd, ok_u);
            if (!ok_u) u = div_native(un, det);
            if (u >= T(0) && u <= T(1)) {
                V3<T> oce1 = cross(v1o, e1);
                T vn = dot(r.d, oce1);
                bool ok_v = ok_r;
                T v = quot(vn, rd, ok_v);
                if (!ok_v) v = div_native(vn, det);
                if (v > T(0) && u + v < T(1)) {
                    T tn = dot(e2, oce1);
                    bool ok_t = ok_r;
                    T t = quot(tn, rd, ok_t);
                    if (!ok_t) t = div_native(tn, det);
                    ts[0] = t;
                    n = 1;
                }
            }
        }
    }
    t0 = ts[0];
    t1 = ts[1];
    t2 = ts[2];
    t3 = ts[3];
    return n;
}

// A cull record (centre[3], radius^2) is 16-byte aligned in both precisions: two 128-bit loads (one in f32)
// instead of four scalar ones.
RT_DEV void load_cull(const double* p, double& x, double& y, double& z, double& w) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    x = a.x; y = a.y; z = b.x; w = b.y;
}
RT_DEV void load_cull(const float* p, float& x, float& y, float& z, float& w) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    x = a.x; y = a.y; z = a.z; w = a.w;
}

// Ray::intersect (ray.rs:35-49) for the shape at sorted position `pos`, of (compile-time) type TYPE:
// object-space ray, local_intersect, then the query's bookkeeping.
template <typename T, int TYPE, bool SMEM>
RT_DEV void test_shape(const SceneView<T, SMEM>& sv, uint32_t pos, const Ray<T>& ray, TraceAcc<T>& acc) {
    const T* g = sv.shape(pos);
    int4 meta = sv.shape_meta(pos);
    Ray<T> local;
    local.o = mat_point(g, ray.o);
    local.d = mat_vector(g, ray.d);
    T t0, t1, t2, t3;
    int n = local_intersect<T, TYPE>(local, g, meta.z, TYPE == 5 ? sv.triangle(pos) : nullptr, t0, t1, t2, t3);
    constexpr int NMAX = (TYPE == 0 || TYPE == 2) ? 2 : (TYPE == 3 || TYPE == 4) ? 4 : 1;
    consume<T, NMAX>(acc, n, t0, t1, t2, t3, (int)pos, meta);
}

// ---- BVH traversal (scenes with many bounded shapes; rt_bvh.h) -----------------------------------
// Each lane walks the hierarchy with its own small stack.  A child is entered iff the ray's parameter
// interval inside its (inflated) box meets the interval the query still cares about:
//   RADIANCE  [0, best_t]   (shrinks as hits are found; `<=` keeps equal-distance candidates, which the
//                            world-order tie-break may still prefer)
//   SHADOW    [0, best_t]   with best_t = light distance; the walk stops at the first blocker
//   CONTAINER (-inf, t_hit] (every intersection before the hit, negative distances included)
// Exact tests of the leaves are batched by shape type across the warp: the lanes that currently hold
// a leaf of the elected type run that type's test together, so a warp never executes two shape types'
// code at once.
template <typename T>
RT_DEV bool box_hit(const T* b, const Ray<T>& ray, V3<T> inv, T t_lo, T t_hi, T& enter) {
    // slab test; fmin/fmax drop the NaN of 0 * inf (ray parallel to a slab and exactly on its face)
    T x1 = (b[0] - ray.o.x) * inv.x, x2 = (b[3] - ray.o.x) * inv.x;
    T y1 = (b[1] - ray.o.y) * inv.y, y2 = (b[4] - ray.o.y) * inv.y;
    T z1 = (b[2] - ray.o.z) * inv.z, z2 = (b[5] - ray.o.z) * inv.z;
    T tn = fmax(fmax(fmin(x1, x2), fmin(y1, y2)), fmax(fmin(z1, z2), t_lo));
    T tf = fmin(fmin(fmax(x1, x2), fmax(y1, y2)), fmin(fmax(z1, z2), t_hi));
    enter = tn;
    return tn <= tf;
}

#ifndef RT_BVH_F32
#define RT_BVH_F32 1  // box tests in single precision with rigorous margins (0: in T, as in round 1)
#endif

// The ray as the single-precision box test sees it.  Every quantity errs on the side of a LARGER parameter interval:
//   o_plus / o_minus  the origin moved by delta = 2^-21 * max(|o|, largest box coordinate) towards +inf / -inf: more
//                     than the rounding of the origin to float plus that of the subtraction below, so
//                     lo_f - o_plus <= lo - o and hi_f - o_minus >= hi - o hold in exact arithmetic (lo_f <= lo and
//                     hi_f >= hi: the packer rounds the boxes outwards);
//   inv               1 / d rounded to float (relative error 2^-24; +-inf for a zero component);
// products and the conversion of the query's distance bounds add relative errors below 2^-22, which the final
// comparison absorbs with a relative slack of 2^-20.
struct BoxRay32 {
    float opx, opy, opz, omx, omy, omz, ix, iy, iz;
};
template <typename T>
RT_DEV BoxRay32 make_box_ray(const Ray<T>& ray, float coord_max) {
    BoxRay32 b;
    const float ox = (float)ray.o.x, oy = (float)ray.o.y, oz = (float)ray.o.z;
    const float delta = 4.76837158e-7f * fmaxf(fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz)), coord_max);
    b.opx = ox + delta; b.opy = oy + delta; b.opz = oz + delta;
    b.omx = ox - delta; b.omy = oy - delta; b.omz = oz - delta;
    b.ix = (float)(T(1) / ray.d.x); b.iy = (float)(T(1) / ray.d.y); b.iz = (float)(T(1) / ray.d.z);
    return b;
}
// lo / hi: one child's box (outward-rounded floats); [t_lo, t_hi]: the distances the query still cares about, already
// rounded outwards.  fminf / fmaxf drop the NaN of 0 * inf (origin on a widened plane of a slab the ray is parallel
// to: outside the true box).
RT_DEV bool box_hit32(float lox, float loy, float loz, float hix, float hiy, float hiz, const BoxRay32& r, float t_lo, float t_hi, float& enter) {
    const float x1 = (lox - r.opx) * r.ix, x2 = (hix - r.omx) * r.ix;
    const float y1 = (loy - r.opy) * r.iy, y2 = (hiy - r.omy) * r.iy;
    const float z1 = (loz - r.opz) * r.iz, z2 = (hiz - r.omz) * r.iz;
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fmaxf(fminf(z1, z2), t_lo));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fminf(fmaxf(z1, z2), t_hi));
    enter = tn;
    return tn <= 3.4028235e38f && tn - tf <= 9.5367432e-7f * (fabsf(tn) + fabsf(tf));
}
RT_DEV float float_up(double v) { return __double2float_ru(v); }
RT_DEV float float_up(float v) { return v; }

template <typename T, bool FULL, bool SMEM>
RT_DEV void trace_bvh(const SceneView<T, SMEM>& sv, const Ray<T>& ray, TraceAcc<T>& acc) {
    const bool active = acc.mode != MODE_IDLE;
#if RT_BVH_F32
    const BoxRay32 bray = make_box_ray(ray, sv.L.bvh_coord_max);
    const float t_lo32 = (acc.mode == MODE_CONTAINER) ? -3.4028235e38f : 0.0f;
#else
    // 1 / d per axis; only used for the conservative box tests, so plain reciprocals are enough.
    // A zero component gives +-inf, which the slab test handles.
    V3<T> inv = mk<T>(T(1) / ray.d.x, T(1) / ray.d.y, T(1) / ray.d.z);
    const T t_lo = (acc.mode == MODE_CONTAINER) ? -Real<T>::max() : T(0);
#endif
    int stack[BVH_MAX_DEPTH + 4];
    int sp = 0;
    int cur = active ? 0 : INT_MIN;  // INT_MIN = nothing left to visit; >= 0 inner node; < 0 leaf ~pos
    int pending = -1;                // leaf waiting for its exact test
#if RT_BVH_WATCHDOG
    unsigned long long wd_outer = 0, wd_inner = 0;
#endif
    for (;;) {
#if RT_BVH_WATCHDOG
        if (++wd_outer > 2000000ull) {
            printf("bvh watchdog outer: blk %d thr %d mode %d cur %d pending %d sp %d best_t %g inner %llu\n", blockIdx.x, threadIdx.x, acc.mode, cur, pending, sp, (double)acc.best_t, wd_inner);
            break;
        }
#endif
        while (pending < 0 && cur != INT_MIN) {
#if RT_BVH_WATCHDOG
            if (++wd_inner > 4000000ull) {
                printf("bvh watchdog inner: blk %d thr %d mode %d cur %d sp %d\n", blockIdx.x, threadIdx.x, acc.mode, cur, sp);
                cur = INT_MIN;
                break;
            }
#endif
            if (cur < 0) {
                pending = ~cur;
                cur = sp > 0 ? stack[--sp] : INT_MIN;
                break;
            }
#if RT_BVH_F32
            const float4* nd = sv.bvh_node32((uint32_t)cur);
            const float4 n0 = nd[0], n1 = nd[1], n2 = nd[2];
            const int2 ch = make_int2(__float_as_int(nd[3].x), __float_as_int(nd[3].y));
            const float t_hi32 = float_up((acc.mode == MODE_CONTAINER) ? acc.c->t_hit : acc.best_t);
            float e0, e1;
            const bool h0 = box_hit32(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, bray, t_lo32, t_hi32, e0);
            const bool h1 = box_hit32(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, bray, t_lo32, t_hi32, e1);
#else
            const T* nb = sv.bvh_boxes((uint32_t)cur);
            const int2 ch = sv.bvh_children((uint32_t)cur);
            const T t_hi = (acc.mode == MODE_CONTAINER) ? acc.c->t_hit : acc.best_t;
            T e0, e1;
            const bool h0 = box_hit(nb, ray, inv, t_lo, t_hi, e0);
            const bool h1 = box_hit(nb + 6, ray, inv, t_lo, t_hi, e1);
#endif
            if (h0 && h1) {
                const bool near0 = e0 <= e1;
                stack[sp++] = near0 ? ch.y : ch.x;
                cur = near0 ? ch.x : ch.y;
            } else if (h0 || h1) {
                cur = h0 ? ch.x : ch.y;
            } else {
                cur = sp > 0 ? stack[--sp] : INT_MIN;
            }
        }
        // exact tests, one shape type at a time across the lanes that walk together.  `group` is whoever is
        // converged here — after the divergent loop above that is normally the whole warp.  Voting on the full
        // mask instead is legal CUDA but deadlocked in one build of the wavefront kernel (10^4 shapes at 1080p,
        // level-0 launch never finished; adding a printf to the loop made it go away): with the group's own mask
        // no lane ever waits for a lane outside its convergence group, whatever ptxas does with the barriers.
        const unsigned group = __activemask();
        unsigned waiting = __ballot_sync(group, pending >= 0);
        if (waiting == 0u) {
            if (__all_sync(group, cur == INT_MIN)) break;
            continue;
        }
        const int my_type = pending >= 0 ? ((sv.shape_meta((uint32_t)pending).z >> FLAG_TYPE_SHIFT) & 7) : -1;
        while (waiting) {
            const int leader = __ffs(waiting) - 1;
            const int type = __shfl_sync(group, my_type, leader);
            if (pending >= 0 && my_type == type) {
                switch (type) {
                case 0: test_shape<T, 0>(sv, (uint32_t)pending, ray, acc); break;
                case 2: test_shape<T, 2>(sv, (uint32_t)pending, ray, acc); break;
                case 3: if (FULL) test_shape<T, 3>(sv, (uint32_t)pending, ray, acc); break;
                case 4: if (FULL) test_shape<T, 4>(sv, (uint32_t)pending, ray, acc); break;
                case 5: if (FULL) test_shape<T, 5>(sv, (uint32_t)pending, ray, acc); break;
                default: break;  // planes are never bounded
                }
                pending = -1;
                // World::is_in_shadow (world.rs:106-111) is an `any`: the first blocker ends the walk
                if (acc.mode == MODE_SHADOW && acc.best_pos >= 0) {
                    cur = INT_MIN;
                    sp = 0;
                }
            }
            waiting = __ballot_sync(group, pending >= 0);
        }
    }
}

// One pixel of the f64 Canvas is 24 bytes at a 24-byte stride: 16-byte aligned for even pixels, 8 mod 16 for odd ones.
// One 128-bit and one 64-bit store instead of three 64-bit ones (f32: 12 bytes, 4-byte aligned: scalar stores).
RT_DEV void store_rgb(double* p, V3<double> c) {
    if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        *reinterpret_cast<double2*>(p) = make_double2(c.x, c.y);
        p[2] = c.z;
    } else {
        p[0] = c.x;
        *reinterpret_cast<double2*>(p + 1) = make_double2(c.y, c.z);
    }
}
RT_DEV void store_rgb(float* p, V3<float> c) {
    p[0] = c.x;
    p[1] = c.y;
    p[2] = c.z;
}

// ptxas prefers recomputing a cheap loop invariant in every iteration to holding it in a register; an empty asm
// makes the value opaque, so it is computed once.
RT_DEV void keep_in_register(double& v) { asm volatile("" : "+d"(v)); }
RT_DEV void keep_in_register(float& v) { asm volatile("" : "+f"(v)); }

// The shape loop's only loop-carried state: where the next cull record is.
template <typename T, bool SMEM>
struct CullCursor;
template <typename T>
struct CullCursor<T, true> {  // tables in shared memory: shared-window byte addresses
    uint32_t at, end;
    template <typename SV>
    RT_DEV CullCursor(const SV& sv, uint32_t first, uint32_t n) {
        at = (uint32_t)__cvta_generic_to_shared(sv.cull(first));
        end = at + n * (uint32_t)(CULL_REALS * sizeof(T));
        asm volatile("" : "+r"(end));  // opaque: a register, not eight uniform instructions per iteration to rebuild it
    }
    RT_DEV bool done() const { return at == end; }
    RT_DEV void next() { at += (uint32_t)(CULL_REALS * sizeof(T)); }
    RT_DEV void finish() { at = end - (uint32_t)(CULL_REALS * sizeof(T)); }  // the loop's increment then ends it
    RT_DEV uint32_t position(uint32_t n) const { return n - (end - at) / (uint32_t)(CULL_REALS * sizeof(T)); }
    RT_DEV void load(double& x, double& y, double& z, double& w) const {
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(at));
        asm("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(z), "=d"(w) : "r"(at));
    }
    RT_DEV void load(float& x, float& y, float& z, float& w) const {
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(at));
    }
};
template <typename T>
struct CullCursor<T, false> {  // tables in global memory
    const T *at, *end;
    template <typename SV>
    RT_DEV CullCursor(const SV& sv, uint32_t first, uint32_t n) {
        at = sv.cull(first);
        end = at + (size_t)n * CULL_REALS;
    }
    RT_DEV bool done() const { return at == end; }
    RT_DEV void next() { at += CULL_REALS; }
    RT_DEV void finish() { at = end - CULL_REALS; }
    RT_DEV uint32_t position(uint32_t n) const { return n - (uint32_t)((end - at) / CULL_REALS); }
    RT_DEV void load(T& x, T& y, T& z, T& w) const { load_cull(at, x, y, z, w); }
};

// The same cursor over the single-precision cull records (16 bytes each, int blob).
template <bool SMEM>
struct Cull32Cursor;
template <>
struct Cull32Cursor<true> {
    uint32_t at, end;
    template <typename SV>
    RT_DEV Cull32Cursor(const SV& sv, uint32_t first, uint32_t n) {
        at = (uint32_t)__cvta_generic_to_shared(sv.cull32(first));
        end = at + n * 16u;
        asm volatile("" : "+r"(end));
    }
    RT_DEV bool done() const { return at == end; }
    RT_DEV void next() { at += 16u; }
    RT_DEV void finish() { at = end - 16u; }
    RT_DEV uint32_t position(uint32_t n) const { return n - (end - at) / 16u; }
    RT_DEV void load(float& x, float& y, float& z, float& w) const {
        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(at));
    }
};
template <>
struct Cull32Cursor<false> {
    const float4 *at, *end;
    template <typename SV>
    RT_DEV Cull32Cursor(const SV& sv, uint32_t first, uint32_t n) {
        at = sv.cull32(first);
        end = at + n;
    }
    RT_DEV bool done() const { return at == end; }
    RT_DEV void next() { ++at; }
    RT_DEV void finish() { at = end - 1; }
    RT_DEV uint32_t position(uint32_t n) const { return n - (uint32_t)(end - at); }
    RT_DEV void load(float& x, float& y, float& z, float& w) const {
        const float4 a = *at;
        x = a.x; y = a.y; z = a.z; w = a.w;
    }
};

// World::collect_intersections (world.rs:25-35) over the uniform list.  One loop with a warp-uniform switch on
// the shape type: the pre-test, the object-space transform and the query bookkeeping exist ONCE in the instruction
// stream instead of once per shape type.  The kernels are bound by instruction supply (GPC instruction cache
// request rate), so code bytes on the hot path matter more than the handful of extra instructions per shape; the
// earlier one-loop-per-type version was 7 % slower.
//
// Conservative pre-test against the shape's world-space bounding sphere (centre c, radius^2 r2, inflated by the
// packer).  With oc = c - o and bq = oc.d, the ray's supporting line misses the sphere iff
// |oc|^2 |d|^2 - bq^2 > r2 |d|^2; if the origin is outside and the centre behind it (bq < 0) every intersection has
// a negative distance, which only the container walk cares about.  The shrink factor dwarfs the rounding of this
// test, and the inflation dwarfs the rounding of the exact test, so a culled shape yields no intersection in the
// reference's arithmetic either.  Shapes without a finite sphere carry r2 = +inf and always pass.
// SHADOW_EXIT: a lane whose shadow query has found a blocker leaves the loop.  Pays when whole warps are in
// the same query (wavefront family: -8 % on cover); in the persistent kernel, where the lanes of a warp are in
// different queries, the extra branch costs more than the idle lanes save.
// The exact test of one shape of the uniform list (Ray::intersect, ray.rs:35-49 + the query's bookkeeping).
template <typename T, bool FULL, bool SMEM>
RT_DEV void exact_test(const SceneView<T, SMEM>& sv, uint32_t pos, const Ray<T>& ray, TraceAcc<T>& acc) {
    const T* g = sv.shape(pos);
    const int4 meta = sv.shape_meta(pos);
    Ray<T> local;  // ray.rs:45-49
    local.o = mat_point(g, ray.o);
    local.d = mat_vector(g, ray.d);
    T t0, t1, t2, t3;  // set by every local_intersect
    int k = 0;
    switch ((meta.z >> FLAG_TYPE_SHIFT) & 7) {
    case 0: k = local_intersect<T, 0>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
    case 1: k = local_intersect<T, 1>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
    case 2: k = local_intersect<T, 2>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
    // !FULL kernels are only launched for scenes without cylinders, cones and triangles: nothing to consume.  (Giving the
    // distances dummy zeros instead put two register-pair clears into every iteration of the caller's shape loop.)
    case 3: if (FULL) k = local_intersect<T, 3>(local, g, meta.z, nullptr, t0, t1, t2, t3); else return; break;
    case 4: if (FULL) k = local_intersect<T, 4>(local, g, meta.z, nullptr, t0, t1, t2, t3); else return; break;
    default: if (FULL) k = local_intersect<T, 5>(local, g, meta.z, sv.triangle(pos), t0, t1, t2, t3); else return; break;
    }
    consume<T, 4>(acc, k, t0, t1, t2, t3, (int)pos, meta);
}

template <typename T, bool FULL, bool SHADOW_EXIT, bool SMEM>
RT_DEV void trace_unified(const SceneView<T, SMEM>& sv, const Ray<T>& ray, TraceAcc<T>& acc) {
    const uint32_t n = sv.L.type_begin[NUM_SHAPE_TYPES];
#if RT_CULL && RT_CANDIDATES
    // Two passes per block of 32 shapes.  (1) Every lane pre-tests ITS ray against the block, branch-free, and keeps
    // the survivors as a bit mask.  (2) Each lane runs the exact tests of its own survivors, first survivor first:
    // in round j every lane is on its j-th candidate, so the warp needs max-over-lanes(#survivors) rounds — 5-6 on
    // the cover frame — where the single loop needed one round per shape that ANY lane's ray survives (12-15 at the
    // deeper levels, each for a handful of lanes: local_intersect ran at 8 of 32 threads per instruction).  The
    // price is per-lane table addresses in the exact test.
    const T behind_below0 = acc.mode == MODE_CONTAINER ? -Real<T>::max() : T(0);
    for (uint32_t base = 0; base < n; base += 32u) {
        const uint32_t count = min(32u, n - base);
        CullCursor<T, SMEM> cur(sv, base, count);
        T behind_below = behind_below0;
        keep_in_register(behind_below);
        uint32_t mask = 0u, bit = 1u;
        for (; !cur.done(); cur.next(), bit <<= 1) {
            T cx, cy, cz, r2;
            cur.load(cx, cy, cz, r2);
            const T ocx = cx - ray.o.x, ocy = cy - ray.o.y, ocz = cz - ray.o.z;
            const T bq = fma(ocz, ray.d.z, fma(ocy, ray.d.y, ocx * ray.d.x));
            const T c2 = fma(ocz, ocz, fma(ocy, ocy, ocx * ocx));
            const T ex = fma(c2, sv.cull_shrink(), -r2);
            const bool outside = ex > T(0), behind = bq < behind_below, misses = ex * acc.dir_sq > bq * bq;
            if (!(outside & (behind | misses))) mask |= bit;
        }
        while (mask) {
            const uint32_t pos = base + (uint32_t)__ffs((int)mask) - 1u;
            mask &= mask - 1u;
            exact_test<T, FULL>(sv, pos, ray, acc);
            // World::is_in_shadow (world.rs:106-111) is an `any`: a lane that has found a blocker is done
            if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) mask = 0u;
        }
        if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) break;
    }
#elif RT_CULL && RT_LEAN_LOOP
    // The loop runs on the cull-record cursor alone: a culled shape (88 % of them on the cover frame) costs the
    // pre-test, one add, one compare and one branch.  Everything else — the shape's position, the addresses of its
    // geometry and meta records — is derived from the cursor on the rare path; the `asm` keeps the compiler from
    // turning those addresses back into loop-carried counters (it emitted five adds per iteration).  The pre-test is
    // evaluated without short-circuits (one data-dependent branch instead of two), and a shadow query that has found
    // its blocker leaves by moving the cursor to the last record instead of a `break` (no per-iteration
    // BSSY / BSYNC pair around the body).
#if RT_CULL_F32
    if constexpr (sizeof(T) == 8) {
        // The f64 kernels run the pre-test in SINGLE precision: its ~15 arithmetic instructions per shape were more than
        // half of all FP64-pipe work of a frame (half rate, 8.4-cycle dependent latency), and a pre-test only has to be
        // safe, not exact.  Safe means: it culls only when, in real arithmetic, the ray's line misses the record's sphere
        // or the sphere lies wholly behind the origin.  With u = 2^-24, M = max(|origin|_inf, |any centre|_inf):
        //   * rounding origin and centre to f32 and subtracting moves oc by at most eps = 4uM per component
        //     (eta = sqrt(3) eps in length); rounding the direction tilts the line by <= 2u, i.e. moves it by <= 2u|oc| at
        //     the centre.  So the true distance line-centre is >= the f32-input one - m, m = eta + 2u|oc|;
        //   * (r + m)^2 <= r^2 (1 + 2^-10) + 1025 m^2 and 1025 m^2 <= 2050 eta^2 + 8200 u^2 |oc|^2 <= 2^-30.4 M^2
        //     (|oc|^2 <= 12 M^2): the packer pads r^2 by 2^-9, the loop adds E = 2^-29 M^2;
        //   * evaluating c2 dd - bq^2 in f32 (fma dots) is off by < 16u c2 dd: the factor 1 - 2^-18 = 1 - 64u on c2
        //     pays for it, including the roundings of the comparison itself;
        //   * "behind": once the origin is outside by these margins (c2 - r^2 >= E/2), a centre whose f32 bq is negative
        //     but whose true bq is not has bq^2 <= (eta + 5u|oc|)^2 dd < E dd / 2 <= dd (c2 - r^2): no real intersection.
        // NaN, infinities and directions whose dd leaves [1e-30, 1e30] make every comparison false or E infinite: no
        // culling.  The exact tests that follow are the f64 ones, so the frame is the same bits with or without this.
        // (Tried: two shapes per iteration with the packed FFMA2 / FADD2 / FMUL2 of sm_100a — 15.5 instead of 24
        // instructions per culled shape, level 0 of the cover frame 7 % faster, but at depth some lane survives in nearly
        // every pair and the two-way rare path costs what the pre-test saves: profiles/r2_notes.md.)
        Cull32Cursor<SMEM> cur(sv, 0u, n);
        float ox = (float)ray.o.x, oy = (float)ray.o.y, oz = (float)ray.o.z;
        float dx = (float)ray.d.x, dy = (float)ray.d.y, dz = (float)ray.d.z;
        // opaque: otherwise ptxas re-converts the six doubles in every iteration (F2F on the FP64 pipe) to save registers
        keep_in_register(ox); keep_in_register(oy); keep_in_register(oz);
        keep_in_register(dx); keep_in_register(dy); keep_in_register(dz);
        float dd = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float m = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fmaxf(fabsf(oz), sv.L.cull_coord_max));
        float pad = (m * m) * 0x1p-29f;
        if (!(dd > 1.0e-30f && dd < 1.0e30f) || !(m < 1.0e18f)) pad = __int_as_float(0x7f800000);
        float behind_below = acc.mode == MODE_CONTAINER ? -__int_as_float(0x7f800000) : 0.0f;
        float shrink = 1.0f - 0x1p-18f;
        keep_in_register(behind_below);
        keep_in_register(pad);
        keep_in_register(dd);
        keep_in_register(shrink);
        for (; !cur.done(); cur.next()) {
            float cx, cy, cz, r2;
            cur.load(cx, cy, cz, r2);
            const float ocx = cx - ox, ocy = cy - oy, ocz = cz - oz;
            const float bq = fmaf(ocz, dz, fmaf(ocy, dy, ocx * dx));
            const float c2 = fmaf(ocz, ocz, fmaf(ocy, ocy, ocx * ocx));
            const float ex = fmaf(c2, shrink, -(r2 + pad));
            const bool outside = ex > 0.0f, behind = bq < behind_below, misses = ex * dd > bq * bq;
            if (outside & (behind | misses)) continue;
            uint32_t pos = cur.position(n);
            asm volatile("" : "+r"(pos));
            exact_test<T, FULL>(sv, pos, ray, acc);
            if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) cur.finish();
        }
        return;
    }
#endif
    CullCursor<T, SMEM> cur(sv, 0u, n);
    // "the centre is behind the origin" only culls when negative distances are of no interest (everything but the
    // container walk): comparing against -max instead of 0 switches it off without a mode test in the loop
    T behind_below = acc.mode == MODE_CONTAINER ? -Real<T>::max() : T(0);
    keep_in_register(behind_below);
    for (; !cur.done(); cur.next()) {
        T cx, cy, cz, r2;
        cur.load(cx, cy, cz, r2);
        const T ocx = cx - ray.o.x, ocy = cy - ray.o.y, ocz = cz - ray.o.z;
        const T bq = fma(ocz, ray.d.z, fma(ocy, ray.d.y, ocx * ray.d.x));
        const T c2 = fma(ocz, ocz, fma(ocy, ocy, ocx * ocx));
        const T ex = fma(c2, sv.cull_shrink(), -r2);
        const bool outside = ex > T(0), behind = bq < behind_below, misses = ex * acc.dir_sq > bq * bq;
        if (outside & (behind | misses)) continue;
        uint32_t pos = cur.position(n);
        asm volatile("" : "+r"(pos));
        exact_test<T, FULL>(sv, pos, ray, acc);
        // World::is_in_shadow (world.rs:106-111) is an `any`: a lane that has found a blocker is done
        if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) cur.finish();
    }
#else
    for (uint32_t pos = 0; pos < n; ++pos) {
#if RT_CULL
        {
            T cx, cy, cz, r2;
            load_cull(sv.cull(pos), cx, cy, cz, r2);
            const T ocx = cx - ray.o.x, ocy = cy - ray.o.y, ocz = cz - ray.o.z;
            const T bq = fma(ocz, ray.d.z, fma(ocy, ray.d.y, ocx * ray.d.x));
            const T c2 = fma(ocz, ocz, fma(ocy, ocy, ocx * ocx));
            const T ex = fma(c2, sv.cull_shrink(), -r2);
            if (ex > T(0) && ((acc.mode != MODE_CONTAINER && bq < T(0)) || ex * acc.dir_sq > bq * bq)) continue;
        }
#endif
        exact_test<T, FULL>(sv, pos, ray, acc);
        if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) break;
    }
#endif
}

// ---- pair-list trace: World::collect_intersections for the 32 rays of a warp at once --------------------------
// trace_unified walks the shape list once per lane: after the first bounce the lanes of a warp disagree about which
// shapes survive the pre-test, so each exact test runs for a handful of lanes (ncu, cover level 5: 8 of 32 threads
// per instruction inside local_intersect) and every pre-test ends in a data-dependent branch that the next one waits
// for.  Here the warp works on (ray, shape) PAIRS instead:
//   A. pre-test: every lane tests ITS ray against a chunk of PAIR_CHUNK shapes, branch-free; survivors are appended to
//      a warp-private list in shared memory (ballot + prefix count), shape-major, i.e. sorted by shape type;
//   B. exact test: whenever 32 pairs are waiting (or the shapes are exhausted) lane j takes pair j — some lane's ray,
//      read from shared memory, against some shape — so the transform and local_intersect run with all 32 lanes and,
//      pairs being sorted by shape, mostly one shape type per round;
//   C. consume: the distances go back through shared memory to the lane that owns the ray, which folds them into its
//      accumulator with the same `consume` as the per-lane loop — (distance, world order) minimum, shadow `any`,
//      or the container bookkeeping — so the result is the same whatever the order of the pairs.
// All 32 lanes of a converged warp must call it (idle lanes contribute no pairs).
#ifndef RT_PAIR_CHUNK
#define RT_PAIR_CHUNK 4
#endif
#ifndef RT_PAIRS_NOINLINE
#define RT_PAIRS_NOINLINE 0
#endif
#if RT_PAIRS_NOINLINE
#define RT_PAIRS_FN __device__ __noinline__
#else
#define RT_PAIRS_FN __device__ __forceinline__
#endif
constexpr int PAIR_CHUNK = RT_PAIR_CHUNK;            // shapes pre-tested between two looks at the list
constexpr int PAIR_RING = PAIR_CHUNK <= 4 ? 256 : 512;  // list capacity (a power of two) >= 31 waiting + PAIR_CHUNK * 32 new ones

template <typename T>
struct PairScratch {  // one per warp, in dynamic shared memory behind the scene tables
    T ox[32], oy[32], oz[32], dx[32], dy[32], dz[32];  // the warp's rays
    T rt[4][32];                                       // distances of the round's pairs, in push order
    int rn[32];                                        // (shape position << 3) | number of distances
    unsigned owner[32];                                // per ray: the pair slots of this round that hold its results
    unsigned list[PAIR_RING + 32];                     // waiting pairs: (shape position << 5) | ray lane; + a dummy slot per lane
};

template <typename T, bool FULL, bool SHADOW_EXIT, bool SMEM>
RT_PAIRS_FN void trace_pairs(const SceneView<T, SMEM>& sv, const Ray<T>& ray, TraceAcc<T>& acc, PairScratch<T>* ws) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned below = (1u << lane) - 1u;
    const uint32_t n = sv.L.type_begin[NUM_SHAPE_TYPES];
    ws->ox[lane] = ray.o.x; ws->oy[lane] = ray.o.y; ws->oz[lane] = ray.o.z;
    ws->dx[lane] = ray.d.x; ws->dy[lane] = ray.d.y; ws->dz[lane] = ray.d.z;
    __syncwarp();
    bool live = acc.mode != MODE_IDLE;
    const bool behind_counts = acc.mode == MODE_CONTAINER;  // negative distances matter to the container walk only
    uint32_t pos = 0, head = 0, tail = 0;
    for (;;) {
        const uint32_t avail = tail - head;
        if (avail < 32u && pos < n) {
            // ---- A: pre-test PAIR_CHUNK shapes (see trace_unified for the test itself) ----
            // Branch-free on purpose: no test depends on another's outcome, so the loads of the whole chunk can be
            // issued first and the chains interleave.  Positions past the end re-test the last shape and are masked.
#pragma unroll
            for (int j = 0; j < PAIR_CHUNK; ++j) {
                const uint32_t p = min(pos + (uint32_t)j, n - 1u);
                bool pass = live & (pos + (uint32_t)j < n);
#if RT_CULL
                T cx, cy, cz, r2;
                load_cull(sv.cull(p), cx, cy, cz, r2);
                const T ocx = cx - ray.o.x, ocy = cy - ray.o.y, ocz = cz - ray.o.z;
                const T bq = fma(ocz, ray.d.z, fma(ocy, ray.d.y, ocx * ray.d.x));
                const T c2 = fma(ocz, ocz, fma(ocy, ocy, ocx * ocx));
                const T ex = fma(c2, sv.cull_shrink(), -r2);
                const bool outside = ex > T(0), behind = bq < T(0), misses = ex * acc.dir_sq > bq * bq;
                pass = pass & !(outside & ((behind & !behind_counts) | misses));
#endif
                const unsigned b = __ballot_sync(0xffffffffu, pass);
                // losers write to a private dummy slot behind the ring: no branch around the store
                const uint32_t slot = pass ? ((tail + (uint32_t)__popc(b & below)) & (PAIR_RING - 1)) : (uint32_t)PAIR_RING + lane;
                ws->list[slot] = (p << 5) | lane;
                tail += (uint32_t)__popc(b);
            }
            pos += PAIR_CHUNK;
            continue;
        }
        if (avail == 0u) break;
        // ---- B: one round of exact tests, lane j on pair head + j ----
        const uint32_t cnt = avail < 32u ? avail : 32u;
        __syncwarp();  // the list entries of this round are visible
        if (lane < cnt) {
            const unsigned e = ws->list[(head + lane) & (PAIR_RING - 1)];
            const uint32_t sp = e >> 5, r = e & 31u;
            Ray<T> wr;
            wr.o = mk<T>(ws->ox[r], ws->oy[r], ws->oz[r]);
            wr.d = mk<T>(ws->dx[r], ws->dy[r], ws->dz[r]);
            const T* g = sv.shape(sp);
            const int4 meta = sv.shape_meta(sp);
            Ray<T> local;  // ray.rs:45-49
            local.o = mat_point(g, wr.o);
            local.d = mat_vector(g, wr.d);
            T t0 = T(0), t1 = T(0), t2 = T(0), t3 = T(0);
            int k = 0;
            switch ((meta.z >> FLAG_TYPE_SHIFT) & 7) {
            case 0: k = local_intersect<T, 0>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
            case 1: k = local_intersect<T, 1>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
            case 2: k = local_intersect<T, 2>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
            case 3: if (FULL) k = local_intersect<T, 3>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
            case 4: if (FULL) k = local_intersect<T, 4>(local, g, meta.z, nullptr, t0, t1, t2, t3); break;
            default: if (FULL) k = local_intersect<T, 5>(local, g, meta.z, sv.triangle(sp), t0, t1, t2, t3); break;
            }
            if (k > 0) {
                ws->rt[0][lane] = t0; ws->rt[1][lane] = t1; ws->rt[2][lane] = t2; ws->rt[3][lane] = t3;
                ws->rn[lane] = (int)((sp << 3) | (uint32_t)k);
                atomicOr(&ws->owner[r], 1u << lane);
            }
        }
        head += cnt;
        __syncwarp();
        // ---- C: every lane folds the results of ITS ray into its accumulator ----
        unsigned mine = ws->owner[lane];
        if (mine) ws->owner[lane] = 0u;
        while (mine) {
            const int j = __ffs((int)mine) - 1;
            mine &= mine - 1u;
            const int rn = ws->rn[j];
            const uint32_t sp = (uint32_t)rn >> 3;
            consume<T, 4>(acc, rn & 7, ws->rt[0][j], ws->rt[1][j], ws->rt[2][j], ws->rt[3][j], (int)sp, sv.shape_meta(sp));
        }
        // World::is_in_shadow (world.rs:106-111) is an `any`: a lane that has found a blocker adds no more pairs
        if (SHADOW_EXIT && acc.mode == MODE_SHADOW && acc.best_pos >= 0) live = false;
        __syncwarp();  // results consumed before the next round overwrites them
    }
}

// utils.rs:16-24
template <typename T> RT_DEV bool coarse_eq(T a, T b) { return a == b || fabs(a - b) < Real<T>::eps(); }

// local_normal_at of the six shapes (sphere.rs:57-59, plane.rs:52-54, cube.rs:89-101,
// cylinder.rs:114-126, cone.rs:116-133, triangle.rs:78-80)
template <typename T, bool SMEM>
RT_COLD V3<T> local_normal_at(const SceneView<T, SMEM>& sv, uint32_t pos, int type, const T* g, V3<T> p) {
    switch (type) {
    case 0: return p;
    case 1: return mk<T>(T(0), T(1), T(0));
    case 2: {
        T ax = fabs(p.x), ay = fabs(p.y), az = fabs(p.z);
        T mx = fmax(fmax(fmax(-Real<T>::max(), ax), ay), az);
        if (coarse_eq(mx, ax)) return mk<T>(p.x, T(0), T(0));
        if (coarse_eq(mx, ay)) return mk<T>(T(0), p.y, T(0));
        return mk<T>(T(0), T(0), p.z);
    }
    case 3: {
        T dist = sq(p.x) + sq(p.z);
        if (dist < T(1) && p.y >= (g[SHAPE_MAX] - Real<T>::eps())) return mk<T>(T(0), T(1), T(0));
        if (dist < T(1) && p.y <= (g[SHAPE_MIN] + Real<T>::eps())) return mk<T>(T(0), T(-1), T(0));
        return mk<T>(p.x, T(0), p.z);
    }
    case 4: {
        T dist = sq(p.x) + sq(p.z);
        if (dist < sq(g[SHAPE_MAX]) && p.y >= (g[SHAPE_MAX] - Real<T>::eps())) return mk<T>(T(0), T(1), T(0));
        if (dist < sq(g[SHAPE_MIN]) && p.y <= (g[SHAPE_MIN] + Real<T>::eps())) return mk<T>(T(0), T(-1), T(0));
        T y = sqrt(dist);
        if (p.y > T(0)) y = -y;
        return mk<T>(p.x, y, p.z);
    }
    default: return ld3(sv.triangle(pos) + 9);
    }
}

// Rust `f64 as i64` (saturating, NaN -> 0) followed by `% 2 == 0`
template <typename T>
RT_COLD bool even_as_i64(T v) {
    if (v != v) return true;                               // NaN as i64 = 0
    if (v >= T(9223372036854775808.0)) return false;       // saturates to i64::MAX (odd)
    if (v <= T(-9223372036854775808.0)) return true;       // saturates to i64::MIN (even)
    return (((long long)v) % 2) == 0;
}

// Pattern::color_at (patterns/*.rs) at pattern-space point p
template <typename T, bool SMEM>
RT_COLD V3<T> pattern_color_at(const SceneView<T, SMEM>& sv, int pattern, V3<T> p) {
    for (;;) {
        const T* pr = sv.pattern(pattern);
        const int* pm = sv.pattern_meta(pattern);
        V3<T> a = ld3(pr), b = ld3(pr + 3);
        switch (pm[0]) {
        case 0: return even_as_i64(floor(p.x)) ? a : b;  // stripe_pattern.rs:24-31
        case 1: {                                        // gradient_pattern.rs:24-31
            V3<T> distance = b - a;
            T fraction = fabs(p.x - trunc(p.x));
            if (!even_as_i64(p.x)) fraction = T(1) - fraction;
            return a + (distance * fraction);
        }
        case 2: return even_as_i64(floor(sqrt_native(sq(p.x) + sq(p.z)))) ? a : b;   // ring_pattern.rs:25-32
        case 3: return even_as_i64(floor(p.x) + floor(p.y) + floor(p.z)) ? a : b;  // checker_pattern.rs:24-31
        case 4: pattern = even_as_i64(floor(p.x)) ? pm[1] : pm[2]; break;  // complex_pattern.rs:24-33
        default: return p;                                                 // TestPattern, pattern.rs:62-66
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The pieces of a node both kernel families (and the known-answer probes of rt_probe.cuh) are built from.

// Image row of the k-th row a launch renders (rt_scene.h CameraParams)
template <typename T>
RT_DEV uint32_t image_row(const CameraParams<T>& cam, uint32_t k) {
    const uint32_t q = k / cam.band_rows;
    return ((q / cam.band_take) * cam.shard_count + cam.shard_index + q % cam.band_take) * cam.band_rows + k % cam.band_rows;
}

// Camera::ray_for_pixel, camera.rs:52-68 (x, y: image column and row)
template <typename T>
RT_DEV Ray<T> camera_ray(const CameraParams<T>& cam, uint32_t x, uint32_t y) {
    T offset_x = (T(x) + T(0.5)) * cam.pixel_size;
    T offset_y = (T(y) + T(0.5)) * cam.pixel_size;
    T world_x = cam.half_width - offset_x;
    T world_y = cam.half_height - offset_y;
    V3<T> pixel = mat_point(cam.inv, mk<T>(world_x, world_y, T(-1)));
    V3<T> origin = ld3(cam.origin);
    Ray<T> ray;
    ray.o = origin;
    ray.d = normalized(pixel - origin);
    return ray;
}

// Shape::normal_at, shape.rs:22-27: world point -> object space -> local normal -> world space, normalised
template <typename T, bool SMEM>
RT_DEV V3<T> world_normal_at(const SceneView<T, SMEM>& sv, uint32_t pos, int type, const T* g, V3<T> point) {
    V3<T> local_point = mat_point(g, point);
    V3<T> local_normal = local_normal_at(sv, pos, type, g, local_point);
    return normalized(mat_transposed_vector(g, local_normal));
}

// World::refracted_color's direction (world.rs:136-150): false on total internal reflection, else the refracted
// direction is normal * a - eye * n_ratio
template <typename T>
RT_DEV bool refraction_coefficients(T n1, T n2, T cos_i, T& a, T& n_ratio) {
    n_ratio = n1 / n2;
    T sin2_t = sq(n_ratio) * (T(1) - sq(cos_i));
    if (sin2_t > T(1)) return false;
    T cos_t = sqrt(T(1) - sin2_t);
    a = fma(n_ratio, cos_i, -cos_t);
    return true;
}

// ComputedHit::schlicks_approximation, computed_hit.rs:50-68
template <typename T>
RT_DEV T schlick_reflectance(T n1, T n2, T cos_i) {
    T c = cos_i;
    if (n1 > n2) {
        T ratio = n1 / n2;
        T sin2_t = sq(ratio) * (T(1) - sq(c));
        if (sin2_t > T(1)) return T(1);
        c = sqrt(T(1) - sin2_t);
    }
    T r0 = sq((n1 - n2) / (n1 + n2));
    T x = T(1) - c;
    T x5 = x * ((x * x) * (x * x));  // powi(5)
    return fma(T(1) - r0, x5, r0);
}

// Material::resolve_color, material.rs:75-80, at the over point (material.rs:116-130)
template <typename T, bool SMEM>
RT_DEV V3<T> resolve_color(const SceneView<T, SMEM>& sv, uint32_t material, uint32_t pos, V3<T> over) {
    const int pat = sv.material_pattern(material);
    if (pat < 0) return ld3(sv.material(material));
    V3<T> object_point = mat_point(sv.shape(pos), over);  // pattern.rs:10-14
    V3<T> pattern_point = mat_point(sv.pattern((uint32_t)pat) + 6, object_point);
    return pattern_color_at(sv, pat, pattern_point);
}

// Material::lighting, material.rs:53-114, for one light.  light_dir = normalized(light.position - over_point).
// specular == 0 (every matte material): (intensity * 0) * pow(..) is an exact zero for the finite factor pow returns
// on (0, ~1], so the whole term — and the pow call — is skipped.
template <typename T>
RT_DEV V3<T> phong_lighting(const T* m, V3<T> base, V3<T> intensity, V3<T> light_dir, V3<T> eye, V3<T> normal, bool in_shadow) {
    V3<T> effective = hadamard(base, intensity);
    V3<T> ambient = effective * m[MAT_AMBIENT];
    if (in_shadow) return ambient;
    T ldn = dot(light_dir, normal);
    if (ldn < T(0)) return ambient;
    V3<T> diffuse = (effective * m[MAT_DIFFUSE]) * ldn;
    V3<T> refl = reflect(neg(light_dir), normal);
    T rde = dot(refl, eye);
    if (rde <= T(0) || m[MAT_SPECULAR] == T(0)) return ambient + diffuse;
    T factor = pow(rde, m[MAT_SHININESS]);
    V3<T> specular = (intensity * m[MAT_SPECULAR]) * factor;
    return (ambient + diffuse) + specular;
}

// Canvas::to_png_file, canvas.rs:117-123: clamp to [0, 1], * 255, round half away from zero, NaN -> 0
template <typename T>
RT_DEV uint8_t quantise(T v) {
    v = (v < T(0)) ? T(0) : v;
    v = (v > T(1)) ? T(1) : v;
    v = round(v * T(255));
    return (v != v) ? (uint8_t)0 : (uint8_t)v;
}

// One frame of the explicit recursion stack = one World::shade_hit in flight (world.rs:38-67).
template <typename T>
struct Frame {
    V3<T> a;         // reflect direction until the reflect child is launched, then the reflected colour
    V3<T> surface;   // sum over lights (world.rs:43-53)
    V3<T> refr_o;    // refracted ray origin (world.rs:152) until that child is launched, then the refracted colour
    V3<T> refr_d;    // refracted ray direction
    T reflectance;   // Schlick (computed_hit.rs:50-68), only when reflective && transparent
    T k_reflect, k_transparent;
    int flags;
};
enum : int { FR_REFLECT = 1, FR_REFRACT = 2, FR_SCHLICK = 4, FR_WAIT_REFLECT = 8, FR_WAIT_REFRACT = 16 };

enum : int { ST_FETCH = 0, ST_RADIANCE = 1, ST_CONTAINER = 2, ST_SHADOW = 3, ST_DONE = 4 };

#ifndef RT_BLOCK_THREADS
#define RT_BLOCK_THREADS 128
#endif
#ifndef RT_MIN_BLOCKS_PER_SM
#define RT_MIN_BLOCKS_PER_SM 4
#endif

#ifndef RT_PHASE_SYNC
#define RT_PHASE_SYNC 0
#endif

#ifndef RT_TILE_ORDER
// 0 scanline, 1 middle-out, 2 bottom-up, 3 scattered.  Per-pixel cost varies by 50x and is spatially
// clustered; in scanline order the expensive rows of a typical frame (floor reflections, glass) come
// last and the launch ends with a long thin tail (measured: 5.17 ms scanline vs 4.39 ms heavy-first
// vs 4.53 ms scattered on cover@1080p).  Scattered needs no knowledge of the scene.
#define RT_TILE_ORDER 3
#endif

constexpr int TILE_W = 8, TILE_H = 4;   // a warp's 32 pixel slots = one 8x4 tile
#ifndef RT_CHUNK_SLOTS
#define RT_CHUNK_SLOTS 64
#endif
constexpr int CHUNK_SLOTS = RT_CHUNK_SLOTS;  // slots a warp takes from the global counter at a time

template <typename T, int MAX_FRAMES, bool FULL, bool BVH, bool SMEM>
__global__ void __launch_bounds__(RT_BLOCK_THREADS, RT_MIN_BLOCKS_PER_SM)
render_kernel(const T* __restrict__ g_reals, const int* __restrict__ g_ints, SceneLayout layout, CameraParams<T> cam,
              T* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8, unsigned long long* __restrict__ counters,
              unsigned int* __restrict__ work_counter) {
    SceneView<T, SMEM> sv;
    sv.L = layout;
    sv.reals = g_reals;
    sv.ints = g_ints;
    if constexpr (SMEM) stage_scene<T>(layout, g_reals, g_ints);

    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (cam.hsize + TILE_W - 1) / TILE_W;
    const uint32_t tiles_y = (cam.n_rows + TILE_H - 1) / TILE_H;
    const uint32_t total_slots = tiles_x * tiles_y * (TILE_W * TILE_H);
    const int n_lights = (int)layout.n_lights;

    Frame<T> stack[MAX_FRAMES];

    // lane state
    int state = ST_FETCH;
    int depth = 0;
    Ray<T> ray;
    ray.o = mk<T>(T(0), T(0), T(0));
    ray.d = mk<T>(T(0), T(0), T(1));
    size_t out_index = 0;
    // current node
    V3<T> over = mk<T>(T(0), T(0), T(0)), under = over, normal = over, eye = over, base = over, surface = over;
    T t_hit = T(0), shadow_distance = T(0);
    int hit_pos = -1, hit_material = 0, light = 0;
    V3<T> node_dir = over;  // direction of the node's radiance ray (for the reflect vector)
    // warp-private chunk of pixel slots
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;
    // counters
    unsigned int c_primary = 0, c_shadow = 0, c_reflect = 0, c_refract = 0, c_nodes = 0;

    for (;;) {
        // ---- phase A: idle lanes take the next pixel slot of the warp's chunk ---------------------
        {
            if (exhausted && state == ST_FETCH) state = ST_DONE;
            unsigned need = __ballot_sync(0xffffffffu, state == ST_FETCH);
            while (need) {
                if (chunk_next >= chunk_end) {
                    uint32_t base_slot = 0;
                    if (lane == 0) base_slot = atomicAdd(work_counter, (unsigned)CHUNK_SLOTS);
                    base_slot = __shfl_sync(0xffffffffu, base_slot, 0);
                    chunk_next = base_slot;
                    chunk_end = base_slot + CHUNK_SLOTS;
                    if (base_slot >= total_slots) {  // frame exhausted
                        if (state == ST_FETCH) state = ST_DONE;
                        exhausted = true;
                        break;
                    }
                }
                unsigned rank = __popc(need & ((1u << lane) - 1u));
                uint32_t avail = chunk_end - chunk_next;
                bool mine = (state == ST_FETCH) && rank < avail;
                if (mine) {
                    uint32_t slot = chunk_next + rank;
                    state = ST_DONE;  // unless the slot is a real pixel
                    if (slot < total_slots) {
                        uint32_t tile = slot / (TILE_W * TILE_H), in = slot % (TILE_W * TILE_H);
                        uint32_t tile_row = tile / tiles_x;
#if RT_TILE_ORDER == 1
                        // rows of tiles from the middle of the frame outwards: the expensive pixels of a
                        // typical scene (glass, mirrors) sit near the centre and should start first
                        {
                            const uint32_t mid = tiles_y / 2;
                            const uint32_t h = (tile_row + 1) / 2;
                            tile_row = (tile_row & 1u) ? (mid >= h ? mid - h : tiles_y - 1 - (h - mid - 1)) : (mid + h < tiles_y ? mid + h : (tiles_y - 1) - (mid + h - tiles_y));
                        }
#elif RT_TILE_ORDER == 2
                        tile_row = tiles_y - 1 - tile_row;  // bottom-up (experiment)
#elif RT_TILE_ORDER == 3
                        // scattered: a multiplicative permutation of the tile index (stride coprime to the
                        // tile count), so that every part of the frame is sampled all along the launch
                        {
                            const uint32_t n_tiles = tiles_x * tiles_y;
                            tile = (uint32_t)(((unsigned long long)tile * cam.tile_stride) % n_tiles);
                            tile_row = tile / tiles_x;
                        }
#endif
                        uint32_t x = (tile % tiles_x) * TILE_W + in % TILE_W;
                        uint32_t k = tile_row * TILE_H + in / TILE_W;
                        state = ST_FETCH;  // padding slot: try again on the next round
                        if (x < cam.hsize && k < cam.n_rows) {
                            uint32_t y = image_row(cam, k);
                            ray = camera_ray(cam, x, y);
                            if (cam.probe_ray) ray.d = mk<T>(cam.inv[0], cam.inv[1], cam.inv[2]);
                            out_index = (size_t)(cam.out_full_frame ? y : k) * cam.hsize + x;
                            depth = 0;
                            state = ST_RADIANCE;
                            ++c_primary;
                        }
                    } else {
                        state = ST_DONE;
                    }
                }
                uint32_t taken = min(avail, (uint32_t)__popc(need));
                chunk_next += taken;
                need = __ballot_sync(0xffffffffu, state == ST_FETCH);
            }
        }
#if RT_PHASE_SYNC
        // CTA-wide phase lockstep: every warp of the CTA traces, then every warp shades.  The kernel's
        // code does not fit the instruction cache; keeping the warps of an SM in the same phase makes
        // them share the lines they fetch.
        if (__syncthreads_and(state == ST_DONE)) break;
#else
        if (__all_sync(0xffffffffu, state == ST_DONE)) break;
#endif

        // ---- phase B: one trace for every lane that has a ray ----------------------------------------
        TraceAcc<T> acc;
        ContainerAcc<T> cacc;
        acc.c = RT_ACC_SPLIT ? &cacc : acc_store(acc);
        acc.mode = (state == ST_RADIANCE) ? MODE_RADIANCE : (state == ST_SHADOW) ? MODE_SHADOW : (state == ST_CONTAINER) ? MODE_CONTAINER : MODE_IDLE;
        acc.best_t = (state == ST_SHADOW) ? shadow_distance : Real<T>::max();
        acc.dir_sq = fma(ray.d.z, ray.d.z, fma(ray.d.y, ray.d.y, ray.d.x * ray.d.x));
        // intersection.rs:77-79: a blocker needs t < light distance, strictly: with the seed order -1 the
        // `t == best_t && order < best_orig` arm of consume() can never accept a hit AT the light's distance
        acc.best_orig = (state == ST_SHADOW) ? -1 : 0x7fffffff;
        acc.best_pos = -1;
        if (state == ST_CONTAINER) {  // the container bookkeeping lives in local memory: only touch it when it is used
            acc.c->t_hit = t_hit;
            acc.c->hit_class = sv.shape_meta((uint32_t)hit_pos).w;
            acc.c->hit_class_inside = false;
            acc.c->all_pos = acc.c->excl_pos = -1;
            acc.c->all_t = acc.c->excl_t = T(0);
            acc.c->all_orig = acc.c->excl_orig = 0;
        }
        if (acc.mode != MODE_IDLE) trace_unified<T, FULL, false>(sv, ray, acc);  // uniform list (BVH scenes: the unbounded shapes)
        if (BVH) trace_bvh<T, FULL>(sv, ray, acc);  // every lane takes part: warp votes inside

#if RT_PHASE_SYNC
        __syncthreads();
#endif
        // ---- phase C: consume the result ------------------------------------------------------------
        bool finish_hit = false;     // ComputedHit complete -> start the light loop
        bool after_lights = false;   // surface colour complete -> children
        bool returning = false;      // a colour is ready for the parent
        V3<T> colour = mk<T>(T(0), T(0), T(0));
        T n1 = T(1), n2 = T(1);  // Material::DEFAULT_REFRACTIVE_INDEX, material.rs:24

        if (state == ST_RADIANCE) {
            // World::internal_color_at, world.rs:70-86
            if (acc.best_pos < 0) {
                returning = true;  // World::DEFAULT_COLOR
            } else {
                ++c_nodes;
                hit_pos = acc.best_pos;
                t_hit = acc.best_t;
                const T* g = sv.shape((uint32_t)hit_pos);
                int4 meta = sv.shape_meta((uint32_t)hit_pos);
                hit_material = meta.y;
                // Intersection::prepare_computations, intersection.rs:21-31
                V3<T> point = ray.o + ray.d * t_hit;
                normal = world_normal_at(sv, (uint32_t)hit_pos, (meta.z >> FLAG_TYPE_SHIFT) & 7, g, point);
                eye = neg(ray.d);
                if (dot(normal, eye) < T(0)) normal = neg(normal);
                node_dir = ray.d;
                // computed_hit.rs:33-34
                const T off = Real<T>::offset(point.x, point.y, point.z, t_hit);
                over = point + (normal * off);
                under = point - (normal * off);
                const T* m = sv.material((uint32_t)hit_material);
                // n1 / n2 feed refracted_color (world.rs:136, needs remaining > 0) and Schlick (world.rs:59,
                // whose result multiplies black children when remaining == 0): only then walk containers
                bool need_containers = (m[MAT_TRANSPARENCY] != T(0)) && ((int)cam.max_depth - depth > 0);
                if (need_containers) state = ST_CONTAINER;  // same ray, container query
                else finish_hit = true;
            }
        } else if (state == ST_CONTAINER) {
            // intersection.rs:33-62
            const int hit_mat = hit_material;
            n1 = (acc.c->all_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)acc.c->all_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
            if (acc.c->hit_class_inside)
                n2 = (acc.c->excl_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)acc.c->excl_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
            else
                n2 = sv.material((uint32_t)hit_mat)[MAT_REFRACTIVE_INDEX];
            finish_hit = true;
        } else if (state == ST_SHADOW) {
            // Material::lighting for light `light`, material.rs:53-114, evaluated at over_point (material.rs:116-130)
            const T* m = sv.material((uint32_t)hit_material);
            const T* lt = sv.light((uint32_t)light);
            const bool in_shadow = acc.best_pos >= 0;
            // the shadow ray's direction is normalized(light.position - over_point): the light vector of material.rs:88
            const V3<T> lit = phong_lighting(m, base, ld3(lt + 3), ray.d, eye, normal, in_shadow);
            surface = surface + lit;  // fold(Color::BLACK, Color::add), world.rs:53
            ++light;
            if (light >= n_lights) after_lights = true;
        }

        if (finish_hit) {
            const T* m = sv.material((uint32_t)hit_material);
            const int remaining = (int)cam.max_depth - depth;
            Frame<T>& f = stack[depth];
            f.flags = 0;
            f.k_reflect = m[MAT_REFLECTIVENESS];
            f.k_transparent = m[MAT_TRANSPARENCY];
            f.a = mk<T>(T(0), T(0), T(0));       // reflected colour unless a reflect child runs (world.rs:121)
            f.refr_o = mk<T>(T(0), T(0), T(0));  // refracted colour unless a refract child runs (world.rs:137,146)
            if (remaining > 0 && m[MAT_REFLECTIVENESS] != T(0)) {  // world.rs:120
                f.a = reflect(node_dir, normal);  // intersection.rs:31; replaced by the colour on return
                f.flags |= FR_REFLECT;
            }
            T cos_i = dot(eye, normal);
            if (remaining > 0 && m[MAT_TRANSPARENCY] != T(0)) {  // world.rs:136-154
                T a, n_ratio;
                if (refraction_coefficients(n1, n2, cos_i, a, n_ratio)) {
                    f.refr_o = under;
                    f.refr_d = (normal * a) - (eye * n_ratio);
                    f.flags |= FR_REFRACT;
                }
            }
            if (m[MAT_REFLECTIVENESS] > T(0) && m[MAT_TRANSPARENCY] > T(0)) {  // world.rs:59
                f.flags |= FR_SCHLICK;
                f.reflectance = schlick_reflectance(n1, n2, cos_i);
            }
            base = resolve_color(sv, (uint32_t)hit_material, (uint32_t)hit_pos, over);  // the same for every light of this node
            surface = mk<T>(T(0), T(0), T(0));
            light = 0;
            if (n_lights == 0) after_lights = true;
            else state = ST_SHADOW;
        }

        if (state == ST_SHADOW && !after_lights) {
            // World::is_in_shadow, world.rs:98-112: next shadow ray
            V3<T> to_light = ld3(sv.light((uint32_t)light)) - over;
            ray.o = over;
            Normalized<T> nl = normalize_full(to_light);  // magnitude() and normalized() take the same sqrt
            ray.d = nl.v;
            shadow_distance = nl.magnitude;
            ++c_shadow;
        }

        if (after_lights) stack[depth].surface = surface;

        // ---- children and returns: World::shade_hit's tail (world.rs:55-66) as a post-order walk ----
        bool advance = after_lights;
        while (advance || returning) {
            if (returning) {
                if (depth == 0) {  // Camera::render_parallel writes the pixel, camera.rs:108
                    if (out_rgb) store_rgb(out_rgb + out_index * 3, colour);
                    if (out_rgb8) {
                        out_rgb8[out_index * 3 + 0] = quantise(colour.x);
                        out_rgb8[out_index * 3 + 1] = quantise(colour.y);
                        out_rgb8[out_index * 3 + 2] = quantise(colour.z);
                    }
                    state = ST_FETCH;
                    break;
                }
                --depth;
                Frame<T>& p = stack[depth];
                if (p.flags & FR_WAIT_REFLECT) {
                    p.flags &= ~FR_WAIT_REFLECT;
                    p.a = colour * p.k_reflect;  // world.rs:127
                } else {
                    p.flags &= ~FR_WAIT_REFRACT;
                    p.refr_o = colour * p.k_transparent;  // world.rs:156
                }
                returning = false;
            }
            advance = false;
            Frame<T>& f = stack[depth];
            if (f.flags & FR_REFLECT) {
                // World::reflected_color, world.rs:114-128.  Only reachable straight after this node's
                // light loop, so `over` still is this node's over_point.
                ray.o = over;
                ray.d = f.a;
                f.flags = (f.flags & ~FR_REFLECT) | FR_WAIT_REFLECT;
                ++depth;
                state = ST_RADIANCE;
                ++c_reflect;
                break;
            }
            if (f.flags & FR_REFRACT) {
                // World::refracted_color, world.rs:130-157
                ray.o = f.refr_o;
                ray.d = f.refr_d;
                f.flags = (f.flags & ~FR_REFRACT) | FR_WAIT_REFRACT;
                ++depth;
                state = ST_RADIANCE;
                ++c_refract;
                break;
            }
            // world.rs:59-66
            if (f.flags & FR_SCHLICK)
                colour = (f.surface + (f.a * f.reflectance)) + (f.refr_o * (T(1) - f.reflectance));
            else
                colour = (f.surface + f.a) + f.refr_o;
            returning = true;
        }
    }

    // ---- work counters (rays = World::collect_intersections calls) ---------------------------------
    if (counters) {
        unsigned int vals[5] = {c_primary, c_shadow, c_reflect, c_refract, c_nodes};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            unsigned int v = vals[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(&counters[k], (unsigned long long)v);
        }
        unsigned int px = c_primary;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) px += __shfl_xor_sync(0xffffffffu, px, o);
        if (lane == 0 && px) atomicAdd(&counters[COUNTER_PIXELS], (unsigned long long)px);
    }
}

}  // namespace rt
