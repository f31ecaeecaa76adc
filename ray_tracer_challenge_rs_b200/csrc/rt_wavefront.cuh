// rt_wavefront.cuh — the render pass as a wavefront over recursion LEVELS (the second kernel family).
//
// The persistent kernel of rt_kernel.cuh evaluates a pixel's whole recursion tree (world.rs:70-157) on one
// lane, one ray at a time: a camera ray that meets glass grows a tree of ~100 dependent rays, and the launch
// cannot end before the slowest such chain does (profiles/r1_notes.md: a 1080p frame stops scaling at ~4
// GPUs).  Here the unit of work is a NODE of some pixel's tree, and all nodes of one depth are processed by
// one launch:
//
//   level d kernel : one thread per HIT of depth d (level 0: per pixel, the camera ray is traced first).
//                    Create the node: prepare_computations (intersection.rs:21-75, with the container
//                    re-trace when the material is transparent), the light loop with its shadow rays
//                    (world.rs:43-53); then trace the reflected / refracted rays (world.rs:114-157) right
//                    away and append only those that HIT something to the queue of level d+1 — a miss is
//                    World::DEFAULT_COLOR, i.e. the black the parent's slot already holds.  So every queue
//                    entry is a node-to-be and no lane of a deeper launch idles on a miss.
//   combine kernel : levels from the deepest up.  A node's colour is `surface + reflected + refracted`
//                    (Schlick-weighted when reflective and transparent, world.rs:59-66) and is written, scaled
//                    by the parent's reflectiveness / transparency (world.rs:127,156), into the parent's
//                    slot — or into the pixel for level 0.
//
// Two refinements keep warps full and HBM traffic low:
//   * a queue is filled from both ends — hits on opaque materials from the front, hits on transparent
//     materials from the back — so a warp of a deeper launch is all-opaque or all-glass and the container
//     and refraction phases are skipped by whole warps instead of running for one or two lanes;
//   * a node without children (no iterations left, or a matte material) never becomes a record: its colour
//     is final when its light loop ends and goes straight into the parent's slot (or the pixel).
//
// Every arithmetic operation of a node is the one the persistent kernel (and the reference) performs, in
// the same order; only WHEN a node is evaluated changes.  The critical path is max_depth+1 launches
// instead of the longest ray chain, lanes of a warp are always in the same phase, and sibling rays stay
// together in the queues.  Queues and node records live in HBM: 64 B per queued ray, 112 B per node (f64).
#pragma once

#include "rt_kernel.cuh"

namespace rt {

// A queued hit: the radiance ray, where it hit, and whose child it is.
#ifndef RT_WF_RAY_ALIGN
#define RT_WF_RAY_ALIGN 16
#endif
template <typename T>
struct alignas(RT_WF_RAY_ALIGN) WfRay {
    T ox, oy, oz, dx, dy, dz;
    T t;         // hit distance (Intersections::hit, intersections.rs:13-18)
    int pos;     // sorted position of the hit shape
    int pad;
    // the tail is read again when the node is finished (it is not kept in registers across the node's traces)
    T k;         // the parent's reflectiveness (slot 0) / transparency (slot 1): scales the colour sent back
    int parent;  // node that spawned the ray
    int slot;    // 0: its reflected colour, 1: its refracted colour
};

template <typename T>
struct alignas(32) WfNode {  // f64: 96 bytes = three 32-byte sectors exactly
    int link;         // >= 0: the parent node; < 0: a level-0 node, ~link is the output index of its pixel
    int slot_flags;   // bit 0: slot in the parent (0 reflected, 1 refracted); bit 1: Schlick blend (world.rs:59)
    T k_parent;       // scale applied to this node's colour when it is handed to the parent (world.rs:127,156)
    T reflectance;
    T surface[3];
    T reflected[3];   // already scaled by k_reflect (world.rs:127); black until a child reports
    T refracted[3];   // already scaled by k_transparent (world.rs:156)
};

// Binned queues (scenes with a short uniform shape list): a level's entries are CONSUMED grouped by (shape they hit,
// reflected / refracted), through a permutation that wf_bin_kernel builds between two level launches.  Entries carry
// their parent link, so the order changes no bit of the frame; it changes which 32 entries share a warp: entries that
// hit the same shape from the same kind of ray have similar origins, normals and — mostly — directions, so their shadow
// and child rays survive the pre-test at the same shapes.  (Measured first with a host-side sort between the launches,
// benchmarks/sort_experiment.py: levels 1-6 of the cover frame 11 % faster; sorting by kind alone is slower than no
// sort.)  Bin = min(hit position, RT_WF_BINS / 2 - 1) * 2 + slot.
#ifndef RT_WF_BINS
#define RT_WF_BINS 64
#endif

// Device-side bookkeeping of one frame.
struct WfCounts {
    unsigned n_rays[16 + 2];   // hits queued for level d at the FRONT of its queue: opaque materials
    unsigned n_back[16 + 2];   // hits queued for level d at the BACK of its queue: transparent materials
    unsigned node_end[16 + 2]; // nodes created by levels 0..d end at node_end[d] (node_end[-1] = 0 implied)
    unsigned bins[16 + 2][2][RT_WF_BINS];  // binned queues: entries of level d's front / back end per (hit shape, reflect / refract) bin
    unsigned n_nodes;          // nodes allocated so far
    unsigned work;             // chunk cursor of the running launch
    unsigned done;             // CTAs of the running launch that have finished (the last one closes the level)
    unsigned overflow;         // a queue or the node array was too small: the frame must be re-rendered
};

// Programmatic dependent launch, device side (sm_90+; both are no-ops for a kernel launched the ordinary way).
RT_DEV void wf_release_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
RT_DEV void wf_wait_for_previous() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifndef RT_WF_SYNC
#define RT_WF_SYNC 0
#endif
#ifndef RT_WF_THREADS
#define RT_WF_THREADS 128
#endif
#ifndef RT_WF_MIN_BLOCKS
#define RT_WF_MIN_BLOCKS 6  // sphere / plane / cube kernels: 80 registers, no spills worth mentioning
#endif
#ifndef RT_WF_MIN_BLOCKS_FULL
#define RT_WF_MIN_BLOCKS_FULL 4  // kernels that also hold the cylinder / cone / triangle tests spill below ~120 registers (cylinders: 1.07 ms at 6, 0.93 at 4)
#endif

template <typename T>
RT_DEV void wf_store_pixel(T* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8, size_t out_index, V3<T> colour) {
    if (out_rgb) store_rgb(out_rgb + out_index * 3, colour);
    if (out_rgb8) {
        out_rgb8[out_index * 3 + 0] = quantise(colour.x);
        out_rgb8[out_index * 3 + 1] = quantise(colour.y);
        out_rgb8[out_index * 3 + 2] = quantise(colour.z);
    }
}

template <typename T>
RT_DEV void wf_reset_acc(TraceAcc<T>& acc, int mode, const Ray<T>& ray, T best_t) {
    acc.mode = mode;
    acc.best_t = best_t;
    acc.dir_sq = fma(ray.d.z, ray.d.z, fma(ray.d.y, ray.d.y, ray.d.x * ray.d.x));
    // strict `t < light distance` for shadow queries (intersection.rs:77-79): see render_kernel
    acc.best_orig = (mode == MODE_SHADOW) ? -1 : 0x7fffffff;
    acc.best_pos = -1;
    if (mode == MODE_CONTAINER) {  // the container bookkeeping lives in local memory: only touch it when it is used
        acc.c->t_hit = T(0);
        acc.c->hit_class = -1;
        acc.c->hit_class_inside = false;
        acc.c->all_pos = acc.c->excl_pos = -1;
        acc.c->all_t = acc.c->excl_t = T(0);
        acc.c->all_orig = acc.c->excl_orig = 0;
    }
}

constexpr int WF_PARK_REALS = 19;  // per-thread node state parked in shared memory (wf_level_kernel, PK_*)

#ifndef RT_WF_PAIRS
#define RT_WF_PAIRS 0  // 1: the uniform shape list is traced as compacted (ray, shape) pairs (trace_pairs; measured slower, profiles/r2_notes.md); 0: per-lane loop
#endif

// Called by all 32 lanes of a converged warp.
template <typename T, bool FULL, bool BVH, bool SMEM>
RT_DEV void wf_trace(const SceneView<T, SMEM>& sv, const Ray<T>& ray, TraceAcc<T>& acc, PairScratch<T>* ws) {
    if (RT_WF_PAIRS && !BVH) {
        trace_pairs<T, FULL, true>(sv, ray, acc, ws);
    } else {
        if (acc.mode != MODE_IDLE) trace_unified<T, FULL, true>(sv, ray, acc);  // BVH scenes: the unbounded shapes
        if (BVH) trace_bvh<T, FULL>(sv, ray, acc);  // warp votes inside: every lane takes part
    }
}

template <typename T, bool FULL, bool BVH, bool SMEM>
__global__ void __launch_bounds__(RT_WF_THREADS, FULL ? RT_WF_MIN_BLOCKS_FULL : RT_WF_MIN_BLOCKS)
wf_level_kernel(const T* __restrict__ g_reals, const int* __restrict__ g_ints, SceneLayout layout, CameraParams<T> cam, int level,
                const WfRay<T>* __restrict__ rays_in, WfRay<T>* __restrict__ rays_out, unsigned cap_rays, WfNode<T>* __restrict__ nodes,
                unsigned cap_nodes, WfCounts* __restrict__ counts, T* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8,
                unsigned long long* __restrict__ counters, const unsigned* __restrict__ perm_in, unsigned long long* __restrict__ keys_out) {
    // perm_in: consume the input queue through this permutation (nullptr: in arrival order); keys_out: record (bin, rank
    // within the bin) of every entry appended to the output queue, for wf_bin_kernel (nullptr: the next level is not binned)
    SceneView<T, SMEM> sv;
    sv.L = layout;
    sv.reals = g_reals;
    sv.ints = g_ints;

    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tiles_x = (cam.hsize + TILE_W - 1) / TILE_W;
    const uint32_t tiles_y = (cam.n_rows + TILE_H - 1) / TILE_H;
    // Programmatic dependent launch (rtgpu.cu launch_chain): the next kernel of the frame may be scheduled as soon as
    // this grid's CTAs leave the SMs; it stages its scene tables while our tail is still running and then waits, as we do
    // here, for the previous grid to have completed and flushed before it touches anything that grid wrote.
    wf_release_dependents();
    if constexpr (SMEM) stage_scene<T>(layout, g_reals, g_ints);  // reads only the scene blobs: before the wait
    wf_wait_for_previous();
    // level > 0: front entries, padding to a warp boundary, then back entries
    unsigned n_front = 0, n_back = 0;
    if (level > 0) {
        n_front = counts->n_rays[level];
        n_back = counts->n_back[level];
        if ((unsigned long long)n_front + n_back > cap_rays) {  // the producer overflowed (it has flagged it): stay in bounds
            n_front = min(n_front, cap_rays);
            n_back = min(n_back, cap_rays - n_front);
        }
    }
    const unsigned front_padded = (n_front + 31u) & ~31u;
    const unsigned n_items = level == 0 ? tiles_x * tiles_y * (TILE_W * TILE_H) : front_padded + n_back;
    const int n_lights = (int)layout.n_lights;
    const int remaining = (int)cam.max_depth - level;
    // work counters of the warp (primary, reflect, refract, nodes; shadow rays = nodes x lights): in shared memory, fed
    // by ballots once per work item, instead of five registers per lane that would be live across every trace
    __shared__ unsigned s_count[RT_WF_THREADS / 32][4];
    unsigned* const my_count = s_count[threadIdx.x >> 5];
    if (lane < 4u) my_count[lane] = 0u;
    __syncwarp();
    // behind the staged tables: the warps' scratch for the pair-list trace (if compiled in), then WF_PARK_REALS columns
    // of per-thread node state
    PairScratch<T>* const ws = reinterpret_cast<PairScratch<T>*>(sv.scratch()) + (threadIdx.x >> 5);
    // The thread's column base as an opaque 32-bit offset into the dynamic shared window: ptxas otherwise rebuilds the
    // address from the layout constants at every use (38 sites, 3.4 % of the level-0 instructions), and an opaque POINTER
    // would turn the accesses into generic loads.
    extern __shared__ __align__(16) unsigned char rt_scene_smem[];
    uint32_t park_at = (uint32_t)(sv.scratch() - rt_scene_smem) + (uint32_t)((RT_WF_PAIRS && !BVH) ? (RT_WF_THREADS / 32) * sizeof(PairScratch<T>) : 0) +
                       threadIdx.x * (uint32_t)sizeof(T);
    asm volatile("" : "+r"(park_at));
    T* const park = reinterpret_cast<T*>(rt_scene_smem + park_at);
    if (RT_WF_PAIRS && !BVH) {
        ws->owner[lane] = 0u;
        __syncwarp();
    }

    for (;;) {
        // one warp = 32 consecutive work items (one 8x4 tile of pixels, or 32 neighbouring queue entries)
#if RT_WF_SYNC
        // CTA-wide lockstep: the whole CTA takes blockDim.x consecutive items and walks the phases together
        // (barrier per phase), so that the warps of an SM execute — and fetch — the same code at the same time
        __shared__ unsigned cta_first;
        __syncthreads();
        if (threadIdx.x == 0) cta_first = atomicAdd(&counts->work, blockDim.x);
        __syncthreads();
        const unsigned first = cta_first + (threadIdx.x & ~31u);
        if (cta_first >= n_items) break;
        const unsigned item = first + lane;
#else
        unsigned first = 0;
        if (lane == 0) first = atomicAdd(&counts->work, 32u);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= n_items) break;
        const unsigned item = first + lane;
#endif

        // ---- the radiance ray of this work item -------------------------------------------------
        // The kernel is register-bound (occupancy) and a node's state would be live across every trace of the node,
        // so the state lives in SHARED memory, one column per thread (PK(i)): each phase loads what it needs before
        // its trace and again after it.  P is the ray's origin until the hit is known and the hit POINT afterwards, D
        // the ray's direction; over / under points, the eye and reflect vectors and the refracted direction are
        // recomputed from (P, D, normal) where they are used — the same expressions as before, hence the same bits.
#define PK(i) park[(i) * RT_WF_THREADS]
        enum { PK_SURFACE = 0, PK_BASE = 3, PK_P = 6, PK_D = 9, PK_NORMAL = 12, PK_T_HIT = 15, PK_REFLECTANCE = 16, PK_REFR_A = 17, PK_REFR_N = 18 };
        static_assert(PK_REFR_N < WF_PARK_REALS, "WF_PARK_REALS too small");
        bool active = item < n_items;
        unsigned out_index = 0;
        int hit_pos = -1;      // deeper levels: the hit their parent's launch found
        unsigned q_index = 0;  // deeper levels: where the item sits in the input queue
        if (level == 0) {
            if (active) {
                const uint32_t tile = item / (TILE_W * TILE_H), in = item % (TILE_W * TILE_H);
                const uint32_t x = (tile % tiles_x) * TILE_W + in % TILE_W;
                const uint32_t k = (tile / tiles_x) * TILE_H + in / TILE_W;
                active = x < cam.hsize && k < cam.n_rows;
                if (active) {
                    const uint32_t y = image_row(cam, k);
                    const Ray<T> cr = camera_ray(cam, x, y);
                    const V3<T> origin = cr.o, d = cam.probe_ray ? mk<T>(cam.inv[0], cam.inv[1], cam.inv[2]) : cr.d;
                    PK(PK_P + 0) = origin.x; PK(PK_P + 1) = origin.y; PK(PK_P + 2) = origin.z;
                    PK(PK_D + 0) = d.x; PK(PK_D + 1) = d.y; PK(PK_D + 2) = d.z;
                    out_index = (cam.out_full_frame ? y : k) * cam.hsize + x;
                }
            }
        } else if (active && (item < n_front || item >= front_padded)) {
            // the parent's launch already traced this ray and only queued it because it hit
            q_index = item < n_front ? item : cap_rays - 1u - (item - front_padded);
            RT_CHECK(q_index < cap_rays);
            if (perm_in) q_index = perm_in[q_index];
            RT_CHECK(q_index < cap_rays);
            const WfRay<T>& r = rays_in[q_index];
            PK(PK_P + 0) = r.ox; PK(PK_P + 1) = r.oy; PK(PK_P + 2) = r.oz;
            PK(PK_D + 0) = r.dx; PK(PK_D + 1) = r.dy; PK(PK_D + 2) = r.dz;
            PK(PK_T_HIT) = r.t;
            hit_pos = r.pos;
        } else {
            active = false;  // padding between the two ends of the queue
        }

        // per-node state, filled phase by phase
        bool alive = false, need_containers = false;
        unsigned node_index = 0;
        PK(PK_SURFACE + 0) = T(0); PK(PK_SURFACE + 1) = T(0); PK(PK_SURFACE + 2) = T(0);  // Color::BLACK
        int hit_material = 0, flags = 0;
        const bool is_primary = level == 0 && active;

        // Every query of a node goes through ONE copy of the intersection code: the phases below only differ in
        // the ray they set up before it and in what they do with the answer after it.  (Five inlined copies
        // made this kernel instruction-fetch bound: GPC instruction cache at 90 % of its request rate.)
        //   phase 0        the radiance ray itself (level 0 only; deeper levels were traced by their parent)
        //   phase 1        refraction containers, same ray (intersection.rs:33-62)
        //   phase 2..1+L   one shadow ray per light (world.rs:43-53)
        //   phase 2+L, 3+L the reflected / refracted child rays (world.rs:114-157)
        const int n_phases = 4 + n_lights;
#pragma unroll 1
        for (int phase = 0; phase < n_phases; ++phase) {
#if RT_WF_SYNC
            __syncthreads();
#endif
            Ray<T> tray;
            tray.o = mk<T>(T(0), T(0), T(0));
            tray.d = mk<T>(T(0), T(0), T(1));
            int mode = MODE_IDLE;
            T seed = Real<T>::max();
            const int light = phase - 2;
            const int child = phase - 2 - n_lights;
            bool spawn = false;
            if (phase <= 1) {
                // phase 0: the camera ray; phase 1: the same ray again for the containers (P is still its origin)
                if (phase == 0 ? (level == 0 && active) : need_containers) {
                    tray.o = mk<T>(PK(PK_P + 0), PK(PK_P + 1), PK(PK_P + 2));
                    tray.d = mk<T>(PK(PK_D + 0), PK(PK_D + 1), PK(PK_D + 2));
                    mode = phase == 0 ? MODE_RADIANCE : MODE_CONTAINER;
                }
            } else {
                spawn = alive && (child < 0 || (flags & (child == 0 ? FR_REFLECT : FR_REFRACT)));
                if (spawn) {
                    const V3<T> P = mk<T>(PK(PK_P + 0), PK(PK_P + 1), PK(PK_P + 2));
                    const V3<T> normal = mk<T>(PK(PK_NORMAL + 0), PK(PK_NORMAL + 1), PK(PK_NORMAL + 2));
                    const V3<T> off = normal * Real<T>::offset(P.x, P.y, P.z, PK(PK_T_HIT));  // computed_hit.rs:33-34
                    if (child < 0) {  // World::is_in_shadow, world.rs:98-112
                        const V3<T> over = P + off;
                        Normalized<T> nl = normalize_full(ld3(sv.light((uint32_t)light)) - over);
                        tray.o = over;
                        tray.d = nl.v;
                        seed = nl.magnitude;
                        mode = MODE_SHADOW;
                    } else {
                        const V3<T> D = mk<T>(PK(PK_D + 0), PK(PK_D + 1), PK(PK_D + 2));
                        if (child == 0) {  // world.rs:124, intersection.rs:31
                            tray.o = P + off;
                            tray.d = reflect(D, normal);
                        } else {  // world.rs:150-152
                            tray.o = P - off;
                            tray.d = (normal * PK(PK_REFR_A)) - (neg(D) * PK(PK_REFR_N));
                        }
                        mode = MODE_RADIANCE;
                    }
                }
            }

            TraceAcc<T> acc;
            ContainerAcc<T> cacc;
            acc.c = RT_ACC_SPLIT ? &cacc : acc_store(acc);
            wf_reset_acc(acc, mode, tray, seed);
            if (phase == 1 && need_containers) {
                acc.c->t_hit = PK(PK_T_HIT);
                acc.c->hit_class = sv.shape_meta((uint32_t)hit_pos).w;
            }
            if (__any_sync(0xffffffffu, mode != MODE_IDLE)) wf_trace<T, FULL, BVH>(sv, tray, acc, ws);
            asm volatile("" ::: "memory");  // parked state is loaded again below, not carried in registers across the trace

            if (phase == 0) {
                // ---- World::internal_color_at (world.rs:70-86): the hit, the node, prepare_computations ----
                if (level == 0) {
                    hit_pos = acc.best_pos;
                    PK(PK_T_HIT) = acc.best_t;
                }
                const bool hit = active && hit_pos >= 0;
                if (level == 0 && active && !hit) wf_store_pixel(out_rgb, out_rgb8, (size_t)out_index, mk<T>(T(0), T(0), T(0)));  // World::DEFAULT_COLOR
                alive = hit;
                if (alive) {  // intersection.rs:21-31
                    const T* g = sv.shape((uint32_t)hit_pos);
                    const int4 meta = sv.shape_meta((uint32_t)hit_pos);
                    hit_material = meta.y;
                    const V3<T> P = mk<T>(PK(PK_P + 0), PK(PK_P + 1), PK(PK_P + 2)), D = mk<T>(PK(PK_D + 0), PK(PK_D + 1), PK(PK_D + 2));
                    V3<T> point = P + D * PK(PK_T_HIT);
                    V3<T> normal = world_normal_at(sv, (uint32_t)hit_pos, (meta.z >> FLAG_TYPE_SHIFT) & 7, g, point);
                    if (dot(normal, neg(D)) < T(0)) normal = neg(normal);
                    PK(PK_NORMAL + 0) = normal.x; PK(PK_NORMAL + 1) = normal.y; PK(PK_NORMAL + 2) = normal.z;
                    // n1 / n2 only feed refracted_color and Schlick, both irrelevant without iterations left
                    need_containers = sv.material((uint32_t)hit_material)[MAT_TRANSPARENCY] != T(0) && remaining > 0;
                }
            } else if (phase == 1) {
                // ---- n1 / n2 (intersection.rs:39-58), children to spawn, Schlick, base colour --------------
                T n1 = T(1), n2 = T(1);  // Material::DEFAULT_REFRACTIVE_INDEX
                if (need_containers) {
                    n1 = (acc.c->all_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)acc.c->all_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
                    if (acc.c->hit_class_inside)
                        n2 = (acc.c->excl_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)acc.c->excl_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
                    else
                        n2 = sv.material((uint32_t)hit_material)[MAT_REFRACTIVE_INDEX];
                }
                if (alive) {
                    const T* m = sv.material((uint32_t)hit_material);
                    const V3<T> D = mk<T>(PK(PK_D + 0), PK(PK_D + 1), PK(PK_D + 2));
                    const V3<T> normal = mk<T>(PK(PK_NORMAL + 0), PK(PK_NORMAL + 1), PK(PK_NORMAL + 2));
                    // from here on P is the hit point (the container query above was the last user of the origin)
                    const V3<T> P = mk<T>(PK(PK_P + 0), PK(PK_P + 1), PK(PK_P + 2)) + D * PK(PK_T_HIT);
                    PK(PK_P + 0) = P.x; PK(PK_P + 1) = P.y; PK(PK_P + 2) = P.z;
                    const V3<T> eye = neg(D);
                    if (remaining > 0 && m[MAT_REFLECTIVENESS] != T(0)) flags |= FR_REFLECT;  // world.rs:120
                    const T cos_i = dot(eye, normal);
                    if (remaining > 0 && m[MAT_TRANSPARENCY] != T(0)) {  // world.rs:136-154
                        T a, n_ratio;
                        if (refraction_coefficients(n1, n2, cos_i, a, n_ratio)) {
                            PK(PK_REFR_A) = a;  // refracted direction = normal * a - eye * n_ratio (world.rs:150)
                            PK(PK_REFR_N) = n_ratio;
                            flags |= FR_REFRACT;
                        }
                    }
                    if (m[MAT_REFLECTIVENESS] > T(0) && m[MAT_TRANSPARENCY] > T(0)) {  // world.rs:59
                        flags |= FR_SCHLICK;
                        PK(PK_REFLECTANCE) = schlick_reflectance(n1, n2, cos_i);
                    }
                    const V3<T> base = resolve_color(sv, (uint32_t)hit_material, (uint32_t)hit_pos, P + (normal * Real<T>::offset(P.x, P.y, P.z, PK(PK_T_HIT))));
                    PK(PK_BASE + 0) = base.x; PK(PK_BASE + 1) = base.y; PK(PK_BASE + 2) = base.z;
                }
            } else if (child < 0) {
                // ---- Material::lighting, material.rs:53-114, at over_point (material.rs:116-130) ------------
                if (alive) {
                    const T* lt = sv.light((uint32_t)light);
                    const T* m = sv.material((uint32_t)hit_material);
                    const V3<T> base = mk<T>(PK(PK_BASE + 0), PK(PK_BASE + 1), PK(PK_BASE + 2));
                    const V3<T> normal = mk<T>(PK(PK_NORMAL + 0), PK(PK_NORMAL + 1), PK(PK_NORMAL + 2));
                    const V3<T> eye = neg(mk<T>(PK(PK_D + 0), PK(PK_D + 1), PK(PK_D + 2)));
                    // the shadow ray's direction is normalized(light.position - over_point): the light vector of material.rs:88
                    const V3<T> lit = phong_lighting(m, base, ld3(lt + 3), tray.d, eye, normal, acc.best_pos >= 0);
                    // fold(Color::BLACK, Color::add)
                    PK(PK_SURFACE + 0) += lit.x; PK(PK_SURFACE + 1) += lit.y; PK(PK_SURFACE + 2) += lit.z;
                }
            } else {
                // ---- a child that hit something becomes a work item of the next level -----------------------
                const bool child_hit = spawn && acc.best_pos >= 0;  // (spawn: this lane traced a child in this phase)
                // which end of the next queue: does the child's hit need the container / refraction phases?
                const bool glass = child_hit && sv.material((uint32_t)sv.shape_meta((uint32_t)acc.best_pos).y)[MAT_TRANSPARENCY] != T(0);
                const unsigned queued_front = __ballot_sync(0xffffffffu, child_hit && !glass);
                const unsigned queued_back = __ballot_sync(0xffffffffu, glass);
                if (queued_front | queued_back) {
                    unsigned fbase = 0, bbase = 0;
                    if (lane == 0) {
                        if (queued_front) fbase = atomicAdd(&counts->n_rays[level + 1], (unsigned)__popc(queued_front));
                        if (queued_back) bbase = atomicAdd(&counts->n_back[level + 1], (unsigned)__popc(queued_back));
                    }
                    fbase = __shfl_sync(0xffffffffu, fbase, 0);
                    bbase = __shfl_sync(0xffffffffu, bbase, 0);
                    unsigned bin = 0, rank = 0;
                    if (keys_out) {
                        // rank within (end, bin): one atomic per distinct bin among the warp's appending lanes
                        bin = child_hit ? (min((unsigned)acc.best_pos, (unsigned)(RT_WF_BINS / 2 - 1)) << 1) | (unsigned)child : 0u;
                        const unsigned peers = __match_any_sync(0xffffffffu, child_hit ? (glass ? 0x100u : 0u) | bin : 0xffffffffu);
                        const int leader = __ffs((int)peers) - 1;
                        unsigned base = 0;
                        if (child_hit && (int)lane == leader) base = atomicAdd(&counts->bins[level + 1][glass ? 1 : 0][bin], (unsigned)__popc(peers));
                        base = __shfl_sync(0xffffffffu, base, leader);
                        rank = base + __popc(peers & ((1u << lane) - 1u));
                    }
                    if (child_hit) {
                        const unsigned below = (1u << lane) - 1u;
                        const unsigned f = fbase + __popc(queued_front & below), bk = bbase + __popc(queued_back & below);
                        // both ends grow towards each other; the last CTA of the launch compares their sum with the capacity
                        // once the launch is over (an overlap garbles entries of a frame that is re-rendered anyway)
                        const unsigned q = glass ? cap_rays - 1u - bk : f;
                        if ((glass ? bk : f) < cap_rays) {
                            WfRay<T> r;
                            r.ox = tray.o.x; r.oy = tray.o.y; r.oz = tray.o.z;
                            r.dx = tray.d.x; r.dy = tray.d.y; r.dz = tray.d.z;
                            r.t = acc.best_t;
                            r.k = sv.material((uint32_t)hit_material)[child == 0 ? MAT_REFLECTIVENESS : MAT_TRANSPARENCY];
                            r.pos = acc.best_pos;
                            r.parent = (int)node_index;
                            r.slot = child;
                            r.pad = 0;
                            RT_CHECK(q < cap_rays && r.pos >= 0 && (uint32_t)r.pos < sv.L.n_shapes && r.parent >= 0 && (unsigned)r.parent < cap_nodes);
                            rays_out[q] = r;
                            if (keys_out) keys_out[q] = ((unsigned long long)bin << 32) | rank;
                        } else {
                            counts->overflow = 1u;
                        }
                    }
                }
            }

            if (phase == 1 + n_lights) {
                // ---- the light loop is over: a node without children is finished, the others get a record ----
                const bool interior = alive && (flags & (FR_REFLECT | FR_REFRACT));
                const unsigned records = __ballot_sync(0xffffffffu, interior);
                if (records) {
                    unsigned node_base = 0;
                    if (lane == 0) node_base = atomicAdd(&counts->n_nodes, (unsigned)__popc(records));
                    node_base = __shfl_sync(0xffffffffu, node_base, 0);
                    node_index = node_base + __popc(records & ((1u << lane) - 1u));
                }
                const V3<T> surface = mk<T>(PK(PK_SURFACE + 0), PK(PK_SURFACE + 1), PK(PK_SURFACE + 2));
                // whose child this node is: read back from the queue entry (level 0: a pixel's root)
                int parent = -1, slot = 0;
                T k_parent = T(1);
                if (level > 0 && alive) {
                    const WfRay<T>& r = rays_in[q_index];
                    k_parent = r.k;
                    parent = r.parent;
                    slot = r.slot;
                }
                if (interior && node_index >= cap_nodes) {
                    counts->overflow = 1u;
                    flags &= ~(FR_REFLECT | FR_REFRACT);  // no record, no children: the frame is re-rendered anyway
                } else if (interior) {
                    WfNode<T> nd;
                    nd.link = parent >= 0 ? parent : ~(int)out_index;
                    nd.slot_flags = slot | ((flags & FR_SCHLICK) ? 2 : 0);
                    nd.k_parent = k_parent;
                    nd.reflectance = (flags & FR_SCHLICK) ? PK(PK_REFLECTANCE) : T(0);
                    nd.surface[0] = surface.x; nd.surface[1] = surface.y; nd.surface[2] = surface.z;
                    nd.reflected[0] = nd.reflected[1] = nd.reflected[2] = T(0);  // world.rs:121
                    nd.refracted[0] = nd.refracted[1] = nd.refracted[2] = T(0);  // world.rs:137,146
                    RT_CHECK(node_index < cap_nodes && (parent < 0 || (unsigned)parent < cap_nodes));
                    nodes[node_index] = nd;
                } else if (alive) {
                    // world.rs:59-66 with black children: (surface + 0) + 0 — and 0 * reflectance is 0 too — is `surface`
                    if (parent < 0) {
                        wf_store_pixel(out_rgb, out_rgb8, (size_t)out_index, surface);
                    } else {
                        const V3<T> c = surface * k_parent;  // world.rs:127 / 156
                        T* dst = slot == 0 ? nodes[parent].reflected : nodes[parent].refracted;
                        dst[0] = c.x; dst[1] = c.y; dst[2] = c.z;
                    }
                }
            }
        }
        // ---- the item's contribution to the work counters ----
        {
            const unsigned b_primary = __ballot_sync(0xffffffffu, is_primary), b_nodes = __ballot_sync(0xffffffffu, alive);
            const unsigned b_reflect = __ballot_sync(0xffffffffu, alive && (flags & FR_REFLECT));
            const unsigned b_refract = __ballot_sync(0xffffffffu, alive && (flags & FR_REFRACT));
            if (lane == 0) {
                my_count[0] += (unsigned)__popc(b_primary);
                my_count[1] += (unsigned)__popc(b_reflect);
                my_count[2] += (unsigned)__popc(b_refract);
                my_count[3] += (unsigned)__popc(b_nodes);
            }
        }
    }

    if (counters && lane == 0) {
        __syncwarp(1u);
        const unsigned c_primary = my_count[0], c_reflect = my_count[1], c_refract = my_count[2], c_nodes = my_count[3];
        if (c_primary) {
            atomicAdd(&counters[COUNTER_PRIMARY], (unsigned long long)c_primary);
            atomicAdd(&counters[COUNTER_PIXELS], (unsigned long long)c_primary);
        }
        if (c_nodes) {
            atomicAdd(&counters[COUNTER_HIT_NODES], (unsigned long long)c_nodes);
            atomicAdd(&counters[COUNTER_SHADOW], (unsigned long long)c_nodes * (unsigned long long)n_lights);  // one per light and node, world.rs:43-53
        }
        if (c_reflect) atomicAdd(&counters[COUNTER_REFLECT], (unsigned long long)c_reflect);
        if (c_refract) atomicAdd(&counters[COUNTER_REFRACT], (unsigned long long)c_refract);
    }

#undef PK
    // The last CTA to finish closes the level (no separate launch): remember where this level's nodes end, reset
    // the chunk cursor for the next launch, and check that the two ends of the next queue did not meet.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&counts->done, 1u) == gridDim.x - 1u) {
            __threadfence();
            volatile WfCounts* c = counts;
            c->node_end[level] = c->n_nodes;
            c->work = 0u;
            c->done = 0u;
            if ((unsigned long long)c->n_rays[level + 1] + c->n_back[level + 1] > cap_rays) c->overflow = 1u;
        }
    }
}

// World::shade_hit's tail (world.rs:59-66) for the nodes of one level, deepest level first.  Memory-latency bound: a
// thread fetches its whole record with 128-bit loads (six per f64 record, all in flight before the first use) and
// handles two records per iteration.
template <typename T>
RT_DEV WfNode<T> wf_load_node(const WfNode<T>* p) {
    static_assert(sizeof(WfNode<T>) % 16 == 0, "node records are read with 128-bit loads");
    union {
        uint4 raw[sizeof(WfNode<T>) / 16];
        WfNode<T> node;
    } u;
    const uint4* src = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (unsigned k = 0; k < sizeof(WfNode<T>) / 16; ++k) u.raw[k] = src[k];
    return u.node;
}

template <typename T>
RT_DEV void wf_combine_node(WfNode<T>* __restrict__ nodes, const WfNode<T>& n, T* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8) {
    const V3<T> surface = mk<T>(n.surface[0], n.surface[1], n.surface[2]);
    const V3<T> reflected = mk<T>(n.reflected[0], n.reflected[1], n.reflected[2]);
    const V3<T> refracted = mk<T>(n.refracted[0], n.refracted[1], n.refracted[2]);
    V3<T> colour;
    if (n.slot_flags & 2) colour = (surface + (reflected * n.reflectance)) + (refracted * (T(1) - n.reflectance));
    else colour = (surface + reflected) + refracted;
    if (n.link < 0) {
        wf_store_pixel(out_rgb, out_rgb8, (size_t)(unsigned)~n.link, colour);  // Camera::render_parallel, camera.rs:108
    } else {
        const V3<T> c = colour * n.k_parent;  // world.rs:127 / 156
        RT_CHECK(n.link >= 0);
        T* dst = (n.slot_flags & 1) == 0 ? nodes[n.link].reflected : nodes[n.link].refracted;
        dst[0] = c.x; dst[1] = c.y; dst[2] = c.z;
    }
}

// Between the launches of levels d-1 and d of a binned frame: turn the (bin, rank) keys that level d-1 recorded into
// the permutation level d consumes its queue through.  Slot s of the front end (s counted in bin order, arrival order
// inside a bin) is the entry perm[s]; slot s of the back end is perm[cap - 1 - s].  A frame whose queue overflowed is
// rendered again anyway: it gets the identity, so that no lane ever follows a stale index.
__global__ void __launch_bounds__(256) wf_bin_kernel(const WfCounts* __restrict__ counts, int level, unsigned cap_rays,
                                                     const unsigned long long* __restrict__ keys, unsigned* __restrict__ perm) {
    static_assert(RT_WF_BINS == 64, "the scan below is written for two warps per end");
    __shared__ unsigned base[2][RT_WF_BINS];
    __shared__ unsigned warp_total[4];
    wf_release_dependents();
    wf_wait_for_previous();
    unsigned n_front = counts->n_rays[level], n_back = counts->n_back[level];
    const bool overflowed = (unsigned long long)n_front + n_back > cap_rays;
    if (overflowed) {
        n_front = min(n_front, cap_rays);
        n_back = min(n_back, cap_rays - n_front);
    }
    // exclusive scan of the bin counts of both ends: threads 0..63 the front end, 64..127 the back end
    if (threadIdx.x < 2 * RT_WF_BINS) {
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        const unsigned mine = counts->bins[level][threadIdx.x / RT_WF_BINS][threadIdx.x % RT_WF_BINS];
        unsigned incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned up = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += up;
        }
        if (lane == 31u) warp_total[warp] = incl;
        __syncwarp();
        base[threadIdx.x / RT_WF_BINS][threadIdx.x % RT_WF_BINS] = incl - mine;  // the second warp of an end adds the first one's total below
    }
    __syncthreads();
    if (threadIdx.x < 2 * RT_WF_BINS && ((threadIdx.x >> 5) & 1u)) base[threadIdx.x / RT_WF_BINS][threadIdx.x % RT_WF_BINS] += warp_total[(threadIdx.x >> 5) - 1u];
    __syncthreads();
    const unsigned n = n_front + n_back, stride = gridDim.x * blockDim.x;
    for (unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4u * stride) {
        // four independent entries per thread and round: the loads are in flight together
        unsigned q[4];
        unsigned long long key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = i0 + (unsigned)u * stride;
            q[u] = i < n ? (i >= n_front ? cap_rays - 1u - (i - n_front) : i) : 0xffffffffu;
            key[u] = (q[u] != 0xffffffffu && !overflowed) ? keys[q[u]] : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (q[u] == 0xffffffffu) continue;
            if (overflowed) {
                perm[q[u]] = q[u];
                continue;
            }
            const bool back = i0 + (unsigned)u * stride >= n_front;
            const unsigned s = base[back ? 1 : 0][(unsigned)(key[u] >> 32) & (RT_WF_BINS - 1)] + (unsigned)key[u];
            RT_CHECK(q[u] < cap_rays && s < (back ? n_back : n_front));
            if (s < (back ? n_back : n_front)) perm[back ? cap_rays - 1u - s : s] = q[u];
        }
    }
}

template <typename T>
__global__ void wf_combine_kernel(WfNode<T>* __restrict__ nodes, const WfCounts* __restrict__ counts, int level, unsigned cap_nodes,
                                  T* __restrict__ out_rgb, uint8_t* __restrict__ out_rgb8, int release) {
    // release: another kernel of THIS library follows (it waits before it reads).  The last kernel of a frame never
    // releases early: whatever the caller launches next on the stream must see a finished frame.
    if (release) wf_release_dependents();
    wf_wait_for_previous();
    const unsigned begin = level == 0 ? 0u : min(counts->node_end[level - 1], cap_nodes);
    const unsigned end = min(counts->node_end[level], cap_nodes);
    const unsigned stride = gridDim.x * blockDim.x;
#ifndef RT_WF_COMBINE_ILP
#define RT_WF_COMBINE_ILP 4  // node records in flight per thread and round (the pass is bound by memory latency: 2 -> 4 and 4 -> 8 CTAs per SM: cover -1.1 %)
#endif
    for (unsigned i = begin + blockIdx.x * blockDim.x + threadIdx.x; i < end; i += RT_WF_COMBINE_ILP * stride) {
        WfNode<T> n[RT_WF_COMBINE_ILP];
#pragma unroll
        for (int u = 0; u < RT_WF_COMBINE_ILP; ++u)
            if (i + (unsigned)u * stride < end) n[u] = wf_load_node(nodes + i + (unsigned)u * stride);
#pragma unroll
        for (int u = 0; u < RT_WF_COMBINE_ILP; ++u)
            if (i + (unsigned)u * stride < end) wf_combine_node(nodes, n[u], out_rgb, out_rgb8);
    }
}

}  // namespace rt
