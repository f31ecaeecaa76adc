// rtgpu.cu — implementation of the C ABI in include/rtgpu.h (host side: scene packer, launcher,
// row-band sharding across devices, measurement helpers).  The kernels live in rt_kernel.cuh.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -O3 -shared (see build.py).
// -fmad=false is REQUIRED: the reference never contracts a*b+c (rt_kernel.cuh, header comment).

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <limits>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rtgpu.h"
#include "rt_bvh.h"
#include "rt_kernel.cuh"
#include "rt_wavefront.cuh"
#include "rt_probe.cuh"
#include "rt_scene.h"

namespace {

thread_local std::string g_last_error;

int fail(int status, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return status;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(_e == cudaErrorMemoryAllocation ? RTGPU_ERR_OUT_OF_MEMORY : RTGPU_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Scene packer: rtgpu_scene (world order, SoA) -> the two device blobs of rt_scene.h (type-sorted).

// resize() without value-initialisation: the packer zeroes (first-touches) the blobs of a large scene on several threads
template <typename T>
struct NoInitAllocator : std::allocator<T> {
    template <typename U>
    struct rebind {
        using other = NoInitAllocator<U>;
    };
    template <typename U>
    void construct(U* p) noexcept {
        ::new (static_cast<void*>(p)) U;
    }
    template <typename U, typename... Args>
    void construct(U* p, Args&&... args) {
        ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
    }
};

struct PackedScene {
    std::vector<double, NoInitAllocator<double>> reals;
    std::vector<int, NoInitAllocator<int>> ints;
    rt::SceneLayout layout;
    bool has_cyl_cone_tri = false;
    bool has_secondary = false;  // some material is reflective or transparent: rays beyond primary + shadow exist
    uint32_t secondary_shapes = 0;  // shapes whose material is reflective or transparent
    uint64_t fingerprint = 0;    // of the packed blobs (sampled for large scenes): the tuner's notion of "same scene"
};

int validate_scene(const rtgpu_scene* s) {
    if (!s) return fail(RTGPU_ERR_INVALID_ARGUMENT, "scene is NULL");
    if (s->abi_version != RTGPU_ABI_VERSION)
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "scene.abi_version %u != %u", s->abi_version, RTGPU_ABI_VERSION);
    const uint32_t S = s->n_shapes, M = s->n_materials, Q = s->n_patterns, L = s->n_lights, NT = s->n_triangles;
    if (S && (!s->shape_type || !s->shape_inv || !s->shape_min || !s->shape_max || !s->shape_closed ||
              !s->shape_material || !s->shape_eq_class))
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "a shape array is NULL");
    if (M && (!s->mat_color || !s->mat_params || !s->mat_casts_shadow || !s->mat_pattern))
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "a material array is NULL");
    if (Q && (!s->pat_type || !s->pat_color_a || !s->pat_color_b || !s->pat_inv || !s->pat_child_a || !s->pat_child_b))
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "a pattern array is NULL");
    if (L && (!s->light_position || !s->light_intensity)) return fail(RTGPU_ERR_INVALID_ARGUMENT, "a light array is NULL");
    if (NT && (!s->tri_vertex_1 || !s->tri_edge_1 || !s->tri_edge_2 || !s->tri_normal || !s->shape_triangle))
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "a triangle array is NULL");
    for (uint32_t i = 0; i < S; ++i) {
        if (s->shape_type[i] >= RTGPU_SHAPE_TYPE_COUNT) return fail(RTGPU_ERR_INVALID_ARGUMENT, "shape %u: type %u", i, s->shape_type[i]);
        if (s->shape_material[i] >= M) return fail(RTGPU_ERR_INVALID_ARGUMENT, "shape %u: material %u out of range", i, s->shape_material[i]);
        if (s->shape_eq_class[i] > i) return fail(RTGPU_ERR_INVALID_ARGUMENT, "shape %u: eq_class %u is not the lowest equal index", i, s->shape_eq_class[i]);
        if (s->shape_type[i] == RTGPU_TRIANGLE && (!s->shape_triangle || s->shape_triangle[i] < 0 || (uint32_t)s->shape_triangle[i] >= NT))
            return fail(RTGPU_ERR_INVALID_ARGUMENT, "shape %u: triangle index out of range", i);
    }
    for (uint32_t m = 0; m < M; ++m)
        if (s->mat_pattern[m] >= (int32_t)Q) return fail(RTGPU_ERR_INVALID_ARGUMENT, "material %u: pattern %d out of range", m, s->mat_pattern[m]);
    for (uint32_t q = 0; q < Q; ++q) {
        if (s->pat_type[q] >= RTGPU_PATTERN_TYPE_COUNT) return fail(RTGPU_ERR_INVALID_ARGUMENT, "pattern %u: type %u", q, s->pat_type[q]);
        if (s->pat_type[q] == RTGPU_PATTERN_COMPLEX) {
            // children are emitted before their parent by every flattener: guarantees termination on device
            if (s->pat_child_a[q] < 0 || s->pat_child_b[q] < 0 || (uint32_t)s->pat_child_a[q] >= q || (uint32_t)s->pat_child_b[q] >= q)
                return fail(RTGPU_ERR_INVALID_ARGUMENT, "pattern %u: complex children must precede their parent", q);
        }
    }
    return RTGPU_OK;
}

// ---- conservative world-space bounding spheres (rt_scene.h CULL_REALS) ---------------------------

// forward linear part M = A^-1 and translation p0 = -M t of the affine transform whose inverse is [A | t]
bool invert_affine(const double* inv, double M[9], double p0[3]) {
    const double a = inv[0], b = inv[1], c = inv[2], d = inv[4], e = inv[5], f = inv[6], g = inv[8], h = inv[9], k = inv[10];
    const double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    M[0] = (e * k - f * h) * id; M[1] = (c * h - b * k) * id; M[2] = (b * f - c * e) * id;
    M[3] = (f * g - d * k) * id; M[4] = (a * k - c * g) * id; M[5] = (c * d - a * f) * id;
    M[6] = (d * h - e * g) * id; M[7] = (b * g - a * h) * id; M[8] = (a * e - b * d) * id;
    const double t[3] = {inv[3], inv[7], inv[11]};
    for (int r = 0; r < 3; ++r) p0[r] = -(M[r * 3 + 0] * t[0] + M[r * 3 + 1] * t[1] + M[r * 3 + 2] * t[2]);
    for (int i = 0; i < 9; ++i)
        if (!std::isfinite(M[i])) return false;
    return std::isfinite(p0[0]) && std::isfinite(p0[1]) && std::isfinite(p0[2]);
}

void mat3_apply(const double M[9], const double v[3], double out[3]) {
    for (int r = 0; r < 3; ++r) out[r] = M[r * 3 + 0] * v[0] + M[r * 3 + 1] * v[1] + M[r * 3 + 2] * v[2];
}

// largest singular value of M (cyclic Jacobi on M^T M), rounded up
double sigma_max(const double M[9]) {
    double S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = M[0 * 3 + i] * M[0 * 3 + j] + M[1 * 3 + i] * M[1 * 3 + j] + M[2 * 3 + i] * M[2 * 3 + j];
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = std::fabs(S[0][1]) + std::fabs(S[0][2]) + std::fabs(S[1][2]);
        if (off <= 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (S[p][q] == 0.0) continue;
                const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int r = 0; r < 3; ++r) {  // S = S * J
                    const double srp = S[r][p], srq = S[r][q];
                    S[r][p] = c * srp - sn * srq;
                    S[r][q] = sn * srp + c * srq;
                }
                for (int r = 0; r < 3; ++r) {  // S = J^T * S
                    const double spr = S[p][r], sqr = S[q][r];
                    S[p][r] = c * spr - sn * sqr;
                    S[q][r] = sn * spr + c * sqr;
                }
            }
    }
    const double lam = std::max(S[0][0], std::max(S[1][1], S[2][2]));
    // Gershgorin slack for whatever off-diagonal mass is left, then a relative safety factor
    const double slack = std::fabs(S[0][1]) + std::fabs(S[0][2]) + std::fabs(S[1][2]);
    return std::sqrt(std::max(0.0, lam + slack)) * (1.0 + 1e-9);
}

void bounding_sphere(const rtgpu_scene* s, uint32_t i, double out[4]) {
    const double inf = std::numeric_limits<double>::infinity();
    out[0] = out[1] = out[2] = 0.0;
    out[3] = inf;  // unbounded: never culled
    const int type = s->shape_type[i];
    if (type == RTGPU_PLANE) return;
    double M[9], p0[3];
    if (!invert_affine(s->shape_inv + (size_t)i * 12, M, p0)) return;
    double centre_local[3] = {0, 0, 0};
    double radius = inf;
    if (type == RTGPU_SPHERE) {
        radius = sigma_max(M);
    } else if (type == RTGPU_CUBE) {
        radius = 0.0;
        for (int c = 0; c < 8; ++c) {
            const double v[3] = {(c & 1) ? 1.0 : -1.0, (c & 2) ? 1.0 : -1.0, (c & 4) ? 1.0 : -1.0};
            double w[3];
            mat3_apply(M, v, w);
            radius = std::max(radius, std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
        }
    } else if (type == RTGPU_CYLINDER || type == RTGPU_CONE) {
        const double mn = s->shape_min[i], mx = s->shape_max[i];
        if (!(std::isfinite(mn) && std::isfinite(mx)) || std::fabs(mn) > 1e150 || std::fabs(mx) > 1e150 || !(mn <= mx)) return;
        const double ym = 0.5 * (mn + mx);
        centre_local[1] = ym;
        double local_r;
        if (type == RTGPU_CYLINDER) {
            const double hgt = 0.5 * (mx - mn);
            local_r = std::sqrt(1.0 + hgt * hgt);
        } else {
            // cone surface: radius |y| at height y; caps included
            const double r_lo = std::sqrt(mn * mn + (mn - ym) * (mn - ym)), r_hi = std::sqrt(mx * mx + (mx - ym) * (mx - ym));
            local_r = std::max(r_lo, r_hi);
        }
        radius = sigma_max(M) * local_r;
    } else if (type == RTGPU_TRIANGLE) {
        const size_t t = (size_t)s->shape_triangle[i] * 3;
        double v[3][3];
        for (int k = 0; k < 3; ++k) {
            v[0][k] = s->tri_vertex_1[t + k];
            v[1][k] = s->tri_vertex_1[t + k] + s->tri_edge_1[t + k];
            v[2][k] = s->tri_vertex_1[t + k] + s->tri_edge_2[t + k];
            centre_local[k] = (v[0][k] + v[1][k] + v[2][k]) / 3.0;
        }
        radius = 0.0;
        for (int j = 0; j < 3; ++j) {
            const double dl[3] = {v[j][0] - centre_local[0], v[j][1] - centre_local[1], v[j][2] - centre_local[2]};
            double w[3];
            mat3_apply(M, dl, w);
            radius = std::max(radius, std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]));
        }
    }
    double cw[3];
    mat3_apply(M, centre_local, cw);
    const double cx = p0[0] + cw[0], cy = p0[1] + cw[1], cz = p0[2] + cw[2];
    // inflate: relative 1e-6 plus an absolute term for the rounding of the centre itself
    const double scale = std::fabs(cx) + std::fabs(cy) + std::fabs(cz) + radius;
    const double r = radius * (1.0 + 1e-6) + scale * 1e-12;
    if (!(std::isfinite(r) && std::isfinite(cx) && std::isfinite(cy) && std::isfinite(cz)) || r * r > 1e300) return;
    out[0] = cx;
    out[1] = cy;
    out[2] = cz;
    out[3] = r * r;
}

// Conservative world-space box of a bounded shape; false = unbounded (planes, untruncated cylinders / cones,
// degenerate transforms): such shapes stay in the flat always-tested list.
bool bounding_box(const rtgpu_scene* s, uint32_t i, rt::Aabb* out) {
    const int type = s->shape_type[i];
    if (type == RTGPU_PLANE) return false;
    double M[9], p0[3];
    if (!invert_affine(s->shape_inv + (size_t)i * 12, M, p0)) return false;
    rt::Aabb box;
    box.reset();
    auto add_local = [&](const double v[3]) {
        double w[3];
        mat3_apply(M, v, w);
        const double p[3] = {p0[0] + w[0], p0[1] + w[1], p0[2] + w[2]};
        box.grow_point(p);
    };
    if (type == RTGPU_SPHERE) {
        // the extent of M * (unit sphere) along axis k is the norm of row k of M
        for (int k = 0; k < 3; ++k) {
            const double e = std::sqrt(M[k * 3] * M[k * 3] + M[k * 3 + 1] * M[k * 3 + 1] + M[k * 3 + 2] * M[k * 3 + 2]);
            box.lo[k] = p0[k] - e;
            box.hi[k] = p0[k] + e;
        }
    } else if (type == RTGPU_CUBE || type == RTGPU_CYLINDER || type == RTGPU_CONE) {
        double ylo = -1.0, yhi = 1.0, r = 1.0;
        if (type != RTGPU_CUBE) {
            ylo = s->shape_min[i];
            yhi = s->shape_max[i];
            if (!(std::isfinite(ylo) && std::isfinite(yhi)) || std::fabs(ylo) > 1e150 || std::fabs(yhi) > 1e150 || !(ylo <= yhi)) return false;
            if (type == RTGPU_CONE) r = std::max(std::fabs(ylo), std::fabs(yhi));  // radius |y| at height y
        }
        for (int c = 0; c < 8; ++c) {
            const double v[3] = {(c & 1) ? r : -r, (c & 2) ? yhi : ylo, (c & 4) ? r : -r};
            add_local(v);
        }
    } else {  // triangle
        const size_t t = (size_t)s->shape_triangle[i] * 3;
        double v[3][3];
        for (int k = 0; k < 3; ++k) {
            v[0][k] = s->tri_vertex_1[t + k];
            v[1][k] = s->tri_vertex_1[t + k] + s->tri_edge_1[t + k];
            v[2][k] = s->tri_vertex_1[t + k] + s->tri_edge_2[t + k];
        }
        for (int j = 0; j < 3; ++j) add_local(v[j]);
    }
    // inflate: relative to the box size and to the magnitude of its coordinates (>> f64 rounding of the slab test)
    double size = 0.0, mag = 0.0;
    for (int k = 0; k < 3; ++k) {
        size = std::max(size, box.hi[k] - box.lo[k]);
        mag = std::max(mag, std::max(std::fabs(box.lo[k]), std::fabs(box.hi[k])));
    }
    const double pad = size * 1e-6 + mag * 1e-9 + 1e-300;
    for (int k = 0; k < 3; ++k) {
        box.lo[k] -= pad;
        box.hi[k] += pad;
        if (!std::isfinite(box.lo[k]) || !std::isfinite(box.hi[k])) return false;
    }
    *out = box;
    return true;
}

// Scenes with at least this many bounded shapes are traversed through a BVH (RTGPU_BVH_MIN overrides;
// 0 disables the hierarchy).
uint32_t bvh_threshold() {
    const char* e = getenv("RTGPU_BVH_MIN");
    if (e && *e) return (uint32_t)strtoul(e, nullptr, 10);
    return 32;
}

double wall_ms();

// fn(begin, end, worker) over [0, n) in contiguous chunks on several threads (the calling one included) when there is
// enough to share out; the packer's per-shape and per-node loops are independent iterations writing disjoint records.
template <typename F>
void parallel_for(uint32_t n, uint32_t min_chunk, F&& fn) {
    unsigned workers = std::max(1u, std::thread::hardware_concurrency());
    workers = std::min<unsigned>(workers, std::max<uint32_t>(1u, n / std::max(1u, min_chunk)));
    workers = std::min(workers, 64u);
    if (workers <= 1) {
        fn(0u, n, 0u);
        return;
    }
    const uint32_t chunk = (n + workers - 1) / workers;
    std::vector<std::thread> pool;
    for (unsigned w = 1; w < workers; ++w) pool.emplace_back([&fn, w, chunk, n] { fn(std::min(n, w * chunk), std::min(n, (w + 1) * chunk), w); });
    fn(0u, std::min(n, chunk), 0u);
    for (auto& t : pool) t.join();
}

int pack_scene(const rtgpu_scene* s, PackedScene* out) {
    // RTGPU_PACK_TRACE=1: where the host time of packing a (large) scene goes, on stderr
    const bool pack_trace = getenv("RTGPU_PACK_TRACE") != nullptr;
    double t_mark = pack_trace ? wall_ms() : 0.0;
    auto mark = [&](const char* what) {
        if (!pack_trace) return;
        const double now = wall_ms();
        fprintf(stderr, "[rtgpu] pack: %-34s %8.2f ms\n", what, now - t_mark);
        t_mark = now;
    };
    int st = validate_scene(s);
    if (st != RTGPU_OK) return st;
    mark("validate");
    const uint32_t S = s->n_shapes, M = s->n_materials, Q = s->n_patterns, L = s->n_lights;
    rt::SceneLayout& lay = out->layout;
    memset(&lay, 0, sizeof(lay));
    lay.n_shapes = S;
    lay.cull_shrink64 = 1.0 - 1.0e-9;
    lay.cull_shrink32 = 1.0f - 1.0e-3f;
    lay.n_materials = M;
    lay.n_patterns = Q;
    lay.n_lights = L;

    // bounded shapes may go into a BVH; unbounded ones always stay in the flat per-type lists
    std::vector<rt::Aabb> boxes(S);
    std::vector<uint8_t> bounded(S, 0);
    uint32_t n_bounded = 0;
    parallel_for(S, 1u << 14, [&](uint32_t begin, uint32_t end, unsigned) {
        for (uint32_t i = begin; i < end; ++i) bounded[i] = bounding_box(s, i, &boxes[i]) ? 1 : 0;
    });
    for (uint32_t i = 0; i < S; ++i) n_bounded += bounded[i];
    mark("bounding boxes");
    const uint32_t threshold = bvh_threshold();
    // (a hierarchy over a single shape would be one half-empty node whose empty box no slab test rejects: keep it flat)
    const bool use_bvh = threshold > 0 && n_bounded >= std::max(threshold, 2u);
    // flat part: stable grouping by type, world order kept inside a type (and carried as `orig` for tie-breaks)
    std::vector<uint32_t> order;
    order.reserve(S);
    for (int t = 0; t < rt::NUM_SHAPE_TYPES; ++t) {
        lay.type_begin[t] = (uint32_t)order.size();
        for (uint32_t i = 0; i < S; ++i)
            if (s->shape_type[i] == t && !(use_bvh && bounded[i])) order.push_back(i);
    }
    lay.type_begin[rt::NUM_SHAPE_TYPES] = (uint32_t)order.size();
    // BVH part: bounded shapes in depth-first leaf order
    rt::Bvh bvh;
    const uint32_t n_flat = (uint32_t)order.size();
    if (use_bvh) {
        std::vector<uint32_t> items;
        std::vector<rt::Aabb> item_boxes;
        items.reserve(n_bounded);
        item_boxes.reserve(n_bounded);
        for (uint32_t i = 0; i < S; ++i)
            if (bounded[i]) {
                items.push_back(i);
                item_boxes.push_back(boxes[i]);
            }
        mark("order + item boxes");
        bvh = rt::build_bvh(item_boxes, rt::BVH_MAX_DEPTH);
        mark("BVH build");
        if (bvh.max_depth >= rt::BVH_MAX_DEPTH) return fail(RTGPU_ERR_UNSUPPORTED, "BVH depth %d exceeds the device stack", bvh.max_depth);
        for (uint32_t k = 0; k < bvh.leaf_order.size(); ++k) order.push_back(items[bvh.leaf_order[k]]);
    }
    uint32_t n_tri = 0;
    for (uint32_t i = 0; i < S; ++i) {
        n_tri += s->shape_type[i] == RTGPU_TRIANGLE ? 1u : 0u;
        if (s->shape_type[i] >= RTGPU_CYLINDER) out->has_cyl_cone_tri = true;
    }
    for (uint32_t m = 0; m < M; ++m) {
        const double* prm = s->mat_params + (size_t)m * RTGPU_MAT_PARAM_COUNT;
        if (prm[4] > 0.0 || prm[5] > 0.0) out->has_secondary = true;  // reflective, transparency (world.rs:121, 137)
    }
    for (uint32_t i = 0; i < S; ++i) {
        const double* prm = s->mat_params + (size_t)s->shape_material[i] * RTGPU_MAT_PARAM_COUNT;
        if (prm[4] > 0.0 || prm[5] > 0.0) out->secondary_shapes++;
    }
    const uint32_t n_nodes = (uint32_t)bvh.nodes.size();
    lay.n_bvh_nodes = use_bvh ? std::max(n_nodes, 1u) : 0u;  // a single bounded shape still gets one (half-empty) node
    lay.bvh_root = 0;

    // value-equal classes: size parity and highest member (rt_scene.h FLAG_CONTAINER_REP)
    std::vector<uint32_t> class_size(S, 0), class_last(S, 0);
    for (uint32_t i = 0; i < S; ++i) {
        class_size[s->shape_eq_class[i]]++;
        class_last[s->shape_eq_class[i]] = i;
    }

    lay.tri_off = S * rt::SHAPE_REALS;
    lay.mat_off = lay.tri_off + n_tri * rt::TRI_REALS;
    lay.pat_off = lay.mat_off + M * rt::MAT_REALS;
    lay.light_off = lay.pat_off + Q * rt::PAT_REALS;
    lay.cull_off = (lay.light_off + L * rt::LIGHT_REALS + 3u) & ~3u;  // 16-byte aligned records in f32 too
    lay.bvh_off = lay.cull_off + S * rt::CULL_REALS;
    lay.n_reals = lay.bvh_off + lay.n_bvh_nodes * rt::BVH_REALS;
    lay.n_reals = (lay.n_reals + 3u) & ~3u;  // 16-byte multiple in f32 too (bulk copies into shared memory)
    lay.mat_meta_off = S * rt::SHAPE_INTS;
    lay.pat_meta_off = lay.mat_meta_off + M * rt::MAT_INTS;
    lay.bvh_meta_off = (lay.pat_meta_off + Q * rt::PAT_INTS + 1u) & ~1u;
    lay.bvh32_off = (lay.bvh_meta_off + lay.n_bvh_nodes * rt::BVH_INTS + 3u) & ~3u;
    lay.cull32_off = (lay.bvh32_off + lay.n_bvh_nodes * rt::BVH32_WORDS + 3u) & ~3u;
    lay.n_ints = lay.cull32_off + S * 4u;
    lay.n_ints = (lay.n_ints + 3u) & ~3u;
    if ((uint64_t)S * rt::SHAPE_REALS + (uint64_t)n_tri * rt::TRI_REALS + (uint64_t)S * rt::CULL_REALS + (uint64_t)lay.n_bvh_nodes * rt::BVH_REALS > 0xF0000000ull)
        return fail(RTGPU_ERR_UNSUPPORTED, "scene too large for 32-bit blob offsets (%u shapes)", S);

    out->reals.clear();
    out->ints.clear();
    out->reals.resize(lay.n_reals);  // no initialisation (NoInitAllocator): zeroed below, in parallel
    out->ints.resize(lay.n_ints);
    double* R = out->reals.data();
    int* I = out->ints.data();
    parallel_for(lay.n_reals, 1u << 18, [&](uint32_t b, uint32_t e, unsigned) { memset(R + b, 0, (size_t)(e - b) * sizeof(double)); });
    parallel_for(lay.n_ints, 1u << 19, [&](uint32_t b, uint32_t e, unsigned) { memset(I + b, 0, (size_t)(e - b) * sizeof(int)); });

    mark("layout + blob allocation");
    // triangles get their slots in sorted order: a prefix count, so that the loop below has independent iterations
    std::vector<uint32_t> tri_slot_at;
    if (n_tri) {
        tri_slot_at.resize(S);
        uint32_t running = 0;
        for (uint32_t pos = 0; pos < S; ++pos) {
            tri_slot_at[pos] = running;
            running += s->shape_type[order[pos]] == RTGPU_TRIANGLE ? 1u : 0u;
        }
    }
    float cull_coord_max_of[64] = {0};
    parallel_for(S, 1u << 13, [&](uint32_t pos_begin, uint32_t pos_end, unsigned worker) {
    float coord_max_local = 0.0f;
    for (uint32_t pos = pos_begin; pos < pos_end; ++pos) {
        const uint32_t i = order[pos];
        double* g = R + (size_t)pos * rt::SHAPE_REALS;
        memcpy(g, s->shape_inv + (size_t)i * 12, 12 * sizeof(double));
        g[rt::SHAPE_MIN] = s->shape_min[i];
        g[rt::SHAPE_MAX] = s->shape_max[i];
        const uint32_t mat = s->shape_material[i];
        const uint32_t cls = s->shape_eq_class[i];
        int flags = 0;
        if (s->shape_closed[i]) flags |= rt::FLAG_CLOSED;
        if (s->mat_casts_shadow[mat]) flags |= rt::FLAG_CASTS_SHADOW;
        if ((class_size[cls] & 1u) && class_last[cls] == i) flags |= rt::FLAG_CONTAINER_REP;
        flags |= (int)s->shape_type[i] << rt::FLAG_TYPE_SHIFT;
        int* m = I + (size_t)pos * rt::SHAPE_INTS;
        m[0] = (int)i;
        m[1] = (int)mat;
        m[2] = flags;
        m[3] = (int)cls;
        m[4] = -1;
        bounding_sphere(s, i, R + lay.cull_off + (size_t)pos * rt::CULL_REALS);
        {
            // single-precision copy for the f64 kernels' pre-test (rt_kernel.cuh trace_unified): {cx, cy, cz, r2}, centre
            // rounded to nearest (the kernel's margin covers that), radius^2 padded by 2^-9 and rounded up; an infinite
            // or unrepresentable record stays "never culled"
            const double* c64 = R + lay.cull_off + (size_t)pos * rt::CULL_REALS;
            float c32[4];
            for (int k = 0; k < 3; ++k) c32[k] = (float)c64[k];
            const double padded = c64[3] * (1.0 + 0x1p-9);
            float r2 = (float)padded;
            if ((double)r2 < padded) r2 = std::nextafterf(r2, std::numeric_limits<float>::infinity());
            c32[3] = r2;
            if (std::isfinite(c64[3]))
                for (int k = 0; k < 3; ++k) {
                    const float a = std::nextafterf(std::fabs(c32[k]), std::numeric_limits<float>::infinity());
                    if (!std::isfinite(a)) coord_max_local = std::numeric_limits<float>::infinity();  // the kernel then never culls
                    else if (a > coord_max_local) coord_max_local = a;
                }
            memcpy(I + lay.cull32_off + (size_t)pos * 4, c32, sizeof(c32));
        }
        if (s->shape_type[i] == RTGPU_TRIANGLE) {
            const size_t t = (size_t)s->shape_triangle[i] * 3;
            const uint32_t tri_slot = tri_slot_at[pos];
            m[4] = (int)tri_slot;
            double* td = R + lay.tri_off + (size_t)tri_slot * rt::TRI_REALS;
            memcpy(td + 0, s->tri_vertex_1 + t, 3 * sizeof(double));
            memcpy(td + 3, s->tri_edge_1 + t, 3 * sizeof(double));
            memcpy(td + 6, s->tri_edge_2 + t, 3 * sizeof(double));
            memcpy(td + 9, s->tri_normal + t, 3 * sizeof(double));
        }
    }
    cull_coord_max_of[worker] = coord_max_local;
    });
    for (float v : cull_coord_max_of) lay.cull_coord_max = std::max(lay.cull_coord_max, v);  // (inf stays inf)
    mark("shapes (records, cull spheres)");
    for (uint32_t m = 0; m < M; ++m) {
        double* d = R + lay.mat_off + (size_t)m * rt::MAT_REALS;
        memcpy(d, s->mat_color + (size_t)m * 3, 3 * sizeof(double));
        memcpy(d + 3, s->mat_params + (size_t)m * RTGPU_MAT_PARAM_COUNT, RTGPU_MAT_PARAM_COUNT * sizeof(double));
        I[lay.mat_meta_off + m * rt::MAT_INTS + 0] = s->mat_pattern[m];
        I[lay.mat_meta_off + m * rt::MAT_INTS + 1] = s->mat_casts_shadow[m] ? 1 : 0;
    }
    for (uint32_t q = 0; q < Q; ++q) {
        double* d = R + lay.pat_off + (size_t)q * rt::PAT_REALS;
        memcpy(d, s->pat_color_a + (size_t)q * 3, 3 * sizeof(double));
        memcpy(d + 3, s->pat_color_b + (size_t)q * 3, 3 * sizeof(double));
        memcpy(d + 6, s->pat_inv + (size_t)q * 12, 12 * sizeof(double));
        int* pm = I + lay.pat_meta_off + q * rt::PAT_INTS;
        pm[0] = s->pat_type[q];
        pm[1] = s->pat_child_a[q];
        pm[2] = s->pat_child_b[q];
    }
    for (uint32_t l = 0; l < L; ++l) {
        double* d = R + lay.light_off + (size_t)l * rt::LIGHT_REALS;
        memcpy(d, s->light_position + (size_t)l * 3, 3 * sizeof(double));
        memcpy(d + 3, s->light_intensity + (size_t)l * 3, 3 * sizeof(double));
    }
    if (use_bvh) {
        const double inf = std::numeric_limits<double>::infinity();
        auto leaf_ref = [&](int32_t ref) { return ref >= 0 ? ref : ~(int32_t)(n_flat + (uint32_t)(~ref)); };  // leaf k -> sorted position
        if (n_nodes == 0) {
            // one bounded shape: a node whose second child is an empty box
            double* nb = R + lay.bvh_off;
            const rt::Aabb& b = boxes[order[n_flat]];
            for (int k = 0; k < 3; ++k) {
                nb[k] = b.lo[k];
                nb[3 + k] = b.hi[k];
                nb[6 + k] = inf;
                nb[9 + k] = -inf;
            }
            I[lay.bvh_meta_off + 0] = ~(int32_t)n_flat;
            I[lay.bvh_meta_off + 1] = ~(int32_t)n_flat;
        }
        for (uint32_t k = 0; k < n_nodes; ++k) {
            double* nb = R + lay.bvh_off + (size_t)k * rt::BVH_REALS;
            for (int c = 0; c < 2; ++c)
                for (int a = 0; a < 3; ++a) {
                    nb[c * 6 + a] = bvh.nodes[k].box[c].lo[a];
                    nb[c * 6 + 3 + a] = bvh.nodes[k].box[c].hi[a];
                }
            I[lay.bvh_meta_off + k * rt::BVH_INTS + 0] = leaf_ref(bvh.nodes[k].child[0]);
            I[lay.bvh_meta_off + k * rt::BVH_INTS + 1] = leaf_ref(bvh.nodes[k].child[1]);
        }
    }
    mark("materials, patterns, lights, nodes");
    if (use_bvh) {
        // single-precision node copies: boxes rounded outwards, children alongside (rt_scene.h BVH32_WORDS)
        auto down = [](double v) { float f = (float)v; return ((double)f > v) ? std::nextafterf(f, -std::numeric_limits<float>::infinity()) : f; };
        auto up = [](double v) { float f = (float)v; return ((double)f < v) ? std::nextafterf(f, std::numeric_limits<float>::infinity()) : f; };
        float coord_max_of[64] = {0};
        parallel_for(lay.n_bvh_nodes, 1u << 14, [&](uint32_t k_begin, uint32_t k_end, unsigned worker) {
        float coord_max = 0.0f;
        for (uint32_t k = k_begin; k < k_end; ++k) {
            const double* nb = R + lay.bvh_off + (size_t)k * rt::BVH_REALS;
            int* w = I + lay.bvh32_off + (size_t)k * rt::BVH32_WORDS;
            float f[12];
            for (int c = 0; c < 2; ++c)
                for (int a = 0; a < 3; ++a) {
                    f[c * 6 + a] = down(nb[c * 6 + a]);
                    f[c * 6 + 3 + a] = up(nb[c * 6 + 3 + a]);
                }
            for (int j = 0; j < 12; ++j)
                if (std::isfinite(f[j])) coord_max = std::max(coord_max, std::fabs(f[j]));
            memcpy(w, f, sizeof(f));
            w[12] = I[lay.bvh_meta_off + k * rt::BVH_INTS + 0];
            w[13] = I[lay.bvh_meta_off + k * rt::BVH_INTS + 1];
            w[14] = w[15] = 0;
        }
        coord_max_of[worker] = coord_max;
        });
        float coord_max = 0.0f;
        for (float v : coord_max_of) coord_max = std::max(coord_max, v);
        lay.bvh_coord_max = coord_max;
    }
    mark("single-precision nodes");
    // FNV-1a over (a sample of) the blobs: tells the family tuner whether two uploads are the same scene
    uint64_t fp = 1469598103934665603ull;
    auto mix = [&fp](uint64_t v) { fp = (fp ^ v) * 1099511628211ull; };
    mix(lay.n_reals);
    mix(lay.n_ints);
    const size_t step_r = std::max<size_t>(1, out->reals.size() >> 15), step_i = std::max<size_t>(1, out->ints.size() >> 15);
    for (size_t k = 0; k < out->reals.size(); k += step_r) {
        uint64_t bits;
        memcpy(&bits, &out->reals[k], sizeof(bits));
        mix(bits);
    }
    for (size_t k = 0; k < out->ints.size(); k += step_i) mix((uint64_t)(uint32_t)out->ints[k]);
    out->fingerprint = fp;
    mark("fingerprint");
    return RTGPU_OK;
}

// 64-bit hash of every byte a scene description points to (and its counts): rtgpu_render renders many frames of
// one scene, and packing + uploading it again each time (seconds for a BVH scene) only makes sense when it changed.
uint64_t hash_words(uint64_t h, const void* data, size_t bytes) {
    if (!data) return h * 0x9E3779B97F4A7C15ull + bytes;
    const unsigned char* p = static_cast<const unsigned char*>(data);
    size_t k = 0;
    for (; k + 8 <= bytes; k += 8) {
        uint64_t w;
        memcpy(&w, p + k, 8);
        h = (h ^ w) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
    }
    uint64_t tail = 0;
    if (k < bytes) memcpy(&tail, p + k, bytes - k);
    h = (h ^ tail ^ ((uint64_t)bytes << 56)) * 0xC2B2AE3D27D4EB4Full;
    return h ^ (h >> 32);
}

uint64_t scene_input_hash(const rtgpu_scene* s) {
    if (!s) return 0;
    uint64_t h = 0x243F6A8885A308D3ull;
    const uint32_t counts[6] = {s->abi_version, s->n_shapes, s->n_triangles, s->n_materials, s->n_patterns, s->n_lights};
    h = hash_words(h, counts, sizeof(counts));
    const size_t S = s->n_shapes, NT = s->n_triangles, M = s->n_materials, Q = s->n_patterns, L = s->n_lights;
    h = hash_words(h, s->shape_type, S);
    h = hash_words(h, s->shape_inv, S * 12 * sizeof(double));
    h = hash_words(h, s->shape_min, S * sizeof(double));
    h = hash_words(h, s->shape_max, S * sizeof(double));
    h = hash_words(h, s->shape_closed, S);
    h = hash_words(h, NT ? s->shape_triangle : nullptr, NT ? S * sizeof(int32_t) : 0);
    h = hash_words(h, s->shape_material, S * sizeof(uint32_t));
    h = hash_words(h, s->shape_eq_class, S * sizeof(uint32_t));
    h = hash_words(h, s->tri_vertex_1, NT * 3 * sizeof(double));
    h = hash_words(h, s->tri_edge_1, NT * 3 * sizeof(double));
    h = hash_words(h, s->tri_edge_2, NT * 3 * sizeof(double));
    h = hash_words(h, s->tri_normal, NT * 3 * sizeof(double));
    h = hash_words(h, s->mat_color, M * 3 * sizeof(double));
    h = hash_words(h, s->mat_params, M * RTGPU_MAT_PARAM_COUNT * sizeof(double));
    h = hash_words(h, s->mat_casts_shadow, M);
    h = hash_words(h, s->mat_pattern, M * sizeof(int32_t));
    h = hash_words(h, s->pat_type, Q);
    h = hash_words(h, s->pat_color_a, Q * 3 * sizeof(double));
    h = hash_words(h, s->pat_color_b, Q * 3 * sizeof(double));
    h = hash_words(h, s->pat_inv, Q * 12 * sizeof(double));
    h = hash_words(h, s->pat_child_a, Q * sizeof(int32_t));
    h = hash_words(h, s->pat_child_b, Q * sizeof(int32_t));
    h = hash_words(h, s->light_position, L * 3 * sizeof(double));
    h = hash_words(h, s->light_intensity, L * 3 * sizeof(double));
    const char* bvh_min = getenv("RTGPU_BVH_MIN");  // changes what pack_scene builds from the same input
    h = hash_words(h, bvh_min, bvh_min ? strlen(bvh_min) : 0);
    return h ? h : 1;
}

// ---------------------------------------------------------------------------------------------
// Rows of a launch

struct RowSel {
    uint32_t band_rows, shard_index, shard_count;
    uint32_t take = 1;  // bands rendered out of every shard_count, starting at band shard_index (public selections: 1)
};

int normalise_rows(const rtgpu_rows* rows, uint32_t vsize, RowSel* out) {
    RowSel r;
    r.band_rows = (rows && rows->band_rows) ? rows->band_rows : (vsize ? vsize : 1u);
    r.shard_count = (rows && rows->shard_count) ? rows->shard_count : 1u;
    r.shard_index = rows ? rows->shard_index : 0u;
    if (r.shard_index >= r.shard_count) return fail(RTGPU_ERR_INVALID_ARGUMENT, "rows.shard_index %u >= shard_count %u", r.shard_index, r.shard_count);
    // a band taller than the image is the whole image; the period (band_rows * shard_count) must fit 32 bits so that
    // the kernels' row arithmetic cannot wrap
    if (vsize && r.band_rows > vsize) r.band_rows = vsize;
    if ((uint64_t)r.band_rows * r.shard_count > 0xFFFFFFFFull)
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "rows.band_rows %u x shard_count %u overflows", r.band_rows, r.shard_count);
    *out = r;
    return RTGPU_OK;
}

uint32_t count_rows(const RowSel& r, uint32_t vsize) {
    if (r.band_rows == 0 || r.shard_count == 0) return 0;
    // bands b with (b % shard_count) in [shard_index, shard_index + take); the last band of the image may be partial
    auto selected_among_first = [&r](uint64_t n_bands) {
        const uint64_t part = n_bands % r.shard_count;
        const uint64_t in_part = part > r.shard_index ? std::min<uint64_t>(part - r.shard_index, r.take) : 0;
        return (n_bands / r.shard_count) * r.take + in_part;
    };
    const uint64_t full_bands = vsize / r.band_rows, rem_rows = vsize % r.band_rows;
    uint64_t n = selected_among_first(full_bands) * r.band_rows;
    const uint64_t last = full_bands % r.shard_count;  // position of the partial band inside its period
    if (rem_rows && last >= r.shard_index && last < (uint64_t)r.shard_index + r.take) n += rem_rows;
    return (uint32_t)n;
}

// image row of the k-th selected row (the kernels' image_row)
uint32_t selected_row(const RowSel& r, uint32_t k) {
    const uint32_t q = k / r.band_rows;
    return ((q / r.take) * r.shard_count + r.shard_index + q % r.take) * r.band_rows + k % r.band_rows;
}

// ---------------------------------------------------------------------------------------------
// Context: one scene resident on one device

}  // namespace

struct rtgpu_context {
    int device = 0;
    rt::SceneLayout layout{};
    double* d_reals64 = nullptr;
    float* d_reals32 = nullptr;
    int* d_ints = nullptr;
    unsigned int* d_work = nullptr;            // pixel-slot counter of the persistent kernel
    unsigned long long* d_counters = nullptr;  // rt::NUM_COUNTERS
    // host-buffer path
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void* d_out = nullptr;
    size_t d_out_bytes = 0;
    uint8_t* d_out8 = nullptr;
    size_t d_out8_bytes = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    bool has_cyl_cone_tri = false;  // selects the kernel instantiated with those shape types
    size_t cap_reals = 0, cap_ints = 0;
    bool zero_copy = false;  // the last host render wrote straight into the caller's pinned buffers
    // wavefront path (rt_wavefront.cuh): two ray queues, the node array and the frame's bookkeeping
    void* d_wf_rays[2] = {nullptr, nullptr};
    void* d_wf_nodes = nullptr;
    rt::WfCounts* d_wf_counts = nullptr;
    unsigned long long* d_wf_priv = nullptr;  // the frame's own work counters, committed to the caller's once it is complete
    unsigned long long* d_wf_keys = nullptr;  // binned queues (rt_wavefront.cuh wf_bin_kernel): (bin, rank) per queue entry ...
    unsigned* d_wf_perm = nullptr;            // ... and the permutation the next level consumes its queue through
    size_t wf_bin_entries = 0;                // entries both arrays hold
    uint64_t launches = 0;                    // kernels launched for this context so far (rtgpu_context_launch_count)
    size_t wf_cap_rays = 0, wf_cap_nodes = 0;  // in elements
    size_t wf_bytes_rays = 0, wf_bytes_nodes = 0;
    bool wf_used = false;  // the last launch took the wavefront path (overflow must be checked after it)
    bool wf_keep_overflow = false;  // this launch continues a frame rendered in chunks: the overflow flag is sticky
    // host-buffer renders in chunks (wavefront family): the copy of one chunk overlaps the kernels of the next
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_chunk = nullptr, ev_copied = nullptr;
    // kernel-family tuner (FAMILY_AUTO): per (scene, frame shape, path) the best time of each family
    bool has_secondary = false;
    uint32_t secondary_shapes = 0;
    uint64_t scene_fingerprint = 0;
    struct TuneEntry {
        uint64_t key = 0;
        float best_ms[2] = {1e30f, 1e30f};
        int runs[2] = {0, 0};
    };
    std::vector<TuneEntry> tune;
    cudaEvent_t tune_ev0 = nullptr, tune_ev1 = nullptr;
    int tune_pending_family = -1;  // a trial whose events have been recorded but not read yet
    uint64_t tune_pending_key = 0;
    uint64_t tune_last_key = 0;    // key of the host-buffer render in flight
    int last_family = 0;
    uint64_t uploaded_input_hash = 0;  // scene_input_hash of what upload_scene last put on the device (0 = unknown)
    // What the host reads after a frame — work counters, overflow flag and the sizes the frame asked for — lands in
    // pinned, device-mapped memory, written by the last (tiny) kernel of the frame: one stream sync, no copies.
    struct HostStatus {
        unsigned long long counters[rt::NUM_COUNTERS];
        unsigned overflow, max_rays, n_nodes, valid;
        unsigned long long queued_rays;  // hits queued over all levels of the frame (wavefront family)
    };
    HostStatus* h_status = nullptr;   // host view
    HostStatus* d_status = nullptr;   // device alias of the same memory
};

namespace {

// Bytes of dynamic shared memory the staged scene needs, or 0 when it does not fit the budget that keeps
// several CTAs resident per SM (then the kernels instantiated with SMEM = false read it from global memory).
template <typename T>
size_t scene_smem_bytes(const rtgpu_context* ctx) {
    const rt::SceneLayout& lay = ctx->layout;
    const size_t smem = (((size_t)lay.n_reals * sizeof(T) + 15) & ~size_t(15)) + (size_t)lay.n_ints * sizeof(int);
    const size_t smem_cap = std::min<size_t>(ctx->smem_optin, 64 * 1024);
    return smem <= smem_cap ? smem : 0;
}

template <typename T, int MAX_FRAMES, bool FULL, bool BVH, bool SMEM>
int launch_kernel_impl(rtgpu_context* ctx, const T* d_reals, const rt::CameraParams<T>& cam, T* d_out, uint8_t* d_out8,
                       unsigned long long* d_counters, cudaStream_t stream) {
    auto kernel = rt::render_kernel<T, MAX_FRAMES, FULL, BVH, SMEM>;
    rt::SceneLayout lay = ctx->layout;
    const size_t smem = SMEM ? scene_smem_bytes<T>(ctx) : 0;
    lay.in_shared = SMEM ? 1u : 0u;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, RT_BLOCK_THREADS, smem));
    if (blocks_per_sm < 1) return fail(RTGPU_ERR_CUDA, "render kernel does not fit on an SM (smem %zu B)", smem);
    // persistent grid: every SM full, no more; never more CTAs than there are 64-slot chunks of work
    const uint64_t slots = (uint64_t)((cam.hsize + rt::TILE_W - 1) / rt::TILE_W) * ((cam.n_rows + rt::TILE_H - 1) / rt::TILE_H) * 32ull;
    const uint64_t warps_needed = (slots + rt::CHUNK_SLOTS - 1) / rt::CHUNK_SLOTS;
    const uint64_t blocks_needed = (warps_needed + (RT_BLOCK_THREADS / 32) - 1) / (RT_BLOCK_THREADS / 32);
    uint64_t grid = (uint64_t)ctx->sm_count * (uint64_t)blocks_per_sm;
    if (grid > blocks_needed) grid = blocks_needed;
    if (grid < 1) grid = 1;
    CUDA_TRY(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int), stream));
    kernel<<<(unsigned)grid, RT_BLOCK_THREADS, smem, stream>>>(d_reals, ctx->d_ints, lay, cam, d_out, d_out8, d_counters, ctx->d_work);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return RTGPU_OK;
}

// ---- wavefront path ------------------------------------------------------------------------------

#ifndef RTGPU_DEFAULT_WAVEFRONT
#define RTGPU_DEFAULT_WAVEFRONT 0
#endif

__global__ void wf_commit_counters_kernel(const unsigned long long* priv, unsigned long long* user) {
    rt::wf_wait_for_previous();  // (no early release: this may be the last kernel of the frame)
    if (threadIdx.x < rt::NUM_COUNTERS && priv[threadIdx.x]) atomicAdd(&user[threadIdx.x], priv[threadIdx.x]);
}

// Last kernel of a host-buffer frame: everything the host wants to know, written into mapped pinned memory.
__global__ void publish_status_kernel(const unsigned long long* counters, const rt::WfCounts* wf, rtgpu_context::HostStatus* out) {
    rt::wf_wait_for_previous();
    if (threadIdx.x < rt::NUM_COUNTERS) out->counters[threadIdx.x] = counters ? counters[threadIdx.x] : 0ull;
    if (threadIdx.x == 0) {
        unsigned max_rays = 0;
        unsigned long long queued = 0;
        if (wf)
            for (int d = 0; d < 18; ++d) {
                max_rays = max(max_rays, wf->n_rays[d] + wf->n_back[d]);
                queued += (unsigned long long)wf->n_rays[d] + wf->n_back[d];
            }
        out->queued_rays = queued;
        out->overflow = wf ? wf->overflow : 0u;
        out->max_rays = max_rays;
        out->n_nodes = wf ? wf->n_nodes : 0u;
        __threadfence_system();
        out->valid = 1u;
    }
}

enum { FAMILY_PERSISTENT = 0, FAMILY_WAVEFRONT = 1, FAMILY_AUTO = 2 };
thread_local int g_last_family = FAMILY_PERSISTENT;

// What the caller asked for: opts.flags, else RTGPU_FAMILY=persistent|wavefront|auto (or RTGPU_WAVEFRONT=0/1), else auto.
int requested_family(const rtgpu_opts* opts) {
    if (opts && (opts->flags & RTGPU_FLAG_PERSISTENT)) return FAMILY_PERSISTENT;
    if (opts && (opts->flags & RTGPU_FLAG_WAVEFRONT)) return FAMILY_WAVEFRONT;
    const char* f = getenv("RTGPU_FAMILY");
    if (f && *f) {
        if (f[0] == 'p') return FAMILY_PERSISTENT;
        if (f[0] == 'w') return FAMILY_WAVEFRONT;
        return FAMILY_AUTO;
    }
    const char* e = getenv("RTGPU_WAVEFRONT");
    if (e && *e) return e[0] != '0' ? FAMILY_WAVEFRONT : FAMILY_PERSISTENT;
    return RTGPU_DEFAULT_WAVEFRONT != 0 ? FAMILY_WAVEFRONT : FAMILY_AUTO;
}

// ---- family tuner ----------------------------------------------------------------------------------
// Which family is faster depends on the scene (how much the per-pixel work varies) and on how many pixels one
// launch covers, and no static rule we tried predicts it.  So FAMILY_AUTO measures: the first TUNE_RUNS frames of
// each family for a given (scene, frame shape, path) are timed with CUDA events on the stream they run on,
// alternating W, P, W, P; after that the faster family renders every frame.  Both families produce the same
// pixels and counters bit for bit (tests/test_gpu_parity.py), so the choice is invisible apart from the time.
constexpr int TUNE_RUNS = 2;

uint64_t tune_key(const rtgpu_context* ctx, const rtgpu_camera* camera, const RowSel& sel, uint32_t precision, uint32_t max_depth, uint32_t path) {
    uint64_t k = ctx->scene_fingerprint;
    auto mix = [&k](uint64_t v) { k = (k ^ v) * 1099511628211ull; };
    mix(camera->hsize);
    mix(camera->vsize);
    mix(sel.band_rows);
    mix(sel.shard_count);
    mix(precision);
    mix(max_depth);
    mix(path);
    return k ? k : 1;
}

void tune_collect(rtgpu_context* ctx) {
    if (ctx->tune_pending_family < 0) return;
    const int family = ctx->tune_pending_family;
    ctx->tune_pending_family = -1;
    float ms = 0.f;
    if (cudaEventSynchronize(ctx->tune_ev1) != cudaSuccess || cudaEventElapsedTime(&ms, ctx->tune_ev0, ctx->tune_ev1) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (auto& e : ctx->tune)
        if (e.key == ctx->tune_pending_key) {
            e.best_ms[family] = std::min(e.best_ms[family], ms);
            e.runs[family]++;
            return;
        }
}

// Family for this frame; *trial = the caller should bracket the frame with tune_begin / tune_end.
int resolve_family(rtgpu_context* ctx, const rtgpu_opts* opts, uint64_t key, bool* trial) {
    *trial = false;
    const int want = requested_family(opts);
    if (want != FAMILY_AUTO) return want;
    if (!ctx->has_secondary) return FAMILY_PERSISTENT;  // primary + shadow rays only: one level, nothing for the queues to even out
    tune_collect(ctx);
    rtgpu_context::TuneEntry* entry = nullptr;
    for (auto& e : ctx->tune)
        if (e.key == key) entry = &e;
    if (!entry) {
        if (ctx->tune.size() >= 64) ctx->tune.erase(ctx->tune.begin());
        ctx->tune.emplace_back();
        entry = &ctx->tune.back();
        entry->key = key;
    }
    if (entry->runs[0] < TUNE_RUNS || entry->runs[1] < TUNE_RUNS) {
        if (!ctx->tune_ev0 && (cudaEventCreate(&ctx->tune_ev0) != cudaSuccess || cudaEventCreate(&ctx->tune_ev1) != cudaSuccess)) {
            cudaGetLastError();
            return FAMILY_PERSISTENT;
        }
        *trial = true;
        // Order W, P, W, P: the first frame of a (scene, frame shape) — all a one-shot CLI call ever renders — takes
        // the family that wins on most scenes with secondary rays (cover, table, cylinders, reflect_refract, refraction;
        // profiles/r2_notes.md), and with the small first-guess queues its cold start costs the same as the persistent
        // kernel's (benchmarks/cold_one_shot.py).
        return (entry->runs[FAMILY_WAVEFRONT] <= entry->runs[FAMILY_PERSISTENT] && entry->runs[FAMILY_WAVEFRONT] < TUNE_RUNS) ? FAMILY_WAVEFRONT : FAMILY_PERSISTENT;
    }
    return entry->best_ms[FAMILY_WAVEFRONT] < entry->best_ms[FAMILY_PERSISTENT] ? FAMILY_WAVEFRONT : FAMILY_PERSISTENT;
}

// The wavefront family cannot get its buffers for this frame shape: automatic mode stays with the persistent kernel.
void tune_rule_out_wavefront(rtgpu_context* ctx, uint64_t key) {
    cudaGetLastError();
    for (auto& e : ctx->tune)
        if (e.key == key) {
            e.runs[FAMILY_PERSISTENT] = e.runs[FAMILY_WAVEFRONT] = TUNE_RUNS;
            e.best_ms[FAMILY_PERSISTENT] = 0.0f;
            e.best_ms[FAMILY_WAVEFRONT] = 1e30f;
        }
}

void tune_begin(rtgpu_context* ctx, cudaStream_t stream) { cudaEventRecord(ctx->tune_ev0, stream); }

void tune_end(rtgpu_context* ctx, cudaStream_t stream, uint64_t key, int family) {
    if (cudaEventRecord(ctx->tune_ev1, stream) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    ctx->tune_pending_family = family;
    ctx->tune_pending_key = key;
}

// First guess of the wavefront buffers per pixel of a launch.  The cover frame peaks at 0.63 queued hits per pixel in one
// level and needs 2.09 node records per pixel over all levels (2.0 made the first frame of every context overflow at
// level 5 and render twice); glass-heavy frames ask for more and get it through the overflow -> enlarge -> render-again path,
// once (the buffers are kept).  Small on purpose: the first frame of a process pays for allocating and first touching
// them (3 + 5 per pixel = 1.7 GB at 1080p cost a cold call ~1 s and its first frame 15 ms instead of 2).
constexpr double WF_RAYS_PER_PIXEL = 1.0, WF_NODES_PER_PIXEL = 2.25;
// Binned queues pay where the deeper launches are large and bound by divergence over a longer shape list: at least half
// of the shapes reflective or transparent (cover 18 of 19: -4.4 %; reflect_refract 7 of 13: -2.5 %; table 6 of 18: +-0;
// cylinders 3 of 11: +5 %), at least 8 shapes (scenes of 3-6 shapes: +6..8 % when forced on) and 2^18 pixels per launch
// (smaller launches are bound by launch latency); everything else stays in arrival order.  profiles/r2_notes.md.
#ifndef RT_WF_COMBINE_CTAS
#define RT_WF_COMBINE_CTAS 8  // CTAs of 256 threads per SM in the combine pass (full occupancy)
#endif
constexpr uint64_t WF_BIN_MIN_PIXELS = 1u << 18;
constexpr uint32_t WF_BIN_MIN_SHAPES = 8;

template <typename T>
int wavefront_reserve(rtgpu_context* ctx, uint64_t pixels, double growth) {
    // first guess: WF_RAYS_PER_PIXEL queued rays and WF_NODES_PER_PIXEL nodes per pixel; after an overflow the caller asks for more.
    // RTGPU_WF_INITIAL_SCALE scales the guess (tests use a small one to exercise the enlarge-and-render-again path).
    if (const char* e = getenv("RTGPU_WF_INITIAL_SCALE"); growth == 1.0 && e && *e && atof(e) > 0.0) growth = atof(e);
    size_t want_rays = std::max<size_t>((size_t)(WF_RAYS_PER_PIXEL * growth * (double)pixels), 1u << 12);
    size_t want_nodes = std::max<size_t>((size_t)(WF_NODES_PER_PIXEL * growth * (double)pixels), 1u << 12);
    want_rays = std::min<size_t>(want_rays, 0xFFFFFF00u);
    want_nodes = std::min<size_t>(want_nodes, 0x7FFFFF00u);
    const size_t bytes_rays = want_rays * sizeof(rt::WfRay<T>), bytes_nodes = want_nodes * sizeof(rt::WfNode<T>);
    // RTGPU_WF_MAX_BYTES caps what the family may hold on a device (two queues + the node array)
    if (const char* cap = getenv("RTGPU_WF_MAX_BYTES"); cap && *cap) {
        const unsigned long long limit = strtoull(cap, nullptr, 10);
        if (2ull * std::max(bytes_rays, ctx->wf_bytes_rays) + std::max(bytes_nodes, ctx->wf_bytes_nodes) > limit)
            return fail(RTGPU_ERR_OUT_OF_MEMORY, "wavefront buffers (%zu B rays x 2 + %zu B nodes) exceed RTGPU_WF_MAX_BYTES=%llu", bytes_rays,
                        bytes_nodes, limit);
    }
    if (bytes_rays > ctx->wf_bytes_rays) {
        for (int k = 0; k < 2; ++k) {
            if (ctx->d_wf_rays[k]) cudaFree(ctx->d_wf_rays[k]);
            ctx->d_wf_rays[k] = nullptr;
        }
        ctx->wf_bytes_rays = 0;
        for (int k = 0; k < 2; ++k) CUDA_TRY(cudaMalloc(&ctx->d_wf_rays[k], bytes_rays));
        ctx->wf_bytes_rays = bytes_rays;
    }
    if (bytes_nodes > ctx->wf_bytes_nodes) {
        if (ctx->d_wf_nodes) cudaFree(ctx->d_wf_nodes);
        ctx->d_wf_nodes = nullptr;
        ctx->wf_bytes_nodes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_wf_nodes, bytes_nodes));
        ctx->wf_bytes_nodes = bytes_nodes;
    }
    ctx->wf_cap_rays = ctx->wf_bytes_rays / sizeof(rt::WfRay<T>);
    ctx->wf_cap_nodes = ctx->wf_bytes_nodes / sizeof(rt::WfNode<T>);
    if (!ctx->d_wf_counts) CUDA_TRY(cudaMalloc(&ctx->d_wf_counts, sizeof(rt::WfCounts)));
    if (!ctx->d_wf_priv) CUDA_TRY(cudaMalloc(&ctx->d_wf_priv, rt::NUM_COUNTERS * sizeof(unsigned long long)));
    return RTGPU_OK;
}

double wall_ms();

// One kernel of a frame's chain.  dependent: launched with programmatic stream serialization — it may start while the
// previous kernel of the stream is still draining; every kernel of the chain waits (griddepcontrol.wait) before it
// reads what its predecessor wrote, and releases its own successor at once (rt_wavefront.cuh wf_release_dependents).
// RTGPU_PDL=0 launches everything the ordinary way.
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, bool dependent, Args&&... args) {
    static const bool allowed = !(getenv("RTGPU_PDL") && getenv("RTGPU_PDL")[0] == '0');
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (dependent && allowed) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

template <typename T, bool FULL, bool BVH, bool SMEM>
int launch_wavefront_impl(rtgpu_context* ctx, const T* d_reals, const rt::CameraParams<T>& cam, T* d_out, uint8_t* d_out8,
                          unsigned long long* d_counters, cudaStream_t stream) {
    auto level_kernel = rt::wf_level_kernel<T, FULL, BVH, SMEM>;
    rt::SceneLayout lay = ctx->layout;
    // staged scene tables (16-byte padded), then one PairScratch per warp for the pair-list trace
    const size_t smem = (SMEM ? ((scene_smem_bytes<T>(ctx) + 15) & ~size_t(15)) : 0) +
                        ((RT_WF_PAIRS && !BVH) ? (RT_WF_THREADS / 32) * sizeof(rt::PairScratch<T>) : 0) +
                        (size_t)rt::WF_PARK_REALS * RT_WF_THREADS * sizeof(T);
    lay.in_shared = SMEM ? 1u : 0u;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, level_kernel, RT_WF_THREADS, smem));
    if (blocks_per_sm < 1) return fail(RTGPU_ERR_CUDA, "wavefront kernel does not fit on an SM (smem %zu B)", smem);
    const uint64_t pixels = (uint64_t)cam.hsize * cam.n_rows;
    // node records address their pixel with 31 bits (rt_wavefront.cuh WfNode::link); reported like a buffer that cannot
    // be had, so that automatic mode renders such a frame with the persistent kernel
    if ((uint64_t)cam.hsize * (cam.out_full_frame ? cam.vsize : cam.n_rows) >= 0x7FFFFFFFull)
        return fail(RTGPU_ERR_OUT_OF_MEMORY, "the wavefront family indexes at most 2^31 - 1 pixels per launch");
    if (ctx->wf_cap_rays == 0 || (double)(ctx->wf_bytes_rays / sizeof(rt::WfRay<T>)) < WF_RAYS_PER_PIXEL * (double)pixels / 2) {
        int st = wavefront_reserve<T>(ctx, pixels, 1.0);
        if (st != RTGPU_OK) return st;
    }
    ctx->wf_cap_rays = ctx->wf_bytes_rays / sizeof(rt::WfRay<T>);
    ctx->wf_cap_nodes = ctx->wf_bytes_nodes / sizeof(rt::WfNode<T>);
    const unsigned grid = (unsigned)(ctx->sm_count * blocks_per_sm);
    const int levels = (int)cam.max_depth + 1;
    const char* dbg = getenv("RTGPU_DEBUG_SYNC");  // wait after every level and report the queue sizes on stderr
    const bool debug_sync = dbg && dbg[0] == '1';
    if (debug_sync)
        fprintf(stderr, "[rtgpu] wavefront: %llu pixels, grid %u x %d, smem %zu B, queues %zu rays / %zu nodes, FULL %d BVH %d SMEM %d\n",
                (unsigned long long)pixels, grid, RT_WF_THREADS, smem, ctx->wf_cap_rays, ctx->wf_cap_nodes, (int)FULL, (int)BVH, (int)SMEM);
    if (!ctx->d_wf_counts || !ctx->d_wf_priv || !ctx->d_wf_nodes || !ctx->d_wf_rays[0] || !ctx->d_wf_rays[1])
        return fail(RTGPU_ERR_CUDA, "wavefront buffers missing (counts %p priv %p nodes %p rays %p %p, cap %zu %zu)", (void*)ctx->d_wf_counts,
                    (void*)ctx->d_wf_priv, ctx->d_wf_nodes, ctx->d_wf_rays[0], ctx->d_wf_rays[1], ctx->wf_cap_rays, ctx->wf_cap_nodes);
    CUDA_TRY(cudaMemsetAsync(ctx->d_wf_counts, 0, ctx->wf_keep_overflow ? offsetof(rt::WfCounts, overflow) : sizeof(rt::WfCounts), stream));
    CUDA_TRY(cudaMemsetAsync(ctx->d_wf_priv, 0, rt::NUM_COUNTERS * sizeof(unsigned long long), stream));
    rt::WfNode<T>* nodes = reinterpret_cast<rt::WfNode<T>*>(ctx->d_wf_nodes);
    // Binned queues for scenes whose whole shape list fits the bins (rt_wavefront.cuh RT_WF_BINS); small launches are
    // bound by launch latency, not by divergence, and skip the extra kernel per level.  RTGPU_WF_BINS=0 / 1 forces it.
    bool binned = !BVH && lay.type_begin[rt::NUM_SHAPE_TYPES] <= RT_WF_BINS / 2 && lay.type_begin[rt::NUM_SHAPE_TYPES] >= WF_BIN_MIN_SHAPES && pixels >= WF_BIN_MIN_PIXELS &&
                  2u * ctx->secondary_shapes >= lay.n_shapes;
    if (const char* e = getenv("RTGPU_WF_BINS"); e && *e) binned = !BVH && e[0] == '1';
    if (binned && ctx->wf_bin_entries < ctx->wf_cap_rays) {
        if (ctx->d_wf_keys) cudaFree(ctx->d_wf_keys);
        if (ctx->d_wf_perm) cudaFree(ctx->d_wf_perm);
        ctx->d_wf_keys = nullptr;
        ctx->d_wf_perm = nullptr;
        ctx->wf_bin_entries = 0;
        const size_t entries = ctx->wf_bytes_rays / sizeof(rt::WfRay<float>);  // enough for either precision
        CUDA_TRY(cudaMalloc(&ctx->d_wf_keys, entries * sizeof(unsigned long long)));
        CUDA_TRY(cudaMalloc(&ctx->d_wf_perm, entries * sizeof(unsigned)));
        ctx->wf_bin_entries = entries;
    }
    for (int level = 0; level < levels; ++level) {
        const rt::WfRay<T>* in = reinterpret_cast<const rt::WfRay<T>*>(ctx->d_wf_rays[level & 1]);
        rt::WfRay<T>* out = reinterpret_cast<rt::WfRay<T>*>(ctx->d_wf_rays[(level + 1) & 1]);
        // level 0 follows the frame's memsets (an ordinary dependency); everything after it is chained
        CUDA_TRY(launch_chain(level_kernel, grid, RT_WF_THREADS, smem, stream, level > 0, d_reals, (const int*)ctx->d_ints, lay, cam, level, in, out,
                              (unsigned)ctx->wf_cap_rays, nodes, (unsigned)ctx->wf_cap_nodes, ctx->d_wf_counts, d_out, d_out8, ctx->d_wf_priv,
                              (const unsigned*)((binned && level > 0) ? ctx->d_wf_perm : nullptr),
                              (unsigned long long*)((binned && level + 1 < levels) ? ctx->d_wf_keys : nullptr)));
        ctx->launches++;
        if (binned && level + 1 < levels) {
            CUDA_TRY(launch_chain(rt::wf_bin_kernel, (unsigned)ctx->sm_count * 4u, 256u, 0, stream, true, (const rt::WfCounts*)ctx->d_wf_counts, level + 1,
                                  (unsigned)ctx->wf_cap_rays, (const unsigned long long*)ctx->d_wf_keys, ctx->d_wf_perm));
            ctx->launches++;
        }
        if (debug_sync) {
            const double t0 = wall_ms();
            CUDA_TRY(cudaStreamSynchronize(stream));
            rt::WfCounts h;
            CUDA_TRY(cudaMemcpy(&h, ctx->d_wf_counts, sizeof(h), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[rtgpu] wavefront level %d done (+%.2f ms wait): next queue %u front + %u back, %u nodes, overflow %u\n", level,
                    wall_ms() - t0, h.n_rays[level + 1], h.n_back[level + 1], h.n_nodes, h.overflow);
        }
    }
    for (int level = levels - 1; level >= 0; --level)
        CUDA_TRY(launch_chain(rt::wf_combine_kernel<T>, (unsigned)ctx->sm_count * RT_WF_COMBINE_CTAS, 256u, 0, stream, !debug_sync, nodes,
                              (const rt::WfCounts*)ctx->d_wf_counts, level, (unsigned)ctx->wf_cap_nodes, d_out, d_out8,
                              (level > 0 || d_counters != nullptr) ? 1 : 0));  // the counter commit follows the last pass
    ctx->launches += (uint64_t)levels;
    CUDA_TRY(cudaGetLastError());
    ctx->wf_used = true;
    return RTGPU_OK;
}

template <typename T>
int launch_wavefront(rtgpu_context* ctx, const T* d_reals, const rt::CameraParams<T>& cam, T* d_out, uint8_t* d_out8,
                     unsigned long long* d_counters, cudaStream_t stream) {
    const int variant = (ctx->has_cyl_cone_tri ? 4 : 0) | (ctx->layout.n_bvh_nodes > 0 ? 2 : 0) | (scene_smem_bytes<T>(ctx) ? 1 : 0);
    switch (variant) {
    case 0: return launch_wavefront_impl<T, false, false, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 1: return launch_wavefront_impl<T, false, false, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 2: return launch_wavefront_impl<T, false, true, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 3: return launch_wavefront_impl<T, false, true, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 4: return launch_wavefront_impl<T, true, false, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 5: return launch_wavefront_impl<T, true, false, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 6: return launch_wavefront_impl<T, true, true, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    default: return launch_wavefront_impl<T, true, true, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    }
}

// After a wavefront launch has completed on `stream`: did a queue or the node array overflow?
// Returns 1 (and enlarges the buffers) when the frame has to be rendered again, 0 when it is complete.
// published: the frame ended with publish_status_kernel and the stream has been synchronised since — the answer is
// already in mapped host memory; otherwise it is published and waited for here.
template <typename T>
int wavefront_check(rtgpu_context* ctx, uint64_t pixels, cudaStream_t stream, bool published = false) {
    if (!ctx->wf_used) return 0;
    ctx->wf_used = false;
    if (!published || !ctx->h_status->valid) {
        ctx->h_status->valid = 0u;
        CUDA_TRY(launch_chain(publish_status_kernel, 1u, 32u, 0, stream, true, (const unsigned long long*)nullptr, (const rt::WfCounts*)ctx->d_wf_counts, ctx->d_status));
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    const rtgpu_context::HostStatus h = *ctx->h_status;
    if (!h.overflow) return 0;
    // what the frame actually asked for, with headroom
    const double g_rays = (double)h.max_rays / (WF_RAYS_PER_PIXEL * (double)pixels), g_nodes = (double)h.n_nodes / (WF_NODES_PER_PIXEL * (double)pixels);
    const double growth = std::max(1.25 * std::max(g_rays, g_nodes), 2.0 * (double)ctx->wf_cap_rays / (WF_RAYS_PER_PIXEL * (double)pixels));
    int st = wavefront_reserve<T>(ctx, pixels, growth);
    if (st != RTGPU_OK) return st;
    return 1;
}

template <typename T, int MAX_FRAMES>
int launch_kernel(rtgpu_context* ctx, const T* d_reals, const rt::CameraParams<T>& cam, T* d_out, uint8_t* d_out8,
                  unsigned long long* d_counters, cudaStream_t stream) {
    // scenes without cylinders, cones and triangles run the kernel instantiated without those loops; scenes
    // below the BVH threshold the one without the traversal; small scenes the one that reads shared memory
    const int variant = (ctx->has_cyl_cone_tri ? 4 : 0) | (ctx->layout.n_bvh_nodes > 0 ? 2 : 0) | (scene_smem_bytes<T>(ctx) ? 1 : 0);
    switch (variant) {
    case 0: return launch_kernel_impl<T, MAX_FRAMES, false, false, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 1: return launch_kernel_impl<T, MAX_FRAMES, false, false, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 2: return launch_kernel_impl<T, MAX_FRAMES, false, true, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 3: return launch_kernel_impl<T, MAX_FRAMES, false, true, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 4: return launch_kernel_impl<T, MAX_FRAMES, true, false, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 5: return launch_kernel_impl<T, MAX_FRAMES, true, false, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    case 6: return launch_kernel_impl<T, MAX_FRAMES, true, true, false>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    default: return launch_kernel_impl<T, MAX_FRAMES, true, true, true>(ctx, d_reals, cam, d_out, d_out8, d_counters, stream);
    }
}

template <typename T>
void fill_camera(const rtgpu_camera* c, const RowSel& rows, uint32_t n_rows, uint32_t max_depth, bool full_frame_out, rt::CameraParams<T>* out) {
    out->out_full_frame = full_frame_out ? 1u : 0u;
    out->probe_ray = 0u;
    out->half_width = (T)c->half_width;
    out->half_height = (T)c->half_height;
    out->pixel_size = (T)c->pixel_size;
    for (int i = 0; i < 12; ++i) out->inv[i] = (T)c->inv[i];
    for (int i = 0; i < 3; ++i) out->origin[i] = (T)c->origin[i];
    out->hsize = c->hsize;
    out->vsize = c->vsize;
    out->n_rows = n_rows;
    out->band_rows = rows.band_rows;
    out->shard_index = rows.shard_index;
    out->shard_count = rows.shard_count;
    out->band_take = rows.take;
    out->max_depth = max_depth;
    // stride for the scattered tile order: near the golden ratio of the tile count, made coprime to it
    const uint64_t n_tiles = (uint64_t)((c->hsize + rt::TILE_W - 1) / rt::TILE_W) * ((n_rows + rt::TILE_H - 1) / rt::TILE_H);
    uint64_t stride = (uint64_t)((double)n_tiles * 0.6180339887498949);
    if (stride < 1) stride = 1;
    auto gcd = [](uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; };
    while (n_tiles > 1 && gcd(stride, n_tiles) != 1) ++stride;
    out->tile_stride = (uint32_t)(n_tiles > 1 ? stride % n_tiles : 0);
    if (n_tiles > 1 && out->tile_stride == 0) out->tile_stride = 1;
}

int ensure_f32_blob(rtgpu_context* ctx, cudaStream_t stream) {
    if (ctx->d_reals32 || ctx->layout.n_reals == 0) return RTGPU_OK;
    std::vector<double> h(ctx->layout.n_reals);
    CUDA_TRY(cudaMemcpyAsync(h.data(), ctx->d_reals64, h.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    std::vector<float> f(h.size());
    for (size_t i = 0; i < h.size(); ++i) f[i] = (float)h[i];
    CUDA_TRY(cudaMalloc(&ctx->d_reals32, f.size() * sizeof(float)));
    CUDA_TRY(cudaMemcpy(ctx->d_reals32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
    return RTGPU_OK;
}

int check_opts(const rtgpu_opts* opts, uint32_t* precision, uint32_t* max_depth) {
    *precision = opts ? opts->precision : (uint32_t)RTGPU_PRECISION_F64;
    *max_depth = opts ? opts->max_depth : 6u;
    if (*precision != RTGPU_PRECISION_F64 && *precision != RTGPU_PRECISION_F32)
        return fail(RTGPU_ERR_INVALID_ARGUMENT, "opts.precision %u", *precision);
    if (*max_depth > 15u) return fail(RTGPU_ERR_INVALID_ARGUMENT, "opts.max_depth %u > 15", *max_depth);
    return RTGPU_OK;
}

// One frame through the wavefront family.  blocking: wait, verify the buffers were large enough, render again
// with larger ones otherwise.  Non-blocking callers must run wavefront_check themselves after the stream drains.
template <typename T>
int render_wavefront(rtgpu_context* ctx, const T* d_reals, const rt::CameraParams<T>& cam, T* d_out, uint8_t* d_out8,
                     unsigned long long* counters, cudaStream_t stream, bool blocking) {
    const uint64_t pixels = (uint64_t)cam.hsize * cam.n_rows;
    for (int attempt = 0; attempt < 8; ++attempt) {
        int st = launch_wavefront<T>(ctx, d_reals, cam, d_out, d_out8, counters, stream);
        if (st != RTGPU_OK) return st;
        if (!blocking) {
            if (counters) {
                CUDA_TRY(launch_chain(wf_commit_counters_kernel, 1u, 32u, 0, stream, true, (const unsigned long long*)ctx->d_wf_priv, counters));
                ctx->launches++;
            }
            return RTGPU_OK;
        }
        st = wavefront_check<T>(ctx, pixels, stream);
        if (st < 0) return st;
        if (st == 0) {
            if (counters) {
                CUDA_TRY(launch_chain(wf_commit_counters_kernel, 1u, 32u, 0, stream, true, (const unsigned long long*)ctx->d_wf_priv, counters));
                ctx->launches++;
            }
            CUDA_TRY(cudaGetLastError());
            return RTGPU_OK;
        }
    }
    return fail(RTGPU_ERR_OUT_OF_MEMORY, "wavefront buffers kept overflowing");
}

int render_device_impl(rtgpu_context* ctx, const rtgpu_camera* camera, const rtgpu_opts* opts, const rtgpu_rows* rows,
                       void* d_out_rgb, uint8_t* d_out_rgb8, uint64_t* d_counters, cudaStream_t stream, uint32_t* out_n_rows,
                       int family, bool full_frame_out = false, bool wavefront_blocking = true, const RowSel* internal_sel = nullptr) {
    if (!ctx || !camera) return fail(RTGPU_ERR_INVALID_ARGUMENT, "context or camera is NULL");
    if (!d_out_rgb && !d_out_rgb8) return fail(RTGPU_ERR_INVALID_ARGUMENT, "both output pointers are NULL");
    uint32_t precision, max_depth;
    int st = check_opts(opts, &precision, &max_depth);
    if (st != RTGPU_OK) return st;
    RowSel sel;
    if (internal_sel) {
        sel = *internal_sel;
    } else {
        st = normalise_rows(rows, camera->vsize, &sel);
        if (st != RTGPU_OK) return st;
    }
    const uint32_t n_rows = count_rows(sel, camera->vsize);
    if (out_n_rows) *out_n_rows = n_rows;
    if (n_rows == 0 || camera->hsize == 0) return RTGPU_OK;  // nothing to render
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long* counters = reinterpret_cast<unsigned long long*>(d_counters);
    const bool wavefront = family == FAMILY_WAVEFRONT;
    ctx->last_family = g_last_family = wavefront ? FAMILY_WAVEFRONT : FAMILY_PERSISTENT;
    if (precision == RTGPU_PRECISION_F64) {
        rt::CameraParams<double> cam;
        fill_camera(camera, sel, n_rows, max_depth, full_frame_out, &cam);
        if (wavefront) return render_wavefront<double>(ctx, ctx->d_reals64, cam, (double*)d_out_rgb, d_out_rgb8, counters, stream, wavefront_blocking);
        if (max_depth <= 7) return launch_kernel<double, 8>(ctx, ctx->d_reals64, cam, (double*)d_out_rgb, d_out_rgb8, counters, stream);
        return launch_kernel<double, 16>(ctx, ctx->d_reals64, cam, (double*)d_out_rgb, d_out_rgb8, counters, stream);
    }
    st = ensure_f32_blob(ctx, stream);
    if (st != RTGPU_OK) return st;
    rt::CameraParams<float> cam;
    fill_camera(camera, sel, n_rows, max_depth, full_frame_out, &cam);
    if (wavefront) return render_wavefront<float>(ctx, ctx->d_reals32, cam, (float*)d_out_rgb, d_out_rgb8, counters, stream, wavefront_blocking);
    if (max_depth <= 7) return launch_kernel<float, 8>(ctx, ctx->d_reals32, cam, (float*)d_out_rgb, d_out_rgb8, counters, stream);
    return launch_kernel<float, 16>(ctx, ctx->d_reals32, cam, (float*)d_out_rgb, d_out_rgb8, counters, stream);
}

int upload_scene(rtgpu_context* ctx, const PackedScene& packed) {
    const rt::SceneLayout& lay = packed.layout;
    if (ctx->d_reals32) cudaFree(ctx->d_reals32);  // the f32 copy is rebuilt on demand
    ctx->d_reals32 = nullptr;
    ctx->layout = lay;
    ctx->has_cyl_cone_tri = packed.has_cyl_cone_tri;
    ctx->has_secondary = packed.has_secondary;
    ctx->secondary_shapes = packed.secondary_shapes;
    ctx->scene_fingerprint = packed.fingerprint;
    // repeated frames of similar scenes reuse the allocations (cudaFree / cudaMalloc synchronise the device)
    const size_t need_reals = std::max<size_t>(16, (size_t)lay.n_reals * sizeof(double));
    const size_t need_ints = std::max<size_t>(16, (size_t)lay.n_ints * sizeof(int));
    if (need_reals > ctx->cap_reals) {
        if (ctx->d_reals64) cudaFree(ctx->d_reals64);
        ctx->d_reals64 = nullptr;
        ctx->cap_reals = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_reals64, need_reals));
        ctx->cap_reals = need_reals;
    }
    if (need_ints > ctx->cap_ints) {
        if (ctx->d_ints) cudaFree(ctx->d_ints);
        ctx->d_ints = nullptr;
        ctx->cap_ints = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_ints, need_ints));
        ctx->cap_ints = need_ints;
    }
    if (lay.n_reals) CUDA_TRY(cudaMemcpyAsync(ctx->d_reals64, packed.reals.data(), (size_t)lay.n_reals * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (lay.n_ints) CUDA_TRY(cudaMemcpyAsync(ctx->d_ints, packed.ints.data(), (size_t)lay.n_ints * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return RTGPU_OK;
}

int context_init(rtgpu_context* ctx, int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RTGPU_ERR_NO_DEVICE, "no CUDA device available (gpu mode has no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(RTGPU_ERR_INVALID_ARGUMENT, "device %d out of range (0..%d)", device, n - 1);
    ctx->device = device;
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ctx->ev0));
    CUDA_TRY(cudaEventCreate(&ctx->ev1));
    CUDA_TRY(cudaMalloc(&ctx->d_work, sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&ctx->d_counters, rt::NUM_COUNTERS * sizeof(unsigned long long)));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_status), sizeof(rtgpu_context::HostStatus), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(ctx->h_status, 0, sizeof(rtgpu_context::HostStatus));
    CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_status), ctx->h_status, 0));
    return RTGPU_OK;
}

void context_release(rtgpu_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->d_reals64) cudaFree(ctx->d_reals64);
    if (ctx->d_reals32) cudaFree(ctx->d_reals32);
    if (ctx->d_ints) cudaFree(ctx->d_ints);
    if (ctx->d_work) cudaFree(ctx->d_work);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->d_out) cudaFree(ctx->d_out);
    if (ctx->d_out8) cudaFree(ctx->d_out8);
    for (int k = 0; k < 2; ++k)
        if (ctx->d_wf_rays[k]) cudaFree(ctx->d_wf_rays[k]);
    if (ctx->d_wf_nodes) cudaFree(ctx->d_wf_nodes);
    if (ctx->d_wf_counts) cudaFree(ctx->d_wf_counts);
    if (ctx->d_wf_priv) cudaFree(ctx->d_wf_priv);
    if (ctx->d_wf_keys) cudaFree(ctx->d_wf_keys);
    if (ctx->d_wf_perm) cudaFree(ctx->d_wf_perm);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_chunk) cudaEventDestroy(ctx->ev_chunk);
    if (ctx->ev_copied) cudaEventDestroy(ctx->ev_copied);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->tune_ev0) cudaEventDestroy(ctx->tune_ev0);
    if (ctx->tune_ev1) cudaEventDestroy(ctx->tune_ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int ensure_out_buffers(rtgpu_context* ctx, size_t rgb_bytes, size_t rgb8_bytes) {
    if (rgb_bytes > ctx->d_out_bytes) {
        if (ctx->d_out) cudaFree(ctx->d_out);
        ctx->d_out = nullptr;
        ctx->d_out_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_out, rgb_bytes));
        ctx->d_out_bytes = rgb_bytes;
    }
    if (rgb8_bytes > ctx->d_out8_bytes) {
        if (ctx->d_out8) cudaFree(ctx->d_out8);
        ctx->d_out8 = nullptr;
        ctx->d_out8_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->d_out8, rgb8_bytes));
        ctx->d_out8_bytes = rgb8_bytes;
    }
    return RTGPU_OK;
}

// Device-side alias of a pinned, mapped host allocation (cudaHostAlloc / cudaHostRegister / torch pin_memory),
// or nullptr for pageable memory.
void* mapped_device_pointer(const void* host) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) return nullptr;
    return attr.devicePointer;
}

// The frame's kernels are all enqueued: publish what the host will want to read once the stream has drained.
int publish_status(rtgpu_context* ctx) {
    ctx->h_status->valid = 0u;
    CUDA_TRY(launch_chain(publish_status_kernel, 1u, 32u, 0, ctx->stream, true, (const unsigned long long*)ctx->d_counters,
                          (const rt::WfCounts*)(ctx->wf_used ? ctx->d_wf_counts : nullptr), ctx->d_status));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return RTGPU_OK;
}

// Issue (do not wait for) everything one device does for a host-buffer render: kernel + D2H of its
// row bands into the full-frame host buffers.
int enqueue_host_render(rtgpu_context* ctx, const rtgpu_camera* camera, const rtgpu_opts* opts, const rtgpu_rows* rows,
                        void* out_rgb, uint8_t* out_rgb8, int forced_family = -1) {
    uint32_t precision, max_depth;
    int st = check_opts(opts, &precision, &max_depth);
    if (st != RTGPU_OK) return st;
    RowSel sel;
    st = normalise_rows(rows, camera->vsize, &sel);
    if (st != RTGPU_OK) return st;
    const uint32_t n_rows = count_rows(sel, camera->vsize);
    const size_t elem = precision == RTGPU_PRECISION_F64 ? sizeof(double) : sizeof(float);
    const size_t row_rgb = (size_t)camera->hsize * 3 * elem;
    const size_t row_rgb8 = (size_t)camera->hsize * 3;
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (n_rows == 0 || camera->hsize == 0) {
        // a shard without rows (more shards than bands): nothing to render or copy, but the caller still waits on
        // this context's events and reads its (zero) counters
        ctx->wf_used = false;
        ctx->zero_copy = false;
        CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, rt::NUM_COUNTERS * sizeof(unsigned long long), ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        return publish_status(ctx);
    }
    // Zero-copy: when the caller's buffers are pinned, device-mapped host memory the persistent kernel writes each
    // finished pixel straight into the caller's Canvas (full-frame indexing); the PCIe traffic then
    // overlaps the render instead of following it.  Pageable buffers take the staging path below, and so does
    // the wavefront family: it finishes most pixels in its last launches, and that burst of small stores
    // crosses PCIe far slower than one copy-engine transfer of the finished rows.
    void* map_rgb = out_rgb ? mapped_device_pointer(out_rgb) : nullptr;
    void* map_rgb8 = out_rgb8 ? mapped_device_pointer(out_rgb8) : nullptr;
    const char* zc = getenv("RTGPU_ZEROCOPY");
    const bool mappable = !(zc && zc[0] == '0') && (!out_rgb || map_rgb) && (!out_rgb8 || map_rgb8);
    bool trial = false;
    const uint64_t key = tune_key(ctx, camera, sel, precision, max_depth, 1u + (mappable ? 1u : 0u) + (out_rgb ? 2u : 0u) + (out_rgb8 ? 4u : 0u));
    int family = forced_family >= 0 ? forced_family : resolve_family(ctx, opts, key, &trial);
    ctx->tune_last_key = key;
    if (family == FAMILY_WAVEFRONT && n_rows && camera->hsize) {
        // allocate the queues before the timed region (cudaMalloc waits for the device)
        const uint64_t pixels = (uint64_t)camera->hsize * n_rows;
        const size_t ray_bytes = precision == RTGPU_PRECISION_F64 ? sizeof(rt::WfRay<double>) : sizeof(rt::WfRay<float>);
        if ((double)(ctx->wf_bytes_rays / ray_bytes) < WF_RAYS_PER_PIXEL * (double)pixels / 2) {
            st = precision == RTGPU_PRECISION_F64 ? wavefront_reserve<double>(ctx, pixels, 1.0) : wavefront_reserve<float>(ctx, pixels, 1.0);
            if (st == RTGPU_ERR_OUT_OF_MEMORY && forced_family < 0 && requested_family(opts) == FAMILY_AUTO) {
                tune_rule_out_wavefront(ctx, key);
                family = FAMILY_PERSISTENT;
                trial = false;
            } else if (st != RTGPU_OK) {
                return st;
            }
        }
    }
    ctx->zero_copy = mappable && family == FAMILY_PERSISTENT;
    if (!ctx->zero_copy) {
        st = ensure_out_buffers(ctx, out_rgb ? row_rgb * n_rows : 0, out_rgb8 ? row_rgb8 * n_rows : 0);
        if (st != RTGPU_OK) return st;
    }
    CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, rt::NUM_COUNTERS * sizeof(unsigned long long), ctx->stream));
    if (trial) tune_begin(ctx, ctx->stream);
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    if (ctx->zero_copy) {
        st = render_device_impl(ctx, camera, opts, rows, map_rgb, (uint8_t*)map_rgb8, reinterpret_cast<uint64_t*>(ctx->d_counters),
                                ctx->stream, nullptr, family, /*full_frame_out=*/true, /*wavefront_blocking=*/false);
        if (st != RTGPU_OK) return st;
        CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
        if ((st = publish_status(ctx)) != RTGPU_OK) return st;
        if (trial) tune_end(ctx, ctx->stream, key, family);
        return RTGPU_OK;
    }
    // Wavefront family, whole frame: render it in two interleaved parts (16-row bands) and copy the first part while
    // the second one renders; the queues are reused.  The parts are UNEQUAL: with T(f) ~ 0.25 + 1.62 f ms for a
    // fraction f of the 1080p cover frame and 0.94 f ms for its copy, the first part should be as large as the
    // second part's kernels can still hide its copy (f ~ 0.7): 7 of every 10 bands, then the other three — 2.60 -> 2.42 ms
    // end to end against equal halves at the time.  RTGPU_E2E_SPLIT=<take>/<period> overrides, RTGPU_E2E_CHUNKS=1 disables.
    const char* chunks_env = getenv("RTGPU_E2E_CHUNKS");
    // Only for pinned host buffers: a device-to-host copy into pageable memory blocks the submitting thread, so the
    // second part's kernels would not even be enqueued before the first part's copy has finished.
    const bool pinned = (!out_rgb || map_rgb) && (!out_rgb8 || map_rgb8);
    const bool chunked = family == FAMILY_WAVEFRONT && pinned && sel.shard_count == 1 && sel.band_rows >= camera->vsize && camera->vsize >= 64 &&
                         (uint64_t)camera->hsize * camera->vsize >= (1u << 18) && !(chunks_env && chunks_env[0] == '1');
    if (chunked) {
        if (!ctx->copy_stream) {
            CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_chunk, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_copied, cudaEventDisableTiming));
        }
        constexpr uint32_t BAND = 16, CHUNKS = 2;
        // cover@1080p end to end, ms — kernels 1.875: 1/2 2.60, 3/5 2.50, 2/3 2.46, 7/10 2.42, 3/4 2.42, 4/5 2.45; kernels 1.67
        // (end of round 2): 2/3 2.19, 7/10 2.17, 5/7 2.20, 3/4 2.24: the faster the kernels, the smaller the first part
        // whose copy the second part's kernels can still hide
        uint32_t take0 = 7, period = 10;
        if (const char* sp = getenv("RTGPU_E2E_SPLIT"); sp && *sp) {
            unsigned a = 0, b = 0;
            if (sscanf(sp, "%u/%u", &a, &b) == 2 && a >= 1 && b >= 2 && a < b && b <= 64) {
                take0 = a;
                period = b;
            }
        }
        size_t rows_before = 0;
        for (uint32_t c = 0; c < CHUNKS; ++c) {
            RowSel sub_sel{BAND, c == 0 ? 0u : take0, period, c == 0 ? take0 : period - take0};
            const uint32_t rows_c = count_rows(sub_sel, camera->vsize);
            char* d_rgb_c = out_rgb ? (char*)ctx->d_out + rows_before * row_rgb : nullptr;
            uint8_t* d_rgb8_c = out_rgb8 ? ctx->d_out8 + rows_before * row_rgb8 : nullptr;
            ctx->wf_keep_overflow = c > 0;
            st = render_device_impl(ctx, camera, opts, nullptr, d_rgb_c, d_rgb8_c, reinterpret_cast<uint64_t*>(ctx->d_counters), ctx->stream, nullptr,
                                    family, false, /*wavefront_blocking=*/false, &sub_sel);
            ctx->wf_keep_overflow = false;
            if (st != RTGPU_OK) return st;
            // the last part's copies follow its kernels on the main stream; earlier ones go to the copy stream
            cudaStream_t cs = ctx->stream;
            if (c + 1 < CHUNKS) {
                CUDA_TRY(cudaEventRecord(ctx->ev_chunk, ctx->stream));
                CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk, 0));
                cs = ctx->copy_stream;
            } else {
                CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
                if ((st = publish_status(ctx)) != RTGPU_OK) return st;
            }
            // every period contributes `take` consecutive bands, contiguous in the compact buffer AND in the image: one
            // strided copy for the complete groups, one plain copy for what is left of the last group
            const uint32_t group_rows = sub_sel.take * BAND;
            const uint32_t full_groups = rows_c / group_rows, tail_rows = rows_c % group_rows;
            const size_t first_row = (size_t)sub_sel.shard_index * BAND, period_rows = (size_t)period * BAND;
            if (out_rgb) {
                if (full_groups)
                    CUDA_TRY(cudaMemcpy2DAsync((char*)out_rgb + first_row * row_rgb, period_rows * row_rgb, d_rgb_c, (size_t)group_rows * row_rgb,
                                               (size_t)group_rows * row_rgb, full_groups, cudaMemcpyDeviceToHost, cs));
                if (tail_rows)
                    CUDA_TRY(cudaMemcpyAsync((char*)out_rgb + (first_row + (size_t)full_groups * period_rows) * row_rgb,
                                             d_rgb_c + (size_t)full_groups * group_rows * row_rgb, (size_t)tail_rows * row_rgb, cudaMemcpyDeviceToHost, cs));
            }
            if (out_rgb8) {
                if (full_groups)
                    CUDA_TRY(cudaMemcpy2DAsync(out_rgb8 + first_row * row_rgb8, period_rows * row_rgb8, d_rgb8_c, (size_t)group_rows * row_rgb8,
                                               (size_t)group_rows * row_rgb8, full_groups, cudaMemcpyDeviceToHost, cs));
                if (tail_rows)
                    CUDA_TRY(cudaMemcpyAsync(out_rgb8 + (first_row + (size_t)full_groups * period_rows) * row_rgb8,
                                             d_rgb8_c + (size_t)full_groups * group_rows * row_rgb8, (size_t)tail_rows * row_rgb8, cudaMemcpyDeviceToHost, cs));
            }
            if (c + 1 < CHUNKS) CUDA_TRY(cudaEventRecord(ctx->ev_copied, ctx->copy_stream));
            rows_before += rows_c;
        }
        CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_copied, 0));  // the main stream ends after every copy
        if (trial) tune_end(ctx, ctx->stream, key, family);
        return RTGPU_OK;
    }
    st = render_device_impl(ctx, camera, opts, rows, out_rgb ? ctx->d_out : nullptr, out_rgb8 ? ctx->d_out8 : nullptr,
                            reinterpret_cast<uint64_t*>(ctx->d_counters), ctx->stream, nullptr, family, false, /*wavefront_blocking=*/false);
    if (st != RTGPU_OK) return st;
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    if ((st = publish_status(ctx)) != RTGPU_OK) return st;
    // Pageable destination: both families stage and copy the same bytes, and a pageable copy (5-10 ms for a 1080p f64
    // frame, +-1 ms from run to run) would drown the difference between the kernels: the calibration times the kernels.
    const bool trial_on_kernels = trial && !pinned;
    if (trial_on_kernels) tune_end(ctx, ctx->stream, key, family);
    // compact band q of this shard -> image band (q * shard_count + shard_index): the bands are periodic in the image,
    // so ONE strided copy moves all complete bands (a call per band cost ~3 µs each: 34 calls per device for 4-row
    // bands of a 1080p frame on 8 devices), and one plain copy the partial last band
    static const bool per_band_copies = getenv("RTGPU_PER_BAND_COPIES") != nullptr;  // A/B switch, off by default
    if (per_band_copies) {
        for (uint32_t k = 0; k < n_rows; k += sel.band_rows) {
            const uint32_t rows_here = std::min(sel.band_rows, n_rows - k);
            const uint32_t y = selected_row(sel, k);
            if (out_rgb)
                CUDA_TRY(cudaMemcpyAsync((char*)out_rgb + (size_t)y * row_rgb, (char*)ctx->d_out + (size_t)k * row_rgb, row_rgb * rows_here,
                                         cudaMemcpyDeviceToHost, ctx->stream));
            if (out_rgb8)
                CUDA_TRY(cudaMemcpyAsync(out_rgb8 + (size_t)y * row_rgb8, ctx->d_out8 + (size_t)k * row_rgb8, row_rgb8 * rows_here, cudaMemcpyDeviceToHost,
                                         ctx->stream));
        }
    } else {
        const uint32_t full_bands = n_rows / sel.band_rows, tail_rows = n_rows % sel.band_rows;
        const size_t first_row = (size_t)sel.shard_index * sel.band_rows, period_rows = (size_t)sel.shard_count * sel.band_rows;
        if (out_rgb) {
            if (full_bands)
                CUDA_TRY(cudaMemcpy2DAsync((char*)out_rgb + first_row * row_rgb, period_rows * row_rgb, ctx->d_out, (size_t)sel.band_rows * row_rgb,
                                           (size_t)sel.band_rows * row_rgb, full_bands, cudaMemcpyDeviceToHost, ctx->stream));
            if (tail_rows)
                CUDA_TRY(cudaMemcpyAsync((char*)out_rgb + (first_row + (size_t)full_bands * period_rows) * row_rgb,
                                         (char*)ctx->d_out + (size_t)full_bands * sel.band_rows * row_rgb, (size_t)tail_rows * row_rgb, cudaMemcpyDeviceToHost,
                                         ctx->stream));
        }
        if (out_rgb8) {
            if (full_bands)
                CUDA_TRY(cudaMemcpy2DAsync(out_rgb8 + first_row * row_rgb8, period_rows * row_rgb8, ctx->d_out8, (size_t)sel.band_rows * row_rgb8,
                                           (size_t)sel.band_rows * row_rgb8, full_bands, cudaMemcpyDeviceToHost, ctx->stream));
            if (tail_rows)
                CUDA_TRY(cudaMemcpyAsync(out_rgb8 + (first_row + (size_t)full_bands * period_rows) * row_rgb8,
                                         ctx->d_out8 + (size_t)full_bands * sel.band_rows * row_rgb8, (size_t)tail_rows * row_rgb8, cudaMemcpyDeviceToHost,
                                         ctx->stream));
        }
    }
    if (trial && !trial_on_kernels) tune_end(ctx, ctx->stream, key, family);  // pinned: the copies belong to this path's cost
    return RTGPU_OK;
}

int finish_host_render(rtgpu_context* ctx, const rtgpu_camera* camera, const rtgpu_opts* opts, const rtgpu_rows* rows, void* out_rgb,
                       uint8_t* out_rgb8, rtgpu_stats* stats) {
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    // wavefront family: if a queue or the node array was too small the buffers have been enlarged: render again
    for (int attempt = 0; ctx->wf_used; ++attempt) {
        RowSel sel;
        int st = normalise_rows(rows, camera->vsize, &sel);
        if (st != RTGPU_OK) return st;
        const uint64_t pixels = (uint64_t)camera->hsize * count_rows(sel, camera->vsize);
        const bool f64 = !opts || opts->precision == RTGPU_PRECISION_F64;
        st = f64 ? wavefront_check<double>(ctx, pixels, ctx->stream, true) : wavefront_check<float>(ctx, pixels, ctx->stream, true);
        if (st == RTGPU_ERR_OUT_OF_MEMORY && requested_family(opts) == FAMILY_AUTO) {
            // the frame needs larger queues than the device can give: the persistent kernel needs none
            tune_rule_out_wavefront(ctx, ctx->tune_last_key);
            st = enqueue_host_render(ctx, camera, opts, rows, out_rgb, out_rgb8, FAMILY_PERSISTENT);
            if (st != RTGPU_OK) return st;
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            break;
        }
        if (st < 0) return st;
        if (st == 0) break;
        if (attempt >= 8) return fail(RTGPU_ERR_OUT_OF_MEMORY, "wavefront buffers kept overflowing");
        st = enqueue_host_render(ctx, camera, opts, rows, out_rgb, out_rgb8, FAMILY_WAVEFRONT);
        if (st != RTGPU_OK) return st;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (stats) {
        if (!ctx->h_status->valid) return fail(RTGPU_ERR_CUDA, "the frame's status block was not published");
        const unsigned long long* c = ctx->h_status->counters;
        stats->rays_primary += c[rt::COUNTER_PRIMARY];
        stats->rays_shadow += c[rt::COUNTER_SHADOW];
        stats->rays_reflect += c[rt::COUNTER_REFLECT];
        stats->rays_refract += c[rt::COUNTER_REFRACT];
        stats->hit_nodes += c[rt::COUNTER_HIT_NODES];
        stats->pixels += c[rt::COUNTER_PIXELS];
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);
    }
    return RTGPU_OK;
}

double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// FMA-chain peak kernels: 8 independent chains per thread
template <typename T>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + T(1), x2 = x0 + T(2), x3 = x0 + T(3), x4 = x0 + T(4), x5 = x0 + T(5), x6 = x0 + T(6), x7 = x0 + T(7);
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b);
        x1 = fma(x1, a, b);
        x2 = fma(x2, a, b);
        x3 = fma(x3, a, b);
        x4 = fma(x4, a, b);
        x5 = fma(x5, a, b);
        x6 = fma(x6, a, b);
        x7 = fma(x7, a, b);
    }
    T s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == T(-12345.678)) out[0] = s;  // never true: keeps the chains alive
}

template <typename T>
int measure_peak(int device, double* out_tflops, double* out_ms) {
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    T* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, sizeof(T)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    fma_peak_kernel<T><<<blocks, threads>>>(d, 1024, (T)0.999, (T)0.001);  // warm-up
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0));
        fma_peak_kernel<T><<<blocks, threads>>>(d, iters, (T)0.999, (T)0.001);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, (double)ms);
    }
    CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    const double flops = 2.0 * 8.0 * (double)iters * (double)threads * (double)blocks;
    if (out_tflops) *out_tflops = flops / (best * 1e-3) / 1e12;
    if (out_ms) *out_ms = best;
    return RTGPU_OK;
}

__global__ void selftest_arith_kernel(const double* a, const double* b, size_t n, unsigned long long* out) {
    unsigned long long div_bad = 0, sqrt_bad = 0, div_fb = 0, sqrt_fb = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = a[i], y = b[i];
        bool ok = true;
        rt::Recip<double> r = rt::recip(y, ok);
        const double q = rt::quot(x, r, ok);
        const double q_native = x / y;
        if (!ok) ++div_fb;
        else if (__double_as_longlong(q) != __double_as_longlong(q_native)) ++div_bad;
        bool ok0 = true;
        rt::Recip<double> r0 = rt::recip(y, ok0);
        const double q0 = rt::quot0(x * 0.0, r0, ok0);  // zero numerators keep the IEEE sign of zero
        if (ok0 && __double_as_longlong(q0) != __double_as_longlong((x * 0.0) / y)) ++div_bad;
        bool oks = true;
        const double s = rt::sqrt_fast(x, oks);
        const double s_native = sqrt(x);
        if (!oks) ++sqrt_fb;
        else if (__double_as_longlong(s) != __double_as_longlong(s_native)) ++sqrt_bad;
    }
    atomicAdd(&out[0], div_bad);
    atomicAdd(&out[1], sqrt_bad);
    atomicAdd(&out[2], div_fb);
    atomicAdd(&out[3], sqrt_fb);
}

// Resident helper threads for the multi-device one-shot call: job g runs on helper g - 1 (job 0 on the caller's
// thread), so a frame costs a wake-up per device instead of a thread creation.
class WorkerPool {
public:
    template <typename F>
    void run(int n_jobs, F&& fn) {
        if (n_jobs <= 1) {
            fn(0);
            return;
        }
        std::function<void(int)> f = fn;
        {
            std::unique_lock<std::mutex> lk(mu_);
            while ((int)threads_.size() < n_jobs - 1) {
                const int index = (int)threads_.size();
                slots_.emplace_back(new Slot());
                threads_.emplace_back([this, index] { loop(index); });
            }
            job_ = &f;
            pending_ = n_jobs - 1;
            for (int k = 0; k < n_jobs - 1; ++k) slots_[k]->go = true;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
    ~WorkerPool() {
        {
            std::unique_lock<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }

private:
    struct Slot {
        bool go = false;
    };
    void loop(int index) {
        for (;;) {
            std::function<void(int)>* job = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || slots_[index]->go; });
                if (stop_) return;
                slots_[index]->go = false;
                job = job_;
            }
            (*job)(index + 1);
            {
                std::unique_lock<std::mutex> lk(mu_);
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> threads_;
    std::vector<std::unique_ptr<Slot>> slots_;
    std::function<void(int)>* job_ = nullptr;
    int pending_ = 0;
    bool stop_ = false;
};

WorkerPool& worker_pool() {
    static WorkerPool* pool = new WorkerPool();  // leaked on purpose: no thread joins during static destruction
    return *pool;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================

extern "C" {

uint32_t rtgpu_abi_version(void) { return RTGPU_ABI_VERSION; }

const char* rtgpu_last_error(void) { return g_last_error.c_str(); }

int rtgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n < 0 ? 0 : n;
}

uint32_t rtgpu_rows_count(const rtgpu_rows* rows, uint32_t vsize) {
    RowSel sel;
    if (normalise_rows(rows, vsize, &sel) != RTGPU_OK) return 0;
    return count_rows(sel, vsize);
}

uint32_t rtgpu_rows_list(const rtgpu_rows* rows, uint32_t vsize, uint32_t* out_rows, uint32_t capacity) {
    RowSel sel;
    if (normalise_rows(rows, vsize, &sel) != RTGPU_OK) return 0;
    const uint32_t n = count_rows(sel, vsize);
    if (out_rows)
        for (uint32_t k = 0; k < n && k < capacity; ++k)
            out_rows[k] = selected_row(sel, k);
    return n;
}

int rtgpu_context_create(const rtgpu_scene* scene, int device, rtgpu_context** out_context) {
    if (!out_context) return fail(RTGPU_ERR_INVALID_ARGUMENT, "out_context is NULL");
    *out_context = nullptr;
    PackedScene packed;
    int st = pack_scene(scene, &packed);
    if (st != RTGPU_OK) return st;
    rtgpu_context* ctx = new rtgpu_context();
    st = context_init(ctx, device);
    if (st == RTGPU_OK) st = upload_scene(ctx, packed);
    if (st != RTGPU_OK) {
        std::string keep = g_last_error;
        context_release(ctx);
        g_last_error = keep;
        return st;
    }
    *out_context = ctx;
    return RTGPU_OK;
}

void rtgpu_context_destroy(rtgpu_context* context) { context_release(context); }

int rtgpu_context_render_device(rtgpu_context* context, const rtgpu_camera* camera, const rtgpu_opts* opts,
                                const rtgpu_rows* rows, void* d_out_rgb, uint8_t* d_out_rgb8, uint64_t* d_counters,
                                void* cuda_stream) {
    if (!context || !camera) return fail(RTGPU_ERR_INVALID_ARGUMENT, "context or camera is NULL");
    uint32_t precision, max_depth;
    int st = check_opts(opts, &precision, &max_depth);
    if (st != RTGPU_OK) return st;
    RowSel sel;
    st = normalise_rows(rows, camera->vsize, &sel);
    if (st != RTGPU_OK) return st;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    bool trial = false;
    const uint64_t key = tune_key(context, camera, sel, precision, max_depth, 0u);
    CUDA_TRY(cudaSetDevice(context->device));
    const int family = resolve_family(context, opts, key, &trial);
    if (trial) tune_begin(context, stream);
    st = render_device_impl(context, camera, opts, rows, d_out_rgb, d_out_rgb8, d_counters, stream, nullptr, family);
    if (st == RTGPU_ERR_OUT_OF_MEMORY && family == FAMILY_WAVEFRONT && requested_family(opts) == FAMILY_AUTO) {
        tune_rule_out_wavefront(context, key);  // no room for the queues: the persistent kernel needs none
        return render_device_impl(context, camera, opts, rows, d_out_rgb, d_out_rgb8, d_counters, stream, nullptr, FAMILY_PERSISTENT);
    }
    if (trial && st == RTGPU_OK) tune_end(context, stream, key, family);
    return st;
}

int rtgpu_last_family(void) { return g_last_family; }

uint64_t rtgpu_context_launch_count(rtgpu_context* context) { return context ? context->launches : 0u; }

int rtgpu_context_frame_records(rtgpu_context* context, uint64_t out[4]) {
    if (!context || !out) return fail(RTGPU_ERR_INVALID_ARGUMENT, "context or out is NULL");
    if (!context->h_status || !context->h_status->valid) return fail(RTGPU_ERR_INVALID_ARGUMENT, "no host-buffer frame has been rendered on this context yet");
    const bool f32 = context->d_reals32 != nullptr && context->wf_cap_rays && context->wf_bytes_rays / context->wf_cap_rays == sizeof(rt::WfRay<float>);
    out[0] = context->h_status->queued_rays;
    out[1] = context->h_status->n_nodes;
    out[2] = f32 ? sizeof(rt::WfRay<float>) : sizeof(rt::WfRay<double>);
    out[3] = f32 ? sizeof(rt::WfNode<float>) : sizeof(rt::WfNode<double>);
    return RTGPU_OK;
}

int rtgpu_context_render(rtgpu_context* context, const rtgpu_camera* camera, const rtgpu_opts* opts, const rtgpu_rows* rows,
                         void* out_rgb, uint8_t* out_rgb8, rtgpu_stats* stats) {
    if (!context || !camera) return fail(RTGPU_ERR_INVALID_ARGUMENT, "context or camera is NULL");
    if (!out_rgb && !out_rgb8) return fail(RTGPU_ERR_INVALID_ARGUMENT, "both output pointers are NULL");
    const double t0 = wall_ms();
    if (stats) memset(stats, 0, sizeof(*stats));
    int st = enqueue_host_render(context, camera, opts, rows, out_rgb, out_rgb8);
    if (st != RTGPU_OK) return st;
    st = finish_host_render(context, camera, opts, rows, out_rgb, out_rgb8, stats);
    if (st != RTGPU_OK) return st;
    if (stats) stats->total_ms = wall_ms() - t0;
    return RTGPU_OK;
}

int rtgpu_render(const rtgpu_scene* scene, const rtgpu_camera* camera, const rtgpu_opts* opts, void* out_rgb,
                 uint8_t* out_rgb8, rtgpu_stats* stats) {
    if (!camera) return fail(RTGPU_ERR_INVALID_ARGUMENT, "camera is NULL");
    if (!out_rgb && !out_rgb8) return fail(RTGPU_ERR_INVALID_ARGUMENT, "both output pointers are NULL");
    const double t0 = wall_ms();
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!scene) return fail(RTGPU_ERR_INVALID_ARGUMENT, "scene is NULL");
    int st = RTGPU_OK;
    // One resident scene per device, kept for the lifetime of the process so that repeated frames
    // pay for neither cudaMalloc / stream creation nor — when the scene description is byte-for-byte the one the
    // device already holds — for validation, packing (BVH build) and upload.
    static std::mutex mu;
    static std::vector<rtgpu_context*> cache;
    std::lock_guard<std::mutex> lock(mu);
    int n_gpus = (opts && opts->n_gpus > 0) ? opts->n_gpus : 1;
    const uint64_t input_hash = scene_input_hash(scene);
    bool need_pack = (int)cache.size() < n_gpus;
    for (int g = 0; g < n_gpus && !need_pack; ++g) need_pack = !cache[g] || cache[g]->uploaded_input_hash != input_hash;
    PackedScene packed;
    if (need_pack) {
        st = pack_scene(scene, &packed);  // validates: a malformed scene is reported before anything touches a device
        if (st != RTGPU_OK) return st;
    }
    const int available = rtgpu_device_count();
    if (available <= 0) return fail(RTGPU_ERR_NO_DEVICE, "no CUDA device available (gpu mode has no CPU fallback)");
    if (n_gpus > available) return fail(RTGPU_ERR_INVALID_ARGUMENT, "opts.n_gpus %d > %d visible devices", n_gpus, available);
    uint32_t band_rows = (opts && opts->band_rows) ? opts->band_rows : 16u;
    band_rows = (band_rows + 3u) & ~3u;  // whole 8x4 tiles per band
    if ((int)cache.size() < n_gpus) cache.resize(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; ++g) {
        if (!cache[g]) {
            rtgpu_context* ctx = new rtgpu_context();
            st = context_init(ctx, g);
            if (st != RTGPU_OK) {
                std::string keep = g_last_error;
                context_release(ctx);
                g_last_error = keep;
                return st;
            }
            cache[g] = ctx;
        }
    }
    // row bands: device g renders bands g, g+G, g+2G, ... (interleaved: per-row cost varies a lot).
    // One host thread per device (the caller's for device 0, resident workers for the others): upload, launches,
    // copies and the final wait of the devices proceed side by side instead of queueing behind one another.
    struct DeviceJob {
        int status = RTGPU_OK;
        std::string error;
        rtgpu_stats stats{};
        double t_start = 0, t_enqueued = 0, t_done = 0;  // ms since the call began (RTGPU_TRACE=1 prints them)
    };
    static const bool trace = getenv("RTGPU_TRACE") != nullptr;
    std::vector<DeviceJob> jobs(n_gpus);
    auto work = [&](int g) {
        DeviceJob& job = jobs[g];
        job.t_start = wall_ms() - t0;
        rtgpu_rows rows;
        rows.band_rows = n_gpus == 1 ? 0u : band_rows;
        rows.shard_index = (uint32_t)g;
        rows.shard_count = (uint32_t)n_gpus;
        int rc = cudaSetDevice(g) == cudaSuccess ? RTGPU_OK : fail(RTGPU_ERR_CUDA, "cudaSetDevice(%d) failed", g);
        if (rc == RTGPU_OK && cache[g]->uploaded_input_hash != input_hash) {
            cache[g]->uploaded_input_hash = 0;
            rc = upload_scene(cache[g], packed);
            if (rc == RTGPU_OK) cache[g]->uploaded_input_hash = input_hash;
        }
        if (rc == RTGPU_OK) rc = enqueue_host_render(cache[g], camera, opts, &rows, out_rgb, out_rgb8);
        job.t_enqueued = wall_ms() - t0;
        if (rc == RTGPU_OK) rc = finish_host_render(cache[g], camera, opts, &rows, out_rgb, out_rgb8, stats ? &job.stats : nullptr);
        job.t_done = wall_ms() - t0;
        job.status = rc;
        if (rc != RTGPU_OK) job.error = g_last_error;
    };
    const double t_dispatch = wall_ms() - t0;
    worker_pool().run(n_gpus, work);
    if (trace) {
        fprintf(stderr, "[rtgpu] one-shot: dispatch at %.3f ms, joined at %.3f ms\n", t_dispatch, wall_ms() - t0);
        for (int g = 0; g < n_gpus; ++g)
            fprintf(stderr, "[rtgpu]   device %d: start %.3f enqueued %.3f done %.3f kernels %.3f ms\n", g, jobs[g].t_start, jobs[g].t_enqueued, jobs[g].t_done,
                    jobs[g].stats.kernel_ms);
    }
    for (int g = 0; g < n_gpus; ++g) {
        if (jobs[g].status != RTGPU_OK) {
            g_last_error = jobs[g].error;
            return jobs[g].status;
        }
        if (stats) {
            stats->rays_primary += jobs[g].stats.rays_primary;
            stats->rays_shadow += jobs[g].stats.rays_shadow;
            stats->rays_reflect += jobs[g].stats.rays_reflect;
            stats->rays_refract += jobs[g].stats.rays_refract;
            stats->hit_nodes += jobs[g].stats.hit_nodes;
            stats->pixels += jobs[g].stats.pixels;
            stats->kernel_ms = std::max(stats->kernel_ms, jobs[g].stats.kernel_ms);
        }
    }
    g_last_family = cache[0]->last_family;
    if (stats) stats->total_ms = wall_ms() - t0;
    return RTGPU_OK;
}

int rtgpu_debug_probe(rtgpu_context* context, const rtgpu_camera* camera, uint32_t kind, const double* in, size_t n_in, double* out,
                      size_t n_out) {
    if (!context || !out || n_out == 0 || (n_in && !in)) return fail(RTGPU_ERR_INVALID_ARGUMENT, "context, in or out is NULL");
    if (kind == RTGPU_PROBE_RAY_FOR_PIXEL && !camera) return fail(RTGPU_ERR_INVALID_ARGUMENT, "this probe needs a camera");
    if (context->layout.n_bvh_nodes > 0 && (kind == RTGPU_PROBE_IN_SHADOW || kind == RTGPU_PROBE_PREPARE))
        return fail(RTGPU_ERR_UNSUPPORTED, "scene queries through the probes cover uniform shape lists only (this scene uses a BVH)");
    CUDA_TRY(cudaSetDevice(context->device));
    rt::CameraParams<double> cam;
    memset(&cam, 0, sizeof(cam));
    if (camera) {
        RowSel whole{camera->vsize ? camera->vsize : 1u, 0u, 1u};
        fill_camera(camera, whole, camera->vsize, 6u, false, &cam);
    }
    double *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, std::max<size_t>(1, n_in) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_out, n_out * sizeof(double)));
    if (n_in) CUDA_TRY(cudaMemcpy(d_in, in, n_in * sizeof(double), cudaMemcpyHostToDevice));
    rt::SceneLayout lay = context->layout;
    lay.in_shared = 0u;
    rt::probe_kernel<double><<<1, 32, 0, context->stream>>>(context->d_reals64, context->d_ints, lay, cam, (int)kind, d_in, (int)n_in, d_out, (int)n_out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(context->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n_out * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(RTGPU_ERR_CUDA, "probe kernel failed: %s", cudaGetErrorString(e));
    return RTGPU_OK;
}

int rtgpu_debug_color_at(rtgpu_context* context, const double origin[3], const double direction[3], const rtgpu_opts* opts, double out_rgb[3]) {
    if (!context || !origin || !direction || !out_rgb) return fail(RTGPU_ERR_INVALID_ARGUMENT, "NULL argument");
    uint32_t precision, max_depth;
    int st = check_opts(opts, &precision, &max_depth);
    if (st != RTGPU_OK) return st;
    if (precision != RTGPU_PRECISION_F64) return fail(RTGPU_ERR_UNSUPPORTED, "colour probes run in the f64 parity mode");
    CUDA_TRY(cudaSetDevice(context->device));
    // a 1 x 1 frame whose only ray is the caller's, through the unmodified kernels of the requested family
    rt::CameraParams<double> cam;
    memset(&cam, 0, sizeof(cam));
    cam.hsize = cam.vsize = cam.n_rows = cam.band_rows = 1u;
    cam.shard_count = cam.band_take = 1u;
    cam.max_depth = max_depth;
    cam.tile_stride = 0u;
    cam.probe_ray = 1u;
    for (int k = 0; k < 3; ++k) {
        cam.origin[k] = origin[k];
        cam.inv[k] = direction[k];
    }
    st = ensure_out_buffers(context, 3 * sizeof(double), 0);
    if (st != RTGPU_OK) return st;
    const int family = requested_family(opts) == FAMILY_WAVEFRONT ? FAMILY_WAVEFRONT : FAMILY_PERSISTENT;
    context->last_family = g_last_family = family;
    if (family == FAMILY_WAVEFRONT) st = render_wavefront<double>(context, context->d_reals64, cam, (double*)context->d_out, nullptr, nullptr, context->stream, true);
    else if (max_depth <= 7) st = launch_kernel<double, 8>(context, context->d_reals64, cam, (double*)context->d_out, nullptr, nullptr, context->stream);
    else st = launch_kernel<double, 16>(context, context->d_reals64, cam, (double*)context->d_out, nullptr, nullptr, context->stream);
    if (st != RTGPU_OK) return st;
    CUDA_TRY(cudaMemcpyAsync(out_rgb, context->d_out, 3 * sizeof(double), cudaMemcpyDeviceToHost, context->stream));
    CUDA_TRY(cudaStreamSynchronize(context->stream));
    return RTGPU_OK;
}

void* rtgpu_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (rtgpu_device_count() <= 0) {
        fail(RTGPU_ERR_NO_DEVICE, "no CUDA device available");
        return nullptr;
    }
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        fail(RTGPU_ERR_OUT_OF_MEMORY, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void rtgpu_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

int rtgpu_measure_fma_peak(int device, uint32_t precision, double* out_tflops, double* out_ms) {
    if (rtgpu_device_count() <= 0) return fail(RTGPU_ERR_NO_DEVICE, "no CUDA device available");
    if (precision == RTGPU_PRECISION_F64) return measure_peak<double>(device, out_tflops, out_ms);
    if (precision == RTGPU_PRECISION_F32) return measure_peak<float>(device, out_tflops, out_ms);
    return fail(RTGPU_ERR_INVALID_ARGUMENT, "precision %u", precision);
}

int rtgpu_selftest_arith(int device, const double* a, const double* b, size_t n, uint64_t* out_div_mismatches,
                         uint64_t* out_sqrt_mismatches, uint64_t* out_div_fallbacks, uint64_t* out_sqrt_fallbacks) {
    if (!a || !b) return fail(RTGPU_ERR_INVALID_ARGUMENT, "operand array is NULL");
    if (rtgpu_device_count() <= 0) return fail(RTGPU_ERR_NO_DEVICE, "no CUDA device available");
    CUDA_TRY(cudaSetDevice(device));
    double *da = nullptr, *db = nullptr;
    unsigned long long* dout = nullptr;
    CUDA_TRY(cudaMalloc(&da, std::max<size_t>(8, n * sizeof(double))));
    CUDA_TRY(cudaMalloc(&db, std::max<size_t>(8, n * sizeof(double))));
    CUDA_TRY(cudaMalloc(&dout, 4 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemcpy(da, a, n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(db, b, n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(dout, 0, 4 * sizeof(unsigned long long)));
    selftest_arith_kernel<<<1184, 256>>>(da, db, n, dout);
    CUDA_TRY(cudaGetLastError());
    unsigned long long h[4];
    CUDA_TRY(cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    if (out_div_mismatches) *out_div_mismatches = h[0];
    if (out_sqrt_mismatches) *out_sqrt_mismatches = h[1];
    if (out_div_fallbacks) *out_div_fallbacks = h[2];
    if (out_sqrt_fallbacks) *out_sqrt_fallbacks = h[3];
    return RTGPU_OK;
}

}  // extern "C"
