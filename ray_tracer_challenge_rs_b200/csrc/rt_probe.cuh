// rt_probe.cuh — known-answer probes: ONE device thread runs the device functions the render kernels are built
// from (local_intersect, normals, patterns, Phong lighting, the hit rule, the refraction-container pass, Schlick) on
// caller-supplied inputs, so that the reference's own unit tests can be replayed on the GPU
// bit for bit, not only through whole frames.  Test infrastructure behind rtgpu_debug_probe; not on the render path.
#pragma once

#include "rt_kernel.cuh"

namespace rt {

// kinds = rtgpu_probe_kind (include/rtgpu.h); `in` / `out` are arrays of doubles (integers travel as doubles)
enum : int {
    PROBE_RAY_FOR_PIXEL = 0, PROBE_INTERSECT = 1, PROBE_LOCAL_NORMAL = 2, PROBE_NORMAL = 3, PROBE_PATTERN = 4,
    PROBE_LIGHTING = 5, PROBE_IN_SHADOW = 6, PROBE_PREPARE = 7, PROBE_QUANTISE = 8, PROBE_COLLECT = 9,
};

template <typename T>
RT_DEV int probe_find(const SceneView<T, false>& sv, int world_index) {
    for (uint32_t pos = 0; pos < sv.L.n_shapes; ++pos)
        if (sv.shape_meta(pos).x == world_index) return (int)pos;
    return -1;
}

// Ray::intersect (ray.rs:35-49) of the shape at sorted position pos
template <typename T>
RT_DEV int probe_intersect(const SceneView<T, false>& sv, uint32_t pos, const Ray<T>& ray, T* t) {
    const T* g = sv.shape(pos);
    const int4 meta = sv.shape_meta(pos);
    Ray<T> local;
    local.o = mat_point(g, ray.o);
    local.d = mat_vector(g, ray.d);
    switch ((meta.z >> FLAG_TYPE_SHIFT) & 7) {
    case 0: return local_intersect<T, 0>(local, g, meta.z, nullptr, t[0], t[1], t[2], t[3]);
    case 1: return local_intersect<T, 1>(local, g, meta.z, nullptr, t[0], t[1], t[2], t[3]);
    case 2: return local_intersect<T, 2>(local, g, meta.z, nullptr, t[0], t[1], t[2], t[3]);
    case 3: return local_intersect<T, 3>(local, g, meta.z, nullptr, t[0], t[1], t[2], t[3]);
    case 4: return local_intersect<T, 4>(local, g, meta.z, nullptr, t[0], t[1], t[2], t[3]);
    default: return local_intersect<T, 5>(local, g, meta.z, sv.triangle(pos), t[0], t[1], t[2], t[3]);
    }
}

template <typename T>
__global__ void probe_kernel(const T* __restrict__ g_reals, const int* __restrict__ g_ints, SceneLayout layout, CameraParams<T> cam, int kind,
                             const double* __restrict__ in, int n_in, double* __restrict__ out, int n_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    SceneView<T, false> sv;
    sv.L = layout;
    sv.reals = g_reals;
    sv.ints = g_ints;
    for (int k = 0; k < n_out; ++k) out[k] = 0.0;
    auto v3 = [&](int at) { return mk<T>((T)in[at], (T)in[at + 1], (T)in[at + 2]); };
    auto put3 = [&](int at, V3<T> v) { out[at] = (double)v.x; out[at + 1] = (double)v.y; out[at + 2] = (double)v.z; };
    switch (kind) {
    case PROBE_RAY_FOR_PIXEL: {  // in: px, py
        const Ray<T> r = camera_ray(cam, (uint32_t)in[0], (uint32_t)in[1]);
        put3(0, r.o);
        put3(3, r.d);
        break;
    }
    case PROBE_INTERSECT: {  // in: shape, origin, direction -> n, t[4]
        const int pos = probe_find(sv, (int)in[0]);
        if (pos < 0) break;
        Ray<T> ray;
        ray.o = v3(1);
        ray.d = v3(4);
        T t[4] = {T(0), T(0), T(0), T(0)};
        const int n = probe_intersect(sv, (uint32_t)pos, ray, t);
        out[0] = n;
        for (int k = 0; k < 4; ++k) out[1 + k] = (double)t[k];
        break;
    }
    case PROBE_LOCAL_NORMAL:
    case PROBE_NORMAL: {  // in: shape, point
        const int pos = probe_find(sv, (int)in[0]);
        if (pos < 0) break;
        const T* g = sv.shape((uint32_t)pos);
        const int type = (sv.shape_meta((uint32_t)pos).z >> FLAG_TYPE_SHIFT) & 7;
        put3(0, kind == PROBE_LOCAL_NORMAL ? local_normal_at(sv, (uint32_t)pos, type, g, v3(1)) : world_normal_at(sv, (uint32_t)pos, type, g, v3(1)));
        break;
    }
    case PROBE_PATTERN: {  // in: pattern, shape, world point: Pattern::color_at_shape, pattern.rs:10-14
        const int pos = probe_find(sv, (int)in[1]);
        if (pos < 0) break;
        V3<T> object_point = mat_point(sv.shape((uint32_t)pos), v3(2));
        V3<T> pattern_point = mat_point(sv.pattern((uint32_t)in[0]) + 6, object_point);
        put3(0, pattern_color_at(sv, (int)in[0], pattern_point));
        break;
    }
    case PROBE_LIGHTING: {  // in: material, shape, light position, light intensity, point, eye, normal, in_shadow
        const int pos = probe_find(sv, (int)in[1]);
        if (pos < 0) break;
        const uint32_t material = (uint32_t)in[0];
        const V3<T> point = v3(8);
        const V3<T> base = resolve_color(sv, material, (uint32_t)pos, point);
        const V3<T> light_dir = normalized(v3(2) - point);  // material.rs:88
        put3(0, phong_lighting(sv.material(material), base, v3(5), light_dir, v3(11), v3(14), in[17] != 0.0));
        break;
    }
    case PROBE_IN_SHADOW: {  // in: light, point: World::is_in_shadow, world.rs:98-112
        const V3<T> point = v3(1);
        Normalized<T> nl = normalize_full(ld3(sv.light((uint32_t)in[0])) - point);
        Ray<T> ray;
        ray.o = point;
        ray.d = nl.v;
        TraceAcc<T> acc;
        ContainerAcc<T> cacc;
        acc.c = &cacc;
        acc.mode = MODE_SHADOW;
        acc.best_t = nl.magnitude;
        acc.dir_sq = fma(ray.d.z, ray.d.z, fma(ray.d.y, ray.d.y, ray.d.x * ray.d.x));
        acc.best_orig = -1;
        acc.best_pos = -1;
        trace_unified<T, true, false>(sv, ray, acc);
        out[0] = acc.best_pos >= 0 ? 1.0 : 0.0;
        break;
    }
    case PROBE_COLLECT: {  // in: origin, direction -> count, then (t, shape) pairs in world push order (world.rs:25-33)
        Ray<T> ray;
        ray.o = v3(0);
        ray.d = v3(3);
        int count = 0;
        for (int world = 0; world < (int)sv.L.n_shapes; ++world) {
            const int pos = probe_find(sv, world);
            if (pos < 0) continue;
            T t[4] = {T(0), T(0), T(0), T(0)};
            const int n = probe_intersect(sv, (uint32_t)pos, ray, t);
            for (int k = 0; k < n; ++k) {
                if (1 + 2 * count + 1 < n_out) {
                    out[1 + 2 * count] = (double)t[k];
                    out[2 + 2 * count] = world;
                }
                ++count;
            }
        }
        out[0] = count;
        break;
    }
    case PROBE_PREPARE: {
        // in: origin, direction, k (-1 = Intersections::hit), n, then n x (t, shape) as the caller's (sorted) list.
        // out: found, distance, shape, inside, point, over, under, eye, normal, reflect, n1, n2, schlick
        Ray<T> ray;
        ray.o = v3(0);
        ray.d = v3(3);
        const int k = (int)in[6], n = (int)in[7];
        T t_hit = T(0);
        int world = -1;
        if (k >= 0) {
            if (k >= n) break;
            t_hit = (T)in[8 + 2 * k];
            world = (int)in[9 + 2 * k];
        } else {
            // Intersections::hit (intersections.rs:13-18) through the kernels' own rule: the running
            // (distance, order) minimum of consume(), one list entry at a time
            TraceAcc<T> acc;
            ContainerAcc<T> cacc;
            acc.c = &cacc;
            acc.mode = MODE_RADIANCE;
            acc.best_t = Real<T>::max();
            acc.dir_sq = T(1);
            acc.best_orig = 0x7fffffff;
            acc.best_pos = -1;
            for (int e = 0; e < n; ++e) {
                int4 meta = make_int4(e, 0, FLAG_CASTS_SHADOW, 0);  // order in the (stable-sorted) list
                consume<T, 4>(acc, 1, (T)in[8 + 2 * e], T(0), T(0), T(0), e, meta);
            }
            if (acc.best_pos < 0) break;
            t_hit = acc.best_t;
            world = (int)in[9 + 2 * acc.best_pos];
        }
        const int pos = probe_find(sv, world);
        if (pos < 0) break;
        const T* g = sv.shape((uint32_t)pos);
        const int4 meta = sv.shape_meta((uint32_t)pos);
        const uint32_t material = (uint32_t)meta.y;
        // Intersection::prepare_computations, intersection.rs:21-31; computed_hit.rs:33-34
        const V3<T> point = ray.o + ray.d * t_hit;
        V3<T> normal = world_normal_at(sv, (uint32_t)pos, (meta.z >> FLAG_TYPE_SHIFT) & 7, g, point);
        const V3<T> eye = neg(ray.d);
        const bool inside = dot(normal, eye) < T(0);
        if (inside) normal = neg(normal);
        const T off = Real<T>::offset(point.x, point.y, point.z, t_hit);
        const V3<T> over = point + (normal * off), under = point - (normal * off);
        // refraction containers (intersection.rs:33-62): the kernels' list-free bookkeeping — per shape, the parity of
        // its intersections before the hit and the last of them — fed with the CALLER's list, shape by shape, exactly
        // as the shape loop feeds it with what local_intersect returns.  (A hand-built list need not agree with the
        // geometry to the last bit: the reference's own tests use rounded distances.)
        TraceAcc<T> acc;
        ContainerAcc<T> cacc;
        acc.c = &cacc;
        acc.mode = MODE_CONTAINER;
        acc.best_t = Real<T>::max();
        acc.dir_sq = fma(ray.d.z, ray.d.z, fma(ray.d.y, ray.d.y, ray.d.x * ray.d.x));
        acc.best_orig = 0x7fffffff;
        acc.best_pos = -1;
        cacc.t_hit = t_hit;
        cacc.hit_class = meta.w;
        cacc.hit_class_inside = false;
        cacc.all_pos = cacc.excl_pos = -1;
        cacc.all_t = cacc.excl_t = T(0);
        cacc.all_orig = cacc.excl_orig = 0;
        for (uint32_t p2 = 0; p2 < sv.L.n_shapes; ++p2) {
            const int4 m2 = sv.shape_meta(p2);
            T ts[4] = {T(0), T(0), T(0), T(0)};
            int cnt = 0;
            for (int e = 0; e < n && cnt < 4; ++e)
                if ((int)in[9 + 2 * e] == m2.x) ts[cnt++] = (T)in[8 + 2 * e];
            if (cnt) consume<T, 4>(acc, cnt, ts[0], ts[1], ts[2], ts[3], (int)p2, m2);
        }
        T n1 = (cacc.all_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)cacc.all_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
        T n2;
        if (cacc.hit_class_inside)
            n2 = (cacc.excl_pos >= 0) ? sv.material((uint32_t)sv.shape_meta((uint32_t)cacc.excl_pos).y)[MAT_REFRACTIVE_INDEX] : T(1);
        else
            n2 = sv.material(material)[MAT_REFRACTIVE_INDEX];
        out[0] = 1.0;
        out[1] = (double)t_hit;
        out[2] = world;
        out[3] = inside ? 1.0 : 0.0;
        put3(4, point);
        put3(7, over);
        put3(10, under);
        put3(13, eye);
        put3(16, normal);
        put3(19, reflect(ray.d, normal));  // intersection.rs:31
        out[22] = (double)n1;
        out[23] = (double)n2;
        out[24] = (double)schlick_reflectance(n1, n2, dot(eye, normal));
        break;
    }
    case PROBE_QUANTISE: out[0] = quantise((T)in[0]); break;
    default: break;
    }
}

}  // namespace rt
