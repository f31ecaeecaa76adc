/*
 * rt_oracle.c — CPU oracle for the camera render pass.  TEST INFRASTRUCTURE, NOT PRODUCT
 * (see rt_oracle.h for who may call it).  Parity PINNED against the reference's known-answer
 * tests and golden renders (tests/test_oracle_kat.py, tests/test_oracle_golden.py).
 *
 * Restates, operation for operation, the reference's per-pixel algorithm.  All citations are
 * paths under the reference tree `ray-tracer/src/` unless they start with `ray-tracer-cli/`.
 * Rust never contracts `a*b+c`, and `mul_add` is a true fused multiply-add, so:
 *   - this file must be built with -ffp-contract=off;
 *   - `fma()` appears exactly where the reference says `mul_add`.
 * The structure deliberately mirrors the reference (materialised intersection list, stable sort,
 * recursive colour evaluation, literal container walk) — it is NOT how the device path works,
 * which is what makes it an independent check.
 */
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RTO_EPSILON 0.00000008 /* consts.rs:2 */
#define RTO_MAX DBL_MAX        /* consts.rs:6  (f64::MAX) */
#define RTO_MIN (-DBL_MAX)     /* consts.rs:4  (f64::MIN) */
#define RTO_DEFAULT_RI 1.0     /* composites/material.rs:24 */

typedef struct v3 {
    double x, y, z;
} v3;

static inline v3 v3_make(double x, double y, double z) {
    v3 r = {x, y, z};
    return r;
}
static inline v3 v3_from(const double *p) { return v3_make(p[0], p[1], p[2]); }
static inline void v3_store(double *p, v3 a) {
    p[0] = a.x;
    p[1] = a.y;
    p[2] = a.z;
}
/* primitives/vector.rs:143-186, point.rs (componentwise operators) */
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_scale(v3 a, double s) { return v3_make(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); } /* color.rs Mul<Color> */

/* utils.rs:27-32 */
static inline double squared(double v) { return v * v; }

/* vector.rs:93-95: z.mul_add(rz, x.mul_add(rx, y * ry)) */
static inline double v3_dot(v3 a, v3 b) { return fma(a.z, b.z, fma(a.x, b.x, a.y * b.y)); }

/* vector.rs:97-103 */
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(fma(a.y, b.z, -a.z * b.y), fma(a.z, b.x, -a.x * b.z), fma(a.x, b.y, -a.y * b.x));
}

/* vector.rs:84-86 */
static inline double v3_magnitude(v3 a) { return sqrt(squared(a.x) + squared(a.y) + squared(a.z)); }

/* vector.rs:88-91: divides each component */
static inline v3 v3_normalized(v3 a) {
    double m = v3_magnitude(a);
    return v3_make(a.x / m, a.y / m, a.z / m);
}

/* vector.rs:105-107: self - (normal * 2.0 * self.dot(normal)) */
static inline v3 v3_reflect(v3 v, v3 n) { return v3_sub(v, v3_scale(v3_scale(n, 2.0), v3_dot(v, n))); }

/* matrix.rs:332-346: fold from 0.0 over [x, y, z, 1.0], rows 0..2 */
static inline v3 mat_point(const double *m, v3 p) {
    v3 r;
    r.x = (((0.0 + m[0] * p.x) + m[1] * p.y) + m[2] * p.z) + m[3] * 1.0;
    r.y = (((0.0 + m[4] * p.x) + m[5] * p.y) + m[6] * p.z) + m[7] * 1.0;
    r.z = (((0.0 + m[8] * p.x) + m[9] * p.y) + m[10] * p.z) + m[11] * 1.0;
    return r;
}

/* matrix.rs:348-362: fold from 0.0 over [x, y, z, 0.0] */
static inline v3 mat_vector(const double *m, v3 v) {
    v3 r;
    r.x = (((0.0 + m[0] * v.x) + m[1] * v.y) + m[2] * v.z) + m[3] * 0.0;
    r.y = (((0.0 + m[4] * v.x) + m[5] * v.y) + m[6] * v.z) + m[7] * 0.0;
    r.z = (((0.0 + m[8] * v.x) + m[9] * v.y) + m[10] * v.z) + m[11] * 0.0;
    return r;
}

/* shapes/shape.rs:25: transformation_inverse().transpose() * local_normal.  Row r of the transpose
 * is column r of the inverse; its 4th entry is inverse[3][r] = 0.0 for every affine transform (the
 * flattened scene carries rows 0..2 only), times the vector's w = 0.0. */
static inline v3 mat_transposed_vector(const double *m, v3 v) {
    v3 r;
    r.x = (((0.0 + m[0] * v.x) + m[4] * v.y) + m[8] * v.z) + 0.0 * 0.0;
    r.y = (((0.0 + m[1] * v.x) + m[5] * v.y) + m[9] * v.z) + 0.0 * 0.0;
    r.z = (((0.0 + m[2] * v.x) + m[6] * v.y) + m[10] * v.z) + 0.0 * 0.0;
    return r;
}

typedef struct ray_t {
    v3 origin, direction;
} ray_t;

/* ray.rs:30-32 */
static inline v3 ray_position(const ray_t *r, double t) { return v3_add(r->origin, v3_scale(r->direction, t)); }

/* ---------------------------------------------------------------------------------------- */
/* Intersections (composites/intersections.rs:6, a Vec<Intersection>)                         */

typedef struct isect_t {
    double distance;
    uint32_t shape;
} isect_t;

typedef struct isect_list {
    isect_t *v;
    size_t n, cap;
    isect_t *tmp; /* merge-sort scratch */
    size_t tmp_cap;
} isect_list;

static void list_push(isect_list *l, double t, uint32_t shape) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 64;
        l->v = (isect_t *)realloc(l->v, l->cap * sizeof(isect_t));
    }
    l->v[l->n].distance = t;
    l->v[l->n].shape = shape;
    l->n++;
}

static void list_free(isect_list *l) {
    free(l->v);
    free(l->tmp);
    memset(l, 0, sizeof(*l));
}

/* intersection.rs:91-101: Ord = partial_cmp on distance, incomparable (NaN) -> Equal.
 * `a sorts strictly after b` */
static inline int isect_greater(const isect_t *a, const isect_t *b) { return a->distance > b->distance; }

/* world.rs:34: Vec::sort — a STABLE sort.  Insertion sort for short lists, top-down merge above. */
static void list_sort_range(isect_t *v, isect_t *tmp, size_t n) {
    if (n <= 24) {
        for (size_t i = 1; i < n; i++) {
            isect_t key = v[i];
            size_t j = i;
            while (j > 0 && isect_greater(&v[j - 1], &key)) {
                v[j] = v[j - 1];
                j--;
            }
            v[j] = key;
        }
        return;
    }
    size_t h = n / 2;
    list_sort_range(v, tmp, h);
    list_sort_range(v + h, tmp, n - h);
    size_t i = 0, j = h, k = 0;
    while (i < h && j < n) {
        if (isect_greater(&v[i], &v[j])) tmp[k++] = v[j++];
        else tmp[k++] = v[i++];
    }
    while (i < h) tmp[k++] = v[i++];
    while (j < n) tmp[k++] = v[j++];
    memcpy(v, tmp, n * sizeof(isect_t));
}

static void list_sort(isect_list *l) {
    if (l->n > 24 && l->tmp_cap < l->n) {
        l->tmp_cap = l->cap;
        l->tmp = (isect_t *)realloc(l->tmp, l->tmp_cap * sizeof(isect_t));
    }
    list_sort_range(l->v, l->tmp, l->n);
}

/* ---------------------------------------------------------------------------------------- */
/* local_intersect of the six shapes                                                          */

/* utils.rs:47-57 */
static inline int solve_quadratic(double a, double b, double c, double *s1, double *s2) {
    double discriminant = fma(4.0 * a, -c, squared(b));
    if (discriminant < 0.0) return 0;
    double double_a = 2.0 * a;
    double discriminant_root = sqrt(discriminant);
    *s1 = (-b - discriminant_root) / double_a;
    *s2 = (-b + discriminant_root) / double_a;
    return 1;
}

/* shapes/sphere.rs:41-53 */
static void sphere_local_intersect(const ray_t *ray, uint32_t shape, isect_list *out) {
    v3 sphere_to_ray = ray->origin;
    double a = v3_dot(ray->direction, ray->direction);
    double b = 2.0 * v3_dot(ray->direction, sphere_to_ray);
    double c = v3_dot(sphere_to_ray, sphere_to_ray) - 1.0;
    double t1, t2;
    if (solve_quadratic(a, b, c, &t1, &t2)) {
        list_push(out, t1, shape);
        list_push(out, t2, shape);
    }
}

/* shapes/plane.rs:42-48 */
static void plane_local_intersect(const ray_t *ray, uint32_t shape, isect_list *out) {
    if (fabs(ray->direction.y) < RTO_EPSILON) return;
    double distance = -ray->origin.y / ray->direction.y;
    list_push(out, distance, shape);
}

/* shapes/cube.rs:22-43 */
static void cube_check_axis(double origin, double direction, double *tmin, double *tmax) {
    double distance_min_numerator = -1.0 - origin;
    double distance_max_numerator = 1.0 - origin;
    double distance_min, distance_max;
    if (fabs(direction) >= RTO_EPSILON) {
        distance_min = distance_min_numerator / direction;
        distance_max = distance_max_numerator / direction;
    } else {
        distance_min = distance_min_numerator * RTO_MAX;
        distance_max = distance_max_numerator * RTO_MAX;
    }
    if (distance_min > distance_max) {
        double t = distance_min;
        distance_min = distance_max;
        distance_max = t;
    }
    *tmin = distance_min;
    *tmax = distance_max;
}

/* f64::max / f64::min (IEEE maxNum/minNum: a NaN operand is ignored) = C fmax / fmin */

/* shapes/cube.rs:65-85 */
static void cube_local_intersect(const ray_t *ray, uint32_t shape, isect_list *out) {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    cube_check_axis(ray->origin.x, ray->direction.x, &xmin, &xmax);
    cube_check_axis(ray->origin.y, ray->direction.y, &ymin, &ymax);
    cube_check_axis(ray->origin.z, ray->direction.z, &zmin, &zmax);
    double distance_min = fmax(fmax(fmax(RTO_MIN, xmin), ymin), zmin);
    double distance_max = fmin(fmin(fmin(RTO_MAX, xmax), ymax), zmax);
    if (distance_min < distance_max && distance_max > 0.0) {
        list_push(out, distance_min, shape);
        list_push(out, distance_max, shape);
    }
}

/* shapes/cylinder.rs:34-39 */
static int cylinder_check_cap(const ray_t *ray, double distance) {
    double x = fma(ray->direction.x, distance, ray->origin.x);
    double z = fma(ray->direction.z, distance, ray->origin.z);
    return (squared(x) + squared(z)) <= 1.0;
}

/* shapes/cylinder.rs:41-59 */
static void cylinder_intersect_caps(const ray_t *ray, uint32_t shape, double min, double max, int closed,
                                    isect_list *out) {
    if (!closed || fabs(ray->direction.y) < RTO_EPSILON) return;
    double distance = (min - ray->origin.y) / ray->direction.y;
    if (cylinder_check_cap(ray, distance)) list_push(out, distance, shape);
    distance = (max - ray->origin.y) / ray->direction.y;
    if (cylinder_check_cap(ray, distance)) list_push(out, distance, shape);
}

/* shapes/cylinder.rs:81-110 */
static void cylinder_local_intersect(const ray_t *ray, uint32_t shape, double min, double max, int closed,
                                     isect_list *out) {
    double a = squared(ray->direction.x) + squared(ray->direction.z);
    if (fabs(a) > 0.0) {
        double b = 2.0 * fma(ray->origin.x, ray->direction.x, ray->origin.z * ray->direction.z);
        double c = squared(ray->origin.x) + squared(ray->origin.z) - 1.0;
        double distance_1, distance_2;
        if (solve_quadratic(a, b, c, &distance_1, &distance_2)) {
            if (distance_1 > distance_2) {
                double t = distance_1;
                distance_1 = distance_2;
                distance_2 = t;
            }
            double y1 = fma(distance_1, ray->direction.y, ray->origin.y);
            if (min < y1 && y1 < max) list_push(out, distance_1, shape);
            double y2 = fma(distance_2, ray->direction.y, ray->origin.y);
            if (min < y2 && y2 < max) list_push(out, distance_2, shape);
        }
    }
    cylinder_intersect_caps(ray, shape, min, max, closed, out);
}

/* shapes/cone.rs:34-39 */
static int cone_check_caps(const ray_t *ray, double distance, double radius) {
    double x = fma(ray->direction.x, distance, ray->origin.x);
    double z = fma(ray->direction.z, distance, ray->origin.z);
    return (squared(x) + squared(z)) <= squared(radius);
}

/* shapes/cone.rs:41-59 */
static void cone_intersect_caps(const ray_t *ray, uint32_t shape, double min, double max, int closed,
                                isect_list *out) {
    if (!closed || fabs(ray->direction.y) < RTO_EPSILON) return;
    double distance = (min - ray->origin.y) / ray->direction.y;
    if (cone_check_caps(ray, distance, min)) list_push(out, distance, shape);
    distance = (max - ray->origin.y) / ray->direction.y;
    if (cone_check_caps(ray, distance, max)) list_push(out, distance, shape);
}

/* shapes/cone.rs:81-112 */
static void cone_local_intersect(const ray_t *ray, uint32_t shape, double min, double max, int closed,
                                 isect_list *out) {
    v3 o = ray->origin, d = ray->direction;
    double a = squared(d.x) - squared(d.y) + squared(d.z);
    double b = 2.0 * fma(o.z, d.z, fma(o.x, d.x, -o.y * d.y));
    double c = squared(o.x) - squared(o.y) + squared(o.z);
    double distance_1, distance_2;
    if (fabs(a) < RTO_EPSILON && fabs(b) > RTO_EPSILON) {
        double distance = -c / (2.0 * b);
        list_push(out, distance, shape);
    } else if (solve_quadratic(a, b, c, &distance_1, &distance_2)) {
        if (distance_1 > distance_2) {
            double t = distance_1;
            distance_1 = distance_2;
            distance_2 = t;
        }
        double y1 = fma(d.y, distance_1, o.y);
        if (min < y1 && y1 < max) list_push(out, distance_1, shape);
        double y2 = fma(d.y, distance_2, o.y);
        if (min < y2 && y2 < max) list_push(out, distance_2, shape);
    }
    cone_intersect_caps(ray, shape, min, max, closed, out);
}

/* shapes/triangle.rs:39-56 */
static void triangle_local_intersect(const ray_t *ray, uint32_t shape, v3 vertex_1, v3 edge_1, v3 edge_2,
                                     isect_list *out) {
    v3 direction_cross_edge2 = v3_cross(ray->direction, edge_2);
    double determinant = v3_dot(edge_1, direction_cross_edge2);
    if (fabs(determinant) < RTO_EPSILON) return;
    v3 vertex1_to_origin = v3_sub(ray->origin, vertex_1);
    double u = v3_dot(vertex1_to_origin, direction_cross_edge2) / determinant;
    if (!(u >= 0.0 && u <= 1.0)) return; /* !(0.0..=1.0).contains(&u) */
    v3 origin_cross_edge1 = v3_cross(vertex1_to_origin, edge_1);
    double v = v3_dot(ray->direction, origin_cross_edge1) / determinant;
    if (v > 0.0 && u + v < 1.0) {
        double distance = v3_dot(edge_2, origin_cross_edge1) / determinant;
        list_push(out, distance, shape);
    }
}

/* ray.rs:35-49: transform the ray into object space, then the shape's local_intersect */
static void ray_intersect(const rtgpu_scene *s, const ray_t *ray, uint32_t shape, isect_list *out) {
    const double *inv = s->shape_inv + (size_t)shape * 12;
    ray_t local;
    local.origin = mat_point(inv, ray->origin);
    local.direction = mat_vector(inv, ray->direction);
    switch (s->shape_type[shape]) {
    case RTGPU_SPHERE: sphere_local_intersect(&local, shape, out); break;
    case RTGPU_PLANE: plane_local_intersect(&local, shape, out); break;
    case RTGPU_CUBE: cube_local_intersect(&local, shape, out); break;
    case RTGPU_CYLINDER:
        cylinder_local_intersect(&local, shape, s->shape_min[shape], s->shape_max[shape], s->shape_closed[shape], out);
        break;
    case RTGPU_CONE:
        cone_local_intersect(&local, shape, s->shape_min[shape], s->shape_max[shape], s->shape_closed[shape], out);
        break;
    case RTGPU_TRIANGLE: {
        size_t t = (size_t)s->shape_triangle[shape] * 3;
        triangle_local_intersect(&local, shape, v3_from(s->tri_vertex_1 + t), v3_from(s->tri_edge_1 + t),
                                 v3_from(s->tri_edge_2 + t), out);
        break;
    }
    default: break;
    }
}

/* ---------------------------------------------------------------------------------------- */
/* local_normal_at of the six shapes + Shape::normal_at                                      */

/* utils.rs:16-24 */
static inline int coarse_eq(double a, double b) {
    if (a == b) return 1;
    return fabs(a - b) < RTO_EPSILON;
}

static v3 local_normal_at(const rtgpu_scene *s, uint32_t shape, v3 p) {
    switch (s->shape_type[shape]) {
    case RTGPU_SPHERE: /* sphere.rs:57-59 */ return p;
    case RTGPU_PLANE: /* plane.rs:52-54 */ return v3_make(0.0, 1.0, 0.0);
    case RTGPU_CUBE: { /* cube.rs:89-101 */
        double ax = fabs(p.x), ay = fabs(p.y), az = fabs(p.z);
        double max_value = fmax(fmax(fmax(RTO_MIN, ax), ay), az);
        if (coarse_eq(max_value, ax)) return v3_make(p.x, 0.0, 0.0);
        else if (coarse_eq(max_value, ay)) return v3_make(0.0, p.y, 0.0);
        return v3_make(0.0, 0.0, p.z);
    }
    case RTGPU_CYLINDER: { /* cylinder.rs:114-126 */
        double distance = squared(p.x) + squared(p.z);
        if (distance < 1.0 && p.y >= (s->shape_max[shape] - RTO_EPSILON)) return v3_make(0.0, 1.0, 0.0);
        if (distance < 1.0 && p.y <= (s->shape_min[shape] + RTO_EPSILON)) return v3_make(0.0, -1.0, 0.0);
        return v3_make(p.x, 0.0, p.z);
    }
    case RTGPU_CONE: { /* cone.rs:116-133 */
        double distance = squared(p.x) + squared(p.z);
        if (distance < squared(s->shape_max[shape]) && p.y >= (s->shape_max[shape] - RTO_EPSILON))
            return v3_make(0.0, 1.0, 0.0);
        if (distance < squared(s->shape_min[shape]) && p.y <= (s->shape_min[shape] + RTO_EPSILON))
            return v3_make(0.0, -1.0, 0.0);
        double y = sqrt(distance);
        if (p.y > 0.0) y = -y;
        return v3_make(p.x, y, p.z);
    }
    case RTGPU_TRIANGLE: /* triangle.rs:78-80 */
        return v3_from(s->tri_normal + (size_t)s->shape_triangle[shape] * 3);
    default: return v3_make(0.0, 0.0, 0.0);
    }
}

/* shapes/shape.rs:22-27 */
static v3 shape_normal_at(const rtgpu_scene *s, uint32_t shape, v3 point) {
    const double *inv = s->shape_inv + (size_t)shape * 12;
    v3 local_point = mat_point(inv, point);
    v3 local_normal = local_normal_at(s, shape, local_point);
    v3 world_normal = mat_transposed_vector(inv, local_normal);
    return v3_normalized(world_normal);
}

/* ---------------------------------------------------------------------------------------- */
/* Patterns                                                                                  */

/* Rust `f64 as i64`: truncates toward zero, saturates, NaN -> 0 */
static inline int64_t f64_as_i64(double v) {
    if (v != v) return 0;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}

static v3 pattern_color_at(const rtgpu_scene *s, uint32_t pattern, v3 p) {
    v3 a = v3_from(s->pat_color_a + (size_t)pattern * 3);
    v3 b = v3_from(s->pat_color_b + (size_t)pattern * 3);
    switch (s->pat_type[pattern]) {
    case RTGPU_PATTERN_STRIPE: { /* stripe_pattern.rs:24-31 */
        int64_t distance = f64_as_i64(floor(p.x));
        return (distance % 2 == 0) ? a : b;
    }
    case RTGPU_PATTERN_GRADIENT: { /* gradient_pattern.rs:24-31 */
        v3 distance = v3_sub(b, a);
        double fraction = fabs(p.x - trunc(p.x)); /* f64::fract */
        if (f64_as_i64(p.x) % 2 != 0) fraction = 1.0 - fraction;
        return v3_add(a, v3_scale(distance, fraction));
    }
    case RTGPU_PATTERN_RING: { /* ring_pattern.rs:25-32 */
        int64_t distance = f64_as_i64(floor(sqrt(squared(p.x) + squared(p.z))));
        return (distance % 2 == 0) ? a : b;
    }
    case RTGPU_PATTERN_CHECKER: { /* checker_pattern.rs:24-31 */
        int64_t distance = f64_as_i64(floor(p.x) + floor(p.y) + floor(p.z));
        return (distance % 2 == 0) ? a : b;
    }
    case RTGPU_PATTERN_COMPLEX: { /* complex_pattern.rs:24-33: children see the SAME point */
        int64_t distance = f64_as_i64(floor(p.x));
        int32_t child = (distance % 2 == 0) ? s->pat_child_a[pattern] : s->pat_child_b[pattern];
        return pattern_color_at(s, (uint32_t)child, p);
    }
    case RTGPU_PATTERN_TEST: /* pattern.rs:29-60 (TestPattern): colour = the point */
        return p;
    default: return v3_make(0.0, 0.0, 0.0);
    }
}

/* patterns/pattern.rs:10-14 */
static v3 pattern_color_at_shape(const rtgpu_scene *s, uint32_t pattern, uint32_t shape, v3 point) {
    v3 object_point = mat_point(s->shape_inv + (size_t)shape * 12, point);
    v3 pattern_point = mat_point(s->pat_inv + (size_t)pattern * 12, object_point);
    return pattern_color_at(s, pattern, pattern_point);
}

/* ---------------------------------------------------------------------------------------- */
/* Material::lighting                                                                        */

/* composites/material.rs:53-114 */
static v3 material_lighting(const rtgpu_scene *s, uint32_t material, uint32_t shape, v3 light_position,
                            v3 light_intensity, v3 point, v3 camera_direction, v3 normal, int in_shadow) {
    const double *mp = s->mat_params + (size_t)material * RTGPU_MAT_PARAM_COUNT;
    /* resolve_color, material.rs:75-80 */
    v3 base = (s->mat_pattern[material] >= 0)
                  ? pattern_color_at_shape(s, (uint32_t)s->mat_pattern[material], shape, point)
                  : v3_from(s->mat_color + (size_t)material * 3);
    v3 effective_color = v3_mul(base, light_intensity);
    /* calculate_lighting, material.rs:83-114 */
    v3 ambient = v3_scale(effective_color, mp[RTGPU_MAT_AMBIENT]);
    if (in_shadow) return ambient;
    v3 light_direction = v3_normalized(v3_sub(light_position, point));
    double light_dot_normal = v3_dot(light_direction, normal);
    if (light_dot_normal < 0.0) return ambient;
    v3 diffuse = v3_scale(v3_scale(effective_color, mp[RTGPU_MAT_DIFFUSE]), light_dot_normal);
    v3 reflect_direction = v3_reflect(v3_neg(light_direction), normal);
    double reflect_dot_camera = v3_dot(reflect_direction, camera_direction);
    if (reflect_dot_camera <= 0.0) return v3_add(ambient, diffuse);
    double factor = pow(reflect_dot_camera, mp[RTGPU_MAT_SHININESS]);
    v3 specular = v3_scale(v3_scale(light_intensity, mp[RTGPU_MAT_SPECULAR]), factor);
    return v3_add(v3_add(ambient, diffuse), specular);
}

/* ---------------------------------------------------------------------------------------- */
/* World                                                                                     */

typedef struct world_ctx {
    const rtgpu_scene *s;
    isect_list *lists; /* scratch Intersections buffers, 2 per recursion level */
    size_t n_lists;
    uint32_t *containers; /* scratch Vec<&dyn Shape> of prepare_computations */
    size_t containers_cap;
    rtgpu_stats stats;
} world_ctx;

static void ctx_init(world_ctx *c, const rtgpu_scene *s, uint32_t max_depth) {
    memset(c, 0, sizeof(*c));
    c->s = s;
    c->n_lists = 2 * ((size_t)max_depth + 2);
    c->lists = (isect_list *)calloc(c->n_lists, sizeof(isect_list));
}

static void ctx_free(world_ctx *c) {
    for (size_t i = 0; i < c->n_lists; i++) list_free(&c->lists[i]);
    free(c->lists);
    free(c->containers);
}

/* world.rs:25-35 */
static void world_collect_intersections(world_ctx *c, const ray_t *ray, isect_list *out) {
    out->n = 0;
    for (uint32_t i = 0; i < c->s->n_shapes; i++) ray_intersect(c->s, ray, i, out);
    list_sort(out);
}

/* intersections.rs:13-18: filter(distance >= 0.0).min() — Iterator::min keeps the FIRST minimum */
static const isect_t *list_hit(const isect_list *l) {
    const isect_t *best = NULL;
    for (size_t i = 0; i < l->n; i++) {
        const isect_t *e = &l->v[i];
        if (!(e->distance >= 0.0)) continue;
        if (best == NULL || e->distance < best->distance) best = e;
    }
    return best;
}

typedef struct computed_hit {
    double distance;
    uint32_t shape;
    int is_inside;
    v3 point, over_point, under_point, camera_direction, normal, reflect_direction;
    double refractive_index_1, refractive_index_2;
} computed_hit;

static inline double shape_refractive_index(const rtgpu_scene *s, uint32_t shape) {
    return s->mat_params[(size_t)s->shape_material[shape] * RTGPU_MAT_PARAM_COUNT + RTGPU_MAT_REFRACTIVE_INDEX];
}

/* `*shape == intersection.shape` on `dyn Shape` is equality BY VALUE (shapes/shape.rs:34-38,
 * dyn_partial_eq.rs:9-16); the flattener encodes it as shape_eq_class. */
static inline int shapes_equal(const rtgpu_scene *s, uint32_t a, uint32_t b) {
    return s->shape_eq_class[a] == s->shape_eq_class[b];
}

/* composites/intersection.rs:21-75 + computed_hit.rs:22-48 */
static void prepare_computations(world_ctx *c, const isect_t *self, const ray_t *ray, const isect_list *xs,
                                 computed_hit *h) {
    const rtgpu_scene *s = c->s;
    v3 point = ray_position(ray, self->distance);
    v3 normal = shape_normal_at(s, self->shape, point);
    v3 camera_direction = v3_neg(ray->direction);
    int is_inside = v3_dot(normal, camera_direction) < 0.0;
    if (is_inside) normal = v3_neg(normal);
    v3 reflect_direction = v3_reflect(ray->direction, normal);

    if (c->containers_cap < xs->n + 1) {
        c->containers_cap = xs->n + 64;
        c->containers = (uint32_t *)realloc(c->containers, c->containers_cap * sizeof(uint32_t));
    }
    uint32_t *shapes = c->containers;
    size_t n_shapes = 0;
    double refractive_index_1 = RTO_DEFAULT_RI;
    double refractive_index_2 = RTO_DEFAULT_RI;
    for (size_t i = 0; i < xs->n; i++) {
        const isect_t *x = &xs->v[i];
        /* derived PartialEq of Intersection: distance == distance && shape == shape (by value) */
        int is_self = (self->distance == x->distance) && shapes_equal(s, self->shape, x->shape);
        if (is_self)
            refractive_index_1 = n_shapes ? shape_refractive_index(s, shapes[n_shapes - 1]) : RTO_DEFAULT_RI;
        size_t position = n_shapes;
        for (size_t k = 0; k < n_shapes; k++)
            if (shapes_equal(s, shapes[k], x->shape)) {
                position = k;
                break;
            }
        if (position < n_shapes) {
            memmove(shapes + position, shapes + position + 1, (n_shapes - position - 1) * sizeof(uint32_t));
            n_shapes--;
        } else {
            shapes[n_shapes++] = x->shape;
        }
        if (is_self) {
            refractive_index_2 = n_shapes ? shape_refractive_index(s, shapes[n_shapes - 1]) : RTO_DEFAULT_RI;
            break;
        }
    }

    h->distance = self->distance;
    h->shape = self->shape;
    h->is_inside = is_inside;
    h->point = point;
    h->camera_direction = camera_direction;
    h->normal = normal;
    h->reflect_direction = reflect_direction;
    h->refractive_index_1 = refractive_index_1;
    h->refractive_index_2 = refractive_index_2;
    /* computed_hit.rs:33-34 */
    h->over_point = v3_add(point, v3_scale(normal, RTO_EPSILON));
    h->under_point = v3_sub(point, v3_scale(normal, RTO_EPSILON));
}

/* computed_hit.rs:50-68 */
static double schlicks_approximation(const computed_hit *h) {
    double cos = v3_dot(h->camera_direction, h->normal);
    if (h->refractive_index_1 > h->refractive_index_2) {
        double refraction_ratio = h->refractive_index_1 / h->refractive_index_2;
        double sin2_t = squared(refraction_ratio) * (1.0 - squared(cos));
        if (sin2_t > 1.0) return 1.0;
        cos = sqrt(1.0 - sin2_t);
    }
    double reflection_coefficient =
        squared((h->refractive_index_1 - h->refractive_index_2) / (h->refractive_index_1 + h->refractive_index_2));
    double x = 1.0 - cos;
    double x5 = x * ((x * x) * (x * x)); /* powi(5): compiler-rt __powidf2 square-and-multiply order */
    return fma(1.0 - reflection_coefficient, x5, reflection_coefficient);
}

static v3 world_internal_color_at(world_ctx *c, const ray_t *ray, size_t level, uint32_t remaining);

/* world.rs:98-112 */
static int world_is_in_shadow(world_ctx *c, v3 light_position, v3 point, isect_list *xs) {
    v3 light_direction = v3_sub(light_position, point);
    double light_distance = v3_magnitude(light_direction);
    ray_t shadow_ray;
    shadow_ray.origin = point;
    shadow_ray.direction = v3_normalized(light_direction);
    c->stats.rays_shadow++;
    world_collect_intersections(c, &shadow_ray, xs);
    for (size_t i = 0; i < xs->n; i++) {
        const isect_t *x = &xs->v[i];
        /* intersection.rs:77-79 */
        if (c->s->mat_casts_shadow[c->s->shape_material[x->shape]] && x->distance >= 0.0 &&
            x->distance < light_distance)
            return 1;
    }
    return 0;
}

/* world.rs:114-128 */
static v3 world_reflected_color(world_ctx *c, const computed_hit *h, size_t level, uint32_t remaining) {
    const double *mp = c->s->mat_params + (size_t)c->s->shape_material[h->shape] * RTGPU_MAT_PARAM_COUNT;
    if (remaining == 0 || mp[RTGPU_MAT_REFLECTIVENESS] == 0.0) return v3_make(0.0, 0.0, 0.0);
    ray_t reflected_ray;
    reflected_ray.origin = h->over_point;
    reflected_ray.direction = h->reflect_direction;
    c->stats.rays_reflect++;
    v3 reflected_color = world_internal_color_at(c, &reflected_ray, level + 1, remaining - 1);
    return v3_scale(reflected_color, mp[RTGPU_MAT_REFLECTIVENESS]);
}

/* world.rs:130-157 */
static v3 world_refracted_color(world_ctx *c, const computed_hit *h, size_t level, uint32_t remaining) {
    const double *mp = c->s->mat_params + (size_t)c->s->shape_material[h->shape] * RTGPU_MAT_PARAM_COUNT;
    if (remaining == 0 || mp[RTGPU_MAT_TRANSPARENCY] == 0.0) return v3_make(0.0, 0.0, 0.0);
    double n_ratio = h->refractive_index_1 / h->refractive_index_2;
    double cos_i = v3_dot(h->camera_direction, h->normal);
    double sin2_t = squared(n_ratio) * (1.0 - squared(cos_i));
    if (sin2_t > 1.0) return v3_make(0.0, 0.0, 0.0);
    double cos_t = sqrt(1.0 - sin2_t);
    v3 direction = v3_sub(v3_scale(h->normal, fma(n_ratio, cos_i, -cos_t)), v3_scale(h->camera_direction, n_ratio));
    ray_t refracted_ray;
    refracted_ray.origin = h->under_point;
    refracted_ray.direction = direction;
    c->stats.rays_refract++;
    v3 refracted_color = world_internal_color_at(c, &refracted_ray, level + 1, remaining - 1);
    return v3_scale(refracted_color, mp[RTGPU_MAT_TRANSPARENCY]);
}

/* world.rs:38-67 */
static v3 world_shade_hit(world_ctx *c, const computed_hit *h, size_t level, uint32_t remaining) {
    const rtgpu_scene *s = c->s;
    uint32_t material = s->shape_material[h->shape];
    const double *mp = s->mat_params + (size_t)material * RTGPU_MAT_PARAM_COUNT;
    isect_list *shading = &c->lists[2 * level + 1];
    v3 surface_color = v3_make(0.0, 0.0, 0.0); /* fold(Color::BLACK, Color::add) */
    for (uint32_t l = 0; l < s->n_lights; l++) {
        v3 light_position = v3_from(s->light_position + (size_t)l * 3);
        v3 light_intensity = v3_from(s->light_intensity + (size_t)l * 3);
        int in_shadow = world_is_in_shadow(c, light_position, h->over_point, shading);
        /* lighting_from_computed_hit, material.rs:116-130: evaluated at over_point */
        v3 lit = material_lighting(s, material, h->shape, light_position, light_intensity, h->over_point,
                                   h->camera_direction, h->normal, in_shadow);
        surface_color = v3_add(surface_color, lit);
    }
    v3 reflected_color = world_reflected_color(c, h, level, remaining);
    v3 refracted_color = world_refracted_color(c, h, level, remaining);
    if (mp[RTGPU_MAT_REFLECTIVENESS] > 0.0 && mp[RTGPU_MAT_TRANSPARENCY] > 0.0) {
        double reflectance = schlicks_approximation(h);
        return v3_add(v3_add(surface_color, v3_scale(reflected_color, reflectance)),
                      v3_scale(refracted_color, 1.0 - reflectance));
    }
    return v3_add(v3_add(surface_color, reflected_color), refracted_color);
}

/* world.rs:70-86 */
static v3 world_internal_color_at(world_ctx *c, const ray_t *ray, size_t level, uint32_t remaining) {
    isect_list *xs = &c->lists[2 * level];
    world_collect_intersections(c, ray, xs);
    const isect_t *hit = list_hit(xs);
    if (hit == NULL) return v3_make(0.0, 0.0, 0.0); /* World::DEFAULT_COLOR */
    computed_hit h;
    prepare_computations(c, hit, ray, xs, &h);
    c->stats.hit_nodes++;
    return world_shade_hit(c, &h, level, remaining);
}

/* camera.rs:52-68 */
static ray_t camera_ray_for_pixel(const rtgpu_camera *cam, uint32_t px, uint32_t py) {
    double offset_x = ((double)px + 0.5) * cam->pixel_size;
    double offset_y = ((double)py + 0.5) * cam->pixel_size;
    double world_x = cam->half_width - offset_x;
    double world_y = cam->half_height - offset_y;
    v3 pixel = mat_point(cam->inv, v3_make(world_x, world_y, -1.0));
    v3 origin = v3_from(cam->origin);
    ray_t r;
    r.origin = origin;
    r.direction = v3_normalized(v3_sub(pixel, origin));
    return r;
}

/* canvas.rs:117-123 + color.rs:84-90: clamp(0,1) * 255 -> round (half away from zero) -> as u8 */
uint8_t rto_quantise(double channel) {
    double v = channel;
    if (v < 0.0) v = 0.0;
    if (v > 1.0) v = 1.0; /* f64::clamp leaves NaN as NaN */
    v = round(v * 255.0);
    if (v != v) return 0; /* NaN as u8 = 0 */
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

/* ---------------------------------------------------------------------------------------- */
/* Camera::render / render_parallel                                                           */

static void stats_add(rtgpu_stats *a, const rtgpu_stats *b) {
    a->rays_primary += b->rays_primary;
    a->rays_shadow += b->rays_shadow;
    a->rays_reflect += b->rays_reflect;
    a->rays_refract += b->rays_refract;
    a->hit_nodes += b->hit_nodes;
    a->pixels += b->pixels;
}

static double now_ms(void) {
#ifdef _OPENMP
    return omp_get_wtime() * 1e3;
#else
    return 0.0;
#endif
}

int rto_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* pixel k of the job -> (canvas index, output slot) */
typedef struct pixel_job {
    const rtgpu_scene *scene;
    const rtgpu_camera *camera;
    uint32_t max_depth;
    const uint64_t *pixels; /* explicit list, or NULL */
    const uint32_t *rows;   /* selected rows (full-frame output), or NULL */
    size_t n;
    double *out_rgb;
    uint8_t *out_rgb8;
} pixel_job;

static void run_job(const pixel_job *job, int threads, rtgpu_stats *stats) {
    rtgpu_stats total;
    memset(&total, 0, sizeof(total));
    double t0 = now_ms();
    const uint32_t hsize = job->camera->hsize;
    if (threads <= 0) threads = rto_max_threads();
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        world_ctx c;
        ctx_init(&c, job->scene, job->max_depth);
        /* rayon hands out pixel ranges adaptively (camera.rs:101-110); dynamic chunks play that role */
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 256)
#endif
        for (long long k = 0; k < (long long)job->n; k++) {
            uint64_t index, slot;
            if (job->pixels) {
                index = job->pixels[k];
                slot = (uint64_t)k;
            } else if (job->rows) {
                index = (uint64_t)job->rows[(uint64_t)k / hsize] * hsize + (uint64_t)k % hsize;
                slot = index;
            } else {
                index = (uint64_t)k;
                slot = index;
            }
            /* canvas.rs:53-55 */
            uint32_t x = (uint32_t)(index % hsize);
            uint32_t y = (uint32_t)(index / hsize);
            ray_t ray = camera_ray_for_pixel(job->camera, x, y);
            c.stats.rays_primary++;
            c.stats.pixels++;
            v3 colour = world_internal_color_at(&c, &ray, 0, job->max_depth); /* world.rs:89-95 */
            if (job->out_rgb) v3_store(job->out_rgb + slot * 3, colour);
            if (job->out_rgb8) {
                job->out_rgb8[slot * 3 + 0] = rto_quantise(colour.x);
                job->out_rgb8[slot * 3 + 1] = rto_quantise(colour.y);
                job->out_rgb8[slot * 3 + 2] = rto_quantise(colour.z);
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        stats_add(&total, &c.stats);
        ctx_free(&c);
    }
    if (stats) {
        *stats = total;
        stats->total_ms = now_ms() - t0;
        stats->kernel_ms = stats->total_ms;
    }
}

static uint32_t rows_list(const rtgpu_rows *rows, uint32_t vsize, uint32_t *out) {
    uint32_t band = (rows && rows->band_rows) ? rows->band_rows : vsize ? vsize : 1;
    uint32_t count = (rows && rows->shard_count) ? rows->shard_count : 1;
    uint32_t index = rows ? rows->shard_index : 0;
    uint32_t n = 0;
    for (uint32_t y = 0; y < vsize; y++)
        if ((y / band) % count == index) {
            if (out) out[n] = y;
            n++;
        }
    return n;
}

int rto_render(const rtgpu_scene *scene, const rtgpu_camera *camera, uint32_t max_depth, const rtgpu_rows *rows,
               int threads, double *out_rgb, uint8_t *out_rgb8, rtgpu_stats *stats) {
    if (!scene || !camera || (!out_rgb && !out_rgb8)) return RTGPU_ERR_INVALID_ARGUMENT;
    pixel_job job;
    memset(&job, 0, sizeof(job));
    job.scene = scene;
    job.camera = camera;
    job.max_depth = max_depth;
    job.out_rgb = out_rgb;
    job.out_rgb8 = out_rgb8;
    uint32_t *row_list = NULL;
    if (rows && rows->shard_count > 1) {
        row_list = (uint32_t *)malloc(sizeof(uint32_t) * (camera->vsize ? camera->vsize : 1));
        uint32_t n_rows = rows_list(rows, camera->vsize, row_list);
        job.rows = row_list;
        job.n = (size_t)n_rows * camera->hsize;
    } else {
        job.n = (size_t)camera->hsize * camera->vsize;
    }
    run_job(&job, threads, stats);
    free(row_list);
    return RTGPU_OK;
}

int rto_render_pixels(const rtgpu_scene *scene, const rtgpu_camera *camera, uint32_t max_depth,
                      const uint64_t *pixels, size_t n, int threads, double *out_rgb, uint8_t *out_rgb8,
                      rtgpu_stats *stats) {
    if (!scene || !camera || !pixels || (!out_rgb && !out_rgb8)) return RTGPU_ERR_INVALID_ARGUMENT;
    pixel_job job;
    memset(&job, 0, sizeof(job));
    job.scene = scene;
    job.camera = camera;
    job.max_depth = max_depth;
    job.pixels = pixels;
    job.n = n;
    job.out_rgb = out_rgb;
    job.out_rgb8 = out_rgb8;
    run_job(&job, threads, stats);
    return RTGPU_OK;
}

/* ---------------------------------------------------------------------------------------- */
/* micro entry points                                                                        */

static ray_t ray_from(const double r[6]) {
    ray_t ray;
    ray.origin = v3_from(r);
    ray.direction = v3_from(r + 3);
    return ray;
}

void rto_ray_for_pixel(const rtgpu_camera *camera, uint32_t px, uint32_t py, double out[6]) {
    ray_t r = camera_ray_for_pixel(camera, px, py);
    v3_store(out, r.origin);
    v3_store(out + 3, r.direction);
}

int rto_intersect_shape(const rtgpu_scene *scene, uint32_t shape, const double ray[6], double out_t[4]) {
    isect_list l;
    memset(&l, 0, sizeof(l));
    ray_t r = ray_from(ray);
    ray_intersect(scene, &r, shape, &l);
    int n = (int)l.n;
    for (int i = 0; i < n && i < 4; i++) out_t[i] = l.v[i].distance;
    list_free(&l);
    return n;
}

size_t rto_collect_intersections(const rtgpu_scene *scene, const double ray[6], double *out_t, uint32_t *out_shape,
                                 size_t capacity) {
    world_ctx c;
    ctx_init(&c, scene, 0);
    ray_t r = ray_from(ray);
    world_collect_intersections(&c, &r, &c.lists[0]);
    size_t n = c.lists[0].n;
    for (size_t i = 0; i < n && i < capacity; i++) {
        if (out_t) out_t[i] = c.lists[0].v[i].distance;
        if (out_shape) out_shape[i] = c.lists[0].v[i].shape;
    }
    ctx_free(&c);
    return n;
}

void rto_normal_at(const rtgpu_scene *scene, uint32_t shape, const double point[3], double out[3]) {
    v3_store(out, shape_normal_at(scene, shape, v3_from(point)));
}

void rto_local_normal_at(const rtgpu_scene *scene, uint32_t shape, const double point[3], double out[3]) {
    v3_store(out, local_normal_at(scene, shape, v3_from(point)));
}

void rto_pattern_at_shape(const rtgpu_scene *scene, uint32_t pattern, uint32_t shape, const double point[3],
                          double out[3]) {
    v3_store(out, pattern_color_at_shape(scene, pattern, shape, v3_from(point)));
}

void rto_lighting(const rtgpu_scene *scene, uint32_t material, uint32_t shape, const double light_position[3],
                  const double light_intensity[3], const double point[3], const double eye[3],
                  const double normal[3], int in_shadow, double out[3]) {
    v3_store(out, material_lighting(scene, material, shape, v3_from(light_position), v3_from(light_intensity),
                                    v3_from(point), v3_from(eye), v3_from(normal), in_shadow));
}

int rto_is_in_shadow(const rtgpu_scene *scene, uint32_t light, const double point[3]) {
    world_ctx c;
    ctx_init(&c, scene, 0);
    int r = world_is_in_shadow(&c, v3_from(scene->light_position + (size_t)light * 3), v3_from(point), &c.lists[0]);
    ctx_free(&c);
    return r;
}

static const isect_t *pick_entry(const isect_list *xs, int k) {
    if (k < 0) return list_hit(xs);
    if ((size_t)k >= xs->n) return NULL;
    return &xs->v[k];
}

/* The sorted world list of `ray`, or the caller's hand-built list (the reference's unit tests build
 * Intersections by hand, e.g. world.rs:547-571). */
static void fill_list(world_ctx *c, const ray_t *r, const double *list_t, const uint32_t *list_shape, size_t list_n) {
    if (list_t && list_shape) {
        c->lists[0].n = 0;
        for (size_t i = 0; i < list_n; i++) list_push(&c->lists[0], list_t[i], list_shape[i]);
    } else {
        world_collect_intersections(c, r, &c->lists[0]);
    }
}

int rto_prepare_computations(const rtgpu_scene *scene, const double ray[6], const double *list_t,
                             const uint32_t *list_shape, size_t list_n, int k, rto_computed_hit *out) {
    world_ctx c;
    ctx_init(&c, scene, 0);
    ray_t r = ray_from(ray);
    fill_list(&c, &r, list_t, list_shape, list_n);
    const isect_t *e = pick_entry(&c.lists[0], k);
    int ok = 0;
    if (e) {
        computed_hit h;
        prepare_computations(&c, e, &r, &c.lists[0], &h);
        out->distance = h.distance;
        out->shape = h.shape;
        out->is_inside = (uint32_t)h.is_inside;
        v3_store(out->point, h.point);
        v3_store(out->over_point, h.over_point);
        v3_store(out->under_point, h.under_point);
        v3_store(out->camera_direction, h.camera_direction);
        v3_store(out->normal, h.normal);
        v3_store(out->reflect_direction, h.reflect_direction);
        out->refractive_index_1 = h.refractive_index_1;
        out->refractive_index_2 = h.refractive_index_2;
        out->schlick = schlicks_approximation(&h);
        ok = 1;
    }
    ctx_free(&c);
    return ok;
}

void rto_color_at(const rtgpu_scene *scene, const double ray[6], uint32_t remaining, double out[3]) {
    world_ctx c;
    ctx_init(&c, scene, remaining);
    ray_t r = ray_from(ray);
    v3_store(out, world_internal_color_at(&c, &r, 0, remaining));
    ctx_free(&c);
}

int rto_shade_entry(const rtgpu_scene *scene, const double ray[6], const double *list_t,
                    const uint32_t *list_shape, size_t list_n, int k, uint32_t remaining, int which,
                    double out[3]) {
    world_ctx c;
    ctx_init(&c, scene, remaining + 1);
    ray_t r = ray_from(ray);
    fill_list(&c, &r, list_t, list_shape, list_n);
    const isect_t *e = pick_entry(&c.lists[0], k);
    int ok = 0;
    if (e) {
        computed_hit h;
        prepare_computations(&c, e, &r, &c.lists[0], &h);
        v3 colour;
        /* level 1: the children must not reuse lists[0], which `h` was prepared from */
        if (which == 0) colour = world_shade_hit(&c, &h, 1, remaining);
        else if (which == 1) colour = world_reflected_color(&c, &h, 1, remaining);
        else colour = world_refracted_color(&c, &h, 1, remaining);
        v3_store(out, colour);
        ok = 1;
    }
    ctx_free(&c);
    return ok;
}
