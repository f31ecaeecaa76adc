// rt_host.hpp — C++ host-side mirror of the reference's scene API, above the C ABI (include/rtgpu.h).
//
// The reference is compiled (Rust) code and its toolchain is absent from this image, so the host
// side that a `RenderingMode::Gpu` arm needs is provided in C++ with the reference's names and
// semantics (paths under the reference's `ray-tracer/src/`):
//
//   Matrix4 / transformations      primitives/matrix.rs:8-258,317-362, primitives/transformations.rs:5-87
//   Material / Light               composites/material.rs:9-20,157-161, primitives/light.rs:6-9,44-48
//   Pattern (stripe, gradient, ring, checker, complex, test)   patterns/*.rs
//   Shape (sphere, plane, cube, cylinder, cone, triangle)      shapes/*.rs
//   World / Camera / Canvas        composites/world.rs:9-22, camera.rs:10-49,114-127, canvas.rs:13-137
//
// Host math is restated operation for operation (cofactor inverse, fold-from-0.0 products, fma only
// where the reference says mul_add) so that the flattened scene is bit-identical to what the
// reference holds; build with -ffp-contract=off.  `Camera::render_gpu` is the only render method:
// the CPU render loops of the reference (camera.rs:79-112) are what the CUDA library replaces, and
// gpu mode has no CPU fallback.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtgpu.h"

namespace rt_host {

constexpr double EPSILON = 0.00000008;              // consts.rs:2
constexpr double F64_MAX = 1.7976931348623157e308;  // consts.rs:6
constexpr double F64_MIN = -F64_MAX;                // consts.rs:4

using Vec3 = std::array<double, 3>;  // Point / Vector / Color: 3 x f64 (point.rs:8, vector.rs:7, color.rs:7)

inline bool coarse_eq(double a, double b) { return a == b || std::fabs(a - b) < EPSILON; }  // utils.rs:16-24

// vector.rs:84-103
inline double dot(const Vec3& a, const Vec3& b) { return std::fma(a[2], b[2], std::fma(a[0], b[0], a[1] * b[1])); }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {std::fma(a[1], b[2], -a[2] * b[1]), std::fma(a[2], b[0], -a[0] * b[2]), std::fma(a[0], b[1], -a[1] * b[0])};
}
inline double magnitude(const Vec3& a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
inline Vec3 normalized(const Vec3& a) {
    const double m = magnitude(a);
    return {a[0] / m, a[1] / m, a[2] / m};
}
inline Vec3 sub(const Vec3& a, const Vec3& b) { return {a[0] - b[0], a[1] - b[1], a[2] - b[2]}; }

// ---------------------------------------------------------------------------------------------
// Matrix<4> (matrix.rs:8), row-major
struct Matrix4 {
    double m[4][4];

    static Matrix4 identity() {
        Matrix4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = i == j ? 1.0 : 0.0;
        return r;
    }
    bool operator==(const Matrix4& o) const {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                if (!(m[i][j] == o.m[i][j])) return false;
        return true;
    }
    // matrix.rs:45-51
    bool is_identity() const {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                if (!coarse_eq(m[i][j], i == j ? 1.0 : 0.0)) return false;
        return true;
    }
    // matrix.rs:317-330: fold from 0.0 over k
    Matrix4 operator*(const Matrix4& rhs) const {
        Matrix4 r{};
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double acc = 0.0;
                for (int k = 0; k < 4; ++k) acc = acc + (m[i][k] * rhs.m[k][j]);
                r.m[i][j] = acc;
            }
        return r;
    }
    // matrix.rs:332-346: rows 0..2 over [x, y, z, 1.0]
    Vec3 mul_point(const Vec3& p) const {
        const double v[4] = {p[0], p[1], p[2], 1.0};
        Vec3 out{};
        for (int r = 0; r < 3; ++r) {
            double acc = 0.0;
            for (int c = 0; c < 4; ++c) acc = acc + (m[r][c] * v[c]);
            out[r] = acc;
        }
        return out;
    }

   private:
    static double det2(const double a[2][2]) { return (a[0][0] * a[1][1]) - (a[0][1] * a[1][0]); }  // matrix.rs:83-85
    static double det3(const double a[3][3]) {                                                        // matrix.rs:151-155
        double acc = 0.0;
        for (int i = 0; i < 3; ++i) {
            double s[2][2];
            int rr = 0;
            for (int r = 1; r < 3; ++r, ++rr) {
                int cc = 0;
                for (int c = 0; c < 3; ++c)
                    if (c != i) s[rr][cc++] = a[r][c];
            }
            const double minor = det2(s);
            const double cof = (i % 2 == 0) ? minor : -minor;  // matrix.rs:161-168
            acc = acc + (a[0][i] * cof);
        }
        return acc;
    }

   public:
    // matrix.rs:190-240: minor via the 3x3 submatrix (the identity short-cut of :191-193 cannot trigger
    // here because inverse() has already returned for identity-like matrices)
    double cofactor(int row, int col) const {
        double s[3][3];
        int rr = 0;
        for (int r = 0; r < 4; ++r) {
            if (r == row) continue;
            int cc = 0;
            for (int c = 0; c < 4; ++c)
                if (c != col) s[rr][cc++] = m[r][c];
            ++rr;
        }
        const double minor = is_identity() ? 1.0 : det3(s);
        return ((row + col) % 2 == 0) ? minor : -minor;
    }
    double determinant() const {  // matrix.rs:223-227
        double acc = 0.0;
        for (int i = 0; i < 4; ++i) acc = acc + (m[0][i] * cofactor(0, i));
        return acc;
    }
    // matrix.rs:246-258: exact IDENTITY when identity within EPSILON, else cofactor(col,row)/det
    Matrix4 inverse() const {
        if (is_identity()) return identity();
        Matrix4 r{};
        const double det = determinant();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r.m[i][j] = cofactor(j, i) / det;
        return r;
    }
};
using Transformation = Matrix4;

namespace transformations {  // transformations.rs:5-87
inline Transformation translation(double x, double y, double z) {
    Transformation r = Transformation::identity();
    r.m[0][3] = x;
    r.m[1][3] = y;
    r.m[2][3] = z;
    return r;
}
inline Transformation scaling(double x, double y, double z) {
    Transformation r = Transformation::identity();
    r.m[0][0] = x;
    r.m[1][1] = y;
    r.m[2][2] = z;
    return r;
}
inline Transformation rotation_x(double theta) {
    Transformation r = Transformation::identity();
    const double c = std::cos(theta), s = std::sin(theta);
    r.m[1][1] = c;
    r.m[1][2] = -s;
    r.m[2][1] = s;
    r.m[2][2] = c;
    return r;
}
inline Transformation rotation_y(double theta) {
    Transformation r = Transformation::identity();
    const double c = std::cos(theta), s = std::sin(theta);
    r.m[0][0] = c;
    r.m[0][2] = s;
    r.m[2][0] = -s;
    r.m[2][2] = c;
    return r;
}
inline Transformation rotation_z(double theta) {
    Transformation r = Transformation::identity();
    const double c = std::cos(theta), s = std::sin(theta);
    r.m[0][0] = c;
    r.m[0][1] = -s;
    r.m[1][0] = s;
    r.m[1][1] = c;
    return r;
}
inline Transformation shearing(double xy, double xz, double yx, double yz, double zx, double zy) {
    Transformation r = Transformation::identity();
    r.m[0][1] = xy;
    r.m[0][2] = xz;
    r.m[1][0] = yx;
    r.m[1][2] = yz;
    r.m[2][0] = zx;
    r.m[2][1] = zy;
    return r;
}
inline Transformation view_transform(const Vec3& from, const Vec3& to, const Vec3& up) {
    const Vec3 forward = normalized(sub(to, from));
    const Vec3 up_normalized = normalized(up);
    const Vec3 left = cross(forward, up_normalized);
    const Vec3 true_up = cross(left, forward);
    Transformation o = Transformation::identity();
    for (int k = 0; k < 3; ++k) {
        o.m[0][k] = left[k];
        o.m[1][k] = true_up[k];
        o.m[2][k] = -forward[k];
    }
    return o * translation(-from[0], -from[1], -from[2]);
}
}  // namespace transformations

// ---------------------------------------------------------------------------------------------
// patterns/*.rs — every pattern stores only its transformation_inverse
struct Pattern {
    rtgpu_pattern_type type = RTGPU_PATTERN_STRIPE;
    Vec3 color_a{0, 0, 0}, color_b{0, 0, 0};
    Transformation transformation_inverse = Transformation::identity();
    std::shared_ptr<Pattern> pattern_a, pattern_b;  // ComplexPattern (complex_pattern.rs:8-12)

    static std::shared_ptr<Pattern> two_color(rtgpu_pattern_type t, const Vec3& a, const Vec3& b) {
        auto p = std::make_shared<Pattern>();
        p->type = t;
        p->color_a = a;
        p->color_b = b;
        return p;
    }
    void set_transformation(const Transformation& t) { transformation_inverse = t.inverse(); }
    bool equals(const Pattern& o) const {  // derived PartialEq + dyn_partial_eq.rs:9-16
        if (type != o.type || color_a != o.color_a || color_b != o.color_b || !(transformation_inverse == o.transformation_inverse)) return false;
        if ((pattern_a == nullptr) != (o.pattern_a == nullptr) || (pattern_b == nullptr) != (o.pattern_b == nullptr)) return false;
        if (pattern_a && !pattern_a->equals(*o.pattern_a)) return false;
        if (pattern_b && !pattern_b->equals(*o.pattern_b)) return false;
        return true;
    }
};

// composites/material.rs:9-20; Default = material.rs:157-161
struct Material {
    Vec3 color{1, 1, 1};
    std::shared_ptr<Pattern> pattern;
    double ambient = 0.1, diffuse = 0.9, specular = 0.9, shininess = 200.0;
    double reflectiveness = 0.0, refractive_index = 1.0, transparency = 0.0;
    bool casts_shadow = true;

    bool equals(const Material& o) const {
        if (color != o.color || ambient != o.ambient || diffuse != o.diffuse || specular != o.specular || shininess != o.shininess ||
            reflectiveness != o.reflectiveness || refractive_index != o.refractive_index || transparency != o.transparency ||
            casts_shadow != o.casts_shadow)
            return false;
        if ((pattern == nullptr) != (o.pattern == nullptr)) return false;
        return !pattern || pattern->equals(*o.pattern);
    }
};

struct Light {  // primitives/light.rs:6-9
    Vec3 position{-10, 10, -10};
    Vec3 intensity{1, 1, 1};
};

// shapes/*.rs — one struct, tagged (the flattener removes the trait objects anyway)
struct Shape {
    rtgpu_shape_type type = RTGPU_SPHERE;
    Material material;
    Transformation transformation_inverse = Transformation::identity();
    double min = F64_MIN, max = F64_MAX;  // cylinder.rs:133-143, cone.rs:140-150 (Default)
    bool closed = false;
    Vec3 vertex_1{}, vertex_2{}, vertex_3{}, edge_1{}, edge_2{}, normal{};  // triangle.rs:9-18

    static Shape make(rtgpu_shape_type t, const Material& m, const Transformation& transformation) {
        Shape s;
        s.type = t;
        s.material = m;
        s.transformation_inverse = transformation.inverse();  // e.g. sphere.rs:14-19
        return s;
    }
    static Shape triangle(const Vec3& p1, const Vec3& p2, const Vec3& p3) {  // triangle.rs:21-35
        Shape s;
        s.type = RTGPU_TRIANGLE;
        s.vertex_1 = p1;
        s.vertex_2 = p2;
        s.vertex_3 = p3;
        s.edge_1 = sub(p2, p1);
        s.edge_2 = sub(p3, p1);
        s.normal = normalized(cross(s.edge_2, s.edge_1));
        return s;
    }
    void set_transformation(const Transformation& t) { transformation_inverse = t.inverse(); }
    bool equals(const Shape& o) const {  // `dyn Shape == dyn Shape`, shapes/shape.rs:34-38
        if (type != o.type || !material.equals(o.material) || !(transformation_inverse == o.transformation_inverse)) return false;
        if ((type == RTGPU_CYLINDER || type == RTGPU_CONE) && (min != o.min || max != o.max || closed != o.closed)) return false;
        if (type == RTGPU_TRIANGLE && (vertex_1 != o.vertex_1 || vertex_2 != o.vertex_2 || vertex_3 != o.vertex_3)) return false;
        return true;
    }
};

struct World {  // composites/world.rs:9-12
    std::vector<Light> lights;
    std::vector<Shape> shapes;
    static constexpr uint32_t MAX_REFLECTION_ITERATIONS = 6;  // world.rs:15
};

// ---------------------------------------------------------------------------------------------
// Scene flattener: World -> the arrays of rtgpu_scene (owning)
struct FlatScene {
    std::vector<uint8_t> shape_type, shape_closed, mat_casts_shadow, pat_type;
    std::vector<double> shape_inv, shape_min, shape_max, tri_v1, tri_e1, tri_e2, tri_n, mat_color, mat_params, pat_a, pat_b, pat_inv,
        light_position, light_intensity;
    std::vector<int32_t> shape_triangle, mat_pattern, pat_child_a, pat_child_b;
    std::vector<uint32_t> shape_material, shape_eq_class;

    rtgpu_scene view() const {
        rtgpu_scene s{};
        s.abi_version = RTGPU_ABI_VERSION;
        s.n_shapes = (uint32_t)shape_type.size();
        s.shape_type = shape_type.data();
        s.shape_inv = shape_inv.data();
        s.shape_min = shape_min.data();
        s.shape_max = shape_max.data();
        s.shape_closed = shape_closed.data();
        s.shape_triangle = shape_triangle.data();
        s.shape_material = shape_material.data();
        s.shape_eq_class = shape_eq_class.data();
        s.n_triangles = (uint32_t)(tri_v1.size() / 3);
        s.tri_vertex_1 = tri_v1.data();
        s.tri_edge_1 = tri_e1.data();
        s.tri_edge_2 = tri_e2.data();
        s.tri_normal = tri_n.data();
        s.n_materials = (uint32_t)mat_pattern.size();
        s.mat_color = mat_color.data();
        s.mat_params = mat_params.data();
        s.mat_casts_shadow = mat_casts_shadow.data();
        s.mat_pattern = mat_pattern.data();
        s.n_patterns = (uint32_t)pat_type.size();
        s.pat_type = pat_type.data();
        s.pat_color_a = pat_a.data();
        s.pat_color_b = pat_b.data();
        s.pat_inv = pat_inv.data();
        s.pat_child_a = pat_child_a.data();
        s.pat_child_b = pat_child_b.data();
        s.n_lights = (uint32_t)(light_position.size() / 3);
        s.light_position = light_position.data();
        s.light_intensity = light_intensity.data();
        return s;
    }
};

inline void push_rows012(std::vector<double>& out, const Transformation& t) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) out.push_back(t.m[r][c]);
}

inline FlatScene flatten(const World& world) {
    FlatScene f;
    std::vector<const Material*> materials;
    std::vector<const Pattern*> patterns;
    // children before parents (the ABI requires child index < parent index)
    struct Rec {
        static int32_t pattern(FlatScene& f, std::vector<const Pattern*>& seen, const Pattern& p) {
            for (size_t i = 0; i < seen.size(); ++i)
                if (seen[i]->equals(p)) return (int32_t)i;
            int32_t ca = -1, cb = -1;
            if (p.type == RTGPU_PATTERN_COMPLEX) {
                ca = pattern(f, seen, *p.pattern_a);
                cb = pattern(f, seen, *p.pattern_b);
            }
            f.pat_type.push_back((uint8_t)p.type);
            for (int k = 0; k < 3; ++k) {
                f.pat_a.push_back(p.color_a[k]);
                f.pat_b.push_back(p.color_b[k]);
            }
            push_rows012(f.pat_inv, p.transformation_inverse);
            f.pat_child_a.push_back(ca);
            f.pat_child_b.push_back(cb);
            seen.push_back(&p);
            return (int32_t)seen.size() - 1;
        }
    };
    for (size_t i = 0; i < world.shapes.size(); ++i) {
        const Shape& s = world.shapes[i];
        const double* row3 = s.transformation_inverse.m[3];
        if (!(row3[0] == 0.0 && row3[1] == 0.0 && row3[2] == 0.0)) throw std::runtime_error("rtgpu: non-affine transformation_inverse");
        f.shape_type.push_back((uint8_t)s.type);
        push_rows012(f.shape_inv, s.transformation_inverse);
        const bool cc = s.type == RTGPU_CYLINDER || s.type == RTGPU_CONE;
        f.shape_min.push_back(cc ? s.min : 0.0);
        f.shape_max.push_back(cc ? s.max : 0.0);
        f.shape_closed.push_back(cc && s.closed ? 1 : 0);
        if (s.type == RTGPU_TRIANGLE) {
            f.shape_triangle.push_back((int32_t)(f.tri_v1.size() / 3));
            for (int k = 0; k < 3; ++k) {
                f.tri_v1.push_back(s.vertex_1[k]);
                f.tri_e1.push_back(s.edge_1[k]);
                f.tri_e2.push_back(s.edge_2[k]);
                f.tri_n.push_back(s.normal[k]);
            }
        } else {
            f.shape_triangle.push_back(-1);
        }
        // material de-duplicated by value
        int32_t mi = -1;
        for (size_t k = 0; k < materials.size(); ++k)
            if (materials[k]->equals(s.material)) {
                mi = (int32_t)k;
                break;
            }
        if (mi < 0) {
            const Material& m = s.material;
            const int32_t pat = m.pattern ? Rec::pattern(f, patterns, *m.pattern) : -1;
            for (int k = 0; k < 3; ++k) f.mat_color.push_back(m.color[k]);
            const double params[RTGPU_MAT_PARAM_COUNT] = {m.ambient, m.diffuse, m.specular, m.shininess, m.reflectiveness, m.transparency, m.refractive_index};
            f.mat_params.insert(f.mat_params.end(), params, params + RTGPU_MAT_PARAM_COUNT);
            f.mat_casts_shadow.push_back(m.casts_shadow ? 1 : 0);
            f.mat_pattern.push_back(pat);
            materials.push_back(&s.material);
            mi = (int32_t)materials.size() - 1;
        }
        f.shape_material.push_back((uint32_t)mi);
        uint32_t cls = (uint32_t)i;  // lowest index of a value-equal shape (shape.rs:34-38)
        for (size_t k = 0; k < i; ++k)
            if (world.shapes[k].equals(s)) {
                cls = (uint32_t)k;
                break;
            }
        f.shape_eq_class.push_back(cls);
    }
    for (const Light& l : world.lights)
        for (int k = 0; k < 3; ++k) {
            f.light_position.push_back(l.position[k]);
            f.light_intensity.push_back(l.intensity[k]);
        }
    return f;
}

// ---------------------------------------------------------------------------------------------
// composites/canvas.rs:13-17
struct Canvas {
    uint32_t width = 0, height = 0;
    std::vector<double> pixels;  // width*height*3, row-major, index = x + y*width (canvas.rs:44-55)
    std::vector<uint8_t> rgb8;   // the bytes Canvas::to_png_file encodes (canvas.rs:117-123), from the device

    std::string to_ppm() const;                         // canvas.rs:68-97
    void to_ppm_file(const std::string& path) const;    // canvas.rs:107-112
    void to_png_file(const std::string& path) const;    // canvas.rs:114-137 (RGB8, zlib)
};

// composites/camera.rs:10-19
struct Camera {
    uint32_t horizontal_size = 0, vertical_size = 0;
    double field_of_view = 0, half_width = 0, half_height = 0, pixel_size = 0;
    Transformation transformation_inverse = Transformation::identity();
    Vec3 origin{0, 0, 0};

    Camera() = default;
    Camera(uint32_t hsize, uint32_t vsize, double fov) : horizontal_size(hsize), vertical_size(vsize), field_of_view(fov) {  // camera.rs:25-49
        const double half_view = std::tan(fov / 2.0);
        const double aspect = (double)hsize / (double)vsize;
        if (aspect >= 1.0) {
            half_width = half_view;
            half_height = half_view / aspect;
        } else {
            half_width = half_view * aspect;
            half_height = half_view;
        }
        pixel_size = (half_width * 2.0) / (double)hsize;
    }
    void set_transformation(const Transformation& t) {  // camera.rs:124-127
        transformation_inverse = t.inverse();
        origin = transformation_inverse.mul_point({0, 0, 0});
    }
    void set_transformation_inverse(const Transformation& t) {  // camera.rs:133-136
        transformation_inverse = t;
        origin = transformation_inverse.mul_point({0, 0, 0});
    }
    Camera resized(uint32_t hsize, uint32_t vsize) const {
        Camera c(hsize, vsize, field_of_view);
        c.set_transformation_inverse(transformation_inverse);
        return c;
    }
    rtgpu_camera view() const {
        rtgpu_camera c{};
        c.hsize = horizontal_size;
        c.vsize = vertical_size;
        c.half_width = half_width;
        c.half_height = half_height;
        c.pixel_size = pixel_size;
        for (int r = 0; r < 3; ++r)
            for (int k = 0; k < 4; ++k) c.inv[r * 4 + k] = transformation_inverse.m[r][k];
        for (int k = 0; k < 3; ++k) c.origin[k] = origin[k];
        return c;
    }

    // The `RenderingMode::Gpu` arm (ray-tracer-cli/src/main.rs:18-21): same contract as
    // Camera::render_parallel (camera.rs:97-112).  Throws if the CUDA path is unavailable.
    Canvas render_gpu(const World& world, int n_gpus = 1, rtgpu_stats* stats = nullptr) const {
        const FlatScene flat = flatten(world);
        const rtgpu_scene scene = flat.view();
        const rtgpu_camera cam = view();
        rtgpu_opts opts{};
        opts.precision = RTGPU_PRECISION_F64;
        opts.max_depth = World::MAX_REFLECTION_ITERATIONS;
        opts.n_gpus = n_gpus;
        opts.band_rows = 16;
        Canvas canvas;
        canvas.width = horizontal_size;
        canvas.height = vertical_size;
        const size_t n = (size_t)horizontal_size * vertical_size;
        canvas.pixels.assign(n * 3, 0.0);
        canvas.rgb8.assign(n * 3, 0);
        const int st = rtgpu_render(&scene, &cam, &opts, canvas.pixels.data(), canvas.rgb8.data(), stats);
        if (st != RTGPU_OK) throw std::runtime_error(std::string("rtgpu_render failed: ") + rtgpu_last_error());
        return canvas;
    }
};

}  // namespace rt_host
