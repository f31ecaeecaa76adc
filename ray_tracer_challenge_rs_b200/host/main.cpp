// main.cpp — `ray-tracer-cli` of the C++ host: the reference's CLI (ray-tracer-cli/src/main.rs:11-31,
// cli/cli_arguments.rs:6-13, cli/rendering_mode.rs:3-7) with the new `gpu` rendering mode.
//
//   ray-tracer-cli <SCENE_PATH> <IMAGE_OUTPUT_PATH> [-r|--rendering-mode gpu] [-q|--quiet]
//                  [--width W --height H] [--gpus N] [--dump-flat FILE]
//
// `serial` and `parallel` are the reference's CPU loops (camera.rs:79-112), which this build does not
// contain: asking for them is an error, never a silent fallback.  --width/--height override the scene's
// camera size (the reference has no size flag; benches need one).  --dump-flat writes the flattened scene
// (the exact bytes handed to rtgpu_render) and exits without rendering — used by the CPU tests.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>

#include "rt_host.hpp"
#include "scene_loader.hpp"

using namespace rt_host;

namespace {

template <typename T>
void dump(std::ofstream& f, const char* name, const std::vector<T>& v) {
    const uint64_t n = v.size(), width = sizeof(T);
    const uint32_t name_len = (uint32_t)strlen(name);
    f.write((const char*)&name_len, 4);
    f.write(name, name_len);
    f.write((const char*)&width, 8);
    f.write((const char*)&n, 8);
    f.write((const char*)v.data(), (std::streamsize)(n * sizeof(T)));
}

void dump_flat(const std::string& path, const FlatScene& s, const Camera& cam) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    dump(f, "shape_type", s.shape_type);
    dump(f, "shape_inv", s.shape_inv);
    dump(f, "shape_min", s.shape_min);
    dump(f, "shape_max", s.shape_max);
    dump(f, "shape_closed", s.shape_closed);
    dump(f, "shape_triangle", s.shape_triangle);
    dump(f, "shape_material", s.shape_material);
    dump(f, "shape_eq_class", s.shape_eq_class);
    dump(f, "tri_vertex_1", s.tri_v1);
    dump(f, "tri_edge_1", s.tri_e1);
    dump(f, "tri_edge_2", s.tri_e2);
    dump(f, "tri_normal", s.tri_n);
    dump(f, "mat_color", s.mat_color);
    dump(f, "mat_params", s.mat_params);
    dump(f, "mat_casts_shadow", s.mat_casts_shadow);
    dump(f, "mat_pattern", s.mat_pattern);
    dump(f, "pat_type", s.pat_type);
    dump(f, "pat_color_a", s.pat_a);
    dump(f, "pat_color_b", s.pat_b);
    dump(f, "pat_inv", s.pat_inv);
    dump(f, "pat_child_a", s.pat_child_a);
    dump(f, "pat_child_b", s.pat_child_b);
    dump(f, "light_position", s.light_position);
    dump(f, "light_intensity", s.light_intensity);
    const rtgpu_camera c = cam.view();
    std::vector<uint8_t> cam_bytes((const uint8_t*)&c, (const uint8_t*)&c + sizeof(c));
    dump(f, "camera", cam_bytes);
}

int usage(const char* argv0) {
    fprintf(stderr, "Usage: %s <SCENE_PATH> <IMAGE_OUTPUT_PATH> [-r gpu] [-q] [--width W --height H] [--gpus N] [--dump-flat FILE]\n", argv0);
    return 2;
}

}  // namespace

int main(int argc, char** argv) {
    std::string scene_path, output_path, mode = "gpu", dump_path;
    bool quiet = false;
    long width = 0, height = 0, gpus = 1;
    std::vector<std::string> positional;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&](const char* what) -> std::string {
            if (i + 1 >= argc) {
                fprintf(stderr, "error: %s needs a value\n", what);
                exit(2);
            }
            return argv[++i];
        };
        if (a == "-r" || a == "--rendering-mode") mode = next("--rendering-mode");
        else if (a.rfind("--rendering-mode=", 0) == 0) mode = a.substr(17);
        else if (a == "-q" || a == "--quiet") quiet = true;
        else if (a == "--width") width = atol(next("--width").c_str());
        else if (a == "--height") height = atol(next("--height").c_str());
        else if (a == "--gpus") gpus = atol(next("--gpus").c_str());
        else if (a == "--dump-flat") dump_path = next("--dump-flat");
        else if (a == "-h" || a == "--help") return usage(argv[0]);
        else positional.push_back(a);
    }
    if (positional.size() < (dump_path.empty() ? 2u : 1u)) return usage(argv[0]);
    scene_path = positional[0];
    if (positional.size() > 1) output_path = positional[1];
    try {
        auto loaded = load_scene_description(scene_path);  // main.rs:13
        World& world = loaded.first;
        Camera camera = loaded.second;
        if (width > 0 && height > 0) camera = camera.resized((uint32_t)width, (uint32_t)height);
        if (!dump_path.empty()) {
            dump_flat(dump_path, flatten(world), camera);
            return 0;
        }
        if (mode == "serial" || mode == "parallel") {
            fprintf(stderr, "error: rendering mode '%s' is the reference's CPU loop and is not part of this build; use --rendering-mode gpu\n", mode.c_str());
            return 2;
        }
        if (mode != "gpu") {
            fprintf(stderr, "error: invalid value '%s' for '--rendering-mode' [possible values: serial, parallel, gpu]\n", mode.c_str());
            return 2;
        }
        if (!quiet) printf("Rendering image using scene at %s\n", scene_path.c_str());  // main.rs:14-16
        const auto t0 = std::chrono::steady_clock::now();
        rtgpu_stats stats{};
        const Canvas canvas = camera.render_gpu(world, (int)gpus, &stats);
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (!quiet) {
            printf("Image rendered in: %.3fs\n", seconds);  // main.rs:23-25
            const double rays = (double)(stats.rays_primary + stats.rays_shadow + stats.rays_reflect + stats.rays_refract);
            printf("  %.0f rays, kernel %.3f ms, %.1f Mrays/s\n", rays, stats.kernel_ms, rays / (stats.kernel_ms * 1e-3) / 1e6);
        }
        const bool ppm = output_path.size() >= 4 && output_path.compare(output_path.size() - 4, 4, ".ppm") == 0;
        if (ppm) canvas.to_ppm_file(output_path);
        else canvas.to_png_file(output_path);  // main.rs:26
        if (!quiet) printf("Image saved at %s\n", output_path.c_str());
    } catch (const std::exception& e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
    return 0;
}
