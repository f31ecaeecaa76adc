// scene_loader.hpp — YAML scene loader of the C++ host: restates ray-tracer-cli/src/scene_loader.rs
// (line numbers below are in that file) on top of a small YAML-subset reader.
//
// The reference uses yaml-rust; neither it nor yaml-cpp exists here, so `yaml_lite` parses the subset
// the scene descriptions use: block sequences of block mappings, nested block mappings and sequences
// by indentation, flow sequences (`[ translate, 1, -1, 1 ]`, nested), plain / quoted scalars,
// comments.  Scalars are typed like yaml-rust types them: integers, reals (anything `strtod` accepts
// completely, e.g. `1e3`), `true` / `false`, else strings.
#pragma once

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "rt_host.hpp"

namespace rt_host {

namespace yaml_lite {

struct Node {
    enum Kind { Bad, Null, Bool, Integer, Real, String, Array, Hash } kind = Bad;
    bool b = false;
    long long i = 0;
    double r = 0.0;
    std::string s;
    std::vector<Node> items;
    std::vector<std::pair<std::string, Node>> fields;  // insertion order

    const Node& operator[](const std::string& key) const;  // yaml[key]: Bad unless a hash with that key
    bool is_bad() const { return kind == Bad; }
};

Node parse(const std::string& text);  // throws std::runtime_error on syntax the subset does not cover

}  // namespace yaml_lite

// load_scene_description (:361-367).  Throws std::runtime_error where the reference returns Err / panics.
std::pair<World, Camera> load_scene_description(const std::string& path);
std::pair<World, Camera> load_scene_from_string(const std::string& text);

}  // namespace rt_host
