// scene_loader.cpp — see scene_loader.hpp.  Line numbers in comments: ray-tracer-cli/src/scene_loader.rs.
#include "scene_loader.hpp"

#include <cctype>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rt_host {
namespace yaml_lite {

namespace {

const Node& bad_node() {
    static const Node n;
    return n;
}

struct Line {
    int indent;
    std::string text;  // without indentation, comment and trailing blanks
    int number;
};

std::string rstrip(std::string s) {
    while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
    return s;
}

std::string strip(const std::string& s) {
    size_t a = 0;
    while (a < s.size() && (s[a] == ' ' || s[a] == '\t')) ++a;
    return rstrip(s.substr(a));
}

// a '#' starts a comment at the beginning of the text or after white space, outside quotes
std::string strip_comment(const std::string& s) {
    char quote = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (quote) {
            if (c == quote) quote = 0;
        } else if (c == '"' || c == '\'') {
            quote = c;
        } else if (c == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) {
            return s.substr(0, i);
        }
    }
    return s;
}

[[noreturn]] void fail(const Line& l, const std::string& what) {
    throw std::runtime_error("yaml line " + std::to_string(l.number) + ": " + what);
}

// yaml-rust scalar typing: integers, reals (str::parse::<f64> succeeds), booleans, null, else string
Node scalar(const std::string& raw) {
    Node n;
    std::string s = strip(raw);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
        n.kind = Node::String;
        n.s = s.substr(1, s.size() - 2);
        return n;
    }
    if (s.empty() || s == "~" || s == "null") {
        n.kind = Node::Null;
        return n;
    }
    if (s == "true" || s == "false") {
        n.kind = Node::Bool;
        n.b = s == "true";
        return n;
    }
    char* end = nullptr;
    errno = 0;
    const long long iv = std::strtoll(s.c_str(), &end, 10);
    if (errno == 0 && end && *end == '\0' && end != s.c_str()) {
        n.kind = Node::Integer;
        n.i = iv;
        return n;
    }
    // reals: what Rust's f64::from_str accepts; exclude the hex / "nan(...)" forms strtod also takes
    bool plausible = true;
    for (char c : s)
        if (!(std::isdigit((unsigned char)c) || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-')) plausible = false;
    if (plausible) {
        const double rv = std::strtod(s.c_str(), &end);  // correctly rounded, like str::parse::<f64>
        if (end && *end == '\0' && end != s.c_str()) {
            n.kind = Node::Real;
            n.r = rv;
            return n;
        }
    }
    n.kind = Node::String;
    n.s = s;
    return n;
}

// flow sequence "[ a, [b, c], d ]" starting at s[pos] == '['; pos ends after the closing bracket
Node flow_sequence(const std::string& s, size_t& pos, const Line& l) {
    Node n;
    n.kind = Node::Array;
    ++pos;  // '['
    std::string cur;
    bool have = false;
    for (;;) {
        if (pos >= s.size()) fail(l, "unterminated flow sequence");
        const char c = s[pos];
        if (c == '[') {
            n.items.push_back(flow_sequence(s, pos, l));
            have = false;
            cur.clear();
            // skip to ',' or ']'
            while (pos < s.size() && (s[pos] == ' ' || s[pos] == '\t')) ++pos;
            if (pos < s.size() && s[pos] == ',') ++pos;
            continue;
        }
        if (c == ',' || c == ']') {
            if (have || !strip(cur).empty()) n.items.push_back(scalar(cur));
            cur.clear();
            have = false;
            ++pos;
            if (c == ']') return n;
            continue;
        }
        cur.push_back(c);
        if (c != ' ' && c != '\t') have = true;
        ++pos;
    }
}

Node value_of(const std::string& text, const Line& l) {
    const std::string s = strip(text);
    if (!s.empty() && s[0] == '[') {
        size_t pos = 0;
        Node n = flow_sequence(s, pos, l);
        if (!strip(s.substr(pos)).empty()) fail(l, "trailing characters after flow sequence");
        return n;
    }
    if (!s.empty() && s[0] == '{') fail(l, "flow mappings are not supported");
    return scalar(s);
}

// position of the ':' that ends a mapping key ("key: value" or "key:"), or npos
size_t key_colon(const std::string& s) {
    if (s.empty() || s[0] == '[' || s[0] == '"' || s[0] == '\'') return std::string::npos;
    for (size_t i = 0; i < s.size(); ++i)
        if (s[i] == ':' && (i + 1 == s.size() || s[i + 1] == ' ')) return i;
    return std::string::npos;
}

struct Parser {
    std::vector<Line> lines;
    size_t at = 0;

    Node block(int indent) {
        if (at >= lines.size() || lines[at].indent < indent) return Node{};  // Bad: nothing there
        const int here = lines[at].indent;
        const std::string& t = lines[at].text;
        if (t == "-" || t.rfind("- ", 0) == 0) return sequence(here);
        if (key_colon(t) != std::string::npos) return mapping(here);
        Node n = value_of(t, lines[at]);
        ++at;
        return n;
    }

    Node sequence(int indent) {
        Node n;
        n.kind = Node::Array;
        while (at < lines.size() && lines[at].indent == indent && (lines[at].text == "-" || lines[at].text.rfind("- ", 0) == 0)) {
            Line& l = lines[at];
            std::string rest = l.text.size() > 1 ? l.text.substr(2) : std::string();
            size_t lead = 0;
            while (lead < rest.size() && rest[lead] == ' ') ++lead;
            rest = rest.substr(lead);
            if (rest.empty()) {
                ++at;
                n.items.push_back(block(indent + 1));
            } else if (key_colon(rest) != std::string::npos || rest == "-" || rest.rfind("- ", 0) == 0) {
                // the item is a block (mapping or nested sequence) whose first line shares this line
                l.indent = indent + 2 + (int)lead;
                l.text = rest;
                n.items.push_back(block(l.indent));
            } else {
                n.items.push_back(value_of(rest, l));
                ++at;
            }
        }
        return n;
    }

    Node mapping(int indent) {
        Node n;
        n.kind = Node::Hash;
        while (at < lines.size() && lines[at].indent == indent) {
            const Line& l = lines[at];
            const size_t colon = key_colon(l.text);
            if (colon == std::string::npos) break;
            std::string key = strip(l.text.substr(0, colon));
            const std::string rest = strip(l.text.substr(colon + 1));
            Node value;
            ++at;
            if (!rest.empty()) {
                value = value_of(rest, l);
            } else if (at < lines.size() && lines[at].indent > indent) {
                value = block(lines[at].indent);
            } else if (at < lines.size() && lines[at].indent == indent && (lines[at].text == "-" || lines[at].text.rfind("- ", 0) == 0)) {
                value = sequence(indent);  // "key:" followed by a sequence at the same indentation
            } else {
                value.kind = Node::Null;
            }
            n.fields.emplace_back(std::move(key), std::move(value));
        }
        return n;
    }
};

}  // namespace

const Node& Node::operator[](const std::string& key) const {
    if (kind != Hash) return bad_node();
    for (const auto& f : fields)
        if (f.first == key) return f.second;
    return bad_node();
}

Node parse(const std::string& text) {
    Parser p;
    std::istringstream in(text);
    std::string raw;
    int number = 0;
    while (std::getline(in, raw)) {
        ++number;
        std::string s = rstrip(strip_comment(raw));
        if (strip(s).empty()) continue;
        if (s.rfind("---", 0) == 0 || s.rfind("...", 0) == 0) continue;  // single-document files only
        int indent = 0;
        while ((size_t)indent < s.size() && s[indent] == ' ') ++indent;
        if ((size_t)indent < s.size() && s[indent] == '\t') throw std::runtime_error("yaml line " + std::to_string(number) + ": tab indentation");
        p.lines.push_back({indent, s.substr(indent), number});
    }
    if (p.lines.empty()) return Node{};
    Node root = p.block(p.lines[0].indent);
    if (p.at != p.lines.size()) throw std::runtime_error("yaml line " + std::to_string(p.lines[p.at].number) + ": unexpected indentation");
    return root;
}

}  // namespace yaml_lite

// ---------------------------------------------------------------------------------------------

namespace {

using yaml_lite::Node;

[[noreturn]] void parse_float_error() { throw std::runtime_error("cannot parse float from empty string"); }

// :338-344
double parse_f64(const Node& n) {
    if (n.kind == Node::Integer) return (double)n.i;
    if (n.kind == Node::Real) return n.r;
    parse_float_error();
}

// :346-352
Vec3 parse_array_of_3(const std::vector<Node>& v, size_t first = 0) {
    if (v.size() < first + 3) throw std::runtime_error("expected three numbers");
    return {parse_f64(v[first]), parse_f64(v[first + 1]), parse_f64(v[first + 2])};
}

const std::vector<Node>& as_vec(const Node& n) {
    if (n.kind != Node::Array) throw std::runtime_error("expected a sequence");  // .as_vec().unwrap()
    return n.items;
}

uint32_t as_u32(double v) {  // Rust `f64 as u32`
    if (v != v || v <= 0.0) return 0;
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;
}

bool ends_with(const std::string& s, const char* suffix) {
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

struct SceneParser {
    std::map<std::string, Vec3> colors;
    std::map<std::string, Material> materials;
    std::map<std::string, Transformation> transformations;

    template <typename M>
    static const typename M::mapped_type& lookup(const M& m, const std::string& key, const char* what) {
        auto it = m.find(key);
        if (it == m.end()) throw std::runtime_error(std::string("undefined ") + what + " '" + key + "'");  // HashMap index panics
        return it->second;
    }

    // :46-62
    void process_definitions(const Node& doc) {
        if (doc.kind != Node::Array) return;
        for (const Node& entry : doc.items) {
            const Node& name = entry["define"];
            if (name.kind != Node::String) continue;
            if (ends_with(name.s, "-color")) colors[name.s] = parse_color(entry);
            else if (ends_with(name.s, "-material")) materials[name.s] = parse_material(entry);
            else if (ends_with(name.s, "-transform") || ends_with(name.s, "-object")) transformations[name.s] = parse_transformation(entry);
        }
    }

    // :64-90
    Vec3 parse_color(const Node& n) const {
        if (n.kind == Node::Hash) return parse_color(!n["color"].is_bad() ? n["color"] : n["value"]);
        if (n.kind == Node::Array) return parse_array_of_3(n.items);
        if (n.kind == Node::String) return lookup(colors, n.s, "color");
        throw std::runtime_error("Incorrect color value");
    }

    // :92-149
    Material parse_material(const Node& node) const {
        if (node.kind == Node::String) return lookup(materials, node.s, "material");
        Material material;
        if (!node["extend"].is_bad()) {
            if (node["extend"].kind != Node::String) throw std::runtime_error("extend: expected a name");
            material = lookup(materials, node["extend"].s, "material");
        }
        const Node& y = !node["value"].is_bad() ? node["value"] : node;
        if (!y["color"].is_bad()) material.color = parse_color(y["color"]);
        if (!y["pattern"].is_bad()) material.pattern = parse_pattern(y["pattern"]);
        if (!y["ambient"].is_bad()) material.ambient = parse_f64(y["ambient"]);
        if (!y["diffuse"].is_bad()) material.diffuse = parse_f64(y["diffuse"]);
        if (!y["specular"].is_bad()) material.specular = parse_f64(y["specular"]);
        if (!y["shininess"].is_bad()) material.shininess = parse_f64(y["shininess"]);
        if (!y["reflective"].is_bad()) material.reflectiveness = parse_f64(y["reflective"]);
        if (!y["transparency"].is_bad()) material.transparency = parse_f64(y["transparency"]);
        if (!y["refractive-index"].is_bad()) material.refractive_index = parse_f64(y["refractive-index"]);
        if (y["casts-shadow"].kind == Node::Bool) material.casts_shadow = y["casts-shadow"].b;
        return material;
    }

    // :151-193
    std::shared_ptr<Pattern> parse_pattern(const Node& n) const {
        const std::vector<Node>& cols = as_vec(n["colors"]);
        if (cols.size() < 2) throw std::runtime_error("pattern: two colors expected");
        const Vec3 a = parse_color(cols[0]), b = parse_color(cols[1]);
        const bool has_t = !n["transform"].is_bad();
        Transformation t = Transformation::identity();
        if (has_t) t = parse_transformation(n["transform"]);
        rtgpu_pattern_type type;
        const Node& ty = n["type"];
        if (ty.kind != Node::String) throw std::runtime_error("Incorrect pattern type");
        if (ty.s == "stripes") type = RTGPU_PATTERN_STRIPE;
        else if (ty.s == "gradient") type = RTGPU_PATTERN_GRADIENT;
        else if (ty.s == "rings") type = RTGPU_PATTERN_RING;
        else if (ty.s == "checkers") type = RTGPU_PATTERN_CHECKER;
        else throw std::runtime_error("Incorrect pattern type");
        auto p = Pattern::two_color(type, a, b);
        if (has_t) p->set_transformation(t);
        return p;
    }

    // :195-238: a NAMED transform is right-multiplied, an inline op is left-multiplied
    Transformation parse_transformation(const Node& node) const {
        Transformation t = Transformation::identity();
        const Node& y = !node["value"].is_bad() ? node["value"] : node;
        if (y.kind != Node::Array) return t;
        for (const Node& tr : y.items) {
            if (tr.kind == Node::String) {
                t = t * lookup(transformations, tr.s, "transform");
            } else if (tr.kind == Node::Array) {
                if (tr.items.empty() || tr.items[0].kind != Node::String) throw std::runtime_error("transform: operation name expected");
                const std::string& op = tr.items[0].s;
                if (op == "scale") {
                    const Vec3 v = parse_array_of_3(tr.items, 1);
                    t = transformations::scaling(v[0], v[1], v[2]) * t;
                } else if (op == "translate") {
                    const Vec3 v = parse_array_of_3(tr.items, 1);
                    t = transformations::translation(v[0], v[1], v[2]) * t;
                } else if (op == "rotate-x" || op == "rotate-y" || op == "rotate-z") {
                    if (tr.items.size() < 2) parse_float_error();
                    const double a = parse_f64(tr.items[1]);
                    const Transformation r = op == "rotate-x" ? transformations::rotation_x(a) : op == "rotate-y" ? transformations::rotation_y(a) : transformations::rotation_z(a);
                    t = r * t;
                }  // unknown operations are ignored (:232)
            }
        }
        return t;
    }

    // :249-335
    std::pair<World, Camera> parse_scene(const Node& doc) const {
        World world;
        Camera camera(0, 0, 0.0);
        if (doc.kind != Node::Array) return {world, camera};
        for (const Node& entry : doc.items) {
            const Node& add = entry["add"];
            if (add.kind != Node::String) continue;
            const std::string& name = add.s;
            if (name == "camera") {
                const uint32_t w = as_u32(parse_f64(entry["width"])), h = as_u32(parse_f64(entry["height"]));
                const double fov = parse_f64(entry["field-of-view"]);
                const Vec3 from = parse_array_of_3(as_vec(entry["from"])), to = parse_array_of_3(as_vec(entry["to"])), up = parse_array_of_3(as_vec(entry["up"]));
                camera = Camera(w, h, fov);
                camera.set_transformation(transformations::view_transform(from, to, up));
            } else if (name == "light") {
                Light l;
                l.position = parse_array_of_3(as_vec(entry["at"]));
                l.intensity = parse_array_of_3(as_vec(entry["intensity"]));
                world.lights.push_back(l);
            } else if (name == "plane" || name == "sphere" || name == "cube" || name == "cone" || name == "cylinder") {
                const Material material = parse_material(entry["material"]);
                const Transformation transformation = parse_transformation(entry["transform"]);
                const rtgpu_shape_type type = name == "plane" ? RTGPU_PLANE : name == "sphere" ? RTGPU_SPHERE : name == "cube" ? RTGPU_CUBE : name == "cone" ? RTGPU_CONE : RTGPU_CYLINDER;
                Shape s = Shape::make(type, material, transformation);
                if (type == RTGPU_CONE || type == RTGPU_CYLINDER) {  // :296-329
                    if (entry["closed"].kind == Node::Bool) s.closed = entry["closed"].b;
                    if (!entry["max"].is_bad()) s.max = parse_f64(entry["max"]);
                    if (!entry["min"].is_bad()) s.min = parse_f64(entry["min"]);
                }
                world.shapes.push_back(s);
            }  // anything else is ignored (:330)
        }
        return {world, camera};
    }
};

}  // namespace

std::pair<World, Camera> load_scene_from_string(const std::string& text) {
    const Node doc = yaml_lite::parse(text);
    SceneParser parser;
    parser.process_definitions(doc);
    return parser.parse_scene(doc);
}

std::pair<World, Camera> load_scene_description(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return load_scene_from_string(ss.str());
}

}  // namespace rt_host
