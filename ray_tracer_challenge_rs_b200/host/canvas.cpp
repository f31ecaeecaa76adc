// canvas.cpp — Canvas output of the C++ host: PPM (canvas.rs:68-112) and RGB8 PNG (canvas.rs:114-137).
#include <zlib.h>

#include <cstring>
#include <fstream>

#include "rt_host.hpp"

namespace rt_host {

namespace {

// canvas.rs:117-123 for callers that filled `pixels` themselves; render_gpu already returns the device's bytes
std::vector<uint8_t> quantise(const std::vector<double>& pixels) {
    std::vector<uint8_t> out(pixels.size());
    for (size_t i = 0; i < pixels.size(); ++i) {
        double v = pixels[i];
        v = v < 0.0 ? 0.0 : v;
        v = v > 1.0 ? 1.0 : v;
        v = std::round(v * 255.0);
        out[i] = (v != v) ? 0 : (uint8_t)v;
    }
    return out;
}

void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}

void chunk(std::vector<uint8_t>& png, const char type[4], const std::vector<uint8_t>& data) {
    put_be32(png, (uint32_t)data.size());
    const size_t start = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    put_be32(png, (uint32_t)crc32(0L, png.data() + start, (uInt)(png.size() - start)));
}

}  // namespace

std::string Canvas::to_ppm() const {
    const std::vector<uint8_t> bytes = rgb8.empty() ? quantise(pixels) : rgb8;
    std::string out = "P3\n" + std::to_string(width) + " " + std::to_string(height) + "\n255";
    const size_t pixels_per_line = 5;  // floor(70 / (3 * 4)), canvas.rs:76
    const size_t n = (size_t)width * height;
    char buf[8];
    for (size_t p = 0; p < n; ++p) {
        out += (p % pixels_per_line == 0) ? "\n" : " ";
        for (int c = 0; c < 3; ++c) {
            snprintf(buf, sizeof(buf), "%3d", (int)bytes[p * 3 + c]);  // right-aligned to width 3 (canvas.rs:83-91)
            if (c) out += " ";
            out += buf;
        }
    }
    return out;
}

void Canvas::to_ppm_file(const std::string& path) const {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    const std::string s = to_ppm();
    f.write(s.data(), (std::streamsize)s.size());
}

void Canvas::to_png_file(const std::string& path) const {
    const std::vector<uint8_t> bytes = rgb8.empty() ? quantise(pixels) : rgb8;
    // filter type 0 ("NoFilter", canvas.rs:128) in front of every scanline
    std::vector<uint8_t> raw;
    raw.reserve((size_t)height * ((size_t)width * 3 + 1));
    for (uint32_t y = 0; y < height; ++y) {
        raw.push_back(0);
        raw.insert(raw.end(), bytes.begin() + (size_t)y * width * 3, bytes.begin() + (size_t)(y + 1) * width * 3);
    }
    uLongf bound = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(bound);
    if (compress2(z.data(), &bound, raw.data(), (uLong)raw.size(), Z_BEST_COMPRESSION) != Z_OK) throw std::runtime_error("zlib compress failed");
    z.resize(bound);
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width);
    put_be32(ihdr, height);
    ihdr.push_back(8);  // bit depth
    ihdr.push_back(2);  // colour type RGB (ExtendedColorType::Rgb8, canvas.rs:135)
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", z);
    chunk(png, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write " + path);
    f.write((const char*)png.data(), (std::streamsize)png.size());
}

}  // namespace rt_host
