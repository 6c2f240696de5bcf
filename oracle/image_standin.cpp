// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the reference's src/image.cpp, which needs OpenCV and OIDN (neither is
// installed here, and both are out of scope: image.cpp:7-19 saves a PNG, :47-95 runs
// the OIDN denoiser -- an untimed post-process per BASELINE.json).  save() writes a
// little-endian PFM instead of a PNG; denoise() is a no-op.
#include <cstdio>

#include "image.hpp"

static void write_pfm(const std::string& filename, const std::vector<float>& buf, size_t w, size_t h) {
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return;
    std::fprintf(f, "PF\n%zu %zu\n-1.0\n", w, h);
    for (size_t row = h; row-- > 0;) std::fwrite(buf.data() + row * w * 3, sizeof(float), w * 3, f);
    std::fclose(f);
}

void Image::save(const std::string& filename, float) const { write_pfm(filename, color_buffer, width, height); }
void Image::denoise(bool) {}
void RenderResult::denoise(bool) {}
void RenderResult::save_normal(const std::string& filename) const { write_pfm(filename, normal_buffer, width, height); }
void RenderResult::save_albedo(const std::string& filename) const { write_pfm(filename, albedo_buffer, width, height); }
