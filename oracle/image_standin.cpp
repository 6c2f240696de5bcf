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

// The arithmetic of Image::save (image.cpp:10-15), restated because image.cpp itself needs OpenCV: x255, powf gamma,
// B and R swapped.  bgr8 is what cv::imwrite stores for an 8-bit file from that CV_32FC3 matrix (OpenCV converts with
// saturate_cast<uchar>(float) = cvRound, i.e. lrint, clamped to 0..255) -- OpenCV is absent: that conversion is
// recalled from its documentation, parity at the OpenCV boundary is UNPINNED.
#include <cmath>
extern "C" int orc_tone(const float* rgb, int n_pixels, float gamma, float* bgr255, unsigned char* bgr8) {
    for (long i = 0; i < n_pixels; i++) {
        for (int k = 0; k < 3; k++) {
            const float v = 255.0f * powf(rgb[3 * i + 2 - k], gamma);
            if (bgr255) bgr255[3 * i + k] = v;
            if (bgr8) {
                const long r = v != v ? 0 : lrintf(v < -1.0f ? -1.0f : (v > 256.0f ? 256.0f : v));
                bgr8[3 * i + k] = (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    }
    return 0;
}
