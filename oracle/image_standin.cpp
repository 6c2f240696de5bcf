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
// B and R swapped.  bgr8 is what cv::imwrite stores for an 8-bit file from that CV_32FC3 matrix: OpenCV converts with
// convertTo(CV_8U) -- cvRound (the x86 float -> int32 conversion: nearest even; INT_MIN for NaN, infinities and values
// beyond int32), then saturation to 0..255.  The C++ OpenCV is absent here, but the image has its Python build:
// tests/test_abi_and_host.py pins this function against cv2.imwrite / cv2.imread (OpenCV 4.13) on the same matrix.
#include <cmath>
extern "C" int orc_tone(const float* rgb, int n_pixels, float gamma, float* bgr255, unsigned char* bgr8) {
    for (long i = 0; i < n_pixels; i++) {
        for (int k = 0; k < 3; k++) {
            const float v = 255.0f * powf(rgb[3 * i + 2 - k], gamma);
            if (bgr255) bgr255[3 * i + k] = v;
            if (bgr8) {
                const bool in_int32 = std::fabs(v) < 2147483648.0f;   // false for NaN too
                const long r = in_int32 ? lrintf(v) : 0;
                bgr8[3 * i + k] = (unsigned char)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    }
    return 0;
}
