// Image / RenderResult (reference image.hpp:8-45): float RGB planes, row 0 = image top.
// save() writes an 8-bit RGB PNG (".png": the pixels cv::imwrite stores for the reference's matrix, image.cpp:7-19; own
// encoder, host/src/png.cpp), a binary PPM (".ppm") or a PFM (anything else).  The reference's denoiser is OIDN
// (image.cpp:21-95), out of scope and absent here: denoise() only reports that.
#pragma once

#include <cstddef>
#include <string>
#include <vector>

class Image {
public:
    Image(size_t height_, size_t width_) : height(height_), width(width_), color_buffer(height_ * width_ * 3) {}
    Image(const std::vector<float>& buffer, size_t height_, size_t width_) : height(height_), width(width_), color_buffer(buffer) {}
    Image(std::vector<float>&& buffer, size_t height_, size_t width_) : height(height_), width(width_), color_buffer(buffer) {}
    virtual ~Image() {}
    // (declaring the destructor would otherwise take the implicit move operations away, which the reference's Image --
    // no declared destructor -- has: returning a RenderResult by value must not copy its planes)
    Image(const Image&) = default;
    Image(Image&&) = default;
    Image& operator=(const Image&) = default;
    Image& operator=(Image&&) = default;

    void save(const std::string& filename, float gamma = 1.0) const;
    virtual void denoise(bool verbose = false);

    size_t height;
    size_t width;
    std::vector<float> color_buffer;
};

class RenderResult : public Image {
public:
    RenderResult(size_t height_, size_t width_)
        : Image(height_, width_), normal_buffer(height_ * width_ * 3), albedo_buffer(height_ * width_ * 3) {}
    RenderResult(std::vector<float>&& color, std::vector<float>&& normal, std::vector<float>&& albedo, size_t height_, size_t width_)
        : Image(std::move(color), height_, width_), normal_buffer(std::move(normal)), albedo_buffer(std::move(albedo)) {}

    void denoise(bool verbose = false) override;
    void save_normal(const std::string& filename) const;
    void save_albedo(const std::string& filename) const;

    std::vector<float> normal_buffer;
    std::vector<float> albedo_buffer;
};
