// 8-bit RGB PNG writer behind Image::save (the reference encodes through cv::imwrite, image.cpp:16-18).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>

namespace qzhost {

// rows top-down, 3 bytes per pixel in R, G, B order; false if the file cannot be written
bool write_png_rgb8(const std::string& filename, const unsigned char* rgb, size_t width, size_t height);

}  // namespace qzhost
