// Minimal PNG writer for Image::save (reference image.cpp:7-19 hands a CV_32FC3 matrix to cv::imwrite, which converts it
// to 8 bits and encodes an RGB PNG; OpenCV is not part of this build).  8-bit RGB, no interlace, filter 0 on every row,
// the zlib stream made of stored (uncompressed) deflate blocks: any decoder returns exactly the bytes handed in, which
// is all the drop-in needs -- the files are larger than OpenCV's, the pixels are the same.
#include "../png.hpp"

#include <algorithm>
#include <cstdio>
#include <vector>

namespace qzhost {

namespace {

uint32_t crc32(const unsigned char* p, size_t n, uint32_t crc = 0) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}

void put_be32(std::vector<unsigned char>& out, uint32_t v) {
    for (int shift = 24; shift >= 0; shift -= 8) out.push_back((unsigned char)(v >> shift));
}

bool write_chunk(std::FILE* f, const char type[4], const std::vector<unsigned char>& data) {
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)data.size());
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32(head.data() + 4, 4);
    crc = crc32(data.data(), data.size(), crc);
    std::vector<unsigned char> tail;
    put_be32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (data.empty() || std::fwrite(data.data(), 1, data.size(), f) == data.size()) &&
           std::fwrite(tail.data(), 1, tail.size(), f) == tail.size();
}

}  // namespace

bool write_png_rgb8(const std::string& filename, const unsigned char* rgb, size_t width, size_t height) {
    if (width == 0 || height == 0 || width > 0x7fffffffu || height > 0x7fffffffu) return false;
    std::FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    static const unsigned char signature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = std::fwrite(signature, 1, 8, f) == 8;

    std::vector<unsigned char> ihdr;
    put_be32(ihdr, (uint32_t)width);
    put_be32(ihdr, (uint32_t)height);
    const unsigned char format[5] = {8 /* bits */, 2 /* RGB */, 0 /* deflate */, 0 /* adaptive filtering */, 0 /* no interlace */};
    ihdr.insert(ihdr.end(), format, format + 5);
    ok = ok && write_chunk(f, "IHDR", ihdr);

    // scanlines: filter type 0, then the row's bytes
    const size_t stride = 1 + 3 * width;
    std::vector<unsigned char> raw(stride * height);
    for (size_t y = 0; y < height; y++) {
        raw[y * stride] = 0;
        std::copy(rgb + y * 3 * width, rgb + (y + 1) * 3 * width, raw.begin() + y * stride + 1);
    }
    // zlib: header, stored blocks of at most 65535 bytes, Adler-32 of the raw data
    std::vector<unsigned char> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78);
    z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t pos = 0; pos < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        z.push_back(pos + n == raw.size() ? 1 : 0);   // BFINAL, BTYPE = 00
        z.push_back((unsigned char)(n & 0xff));
        z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff));
        z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = pos; i < pos + n;) {   // (5552 bytes is the longest run before the sums can overflow 32 bits)
            const size_t run = std::min<size_t>(5552, pos + n - i);
            for (size_t k = 0; k < run; k++) { a += raw[i + k]; b += a; }
            a %= 65521u; b %= 65521u;
            i += run;
        }
        pos += n;
    }
    put_be32(z, (b << 16) | a);
    ok = ok && z.size() <= 0x7fffffffu;   // (one IDAT chunk: a film beyond ~700 megapixels is refused, not truncated)
    ok = ok && write_chunk(f, "IDAT", z);
    ok = ok && write_chunk(f, "IEND", {});
    return (std::fclose(f) == 0) && ok;
}

}  // namespace qzhost
