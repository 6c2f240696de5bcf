// Textures (reference texture.hpp:10-71, texture.cpp).  value() is evaluated on the GPU
// (csrc/shading.cuh: texture_value); the host classes only carry parameters and flatten.
#pragma once

#include <cmath>
#include <memory>

#include "color/color.hpp"
#include "flatten.hpp"
#include "image.hpp"
#include "vec.hpp"

class Texture {
public:
    virtual ~Texture() {}
    // B200 addition: append to the device texture table, return the id
    virtual int32_t flatten(qzhost::Flattener& f) const = 0;
};

// one spectrum everywhere; RGB colours go through the sRGB table at construction (texture.cpp:3)
class SolidColor : public Texture {
public:
    explicit SolidColor(const RGB& color, const RGBColorSpace& cs = *RGBColorSpace::sRGB())
        : m_spectrum(std::make_shared<RGBSigmoidPolynomial>(cs.to_spectrum(color))) {}
    SolidColor(float r, float g, float b, const RGBColorSpace& cs = *RGBColorSpace::sRGB()) : SolidColor(RGB(r, g, b), cs) {}
    explicit SolidColor(const std::shared_ptr<const Spectrum>& spectrum) : m_spectrum(spectrum) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    std::shared_ptr<const Spectrum> m_spectrum;
};

// 10x10 uv checker of white/black (texture.cpp:15-37)
class DummyTexture : public Texture {
public:
    DummyTexture()
        : white(RGBColorSpace::sRGB()->to_spectrum(RGB(1., 1., 1.))), black(RGBColorSpace::sRGB()->to_spectrum(RGB(0., 0., 0.))) {}

    int32_t flatten(qzhost::Flattener& f) const override;

private:
    RGBSigmoidPolynomial white;
    RGBSigmoidPolynomial black;
};

// nearest-texel RGB image; every lookup converts RGB -> spectrum through the table (texture.cpp:40-61)
class ImageTexture : public Texture {
public:
    explicit ImageTexture(Image&& image_) : image(std::move(image_)) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    Image image;
};
