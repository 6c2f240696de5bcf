// RGB colour, sigmoid-polynomial spectra and the sRGB colour space (reference
// color/rgb.hpp:10-105, rgb.cpp).  RGB -> spectrum uses the 3 x 32^3 x 3 coefficient
// table the reference caches as coeffs_SRGB_32.dat (rgb_to_spectrum_opt.cpp:876-898);
// this library loads that same file from its data directory (the optimiser that
// generates it is host set-up, out of scope here -- SURVEY.md section 2 row 7).
#pragma once

#include <memory>
#include <vector>

#include "spectrum.hpp"
#include "xyz.hpp"
#include "../transform.hpp"
#include "../vec.hpp"

class RGB : public Vec3 {
public:
    RGB() : Vec3(0.0f, 0.0f, 0.0f) {}
    explicit RGB(Vec3&& v) : Vec3(std::move(v)) {}
    RGB(float r_, float g_, float b_) : Vec3(r_, g_, b_) {}
};

// s(c0 + c1*lambda + c2*lambda^2), s(x) = 1/2 + x / (2 sqrt(1 + x^2))   (rgb.cpp:51-67)
class RGBSigmoidPolynomial : public Spectrum {
public:
    RGBSigmoidPolynomial() {}
    RGBSigmoidPolynomial(float c0_, float c1_, float c2_) : c0(c0_), c1(c1_), c2(c2_) {}

    float operator()(float lambda) const override;
    float max_value() const;
    int32_t flatten(qzhost::Flattener& f) const override;

    float c0, c1, c2;
};

class RGBToSpectrumTable {
public:
    static std::shared_ptr<const RGBToSpectrumTable> sRGB();

    RGBToSpectrumTable(std::vector<float>&& z_nodes, std::vector<float>&& coeffs)
        : m_z_nodes(std::move(z_nodes)), m_coeffs(std::move(coeffs)) {}

    RGBSigmoidPolynomial operator()(const RGB& rgb) const;

    std::vector<float> m_z_nodes;
    std::vector<float> m_coeffs;
};

class RGBColorSpace {
public:
    RGBColorSpace(Vec2 r, Vec2 g, Vec2 b, std::shared_ptr<const Spectrum> illuminant,
                  std::shared_ptr<const RGBToSpectrumTable> table);

    RGB rgb_from_xyz(const XYZ& xyz) const { return RGB(m_rgb_from_xyz * xyz); }
    XYZ rgb_to_xyz(const RGB& rgb) const { return XYZ(m_xyz_from_rgb * rgb); }
    RGB rgb_from_sample(const SpectrumSample& ss, const WavelengthSample& wl) const {
        return rgb_from_xyz(XYZ::from_sample(ss, wl));
    }
    RGBSigmoidPolynomial to_spectrum(const RGB& rgb) const;
    Vec2 whitepoint() const { return m_white; }

    static std::shared_ptr<const RGBColorSpace> sRGB();

    const Vec2 m_r, m_g, m_b;
    const std::shared_ptr<const Spectrum> m_illuminant;
    const std::shared_ptr<const RGBToSpectrumTable> m_table;

private:
    Vec2 m_white;
    Mat3 m_xyz_from_rgb;
    Mat3 m_rgb_from_xyz;
};

class RGBUnboundedSpectrum : public Spectrum {
public:
    RGBUnboundedSpectrum(const RGBSigmoidPolynomial& polynomial, float scale) : m_scale(scale), m_polynomial(polynomial) {}
    explicit RGBUnboundedSpectrum(RGB rgb, const RGBColorSpace& cs = *RGBColorSpace::sRGB());
    RGBUnboundedSpectrum(float r, float g, float b, const RGBColorSpace& cs = *RGBColorSpace::sRGB())
        : RGBUnboundedSpectrum(RGB(r, g, b), cs) {}

    float operator()(float lambda) const override { return m_scale * m_polynomial(lambda); }
    int32_t flatten(qzhost::Flattener& f) const override;

private:
    float m_scale;
    RGBSigmoidPolynomial m_polynomial;
};

class RGBIlluminantSpectrum : public Spectrum {
public:
    explicit RGBIlluminantSpectrum(RGB rgb, const RGBColorSpace& cs = *RGBColorSpace::sRGB());
    RGBIlluminantSpectrum(float r, float g, float b, const RGBColorSpace& cs = *RGBColorSpace::sRGB())
        : RGBIlluminantSpectrum(RGB(r, g, b), cs) {}

    float operator()(float lambda) const override {
        if (!m_illuminant) return 0.0f;
        return m_scale * m_polynomial(lambda) * (*m_illuminant)(lambda);
    }
    int32_t flatten(qzhost::Flattener& f) const override;

    const std::shared_ptr<const Spectrum> m_illuminant;

private:
    float m_scale;
    RGBSigmoidPolynomial m_polynomial;
};
