// CIE XYZ tristimulus helper (reference color/xyz.hpp:8-21, xyz.cpp).
#pragma once

#include "spectra.hpp"
#include "spectrum.hpp"
#include "spectrum_sample.hpp"
#include "../vec.hpp"

class XYZ : public Vec3 {
public:
    explicit XYZ(Vec3&& v) : Vec3(std::move(v)) {}
    XYZ(float x_, float y_, float z_) : Vec3(x_, y_, z_) {}

    static XYZ from_spectrum(const Spectrum& s) {
        return XYZ(spectra::X()->inner_product(s) / spectra::CIE_Y_INTEGRAL,
                   spectra::Y()->inner_product(s) / spectra::CIE_Y_INTEGRAL,
                   spectra::Z()->inner_product(s) / spectra::CIE_Y_INTEGRAL);
    }

    static XYZ from_sample(const SpectrumSample& ss, const WavelengthSample& wl) {
        auto pdf = SpectrumSample::from_wavelengths_pdf(wl);
        auto tri = [&](const Spectrum& cmf) {
            return ((SpectrumSample::from_spectrum(cmf, wl) * ss) / pdf).average() / spectra::CIE_Y_INTEGRAL;
        };
        float x = tri(*spectra::X()), y = tri(*spectra::Y()), z = tri(*spectra::Z());
        return XYZ(x, y, z);
    }

    Vec2 xy() const { return Vec2(x / (x + y + z), y / (x + y + z)); }

    static XYZ from_xyY(float x, float y, float Y = 1.0f) {
        if (y == 0.0f) return XYZ(0.0f, 0.0f, 0.0f);
        return XYZ(x * Y / y, Y, (1.0f - x - y) * Y / y);
    }
};
