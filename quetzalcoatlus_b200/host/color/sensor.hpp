// PixelSensor (reference color/sensor.hpp:8-36, sensor.cpp): three sensor response
// curves resampled to 1 nm, an imaging ratio, and the per-sample saturation clamp at 40.
// to_sensor_rgb on the host exists for the colour known-answer tests; the render path
// evaluates it on the GPU (csrc/film.cuh) from the curves baked into qz_camera.
#pragma once

#include "rgb.hpp"
#include "spectra.hpp"
#include "spectrum.hpp"

class PixelSensor {
public:
    PixelSensor(const RGBColorSpace& cs, const Spectrum& illuminant, float imaging_ratio = 1.0f);
    PixelSensor(const Spectrum& r, const Spectrum& g, const Spectrum& b, const RGBColorSpace& cs,
                const Spectrum& illuminant, float imaging_ratio = 1.0f);

    RGB to_sensor_rgb(const SpectrumSample& sample, const WavelengthSample& wavelengths) const;

    static PixelSensor CIE_XYZ(float imaging_ratio = 1.0f / spectra::CIE_Y_INTEGRAL);
    static PixelSensor CANON_EOS(float imaging_ratio = 1.0f / spectra::CANON_EOS_R()->integral());

    // B200 additions: read access for flattening into qz_camera
    const DenselySampledSpectrum& curve_r() const { return m_r; }
    const DenselySampledSpectrum& curve_g() const { return m_g; }
    const DenselySampledSpectrum& curve_b() const { return m_b; }
    float imaging_ratio() const { return m_imaging_ratio; }

private:
    DenselySampledSpectrum m_r, m_g, m_b;
    float m_imaging_ratio;
    Mat3 m_xyz_from_sensor_rgb;
};
