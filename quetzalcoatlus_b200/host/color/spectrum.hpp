// Spectrum hierarchy with the reference's names and evaluation semantics
// (color/spectrum.hpp:10-93, spectrum.cpp).  Host evaluation is used only while building
// a scene (sensor curves, D65 normalisation, colour-space set-up); at render time the same
// records are evaluated on the GPU (csrc/spectra.cuh) from the flattened table.
//
// Quirks of the reference that are reproduced on purpose:
//  * PiecewiseLinearSpectrum::operator() takes i = first knot >= lambda and interpolates
//    i -> i+1 with t <= 0, i.e. it extrapolates backwards from the NEXT segment
//    (spectrum.cpp:101-109);
//  * when i is the last knot it reads one element past the end of both vectors (undefined
//    behaviour in the reference; observed to read 0 in the oracle build).  Here the read is
//    DEFINED to return 0, on the host and on the device.
//  * from_interleaved appends (831, v_last) when the last VALUE (not wavelength) is < 830
//    (spectrum.cpp:88).
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "../flatten.hpp"

const int LAMBDA_MIN = 360;
const int LAMBDA_MAX = 830;

class Spectrum {
public:
    virtual ~Spectrum() {}

    virtual float operator()(float lambda) const = 0;

    // sum over integer wavelengths 360..830 (spectrum.cpp:12-26)
    float integral() const;
    float inner_product(const Spectrum& other) const;

    // B200 addition: append this spectrum to the device table, return its id
    virtual int32_t flatten(qzhost::Flattener& f) const = 0;
};

class ConstantSpectrum : public Spectrum {
public:
    explicit ConstantSpectrum(float value) : m_value(value) {}
    float operator()(float) const override { return m_value; }
    int32_t flatten(qzhost::Flattener& f) const override;

    float m_value;
};

class DenselySampledSpectrum : public Spectrum {
public:
    explicit DenselySampledSpectrum(std::vector<float>&& values, int lambda_min = LAMBDA_MIN);
    explicit DenselySampledSpectrum(const Spectrum& other, int lambda_min = LAMBDA_MIN, int lambda_max = LAMBDA_MAX);

    float operator()(float lambda) const override;
    int32_t flatten(qzhost::Flattener& f) const override;

    float lambda_min() const { return m_lambda_min; }
    float lambda_max() const { return m_lambda_max; }
    const std::vector<float>& values() const { return m_values; }

private:
    int m_lambda_min;
    int m_lambda_max;
    std::vector<float> m_values;
};

class PiecewiseLinearSpectrum : public Spectrum {
public:
    PiecewiseLinearSpectrum(std::vector<float>&& lambdas, std::vector<float>&& values);
    static PiecewiseLinearSpectrum from_interleaved(const std::vector<float>& interleaved, bool normalize = true);

    float operator()(float lambda) const override;
    int32_t flatten(qzhost::Flattener& f) const override;

private:
    std::vector<float> m_lambdas;
    std::vector<float> m_values;
};

class BlackbodySpectrum : public Spectrum {
public:
    explicit BlackbodySpectrum(float t);
    float operator()(float lambda) const override;
    int32_t flatten(qzhost::Flattener& f) const override;

private:
    float m_t;
    float m_normalization_factor;
};
