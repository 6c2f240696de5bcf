// WavelengthSample / SpectrumSample (reference color/spectrum_sample.hpp:11-101): the
// 4-wavelength sample types that appear in the reference's public signatures.  On the
// GPU these are a float4 each (csrc/spectra.cuh); the host versions exist for API
// completeness and for the colour known-answer tests.
#pragma once

#include <array>
#include <cstddef>

#include "spectrum.hpp"

const size_t N_SPECTRUM_SAMPLES = 4;

class WavelengthSample {
public:
    using SampleArray = std::array<float, N_SPECTRUM_SAMPLES>;

    WavelengthSample() {}
    WavelengthSample(SampleArray&& lambdas, SampleArray&& pdf) : m_lambdas(std::move(lambdas)), m_pdf(std::move(pdf)) {}

    // stratified over [lambda_min, lambda_max): first wavelength from u, the others at
    // equal spacing with wrap-around (spectrum_sample.cpp:11-24)
    static WavelengthSample uniform(float u, float lambda_min = LAMBDA_MIN, float lambda_max = LAMBDA_MAX) {
        SampleArray l;
        l[0] = (1.0f - u) * lambda_min + u * lambda_max;
        float delta = (lambda_max - lambda_min) / N_SPECTRUM_SAMPLES;
        for (size_t i = 1; i < N_SPECTRUM_SAMPLES; i++) {
            l[i] = l[i - 1] + delta;
            if (l[i] > lambda_max) l[i] = lambda_min + (l[i] - lambda_max);
        }
        SampleArray pdf;
        pdf.fill(1.0f / (lambda_max - lambda_min));
        return WavelengthSample(std::move(l), std::move(pdf));
    }

    bool secondary_terminated() const {
        for (size_t i = 1; i < N_SPECTRUM_SAMPLES; i++)
            if (m_pdf[i] != 0.0f) return false;
        return true;
    }
    void terminate_secondary() {
        if (secondary_terminated()) return;
        for (size_t i = 1; i < N_SPECTRUM_SAMPLES; i++) m_pdf[i] = 0.0f;
        m_pdf[0] /= N_SPECTRUM_SAMPLES;
    }

    float operator[](size_t i) const { return m_lambdas[i]; }
    bool operator==(const WavelengthSample& o) const { return m_lambdas == o.m_lambdas && m_pdf == o.m_pdf; }

    SampleArray m_lambdas;
    SampleArray m_pdf;
};

class SpectrumSample {
public:
    using SampleArray = std::array<float, N_SPECTRUM_SAMPLES>;

    SpectrumSample() : m_values({}) {}
    explicit SpectrumSample(const SampleArray& v) : m_values(v) {}
    explicit SpectrumSample(float c) { m_values.fill(c); }

    static SpectrumSample from_spectrum(const Spectrum& s, const WavelengthSample& wl) {
        SampleArray v;
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; i++) v[i] = s(wl.m_lambdas[i]);
        return SpectrumSample(v);
    }
    static SpectrumSample from_wavelengths_pdf(const WavelengthSample& wl) { return SpectrumSample(wl.m_pdf); }

    float operator[](size_t i) const { return m_values[i]; }
    float& operator[](size_t i) { return m_values[i]; }

    bool is_zero() const {
        for (float v : m_values) if (v != 0.0f) return false;
        return true;
    }
    float max_component() const {
        float m = m_values[0];
        for (size_t i = 1; i < N_SPECTRUM_SAMPLES; i++) if (m_values[i] > m) m = m_values[i];
        return m;
    }
    float average() const {
        float sum = 0.0f;
        for (float v : m_values) sum += v;
        return sum / N_SPECTRUM_SAMPLES;
    }

#define QZ_SS_BINOP(op)                                                                  \
    SpectrumSample operator op(const SpectrumSample& o) const {                          \
        SampleArray v;                                                                   \
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; i++) v[i] = m_values[i] op o.m_values[i]; \
        return SpectrumSample(v);                                                        \
    }                                                                                    \
    SpectrumSample operator op(float c) const {                                          \
        SampleArray v;                                                                   \
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; i++) v[i] = m_values[i] op c;         \
        return SpectrumSample(v);                                                        \
    }
    QZ_SS_BINOP(+)
    QZ_SS_BINOP(-)
    QZ_SS_BINOP(*)
#undef QZ_SS_BINOP
    // division yields 0 where the divisor is 0 (spectrum_sample.cpp:119-136, 183-201)
    SpectrumSample operator/(const SpectrumSample& o) const {
        SampleArray v;
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; i++) v[i] = o.m_values[i] == 0.0f ? 0.0f : m_values[i] / o.m_values[i];
        return SpectrumSample(v);
    }
    SpectrumSample operator/(float c) const {
        SampleArray v;
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; i++) v[i] = c == 0.0f ? 0.0f : m_values[i] / c;
        return SpectrumSample(v);
    }
    SpectrumSample& operator+=(const SpectrumSample& o) { return *this = *this + o; }
    SpectrumSample& operator-=(const SpectrumSample& o) { return *this = *this - o; }
    SpectrumSample& operator*=(const SpectrumSample& o) { return *this = *this * o; }
    SpectrumSample& operator/=(const SpectrumSample& o) { return *this = *this / o; }
    SpectrumSample& operator+=(float c) { return *this = *this + c; }
    SpectrumSample& operator-=(float c) { return *this = *this - c; }
    SpectrumSample& operator*=(float c) { return *this = *this * c; }
    SpectrumSample& operator/=(float c) { return *this = *this / c; }

    template <typename F>
    SpectrumSample map(F&& f) const {
        SampleArray v;
        for (size_t i = 0; i < N_SPECTRUM_SAMPLES; ++i) v[i] = f(m_values[i]);
        return SpectrumSample(v);
    }
    template <typename F>
    void map_inplace(F&& f) {
        for (float& v : m_values) v = f(v);
    }

    SampleArray m_values;
};
