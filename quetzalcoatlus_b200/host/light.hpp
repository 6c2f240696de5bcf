// Lights (reference light.hpp:17-89, light.cpp): PointLight and AreaLight over a Quad or
// Sphere.  Light::sample / pdf / emission run on the GPU (csrc/shading.cuh); the
// host classes keep parameters and flatten into qz_light.
#pragma once

#include <memory>

#include "color/spectrum_sample.hpp"
#include "flatten.hpp"
#include "shape.hpp"
#include "transform.hpp"
#include "vec.hpp"

enum LightType { POINT, DIRECTIONAL, AREA };

class Light {
public:
    virtual ~Light() {}
    Light(std::shared_ptr<const Spectrum> spectrum, float scale, LightType type) : m_spectrum(spectrum), m_scale(scale), m_type(type) {}

    virtual SpectrumSample total_emission(const WavelengthSample& wavelengths) const = 0;
    LightType type() const { return m_type; }

    // B200 addition
    virtual int32_t flatten(qzhost::Flattener& f) const = 0;

protected:
    std::shared_ptr<const Spectrum> m_spectrum;
    float m_scale;
    LightType m_type;
};

class PointLight : public Light {
public:
    PointLight(const Pt3& point, std::shared_ptr<const Spectrum> spectrum, float scale = 1.0f) : Light(spectrum, scale, POINT), m_point(point) {}

    SpectrumSample total_emission(const WavelengthSample& wl) const override {
        return SpectrumSample::from_spectrum(*m_spectrum, wl) * (4.0f * M_PI * m_scale);
    }
    int32_t flatten(qzhost::Flattener& f) const override;

    Pt3 m_point;
};

class AreaLight : public Light {
public:
    AreaLight(std::unique_ptr<Shape>&& shape, std::shared_ptr<const Spectrum> spectrum, float scale = 1.0f, bool two_sided = false)
        : Light(spectrum, scale, AREA), m_shape(std::move(shape)), m_two_sided(two_sided) {}

    SpectrumSample total_emission(const WavelengthSample& wl) const override {
        return SpectrumSample::from_spectrum(*m_spectrum, wl) * (M_PI * (m_two_sided ? 2.0f : 1.0f) * m_shape->area() * m_scale);
    }
    int32_t flatten(qzhost::Flattener& f) const override;

    const Shape* shape() const { return m_shape.get(); }

private:
    std::unique_ptr<Shape> m_shape;
    bool m_two_sided;
};
