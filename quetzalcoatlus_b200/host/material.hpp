// Materials (reference material.hpp:16-101, material.cpp).  Same constructors, presets and
// public members; Material::bsdf() -- the per-hit BSDF construction -- runs on the GPU
// (csrc/shading.cuh), so the host classes carry parameters and a flatten() hook instead.
#pragma once

#include <algorithm>
#include <array>
#include <cassert>
#include <memory>
#include <numeric>

#include "bxdf.hpp"
#include "color/color.hpp"
#include "flatten.hpp"
#include "texture.hpp"
#include "vec.hpp"

class Material {
public:
    virtual ~Material() {}
    // B200 addition: append to the device material table (de-duplicated by address)
    virtual int32_t flatten(qzhost::Flattener& f) const = 0;
};

class DiffuseMaterial : public Material {
public:
    explicit DiffuseMaterial(std::unique_ptr<Texture>&& texture) : m_texture(std::move(texture)) {}

    template <typename T>
    explicit DiffuseMaterial(T&& texture) : m_texture(std::make_unique<T>(std::forward<T>(texture))) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    std::unique_ptr<Texture> m_texture;
};

class ConductiveMaterial : public Material {
public:
    ConductiveMaterial(float ior, float absorption)
        : m_ior(std::make_shared<ConstantSpectrum>(ior)), m_absorption(std::make_shared<ConstantSpectrum>(absorption)), m_roughness(0.0f, 0.0f) {}
    ConductiveMaterial(std::shared_ptr<const Spectrum> ior, std::shared_ptr<const Spectrum> absorption,
                       TrowbridgeReitzDistribution roughness = TrowbridgeReitzDistribution(0.0f, 0.0f))
        : m_ior(ior), m_absorption(absorption), m_roughness(roughness) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    // measured aluminium / copper optical constants with an anisotropic GGX roughness (material.cpp:22-36)
    static ConductiveMaterial alluminum(float roughness_a = 0.0f, float roughness_b = 0.0f) {
        return ConductiveMaterial(spectra::AL_IOR(), spectra::AL_ABSORPTION(), TrowbridgeReitzDistribution(roughness_a, roughness_b));
    }
    static ConductiveMaterial copper(float roughness_a = 0.0f, float roughness_b = 0.0f) {
        return ConductiveMaterial(spectra::CU_IOR(), spectra::CU_ABSORPTION(), TrowbridgeReitzDistribution(roughness_a, roughness_b));
    }

    std::shared_ptr<const Spectrum> m_ior;
    std::shared_ptr<const Spectrum> m_absorption;
    TrowbridgeReitzDistribution m_roughness;
};

class DielectricMaterial : public Material {
public:
    explicit DielectricMaterial(float ior) : is_constant(true), m_ior(std::make_shared<ConstantSpectrum>(ior)) {}
    // an IOR given as a Spectrum is treated as dispersive: secondary wavelengths terminate (material.cpp:38-48)
    explicit DielectricMaterial(std::shared_ptr<const Spectrum> ior) : is_constant(false), m_ior(ior) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    bool is_constant;
    std::shared_ptr<const Spectrum> m_ior;
};

class ThinDielectricMaterial : public Material {
public:
    explicit ThinDielectricMaterial(float ior) : is_constant(true), m_ior(std::make_shared<ConstantSpectrum>(ior)) {}
    explicit ThinDielectricMaterial(std::shared_ptr<const Spectrum> ior) : is_constant(false), m_ior(ior) {}

    int32_t flatten(qzhost::Flattener& f) const override;

    bool is_constant;
    std::shared_ptr<const Spectrum> m_ior;
};

namespace qzhost {
int32_t flatten_mixed(Flattener& f, const void* key, const Material* const* parts, size_t n);
}

// Picks child floor(sample * N) -- uniformly; the normalised weights are stored but, as in
// the reference (material.hpp:94-97), never used.
template <size_t N>
class MixedMaterial : public Material {
public:
    explicit MixedMaterial(std::array<std::unique_ptr<Material>, N>&& materials, std::array<float, N>&& weights)
        : m_materials(std::move(materials)), m_weights(std::move(weights)) {
        float weight_sum = std::accumulate(m_weights.begin(), m_weights.end(), 0.0f);
        assert(weight_sum > 0.0f);
        for (float& w : m_weights) w = w / weight_sum;
    }

    int32_t flatten(qzhost::Flattener& f) const override {
        std::array<const Material*, N> parts;
        for (size_t i = 0; i < N; i++) parts[i] = m_materials[i].get();
        return qzhost::flatten_mixed(f, this, parts.data(), N);
    }

    std::array<std::unique_ptr<Material>, N> m_materials;
    std::array<float, N> m_weights;
};
