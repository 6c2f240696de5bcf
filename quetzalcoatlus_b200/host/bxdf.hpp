// Only the PARAMETER type of the reference's bxdf.hpp survives on the host:
// TrowbridgeReitzDistribution (bxdf.hpp:134-152, bxdf.cpp:210-215, :234), which users
// pass to ConductiveMaterial.  The BxDF/BSDF evaluation classes (bxdf.hpp:28-199) are the
// rendering path itself and exist only as CUDA device code (csrc/bxdf.cuh); a host copy
// would be a CPU fallback, which this library deliberately does not have.
#pragma once

#include <algorithm>

struct TrowbridgeReitzDistribution {
    TrowbridgeReitzDistribution(float alpha_x, float alpha_y) : m_alpha_x(alpha_x), m_alpha_y(alpha_y) {
        if (!is_smooth() && (alpha_x == 0. || alpha_y == 0.)) {
            m_alpha_x = std::max(1e-5f, alpha_x);
            m_alpha_y = std::max(1e-5f, alpha_y);
        }
    }

    bool is_smooth() const { return m_alpha_x < 1e-3f && m_alpha_y < 1e-3f; }

    float m_alpha_x;
    float m_alpha_y;
};
