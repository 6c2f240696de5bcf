// The reference's interaction.hpp (SurfaceInteraction: hit point, wo, normal, uv, material,
// light) is integrator-internal state; on the B200 it is the hit record of the wavefront
// pipeline (csrc/wf_types.cuh: WfBuffers).  This header exists so that code including it still compiles.
#pragma once

#include "light.hpp"
#include "material.hpp"
#include "vec.hpp"

struct Interaction {
    Pt3 point;
    Vec3 wo;
    Vec3 normal;
    Vec2 uv;
};
