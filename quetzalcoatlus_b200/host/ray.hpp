// Ray (reference ray.hpp:6-17): origin + unnormalised direction.
#pragma once

#include "vec.hpp"

class Ray {
public:
    Pt3 o;
    Vec3 d;

    Ray() : o(0.0f, 0.0f, 0.0f), d(0.0f, 0.0f, 0.0f) {}
    Ray(const Pt3& o_, const Vec3& d_) : o(o_), d(d_) {}

    Pt3 at(float t) const { return o + d * t; }
};
