// Orthonormal basis from one vector (reference onb.hpp:7-43; branchless Frisvad/Duff
// construction with sign = n.z > 0 ? 1 : -1).  Host use: Scene::add_plane (scene.cpp:243-253).
// The device twin lives in csrc/shading.cuh.
#pragma once

#include "vec.hpp"

class OrthonormalBasis {
public:
    Vec3 u[3];

    explicit OrthonormalBasis(Vec3 n) {
        n = n.normalized();
        const float s = n.z > 0.0f ? 1.0f : -1.0f;
        const float a = -1.0f / (s + n.z);
        const float b = n.x * n.y * a;
        u[0] = Vec3(1.0f + s * n.x * n.x * a, s * b, -s * n.x);
        u[1] = Vec3(b, s + n.y * n.y * a, -n.y);
        u[2] = n;
    }

    Vec3 from_local(const Vec3& v) const { return u[0] * v.x + u[1] * v.y + u[2] * v.z; }
    Vec3 to_local(const Vec3& v) const { return Vec3(u[0].dot(v), u[1].dot(v), u[2].dot(v)); }
};
