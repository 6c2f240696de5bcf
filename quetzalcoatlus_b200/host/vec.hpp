// Host-side vector types with the reference's public names and semantics
// (vec.hpp:6-127): Vec3 / Pt3 / Vec2, plain IEEE float32 component arithmetic, no FMA
// (the host library is built with -ffp-contract=off), normalized() divides by sqrt.
// These run on the CPU only while a scene is being BUILT; nothing here is on the
// rendering path.
#pragma once

#include <array>
#include <cmath>
#include <string>

class Vec3 {
public:
    float x, y, z;

    Vec3() : x(0), y(0), z(0) {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}

    float r() const { return x; }
    float g() const { return y; }
    float b() const { return z; }

    std::string str() const { return std::to_string(x) + " " + std::to_string(y) + " " + std::to_string(z); }
    bool is_zero() const { return x == 0 && y == 0 && z == 0; }

    const Vec3& operator+() const { return *this; }
    Vec3 operator-() const { return {-x, -y, -z}; }
    bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }

    Vec3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Vec3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Vec3 operator*(const Vec3& o) const { return {x * o.x, y * o.y, z * o.z}; }
    Vec3 operator/(const Vec3& o) const { return {x / o.x, y / o.y, z / o.z}; }
    Vec3 operator*(float t) const { return {x * t, y * t, z * t}; }
    Vec3 operator/(float t) const { return {x / t, y / t, z / t}; }
    Vec3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Vec3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    Vec3& operator*=(const Vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    Vec3& operator/=(const Vec3& o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
    Vec3& operator*=(float t) { x *= t; y *= t; z *= t; return *this; }
    Vec3& operator/=(float t) { x /= t; y /= t; z /= t; return *this; }

    float norm_squared() const { return x * x + y * y + z * z; }
    float norm() const { return std::sqrt(x * x + y * y + z * z); }
    Vec3 normalized() const { float n = norm(); return {x / n, y / n, z / n}; }
    Vec3& normalize() { float n = norm(); x /= n; y /= n; z /= n; return *this; }

    float dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
    Vec3 cross(const Vec3& o) const { return {y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x}; }

    std::array<float, 4> to_homog() const { return {x, y, z, 0}; }
    static Vec3 from_homog(const std::array<float, 4>& v) { return {v[0], v[1], v[2]}; }

    template <typename F>
    Vec3 map(F f) const { return Vec3(f(x), f(y), f(z)); }
};

inline Vec3 operator*(float t, const Vec3& v) { return v * t; }

class Pt3 : public Vec3 {
public:
    Pt3() : Vec3() {}
    Pt3(float x_, float y_, float z_) : Vec3(x_, y_, z_) {}
    explicit Pt3(Vec3&& v) : Vec3(std::move(v)) {}

    Pt3 operator+(const Vec3& o) const { return {x + o.x, y + o.y, z + o.z}; }
    Pt3 operator-(const Vec3& o) const { return {x - o.x, y - o.y, z - o.z}; }
    Pt3& operator+=(const Vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    Pt3& operator-=(const Vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }

    std::array<float, 4> to_homog() const { return {x, y, z, 1}; }
    static Pt3 from_homog(const std::array<float, 4>& v) { return {v[0] / v[3], v[1] / v[3], v[2] / v[3]}; }
};

class Vec2 {
public:
    float x, y;

    Vec2() : x(0.0f), y(0.0f) {}
    Vec2(float x_, float y_) : x(x_), y(y_) {}

    Vec2 operator-() const { return {-x, -y}; }
    Vec2 operator+(const Vec2& o) const { return {x + o.x, y + o.y}; }
    Vec2 operator-(const Vec2& o) const { return {x - o.x, y - o.y}; }
    Vec2 operator*(float t) const { return {x * t, y * t}; }
    Vec2 operator/(float t) const { return {x / t, y / t}; }
    Vec2& operator+=(const Vec2& o) { x += o.x; y += o.y; return *this; }
    Vec2& operator-=(const Vec2& o) { x -= o.x; y -= o.y; return *this; }
    Vec2& operator*=(float t) { x *= t; y *= t; return *this; }
    Vec2& operator/=(float t) { x /= t; y /= t; return *this; }

    float norm_squared() const { return x * x + y * y; }
    float norm() const { return std::sqrt(x * x + y * y); }
};

inline Vec2 operator*(float t, const Vec2& v) { return v * t; }
