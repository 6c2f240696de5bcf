// Mat3 / Mat4 / Transform with the reference's public surface (transform.hpp:11-109,
// transform.cpp).  Scene-build time only.  Accumulation order matters for bit parity of
// the flattened scene (camera frame, transformed mesh vertices): every matrix product
// starts from 0.0f and adds terms in index order, as the reference's loops do
// (transform.cpp:5-27, 79-102); points are divided by the homogeneous w (vec.cpp:165).
#pragma once

#include <array>
#include <cmath>
#include <cstddef>
#include <optional>

#include "ray.hpp"
#include "vec.hpp"

struct Mat3 {
    std::array<float, 9> data;  // row-major

    Mat3() : data{} {}
    explicit Mat3(std::array<float, 9>&& d) : data(std::move(d)) {}

    float operator[](size_t i) const { return data[i]; }
    float& operator[](size_t i) { return data[i]; }

    Mat3 operator*(const Mat3& o) const {
        Mat3 r;
        for (size_t i = 0; i < 3; i++)
            for (size_t j = 0; j < 3; j++)
                for (size_t k = 0; k < 3; k++) r[i * 3 + j] += data[i * 3 + k] * o[k * 3 + j];
        return r;
    }

    std::array<float, 3> operator*(const std::array<float, 3>& v) const {
        std::array<float, 3> r{};
        for (size_t i = 0; i < 3; i++)
            for (size_t j = 0; j < 3; j++) r[i] += data[i * 3 + j] * v[j];
        return r;
    }

    Vec3 operator*(const Vec3& v) const {
        return Vec3(data[0] * v.x + data[1] * v.y + data[2] * v.z,
                    data[3] * v.x + data[4] * v.y + data[5] * v.z,
                    data[6] * v.x + data[7] * v.y + data[8] * v.z);
    }

    // adjugate / determinant inverse; nullopt when det == 0 (transform.cpp:38-62)
    std::optional<Mat3> invert() const {
        const auto& m = data;
        Mat3 adj({m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[4] * m[2],
                  m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
                  m[3] * m[7] - m[6] * m[4], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]});
        float det = m[0] * adj[0] + m[1] * adj[3] + m[2] * adj[6];
        if (det == 0.0f) return std::nullopt;
        float inv_det = 1.0f / det;
        for (float& v : adj.data) v = v * inv_det;
        return adj;
    }

    static Mat3 identity() { return diagonal({1, 1, 1}); }
    static Mat3 diagonal(const std::array<float, 3>& v) {
        Mat3 r;
        r[0] = v[0]; r[4] = v[1]; r[8] = v[2];
        return r;
    }
};

struct Mat4 {
    std::array<float, 16> data;  // row-major

    Mat4() : data{} {}
    explicit Mat4(std::array<float, 16>&& d) : data(std::move(d)) {}

    float operator[](size_t i) const { return data[i]; }
    float& operator[](size_t i) { return data[i]; }

    Mat4 operator*(const Mat4& o) const {
        Mat4 r;
        for (size_t i = 0; i < 4; i++)
            for (size_t j = 0; j < 4; j++)
                for (size_t k = 0; k < 4; k++) r[i * 4 + j] += data[i * 4 + k] * o[k * 4 + j];
        return r;
    }

    std::array<float, 4> operator*(const std::array<float, 4>& v) const {
        std::array<float, 4> r{};
        for (size_t i = 0; i < 4; i++)
            for (size_t j = 0; j < 4; j++) r[i] += data[i * 4 + j] * v[j];
        return r;
    }

    Vec3 operator*(const Vec3& v) const { return Vec3::from_homog((*this) * v.to_homog()); }  // w = 0
    Pt3 operator*(const Pt3& p) const { return Pt3::from_homog((*this) * p.to_homog()); }      // w = 1

    static Mat4 identity() { return diagonal({1, 1, 1, 1}); }
    static Mat4 diagonal(const std::array<float, 4>& v) {
        Mat4 r;
        r[0] = v[0]; r[5] = v[1]; r[10] = v[2]; r[15] = v[3];
        return r;
    }

    static Mat4 rotate_x(float angle) {
        Mat4 r;
        r[0] = 1; r[5] = std::cos(angle); r[6] = -std::sin(angle);
        r[9] = std::sin(angle); r[10] = std::cos(angle); r[15] = 1;
        return r;
    }
    static Mat4 rotate_y(float angle) {
        Mat4 r;
        r[0] = std::cos(angle); r[2] = std::sin(angle); r[5] = 1;
        r[8] = -std::sin(angle); r[10] = std::cos(angle); r[15] = 1;
        return r;
    }
    static Mat4 rotate_z(float angle) {
        Mat4 r;
        r[0] = std::cos(angle); r[1] = -std::sin(angle);
        r[4] = std::sin(angle); r[5] = std::cos(angle); r[10] = 1; r[15] = 1;
        return r;
    }
    // rotation about a unit axis (Rodrigues), transform.cpp:150-163
    static Mat4 rotation(const Vec3& axis, float angle) {
        float ux = axis.x, uy = axis.y, uz = axis.z;
        float c = std::cos(angle), s = std::sin(angle);
        return Mat4({c + ux * ux * (1 - c), ux * uy * (1 - c) - uz * s, ux * uz * (1 - c) + uy * s, 0,
                     uy * ux * (1 - c) + uz * s, c + uy * uy * (1 - c), uy * uz * (1 - c) - ux * s, 0,
                     uz * ux * (1 - c) - uy * s, uz * uy * (1 - c) + ux * s, c + uz * uz * (1 - c), 0,
                     0, 0, 0, 1});
    }
};

class Transform {
public:
    Mat4 m_mat;
    Mat4 m_inv_mat;

    Transform(Mat4&& matrix, Mat4&& inverse_matrix) : m_mat(matrix), m_inv_mat(inverse_matrix) {}
    virtual ~Transform() {}

    Vec3 apply(const Vec3& v) const { return m_mat * v; }
    Pt3 apply(const Pt3& p) const { return m_mat * p; }
    Ray apply(const Ray& r) const { return Ray(m_mat * r.o, m_mat * r.d); }
    Vec3 apply_inverse(const Vec3& v) const { return m_inv_mat * v; }
    Pt3 apply_inverse(const Pt3& p) const { return m_inv_mat * p; }
    Ray apply_inverse(const Ray& r) const { return Ray(m_inv_mat * r.o, m_inv_mat * r.d); }
    Transform apply(const Transform& t) const { return Transform(m_mat * t.m_mat, t.m_inv_mat * m_inv_mat); }

    template <typename T>
    T operator*(const T& t) const { return apply(t); }

    static Transform identity() { return Transform(Mat4::identity(), Mat4::identity()); }

    static Transform translation(const Vec3& v) { return translation(v.x, v.y, v.z); }
    static Transform translation(float x, float y, float z) {
        Mat4 m = Mat4::identity(), inv = Mat4::identity();
        m[3] = x; m[7] = y; m[11] = z;
        inv[3] = -x; inv[7] = -y; inv[11] = -z;
        return Transform(std::move(m), std::move(inv));
    }

    static Transform scale(const Vec3& v) { return scale(v.x, v.y, v.z); }
    static Transform scale(float x, float y, float z) {
        return Transform(Mat4::diagonal({x, y, z, 1}), Mat4::diagonal({1 / x, 1 / y, 1 / z, 1}));
    }
    static Transform scale(float c) { return scale(c, c, c); }

    static Transform rotate_x(float a) { return Transform(Mat4::rotate_x(a), Mat4::rotate_x(-a)); }
    static Transform rotate_y(float a) { return Transform(Mat4::rotate_y(a), Mat4::rotate_y(-a)); }
    static Transform rotate_z(float a) { return Transform(Mat4::rotate_z(a), Mat4::rotate_z(-a)); }
    static Transform rotation(const Vec3& axis, float a) {
        return Transform(Mat4::rotation(axis, a), Mat4::rotation(axis, -a));
    }
};
