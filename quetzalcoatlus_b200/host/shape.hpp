// Emitter shapes (reference shape.hpp:9-99): Quad (parallelogram p00, du, dv; normal along
// du x dv) and Sphere.  sample_point() is evaluated on the GPU (csrc/shading.cuh);
// the host keeps the parameters, area() and the vertex list Scene::add_light needs.
#pragma once

#include <cmath>
#include <tuple>

#include "vec.hpp"

enum ShapeType { SPHERE, TRIANGLE, QUAD, OBJ, GRID };

class Shape {
public:
    virtual ~Shape() {}
    virtual float area() const = 0;
    virtual float pdf(const Pt3&) const { return 1.0f / area(); }
    virtual ShapeType type() const = 0;
};

class Quad : public Shape {
public:
    Quad(const Pt3& p00, const Vec3& du, const Vec3& dv)
        : m_normal(du.cross(dv).normalized()), m_p00(p00), m_du(du), m_dv(dv), m_area(du.cross(dv).norm()) {}

    float area() const override { return m_area; }
    float pdf(const Pt3&) const override { return 1.0f / area(); }
    ShapeType type() const override { return QUAD; }

    std::tuple<Pt3, Pt3, Pt3, Pt3> get_vertices() const { return {m_p00, m_p00 + m_du, m_p00 + m_du + m_dv, m_p00 + m_dv}; }

    // B200 additions: read access for flattening
    const Vec3& normal() const { return m_normal; }
    const Pt3& p00() const { return m_p00; }
    const Vec3& du() const { return m_du; }
    const Vec3& dv() const { return m_dv; }

private:
    Vec3 m_normal;
    Pt3 m_p00;
    Vec3 m_du;
    Vec3 m_dv;
    float m_area;
};

class Sphere : public Shape {
public:
    Sphere(const Pt3& center, float radius) : m_center(center), m_radius(radius) {}

    float area() const override { return 4.0f * M_PI * m_radius * m_radius; }
    float pdf(const Pt3&) const override { return 1.0f / area(); }
    ShapeType type() const override { return SPHERE; }

    Pt3 m_center;
    float m_radius;
};
