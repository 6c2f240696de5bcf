// Scene (reference scene.hpp:21-108, scene.cpp): the same scene-building surface --
// add_triangle / add_sphere / add_quad / add_plane / add_obj / add_grid / add_light /
// set_bg_light / commit / ready -- but commit() flattens the object graph into POD tables
// and hands them to the CUDA library (qz_scene_commit builds the wide BVH on the GPU)
// instead of calling rtcCommitScene.  There is no Embree and no host-side intersection:
// ray_intersect / occluded / sample_lights (scene.hpp:55-62) are the rendering path and
// exist only as device code.
//
// Error convention as in the reference: failures print to std::cerr and return nullptr
// (add_*) or leave the scene not ready (commit), in which case render() prints
// "Scene must be committed before rendering." and returns a zero-filled RenderResult.
#pragma once

#include <array>
#include <deque>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "color/color.hpp"
#include "flatten.hpp"
#include "image.hpp"
#include "light.hpp"
#include "material.hpp"
#include "ray.hpp"
#include "vec.hpp"

// stands where the reference has Embree's RTCDevice: which CUDA device the scene lives on
struct QzDevice {
    int ordinal = 0;
    bool ok = false;
};

// picks the current CUDA device ($QZ_DEVICE or LOCAL_RANK, else 0); prints and returns a
// not-ok handle when no device is usable (reference: scene.cpp:12-20)
QzDevice initialize_device();

struct BackgroundLight {
    std::shared_ptr<const Spectrum> spectrum;
    float scale = 1.0f;
};

struct NormalData {
    std::vector<Vec3> normals;              // vertex normals of the mesh
    std::vector<std::array<int, 4>> faces;  // per face: indices into `normals`
};

struct GeometryData {
    ShapeType shape;
    const Material* material;
    const AreaLight* light;
    std::unique_ptr<NormalData> normals;
    // B200 additions (filled by Scene::add_*): the primitives of this geometry in Embree
    // primID order, and the grid resolution for add_grid
    std::vector<qz_prim> prims;
    uint32_t grid_dims = 0;
};

class Scene {
public:
    explicit Scene(QzDevice&& device);
    ~Scene();
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    void commit();
    bool ready() const { return m_ready; }

    // points in clockwise order around the outward face, as in the reference
    GeometryData* add_triangle(const Pt3& a, const Pt3& b, const Pt3& c, const Material* material);
    GeometryData* add_sphere(const Pt3& center, float radius, const Material* material);
    GeometryData* add_quad(const Pt3& a, const Pt3& b, const Pt3& c, const Pt3& d, const Material* material);
    GeometryData* add_plane(const Pt3& p, const Vec3& n, const Material* material, float half_size = 1000.0f);
    GeometryData* add_obj(const std::string& filename, const Material* material, const Transform& transform = Transform::identity());
    GeometryData* add_grid(const Image& image, const Material* material, const Transform& transform = Transform::identity());

    void add_light(std::unique_ptr<Light>&& light);
    void set_bg_light(std::shared_ptr<const Spectrum> spectrum, float scale = 1.0f);

    const GeometryData* get_geom_data(unsigned int geom_id) const {
        return geom_id < m_geom_data.size() ? &m_geom_data[geom_id] : nullptr;
    }
    const BackgroundLight& get_bg_light() const { return m_bg_light; }

    // B200 additions
    qz_scene handle() const { return m_handle; }      // the C-ABI scene behind this object
    QzDevice get_device() const { return m_device; }
    size_t n_lights() const { return m_lights.size(); }
    size_t n_primitives() const;

private:
    GeometryData* new_geometry(ShapeType shape, const Material* material);

    QzDevice m_device;
    qz_scene m_handle = nullptr;
    BackgroundLight m_bg_light;
    std::deque<GeometryData> m_geom_data;  // stable addresses: callers hold GeometryData*
    std::vector<std::unique_ptr<Light>> m_lights;
    bool m_ready = false;
};
