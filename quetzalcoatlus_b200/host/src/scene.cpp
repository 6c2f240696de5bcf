// Scene building, flattening and the render() entry of the host library.
//
// Reference behaviour followed (file:line into /root/reference/src):
//   scene.cpp:145-188 add_triangle; :190-241 add_quad; :243-253 add_plane (quad from an
//   OrthonormalBasis); :255-287 add_sphere; :289-373 add_obj (vertices transformed,
//   1-based -> 0-based indices, triangles as degenerate quads, normals NOT transformed);
//   :375-430 add_grid (vertex = pixel RGB as xyz, <= 65535 per side); :432-459 add_light
//   (an area light also inserts its quad/sphere with a null material); :461-464
//   set_bg_light; :22-25 commit; render.cpp:321-397 render().
#include "../scene.hpp"

#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <exception>
#include <iomanip>
#include <thread>

#include "../obj/obj.hpp"
#include "../onb.hpp"
#include "../png.hpp"
#include "../render.hpp"

namespace {

inline float bits_to_float(uint32_t b) {
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

// f(begin, end) over [0, n) cut into one range per host thread (small n: the calling thread alone)
template <class F>
void parallel_ranges(size_t n, F f) {
    const size_t parts = std::max<size_t>(1, std::min<size_t>({std::thread::hardware_concurrency(), 16, n / 65536}));
    std::vector<std::thread> workers;
    for (size_t k = 1; k < parts; k++) workers.emplace_back([&f, n, parts, k] { f(n * k / parts, n * (k + 1) / parts); });
    f(0, n / parts);
    for (auto& w : workers) w.join();
}

qz_prim make_prim(uint32_t kind, uint32_t geom_id, uint32_t prim_id, uint32_t cell, const Pt3* v, int nv) {
    qz_prim p{};
    for (int i = 0; i < nv; i++) {
        p.v[i][0] = v[i].x;
        p.v[i][1] = v[i].y;
        p.v[i][2] = v[i].z;
    }
    p.v[0][3] = bits_to_float(geom_id);
    p.v[1][3] = bits_to_float(prim_id);
    p.v[2][3] = bits_to_float(kind);
    p.v[3][3] = bits_to_float(cell);
    return p;
}

}  // namespace

// ---------------------------------------------------------------- device handle
QzDevice initialize_device() {
    QzDevice d;
    const char* env = std::getenv("QZ_DEVICE");
    if (!env) env = std::getenv("LOCAL_RANK");
    d.ordinal = env ? std::atoi(env) : 0;
    int rc = qz_init(d.ordinal);
    if (rc != QZ_OK) {
        std::cerr << "error code " << rc << ": cannot create device: " << qz_last_error() << std::endl;
        return d;
    }
    d.ok = true;
    return d;
}

// ---------------------------------------------------------------- Scene
Scene::Scene(QzDevice&& device) : m_device(device) {
    if (m_device.ok && qz_scene_create(&m_handle) != QZ_OK) {
        std::cerr << "error: " << qz_last_error() << std::endl;
        m_handle = nullptr;
    }
}

Scene::~Scene() {
    if (m_handle) qz_scene_destroy(m_handle);
}

size_t Scene::n_primitives() const {
    size_t n = 0;
    for (const auto& g : m_geom_data) n += g.prims.size();
    return n;
}

GeometryData* Scene::new_geometry(ShapeType shape, const Material* material) {
    m_geom_data.emplace_back();
    GeometryData* g = &m_geom_data.back();
    g->shape = shape;
    g->material = material;
    g->light = nullptr;
    return g;
}

GeometryData* Scene::add_triangle(const Pt3& a, const Pt3& b, const Pt3& c, const Material* material) {
    GeometryData* g = new_geometry(ShapeType::TRIANGLE, material);
    const Pt3 v[3] = {a, b, c};
    g->prims.push_back(make_prim(QZ_PRIM_TRIANGLE, uint32_t(m_geom_data.size() - 1), 0, 0, v, 3));
    return g;
}

GeometryData* Scene::add_quad(const Pt3& a, const Pt3& b, const Pt3& c, const Pt3& d, const Material* material) {
    GeometryData* g = new_geometry(ShapeType::QUAD, material);
    const Pt3 v[4] = {a, b, c, d};
    g->prims.push_back(make_prim(QZ_PRIM_QUAD, uint32_t(m_geom_data.size() - 1), 0, 0, v, 4));
    return g;
}

GeometryData* Scene::add_plane(const Pt3& p, const Vec3& n, const Material* material, float half_size) {
    OrthonormalBasis basis(n);
    Pt3 a = p - basis.u[0] * half_size - basis.u[1] * half_size;
    Pt3 b = p + basis.u[0] * half_size - basis.u[1] * half_size;
    Pt3 c = p + basis.u[0] * half_size + basis.u[1] * half_size;
    Pt3 d = p - basis.u[0] * half_size + basis.u[1] * half_size;
    return add_quad(a, b, c, d, material);
}

GeometryData* Scene::add_sphere(const Pt3& center, float radius, const Material* material) {
    GeometryData* g = new_geometry(ShapeType::SPHERE, material);
    const Pt3 v[1] = {center};
    qz_prim p = make_prim(QZ_PRIM_SPHERE, uint32_t(m_geom_data.size() - 1), 0, 0, v, 1);
    p.v[1][0] = radius;
    g->prims.push_back(p);
    return g;
}

namespace qzhost {
static thread_local BuildTimes g_build_times{};
const BuildTimes& build_times() { return g_build_times; }
void reset_build_times() { g_build_times = BuildTimes(); }
}  // namespace qzhost

GeometryData* Scene::add_obj(const std::string& filename, const Material* material, const Transform& transform) {
    const auto parse_t0 = std::chrono::steady_clock::now();
    auto obj_data = obj::load_obj(filename);
    qzhost::g_build_times.obj_parse_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - parse_t0).count();
    if (!obj_data) {
        std::cerr << "Failed to load " << filename << std::endl;
        return nullptr;
    }
    const obj::ObjData& mesh = *obj_data;
    if (mesh.vertices.empty() || mesh.faces.empty()) return nullptr;

    // every loop below is independent per element: one range per host thread (a 1M-face mesh is 64 MB of primitives)
    std::vector<Pt3> verts(mesh.vertices.size());
    parallel_ranges(verts.size(), [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++) verts[i] = transform * Pt3(mesh.vertices[i].x, mesh.vertices[i].y, mesh.vertices[i].z);
    });

    std::atomic<bool> in_range{true};
    parallel_ranges(mesh.faces.size(), [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++)
            for (int k = 0; k < 4; k++)
                if (mesh.faces[i].vertices[k] < 1 || size_t(mesh.faces[i].vertices[k]) > verts.size()) in_range = false;
    });
    if (!in_range) {
        std::cerr << "Failed to create buffers for " << filename << " (face index out of range)" << std::endl;
        return nullptr;
    }

    GeometryData* g = new_geometry(ShapeType::OBJ, material);
    const uint32_t geom_id = uint32_t(m_geom_data.size() - 1);
    g->prims.resize(mesh.faces.size());
    parallel_ranges(mesh.faces.size(), [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++) {
            const auto& idx = mesh.faces[i].vertices;
            const Pt3 v[4] = {verts[idx[0] - 1], verts[idx[1] - 1], verts[idx[2] - 1], verts[idx[3] - 1]};
            g->prims[i] = make_prim(QZ_PRIM_QUAD, geom_id, uint32_t(i), 0, v, 4);
        }
    });
    if (!mesh.vertex_normals.empty()) {
        auto nd = std::make_unique<NormalData>();
        nd->normals.resize(mesh.vertex_normals.size());
        parallel_ranges(nd->normals.size(), [&](size_t begin, size_t end) {
            for (size_t i = begin; i < end; i++) nd->normals[i] = Vec3(mesh.vertex_normals[i].x, mesh.vertex_normals[i].y, mesh.vertex_normals[i].z);
        });
        nd->faces.resize(mesh.faces.size());
        parallel_ranges(nd->faces.size(), [&](size_t begin, size_t end) {
            for (size_t i = begin; i < end; i++) {
                const auto& ns = mesh.faces[i].normals;
                nd->faces[i] = {ns[0] - 1, ns[1] - 1, ns[2] - 1, ns[3] - 1};
            }
        });
        g->normals = std::move(nd);
    }
    return g;
}

GeometryData* Scene::add_grid(const Image& image, const Material* material, const Transform& transform) {
    if (image.width > std::numeric_limits<unsigned short>::max() || image.height > std::numeric_limits<unsigned short>::max()) {
        std::cerr << "Grid too large" << std::endl;
        return nullptr;
    }
    const size_t W = image.width, H = image.height;
    if (W < 2 || H < 2) {
        // a grid without a single cell: the reference still attaches it (scene.cpp:376-430 has no lower limit), so it takes
        // a geometry ID and hands back its record; there is nothing to intersect
        GeometryData* g = new_geometry(ShapeType::GRID, material);
        g->grid_dims = 0;
        return g;
    }
    std::vector<Pt3> verts(W * H);
    parallel_ranges(W * H, [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++)
            verts[i] = transform * Pt3(image.color_buffer[i * 3 + 0], image.color_buffer[i * 3 + 1], image.color_buffer[i * 3 + 2]);
    });

    GeometryData* g = new_geometry(ShapeType::GRID, material);
    const uint32_t geom_id = uint32_t(m_geom_data.size() - 1);
    g->grid_dims = uint32_t(W - 1) | (uint32_t(H - 1) << 16);
    g->prims.resize((W - 1) * (H - 1));
    // one cell = one quad (p[y][x], p[y][x+1], p[y+1][x+1], p[y+1][x]); primID 0 (a single RTCGrid); cells in row-major order
    parallel_ranges(g->prims.size(), [&](size_t begin, size_t end) {
        for (size_t i = begin; i < end; i++) {
            const size_t y = i / (W - 1), x = i - y * (W - 1);
            const Pt3 v[4] = {verts[y * W + x], verts[y * W + x + 1], verts[(y + 1) * W + x + 1], verts[(y + 1) * W + x]};
            g->prims[i] = make_prim(QZ_PRIM_GRIDCELL, geom_id, 0, uint32_t(x) | (uint32_t(y) << 16), v, 4);
        }
    });
    return g;
}

void Scene::add_light(std::unique_ptr<Light>&& light) {
    if (light->type() == LightType::AREA) {
        auto area_light = static_cast<const AreaLight*>(light.get());
        const Shape* shape = area_light->shape();
        auto shape_type = shape->type();
        if (shape_type == ShapeType::SPHERE) {
            const Sphere* sphere = static_cast<const Sphere*>(shape);
            if (auto g = add_sphere(sphere->m_center, sphere->m_radius, nullptr)) g->light = area_light;
        } else if (shape_type == ShapeType::QUAD) {
            const Quad* quad = static_cast<const Quad*>(shape);
            auto [a, b, c, d] = quad->get_vertices();
            if (auto g = add_quad(a, b, c, d, nullptr)) g->light = area_light;
        } else {
            std::cerr << "Shape type not yet supported as an area light: " << shape->type() << std::endl;
            return;
        }
    }
    m_lights.push_back(std::move(light));
}

void Scene::set_bg_light(std::shared_ptr<const Spectrum> spectrum, float scale) {
    m_bg_light.spectrum = spectrum;
    m_bg_light.scale = scale;
}

void Scene::commit() {
    if (!m_handle) {
        std::cerr << "error: scene has no CUDA device (initialize_device failed)" << std::endl;
        return;
    }
    const auto commit_t0 = std::chrono::steady_clock::now();
    qzhost::Flattener f;
    std::unordered_map<const Light*, int32_t> light_ids;
    for (const auto& l : m_lights) light_ids[l.get()] = l->flatten(f);

    for (const auto& g : m_geom_data) {
        qz_geometry q{};
        switch (g.shape) {
            case ShapeType::SPHERE: q.shape = QZ_SHAPE_SPHERE; break;
            case ShapeType::TRIANGLE: q.shape = QZ_SHAPE_TRIANGLE; break;
            case ShapeType::QUAD: q.shape = QZ_SHAPE_QUAD; break;
            case ShapeType::OBJ: q.shape = QZ_SHAPE_OBJ; break;
            case ShapeType::GRID: q.shape = QZ_SHAPE_GRID; break;
        }
        q.material = g.material ? g.material->flatten(f) : -1;
        q.light = -1;
        if (g.light) {
            auto it = light_ids.find(g.light);
            if (it != light_ids.end()) q.light = it->second;
        }
        q.normal_offset = -1;
        q.nindex_offset = -1;
        if (g.normals && g.shape == ShapeType::OBJ) {
            q.normal_offset = int32_t(f.normals.size() / 3);
            f.normals.reserve(f.normals.size() + 3 * g.normals->normals.size());
            for (const auto& n : g.normals->normals) {
                f.normals.push_back(n.x); f.normals.push_back(n.y); f.normals.push_back(n.z);
            }
            q.nindex_offset = int32_t(f.normal_indices.size() / 4);
            const auto& nf = g.normals->faces;   // (std::array<int, 4>: contiguous ints)
            if (!nf.empty()) f.normal_indices.insert(f.normal_indices.end(), nf.front().data(), nf.front().data() + 4 * nf.size());
        }
        q.first_prim = uint32_t(f.prims.size());
        q.prim_count = uint32_t(g.prims.size());
        f.prims.insert(f.prims.end(), g.prims.begin(), g.prims.end());
        f.geometries.push_back(q);
        f.grid_dims.push_back(g.grid_dims);
    }

    qz_scene_tables t{};
    t.spectra = f.spectra.data(); t.n_spectra = uint32_t(f.spectra.size());
    t.textures = f.textures.data(); t.n_textures = uint32_t(f.textures.size());
    t.materials = f.materials.data(); t.n_materials = uint32_t(f.materials.size());
    t.mixed_children = f.mixed_children.data(); t.n_mixed_children = uint32_t(f.mixed_children.size());
    t.lights = f.lights.data(); t.n_lights = uint32_t(f.lights.size());
    t.bg_spectrum = m_bg_light.spectrum ? m_bg_light.spectrum->flatten(f) : -1;
    t.bg_scale = m_bg_light.scale;
    // (flatten of the background may have grown the spectrum table / pool: take pointers last)
    t.spectra = f.spectra.data(); t.n_spectra = uint32_t(f.spectra.size());
    t.geometries = f.geometries.data(); t.n_geometries = uint32_t(f.geometries.size());
    t.prims = f.prims.data(); t.n_prims = uint32_t(f.prims.size());
    t.pool = f.pool.data(); t.n_pool = uint32_t(f.pool.size());
    t.normals = f.normals.data(); t.n_normals = uint32_t(f.normals.size() / 3);
    t.normal_indices = f.normal_indices.data(); t.n_normal_indices = uint32_t(f.normal_indices.size() / 4);
    t.grid_dims = f.grid_dims.data();
    auto table = RGBToSpectrumTable::sRGB();
    t.rgb2spec_z = table->m_z_nodes.data();
    t.rgb2spec_coeffs = table->m_coeffs.data();

    if (qz_scene_commit(m_handle, &t) != QZ_OK) {
        std::cerr << "error: " << qz_last_error() << std::endl;
        return;
    }
    qzhost::g_build_times.commit_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - commit_t0).count();
    float bvh_ms = 0.0f;
    if (qz_scene_build_ms(m_handle, &bvh_ms) == QZ_OK) qzhost::g_build_times.bvh_build_ms = bvh_ms;
    m_ready = true;
}

// ---------------------------------------------------------------- flatten hooks
int32_t SolidColor::flatten(qzhost::Flattener& f) const {
    qz_texture t{};
    t.kind = QZ_TEX_SOLID;
    t.a = m_spectrum->flatten(f);
    f.textures.push_back(t);
    return int32_t(f.textures.size()) - 1;
}

int32_t DummyTexture::flatten(qzhost::Flattener& f) const {
    qz_texture t{};
    t.kind = QZ_TEX_DUMMY;
    t.a = white.flatten(f);
    t.b = black.flatten(f);
    f.textures.push_back(t);
    return int32_t(f.textures.size()) - 1;
}

int32_t ImageTexture::flatten(qzhost::Flattener& f) const {
    qz_texture t{};
    t.kind = QZ_TEX_IMAGE;
    t.offset = f.add_pool(image.color_buffer.data(), image.color_buffer.size());
    t.width = uint32_t(image.width);
    t.height = uint32_t(image.height);
    f.textures.push_back(t);
    return int32_t(f.textures.size()) - 1;
}

static int32_t add_material(qzhost::Flattener& f, const void* key, const qz_material& m) {
    f.materials.push_back(m);
    int32_t id = int32_t(f.materials.size()) - 1;
    f.seen_materials[key] = id;
    return id;
}
static int32_t find_material(const qzhost::Flattener& f, const void* key) {
    auto it = f.seen_materials.find(key);
    return it == f.seen_materials.end() ? -1 : it->second;
}

int32_t DiffuseMaterial::flatten(qzhost::Flattener& f) const {
    int32_t id = find_material(f, this);
    if (id >= 0) return id;
    qz_material m{};
    m.kind = QZ_MAT_DIFFUSE;
    m.a = m_texture->flatten(f);
    return add_material(f, this, m);
}

int32_t ConductiveMaterial::flatten(qzhost::Flattener& f) const {
    int32_t id = find_material(f, this);
    if (id >= 0) return id;
    qz_material m{};
    m.kind = QZ_MAT_CONDUCTOR;
    m.a = m_ior->flatten(f);
    m.b = m_absorption->flatten(f);
    m.alpha_x = m_roughness.m_alpha_x;
    m.alpha_y = m_roughness.m_alpha_y;
    return add_material(f, this, m);
}

int32_t DielectricMaterial::flatten(qzhost::Flattener& f) const {
    int32_t id = find_material(f, this);
    if (id >= 0) return id;
    qz_material m{};
    m.kind = QZ_MAT_DIELECTRIC;
    m.a = m_ior->flatten(f);
    m.is_constant = is_constant ? 1u : 0u;
    return add_material(f, this, m);
}

int32_t ThinDielectricMaterial::flatten(qzhost::Flattener& f) const {
    int32_t id = find_material(f, this);
    if (id >= 0) return id;
    qz_material m{};
    m.kind = QZ_MAT_THIN_DIELECTRIC;
    m.a = m_ior->flatten(f);
    m.is_constant = is_constant ? 1u : 0u;
    return add_material(f, this, m);
}

int32_t qzhost::flatten_mixed(Flattener& f, const void* key, const Material* const* parts, size_t n) {
    int32_t id = find_material(f, key);
    if (id >= 0) return id;
    std::vector<int32_t> children(n);
    for (size_t i = 0; i < n; i++) children[i] = parts[i]->flatten(f);
    qz_material m{};
    m.kind = QZ_MAT_MIXED;
    m.a = int32_t(f.mixed_children.size());
    m.count = uint32_t(n);
    f.mixed_children.insert(f.mixed_children.end(), children.begin(), children.end());
    return add_material(f, key, m);
}

int32_t PointLight::flatten(qzhost::Flattener& f) const {
    qz_light l{};
    l.kind = QZ_LIGHT_POINT;
    l.spectrum = m_spectrum->flatten(f);
    l.scale = m_scale;
    l.p[0] = m_point.x; l.p[1] = m_point.y; l.p[2] = m_point.z;
    f.lights.push_back(l);
    return int32_t(f.lights.size()) - 1;
}

int32_t AreaLight::flatten(qzhost::Flattener& f) const {
    qz_light l{};
    l.spectrum = m_spectrum->flatten(f);
    l.scale = m_scale;
    l.two_sided = m_two_sided ? 1u : 0u;
    // ShapeSample::pdf and AreaLight::pdf are both 1.0f / area() (shape.hpp:47,52-55,82,89-91)
    l.inv_area = 1.0f / m_shape->area();
    if (m_shape->type() == ShapeType::QUAD) {
        const Quad* q = static_cast<const Quad*>(m_shape.get());
        l.kind = QZ_LIGHT_AREA_QUAD;
        l.p[0] = q->p00().x; l.p[1] = q->p00().y; l.p[2] = q->p00().z;
        l.du[0] = q->du().x; l.du[1] = q->du().y; l.du[2] = q->du().z;
        l.dv[0] = q->dv().x; l.dv[1] = q->dv().y; l.dv[2] = q->dv().z;
        l.normal[0] = q->normal().x; l.normal[1] = q->normal().y; l.normal[2] = q->normal().z;
    } else {
        const Sphere* s = static_cast<const Sphere*>(m_shape.get());
        l.kind = QZ_LIGHT_AREA_SPHERE;
        l.p[0] = s->m_center.x; l.p[1] = s->m_center.y; l.p[2] = s->m_center.z;
        l.radius = s->m_radius;
    }
    f.lights.push_back(l);
    return int32_t(f.lights.size()) - 1;
}

// ---------------------------------------------------------------- render()
namespace qzhost {

static thread_local qz_stats g_last_stats{};
const qz_stats& last_render_stats() { return g_last_stats; }

qz_camera flatten_camera(const Camera& camera, std::vector<float>& sensor_storage) {
    qz_camera c{};
    c.image_width = uint32_t(camera.image_width);
    c.image_height = uint32_t(camera.image_height);
    const Vec3* src[4] = {&camera.pos, &camera.viewport_bottom_left, &camera.pixel_delta_u, &camera.pixel_delta_v};
    float* dst[4] = {c.pos, c.viewport_bottom_left, c.pixel_delta_u, c.pixel_delta_v};
    for (int i = 0; i < 4; i++) { dst[i][0] = src[i]->x; dst[i][1] = src[i]->y; dst[i][2] = src[i]->z; }
    // the sensor's curves are DenselySampledSpectrum copies over 360..830 nm (sensor.hpp:31-35)
    const DenselySampledSpectrum* curves[3] = {&camera.sensor.curve_r(), &camera.sensor.curve_g(), &camera.sensor.curve_b()};
    const size_t n = size_t(LAMBDA_MAX - LAMBDA_MIN + 1);
    sensor_storage.assign(3 * n, 0.0f);
    for (int k = 0; k < 3; k++)
        for (size_t i = 0; i < n; i++) sensor_storage[k * n + i] = (*curves[k])(float(LAMBDA_MIN + int(i)));
    c.sensor_rgb = sensor_storage.data();
    c.imaging_ratio = camera.sensor.imaging_ratio();
    return c;
}

}  // namespace qzhost

RenderResult render(const Camera& camera, const Scene& scene, size_t n_samples, size_t max_bounces) {
    const size_t H = camera.image_height, W = camera.image_width, n = H * W * 3;
    if (!scene.ready()) {
        std::cout << "Scene must be committed before rendering." << std::endl;
        return RenderResult(H, W);
    }
    const auto start_time = std::chrono::steady_clock::now();
    // the reference's closing lines, character for character (render.cpp:392-394: the end of the progress-bar line, then
    // wall-clock seconds of the render with three decimals and the unit chrono's operator<< appends) -- including what it
    // leaves behind: std::cout stays in fixed notation with precision 3 for whatever the application prints next
    auto render_time_line = [&] {
        const std::chrono::duration<float> duration = std::chrono::steady_clock::now() - start_time;
        std::cout << "[" << std::string(40, '=') << "] 100%\r";   // the final state of the reference's progress bar (render.cpp:220-236)
        std::cout << std::endl << "Render time: " << std::fixed << std::setprecision(3) << duration.count() << "s" << std::endl;
    };
    if (n_samples == 0) {
        // the reference divides the empty per-pixel sums by float(0) (render.cpp:280-282): every film value is 0.0f / 0.0f,
        // the default NaN of the host's division.  (The library itself refuses a render without samples.)
        RenderResult result(H, W);
        volatile float zero = 0.0f;
        const float nan = zero / zero;
        for (auto* plane : {&result.color_buffer, &result.normal_buffer, &result.albedo_buffer}) std::fill(plane->begin(), plane->end(), nan);
        render_time_line();
        return result;
    }
    std::vector<float> sensor;
    qz_camera cam = qzhost::flatten_camera(camera, sensor);
    qz_stats stats{};
    // The RenderResult's three planes are fresh memory every call -- tens of megabytes whose first touch (the vectors'
    // zero fill) costs milliseconds of page faults, a third of a 1920 x 1080 render.  A helper thread takes them while the
    // GPU renders; the library writes into a scratch area this thread keeps between calls (its pages stay resident), and
    // the planes are copied over once both are done.
    static thread_local std::vector<float> scratch;
    if (scratch.size() < 3 * n) scratch.resize(3 * n);
    std::unique_ptr<RenderResult> result;
    std::exception_ptr alloc_failure;
    std::thread prepare([&] {
        try { result = std::make_unique<RenderResult>(H, W); } catch (...) { alloc_failure = std::current_exception(); }
    });
    float* const film = scratch.data();   // (a local: `scratch` named inside another thread would be THAT thread's own)
    const int rc = qz_render(scene.handle(), &cam, uint32_t(n_samples), uint32_t(max_bounces), nullptr, nullptr,
                             film, film + n, film + 2 * n, &stats);
    prepare.join();
    if (alloc_failure) std::rethrow_exception(alloc_failure);
    if (rc != QZ_OK) {
        std::cerr << "error: render failed: " << qz_last_error() << std::endl;
        return std::move(*result);   // (zeros)
    }
    float* const planes[3] = {result->color_buffer.data(), result->normal_buffer.data(), result->albedo_buffer.data()};
    parallel_ranges(n, [&](size_t begin, size_t end) {
        for (int k = 0; k < 3; k++) std::memcpy(planes[k] + begin, film + size_t(k) * n + begin, (end - begin) * sizeof(float));
    });
    qzhost::g_last_stats = stats;
    render_time_line();
    return std::move(*result);
}

// ---------------------------------------------------------------- Image output
static bool has_suffix(const std::string& name, const char* suffix) {
    const size_t n = std::strlen(suffix);
    if (name.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (std::tolower((unsigned char)name[name.size() - n + i]) != suffix[i]) return false;
    return true;
}

static void write_image(const std::string& filename, const std::vector<float>& buf, size_t w, size_t h, float gamma) {
    if (has_suffix(filename, ".png")) {
        // what the reference's examples save: the tone path of image.cpp:10-15 (x255, powf gamma, B and R swapped for
        // OpenCV) runs as a kernel (qz_tone), the 8-bit values are the ones cv::imwrite stores for that CV_32FC3 matrix;
        // OpenCV reads the matrix as BGR and writes RGB, so the PNG's pixels are in the film's own channel order
        std::vector<unsigned char> px(w * h * 3);
        if (qz_tone(buf.data(), uint32_t(w * h), gamma, nullptr, px.data()) != QZ_OK) {
            std::cerr << "error: save failed: " << qz_last_error() << std::endl;
            return;
        }
        for (size_t i = 0; i < w * h; i++) std::swap(px[3 * i], px[3 * i + 2]);
        if (!qzhost::write_png_rgb8(filename, px.data(), w, h)) std::cerr << "Unable to write file " << filename << std::endl;
        return;
    }
    FILE* f = std::fopen(filename.c_str(), "wb");
    if (!f) {
        std::cerr << "Unable to open file " << filename << std::endl;
        return;
    }
    bool ppm = has_suffix(filename, ".ppm");
    if (ppm) {
        // 8-bit: the tone path of image.cpp:10-15 (x255, powf gamma, BGR) runs as a kernel (qz_tone); a PPM stores RGB,
        // so the channel swap the reference does for OpenCV is undone when the rows are written
        std::vector<unsigned char> bgr(w * h * 3);
        if (qz_tone(buf.data(), uint32_t(w * h), gamma, nullptr, bgr.data()) != QZ_OK) {
            std::cerr << "error: save failed: " << qz_last_error() << std::endl;
            std::fclose(f);
            return;
        }
        std::fprintf(f, "P6\n%zu %zu\n255\n", w, h);
        for (size_t i = 0; i < w * h; i++) std::swap(bgr[3 * i], bgr[3 * i + 2]);
        std::fwrite(bgr.data(), 1, bgr.size(), f);
    } else {
        std::fprintf(f, "PF\n%zu %zu\n-1.0\n", w, h);
        for (size_t row = h; row-- > 0;) std::fwrite(buf.data() + row * w * 3, sizeof(float), w * 3, f);
    }
    std::fclose(f);
}

void Image::save(const std::string& filename, float gamma) const { write_image(filename, color_buffer, width, height, gamma); }
void Image::denoise(bool) { std::cerr << "denoise: OpenImageDenoise is not part of this build; image left unchanged" << std::endl; }
void RenderResult::denoise(bool) { std::cerr << "denoise: OpenImageDenoise is not part of this build; image left unchanged" << std::endl; }
void RenderResult::save_normal(const std::string& filename) const { write_image(filename, normal_buffer, width, height, 1.0f); }
void RenderResult::save_albedo(const std::string& filename) const { write_image(filename, albedo_buffer, width, height, 1.0f); }
