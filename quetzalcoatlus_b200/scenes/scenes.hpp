// Benchmark / parity scene definitions, written ONLY against the reference's public
// scene API (Scene / Camera / Material / Light / Texture / Spectrum; scene.hpp:42-108,
// camera.hpp:11-35, material.hpp, light.hpp, texture.hpp, color/*.hpp).
//
// This header is deliberately API-agnostic: it is compiled twice --
//   * against /root/reference/src (oracle/ref_harness.cpp  -> oracle/_ref/liboracle_ref.so)
//   * against quetzalcoatlus_b200/host (harness/qz_harness.cpp -> libqz_harness.so)
// -- which is the drop-in claim in executable form: the same scene-building code
// drives the reference CPU renderer and the B200 renderer.
//
// Scene parameters restate the reference's example programs (the BASELINE.json
// configs): examples/cornell_box.cpp:12-87, glass_spheres.cpp:14-49,
// textures.cpp:12-75, opposing_planes.cpp:20-61, obj_viewer.cpp:51-134,
// mandelbrot.cpp:10-57.  Literal types (float vs double) follow the examples because
// double->float conversion points decide the last bit of the scene data.
#pragma once

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "camera.hpp"
#include "image.hpp"
#include "render.hpp"
#include "scene.hpp"
#include "vec.hpp"

namespace qzscenes {

struct Bundle {
    std::unique_ptr<Scene> scene;
    std::unique_ptr<Camera> camera;
    // materials are caller-owned in the reference API (scene.hpp:67-76): keep them alive here
    // (shared_ptr made from the concrete type: the reference's Material has no virtual destructor)
    std::vector<std::shared_ptr<const Material>> materials;
    size_t n_samples = 1;
    size_t max_bounces = 1;

    template <typename M>
    const Material* keep(M&& m) {
        materials.push_back(std::make_shared<std::decay_t<M>>(std::forward<M>(m)));
        return materials.back().get();
    }
};

struct Options {
    int width = 0;        // 0 = the example's own resolution
    int height = 0;
    std::string obj_path; // mesh scene only
    std::string obj_material = "alluminum";
    std::string obj_light = "point";
};

inline void cornell_walls(Bundle& b, bool mixed_back_wall) {
    Scene& scene = *b.scene;
    scene.add_quad(Pt3(-2.f, 2.f, -7.f), Pt3(2.f, 2.f, -7.f), Pt3(2.f, 2.f, -3.f), Pt3(-2.f, 2.f, -3.f),
                   b.keep(DiffuseMaterial(SolidColor(0.8f, 0.4f, 0.1f))));
    scene.add_quad(Pt3(-2.f, -2.f, -3.f), Pt3(2.f, -2.f, -3.f), Pt3(2.f, -2.f, -7.f), Pt3(-2.f, -2.f, -7.f),
                   b.keep(DiffuseMaterial(SolidColor(0.1f, 0.6f, 0.8f))));
    scene.add_quad(Pt3(-2.f, -2.f, -3.f), Pt3(-2.f, -2.f, -7.f), Pt3(-2.f, 2.f, -7.f), Pt3(-2.f, 2.f, -3.f),
                   b.keep(DiffuseMaterial(SolidColor(0.8f, 0.0f, 0.1f))));
    scene.add_quad(Pt3(2.f, -2.f, -3.f), Pt3(2.f, 2.f, -3.f), Pt3(2.f, 2.f, -7.f), Pt3(2.f, -2.f, -7.f),
                   b.keep(DiffuseMaterial(SolidColor(0.1f, 0.1f, 0.8f))));
    const Material* back;
    if (mixed_back_wall) {
        // no shipped example uses MixedMaterial (material.hpp:80-101); this variant covers it
        std::array<std::unique_ptr<Material>, 2> parts = {
            std::make_unique<DiffuseMaterial>(SolidColor(0.1f, 0.8f, 0.1f)),
            std::make_unique<ConductiveMaterial>(ConductiveMaterial::copper(0.3, 0.15))
        };
        std::array<float, 2> weights = {3.0f, 1.0f};
        b.materials.push_back(std::make_shared<MixedMaterial<2>>(std::move(parts), std::move(weights)));
        back = b.materials.back().get();
    } else {
        back = b.keep(DiffuseMaterial(SolidColor(0.1f, 0.8f, 0.1f)));
    }
    scene.add_quad(Pt3(-2.f, -2.f, -7.f), Pt3(2.f, -2.f, -7.f), Pt3(2.f, 2.f, -7.f), Pt3(-2.f, 2.f, -7.f), back);
}

// examples/cornell_box.cpp:12-87
inline void build_cornell(Bundle& b, const Options& o, bool mixed) {
    Scene& scene = *b.scene;
    auto light_spectrum = spectra::ILLUM_D65();
    auto light_shape = std::make_unique<Quad>(Pt3(-1.f, 1.9999f, -4.f), Vec3(0.f, 0.f, -2.f), Vec3(2.f, 0.f, 0.f));
    scene.add_light(std::make_unique<AreaLight>(std::move(light_shape), light_spectrum, 12.0f, false));
    cornell_walls(b, mixed);
    scene.add_sphere(Pt3(-0.8f, -1.25f, -4.4f), 0.75f,
                     b.keep(DielectricMaterial(std::make_shared<RGBUnboundedSpectrum>(RGB(1.1f, 1.8f, 3.0f)))));
    scene.add_sphere(Pt3(0.6f, -1.0f, -5.5f), 1.0f, b.keep(ConductiveMaterial::copper(0.1, 0.06)));
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 800, o.height ? o.height : 800, M_PI / 3.0f);
    b.n_samples = 128;
    b.max_bounces = 64;
}

// examples/glass_spheres.cpp:14-49
inline void build_glass_spheres(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    auto light_shape = std::make_unique<Quad>(Pt3(-3.f, -3.f, -12.f), Vec3(6.f, 0.f, 0.f), Vec3(0.f, 6.f, 0.f));
    scene.add_light(std::make_unique<AreaLight>(std::move(light_shape), spectra::ILLUM_D65(), 8.0f, false));
    const Material* glass = b.keep(DielectricMaterial(spectra::GLASS_SF11_IOR()));
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++)
                scene.add_sphere(Pt3(-1.5 + i, -1.5 + j, -6 + k), 0.45, glass);
    scene.add_sphere(Pt3(-0.8f, -1.25f, -4.4f), 0.75f, glass);
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 800, o.height ? o.height : 800, M_PI / 3.0f);
    b.n_samples = 512;
    b.max_bounces = 64;
}

// examples/textures.cpp:12-75
inline void build_textures(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    scene.add_light(std::make_unique<PointLight>(
        Pt3(0., 5., 5.), std::make_shared<RGBIlluminantSpectrum>(RGB(8., 2., 4.)), 60.0f));
    const Material* dummy = b.keep(DiffuseMaterial(DummyTexture{}));
    scene.add_sphere(Pt3(-1.25, 0., -5.), 1., dummy);
    scene.add_grid(Image({-2, -2, -3, 0, -2.8, -3, 2, -2, -3,
                          -2, -2, -5, 0, -2.6, -5, 2, -2, -5,
                          -2, -2, -7, 0, -2, -7, 2, -2, -7}, 3, 3), dummy);
    Image image(800, 800);
    for (int y = 0; y < int(image.height); ++y) {
        for (int x = 0; x < int(image.width); ++x) {
            int index = y * int(image.width) + x;
            image.color_buffer[3 * index + 0] = powf((float(y) / image.height - 0.5f) * 2.f, 2.f);
            image.color_buffer[3 * index + 1] = powf((float(x) / image.width - 0.5f) * 2.f, 2.f);
            image.color_buffer[3 * index + 2] = 1.f;
        }
    }
    const Material* image_material = b.keep(DiffuseMaterial(ImageTexture(std::move(image))));
    scene.add_sphere(Pt3(1.25, 0., -5.), 1., image_material);
    scene.add_quad(Pt3(-2.f, -2.f, -7.f), Pt3(2.f, -2.f, -7.f), Pt3(2.f, 2.f, -7.f), Pt3(-2.f, 2.f, -7.f), image_material);
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 800, o.height ? o.height : 800, M_PI / 3.0f);
    b.n_samples = 36;
    b.max_bounces = 32;
}

// examples/opposing_planes.cpp:20-61
inline void build_opposing_planes(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    scene.add_light(std::make_unique<AreaLight>(
        std::make_unique<Quad>(Pt3(-12, -4, 4), Vec3(0, 12, 2), Vec3(12, 0, 0)),
        std::make_shared<RGBIlluminantSpectrum>(RGB(2.0, 1.0, 2.0)), 2.0f));
    scene.add_plane(Pt3(0., 0., -10.), Vec3(-0.5, 0.5, 1.).normalized(), b.keep(ConductiveMaterial::copper(0.4, 0.2)));
    scene.add_plane(Pt3(0., 0., -10.), Vec3(0.5, -0.5, 1.).normalized(), b.keep(ConductiveMaterial::alluminum(0.3, 0.6)));
    const Material* glass = b.keep(DielectricMaterial(spectra::GLASS_SF11_IOR()));
    scene.add_sphere(Pt3(1., 1., -5.), 0.8, glass);
    scene.add_sphere(Pt3(0., 0., -6.), 0.8, glass);
    scene.add_sphere(Pt3(-1., -1., -7.), 0.8, glass);
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 1920, o.height ? o.height : 1080, M_PI_4);
    b.n_samples = 256;
    b.max_bounces = 64;
}

// examples/obj_viewer.cpp:51-134 with its default flags (position 0, scale 1, rotation 0)
inline bool build_obj_viewer(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    const Material* material;
    if (o.obj_material == "diffuse") material = b.keep(DiffuseMaterial(SolidColor(0.6, 0.8, 0.8)));
    else if (o.obj_material == "copper") material = b.keep(ConductiveMaterial::copper(0.12, 0.2));
    else if (o.obj_material == "alluminum") material = b.keep(ConductiveMaterial::alluminum(0.12, 0.2));
    else if (o.obj_material == "glass") material = b.keep(DielectricMaterial(spectra::GLASS_BK7_IOR()));
    else return false;

    auto light_spectrum = spectra::ILLUM_D65();
    if (o.obj_light == "ambient") scene.set_bg_light(light_spectrum, 0.2f);
    else if (o.obj_light == "point") scene.add_light(std::make_unique<PointLight>(Pt3(4., 6., 8.), light_spectrum, 100.0f));
    else if (o.obj_light == "area")
        scene.add_light(std::make_unique<AreaLight>(std::make_unique<Sphere>(Pt3(4., 6., 8.), 1.0), light_spectrum, 12.0f));
    else return false;

    Vec3 position(0.f, 0.f, 0.f);
    Transform transform = Transform::translation(position) * Transform::rotate_x(0.f) * Transform::rotate_y(0.f) *
                          Transform::rotate_z(0.f) * Transform::scale(1.f);
    if (!scene.add_obj(o.obj_path, material, transform)) return false;

    const Material* floor = b.keep(DiffuseMaterial(SolidColor(1.0, 0.1, 0.9)));
    scene.add_plane(Pt3(0., -0.1, 0.), Vec3(0., 1., 0.), floor);
    scene.add_plane(Pt3(0., 0., -5.), Vec3(0., 0., 1.), floor);
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 800, o.height ? o.height : 600, M_PI / 3.0,
                                        Transform::translation(0., 4., 6.) * Transform::rotate_x(-M_PI / 8.0));
    b.n_samples = 24;
    b.max_bounces = 32;
    return true;
}

// examples/mandelbrot.cpp:10-57 at a reduced grid (the shipped one is 1200x1200)
inline Image mandelbrot_grid(unsigned rows, unsigned cols) {
    const int max_iter = 256;
    Image out(rows, cols);
    for (unsigned i = 0; i < rows; i++) {
        for (unsigned j = 0; j < cols; j++) {
            float x = -2.0f + 3.0f * j / cols;
            float y = -1.5f + 3.0f * i / rows;
            std::complex<float> c(x, y), z(0.0f, 0.0f);
            int it = 0;
            while (it < max_iter && std::norm(z) < 2.0f) { z = z * z + c; it++; }
            auto base = 3 * (i * cols + j);
            out.color_buffer[base + 0] = x;
            out.color_buffer[base + 1] = y;
            out.color_buffer[base + 2] = std::pow(it / (float)max_iter, 0.05);
        }
    }
    return out;
}

inline void build_mandelbrot(Bundle& b, const Options& o, unsigned grid_res) {
    Scene& scene = *b.scene;
    scene.add_light(std::make_unique<PointLight>(
        Pt3(5., 4., 2.), std::make_shared<RGBIlluminantSpectrum>(RGB(8., 2., 4.)), 60.0f));
    Image image = mandelbrot_grid(grid_res, grid_res);
    scene.add_grid(image, b.keep(ConductiveMaterial::copper(0.2, 0.12)),
                   Transform::translation(0.5, 0., -5.) * Transform::rotate_x(-0.5));
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 800, o.height ? o.height : 800, M_PI_4);
    b.n_samples = 32;
    b.max_bounces = 12;
}

// Synthetic coverage scene: every API feature the shipped examples leave out --
// add_triangle, ThinDielectricMaterial, constant-IOR dielectric, smooth conductor,
// two-sided sphere area light, a second (point) light so the light pick matters,
// constant background light (set_bg_light), SolidColor from a Spectrum.
inline void build_kitchen_sink(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    scene.set_bg_light(spectra::ILLUM_D65(), 0.05f);
    scene.add_light(std::make_unique<AreaLight>(std::make_unique<Sphere>(Pt3(0.f, 2.5f, -5.f), 0.5f),
                                                spectra::ILLUM_D65(), 6.0f, true));
    scene.add_light(std::make_unique<PointLight>(Pt3(-3.f, 1.f, -2.f),
                                                 std::make_shared<RGBIlluminantSpectrum>(RGB(1.f, 3.f, 2.f)), 10.0f));
    scene.add_light(std::make_unique<AreaLight>(
        std::make_unique<Quad>(Pt3(2.5f, -1.f, -6.f), Vec3(0.f, 1.5f, 0.3f), Vec3(-0.4f, 0.f, 1.5f)),
        std::make_shared<RGBIlluminantSpectrum>(RGB(2.0f, 1.5f, 1.0f)), 3.0f, false));
    scene.add_plane(Pt3(0.f, -2.f, 0.f), Vec3(0.f, 1.f, 0.f), b.keep(DiffuseMaterial(SolidColor(0.7f, 0.7f, 0.6f))), 50.0f);
    scene.add_triangle(Pt3(-3.f, -2.f, -8.f), Pt3(3.f, -2.f, -8.f), Pt3(0.f, 3.f, -8.5f),
                       b.keep(DiffuseMaterial(DummyTexture{})));
    scene.add_triangle(Pt3(-3.5f, -2.f, -7.f), Pt3(-3.f, 2.f, -7.5f), Pt3(-3.5f, -2.f, -3.f),
                       b.keep(ConductiveMaterial(0.2f, 3.9f)));
    scene.add_sphere(Pt3(-1.5f, -1.2f, -5.f), 0.8f, b.keep(ThinDielectricMaterial(1.5f)));
    scene.add_sphere(Pt3(0.3f, -1.3f, -4.2f), 0.7f, b.keep(DielectricMaterial(1.33f)));
    scene.add_sphere(Pt3(1.8f, -1.1f, -5.5f), 0.9f, b.keep(ThinDielectricMaterial(spectra::GLASS_BK7_IOR())));
    scene.add_sphere(Pt3(0.0f, 0.4f, -6.5f), 0.8f, b.keep(ConductiveMaterial::alluminum(0.0f, 0.0f)));
    scene.add_sphere(Pt3(-2.2f, 0.6f, -6.0f), 0.6f, b.keep(DiffuseMaterial(SolidColor(spectra::CU_IOR()))));
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 400, o.height ? o.height : 300, M_PI / 3.0f,
                                        Transform::translation(0.2f, 0.3f, 1.0f) * Transform::rotate_y(0.1f));
    b.n_samples = 16;
    b.max_bounces = 24;
}

// Edge cases the shipped examples never reach: a scene WITHOUT lights (sample_lights draws its two
// 2-D samples without using them, render.cpp:62-66; all radiance comes from the background) ...
inline void build_no_lights(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    scene.set_bg_light(spectra::ILLUM_D65(), 0.8f);
    scene.add_plane(Pt3(0.f, -1.f, 0.f), Vec3(0.f, 1.f, 0.f), b.keep(DiffuseMaterial(SolidColor(0.6f, 0.5f, 0.4f))), 40.0f);
    scene.add_sphere(Pt3(-1.1f, 0.f, -4.f), 1.0f, b.keep(ConductiveMaterial::copper(0.25, 0.1)));
    scene.add_sphere(Pt3(1.1f, 0.f, -4.5f), 1.0f, b.keep(DiffuseMaterial(DummyTexture{})));
    scene.add_sphere(Pt3(0.f, -0.5f, -2.5f), 0.5f, b.keep(DielectricMaterial(spectra::GLASS_BK7_IOR())));
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 200, o.height ? o.height : 150, M_PI / 3.0f);
    b.n_samples = 8;
    b.max_bounces = 16;
}

// ... and a scene without any geometry: every path misses at depth 0 and returns the background.
inline void build_empty_sky(Bundle& b, const Options& o) {
    Scene& scene = *b.scene;
    scene.set_bg_light(std::make_shared<RGBIlluminantSpectrum>(RGB(0.3f, 0.5f, 0.9f)), 1.5f);
    scene.commit();
    b.camera = std::make_unique<Camera>(o.width ? o.width : 64, o.height ? o.height : 48, M_PI / 3.0f);
    b.n_samples = 4;
    b.max_bounces = 8;
}

// Seeded random scenes for the parity tests ("fuzz:<seed>"): every material, texture, light and shape type of the
// reference's scene API in combinations none of the shipped examples has -- rough and anisotropic conductors next to
// dielectrics with dispersion, mixed materials of three children, image textures, bumpy height grids, emitters of
// every kind, with or without a background; grids under rotated and non-uniformly scaled transforms, triangles without
// area, cameras with a field of view of 0.4 to 2.2 rad under yaw, pitch and roll, both sensor presets with and without
// their own imaging ratio.  Integer generator, float arithmetic and one draw per statement (argument evaluation order is
// unspecified), so both builds of this header make the same scene.
struct FuzzRng {
    uint32_t state;
    uint32_t next() { state = state * 1664525u + 1013904223u; return state >> 8; }
    float uniform(float lo, float hi) { return lo + (hi - lo) * (float(next()) * (1.0f / 16777216.0f)); }
    int pick(int n) { return int(next() % uint32_t(n)); }
    Vec3 vec(float x0, float x1, float y0, float y1, float z0, float z1) {
        const float x = uniform(x0, x1);
        const float y = uniform(y0, y1);
        const float z = uniform(z0, z1);
        return Vec3(x, y, z);
    }
    Pt3 point(float x0, float x1, float y0, float y1, float z0, float z1) {
        const Vec3 v = vec(x0, x1, y0, y1, z0, z1);
        return Pt3(v.x, v.y, v.z);
    }
};

inline std::shared_ptr<const Spectrum> fuzz_illuminant(FuzzRng& r) {
    const int kind = r.pick(8);
    if (kind < 3) return spectra::ILLUM_D65();
    if (kind < 7) {
        const Vec3 c = r.vec(0.2f, 3.f, 0.2f, 3.f, 0.2f, 3.f);
        return std::make_shared<RGBIlluminantSpectrum>(RGB(c.x, c.y, c.z));
    }
    // (one emitter in eight: the reference's BlackbodySpectrum is +inf at every wavelength -- its normalisation divides
    // by blackbody(2.897721e-12 / T), which is 0 (spectrum.cpp:126-129) -- and so are the paths it lights)
    const float kelvin = r.uniform(2500.f, 9000.f);
    return std::make_shared<BlackbodySpectrum>(kelvin);
}

// one non-mixed material, handed to `sink` as its concrete type (the reference's Material has no virtual destructor:
// whoever stores it must know the type)
template <class Sink>
inline void fuzz_simple_material(FuzzRng& r, Sink&& sink) {
    const int kind = r.pick(10);
    const float a = r.uniform(0.f, 1.f);
    const float c = r.uniform(0.f, 1.f);
    const float d = r.uniform(0.f, 1.f);
    const bool flag = r.pick(2) != 0;
    if (kind == 9) {
        // an image texture of a few random texels (texture.cpp: ImageTexture::value -> RGB-to-spectrum lookup per texel)
        const int rows = 1 + r.pick(7);
        const int cols = 1 + r.pick(7);
        Image image{size_t(rows), size_t(cols)};
        for (size_t i = 0; i < image.color_buffer.size(); i++) image.color_buffer[i] = r.uniform(0.f, 1.f);
        sink(DiffuseMaterial(ImageTexture(std::move(image))));
        return;
    }
    switch (kind) {
        case 0: sink(DiffuseMaterial(SolidColor(a, c, d))); break;
        case 1: sink(DiffuseMaterial(DummyTexture{})); break;
        case 2: sink(ConductiveMaterial(0.1f + 1.9f * a, 1.0f + 4.0f * c)); break;
        case 3: sink(ConductiveMaterial::copper(0.6f * a, 0.6f * c)); break;
        case 4: sink(ConductiveMaterial::alluminum(flag ? 0.01f + 0.5f * a : 0.0f, 0.5f * c)); break;
        case 5: sink(DielectricMaterial(1.1f + 1.3f * a)); break;
        case 6: sink(DielectricMaterial(flag ? spectra::GLASS_BK7_IOR() : spectra::GLASS_SF11_IOR())); break;
        case 7: sink(ThinDielectricMaterial(1.1f + 0.9f * a)); break;
        default: sink(DiffuseMaterial(SolidColor(spectra::CU_IOR()))); break;
    }
}

inline std::unique_ptr<Material> fuzz_child(FuzzRng& r) {
    std::unique_ptr<Material> out;
    fuzz_simple_material(r, [&](auto&& m) { out = std::make_unique<std::decay_t<decltype(m)>>(std::move(m)); });
    return out;
}

inline const Material* fuzz_material(Bundle& b, FuzzRng& r) {
    const int kind = r.pick(6);
    if (kind == 0) {
        std::unique_ptr<Material> m0 = fuzz_child(r);
        std::unique_ptr<Material> m1 = fuzz_child(r);
        const float w0 = r.uniform(0.1f, 2.f);
        const float w1 = r.uniform(0.1f, 2.f);
        std::array<std::unique_ptr<Material>, 2> parts = {std::move(m0), std::move(m1)};
        std::array<float, 2> weights = {w0, w1};
        b.materials.push_back(std::make_shared<MixedMaterial<2>>(std::move(parts), std::move(weights)));
        return b.materials.back().get();
    }
    if (kind == 1) {
        std::unique_ptr<Material> m0 = fuzz_child(r);
        std::unique_ptr<Material> m1 = fuzz_child(r);
        std::unique_ptr<Material> m2 = fuzz_child(r);
        const float w1 = r.uniform(0.1f, 2.f);
        const float w2 = r.uniform(0.1f, 2.f);
        std::array<std::unique_ptr<Material>, 3> parts = {std::move(m0), std::move(m1), std::move(m2)};
        std::array<float, 3> weights = {1.0f, w1, w2};
        b.materials.push_back(std::make_shared<MixedMaterial<3>>(std::move(parts), std::move(weights)));
        return b.materials.back().get();
    }
    const Material* out = nullptr;
    fuzz_simple_material(r, [&](auto&& m) { out = b.keep(std::move(m)); });
    return out;
}

inline void build_fuzz(Bundle& b, const Options& o, uint32_t seed) {
    FuzzRng r{seed * 2654435761u + 12345u};
    for (int i = 0; i < 4; i++) r.next();
    Scene& scene = *b.scene;
    if (r.pick(3) != 0) {
        auto spectrum = fuzz_illuminant(r);
        const float scale = r.uniform(0.02f, 0.6f);
        scene.set_bg_light(spectrum, scale);
    }
    const int n_lights = r.pick(4);   // (0: only the background, if any, lights the scene)
    for (int i = 0; i < n_lights; i++) {
        const Pt3 p = r.point(-3.f, 3.f, 0.5f, 3.5f, -8.f, -2.f);
        const int kind = r.pick(3);
        auto spectrum = fuzz_illuminant(r);
        const float scale = r.uniform(1.f, 10.f);
        const bool two_sided = r.pick(2) == 0;
        if (kind == 0) {
            scene.add_light(std::make_unique<PointLight>(p, spectrum, 3.0f * scale));
        } else if (kind == 1) {
            const float radius = r.uniform(0.15f, 0.6f);
            scene.add_light(std::make_unique<AreaLight>(std::make_unique<Sphere>(p, radius), spectrum, scale, two_sided));
        } else {
            const Vec3 u = r.vec(0.3f, 1.5f, -0.3f, 0.3f, 0.f, 0.f);
            const Vec3 v = r.vec(0.f, 0.f, -0.3f, 0.3f, 0.3f, 1.5f);
            scene.add_light(std::make_unique<AreaLight>(std::make_unique<Quad>(p, u, v), spectrum, scale, two_sided));
        }
    }
    scene.add_plane(Pt3(0.f, -2.f, 0.f), Vec3(0.f, 1.f, 0.f), fuzz_material(b, r), 40.0f);
    if (r.pick(2)) {
        const float tilt = r.uniform(-0.2f, 0.2f);
        scene.add_plane(Pt3(0.f, 0.f, -10.f), Vec3(tilt, 0.f, 1.f), fuzz_material(b, r), 12.0f);
    }
    const int n_shapes = 3 + r.pick(6);
    for (int i = 0; i < n_shapes; i++) {
        const Pt3 c = r.point(-3.f, 3.f, -1.8f, 1.5f, -8.5f, -3.f);
        const Material* m = fuzz_material(b, r);
        const int kind = r.pick(4);
        const Vec3 u = r.vec(0.5f, 2.f, -0.4f, 0.4f, -0.6f, 0.6f);
        const Vec3 v = r.vec(-0.4f, 0.4f, 0.5f, 2.f, -0.6f, 0.6f);
        if (kind == 0) {
            scene.add_sphere(c, 0.5f * u.x, m);
        } else if (kind == 1) {
            // (one triangle in eight has no area: its three corners lie on a line)
            const bool flat = r.pick(8) == 0;
            scene.add_triangle(c, c + u, flat ? c + u * 2.5f : c + v, m);
        } else if (kind == 2) {
            scene.add_quad(c, c + u, c + u + v, c + v, m);
        } else {
            // a height grid (Scene::add_grid, scene.cpp:375-430): rows x cols vertices over the patch u x v, bumpy
            const int rows = 2 + r.pick(6);
            const int cols = 2 + r.pick(6);
            Image grid{size_t(rows), size_t(cols)};
            for (int y = 0; y < rows; y++) {
                for (int x = 0; x < cols; x++) {
                    const float fu = float(x) / float(cols - 1);
                    const float fv = float(y) / float(rows - 1);
                    const float bump = r.uniform(-0.25f, 0.25f);
                    const size_t at = 3 * (size_t(y) * size_t(cols) + size_t(x));
                    grid.color_buffer[at + 0] = fu * u.x + fv * v.x;
                    grid.color_buffer[at + 1] = fu * u.y + fv * v.y;
                    grid.color_buffer[at + 2] = fu * u.z + fv * v.z + bump;
                }
            }
            const float turn = r.uniform(-0.6f, 0.6f);
            const float roll = r.uniform(-0.8f, 0.8f);
            const Vec3 stretch = r.vec(0.5f, 1.6f, 0.5f, 1.6f, 0.5f, 1.6f);
            scene.add_grid(grid, m, Transform::translation(c.x, c.y, c.z) * Transform::rotate_x(turn) * Transform::rotate_z(roll) *
                                        Transform::scale(stretch.x, stretch.y, stretch.z));
        }
    }
    scene.commit();
    const Vec3 eye = r.vec(-0.5f, 0.5f, -0.3f, 0.6f, 0.f, 1.5f);
    const float yaw = r.uniform(-0.2f, 0.2f);
    const float pitch = r.uniform(-0.15f, 0.15f);
    const float roll = r.uniform(-0.3f, 0.3f);
    const float fov = r.uniform(0.4f, 2.2f);
    const Transform pose = Transform::translation(eye.x, eye.y, eye.z) * Transform::rotate_y(yaw) * Transform::rotate_x(pitch) * Transform::rotate_z(roll);
    const size_t cam_w = o.width ? o.width : 96, cam_h = o.height ? o.height : 72;
    // (one camera in four carries the CIE XYZ sensor instead of the default Canon curves, half of those with their own
    // imaging ratio: sensor.cpp:72-89)
    const int sensor_kind = r.pick(8);
    const float ratio = r.uniform(0.005f, 0.05f);
    if (sensor_kind == 0) b.camera = std::make_unique<Camera>(cam_w, cam_h, fov, pose, PixelSensor::CIE_XYZ());
    else if (sensor_kind == 1) b.camera = std::make_unique<Camera>(cam_w, cam_h, fov, pose, PixelSensor::CIE_XYZ(ratio));
    else if (sensor_kind == 2) b.camera = std::make_unique<Camera>(cam_w, cam_h, fov, pose, PixelSensor::CANON_EOS(ratio));
    else b.camera = std::make_unique<Camera>(cam_w, cam_h, fov, pose);
    b.n_samples = 4;
    b.max_bounces = 24;
}

inline const char* const* scene_names(int* n) {
    static const char* const names[] = {"cornell_box", "glass_spheres", "textures", "opposing_planes",
                                        "obj_viewer", "cornell_mixed", "mandelbrot", "kitchen_sink"};
    *n = int(sizeof(names) / sizeof(names[0]));
    return names;
}

// returns nullptr for an unknown name or a mesh that failed to load
inline std::unique_ptr<Bundle> build(const std::string& name, const Options& o) {
    auto b = std::make_unique<Bundle>();
    b->scene = std::make_unique<Scene>(initialize_device());
    if (name == "cornell_box") build_cornell(*b, o, false);
    else if (name == "cornell_mixed") build_cornell(*b, o, true);
    else if (name == "glass_spheres") build_glass_spheres(*b, o);
    else if (name == "textures") build_textures(*b, o);
    else if (name == "opposing_planes") build_opposing_planes(*b, o);
    else if (name == "mandelbrot") build_mandelbrot(*b, o, 96);
    else if (name == "mandelbrot_full") build_mandelbrot(*b, o, 1200);  // the grid as shipped: 1199 x 1199 cells = 2.87 M triangles
    else if (name == "kitchen_sink") build_kitchen_sink(*b, o);
    else if (name == "no_lights") build_no_lights(*b, o);
    else if (name == "empty_sky") build_empty_sky(*b, o);
    else if (name == "obj_viewer") { if (!build_obj_viewer(*b, o)) return nullptr; }
    else if (name.rfind("fuzz:", 0) == 0) build_fuzz(*b, o, uint32_t(std::strtoul(name.c_str() + 5, nullptr, 10)));
    else return nullptr;
    return b;
}

}  // namespace qzscenes
