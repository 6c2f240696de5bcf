// The same program against two libraries: the reference's sources (oracle/_ref) and the host library.  Prints what a
// user of the scene API observes in the corner cases; the test compares the two outputs line by line.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "render.hpp"

static uint64_t plane_hash(const std::vector<float>& v) {
    uint64_t h = 1469598103934665603ull;
    for (float f : v) {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        for (int k = 0; k < 4; k++) { h ^= (u >> (8 * k)) & 0xffu; h *= 1099511628211ull; }
    }
    return h;
}

int main() {
    Scene scene(initialize_device());
    DiffuseMaterial grey(SolidColor(0.5f, 0.5f, 0.5f));
    std::cout << "ready before commit: " << scene.ready() << std::endl;
    auto* missing = scene.add_obj("/nonexistent/mesh.obj", &grey);
    std::cout << "add_obj(missing file) -> " << (missing ? "geometry" : "nullptr") << std::endl;
    Camera camera(8, 6, M_PI / 3.0f);
    RenderResult early = render(camera, scene, 2, 4);
    double sum = 0.0;
    for (float f : early.color_buffer) sum += f;
    std::cout << "render(not committed): " << early.width << "x" << early.height << " planes " << early.color_buffer.size() << " "
              << early.normal_buffer.size() << " " << early.albedo_buffer.size() << " sum " << sum << std::endl;
    scene.add_sphere(Pt3(0.f, 0.f, -4.f), 1.0f, &grey);
    scene.add_plane(Pt3(0.f, -1.f, 0.f), Vec3(0.f, 1.f, 0.f), &grey, 20.0f);
    scene.add_light(std::make_unique<PointLight>(Pt3(2.f, 3.f, -1.f), spectra::ILLUM_D65(), 20.0f));
    scene.commit();
    std::cout << "ready after commit: " << scene.ready() << std::endl;
    RenderResult film = render(camera, scene, 3, 5);
    std::cout << "film " << film.width << "x" << film.height << " color " << std::hex << plane_hash(film.color_buffer) << " normal "
              << plane_hash(film.normal_buffer) << " albedo " << plane_hash(film.albedo_buffer) << std::dec << std::endl;
    // render() leaves std::cout the way its "Render time" line set it (render.cpp:394): fixed notation, three decimals
    std::cout << "floats after render(): " << 0.2f << " " << 1234.56789 << " " << 3 << std::endl;
    RenderResult none = render(camera, scene, 1, 0);
    std::cout << "zero bounces color " << std::hex << plane_hash(none.color_buffer) << std::dec << std::endl;
    // no samples at all: the reference divides the empty sums by float(0) (render.cpp:280-282)
    RenderResult empty = render(camera, scene, 0, 4);
    std::cout << "zero samples color " << std::hex << plane_hash(empty.color_buffer) << " normal " << plane_hash(empty.normal_buffer)
              << " albedo " << plane_hash(empty.albedo_buffer) << std::dec << std::endl;
    // a second render of the same scene, another camera: nothing is left over from the first
    Camera wide(5, 9, M_PI / 2.0f);
    RenderResult again = render(wide, scene, 2, 3);
    std::cout << "second camera " << again.width << "x" << again.height << " color " << std::hex << plane_hash(again.color_buffer)
              << std::dec << std::endl;
    // geometry data handed back by add_*: stays valid, carries the material
    Scene other(initialize_device());
    auto* tri = other.add_triangle(Pt3(0.f, 0.f, -3.f), Pt3(1.f, 0.f, -3.f), Pt3(0.f, 1.f, -3.f), &grey);
    auto* quad = other.add_quad(Pt3(-1.f, -1.f, -5.f), Pt3(1.f, -1.f, -5.f), Pt3(1.f, 1.f, -5.f), Pt3(-1.f, 1.f, -5.f), &grey);
    std::cout << "geometry material kept: " << (tri && tri->material == &grey) << (quad && quad->material == &grey) << std::endl;
    other.set_bg_light(spectra::ILLUM_D65(), 0.5f);
    other.commit();
    RenderResult sky = render(camera, other, 2, 2);
    std::cout << "background light color " << std::hex << plane_hash(sky.color_buffer) << " albedo " << plane_hash(sky.albedo_buffer)
              << std::dec << std::endl;
    // an empty committed scene
    Scene nothing(initialize_device());
    nothing.commit();
    std::cout << "empty scene ready: " << nothing.ready() << std::endl;
    RenderResult black = render(camera, nothing, 2, 2);
    std::cout << "empty scene color " << std::hex << plane_hash(black.color_buffer) << " normal " << plane_hash(black.normal_buffer)
              << std::dec << std::endl;
    // sensors, fields of view, camera transforms
    ConductiveMaterial metal = ConductiveMaterial::copper(0.2f, 0.1f);
    Scene lit(initialize_device());
    lit.add_sphere(Pt3(0.f, 0.f, -4.f), 1.0f, &metal);
    lit.add_plane(Pt3(0.f, -1.f, 0.f), Vec3(0.f, 1.f, 0.f), &grey, 20.0f);
    lit.add_light(std::make_unique<PointLight>(Pt3(2.f, 3.f, -1.f), spectra::ILLUM_D65(), 20.0f));
    lit.set_bg_light(spectra::ILLUM_D65(), 0.1f);
    lit.commit();
    {
        Camera c(9, 7, M_PI / 3.0f, Transform::identity(), PixelSensor::CIE_XYZ());
        RenderResult f = render(c, lit, 3, 6);
        std::cout << "XYZ sensor " << std::hex << plane_hash(f.color_buffer) << " " << plane_hash(f.albedo_buffer) << std::dec << std::endl;
        Camera c2(9, 7, M_PI / 3.0f, Transform::identity(), PixelSensor::CIE_XYZ(0.01f));
        RenderResult f2 = render(c2, lit, 3, 6);
        std::cout << "XYZ sensor, imaging ratio " << std::hex << plane_hash(f2.color_buffer) << std::dec << std::endl;
        Camera c3(9, 7, M_PI / 3.0f, Transform::identity(), PixelSensor::CANON_EOS(0.05f));
        RenderResult f3 = render(c3, lit, 3, 6);
        std::cout << "Canon sensor, imaging ratio " << std::hex << plane_hash(f3.color_buffer) << std::dec << std::endl;
    }
    for (float fov : {0.2f, 1.0f, 2.5f, 3.0f}) {
        Camera c(7, 11, fov, Transform::translation(0.3f, 0.5f, 1.0f) * Transform::rotate_z(0.3f) * Transform::rotate_x(-0.1f));
        RenderResult f = render(c, lit, 2, 4);
        std::cout << "fov " << fov << " " << std::hex << plane_hash(f.color_buffer) << " " << plane_hash(f.normal_buffer) << std::dec << std::endl;
    }
    {
        Camera c(6, 6, 1.0f, Transform::scale(2.0f, 0.5f, 1.5f) * Transform::rotate_y(0.2f));
        RenderResult f = render(c, lit, 2, 4);
        std::cout << "scaled camera " << std::hex << plane_hash(f.color_buffer) << std::dec << std::endl;
    }
    // degenerate inputs of add_grid / add_obj: which ones yield a geometry, and which geometry IDs the survivors take
    Scene odd(initialize_device());
    Image thin{size_t(5), size_t(1)};
    std::cout << "grid without cells -> " << (odd.add_grid(thin, &grey) ? "geometry" : "nullptr") << std::endl;
    Image too_wide{size_t(2), size_t(70000)};
    std::cout << "grid 70000 wide -> " << (odd.add_grid(too_wide, &grey) ? "geometry" : "nullptr") << std::endl;
    const std::string tmp = std::getenv("API_TMP") ? std::getenv("API_TMP") : "/tmp";
    { std::ofstream f(tmp + "/api_empty.obj"); }
    std::cout << "empty obj -> " << (odd.add_obj(tmp + "/api_empty.obj", &grey) ? "geometry" : "nullptr") << std::endl;
    { std::ofstream f(tmp + "/api_verts.obj"); f << "v 0 0 -3\nv 1 0 -3\nv 0 1 -3\n"; }
    std::cout << "obj without faces -> " << (odd.add_obj(tmp + "/api_verts.obj", &grey) ? "geometry" : "nullptr") << std::endl;
    { std::ofstream f(tmp + "/api_tri.obj"); f << "v 0 0 -3\nv 1 0 -3\nv 0 1 -3\nf 1 2 3\n"; }
    std::cout << "obj triangle -> " << (odd.add_obj(tmp + "/api_tri.obj", &grey) ? "geometry" : "nullptr") << std::endl;
    for (unsigned id = 0; id < 2; id++) std::cout << "geometry " << id << " shape " << int(odd.get_geom_data(id)->shape) << std::endl;
    odd.add_light(std::make_unique<PointLight>(Pt3(2.f, 3.f, -1.f), spectra::ILLUM_D65(), 20.0f));
    odd.commit();
    RenderResult odd_film = render(camera, odd, 3, 5);
    std::cout << "degenerate scene color " << std::hex << plane_hash(odd_film.color_buffer) << " normal " << plane_hash(odd_film.normal_buffer)
              << std::dec << std::endl;
    return 0;
}
