// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C entry points (for ctypes) over the reference's UNMODIFIED sources, compiled in
// place from /root/reference by oracle/Makefile into oracle/_ref/liboracle_ref.so.
// This TU #includes the reference's render.cpp so that its file-local symbols
// sample_pixel() (render.cpp:91) and PixelSample (render.cpp:19) are reachable for
// per-path replay; everything else links against the reference's other objects.
// Ray/scene intersection goes through the Embree shim in oracle/embree_shim
// (Embree 4.3 is absent -- see rtcore_shim.cpp; parity vs real Embree UNPINNED).
//
// The same entry points, with the prefix qzh_ instead of orc_, are exported by the
// product harness (quetzalcoatlus_b200/harness/qz_harness.cpp) so that the tests
// drive both sides through one Python wrapper.

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <iomanip>
#include <numeric>
#include <string>
#include <unistd.h>

#include "render.cpp"  // the reference integrator, verbatim (-I/root/reference/src)
#include "obj/obj.hpp"

#include "scenes.hpp"

// defined (non-static, undeclared in any header) at scene.cpp:41
RTCRayHit create_rayhit(const Ray& ray, RTCScene scene);

namespace {

struct StdoutSilencer {
    int saved = -1;
    StdoutSilencer() {
        fflush(stdout); std::cout.flush();
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        if (nul >= 0) { dup2(nul, 1); close(nul); }
    }
    ~StdoutSilencer() {
        fflush(stdout); std::cout.flush();
        if (saved >= 0) { dup2(saved, 1); close(saved); }
    }
};

std::shared_ptr<const Spectrum> spectrum_by_name(const std::string& name) {
    if (name == "X") return spectra::X();
    if (name == "Y") return spectra::Y();
    if (name == "Z") return spectra::Z();
    if (name == "D65") return spectra::ILLUM_D65();
    if (name == "CANON_R") return spectra::CANON_EOS_R();
    if (name == "CANON_G") return spectra::CANON_EOS_G();
    if (name == "CANON_B") return spectra::CANON_EOS_B();
    if (name == "AL_IOR") return spectra::AL_IOR();
    if (name == "AL_ABSORPTION") return spectra::AL_ABSORPTION();
    if (name == "CU_IOR") return spectra::CU_IOR();
    if (name == "CU_ABSORPTION") return spectra::CU_ABSORPTION();
    if (name == "GLASS_BK7_IOR") return spectra::GLASS_BK7_IOR();
    if (name == "GLASS_SF11_IOR") return spectra::GLASS_SF11_IOR();
    float r, g, b;
    if (sscanf(name.c_str(), "rgb:%f,%f,%f", &r, &g, &b) == 3)
        return std::make_shared<RGBSigmoidPolynomial>(RGBColorSpace::sRGB()->to_spectrum(RGB(r, g, b)));
    if (sscanf(name.c_str(), "rgbu:%f,%f,%f", &r, &g, &b) == 3) return std::make_shared<RGBUnboundedSpectrum>(RGB(r, g, b));
    if (sscanf(name.c_str(), "rgbi:%f,%f,%f", &r, &g, &b) == 3) return std::make_shared<RGBIlluminantSpectrum>(RGB(r, g, b));
    if (sscanf(name.c_str(), "const:%f", &r) == 1) return std::make_shared<ConstantSpectrum>(r);
    if (sscanf(name.c_str(), "blackbody:%f", &r) == 1) return std::make_shared<BlackbodySpectrum>(r);
    return nullptr;
}

}  // namespace

extern "C" {

// The reference caches its RGB->spectrum table as coeffs_SRGB_32.dat in the CWD
// (rgb_to_spectrum_opt.cpp:876-898).  Build the table with CWD = data_dir so that the
// oracle and the product read the very same file.
int orc_init(const char* data_dir) {
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) return 1;
    if (data_dir && *data_dir && chdir(data_dir) != 0) return 2;
    {
        StdoutSilencer quiet;
        RGBColorSpace::sRGB();
    }
    if (chdir(cwd) != 0) return 3;
    return 0;
}

const char* orc_impl(void) { return "reference sources + embree shim (oracle/_ref)"; }

void* orc_scene_build(const char* name, int width, int height, const char* obj_path, const char* obj_material,
                      const char* obj_light) {
    qzscenes::Options o;
    o.width = width; o.height = height;
    if (obj_path) o.obj_path = obj_path;
    if (obj_material && *obj_material) o.obj_material = obj_material;
    if (obj_light && *obj_light) o.obj_light = obj_light;
    StdoutSilencer quiet;
    return qzscenes::build(name, o).release();
}

void orc_scene_free(void* h) { delete static_cast<qzscenes::Bundle*>(h); }

void orc_scene_info(void* h, int* w, int* hgt, int* spp, int* max_bounces) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    *w = int(b->camera->image_width); *hgt = int(b->camera->image_height);
    *spp = int(b->n_samples); *max_bounces = int(b->max_bounces);
}

// full render through the reference's own render() (render.cpp:321); seconds = wall
// time around the call, rays = rtcIntersect1 calls (closest-hit + occlusion)
int orc_render(void* h, int spp, int max_bounces, float* color, float* normal, float* albedo, double* seconds,
               unsigned long long* rays, int* n_threads) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    rtcShimResetRayCount();
    auto t0 = std::chrono::steady_clock::now();
    RenderResult r = [&] {
        StdoutSilencer quiet;
        return render(*b->camera, *b->scene, size_t(spp), size_t(max_bounces));
    }();
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (rays) *rays = rtcShimRayCount();
    if (n_threads) {
        size_t px = r.width * r.height;
        *n_threads = int(std::clamp<size_t>(std::thread::hardware_concurrency(), 1, (px + THREAD_JOB_SIZE - 1) / THREAD_JOB_SIZE));
    }
    size_t n = r.width * r.height * 3;
    if (color) std::memcpy(color, r.color_buffer.data(), n * sizeof(float));
    if (normal) std::memcpy(normal, r.normal_buffer.data(), n * sizeof(float));
    if (albedo) std::memcpy(albedo, r.albedo_buffer.data(), n * sizeof(float));
    return 0;
}

// Replay n pixel-samples (x, y, s) exactly as render_pixels() does (render.cpp:267-278;
// NOTE y is the already-flipped sampler/camera y, i.e. y = H-1-row) and dump one
// 32-float record per path:
//   [0..3] lambda  [4..7] lambda pdf (final)  [8..11] radiance L  [12..14] normal
//   [15] rays issued  [16..19] albedo (spectral)  [20..22] sensor rgb of L
//   [23..25] sensor rgb of albedo  [26..31] 0
int orc_trace_paths(void* h, int spp, int max_bounces, int n, const int* xys, float* records) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    const Camera& camera = *b->camera;
    Sampler sampler(spp, int(camera.image_width), int(camera.image_height), 0);
    for (int i = 0; i < n; i++) {
        int x = xys[3 * i], y = xys[3 * i + 1], s = xys[3 * i + 2];
        float* rec = records + size_t(i) * 32;
        std::memset(rec, 0, 32 * sizeof(float));
        unsigned long long rays0 = rtcShimThreadRayCount();
        sampler.start_pixel_sample(x, y, s);
        auto jitter = sampler.sample_pixel();
        float u = float(x) + jitter.x;
        float v = float(y) + jitter.y;
        Ray r = camera.cast_ray(u, v);
        WavelengthSample wavelengths = WavelengthSample::uniform(sampler.sample_1d());
        for (int k = 0; k < 4; k++) rec[k] = wavelengths.m_lambdas[k];
        auto pxs = sample_pixel(r, *b->scene, wavelengths, sampler, size_t(max_bounces));
        RGB c = camera.sensor.to_sensor_rgb(pxs.color, wavelengths);
        RGB a = camera.sensor.to_sensor_rgb(pxs.albedo, wavelengths);
        for (int k = 0; k < 4; k++) {
            rec[4 + k] = wavelengths.m_pdf[k];
            rec[8 + k] = pxs.color[k];
            rec[16 + k] = pxs.albedo[k];
        }
        rec[12] = pxs.normal.x; rec[13] = pxs.normal.y; rec[14] = pxs.normal.z;
        rec[15] = float(rtcShimThreadRayCount() - rays0);
        rec[20] = c.x; rec[21] = c.y; rec[22] = c.z;
        rec[23] = a.x; rec[24] = a.y; rec[25] = a.z;
    }
    return 0;
}

// q = (x, y, s, dim): dim >= 2 -> the Owen-scrambled value of that dimension;
// dim == 0 / 1 -> pixel jitter x / y (sampler.cpp:449-454)
int orc_sampler_eval(int spp, int w, int h, int n, const int* q, float* out) {
    Sampler sampler(spp, w, h, 0);
    for (int i = 0; i < n; i++) {
        int x = q[4 * i], y = q[4 * i + 1], s = q[4 * i + 2], dim = q[4 * i + 3];
        sampler.start_pixel_sample(x, y, s, dim);
        if (dim < 2) {
            auto j = sampler.sample_pixel();
            out[i] = dim == 0 ? j.x : j.y;
        } else {
            out[i] = sampler.sample_1d();
        }
    }
    return 0;
}

int orc_eval_spectrum(const char* name, int n, const float* lambdas, float* out) {
    auto sp = spectrum_by_name(name);
    if (!sp) return 1;
    for (int i = 0; i < n; i++) out[i] = (*sp)(lambdas[i]);
    return 0;
}

// Camera fields in declaration order (camera.hpp:24-31): pos, look_at, up, right,
// viewport_bottom_left, pixel_delta_u, pixel_delta_v -> 21 floats
int orc_camera_fields(void* h, float* out) {
    const Camera& c = *static_cast<qzscenes::Bundle*>(h)->camera;
    const Vec3* f[7] = {&c.pos, &c.look_at, &c.up, &c.right, &c.viewport_bottom_left, &c.pixel_delta_u, &c.pixel_delta_v};
    for (int i = 0; i < 7; i++) { out[3 * i] = f[i]->x; out[3 * i + 1] = f[i]->y; out[3 * i + 2] = f[i]->z; }
    return 0;
}

// in: n x (u, L0..L3); out: n x rgb through WavelengthSample::uniform(u) and the
// camera's PixelSensor::to_sensor_rgb (sensor.cpp:57-70)
int orc_sensor_eval(void* h, int n, const float* in, float* out) {
    const Camera& c = *static_cast<qzscenes::Bundle*>(h)->camera;
    for (int i = 0; i < n; i++) {
        auto wl = WavelengthSample::uniform(in[5 * i]);
        SpectrumSample L(std::array<float, 4>{in[5 * i + 1], in[5 * i + 2], in[5 * i + 3], in[5 * i + 4]});
        RGB rgb = c.sensor.to_sensor_rgb(L, wl);
        out[3 * i] = rgb.x; out[3 * i + 1] = rgb.y; out[3 * i + 2] = rgb.z;
    }
    return 0;
}

// closest-hit probe straight through Scene::ray_intersect's Embree call
// (scene.cpp:41-59): in n x (o, d); out n x (t, u, v, Ng.xyz, geomID, primID) with
// t = -1 on a miss.  Used to pin the GPU traversal kernel against the shim.
int orc_intersect(void* h, int n, const float* rays, float* out) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    for (int i = 0; i < n; i++) {
        const float* r = rays + 6 * i;
        auto rh = create_rayhit(Ray(Pt3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5])), b->scene->get_scene());
        float* o = out + 8 * i;
        if (rh.hit.geomID == RTC_INVALID_GEOMETRY_ID) {
            o[0] = -1.0f; for (int k = 1; k < 8; k++) o[k] = 0.0f;
        } else {
            o[0] = rh.ray.tfar; o[1] = rh.hit.u; o[2] = rh.hit.v;
            o[3] = rh.hit.Ng_x; o[4] = rh.hit.Ng_y; o[5] = rh.hit.Ng_z;
            o[6] = float(rh.hit.geomID); o[7] = float(rh.hit.primID);
        }
    }
    return 0;
}

void orc_force_brute_force(int on) { rtcShimForceBruteForce(on); }

// ObjData of a file as the reference's own loader (obj/obj.cpp:9-175) reads it, flattened for the loader tests: counts =
// {vertices, normals, faces}; arrays may be null (sizing call) and otherwise hold 4 floats per vertex (x y z w), 3 per
// normal, 13 ints per face (vertices[4], textures[4], normals[4], n_vertices).  Returns 1 when the loader returns nullopt.
int orc_obj_load(const char* path, long* counts, float* vertices, float* normals, int* faces) {
    auto data = obj::load_obj(path);
    if (!data) return 1;
    counts[0] = (long)data->vertices.size(); counts[1] = (long)data->vertex_normals.size(); counts[2] = (long)data->faces.size();
    for (size_t i = 0; vertices && i < data->vertices.size(); i++) {
        const auto& v = data->vertices[i];
        vertices[4 * i] = v.x; vertices[4 * i + 1] = v.y; vertices[4 * i + 2] = v.z; vertices[4 * i + 3] = v.w;
    }
    for (size_t i = 0; normals && i < data->vertex_normals.size(); i++) {
        const auto& n = data->vertex_normals[i];
        normals[3 * i] = n.x; normals[3 * i + 1] = n.y; normals[3 * i + 2] = n.z;
    }
    for (size_t i = 0; faces && i < data->faces.size(); i++) {
        const auto& f = data->faces[i];
        for (int k = 0; k < 4; k++) { faces[13 * i + k] = f.vertices[k]; faces[13 * i + 4 + k] = f.textures[k]; faces[13 * i + 8 + k] = f.normals[k]; }
        faces[13 * i + 12] = (int)f.n_vertices;
    }
    return 0;
}

}  // extern "C"
