// C entry points (for ctypes) over the host library: the same functions, with the same
// signatures, that oracle/ref_harness.cpp exports for the reference (prefix orc_), here with
// the prefix qzh_ and running on the B200 through the C ABI of include/qz_b200.h.  The scene
// definitions are the shared header scenes/scenes.hpp, compiled here against
// quetzalcoatlus_b200/host instead of /root/reference/src.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "obj/obj.hpp"
#include "render.hpp"
#include "scenes.hpp"

extern "C" {

int qzh_init(const char* data_dir) {
    if (data_dir && *data_dir) qzhost::set_data_dir(data_dir);
    try {
        RGBColorSpace::sRGB();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "qzh_init: %s\n", e.what());
        return 1;
    }
    return 0;
}

const char* qzh_impl(void) { return "quetzalcoatlus_b200 host library over the qz_b200 C ABI"; }

void* qzh_scene_build(const char* name, int width, int height, const char* obj_path, const char* obj_material,
                      const char* obj_light) {
    qzscenes::Options o;
    o.width = width; o.height = height;
    if (obj_path) o.obj_path = obj_path;
    if (obj_material && *obj_material) o.obj_material = obj_material;
    if (obj_light && *obj_light) o.obj_light = obj_light;
    try {
        qzhost::reset_build_times();
        auto b = qzscenes::build(name, o);
        if (b && !b->scene->ready()) return nullptr;  // commit failed (e.g. no CUDA device)
        return b.release();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "qzh_scene_build: %s\n", e.what());
        return nullptr;
    }
}

void qzh_scene_free(void* h) { delete static_cast<qzscenes::Bundle*>(h); }

void qzh_scene_info(void* h, int* w, int* hgt, int* spp, int* max_bounces) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    *w = int(b->camera->image_width); *hgt = int(b->camera->image_height);
    *spp = int(b->n_samples); *max_bounces = int(b->max_bounces);
}

// full render through render() (host buffers in and out); seconds = wall time around the
// call INCLUDING host<->device copies, rays = closest-hit + shadow queries
int qzh_render(void* h, int spp, int max_bounces, float* color, float* normal, float* albedo, double* seconds,
               unsigned long long* rays, int* n_threads) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    auto t0 = std::chrono::steady_clock::now();
    RenderResult r = render(*b->camera, *b->scene, size_t(spp), size_t(max_bounces));
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const qz_stats& st = qzhost::last_render_stats();
    if (rays) *rays = st.rays_closest + st.rays_shadow;
    if (n_threads) *n_threads = 0;
    size_t n = r.width * r.height * 3;
    if (color) std::memcpy(color, r.color_buffer.data(), n * sizeof(float));
    if (normal) std::memcpy(normal, r.normal_buffer.data(), n * sizeof(float));
    if (albedo) std::memcpy(albedo, r.albedo_buffer.data(), n * sizeof(float));
    return 0;
}

// render() exactly as an application of the reference's API calls it: the RenderResult stays on the C++ side (one value
// of it is read back), so the time is render()'s own -- table uploads, kernels, the three planes device->host into the
// RenderResult's pageable vectors -- without this harness's copy into caller buffers (bench.py: e2e)
int qzh_render_only(void* h, int spp, int max_bounces, double* seconds, float* first_value) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    auto t0 = std::chrono::steady_clock::now();
    RenderResult r = render(*b->camera, *b->scene, size_t(spp), size_t(max_bounces));
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (first_value) *first_value = r.color_buffer.empty() ? 0.0f : r.color_buffer[0];
    return r.color_buffer.empty() ? 1 : 0;
}

// scene-build times of this thread since the last qzh_scene_build (host/render.hpp: BuildTimes)
void qzh_build_times(double* obj_parse_ms, double* commit_ms, double* bvh_build_ms) {
    const qzhost::BuildTimes& t = qzhost::build_times();
    if (obj_parse_ms) *obj_parse_ms = t.obj_parse_ms;
    if (commit_ms) *commit_ms = t.commit_ms;
    if (bvh_build_ms) *bvh_build_ms = t.bvh_build_ms;
}

// stats of the last qzh_render on this thread (the struct of include/qz_b200.h)
void qzh_last_stats(qz_stats* out) { *out = qzhost::last_render_stats(); }

// the C-ABI scene handle and flattened camera of a built scene, for callers (bench.py, the
// multi-GPU driver) that talk to qz_render_device() directly
void* qzh_scene_handle(void* h) { return static_cast<qzscenes::Bundle*>(h)->scene->handle(); }

// fills *cam; the sensor curves it points to stay valid until the scene is freed
int qzh_camera(void* h, qz_camera* cam) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    static thread_local std::vector<float> sensor;  // one live camera per thread is enough for the harness
    *cam = qzhost::flatten_camera(*b->camera, sensor);
    return 0;
}

int qzh_trace_paths(void* h, int spp, int max_bounces, int n, const int* xys, float* records) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    std::vector<float> sensor;
    qz_camera cam = qzhost::flatten_camera(*b->camera, sensor);
    return qz_trace_paths(b->scene->handle(), &cam, uint32_t(spp), uint32_t(max_bounces), uint32_t(n), xys, records);
}

int qzh_sampler_eval(int spp, int w, int h, int n, const int* q, float* out) {
    return qz_sampler_eval(uint32_t(spp), uint32_t(w), uint32_t(h), uint32_t(n), q, out);
}

// Spectrum by name through the host classes AND the device table: builds a one-spectrum
// scene, commits it and evaluates on the device, so that flatten + upload + device
// evaluation are all covered.
int qzh_eval_spectrum(const char* name_c, int n, const float* lambdas, float* out) {
    std::string name = name_c;
    std::shared_ptr<const Spectrum> sp;
    float r, g, b;
    if (name == "X") sp = spectra::X();
    else if (name == "Y") sp = spectra::Y();
    else if (name == "Z") sp = spectra::Z();
    else if (name == "D65") sp = spectra::ILLUM_D65();
    else if (name == "CANON_R") sp = spectra::CANON_EOS_R();
    else if (name == "CANON_G") sp = spectra::CANON_EOS_G();
    else if (name == "CANON_B") sp = spectra::CANON_EOS_B();
    else if (name == "AL_IOR") sp = spectra::AL_IOR();
    else if (name == "AL_ABSORPTION") sp = spectra::AL_ABSORPTION();
    else if (name == "CU_IOR") sp = spectra::CU_IOR();
    else if (name == "CU_ABSORPTION") sp = spectra::CU_ABSORPTION();
    else if (name == "GLASS_BK7_IOR") sp = spectra::GLASS_BK7_IOR();
    else if (name == "GLASS_SF11_IOR") sp = spectra::GLASS_SF11_IOR();
    else if (sscanf(name_c, "rgb:%f,%f,%f", &r, &g, &b) == 3)
        sp = std::make_shared<RGBSigmoidPolynomial>(RGBColorSpace::sRGB()->to_spectrum(RGB(r, g, b)));
    else if (sscanf(name_c, "rgbu:%f,%f,%f", &r, &g, &b) == 3) sp = std::make_shared<RGBUnboundedSpectrum>(RGB(r, g, b));
    else if (sscanf(name_c, "rgbi:%f,%f,%f", &r, &g, &b) == 3) sp = std::make_shared<RGBIlluminantSpectrum>(RGB(r, g, b));
    else if (sscanf(name_c, "const:%f", &r) == 1) sp = std::make_shared<ConstantSpectrum>(r);
    else if (sscanf(name_c, "blackbody:%f", &r) == 1) sp = std::make_shared<BlackbodySpectrum>(r);
    if (!sp) return 1;
    Scene scene(initialize_device());
    scene.set_bg_light(sp, 1.0f);
    scene.commit();
    if (!scene.ready()) return 2;
    qzhost::Flattener probe;  // the background spectrum is flattened last: its id is the table size - 1
    int32_t id = sp->flatten(probe);
    (void)id;
    return qz_eval_spectrum(scene.handle(), -1 /* the background spectrum */, uint32_t(n), lambdas, out);
}

int qzh_camera_fields(void* h, float* out) {
    const Camera& c = *static_cast<qzscenes::Bundle*>(h)->camera;
    const Vec3* f[7] = {&c.pos, &c.look_at, &c.up, &c.right, &c.viewport_bottom_left, &c.pixel_delta_u, &c.pixel_delta_v};
    for (int i = 0; i < 7; i++) { out[3 * i] = f[i]->x; out[3 * i + 1] = f[i]->y; out[3 * i + 2] = f[i]->z; }
    return 0;
}

int qzh_sensor_eval(void* h, int n, const float* in, float* out) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    std::vector<float> sensor;
    qz_camera cam = qzhost::flatten_camera(*b->camera, sensor);
    return qz_sensor_eval(&cam, uint32_t(n), in, out);
}

int qzh_intersect(void* h, int n, const float* rays, float* out) {
    auto* b = static_cast<qzscenes::Bundle*>(h);
    return qz_intersect(b->scene->handle(), uint32_t(n), rays, out);
}

void qzh_force_brute_force(int) {}

// Image::save's tone path (image.cpp:7-19) through the C ABI
int qzh_tone(const float* rgb, int n_pixels, float gamma, float* bgr255, unsigned char* bgr8) {
    return qz_tone(rgb, (uint32_t)n_pixels, gamma, bgr255, bgr8);
}

// Image::save (image.cpp:7-19) of an H x W x 3 float plane: "*.png" (what the reference's examples write), "*.ppm" or PFM
int qzh_image_save(const float* rgb, int height, int width, const char* filename, float gamma) {
    Image image(std::vector<float>(rgb, rgb + (size_t)height * width * 3), (size_t)height, (size_t)width);
    image.save(filename, gamma);
    return 0;
}

// ObjData of a file (obj/obj.hpp), flattened for the loader tests: counts = {vertices, normals, faces}; arrays may be null
// (sizing call) and otherwise hold 4 floats per vertex (x y z w), 3 per normal, 13 ints per face (vertices[4], textures[4],
// normals[4], n_vertices).  Returns 1 when the loader returns nullopt.
int qzh_obj_load(const char* path, long* counts, float* vertices, float* normals, int* faces) {
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
