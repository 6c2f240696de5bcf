// TEST INFRASTRUCTURE ONLY -- never built into, shipped with or loaded by the product.
//
// Host emulation of the CUDA library's C ABI (include/qz_b200.h): the device headers of
// quetzalcoatlus_b200/csrc are compiled here with plain g++ (QZ_HD = inline) and every
// kernel body is run in a sequential loop.  Purpose: let the CPU-only test-suite (no GPU in
// the development container) check the device-code restatement -- sampler, spectra, BxDFs,
// intersection, the LBVH build and traversal, the path loop -- against the oracle, bit for
// bit except for nothing: in this build the transcendental functions are the host libm's.
// What it cannot cover is the wavefront scheduling, queues and film kernels of
// csrc/wf_*.cuh; those are covered by the -m gpu tests.
//
// The product never routes through this file: bench.py, __graft_entry__.smoke() and the
// quetzalcoatlus_b200 package load libqz_b200.so only, which fails loudly without a GPU.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "memo_plan.h"
#include "scene_store.cuh"

using namespace qz;

namespace {

struct SeqExec {
    template <class T> T* alloc(size_t n) { return static_cast<T*>(std::malloc((n ? n : 1) * sizeof(T))); }
    void free(void* p) { std::free(p); }
    template <class T> void upload(T* dst, const T* src, size_t n) { std::memcpy(dst, src, n * sizeof(T)); }
    template <class T> void download(T* dst, const T* src, size_t n) { std::memcpy(dst, src, n * sizeof(T)); }
    void zero(void* p, size_t bytes) { std::memset(p, 0, bytes); }
    template <class F> void parallel_for(uint32_t n, F f) { for (uint32_t i = 0; i < n; i++) f(i); }
    void sort_u64(uint64_t* keys, uint32_t n) { std::sort(keys, keys + n); }
    void mark(const char*) {}
};

thread_local std::string g_error;
SeqExec g_exec;
// records followed by the prefix tables (sampler.cuh); the emulation tabulates a few dimensions
// with a small cap -- enough to run the prefixed evaluation path through every CPU parity test
std::vector<SamplerDim> g_sampler_storage;
SamplerDim* g_sampler_table = nullptr;
float g_rho_tab[QZ_RHO_TAB_FLOATS];
bool g_tables_ready = false;

void ensure_tables() {
    if (g_tables_ready) return;
    std::vector<SamplerDim> recs(QZ_N_PRIMES);
    build_sampler_table(recs.data());
    const uint32_t n_prefix = plan_sampler_prefix(recs.data(), 40, 4096);
    g_sampler_storage.assign(QZ_N_PRIMES + (n_prefix * sizeof(uint16_t) + sizeof(SamplerDim) - 1) / sizeof(SamplerDim) + 1, SamplerDim{});
    std::copy(recs.begin(), recs.end(), g_sampler_storage.begin());
    g_sampler_table = g_sampler_storage.data();
    uint16_t* prefix = reinterpret_cast<uint16_t*>(g_sampler_table + QZ_N_PRIMES);
    for (uint32_t d = 0; d < QZ_N_PRIMES; d++)
        for (uint32_t j = 0; j < recs[d].pre_pow; j++) prefix[recs[d].pre_offset + j] = sampler_prefix_entry(recs[d], j);
    build_rho_table(g_rho_tab);
    g_tables_ready = true;
}

DCamera make_camera(const qz_camera* c) {
    DCamera d;
    d.width = c->image_width; d.height = c->image_height;
    d.pos = v3(c->pos[0], c->pos[1], c->pos[2]);
    d.bottom_left = v3(c->viewport_bottom_left[0], c->viewport_bottom_left[1], c->viewport_bottom_left[2]);
    d.du = v3(c->pixel_delta_u[0], c->pixel_delta_u[1], c->pixel_delta_u[2]);
    d.dv = v3(c->pixel_delta_v[0], c->pixel_delta_v[1], c->pixel_delta_v[2]);
    d.sensor = c->sensor_rgb;
    d.imaging_ratio = c->imaging_ratio;
    return d;
}

}  // namespace

struct qz_scene_t {
    SceneStore<SeqExec> store;
};

extern "C" {

const char* qz_last_error(void) { return g_error.c_str(); }
uint32_t qz_set_default_flags(uint32_t) { return 0; }
int qz_abi_version(void) { return QZ_ABI_VERSION; }
int qz_init(int) { ensure_tables(); return QZ_OK; }
int qz_device_name(char* buf, size_t n) { std::snprintf(buf, n, "host emulation (tests only)"); return QZ_OK; }

int qz_scene_create(qz_scene* out) { ensure_tables(); *out = new qz_scene_t(); return QZ_OK; }
int qz_scene_destroy(qz_scene s) { if (s) { s->store.release(g_exec); delete s; } return QZ_OK; }

int qz_scene_commit(qz_scene s, const qz_scene_tables* t) {
    std::string why = SceneStore<SeqExec>::validate(*t);
    if (!why.empty()) { g_error = why; return QZ_ERR_INVALID; }
    if (!s->store.commit(g_exec, *t, g_sampler_table, g_rho_tab)) { g_error = "BVH build failed"; return QZ_ERR_INVALID; }
    return QZ_OK;
}

// same rule as the product's probes (csrc/qz_b200.cu): the index is defined while (s + 1) * stride < 2^31 (sampler.cpp:419)
static bool sample_numbers_defined(const int32_t* q, uint32_t n, int words, const SamplerParams& spar) {
    for (uint32_t i = 0; i < n; i++) {
        const int64_t s = q[(size_t)i * words + 2];
        if (s < 0 || (uint64_t)(s + 1) * spar.stride >= (1ull << 31)) return false;
    }
    return true;
}

int qz_trace_paths(qz_scene s, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces, uint32_t n,
                   const int32_t* xys, float* records) {
    (void)n_samples;
    if (!s->store.committed) { g_error = "scene not committed"; return QZ_ERR_NOT_COMMITTED; }
    DCamera cam = make_camera(camera);
    SamplerParams spar = make_sampler_params((int)cam.width, (int)cam.height);
    if (!sample_numbers_defined(xys, n, 3, spar)) { g_error = "sample number negative or too large for the 32-bit Halton index"; return QZ_ERR_INVALID; }
    for (uint32_t i = 0; i < n; i++) {
        PathState ps; PathAov aov; Spec4 lambda0;
        run_path<false>(s->store.view, cam, spar, (uint32_t)xys[3 * i], (uint32_t)xys[3 * i + 1], (uint32_t)xys[3 * i + 2],
                        max_bounces, ps, aov, lambda0, nullptr);
        V3 rgb = to_sensor_rgb(cam, ps.L, ps.lambda, ps.pdf);
        V3 argb = to_sensor_rgb(cam, aov.albedo, ps.lambda, ps.pdf);
        write_trace_record(records + (size_t)i * 32, ps, aov, lambda0, rgb, argb);
    }
    return QZ_OK;
}

// The SAMPLE MEMO of a render (csrc/sampler.cuh) built on the host from the product's own pieces: the class tables and
// the row layout of csrc/memo_plan.h, every entry by memo_entry_bits (what k_memo_fill stores), the hot-spectra block by
// memo_fill_hot (what k_memo_spectra runs).  One pass over all sample numbers [s_begin, s_begin + s_count).
struct HostMemo {
    std::vector<uint32_t> rank, idx, tab;
    SampleMemo m{};
    uint64_t filled = 0;
    void build(const qz_scene_t& s, const SamplerParams& spar, uint32_t W, uint32_t H, const std::vector<uint32_t>& rows, uint32_t bounces,
               uint32_t s_begin, uint32_t s_count) {
        memo_owned_classes(spar, W, H, rows, rank, idx);
        const MemoLayout lay = memo_layout(bounces);
        tab.assign((size_t)lay.stride * s_count * idx.size(), QZ_MEMO_EMPTY);
        m = memo_describe(lay, spar, tab.data(), rank.data(), idx.data(), (uint32_t)idx.size(), s_begin, s_count);
        for (uint32_t sn = 0; sn < s_count; sn++)
            for (uint32_t c = 0; c < m.n_cls; c++)
                for (uint32_t d = 0; d < m.dims; d++, filled++)
                    tab[((size_t)sn * m.n_cls + c) * m.stride + m.dim_off + d] = memo_entry_bits(g_sampler_table, spar, m, sn, c, d);
        const uint32_t n_hot = (uint32_t)s.store.hot_spectra.size();
        if (n_hot) {
            DScene fill = s.store.view;
            fill.memo = m;
            fill.memo.n_hot = n_hot;
            for (uint32_t k = 0; k < n_hot; k++) fill.memo.hot_id[k] = m.hot_id[k] = s.store.hot_spectra[k];
            for (size_t r = 0; r < (size_t)s_count * m.n_cls; r++) memo_fill_hot(fill, tab.data() + r * m.stride);
            m.n_hot = n_hot;
        }
        // negative control of the tests: a table whose entries are off in mantissa bit 12 (5e-4 relative) must change the film, or the paths
        // are not reading it.  QZ_EMU_MEMO_CORRUPT = "dims" | "hot".
        if (const char* e = std::getenv("QZ_EMU_MEMO_CORRUPT")) {
            const bool hot = std::strcmp(e, "hot") == 0;
            for (size_t r = 0; r < (size_t)s_count * m.n_cls; r++) {
                uint32_t* row = tab.data() + r * m.stride;
                if (hot) for (uint32_t k = 0; k < 4 * m.n_hot; k++) row[k] ^= 0x1000u;
                else for (uint32_t d = 0; d < m.dims; d++) row[m.dim_off + d] ^= 0x1000u;
            }
        }
    }
};

// per-pixel loop in the reference's order (render.cpp:260-294); no wavefront here.  With QZ_FLAG_FORCE_MEMO the paths
// read their draws and hot spectra from a sample memo, as the wavefront's do (options->reserved: bounces per row, 0 = 8;
// options->samples_per_pass: sample numbers the table covers, 0 = all -- the rest evaluate directly, like a later pass).
int qz_render(qz_scene s, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces, const qz_region* region,
              const qz_render_options* options, float* color, float* normal, float* albedo, qz_stats* stats) {
    if (!s->store.committed) { g_error = "scene not committed"; return QZ_ERR_NOT_COMMITTED; }
    DCamera cam = make_camera(camera);
    SamplerParams spar = make_sampler_params((int)cam.width, (int)cam.height);
    qz_stats st{};
    HostMemo memo;
    struct Restore {   // the scene view carries the memo only for the duration of this call
        DScene& v;
        ~Restore() { v.memo = SampleMemo{}; }
    } restore{s->store.view};
    if (options && (options->flags & QZ_FLAG_FORCE_MEMO)) {
        std::vector<uint32_t> rows;
        for (uint32_t row = 0; row < cam.height; row++)
            if (!(region && region->strip_rows && region->n_shards > 1 && (row / region->strip_rows) % region->n_shards != region->shard)) rows.push_back(row);
        const uint32_t covered = options->samples_per_pass ? std::min(options->samples_per_pass, n_samples) : n_samples;
        memo.build(*s, spar, cam.width, cam.height, rows, options->reserved ? options->reserved : 8u, 0, covered);
        s->store.view.memo = memo.m;
        st.shade_calls = memo.filled;   // (reported so that a test can see the table was really built)
        st.iterations = memo.idx.size();
    }
    for (uint32_t row = 0; row < cam.height; row++) {
        if (region && region->strip_rows && region->n_shards > 1 && (row / region->strip_rows) % region->n_shards != region->shard) continue;
        for (uint32_t x = 0; x < cam.width; x++) {
            uint32_t y = cam.height - row - 1;
            V3 c = v3(0, 0, 0), nn = v3(0, 0, 0), a = v3(0, 0, 0);
            for (uint32_t sidx = 0; sidx < n_samples; sidx++) {
                PathState ps; PathAov aov; Spec4 lambda0;
                run_path<false>(s->store.view, cam, spar, x, y, sidx, max_bounces, ps, aov, lambda0, nullptr);
                c = c + to_sensor_rgb(cam, ps.L, ps.lambda, ps.pdf);
                a = a + to_sensor_rgb(cam, aov.albedo, ps.lambda, ps.pdf);
                nn = nn + aov.normal;
                st.paths++; st.rays_closest += ps.n_rays;
            }
            c = c / (float)n_samples; a = a / (float)n_samples; nn = nn / (float)n_samples;
            size_t i = ((size_t)row * cam.width + x) * 3;
            color[i] = c.x; color[i + 1] = c.y; color[i + 2] = c.z;
            if (normal) { normal[i] = nn.x; normal[i + 1] = nn.y; normal[i + 2] = nn.z; }
            if (albedo) { albedo[i] = a.x; albedo[i + 1] = a.y; albedo[i + 2] = a.z; }
        }
    }
    if (stats) *stats = st;
    return QZ_OK;
}

int qz_render_device(qz_scene, const qz_camera*, uint32_t, uint32_t, const qz_region*, const qz_render_options*, float*,
                     float*, float*, void*, qz_stats*) {
    g_error = "host emulation has no device path";
    return QZ_ERR_NO_DEVICE;
}

int qz_sampler_eval(uint32_t, uint32_t width, uint32_t height, uint32_t n, const int32_t* q, float* out) {
    ensure_tables();
    SamplerParams spar = make_sampler_params((int)width, (int)height);
    if (!sample_numbers_defined(q, n, 4, spar)) { g_error = "sample number negative or too large for the 32-bit Halton index"; return QZ_ERR_INVALID; }
    for (uint32_t i = 0; i < n; i++) {
        Sampler smp = sampler_start(spar, (uint32_t)q[4 * i], (uint32_t)q[4 * i + 1], (uint32_t)q[4 * i + 2]);
        int dim = q[4 * i + 3];
        if (dim < 2) {
            V2 j = sampler_pixel_jitter(spar, smp);
            out[i] = dim == 0 ? j.x : j.y;
        } else {
            smp.dim = (uint32_t)dim;
            out[i] = sample_1d(g_sampler_table, smp);
        }
    }
    return QZ_OK;
}

int qz_intersect(qz_scene s, uint32_t n, const float* rays, float* out) {
    if (!s->store.committed) { g_error = "scene not committed"; return QZ_ERR_NOT_COMMITTED; }
    const DScene& sc = s->store.view;
    for (uint32_t i = 0; i < n; i++) {
        Ray r;
        r.o = v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
        r.d = v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
        Hit h;
        float* o = out + 8 * i;
        if (!closest_hit<false>(sc, r, h, nullptr)) {
            o[0] = -1.0f;
            for (int k = 1; k < 8; k++) o[k] = 0.0f;
        } else {
            const F4* rec = sc.prims + (size_t)h.prim * 4;
            o[0] = h.t; o[1] = h.u; o[2] = h.v; o[3] = h.ng.x; o[4] = h.ng.y; o[5] = h.ng.z;
            o[6] = (float)float_as_u32(rec[0].w);
            o[7] = (float)float_as_u32(rec[1].w);
        }
    }
    return QZ_OK;
}

int qz_eval_spectrum(qz_scene s, int32_t id, uint32_t n, const float* lambdas, float* out) {
    if (!s->store.committed) { g_error = "scene not committed"; return QZ_ERR_NOT_COMMITTED; }
    const DScene& sc = s->store.view;
    if (id < 0) id = sc.bg_spectrum;
    if (id < 0) { g_error = "no such spectrum"; return QZ_ERR_INVALID; }
    for (uint32_t i = 0; i < n; i++) out[i] = eval_spectrum(sc, id, lambdas[i]);
    return QZ_OK;
}

int qz_math_probe(int op, uint32_t n, const float* in, float* out) {
    if (op != 0) return QZ_ERR_INVALID;
    for (uint32_t i = 0; i < n; i++) qz_sincosf(in[i], out[2 * i], out[2 * i + 1]);
    return QZ_OK;
}

int qz_set_device_count(int) { return 0; }
uint32_t qz_strip_rows(uint32_t, uint32_t) { return 1; }   // (no multi-device render in the emulation)
int qz_scene_build_ms(qz_scene, float* ms) { if (ms) *ms = 0.0f; return QZ_OK; }
int qz_film_device(qz_scene, float**, float**, float**, uint32_t*, uint32_t*) { g_error = "host emulation: no device film"; return QZ_ERR_NO_DEVICE; }
int qz_tone_device(const float*, uint32_t, float, float*, uint8_t*, void*) { g_error = "host emulation: no device"; return QZ_ERR_NO_DEVICE; }
int qz_tone(const float* rgb, uint32_t n_pixels, float gamma, float* bgr255, uint8_t* bgr8) {
    for (uint64_t i = 0; i < (uint64_t)n_pixels * 3; i++) {
        const float v = tone_value(rgb[3 * (i / 3) + (2 - i % 3)], gamma);
        if (bgr255) bgr255[i] = v;
        if (bgr8) bgr8[i] = tone_u8(v);
    }
    return QZ_OK;
}

int qz_sensor_eval(const qz_camera* camera, uint32_t n, const float* in, float* out) {
    DCamera cam = make_camera(camera);
    for (uint32_t i = 0; i < n; i++) {
        Spec4 lambda, pdf;
        sample_wavelengths(in[5 * i], lambda, pdf);
        V3 rgb = to_sensor_rgb(cam, spec4(in[5 * i + 1], in[5 * i + 2], in[5 * i + 3], in[5 * i + 4]), lambda, pdf);
        out[3 * i] = rgb.x; out[3 * i + 1] = rgb.y; out[3 * i + 2] = rgb.z;
    }
    return QZ_OK;
}

}  // extern "C"
