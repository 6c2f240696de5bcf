// C ABI of the B200 path-tracing hot path (include/qz_b200.h) and the host-side driver of
// the wavefront pipeline.  One translation unit, compiled for sm_100a only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false --extended-lambda -lineinfo
// -fmad=false is part of the arithmetic contract (common.cuh): with contraction off the
// float32 path is bit-identical to the reference's x86-64 build.
//
// No CPU fallback exists in this library: every entry point that computes needs a CUDA
// device and fails with QZ_ERR_NO_DEVICE / QZ_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "launch.h"
#include "memo_plan.h"
#include "scene_store.cuh"
#include "wf_types.cuh"

using namespace qz;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define QZ_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_error = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? QZ_ERR_NO_DEVICE \
                   : (e_ == cudaErrorMemoryAllocation ? QZ_ERR_OOM : QZ_ERR_CUDA);                 \
        }                                                                                          \
    } while (0)

template <class F>
__global__ void k_parallel_for(uint32_t n, F f) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) f(i);
}

// Executor of the shared build / commit code (bvh_build.cuh, scene_store.cuh) on the GPU.
// Errors are sticky: the first failing CUDA call is remembered and reported by commit.
struct CudaExec {
    cudaError_t err = cudaSuccess;
    void note(cudaError_t e) { if (err == cudaSuccess && e != cudaSuccess) err = e; }

    template <class T> T* alloc(size_t n) {
        void* p = nullptr;
        note(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
        return static_cast<T*>(p);
    }
    void free(void* p) { if (p) cudaFree(p); }
    template <class T> void upload(T* dst, const T* src, size_t n) { if (n) note(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyHostToDevice)); }
    template <class T> void download(T* dst, const T* src, size_t n) { if (n) note(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost)); }
    void zero(void* p, size_t bytes) { if (bytes) note(cudaMemset(p, 0, bytes)); }
    // QZ_BUILD_TRACE=1: device-synchronised wall time of every build step on stderr
    std::chrono::steady_clock::time_point t_mark = std::chrono::steady_clock::now();
    void mark(const char* what) {
        static const bool on = [] { const char* e = std::getenv("QZ_BUILD_TRACE"); return e && std::atoi(e) != 0; }();
        if (!on) return;
        cudaDeviceSynchronize();
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[qz build] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_mark).count());
        t_mark = now;
    }
    template <class F> void parallel_for(uint32_t n, F f) {
        if (!n || err != cudaSuccess) return;
        uint32_t blocks = std::min<uint32_t>((n + 255u) / 256u, 148u * 16u);
        k_parallel_for<<<blocks, 256>>>(n, f);
        note(cudaGetLastError());
    }
    void sort_u64(uint64_t* keys, uint32_t n) {
        if (err != cudaSuccess) return;
        uint64_t* tmp_keys = alloc<uint64_t>(n);
        size_t bytes = 0;
        note(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys, tmp_keys, (int)n));
        void* scratch = nullptr;
        note(cudaMalloc(&scratch, bytes ? bytes : 1));
        note(cub::DeviceRadixSort::SortKeys(scratch, bytes, keys, tmp_keys, (int)n));
        note(cudaMemcpy(keys, tmp_keys, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice));
        cudaFree(scratch);
        cudaFree(tmp_keys);
    }
};

// entry e of the concatenated prefix tables: find its dimension (dim_start is ascending), fill it
__global__ void k_build_sampler_prefix(SamplerDim* table, const uint32_t* __restrict__ dim_start, uint32_t n) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    // last dimension d with dim_start[d] <= e and a table of its own
    uint32_t lo = 0, hi = QZ_N_PRIMES;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (dim_start[mid] <= e) lo = mid; else hi = mid;
    }
    const SamplerDim rec = table[lo];
    uint16_t* prefix = reinterpret_cast<uint16_t*>(table + QZ_N_PRIMES);
    prefix[e] = sampler_prefix_entry(rec, e - rec.pre_offset);
}

// per-device constant tables (sampler dimensions, depth-0 albedo sample warps)
struct DeviceTables {
    SamplerDim* sampler = nullptr;
    float* rho = nullptr;
};
std::mutex g_tables_mutex;
DeviceTables g_tables[64];

int device_tables(int dev, DeviceTables& out) {
    std::lock_guard<std::mutex> lock(g_tables_mutex);
    if (dev < 0 || dev >= 64) return fail(QZ_ERR_INVALID, "device ordinal out of range");
    DeviceTables& t = g_tables[dev];
    if (!t.sampler) {
        std::vector<SamplerDim> host(QZ_N_PRIMES);
        build_sampler_table(host.data());
        float rho[QZ_RHO_TAB_FLOATS];
        build_rho_table(rho);
        // records, then the prefix tables of every dimension (sampler.cuh), filled by a kernel
        const uint32_t n_prefix = plan_sampler_prefix(host.data(), QZ_N_PRIMES, QZ_PREFIX_CAP);
        std::vector<uint32_t> dim_start(QZ_N_PRIMES + 1);
        // dim_start[d] = first entry of the first dimension >= d that has a table (ascending, for the kernel's search)
        {
            uint32_t next = n_prefix;
            for (int d = QZ_N_PRIMES - 1; d >= 0; d--) { if (host[d].pre_pow) next = host[d].pre_offset; dim_start[d] = next; }
            dim_start[QZ_N_PRIMES] = n_prefix;
        }
        // built into locals and published only when every step has succeeded: a half-built entry would make every
        // later call return QZ_OK with a null table
        SamplerDim* d_sampler = nullptr;
        uint32_t* d_start = nullptr;
        float* d_rho = nullptr;
        auto build = [&]() -> cudaError_t {
            cudaError_t e;
            if ((e = cudaMalloc(&d_sampler, host.size() * sizeof(SamplerDim) + (size_t)n_prefix * sizeof(uint16_t) + 16)) != cudaSuccess) return e;
            if ((e = cudaMemcpy(d_sampler, host.data(), host.size() * sizeof(SamplerDim), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&d_start, dim_start.size() * sizeof(uint32_t))) != cudaSuccess) return e;
            if ((e = cudaMemcpy(d_start, dim_start.data(), dim_start.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
            k_build_sampler_prefix<<<(n_prefix + 255) / 256, 256>>>(d_sampler, d_start, n_prefix);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            if ((e = cudaDeviceSynchronize()) != cudaSuccess) return e;
            if ((e = cudaMalloc(&d_rho, sizeof(rho))) != cudaSuccess) return e;
            return cudaMemcpy(d_rho, rho, sizeof(rho), cudaMemcpyHostToDevice);
        };
        const cudaError_t e = build();
        if (d_start) cudaFree(d_start);
        if (e != cudaSuccess) {
            if (d_sampler) cudaFree(d_sampler);
            if (d_rho) cudaFree(d_rho);
            QZ_CUDA(e);
        }
        t.sampler = d_sampler;
        t.rho = d_rho;
    }
    out = t;
    return QZ_OK;
}

thread_local int g_device = -1;
thread_local uint32_t g_default_flags = 0;

bool env_exact() { const char* e = std::getenv("QZ_EXACT"); return e && std::atoi(e) != 0; }

int ensure_device() {
    if (g_device >= 0) {
        QZ_CUDA(cudaSetDevice(g_device));
        return QZ_OK;
    }
    return qz_init(0);
}

DCamera make_camera(const qz_camera* c, const float* d_sensor) {
    DCamera d;
    d.width = c->image_width; d.height = c->image_height;
    d.pos = v3(c->pos[0], c->pos[1], c->pos[2]);
    d.bottom_left = v3(c->viewport_bottom_left[0], c->viewport_bottom_left[1], c->viewport_bottom_left[2]);
    d.du = v3(c->pixel_delta_u[0], c->pixel_delta_u[1], c->pixel_delta_u[2]);
    d.dv = v3(c->pixel_delta_v[0], c->pixel_delta_v[1], c->pixel_delta_v[2]);
    d.sensor = d_sensor;
    d.imaging_ratio = c->imaging_ratio;
    return d;
}

// ------------------------------------------------------------------ small kernels of the probes
__global__ void k_sampler_eval(const SamplerDim* table, SamplerParams spar, uint32_t n, const int32_t* q, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sampler smp = sampler_start(spar, (uint32_t)q[4 * i], (uint32_t)q[4 * i + 1], (uint32_t)q[4 * i + 2]);
    const int dim = q[4 * i + 3];
    if (dim < 2) {
        V2 j = sampler_pixel_jitter(spar, smp);
        out[i] = dim == 0 ? j.x : j.y;
    } else {
        smp.dim = (uint32_t)dim;
        out[i] = sample_1d(table, smp);
    }
}

__global__ void k_intersect(DScene sc, uint32_t n, const float* rays, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r;
    r.o = v3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]);
    r.d = v3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]);
    Hit h;
    float* o = out + 8 * (size_t)i;
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

__global__ void k_math_probe(int op, uint32_t n, const float* in, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op == 0) qz_sincosf(in[i], out[2 * i], out[2 * i + 1]);
}

__global__ void k_eval_spectrum(DScene sc, int32_t id, uint32_t n, const float* lambdas, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = eval_spectrum(sc, id, lambdas[i]);
}

__global__ void k_sensor_eval(DCamera cam, uint32_t n, const float* in, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Spec4 lambda, pdf;
    sample_wavelengths(in[5 * i], lambda, pdf);
    V3 rgb = to_sensor_rgb(cam, spec4(in[5 * i + 1], in[5 * i + 2], in[5 * i + 3], in[5 * i + 4]), lambda, pdf);
    out[3 * i] = rgb.x; out[3 * i + 1] = rgb.y; out[3 * i + 2] = rgb.z;
}

// RAII device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n ? n : 1;
        return cudaMalloc(&p, bytes);
    }
    // grow-only variant for buffers that live across calls
    cudaError_t reserve(size_t n) { return (p && bytes >= n) ? cudaSuccess : alloc(n); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

}  // namespace

// Working memory of render(): allocated on first use, kept with the scene and re-used by later
// calls of the same or a smaller size (cudaMalloc/cudaFree of gigabytes per call would otherwise
// dominate the end-to-end time of short renders).
#define QZ_MAX_PIPELINES 4

struct WorkMem {
    DevBuf rec_hot, rec_side, qbufs[8], q_late, tags[3], counters, statsb, res_a, res_b, res_c, rowsb, sensor, acc, memo, memo_rank, memo_idx;
    uint32_t pool = 0;
    uint64_t cells = 0, acc_pix = 0, rows = 0;
    std::vector<uint32_t> cls_rank, cls_idx;   // pixel classes of the last render's image size and region (sample memo)
    uint64_t cls_key[5] = {};
    bool cls_uploaded = false;
    uint32_t* h_counters = nullptr;  // pinned
    DevBuf film[3];                  // device film planes of the host-buffer entry (qz_render)
    bool film_valid[3] = {false, false, false};
    uint32_t film_w = 0, film_h = 0;
    float* h_film = nullptr;         // pinned staging for pageable caller buffers (three planes)
    cudaEvent_t copy_done[3] = {};
    size_t h_film_bytes = 0;
    std::vector<cudaEvent_t> stage_events;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_a = nullptr, ev_b = nullptr;
    cudaStream_t pipe_stream[QZ_MAX_PIPELINES] = {};
    cudaEvent_t pipe_done[QZ_MAX_PIPELINES] = {};
    cudaEvent_t fork = nullptr;
    ~WorkMem() {
        if (h_counters) cudaFreeHost(h_counters);
        if (h_film) cudaFreeHost(h_film);
        for (cudaEvent_t e : {ev_begin, ev_end, ev_a, ev_b}) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : stage_events) cudaEventDestroy(e);
        for (cudaStream_t q : pipe_stream) if (q) cudaStreamDestroy(q);
        for (cudaEvent_t e : pipe_done) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : copy_done) if (e) cudaEventDestroy(e);
        if (fork) cudaEventDestroy(fork);
    }
};

struct qz_scene_t {
    SceneStore<CudaExec> store;
    int device = 0;
    DeviceTables tables;
    WorkMem work;
    float build_ms = 0.0f;   // table upload + BVH build of the last commit
    // MULTI-GPU behind render(): with QZ_DEVICES = N > 1 the handle owns N - 1 replicas of the committed scene on the
    // following device ordinals (tables uploaded and BVH built on each); qz_render() then renders interleaved row
    // strips on all N devices at once, one host thread per device (the reference fans render() out over host threads,
    // render.cpp:339-382).
    std::vector<qz_scene_t*> replicas;
};

namespace {
thread_local int g_device_count = 0;   // qz_set_device_count(); 0 = take QZ_DEVICES from the environment
int env_devices() {
    if (g_device_count > 0) return g_device_count;
    static const int n = [] { const char* e = std::getenv("QZ_DEVICES"); int v = e ? std::atoi(e) : 1; return v < 1 ? 1 : v; }();
    return n;
}
void destroy_replicas(qz_scene_t* s) {
    for (qz_scene_t* r : s->replicas) {
        cudaSetDevice(r->device);
        CudaExec ex;
        r->store.release(ex);
        delete r;
    }
    s->replicas.clear();
    cudaSetDevice(s->device);
}
// rows per interleaved strip: a height <= 8 that gives every device the same number of rows, powers of two first (with
// strip * n dividing 128 a device owns 1/n of the (y mod 128) pixel classes: its sample memo tabulates only those)
uint32_t balanced_strip_rows(uint32_t height, uint32_t n) {
    const uint32_t order[8] = {8, 4, 2, 1, 7, 6, 5, 3};
    for (uint32_t rows : order)
        if (height % rows == 0 && (height / rows) % n == 0) return rows;
    return 1;
}
}  // namespace

extern "C" {

const char* qz_last_error(void) { return g_error.c_str(); }
uint32_t qz_strip_rows(uint32_t height, uint32_t n_shards) { return balanced_strip_rows(height, n_shards ? n_shards : 1u); }
uint32_t qz_set_default_flags(uint32_t flags) { const uint32_t old = g_default_flags; g_default_flags = flags; return old; }
int qz_abi_version(void) { return QZ_ABI_VERSION; }

int qz_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(QZ_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                          "); this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(QZ_ERR_INVALID, "device ordinal out of range");
    QZ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QZ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(QZ_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not a Blackwell (sm_100a) GPU; this library is built for sm_100a only");
    g_device = device;
    DeviceTables t;
    return device_tables(device, t);
}

int qz_device_name(char* buf, size_t n) {
    int rc = ensure_device();
    if (rc) return rc;
    cudaDeviceProp prop;
    QZ_CUDA(cudaGetDeviceProperties(&prop, g_device));
    std::snprintf(buf, n, "%s", prop.name);
    return QZ_OK;
}

int qz_scene_create(qz_scene* out) {
    if (!out) return fail(QZ_ERR_INVALID, "null output pointer");
    int rc = ensure_device();
    if (rc) return rc;
    qz_scene s = new qz_scene_t();
    s->device = g_device;
    rc = device_tables(g_device, s->tables);
    if (rc) { delete s; return rc; }
    *out = s;
    return QZ_OK;
}

int qz_scene_destroy(qz_scene s) {
    if (!s) return QZ_OK;
    destroy_replicas(s);
    cudaSetDevice(s->device);
    CudaExec ex;
    s->store.release(ex);
    delete s;
    return QZ_OK;
}

int qz_scene_commit(qz_scene s, const qz_scene_tables* t) {
    if (!s || !t) return fail(QZ_ERR_INVALID, "null argument");
    QZ_CUDA(cudaSetDevice(s->device));
    std::string why = SceneStore<CudaExec>::validate(*t);
    if (!why.empty()) return fail(QZ_ERR_INVALID, "invalid scene tables: " + why);
    CudaExec ex;
    // (the build is bracketed on the host around a synchronised device: uploads of the tables are included, the host-side
    // flattening is not)
    cudaDeviceSynchronize();
    const auto t0 = std::chrono::steady_clock::now();
    bool ok = s->store.commit(ex, *t, s->tables.sampler, s->tables.rho);
    ex.note(cudaDeviceSynchronize());
    s->build_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!ok || ex.err != cudaSuccess) {
        s->store.committed = false;
        return fail(ex.err == cudaErrorMemoryAllocation ? QZ_ERR_OOM : QZ_ERR_CUDA,
                    std::string("scene commit failed: ") + cudaGetErrorString(ex.err));
    }
    // replicas for the in-library multi-GPU render (QZ_DEVICES)
    destroy_replicas(s);
    const int n_dev = env_devices();
    if (n_dev > 1) {
        int count = 0;
        QZ_CUDA(cudaGetDeviceCount(&count));
        if (s->device + n_dev > count)
            return fail(QZ_ERR_INVALID, "QZ_DEVICES asks for more devices than are visible from device " + std::to_string(s->device));
        for (int d = 1; d < n_dev; d++) {
            const int dev = s->device + d;
            int can = 0;
            QZ_CUDA(cudaDeviceCanAccessPeer(&can, dev, s->device));
            if (!can) { destroy_replicas(s); return fail(QZ_ERR_CUDA, "QZ_DEVICES: device " + std::to_string(dev) + " has no peer access to device " + std::to_string(s->device) + " (the film is written over NVLink peer memory)"); }
            QZ_CUDA(cudaSetDevice(dev));
            cudaError_t pe = cudaDeviceEnablePeerAccess(s->device, 0);
            if (pe == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); pe = cudaSuccess; }
            QZ_CUDA(pe);
            qz_scene_t* r = new qz_scene_t();
            r->device = dev;
            s->replicas.push_back(r);
            int rc = device_tables(dev, r->tables);
            if (rc) { destroy_replicas(s); return rc; }
            CudaExec rex;
            const bool rok = r->store.commit(rex, *t, r->tables.sampler, r->tables.rho);
            rex.note(cudaDeviceSynchronize());
            if (!rok || rex.err != cudaSuccess) {
                const std::string why_r = cudaGetErrorString(rex.err);
                destroy_replicas(s);
                return fail(rex.err == cudaErrorMemoryAllocation ? QZ_ERR_OOM : QZ_ERR_CUDA, "scene commit failed on device " + std::to_string(dev) + ": " + why_r);
            }
        }
        QZ_CUDA(cudaSetDevice(s->device));
    }
    return QZ_OK;
}

int qz_set_device_count(int n) {
    if (n < 0) return fail(QZ_ERR_INVALID, "negative device count");
    const int old = g_device_count;
    g_device_count = n;
    return old;
}

int qz_scene_build_ms(qz_scene s, float* ms) {
    if (!s || !ms) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    *ms = s->build_ms;
    return QZ_OK;
}

// ------------------------------------------------------------------ render
static int render_impl(qz_scene s, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces, const qz_region* region,
                       const qz_render_options* options, float* d_color, float* d_normal, float* d_albedo,
                       cudaStream_t stream, qz_stats* stats_out) {
    // the kernels receive a per-call copy of the scene view: it carries the pass's sample memo (sampler.cuh)
    DScene sc = s->store.view;
    sc.memo = SampleMemo{};
    const uint32_t W = camera->image_width, H = camera->image_height;
    if (!W || !H || !n_samples) return fail(QZ_ERR_INVALID, "empty image or zero samples");
    if (max_bounces > 65535) return fail(QZ_ERR_INVALID, "max_bounces > 65535 is not supported (the path depth is a 16-bit field)");
    SamplerParams spar = make_sampler_params((int)W, (int)H);
    if ((uint64_t)n_samples * spar.stride >= (1ull << 31)) return fail(QZ_ERR_INVALID, "n_samples too large for the 32-bit Halton index");
    const uint32_t flags = (options ? options->flags : 0u) | g_default_flags;
    if (flags & (QZ_FLAG_LANE_TRAVERSAL | QZ_FLAG_OCTET_TRAVERSAL))
        return fail(QZ_ERR_INVALID, "the first-generation traversal kernels (evidence arms of round 1) are no longer built");

    // rows owned by this call
    std::vector<uint32_t> rows;
    for (uint32_t r = 0; r < H; r++) {
        if (region && region->strip_rows && region->n_shards > 1) {
            if ((r / region->strip_rows) % region->n_shards != region->shard) continue;
        }
        rows.push_back(r);
    }
    qz_stats st{};
    st.bvh_nodes = s->store.bvh.n_nodes;
    st.bvh_bytes = (uint32_t)std::min<uint64_t>(0xffffffffull, (uint64_t)s->store.bvh.n_nodes * sizeof(BvhNode) + (uint64_t)sc.n_prims * 64);
    if (rows.empty()) { if (stats_out) *stats_out = st; return QZ_OK; }
    const uint64_t n_pix64 = (uint64_t)rows.size() * W;
    if (n_pix64 >= (1ull << 31)) return fail(QZ_ERR_INVALID, "image too large");
    const uint32_t n_pix = (uint32_t)n_pix64;

    // pass size: result cells are 36 bytes per pixel-sample; keep a pass under ~6 GB and 2^31 cells
    uint32_t s_pass = options && options->samples_per_pass ? options->samples_per_pass : 0;
    if (!s_pass) {
        uint64_t budget_cells = (6ull << 30) / 36ull;
        s_pass = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_samples, budget_cells / n_pix));
    }
    s_pass = std::min(s_pass, n_samples);
    while ((uint64_t)s_pass * n_pix >= (1ull << 31)) s_pass = std::max(1u, s_pass / 2);
    // SAMPLE MEMO (sampler.cuh): worth its fill when several pixels share a Halton index, i.e. when the call owns several
    // times more pixels than (x mod 128, y mod 128) classes.  Rows: hot spectra, the jitter, the wavelength draw and eight
    // bounces' worth of dimensions (later bounces go through the sampler stage).
    static const int env_memo = [] { const char* e = std::getenv("QZ_MEMO"); return e ? std::atoi(e) : 1; }();
    uint32_t memo_dims = 0, memo_stride = 0, memo_dim_off = 0;
    // (x mod 128, y mod 128) classes the call owns (sampler.cuh, SampleMemo): a function of the image size and the region
    // only, kept with the handle between calls (host: 16 k sampler_start evaluations, ~1.5 ms -- 10 % of a textures frame)
    std::vector<uint32_t>& cls_rank = s->work.cls_rank;
    std::vector<uint32_t>& cls_idx = s->work.cls_idx;
    bool cls_fresh = false;
    if ((env_memo || (flags & QZ_FLAG_FORCE_MEMO)) && !(flags & QZ_FLAG_NO_MEMO)) {
        const uint64_t key[5] = {((uint64_t)W << 32) | H, region ? region->strip_rows : 0u, region ? region->n_shards : 0u, region ? region->shard : 0u, spar.stride};
        if (cls_rank.empty() || std::memcmp(key, s->work.cls_key, sizeof(key)) != 0) {
            memo_owned_classes(spar, W, H, rows, cls_rank, cls_idx);
            std::memcpy(s->work.cls_key, key, sizeof(key));
            cls_fresh = true;
            s->work.cls_uploaded = false;   // (the device copies belong to the previous key until this one is uploaded)
        }
        const uint64_t n_cls = cls_idx.size();
        if ((flags & QZ_FLAG_FORCE_MEMO) || (n_pix64 >= 4 * n_cls && n_pix64 >= 65536)) {
            static const int env_memo_bounces = [] { const char* e = std::getenv("QZ_MEMO_BOUNCES"); int v = e ? std::atoi(e) : 0; return v; }();
            const MemoLayout lay = memo_layout((uint32_t)(env_memo_bounces > 0 ? std::min(env_memo_bounces, 32) : 8));
            memo_dim_off = lay.dim_off, memo_dims = lay.dims, memo_stride = lay.stride;
            // a pass must also fit its rows into the memo's budget (3 GB)
            const uint64_t rows_fit = (3ull << 30) / ((uint64_t)memo_stride * 4 * n_cls);
            if (rows_fit == 0) memo_dims = 0;
            else s_pass = (uint32_t)std::min<uint64_t>(s_pass, rows_fit);
        }
    }
    const uint64_t cells = (uint64_t)s_pass * n_pix;
    // paths in flight: 2^23 for the flat scenes, 2^24 where the BVH is traversed (fewer, fuller iterations: obj_viewer +4 %;
    // cornell_box -1 %)
    const bool flat_scene = sc.n_prims <= QZ_FLAT_MAX_PRIMS && sc.n_prims > 0 && !(flags & (QZ_FLAG_COUNT_TRAVERSAL | QZ_FLAG_FORCE_BVH));
    uint32_t pool = options && options->pool_paths ? options->pool_paths : (flat_scene ? (1u << 23) : (1u << 24));
    pool = (uint32_t)std::min<uint64_t>(pool, cells);
    pool = std::max(pool, 1u);
    const bool count_trav = (flags & QZ_FLAG_COUNT_TRAVERSAL) != 0;
    const bool stage_timing = (flags & QZ_FLAG_STAGE_TIMING) != 0;
    // ARITHMETIC MODE (common.cuh): radiometric values fused / approximate by default; QZ_FLAG_EXACT_ARITHMETIC (or
    // QZ_EXACT=1 in the environment) selects the kernels that keep the reference's arithmetic throughout
    const bool exact = env_exact() || (flags & QZ_FLAG_EXACT_ARITHMETIC) != 0;

    // PIPELINES.  The pool is cut into P independent sub-pools, each with its own queues, tags and
    // counters and its own stream, all drawing new paths from one shared cursor.  Every stage kernel
    // is a persistent grid-stride loop, so it is launched with 1/P of the CTAs an SM can hold: the
    // stages of different pipelines then run SIDE BY SIDE on every SM -- the integer-bound sampler
    // and the slab tests of one pipeline issue in the slots the latency-bound shading kernels of
    // the other leave empty, and one pipeline's tail overlaps the other's next stage.  Stage timing
    // (one event pair per stage) only makes sense serially, so it uses one pipeline.
    static const int env_pipes = [] { const char* e = std::getenv("QZ_PIPELINES"); int v = e ? std::atoi(e) : 0; return v; }();
    int P = env_pipes > 0 ? env_pipes : 3;
    if (stage_timing || pool < 65536u) P = 1;
    P = std::min(P, QZ_MAX_PIPELINES);
    const uint32_t sub_pool = (pool + (uint32_t)P - 1) / (uint32_t)P;
    const uint32_t tag_stride = ((sub_pool + 15u) & ~15u) + 16u;   // per-pipeline slice of the tag arrays (16-byte aligned, padded)
    const uint32_t q_stride = (sub_pool + 3u) & ~3u;

    // ---- working memory (cached in the scene handle)
    WorkMem& wm = s->work;
    DevBuf(&qbufs)[SQ_COUNT] = wm.qbufs;
    DevBuf &counters = wm.counters, &statsb = wm.statsb, &res_a = wm.res_a, &res_b = wm.res_b, &res_c = wm.res_c,
           &rowsb = wm.rowsb, &sensor = wm.sensor, &acc = wm.acc;
    QZ_CUDA(wm.rec_hot.reserve((size_t)sub_pool * P * QZ_REC_BYTES));
    QZ_CUDA(wm.rec_side.reserve((size_t)sub_pool * P * QZ_REC_BYTES));
    for (auto& buf : qbufs) QZ_CUDA(buf.reserve((size_t)q_stride * P * 4));
    QZ_CUDA(wm.q_late.reserve((size_t)q_stride * P * 4));
    for (auto& buf : wm.tags) QZ_CUDA(buf.reserve((size_t)tag_stride * P));
    QZ_CUDA(counters.reserve((size_t)C_WORDS * 4 * QZ_MAX_PIPELINES));
    QZ_CUDA(statsb.reserve(S_WORDS * 8));
    QZ_CUDA(res_a.reserve(cells * 16));
    QZ_CUDA(res_b.reserve(cells * 16));
    QZ_CUDA(res_c.reserve(cells * 4));
    QZ_CUDA(rowsb.reserve(rows.size() * 4));
    QZ_CUDA(sensor.reserve(3 * 471 * 4));
    if (memo_dims) {   // the pass's memo rows and the class tables the kernels address them through
        QZ_CUDA(wm.memo.reserve((size_t)memo_stride * 4 * s_pass * cls_idx.size()));
        const bool realloc = wm.memo_rank.bytes < cls_rank.size() * 4 || wm.memo_idx.bytes < cls_idx.size() * 4 || !wm.memo_rank.p || !wm.memo_idx.p;
        QZ_CUDA(wm.memo_rank.reserve(cls_rank.size() * 4));
        QZ_CUDA(wm.memo_idx.reserve(cls_idx.size() * 4));
        if (cls_fresh || realloc || !wm.cls_uploaded) {
            QZ_CUDA(cudaMemcpyAsync(wm.memo_rank.p, cls_rank.data(), cls_rank.size() * 4, cudaMemcpyHostToDevice, stream));
            QZ_CUDA(cudaMemcpyAsync(wm.memo_idx.p, cls_idx.data(), cls_idx.size() * 4, cudaMemcpyHostToDevice, stream));
            wm.cls_uploaded = true;
        }
    }
    const bool multi_pass = s_pass < n_samples;
    if (multi_pass) QZ_CUDA(acc.reserve((size_t)n_pix * 9 * 4));
    QZ_CUDA(cudaMemcpyAsync(rowsb.p, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice, stream));
    QZ_CUDA(cudaMemcpyAsync(sensor.p, camera->sensor_rgb, 3 * 471 * 4, cudaMemcpyHostToDevice, stream));
    QZ_CUDA(cudaMemsetAsync(statsb.p, 0, S_WORDS * 8, stream));

    static_assert(R_COUNT * sizeof(float) == 32, "the samples of a bounce fill the last two elements of the side line");
    WfBuffers bufs[QZ_MAX_PIPELINES];
    for (int p = 0; p < P; p++) {
        WfBuffers& b = bufs[p];
        b = WfBuffers{};
        char* hot = wm.rec_hot.as<char>() + (size_t)p * sub_pool * QZ_REC_BYTES;
        char* side = wm.rec_side.as<char>() + (size_t)p * sub_pool * QZ_REC_BYTES;
        b.ray_o.base = hot + 0;   b.ray_d.base = hot + 16;  b.weight.base = hot + 32;    b.lambda.base = hot + 48;
        b.hit_a.base = hot + 64;  b.hit_b.base = hot + 80;  b.misc.base = hot + 96;      b.lpdf.base = hot + 112;
        b.sh_o.base = side + 0;   b.sh_d.base = side + 16;  b.sh_c.base = side + 32;     b.radiance.base = side + 48;
        b.aov_n.base = side + 64; b.aov_a.base = side + 80; b.samples.base = side + 96;
        b.stage = wm.tags[0].as<uint8_t>() + (size_t)p * tag_stride;
        b.fam = wm.tags[1].as<uint8_t>() + (size_t)p * tag_stride;
        b.post = wm.tags[2].as<uint8_t>() + (size_t)p * tag_stride;
        for (int k = 0; k < SQ_COUNT; k++) b.q_shade[k] = qbufs[k].as<uint32_t>() + (size_t)p * q_stride;
        b.q_late = wm.q_late.as<uint32_t>() + (size_t)p * q_stride;
        b.counters = counters.as<uint32_t>() + (size_t)p * C_WORDS;
        b.next_path = counters.as<uint32_t>() + C_NEXT_PATH;   // pipeline 0's block holds the shared cursor
        b.stats = statsb.as<unsigned long long>();
        b.res_a = res_a.as<float4>(); b.res_b = res_b.as<float4>(); b.res_c = res_c.as<float>();
        b.pool = sub_pool;
    }

    DCamera cam = make_camera(camera, sensor.as<float>());
    // tiny scenes need no BVH (k_step_flat); counting runs and QZ_FLAG_FORCE_BVH keep the traversal kernels
    const bool flat = sc.n_prims <= QZ_FLAT_MAX_PRIMS && sc.n_prims > 0 && !count_trav && !(flags & QZ_FLAG_FORCE_BVH);

    PassParams pp_cur{};   // the pass being rendered
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    static const int env_per_sm = [] { const char* e = std::getenv("QZ_BLOCKS_PER_SM"); int v = e ? std::atoi(e) : 0; return v; }();
    const int per_sm = env_per_sm > 0 ? env_per_sm : (P == 1 ? 8 : (P == 2 ? 4 : (P == 3 ? 3 : 2)));
    // the lean stages (sampler, flat step, albedo, finish, generate: 32-96 registers, 256-thread CTAs)
    static const int env_lean = [] { const char* e = std::getenv("QZ_LEAN_BLOCKS_PER_SM"); int v = e ? std::atoi(e) : 0; return v; }();

    if (!wm.ev_begin) {
        QZ_CUDA(cudaEventCreate(&wm.ev_begin)); QZ_CUDA(cudaEventCreate(&wm.ev_end));
        QZ_CUDA(cudaEventCreate(&wm.ev_a)); QZ_CUDA(cudaEventCreate(&wm.ev_b));
        QZ_CUDA(cudaMallocHost(&wm.h_counters, (size_t)C_WORDS * 4 * QZ_MAX_PIPELINES));
        for (int p = 0; p < QZ_MAX_PIPELINES; p++) {
            QZ_CUDA(cudaStreamCreateWithFlags(&wm.pipe_stream[p], cudaStreamNonBlocking));
            QZ_CUDA(cudaEventCreateWithFlags(&wm.pipe_done[p], cudaEventDisableTiming));
        }
        QZ_CUDA(cudaEventCreateWithFlags(&wm.fork, cudaEventDisableTiming));
    }
    cudaEvent_t ev_begin = wm.ev_begin, ev_end = wm.ev_end;
    QZ_CUDA(cudaEventRecord(ev_begin, stream));
    // the pipelines run on the handle's own streams, forked from the caller's stream and joined back into it (also a
    // single pipeline: the iteration is captured into a CUDA graph, which the legacy default stream does not allow)
    cudaStream_t ps[QZ_MAX_PIPELINES];
    for (int p = 0; p < P; p++) ps[p] = wm.pipe_stream[p];

    qzl::Stage stg[QZ_MAX_PIPELINES];
    for (int p = 0; p < P; p++) {
        qzl::Stage& g = stg[p];
        g.scene = &sc; g.cam = &cam; g.bufs = &bufs[p]; g.pass = &pp_cur;
        g.flags = flags; g.max_bounces = max_bounces;
        // (the lean stages with fewer CTAs than the others when three pipelines share the SMs: +1 % on cornell_box)
        g.lean_blocks = n_sm * (env_lean > 0 ? env_lean : (P == 3 ? 2 : per_sm));
        g.shade_blocks = n_sm * per_sm;   // persistent CTAs of 4 warps
        g.trav_blocks = n_sm * per_sm;
        g.bin_blocks = (int)std::min<uint32_t>((sub_pool + 2047u) / 2048u, (uint32_t)n_sm * 8u);
        g.stream = ps[p];
    }

    // Stage timing: events are recorded asynchronously between the stages (no host sync inside
    // the pipeline) and read back once per pass, so the figures are device time per stage of the
    // very run that is being timed.
    std::vector<cudaEvent_t>& evpool = wm.stage_events;
    std::vector<float*> ev_target;
    size_t ev_used = 0;
    auto next_event = [&]() -> cudaEvent_t {
        if (ev_used == evpool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            evpool.push_back(e);
        }
        return evpool[ev_used++];
    };
    auto flush_events = [&]() {
        for (int p = 0; p < P; p++) cudaStreamSynchronize(ps[p]);
        cudaStreamSynchronize(stream);
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, evpool[i], evpool[i + 1]);
            *ev_target[i / 2] += ms;
        }
        ev_used = 0;
        ev_target.clear();
    };
    auto timed = [&](cudaStream_t q, float& acc_ms, auto&& launch) {
        if (!stage_timing) { launch(); return; }
        cudaEvent_t a = next_event(), c = next_event();
        cudaEventRecord(a, q);
        launch();
        cudaEventRecord(c, q);
        ev_target.push_back(&acc_ms);
        if (ev_used >= 8192) flush_events();
    };

    // one wavefront iteration of pipeline p (wf_types.cuh): 8 launches for flat scenes, 10 with the BVH
    const int launches_per_iteration = flat ? 8 : 10;
    auto enqueue_iteration = [&](int p) -> cudaError_t {
        const qzl::Stage& g = stg[p];
        cudaStream_t q = ps[p];
        if (flat) {
            timed(q, st.ms_closest, [&] { exact ? qzl::exact::step_flat(g) : qzl::fast::step_flat(g); });
        } else {
            timed(q, st.ms_shadow, [&] { qzl::trace_shadow(g, count_trav); });
            timed(q, st.ms_other, [&] { exact ? qzl::exact::finish(g) : qzl::fast::finish(g); });
            timed(q, st.ms_closest, [&] { qzl::trace_closest(g, count_trav); });
        }
        timed(q, st.ms_other, [&] { qzl::bin(g); });
        timed(q, st.ms_sample, [&] { qzl::sample(g); });
        timed(q, st.ms_shade, [&] {
            exact ? qzl::exact::albedo(g) : qzl::fast::albedo(g);
            for (int fam = 0; fam < 4; fam++) exact ? qzl::exact::shade(g, fam) : qzl::fast::shade(g, fam);
        });
        return cudaGetLastError();
    };

    // The iteration of a pipeline is the same launch sequence with the same arguments for a whole pass: it is
    // captured once per pass into a CUDA graph and replayed (QZ_GRAPH=0 launches kernel by kernel).
    static const bool env_graph = [] { const char* e = std::getenv("QZ_GRAPH"); return !e || std::atoi(e) != 0; }();
    const bool use_graph = env_graph && !stage_timing;
    cudaGraphExec_t gexec[QZ_MAX_PIPELINES] = {};
    auto drop_graphs = [&]() {
        for (int p = 0; p < QZ_MAX_PIPELINES; p++) if (gexec[p]) { cudaGraphExecDestroy(gexec[p]); gexec[p] = nullptr; }
    };
    // on any failure: nothing of this call may still be running on the cached buffers when the caller comes back
    auto quiesce = [&]() {
        for (int p = 0; p < P; p++) cudaStreamSynchronize(ps[p]);
        cudaStreamSynchronize(stream);
        drop_graphs();
    };
#define QZ_RENDER_CUDA(call)                                                                 \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            quiesce();                                                                       \
            g_error = std::string(#call) + ": " + cudaGetErrorString(e_);                    \
            return e_ == cudaErrorMemoryAllocation ? QZ_ERR_OOM : QZ_ERR_CUDA;               \
        }                                                                                    \
    } while (0)

    int rc = QZ_OK;
    for (uint32_t s_begin = 0; s_begin < n_samples && rc == QZ_OK; s_begin += s_pass) {
        PassParams pp{};
        pp.n_pix = n_pix;
        pp.s_begin = s_begin;
        pp.s_count = std::min(s_pass, n_samples - s_begin);
        pp.total = pp.n_pix * pp.s_count;
        pp.max_bounces = max_bounces;
        pp.owned_rows = rowsb.as<uint32_t>();
        pp.width = W; pp.height = H;
        pp.spar = spar;
        pp_cur = pp;
        if (memo_dims) {
            // (n_hot is set after the fill: k_memo_spectra itself must evaluate, not look up)
            sc.memo = memo_describe(MemoLayout{memo_dim_off, memo_dims, memo_stride}, spar, wm.memo.as<uint32_t>(), wm.memo_rank.as<uint32_t>(),
                                    wm.memo_idx.as<uint32_t>(), (uint32_t)cls_idx.size(), s_begin, pp.s_count);
            QZ_RENDER_CUDA(cudaMemsetAsync(sc.memo.tab, 0xff, (size_t)memo_stride * 4 * pp.s_count * cls_idx.size(), stream));
            qzl::memo_fill(sc.sampler_table, &spar, &sc.memo, n_sm * 8, stream);
            st.kernel_launches++;
            const uint32_t n_hot = (uint32_t)s->store.hot_spectra.size();
            if (n_hot) {
                for (uint32_t k = 0; k < n_hot; k++) sc.memo.hot_id[k] = s->store.hot_spectra[k];
                DScene sc_fill = sc;
                sc_fill.memo.n_hot = n_hot;   // the kernel reads the list; from_spectrum() never consults the memo
                qzl::Stage fs = stg[0];
                fs.scene = &sc_fill;
                fs.stream = stream;
                exact ? qzl::exact::memo_spectra(fs) : qzl::fast::memo_spectra(fs);
                st.kernel_launches++;
                sc.memo.n_hot = n_hot;
            }
            QZ_RENDER_CUDA(cudaGetLastError());
        }

        // initial fill: pipeline p starts with the next `first[p]` path ids; the shared cursor continues behind them
        uint32_t init[C_WORDS * QZ_MAX_PIPELINES] = {0};
        uint32_t first[QZ_MAX_PIPELINES], handed = 0;
        for (int p = 0; p < P; p++) { first[p] = std::min<uint32_t>(sub_pool, pp.total - handed); handed += first[p]; }
        init[C_NEXT_PATH] = handed;
        for (int p = 0; p < P; p++) init[(size_t)p * C_WORDS + C_WORK0] = first[p];   // k_generate seeds queue 0 with the slots it fills
        QZ_RENDER_CUDA(cudaMemcpyAsync(counters.p, init, sizeof(uint32_t) * C_WORDS * P, cudaMemcpyHostToDevice, stream));
        QZ_RENDER_CUDA(cudaEventRecord(wm.fork, stream));
        for (int p = 0; p < P; p++) QZ_RENDER_CUDA(cudaStreamWaitEvent(ps[p], wm.fork, 0));
        uint32_t id0 = 0;
        for (int p = 0; p < P; p++) {
            exact ? qzl::exact::generate(stg[p], id0, first[p]) : qzl::fast::generate(stg[p], id0, first[p]);
            id0 += first[p];
            st.kernel_launches++;
        }
        QZ_RENDER_CUDA(cudaGetLastError());
        if (use_graph) {
            drop_graphs();
            for (int p = 0; p < P; p++) {
                if (!first[p]) continue;
                cudaGraph_t graph = nullptr;
                QZ_RENDER_CUDA(cudaStreamBeginCapture(ps[p], cudaStreamCaptureModeThreadLocal));
                const cudaError_t e1 = enqueue_iteration(p);
                const cudaError_t e2 = cudaStreamEndCapture(ps[p], &graph);
                QZ_RENDER_CUDA(e1);
                QZ_RENDER_CUDA(e2);
                const cudaError_t e3 = cudaGraphInstantiate(&gexec[p], graph, 0);
                cudaGraphDestroy(graph);
                QZ_RENDER_CUDA(e3);
            }
        }

        bool live[QZ_MAX_PIPELINES];
        for (int p = 0; p < P; p++) live[p] = first[p] > 0;
        uint64_t it = 0;
        for (;;) {
            bool any = false;
            for (int p = 0; p < P; p++) {
                if (!live[p]) continue;
                any = true;
                if (use_graph) QZ_RENDER_CUDA(cudaGraphLaunch(gexec[p], ps[p]));
                else QZ_RENDER_CUDA(enqueue_iteration(p));
                st.kernel_launches += launches_per_iteration;
            }
            if (!any) break;
            st.iterations++;
            it++;
            // the active counts are read back every 4th iteration (every iteration would serialise host and device);
            // while the host waits for one pipeline the others still have their four iterations queued
            if ((it & 3u) == 0) {
                for (int p = 0; p < P; p++)
                    if (live[p]) QZ_RENDER_CUDA(cudaMemcpyAsync(wm.h_counters + (size_t)p * C_WORDS, bufs[p].counters, C_WORDS * 4, cudaMemcpyDeviceToHost, ps[p]));
                for (int p = 0; p < P; p++) {
                    if (!live[p]) continue;
                    QZ_RENDER_CUDA(cudaStreamSynchronize(ps[p]));
                    if (wm.h_counters[(size_t)p * C_WORDS + C_ACTIVE] == 0) live[p] = false;
                }
            }
            if (it > 1000000ull) { rc = fail(QZ_ERR_CUDA, "wavefront did not terminate"); break; }
        }
        if (rc != QZ_OK) break;
        // (an iteration whose trace stage found no live slot has finished every path: the finish stage runs FIRST in an
        // iteration, on the post tags the previous shading stage left)
        for (int p = 0; p < P; p++) {
            QZ_RENDER_CUDA(cudaEventRecord(wm.pipe_done[p], ps[p]));
            QZ_RENDER_CUDA(cudaStreamWaitEvent(stream, wm.pipe_done[p], 0));
        }
        const bool first_pass = s_begin == 0, last_pass = s_begin + pp.s_count >= n_samples;
        qzl::Stage fs = stg[0];
        fs.stream = stream;
        qzl::film(fs, acc.as<float>(), first_pass, last_pass, n_samples, d_color, d_normal, d_albedo, n_sm * 8);
        st.kernel_launches++;
        QZ_RENDER_CUDA(cudaGetLastError());
    }
    if (rc != QZ_OK) { quiesce(); return rc; }
    QZ_RENDER_CUDA(cudaEventRecord(ev_end, stream));
    QZ_RENDER_CUDA(cudaEventSynchronize(ev_end));
    drop_graphs();
    if (stage_timing) flush_events();
    QZ_CUDA(cudaEventElapsedTime(&st.ms_total, ev_begin, ev_end));
    unsigned long long h_stats[S_WORDS];
    QZ_CUDA(cudaMemcpy(h_stats, statsb.p, sizeof(h_stats), cudaMemcpyDeviceToHost));
    st.paths = (uint64_t)n_pix * n_samples;
    st.rays_closest = h_stats[S_RAYS_CLOSEST];
    st.rays_shadow = h_stats[S_RAYS_SHADOW];
    st.shade_calls = h_stats[S_SHADE];
    st.node_visits = h_stats[S_NODES];
    st.prim_tests = h_stats[S_PRIMS];
    st.stack_overflows = (uint32_t)std::min<unsigned long long>(h_stats[S_OVERFLOW], 0xffffffffull);
    if (st.stack_overflows)
        rc = fail(QZ_ERR_CUDA, "BVH traversal stack overflow: the tree is deeper than the per-lane stack (QZ_STACK entries, bvh.cuh)");
    if (rc == QZ_OK && h_stats[S_PATHS_DONE] != st.paths) rc = fail(QZ_ERR_CUDA, "internal error: finished path count does not match");
    if (stats_out) *stats_out = st;
    return rc;
#undef QZ_RENDER_CUDA
}

int qz_render_device(qz_scene s, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces, const qz_region* region,
                     const qz_render_options* options, float* d_color, float* d_normal, float* d_albedo, void* cuda_stream,
                     qz_stats* stats) {
    if (!s || !camera || !d_color) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    QZ_CUDA(cudaSetDevice(s->device));
    return render_impl(s, camera, n_samples, max_bounces, region, options, d_color, d_normal, d_albedo,
                       static_cast<cudaStream_t>(cuda_stream), stats);
}

int qz_render(qz_scene s, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces, const qz_region* region,
              const qz_render_options* options, float* color, float* normal, float* albedo, qz_stats* stats) {
    if (!s || !camera || !color) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    QZ_CUDA(cudaSetDevice(s->device));
    const size_t n = (size_t)camera->image_width * camera->image_height * 3;
    // device planes and the pinned staging area live with the scene handle: a render() per frame
    // must not pay cudaMalloc / cudaFree / first-touch of tens of megabytes every call
    WorkMem& wm = s->work;
    DevBuf &dc = wm.film[0], &dn = wm.film[1], &da = wm.film[2];
    QZ_CUDA(dc.reserve(n * 4));
    if (normal) QZ_CUDA(dn.reserve(n * 4));
    if (albedo) QZ_CUDA(da.reserve(n * 4));
    float* host[3] = {color, normal, albedo};
    DevBuf* dev[3] = {&dc, &dn, &da};
    // caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pinned
    // memory) are copied to directly; pageable ones go through the pinned staging area
    bool pinned[3] = {false, false, false};
    bool need_staging = false;
    for (int k = 0; k < 3; k++) {
        if (!host[k]) continue;
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, host[k]) == cudaSuccess && attr.type == cudaMemoryTypeHost) pinned[k] = true;
        else { cudaGetLastError(); need_staging = true; }
    }
    // (staging holds all three planes: the device->host copies are queued back to back and each plane is copied out to
    // the caller's pageable buffer, by a few host threads, while the next one is still arriving)
    if (need_staging && wm.h_film_bytes < 3 * n * 4) {
        if (wm.h_film) cudaFreeHost(wm.h_film);
        wm.h_film = nullptr; wm.h_film_bytes = 0;
        QZ_CUDA(cudaMallocHost(&wm.h_film, 3 * n * 4));
        wm.h_film_bytes = 3 * n * 4;
    }
    if (need_staging && !wm.copy_done[0]) {
        for (int k = 0; k < 3; k++) QZ_CUDA(cudaEventCreateWithFlags(&wm.copy_done[k], cudaEventDisableTiming));
    }
    const bool sharded = region && region->strip_rows && region->n_shards > 1;
    if (sharded) {
        // rows this call does not own keep the caller's contents
        for (int k = 0; k < 3; k++)
            if (host[k]) QZ_CUDA(cudaMemcpy(dev[k]->p, host[k], n * 4, cudaMemcpyHostToDevice));
    }
    wm.film_valid[0] = wm.film_valid[1] = wm.film_valid[2] = false;
    int rc = QZ_OK;
    if (!s->replicas.empty() && !sharded) {
        // MULTI-GPU: device d renders the strips (row / strip_rows) % N == d, all samples of its pixels, and its film
        // kernel writes the finished rows STRAIGHT INTO THIS DEVICE'S FILM PLANES through NVLink peer memory -- the
        // film accumulation and the gather are one kernel, there is no separate collective and no host bounce.  Every
        // pixel is owned by exactly one device and keeps its sample order: the film is bit-identical to a 1-GPU render.
        const uint32_t n_dev = 1u + (uint32_t)s->replicas.size();
        const uint32_t strip = balanced_strip_rows(camera->image_height, n_dev);
        std::vector<int> rcs(n_dev, QZ_OK);
        std::vector<std::string> errs(n_dev);
        std::vector<qz_stats> sts(n_dev);
        const uint32_t caller_flags = g_default_flags;
        float* const pc = dc.as<float>();
        float* const pn = normal ? dn.as<float>() : nullptr;
        float* const pa = albedo ? da.as<float>() : nullptr;
        auto work = [&](uint32_t d) {
            qz_scene_t* sd = d == 0 ? s : s->replicas[d - 1];
            qz_region reg{strip, n_dev, d};
            if (cudaSetDevice(sd->device) != cudaSuccess) { rcs[d] = QZ_ERR_CUDA; errs[d] = "cudaSetDevice failed"; return; }
            g_default_flags = caller_flags;
            rcs[d] = render_impl(sd, camera, n_samples, max_bounces, &reg, options, pc, pn, pa, nullptr, &sts[d]);
            if (rcs[d] != QZ_OK) errs[d] = g_error;
        };
        std::vector<std::thread> threads;
        for (uint32_t d = 1; d < n_dev; d++) threads.emplace_back(work, d);
        work(0);
        for (auto& th : threads) th.join();
        QZ_CUDA(cudaSetDevice(s->device));
        for (uint32_t d = 0; d < n_dev; d++)
            if (rcs[d] != QZ_OK) return fail(rcs[d], "device " + std::to_string(d) + ": " + errs[d]);
        if (stats) {
            qz_stats total = sts[0];
            for (uint32_t d = 1; d < n_dev; d++) {
                total.paths += sts[d].paths; total.rays_closest += sts[d].rays_closest; total.rays_shadow += sts[d].rays_shadow;
                total.shade_calls += sts[d].shade_calls; total.kernel_launches += sts[d].kernel_launches;
                total.node_visits += sts[d].node_visits; total.prim_tests += sts[d].prim_tests;
                total.iterations = std::max(total.iterations, sts[d].iterations);
                total.ms_total = std::max(total.ms_total, sts[d].ms_total);
            }
            *stats = total;
        }
    } else {
        rc = render_impl(s, camera, n_samples, max_bounces, region, options, dc.as<float>(), normal ? dn.as<float>() : nullptr,
                         albedo ? da.as<float>() : nullptr, nullptr, stats);
    }
    if (rc != QZ_OK) return rc;
    wm.film_valid[0] = true; wm.film_valid[1] = normal != nullptr; wm.film_valid[2] = albedo != nullptr;
    wm.film_w = camera->image_width; wm.film_h = camera->image_height;
    for (int k = 0; k < 3; k++) {
        if (!host[k]) continue;
        if (pinned[k]) {
            QZ_CUDA(cudaMemcpyAsync(host[k], dev[k]->p, n * 4, cudaMemcpyDeviceToHost, nullptr));
        } else {
            QZ_CUDA(cudaMemcpyAsync(wm.h_film + (size_t)k * n, dev[k]->p, n * 4, cudaMemcpyDeviceToHost, nullptr));
            QZ_CUDA(cudaEventRecord(wm.copy_done[k], nullptr));
        }
    }
    for (int k = 0; k < 3; k++) {
        if (!host[k] || pinned[k]) continue;
        QZ_CUDA(cudaEventSynchronize(wm.copy_done[k]));
        const float* src = wm.h_film + (size_t)k * n;
        float* dst = host[k];
        const int n_thr = n >= (1u << 20) ? 4 : 1;
        const size_t chunk = (n + n_thr - 1) / n_thr;
        std::vector<std::thread> helpers;
        for (int t = 1; t < n_thr; t++) {
            const size_t a = (size_t)t * chunk, e = std::min<size_t>(n, a + chunk);
            if (a < e) helpers.emplace_back([=] { std::memcpy(dst + a, src + a, (e - a) * 4); });
        }
        std::memcpy(dst, src, std::min<size_t>(n, chunk) * 4);
        for (auto& h : helpers) h.join();
    }
    QZ_CUDA(cudaStreamSynchronize(nullptr));
    return QZ_OK;
}

// ------------------------------------------------------------------ output step (image.cpp)
int qz_film_device(qz_scene s, float** d_color, float** d_normal, float** d_albedo, uint32_t* width, uint32_t* height) {
    if (!s) return fail(QZ_ERR_INVALID, "null argument");
    WorkMem& wm = s->work;
    if (!wm.film_valid[0]) return fail(QZ_ERR_INVALID, "no film on the device: qz_render() has not completed on this handle");
    if (d_color) *d_color = wm.film[0].as<float>();
    if (d_normal) *d_normal = wm.film_valid[1] ? wm.film[1].as<float>() : nullptr;
    if (d_albedo) *d_albedo = wm.film_valid[2] ? wm.film[2].as<float>() : nullptr;
    if (width) *width = wm.film_w;
    if (height) *height = wm.film_h;
    return QZ_OK;
}

int qz_tone_device(const float* d_rgb, uint32_t n_pixels, float gamma, float* d_bgr255, uint8_t* d_bgr8, void* cuda_stream) {
    if ((!d_rgb && n_pixels) || (!d_bgr255 && !d_bgr8)) return fail(QZ_ERR_INVALID, "null argument");
    if (n_pixels > 0xffffffffu / 3u) return fail(QZ_ERR_INVALID, "image too large");
    int rc = ensure_device();
    if (rc) return rc;
    if (!n_pixels) return QZ_OK;
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g_device);
    const int blocks = (int)std::min<uint64_t>(((uint64_t)n_pixels * 3 + 255) / 256, (uint64_t)n_sm * 8);
    qzl::tone(d_rgb, n_pixels, gamma, d_bgr255, d_bgr8, blocks, static_cast<cudaStream_t>(cuda_stream));
    QZ_CUDA(cudaGetLastError());
    return QZ_OK;
}

int qz_tone(const float* rgb, uint32_t n_pixels, float gamma, float* bgr255, uint8_t* bgr8) {
    if ((!rgb && n_pixels) || (!bgr255 && !bgr8)) return fail(QZ_ERR_INVALID, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (!n_pixels) return QZ_OK;
    const size_t n = (size_t)n_pixels * 3;
    DevBuf din, dout, d8;
    QZ_CUDA(din.alloc(n * 4));
    if (bgr255) QZ_CUDA(dout.alloc(n * 4));
    if (bgr8) QZ_CUDA(d8.alloc(n));
    QZ_CUDA(cudaMemcpy(din.p, rgb, n * 4, cudaMemcpyHostToDevice));
    rc = qz_tone_device(din.as<float>(), n_pixels, gamma, bgr255 ? dout.as<float>() : nullptr, bgr8 ? d8.as<uint8_t>() : nullptr, nullptr);
    if (rc) return rc;
    if (bgr255) QZ_CUDA(cudaMemcpy(bgr255, dout.p, n * 4, cudaMemcpyDeviceToHost));
    if (bgr8) QZ_CUDA(cudaMemcpy(bgr8, d8.p, n, cudaMemcpyDeviceToHost));
    return QZ_OK;
}

// ------------------------------------------------------------------ probes
// The sampler's index is defined while (s + 1) * stride < 2^31: the reference multiplies two ints (sampler.cpp:419), beyond
// that its result is undefined.  render_impl() refuses such sample counts; the probes refuse such sample numbers.
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
    if (!s || !camera || (!xys && n) || (!records && n)) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    QZ_CUDA(cudaSetDevice(s->device));
    if (!n) return QZ_OK;
    if (!sample_numbers_defined(xys, n, 3, make_sampler_params((int)camera->image_width, (int)camera->image_height)))
        return fail(QZ_ERR_INVALID, "sample number negative or too large for the 32-bit Halton index");
    DevBuf dx, dr, sensor;
    QZ_CUDA(dx.alloc((size_t)n * 12));
    QZ_CUDA(dr.alloc((size_t)n * 128));
    QZ_CUDA(sensor.alloc(3 * 471 * 4));
    QZ_CUDA(cudaMemcpy(dx.p, xys, (size_t)n * 12, cudaMemcpyHostToDevice));
    QZ_CUDA(cudaMemcpy(sensor.p, camera->sensor_rgb, 3 * 471 * 4, cudaMemcpyHostToDevice));
    DCamera cam = make_camera(camera, sensor.as<float>());
    SamplerParams spar = make_sampler_params((int)cam.width, (int)cam.height);
    QZ_CUDA(cudaMemset(s->store.view.overflow, 0, sizeof(uint32_t)));
    if (env_exact() || (g_default_flags & QZ_FLAG_EXACT_ARITHMETIC))
        qzl::exact::trace_paths(&s->store.view, &cam, &spar, max_bounces, n, dx.as<int32_t>(), dr.as<float>());
    else
        qzl::fast::trace_paths(&s->store.view, &cam, &spar, max_bounces, n, dx.as<int32_t>(), dr.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaDeviceSynchronize());
    QZ_CUDA(cudaMemcpy(records, dr.p, (size_t)n * 128, cudaMemcpyDeviceToHost));
    uint32_t overflow = 0;
    QZ_CUDA(cudaMemcpy(&overflow, s->store.view.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (overflow) return fail(QZ_ERR_CUDA, "BVH traversal stack overflow in the per-path replay (QZ_STACK entries, bvh.cuh)");
    return QZ_OK;
}

int qz_sampler_eval(uint32_t n_samples, uint32_t width, uint32_t height, uint32_t n, const int32_t* q, float* out) {
    (void)n_samples;
    int rc = ensure_device();
    if (rc) return rc;
    if (!n) return QZ_OK;
    if (!q || !out) return fail(QZ_ERR_INVALID, "null argument");
    if (!width || !height) return fail(QZ_ERR_INVALID, "empty image");
    if (!sample_numbers_defined(q, n, 4, make_sampler_params((int)width, (int)height)))
        return fail(QZ_ERR_INVALID, "sample number negative or too large for the 32-bit Halton index");
    DeviceTables t;
    rc = device_tables(g_device, t);
    if (rc) return rc;
    DevBuf dq, dout;
    QZ_CUDA(dq.alloc((size_t)n * 16));
    QZ_CUDA(dout.alloc((size_t)n * 4));
    QZ_CUDA(cudaMemcpy(dq.p, q, (size_t)n * 16, cudaMemcpyHostToDevice));
    SamplerParams spar = make_sampler_params((int)width, (int)height);
    k_sampler_eval<<<(n + 127) / 128, 128>>>(t.sampler, spar, n, dq.as<int32_t>(), dout.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return QZ_OK;
}

int qz_intersect(qz_scene s, uint32_t n, const float* rays, float* out) {
    if (!s || (!rays && n) || (!out && n)) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    QZ_CUDA(cudaSetDevice(s->device));
    if (!n) return QZ_OK;
    DevBuf dr, dout;
    QZ_CUDA(dr.alloc((size_t)n * 24));
    QZ_CUDA(dout.alloc((size_t)n * 32));
    QZ_CUDA(cudaMemcpy(dr.p, rays, (size_t)n * 24, cudaMemcpyHostToDevice));
    QZ_CUDA(cudaMemset(s->store.view.overflow, 0, sizeof(uint32_t)));
    k_intersect<<<(n + 127) / 128, 128>>>(s->store.view, n, dr.as<float>(), dout.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
    uint32_t overflow = 0;
    QZ_CUDA(cudaMemcpy(&overflow, s->store.view.overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (overflow) return fail(QZ_ERR_CUDA, "BVH traversal stack overflow in the intersection probe (QZ_STACK entries, bvh.cuh)");
    return QZ_OK;
}

int qz_eval_spectrum(qz_scene s, int32_t id, uint32_t n, const float* lambdas, float* out) {
    if (!s || (!lambdas && n) || (!out && n)) return fail(QZ_ERR_INVALID, "null argument");
    if (!s->store.committed) return fail(QZ_ERR_NOT_COMMITTED, "Scene must be committed before rendering.");
    QZ_CUDA(cudaSetDevice(s->device));
    if (id < 0) id = s->store.view.bg_spectrum;
    if (id < 0) return fail(QZ_ERR_INVALID, "no such spectrum");
    if (!n) return QZ_OK;
    DevBuf dl, dout;
    QZ_CUDA(dl.alloc((size_t)n * 4));
    QZ_CUDA(dout.alloc((size_t)n * 4));
    QZ_CUDA(cudaMemcpy(dl.p, lambdas, (size_t)n * 4, cudaMemcpyHostToDevice));
    k_eval_spectrum<<<(n + 127) / 128, 128>>>(s->store.view, id, n, dl.as<float>(), dout.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return QZ_OK;
}

int qz_math_probe(int op, uint32_t n, const float* in, float* out) {
    if (op != 0 || (!in && n) || (!out && n)) return fail(QZ_ERR_INVALID, "bad argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (!n) return QZ_OK;
    DevBuf din, dout;
    QZ_CUDA(din.alloc((size_t)n * 4));
    QZ_CUDA(dout.alloc((size_t)n * 8));
    QZ_CUDA(cudaMemcpy(din.p, in, (size_t)n * 4, cudaMemcpyHostToDevice));
    k_math_probe<<<(n + 127) / 128, 128>>>(op, n, din.as<float>(), dout.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return QZ_OK;
}

int qz_sensor_eval(const qz_camera* camera, uint32_t n, const float* in, float* out) {
    if (!camera || (!in && n) || (!out && n)) return fail(QZ_ERR_INVALID, "null argument");
    int rc = ensure_device();
    if (rc) return rc;
    if (!n) return QZ_OK;
    DevBuf din, dout, sensor;
    QZ_CUDA(din.alloc((size_t)n * 20));
    QZ_CUDA(dout.alloc((size_t)n * 12));
    QZ_CUDA(sensor.alloc(3 * 471 * 4));
    QZ_CUDA(cudaMemcpy(din.p, in, (size_t)n * 20, cudaMemcpyHostToDevice));
    QZ_CUDA(cudaMemcpy(sensor.p, camera->sensor_rgb, 3 * 471 * 4, cudaMemcpyHostToDevice));
    DCamera cam = make_camera(camera, sensor.as<float>());
    k_sensor_eval<<<(n + 127) / 128, 128>>>(cam, n, din.as<float>(), dout.as<float>());
    QZ_CUDA(cudaGetLastError());
    QZ_CUDA(cudaMemcpy(out, dout.p, (size_t)n * 12, cudaMemcpyDeviceToHost));
    return QZ_OK;
}

}  // extern "C"
