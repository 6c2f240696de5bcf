// Launch interface between the translation units of libqz_b200.so.  Plain data only: the two shading units
// (k_shade_exact.cu, k_shade_fast.cu) compile the same headers in different ARITHMETIC MODES (common.cuh) and under
// different namespace names, so nothing typed by those headers crosses this boundary -- the structs travel as
// `const void*` and are layout-identical in every unit (same headers, same compiler).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qzl {

struct Stage {
    const void* scene;    // qz::DScene
    const void* cam;      // qz::DCamera
    const void* bufs;     // qz::WfBuffers of the pipeline
    const void* pass;     // qz::PassParams
    uint32_t flags;       // QZ_FLAG_*
    uint32_t max_bounces;
    int lean_blocks, shade_blocks, trav_blocks, bin_blocks;
    cudaStream_t stream;
};

// k_trace.cu (the reference's arithmetic in both modes)
void trace_shadow(const Stage& s, bool count);
void trace_closest(const Stage& s, bool count);
// k_stage.cu (integer work)
void bin(const Stage& s);
void sample(const Stage& s);
void film(const Stage& s, float* acc, bool first_pass, bool last_pass, uint32_t n_samples_total, float* color, float* normal, float* albedo, int blocks);

void memo_fill(const void* sampler_table /* qz::SamplerDim */, const void* sampler_params /* qz::SamplerParams */, const void* memo /* qz::SampleMemo */,
               int blocks, cudaStream_t stream);
void tone(const float* rgb, uint32_t n_pixels, float gamma, float* bgr255, uint8_t* bgr8, int blocks, cudaStream_t stream);

// one set per arithmetic mode
#define QZL_MODE_API                                                                                          \
    void generate(const Stage& s, uint32_t first_id, uint32_t n);                                           \
    void memo_spectra(const Stage& s);                                                                      \
    void albedo(const Stage& s);                                                                            \
    void shade(const Stage& s, int family /* 0 misc, 1 diffuse, 2 conductor, 3 dielectric */);              \
    void finish(const Stage& s);                                                                            \
    void step_flat(const Stage& s);                                                                         \
    void trace_paths(const void* scene, const void* cam, const void* sampler_params, uint32_t max_bounces, uint32_t n, \
                     const int32_t* xys, float* records);
namespace exact { QZL_MODE_API }
namespace fast { QZL_MODE_API }

}  // namespace qzl
