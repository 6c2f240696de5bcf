// Queue builder, sampler stage and film (wf_stage.cuh) behind the launch interface.
#include "launch.h"
#include "wf_stage.cuh"

namespace qzl {
using namespace qz;

void bin(const Stage& s) {
    k_bin<SQ_COUNT><<<s.bin_blocks, 256, 0, s.stream>>>(*static_cast<const WfBuffers*>(s.bufs));
}
void sample(const Stage& s) {
    k_sample<<<s.lean_blocks * 2, 256, 0, s.stream>>>(*static_cast<const DScene*>(s.scene), *static_cast<const WfBuffers*>(s.bufs));
}
void film(const Stage& s, float* acc, bool first_pass, bool last_pass, uint32_t n_samples_total, float* color, float* normal,
          float* albedo, int blocks) {
    k_film<<<blocks, 256, 0, s.stream>>>(*static_cast<const WfBuffers*>(s.bufs), *static_cast<const PassParams*>(s.pass), acc, first_pass,
                                         last_pass, n_samples_total, color, normal, albedo);
}
void memo_fill(const void* sampler_table, const void* sampler_params, const void* memo, int blocks, cudaStream_t stream) {
    k_memo_fill<<<blocks, 256, 0, stream>>>(static_cast<const SamplerDim*>(sampler_table), *static_cast<const SamplerParams*>(sampler_params),
                                           *static_cast<const SampleMemo*>(memo));
}
void tone(const float* rgb, uint32_t n_pixels, float gamma, float* bgr255, uint8_t* bgr8, int blocks, cudaStream_t stream) {
    k_tone<<<blocks, 256, 0, stream>>>(rgb, n_pixels, gamma, bgr255, bgr8);
}
}  // namespace qzl
