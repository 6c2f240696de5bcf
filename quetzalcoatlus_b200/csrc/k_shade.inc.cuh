// Body of the two shading translation units (k_shade_exact.cu / k_shade_fast.cu): kernels of wf_shade.cuh in the
// unit's arithmetic mode, behind the plain-data launch interface of launch.h.
#include "launch.h"
#include "wf_shade.cuh"

namespace qzl {
namespace QZL_MODE {

using namespace qz;

#define QZL_UNPACK                                                         \
    const DScene& sc = *static_cast<const DScene*>(s.scene);               \
    const WfBuffers& b = *static_cast<const WfBuffers*>(s.bufs);           \
    (void)sc; (void)b;

void generate(const Stage& s, uint32_t first_id, uint32_t n) {
    QZL_UNPACK
    k_generate<<<s.lean_blocks, 256, 0, s.stream>>>(sc, *static_cast<const DCamera*>(s.cam), b, *static_cast<const PassParams*>(s.pass), first_id, n);
}
void memo_spectra(const Stage& s) {
    QZL_UNPACK
    k_memo_spectra<<<s.lean_blocks, 256, 0, s.stream>>>(sc);
}
void albedo(const Stage& s) {
    QZL_UNPACK
    k_albedo_conductor<<<s.lean_blocks * 2, 256, 0, s.stream>>>(sc, b, s.max_bounces);
}
void shade(const Stage& s, int family) {
    QZL_UNPACK
    switch (family) {
        case 0: k_shade<KH_ANY><<<s.shade_blocks, 128, 0, s.stream>>>(sc, b, s.max_bounces); break;
        case 1: k_shade<KH_DIFFUSE><<<s.shade_blocks, 128, 0, s.stream>>>(sc, b, s.max_bounces); break;
        case 2: k_shade<KH_CONDUCTOR><<<s.shade_blocks, 128, 0, s.stream>>>(sc, b, s.max_bounces); break;
        default: k_shade<KH_DIELECTRIC><<<s.shade_blocks, 128, 0, s.stream>>>(sc, b, s.max_bounces); break;
    }
}
void finish(const Stage& s) {
    QZL_UNPACK
    k_finish<<<s.lean_blocks, 256, 0, s.stream>>>(sc, *static_cast<const DCamera*>(s.cam), b, *static_cast<const PassParams*>(s.pass));
}
void step_flat(const Stage& s) {
    QZL_UNPACK
    k_step_flat<<<s.lean_blocks * 2, 128, 0, s.stream>>>(sc, *static_cast<const DCamera*>(s.cam), b, *static_cast<const PassParams*>(s.pass), s.flags);
}
void trace_paths(const void* scene, const void* cam, const void* sampler_params, uint32_t max_bounces, uint32_t n, const int32_t* xys,
                 float* records) {
    k_trace_paths<<<(n + 63) / 64, 64>>>(*static_cast<const DScene*>(scene), *static_cast<const DCamera*>(cam),
                                         *static_cast<const SamplerParams*>(sampler_params), max_bounces, n, xys, records);
}

}  // namespace QZL_MODE
}  // namespace qzl
