// Traversal stage (wf_trace.cuh) behind the launch interface.
#include "launch.h"
#include "wf_trace.cuh"

namespace qzl {
using namespace qz;

void trace_shadow(const Stage& s, bool count) {
    const DScene& sc = *static_cast<const DScene*>(s.scene);
    const WfBuffers& b = *static_cast<const WfBuffers*>(s.bufs);
    if (count) k_trace_lane<true, true><<<s.trav_blocks, 128, 0, s.stream>>>(sc, b, s.flags);
    else k_trace_lane<true, false><<<s.trav_blocks, 128, 0, s.stream>>>(sc, b, s.flags);
}
void trace_closest(const Stage& s, bool count) {
    const DScene& sc = *static_cast<const DScene*>(s.scene);
    const WfBuffers& b = *static_cast<const WfBuffers*>(s.bufs);
    if (count) k_trace_lane<false, true><<<s.trav_blocks, 128, 0, s.stream>>>(sc, b, s.flags);
    else k_trace_lane<false, false><<<s.trav_blocks, 128, 0, s.stream>>>(sc, b, s.flags);
}
}  // namespace qzl
