// Wavefront path-tracing pipeline: the reference's per-pixel loop (render.cpp:247-319
// render_pixels, :91-212 sample_pixel) re-cut into stages connected by per-slot tags and index queues.
// This header holds the data layout shared by every stage; the kernels live in
//   wf_trace.cuh   k_trace_lane<ANY_HIT>: persistent, phase-scheduled BVH traversal (closest hit: scene.cpp:61-117 /
//                  rtcIntersect1; any hit: scene.cpp:136-143)
//   wf_stage.cuh   k_bin (ordered compaction of the family tags into one queue per family x {first hit, later}),
//                  k_sample (this bounce's Owen-scrambled Halton draws), k_film (ordered per-pixel sum, render.cpp:264-294)
//   wf_shade.cuh   k_generate (camera ray + wavelengths, render.cpp:268-273), k_albedo_conductor (depth-0 albedo of
//                  conductors, render.cpp:150-170), k_shade<family> (emission + MIS, BSDF, light sampling, BSDF
//                  sampling, throughput, Russian roulette), k_finish (PixelSensor::to_sensor_rgb of finished paths,
//                  sensor.cpp:57-70, then a new path in the freed slot), and k_step_flat: shadow test + finish +
//                  regenerate + closest hit in ONE pass over the slots for scenes small enough to need no BVH.
//
// One wavefront iteration of a pipeline (host loop: qz_b200.cu render_impl):
//   BVH scenes : k_trace_lane<any hit> -> k_finish -> k_trace_lane<closest> -> k_bin -> k_sample -> k_albedo_conductor
//                -> k_shade x 4 families                                                          (10 launches)
//   flat scenes: k_step_flat -> k_bin -> k_sample -> k_albedo_conductor -> k_shade x 4             (8 launches)
//
// Path state lives in two 128-byte records per slot (WfBuffers below), every field a 16-byte element.  The stages
// communicate through ONE BYTE PER SLOT AND STAGE (`stage`, `fam`, `post`); only the shading stage needs index queues
// (it is the one stage whose work differs per material family), and those are built IN SLOT ORDER by k_bin -- an
// ordered compaction, one block per 2048 consecutive slots, one atomic per block and queue -- so consecutive lanes of
// a shading kernel touch neighbouring slots.  The intersection and finish stages simply walk the slots densely and
// look at the tag.  Russian roulette stays fused at the end of k_shade: it needs the freshly updated throughput.
//
// DETERMINISM: a path's arithmetic depends only on (x, y, s); queue order varies from run to run but no result
// depends on it.  Every pixel-sample writes its sensor RGB into its own result cell and k_film adds the cells of a
// pixel in ascending s, exactly the reference's summation order, so the film is bit-stable and independent of the
// pool size, the pass size, the number of pipelines and the number of GPUs.
#pragma once

#include "shading.cuh"

namespace qz {

// shade queues: material family x {later bounce, first hit}
enum ShadeQueue { SQ_MISC = 0, SQ_DIFFUSE = 1, SQ_CONDUCTOR = 2, SQ_DIELECTRIC = 3, SQ_FAMILIES = 4, SQ_COUNT = 8 };

// per-slot stage tag (what the next closest-hit stage does with the slot)
// Bits: live | first hit (depth 0) | LATE = the next bounce's draws are not in the path's memo row (sampler.cuh: it
// starts beyond the row's dimensions, or the render has no memo), so they come from the sampler stage.  The trace stage
// passes that on as QZ_FAM_LATE in the family tag and k_bin lists such slots in the late queue as well, so k_sample
// visits only them.
enum StageTag : uint8_t { ST_EMPTY = 0, ST_TRACE = 1, ST_FIRST = 2, ST_TRACE_FIRST = 3, ST_LATE = 4 };
#define QZ_FAM_LATE 0x10u    /* family tag bit: the bounce needs the sampler stage */
#define QZ_FAM_NONE 0xffu   /* family tag of a slot that was not traced this iteration */
// post-shade tag bits: consumed (and cleared) by the shadow / finish stage of the next iteration
#define QZ_POST_SHADOW 1u
#define QZ_POST_DONE 2u
#define QZ_FLAT_MAX_PRIMS 96  /* scenes this small are intersected without a BVH (k_step_flat) */

// counter block layout (uint32 words), one block per pipeline
enum Counter {
    C_ACTIVE = 0,                         // paths shaded in the last completed iteration (termination test)
    C_SHADE0 = 2,                         // .. C_SHADE0 + SQ_COUNT - 1: queue lengths of this iteration
    C_LATE = 10,                          // length of the late queue (bounces that need the sampler stage)
    C_CURSOR_TRACE = 12, C_CURSOR_SHADOW = 13,   // work cursors of the persistent traversal kernels
    C_NEXT_PATH = 14,                     // (pipeline 0's block only) next path id of the pass to hand out
    C_WORK0 = 16,                         // .. C_WORK0 + SQ_COUNT - 1: queue lengths of the PREVIOUS iteration = this iteration's work list
    C_WORDS = 32
};
// 64-bit statistics block (shared by all pipelines, atomics)
enum Stat { S_RAYS_CLOSEST = 0, S_RAYS_SHADOW = 1, S_SHADE = 2, S_NODES = 3, S_PRIMS = 4, S_PATHS_DONE = 5, S_OVERFLOW = 6, S_WORDS = 8 };

// A field of the per-slot path record: 16-byte elements, QZ_REC_BYTES apart.  Records stream through the SMs once
// per stage, so they are read and written with the cache-global (L2 only) forms: explicit global-space accesses that
// leave L1 to the scene tables.
#define QZ_REC_BYTES 128
template <class T>
struct RecField {
    char* base;
    __device__ __forceinline__ T* at(uint32_t slot) const { return reinterpret_cast<T*>(base + (size_t)slot * QZ_REC_BYTES); }
    __device__ __forceinline__ T get(uint32_t slot) const { return __ldcg(at(slot)); }
    // through L1: for a line the same thread prefetched into L1 a trip earlier (k_shade); a slot is read by one thread
    // per kernel and L1 starts every kernel empty, so the line cannot be stale
    __device__ __forceinline__ T get_l1(uint32_t slot) const { return __ldca(at(slot)); }
    __device__ __forceinline__ void set(uint32_t slot, const T& v) const { __stcg(at(slot), v); }
};

// PATH STATE LAYOUT.  A slot owns two 128-byte records (two arrays of cache lines), every field a
// 16-byte element at a fixed offset:
//   hot line :  ray_o | ray_d | weight | lambda | hit_a | hit_b | misc | lpdf
//   side line:  sh_o  | sh_d  | sh_c   | radiance | aov_n | aov_a | samples (8 floats)
// The shade queue of one material family holds a THIRD of the slots, in slot order but with gaps: with one array
// per field (the first layout of this pipeline) every 16-byte access of a lane sat alone in its 32-byte sector and
// 64-byte DRAM burst, nine different DRAM pages per path.  With the record layout a bounce reads ONE full line and
// writes back three of its four sectors; what a kernel does not need it does not touch, at sector granularity.
struct WfBuffers {
    RecField<float4> ray_o, ray_d;      // o.xyz | ior_scale ; d.xyz | p_b
    RecField<float4> hit_a, hit_b;      // t, u, v, primID ; Ng.xyz, geomID (0xffffffff = miss)
    RecField<float4> weight, radiance, lambda, lpdf;
    RecField<uint4> misc;               // path id, halton index, dim | depth << 16, flags | rays issued << 8
    RecField<float4> aov_n, aov_a;
    RecField<float4> sh_o, sh_d, sh_c;
    RecField<float4> samples;           // R_COUNT floats per slot (two elements): this bounce's draws, written by k_sample
    uint8_t *stage, *fam, *post;  // per-slot tags (StageTag, family queue id or QZ_FAM_NONE, QZ_POST_* bits)
    uint32_t* q_shade[SQ_COUNT];
    uint32_t* q_late;           // slots of this iteration whose bounce needs k_sample (QZ_FAM_LATE)
    uint32_t* counters;         // this pipeline's counter block
    uint32_t* next_path;        // next path id of the pass to hand out (shared by all pipelines)
    unsigned long long* stats;  // shared by all pipelines (atomics)
    float4 *res_a, *res_b;      // per pixel-sample: (color rgb, normal.x), (albedo rgb, normal.y)
    float* res_c;               // normal.z
    uint32_t pool;
};

struct PassParams {
    uint32_t n_pix;             // owned pixels
    uint32_t s_begin, s_count;  // sample indices of this pass
    uint32_t total;             // n_pix * s_count
    uint32_t max_bounces;
    const uint32_t* owned_rows; // film rows owned by this call, ascending
    uint32_t width, height;
    SamplerParams spar;
};

#if defined(__CUDACC__)
// WORK LIST of the intersection / finish stages.  Everything those stages have to do in an iteration was left by the
// shading stage of the previous one (a shadow request, a finished path whose slot takes the next pixel-sample, a
// continued ray), so their work list IS the previous iteration's shade queues, one after the other -- compact, in slot
// order within a family, and already there.  (The first version walked every slot of the pool and looked at its tag:
// in the tail of a pass, when a few long paths are left in a pool of millions, a warp found 2 live slots among its 32
// and the stage cost as much as with a full pool -- profiles/r02_summary.md.)  k_generate seeds the list with the
// slots it filled; k_sample's bookkeeping thread publishes the lengths (C_WORK0) after k_bin has built the queues.
struct WorkList {
    uint32_t start[SQ_COUNT + 1];
    __device__ __forceinline__ void load(const uint32_t* counters) {
        start[0] = 0;
#pragma unroll
        for (int q = 0; q < SQ_COUNT; q++) start[q + 1] = start[q] + counters[C_WORK0 + q];
    }
    __device__ __forceinline__ uint32_t total() const { return start[SQ_COUNT]; }
    __device__ __forceinline__ uint32_t slot(uint32_t* const* queues, uint32_t i) const {
        int q = 0;
#pragma unroll
        for (int k = 1; k < SQ_COUNT; k++) q += (i >= start[k]) ? 1 : 0;
        return __ldcg(queues[q] + (i - start[q]));
    }
};

// start fetching the 128-byte line that holds p
__device__ __forceinline__ void prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_line_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
__device__ __forceinline__ float4 f4(const Spec4& s) { return make_float4(s.v[0], s.v[1], s.v[2], s.v[3]); }
__device__ __forceinline__ Spec4 s4(const float4& f) { return spec4(f.x, f.y, f.z, f.w); }

// misc.z = dim | depth << 16 (16 bits each); misc.w = flags | rays issued << 8
__device__ __forceinline__ uint4 pack_misc(uint32_t path_id, const PathState& ps) {
    return make_uint4(path_id, ps.smp.index, ps.smp.dim | (ps.depth << 16), (ps.flags & 0xffu) | (ps.n_rays << 8));
}
__device__ __forceinline__ void unpack_misc(const uint4& m, PathState& ps) {
    ps.smp.index = m.y;
    ps.smp.dim = m.z & 0xffffu;
    ps.depth = m.z >> 16;
    ps.flags = m.w & 0xffu;
    ps.n_rays = m.w >> 8;
}

// sum over the warp's active lanes, one atomic: statistics that a grid-stride kernel accumulated in a register
__device__ __forceinline__ void stat_add(unsigned long long* counter, uint32_t v) {
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(full, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(counter, (unsigned long long)v);
}
#endif

}  // namespace qz
