// Wavefront path-tracing pipeline: the reference's per-pixel loop (render.cpp:247-319
// render_pixels, :91-212 sample_pixel) re-cut into stages connected by index queues.
//
//   k_generate          camera ray + wavelengths for new pixel-samples (render.cpp:268-273)
//   k_trace_lane<false> persistent, phase-scheduled BVH traversal, closest hit (scene.cpp:61-117 /
//                       rtcIntersect1); tags each hit with its material family.  k_closest_flat:
//                       BVH-less variant for scenes of <= 96 primitives.  (k_closest_hit, k_trace_oct:
//                       earlier generations, kept behind flags as evidence arms.)
//   k_bin               ordered compaction of the tags into one queue per family x {first hit, later}
//   k_sample            this bounce's Owen-scrambled Halton draws, one thread per dimension
//   k_albedo_conductor  depth-0 albedo of conductors, sixteen lanes per path (render.cpp:150-170)
//   k_shade<family>     emission + MIS, BSDF construction, depth-0 albedo, light sampling,
//                       BSDF sampling, throughput update and Russian roulette for ONE family
//                       (diffuse / conductor / dielectric); a fourth "misc" kernel takes misses,
//                       emitter pass-throughs and MixedMaterial hits with run-time dispatch
//   k_trace_lane<true>  any-hit traversal of the next-event shadow rays (scene.cpp:136-143); adds the
//                       pending contribution when unoccluded (k_shadow_flat for tiny scenes)
//   k_finish            PixelSensor::to_sensor_rgb of finished paths (sensor.cpp:57-70) into the
//                       per-sample result buffer, then regenerates a new path in the freed slot
//   k_film              ordered per-pixel sum over the sample index (render.cpp:264-294)
//
// Path state lives in two 128-byte records per slot (WfBuffers below), every field a 16-byte element;
// the queues carry 4-byte slot indices.  QUEUES ARE BUILT IN SLOT ORDER: the stages do not append to
// queues with atomics (which scatters neighbouring slots across a queue); they write a one-byte tag
// per slot, and k_bin -- an ordered compaction, one block per 2048 consecutive slots, one atomic
// per block and queue -- turns the tags into queues whose entries ascend within each 2048-slot
// chunk.  Consecutive lanes of the shading kernels then touch neighbouring slots, and the
// closest-hit stage simply walks the slots densely.  Russian roulette stays fused at the end of
// k_shade: it needs the freshly updated throughput and the next sampler dimension, so a separate
// kernel would only re-read what is in registers.  The host (qz_b200.cu: render_impl) runs several
// such pipelines side by side, each on a sub-pool with its own queues and stream.
//
// DETERMINISM: a path's arithmetic depends only on (x, y, s); queue order varies from run to
// run but no result depends on it.  Every pixel-sample writes its sensor RGB into its own
// result cell and k_film adds the cells of a pixel in ascending s, exactly the reference's
// summation order, so the film is bit-stable and independent of the pool size, the pass
// size, the number of pipelines and the number of GPUs.
#pragma once

#include <cooperative_groups.h>

#include "shading.cuh"

namespace qz {

namespace cg = cooperative_groups;

// shade queues: material family x {later bounce, first hit}
enum ShadeQueue { SQ_MISC = 0, SQ_DIFFUSE = 1, SQ_CONDUCTOR = 2, SQ_DIELECTRIC = 3, SQ_FAMILIES = 4, SQ_COUNT = 8 };

// per-slot stage tag (what the next closest-hit stage does with the slot)
enum StageTag : uint8_t { ST_EMPTY = 0, ST_TRACE = 1, ST_TRACE_FIRST = 2 };
#define QZ_FAM_NONE 0xffu   /* family tag of a slot that was not traced this iteration */
// post-shade tag bits
#define QZ_POST_SHADOW 1u
#define QZ_POST_DONE 2u
#define QZ_FLAT_MAX_PRIMS 96  /* scenes this small are intersected without a BVH (k_closest_flat) */

// counter block layout (uint32 words)
enum Counter {
    C_ACTIVE = 0,                         // paths shaded in the last completed iteration (termination test)
    C_SHADE0 = 2,                         // .. C_SHADE0 + SQ_COUNT - 1
    C_SHADOW = 10, C_DONE = 11,           // (adjacent: k_bin<2, true> fills both)
    C_CURSOR_TRACE = 12, C_CURSOR_SHADOW = 13,
    C_NEXT_PATH = 14,                     // (pipeline 0's block only) next path id of the pass to hand out
    C_WORDS = 16
};
// 64-bit statistics block
enum Stat { S_RAYS_CLOSEST = 0, S_RAYS_SHADOW = 1, S_SHADE = 2, S_NODES = 3, S_PRIMS = 4, S_PATHS_DONE = 5, S_OVERFLOW = 6, S_WORDS = 8 };

// A field of the per-slot path record: 16-byte elements, QZ_REC_BYTES apart.  b.field[slot]
// reads and writes like a plain array; the stride is what the layout below is about.
#define QZ_REC_BYTES 128
template <class T>
struct RecField {
    char* base;
    __device__ __forceinline__ T& operator[](uint32_t slot) const {
        return *reinterpret_cast<T*>(base + (size_t)slot * QZ_REC_BYTES);
    }
};

// PATH STATE LAYOUT.  A slot owns two 128-byte records (two arrays of cache lines), every field a
// 16-byte element at a fixed offset:
//   hot line :  ray_o | ray_d | weight | lambda | hit_a | hit_b | misc | lpdf
//   side line:  sh_o  | sh_d  | sh_c   | radiance | aov_n | aov_a | samples (8 floats)
// The shade queue of one material family holds a THIRD of the slots, in slot order but with
// gaps: with one array per field (the first layout of this pipeline) every 16-byte access of a
// lane then sat alone in its 32-byte sector and 64-byte DRAM burst, nine different DRAM pages per
// path, and the shading kernels ran at 2.8 TB/s of mostly wasted sectors, insensitive to
// occupancy and instruction count (profiles/r01_shading.md).  With the record layout a bounce reads
// ONE full line and writes back three of its four sectors; what a kernel does not need (the
// closest-hit stage reads 32 bytes and writes 32) it does not touch, at sector granularity.
struct WfBuffers {
    RecField<float4> ray_o, ray_d;      // o.xyz | ior_scale ; d.xyz | p_b
    RecField<float4> hit_a, hit_b;      // t, u, v, primID ; Ng.xyz, geomID (0xffffffff = miss)
    RecField<float4> weight, radiance, lambda, lpdf;
    RecField<uint4> misc;               // path id, halton index, dim | depth << 16 | flags << 24, rays issued
    RecField<float4> aov_n, aov_a;
    RecField<float4> sh_o, sh_d, sh_c;
    RecField<float4> samples;           // R_COUNT floats per slot (two elements): this bounce's draws, written by k_sample
    uint8_t *stage, *fam, *post;  // per-slot tags (StageTag, family queue id or QZ_FAM_NONE, QZ_POST_* bits)
    uint32_t* q_shade[SQ_COUNT];
    uint32_t *q_shadow, *q_done;  // (adjacent in memory order is not required)
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

// start fetching the 128-byte line that holds p
__device__ __forceinline__ void prefetch_line(const void* p) {
#ifndef QZ_PREFETCH_L2
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// ------------------------------------------------------------------ SoA helpers
__device__ __forceinline__ float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
__device__ __forceinline__ float4 f4(const Spec4& s) { return make_float4(s.v[0], s.v[1], s.v[2], s.v[3]); }
__device__ __forceinline__ Spec4 s4(const float4& f) { return spec4(f.x, f.y, f.z, f.w); }

__device__ __forceinline__ void store_state(const WfBuffers& b, uint32_t slot, const PathState& ps, uint32_t path_id) {
    b.ray_o[slot] = f4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ior_scale);
    b.ray_d[slot] = f4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, ps.p_b);
    b.weight[slot] = f4(ps.weight);
    b.radiance[slot] = f4(ps.L);
    b.lpdf[slot] = f4(ps.pdf);
    b.misc[slot] = make_uint4(path_id, ps.smp.index, ps.smp.dim | (ps.depth << 16) | (ps.flags << 24), ps.n_rays);
}

__device__ __forceinline__ uint32_t load_state(const WfBuffers& b, uint32_t slot, PathState& ps) {
    const float4 o = b.ray_o[slot], d = b.ray_d[slot];
    ps.ray.o = v3(o.x, o.y, o.z); ps.ior_scale = o.w;
    ps.ray.d = v3(d.x, d.y, d.z); ps.p_b = d.w;
    ps.weight = s4(b.weight[slot]);
    ps.L = s4(b.radiance[slot]);
    ps.lambda = s4(b.lambda[slot]);
    ps.pdf = s4(b.lpdf[slot]);
    const uint4 m = b.misc[slot];
    ps.smp.index = m.y;
    ps.smp.dim = m.z & 0xffffu;
    ps.depth = (m.z >> 16) & 0xffu;
    ps.flags = m.z >> 24;
    ps.n_rays = m.w;
    return m.x;
}

// append `slot` to a queue; lanes of the warp that push to the same queue share one atomic
__device__ __forceinline__ void queue_push(uint32_t* counter, uint32_t* queue, uint32_t slot) {
    const unsigned peers = __match_any_sync(__activemask(), (unsigned long long)counter);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    queue[base + __popc(peers & ((1u << lane) - 1u))] = slot;
}

// path id of the pass -> pixel and sample; initialises the slot (render.cpp:261-273)
__device__ __forceinline__ void init_slot(const DScene& sc, const DCamera& cam, const WfBuffers& b, const PassParams& pp,
                                          uint32_t slot, uint32_t path_id) {
    const uint32_t pix = path_id % pp.n_pix;
    const uint32_t s = pp.s_begin + path_id / pp.n_pix;
    const uint32_t row = pp.owned_rows[pix / pp.width];
    const uint32_t x = pix % pp.width;
    const uint32_t y = pp.height - row - 1;
    PathState ps;
    PathAov aov;
    start_path(sc, cam, pp.spar, x, y, s, ps, aov);
    store_state(b, slot, ps, path_id);
    b.lambda[slot] = f4(ps.lambda);
    b.aov_n[slot] = f4(0.0f, 0.0f, 0.0f, 0.0f);
    b.aov_a[slot] = f4(0.0f, 0.0f, 0.0f, 0.0f);
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) k_generate(DScene sc, DCamera cam, WfBuffers b, PassParams pp, uint32_t first_id, uint32_t n) {
    for (uint32_t i = n + blockIdx.x * blockDim.x + threadIdx.x; i < b.pool; i += gridDim.x * blockDim.x) b.stage[i] = ST_EMPTY;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        init_slot(sc, cam, b, pp, i, first_id + i);
        b.stage[i] = ST_TRACE_FIRST;
    }
}

// family of the surface a closest-hit result lands on (the branch-sorting key)
__device__ __forceinline__ int classify_hit(const DScene& sc, const Hit& hit, bool unsorted) {
    if (unsorted || hit.geom_id == QZ_NO_HIT) return SQ_MISC;
    const int32_t mat = sc.geoms[hit.geom_id].material;
    if (mat < 0) return SQ_MISC;
    const uint32_t kind = sc.materials[mat].kind;
    if (kind == QZ_MAT_DIFFUSE) return SQ_DIFFUSE;
    if (kind == QZ_MAT_CONDUCTOR) return SQ_CONDUCTOR;
    if (kind == QZ_MAT_DIELECTRIC || kind == QZ_MAT_THIN_DIELECTRIC) return SQ_DIELECTRIC;
    return SQ_MISC;  // MixedMaterial: the family depends on the bounce's material sample
}

// closest-hit epilogue shared by the BVH and the flat kernel: write the hit record and the tag
// of the shade queue the path belongs to (k_bin builds the queues from the tags)
__device__ __forceinline__ void finish_closest(const DScene& sc, const WfBuffers& b, uint32_t slot, bool first, const Hit& h, uint32_t flags) {
    b.hit_a[slot] = f4(h.t, h.u, h.v, __uint_as_float(h.prim_id));
    b.hit_b[slot] = f4(h.ng.x, h.ng.y, h.ng.z, __uint_as_float(h.geom_id));
    const bool unsorted = (flags & QZ_FLAG_UNSORTED_SHADING) != 0;
    const int fam = classify_hit(sc, h, unsorted);
    b.fam[slot] = (uint8_t)(fam + (first && !unsorted ? SQ_FAMILIES : 0));
}

// Ordered compaction of per-slot tags into queues.  EXCLUSIVE: tag == q selects queue q (shade
// families); otherwise bit q of the tag selects queue q (shadow / done: a slot can be in both).
// A block owns 2048 consecutive slots, each thread 8 consecutive ones, so a queue's entries
// ascend within every block's share; one atomicAdd per block and queue reserves the share.
template <int NQ, bool EXCLUSIVE>
__global__ void __launch_bounds__(256) k_bin(const uint8_t* __restrict__ tags, const uint8_t* __restrict__ traced,
                                             uint32_t n_slots, uint32_t* counters, WfBuffers b, int which) {
    __shared__ uint32_t s_warp[8][NQ];
    __shared__ uint32_t s_base[NQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x * 2048u; tile < n_slots; tile += gridDim.x * 2048u) {
        const uint32_t first_slot = tile + threadIdx.x * 8u;
        // this thread's 8 tags; a slot takes part only if it exists and (when `traced` is given) was
        // traced this iteration
        uint32_t tg[8];
        bool ok[8];
        if (first_slot + 8u <= n_slots) {
            const uint2 raw = *reinterpret_cast<const uint2*>(tags + first_slot);
            uint2 tr = make_uint2(0u, 0u);
            if (traced) tr = *reinterpret_cast<const uint2*>(traced + first_slot);
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                tg[k] = ((k < 4 ? raw.x : raw.y) >> (8 * (k & 3))) & 0xffu;
                ok[k] = (((k < 4 ? tr.x : tr.y) >> (8 * (k & 3))) & 0xffu) != QZ_FAM_NONE;
            }
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const bool in_range = first_slot + k < n_slots;
                tg[k] = in_range ? tags[first_slot + k] : QZ_FAM_NONE;
                ok[k] = in_range && (!traced || traced[first_slot + k] != QZ_FAM_NONE);
            }
        }
        // membership masks (bit k = this thread's k-th slot) and per-thread counts
        uint32_t member[NQ], cnt[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            uint32_t m = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const bool in = ok[k] && (EXCLUSIVE ? tg[k] == (uint32_t)q : ((tg[k] >> q) & 1u) != 0u);
                m |= (in ? 1u : 0u) << k;
            }
            member[q] = m;
            cnt[q] = __popc(m);
        }
        // exclusive scan of the counts over the block, per queue
        uint32_t excl[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            uint32_t x = cnt[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            excl[q] = x - cnt[q];
            if (lane == 31) s_warp[warp][q] = x;
        }
        __syncthreads();
        if (threadIdx.x < NQ) {
            uint32_t total = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = total; total += c; }
            s_base[threadIdx.x] = total ? atomicAdd(&counters[threadIdx.x], total) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            uint32_t* queue = which == 0 ? b.q_shade[q] : (q == 0 ? b.q_shadow : b.q_done);
            uint32_t pos = s_base[q] + s_warp[warp][q] + excl[q];
            uint32_t m = member[q];
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                queue[pos++] = first_slot + (uint32_t)k;
            }
        }
        __syncthreads();
    }
}

#ifndef QZ_SHADE_MIN_BLOCKS_LIGHT
#define QZ_SHADE_MIN_BLOCKS_LIGHT 5   /* diffuse / dielectric shade kernels: 96 registers */
#endif
#ifndef QZ_SHADE_MIN_BLOCKS_HEAVY
#define QZ_SHADE_MIN_BLOCKS_HEAVY 4   /* conductor / run-time-dispatch shade kernels */
#endif

#ifndef QZ_REFILL_MIN
#define QZ_REFILL_MIN 8   /* idle lanes that trigger a refill from the queue */
#endif
#ifndef QZ_TRACE_MIN_BLOCKS
#define QZ_TRACE_MIN_BLOCKS 5
#endif
#define QZ_STEPS_PER_ROUND 4

// Persistent closest-hit traversal.  Each warp owns 32 lanes of traversal state; lanes whose
// ray has finished are refilled from the queue cursor as soon as QZ_REFILL_MIN of them are
// idle (warp vote), so a warp keeps working at high lane occupancy on incoherent rays.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_closest_hit(DScene sc, WfBuffers b, uint32_t flags) {
    const uint32_t count = b.pool;  // the slots are walked densely; the stage tag says which are live
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    Trav tv;
    tv.done = true;
    bool active = false;
    bool first = false;
    bool exhausted = false;
    uint32_t slot = 0;
    TraversalCounters cnt;
    cnt.nodes = 0; cnt.prims = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);
        if (!exhausted && idle && (__popc(idle) >= QZ_REFILL_MIN || idle == full)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&b.counters[C_CURSOR_TRACE], (uint32_t)__popc(idle));
            base = __shfl_sync(full, base, leader);
            if (!active) {
                const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < count) {
                    const uint8_t st = b.stage[idx];
                    if (st == ST_EMPTY) {
                        b.fam[idx] = QZ_FAM_NONE;
                    } else {
                        slot = idx;
                        first = st == ST_TRACE_FIRST;
                        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
                        Ray r;
                        r.o = v3(o.x, o.y, o.z);
                        r.d = v3(d.x, d.y, d.z);
                        trav_init(tv, r, INFINITY);
                        active = true;
                    }
                }
            }
            if (base + __popc(idle) >= count) exhausted = true;
        }
        if (!__any_sync(full, active)) { if (exhausted) break; else continue; }
#pragma unroll 1
        for (int k = 0; k < QZ_STEPS_PER_ROUND; k++)
            if (active && !tv.done) trav_step<false, COUNT>(sc, tv, &cnt);
        if (active && tv.done) {
            finish_closest(sc, b, slot, first, tv.best, flags);
            active = false;
        }
    }
    if (COUNT) {
        atomicAdd(&b.stats[S_NODES], (unsigned long long)cnt.nodes);
        atomicAdd(&b.stats[S_PRIMS], (unsigned long long)cnt.prims);
    }
}

// Tiny scenes (<= QZ_FLAT_MAX_PRIMS primitives: every shipped analytic scene): no BVH at all.
// The primitive records are staged once per CTA in shared memory and every lane walks the
// same list in lockstep -- no stack, no divergence between lanes, broadcast shared-memory
// reads.  The result is the brute-force minimum under the (t, key) order, i.e. exactly what the
// BVH traversal is defined to return.
__global__ void __launch_bounds__(256) k_closest_flat(DScene sc, WfBuffers b, uint32_t flags) {
    __shared__ FlatPrim s_prims[QZ_FLAT_MAX_PRIMS];
    const uint32_t n_prims = sc.n_prims;
    for (uint32_t i = threadIdx.x; i < n_prims; i += blockDim.x)
        s_prims[i] = make_flat_prim(sc.prims[4 * i], sc.prims[4 * i + 1], sc.prims[4 * i + 2], sc.prims[4 * i + 3]);
    __syncthreads();
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < b.pool; slot += gridDim.x * blockDim.x) {
        const uint8_t st = b.stage[slot];
        if (st == ST_EMPTY) { b.fam[slot] = QZ_FAM_NONE; continue; }
        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
        const V3 O = v3(o.x, o.y, o.z), D = v3(d.x, d.y, d.z);
        const float rd2 = 1.0f / dot(D, D);   // the ray's share of every sphere test
        Hit best;
        best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
        best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
#pragma unroll 1
        for (uint32_t p = 0; p < n_prims; p++)
            flat_prim_test(sc, s_prims[p], p, O, D, rd2, QZ_TNEAR, INFINITY, best);
        finish_closest(sc, b, slot, st == ST_TRACE_FIRST, best, flags);
    }
}

__global__ void __launch_bounds__(256) k_shadow_flat(DScene sc, WfBuffers b) {
    __shared__ FlatPrim s_prims[QZ_FLAT_MAX_PRIMS];
    const uint32_t n_prims = sc.n_prims;
    for (uint32_t i = threadIdx.x; i < n_prims; i += blockDim.x)
        s_prims[i] = make_flat_prim(sc.prims[4 * i], sc.prims[4 * i + 1], sc.prims[4 * i + 2], sc.prims[4 * i + 3]);
    __syncthreads();
    const uint32_t count = b.counters[C_SHADOW];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t slot = b.q_shadow[i];
        const float4 o = b.sh_o[slot], d = b.sh_d[slot];
        const V3 O = v3(o.x, o.y, o.z), D = v3(d.x, d.y, d.z);
        const float rd2 = 1.0f / dot(D, D);   // the ray's share of every sphere test
        Hit best;
        best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
        best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
        // occluded iff the closest hit has t <= 1 (scene.cpp:136-143)
#pragma unroll 1
        for (uint32_t p = 0; p < n_prims; p++)
            flat_prim_test(sc, s_prims[p], p, O, D, rd2, QZ_TNEAR, INFINITY, best);
        if (!(best.prim != QZ_NO_HIT && best.t <= 1.0f)) {
            const float4 L = b.radiance[slot], c = b.sh_c[slot];
            b.radiance[slot] = f4(L.x + c.x, L.y + c.y, L.z + c.z, L.w + c.w);
        }
    }
}

// Persistent any-hit traversal of the shadow queue; resolves next-event estimation.
template <bool COUNT>
__global__ void __launch_bounds__(128) k_shadow(DScene sc, WfBuffers b) {
    const uint32_t count = b.counters[C_SHADOW];
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    Trav tv;
    tv.done = true;
    bool active = false;
    bool exhausted = false;
    uint32_t slot = 0;
    TraversalCounters cnt;
    cnt.nodes = 0; cnt.prims = 0;
    for (;;) {
        const unsigned idle = __ballot_sync(full, !active);
        if (!exhausted && idle && (__popc(idle) >= QZ_REFILL_MIN || idle == full)) {
            const int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(&b.counters[C_CURSOR_SHADOW], (uint32_t)__popc(idle));
            base = __shfl_sync(full, base, leader);
            if (!active) {
                const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                if (idx < count) {
                    slot = b.q_shadow[idx];
                    const float4 o = b.sh_o[slot], d = b.sh_d[slot];
                    Ray r;
                    r.o = v3(o.x, o.y, o.z);
                    r.d = v3(d.x, d.y, d.z);
                    trav_init(tv, r, 1.0f);
                    active = true;
                }
            }
            if (base + __popc(idle) >= count) exhausted = true;
        }
        if (!__any_sync(full, active)) break;
#pragma unroll 1
        for (int k = 0; k < QZ_STEPS_PER_ROUND; k++)
            if (active && !tv.done) trav_step<true, COUNT>(sc, tv, &cnt);
        if (active && tv.done) {
            if (!tv.occluded) {
                const float4 L = b.radiance[slot], c = b.sh_c[slot];
                b.radiance[slot] = f4(L.x + c.x, L.y + c.y, L.z + c.z, L.w + c.w);
            }
            active = false;
        }
    }
    if (COUNT) {
        atomicAdd(&b.stats[S_NODES], (unsigned long long)cnt.nodes);
        atomicAdd(&b.stats[S_PRIMS], (unsigned long long)cnt.prims);
    }
}

// ------------------------------------------------------------------ phase-scheduled traversal
// One ray per lane, but the WARP decides what every lane does next.  A lane is in one of four
// states -- NODE (has a wide node to open), PRIM (has leaf primitives of the last node pending),
// DONE (result to write), IDLE (no ray) -- and each trip of the persistent loop runs exactly ONE
// of three instruction streams for all lanes that are in the matching state:
//   * finish + refill, when >= QZ_REFILL_MIN lanes wait for it (warp vote), or nothing else can run;
//   * the primitive stream (one primitive per pending lane), when more lanes are in PRIM than in NODE;
//   * the node stream otherwise: one 128-byte node per lane, all eight child slabs tested
//     branch-free, hit leaf children recorded as a bit mask over the node's contiguous leaf
//     block (NOT tested here), hit internal children sorted far-to-near by a register sorting
//     network and pushed; then the lane pops its next node, culling against its current best.
// So the long, expensive streams (8 slab tests; Moeller-Trumbore) always run with the majority of the
// warp's lanes, instead of every lane dragging the other 31 through its own private
// node / primitive / pop sequence (the first version of this kernel: 6.9 of 32 lanes active per
// instruction on the 1M-triangle scene, profiles/r01_traversal.md).  Results are unchanged: same
// slab and primitive arithmetic, closest hit = minimum under (t, key), which does not depend on
// the visiting order.
enum LaneState { LS_IDLE = 0, LS_NODE = 1, LS_PRIM = 2, LS_DONE = 3 };

#define QZ_CSWAP_DESC(a, b) { const uint32_t hi_ = a > b ? a : b, lo_ = a > b ? b : a; a = hi_; b = lo_; }

template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128, QZ_TRACE_MIN_BLOCKS) k_trace_lane(DScene sc, WfBuffers b, uint32_t flags) {
    const uint32_t count = ANY_HIT ? b.counters[C_SHADOW] : b.pool;
    uint32_t* cursor = &b.counters[ANY_HIT ? C_CURSOR_SHADOW : C_CURSOR_TRACE];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int state = LS_IDLE;
    bool exhausted = false;  // warp-uniform
    uint32_t slot = 0, cur = 0, leafbits = 0, leaf_base = 0;
    bool first = false, occl = false;
    V3 O = v3(0.0f, 0.0f, 0.0f), D = v3(0.0f, 0.0f, 0.0f);
    float inv[3] = {0.0f, 0.0f, 0.0f};
    float limit = 0.0f;   // closest hit: best t so far; any hit: the end of the segment
    Hit best;
    best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
    best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
    uint2 stack[QZ_STACK];
    int sp = 0;
    uint32_t cnt_nodes = 0, cnt_prims = 0, overflow = 0;

    // next node for this lane: nearest entry still within reach, or the ray is done
    auto pop_next = [&]() {
        for (;;) {
            if (sp == 0) { state = LS_DONE; return; }
            sp--;
            const uint2 e = stack[sp];
            if (__uint_as_float(e.x & ~7u) <= limit) {
                cur = e.y; state = LS_NODE;
                prefetch_line(sc.nodes + cur);   // the node stream that opens it is at least a trip away
                return;
            }
        }
    };

    for (;;) {
        const unsigned m_node = __ballot_sync(full, state == LS_NODE);
        const unsigned m_prim = __ballot_sync(full, state == LS_PRIM);
        const unsigned m_done = __ballot_sync(full, state == LS_DONE);
        const unsigned m_idle = ~(m_node | m_prim | m_done);
        const bool busy = (m_node | m_prim) != 0u;
        const bool want_finish = m_done != 0u && (__popc(m_done) >= QZ_REFILL_MIN || !busy);
        const bool want_refill = !exhausted && (__popc(m_done | m_idle) >= QZ_REFILL_MIN || !busy);
        if (want_finish || want_refill) {
            // ---- finish + refill stream
            if (state == LS_DONE) {
                if (ANY_HIT) {
                    if (!occl) {
                        const float4 L = b.radiance[slot], c = b.sh_c[slot];
                        b.radiance[slot] = f4(L.x + c.x, L.y + c.y, L.z + c.z, L.w + c.w);
                    }
                } else {
                    finish_closest(sc, b, slot, first, best, flags);
                }
                state = LS_IDLE;
            }
            if (!exhausted) {
                const unsigned idle = m_done | m_idle;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(full, base, 0);
                if (state == LS_IDLE) {
                    const uint32_t idx = base + __popc(idle & ((1u << lane) - 1u));
                    if (idx < count) {
                        float4 ro, rd;
                        bool live = true;
                        if (ANY_HIT) {
                            slot = b.q_shadow[idx];
                            ro = b.sh_o[slot]; rd = b.sh_d[slot];
                            limit = 1.0f;
                        } else {
                            const uint8_t st = b.stage[idx];
                            live = st != ST_EMPTY;
                            if (live) {
                                slot = idx;
                                first = st == ST_TRACE_FIRST;
                                ro = b.ray_o[slot]; rd = b.ray_d[slot];
                                limit = INFINITY;
                            } else {
                                b.fam[idx] = QZ_FAM_NONE;
                            }
                        }
                        if (live) {
                            O = v3(ro.x, ro.y, ro.z); D = v3(rd.x, rd.y, rd.z);
                            inv[0] = 1.0f / D.x; inv[1] = 1.0f / D.y; inv[2] = 1.0f / D.z;
                            best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
                            best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
                            occl = false;
                            cur = 0; sp = 0; leafbits = 0;
                            state = LS_NODE;
                        }
                    }
                }
                if (base + __popc(idle) >= count) exhausted = true;
            }
            continue;
        }
        if (!busy) break;  // nothing in flight, nothing to write, no slots left

#ifndef QZ_PRIM_FIRST_MIN
#define QZ_PRIM_FIRST_MIN 33   /* run the primitive stream whenever at least this many lanes wait for it (33 = majority rule only) */
#endif
        if (__popc(m_prim) > __popc(m_node) || __popc(m_prim) >= QZ_PRIM_FIRST_MIN) {
            // ---- primitive stream: one pending primitive per lane
            if (state == LS_PRIM) {
                const uint32_t p = leaf_base + (uint32_t)(__ffs(leafbits) - 1);
                leafbits &= leafbits - 1u;
                if (COUNT) cnt_prims++;
                if (ANY_HIT) {
                    best.t = INFINITY; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
                    prim_test(sc, p, O, D, QZ_TNEAR, INFINITY, best);
                    // a hit beyond the segment does not occlude (scene.cpp:136-143)
                    if (best.prim != QZ_NO_HIT && best.t <= limit) { occl = true; state = LS_DONE; }
                } else {
                    prim_test(sc, p, O, D, QZ_TNEAR, INFINITY, best);
                    limit = best.t;
                }
                if (state == LS_PRIM) {
                    if (leafbits == 0u) pop_next();
                    else prefetch_line(sc.prims + (size_t)(leaf_base + (uint32_t)(__ffs(leafbits) - 1)) * 4);
                }
            }
            continue;
        }

        // ---- node stream: open one node per lane
        if (state == LS_NODE) {
            const uint4* np = reinterpret_cast<const uint4*>(sc.nodes + cur);
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = __ldg(np + i);
            if (COUNT) cnt_nodes++;
            const float org[3] = {__uint_as_float(w[0].x), __uint_as_float(w[0].y), __uint_as_float(w[0].z)};
            const float scl[3] = {exp_scale(w[0].w & 0xffu), exp_scale((w[0].w >> 8) & 0xffu), exp_scale((w[0].w >> 16) & 0xffu)};
            const uint32_t child_base = w[1].x;
            leaf_base = w[1].y;
            const uint32_t metas[2] = {w[1].z, w[1].w};
            // quantised planes: words 2..4 = qlo[x,y,z][8], words 5..7 = qhi[x,y,z][8] (u16 each).  Per axis
            // the ray's direction sign says which of the two is the entry plane -- chosen once per
            // node on the packed words, not per child on the decoded distances.
            uint32_t qn[3][4], qf[3][4];
            float bias[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const bool fwd = inv[a] >= 0.0f;
                const uint32_t* lo4 = reinterpret_cast<const uint32_t*>(&w[2 + a]);
                const uint32_t* hi4 = reinterpret_cast<const uint32_t*>(&w[5 + a]);
#pragma unroll
                for (int j = 0; j < 4; j++) { qn[a][j] = fwd ? lo4[j] : hi4[j]; qf[a][j] = fwd ? hi4[j] : lo4[j]; }
                bias[a] = a == 0 ? O.x : (a == 1 ? O.y : O.z);
            }
            uint32_t key[8];
            uint32_t lb = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t m = (metas[k >> 2] >> (8 * (k & 3))) & 0xffu;
                float t0 = QZ_TNEAR, t1 = limit;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const uint32_t q0 = (qn[a][k >> 1] >> (16 * (k & 1))) & 0xffffu;
                    const uint32_t q1 = (qf[a][k >> 1] >> (16 * (k & 1))) & 0xffffu;
                    // plane = org + q * 2^e: the product is exact, so the fused form rounds once, to the same value
                    const float pn = __fmaf_rn((float)q0, scl[a], org[a]);
                    const float pf = __fmaf_rn((float)q1, scl[a], org[a]);
                    t0 = fmaxf(t0, (pn - bias[a]) * inv[a]);   // fmaxf / fminf drop NaN (0 * inf)
                    t1 = fminf(t1, (pf - bias[a]) * inv[a]);
                }
                // box_hit's conservative widening (bvh.cuh) is a monotone map of each slab distance, so it
                // is applied once, to the max / min of them; QZ_TNEAR and the limit stay as they are
                float e0 = t0 * (t0 >= 0.0f ? 0.9999995f : 1.0000005f);
                float e1 = t1 * (t1 >= 0.0f ? 1.0000005f : 0.9999995f);
                e0 = fmaxf(e0, QZ_TNEAR);
                e1 = fminf(e1, limit);
                const bool hit = m != 0u && e0 <= e1;
                const bool internal = (m & 0x80u) != 0u;
                key[k] = (hit && internal) ? ((__float_as_uint(e0) & ~7u) | (m & 7u)) : 0u;
                if (hit && !internal) lb |= ((1u << (m >> 5)) - 1u) << (m & 31u);
            }
            // sort the (distance | child slot) keys descending: Batcher's 19-comparator network.  Shadow rays stop at
            // the first occluder wherever it is, so any-hit traversal pushes the children as they come.
            if (!ANY_HIT) {
                QZ_CSWAP_DESC(key[0], key[1]); QZ_CSWAP_DESC(key[2], key[3]); QZ_CSWAP_DESC(key[4], key[5]); QZ_CSWAP_DESC(key[6], key[7]);
                QZ_CSWAP_DESC(key[0], key[2]); QZ_CSWAP_DESC(key[1], key[3]); QZ_CSWAP_DESC(key[4], key[6]); QZ_CSWAP_DESC(key[5], key[7]);
                QZ_CSWAP_DESC(key[1], key[2]); QZ_CSWAP_DESC(key[5], key[6]);
                QZ_CSWAP_DESC(key[0], key[4]); QZ_CSWAP_DESC(key[1], key[5]); QZ_CSWAP_DESC(key[2], key[6]); QZ_CSWAP_DESC(key[3], key[7]);
                QZ_CSWAP_DESC(key[2], key[4]); QZ_CSWAP_DESC(key[3], key[5]);
                QZ_CSWAP_DESC(key[1], key[2]); QZ_CSWAP_DESC(key[3], key[4]); QZ_CSWAP_DESC(key[5], key[6]);
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (key[k]) {
                    if (sp < QZ_STACK) { stack[sp] = make_uint2(key[k], child_base + (key[k] & 7u)); sp++; }
                    else overflow = 1u;
                }
            }
            leafbits = lb;
            if (lb) {
                state = LS_PRIM;
                prefetch_line(sc.prims + (size_t)(leaf_base + (uint32_t)(__ffs(lb) - 1)) * 4);  // first pending primitive record
            } else {
                pop_next();
            }
        }
    }
    if (COUNT) {
        atomicAdd(&b.stats[S_NODES], (unsigned long long)cnt_nodes);
        atomicAdd(&b.stats[S_PRIMS], (unsigned long long)cnt_prims);
    }
    if (overflow) atomicAdd(&b.stats[S_OVERFLOW], 1ull);
}

// ------------------------------------------------------------------ octet traversal
// Warp-cooperative BVH traversal: EIGHT LANES PER RAY, one lane per child of the wide node, four
// rays per warp advancing in lockstep (node phase, leaf phase, push/pop phase), so that lanes of a
// warp never sit in different parts of the state machine:
//   * node phase: the octet reads ONE 128-byte node line (header broadcast + each lane its own
//     child's six quantised planes) and every lane slab-tests its child -- 8 box tests per
//     instruction stream instead of 8 serial ones per lane;
//   * leaf phase: the primitives of all leaf children that were hit form a bit mask over the
//     node's contiguous leaf block; the octet's lanes take one primitive each;
//   * push/pop: hit internal children are ranked by entry distance with 7 shuffles and written,
//     far to near, to the octet's SHARED-MEMORY stack; the next node is popped with culling
//     against the current best distance;
//   * an octet whose ray has finished takes the next live slot of its private 16-slot chunk
//     (one atomic per chunk) -- the persistent, vote-driven refill of the per-lane kernel at
//     octet granularity.
// The answer is the brute-force minimum under the (t, key) order, exactly as the per-lane
// traversal (bvh.cuh) defines it: the same box and primitive arithmetic, a lane-local best per
// lane and one (t, key) arg-min over the octet when the ray is done.
#define QZ_OCT_STACK 96   /* shared stack entries per octet (internal nodes only) */
#define QZ_OCT_CHUNK 16   /* slots an octet reserves with one atomic */

__device__ __forceinline__ uint32_t nonzero_byte_nibble(uint32_t w) {
    return ((__vcmpne4(w, 0u) & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ uint32_t equal_byte_nibble(uint32_t w, uint32_t pattern) {
    return ((__vcmpeq4(w, pattern) & 0x08040201u) * 0x01010101u) >> 24;
}

template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128, ANY_HIT ? 8 : 6) k_trace_oct(DScene sc, WfBuffers b, uint32_t flags) {
    __shared__ uint2 s_stack[16][QZ_OCT_STACK + 1];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int k = lane & 7;            // the child this lane tests
    const int obase = lane & ~7;       // first lane of the octet
    uint2* stack = s_stack[threadIdx.x >> 3];
    const uint32_t count = ANY_HIT ? b.counters[C_SHADOW] : b.pool;
    uint32_t* cursor = &b.counters[ANY_HIT ? C_CURSOR_SHADOW : C_CURSOR_TRACE];

    // octet-uniform state
    bool active = false, exhausted = false;
    uint32_t chunk_base = 0, live = 0, firsts = 0;
    uint32_t slot = 0, cur = 0;
    bool first = false;
    V3 O = v3(0.0f, 0.0f, 0.0f), D = v3(0.0f, 0.0f, 0.0f);
    float o[3] = {0.0f, 0.0f, 0.0f}, inv[3] = {0.0f, 0.0f, 0.0f};
    float limit = 0.0f;   // closest hit: best t of the octet so far; any hit: the segment end
    int sp = 0;
    // lane-local
    Hit best;
    best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
    best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
    uint32_t cnt_nodes = 0, cnt_prims = 0, overflow = 0;

    for (;;) {
        // ---- refill: idle octets take the next live slot; an empty chunk is replaced first
        while (__any_sync(full, !active && !(exhausted && live == 0u))) {
            const bool want = !active && !(exhausted && live == 0u);
            const bool fetch = want && live == 0u;
            uint32_t base = 0;
            if (fetch && k == 0) base = atomicAdd(cursor, (uint32_t)QZ_OCT_CHUNK);
            base = __shfl_sync(full, base, obase);
            if (fetch) {
                if (base >= count) {
                    exhausted = true;
                } else {
                    chunk_base = base;
                    const uint32_t n = count - base < (uint32_t)QZ_OCT_CHUNK ? count - base : (uint32_t)QZ_OCT_CHUNK;
                    uint32_t lv = (1u << n) - 1u;
                    if (!ANY_HIT) {
                        // the stage tags of the chunk say which slots carry a ray this iteration
                        const uint4 tg = *reinterpret_cast<const uint4*>(b.stage + base);
                        const uint32_t nz = nonzero_byte_nibble(tg.x) | (nonzero_byte_nibble(tg.y) << 4) |
                                            (nonzero_byte_nibble(tg.z) << 8) | (nonzero_byte_nibble(tg.w) << 12);
                        const uint32_t pat = 0x01010101u * (uint32_t)ST_TRACE_FIRST;
                        firsts = equal_byte_nibble(tg.x, pat) | (equal_byte_nibble(tg.y, pat) << 4) |
                                 (equal_byte_nibble(tg.z, pat) << 8) | (equal_byte_nibble(tg.w, pat) << 12);
                        lv &= nz;
                        for (uint32_t i = (uint32_t)k; i < n; i += 8u)
                            if (!((lv >> i) & 1u)) b.fam[base + i] = QZ_FAM_NONE;
                    }
                    live = lv;
                }
            } else if (want) {
                const int i = __ffs(live) - 1;
                live &= live - 1u;
                float4 ro, rd;
                if (ANY_HIT) {
                    slot = b.q_shadow[chunk_base + (uint32_t)i];
                    ro = b.sh_o[slot]; rd = b.sh_d[slot];
                    limit = 1.0f;
                } else {
                    slot = chunk_base + (uint32_t)i;
                    first = ((firsts >> i) & 1u) != 0u;
                    ro = b.ray_o[slot]; rd = b.ray_d[slot];
                    limit = INFINITY;
                }
                O = v3(ro.x, ro.y, ro.z); D = v3(rd.x, rd.y, rd.z);
                o[0] = O.x; o[1] = O.y; o[2] = O.z;
                inv[0] = 1.0f / D.x; inv[1] = 1.0f / D.y; inv[2] = 1.0f / D.z;
                best.t = INFINITY; best.prim = QZ_NO_HIT; best.key = 0xffffffffu; best.geom_id = QZ_NO_HIT;
                best.u = 0.0f; best.v = 0.0f; best.ng = v3(0.0f, 0.0f, 0.0f); best.prim_id = 0;
                cur = 0; sp = 0;
                active = true;
            }
        }
        if (!__any_sync(full, active)) break;

        // ---- node phase: lane k slab-tests child k of the octet's current node
        bool hit_child = false;
        float dist = 0.0f;
        uint32_t meta = 0, child_base = 0, leaf_base = 0;
        if (active) {
            const uint4* np = reinterpret_cast<const uint4*>(sc.nodes + cur);
            const uint4 h0 = __ldg(np), h1 = __ldg(np + 1);
            meta = ((k < 4 ? h1.z : h1.w) >> (8 * (k & 3))) & 0xffu;
            child_base = h1.x; leaf_base = h1.y;
            if (COUNT && k == 0) cnt_nodes++;
            if (meta) {
                const unsigned short* q = reinterpret_cast<const unsigned short*>(np + 2);
                const float org[3] = {__uint_as_float(h0.x), __uint_as_float(h0.y), __uint_as_float(h0.z)};
                const float scl[3] = {exp_scale(h0.w & 0xffu), exp_scale((h0.w >> 8) & 0xffu), exp_scale((h0.w >> 16) & 0xffu)};
                float lo[3], hi[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    lo[a] = org[a] + (float)__ldg(q + a * 8 + k) * scl[a];
                    hi[a] = org[a] + (float)__ldg(q + 24 + a * 8 + k) * scl[a];
                }
                hit_child = box_hit(lo, hi, o, inv, QZ_TNEAR, limit, dist);
            }
        }

        // ---- leaf phase: the primitives of the hit leaf children, one per lane and round
        uint32_t bits = (hit_child && !(meta & 0x80u)) ? (((1u << (meta >> 5)) - 1u) << (meta & 31u)) : 0u;
        bits |= __shfl_xor_sync(full, bits, 1);
        bits |= __shfl_xor_sync(full, bits, 2);
        bits |= __shfl_xor_sync(full, bits, 4);
        bool occl = false;
        if (__any_sync(full, bits != 0u)) {
            do {
                uint32_t m = bits;
                for (int i = 0; i < k; i++) m &= m - 1u;
                if (m) {
                    const uint32_t p = leaf_base + (uint32_t)(__ffs(m) - 1);
                    if (COUNT) cnt_prims++;
                    if (ANY_HIT) {
                        best.t = INFINITY; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
                        prim_test(sc, p, O, D, QZ_TNEAR, INFINITY, best);
                        occl = occl || (best.prim != QZ_NO_HIT && best.t <= limit);
                    } else {
                        prim_test(sc, p, O, D, QZ_TNEAR, INFINITY, best);
                    }
                }
                // the lowest eight set bits are done
                if (__popc(bits) <= 8) bits = 0u;
                else {
#pragma unroll
                    for (int i = 0; i < 8; i++) bits &= bits - 1u;
                }
            } while (__any_sync(full, bits != 0u));
            if (ANY_HIT) {
                occl = ((__ballot_sync(full, occl) >> obase) & 0xffu) != 0u;
            } else {
                float t = best.t;
                t = fminf(t, __shfl_xor_sync(full, t, 1));
                t = fminf(t, __shfl_xor_sync(full, t, 2));
                t = fminf(t, __shfl_xor_sync(full, t, 4));
                if (active) limit = t;
            }
        }

        // ---- push the hit internal children far-to-near, pop the next node
        const bool push = hit_child && (meta & 0x80u) && (ANY_HIT || dist <= limit);
        const uint32_t key = push ? ((__float_as_uint(dist) & ~7u) | (uint32_t)k) : 0u;
        int rank = 0;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            const uint32_t other = __shfl_sync(full, key, obase | ((k + j) & 7));
            rank += other > key ? 1 : 0;
        }
        const int n_push = __popc((__ballot_sync(full, push) >> obase) & 0xffu);
        if (push) {
            if (sp + rank < QZ_OCT_STACK) stack[sp + rank] = make_uint2(__float_as_uint(dist), child_base + (meta & 0x7fu));
            else overflow = 1u;
        }
        __syncwarp();
        bool done = false;
        if (active) {
            sp = sp + n_push < QZ_OCT_STACK ? sp + n_push : QZ_OCT_STACK;
            if (ANY_HIT && occl) {
                done = true;
            } else {
                for (;;) {
                    if (sp == 0) { done = true; break; }
                    sp--;
                    const uint2 e = stack[sp];
                    if (__uint_as_float(e.x) <= limit) { cur = e.y; break; }
                }
            }
        }
        __syncwarp();

        // ---- finished rays
        if (__any_sync(full, done)) {
            if (ANY_HIT) {
                if (done && !occl && k == 0) {
                    const float4 L = b.radiance[slot], c = b.sh_c[slot];
                    b.radiance[slot] = f4(L.x + c.x, L.y + c.y, L.z + c.z, L.w + c.w);
                }
            } else {
                // arg-min of (t, key) over the octet; t > 0 or +inf, so its bit pattern orders like the value
                uint32_t bt = __float_as_uint(best.t), bk = best.key;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    const uint32_t ot = __shfl_xor_sync(full, bt, d), ok = __shfl_xor_sync(full, bk, d);
                    if (ot < bt || (ot == bt && ok < bk)) { bt = ot; bk = ok; }
                }
                if (done) {
                    const bool miss = bk == 0xffffffffu;  // no lane found a hit
                    const bool winner = miss ? k == 0 : (best.prim != QZ_NO_HIT && __float_as_uint(best.t) == bt && best.key == bk);
                    if (winner) finish_closest(sc, b, slot, first, best, flags);
                }
            }
            if (done) active = false;
        }
    }
    if (COUNT) {
        if (cnt_nodes) atomicAdd(&b.stats[S_NODES], (unsigned long long)cnt_nodes);
        if (cnt_prims) atomicAdd(&b.stats[S_PRIMS], (unsigned long long)cnt_prims);
    }
    if (overflow) atomicAdd(&b.stats[S_OVERFLOW], 1ull);
}

// roles (bit mask over SampleRole) a queue's paths will read this bounce
__device__ __forceinline__ uint32_t queue_roles(int queue_id, uint32_t n_lights) {
    const int fam = queue_id % SQ_FAMILIES;
    const bool first = queue_id >= SQ_FAMILIES;
    if (fam == SQ_MISC) return 0u;  // run-time dispatch: evaluated on the fly inside k_shade<KH_ANY>
    uint32_t m = 0;
    if (fam == SQ_DIFFUSE || fam == SQ_CONDUCTOR) {
        m |= (3u << R_LIGHT) | (3u << R_BSDF);
        if (n_lights > 1) m |= 1u << R_PICK;
    } else {
        m |= 1u << R_U1;
    }
    if (!first) m |= 1u << R_RR;  // roulette cannot apply at depth 1 (render.cpp:202)
    return m;
}

// The sampler as its own wavefront stage.  The Owen-scrambled Halton evaluation is a long,
// purely integer, dependent chain per digit (~75 instructions); inside the shading kernels it
// runs at low occupancy next to float-heavy code and diverges with every lane's bounce depth.
// Here one THREAD evaluates ONE dimension of one path: 32 registers, full occupancy, lanes of a
// warp work on neighbouring dimensions of the same few paths, nothing else in the kernel.
// k_shade then reads eight floats per path instead of running the sampler.
__global__ void __launch_bounds__(256) k_sample(DScene sc, WfBuffers b) {
    // prefix sums of (entries x roles) over the six sampled queues
    uint32_t start[SQ_COUNT + 1];
    uint32_t nrole[SQ_COUNT], rmagic[SQ_COUNT], rlist[SQ_COUNT];
    start[0] = 0;
#pragma unroll
    for (int q = 0; q < SQ_COUNT; q++) {
        const uint32_t roles = queue_roles(q, sc.n_lights);
        nrole[q] = __popc(roles);
        // item -> (entry, role) without an integer division: ceil(2^32 / n) as a multiply-high magic (exact for
        // items < 2^32 / 7), and the queue's roles listed in a nibble string
        rmagic[q] = nrole[q] ? (uint32_t)(((1ull << 32) + nrole[q] - 1u) / nrole[q]) : 0u;
        uint32_t list = 0, m = roles;
#pragma unroll
        for (int j = 0; j < R_COUNT; j++) {
            if (m) { list |= (uint32_t)(__ffs(m) - 1) << (4 * j); m &= m - 1u; }
        }
        rlist[q] = list;
        start[q + 1] = start[q] + b.counters[C_SHADE0 + q] * nrole[q];
    }
    const uint32_t total = start[SQ_COUNT];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < SQ_COUNT; k++) q += (t >= start[k]) ? 1 : 0;
        const uint32_t local = t - start[q];
        const uint32_t entry = nrole[q] == 1u ? local : __umulhi(local, rmagic[q]);
        const uint32_t r = local - entry * nrole[q];
        const int role = (int)((rlist[q] >> (4u * r)) & 15u);
        const uint32_t slot = b.q_shade[q][entry];
        const uint4 misc = b.misc[slot];
        const int fam = q % SQ_FAMILIES;
        bool nee = fam == SQ_DIFFUSE;
        if (fam == SQ_CONDUCTOR) {
            const uint32_t geom = __float_as_uint(b.hit_b[slot].w);
            const qz_material mat = sc.materials[sc.geoms[geom].material];
            nee = !(mat.alpha_x < 1e-3f && mat.alpha_y < 1e-3f);
        }
        // the dimension of this role: bounce_dims() in closed form while the bounce cannot wrap past the prime table
        const uint32_t d0 = misc.z & 0xffffu;
        const bool has_lights = sc.n_lights != 0;
        uint32_t dim;
        if (d0 + 10u < QZ_N_PRIMES) {
            const uint32_t light0 = d0 + 1u + ((nee && has_lights) ? 1u : 0u);
            const uint32_t bsdf0 = nee ? light0 + (has_lights ? 2u : 4u) : d0 + 1u;
            dim = role == R_MAT ? d0 : role == R_PICK ? ((nee && has_lights) ? d0 + 1u : 2u) : role == R_LIGHT ? light0 : role == R_LIGHT + 1 ? light0 + 1u
                : role == R_BSDF ? bsdf0 : role == R_BSDF + 1 ? bsdf0 + 1u : role == R_U1 ? bsdf0 + 2u : bsdf0 + 3u;
        } else {
            uint32_t dims[R_COUNT];
            bounce_dims(d0, nee, has_lights, dims);
            dim = dims[role];
        }
        Sampler smp;
        smp.index = misc.y; smp.dim = 0;
        reinterpret_cast<float*>(&b.samples[slot])[role] = sample_dimension(sc.sampler_table, smp, dim);
    }
}

// Depth-0 albedo of conductors as its own stage.  The reference estimates the albedo AOV at the first
// hit with 16 fixed BxDF samples (render.cpp:150-170, bxdf.hpp:46-56); for a rough conductor that is
// 16 x (visible-normal sample, D, G, complex Fresnel at 4 wavelengths) -- about three times the rest
// of the bounce, executed as one serial chain per thread inside a 128-register kernel.  Here SIXTEEN
// LANES share a path, one sample each, in a lean kernel at high occupancy; the 16 terms are then
// added in sample order by every lane of the group (shuffles), which is the reference's
// summation order, so the value is unchanged bit for bit.
__global__ void __launch_bounds__(256) k_albedo_conductor(DScene sc, WfBuffers b, uint32_t max_bounces) {
    if (max_bounces == 0) return;  // the path loop breaks before the estimate (render.cpp:137): the AOV stays zero
    const uint32_t count = b.counters[C_SHADE0 + SQ_FAMILIES + SQ_CONDUCTOR];
    const uint32_t* queue = b.q_shade[SQ_FAMILIES + SQ_CONDUCTOR];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int i = lane & 15;          // this lane's sample
    const int gbase = lane & 16;      // first lane of the 16-lane group
    const uint32_t n_groups = gridDim.x * blockDim.x / 16u;
    const uint32_t rounds = (count + n_groups - 1u) / n_groups;
    uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) / 16u;
    for (uint32_t r = 0; r < rounds; r++, e += n_groups) {
        const bool live = e < count;   // uniform within the group; the warp stays converged for the shuffles
        Spec4 term = spec4(0.0f);
        bool valid = false;
        uint32_t slot = 0;
        SurfacePoint sp;
        qz_material mat;
        float coeff = 0.0f;
        if (live) {
            slot = queue[e];
            const float4 o = b.ray_o[slot], d = b.ray_d[slot], ha = b.hit_a[slot], hb = b.hit_b[slot];
            Ray ray;
            ray.o = v3(o.x, o.y, o.z); ray.d = v3(d.x, d.y, d.z);
            Hit hit;
            hit.t = ha.x; hit.u = ha.y; hit.v = ha.z; hit.prim_id = __float_as_uint(ha.w);
            hit.ng = v3(hb.x, hb.y, hb.z); hit.geom_id = __float_as_uint(hb.w); hit.prim = hit.geom_id; hit.key = 0;
            const Spec4 lambda = s4(b.lambda[slot]);
            sp = make_surface_point(sc, ray, hit);
            // Material::bsdf for a conductor (material.cpp:16-20) is eight spectrum lookups (eta and k at four
            // wavelengths): lane j of the group does lookup j, the group exchanges them below
            mat = sc.materials[sp.material];
            const int j = i & 7;
            const float lam = j & 2 ? (j & 1 ? lambda.v[3] : lambda.v[2]) : (j & 1 ? lambda.v[1] : lambda.v[0]);
            coeff = eval_spectrum(sc, j < 4 ? mat.a : mat.b, lam);
        }
        Bsdf f;
        f.kind = BX_CONDUCTOR; f.ior = 1.0f;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            f.a.v[c] = __shfl_sync(full, coeff, gbase + c);
            f.b.v[c] = __shfl_sync(full, coeff, gbase + 4 + c);
        }
        if (live) {
            f.rough.ax = mat.alpha_x; f.rough.ay = mat.alpha_y;
            make_basis(sp.normal, f.u0, f.u1, f.u2);
            const V3 wo = to_local(f, sp.wo);
            const float* t = sc.rho_tab + i * 8;
            const BsdfSample smp = bxdf_sample<KH_CONDUCTOR>(f, wo, t[0], v2(t[1], t[2]), true, v3(t[6], t[7], 0.0f));
            valid = smp.valid;
            if (valid) term = smp.spec * fabsf(smp.wi.z) / smp.pdf;
        }
        Spec4 acc = spec4(0.0f);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const bool ok = __shfl_sync(full, (int)valid, gbase + j) != 0;
            Spec4 v;
#pragma unroll
            for (int c = 0; c < 4; c++) v.v[c] = __shfl_sync(full, term.v[c], gbase + j);
            if (ok) acc = acc + v;
        }
        if (live && i == 0) b.aov_a[slot] = f4(acc / 16.0f);
    }
}

// One bounce for every path of one (family, first-hit?) queue.  Only what the family can change
// is loaded and stored: the radiance buffer is touched only when the hit itself adds radiance
// (emitters and misses live in the run-time-dispatch queue), the wavelength pdf only by the
// dielectric and run-time-dispatch kernels (dispersion), the AOVs only by first-hit kernels.
template <int KH, int FIRST>
__global__ void __launch_bounds__(128, (KH == KH_DIFFUSE || KH == KH_DIELECTRIC) ? QZ_SHADE_MIN_BLOCKS_LIGHT : QZ_SHADE_MIN_BLOCKS_HEAVY)
k_shade(DScene sc, WfBuffers b, int queue_id, uint32_t max_bounces) {
    const uint32_t count = b.counters[C_SHADE0 + queue_id];
    const uint32_t* queue = b.q_shade[queue_id];
    // A bounce is: read one record, ~2000 dependent instructions, write it back -- nothing in
    // it overlaps the read.  So the read of the thread's NEXT path is started a whole bounce
    // early: its queue entry is loaded two trips ahead, its record lines prefetched one trip ahead.
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_cur = i < count ? queue[i] : 0u;
    uint32_t slot_next = i + stride < count ? queue[i + stride] : 0u;
    for (; i < count; i += stride) {
        const uint32_t slot = slot_cur;
        const uint32_t slot_after = (i + 2u * stride < count && i + 2u * stride >= i) ? queue[i + 2u * stride] : 0u;
        if (i + stride < count) {
            prefetch_line(&b.ray_o[slot_next]);
            if (KH != KH_ANY) prefetch_line(&b.samples[slot_next]);
        }
        slot_cur = slot_next;
        slot_next = slot_after;
        PathState ps;
        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
        ps.ray.o = v3(o.x, o.y, o.z); ps.ior_scale = o.w;
        ps.ray.d = v3(d.x, d.y, d.z); ps.p_b = d.w;
        ps.weight = s4(b.weight[slot]);
        ps.lambda = s4(b.lambda[slot]);
        ps.pdf = s4(b.lpdf[slot]);   // same 32-byte sector as misc: free to read, and written back whole
        ps.L = spec4(0.0f);
        const uint4 m = b.misc[slot];
        ps.smp.index = m.y;
        ps.smp.dim = m.z & 0xffffu;
        ps.depth = (m.z >> 16) & 0xffu;
        ps.flags = m.z >> 24;
        ps.n_rays = m.w + 1;  // + the closest-hit query that produced this hit
        const float4 ha = b.hit_a[slot], hb = b.hit_b[slot];
        Hit hit;
        hit.t = ha.x; hit.u = ha.y; hit.v = ha.z; hit.prim_id = __float_as_uint(ha.w);
        hit.ng = v3(hb.x, hb.y, hb.z);
        hit.geom_id = __float_as_uint(hb.w);
        hit.prim = hit.geom_id;  // only compared against QZ_NO_HIT from here on
        hit.key = 0;
        PathAov aov;
        aov.normal = v3(0.0f, 0.0f, 0.0f);
        aov.albedo = spec4(0.0f);
        const bool first = FIRST < 0 ? ps.depth == 0 : FIRST != 0;
        ShadowRequest sh;
        Spec4 gain;
        bool has_gain, alive;
        if (KH == KH_ANY) {
            SamplesOnTheFly src;
            src.tab = sc.sampler_table; src.index = ps.smp.index;
            alive = shade_bounce<KH, FIRST>(sc, ps, aov, hit, max_bounces, sh, src, gain, has_gain);
        } else {
            const float4* sv = &b.samples[slot];
            const float4 s0 = sv[0], s1 = sv[1];
            SamplesPrecomputed src;
            src.v[0] = s0.x; src.v[1] = s0.y; src.v[2] = s0.z; src.v[3] = s0.w;
            src.v[4] = s1.x; src.v[5] = s1.y; src.v[6] = s1.z; src.v[7] = s1.w;
            alive = shade_bounce<KH, FIRST>(sc, ps, aov, hit, max_bounces, sh, src, gain, has_gain);
        }
        if (has_gain) b.radiance[slot] = f4(s4(b.radiance[slot]) + gain);
        if (first) {
            // depth is still 0 after an emitter pass-through, so these may be written more than
            // once per path; the last write (the first real surface) wins, as in the reference
            if (hit.prim != QZ_NO_HIT) b.aov_n[slot] = f4(aov.normal.x, aov.normal.y, aov.normal.z, 0.0f);
            // (first-hit conductors: the albedo comes from k_albedo_conductor)
            if (!(KH == KH_CONDUCTOR && FIRST == 1) && (ps.depth != 0 || !alive)) b.aov_a[slot] = f4(aov.albedo);
        }
        uint32_t post = alive ? 0u : QZ_POST_DONE;
        if (ps.flags & QZ_FLAG_HAS_SHADOW) {
            ps.n_rays++;
            b.sh_o[slot] = f4(sh.o.x, sh.o.y, sh.o.z, 0.0f);
            b.sh_d[slot] = f4(sh.d.x, sh.d.y, sh.d.z, 0.0f);
            b.sh_c[slot] = f4(sh.contrib);
            post |= QZ_POST_SHADOW;
        }
        if (alive) {
            b.ray_o[slot] = f4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ior_scale);
            b.ray_d[slot] = f4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, ps.p_b);
            // weight and lambda share a sector: writing both keeps it a full-sector store
            b.weight[slot] = f4(ps.weight);
            b.lambda[slot] = f4(ps.lambda);
        }
        b.lpdf[slot] = f4(ps.pdf);
        b.misc[slot] = make_uint4(m.x, ps.smp.index, ps.smp.dim | (ps.depth << 16) | (ps.flags << 24), ps.n_rays);
        b.post[slot] = (uint8_t)post;
        b.stage[slot] = alive ? (ps.depth == 0 ? ST_TRACE_FIRST : ST_TRACE) : ST_EMPTY;
    }
}

// Finished paths: sensor conversion into the result cells, then a new path in the same slot.
__global__ void __launch_bounds__(256) k_finish(DScene sc, DCamera cam, WfBuffers b, PassParams pp) {
    const uint32_t count = b.counters[C_DONE];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t slot = b.q_done[i];
        const uint32_t path_id = b.misc[slot].x;
        const Spec4 L = s4(b.radiance[slot]), lambda = s4(b.lambda[slot]), pdf = s4(b.lpdf[slot]);
        const float4 n = b.aov_n[slot];
        const V3 rgb = to_sensor_rgb(cam, L, lambda, pdf);
        const V3 argb = to_sensor_rgb(cam, s4(b.aov_a[slot]), lambda, pdf);
        b.res_a[path_id] = f4(rgb.x, rgb.y, rgb.z, n.x);
        b.res_b[path_id] = f4(argb.x, argb.y, argb.z, n.y);
        b.res_c[path_id] = n.z;
        // regenerate
        const unsigned peers = __activemask();
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(b.next_path, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        const uint32_t next_id = base + __popc(peers & ((1u << lane) - 1u));
        if (next_id < pp.total) {
            init_slot(sc, cam, b, pp, slot, next_id);
            b.stage[slot] = ST_TRACE_FIRST;
        }
    }
}

// Between iterations: account the queue sizes, clear the consumed queues and the cursors.
__global__ void k_next_iteration(WfBuffers b) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        uint32_t* c = b.counters;
        unsigned long long shade = 0;
        for (int k = 0; k < SQ_COUNT; k++) shade += c[C_SHADE0 + k];
        atomicAdd(&b.stats[S_RAYS_CLOSEST], shade);  // every traced slot lands in exactly one shade queue
        atomicAdd(&b.stats[S_RAYS_SHADOW], (unsigned long long)c[C_SHADOW]);
        atomicAdd(&b.stats[S_SHADE], shade);
        atomicAdd(&b.stats[S_PATHS_DONE], (unsigned long long)c[C_DONE]);
        c[C_ACTIVE] = (uint32_t)shade;
        for (int k = 0; k < SQ_COUNT; k++) c[C_SHADE0 + k] = 0;
        c[C_SHADOW] = 0;
        c[C_DONE] = 0;
        c[C_CURSOR_TRACE] = 0;
        c[C_CURSOR_SHADOW] = 0;
    }
}

// Ordered accumulation of one pass into the running per-pixel sums; on the last pass divide
// by the sample count and write the three film planes (render.cpp:264-294).
__global__ void __launch_bounds__(256) k_film(WfBuffers b, PassParams pp, float* acc /* 9 floats per owned pixel */,
                                               bool first_pass, bool last_pass, uint32_t n_samples_total, float* color,
                                               float* normal, float* albedo) {
    for (uint32_t pix = blockIdx.x * blockDim.x + threadIdx.x; pix < pp.n_pix; pix += gridDim.x * blockDim.x) {
        float c[3], a[3], n[3];
        if (first_pass) {
            for (int k = 0; k < 3; k++) { c[k] = 0.0f; a[k] = 0.0f; n[k] = 0.0f; }
        } else {
            for (int k = 0; k < 3; k++) {
                c[k] = acc[(size_t)k * pp.n_pix + pix];
                a[k] = acc[(size_t)(3 + k) * pp.n_pix + pix];
                n[k] = acc[(size_t)(6 + k) * pp.n_pix + pix];
            }
        }
        for (uint32_t s = 0; s < pp.s_count; s++) {
            const size_t cell = (size_t)s * pp.n_pix + pix;
            const float4 ra = b.res_a[cell], rb = b.res_b[cell];
            const float rc = b.res_c[cell];
            c[0] += ra.x; c[1] += ra.y; c[2] += ra.z;
            a[0] += rb.x; a[1] += rb.y; a[2] += rb.z;
            n[0] += ra.w; n[1] += rb.w; n[2] += rc;
        }
        if (!last_pass) {
            for (int k = 0; k < 3; k++) {
                acc[(size_t)k * pp.n_pix + pix] = c[k];
                acc[(size_t)(3 + k) * pp.n_pix + pix] = a[k];
                acc[(size_t)(6 + k) * pp.n_pix + pix] = n[k];
            }
        } else {
            const float inv = (float)n_samples_total;
            const uint32_t row = pp.owned_rows[pix / pp.width];
            const size_t o = ((size_t)row * pp.width + pix % pp.width) * 3;
            for (int k = 0; k < 3; k++) {
                color[o + k] = c[k] / inv;
                if (normal) normal[o + k] = n[k] / inv;
                if (albedo) albedo[o + k] = a[k] / inv;
            }
        }
    }
}

}  // namespace qz
