// Wavefront path-tracing pipeline: the reference's per-pixel loop (render.cpp:247-319
// render_pixels, :91-212 sample_pixel) re-cut into stages connected by SoA queues.
//
//   k_generate        camera ray + wavelengths for new pixel-samples (render.cpp:268-273)
//   k_closest_hit     persistent, warp-scheduled BVH traversal (scene.cpp:61-117 / rtcIntersect1);
//                     tags each hit with its material family (k_closest_flat: BVH-less variant for
//                     scenes of <= 96 primitives)
//   k_bin             ordered compaction of the tags into one queue per family x {first hit, later}
//   k_sample          this bounce's Owen-scrambled Halton draws, one thread per dimension
//   k_shade<family>   emission + MIS, BSDF construction, depth-0 albedo, light sampling,
//                     BSDF sampling, throughput update and Russian roulette for ONE family
//                     (diffuse / conductor / dielectric); a fourth "misc" kernel takes misses,
//                     emitter pass-throughs and MixedMaterial hits with run-time dispatch
//   k_shadow          persistent any-hit traversal of the next-event shadow rays
//                     (scene.cpp:136-143); adds the pending contribution when unoccluded
//   k_finish          PixelSensor::to_sensor_rgb of finished paths (sensor.cpp:57-70) into the
//                     per-sample result buffer, then regenerates a new path in the freed slot
//   k_film            ordered per-pixel sum over the sample index (render.cpp:264-294)
//
// Path state lives in structure-of-arrays buffers of 16-byte elements indexed by slot; the
// queues carry 4-byte slot indices.  QUEUES ARE BUILT IN SLOT ORDER: the stages do not append to
// queues with atomics (which scatters neighbouring slots across a queue and turns every 16-byte
// state access into its own 64-byte DRAM burst); they write a one-byte tag per slot, and k_bin
// -- an ordered compaction, one block per 2048 consecutive slots, one atomic per block and queue
// -- turns the tags into queues whose entries ascend within each 2048-slot chunk.  Consecutive
// lanes of the shading kernels then touch neighbouring slots, and the closest-hit stage simply
// walks the slots densely.  Russian roulette stays fused at the end of k_shade: it
// needs the freshly updated throughput and the next sampler dimension, so a separate kernel
// would only re-read what is in registers.
//
// DETERMINISM: a path's arithmetic depends only on (x, y, s); queue order varies from run to
// run but no result depends on it.  Every pixel-sample writes its sensor RGB into its own
// result cell and k_film adds the cells of a pixel in ascending s, exactly the reference's
// summation order, so the film is bit-stable and independent of the pool size, the pass
// size and the number of GPUs.
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
    C_NEXT_PATH = 14,                     // next path id of the pass to hand out
    C_WORDS = 16
};
// 64-bit statistics block
enum Stat { S_RAYS_CLOSEST = 0, S_RAYS_SHADOW = 1, S_SHADE = 2, S_NODES = 3, S_PRIMS = 4, S_PATHS_DONE = 5, S_WORDS = 8 };

struct WfBuffers {
    float4 *ray_o, *ray_d;      // o.xyz | ior_scale ; d.xyz | p_b
    float4 *hit_a, *hit_b;      // t, u, v, primID ; Ng.xyz, geomID (0xffffffff = miss)
    float4 *weight, *radiance, *lambda, *lpdf;
    uint4* misc;                // path id, halton index, dim | depth << 16 | flags << 24, rays issued
    float4 *aov_n, *aov_a;
    float4 *sh_o, *sh_d, *sh_c;
    float* samples;             // R_COUNT floats per slot: this bounce's draws, written by k_sample
    uint8_t *stage, *fam, *post;  // per-slot tags (StageTag, family queue id or QZ_FAM_NONE, QZ_POST_* bits)
    uint32_t* q_shade[SQ_COUNT];
    uint32_t *q_shadow, *q_done;  // (adjacent in memory order is not required)
    uint32_t* counters;
    unsigned long long* stats;
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
__global__ void __launch_bounds__(256) k_generate(DScene sc, DCamera cam, WfBuffers b, PassParams pp, uint32_t n) {
    for (uint32_t i = n + blockIdx.x * blockDim.x + threadIdx.x; i < b.pool; i += gridDim.x * blockDim.x) b.stage[i] = ST_EMPTY;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        init_slot(sc, cam, b, pp, i, i);
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

#define QZ_REFILL_MIN 8   /* idle lanes that trigger a refill from the queue */
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
    __shared__ F4 s_prims[QZ_FLAT_MAX_PRIMS * 4];
    const uint32_t n_prims = sc.n_prims;
    for (uint32_t i = threadIdx.x; i < n_prims * 4; i += blockDim.x) s_prims[i] = sc.prims[i];
    __syncthreads();
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < b.pool; slot += gridDim.x * blockDim.x) {
        const uint8_t st = b.stage[slot];
        if (st == ST_EMPTY) { b.fam[slot] = QZ_FAM_NONE; continue; }
        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
        const V3 O = v3(o.x, o.y, o.z), D = v3(d.x, d.y, d.z);
        Hit best;
        best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
        best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
#pragma unroll 1
        for (uint32_t p = 0; p < n_prims; p++)
            prim_test_rec(sc, s_prims[4 * p], s_prims[4 * p + 1], s_prims[4 * p + 2], s_prims[4 * p + 3], p, O, D, QZ_TNEAR, INFINITY, best);
        finish_closest(sc, b, slot, st == ST_TRACE_FIRST, best, flags);
    }
}

__global__ void __launch_bounds__(256) k_shadow_flat(DScene sc, WfBuffers b) {
    __shared__ F4 s_prims[QZ_FLAT_MAX_PRIMS * 4];
    const uint32_t n_prims = sc.n_prims;
    for (uint32_t i = threadIdx.x; i < n_prims * 4; i += blockDim.x) s_prims[i] = sc.prims[i];
    __syncthreads();
    const uint32_t count = b.counters[C_SHADOW];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t slot = b.q_shadow[i];
        const float4 o = b.sh_o[slot], d = b.sh_d[slot];
        const V3 O = v3(o.x, o.y, o.z), D = v3(d.x, d.y, d.z);
        Hit best;
        best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
        best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
        // occluded iff the closest hit has t <= 1 (scene.cpp:136-143)
#pragma unroll 1
        for (uint32_t p = 0; p < n_prims; p++)
            prim_test_rec(sc, s_prims[4 * p], s_prims[4 * p + 1], s_prims[4 * p + 2], s_prims[4 * p + 3], p, O, D, QZ_TNEAR, INFINITY, best);
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
    uint32_t roles[SQ_COUNT], nrole[SQ_COUNT];
    start[0] = 0;
#pragma unroll
    for (int q = 0; q < SQ_COUNT; q++) {
        roles[q] = queue_roles(q, sc.n_lights);
        nrole[q] = __popc(roles[q]);
        start[q + 1] = start[q] + b.counters[C_SHADE0 + q] * nrole[q];
    }
    const uint32_t total = start[SQ_COUNT];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int q = 0;
#pragma unroll
        for (int k = 1; k < SQ_COUNT; k++) q += (t >= start[k]) ? 1 : 0;
        const uint32_t local = t - start[q];
        const uint32_t entry = local / nrole[q], r = local % nrole[q];
        // r-th set bit of the role mask
        uint32_t m = roles[q];
        for (uint32_t k = 0; k < r; k++) m &= m - 1;
        const int role = __ffs(m) - 1;
        const uint32_t slot = b.q_shade[q][entry];
        const uint4 misc = b.misc[slot];
        const int fam = q % SQ_FAMILIES;
        bool nee = fam == SQ_DIFFUSE;
        if (fam == SQ_CONDUCTOR) {
            const uint32_t geom = __float_as_uint(b.hit_b[slot].w);
            const qz_material mat = sc.materials[sc.geoms[geom].material];
            nee = !(mat.alpha_x < 1e-3f && mat.alpha_y < 1e-3f);
        }
        uint32_t dims[R_COUNT];
        bounce_dims(misc.z & 0xffffu, nee, sc.n_lights != 0, dims);
        Sampler smp;
        smp.index = misc.y; smp.dim = 0;
        b.samples[(size_t)slot * R_COUNT + role] = sample_dimension(sc.sampler_table, smp, dims[role]);
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
    constexpr bool kPdf = KH == KH_DIELECTRIC || KH == KH_ANY;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t slot = queue[i];
        PathState ps;
        const float4 o = b.ray_o[slot], d = b.ray_d[slot];
        ps.ray.o = v3(o.x, o.y, o.z); ps.ior_scale = o.w;
        ps.ray.d = v3(d.x, d.y, d.z); ps.p_b = d.w;
        ps.weight = s4(b.weight[slot]);
        ps.lambda = s4(b.lambda[slot]);
        ps.pdf = kPdf ? s4(b.lpdf[slot]) : spec4(0.0f);
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
            const float4* sv = reinterpret_cast<const float4*>(b.samples + (size_t)slot * R_COUNT);
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
            if (ps.depth != 0 || !alive) b.aov_a[slot] = f4(aov.albedo);
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
            b.weight[slot] = f4(ps.weight);
        }
        if (kPdf) b.lpdf[slot] = f4(ps.pdf);
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
        if (lane == leader) base = atomicAdd(&b.counters[C_NEXT_PATH], (uint32_t)__popc(peers));
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
        b.stats[S_RAYS_CLOSEST] += shade;  // every traced slot lands in exactly one shade queue
        b.stats[S_RAYS_SHADOW] += c[C_SHADOW];
        b.stats[S_SHADE] += shade;
        b.stats[S_PATHS_DONE] += c[C_DONE];
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
