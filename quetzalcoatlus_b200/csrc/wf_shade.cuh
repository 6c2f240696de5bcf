// Arithmetic-bearing stages of the wavefront (wf_types.cuh).  This header is compiled twice -- csrc/k_shade_exact.cu
// and csrc/k_shade_fast.cu, the two ARITHMETIC MODES of common.cuh -- and the host picks one set per render call.
#pragma once

#include "wf_types.cuh"

namespace qz {

__device__ __forceinline__ void store_state(const WfBuffers& b, uint32_t slot, const PathState& ps, uint32_t path_id) {
    b.ray_o.set(slot, f4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ior_scale));
    b.ray_d.set(slot, f4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, ps.p_b));
    b.weight.set(slot, f4(ps.weight));
    b.radiance.set(slot, f4(ps.L));
    b.lpdf.set(slot, f4(ps.pdf));
    b.misc.set(slot, pack_misc(path_id, ps));
}

// path id of the pass -> pixel and sample; initialises the slot (render.cpp:261-273)
__device__ __forceinline__ void init_slot(const DScene& sc, const DCamera& cam, const WfBuffers& b, const PassParams& pp,
                                          uint32_t slot, uint32_t path_id, float4& ray_o, float4& ray_d) {
    const uint32_t pix = path_id % pp.n_pix;
    const uint32_t s = pp.s_begin + path_id / pp.n_pix;
    const uint32_t row = pp.owned_rows[pix / pp.width];
    const uint32_t x = pix % pp.width;
    const uint32_t y = pp.height - row - 1;
    PathState ps;
    PathAov aov;
    start_path(sc, cam, pp.spar, x, y, s, ps, aov);
    store_state(b, slot, ps, path_id);
    b.lambda.set(slot, f4(ps.lambda));
    b.aov_n.set(slot, f4(0.0f, 0.0f, 0.0f, 0.0f));
    b.aov_a.set(slot, f4(0.0f, 0.0f, 0.0f, 0.0f));
    ray_o = f4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ior_scale);
    ray_d = f4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, ps.p_b);
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256) k_generate(DScene sc, DCamera cam, WfBuffers b, PassParams pp, uint32_t first_id, uint32_t n) {
    // (the work list of the first iteration -- slots 0 .. n-1 in queue 0 -- gets its length from the host's counter block)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < b.pool; i += gridDim.x * blockDim.x) {
        b.post[i] = 0;
        b.fam[i] = QZ_FAM_NONE;
        if (i < n) {
            float4 o, d;
            init_slot(sc, cam, b, pp, i, first_id + i, o, d);
            b.stage[i] = (uint8_t)(ST_TRACE_FIRST | (first_bounce_in_memo(sc.memo) ? 0 : ST_LATE));
            b.q_shade[0][i] = i;
        } else {
            b.stage[i] = ST_EMPTY;
        }
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
            const float4 o = b.ray_o.get(slot), d = b.ray_d.get(slot), ha = b.hit_a.get(slot), hb = b.hit_b.get(slot);
            Ray ray;
            ray.o = v3(o.x, o.y, o.z); ray.d = v3(d.x, d.y, d.z);
            Hit hit;
            hit.t = ha.x; hit.u = ha.y; hit.v = ha.z; hit.prim_id = __float_as_uint(ha.w);
            hit.ng = v3(hb.x, hb.y, hb.z); hit.geom_id = __float_as_uint(hb.w); hit.prim = hit.geom_id; hit.key = 0;
            const Spec4 lambda = s4(b.lambda.get(slot));
            sp = make_surface_point(sc, ray, hit);
            // Material::bsdf for a conductor (material.cpp:16-20) is eight spectrum lookups (eta and k at four
            // wavelengths): lane j of the group does lookup j, the group exchanges them below
            mat = sc.materials[sp.material];
            const int j = i & 7;
            const float lam = j & 2 ? (j & 1 ? lambda.v[3] : lambda.v[2]) : (j & 1 ? lambda.v[1] : lambda.v[0]);
            {
                const int32_t sid = j < 4 ? mat.a : mat.b;
                const int hot = sc.memo.n_hot ? memo_hot_slot(sc.memo, sid) : -1;
                const uint32_t* row = hot >= 0 ? memo_row(sc.memo, b.misc.get(slot).y) : nullptr;
                if (row) coeff = __uint_as_float(__ldcg(row + 4 * hot + (j & 3)));
                else coeff = eval_spectrum_rec<true>(sc, load_spectrum(sc, sid), lam);
            }
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
            if (valid) term = r_div(smp.spec * fabsf(smp.wi.z), smp.pdf);
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
        if (live && i == 0) b.aov_a.set(slot, f4(acc / 16.0f));
    }
}

// Fills the hot-spectra block of the memo rows (sampler.cuh): one thread per (sample number, pixel class) derives the
// row's wavelengths from its dimension-2 value (already filled by k_memo_fill) exactly as start_path does and evaluates
// every hot spectrum with from_spectrum -- the function the shading code would otherwise call per bounce.
__global__ void __launch_bounds__(256) k_memo_spectra(DScene sc) {
    const SampleMemo& memo = sc.memo;
    const uint64_t total = (uint64_t)memo.s_count * memo.n_cls;
    for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (uint64_t)gridDim.x * blockDim.x) {
        memo_fill_hot(sc, memo.tab + (size_t)t * memo.stride);
    }
}

#define QZ_REC_GET get_l1   /* k_shade reads the hot line it prefetched into L1 */
#ifndef QZ_SHADE_MIN_BLOCKS_LIGHT
#define QZ_SHADE_MIN_BLOCKS_LIGHT 4   /* diffuse / dielectric shade kernels: 128 registers (5 blocks = 96 registers spills: -12 % on the stage) */
#endif
#ifndef QZ_SHADE_MIN_BLOCKS_HEAVY
#define QZ_SHADE_MIN_BLOCKS_HEAVY 4   /* conductor / run-time-dispatch shade kernels */
#endif

// One bounce for every path of one material family: its first-hit queue, then its later-bounce queue, as one index
// space (a warp straddles the boundary at most once per kernel, so `first` is warp-uniform in practice).  Only
// what the family can change is loaded and stored: the radiance buffer is touched only when the hit itself adds
// radiance (emitters and misses live in the run-time-dispatch queue), the AOVs only at first hits.
template <int KH>
__global__ void __launch_bounds__(128, (KH == KH_DIFFUSE || KH == KH_DIELECTRIC) ? QZ_SHADE_MIN_BLOCKS_LIGHT : QZ_SHADE_MIN_BLOCKS_HEAVY)
k_shade(DScene sc, WfBuffers b, uint32_t max_bounces) {
    // (tried: a second, lean kernel over the run-time-dispatch queue for the hits without a material -- misses, emitter
    // geometry -- next to the MixedMaterial one: 80 registers instead of 128, but one more launch and two walks over the
    // queue; shading stage 45.1 -> 47.0 ms on cornell_box, 13.1 -> 16.6 on mandelbrot: dropped)
    constexpr int FAM = KH == KH_DIFFUSE ? SQ_DIFFUSE : (KH == KH_CONDUCTOR ? SQ_CONDUCTOR : (KH == KH_DIELECTRIC ? SQ_DIELECTRIC : SQ_MISC));
    const uint32_t n_first = b.counters[C_SHADE0 + SQ_FAMILIES + FAM];
    const uint32_t count = n_first + b.counters[C_SHADE0 + FAM];
    const uint32_t* q_first = b.q_shade[SQ_FAMILIES + FAM];
    const uint32_t* q_later = b.q_shade[FAM];
    auto entry = [&](uint32_t i) -> uint32_t { return i < n_first ? q_first[i] : q_later[i - n_first]; };
    // A bounce is: read one record, many dependent instructions, write it back -- nothing in it overlaps the
    // read.  So the read of the thread's NEXT path is started a whole bounce early: its queue entry is loaded two
    // trips ahead, its record lines prefetched one trip ahead.
    const uint32_t stride = gridDim.x * blockDim.x;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot_cur = i < count ? entry(i) : 0u;
    uint32_t slot_next = i + stride < count ? entry(i + stride) : 0u;
    for (; i < count; i += stride) {
        const uint32_t slot = slot_cur;
        const uint32_t slot_after = (i + 2u * stride < count && i + 2u * stride >= i) ? entry(i + 2u * stride) : 0u;
        // (tried: loading the next path's misc word here and prefetching its memo row at the bottom of the trip -- the
        // extra live registers spilled and the stage got 16 % slower, profiles/r02_summary.md)
        // (the side line is mostly written -- shadow request, radiance -- yet fetching it ahead pays: without it the
        // partial-line stores wait for their fill, step 93.2 -> 95.3 ms)
        // The hot line goes to L1 and is then read through L1 (RecField::get_l1): a slot is read by one thread per
        // kernel and L1 starts every kernel empty, so the line cannot be stale (+2 % on the step over an L2 prefetch).
        if (i + stride < count) {
            prefetch_line_l1(b.ray_o.at(slot_next));
            if (KH != KH_ANY) prefetch_line(b.samples.at(slot_next));
        }
        slot_cur = slot_next;
        slot_next = slot_after;
        const bool first = i < n_first;
        PathState ps;
        const float4 o = b.ray_o.QZ_REC_GET(slot), d = b.ray_d.QZ_REC_GET(slot);
        ps.ray.o = v3(o.x, o.y, o.z); ps.ior_scale = o.w;
        ps.ray.d = v3(d.x, d.y, d.z); ps.p_b = d.w;
        ps.weight = s4(b.weight.QZ_REC_GET(slot));
        ps.lambda = s4(b.lambda.QZ_REC_GET(slot));
        // only a dispersive dielectric changes the wavelength pdf (terminate_secondary): the other families neither load nor store it
        constexpr bool TOUCHES_PDF = KH == KH_DIELECTRIC || KH == KH_ANY;
        ps.pdf = TOUCHES_PDF ? s4(b.lpdf.QZ_REC_GET(slot)) : spec4(0.0f);
        ps.L = spec4(0.0f);
        const uint4 m = b.misc.QZ_REC_GET(slot);
        unpack_misc(m, ps);
        ps.n_rays += 1;  // + the closest-hit query that produced this hit
        const float4 ha = b.hit_a.QZ_REC_GET(slot), hb = b.hit_b.QZ_REC_GET(slot);
        Hit hit;
        hit.t = ha.x; hit.u = ha.y; hit.v = ha.z; hit.prim_id = __float_as_uint(ha.w);
        hit.ng = v3(hb.x, hb.y, hb.z);
        hit.geom_id = __float_as_uint(hb.w);
        hit.prim = hit.geom_id;  // only compared against QZ_NO_HIT from here on
        hit.key = 0;
        PathAov aov;
        aov.normal = v3(0.0f, 0.0f, 0.0f);
        aov.albedo = spec4(0.0f);
        ShadowRequest sh;
        Spec4 gain;
        bool has_gain, alive;
        if (KH == KH_ANY) {
            SamplesOnTheFly src;
            src.tab = sc.sampler_table; src.index = ps.smp.index; src.memo = sc.memo;
            alive = shade_bounce<KH, -1, true>(sc, ps, aov, hit, max_bounces, sh, src, gain, has_gain);
        } else {
            SamplesPrecomputed src;
            const uint32_t d0 = m.z & 0xffffu;
            if (bounce_in_memo(sc.memo, m.y, d0)) {
                // this bounce's draws sit in the path's memo row (sampler.cuh), one 32-byte sector: no sampler stage
                bool nee = KH == KH_DIFFUSE;
                if (KH == KH_CONDUCTOR) {
                    const qz_material mat = sc.materials[sc.geoms[hit.geom_id].material];
                    nee = !(mat.alpha_x < 1e-3f && mat.alpha_y < 1e-3f);
                }
                uint32_t dims[R_COUNT];
                bounce_dims(d0, nee, sc.n_lights != 0, dims);
                // (tried: prefetching the row's hot-spectra line into L1 here -- no change)
                const uint32_t* row = memo_row(sc.memo, m.y) + sc.memo.dim_off;
#pragma unroll
                for (int k = 0; k < R_COUNT; k++) src.v[k] = 0.0f;
                if (KH == KH_DIELECTRIC) {
                    src.v[R_U1] = __uint_as_float(__ldcg(row + dims[R_U1]));
                } else {
                    if (sc.n_lights > 1) src.v[R_PICK] = __uint_as_float(__ldcg(row + dims[R_PICK]));
                    src.v[R_LIGHT] = __uint_as_float(__ldcg(row + dims[R_LIGHT]));
                    src.v[R_LIGHT + 1] = __uint_as_float(__ldcg(row + dims[R_LIGHT + 1]));
                    src.v[R_BSDF] = __uint_as_float(__ldcg(row + dims[R_BSDF]));
                    src.v[R_BSDF + 1] = __uint_as_float(__ldcg(row + dims[R_BSDF + 1]));
                }
                src.v[R_RR] = __uint_as_float(__ldcg(row + dims[R_RR]));
            } else {
                const float4* sv = b.samples.at(slot);
                const float4 s0 = __ldcg(sv), s1 = __ldcg(sv + 1);
                src.v[0] = s0.x; src.v[1] = s0.y; src.v[2] = s0.z; src.v[3] = s0.w;
                src.v[4] = s1.x; src.v[5] = s1.y; src.v[6] = s1.z; src.v[7] = s1.w;
            }
            alive = shade_bounce<KH, -1, true>(sc, ps, aov, hit, max_bounces, sh, src, gain, has_gain);
        }
        if (has_gain) b.radiance.set(slot, f4(s4(b.radiance.get(slot)) + gain));
        if (first) {
            // depth is still 0 after an emitter pass-through, so these may be written more than
            // once per path; the last write (the first real surface) wins, as in the reference
            if (hit.prim != QZ_NO_HIT) b.aov_n.set(slot, f4(aov.normal.x, aov.normal.y, aov.normal.z, 0.0f));
            // (conductors: the albedo comes from k_albedo_conductor)
            if (KH != KH_CONDUCTOR && (ps.depth != 0 || !alive)) b.aov_a.set(slot, f4(aov.albedo));
        }
        uint32_t post = alive ? 0u : QZ_POST_DONE;
        if (ps.flags & QZ_FLAG_HAS_SHADOW) {
            ps.n_rays++;
            b.sh_o.set(slot, f4(sh.o.x, sh.o.y, sh.o.z, 0.0f));
            b.sh_d.set(slot, f4(sh.d.x, sh.d.y, sh.d.z, 0.0f));
            b.sh_c.set(slot, f4(sh.contrib));
            post |= QZ_POST_SHADOW;
        }
        if (alive) {
            b.ray_o.set(slot, f4(ps.ray.o.x, ps.ray.o.y, ps.ray.o.z, ps.ior_scale));
            b.ray_d.set(slot, f4(ps.ray.d.x, ps.ray.d.y, ps.ray.d.z, ps.p_b));
            // weight and lambda share a sector: writing both keeps it a full-sector store
            b.weight.set(slot, f4(ps.weight));
            b.lambda.set(slot, f4(ps.lambda));
        }
        if (TOUCHES_PDF) b.lpdf.set(slot, f4(ps.pdf));
        b.misc.set(slot, pack_misc(m.x, ps));
        b.post[slot] = (uint8_t)post;
        // (a path at depth 0 after an emitter pass-through is a first hit again; a continued path whose next bounce lies
        // beyond the memo row goes through the sampler stage)
        const bool late = !bounce_in_memo(sc.memo, ps.smp.index, ps.smp.dim);
        b.stage[slot] = alive ? (uint8_t)((ps.depth == 0 ? ST_TRACE_FIRST : ST_TRACE) | (late ? ST_LATE : 0)) : (uint8_t)ST_EMPTY;
    }
}

// A finished path: sensor conversion into its result cell (sensor.cpp:57-70), then the next pixel-sample of the
// pass in the same slot.  Returns the slot's new stage tag; a regenerated path's ray comes back in (ray_o, ray_d).
__device__ __forceinline__ uint8_t finish_and_regenerate(const DScene& sc, const DCamera& cam, const WfBuffers& b, const PassParams& pp,
                                                         uint32_t slot, const Spec4& L, float4& ray_o, float4& ray_d) {
    const uint32_t path_id = b.misc.get(slot).x;
    const Spec4 lambda = s4(b.lambda.get(slot)), pdf = s4(b.lpdf.get(slot));
    const float4 n = b.aov_n.get(slot);
    const V3 rgb = to_sensor_rgb(cam, L, lambda, pdf);
    const V3 argb = to_sensor_rgb(cam, s4(b.aov_a.get(slot)), lambda, pdf);
    __stcs(b.res_a + path_id, f4(rgb.x, rgb.y, rgb.z, n.x));
    __stcs(b.res_b + path_id, f4(argb.x, argb.y, argb.z, n.y));
    __stcs(b.res_c + path_id, n.z);
    // regenerate
    const unsigned peers = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(b.next_path, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const uint32_t next_id = base + __popc(peers & ((1u << lane) - 1u));
    if (next_id >= pp.total) return ST_EMPTY;
    init_slot(sc, cam, b, pp, slot, next_id, ray_o, ray_d);
    return (uint8_t)(ST_TRACE_FIRST | (first_bounce_in_memo(sc.memo) ? 0 : ST_LATE));
}

// BVH scenes: the finish stage between the two traversal kernels.  Walks the work list, consumes the post tags.
__global__ void __launch_bounds__(256) k_finish(DScene sc, DCamera cam, WfBuffers b, PassParams pp) {
    if (blockIdx.x == 0 && threadIdx.x < SQ_COUNT) b.counters[C_SHADE0 + threadIdx.x] = 0;   // consumed by the shading stage; k_bin refills them
    if (blockIdx.x == 0 && threadIdx.x == SQ_COUNT) b.counters[C_LATE] = 0;
    WorkList wl;
    wl.load(b.counters);
    const uint32_t n_work = wl.total();
    uint32_t n_done = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_work; i += gridDim.x * blockDim.x) {
        const uint32_t slot = wl.slot(b.q_shade, i);
        const uint8_t po = b.post[slot];
        if (!po) continue;
        if (po & QZ_POST_DONE) {
            float4 o, d;
            const uint8_t st = finish_and_regenerate(sc, cam, b, pp, slot, s4(b.radiance.get(slot)), o, d);
            if (st != ST_EMPTY) b.stage[slot] = st;
            n_done++;
        }
        b.post[slot] = 0;
    }
    stat_add(&b.stats[S_PATHS_DONE], n_done);
}

// Flat scenes (<= QZ_FLAT_MAX_PRIMS primitives: every shipped analytic scene): no BVH and no separate shadow /
// finish / closest-hit kernels.  ONE kernel does, per slot of the work list (wf_types.cuh), what its tags ask for:
//   post & SHADOW : the next-event shadow test of the last bounce (scene.cpp:136-143), radiance += contribution;
//   post & DONE   : finish the path and start the next pixel-sample of the pass in the slot;
//   stage != EMPTY: closest hit of the slot's ray (scene.cpp:61-117) and the family tag for k_bin.
// The primitive records are staged once per CTA in shared memory together with each triangle's edges and normal
// (make_flat_prim: the expressions tri_test uses), and every lane walks the same list in lockstep -- no stack, no
// divergence inside a test, broadcast shared-memory reads.  The answer is the brute-force minimum under the (t, key)
// order, i.e. exactly what the BVH traversal is defined to return.
//
// A CTA works on TILES of QZ_FLAT_TILE consecutive work-list entries (drawn from an atomic cursor) in three dense
// phases, because the three jobs are needed by different subsets of the lanes:
//   1. shadow tests; slots whose path is finished are collected in a shared-memory list;
//   2. that list, densely: sensor conversion, result cell, new path -- the sampler's integer chains of a new path
//      (Halton index, pixel jitter, the wavelength draw) used to run inside the per-slot pass with a third of the
//      lanes (10 of 32 per instruction, a quarter of the kernel's instructions: profiles/r02_summary.md);
//   3. closest hit for every slot of the tile that carries a ray.
// Shadow rays ask "is there any hit with t <= 1" (the closest hit has t <= 1 iff some hit has): emitter primitives are
// staged first and a lane leaves the loop at its first occluder, so the rays that end on the light's own geometry
// (all of them for an axis-aligned emitter) test one primitive instead of all.
#ifndef QZ_FLAT_TILE
#define QZ_FLAT_TILE 1024
#endif

// any valid hit with t <= 1?  The validity conditions are tri_test_pre's; T / absDen <= 1 iff T <= absDen (round to
// nearest cannot carry a quotient > 1 down to 1: that needs ulp(absDen) / absDen == 2^-24 exactly), so no division.
__device__ __forceinline__ bool tri_occludes(V3 O, V3 D, float tnear, V3 v0, V3 e1, V3 e2, V3 ng) {
    const V3 C = v0 - O;
    const V3 R = cross(C, D);
    const float den = dot(ng, D);
    const float absDen = fabsf(den);
    const uint32_t s = float_as_u32(den) & 0x80000000u;
    const float U = xor_sign(dot(R, e2), s);
    const float V = xor_sign(dot(R, e1), s);
    if (!(den != 0.0f) || !(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= absDen)) return false;
    const float T = xor_sign(dot(ng, C), s);
    return absDen * tnear < T && T <= absDen;
}
__device__ __forceinline__ bool flat_prim_occludes(const FlatPrim& f, V3 O, V3 D, float rd2) {
    const uint32_t kind = prim_kind(float_as_u32(f.e1a.w));
    if (kind == QZ_PRIM_SPHERE) {
        PrimHit h;
        return sphere_test_rd2(O, D, rd2, QZ_TNEAR, INFINITY, xyz(f.a), f.b.x, h) && h.t <= 1.0f;
    }
    if (tri_occludes(O, D, QZ_TNEAR, xyz(f.a), xyz(f.e1a), xyz(f.e2a), xyz(f.nga))) return true;
    return kind != QZ_PRIM_TRIANGLE && f.c.w != 0.0f &&
           tri_occludes(O, D, QZ_TNEAR, xyz(f.c), xyz(f.e1b), xyz(f.e2b), v3(f.nga.w, f.e1b.w, f.e2b.w));
}

__global__ void __launch_bounds__(128, 4) k_step_flat(DScene sc, DCamera cam, WfBuffers b, PassParams pp, uint32_t flags) {
    __shared__ FlatPrim s_prims[QZ_FLAT_MAX_PRIMS];
    __shared__ uint8_t s_emit[QZ_FLAT_MAX_PRIMS];
    __shared__ uint32_t s_slot[QZ_FLAT_TILE], s_done[QZ_FLAT_TILE];
    __shared__ uint32_t s_ndone, s_tile;
    __shared__ WorkList wl;
    const uint32_t n_prims = sc.n_prims;
    // stage the primitives, emitter geometry first (stable partition; the closest hit does not depend on the order)
    for (uint32_t i = threadIdx.x; i < n_prims; i += blockDim.x)
        s_emit[i] = sc.geoms[float_as_u32(sc.prims[4 * i].w)].light >= 0 ? 1 : 0;
    if (threadIdx.x == 0) wl.load(b.counters);
    if (blockIdx.x == 0 && threadIdx.x < SQ_COUNT) b.counters[C_SHADE0 + threadIdx.x] = 0;   // consumed by the shading stage; k_bin refills them
    if (blockIdx.x == 0 && threadIdx.x == SQ_COUNT) b.counters[C_LATE] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_prims; i += blockDim.x) {
        uint32_t before_same = 0, n_emit = 0;
        for (uint32_t j = 0; j < n_prims; j++) {
            n_emit += s_emit[j];
            if (j < i && s_emit[j] == s_emit[i]) before_same++;
        }
        const uint32_t pos = s_emit[i] ? before_same : n_emit + before_same;
        s_prims[pos] = make_flat_prim(sc.prims[4 * i], sc.prims[4 * i + 1], sc.prims[4 * i + 2], sc.prims[4 * i + 3]);
    }
    const uint32_t n_work = wl.total();
    uint32_t n_closest = 0, n_shadow = 0, n_done = 0;
    for (;;) {
        __syncthreads();   // (first trip: the staged primitives; later trips: everyone is done with the previous tile's lists)
        if (threadIdx.x == 0) { s_tile = atomicAdd(&b.counters[C_CURSOR_TRACE], 1u); s_ndone = 0; }
        __syncthreads();
        const uint32_t tile = s_tile * (uint32_t)QZ_FLAT_TILE;
        if (tile >= n_work) break;
        const uint32_t tile_n = min((uint32_t)QZ_FLAT_TILE, n_work - tile);
        // ---- phase 1: shadow tests, collect the finished paths
        // (tried: the tags of all of a thread's entries loaded up front and the next shadow request loaded before the
        // current primitive loop, as in phase 3 -- 105 registers, one CTA per SM fewer, step 93.2 -> 97.7 ms)
        for (uint32_t k = threadIdx.x; k < tile_n; k += blockDim.x) {
            const uint32_t slot = wl.slot(b.q_shade, tile + k);
            s_slot[k] = slot;
            const uint8_t po = b.post[slot];
            if (!po) continue;
            if (po & QZ_POST_SHADOW) {
                n_shadow++;
                const float4 o = b.sh_o.get(slot), d = b.sh_d.get(slot);
                const V3 O = v3(o.x, o.y, o.z), D = v3(d.x, d.y, d.z);
                const float rd2 = 1.0f / dot(D, D);   // the ray's share of every sphere test
                bool occl = false;
#pragma unroll 1
                for (uint32_t p = 0; p < n_prims && !occl; p++) occl = flat_prim_occludes(s_prims[p], O, D, rd2);
                if (!occl) b.radiance.set(slot, f4(s4(b.radiance.get(slot)) + s4(b.sh_c.get(slot))));
            }
            if (po & QZ_POST_DONE) s_done[atomicAdd(&s_ndone, 1u)] = slot;
            b.post[slot] = 0;
        }
        __syncthreads();
        // ---- phase 2: finished paths -> result cell, next pixel-sample of the pass in the slot
        const uint32_t nd = s_ndone;
        for (uint32_t k = threadIdx.x; k < nd; k += blockDim.x) {
            const uint32_t slot = s_done[k];
            float4 ro, rd;
            const uint8_t st = finish_and_regenerate(sc, cam, b, pp, slot, s4(b.radiance.get(slot)), ro, rd);
            if (st != ST_EMPTY) b.stage[slot] = st;
            n_done++;
        }
        __syncthreads();
        // ---- phase 3: closest hit of every slot of the tile that carries a ray
        // (the next entry's tag and ray are loaded before this entry's primitive loop: the loop hides their latency)
        uint32_t slot_n = 0;
        uint8_t st_n = ST_EMPTY;
        float4 ro_n = f4(0.0f, 0.0f, 0.0f, 0.0f), rd_n = ro_n;
        if (threadIdx.x < tile_n) {
            slot_n = s_slot[threadIdx.x];
            st_n = b.stage[slot_n];
            ro_n = b.ray_o.get(slot_n); rd_n = b.ray_d.get(slot_n);
        }
        for (uint32_t k = threadIdx.x; k < tile_n; k += blockDim.x) {
            const uint32_t slot = slot_n;
            const uint8_t st = st_n;
            const float4 ro = ro_n, rd = rd_n;
            if (k + blockDim.x < tile_n) {
                slot_n = s_slot[k + blockDim.x];
                st_n = b.stage[slot_n];
                ro_n = b.ray_o.get(slot_n); rd_n = b.ray_d.get(slot_n);
            }
            if (st == ST_EMPTY) { b.fam[slot] = QZ_FAM_NONE; continue; }
            n_closest++;
            const V3 O = v3(ro.x, ro.y, ro.z), D = v3(rd.x, rd.y, rd.z);
            const float rd2 = 1.0f / dot(D, D);
            // (per candidate only the distance; u, v, Ng of the winner once per ray: intersect.cuh, FlatBest)
            FlatBest fbest;
            fbest.t = INFINITY; fbest.U = 0.0f; fbest.V = 0.0f; fbest.absDen = 0.0f; fbest.info = 0xffffffffu; fbest.key = 0xffffffffu;
#pragma unroll 1
            for (uint32_t p = 0; p < n_prims; p++)
                flat_prim_test_lazy(s_prims[p], p, O, D, rd2, QZ_TNEAR, INFINITY, fbest);
            Hit best;
            flat_best_finish(sc, s_prims, fbest, O, D, rd2, QZ_TNEAR, INFINITY, best);
            b.hit_a.set(slot, f4(best.t, best.u, best.v, __uint_as_float(best.prim_id)));
            b.hit_b.set(slot, f4(best.ng.x, best.ng.y, best.ng.z, __uint_as_float(best.geom_id)));
            const bool unsorted = (flags & QZ_FLAG_UNSORTED_SHADING) != 0;
            int fam = SQ_MISC;
            if (!unsorted && best.geom_id != QZ_NO_HIT) {
                const int32_t mat = sc.geoms[best.geom_id].material;
                if (mat >= 0) {
                    const uint32_t kind = sc.materials[mat].kind;
                    fam = kind == QZ_MAT_DIFFUSE ? SQ_DIFFUSE : (kind == QZ_MAT_CONDUCTOR ? SQ_CONDUCTOR
                          : ((kind == QZ_MAT_DIELECTRIC || kind == QZ_MAT_THIN_DIELECTRIC) ? SQ_DIELECTRIC : SQ_MISC));
                }
            }
            b.fam[slot] = (uint8_t)(fam + ((st & ST_FIRST) && !unsorted ? SQ_FAMILIES : 0) + ((st & ST_LATE) ? QZ_FAM_LATE : 0u));
        }
    }
    stat_add(&b.stats[S_RAYS_CLOSEST], n_closest);
    stat_add(&b.stats[S_RAYS_SHADOW], n_shadow);
    stat_add(&b.stats[S_PATHS_DONE], n_done);
}

// per-path replay (qz_trace_paths): a whole path in one thread, no wavefront
__global__ void k_trace_paths(DScene sc, DCamera cam, SamplerParams spar, uint32_t max_bounces, uint32_t n,
                              const int32_t* xys, float* records) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PathState ps;
    PathAov aov;
    Spec4 lambda0;
    run_path<false>(sc, cam, spar, (uint32_t)xys[3 * i], (uint32_t)xys[3 * i + 1], (uint32_t)xys[3 * i + 2], max_bounces, ps,
                    aov, lambda0, nullptr);
    V3 rgb = to_sensor_rgb(cam, ps.L, ps.lambda, ps.pdf);
    V3 argb = to_sensor_rgb(cam, aov.albedo, ps.lambda, ps.pdf);
    write_trace_record(records + (size_t)i * 32, ps, aov, lambda0, rgb, argb);
}

}  // namespace qz
