// Persistent, phase-scheduled BVH traversal (replaces rtcIntersect1 behind scene.cpp:56 and Scene::occluded,
// scene.cpp:136-143).
//
// One ray per lane, but the WARP decides what every lane does next.  A lane is in one of four states -- NODE (has a
// wide node to open), PRIM (has leaf primitives of the last node pending), DONE (result to write), IDLE (no ray) --
// and each trip of the persistent loop runs exactly ONE of three instruction streams for all lanes that are in the
// matching state:
//   * finish + refill, when >= QZ_REFILL_MIN lanes wait for it (warp vote), or nothing else can run;
//   * the primitive stream (one primitive per pending lane), when more lanes are in PRIM than in NODE;
//   * the node stream otherwise: one 128-byte node per lane, all eight child slabs tested branch-free, hit leaf
//     children recorded as a bit mask over the node's contiguous leaf block (NOT tested here), hit internal children
//     sorted far-to-near by a register sorting network and pushed; then the lane pops its next node, culling against
//     its current best.
// So the long, expensive streams (8 slab tests; Moeller-Trumbore) always run with the majority of the warp's lanes,
// instead of every lane dragging the other 31 through its own private node / primitive / pop sequence (the first
// version of this kernel: 6.9 of 32 lanes active per instruction on a 1M-triangle scene, profiles/r01_summary.md).
// Results are unchanged: same slab and primitive arithmetic, closest hit = minimum under (t, key), which does not
// depend on the visiting order.
//
// STACK.  The first QZ_SMEM_STACK entries of a lane's stack live in SHARED memory, laid out entry-major
// (s_stack[entry][thread]: a warp's accesses to one level are 256 contiguous bytes, and lanes at different levels
// still hit different banks); only deeper excursions spill to the per-lane local-memory array behind it.  An
// 8-wide tree over a million triangles is ~7 levels deep and a lane rarely holds more than a dozen entries, so the
// local array -- 1 KB per lane that used to be the stack and competed with nodes and primitives for L1 -- is
// untouched on almost every ray.  The kernel draws from the iteration's work list (wf_types.cuh); the per-slot tag
// says which of its slots carry a ray for this stage.
#pragma once

#include "wf_types.cuh"

namespace qz {

#ifndef QZ_REFILL_MIN
#define QZ_REFILL_MIN 16   /* idle lanes that trigger a refill (4: -5 %, 8: -1 %, measured on obj_viewer / mandelbrot) */
#endif
#ifndef QZ_TRACE_MIN_BLOCKS
#define QZ_TRACE_MIN_BLOCKS 5
#endif
#ifndef QZ_SMEM_STACK
#define QZ_SMEM_STACK 16   /* stack entries per lane kept in shared memory (16 x 8 B x 128 threads = 16 KB per CTA) */
#endif
#define QZ_LOCAL_STACK (QZ_STACK - QZ_SMEM_STACK)
#ifndef QZ_PRIM_FIRST_MIN
#define QZ_PRIM_FIRST_MIN 33   /* run the primitive stream whenever at least this many lanes wait for it (33 = majority rule only) */
#endif

// family of the surface a closest-hit result lands on (the branch-sorting key)
__device__ __forceinline__ int classify_hit(const DScene& sc, uint32_t geom_id, bool unsorted) {
    if (unsorted || geom_id == QZ_NO_HIT) return SQ_MISC;
    const int32_t mat = sc.geoms[geom_id].material;
    if (mat < 0) return SQ_MISC;
    const uint32_t kind = sc.materials[mat].kind;
    if (kind == QZ_MAT_DIFFUSE) return SQ_DIFFUSE;
    if (kind == QZ_MAT_CONDUCTOR) return SQ_CONDUCTOR;
    if (kind == QZ_MAT_DIELECTRIC || kind == QZ_MAT_THIN_DIELECTRIC) return SQ_DIELECTRIC;
    return SQ_MISC;  // MixedMaterial: the family depends on the bounce's material sample
}

// closest-hit epilogue: write the hit record and the tag of the shade queue the path belongs to
__device__ __forceinline__ void finish_closest(const DScene& sc, const WfBuffers& b, uint32_t slot, bool first, bool late, const Hit& h, uint32_t flags) {
    b.hit_a.set(slot, f4(h.t, h.u, h.v, __uint_as_float(h.prim_id)));
    b.hit_b.set(slot, f4(h.ng.x, h.ng.y, h.ng.z, __uint_as_float(h.geom_id)));
    const bool unsorted = (flags & QZ_FLAG_UNSORTED_SHADING) != 0;
    const int fam = classify_hit(sc, h.geom_id, unsorted);
    b.fam[slot] = (uint8_t)(fam + (first && !unsorted ? SQ_FAMILIES : 0) + (late ? QZ_FAM_LATE : 0u));
}

// LS_POP: the lane finished its leaf primitives and takes its next node from the stack at the head of the next node
// stream, together with the other lanes in that state (popping right where the primitive stream ends ran the stack
// loop for the three or four lanes that had just finished: 6 % of the kernel's instructions at 4 of 32 lanes)
enum LaneState { LS_IDLE = 0, LS_NODE = 1, LS_PRIM = 2, LS_DONE = 3, LS_POP = 4 };

#define QZ_CSWAP_DESC(a, b) { const uint32_t hi_ = a > b ? a : b, lo_ = a > b ? b : a; a = hi_; b = lo_; }

template <bool ANY_HIT, bool COUNT>
__global__ void __launch_bounds__(128, QZ_TRACE_MIN_BLOCKS) k_trace_lane(DScene sc, WfBuffers b, uint32_t flags) {
    __shared__ uint2 s_stack[QZ_SMEM_STACK][128];
    __shared__ WorkList wl;
    if (threadIdx.x == 0) wl.load(b.counters);
    __syncthreads();
    const uint32_t count = wl.total();
    uint32_t* cursor = &b.counters[ANY_HIT ? C_CURSOR_SHADOW : C_CURSOR_TRACE];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int state = LS_IDLE;
    bool exhausted = false;  // warp-uniform
    uint32_t slot = 0, cur = 0, leafbits = 0, leaf_base = 0;
    bool first = false, late = false, occl = false;
    V3 O = v3(0.0f, 0.0f, 0.0f), D = v3(0.0f, 0.0f, 0.0f);
    float inv[3] = {0.0f, 0.0f, 0.0f};
    float limit = 0.0f;   // closest hit: best t so far; any hit: the end of the segment
    Hit best;
    best.t = INFINITY; best.u = 0.0f; best.v = 0.0f; best.prim = QZ_NO_HIT; best.key = 0xffffffffu;
    best.ng = v3(0.0f, 0.0f, 0.0f); best.geom_id = QZ_NO_HIT; best.prim_id = 0;
    uint2 deep[QZ_LOCAL_STACK];
    int sp = 0;
    uint32_t cnt_nodes = 0, cnt_prims = 0, cnt_rays = 0, overflow = 0;

    auto push = [&](uint2 e) {
        if (sp < QZ_SMEM_STACK) s_stack[sp][threadIdx.x] = e;
        else if (sp < QZ_STACK) deep[sp - QZ_SMEM_STACK] = e;
        else { overflow = 1u; return; }
        sp++;
    };
    // next node for this lane: nearest entry still within reach, or the ray is done
    auto pop_next = [&]() {
        for (;;) {
            if (sp == 0) { state = LS_DONE; return; }
            sp--;
            const uint2 e = sp < QZ_SMEM_STACK ? s_stack[sp][threadIdx.x] : deep[sp - QZ_SMEM_STACK];
            if (__uint_as_float(e.x & ~7u) <= limit) {
                cur = e.y; state = LS_NODE;
                prefetch_line_l1(sc.nodes + cur);   // the node stream that opens it is at least a trip away
                return;
            }
        }
    };

    for (;;) {
        const unsigned m_node = __ballot_sync(full, state == LS_NODE || state == LS_POP);
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
                        const float4 L = b.radiance.get(slot), c = b.sh_c.get(slot);
                        b.radiance.set(slot, f4(L.x + c.x, L.y + c.y, L.z + c.z, L.w + c.w));
                    }
                } else {
                    finish_closest(sc, b, slot, first, late, best, flags);
                }
                state = LS_IDLE;
            }
            if (!exhausted) {
                const unsigned idle = m_done | m_idle;
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, (uint32_t)__popc(idle));
                base = __shfl_sync(full, base, 0);
                if (state == LS_IDLE) {
                    const uint32_t widx = base + __popc(idle & ((1u << lane) - 1u));
                    if (widx < count) {
                        const uint32_t idx = wl.slot(b.q_shade, widx);
                        float4 ro, rd;
                        bool live;
                        if (ANY_HIT) {
                            live = (b.post[idx] & QZ_POST_SHADOW) != 0;
                            if (live) {
                                slot = idx;
                                ro = b.sh_o.get(slot); rd = b.sh_d.get(slot);
                                limit = 1.0f;
                            }
                        } else {
                            const uint8_t st = b.stage[idx];
                            live = st != ST_EMPTY;
                            if (live) {
                                slot = idx;
                                first = (st & ST_FIRST) != 0;
                                late = (st & ST_LATE) != 0;
                                ro = b.ray_o.get(slot); rd = b.ray_d.get(slot);
                                limit = INFINITY;
                            } else {
                                b.fam[idx] = QZ_FAM_NONE;
                            }
                        }
                        if (live) {
                            cnt_rays++;
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
                    if (leafbits == 0u) state = LS_POP;
                    else prefetch_line_l1(sc.prims + (size_t)(leaf_base + (uint32_t)(__ffs(leafbits) - 1)) * 4);
                }
            }
            continue;
        }

        // ---- node stream: open one node per lane
        if (state == LS_POP) pop_next();
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
            //
            // SLAB DISTANCES.  The distance to plane q of an axis is ((org + q * 2^e) - O) * inv = q * A + B with
            // A = 2^e * inv (exact: a power of two) and B = (org - O) * inv -- ONE multiply-add per plane, and no
            // integer-to-float conversion either: 0x4B000000 | q is the float 2^23 + q, so the distance is
            // fma(2^23 + q, A, B - 2^23 * A).  This is not the build's decode expression rounded the build's way;
            // it is within  |A| (the two roundings at magnitude 2^23 |A|: one quantisation step, 1/65535 of the node)
            // + 2^-24 (3 |B| + pmax |inv| + 4 |t|)  of it (B: two roundings; the decoded plane's own rounding, pmax = the
            // node's largest coordinate; the final rounding), so each axis interval is widened by
            // slack = 1.6 |A| + 1e-6 |B| + 4e-7 pmax |inv| (which also covers the few ulps box_hit adds): conservative, and the
            // closest hit does not depend on which boxes are opened beyond those that contain it.  (The first version
            // decoded every plane as the build does -- convert, multiply-add, subtract, multiply: 64 instructions per
            // child, over half of the kernel: profiles/r02_summary.md.)
            uint32_t qn[3][4], qf[3][4];
            float A[3], Bn[3], Bf[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const bool fwd = inv[a] >= 0.0f;
                const uint32_t* lo4 = reinterpret_cast<const uint32_t*>(&w[2 + a]);
                const uint32_t* hi4 = reinterpret_cast<const uint32_t*>(&w[5 + a]);
#pragma unroll
                for (int j = 0; j < 4; j++) { qn[a][j] = fwd ? lo4[j] : hi4[j]; qf[a][j] = fwd ? hi4[j] : lo4[j]; }
                const float Oa = a == 0 ? O.x : (a == 1 ? O.y : O.z);
                A[a] = scl[a] * inv[a];
                const float B = (org[a] - Oa) * inv[a];
                const float Bp = __fmaf_rn(-8388608.0f, A[a], B);
                const float pmax = fmaxf(fabsf(org[a]), fabsf(__fmaf_rn(65535.0f, scl[a], org[a])));
                // (+ the few ulps of the distance itself, |t| <= |B| + 65535 |A|, that box_hit applies multiplicatively)
                const float slack = __fmaf_rn(4e-7f, pmax * fabsf(inv[a]), __fmaf_rn(1e-6f, fabsf(B), 1.6f * fabsf(A[a])));
                Bn[a] = Bp - slack;
                Bf[a] = Bp + slack;
            }
            uint32_t key[8];
            uint32_t lb = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t m = (metas[k >> 2] >> (8 * (k & 3))) & 0xffu;
                float t0 = QZ_TNEAR, t1 = limit;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    // bytes of the u16 under 0x4B00: the float 2^23 + q
                    const float xn = __uint_as_float(__byte_perm(qn[a][k >> 1], 0x4B000000u, (k & 1) ? 0x7632u : 0x7610u));
                    const float xf = __uint_as_float(__byte_perm(qf[a][k >> 1], 0x4B000000u, (k & 1) ? 0x7632u : 0x7610u));
                    t0 = fmaxf(t0, __fmaf_rn(xn, A[a], Bn[a]));   // fmaxf / fminf drop NaN (0 * inf, inf - inf)
                    t1 = fminf(t1, __fmaf_rn(xf, A[a], Bf[a]));
                }
                // (t0 >= QZ_TNEAR > 0 and t1 <= limit by their initial values; the conservative widening is in the slack)
                const bool hit = m != 0u && t0 <= t1;
                const bool internal = (m & 0x80u) != 0u;
                key[k] = (hit && internal) ? ((__float_as_uint(t0) & ~7u) | (m & 7u)) : 0u;
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
            // (tried: a fast path without per-entry range checks when all eight would fit the shared-memory part -- no change)
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (key[k]) push(make_uint2(key[k], child_base + (key[k] & 7u)));
            }
            leafbits = lb;
            if (lb) {
                state = LS_PRIM;
                prefetch_line_l1(sc.prims + (size_t)(leaf_base + (uint32_t)(__ffs(lb) - 1)) * 4);  // first pending primitive record
            } else {
                pop_next();
            }
        }
    }
    stat_add(&b.stats[ANY_HIT ? S_RAYS_SHADOW : S_RAYS_CLOSEST], cnt_rays);
    if (COUNT) {
        stat_add(&b.stats[S_NODES], cnt_nodes);
        stat_add(&b.stats[S_PRIMS], cnt_prims);
    }
    if (overflow) atomicAdd(&b.stats[S_OVERFLOW], 1ull);
}

}  // namespace qz
