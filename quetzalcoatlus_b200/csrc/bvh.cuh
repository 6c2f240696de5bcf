// Wide BVH: node layout, GPU build steps and ray traversal.  Replaces Embree's
// rtcCommitScene / rtcIntersect1 (scene.cpp:23, :56).
//
// NODE = 128 bytes = one cache line = eight 16-byte loads:
//   origin (3 x f32), per-axis power-of-two scale exponents (3 x u8), child_base, leaf_base,
//   8 x meta byte, then 16-bit quantised child boxes qlo[3][8], qhi[3][8] (96 bytes).
//   Child box = origin + q * 2^e, rounded OUTWARD at build time and verified with the very
//   expression traversal uses, so quantisation can only enlarge a box.
//   meta: 0 = empty; 0x80 | slot = internal child (node index child_base + slot);
//   (count << 5) | offset = leaf child holding primitives leaf_base + offset .. + count (count 1..3).
//
// BUILD (all on the GPU; steps are written as per-index bodies so the test-only host
// emulation can run the same code sequentially):
//   1 primitive boxes (padded so that no box test can reject a primitive whose own rounded
//     intersection arithmetic reports a hit) and scene bounds;
//   2 30-bit Morton codes over the centroid bounds, 64-bit key = code << 32 | index,
//     radix sort (cub::DeviceRadixSort);
//   3 Karras 2012 LBVH topology from the sorted keys, one thread per internal node;
//   4 bottom-up AABB fit with one atomic flag per internal node;
//   5 top-down collapse of the binary tree into 8-wide nodes, one wavefront per level:
//     repeatedly open the child with the largest surface area; subtrees of <= 3 primitives
//     become leaf children; the primitives of a node's leaf children get one contiguous
//     block of the final primitive array;
//   6 primitive records are gathered into leaf order with their tie-break key.
//   Primitives whose box is huge next to the rest (the 2000-unit planes of opposing_planes /
//   obj_viewer beside 0.8-unit spheres) are kept out of the Morton build and attached to a
//   super-root, so they cannot bloat every ancestor box on their path.
//
// TRAVERSAL: one ray per thread, near-to-far with a per-thread stack; the kernels in
// csrc/wf_trace.cuh wrap it in a persistent, warp-scheduled loop with warp-vote refill.  The
// answer equals brute force over all primitives under the (t, key) order: boxes are
// conservative, entry distances equal to the current best are still visited.  (For hits that lie inside the primitive's
// padded box, that is: a ray exactly coplanar with a triangle can pass the Moeller-Trumbore test on rounding noise far from
// the triangle, and brute force then reports a "hit" every BVH culls -- profiles/r02_summary.md, special-ray sweep.)
#pragma once

#include "intersect.cuh"

namespace qz {

struct alignas(16) BvhNode {
    float ox, oy, oz;
    uint8_t ex, ey, ez, pad;  // biased exponents: scale = 2^(e - 127)
    uint32_t child_base;
    uint32_t leaf_base;
    uint8_t meta[8];
    uint16_t qlo[3][8];
    uint16_t qhi[3][8];
};
static_assert(sizeof(BvhNode) == 128, "a wide node is exactly one 128-byte line");

#define QZ_LEAF_MAX 3 /* primitives per leaf child: the count shares the meta byte with the internal flag (bit 7), so 3 is the maximum */

struct Aabb {
    float lo[3], hi[3];
};

QZ_HD Aabb aabb_empty() {
    Aabb b;
    for (int a = 0; a < 3; a++) { b.lo[a] = INFINITY; b.hi[a] = -INFINITY; }
    return b;
}
QZ_HD void aabb_grow(Aabb& b, const Aabb& o) {
    for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], o.lo[a]); b.hi[a] = fmaxf(b.hi[a], o.hi[a]); }
}
QZ_HD float aabb_half_area(const Aabb& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// step 1: padded box of primitive record i (host-order records: word 2 = kind only)
QZ_HD Aabb prim_box(const F4* prims, uint32_t i) {
    const F4* r = prims + (size_t)i * 4;
    const uint32_t kind = float_as_u32(r[2].w) & 3u;
    Aabb b;
    if (kind == QZ_PRIM_SPHERE) {
        const float c[3] = {r[0].x, r[0].y, r[0].z};
        const float rad = r[1].x;
        for (int a = 0; a < 3; a++) { b.lo[a] = c[a] - rad; b.hi[a] = c[a] + rad; }
    } else {
        const int nv = kind == QZ_PRIM_TRIANGLE ? 3 : 4;
        b = aabb_empty();
        for (int k = 0; k < nv; k++) {
            const float c[3] = {r[k].x, r[k].y, r[k].z};
            for (int a = 0; a < 3; a++) { b.lo[a] = fminf(b.lo[a], c[a]); b.hi[a] = fmaxf(b.hi[a], c[a]); }
        }
    }
    for (int a = 0; a < 3; a++) {
        float m = fmaxf(fabsf(b.lo[a]), fabsf(b.hi[a]));
        float pad = 1e-5f * m + 1e-7f;
        b.lo[a] -= pad;
        b.hi[a] += pad;
    }
    return b;
}

QZ_HD uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// step 2: Morton key of a box centroid inside the centroid bounds
QZ_HD uint64_t morton_key(const Aabb& box, const Aabb& cb, uint32_t index) {
    uint32_t q[3];
    for (int a = 0; a < 3; a++) {
        float c = 0.5f * (box.lo[a] + box.hi[a]);
        float ext = cb.hi[a] - cb.lo[a];
        float n = ext > 0.0f ? (c - cb.lo[a]) / ext : 0.0f;
        n = fminf(fmaxf(n * 1024.0f, 0.0f), 1023.0f);
        q[a] = (uint32_t)n;
    }
    uint32_t code = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
    return ((uint64_t)code << 32) | index;
}

QZ_HD int clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// Binary LBVH over n sorted keys: internal nodes 0..n-2, leaf i is node (n-1)+i.
struct Lbvh {
    const uint64_t* keys;
    uint32_t* left;       // n-1
    uint32_t* right;      // n-1
    uint32_t* parent;     // 2n-1
    uint32_t* first;      // n-1: range of sorted positions covered
    uint32_t* last;       // n-1
    Aabb* box;            // 2n-1
    uint32_t* flag;       // n-1, zeroed
    uint32_t n;
};

QZ_HD int lbvh_delta(const Lbvh& t, int i, int j) {
    if (j < 0 || j >= (int)t.n) return -1;
    return clz64(t.keys[i] ^ t.keys[j]);  // keys are unique (index in the low word)
}

// step 3 (Karras 2012, "Maximizing parallelism in the construction of BVHs", section 4)
QZ_HD void lbvh_topology_body(const Lbvh& t, uint32_t idx) {
    const int i = (int)idx;
    const int d = (lbvh_delta(t, i, i + 1) - lbvh_delta(t, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(t, i, i - d);
    int lmax = 2;
    while (lbvh_delta(t, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int s = lmax / 2; s >= 1; s /= 2)
        if (lbvh_delta(t, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = lbvh_delta(t, i, j);
    int s = 0;
    int step = l;
    do {
        step = (step + 1) >> 1;
        if (lbvh_delta(t, i, i + (s + step) * d) > dnode) s += step;
    } while (step > 1);
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t n = t.n;
    const uint32_t lc = (lo == gamma) ? (n - 1) + (uint32_t)gamma : (uint32_t)gamma;
    const uint32_t rc = (hi == gamma + 1) ? (n - 1) + (uint32_t)(gamma + 1) : (uint32_t)(gamma + 1);
    t.left[i] = lc;
    t.right[i] = rc;
    t.parent[lc] = (uint32_t)i;
    t.parent[rc] = (uint32_t)i;
    t.first[i] = (uint32_t)lo;
    t.last[i] = (uint32_t)hi;
    if (i == 0) t.parent[0] = 0xffffffffu;
}

QZ_HD uint32_t atomic_inc_flag(uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}
QZ_HD uint32_t atomic_add_u32(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    uint32_t o = *p; *p += v; return o;
#endif
}
QZ_HD void fence() {
#if defined(__CUDA_ARCH__)
    __threadfence();
#endif
}

// step 4: one thread per leaf walks up; the second arrival at a node fits its box
QZ_HD void lbvh_fit_body(const Lbvh& t, uint32_t leaf) {
    uint32_t node = t.parent[(t.n - 1) + leaf];
    while (node != 0xffffffffu) {
        fence();
        if (atomic_inc_flag(&t.flag[node]) == 0) return;  // first arrival: sibling not ready
        fence();
        Aabb b = t.box[t.left[node]];
        aabb_grow(b, t.box[t.right[node]]);
        t.box[node] = b;
        node = t.parent[node];
    }
}

QZ_HD float exp_scale(uint32_t biased) { return u32_as_float(biased << 23); }

// Quantise `nc` child boxes against their union.  Returns the node with boxes/meta filled.
QZ_HD void encode_node(BvhNode& node, const Aabb* cbox, const uint8_t* meta, int nc, uint32_t child_base, uint32_t leaf_base) {
    Aabb u = aabb_empty();
    for (int k = 0; k < nc; k++) aabb_grow(u, cbox[k]);
    node.ox = u.lo[0]; node.oy = u.lo[1]; node.oz = u.lo[2];
    node.pad = 0;
    node.child_base = child_base;
    node.leaf_base = leaf_base;
    uint8_t e[3];
    for (int a = 0; a < 3; a++) {
        float ext = u.hi[a] - u.lo[a];
        // smallest power of two with ext / 2^k <= 32767 (one spare bit for the outward fix-ups)
        int k = -100;
        if (ext > 0.0f && !is_inf(ext)) {
            int ex;
            float m = frexpf(ext / 32767.0f, &ex);  // ext/32767 = m * 2^ex, m in [0.5, 1)
            (void)m;
            k = ex;
        }
        int biased = k + 127;
        if (biased < 1) biased = 1;
        if (biased > 254) biased = 254;
        e[a] = (uint8_t)biased;
    }
    node.ex = e[0]; node.ey = e[1]; node.ez = e[2];
    const float org[3] = {node.ox, node.oy, node.oz};
    for (int k = 0; k < 8; k++) {
        node.meta[k] = k < nc ? meta[k] : 0;
        for (int a = 0; a < 3; a++) {
            uint32_t qlo = 0, qhi = 0;
            if (k < nc) {
                const float s = exp_scale(e[a]);
                float fl = floorf((cbox[k].lo[a] - org[a]) / s);
                float fh = ceilf((cbox[k].hi[a] - org[a]) / s);
                fl = fminf(fmaxf(fl, 0.0f), 65535.0f);
                fh = fminf(fmaxf(fh, 0.0f), 65535.0f);
                qlo = (uint32_t)fl; qhi = (uint32_t)fh;
                // verify with the decode expression of traversal; widen until conservative
                while (qlo > 0 && org[a] + (float)qlo * s > cbox[k].lo[a]) qlo--;
                while (qhi < 65535u && org[a] + (float)qhi * s < cbox[k].hi[a]) qhi++;
            }
            node.qlo[a][k] = (uint16_t)qlo;
            node.qhi[a][k] = (uint16_t)qhi;
        }
    }
}

// step 5: collapse work item -> one wide node
struct CollapseItem {
    uint32_t bin;   // binary internal node to open
    uint32_t wide;  // wide node index to fill
};

struct Collapse {
    Lbvh t;
    BvhNode* nodes;
    uint32_t* node_counter;   // next free wide node
    uint32_t* prim_counter;   // next free slot of the final primitive order
    uint32_t* final_sorted;   // final slot -> sorted position
    CollapseItem* out_queue;
    uint32_t* out_count;
};

QZ_HD uint32_t lbvh_count(const Lbvh& t, uint32_t node) {
    return node >= t.n - 1 ? 1u : t.last[node] - t.first[node] + 1u;
}
QZ_HD uint32_t lbvh_first(const Lbvh& t, uint32_t node) { return node >= t.n - 1 ? node - (t.n - 1) : t.first[node]; }

QZ_HD void collapse_body(const Collapse& c, const CollapseItem item) {
    const Lbvh& t = c.t;
    uint32_t ch[8];
    int nc = 2;
    ch[0] = t.left[item.bin];
    ch[1] = t.right[item.bin];
    while (nc < 8) {
        int best = -1;
        float best_area = -1.0f;
        for (int k = 0; k < nc; k++) {
            if (lbvh_count(t, ch[k]) <= QZ_LEAF_MAX) continue;  // stays a leaf child
            float area = aabb_half_area(t.box[ch[k]]);
            if (area > best_area) { best_area = area; best = k; }
        }
        if (best < 0) break;
        uint32_t open = ch[best];
        ch[best] = t.left[open];
        ch[nc++] = t.right[open];
    }
    int n_int = 0;
    uint32_t n_leaf_prims = 0;
    for (int k = 0; k < nc; k++) {
        uint32_t cnt = lbvh_count(t, ch[k]);
        if (cnt > QZ_LEAF_MAX) n_int++; else n_leaf_prims += cnt;
    }
    uint32_t child_base = n_int ? atomic_add_u32(c.node_counter, (uint32_t)n_int) : 0u;
    uint32_t leaf_base = n_leaf_prims ? atomic_add_u32(c.prim_counter, n_leaf_prims) : 0u;
    uint32_t qbase = n_int ? atomic_add_u32(c.out_count, (uint32_t)n_int) : 0u;
    Aabb cbox[8];
    uint8_t meta[8];
    int slot = 0;
    uint32_t off = 0;
    for (int k = 0; k < nc; k++) {
        cbox[k] = t.box[ch[k]];
        uint32_t cnt = lbvh_count(t, ch[k]);
        if (cnt > QZ_LEAF_MAX) {
            meta[k] = (uint8_t)(0x80u | (uint32_t)slot);
            CollapseItem next;
            next.bin = ch[k];
            next.wide = child_base + (uint32_t)slot;
            c.out_queue[qbase + (uint32_t)slot] = next;
            slot++;
        } else {
            meta[k] = (uint8_t)((cnt << 5) | off);
            uint32_t f = lbvh_first(t, ch[k]);
            for (uint32_t j = 0; j < cnt; j++) c.final_sorted[leaf_base + off + j] = f + j;
            off += cnt;
        }
    }
    BvhNode node;
    encode_node(node, cbox, meta, nc, child_base, leaf_base);
    c.nodes[item.wide] = node;
}

// step 6: write the record of final slot `slot` from host-order primitive `src`, packing the
// tie-break key (= host order index) next to the kind
QZ_HD void gather_prim_body(const F4* src_prims, F4* dst_prims, uint32_t slot, uint32_t src) {
    const F4* r = src_prims + (size_t)src * 4;
    F4* w = dst_prims + (size_t)slot * 4;
    w[0] = r[0]; w[1] = r[1]; w[3] = r[3];
    F4 c = r[2];
    c.w = u32_as_float((float_as_u32(c.w) & 3u) | (src << 2));
    w[2] = c;
}

// ------------------------------------------------------------------ traversal
struct Ray {
    V3 o, d;
};

#define QZ_STACK 128

struct StackEntry {
    float dist;
    uint32_t node;
};

// slab test against a decoded child box: true when the ray interval [tnear, limit] overlaps
// the box, with the entry distance in `entry`.  The interval is widened by a few ulps
// (multiplicatively, so infinities stay infinities) to keep the test conservative.
QZ_HD bool box_hit(const float lo[3], const float hi[3], const float o[3], const float inv[3], float tnear, float limit,
                   float& entry) {
    float t0 = tnear, t1 = limit;
    for (int a = 0; a < 3; a++) {
        float ta = (lo[a] - o[a]) * inv[a];
        float tb = (hi[a] - o[a]) * inv[a];
        float tmn = fminf(ta, tb), tmx = fmaxf(ta, tb);  // fminf/fmaxf drop NaN (0 * inf)
        tmn = tmn * (tmn >= 0.0f ? 0.9999995f : 1.0000005f);
        tmx = tmx * (tmx >= 0.0f ? 1.0000005f : 0.9999995f);
        t0 = fmaxf(t0, tmn);
        t1 = fminf(t1, tmx);
    }
    entry = t0;
    return t0 <= t1;
}

struct TraversalCounters {
    uint32_t nodes, prims;
};

// Traversal is a small state machine so that the persistent kernels can interleave "fetch a
// new ray" with "advance the current one" per lane (warp-vote refill).
struct Trav {
    Ray ray;
    float o[3], inv[3];
    float tmax_any;
    Hit best;
    StackEntry stack[QZ_STACK];
    int sp;
    uint32_t cur;
    float cur_dist;
    bool done;
    bool occluded;
    bool overflow;   // a child was dropped because the stack was full: the result cannot be trusted
};

QZ_HD void trav_init(Trav& tv, const Ray& ray, float tmax_any) {
    tv.ray = ray;
    tv.o[0] = ray.o.x; tv.o[1] = ray.o.y; tv.o[2] = ray.o.z;
    tv.inv[0] = 1.0f / ray.d.x; tv.inv[1] = 1.0f / ray.d.y; tv.inv[2] = 1.0f / ray.d.z;
    tv.tmax_any = tmax_any;
    tv.best.t = INFINITY; tv.best.u = 0.0f; tv.best.v = 0.0f; tv.best.prim = QZ_NO_HIT; tv.best.key = 0xffffffffu;
    tv.best.ng = v3(0.0f, 0.0f, 0.0f);
    tv.best.geom_id = QZ_NO_HIT; tv.best.prim_id = 0;
    tv.sp = 0;
    tv.cur = 0;
    tv.cur_dist = 0.0f;
    tv.done = false;
    tv.occluded = false;
    tv.overflow = false;
}

// ANY_HIT: stop at the first primitive hit with t <= tmax_any (Scene::occluded,
// scene.cpp:136-143: "closest hit has tfar <= 1" == "some hit has t <= 1").
// One call = visit node `cur` (if still within reach), then pop the next one.
template <bool ANY_HIT, bool COUNT>
QZ_HD void trav_step(const DScene& sc, Trav& tv, TraversalCounters* cnt) {
    const float limit = ANY_HIT ? tv.tmax_any : tv.best.t;
    if (tv.cur_dist <= limit) {
        const BvhNode* np = sc.nodes + tv.cur;
#if defined(__CUDA_ARCH__)
        // eight 16-byte loads of one 128-byte line
        uint4 w[8];
        const uint4* src = reinterpret_cast<const uint4*>(np);
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = __ldg(src + k);
        const BvhNode& n = *reinterpret_cast<const BvhNode*>(w);
#else
        const BvhNode& n = *np;
#endif
        if (COUNT) cnt->nodes++;
        const float org[3] = {n.ox, n.oy, n.oz};
        const float scl[3] = {exp_scale(n.ex), exp_scale(n.ey), exp_scale(n.ez)};
        const int base_sp = tv.sp;
        for (int k = 0; k < 8; k++) {
            const uint32_t m = n.meta[k];
            if (m == 0) continue;
            float lo[3], hi[3];
            for (int a = 0; a < 3; a++) {
                lo[a] = org[a] + (float)n.qlo[a][k] * scl[a];
                hi[a] = org[a] + (float)n.qhi[a][k] * scl[a];
            }
            const float lim = ANY_HIT ? tv.tmax_any : tv.best.t;
            float d;
            if (!box_hit(lo, hi, tv.o, tv.inv, QZ_TNEAR, lim, d)) continue;
            if (m & 0x80u) {
                if (tv.sp < QZ_STACK) {
                    tv.stack[tv.sp].dist = d;
                    tv.stack[tv.sp].node = n.child_base + (m & 0x7fu);
                    tv.sp++;
                } else {
                    tv.overflow = true;
                }
            } else {
                const uint32_t first = n.leaf_base + (m & 31u);
                const uint32_t count = m >> 5;
                for (uint32_t j = 0; j < count; j++) {
                    if (COUNT) cnt->prims++;
                    prim_test(sc, first + j, tv.ray.o, tv.ray.d, QZ_TNEAR, INFINITY, tv.best);
                }
                if (ANY_HIT && tv.best.prim != QZ_NO_HIT) {
                    if (tv.best.t <= tv.tmax_any) { tv.occluded = true; tv.done = true; return; }
                    // a hit beyond the segment does not occlude: forget it
                    tv.best.t = INFINITY; tv.best.prim = QZ_NO_HIT; tv.best.key = 0xffffffffu;
                }
            }
        }
        // order the newly pushed children far-to-near so the nearest is popped first
        for (int a = base_sp + 1; a < tv.sp; a++) {
            StackEntry e = tv.stack[a];
            int b = a - 1;
            while (b >= base_sp && tv.stack[b].dist < e.dist) { tv.stack[b + 1] = tv.stack[b]; b--; }
            tv.stack[b + 1] = e;
        }
    }
    if (tv.sp == 0) { tv.done = true; return; }
    tv.sp--;
    tv.cur = tv.stack[tv.sp].node;
    tv.cur_dist = tv.stack[tv.sp].dist;
}

// closest hit (Scene::ray_intersect's rtcIntersect1, scene.cpp:56): returns true on a hit
template <bool COUNT>
QZ_HD bool closest_hit(const DScene& sc, const Ray& ray, Hit& hit, TraversalCounters* cnt) {
    Trav tv;
    trav_init(tv, ray, INFINITY);
    while (!tv.done) trav_step<false, COUNT>(sc, tv, cnt);
    if (tv.overflow && sc.overflow) *sc.overflow = 1u;   // every writer stores the same value
    hit = tv.best;
    return hit.prim != QZ_NO_HIT;
}

// Scene::occluded(start, end) with ray = (start, end - start) (scene.cpp:136-143)
template <bool COUNT>
QZ_HD bool occluded(const DScene& sc, const Ray& ray, TraversalCounters* cnt) {
    Trav tv;
    trav_init(tv, ray, 1.0f);
    while (!tv.done) trav_step<true, COUNT>(sc, tv, cnt);
    if (tv.overflow && sc.overflow) *sc.overflow = 1u;
    return tv.occluded;
}

}  // namespace qz
