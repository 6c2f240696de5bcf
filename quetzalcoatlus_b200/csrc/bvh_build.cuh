// Orchestration of the wide-BVH build (steps 1-6 of bvh.cuh) over an executor.  The CUDA
// library instantiates it with CudaExec (qz_b200.cu: every parallel_for is a kernel launch,
// the sort is cub::DeviceRadixSort); the test-only host emulation instantiates it with a
// sequential executor, so the CPU test-suite exercises exactly this code.
#pragma once

#include <algorithm>
#include <vector>

#include "bvh.cuh"

#if defined(__CUDACC__)
#define QZ_LAMBDA [=] __host__ __device__
#else
#define QZ_LAMBDA [=]
#endif

namespace qz {

struct BvhBuildResult {
    BvhNode* nodes = nullptr;  // executor memory
    F4* prims = nullptr;       // leaf-ordered primitive records (4 x F4 each), executor memory
    uint32_t n_nodes = 0;
    uint32_t n_prims = 0;
    uint32_t n_large = 0;      // primitives hoisted to the super-root
    uint32_t depth = 0;        // collapse levels
};

// monotone float <-> int map for atomicMin / atomicMax on floats
QZ_HD int float_to_ordered(float f) {
    int i = (int)float_as_u32(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
QZ_HD float ordered_to_float(int i) { return u32_as_float((uint32_t)(i >= 0 ? i : i ^ 0x7fffffff)); }

QZ_HD void atomic_min_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
QZ_HD void atomic_max_i(int* p, int v) {
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}

#define QZ_MAX_LARGE 21u  /* 7 leaf children x 3 primitives beside the small-tree child */
#define QZ_ROOT_ONLY 24u  /* 8 leaf children x 3 primitives */

// host helper: one wide node whose children are all leaf children over consecutive final slots
inline void make_leaf_only_node(BvhNode& node, const Aabb* boxes, uint32_t n, uint32_t leaf_base, int first_child,
                                const Aabb* extra_box, uint8_t extra_meta, uint32_t child_base) {
    Aabb cbox[8];
    uint8_t meta[8];
    int nc = 0;
    if (extra_box) { cbox[nc] = *extra_box; meta[nc] = extra_meta; nc++; }
    (void)first_child;
    uint32_t off = 0;
    while (off < n && nc < 8) {
        uint32_t cnt = std::min<uint32_t>(QZ_LEAF_MAX, n - off);
        Aabb b = aabb_empty();
        for (uint32_t j = 0; j < cnt; j++) aabb_grow(b, boxes[off + j]);
        cbox[nc] = b;
        meta[nc] = (uint8_t)((cnt << 5) | off);
        nc++;
        off += cnt;
    }
    encode_node(node, cbox, meta, nc, child_base, leaf_base);
}

template <class Exec>
bool build_wide_bvh(Exec& ex, const F4* d_src_prims, uint32_t n, BvhBuildResult& out) {
    out = BvhBuildResult();
    out.n_prims = n;
    if (n == 0) {
        // an empty scene still needs a root: one node without children
        BvhNode root;
        encode_node(root, nullptr, nullptr, 0, 0, 0);
        root.ox = root.oy = root.oz = 0.0f;
        out.nodes = ex.template alloc<BvhNode>(1);
        ex.upload(out.nodes, &root, 1);
        out.prims = ex.template alloc<F4>(4);
        out.n_nodes = 1;
        return true;
    }

    ex.mark("(start)");
    // ---- 1: boxes, scene bounds
    Aabb* d_box = ex.template alloc<Aabb>(n);
    int* d_bounds = ex.template alloc<int>(12);  // [0..5] all-primitive bounds, [6..11] small-centroid bounds
    {
        int init[12];
        for (int a = 0; a < 3; a++) {
            init[a] = init[6 + a] = float_to_ordered(INFINITY);
            init[3 + a] = init[9 + a] = float_to_ordered(-INFINITY);
        }
        ex.upload(d_bounds, init, 12);
    }
    ex.parallel_for(n, QZ_LAMBDA(uint32_t i) {
        Aabb b = prim_box(d_src_prims, i);
        d_box[i] = b;
        for (int a = 0; a < 3; a++) {
            atomic_min_i(d_bounds + a, float_to_ordered(b.lo[a]));
            atomic_max_i(d_bounds + 3 + a, float_to_ordered(b.hi[a]));
        }
    });
    int h_bounds[12];
    ex.download(h_bounds, d_bounds, 12);
    float scene_ext = 0.0f;
    for (int a = 0; a < 3; a++) scene_ext = std::max(scene_ext, ordered_to_float(h_bounds[3 + a]) - ordered_to_float(h_bounds[a]));
    // prim_box pads a box relative to the primitive's OWN coordinates, but the rounding error of the intersection
    // arithmetic (C = v0 - O, perp = c0 - D * projC0) grows with the distance to the ray's ORIGIN, which may lie
    // anywhere in the scene (a path leaving a 2000-unit plane towards a 0.8-unit sphere).  A second pad proportional to
    // the largest coordinate of the scene keeps "BVH answer == brute force" for such rays too.
    {
        float far = 0.0f;
        for (int a = 0; a < 6; a++) { const float v = std::fabs(ordered_to_float(h_bounds[a])); if (v <= 3.0e38f) far = std::max(far, v); }
        const float far_pad = 3e-7f * far;
        ex.parallel_for(n, QZ_LAMBDA(uint32_t i) {
            Aabb b = d_box[i];
            for (int a = 0; a < 3; a++) { b.lo[a] -= far_pad; b.hi[a] += far_pad; }
            d_box[i] = b;
        });
    }

    ex.mark("1 boxes + bounds");
    // ---- hoisting of huge primitives
    uint32_t* d_large = ex.template alloc<uint32_t>(64);
    uint32_t* d_counters = ex.template alloc<uint32_t>(8);  // 0 large, 1 small, 2 node, 3 prim, 4 queue
    ex.zero(d_counters, 8 * sizeof(uint32_t));
    std::vector<uint32_t> large;
    if (n > QZ_ROOT_ONLY) {
        const float thresh = 0.25f * scene_ext;
        ex.parallel_for(n, QZ_LAMBDA(uint32_t i) {
            const Aabb b = d_box[i];
            float e = fmaxf(b.hi[0] - b.lo[0], fmaxf(b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]));
            if (e > thresh) {
                uint32_t pos = atomic_add_u32(d_counters + 0, 1u);
                if (pos < 64u) d_large[pos] = i;
            }
        });
        uint32_t n_large = 0;
        ex.download(&n_large, d_counters + 0, 1);
        if (n_large > 0 && n_large <= QZ_MAX_LARGE && n_large < n) {
            large.resize(n_large);
            ex.download(large.data(), d_large, n_large);
            std::sort(large.begin(), large.end());
        }
    }
    const uint32_t n_large = (uint32_t)large.size();
    const uint32_t n_small = n - n_large;
    out.n_large = n_large;

    // small-primitive index list (identity when nothing is hoisted)
    uint32_t* d_small = ex.template alloc<uint32_t>(n_small ? n_small : 1);
    if (n_large == 0) {
        ex.parallel_for(n, QZ_LAMBDA(uint32_t i) { d_small[i] = i; });
    } else {
        uint32_t* d_is_large = ex.template alloc<uint32_t>(n);
        ex.zero(d_is_large, n * sizeof(uint32_t));
        const uint32_t nl = n_large;
        ex.upload(d_large, large.data(), n_large);
        ex.parallel_for(nl, QZ_LAMBDA(uint32_t k) { d_is_large[d_large[k]] = 1u; });
        ex.parallel_for(n, QZ_LAMBDA(uint32_t i) {
            if (!d_is_large[i]) d_small[atomic_add_u32(d_counters + 1, 1u)] = i;
        });
        ex.free(d_is_large);
    }

    ex.mark("hoist huge primitives");
    uint32_t* d_final_src = ex.template alloc<uint32_t>(n);
    const uint32_t small_root = n_large ? 1u : 0u;
    BvhNode* d_nodes = ex.template alloc<BvhNode>((size_t)n_small + 2);
    Aabb small_union = aabb_empty();
    uint32_t n_nodes = small_root + 1;

    if (n_small <= QZ_ROOT_ONLY) {
        // ---- tiny scene: a single node of leaf children, built from downloaded boxes
        std::vector<uint32_t> idx(n_small);
        ex.download(idx.data(), d_small, n_small);
        std::sort(idx.begin(), idx.end());
        std::vector<Aabb> boxes(n_small);
        for (uint32_t k = 0; k < n_small; k++) {
            ex.download(&boxes[k], d_box + idx[k], 1);
            aabb_grow(small_union, boxes[k]);
        }
        BvhNode node;
        make_leaf_only_node(node, boxes.data(), n_small, 0, 0, nullptr, 0, 0);
        ex.upload(d_nodes + small_root, &node, 1);
        ex.upload(d_final_src, idx.data(), n_small);
        out.depth = 1;
    } else {
        // ---- 2: Morton keys over the centroid bounds of the small primitives, sort
        ex.parallel_for(n_small, QZ_LAMBDA(uint32_t k) {
            const Aabb b = d_box[d_small[k]];
            for (int a = 0; a < 3; a++) {
                float c = 0.5f * (b.lo[a] + b.hi[a]);
                atomic_min_i(d_bounds + 6 + a, float_to_ordered(c));
                atomic_max_i(d_bounds + 9 + a, float_to_ordered(c));
            }
        });
        ex.download(h_bounds, d_bounds, 12);
        Aabb cb;
        for (int a = 0; a < 3; a++) { cb.lo[a] = ordered_to_float(h_bounds[6 + a]); cb.hi[a] = ordered_to_float(h_bounds[9 + a]); }
        uint64_t* d_keys = ex.template alloc<uint64_t>(n_small);
        ex.parallel_for(n_small, QZ_LAMBDA(uint32_t k) {
            const uint32_t p = d_small[k];
            d_keys[k] = morton_key(d_box[p], cb, p);
        });
        ex.mark("2 morton keys");
        ex.sort_u64(d_keys, n_small);
        ex.mark("2 radix sort");

        // ---- 3, 4: topology and boxes
        Lbvh t;
        t.n = n_small;
        t.keys = d_keys;
        t.left = ex.template alloc<uint32_t>(n_small - 1);
        t.right = ex.template alloc<uint32_t>(n_small - 1);
        t.parent = ex.template alloc<uint32_t>(2 * (size_t)n_small - 1);
        t.first = ex.template alloc<uint32_t>(n_small - 1);
        t.last = ex.template alloc<uint32_t>(n_small - 1);
        t.box = ex.template alloc<Aabb>(2 * (size_t)n_small - 1);
        t.flag = ex.template alloc<uint32_t>(n_small - 1);
        ex.zero(t.flag, (n_small - 1) * sizeof(uint32_t));
        ex.parallel_for(n_small - 1, QZ_LAMBDA(uint32_t i) { lbvh_topology_body(t, i); });
        ex.parallel_for(n_small, QZ_LAMBDA(uint32_t i) { t.box[(t.n - 1) + i] = d_box[(uint32_t)t.keys[i]]; });
        ex.parallel_for(n_small, QZ_LAMBDA(uint32_t i) { lbvh_fit_body(t, i); });
        ex.download(&small_union, t.box, 1);
        ex.mark("3-4 topology + fit");

        // ---- 5: collapse, one launch per tree level
        CollapseItem* d_q[2] = {ex.template alloc<CollapseItem>(n_small), ex.template alloc<CollapseItem>(n_small)};
        CollapseItem first_item;
        first_item.bin = 0;
        first_item.wide = small_root;
        ex.upload(d_q[0], &first_item, 1);
        uint32_t h_counters[8] = {0, 0, small_root + 1, 0, 0, 0, 0, 0};
        ex.upload(d_counters, h_counters, 8);
        uint32_t count = 1;
        int cur = 0;
        while (count) {
            Collapse c;
            c.t = t;
            c.nodes = d_nodes;
            c.node_counter = d_counters + 2;
            c.prim_counter = d_counters + 3;
            c.final_sorted = d_final_src;
            c.out_queue = d_q[cur ^ 1];
            c.out_count = d_counters + 4;
            const CollapseItem* in = d_q[cur];
            ex.parallel_for(count, QZ_LAMBDA(uint32_t i) { collapse_body(c, in[i]); });
            ex.download(h_counters, d_counters, 8);
            count = h_counters[4];
            h_counters[4] = 0;
            ex.upload(d_counters + 4, h_counters + 4, 1);
            cur ^= 1;
            out.depth++;
        }
        ex.mark("5 collapse");
        n_nodes = h_counters[2];
        // collapse wrote sorted positions: turn them into host-order primitive indices
        ex.parallel_for(n_small, QZ_LAMBDA(uint32_t s) { d_final_src[s] = (uint32_t)t.keys[d_final_src[s]]; });
        ex.free(d_q[0]); ex.free(d_q[1]);
        ex.free(t.left); ex.free(t.right); ex.free(t.parent); ex.free(t.first); ex.free(t.last); ex.free(t.box); ex.free(t.flag);
        ex.free(d_keys);
    }

    if (n_large) {
        // super-root: child 0 = the small tree, the rest = leaf children over the hoisted primitives
        std::vector<Aabb> lboxes(n_large);
        for (uint32_t k = 0; k < n_large; k++) ex.download(&lboxes[k], d_box + large[k], 1);
        BvhNode root;
        make_leaf_only_node(root, lboxes.data(), n_large, n_small, 1, &small_union, (uint8_t)0x80u, small_root);
        ex.upload(d_nodes, &root, 1);
        ex.upload(d_final_src + n_small, large.data(), n_large);
    }

    // ---- 6: gather the records into leaf order
    F4* d_prims = ex.template alloc<F4>((size_t)n * 4);
    ex.parallel_for(n, QZ_LAMBDA(uint32_t s) { gather_prim_body(d_src_prims, d_prims, s, d_final_src[s]); });
    ex.mark("6 gather");

    ex.free(d_box); ex.free(d_bounds); ex.free(d_large); ex.free(d_counters); ex.free(d_small); ex.free(d_final_src);
    out.nodes = d_nodes;
    out.prims = d_prims;
    out.n_nodes = n_nodes;
    return true;
}

}  // namespace qz
