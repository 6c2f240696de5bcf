// HOST side of the SAMPLE MEMO (sampler.cuh): which (x mod 128, y mod 128) pixel classes a call owns, how a row is laid
// out and how the device-side descriptor is filled in.  Shared by the CUDA library (qz_b200.cu) and by the host emulation
// of the test-suite (tests/emu), so that the CPU tests walk the very tables the kernels read.
#pragma once
#include <cstdint>
#include <vector>

#include "sampler.cuh"

namespace qz {

// Classes owned by a call that renders image rows `rows` (top-down row numbers, as the film stores them) of a W x H image:
// `cls_rank[idx_pix]` = rank of the class whose sample-0 Halton index is idx_pix (0xffffffff: not owned), `cls_idx` the
// inverse list.  The sampler addresses pixels bottom-up (render.cpp:262), hence H - 1 - row.
inline void memo_owned_classes(const SamplerParams& spar, uint32_t W, uint32_t H, const std::vector<uint32_t>& rows,
                               std::vector<uint32_t>& cls_rank, std::vector<uint32_t>& cls_idx) {
    bool own_py[QZ_MAX_HALTON_RESOLUTION] = {};
    for (uint32_t r : rows) own_py[(H - 1u - r) & (QZ_MAX_HALTON_RESOLUTION - 1)] = true;
    cls_rank.assign(spar.stride, 0xffffffffu);
    cls_idx.clear();
    const uint32_t cw = W < QZ_MAX_HALTON_RESOLUTION ? W : (uint32_t)QZ_MAX_HALTON_RESOLUTION;
    for (uint32_t py = 0; py < QZ_MAX_HALTON_RESOLUTION; py++) {
        if (!own_py[py]) continue;
        for (uint32_t px = 0; px < cw; px++) {
            const uint32_t idx = sampler_start(spar, px, py, 0).index;
            if (cls_rank[idx] == 0xffffffffu) { cls_rank[idx] = (uint32_t)cls_idx.size(); cls_idx.push_back(idx); }
        }
    }
}

// Words of a row: the hot-spectra block, then dimension 0 at a word = 5 mod 8 so that dimension 3 -- the first bounce's
// first draw -- starts a 32-byte sector; 3 + 8 * bounces dimensions, rounded up to whole sectors.
struct MemoLayout {
    uint32_t dim_off, dims, stride;
};
inline MemoLayout memo_layout(uint32_t bounces) {
    MemoLayout l;
    l.dim_off = ((4u * QZ_MEMO_MAX_HOT + 7u) & ~7u) + 5u;
    l.dims = 3u + 8u * bounces;
    l.stride = (l.dim_off + l.dims + 7u) & ~7u;
    return l;
}

// descriptor of one pass's table (n_hot stays 0 until the hot block has been filled: the fill itself must evaluate)
inline SampleMemo memo_describe(const MemoLayout& l, const SamplerParams& spar, uint32_t* tab, const uint32_t* cls_rank, const uint32_t* cls_idx,
                                uint32_t n_cls, uint32_t s_begin, uint32_t s_count) {
    SampleMemo m{};
    m.tab = tab;
    m.dims = l.dims;
    m.n_cls = n_cls;
    m.s_begin = s_begin;
    m.s_count = s_count;
    m.stride = l.stride;
    m.dim_off = l.dim_off;
    m.n_hot = 0;
    m.idx_stride = spar.stride;
    m.idx_magic = (uint32_t)((1ull << 32) / spar.stride);
    m.cls_rank = cls_rank;
    m.cls_idx = cls_idx;
    return m;
}

}  // namespace qz
