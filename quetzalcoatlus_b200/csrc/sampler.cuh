// Owen-scrambled Halton sampler, bit-exact restatement of the reference's Sampler
// (sampler.cpp:161-198 permutation_element / mix_bits, :298-314 radical_inv, :316-324
// inv_radical_inv, :335-352 owen_scrambled_radical_inv, :383-454 Sampler).  Everything up
// to the final float multiply is integer arithmetic, so the device values equal the
// oracle's bit for bit (tests/test_gpu_parity.py::test_sampler_bit_exact* and tests/test_emu_parity.py check the known answers of SURVEY.md
// section 4 and random (x, y, s, dim) tuples).
//
// B200 shaping: per-dimension constants live in a 16-byte record (one LDG.128): the prime
// base with its digit count, floor(2^32/base) for multiply-high division (one correction
// step, exact for every 32-bit numerator because base <= 7919), the per-dimension hash
// uint32(mix_bits(1 + (dim << 4))) and base^-ndigits accumulated in float exactly as the
// reference's loop does.  The Halton index fits 32 bits (spp * 31104 < 2^31, the
// reference's own int overflow limit at sampler.cpp:420).
#pragma once

#include "common.cuh"

namespace qz {

#define QZ_N_PRIMES 1000
#define QZ_MAX_HALTON_RESOLUTION 128
#define QZ_ONE_MINUS_EPS 0.99999994f /* 0x1.fffffep-1 (util.hpp:3) */

struct SamplerDim {
    uint32_t base_nd;  // base | ndigits << 16
    uint32_t magic;    // floor(2^32 / base)
    uint32_t hash;     // uint32(mix_bits(1 + (dim << 4)))
    uint32_t scale;    // float bits of base^-ndigits (sequential float products)
    // prefix table (see below): the scrambled value of the index's lowest pre_k digits
    uint32_t pre_pow;     // base^pre_k; 0 = this dimension has no prefix table
    uint32_t pre_magic;   // floor(2^32 / pre_pow)
    uint32_t pre_offset;  // first entry of this dimension in the prefix array
    uint32_t pre_k;
};
static_assert(sizeof(SamplerDim) == 32, "two 16-byte loads per dimension record");

// PREFIX TABLES.  The scrambled value of the lowest k digits of the Halton index depends on
// nothing but (dimension, index mod base^k): digit j's permutation is keyed by the scrambled
// digits below it.  So for every dimension the first k digits -- as many as fit
// QZ_PREFIX_CAP entries -- are tabulated once per device (16-bit entries, since the value is
// below base^k <= 32768), the table following the QZ_N_PRIMES records in the same allocation.
// An evaluation then costs one division by base^k, one 2-byte gather and the remaining
// nd - k digits: base 5 runs 5 digit rounds instead of 11, base 7 4 instead of 9, bases 37..181
// 3 instead of 5.  The entries are produced by the very digit loop below, so the values are
// the reference's bit for bit.
#define QZ_PREFIX_CAP 32768u
QZ_HD const uint16_t* sampler_prefix_array(const SamplerDim* table) {
    return reinterpret_cast<const uint16_t*>(table + QZ_N_PRIMES);
}

// per-render constants of Sampler's constructor (sampler.cpp:383-402)
struct SamplerParams {
    uint32_t scale0, scale1;  // base_scales: powers of 2 / 3 covering min(res, 128)
    uint32_t exp0, exp1;      // base_exps
    uint32_t mult_inv0, mult_inv1;
    uint32_t stride;          // scale0 * scale1
    uint32_t pad;
};

QZ_HD uint64_t mix_bits(uint64_t v) {
    v ^= (v >> 31);
    v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27);
    v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}

// x / d and x % d for any 32-bit x, with m = floor(2^32 / d): the estimate is low by at most 1
QZ_HD uint32_t div_magic(uint32_t x, uint32_t d, uint32_t m, uint32_t& rem) {
    uint32_t q = umulhi32(x, m);
    uint32_t r = x - q * d;
    if (r >= d) { r -= d; q++; }
    rem = r;
    return q;
}

// sampler.cpp:161-189
QZ_HD uint32_t permutation_element(uint32_t i, uint32_t l, uint32_t magic, uint32_t w, uint32_t p) {
    do {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;
        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;
        i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    uint32_t rem;
    div_magic(i + p, l, magic, rem);
    return rem;
}

// digits [k0, k1) of sampler.cpp:335-352's loop: `a` holds the index with the first k0 digits
// already removed, `reversed` their scrambled value
QZ_HD uint64_t owen_digits(const SamplerDim& dimrec, uint32_t a, uint32_t k0, uint32_t k1, uint64_t reversed) {
    const uint32_t base = dimrec.base_nd & 0xffffu;
    // w = (next power of two >= base) - 1, as the or-shift cascade of permutation_element computes it
    uint32_t w = base - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    for (uint32_t k = k0; k < k1; k++) {
        uint32_t digit;
        a = div_magic(a, base, dimrec.magic, digit);
        uint32_t digit_hash = (uint32_t)mix_bits((uint64_t)dimrec.hash ^ reversed);
        digit = permutation_element(digit, base, dimrec.magic, w, digit_hash);
        reversed = reversed * base + digit;
    }
    return reversed;
}

// sampler.cpp:335-352 for one table entry; `a` is the Halton index
QZ_HD float owen_scrambled_radical_inv(const SamplerDim& dimrec, const uint16_t* prefix, uint32_t a) {
    const uint32_t nd = dimrec.base_nd >> 16;
    uint64_t reversed;
    if (dimrec.pre_pow) {
        uint32_t lo;
        const uint32_t hi = div_magic(a, dimrec.pre_pow, dimrec.pre_magic, lo);
#if defined(__CUDA_ARCH__)
        const uint32_t head = __ldg(prefix + dimrec.pre_offset + lo);
#else
        const uint32_t head = prefix[dimrec.pre_offset + lo];
#endif
        reversed = owen_digits(dimrec, hi, dimrec.pre_k, nd, head);
    } else {
        reversed = owen_digits(dimrec, a, 0, nd, 0);
    }
    float r = u32_as_float(dimrec.scale) * (float)reversed;
    return std_min(r, QZ_ONE_MINUS_EPS);
}

// sampler.cpp:298-314 (unscrambled; used for the pixel jitter only)
QZ_HD float radical_inv(uint32_t base, uint32_t a) {
    float inv_base = 1.0f / (float)base, inv_base_m = 1.0f;
    uint64_t reversed = 0;
    while (a) {
        uint32_t next = a / base;
        uint32_t digit = a - next * base;
        reversed = reversed * base + digit;
        inv_base_m *= inv_base;
        a = next;
    }
    return std_min((float)reversed * inv_base_m, QZ_ONE_MINUS_EPS);
}

struct Sampler {
    uint32_t index;  // halton_index
    uint32_t dim;    // next dimension
};

// Sampler::start_pixel_sample (sampler.cpp:404-422); y is the flipped image y
QZ_HD Sampler sampler_start(const SamplerParams& sp, uint32_t x, uint32_t y, uint32_t s) {
    Sampler smp;
    uint32_t idx = 0;
    if (sp.stride > 1) {
        uint32_t px = x & (QZ_MAX_HALTON_RESOLUTION - 1), py = y & (QZ_MAX_HALTON_RESOLUTION - 1);
        uint32_t off0 = 0, off1 = 0;
        for (uint32_t i = 0; i < sp.exp0; i++) { off0 = off0 * 2 + (px & 1u); px >>= 1; }
        for (uint32_t i = 0; i < sp.exp1; i++) { off1 = off1 * 3 + (py % 3u); py /= 3u; }
        idx = off0 * (sp.stride / sp.scale0) * sp.mult_inv0 + off1 * (sp.stride / sp.scale1) * sp.mult_inv1;
        idx %= sp.stride;
    }
    smp.index = idx + s * sp.stride;
    smp.dim = 2;
    return smp;
}

// Sampler::sample_pixel (sampler.cpp:449-454)
QZ_HD V2 sampler_pixel_jitter(const SamplerParams& sp, const Sampler& smp) {
    return v2(radical_inv(2, smp.index >> sp.exp0), radical_inv(3, smp.index / sp.scale1));
}

QZ_HD SamplerDim load_dim(const SamplerDim* __restrict__ table, uint32_t dim) {
#if defined(__CUDA_ARCH__)
    const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(table + dim)), r1 = __ldg(reinterpret_cast<const uint4*>(table + dim) + 1);
    SamplerDim rec;
    rec.base_nd = r0.x; rec.magic = r0.y; rec.hash = r0.z; rec.scale = r0.w;
    rec.pre_pow = r1.x; rec.pre_magic = r1.y; rec.pre_offset = r1.z; rec.pre_k = r1.w;
    return rec;
#else
    return table[dim];
#endif
}

QZ_HD float sample_dimension(const SamplerDim* __restrict__ table, const Sampler& smp, uint32_t dim) {
    const SamplerDim rec = load_dim(table, dim);
    return owen_scrambled_radical_inv(rec, sampler_prefix_array(table), smp.index);
}

// SAMPLE MEMO.  The value of a dimension depends on nothing but (dimension, Halton index), and the index depends on
// the pixel only through (x mod 128, y mod 128) (sampler_start): in an 800 x 800 image 39 pixels share every index of
// a sample number, in a 4K image 506.  The wavefront therefore keeps, per render pass, one ROW per Halton index of the
// pass -- filled ahead of the wavefront by k_memo_fill, one evaluation per (index, dimension) instead of one per pixel
// and bounce -- and the shading kernels read a bounce's draws straight from the path's row:
//   words [0, 4 * n_hot)      "hot" spectra (the lights' emission, the conductors' eta and k) at the four wavelengths
//                             of the index -- they too depend on nothing but the index (the wavelengths come from
//                             dimension 2) -- 16 bytes per spectrum (wf_shade.cuh: k_memo_spectra);
//   word  dim_off + d         the value of dimension d as a float pattern, d < dims (d = 0, 1: the pixel jitter's two
//                             radical inverses).  dim_off = 5 mod 8, so that dimension 3 -- where the first bounce's
//                             eight draws start -- begins a 32-byte sector: a bounce's draws are one sector.
// All-ones = not computed (a valid value is < 1): kept as a lazy fallback; entries only ever change from "empty" to the
// one value the reference's arithmetic gives, so the samples are the reference's bit for bit.  tab == nullptr
// (per-path replay, host emulation, small images without reuse) evaluates directly.
#define QZ_MEMO_EMPTY 0xffffffffu
#define QZ_MEMO_MAX_HOT 8
// ROWS.  A Halton index is idx_pix + s * 31104 with idx_pix one of 31104 residues, of which an image uses at most
// 128 x 128 -- and a multi-GPU shard, which owns every n-th strip of rows, only its share of the y classes.  The rows of
// a pass are therefore laid out as (sample number, class rank): `cls_rank` maps idx_pix to the rank of the (x mod 128,
// y mod 128) class among those the call owns (0xffffffff: not owned), `cls_idx` is the inverse list.
struct SampleMemo {
    uint32_t* tab;
    uint32_t dims;      // dimensions per row
    uint32_t n_cls;     // owned pixel classes
    uint32_t s_begin, s_count;   // sample numbers of the pass
    uint32_t stride;    // words per row (multiple of 8)
    uint32_t dim_off;   // word of dimension 0
    uint32_t n_hot;     // hot spectra per row
    uint32_t idx_stride, idx_magic;      // the sampler's index stride (31104) and floor(2^32 / stride)
    const uint32_t* cls_rank;            // [idx_stride]
    const uint32_t* cls_idx;             // [n_cls]: idx_pix of each owned class
    int32_t hot_id[QZ_MEMO_MAX_HOT];     // spectrum ids of the hot spectra
};
QZ_HD uint32_t* memo_row(const SampleMemo& m, uint32_t index) {
    if (!m.tab) return nullptr;
    uint32_t idx_pix;
    const uint32_t s = div_magic(index, m.idx_stride, m.idx_magic, idx_pix) - m.s_begin;
    if (s >= m.s_count) return nullptr;
#if defined(__CUDA_ARCH__)
    const uint32_t r = __ldg(m.cls_rank + idx_pix);
#else
    const uint32_t r = m.cls_rank[idx_pix];
#endif
    if (r >= m.n_cls) return nullptr;
    return m.tab + ((size_t)s * m.n_cls + r) * m.stride;
}
QZ_HD uint32_t* memo_slot(const SampleMemo& m, uint32_t dim, uint32_t index) {
    uint32_t* row = memo_row(m, index);
    if (!row || dim >= m.dims) return nullptr;
    return row + m.dim_off + dim;
}
QZ_HD uint32_t memo_load(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return p ? __ldcg(p) : QZ_MEMO_EMPTY;   // L2: the table is shared by all SMs
#else
    return p ? *p : QZ_MEMO_EMPTY;
#endif
}
QZ_HD void memo_store(uint32_t* p, float v) {
#if defined(__CUDA_ARCH__)
    if (p) __stcg(p, __float_as_uint(v));
#else
    if (p) *p = float_as_u32(v);
#endif
}
// slot of a spectrum in the rows' hot block, or -1
QZ_HD int memo_hot_slot(const SampleMemo& m, int32_t spectrum_id) {
    int slot = -1;
    for (int k = 0; k < QZ_MEMO_MAX_HOT; k++)
        if (k < (int)m.n_hot && m.hot_id[k] == spectrum_id) slot = k;
    return slot;
}
QZ_HD float sample_dimension_memo(const SamplerDim* __restrict__ table, const SampleMemo& m, uint32_t index, uint32_t dim) {
    uint32_t* p = memo_slot(m, dim, index);
    const uint32_t bits = memo_load(p);
    if (bits != QZ_MEMO_EMPTY) return u32_as_float(bits);
    Sampler smp; smp.index = index; smp.dim = 0;
    const float v = sample_dimension(table, smp, dim);
    memo_store(p, v);
    return v;
}
QZ_HD V2 sampler_pixel_jitter_memo(const SamplerParams& sp, const SampleMemo& m, const Sampler& smp) {
    uint32_t* p0 = memo_slot(m, 0, smp.index);
    uint32_t* p1 = memo_slot(m, 1, smp.index);
    const uint32_t b0 = memo_load(p0), b1 = memo_load(p1);
    V2 j;
    if (b0 != QZ_MEMO_EMPTY) j.x = u32_as_float(b0);
    else { j.x = radical_inv(2, smp.index >> sp.exp0); memo_store(p0, j.x); }
    if (b1 != QZ_MEMO_EMPTY) j.y = u32_as_float(b1);
    else { j.y = radical_inv(3, smp.index / sp.scale1); memo_store(p1, j.y); }
    return j;
}

// what the eager fill (wf_stage.cuh: k_memo_fill) stores for dimension `dim` of the row (sample number s_begin + s, class
// `cls`): the two radical inverses of the pixel jitter for dimensions 0 and 1, the Owen-scrambled value otherwise
QZ_HD uint32_t memo_entry_bits(const SamplerDim* __restrict__ table, const SamplerParams& spar, const SampleMemo& m, uint32_t s, uint32_t cls,
                               uint32_t dim) {
    Sampler smp;
    smp.index = m.cls_idx[cls] + (m.s_begin + s) * m.idx_stride;
    smp.dim = 0;
    float v;
    if (dim == 0) v = radical_inv(2, smp.index >> spar.exp0);
    else if (dim == 1) v = radical_inv(3, smp.index / spar.scale1);
    else v = sample_dimension(table, smp, dim);
    return float_as_u32(v);
}

// two independent dimensions evaluated in one loop: the digit chains of the two dimensions do
// not depend on each other, so interleaving them doubles the instruction-level parallelism of
// what is otherwise one long dependent integer chain per digit
QZ_HD V2 owen_scrambled_radical_inv_pair(const SamplerDim& r0, const SamplerDim& r1, const uint16_t* prefix, uint32_t a) {
    const uint32_t b0 = r0.base_nd & 0xffffu, b1 = r1.base_nd & 0xffffu;
    const uint32_t n0 = r0.base_nd >> 16, n1 = r1.base_nd >> 16;
    uint32_t w0 = b0 - 1, w1 = b1 - 1;
    w0 |= w0 >> 1; w0 |= w0 >> 2; w0 |= w0 >> 4; w0 |= w0 >> 8; w0 |= w0 >> 16;
    w1 |= w1 >> 1; w1 |= w1 >> 2; w1 |= w1 >> 4; w1 |= w1 >> 8; w1 |= w1 >> 16;
    uint64_t rev0 = 0, rev1 = 0;
    uint32_t a0 = a, a1 = a, k0 = 0, k1 = 0;
    if (r0.pre_pow) {
        uint32_t lo;
        a0 = div_magic(a, r0.pre_pow, r0.pre_magic, lo);
#if defined(__CUDA_ARCH__)
        rev0 = __ldg(prefix + r0.pre_offset + lo);
#else
        rev0 = prefix[r0.pre_offset + lo];
#endif
        k0 = r0.pre_k;
    }
    if (r1.pre_pow) {
        uint32_t lo;
        a1 = div_magic(a, r1.pre_pow, r1.pre_magic, lo);
#if defined(__CUDA_ARCH__)
        rev1 = __ldg(prefix + r1.pre_offset + lo);
#else
        rev1 = prefix[r1.pre_offset + lo];
#endif
        k1 = r1.pre_k;
    }
    // the two chains advance together while both have digits left
    while (k0 < n0 && k1 < n1) {
        uint32_t d0, d1;
        a0 = div_magic(a0, b0, r0.magic, d0);
        a1 = div_magic(a1, b1, r1.magic, d1);
        uint32_t h0 = (uint32_t)mix_bits((uint64_t)r0.hash ^ rev0);
        uint32_t h1 = (uint32_t)mix_bits((uint64_t)r1.hash ^ rev1);
        d0 = permutation_element(d0, b0, r0.magic, w0, h0);
        d1 = permutation_element(d1, b1, r1.magic, w1, h1);
        rev0 = rev0 * b0 + d0;
        rev1 = rev1 * b1 + d1;
        k0++; k1++;
    }
    if (k0 < n0) rev0 = owen_digits(r0, a0, k0, n0, rev0);
    if (k1 < n1) rev1 = owen_digits(r1, a1, k1, n1, rev1);
    return v2(std_min(u32_as_float(r0.scale) * (float)rev0, QZ_ONE_MINUS_EPS),
              std_min(u32_as_float(r1.scale) * (float)rev1, QZ_ONE_MINUS_EPS));
}

// Sampler::sample_1d / sample_2d (sampler.cpp:433-447): dimensions wrap to 2 past the table.
// The *_skip variants advance the dimension counter exactly like the real calls but do not
// evaluate the value: the reference draws several samples whose value it never reads (the
// material-select sample of non-mixed materials, the light pick with a single light, the 2-D
// sample of point lights and specular BxDFs, the 1-D sample of non-dielectric BxDFs, the
// roulette sample when roulette does not apply); skipping the arithmetic is exact.
QZ_HD uint32_t sample_1d_skip(Sampler& smp) {
    if (smp.dim >= QZ_N_PRIMES) smp.dim = 2;
    return smp.dim++;
}
QZ_HD uint32_t sample_2d_skip(Sampler& smp) {
    if (smp.dim + 1 >= QZ_N_PRIMES) smp.dim = 2;
    uint32_t d = smp.dim;
    smp.dim += 2;
    return d;
}
QZ_HD float sample_1d(const SamplerDim* __restrict__ table, Sampler& smp) {
    return sample_dimension(table, smp, sample_1d_skip(smp));
}
QZ_HD V2 sample_2d(const SamplerDim* __restrict__ table, Sampler& smp) {
    const uint32_t d = sample_2d_skip(smp);
    return owen_scrambled_radical_inv_pair(load_dim(table, d), load_dim(table, d + 1), sampler_prefix_array(table), smp.index);
}

// ---- host-side table construction (runs once per process) ---------------------------
inline void build_sampler_table(SamplerDim* out /* QZ_N_PRIMES entries */) {
    uint32_t count = 0;
    for (uint32_t n = 2; count < QZ_N_PRIMES; n++) {
        bool prime = true;
        for (uint32_t d = 2; d * d <= n; d++) if (n % d == 0) { prime = false; break; }
        if (!prime) continue;
        const uint32_t dim = count++;
        // digit count and scale: the loop of sampler.cpp:341-350 in float
        float inv_base = 1.0f / (float)n, inv_base_m = 1.0f;
        uint32_t nd = 0;
        while (1.0f - inv_base_m < 1.0f) { inv_base_m *= inv_base; nd++; }
        SamplerDim rec;
        rec.base_nd = n | (nd << 16);
        rec.magic = (uint32_t)((1ull << 32) / n);
        rec.hash = (uint32_t)mix_bits((uint64_t)(1 + (int)(dim << 4)));
        rec.scale = float_as_u32(inv_base_m);
        rec.pre_pow = 0; rec.pre_magic = 0; rec.pre_offset = 0; rec.pre_k = 0;
        out[dim] = rec;
    }
}

// Lays the prefix tables of dimensions [2, n_dims) out behind each other: per dimension the
// largest k < ndigits with base^k <= cap.  Returns the total number of 16-bit entries.
inline uint32_t plan_sampler_prefix(SamplerDim* recs, uint32_t n_dims, uint32_t cap) {
    uint32_t total = 0;
    for (uint32_t dim = 2; dim < n_dims && dim < QZ_N_PRIMES; dim++) {
        SamplerDim& r = recs[dim];
        const uint32_t base = r.base_nd & 0xffffu, nd = r.base_nd >> 16;
        uint32_t k = 0, pw = 1;
        while (k + 1 < nd && (uint64_t)pw * base <= cap && (uint64_t)pw * base <= 65536ull) { pw *= base; k++; }
        if (k == 0) continue;
        r.pre_pow = pw;
        r.pre_magic = (uint32_t)((1ull << 32) / pw);
        r.pre_offset = total;
        r.pre_k = k;
        total += pw;
    }
    return total;
}

// entry j of a dimension's prefix table
QZ_HD uint16_t sampler_prefix_entry(const SamplerDim& rec, uint32_t j) {
    return (uint16_t)owen_digits(rec, j, 0, rec.pre_k, 0);
}

inline SamplerParams make_sampler_params(int x_res, int y_res) {
    SamplerParams sp;
    const int res[2] = {x_res, y_res};
    uint32_t scales[2], exps[2];
    for (int i = 0; i < 2; i++) {
        int base = i == 0 ? 2 : 3;
        int64_t scale = 1, e = 0;
        int lim = res[i] < QZ_MAX_HALTON_RESOLUTION ? res[i] : QZ_MAX_HALTON_RESOLUTION;
        while (scale < lim) { scale *= base; e++; }
        scales[i] = (uint32_t)scale; exps[i] = (uint32_t)e;
    }
    // multiplicative inverses by brute force (sampler.cpp:354-376 uses extended Euclid)
    auto minv = [](uint32_t a, uint32_t n) -> uint32_t {
        if (n == 1) return 0;
        for (uint32_t x = 0; x < n; x++) if ((uint64_t)a * x % n == 1) return x;
        return 0;
    };
    sp.scale0 = scales[0]; sp.scale1 = scales[1];
    sp.exp0 = exps[0]; sp.exp1 = exps[1];
    sp.mult_inv0 = minv(scales[1] % scales[0], scales[0]);
    sp.mult_inv1 = minv(scales[0] % scales[1], scales[1]);
    sp.stride = scales[0] * scales[1];
    sp.pad = 0;
    return sp;
}

}  // namespace qz
