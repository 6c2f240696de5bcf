// Arithmetic-free stages of the wavefront (wf_types.cuh): the queue builder, the sampler stage and the film.
#pragma once

#include "wf_types.cuh"

namespace qz {

// Ordered compaction of the per-slot family tags into the shade queues: tag == q selects queue q.
// A block owns 2048 consecutive slots, each thread 8 consecutive ones, so a queue's entries
// ascend within every block's share; one atomicAdd per block and queue reserves the share.
template <int NQ>
__global__ void __launch_bounds__(256) k_bin(WfBuffers b) {
    // queue NQ is the late queue (tag bit QZ_FAM_LATE): a slot can be in its family queue and in the late queue
    constexpr int NQL = NQ + 1;
    __shared__ uint32_t s_warp[8][NQL];
    __shared__ uint32_t s_base[NQL];
    const uint8_t* __restrict__ tags = b.fam;
    const uint32_t n_slots = b.pool;
    uint32_t* counters = b.counters + C_SHADE0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t tile = blockIdx.x * 2048u; tile < n_slots; tile += gridDim.x * 2048u) {
        const uint32_t first_slot = tile + threadIdx.x * 8u;
        uint32_t tg[8];
        if (first_slot + 8u <= n_slots) {
            const uint2 raw = *reinterpret_cast<const uint2*>(tags + first_slot);
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) tg[k] = ((k < 4 ? raw.x : raw.y) >> (8 * (k & 3))) & 0xffu;
        } else {
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) tg[k] = first_slot + k < n_slots ? tags[first_slot + k] : QZ_FAM_NONE;
        }
        // membership masks (bit k = this thread's k-th slot) and per-thread counts
        uint32_t member[NQL], cnt[NQL];
#pragma unroll
        for (int q = 0; q < NQL; q++) {
            uint32_t m = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
                const bool in = tg[k] != QZ_FAM_NONE && (q < NQ ? (tg[k] & 0x0fu) == (uint32_t)q : (tg[k] & QZ_FAM_LATE) != 0u);
                m |= (in ? 1u : 0u) << k;
            }
            member[q] = m;
            cnt[q] = __popc(m);
        }
        // exclusive scan of the counts over the block, per queue
        uint32_t excl[NQL];
#pragma unroll
        for (int q = 0; q < NQL; q++) {
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
        if (threadIdx.x < NQL) {
            uint32_t total = 0;
            for (int w = 0; w < 8; w++) { const uint32_t c = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = total; total += c; }
            uint32_t* counter = threadIdx.x < NQ ? &counters[threadIdx.x] : &b.counters[C_LATE];
            s_base[threadIdx.x] = total ? atomicAdd(counter, total) : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NQL; q++) {
            uint32_t* queue = q < NQ ? b.q_shade[q] : b.q_late;
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

// The sampler as its own wavefront stage.  The Owen-scrambled Halton evaluation is a long, purely integer,
// dependent chain per digit (~75 instructions); inside the shading kernels it would run at low occupancy next to
// float-heavy code.  Here ONE THREAD evaluates the draws of ONE PATH-BOUNCE, role after role: the 32 lanes of a warp
// work on 32 consecutive entries of a shade queue, so they evaluate the same role at the same time -- in a first-hit
// queue even the same dimension, hence the same base and digit count -- and each path's state is read once
// (one 32-byte sector) and its eight draws leave as one 32-byte store: 64 bytes per bounce.  (The first version
// used one thread per draw: up to seven threads re-read a path's state and each wrote one scalar -- 2.8 x the
// bytes, profiles/r01_summary.md.)  k_shade then reads eight floats per path instead of running the sampler.
// Thread 0 of the stage also does the pipeline's per-iteration bookkeeping (queue lengths -> statistics and the
// termination count, traversal cursors back to zero): no traversal kernel runs concurrently on this stream.
__global__ void __launch_bounds__(256) k_sample(DScene sc, WfBuffers b) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t shade = 0;
        for (int q = 0; q < SQ_COUNT; q++) shade += b.counters[C_SHADE0 + q];
        atomicAdd(&b.stats[S_SHADE], (unsigned long long)shade);
        b.counters[C_ACTIVE] = shade;
        // the queues of this iteration are the next iteration's work list (wf_types.cuh, WorkList)
        for (int q = 0; q < SQ_COUNT; q++) b.counters[C_WORK0 + q] = b.counters[C_SHADE0 + q];
        b.counters[C_CURSOR_TRACE] = 0;
        b.counters[C_CURSOR_SHADOW] = 0;
    }
    // only the bounces beyond the memo row (the late queue, built by k_bin) get their draws here; all others read them
    // from the row in k_shade
    const uint32_t total = b.counters[C_LATE];
    const bool has_lights = sc.n_lights != 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t slot = b.q_late[t];
        const int q = (int)(b.fam[slot] & 0x0fu);
        if (!queue_roles(q, sc.n_lights)) continue;   // run-time dispatch: evaluated on the fly inside k_shade<KH_ANY>
        const uint4 misc = b.misc.get(slot);
        const int fam = q % SQ_FAMILIES;
        bool nee = fam == SQ_DIFFUSE;
        if (fam == SQ_CONDUCTOR) {
            const uint32_t geom = __float_as_uint(b.hit_b.get(slot).w);
            const qz_material mat = sc.materials[sc.geoms[geom].material];
            nee = !(mat.alpha_x < 1e-3f && mat.alpha_y < 1e-3f);
        }
        uint32_t dims[R_COUNT];
        bounce_dims(misc.z & 0xffffu, nee, has_lights, dims);
        const uint32_t roles = queue_roles(q, sc.n_lights);
        // the roles' values: first from the pass's memo table (all loads in flight together), then the digit loops for
        // the entries nobody has computed yet -- one copy of the loops in the instruction stream, the role's dimension and
        // result register are selected
        float v[R_COUNT];
        uint32_t missing = 0;
#pragma unroll
        for (int k = 0; k < R_COUNT; k++) {
            v[k] = 0.0f;
            if ((roles >> k) & 1u) {
                const uint32_t bits = memo_load(memo_slot(sc.memo, dims[k], misc.y));
                if (bits != QZ_MEMO_EMPTY) v[k] = __uint_as_float(bits);
                else missing |= 1u << k;
            }
        }
        Sampler smp;
        smp.index = misc.y; smp.dim = 0;
#pragma unroll 1
        for (uint32_t m = missing; m; m &= m - 1u) {
            const int r = __ffs(m) - 1;
            uint32_t dim = dims[0];
#pragma unroll
            for (int k = 1; k < R_COUNT; k++) dim = k == r ? dims[k] : dim;
            const float val = sample_dimension(sc.sampler_table, smp, dim);
            memo_store(memo_slot(sc.memo, dim, misc.y), val);
#pragma unroll
            for (int k = 0; k < R_COUNT; k++) v[k] = k == r ? val : v[k];
        }
        float4* out = b.samples.at(slot);
        __stcg(out, make_float4(v[0], v[1], v[2], v[3]));
        __stcg(out + 1, make_float4(v[4], v[5], v[6], v[7]));
    }
}

// Fills the dimension words of the pass's sample memo (sampler.cuh) ahead of the wavefront: every sample number of the
// pass, every (x mod 128, y mod 128) class the call owns, every dimension.  A CTA of eight warps takes 32 neighbouring
// classes of one sample number and eight consecutive words of their rows: warp w evaluates ONE dimension for the 32
// classes -- same base, same digit count, no divergence but the permutation's rejection loop -- and the CTA then writes
// each row's eight words as one 32-byte sector.  Each value is computed once instead of once per pixel that shares
// the index (39 pixels at 800 x 800) and bounce.  (Filled lazily by the paths themselves, the 39 first users of an
// entry all ran in the same wavefront iteration and nearly every warp of the sampler stage dragged a few missing
// lanes through the digit loops: profiles/r02_summary.md.)
__global__ void __launch_bounds__(256) k_memo_fill(const SamplerDim* __restrict__ table, SamplerParams spar, SampleMemo memo) {
    __shared__ uint32_t s_val[32][9];
    const uint32_t ncls = memo.n_cls;
    const uint32_t cls_groups = (ncls + 31u) / 32u;
    const uint32_t word0 = memo.dim_off - 5u;                       // first word of the dimension region, 32-byte aligned
    const uint32_t n_chunks = (memo.dims + 5u + 7u) / 8u;
    const uint64_t total = (uint64_t)memo.s_count * cls_groups * n_chunks;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
        // chunk-major: consecutive CTAs work on the same dimensions
        const uint32_t chunk = (uint32_t)(item / ((uint64_t)memo.s_count * cls_groups));
        const uint32_t rem = (uint32_t)(item - (uint64_t)chunk * memo.s_count * cls_groups);
        const uint32_t s = rem / cls_groups, cls0 = (rem - s * cls_groups) * 32u, cls = cls0 + lane;
        const int dim = (int)(chunk * 8u + warp) - 5;
        uint32_t bits = QZ_MEMO_EMPTY;
        if (cls < ncls && dim >= 0 && dim < (int)memo.dims) bits = memo_entry_bits(table, spar, memo, s, cls, (uint32_t)dim);
        // (tried: every lane storing its own word, no transpose and no barrier, leaving the merge to L2 -- step 91.5 -> 92.3 ms)
        s_val[lane][warp] = bits;
        __syncthreads();
        if (threadIdx.x < 64u) {
            const uint32_t r = threadIdx.x >> 1, half = threadIdx.x & 1u;
            const uint32_t w = word0 + chunk * 8u + half * 4u;
            if (cls0 + r < ncls && w + 4u <= memo.stride) {
                uint32_t* row = memo.tab + ((size_t)s * ncls + cls0 + r) * memo.stride;
                __stcg(reinterpret_cast<uint4*>(row + w), make_uint4(s_val[r][half * 4u], s_val[r][half * 4u + 1u], s_val[r][half * 4u + 2u], s_val[r][half * 4u + 3u]));
            }
        }
        __syncthreads();
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
            const float4 ra = __ldcs(b.res_a + cell), rb = __ldcs(b.res_b + cell);
            const float rc = __ldcs(b.res_c + cell);
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

// Image::save's tone path on the device film (image.cpp:7-19): 255 * powf(c, gamma) per channel with the channel order
// reversed (the BGR float image the reference hands to cv::imwrite), and optionally that image saturated to 8 bits the
// way OpenCV stores a CV_32F matrix into an 8-bit file (saturate_cast<uchar>: round to nearest even, clamp to 0..255).
// The per-channel arithmetic is tone_value / tone_u8 (math.cuh).
__global__ void __launch_bounds__(256) k_tone(const float* __restrict__ rgb, uint32_t n_pixels, float gamma, float* __restrict__ bgr255,
                                               uint8_t* __restrict__ bgr8) {
    const uint32_t n = n_pixels * 3u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t pix = i / 3u, ch = i - 3u * pix;
        const float c = rgb[3u * pix + (2u - ch)];
        const float v = tone_value(c, gamma);
        if (bgr255) bgr255[i] = v;
        if (bgr8) bgr8[i] = tone_u8(v);
    }
}

}  // namespace qz
