// One bounce of the reference's path loop (render.cpp:107-209 sample_pixel, :54-87
// sample_lights, :45-52 power_heuristic) plus what it calls: surface interaction from a hit
// (scene.cpp:61-117), emission (interaction.hpp:32-37, light.cpp:48-53), materials
// (material.cpp:4-60, material.hpp:94-97 MixedMaterial), textures (texture.cpp:5-61) and
// light sampling (light.cpp:8-42, shape.hpp:46-49,78-83, scene.cpp:119-134).
//
// The wavefront kernels and the per-path replay kernel both call shade_bounce(); the only
// difference between them is where the path state lives between bounces.
#pragma once

#include "bvh.cuh"
#include "bxdf.cuh"
#include "spectra.cuh"

namespace qz {

// ------------------------------------------------------------------ path state
#define QZ_FLAG_SPECULAR_BOUNCE 1u
#define QZ_FLAG_HAS_SHADOW 2u     /* a next-event shadow ray is pending for this bounce */

struct PathState {
    Ray ray;
    Spec4 weight, L, lambda, pdf;
    Sampler smp;
    uint32_t depth;
    uint32_t flags;
    float p_b;
    float ior_scale;
    uint32_t n_rays;  // closest-hit + occlusion queries issued (the oracle's per-path ray count)
};

struct PathAov {
    V3 normal;
    Spec4 albedo;
};

struct ShadowRequest {
    V3 o, d;         // Scene::occluded(start, end): ray start -> end - start (scene.cpp:136-138)
    Spec4 contrib;   // weight * sample_lights(...) assuming the light is visible
};

QZ_HD float power_heuristic(float f_pdf, float g_pdf) {
    float f = 1 * f_pdf;
    float g = 1 * g_pdf;
    if (is_inf(f * f)) return 1.0f;
    return r_div(f * f, r_fma(f, f, g * g));
}

// scene.cpp:27-39
QZ_HD V2 sphere_uv(V3 n) {
    float phi = (float)((double)qz_atan2f(n.z, n.x) + QZ_PI);
    float u = (float)((double)phi / (2.0 * QZ_PI));
    if (u >= 1.0f) u -= 1.1920929e-07f;
    float theta = qz_acosf(n.y);
    float v = (float)((double)theta / QZ_PI);
    if (v >= 1.0f) v -= 1.1920929e-07f;
    return v2(u, v);
}

// from_spectrum at the PATH's wavelengths for a spectrum that may be "hot" (sampler.cuh, SAMPLE MEMO: the lights'
// emission, the conductors' eta and k): the four values depend on the Halton index only and sit in the path's memo
// row, computed once per index by k_memo_spectra with this very function -- the same bits as evaluating here.
QZ_HD Spec4 from_spectrum_path(const DScene& sc, uint32_t index, int32_t id, const Spec4& lambda) {
    if (sc.memo.n_hot) {
        const int slot = memo_hot_slot(sc.memo, id);
        const uint32_t* row = memo_row(sc.memo, index);
        if (slot >= 0 && row) {
#if defined(__CUDA_ARCH__)
            const float4 v = __ldcg(reinterpret_cast<const float4*>(row) + slot);
            return spec4(v.x, v.y, v.z, v.w);
#else
            const uint32_t* w = row + 4 * slot;   // (host emulation, tests/emu)
            return spec4(u32_as_float(w[0]), u32_as_float(w[1]), u32_as_float(w[2]), u32_as_float(w[3]));
#endif
        }
    }
    return from_spectrum(sc, id, lambda);
}

// The hot-spectra block of one memo row (wf_shade.cuh: k_memo_spectra): the row's wavelengths come from its dimension-2
// value exactly as in start_path; every hot spectrum is evaluated by from_spectrum, which never consults the memo.
QZ_HD void memo_fill_hot(const DScene& sc, uint32_t* row) {
    const SampleMemo& memo = sc.memo;
    const float ul = u32_as_float(memo_load(row + memo.dim_off + 2));
    Spec4 lambda, pdf;
    sample_wavelengths(ul, lambda, pdf);
    for (uint32_t k = 0; k < memo.n_hot; k++) {
        const Spec4 v = from_spectrum(sc, memo.hot_id[k], lambda);
#if defined(__CUDA_ARCH__)
        __stcg(reinterpret_cast<float4*>(row) + k, make_float4(v.v[0], v.v[1], v.v[2], v.v[3]));
#else
        for (int c = 0; c < 4; c++) row[4 * k + c] = float_as_u32(v.v[c]);
#endif
    }
}

// Texture::value (texture.cpp)
QZ_HD Spec4 texture_value(const DScene& sc, int32_t tex_id, V2 uv, const Spec4& lambda) {
    const qz_texture t = sc.textures[tex_id];
    if (t.kind == QZ_TEX_SOLID) return from_spectrum(sc, t.a, lambda);
    if (t.kind == QZ_TEX_DUMMY) {
        float u = uv.x * 10.0f;  // float(double(uv.x) * 10.) is the same value
        float v = uv.y * 10.0f;
        int cell = (int)(floorf(u) + floorf(v));
        return from_spectrum(sc, (cell % 2 == 0) ? t.a : t.b, lambda);
    }
    // ImageTexture: nearest texel, RGB -> sigmoid spectrum through the table on every lookup
    uint32_t x = (uint32_t)(uv.x * (float)t.width);
    if (x == t.width) x = t.width - 1;
    uint32_t y = (uint32_t)(uv.y * (float)t.height);
    if (y == t.height) y = t.height - 1;
    const float* px = sc.pool + t.offset + 3 * ((size_t)y * t.width + x);
    V3 c = rgb_to_sigmoid(sc, px[0], px[1], px[2]);
    return spec4(sigmoid_poly<true>(c.x, c.y, c.z, lambda.v[0]), sigmoid_poly<true>(c.x, c.y, c.z, lambda.v[1]),
                 sigmoid_poly<true>(c.x, c.y, c.z, lambda.v[2]), sigmoid_poly<true>(c.x, c.y, c.z, lambda.v[3]));
}

// material kind after resolving MixedMaterial with the bounce's material sample
// (material.hpp:94-97: child floor(sample * N), the same sample passed down)
QZ_HD int32_t resolve_material(const DScene& sc, int32_t mat, float sample) {
    for (int guard = 0; guard < 8; guard++) {
        const qz_material m = sc.materials[mat];
        if (m.kind != QZ_MAT_MIXED) return mat;
        uint32_t idx = (uint32_t)(sample * (float)m.count);
        mat = sc.mixed_children[m.a + (int32_t)idx];
    }
    return mat;
}

struct SurfacePoint {
    V3 point, wo, normal;
    V2 uv;
    int32_t material, light;
    bool is_sphere;
};

// Scene::ray_intersect after the Embree call (scene.cpp:72-116)
QZ_HD SurfacePoint make_surface_point(const DScene& sc, const Ray& ray, const Hit& hit) {
    SurfacePoint sp;
    const uint32_t prim_id = hit.prim_id;
    const qz_geometry g = sc.geoms[hit.geom_id];
    sp.uv = v2(hit.u, hit.v);
    if (g.normal_offset >= 0 && g.shape == QZ_SHAPE_OBJ) {
        const int32_t* fi = sc.nidx + 4 * ((size_t)g.nindex_offset + prim_id);
        V3 n[4];
        for (int k = 0; k < 4; k++) {
            const float* p = sc.normals + 3 * ((size_t)g.normal_offset + (size_t)fi[k]);
            n[k] = v3(p[0], p[1], p[2]);
        }
        float u = hit.u, v = hit.v;
        V3 blend = n[0] * ((1.0f - u) * (1.0f - v)) + n[1] * (u * (1.0f - v)) + n[2] * (u * v) + n[3] * ((1.0f - u) * v);
        sp.normal = normalized(blend);
    } else {
        sp.normal = normalized(hit.ng);
    }
    sp.is_sphere = g.shape == QZ_SHAPE_SPHERE;
    sp.point = ray.o + ray.d * hit.t;
    sp.wo = normalized(-ray.d);
    sp.material = g.material;
    sp.light = g.light;
    return sp;
}

// Material::bsdf for a resolved (non-mixed) material; may terminate secondary wavelengths
template <int KH>
QZ_HD Bsdf make_bsdf(const DScene& sc, int32_t mat, SurfacePoint& sp, const Spec4& lambda, Spec4& pdf, uint32_t index) {
    const qz_material m = sc.materials[mat];
    Bsdf f;
    f.a = spec4(0.0f); f.b = spec4(0.0f); f.rough.ax = 0.0f; f.rough.ay = 0.0f; f.ior = 1.0f;
    make_basis(sp.normal, f.u0, f.u1, f.u2);
    if (KH == KH_DIFFUSE || (KH == KH_ANY && m.kind == QZ_MAT_DIFFUSE)) {
        f.kind = BX_DIFFUSE;
        // sphere uv are only ever consumed by textures: evaluate them lazily
        const uint32_t tk = sc.textures[m.a].kind;
        if (sp.is_sphere && tk != QZ_TEX_SOLID) sp.uv = sphere_uv(sp.normal);
        f.a = texture_value(sc, m.a, sp.uv, lambda);
    } else if (KH == KH_CONDUCTOR || (KH == KH_ANY && m.kind == QZ_MAT_CONDUCTOR)) {
        f.kind = BX_CONDUCTOR;
        f.a = from_spectrum_path(sc, index, m.a, lambda);
        f.b = from_spectrum_path(sc, index, m.b, lambda);
        f.rough.ax = m.alpha_x; f.rough.ay = m.alpha_y;
    } else {
        f.kind = m.kind == QZ_MAT_DIELECTRIC ? BX_DIELECTRIC : BX_THIN;
        f.ior = eval_spectrum(sc, m.a, lambda.v[0]);
        if (!m.is_constant) terminate_secondary(pdf);
    }
    return f;
}

// Light emission towards w from a point with normal n (light.cpp:48-53)
QZ_HD Spec4 light_emission(const DScene& sc, const qz_light& l, V3 n, V3 w, const Spec4& lambda, uint32_t index) {
    if (!l.two_sided && dot(n, w) < 0.0f) return spec4(0.0f);
    return from_spectrum_path(sc, index, l.spectrum, lambda) * l.scale;
}

// ------------------------------------------------------------------ where a bounce's samples come from
// The draws of one bounce, in the order the reference makes them (render.cpp:141-200):
//   material select (1), [light pick (1), light uv (2)] if the BSDF is not specular, BSDF uv (2),
//   BSDF u1 (1), roulette (1).  shade_bounce() does the dimension bookkeeping itself and asks a
//   SOURCE for the value of a role at a dimension:
//   * SamplesOnTheFly evaluates the Owen-scrambled dimension right there (per-path replay, the
//     run-time-dispatch kernels, the host emulation);
//   * SamplesPrecomputed returns what k_sample (wf_stage.cuh) computed for this bounce in a
//     separate, fully convergent, high-occupancy integer kernel.
enum SampleRole { R_MAT = 0, R_PICK = 1, R_LIGHT = 2 /* and 3 */, R_BSDF = 4 /* and 5 */, R_U1 = 6, R_RR = 7, R_COUNT = 8 };

struct SamplesOnTheFly {
    const SamplerDim* tab;
    uint32_t index;
    SampleMemo memo;   // the wavefront's table of values already computed this pass (sampler.cuh); none in the replay
    QZ_HD float one(int, uint32_t dim) const { return sample_dimension_memo(tab, memo, index, dim); }
    QZ_HD V2 two(int, uint32_t dim) const {
        if (memo.tab) return v2(sample_dimension_memo(tab, memo, index, dim), sample_dimension_memo(tab, memo, index, dim + 1));
        return owen_scrambled_radical_inv_pair(load_dim(tab, dim), load_dim(tab, dim + 1), sampler_prefix_array(tab), index);
    }
};

struct SamplesPrecomputed {
    float v[R_COUNT];
    QZ_HD float one(int role, uint32_t) const { return v[role]; }
    QZ_HD V2 two(int role, uint32_t) const { return v2(v[role], v[role + 1]); }
};

// A bounce that starts at dimension d0 draws at most nine dimensions (bounce_dims below: nine in a scene without
// lights).  When all of them lie in the path's memo row (sampler.cuh) the shading kernel reads them there and the
// sampler stage skips the bounce; later bounces go through k_sample and the record as before.
QZ_HD bool bounce_in_memo(const SampleMemo& m, uint32_t index, uint32_t d0) {
    (void)index;   // every path of the pass has a row: its sample number and pixel class are the pass's and the call's
    return m.tab && d0 + 9u <= m.dims;
}

// (a new path's first bounce starts at dimension 3, and every Halton index of the pass has a row)
QZ_HD bool first_bounce_in_memo(const SampleMemo& m) { return m.tab && 3u + 9u <= m.dims; }

// dimensions of every role of a bounce that starts at dimension d0 (same skip calls as shade_bounce)
QZ_HD void bounce_dims(uint32_t d0, bool nee, bool has_lights, uint32_t dims[R_COUNT]) {
    Sampler t; t.index = 0; t.dim = d0;
    for (int k = 0; k < R_COUNT; k++) dims[k] = 2;
    dims[R_MAT] = sample_1d_skip(t);
    if (nee) {
        if (has_lights) dims[R_PICK] = sample_1d_skip(t);
        const uint32_t dl = sample_2d_skip(t);
        dims[R_LIGHT] = dl; dims[R_LIGHT + 1] = dl + 1;
        if (!has_lights) sample_2d_skip(t);
    }
    const uint32_t db = sample_2d_skip(t);
    dims[R_BSDF] = db; dims[R_BSDF + 1] = db + 1;
    dims[R_U1] = sample_1d_skip(t);
    dims[R_RR] = sample_1d_skip(t);
}

// sample_lights (render.cpp:54-87) without the occlusion test: returns the contribution for a
// visible light and the shadow segment to test; `has_shadow` false means the term is zero
template <int KH, class SRC>
QZ_HD void sample_lights(const DScene& sc, const SurfacePoint& sp, const Bsdf& f, const Spec4& lambda, Sampler& smp,
                         const SRC& src, bool& has_shadow, V3& p_light_out, Spec4& result) {
    has_shadow = false;
    result = spec4(0.0f);
    int32_t li = -1;
    float proba = 0.0f;
    if (sc.n_lights != 0) {
        // with a single light the pick is index 0 whatever the sample says: draw without evaluating
        const uint32_t dp = sample_1d_skip(smp);
        if (sc.n_lights == 1) li = 0;
        else li = (int32_t)(uint32_t)(src.one(R_PICK, dp) * (float)sc.n_lights);
        proba = 1.0f / (float)sc.n_lights;
    }
    if (li < 0) {
        sample_2d_skip(smp);
        sample_2d_skip(smp);  // keeps the dimension count even (render.cpp:62-66)
        return;
    }
    // point lights never read their 2-D sample (light.cpp:8-17)
    V2 u2 = v2(0.0f, 0.0f);
    const uint32_t dl = sample_2d_skip(smp);
    if (sc.lights[li].kind != QZ_LIGHT_POINT) u2 = src.two(R_LIGHT, dl);
    const qz_light l = sc.lights[li];
    const V3 lp = v3(l.p[0], l.p[1], l.p[2]);
    Spec4 spec;
    V3 wi, p_light;
    float pdf;
    if (l.kind == QZ_LIGHT_POINT) {
        V3 dv = lp - sp.point;
        wi = r_normalized(dv);
        spec = from_spectrum_path(sc, smp.index, l.spectrum, lambda) * r_div(l.scale, norm_squared(dv));
        pdf = 1.0f;
        p_light = lp;
    } else {
        V3 n;
        if (l.kind == QZ_LIGHT_AREA_QUAD) {
            p_light = lp + v3(l.du[0], l.du[1], l.du[2]) * u2.x + v3(l.dv[0], l.dv[1], l.dv[2]) * u2.y;
            n = v3(l.normal[0], l.normal[1], l.normal[2]);
        } else {
            n = sample_uniform_sphere(u2);
            p_light = lp + n * l.radius;
        }
        pdf = l.inv_area;
        V3 dv = p_light - sp.point;
        if (pdf == 0.0f || norm_squared(dv) == 0.0f) return;
        wi = r_normalized(dv);
        spec = light_emission(sc, l, n, -wi, lambda, smp.index);
        if (is_zero(spec)) return;
    }
    if (is_zero(spec) || pdf == 0.0f) return;
    Spec4 fv = bsdf_f<KH>(f, sp.wo, wi) * fabsf(dot(wi, sp.normal));
    if (is_zero(fv)) return;
    float p_l = proba * pdf;
    if (l.kind != QZ_LIGHT_POINT) {
        float p_b = bsdf_pdf<KH>(f, sp.wo, wi);
        float w_l = power_heuristic(p_l, p_b);
        result = spec * fv * r_div(w_l, p_l);
    } else {
        result = r_div(spec * fv, p_l);
    }
    has_shadow = true;
    p_light_out = p_light;
}

// One iteration of the while loop of sample_pixel() AFTER the closest-hit query.
// Returns true when the path continues (ps.ray holds the next ray).
// FIRST: 1 = the path is known to be at depth 0, 0 = known to be deeper, -1 = decide at run time.
// The wavefront sorts first hits into their own queues: only they pay for the 16-sample albedo
// estimate, and the kernels for deeper bounces do not even contain that code.
// Radiance picked up AT the hit (emitter or background) is returned in `gain` (has_gain) rather
// than added to ps.L: a bounce adds at most one such term, and the wavefront kernels only touch
// the radiance buffer when there is one.  ps.L is neither read nor written here.
// ALBEDO_STAGE: the depth-0 albedo of (non-mixed) conductors is estimated by a separate kernel (k_albedo_conductor).
template <int KH, int FIRST, bool ALBEDO_STAGE, class SRC>
QZ_HD bool shade_bounce(const DScene& sc, PathState& ps, PathAov& aov, const Hit& hit, uint32_t max_bounces,
                        ShadowRequest& shadow, const SRC& src, Spec4& gain, bool& has_gain) {
    const bool first = FIRST < 0 ? ps.depth == 0 : FIRST != 0;
    ps.flags &= ~QZ_FLAG_HAS_SHADOW;
    has_gain = false;
    if (hit.prim == QZ_NO_HIT) {
        if (sc.bg_spectrum >= 0) { gain = ps.weight * from_spectrum(sc, sc.bg_spectrum, ps.lambda) * sc.bg_scale; has_gain = true; }
        return false;
    }
    SurfacePoint sp = make_surface_point(sc, ps.ray, hit);
    if (first) aov.normal = sp.normal;

    if (sp.light >= 0) {
        const qz_light l = sc.lights[sp.light];
        Spec4 emitted = light_emission(sc, l, sp.normal, -ps.ray.d, ps.lambda, ps.smp.index);
        if (!is_zero(emitted)) {
            has_gain = true;
            if (first || (ps.flags & QZ_FLAG_SPECULAR_BOUNCE)) {
                gain = ps.weight * emitted;
            } else {
                // light_sample_pmf * light->pdf: uniform light pick, area-measure pdf (scene.cpp:132-134, light.cpp:44-46)
                float light_proba = (1.0f / (float)sc.n_lights) * l.inv_area;
                float light_weight = power_heuristic(ps.p_b, light_proba);
                gain = emitted * (ps.weight * light_weight);
            }
        }
    }
    if (ps.depth == max_bounces) return false;

    // the material-select sample is only read by MixedMaterial (material.hpp:94-97)
    const uint32_t mat_dim = sample_1d_skip(ps.smp);
    if (sp.material < 0) {
        // emitter geometry has no material: pass straight through, depth unchanged (render.cpp:142-148)
        ps.flags |= QZ_FLAG_SPECULAR_BOUNCE;
        ps.ray.o = sp.point;
        return !is_zero(ps.weight);
    }
    float mat_sample = 0.0f;
    if (KH == KH_ANY && sc.materials[sp.material].kind == QZ_MAT_MIXED) mat_sample = src.one(R_MAT, mat_dim);
    const int32_t mat = KH == KH_ANY ? resolve_material(sc, sp.material, mat_sample) : sp.material;
    Bsdf f = make_bsdf<KH>(sc, mat, sp, ps.lambda, ps.pdf, ps.smp.index);

    // the wavefront's conductor kernel leaves the estimate to k_albedo_conductor (wf_shade.cuh)
    if (first && !(KH == KH_CONDUCTOR && ALBEDO_STAGE)) aov.albedo = bsdf_rho_hd<KH>(sc, f, sp.wo);

    if (!bsdf_is_specular<KH>(f)) {
        bool has_shadow;
        V3 p_light;
        Spec4 direct;
        sample_lights<KH>(sc, sp, f, ps.lambda, ps.smp, src, has_shadow, p_light, direct);
        if (has_shadow) {
            shadow.o = sp.point;
            shadow.d = p_light - sp.point;
            shadow.contrib = ps.weight * direct;
            ps.flags |= QZ_FLAG_HAS_SHADOW;
        }
    }

    // g++ evaluates the arguments of bsdf->sample(wo, sample_1d(), sample_2d()) right to left:
    // the 2-D sample takes the next two dimensions, the 1-D sample the one after (render.cpp:177).
    // Only diffuse and rough-conductor BxDFs read the 2-D sample, only the dielectrics the 1-D one.
    const bool needs_u2 = is_kind<KH>(f, BX_DIFFUSE) || (is_kind<KH>(f, BX_CONDUCTOR) && !tr_is_smooth(f.rough));
    const bool needs_u1 = is_kind<KH>(f, BX_DIELECTRIC) || is_kind<KH>(f, BX_THIN);
    V2 u2 = v2(0.0f, 0.0f);
    const uint32_t db = sample_2d_skip(ps.smp);
    if (needs_u2) u2 = src.two(R_BSDF, db);
    float u1 = 0.0f;
    const uint32_t du = sample_1d_skip(ps.smp);
    if (needs_u1) u1 = src.one(R_U1, du);
    BsdfSample bs = bsdf_sample<KH>(f, sp.wo, u1, u2);
    if (!bs.valid) return false;

    ps.weight = ps.weight * r_div(bs.spec * fabsf(dot(bs.wi, sp.normal)), bs.pdf);
    ps.p_b = bs.pdf;  // pdf_is_proportional is never set by any BxDF
    if (bs.specular) ps.flags |= QZ_FLAG_SPECULAR_BOUNCE; else ps.flags &= ~QZ_FLAG_SPECULAR_BOUNCE;
    if (bs.transmission) ps.ior_scale *= bs.ior;
    ps.ray.o = sp.point;
    ps.ray.d = bs.wi;
    ps.depth++;

    // the roulette sample is drawn every bounce but only read when roulette applies (render.cpp:200-208)
    const uint32_t rr_dim = sample_1d_skip(ps.smp);
    Spec4 rr = ps.weight * ps.ior_scale;
    if (max_component(rr) < 1.f && ps.depth > 1) {
        float q = std_max(0.0f, 1.0f - max_component(rr));
        float roulette = src.one(R_RR, rr_dim);
        if (roulette < q) return false;
        ps.weight = r_div(ps.weight, 1.0f - q);
    }
    return !is_zero(ps.weight);
}

// render_pixels() prologue for one pixel-sample (render.cpp:268-273); y is the flipped image y
QZ_HD void start_path(const DScene& sc, const DCamera& cam, const SamplerParams& spar, uint32_t x, uint32_t y, uint32_t s,
                      PathState& ps, PathAov& aov) {
    ps.smp = sampler_start(spar, x, y, s);
    V2 jitter = sampler_pixel_jitter_memo(spar, sc.memo, ps.smp);
    float u = (float)x + jitter.x;
    float v = (float)y + jitter.y;
    ps.ray.o = cam.pos;
    ps.ray.d = cam.bottom_left + cam.du * u + cam.dv * v - cam.pos;
    float ul = sample_dimension_memo(sc.sampler_table, sc.memo, ps.smp.index, sample_1d_skip(ps.smp));
    sample_wavelengths(ul, ps.lambda, ps.pdf);
    ps.weight = spec4(1.0f);
    ps.L = spec4(0.0f);
    ps.depth = 0;
    ps.flags = 0;
    ps.p_b = 1.0f;
    ps.ior_scale = 1.0f;
    ps.n_rays = 0;
    aov.normal = v3(0.0f, 0.0f, 0.0f);
    aov.albedo = spec4(0.0f);
}

// A whole path in one thread: the per-path replay kernel (qz_trace_paths) and the host
// emulation use this; the wavefront pipeline runs the same three calls as separate kernels.
template <bool COUNT>
QZ_HD void run_path(const DScene& sc, const DCamera& cam, const SamplerParams& spar, uint32_t x, uint32_t y, uint32_t s,
                    uint32_t max_bounces, PathState& ps, PathAov& aov, Spec4& lambda0, TraversalCounters* cnt) {
    start_path(sc, cam, spar, x, y, s, ps, aov);
    lambda0 = ps.lambda;
    for (;;) {
        Hit hit;
        closest_hit<COUNT>(sc, ps.ray, hit, cnt);
        ps.n_rays++;
        ShadowRequest sh;
        SamplesOnTheFly src;
        src.tab = sc.sampler_table; src.index = ps.smp.index; src.memo = sc.memo;
        Spec4 gain;
        bool has_gain;
        bool alive = shade_bounce<KH_ANY, -1, false>(sc, ps, aov, hit, max_bounces, sh, src, gain, has_gain);
        if (has_gain) ps.L = ps.L + gain;
        if (ps.flags & QZ_FLAG_HAS_SHADOW) {
            Ray sr;
            sr.o = sh.o; sr.d = sh.d;
            ps.n_rays++;
            if (!occluded<COUNT>(sc, sr, cnt)) ps.L = ps.L + sh.contrib;
        }
        if (!alive) break;
    }
}

// per-path replay record (include/qz_b200.h: qz_trace_paths)
QZ_HD void write_trace_record(float* rec, const PathState& ps, const PathAov& aov, const Spec4& lambda0, V3 rgb, V3 argb) {
    for (int k = 0; k < 32; k++) rec[k] = 0.0f;
    for (int k = 0; k < 4; k++) {
        rec[k] = lambda0.v[k];
        rec[4 + k] = ps.pdf.v[k];
        rec[8 + k] = ps.L.v[k];
        rec[16 + k] = aov.albedo.v[k];
    }
    rec[12] = aov.normal.x; rec[13] = aov.normal.y; rec[14] = aov.normal.z;
    rec[15] = (float)ps.n_rays;
    rec[20] = rgb.x; rec[21] = rgb.y; rec[22] = rgb.z;
    rec[23] = argb.x; rec[24] = argb.y; rec[25] = argb.z;
}

}  // namespace qz
