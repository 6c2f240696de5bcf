// BxDF models and the warps that feed them: restatement of the reference's bxdf.cpp and the
// static warp functions of sampler.cpp, expression by expression (operand order, the places
// where the reference computes in double, and its known quirks are all kept -- SURVEY.md
// section 8.a-6, -11, -12):
//   sampler.cpp:10-25 concentric disk (double M_PI_4 / M_PI_2), :31-35 polar disk (uses
//   M_2_PI = 2/pi where 2*pi was meant), :56-62 uniform sphere, :68-74 cosine hemisphere;
//   bxdf.cpp:116-141 DiffuseBxDF; :144-169 reflect / refract; :171-186 real Fresnel;
//   :188-208 complex Fresnel (std::complex<float>: libgcc __divsc3 divides in double,
//   glibc csqrtf with a double hypot, norm = re^2 + im^2 -- pinned bit-exactly against the
//   reference object); :210-272 TrowbridgeReitzDistribution (std::pow(x, 2)
//   => double squares); :275-338 ConductorBxDF; :341-385 DielectricBxDF; :388-423
//   ThinDielectricBxDF; bxdf.hpp:46-56 rho_hd with the 16 fixed samples of render.cpp:153-167.
#pragma once

#include "math.cuh"
#include "scene.cuh"

namespace qz {

// ------------------------------------------------------------------ warps
QZ_HD V2 sample_uniform_disk(V2 uv) {
    V2 off = v2(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f);
    if (off.x == 0.0f && off.y == 0.0f) return v2(0.0f, 0.0f);
    float theta, r;
    if (fabsf(off.x) > fabsf(off.y)) {
        r = off.x;
        theta = (float)(QZ_PI_4 * (double)(off.y / off.x));
    } else {
        r = off.y;
        theta = (float)(QZ_PI_2 - QZ_PI_4 * (double)(off.x / off.y));
    }
    float sn, cs;
    qz_sincosf(theta, sn, cs);
    return v2(r * cs, r * sn);
}

QZ_HD V2 sample_uniform_disk_polar(V2 uv) {
    float r = sqrtf(uv.x);
    float theta = (float)(QZ_2_PI * (double)uv.y);
    float sn, cs;
    qz_sincosf(theta, sn, cs);
    return v2(r * cs, r * sn);
}

QZ_HD V3 sample_uniform_sphere(V2 uv) {
    float z = 1.0f - 2.0f * uv.x;
    float r = sqrtf(std_max(0.0f, 1.0f - z * z));
    float phi = (float)(2.0 * QZ_PI * (double)uv.y);
    float sn, cs;
    qz_sincosf(phi, sn, cs);
    return v3(r * cs, r * sn, z);
}

QZ_HD V3 sample_cosine_hemisphere(V2 uv) {
    V2 d = sample_uniform_disk(uv);
    float z = sqrtf(std_max(0.0f, 1.0f - d.x * d.x - d.y * d.y));
    return v3(d.x, d.y, z);
}

// sampler.hpp:55-57: cos_theta * M_1_PI evaluated in double
QZ_HD float cosine_hemisphere_pdf(float cos_theta) {
#if QZ_FAST
    return cos_theta * (float)QZ_1_PI;
#else
    return (float)((double)cos_theta * QZ_1_PI);
#endif
}

// ------------------------------------------------------------------ local-frame helpers (bxdf.cpp:8-45)
QZ_HD float cos2_theta(V3 w) { return w.z * w.z; }
QZ_HD float sin2_theta(V3 w) { return std_max(0.0f, 1.0f - cos2_theta(w)); }
QZ_HD float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
QZ_HD float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
QZ_HD float cos_phi(V3 w) {
    float s = sin_theta(w);
    return (s == 0.0f) ? 1.0f : std_clamp(w.x / s, -1.0f, 1.0f);
}
QZ_HD float sin_phi(V3 w) {
    float s = sin_theta(w);
    return (s == 0.0f) ? 0.0f : std_clamp(w.y / s, -1.0f, 1.0f);
}

// ------------------------------------------------------------------ Fresnel
QZ_HD float fresnel_dielectric(float cos_theta_i, float ior) {
    cos_theta_i = std_clamp(cos_theta_i, -1.0f, 1.0f);
    if (cos_theta_i < 0.0f) {
        ior = 1.0f / ior;
        cos_theta_i = -cos_theta_i;
    }
    float sin2_theta_i = 1.0f - cos_theta_i * cos_theta_i;
    float sin2_theta_t = sin2_theta_i / (ior * ior);
    if (sin2_theta_t >= 1.0f) return 1.0f;
    float cos_theta_t = sqrtf(1.0f - sin2_theta_t);
    float r_parallel = (ior * cos_theta_i - cos_theta_t) / (ior * cos_theta_i + cos_theta_t);
    float r_perp = (cos_theta_i - ior * cos_theta_t) / (cos_theta_i + ior * cos_theta_t);
    return 0.5f * (r_parallel * r_parallel + r_perp * r_perp);
}

struct Cx {
    float re, im;
};
QZ_HD Cx cx(float re, float im) { Cx c; c.re = re; c.im = im; return c; }
// libgcc __mulsc3 (finite operands): (ac - bd, ad + bc) in float
QZ_HD Cx cx_mul(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
// libgcc __divsc3: float operands are divided through double, no Smith scaling
QZ_HD Cx cx_div(Cx a, Cx b) {
    double aa = a.re, bb = a.im, cc = b.re, dd = b.im;
    double denom = cc * cc + dd * dd;
    return cx((float)((aa * cc + bb * dd) / denom), (float)((bb * cc - aa * dd) / denom));
}
// glibc hypotf: sqrt of the double sum of squares, rounded once
QZ_HD float hypot_f(float x, float y) { return (float)sqrt((double)x * (double)x + (double)y * (double)y); }
// glibc csqrtf (s_csqrt_template.c), finite operands in normal range
QZ_HD Cx cx_sqrt(Cx z) {
    if (z.im == 0.0f) {
        if (z.re < 0.0f) return cx(0.0f, copysignf(sqrtf(-z.re), z.im));
        return cx(fabsf(sqrtf(z.re)), copysignf(0.0f, z.im));
    }
    if (z.re == 0.0f) {
        float r = sqrtf(0.5f * fabsf(z.im));
        return cx(r, copysignf(r, z.im));
    }
    float d = hypot_f(z.re, z.im), r, s;
    if (z.re > 0.0f) {
        r = sqrtf(0.5f * (d + z.re));
        s = 0.5f * (z.im / r);
    } else {
        s = sqrtf(0.5f * (d - z.re));
        r = fabsf(0.5f * (z.im / s));
    }
    return cx(r, copysignf(s, z.im));
}

// bxdf.cpp:188-200 for cos_theta_i >= 0 (every call site passes an absolute value; the
// negative-cosine branch is kept for completeness)
QZ_HD float fresnel_complex(float cos_theta_i, Cx ior) {
    cos_theta_i = std_clamp(cos_theta_i, -1.0f, 1.0f);
    if (cos_theta_i < 0.0f) {
        ior = cx_div(cx(1.0f, 0.0f), ior);
        cos_theta_i = -cos_theta_i;
    }
    float sin2_theta_i = 1.0f - cos_theta_i * cos_theta_i;
    Cx sin2_theta_t = cx_div(cx(sin2_theta_i, 0.0f), cx_mul(ior, ior));
    Cx cos_theta_t = cx_sqrt(cx(1.0f - sin2_theta_t.re, -sin2_theta_t.im));
    Cx ic = cx(ior.re * cos_theta_i, ior.im * cos_theta_i);
    Cx r_parallel = cx_div(cx(ic.re - cos_theta_t.re, ic.im - cos_theta_t.im), cx(ic.re + cos_theta_t.re, ic.im + cos_theta_t.im));
    Cx it = cx_mul(ior, cos_theta_t);
    Cx r_perp = cx_div(cx(cos_theta_i - it.re, -it.im), cx(cos_theta_i + it.re, it.im));
    float n_par = r_parallel.re * r_parallel.re + r_parallel.im * r_parallel.im;
    float n_perp = r_perp.re * r_perp.re + r_perp.im * r_perp.im;
    return 0.5f * (n_par + n_perp);
}

#if QZ_FAST
// Radiometric build: the same formula in float complex arithmetic (cos_theta_i >= 0 at every call site).
QZ_HD Cx fcx_div(Cx a, Cx b) {
    const float inv = r_rcp(r_fma(b.re, b.re, b.im * b.im));
    return cx(r_fma(a.re, b.re, a.im * b.im) * inv, r_fma(a.im, b.re, -(a.re * b.im)) * inv);
}
QZ_HD Cx fcx_sqrt(Cx z) {
    // principal square root; (re, im) = (r, im / 2r) with r = sqrt((|z| + re) / 2) for re >= 0, mirrored otherwise
    const float d = r_sqrt(r_fma(z.re, z.re, z.im * z.im));
    if (d == 0.0f) return cx(0.0f, 0.0f);
    if (z.re >= 0.0f) {
        const float r = r_sqrt(0.5f * (d + z.re));
        return cx(r, 0.5f * r_div(z.im, r));
    }
    const float sI = r_sqrt(0.5f * (d - z.re));
    return cx(fabsf(0.5f * r_div(z.im, sI)), copysignf(sI, z.im));
}
QZ_HD float fresnel_complex_fast(float cos_theta_i, Cx ior) {
    cos_theta_i = std_clamp(cos_theta_i, 0.0f, 1.0f);
    const float sin2_theta_i = r_fma(-cos_theta_i, cos_theta_i, 1.0f);
    const Cx ior2 = cx(r_fma(ior.re, ior.re, -(ior.im * ior.im)), 2.0f * ior.re * ior.im);
    const Cx sin2_theta_t = fcx_div(cx(sin2_theta_i, 0.0f), ior2);
    const Cx cos_theta_t = fcx_sqrt(cx(1.0f - sin2_theta_t.re, -sin2_theta_t.im));
    const Cx ic = cx(ior.re * cos_theta_i, ior.im * cos_theta_i);
    const Cx r_parallel = fcx_div(cx(ic.re - cos_theta_t.re, ic.im - cos_theta_t.im), cx(ic.re + cos_theta_t.re, ic.im + cos_theta_t.im));
    const Cx it = cx(r_fma(ior.re, cos_theta_t.re, -(ior.im * cos_theta_t.im)), r_fma(ior.re, cos_theta_t.im, ior.im * cos_theta_t.re));
    const Cx r_perp = fcx_div(cx(cos_theta_i - it.re, -it.im), cx(cos_theta_i + it.re, it.im));
    const float n_par = r_fma(r_parallel.re, r_parallel.re, r_parallel.im * r_parallel.im);
    const float n_perp = r_fma(r_perp.re, r_perp.re, r_perp.im * r_perp.im);
    return 0.5f * (n_par + n_perp);
}
#endif

QZ_HD Spec4 fresnel_conductor(float cos_theta_i, const Spec4& eta, const Spec4& k) {
#if QZ_FAST
    return spec4(fresnel_complex_fast(cos_theta_i, cx(eta.v[0], k.v[0])), fresnel_complex_fast(cos_theta_i, cx(eta.v[1], k.v[1])),
                 fresnel_complex_fast(cos_theta_i, cx(eta.v[2], k.v[2])), fresnel_complex_fast(cos_theta_i, cx(eta.v[3], k.v[3])));
#endif
    return spec4(fresnel_complex(cos_theta_i, cx(eta.v[0], k.v[0])), fresnel_complex(cos_theta_i, cx(eta.v[1], k.v[1])),
                 fresnel_complex(cos_theta_i, cx(eta.v[2], k.v[2])), fresnel_complex(cos_theta_i, cx(eta.v[3], k.v[3])));
}

// ------------------------------------------------------------------ Trowbridge-Reitz (GGX)
struct TRDist {
    float ax, ay;
};
QZ_HD bool tr_is_smooth(const TRDist& d) { return d.ax < 1e-3f && d.ay < 1e-3f; }

#if QZ_FAST
// radiometric build of D, Lambda, G1, G, D_visible: float throughout, one reciprocal square root for cos/sin phi
QZ_HD void tr_phi(V3 w, float& cphi, float& sphi, float& sin2) {
    sin2 = std_max(0.0f, r_fma(-w.z, w.z, 1.0f));
    if (sin2 == 0.0f) { cphi = 1.0f; sphi = 0.0f; return; }
    const float inv = r_rsqrt(sin2);
    cphi = std_clamp(w.x * inv, -1.0f, 1.0f);
    sphi = std_clamp(w.y * inv, -1.0f, 1.0f);
}
QZ_HD float tr_D(const TRDist& d, V3 wm) {
    float cphi, sphi, sin2;
    tr_phi(wm, cphi, sphi, sin2);
    const float c2 = wm.z * wm.z;
    const float tan2 = r_div(sin2, c2);
    if (is_inf(tan2) || tan2 != tan2) return 0.0f;
    const float a = r_div(cphi, d.ax), b = r_div(sphi, d.ay);
    const float e1 = r_fma(tan2, r_fma(a, a, b * b), 1.0f);
    return r_rcp((float)QZ_PI * d.ax * d.ay * (c2 * c2) * (e1 * e1));
}
QZ_HD float tr_lambda(const TRDist& d, V3 w) {
    float cphi, sphi, sin2;
    tr_phi(w, cphi, sphi, sin2);
    const float tan2 = r_div(sin2, w.z * w.z);
    if (is_inf(tan2) || tan2 != tan2) return 0.0f;
    const float a = cphi * d.ax, b = sphi * d.ay;
    return 0.5f * (r_sqrt(r_fma(r_fma(a, a, b * b), tan2, 1.0f)) - 1.0f);
}
QZ_HD float tr_G1(const TRDist& d, V3 w) { return r_rcp(1.0f + tr_lambda(d, w)); }
QZ_HD float tr_G(const TRDist& d, V3 wo, V3 wi) { return r_rcp(1.0f + tr_lambda(d, wo) + tr_lambda(d, wi)); }
QZ_HD float tr_D_visible(const TRDist& d, V3 w, V3 wm) { return r_div(tr_G1(d, w), fabsf(w.z)) * tr_D(d, wm) * fabsf(dot(w, wm)); }
#else
QZ_HD float tr_D(const TRDist& d, V3 wm) {
    float tan2 = tan2_theta(wm);
    if (is_inf(tan2)) return 0.0f;
    double c2 = (double)cos2_theta(wm);
    float cos4 = (float)(c2 * c2);
    double a = (double)(cos_phi(wm) / d.ax), b = (double)(sin_phi(wm) / d.ay);
    float e = (float)((double)tan2 * (a * a + b * b));
    double den = QZ_PI * (double)d.ax * (double)d.ay * (double)cos4 * (double)(1.0f + e) * (double)(1.0f + e);
    return (float)(1.0 / den);
}

QZ_HD float tr_lambda(const TRDist& d, V3 w) {
    float tan2 = tan2_theta(w);
    if (is_inf(tan2)) return 0.0f;
    double a = (double)(cos_phi(w) * d.ax), b = (double)(sin_phi(w) * d.ay);
    float alpha2 = (float)(a * a + b * b);
    return (sqrtf(1.0f + alpha2 * tan2) - 1.0f) / 2.0f;
}
QZ_HD float tr_G1(const TRDist& d, V3 w) { return 1.0f / (1.0f + tr_lambda(d, w)); }
QZ_HD float tr_G(const TRDist& d, V3 wo, V3 wi) { return 1.0f / (1.0f + tr_lambda(d, wo) + tr_lambda(d, wi)); }
// distribution of visible normals (bxdf.cpp:230-232)
QZ_HD float tr_D_visible(const TRDist& d, V3 w, V3 wm) { return (tr_G1(d, w) / fabsf(w.z)) * tr_D(d, wm) * fabsf(dot(w, wm)); }
#endif

// visible-normal sampling from an already warped polar-disk point p (bxdf.cpp:255-272)
QZ_HD V3 tr_sample_from_disk(const TRDist& d, V3 w, V2 p) {
    V3 wh = normalized(v3(d.ax * w.x, d.ay * w.y, w.z));
    if (wh.z < 0.0f) wh = -wh;
    V3 t1 = wh.z < 0.99999f ? normalized(cross(v3(0.0f, 0.0f, 1.0f), wh)) : v3(1.0f, 0.0f, 0.0f);
    V3 t2 = cross(wh, t1);
    float h = sqrtf(1.0f - p.x * p.x);
    p.y = lerpf(h, p.y, (float)((1.0 + (double)wh.z) * 0.5));
    float pz = sqrtf(std_max(0.0f, 1.0f - (p.x * p.x + p.y * p.y)));
    V3 nh = t1 * p.x + t2 * p.y + wh * pz;
    return normalized(v3(d.ax * nh.x, d.ay * nh.y, std_max(1e-6f, nh.z)));
}
QZ_HD V3 tr_sample(const TRDist& d, V3 w, V2 u) { return tr_sample_from_disk(d, w, sample_uniform_disk_polar(u)); }

// ------------------------------------------------------------------ BSDF record
enum BxdfKind { BX_DIFFUSE = 0, BX_CONDUCTOR = 1, BX_DIELECTRIC = 2, BX_THIN = 3 };

struct Bsdf {
    int kind;
    V3 u0, u1, u2;  // OrthonormalBasis (onb.hpp:12-28); u2 = shading normal
    Spec4 a;        // diffuse: reflectance; conductor: eta
    Spec4 b;        // conductor: k
    TRDist rough;
    float ior;      // dielectric / thin
};

QZ_HD void make_basis(V3 n, V3& u0, V3& u1, V3& u2) {
    n = normalized(n);
    float sign = n.z > 0.0f ? 1.0f : -1.0f;
    float a = -1.0f / (sign + n.z);
    float b = n.x * n.y * a;
    u0 = v3(1.0f + sign * n.x * n.x * a, sign * b, -sign * n.x);
    u1 = v3(b, sign + n.y * n.y * a, -n.y);
    u2 = n;
}
QZ_HD V3 to_local(const Bsdf& f, V3 v) { return v3(dot(f.u0, v), dot(f.u1, v), dot(f.u2, v)); }
QZ_HD V3 from_local(const Bsdf& f, V3 v) { return f.u0 * v.x + f.u1 * v.y + f.u2 * v.z; }

// Compile-time kind hint of the branch-sorted shading kernels: a kernel that only ever sees
// one BxDF family folds the other branches away.  KH_ANY keeps the run-time dispatch.
enum KindHint { KH_ANY = -1, KH_DIFFUSE = 0, KH_CONDUCTOR = 1, KH_DIELECTRIC = 2 /* dielectric or thin */ };

template <int KH>
QZ_HD bool is_kind(const Bsdf& f, int k) {
    if (KH == KH_ANY) return f.kind == k;
    if (KH == KH_DIELECTRIC) return (k == BX_DIELECTRIC || k == BX_THIN) && f.kind == k;
    return KH == k;
}

template <int KH>
QZ_HD bool bsdf_is_specular(const Bsdf& f) {
    if (is_kind<KH>(f, BX_DIFFUSE)) return false;
    if (is_kind<KH>(f, BX_CONDUCTOR)) return tr_is_smooth(f.rough);
    return true;
}

struct BsdfSample {
    Spec4 spec;
    V3 wi;
    float pdf;
    float ior;
    bool specular, transmission;
    bool valid;
};

QZ_HD V3 reflect(V3 wo, V3 n) { return -wo + n * (2.0f * dot(wo, n)); }

// local-frame evaluation f(wo, wi)
template <int KH>
QZ_HD Spec4 bxdf_f(const Bsdf& f, V3 wo, V3 wi) {
    if (is_kind<KH>(f, BX_DIFFUSE)) {
        if (wo.z * wi.z <= 0.0f) return spec4(0.0f);
        return f.a * (float)QZ_1_PI;
    }
    if (is_kind<KH>(f, BX_CONDUCTOR)) {
        if (wo.z * wi.z <= 0.0f || tr_is_smooth(f.rough)) return spec4(0.0f);
        float cos_o = fabsf(wo.z), cos_i = fabsf(wi.z);
        if (cos_i == 0.0f || cos_o == 0.0f) return spec4(0.0f);
        V3 wm = wi + wo;
        if (norm_squared(wm) == 0.0f) return spec4(0.0f);
        wm = r_normalized(wm);
        Spec4 F = fresnel_conductor(fabsf(dot(wo, wm)), f.a, f.b);
        return F * r_div(tr_D(f.rough, wm) * tr_G(f.rough, wo, wi), 4.0f * cos_i * cos_o);
    }
    return spec4(0.0f);
}

template <int KH>
QZ_HD float bxdf_pdf(const Bsdf& f, V3 wo, V3 wi) {
    if (is_kind<KH>(f, BX_DIFFUSE)) {
        if (wo.z * wi.z <= 0.0f) return 0.0f;
        return cosine_hemisphere_pdf(fabsf(wi.z));
    }
    if (is_kind<KH>(f, BX_CONDUCTOR)) {
        if (wo.z * wi.z <= 0.0f || tr_is_smooth(f.rough)) return 0.0f;
        V3 wm = wo + wi;
        if (norm_squared(wm) == 0.0f) return 0.0f;
        wm = dot(wm, v3(0.0f, 0.0f, 1.0f)) > 0.0f ? r_normalized(wm) : -r_normalized(wm);
        return r_div(tr_D_visible(f.rough, wo, wm), 4.0f * fabsf(dot(wo, wm)));
    }
    return 0.0f;
}

// BxDF::sample in the local frame.  `disk` carries a precomputed warp of u2 (the depth-0
// albedo estimate re-uses 16 constant points): for DIFFUSE the cosine-hemisphere direction,
// for CONDUCTOR the polar-disk point; with use_disk == false the warp is evaluated from u2.
template <int KH>
QZ_HD BsdfSample bxdf_sample(const Bsdf& f, V3 wo, float u1, V2 u2, bool use_disk, V3 disk) {
    BsdfSample s;
    s.valid = false; s.pdf = 0.0f; s.ior = 1.0f; s.specular = false; s.transmission = false;
    s.spec = spec4(0.0f); s.wi = v3(0.0f, 0.0f, 0.0f);
    if (is_kind<KH>(f, BX_DIFFUSE)) {
        V3 wi = use_disk ? disk : sample_cosine_hemisphere(u2);
        if (wo.z < 0.0f) wi.z *= -1.0f;
        s.spec = f.a * (float)QZ_1_PI;
        s.wi = wi;
        s.pdf = cosine_hemisphere_pdf(fabsf(wi.z));
        s.valid = true;
        return s;
    }
    if (is_kind<KH>(f, BX_CONDUCTOR)) {
        if (tr_is_smooth(f.rough)) {
            V3 wi = v3(-wo.x, -wo.y, wo.z);
            s.spec = r_div(fresnel_conductor(fabsf(wi.z), f.a, f.b), fabsf(wi.z));
            s.wi = wi; s.pdf = 1.0f; s.specular = true; s.valid = true;
            return s;
        }
        V3 wm = use_disk ? tr_sample_from_disk(f.rough, wo, v2(disk.x, disk.y)) : tr_sample(f.rough, wo, u2);
        V3 wi = reflect(wo, wm);
        if (wo.z * wi.z <= 0.0f) return s;
        s.pdf = r_div(tr_D_visible(f.rough, wo, wm), 4.0f * fabsf(dot(wo, wm)));
        float cos_o = fabsf(wo.z), cos_i = fabsf(wi.z);
        Spec4 F = fresnel_conductor(fabsf(dot(wo, wm)), f.a, f.b);
        s.spec = F * r_div(tr_D(f.rough, wm) * tr_G(f.rough, wo, wi), 4.0f * cos_i * cos_o);
        s.wi = wi; s.valid = true;
        return s;
    }
    // dielectric interfaces: reflectance at the interface, then the thin-slab series for BX_THIN
    float r = fresnel_dielectric(wo.z, f.ior);
    float t = 1.0f - r;
    if (is_kind<KH>(f, BX_THIN) && r < 1.0f) {
        r += t * t * r / (1.0f - r * r);
        t = 1.0f - r;
    }
    if (u1 < r / (r + t)) {
        V3 wi = v3(-wo.x, -wo.y, wo.z);
        s.spec = spec4(r / fabsf(wi.z));
        s.wi = wi; s.pdf = r / (r + t);
        s.specular = is_kind<KH>(f, BX_DIELECTRIC);  // ThinDielectricBxDF leaves the flags unset (bxdf.cpp:401-417)
        s.valid = true;
        return s;
    }
    if (is_kind<KH>(f, BX_THIN)) {
        V3 wi = -wo;
        s.spec = spec4(t / fabsf(wi.z));
        s.wi = wi; s.pdf = t / (r + t); s.valid = true;
        return s;
    }
    // refract(wo, (0,0,1), ior) (bxdf.cpp:148-169)
    V3 n = v3(0.0f, 0.0f, 1.0f);
    float ior = f.ior;
    float cos_i = dot(wo, n);
    if (cos_i < 0.0f) {
        ior = 1.0f / ior;
        cos_i = -cos_i;
        n = -n;
    }
    float sin2_i = std_max(0.0f, 1.0f - cos_i * cos_i);
    float sin2_t = sin2_i / (ior * ior);
    if (sin2_t >= 1.0f) return s;  // total internal reflection: no sample
    float cos_t = sqrtf(1.0f - sin2_t);
    V3 wt = (-wo) / ior + n * (cos_i / ior - cos_t);
    s.spec = spec4(t / fabsf(wt.z));
    s.wi = wt; s.pdf = t / (r + t); s.ior = ior;
    s.specular = true; s.transmission = true; s.valid = true;
    return s;
}

// BSDF::operator() / pdf / sample in render space (bxdf.cpp:76-108)
template <int KH>
QZ_HD Spec4 bsdf_f(const Bsdf& f, V3 wo_r, V3 wi_r) {
    V3 wi = to_local(f, wi_r);
    V3 wo = to_local(f, wo_r);
    if (wo.z == 0.0f) return spec4(0.0f);
    return bxdf_f<KH>(f, wo, wi);
}
template <int KH>
QZ_HD float bsdf_pdf(const Bsdf& f, V3 wo_r, V3 wi_r) {
    V3 wi = to_local(f, wi_r);
    V3 wo = to_local(f, wo_r);
    if (wo.z == 0.0f) return 0.0f;
    return bxdf_pdf<KH>(f, wo, wi);
}
template <int KH>
QZ_HD BsdfSample bsdf_sample(const Bsdf& f, V3 wo_r, float u1, V2 u2) {
    V3 wo = to_local(f, wo_r);
    BsdfSample s;
    if (wo.z == 0.0f) {
        s.valid = false; s.pdf = 0.0f; s.ior = 1.0f; s.specular = false; s.transmission = false;
        s.spec = spec4(0.0f); s.wi = v3(0.0f, 0.0f, 0.0f);
        return s;
    }
    s = bxdf_sample<KH>(f, wo, u1, u2, false, v3(0.0f, 0.0f, 0.0f));
    if (!s.valid || is_zero(s.spec) || s.pdf == 0.0f || s.wi.z == 0.0f) { s.valid = false; return s; }
    s.wi = from_local(f, s.wi);
    return s;
}

// hemispherical-directional reflectance with the 16 fixed sample points of render.cpp:153-167
// (bxdf.hpp:46-56).  rho_tab (host-built, 16 x 8 floats): uc, u2.x, u2.y, cosine-hemisphere
// warp of u2 (x, y, z), polar-disk warp of u2 (x, y) -- the warps of CONSTANT points are
// evaluated once on the host with the same libm the oracle uses.
template <int KH>
QZ_HD Spec4 bsdf_rho_hd(const DScene& sc, const Bsdf& f, V3 wo_r) {
    V3 wo = to_local(f, wo_r);
#if QZ_FAST
    // diffuse: every one of the 16 terms is reflectance * (1/pi * |z_i|) / (|z_i| * 1/pi) -- the sum over the
    // constant sample points is folded into one factor (rho_tab[128], build_rho_table)
    if (is_kind<KH>(f, BX_DIFFUSE)) return f.a * sc.rho_tab[128];
#endif
    Spec4 acc = spec4(0.0f);
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
        const float* t = sc.rho_tab + i * 8;
        V3 disk = is_kind<KH>(f, BX_DIFFUSE) ? v3(t[3], t[4], t[5]) : v3(t[6], t[7], 0.0f);
        BsdfSample s = bxdf_sample<KH>(f, wo, t[0], v2(t[1], t[2]), true, disk);
        if (s.valid) acc = acc + r_div(s.spec * fabsf(s.wi.z), s.pdf);
    }
    return acc / 16.0f;
}

#define QZ_RHO_TAB_FLOATS (16 * 8 + 1)
inline void build_rho_table(float* out /* QZ_RHO_TAB_FLOATS */) {
    // render.cpp:153-167 (double literals narrowed to float, as the std::array initialisers do)
    const float uc[16] = {0.75741637, 0.37870818, 0.7083487, 0.18935409, 0.9149363, 0.35417435, 0.5990858, 0.09467703,
                          0.8578725, 0.45746812, 0.686759, 0.17708716, 0.9674518, 0.2995429, 0.5083201, 0.047338516};
    const float u2[16][2] = {{0.855985, 0.570367}, {0.381823, 0.851844}, {0.285328, 0.764262}, {0.733380, 0.114073},
                             {0.542663, 0.344465}, {0.127274, 0.414848}, {0.964700, 0.947162}, {0.594089, 0.643463},
                             {0.095109, 0.170369}, {0.825444, 0.263359}, {0.429467, 0.454469}, {0.244460, 0.816459},
                             {0.756135, 0.731258}, {0.516165, 0.152852}, {0.180888, 0.214174}, {0.898579, 0.503897}};
    for (int i = 0; i < 16; i++) {
        float* t = out + i * 8;
        t[0] = uc[i]; t[1] = u2[i][0]; t[2] = u2[i][1];
        V3 c = sample_cosine_hemisphere(v2(u2[i][0], u2[i][1]));
        t[3] = c.x; t[4] = c.y; t[5] = c.z;
        V2 p = sample_uniform_disk_polar(v2(u2[i][0], u2[i][1]));
        t[6] = p.x; t[7] = p.y;
    }
    // [128]: mean over the 16 points of (1/pi * z) / pdf(z) with the reference's float / double roundings -- the
    // factor the radiometric build multiplies a diffuse reflectance with (bsdf_rho_hd)
    float acc = 0.0f;
    for (int i = 0; i < 16; i++) {
        const float z = fabsf(out[i * 8 + 5]);
        const float pdf = (float)((double)z * QZ_1_PI);
        if (pdf != 0.0f) acc += ((float)QZ_1_PI * z) / pdf;
    }
    out[128] = acc / 16.0f;
}

}  // namespace qz
