// Spectrum evaluation on the device: tagged-union restatement of the reference's virtual
// Spectrum::operator() family (color/spectrum.cpp:47-53 dense, :101-109 piecewise linear
// with its two quirks, :113-133 blackbody; color/rgb.cpp:51-67 sigmoid polynomial,
// :176-178 unbounded, :191-196 illuminant), the RGB -> spectrum table lookup used by
// ImageTexture (rgb.cpp:50-57, 78-140), wavelength sampling (spectrum_sample.cpp:11-42)
// and PixelSensor::to_sensor_rgb (sensor.cpp:57-70).
#pragma once

#include "math.cuh"
#include "scene.cuh"

namespace qz {

// R = the value is RADIOMETRIC (common.cuh, ARITHMETIC MODES): evaluated with the r_* primitives, which are
// the plain IEEE operations in the exact build.  R = false is the reference's arithmetic in both builds: the
// index of refraction of a dielectric steers the refracted DIRECTION and is always evaluated that way.
template <bool R> QZ_HD float t_div(float a, float b) { return R ? r_div(a, b) : a / b; }
template <bool R> QZ_HD float t_fma(float a, float b, float c) { return R ? r_fma(a, b, c) : a * b + c; }

template <bool R>
QZ_HD float sigmoidf(float x) {
    if (is_inf(x)) return x > 0.0f ? 1.0f : 0.0f;
    // (IEEE square root and division in the radiometric build as well: for a strongly negative x the result is the
    // small difference 0.5 - 0.4999..., so an ulp of the quotient is 1e-4 of a dark reflectance -- the reference's
    // value can only be matched by rounding the way it does)
    return 0.5f + 0.5f * x / (sqrtf(1.0f + x * x));
}

template <bool R>
QZ_HD float sigmoid_poly(float c0, float c1, float c2, float lambda) {
    // (the polynomial too: saturated colours have coefficients in the thousands that cancel to an argument of order one)
    return sigmoidf<R>(c0 + c1 * lambda + c2 * lambda * lambda);
}

// The three tables a spectrum evaluation reads.  The out-of-line evaluators below take this BY VALUE (six registers),
// not the scene by reference: a reference parameter of a real call forced every shading kernel to keep a copy of the
// whole DScene in local memory (a 200-byte frame per thread that, at 640 threads per SM, competed with the scene
// tables for L1) and the callee to reach the tables through it.
struct SpecView {
    const qz_spectrum* spectra;
    const float* pool;
    const uint8_t* pw_accel;
};
QZ_HD SpecView spec_view(const DScene& sc) { SpecView v; v.spectra = sc.spectra; v.pool = sc.pool; v.pw_accel = sc.pw_accel; return v; }

// lroundf(x) for |x| < 2^22 (wavelength offsets): round half away from zero.  x - trunc(x) is exact, so this is the
// same integer; CUDA's lroundf is a 64-bit software sequence (5 % of the sensor stage's instructions).
QZ_HD int lround_small(float x) {
    const float t = truncf(x);
    const float r = x - t;
    int i = (int)t;
    if (r >= 0.5f) i++;
    if (r <= -0.5f) i--;
    return i;
}

// every kind except RGB_ILLUMINANT (which multiplies by another spectrum)
template <bool R>
QZ_HD float eval_spectrum_leaf(const SpecView& sc, const qz_spectrum& s, float lambda) {
    {
        switch (s.kind) {
            case QZ_SPEC_CONSTANT:
                return s.a;
            case QZ_SPEC_DENSE: {
                const float off = lambda - (float)s.aux;
                if (!(off > -4194304.0f && off < 4194304.0f)) return 0.0f;   // (also NaN) far outside any table
                const int idx = lround_small(off);
                if (idx < 0 || idx >= (int)s.count) return 0.0f;
                return sc.pool[s.offset + idx];
            }
            case QZ_SPEC_PIECEWISE: {
                const float* l = sc.pool + s.offset;
                const float* v = l + s.count + 1;
                if (s.count == 0 || lambda < l[0] || lambda > l[s.count - 1]) return 0.0f;
                // std::lower_bound: first knot >= lambda
                uint32_t first = 0;
                if (s.aux >= 0 && lambda >= 360.0f && lambda < 831.0f) {
                    // bucket table: lower_bound for floor(lambda), then at most a knot or two forward
                    first = sc.pw_accel[(uint32_t)s.aux + (uint32_t)(int)(lambda - 360.0f)];
                    while (l[first] < lambda) first++;
                } else {
                    uint32_t len = s.count;
                    while (len > 0) {
                        uint32_t half = len >> 1;
                        if (l[first + half] < lambda) { first += half + 1; len -= half + 1; }
                        else len = half;
                    }
                }
                // l[count] and v[count] are the zero pads standing in for the reference's
                // one-past-the-end read
                const float l0 = l[first], l1 = l[first + 1], v0 = v[first], v1 = v[first + 1];
                float t = t_div<R>(lambda - l0, l1 - l0);
                if (R && QZ_FAST) return r_fma(t, v1 - v0, v0);
                return v0 * (1.0f - t) + v1 * t;
            }
            case QZ_SPEC_SIGMOID:
                return sigmoid_poly<R>(s.a, s.b, s.c, lambda);
            case QZ_SPEC_RGB_UNBOUNDED:
                return s.scale * sigmoid_poly<R>(s.a, s.b, s.c, lambda);
            case QZ_SPEC_BLACKBODY: {
                // spectrum.cpp:113-133; pow(lambda, 5) is a double pow, expm1f a float one
                if (s.a <= 0.f) return s.b * 0.f;
                const float c = 299792458.f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
                float l = lambda * 1e-9f;
                double den = pow((double)lambda, 5.0) * (double)expm1f(h * c / (l * kb * s.a));
                float le = (float)((double)(2.0f * h * c * c) / den);
                return s.b * le;
            }
            default:
                return 0.0f;
        }
    }
}

QZ_HD qz_spectrum load_spectrum(const SpecView& sc, int32_t id) {
#if defined(__CUDA_ARCH__)
    // one 32-byte record = two 16-byte loads through the read-only path
    const uint4* p = reinterpret_cast<const uint4*>(sc.spectra + id);
    const uint4 r0 = __ldg(p), r1 = __ldg(p + 1);
    qz_spectrum s;
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
    w[0] = r0.x; w[1] = r0.y; w[2] = r0.z; w[3] = r0.w; w[4] = r1.x; w[5] = r1.y; w[6] = r1.z; w[7] = r1.w;
    return s;
#else
    return sc.spectra[id];
#endif
}

QZ_HD qz_spectrum load_spectrum(const DScene& sc, int32_t id) { return load_spectrum(spec_view(sc), id); }

template <bool R>
QZ_HD float eval_spectrum_rec(const SpecView& sc, const qz_spectrum& s, float lambda) {
    if (s.kind == QZ_SPEC_RGB_ILLUMINANT) {
        if (s.aux < 0) return 0.0f;
        // m_scale * m_polynomial(lambda) * illuminant(lambda), left to right (rgb.cpp:191-196);
        // the illuminant is never itself an RGBIlluminantSpectrum (it is the colour space's D65)
        float head = s.scale * sigmoid_poly<R>(s.a, s.b, s.c, lambda);
        const qz_spectrum ill = load_spectrum(sc, s.aux);
        return head * eval_spectrum_leaf<R>(sc, ill, lambda);
    }
    return eval_spectrum_leaf<R>(sc, s, lambda);
}

template <bool R>
QZ_HD float eval_spectrum_rec(const DScene& sc, const qz_spectrum& s, float lambda) { return eval_spectrum_rec<R>(spec_view(sc), s, lambda); }

// Spectrum::operator() in the reference's arithmetic (probes, index of refraction)
QZ_HD_CALL float eval_spectrum_call(SpecView sc, int32_t id, float lambda) {
    return eval_spectrum_rec<false>(sc, load_spectrum(sc, id), lambda);
}
QZ_HD float eval_spectrum(const DScene& sc, int32_t id, float lambda) { return eval_spectrum_call(spec_view(sc), id, lambda); }

// SpectrumSample::from_spectrum (spectrum_sample.cpp:49-58), radiometric: the record (and an illuminant's
// second record) is fetched ONCE for the four wavelengths, and the four evaluations are one instruction
// stream the scheduler can interleave -- as four calls of eval_spectrum() they were four serial chains of
// dependent loads, 12 % of a shading kernel's instructions for the record fetch alone.
// (arguments and result of the real call travel in registers: scalars in, one 16-byte vector out)
struct SpecRet { float x, y, z, w; };
QZ_HD_CALL SpecRet from_spectrum_call(SpecView sc, int32_t id, float la, float lb, float lc, float ld);
QZ_HD Spec4 from_spectrum(const DScene& sc, int32_t id, const Spec4& lambda) {
    const SpecRet r = from_spectrum_call(spec_view(sc), id, lambda.v[0], lambda.v[1], lambda.v[2], lambda.v[3]);
    return spec4(r.x, r.y, r.z, r.w);
}
QZ_HD Spec4 from_spectrum_body(const SpecView& sc, int32_t id, const Spec4& lambda) {
    const qz_spectrum s = load_spectrum(sc, id);
    if (s.kind == QZ_SPEC_RGB_ILLUMINANT) {
        if (s.aux < 0) return spec4(0.0f);
        const qz_spectrum ill = load_spectrum(sc, s.aux);
        Spec4 r;
#pragma unroll
        for (int k = 0; k < 4; k++)
            r.v[k] = (s.scale * sigmoid_poly<true>(s.a, s.b, s.c, lambda.v[k])) * eval_spectrum_leaf<true>(sc, ill, lambda.v[k]);
        return r;
    }
    Spec4 r;
#pragma unroll
    for (int k = 0; k < 4; k++) r.v[k] = eval_spectrum_leaf<true>(sc, s, lambda.v[k]);
    return r;
}
QZ_HD_CALL SpecRet from_spectrum_call(SpecView sc, int32_t id, float la, float lb, float lc, float ld) {
    const Spec4 r = from_spectrum_body(sc, id, spec4(la, lb, lc, ld));
    SpecRet o; o.x = r.v[0]; o.y = r.v[1]; o.z = r.v[2]; o.w = r.v[3];
    return o;
}

// RGBColorSpace::to_spectrum + RGBToSpectrumTable::operator() (rgb.cpp:50-57, 78-140);
// returns (c0, c1, c2) of the sigmoid polynomial.  The reference's arithmetic in both builds: the
// coefficients feed a polynomial whose terms cancel (see sigmoid_poly), so a relative 1e-7 here is 1e-4 there.
struct LutView { const float* lut_z; const float* lut_coeffs; };
QZ_HD_CALL V3 rgb_to_sigmoid_call(LutView sc, float r, float g, float b);
QZ_HD V3 rgb_to_sigmoid(const DScene& sc, float r, float g, float b) {
    LutView v; v.lut_z = sc.lut_z; v.lut_coeffs = sc.lut_coeffs;
    return rgb_to_sigmoid_call(v, r, g, b);
}
QZ_HD_CALL V3 rgb_to_sigmoid_call(LutView sc, float r, float g, float b) {
    r = std_clamp(r, 0.0f, 1.0f); g = std_clamp(g, 0.0f, 1.0f); b = std_clamp(b, 0.0f, 1.0f);
    if (r == g && g == b) {
        return v3(0.0f, 0.0f, (r - 0.5f) / sqrtf(std_max(0.0f, r * (1.0f - r))));
    }
    const float comps[3] = {r, g, b};
    uint32_t maxc = (comps[0] > comps[1]) ? ((comps[0] > comps[2]) ? 0u : 2u) : ((comps[1] > comps[2]) ? 1u : 2u);
    float z = comps[maxc];
    float x = comps[(maxc + 1) % 3] * 31.0f / z;
    float y = comps[(maxc + 2) % 3] * 31.0f / z;
    uint32_t xi = (uint32_t)x; if (xi > 30u) xi = 30u;
    uint32_t yi = (uint32_t)y; if (yi > 30u) yi = 30u;
    // lower_bound over the 32 z nodes, then step back one
    uint32_t first = 0, len = 32;
    while (len > 0) {
        uint32_t half = len >> 1;
        if (sc.lut_z[first + half] < z) { first += half + 1; len -= half + 1; }
        else len = half;
    }
    uint32_t zi = first;
    if (zi != 0) zi--;
    if (zi > 30u) zi = 30u;
    float dx = x - (float)xi, dy = y - (float)yi;
    float dz = (z - sc.lut_z[zi]) / (sc.lut_z[zi + 1] - sc.lut_z[zi]);
    float c[3];
    const float* base = sc.lut_coeffs + (size_t)maxc * 32 * 32 * 32 * 3;
    for (uint32_t i = 0; i < 3; i++) {
#define QZ_CO(a, bb, d) base[((zi + (d)) * 32 * 32 + (yi + (bb)) * 32 + (xi + (a))) * 3 + i]
        c[i] = lerpf(lerpf(lerpf(QZ_CO(0, 0, 0), QZ_CO(1, 0, 0), dx), lerpf(QZ_CO(0, 1, 0), QZ_CO(1, 1, 0), dx), dy),
                      lerpf(lerpf(QZ_CO(0, 0, 1), QZ_CO(1, 0, 1), dx), lerpf(QZ_CO(0, 1, 1), QZ_CO(1, 1, 1), dx), dy), dz);
#undef QZ_CO
    }
    return v3(c[2], c[1], c[0]);
}

// WavelengthSample::uniform (spectrum_sample.cpp:11-24) over [360, 830]
QZ_HD void sample_wavelengths(float u, Spec4& lambda, Spec4& pdf) {
    const float lmin = 360.0f, lmax = 830.0f;
    lambda.v[0] = (1.0f - u) * lmin + u * lmax;
    const float delta = (lmax - lmin) / 4.0f;
    for (int i = 1; i < 4; i++) {
        lambda.v[i] = lambda.v[i - 1] + delta;
        if (lambda.v[i] > lmax) lambda.v[i] = lmin + (lambda.v[i] - lmax);
    }
    pdf = spec4(1.0f / (lmax - lmin));
}

// WavelengthSample::terminate_secondary (spectrum_sample.cpp:34-42)
QZ_HD void terminate_secondary(Spec4& pdf) {
    if (pdf.v[1] == 0.0f && pdf.v[2] == 0.0f && pdf.v[3] == 0.0f) return;
    pdf.v[1] = 0.0f; pdf.v[2] = 0.0f; pdf.v[3] = 0.0f;
    pdf.v[0] /= 4.0f;
}

// DenselySampledSpectrum lookup of one sensor curve (lambda_min 360, 471 samples)
QZ_HD float sensor_curve(const float* curve, float lambda) {
    const int idx = lround_small(lambda - 360.0f);
    if (idx < 0 || idx >= 471) return 0.0f;
    return curve[idx];
}

// PixelSensor::to_sensor_rgb (sensor.cpp:57-70) incl. the per-sample saturation clamp at 40
QZ_HD V3 to_sensor_rgb(const DCamera& cam, const Spec4& L, const Spec4& lambda, const Spec4& pdf) {
    Spec4 l = r_div(L, pdf);
    float rgb[3];
    // the four curve indices are the same for the three curves
    int idx[4];
    for (int j = 0; j < 4; j++) idx[j] = lround_small(lambda.v[j] - 360.0f);   // wavelengths lie in [360, 830]
    for (int k = 0; k < 3; k++) {
        const float* curve = cam.sensor + k * 471;
        Spec4 resp;
        for (int j = 0; j < 4; j++) resp.v[j] = (idx[j] < 0 || idx[j] >= 471) ? 0.0f : curve[idx[j]];
        rgb[k] = average(resp * l) * cam.imaging_ratio;
    }
    // std::max({x, y, z})
    float m = rgb[0];
    if (m < rgb[1]) m = rgb[1];
    if (m < rgb[2]) m = rgb[2];
    if (m > 40.0f) {
        float s = r_div(40.0f, m);
        rgb[0] *= s; rgb[1] *= s; rgb[2] *= s;
    }
    return v3(rgb[0], rgb[1], rgb[2]);
}

}  // namespace qz
