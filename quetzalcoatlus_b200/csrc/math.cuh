// Transcendental functions of the hot path.  These are the ONLY operations that cannot be
// bit-identical between the B200 and the oracle's glibc libm (sampler.cpp:22-23,34 sinf/cosf
// in the disk warps; :44-46,58-61 in the sphere warps; scene.cpp:28,33 atan2f/acosf for
// sphere uv).  glibc's float functions evaluate a double-precision kernel and round once;
// the device versions do the same with CUDA's double functions, so the two agree except
// when the exact value sits within ~1e-9 ulp... of a float rounding boundary of the
// respective polynomial error (measured mismatch rate in DESIGN.md).  In the test-only
// host emulation build the host libm is used, which isolates everything else for bit-exact
// checking on the CPU.
#pragma once

#include "common.cuh"

namespace qz {

QZ_HD_CALL float qz_sinf(float x) {
#if defined(__CUDA_ARCH__)
    return (float)sin((double)x);
#else
    return sinf(x);
#endif
}
QZ_HD_CALL float qz_cosf(float x) {
#if defined(__CUDA_ARCH__)
    return (float)cos((double)x);
#else
    return cosf(x);
#endif
}
// sinf and cosf of the same argument (every warp needs both): glibc's own algorithm.  glibc >= 2.28
// computes sinf / cosf / sincosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h -- the ARM
// optimized routines) in DOUBLE: quadrant n = round(x * 2/pi) by a scaled float-to-int conversion,
// r = x - n * pi/2, then a degree-7 sine or degree-8 cosine polynomial in r, rounded to float once.
// Restated here operation for operation, with the multiply-adds fused as in the FMA build of the functions
// that glibc selects on every x86-64 CPU of the last decade (a build without FMA differs from it in about
// one result in 1e8); IEEE double arithmetic gives the same double on the B200 as on x86-64, hence the
// same float -- including the ~1.4 % of arguments where glibc's result is not the
// correctly rounded one, which the earlier "sin() in double, rounded once" version got "wrong" by being
// right.  (Checked bit for bit against glibc 2.39 on 2e8 arguments; tests/test_gpu_parity.py::
// test_sincos_is_glibc.)  The oracle's g++ -O2 merges the reference's std::sin / std::cos pairs
// (sampler.cpp:22-23,34,58-61) into sincosf, whose two results are these same two polynomials.
// |x| >= 120 (never produced by the warps) falls back to the correctly rounded double functions.
QZ_HD float qz_sincos_poly(double x, double x2, int n, bool negate_cos) {
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = fma(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = x3 * x2;
        const double s = fma(x3, -0x1.555545995a603p-3, x);
        return (float)fma(x7, s1, s);
    }
    // cosine polynomial; glibc's second table entry (used when n & 2) holds the NEGATED coefficients, and
    // since rounding is symmetric that is exactly the negated result
    const double x4 = x2 * x2;
    const double c2 = fma(x2, 0x1.99343027bf8c3p-16, -0x1.6c087e89a359dp-10);
    const double c1 = fma(x2, -0x1.ffffffd0c621cp-2, 0x1p0);
    const double x6 = x4 * x2;
    const double c = fma(x4, 0x1.55553e1068f19p-5, c1);
    const float r = (float)fma(x6, c2, c);
    return negate_cos ? -r : r;
}
QZ_HD void qz_sincosf(float y, float& s, float& c) {
#if defined(__CUDA_ARCH__)
    const uint32_t top = (float_as_u32(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top < 0x3f4u) {                       // |y| < 0.75 (abstop12(pi/4))
        if (top < 0x398u) { s = y; c = 1.0f; return; }   // |y| < 2^-12
        const double x2 = x * x;
        s = qz_sincos_poly(x, x2, 0, false);
        c = qz_sincos_poly(x, x2, 1, false);
        return;
    }
    if (top >= 0x42fu) { s = qz_sinf(y); c = qz_cosf(y); return; }   // |y| >= 120, inf, nan
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = ((int)r + 0x800000) >> 24;
    x = fma(-(double)n, 0x1.921FB54442D18p0, x);
    const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const bool neg = (n & 2) != 0;
    const double xs = x * sign, x2 = x * x;
    s = qz_sincos_poly(xs, x2, n, neg);
    c = qz_sincos_poly(xs, x2, n ^ 1, neg);
#else
    s = sinf(y);
    c = cosf(y);
#endif
}
QZ_HD_CALL float qz_atan2f(float y, float x) {
#if defined(__CUDA_ARCH__)
    return (float)atan2((double)y, (double)x);
#else
    return atan2f(y, x);
#endif
}
QZ_HD_CALL float qz_acosf(float x) {
#if defined(__CUDA_ARCH__)
    return (float)acos((double)x);
#else
    return acosf(x);
#endif
}

#define QZ_PI 3.14159265358979323846       /* M_PI   */
#define QZ_PI_2 1.57079632679489661923     /* M_PI_2 */
#define QZ_PI_4 0.78539816339744830962     /* M_PI_4 */
#define QZ_1_PI 0.31830988618379067154     /* M_1_PI */
#define QZ_2_PI 0.63661977236758134308     /* M_2_PI (= 2/pi; the reference uses it where 2*pi was meant, sampler.cpp:33) */

// util.cpp:3-5
QZ_HD float lerpf(float a, float b, float t) { return a + t * (b - a); }
QZ_HD float r_lerp(float a, float b, float t) { return r_fma(t, b - a, a); }

// Image::save's tone path, one channel (image.cpp:10-15): 255 * powf(c, gamma).  powf is evaluated in double and rounded
// once (glibc's powf does the same internally); gamma == 1 returns c itself, as powf(c, 1) does.
QZ_HD float tone_value(float c, float gamma) {
    const float p = gamma == 1.0f ? c : (float)pow((double)c, (double)gamma);
    return 255.0f * p;
}
// What cv::imwrite stores for a CV_32F matrix in an 8-bit file (it converts with convertTo(CV_8U): cvRound, then
// saturate): round to nearest even, clamp to 0..255.  cvRound is the x86 float -> int32 conversion, which answers
// INT_MIN for NaN, the infinities and everything beyond the int32 range -- those all become 0, not 255 (pinned against
// OpenCV 4.13's imwrite by tests/test_abi_and_host.py).
QZ_HD uint8_t tone_u8(float v) {
    if (!(fabsf(v) < 2147483648.0f)) return 0;
    const float r = rintf(v);
    return (uint8_t)(r >= 255.0f ? 255.0f : (r > 0.0f ? r : 0.0f));
}

}  // namespace qz
