// Shared small types of the device code: 3-vectors, 4-wide spectral samples and the
// comparison helpers whose NaN/tie behaviour must match the C++ standard library calls the
// reference uses.
//
// ARITHMETIC CONTRACT.  The whole translation unit is compiled with FMA contraction off
// (nvcc -fmad=false), IEEE division and square root (nvcc defaults), no flush-to-zero and no
// fast-math; every expression below keeps the operand order of the reference expression it
// restates, including the places where the reference silently computes in double
// (a float multiplied by M_PI-style double constants).  Under that contract float32
// +,-,*,/,sqrt are bit-identical between the B200 and the x86-64 oracle build, so per-path
// differences can only come from transcendental functions (see math.cuh).
//
// Functions are marked QZ_HD: under nvcc that is __host__ __device__; the same headers
// also compile under plain g++ for the TEST-ONLY emulation build (tests/emu), which lets the
// CPU test-suite check this restatement against the oracle without a GPU.  The emulation
// build is never part of the shipped libraries.
#pragma once

#include <stdint.h>
#include <math.h>

// ARITHMETIC MODES.  The device code exists in two builds (two translation units over the same
// headers, csrc/k_shade_exact.cu and csrc/k_shade_fast.cu; the second renames the namespace):
//   QZ_FAST == 0  every float operation as the reference's x86-64 build performs it (see below);
//   QZ_FAST == 1  GEOMETRY and every DISCRETE DECISION still so -- ray origins and directions, hit
//                 points, normals, frames, sampled directions, the dielectric interface (index of
//                 refraction, Fresnel split, reflect-or-refract choice), texel and light picks --
//                 while RADIOMETRIC values (spectra, BSDF values, pdfs, MIS weights, throughput,
//                 sensor response, the depth-0 albedo AOV) use explicit fused multiply-adds, the
//                 hardware reciprocal / reciprocal square root and float instead of double
//                 intermediates: a few 1e-7 relative per operation, nothing fed back into the
//                 geometry.  The only discrete decision that reads a radiometric value is Russian
//                 roulette (throughput against a sample).
// Both builds are compiled with -fmad=false: a fused multiply-add exists only where the source
// says r_fma(), so a function gives the same bits in every kernel it is inlined into (the
// wavefront film stays bit-identical to the per-path replay in either mode).
#ifndef QZ_FAST
#define QZ_FAST 0
#endif

#if defined(__CUDACC__)
#define QZ_HD __host__ __device__ __forceinline__
#define QZ_D __device__ __forceinline__
// QZ_HD_CALL: large, rarely repeated bodies (double-precision transcendentals, the spectrum
// switch) are real calls on the device.  Inlined at every site they made each shading kernel
// 350-700 KB of straight-line code that every warp streams through once per bounce: the
// instruction caches miss constantly (no_instruction stalls, profiles/r01_summary.md).
#ifndef QZ_INLINE_EVERYTHING
#define QZ_HD_CALL inline __host__ __device__ __noinline__
#else
#define QZ_HD_CALL inline __host__ __device__ __forceinline__
#endif
#else
#define QZ_HD inline
#define QZ_D inline
#define QZ_HD_CALL inline
#endif

namespace qz {

struct V3 {
    float x, y, z;
};

QZ_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
QZ_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
QZ_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
QZ_HD V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
QZ_HD V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
// IEEE a / b.  On the device a ZERO NUMERATOR sends the division into its out-of-line slow path (FCHK flags it), ~30
// instructions and a divergent call -- and the components of an axis-aligned normal are zeros: normalising the
// geometric normal of a wall and building its frame took that path four times per bounce, a tenth of the diffuse
// shading kernel's instructions (profiles/r02_summary.md).  0 / b for an ordinary b is a zero with the signs xor-ed;
// everything else is the division itself.
QZ_HD float div_zn(float a, float b) {
#if defined(__CUDA_ARCH__)
    const bool z = a == 0.0f && b != 0.0f && b == b;
    const float q = (z ? 1.0f : a) / b;
    return z ? __uint_as_float(__float_as_uint(a) ^ (__float_as_uint(b) & 0x80000000u)) : q;
#else
    return a / b;
#endif
}
QZ_HD V3 operator/(V3 a, float s) { return v3(div_zn(a.x, s), div_zn(a.y, s), div_zn(a.z, s)); }
// vec.cpp:116-126: (x*x' + y*y') + z*z'
QZ_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
QZ_HD V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
QZ_HD float norm_squared(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
QZ_HD float norm(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
// vec.cpp:99-102: divide each component by the norm (no reciprocal)
QZ_HD V3 normalized(V3 a) { float n = norm(a); return v3(div_zn(a.x, n), div_zn(a.y, n), div_zn(a.z, n)); }

struct V2 {
    float x, y;
};
QZ_HD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }

// std::max / std::min / std::clamp semantics (first argument wins ties and NaN compares false)
QZ_HD float std_max(float a, float b) { return (a < b) ? b : a; }
QZ_HD float std_min(float a, float b) { return (b < a) ? b : a; }
QZ_HD float std_clamp(float v, float lo, float hi) { return (v < lo) ? lo : ((hi < v) ? hi : v); }
QZ_HD bool is_inf(float x) { return x == INFINITY || x == -INFINITY; }

// SpectrumSample (spectrum_sample.hpp:37-101): four floats, component-wise arithmetic
struct Spec4 {
    float v[4];
};

QZ_HD Spec4 spec4(float c) { Spec4 s; s.v[0] = c; s.v[1] = c; s.v[2] = c; s.v[3] = c; return s; }
QZ_HD Spec4 spec4(float a, float b, float c, float d) { Spec4 s; s.v[0] = a; s.v[1] = b; s.v[2] = c; s.v[3] = d; return s; }
QZ_HD Spec4 operator+(Spec4 a, Spec4 b) { return spec4(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]); }
QZ_HD Spec4 operator*(Spec4 a, Spec4 b) { return spec4(a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]); }
QZ_HD Spec4 operator*(Spec4 a, float c) { return spec4(a.v[0] * c, a.v[1] * c, a.v[2] * c, a.v[3] * c); }
// division yields 0 where the divisor is 0 (spectrum_sample.cpp:119-136, 183-201)
QZ_HD Spec4 operator/(Spec4 a, Spec4 b) {
    return spec4(b.v[0] == 0.0f ? 0.0f : a.v[0] / b.v[0], b.v[1] == 0.0f ? 0.0f : a.v[1] / b.v[1],
                 b.v[2] == 0.0f ? 0.0f : a.v[2] / b.v[2], b.v[3] == 0.0f ? 0.0f : a.v[3] / b.v[3]);
}
QZ_HD Spec4 operator/(Spec4 a, float c) {
    if (c == 0.0f) return spec4(0.0f);
    return spec4(a.v[0] / c, a.v[1] / c, a.v[2] / c, a.v[3] / c);
}
QZ_HD bool is_zero(Spec4 a) { return !(a.v[0] != 0.0f || a.v[1] != 0.0f || a.v[2] != 0.0f || a.v[3] != 0.0f); }
QZ_HD float max_component(Spec4 a) {
    float m = a.v[0];
    if (a.v[1] > m) m = a.v[1];
    if (a.v[2] > m) m = a.v[2];
    if (a.v[3] > m) m = a.v[3];
    return m;
}
// spectrum_sample.cpp:203-209: running sum from 0, then / 4
QZ_HD float average(Spec4 a) {
    float sum = 0.0f;
    sum += a.v[0]; sum += a.v[1]; sum += a.v[2]; sum += a.v[3];
    return sum / 4.0f;
}

// ---- radiometric primitives (see ARITHMETIC MODES above).  With QZ_FAST == 0 each is the plain
// IEEE expression, so code written with them is still the reference's arithmetic.
QZ_HD float r_fma(float a, float b, float c) {
#if QZ_FAST
    return fmaf(a, b, c);
#else
    return a * b + c;
#endif
}
// (the .approx.ftz forms are one MUFU instruction each; 1-2 ulp, denormal operands count as zero)
QZ_HD float r_rcp(float x) {
#if QZ_FAST && defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
QZ_HD float r_div(float a, float b) {
#if QZ_FAST && defined(__CUDA_ARCH__)
    return a * r_rcp(b);
#else
    return a / b;
#endif
}
QZ_HD float r_sqrt(float x) {
#if QZ_FAST && defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return sqrtf(x);
#endif
}
QZ_HD float r_rsqrt(float x) {
#if QZ_FAST && defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / sqrtf(x);
#endif
}
// radiometric Spec4 division: 0 where the divisor is 0, like operator/
QZ_HD Spec4 r_div(Spec4 a, float c) {
#if QZ_FAST
    if (c == 0.0f) return spec4(0.0f);
    const float i = r_rcp(c);
    return spec4(a.v[0] * i, a.v[1] * i, a.v[2] * i, a.v[3] * i);
#else
    return a / c;
#endif
}
QZ_HD Spec4 r_div(Spec4 a, Spec4 b) {
#if QZ_FAST
    return spec4(b.v[0] == 0.0f ? 0.0f : r_div(a.v[0], b.v[0]), b.v[1] == 0.0f ? 0.0f : r_div(a.v[1], b.v[1]),
                 b.v[2] == 0.0f ? 0.0f : r_div(a.v[2], b.v[2]), b.v[3] == 0.0f ? 0.0f : r_div(a.v[3], b.v[3]));
#else
    return a / b;
#endif
}
// direction used for radiometric terms only (never stored in the path)
QZ_HD V3 r_normalized(V3 a) {
#if QZ_FAST
    const float i = r_rsqrt(r_fma(a.z, a.z, r_fma(a.y, a.y, a.x * a.x)));
    return v3(a.x * i, a.y * i, a.z * i);
#else
    return normalized(a);
#endif
}

QZ_HD float u32_as_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { uint32_t u; float f; } c; c.u = u; return c.f;
#endif
}
QZ_HD uint32_t float_as_u32(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { uint32_t u; float f; } c; c.f = f; return c.u;
#endif
}
QZ_HD uint32_t umulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

}  // namespace qz
