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

}  // namespace qz
