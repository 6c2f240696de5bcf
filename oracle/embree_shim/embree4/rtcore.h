// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Stand-in for <embree4/rtcore.h> (Intel Embree 4.3, pinned by the reference in
// CMakeLists.txt:12 / readme.md:21; un-vendored, absent from /root/reference and
// from this image).  It declares exactly the 15 entry points and the types the
// reference calls (scene.cpp:13-18,23,41-58,146-185,197-238,256-284,300-370,
// 376-427; scene.hpp:44-48,85) so that the reference's own sources compile and
// link unmodified.  The arithmetic behind rtcIntersect1 is RESTATED from Embree's
// published algorithm in rtcore_shim.cpp -- parity against real Embree is
// UNPINNED (no Embree binary, no reference test vectors at this boundary).
#pragma once

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1

typedef struct RTCDeviceTy* RTCDevice;
typedef struct RTCSceneTy* RTCScene;
typedef struct RTCGeometryTy* RTCGeometry;

enum RTCError {
    RTC_ERROR_NONE = 0,
    RTC_ERROR_UNKNOWN = 1,
    RTC_ERROR_INVALID_ARGUMENT = 2,
    RTC_ERROR_INVALID_OPERATION = 3,
    RTC_ERROR_OUT_OF_MEMORY = 4,
    RTC_ERROR_UNSUPPORTED_CPU = 5,
    RTC_ERROR_CANCELLED = 6
};

enum RTCGeometryType {
    RTC_GEOMETRY_TYPE_TRIANGLE = 0,
    RTC_GEOMETRY_TYPE_QUAD = 1,
    RTC_GEOMETRY_TYPE_GRID = 2,
    RTC_GEOMETRY_TYPE_SPHERE_POINT = 50
};

enum RTCBufferType {
    RTC_BUFFER_TYPE_INDEX = 0,
    RTC_BUFFER_TYPE_VERTEX = 1,
    RTC_BUFFER_TYPE_GRID = 8
};

enum RTCFormat {
    RTC_FORMAT_UNDEFINED = 0,
    RTC_FORMAT_UINT3 = 0x5003,
    RTC_FORMAT_UINT4 = 0x5004,
    RTC_FORMAT_FLOAT3 = 0x9003,
    RTC_FORMAT_FLOAT4 = 0x9004,
    RTC_FORMAT_GRID = 0xA001
};

struct RTCGrid {
    unsigned int startVertexID;
    unsigned int stride;
    unsigned short width, height;
};

struct RTCRay {
    float org_x, org_y, org_z;
    float tnear;
    float dir_x, dir_y, dir_z;
    float time;
    float tfar;
    unsigned int mask;
    unsigned int id;
    unsigned int flags;
};

struct RTCHit {
    float Ng_x, Ng_y, Ng_z;
    float u, v;
    unsigned int primID;
    unsigned int geomID;
    unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};

struct RTCRayHit {
    struct RTCRay ray;
    struct RTCHit hit;
};

struct RTCIntersectArguments;

typedef void (*RTCErrorFunction)(void* userPtr, enum RTCError code, const char* str);

RTCDevice rtcNewDevice(const char* config);
void rtcReleaseDevice(RTCDevice device);
enum RTCError rtcGetDeviceError(RTCDevice device);
void rtcSetDeviceErrorFunction(RTCDevice device, RTCErrorFunction error, void* userPtr);

RTCScene rtcNewScene(RTCDevice device);
void rtcReleaseScene(RTCScene scene);
void rtcCommitScene(RTCScene scene);
void* rtcGetGeometryUserDataFromScene(RTCScene scene, unsigned int geomID);

RTCGeometry rtcNewGeometry(RTCDevice device, enum RTCGeometryType type);
void* rtcSetNewGeometryBuffer(RTCGeometry geometry, enum RTCBufferType type, unsigned int slot,
                              enum RTCFormat format, size_t byteStride, size_t itemCount);
void rtcSetGeometryUserData(RTCGeometry geometry, void* ptr);
void rtcCommitGeometry(RTCGeometry geometry);
unsigned int rtcAttachGeometry(RTCScene scene, RTCGeometry geometry);
void rtcReleaseGeometry(RTCGeometry geometry);

#ifdef __cplusplus
void rtcIntersect1(RTCScene scene, struct RTCRayHit* rayhit, struct RTCIntersectArguments* args = nullptr);
#else
void rtcIntersect1(RTCScene scene, struct RTCRayHit* rayhit, struct RTCIntersectArguments* args);
#endif

// ---- shim-only instrumentation (not part of Embree) ------------------------
// total rtcIntersect1 calls since the last reset, summed over all threads
unsigned long long rtcShimRayCount(void);
void rtcShimResetRayCount(void);
// rays issued by the calling thread only (used for per-path ray counts)
unsigned long long rtcShimThreadRayCount(void);
// force brute-force intersection (no BVH) for scenes committed afterwards; used
// by the tests to check the shim's own BVH against its brute-force definition
void rtcShimForceBruteForce(int on);

#ifdef __cplusplus
}
#endif
