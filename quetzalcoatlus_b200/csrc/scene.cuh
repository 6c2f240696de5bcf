// Device view of a committed scene: raw pointers to the POD tables of include/qz_b200.h
// plus the wide BVH.  Passed to kernels by value (fits the 4 KB kernel parameter space).
#pragma once

#include "common.cuh"
#include "sampler.cuh"
#include "../../include/qz_b200.h"

namespace qz {

struct BvhNode;

#define QZ_PW_BUCKETS 472 /* thresholds 360 .. 831 nm */

struct F4 {
    float x, y, z, w;
};

struct DScene {
    const qz_spectrum* spectra;
    const qz_texture* textures;
    const qz_material* materials;
    const int32_t* mixed_children;
    const qz_light* lights;
    const qz_geometry* geoms;
    const F4* prims;            // 4 x F4 per primitive
    const float* pool;
    const uint8_t* pw_accel;    // 1-nm bucket tables of the piecewise-linear spectra (scene_store.cuh)
    const float* normals;       // float3 packed
    const int32_t* nidx;        // int4 per OBJ face
    const uint32_t* grid_dims;
    const float* lut_z;
    const float* lut_coeffs;
    const BvhNode* nodes;       // wide BVH, node 0 = root
    const uint32_t* leaf_prims; // primitive indices in leaf order
    const SamplerDim* sampler_table;
    const float* rho_tab;       // depth-0 albedo sample tables (see bxdf.cuh)
    SampleMemo memo;            // per-pass table of sampler values (sampler.cuh); tab == nullptr: none
    uint32_t* overflow;         // set to 1 by the per-thread traversal (bvh.cuh) when its stack overflowed
    uint32_t n_lights;
    uint32_t n_prims;
    uint32_t n_nodes;
    int32_t bg_spectrum;
    float bg_scale;
};

// camera + sensor on the device
struct DCamera {
    uint32_t width, height;
    V3 pos, bottom_left, du, dv;
    const float* sensor;  // 3 x 471
    float imaging_ratio;
};

}  // namespace qz
