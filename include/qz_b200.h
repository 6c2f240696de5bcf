/* qz_b200.h -- C ABI of the B200 path-tracing hot path.
 *
 * The reference (quevivasbien/quetzalcoatlus) has no FFI layer: its hot path sits behind the
 * C++ call `RenderResult render(const Camera&, const Scene&, size_t n_samples, size_t
 * max_bounces)` (src/render.hpp:8-13) and the scene-building methods of `Scene`
 * (src/scene.hpp:42-108).  This header is the thin C boundary SURVEY.md section 8.b
 * specifies underneath a C++ host library that mirrors that API name for name
 * (quetzalcoatlus_b200/host/).  Each entry point below cites the reference interface it
 * replaces.  Plain pointers and sizes only; no C++/torch types; `int` return, 0 = ok,
 * otherwise a qz_status and qz_last_error() holds a message.  The caller owns all host
 * buffers; the library owns device memory behind the opaque handle.  One host thread per
 * handle.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * QZ_ERR_NO_DEVICE.
 *
 * All tables are plain-old-data, uploaded once, immutable after qz_scene_commit().
 */
#ifndef QZ_B200_H
#define QZ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QZ_ABI_VERSION 1

typedef enum qz_status {
    QZ_OK = 0,
    QZ_ERR_NO_DEVICE = 1,     /* no CUDA device / driver: the product path refuses to run */
    QZ_ERR_INVALID = 2,       /* bad argument or table cross-reference */
    QZ_ERR_NOT_COMMITTED = 3, /* render() on an uncommitted scene (render.cpp:328-331) */
    QZ_ERR_CUDA = 4,          /* a CUDA call failed; message has the CUDA error string */
    QZ_ERR_OOM = 5
} qz_status;

/* ---- spectra (src/color/spectrum.hpp, rgb.hpp) ------------------------------------ */
typedef enum qz_spectrum_kind {
    QZ_SPEC_CONSTANT = 0,       /* ConstantSpectrum: a = value                               */
    QZ_SPEC_DENSE = 1,          /* DenselySampledSpectrum: pool[offset..offset+count), aux = lambda_min */
    QZ_SPEC_PIECEWISE = 2,      /* PiecewiseLinearSpectrum: pool[offset..) = count lambdas, 0, count values, 0
                                   (the trailing zeros reproduce the reference's one-past-the-end
                                   read at spectrum.cpp:107-108 deterministically)              */
    QZ_SPEC_SIGMOID = 3,        /* RGBSigmoidPolynomial: a,b,c = c0,c1,c2                    */
    QZ_SPEC_RGB_UNBOUNDED = 4,  /* RGBUnboundedSpectrum: a,b,c = c0,c1,c2, scale             */
    QZ_SPEC_RGB_ILLUMINANT = 5, /* RGBIlluminantSpectrum: a,b,c, scale, aux = illuminant id  */
    QZ_SPEC_BLACKBODY = 6       /* BlackbodySpectrum: a = T, b = normalisation factor        */
} qz_spectrum_kind;

typedef struct qz_spectrum {
    uint32_t kind;
    uint32_t offset; /* into the float pool */
    uint32_t count;
    int32_t aux;
    float a, b, c, scale;
} qz_spectrum; /* 32 bytes */

/* ---- textures (src/texture.hpp) ---------------------------------------------------- */
typedef enum qz_texture_kind {
    QZ_TEX_SOLID = 0, /* SolidColor: a = spectrum id                                         */
    QZ_TEX_DUMMY = 1, /* DummyTexture: a = white spectrum id, b = black spectrum id          */
    QZ_TEX_IMAGE = 2  /* ImageTexture: offset into the float pool (RGB interleaved), w, h    */
} qz_texture_kind;

typedef struct qz_texture {
    uint32_t kind;
    int32_t a, b;
    uint32_t offset;
    uint32_t width, height;
    uint32_t pad[2];
} qz_texture; /* 32 bytes */

/* ---- materials (src/material.hpp) -------------------------------------------------- */
typedef enum qz_material_kind {
    QZ_MAT_DIFFUSE = 0,         /* a = texture id                                            */
    QZ_MAT_CONDUCTOR = 1,       /* a = ior spectrum, b = absorption spectrum, alpha_x/y      */
    QZ_MAT_DIELECTRIC = 2,      /* a = ior spectrum, is_constant                             */
    QZ_MAT_THIN_DIELECTRIC = 3, /* a = ior spectrum, is_constant                             */
    QZ_MAT_MIXED = 4            /* children[a .. a+count) in the child-index table           */
} qz_material_kind;

typedef struct qz_material {
    uint32_t kind;
    int32_t a, b;
    uint32_t count;
    float alpha_x, alpha_y; /* TrowbridgeReitzDistribution after its constructor's clamp (bxdf.cpp:210-215) */
    uint32_t is_constant;
    uint32_t pad;
} qz_material; /* 32 bytes */

/* ---- lights (src/light.hpp, shape.hpp) --------------------------------------------- */
typedef enum qz_light_kind { QZ_LIGHT_POINT = 0, QZ_LIGHT_AREA_QUAD = 2, QZ_LIGHT_AREA_SPHERE = 3 } qz_light_kind;

typedef struct qz_light {
    uint32_t kind;
    int32_t spectrum;
    float scale;
    uint32_t two_sided;
    float p[3];        /* point position | quad p00 | sphere centre */
    float radius;
    float du[3];
    float inv_area;    /* 1.0f / area() as the reference computes it (shape.hpp:47,82) */
    float dv[3];
    float pad0;
    float normal[3];   /* Quad::m_normal */
    float pad1;
} qz_light; /* 80 bytes */

/* ---- geometry (src/scene.hpp:35-40 GeometryData, one per Scene::add_* call) --------- */
typedef enum qz_shape_kind { QZ_SHAPE_SPHERE = 0, QZ_SHAPE_TRIANGLE = 1, QZ_SHAPE_QUAD = 2, QZ_SHAPE_OBJ = 3, QZ_SHAPE_GRID = 4 } qz_shape_kind;

typedef struct qz_geometry {
    uint32_t shape;
    int32_t material;      /* -1 = null material (emitter geometry, scene.cpp:440,448) */
    int32_t light;         /* -1 = not an emitter */
    int32_t normal_offset; /* OBJ with vn: first float3 of this mesh in the normal pool, else -1 */
    int32_t nindex_offset; /* OBJ with vn: first int4 (per-face normal indices) in the index pool */
    uint32_t first_prim;   /* index of this geometry's first primitive record */
    uint32_t prim_count;
    uint32_t pad;
} qz_geometry; /* 32 bytes */

/* One primitive = 64 bytes = four float4.  x,y,z of v[0..3] are vertices (sphere: v[0] = centre,
 * v[1].x = radius); the w lanes carry integers (bit patterns):
 *   v[0].w = geomID   v[1].w = primID (as Embree reports it)   v[2].w = qz_prim_kind
 *   v[3].w = grid cell: x | y << 16 (the grid's resolution minus one is in grid_dims[geomID])
 * Primitives are ordered by (geomID, primID, cell); that order is the tie-break of the
 * closest-hit search.                                                                     */
typedef enum qz_prim_kind { QZ_PRIM_TRIANGLE = 0, QZ_PRIM_QUAD = 1, QZ_PRIM_SPHERE = 2, QZ_PRIM_GRIDCELL = 3 } qz_prim_kind;

typedef struct qz_prim {
    float v[4][4];
} qz_prim;

typedef struct qz_scene_tables {
    const qz_spectrum* spectra;   uint32_t n_spectra;
    const qz_texture* textures;   uint32_t n_textures;
    const qz_material* materials; uint32_t n_materials;
    const int32_t* mixed_children; uint32_t n_mixed_children;
    const qz_light* lights;       uint32_t n_lights;
    const qz_geometry* geometries; uint32_t n_geometries;
    const qz_prim* prims;         uint32_t n_prims;
    const float* pool;            uint32_t n_pool;      /* float pool: spectra samples, texture images */
    const float* normals;         uint32_t n_normals;   /* float3 vertex normals of all OBJ meshes */
    const int32_t* normal_indices; uint32_t n_normal_indices; /* int4 per OBJ face */
    const uint32_t* grid_dims;    /* per geometry: (W-1) | (H-1) << 16 for grids, else 0; n_geometries entries */
    int32_t bg_spectrum;          /* Scene::set_bg_light (scene.cpp:461), -1 = none */
    float bg_scale;
    const float* rgb2spec_z;      /* 32 z nodes (rgb.cpp:78-140)           */
    const float* rgb2spec_coeffs; /* 3*32*32*32*3 coefficients             */
} qz_scene_tables;

/* Camera (src/camera.hpp:24-33) with the PixelSensor baked to its three dense curves
 * (sensor.hpp:31-35: m_r/m_g/m_b, lambda 360..830 nm, 471 samples each).                 */
typedef struct qz_camera {
    uint32_t image_width, image_height;
    float pos[3];
    float viewport_bottom_left[3];
    float pixel_delta_u[3];
    float pixel_delta_v[3];
    const float* sensor_rgb; /* 3 x 471 floats: r, g, b */
    float imaging_ratio;
} qz_camera;

/* Which part of the film this call renders.  Rows are film rows (row 0 = image top,
 * render.cpp:261-262).  With n_shards > 1 the call owns the strips
 * (row / strip_rows) % n_shards == shard; other rows are left untouched (zero them first
 * and a sum-reduce over ranks assembles the film bit-identically to a 1-GPU render).     */
typedef struct qz_region {
    uint32_t strip_rows; /* 0 = whole film */
    uint32_t n_shards;
    uint32_t shard;
} qz_region;

#define QZ_FLAG_UNSORTED_SHADING 1u /* one uber shading kernel over the unsorted queue (evidence runs only) */
#define QZ_FLAG_COUNT_TRAVERSAL 2u  /* count wide-node visits and primitive tests (slower; for B_ray)        */
#define QZ_FLAG_FORCE_BVH 8u        /* use the BVH traversal kernels even for scenes small enough for the flat kernel */
#define QZ_FLAG_LANE_TRAVERSAL 16u  /* (round-1 evidence arm, no longer built: QZ_ERR_INVALID) */
#define QZ_FLAG_OCTET_TRAVERSAL 32u /* (round-1 evidence arm, no longer built: QZ_ERR_INVALID) */
#define QZ_FLAG_STAGE_TIMING 4u     /* CUDA events around every stage (serialises the pipeline; for profiles) */
/* Arithmetic mode.  By default radiometric values (spectra, BSDF values, pdfs, MIS weights, throughput, sensor
 * response) use fused multiply-adds and the hardware reciprocal / square root: a few 1e-7 relative per operation,
 * geometry and discrete decisions untouched (per-path radiance within ~1e-6 of the reference, 1e-4 allowed).  With
 * this flag every float operation is performed as the reference's x86-64 build performs it: paths are
 * bit-identical to the reference's, at about twice the shading time.                                             */
#define QZ_FLAG_EXACT_ARITHMETIC 64u
/* The per-pass sample memo (csrc/sampler.cuh: sampler values and hot spectra tabulated per Halton index) is used when the
 * call owns at least four times more pixels than (x mod 128, y mod 128) classes; these two flags force it on or off
 * (parity tests: the film is bit-identical either way).                                                            */
#define QZ_FLAG_FORCE_MEMO 128u
#define QZ_FLAG_NO_MEMO 256u

typedef struct qz_render_options {
    uint32_t flags;
    uint32_t pool_paths;      /* paths in flight, 0 = default */
    uint32_t samples_per_pass; /* film pass granularity, 0 = auto */
    uint32_t reserved;
} qz_render_options;

typedef struct qz_stats {
    uint64_t paths;
    uint64_t rays_closest;  /* closest-hit queries, incl. emitter pass-throughs (render.cpp:108) */
    uint64_t rays_shadow;   /* Scene::occluded calls (render.cpp:75)                             */
    uint64_t shade_calls;   /* path-bounces shaded                                               */
    uint64_t iterations;    /* wavefront iterations                                              */
    uint64_t kernel_launches;
    uint64_t node_visits, prim_tests; /* only with QZ_FLAG_COUNT_TRAVERSAL */
    float ms_total;         /* device time of the whole call (CUDA events)                       */
    float ms_closest, ms_shadow, ms_shade, ms_other; /* per-stage device time (QZ_FLAG_STAGE_TIMING) */
    uint32_t bvh_nodes, bvh_bytes;
    float ms_sample;        /* the sampler stage (k_sample)                                      */
    uint32_t stack_overflows; /* warps that ran out of traversal stack (render fails if non-zero) */
} qz_stats;

typedef struct qz_scene_t* qz_scene;

const char* qz_last_error(void);
int qz_abi_version(void);

/* Selects the CUDA device for subsequent handles created by this thread
 * (replaces initialize_device(), scene.cpp:12-20).                                        */
int qz_init(int device);
int qz_device_name(char* buf, size_t n);

/* Flags OR-ed into the options of every later qz_render* / qz_trace_paths call of this thread (the reference's
 * render() has no options argument: this is how a host application selects QZ_FLAG_EXACT_ARITHMETIC). Returns
 * the previous value.                                                                                     */
uint32_t qz_set_default_flags(uint32_t flags);

/* Multi-GPU behind render() (the reference fans render() out over host threads, render.cpp:339-382; here over GPUs):
 * scenes committed by this thread after qz_set_device_count(n), n > 1, are replicated on the n device ordinals starting
 * at the thread's device (tables uploaded, BVH built on each), and qz_render() then renders interleaved row strips on all
 * of them, one host thread per device; each device's film kernel writes its rows straight into the first device's film
 * planes through NVLink peer memory, so there is no separate collective and no host bounce, and the film is
 * bit-identical to a one-GPU render.  n = 0 (the default) takes the count from the environment variable QZ_DEVICES
 * (so an unmodified application of the reference's API can be switched over from outside).  Returns the previous value. */
int qz_set_device_count(int n);

/* Rows per interleaved strip that qz_render() uses when it shards a film of `height` rows over `n_shards` devices, for a
 * multi-process driver that wants the same partition (qz_region.strip_rows): the largest height <= 8 that gives every
 * shard the same number of rows, powers of two first -- with strip * n_shards dividing 128 a shard owns 1/n_shards of the
 * sampler's (y mod 128) pixel classes.  Pure function; needs no device.  (The reference tiles its film over host threads,
 * render.cpp:339-382.)                                                                                              */
uint32_t qz_strip_rows(uint32_t height, uint32_t n_shards);

/* Scene::Scene / ~Scene (scene.hpp:44-49) */
int qz_scene_create(qz_scene* out);
int qz_scene_destroy(qz_scene scene);

/* Scene::commit (scene.cpp:22-25): uploads the tables and builds the wide BVH on the GPU
 * (replaces rtcCommitScene).                                                              */
int qz_scene_commit(qz_scene scene, const qz_scene_tables* tables);

/* Device time of the BVH build of the last qz_scene_commit() on this handle (CUDA events around the build kernels:
 * primitive boxes, Morton keys, radix sort, LBVH topology + fit, collapse to 8-wide nodes, leaf gather) -- the
 * counterpart of rtcCommitScene's build, reported as a secondary metric.                        */
int qz_scene_build_ms(qz_scene scene, float* ms);

/* render() (render.hpp:8-13, render.cpp:321-397) with HOST output buffers: H*W*3 floats
 * each, RGB interleaved, row 0 = top (image.hpp:26-45).  Any of normal/albedo may be NULL.
 * Limits (QZ_ERR_INVALID beyond them): max_bounces <= 65535; n_samples * sampler stride < 2^31 (stride = 31104 from
 * 128 x 128 pixels on, i.e. n_samples <= 69042) -- past that the reference's own `sample_index * sample_stride`
 * (sampler.cpp:419, int * int) overflows, which is undefined there.  The probes below apply the same rule to `s`.   */
int qz_render(qz_scene scene, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces,
              const qz_region* region, const qz_render_options* options,
              float* color, float* normal, float* albedo, qz_stats* stats);

/* Same, but the three planes are DEVICE pointers (e.g. torch tensors) and the work is
 * enqueued on `cuda_stream` (a cudaStream_t cast to void*, NULL = default stream); the
 * call returns after the stream has been synchronised so that stats are valid.           */
int qz_render_device(qz_scene scene, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces,
                     const qz_region* region, const qz_render_options* options,
                     float* d_color, float* d_normal, float* d_albedo, void* cuda_stream, qz_stats* stats);

/* AOV hand-off to a GPU denoiser (RenderResult::denoise, image.cpp:47-95, copies the three planes into OIDN buffers):
 * DEVICE pointers of the colour / normal / albedo planes of the last qz_render() on this handle (H*W*3 floats each;
 * a plane that was not requested is NULL), valid until the next render or qz_scene_destroy().                    */
int qz_film_device(qz_scene scene, float** d_color, float** d_normal, float** d_albedo, uint32_t* width, uint32_t* height);

/* Image::save's tone path (image.cpp:7-19) as a kernel over a film plane: bgr255[3i + k] = 255 * powf(rgb[3i + 2 - k],
 * gamma) -- the BGR float image the reference hands to cv::imwrite -- and/or bgr8 = that value saturated to 8 bits
 * (round to nearest even, clamp to 0..255: OpenCV's conversion of a CV_32F matrix for an 8-bit file).  Either output
 * may be NULL.  qz_tone_device takes DEVICE pointers and enqueues on `cuda_stream`; qz_tone takes host buffers.      */
int qz_tone_device(const float* d_rgb, uint32_t n_pixels, float gamma, float* d_bgr255, uint8_t* d_bgr8, void* cuda_stream);
int qz_tone(const float* rgb, uint32_t n_pixels, float gamma, float* bgr255, uint8_t* bgr8);

/* Per-path replay (the parity harness; the reference equivalent is the loop body
 * render.cpp:268-277 around sample_pixel(), render.cpp:91): n x (x, y, s) with y the
 * sampler/camera y (= H-1-row); 32 floats per path:
 *   [0..3] lambda  [4..7] final lambda pdf  [8..11] radiance  [12..14] normal
 *   [15] rays issued  [16..19] albedo  [20..22] sensor rgb  [23..25] albedo rgb            */
int qz_trace_paths(qz_scene scene, const qz_camera* camera, uint32_t n_samples, uint32_t max_bounces,
                   uint32_t n, const int32_t* xys, float* records);

/* Sampler known-answer entry (Sampler::start_pixel_sample / sample_1d / sample_pixel,
 * sampler.cpp:404-454): n x (x, y, s, dim); dim 0/1 = pixel jitter, dim >= 2 = that dimension. */
int qz_sampler_eval(uint32_t n_samples, uint32_t width, uint32_t height, uint32_t n, const int32_t* q, float* out);

/* Closest-hit probe of the traversal kernel (replaces create_rayhit/rtcIntersect1,
 * scene.cpp:41-59): n x (o, d) -> n x (t, u, v, Ng.xyz, geomID, primID), t = -1 on a miss. */
int qz_intersect(qz_scene scene, uint32_t n, const float* rays, float* out);

/* Spectrum::operator() on the device, for table parity: evaluates spectrum `id` of the
 * committed scene (id < 0: the background spectrum) at n wavelengths.                     */
int qz_eval_spectrum(qz_scene scene, int32_t id, uint32_t n, const float* lambdas, float* out);

/* PixelSensor::to_sensor_rgb on the device (sensor.cpp:57-70): n x (u, L0..L3) -> n x rgb */
int qz_sensor_eval(const qz_camera* camera, uint32_t n, const float* in, float* out);

/* Device math probe, for parity of the restated libm functions: op 0 = (sinf, cosf) of n arguments -> 2n floats
 * (csrc/math.cuh restates glibc's sincosf, which the oracle's std::sin / std::cos calls compile to).           */
int qz_math_probe(int op, uint32_t n, const float* in, float* out);

#ifdef __cplusplus
}
#endif

#endif /* QZ_B200_H */
