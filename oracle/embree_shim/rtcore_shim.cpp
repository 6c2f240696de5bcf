// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Restatement of the Intel Embree 4.3 entry points the reference uses (see
// embree4/rtcore.h in this directory for the call-site list).  Embree itself is a
// third-party dependency that is NOT under /root/reference and not installed in
// this image, so its algorithm is restated here from its published sources
// (kernels/geometry/triangle_intersector_moeller.h, quad_intersector_moeller.h,
// sphere_intersector.h, grid_soa / subgrid intersectors) and this file DEFINES the
// arithmetic for this project.  PARITY AGAINST REAL EMBREE IS UNPINNED: the
// reference ships no test vectors at this boundary and no Embree binary exists
// here to produce any.
//
// Conventions fixed by this shim (the CUDA traversal kernel in
// quetzalcoatlus_b200/csrc/intersect.cuh follows the same operation order, with
// FMA contraction disabled, so both sides are bit-identical):
//
//  * plain IEEE float32, no FMA; dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z;
//    cross(a,b) = (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x).
//  * Triangle (v0,v1,v2) -- Moeller-Trumbore as in Embree:
//        e1 = v0-v1, e2 = v2-v0, Ng = cross(e2,e1)   (= (v1-v0)x(v2-v0))
//        C = v0-O, R = cross(C,D), den = dot(Ng,D), absDen = |den|, s = sign(den)
//        U = dot(R,e2)^s, V = dot(R,e1)^s, T = dot(Ng,C)^s
//        hit iff den != 0, U >= 0, V >= 0, U+V <= absDen, absDen*tnear < T,
//                T <= absDen*tfar(ray)
//        t = T/absDen, u = U/absDen, v = V/absDen      (true division; Embree uses
//        a Newton-refined rcp -- for the axis-aligned emitters of the shipped scenes
//        both give t == 1 exactly on NEE shadow rays)
//    No back-face culling.
//  * Quad (v0,v1,v2,v3) = triangle (v0,v1,v3) then triangle (v2,v3,v1); on the second
//    u = (absDen-U)/absDen, v = (absDen-V)/absDen.  Ng is the tested triangle's.
//  * Grid: (W-1)(H-1) cells, cell (x,y) = quad (p[y][x], p[y][x+1], p[y+1][x+1],
//    p[y+1][x]); u = (x + u_cell)/(W-1), v = (y + v_cell)/(H-1); primID = grid index.
//  * SPHERE_POINT (c,r): rd2 = 1/dot(D,D), c0 = c-O, projC0 = dot(c0,D)*rd2,
//    perp = c0 - projC0*D, l2 = dot(perp,perp), r2 = r*r, hit iff l2 <= r2,
//    td = sqrt((r2-l2)*rd2), t_front = projC0-td, t_back = projC0+td, each valid iff
//    tnear <= t <= tfar(ray); front preferred; Ng = (-/+td)*D - perp; u = v = 0.
//  * Closest hit = minimum t over all primitives; ties go to the lower
//    (geomID, primID, cell, triangle-half).  This makes the answer independent of
//    traversal order, so brute force, the BVH below and the GPU BVH all agree.
//  * geomIDs are handed out sequentially by rtcAttachGeometry.

#include "embree4/rtcore.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <vector>

namespace {

struct F3 { float x, y, z; };
inline F3 sub(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline F3 mul(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline F3 cross(F3 a, F3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float xor_sign(float v, uint32_t s) {
    uint32_t b; std::memcpy(&b, &v, 4); b ^= s; std::memcpy(&v, &b, 4); return v;
}
inline uint32_t sign_mask(float v) {
    uint32_t b; std::memcpy(&b, &v, 4); return b & 0x80000000u;
}

enum PrimKind : uint32_t { PK_TRI = 0, PK_QUAD = 1, PK_SPHERE = 2, PK_GRIDCELL = 3 };

struct Prim {
    F3 v[4];          // tri: v0..v2; quad/cell: v0..v3; sphere: v[0] = centre
    float radius;
    uint32_t kind;
    uint32_t geomID, primID;
    uint32_t cx, cy;  // grid cell coordinates
    uint32_t gw, gh;  // grid resolution minus one
};

struct Hit {
    float t = std::numeric_limits<float>::infinity();
    float u = 0, v = 0;
    F3 ng{0, 0, 0};
    uint32_t prim = 0xffffffffu;  // index into Scene::prims (ordered by geomID, primID, cell)
};

struct TriHit { float t, u, v; F3 ng; };

inline bool tri_test(F3 O, F3 D, float tnear, float tfar, F3 v0, F3 v1, F3 v2, bool flip, TriHit& h) {
    F3 e1 = sub(v0, v1);
    F3 e2 = sub(v2, v0);
    F3 ng = cross(e2, e1);
    F3 C = sub(v0, O);
    F3 R = cross(C, D);
    float den = dot(ng, D);
    float absDen = std::fabs(den);
    uint32_t s = sign_mask(den);
    float U = xor_sign(dot(R, e2), s);
    float V = xor_sign(dot(R, e1), s);
    if (!(den != 0.0f) || !(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= absDen)) return false;
    float T = xor_sign(dot(ng, C), s);
    if (!(absDen * tnear < T) || !(T <= absDen * tfar)) return false;
    h.t = T / absDen;
    if (flip) {
        h.u = (absDen - U) / absDen;
        h.v = (absDen - V) / absDen;
    } else {
        h.u = U / absDen;
        h.v = V / absDen;
    }
    h.ng = ng;
    return true;
}

inline bool sphere_test(F3 O, F3 D, float tnear, float tfar, F3 c, float r, TriHit& h) {
    float rd2 = 1.0f / dot(D, D);
    F3 c0 = sub(c, O);
    float projC0 = dot(c0, D) * rd2;
    F3 perp = sub(c0, mul(D, projC0));
    float l2 = dot(perp, perp);
    float r2 = r * r;
    if (!(l2 <= r2)) return false;
    float td = std::sqrt((r2 - l2) * rd2);
    float t_front = projC0 - td;
    float t_back = projC0 + td;
    if (tnear <= t_front && t_front <= tfar) {
        h.t = t_front;
        h.ng = sub(mul(D, -td), perp);
    } else if (tnear <= t_back && t_back <= tfar) {
        h.t = t_back;
        h.ng = sub(mul(D, td), perp);
    } else {
        return false;
    }
    h.u = 0.0f;
    h.v = 0.0f;
    return true;
}

// test one primitive; update `best` under the (t, prim index) order
inline void prim_test(const Prim& p, uint32_t index, F3 O, F3 D, float tnear, float tfar, Hit& best) {
    TriHit h;
    bool found = false;
    if (p.kind == PK_SPHERE) {
        found = sphere_test(O, D, tnear, tfar, p.v[0], p.radius, h);
    } else if (p.kind == PK_TRI) {
        found = tri_test(O, D, tnear, tfar, p.v[0], p.v[1], p.v[2], false, h);
    } else {
        TriHit a, b;
        bool fa = tri_test(O, D, tnear, tfar, p.v[0], p.v[1], p.v[3], false, a);
        bool fb = tri_test(O, D, tnear, tfar, p.v[2], p.v[3], p.v[1], true, b);
        if (fa && (!fb || a.t <= b.t)) { h = a; found = true; }
        else if (fb) { h = b; found = true; }
        if (found && p.kind == PK_GRIDCELL) {
            h.u = (float(p.cx) + h.u) / float(p.gw);
            h.v = (float(p.cy) + h.v) / float(p.gh);
        }
    }
    if (!found) return;
    if (h.t < best.t || (h.t == best.t && index < best.prim)) {
        best.t = h.t; best.u = h.u; best.v = h.v; best.ng = h.ng; best.prim = index;
    }
}

struct Box { float lo[3], hi[3]; };

inline Box prim_box(const Prim& p) {
    Box b;
    if (p.kind == PK_SPHERE) {
        const float c[3] = {p.v[0].x, p.v[0].y, p.v[0].z};
        for (int a = 0; a < 3; a++) { b.lo[a] = c[a] - p.radius; b.hi[a] = c[a] + p.radius; }
    } else {
        int n = (p.kind == PK_TRI) ? 3 : 4;
        for (int a = 0; a < 3; a++) { b.lo[a] = std::numeric_limits<float>::infinity(); b.hi[a] = -b.lo[a]; }
        for (int i = 0; i < n; i++) {
            const float c[3] = {p.v[i].x, p.v[i].y, p.v[i].z};
            for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], c[a]); b.hi[a] = std::max(b.hi[a], c[a]); }
        }
    }
    // conservative padding: the box test must never reject a primitive whose own
    // (rounded) intersection arithmetic reports a hit
    for (int a = 0; a < 3; a++) {
        float m = std::max(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
        float pad = 1e-5f * m + 1e-7f;
        b.lo[a] -= pad; b.hi[a] += pad;
    }
    return b;
}

struct Node {
    Box box;
    uint32_t left;   // inner: index of left child (right = left+1); leaf: first prim ref
    uint32_t count;  // 0 = inner, else number of prim refs
};

struct Geometry {
    std::atomic<int> refs{1};
    RTCGeometryType type;
    std::vector<char> vertices; size_t vstride = 0, vcount = 0;
    std::vector<char> indices;  size_t istride = 0, icount = 0;
    std::vector<RTCGrid> grids;
    void* user = nullptr;
};

struct Device { std::atomic<int> refs{1}; };

bool g_force_brute = false;

struct Scene {
    std::atomic<int> refs{1};
    std::vector<Geometry*> geoms;
    std::vector<Prim> prims;
    std::vector<uint32_t> refs_sorted;  // prim indices in BVH leaf order
    std::vector<Node> nodes;
    bool use_bvh = false;

    void build();
    void build_node(uint32_t node, uint32_t begin, uint32_t end, std::vector<Box>& boxes, std::vector<F3>& cent);
    void intersect(RTCRayHit* rh) const;
};

inline F3 load3(const Geometry* g, size_t i) {
    const float* p = reinterpret_cast<const float*>(g->vertices.data() + i * g->vstride);
    return {p[0], p[1], p[2]};
}

void Scene::build() {
    prims.clear();
    for (uint32_t gid = 0; gid < geoms.size(); gid++) {
        const Geometry* g = geoms[gid];
        if (g->type == RTC_GEOMETRY_TYPE_TRIANGLE || g->type == RTC_GEOMETRY_TYPE_QUAD) {
            int nv = g->type == RTC_GEOMETRY_TYPE_TRIANGLE ? 3 : 4;
            for (size_t i = 0; i < g->icount; i++) {
                const unsigned* idx = reinterpret_cast<const unsigned*>(g->indices.data() + i * g->istride);
                Prim p{};
                p.kind = nv == 3 ? PK_TRI : PK_QUAD;
                for (int k = 0; k < nv; k++) p.v[k] = load3(g, idx[k]);
                p.geomID = gid; p.primID = uint32_t(i);
                prims.push_back(p);
            }
        } else if (g->type == RTC_GEOMETRY_TYPE_SPHERE_POINT) {
            for (size_t i = 0; i < g->vcount; i++) {
                const float* v = reinterpret_cast<const float*>(g->vertices.data() + i * g->vstride);
                Prim p{};
                p.kind = PK_SPHERE;
                p.v[0] = {v[0], v[1], v[2]}; p.radius = v[3];
                p.geomID = gid; p.primID = uint32_t(i);
                prims.push_back(p);
            }
        } else if (g->type == RTC_GEOMETRY_TYPE_GRID) {
            for (size_t gi = 0; gi < g->grids.size(); gi++) {
                const RTCGrid& gr = g->grids[gi];
                for (uint32_t y = 0; y + 1 < gr.height; y++) {
                    for (uint32_t x = 0; x + 1 < gr.width; x++) {
                        Prim p{};
                        p.kind = PK_GRIDCELL;
                        size_t base = gr.startVertexID + size_t(y) * gr.stride + x;
                        p.v[0] = load3(g, base);
                        p.v[1] = load3(g, base + 1);
                        p.v[2] = load3(g, base + gr.stride + 1);
                        p.v[3] = load3(g, base + gr.stride);
                        p.geomID = gid; p.primID = uint32_t(gi);
                        p.cx = x; p.cy = y; p.gw = gr.width - 1u; p.gh = gr.height - 1u;
                        prims.push_back(p);
                    }
                }
            }
        }
    }
    use_bvh = !g_force_brute && prims.size() > 16;
    nodes.clear();
    refs_sorted.clear();
    if (!use_bvh) return;
    std::vector<Box> boxes(prims.size());
    std::vector<F3> cent(prims.size());
    refs_sorted.resize(prims.size());
    for (size_t i = 0; i < prims.size(); i++) {
        boxes[i] = prim_box(prims[i]);
        cent[i] = {0.5f * (boxes[i].lo[0] + boxes[i].hi[0]), 0.5f * (boxes[i].lo[1] + boxes[i].hi[1]),
                   0.5f * (boxes[i].lo[2] + boxes[i].hi[2])};
        refs_sorted[i] = uint32_t(i);
    }
    nodes.reserve(prims.size());
    nodes.push_back(Node{});
    build_node(0, 0, uint32_t(prims.size()), boxes, cent);
}

inline void grow(Box& b, const Box& o) {
    for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], o.lo[a]); b.hi[a] = std::max(b.hi[a], o.hi[a]); }
}
inline Box empty_box() {
    Box b;
    for (int a = 0; a < 3; a++) { b.lo[a] = std::numeric_limits<float>::infinity(); b.hi[a] = -b.lo[a]; }
    return b;
}
inline float half_area(const Box& b) {
    float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}

// binned-SAH top-down build (16 bins); only the tree SHAPE depends on it, never the hit result
void Scene::build_node(uint32_t ni, uint32_t begin, uint32_t end, std::vector<Box>& boxes, std::vector<F3>& cent) {
    Box nb = empty_box();
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (uint32_t i = begin; i < end; i++) {
        uint32_t r = refs_sorted[i];
        grow(nb, boxes[r]);
        const float c[3] = {cent[r].x, cent[r].y, cent[r].z};
        for (int a = 0; a < 3; a++) { clo[a] = std::min(clo[a], c[a]); chi[a] = std::max(chi[a], c[a]); }
    }
    nodes[ni].box = nb;
    uint32_t n = end - begin;
    if (n <= 4) { nodes[ni].left = begin; nodes[ni].count = n; return; }

    const int NB = 16;
    int best_axis = -1, best_split = -1;
    float best_cost = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) {
        float ext = chi[a] - clo[a];
        if (!(ext > 0.0f)) continue;
        Box bb[NB]; uint32_t bc[NB];
        for (int b = 0; b < NB; b++) { bb[b] = empty_box(); bc[b] = 0; }
        float scale = float(NB) / ext;
        for (uint32_t i = begin; i < end; i++) {
            uint32_t r = refs_sorted[i];
            float c = a == 0 ? cent[r].x : (a == 1 ? cent[r].y : cent[r].z);
            int b = std::min(NB - 1, std::max(0, int((c - clo[a]) * scale)));
            grow(bb[b], boxes[r]); bc[b]++;
        }
        float la[NB]; uint32_t lc[NB];
        Box acc = empty_box(); uint32_t cnt = 0;
        for (int b = 0; b < NB; b++) { grow(acc, bb[b]); cnt += bc[b]; la[b] = cnt ? half_area(acc) : 0.0f; lc[b] = cnt; }
        acc = empty_box(); cnt = 0;
        for (int b = NB - 1; b > 0; b--) {
            grow(acc, bb[b]); cnt += bc[b];
            if (cnt == 0 || lc[b - 1] == 0) continue;
            float cost = la[b - 1] * float(lc[b - 1]) + half_area(acc) * float(cnt);
            if (cost < best_cost) { best_cost = cost; best_axis = a; best_split = b; }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        mid = begin + n / 2;  // all centroids coincide
    } else {
        int a = best_axis;
        float scale = float(NB) / (chi[a] - clo[a]);
        auto it = std::partition(refs_sorted.begin() + begin, refs_sorted.begin() + end, [&](uint32_t r) {
            float c = a == 0 ? cent[r].x : (a == 1 ? cent[r].y : cent[r].z);
            int b = std::min(NB - 1, std::max(0, int((c - clo[a]) * scale)));
            return b < best_split;
        });
        mid = uint32_t(it - refs_sorted.begin());
        if (mid == begin || mid == end) mid = begin + n / 2;
    }
    uint32_t left = uint32_t(nodes.size());
    nodes.push_back(Node{});
    nodes.push_back(Node{});
    nodes[ni].left = left; nodes[ni].count = 0;
    build_node(left, begin, mid, boxes, cent);
    build_node(left + 1, mid, end, boxes, cent);
}

// slab test: true when [tnear, tbest] overlaps the box; entry distance in `entry`.
// Conservative: the interval is widened by a few ulps.
inline bool box_hit(const Box& b, const float o[3], const float inv[3], float tnear, float tbest, float& entry) {
    float t0 = tnear, t1 = tbest;
    for (int a = 0; a < 3; a++) {
        float ta = (b.lo[a] - o[a]) * inv[a];
        float tb = (b.hi[a] - o[a]) * inv[a];
        float tmn = std::fmin(ta, tb), tmx = std::fmax(ta, tb);  // fmin/fmax drop NaN (0*inf)
        tmn = tmn * (tmn >= 0.0f ? 0.9999995f : 1.0000005f);
        tmx = tmx * (tmx >= 0.0f ? 1.0000005f : 0.9999995f);
        t0 = std::fmax(t0, tmn);
        t1 = std::fmin(t1, tmx);
    }
    entry = t0;
    return t0 <= t1;
}

thread_local unsigned long long t_rays = 0;
std::atomic<unsigned long long> g_rays_retired{0};
std::mutex g_live_mutex;
std::vector<unsigned long long*> g_live;  // counters of threads still alive

struct ThreadCounterGuard {
    ThreadCounterGuard() { std::lock_guard<std::mutex> l(g_live_mutex); g_live.push_back(&t_rays); }
    ~ThreadCounterGuard() {
        std::lock_guard<std::mutex> l(g_live_mutex);
        g_rays_retired += t_rays;
        g_live.erase(std::find(g_live.begin(), g_live.end(), &t_rays));
    }
};

void Scene::intersect(RTCRayHit* rh) const {
    static thread_local ThreadCounterGuard guard;
    t_rays++;
    F3 O{rh->ray.org_x, rh->ray.org_y, rh->ray.org_z};
    F3 D{rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z};
    float tnear = rh->ray.tnear, tfar = rh->ray.tfar;
    Hit best;
    if (!use_bvh) {
        for (uint32_t i = 0; i < prims.size(); i++) prim_test(prims[i], i, O, D, tnear, tfar, best);
    } else {
        const float o[3] = {O.x, O.y, O.z};
        const float inv[3] = {1.0f / D.x, 1.0f / D.y, 1.0f / D.z};
        uint32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            // ties (entry == best.t) must still be visited: a lower-index primitive may sit there
            float limit = std::fmin(tfar, best.t);
            float dn;
            if (!box_hit(n.box, o, inv, tnear, limit, dn)) continue;
            if (n.count) {
                for (uint32_t i = 0; i < n.count; i++) {
                    uint32_t r = refs_sorted[n.left + i];
                    prim_test(prims[r], r, O, D, tnear, tfar, best);
                }
            } else {
                float dl, dr;
                bool hl = box_hit(nodes[n.left].box, o, inv, tnear, limit, dl);
                bool hr = box_hit(nodes[n.left + 1].box, o, inv, tnear, limit, dr);
                // push the far child first
                if (hl && hr) {
                    if (dl <= dr) { stack[sp++] = n.left + 1; stack[sp++] = n.left; }
                    else { stack[sp++] = n.left; stack[sp++] = n.left + 1; }
                } else if (hl) stack[sp++] = n.left;
                else if (hr) stack[sp++] = n.left + 1;
            }
        }
    }
    if (best.prim != 0xffffffffu) {
        const Prim& p = prims[best.prim];
        rh->ray.tfar = best.t;
        rh->hit.u = best.u; rh->hit.v = best.v;
        rh->hit.Ng_x = best.ng.x; rh->hit.Ng_y = best.ng.y; rh->hit.Ng_z = best.ng.z;
        rh->hit.geomID = p.geomID; rh->hit.primID = p.primID;
        rh->hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
    }
}

}  // namespace

struct RTCDeviceTy : Device {};
struct RTCSceneTy : Scene {};
struct RTCGeometryTy : Geometry {};

extern "C" {

RTCDevice rtcNewDevice(const char*) { return new RTCDeviceTy(); }
void rtcReleaseDevice(RTCDevice d) { if (d && --d->refs == 0) delete d; }
RTCError rtcGetDeviceError(RTCDevice) { return RTC_ERROR_NONE; }
void rtcSetDeviceErrorFunction(RTCDevice, RTCErrorFunction, void*) {}

RTCScene rtcNewScene(RTCDevice) { return new RTCSceneTy(); }
void rtcReleaseScene(RTCScene s) {
    if (s && --s->refs == 0) {
        for (Geometry* g : s->geoms) rtcReleaseGeometry(static_cast<RTCGeometry>(g));
        delete s;
    }
}
void rtcCommitScene(RTCScene s) { s->build(); }
void* rtcGetGeometryUserDataFromScene(RTCScene s, unsigned int geomID) {
    return geomID < s->geoms.size() ? s->geoms[geomID]->user : nullptr;
}

RTCGeometry rtcNewGeometry(RTCDevice, RTCGeometryType type) {
    RTCGeometry g = new RTCGeometryTy();
    g->type = type;
    return g;
}
void* rtcSetNewGeometryBuffer(RTCGeometry g, RTCBufferType type, unsigned int, RTCFormat, size_t byteStride, size_t itemCount) {
    if (type == RTC_BUFFER_TYPE_VERTEX) {
        g->vertices.assign(byteStride * itemCount + 16, 0);  // Embree pads vertex buffers for 16-byte loads
        g->vstride = byteStride; g->vcount = itemCount;
        return g->vertices.data();
    }
    if (type == RTC_BUFFER_TYPE_INDEX) {
        g->indices.assign(byteStride * itemCount, 0);
        g->istride = byteStride; g->icount = itemCount;
        return g->indices.data();
    }
    if (type == RTC_BUFFER_TYPE_GRID) {
        g->grids.assign(itemCount, RTCGrid{});
        return g->grids.data();
    }
    return nullptr;
}
void rtcSetGeometryUserData(RTCGeometry g, void* p) { g->user = p; }
void rtcCommitGeometry(RTCGeometry) {}
unsigned int rtcAttachGeometry(RTCScene s, RTCGeometry g) {
    g->refs++;
    s->geoms.push_back(g);
    return unsigned(s->geoms.size() - 1);
}
void rtcReleaseGeometry(RTCGeometry g) { if (g && --g->refs == 0) delete g; }

void rtcIntersect1(RTCScene s, RTCRayHit* rh, RTCIntersectArguments*) { s->intersect(rh); }

unsigned long long rtcShimRayCount(void) {
    std::lock_guard<std::mutex> l(g_live_mutex);
    unsigned long long n = g_rays_retired;
    for (auto* p : g_live) n += *p;
    return n;
}
void rtcShimResetRayCount(void) {
    std::lock_guard<std::mutex> l(g_live_mutex);
    g_rays_retired = 0;
    for (auto* p : g_live) *p = 0;
}
unsigned long long rtcShimThreadRayCount(void) { return t_rays; }
void rtcShimForceBruteForce(int on) { g_force_brute = on != 0; }

}  // extern "C"
