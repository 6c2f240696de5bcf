// Ray / primitive intersection.  The reference delegates this to Embree's rtcIntersect1
// (scene.cpp:41-59); the arithmetic is the one the oracle's Embree shim defines
// (oracle/embree_shim/rtcore_shim.cpp, restating Embree 4.3's Moeller-Trumbore triangle /
// quad / grid and sphere-point intersectors) and is repeated here operation for operation so
// that t, u, v and Ng are bit-identical.  Closest hit = minimum t, ties to the lower
// primitive index (geomID, primID, cell order), which makes the result independent of the
// BVH traversal order.
#pragma once

#include "scene.cuh"

namespace qz {

struct Hit {
    float t, u, v;
    uint32_t prim;     // slot in the (leaf-ordered) primitive array; 0xffffffff = miss
    V3 ng;             // unnormalised geometric normal as Embree reports it
    uint32_t key;      // tie-break key: primitive index in (geomID, primID, cell) order
    uint32_t geom_id;  // RTCHit::geomID
    uint32_t prim_id;  // RTCHit::primID
};

#define QZ_NO_HIT 0xffffffffu
#define QZ_TNEAR 0.0001f /* scene.cpp:49; in ray-parameter units */

QZ_HD float xor_sign(float v, uint32_t s) { return u32_as_float(float_as_u32(v) ^ s); }

struct PrimHit {
    float t, u, v;
    V3 ng;
};

// Moeller-Trumbore, Embree operation order, on a triangle given by v0 and its precomputed
// e1 = v0 - v1, e2 = v2 - v0, ng = e2 x e1; `flip` = second triangle of a quad
QZ_HD bool tri_test_pre(V3 O, V3 D, float tnear, float tfar, V3 v0, V3 e1, V3 e2, V3 ng, bool flip, PrimHit& h) {
    V3 C = v0 - O;
    V3 R = cross(C, D);
    float den = dot(ng, D);
    float absDen = fabsf(den);
    uint32_t s = float_as_u32(den) & 0x80000000u;
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

QZ_HD bool tri_test(V3 O, V3 D, float tnear, float tfar, V3 v0, V3 v1, V3 v2, bool flip, PrimHit& h) {
    V3 e1 = v0 - v1;
    V3 e2 = v2 - v0;
    V3 ng = cross(e2, e1);
    return tri_test_pre(O, D, tnear, tfar, v0, e1, e2, ng, flip, h);
}

// rd2 = 1 / dot(D, D) depends on the ray only: callers that test many spheres per ray pass it in
QZ_HD bool sphere_test_rd2(V3 O, V3 D, float rd2, float tnear, float tfar, V3 c, float r, PrimHit& h) {
    V3 c0 = c - O;
    float projC0 = dot(c0, D) * rd2;
    V3 perp = c0 - D * projC0;
    float l2 = dot(perp, perp);
    float r2 = r * r;
    if (!(l2 <= r2)) return false;
    float td = sqrtf((r2 - l2) * rd2);
    float t_front = projC0 - td;
    float t_back = projC0 + td;
    if (tnear <= t_front && t_front <= tfar) {
        h.t = t_front;
        h.ng = D * (-td) - perp;
    } else if (tnear <= t_back && t_back <= tfar) {
        h.t = t_back;
        h.ng = D * td - perp;
    } else {
        return false;
    }
    h.u = 0.0f;
    h.v = 0.0f;
    return true;
}

QZ_HD bool sphere_test(V3 O, V3 D, float tnear, float tfar, V3 c, float r, PrimHit& h) {
    float rd2 = 1.0f / dot(D, D);
    return sphere_test_rd2(O, D, rd2, tnear, tfar, c, r, h);
}

QZ_HD V3 xyz(const F4& f) { return v3(f.x, f.y, f.z); }

QZ_HD F4 load_f4(const F4* p) {
#if defined(__CUDA_ARCH__)
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r;
#else
    return *p;
#endif
}

// word 2 of a primitive record: kind in the low 2 bits, tie-break key above
QZ_HD uint32_t prim_kind(uint32_t w2) { return w2 & 3u; }
QZ_HD uint32_t prim_key(uint32_t w2) { return w2 >> 2; }

// Tests primitive slot `slot`; updates `best` under the (t, key) order.  tfar is the ray's
// own far limit (infinity for path rays): validity never depends on the running best.
QZ_HD void prim_test_rec(const DScene& sc, const F4& a, const F4& b, const F4& c, const F4& d, uint32_t slot, V3 O, V3 D,
                         float tnear, float tfar, Hit& best) {
    const uint32_t w2 = float_as_u32(c.w);
    const uint32_t kind = prim_kind(w2);
    PrimHit h;
    bool found = false;
    if (kind == QZ_PRIM_SPHERE) {
        found = sphere_test(O, D, tnear, tfar, xyz(a), b.x, h);
    } else if (kind == QZ_PRIM_TRIANGLE) {
        found = tri_test(O, D, tnear, tfar, xyz(a), xyz(b), xyz(c), false, h);
    } else {
        PrimHit ha, hb;
        bool fa = tri_test(O, D, tnear, tfar, xyz(a), xyz(b), xyz(d), false, ha);
        // OBJ triangles arrive as quads with the last vertex repeated (obj.cpp:41-115): the second
        // triangle (c, d, b) then has a zero edge, hence a zero normal and den == 0 -- rejected by
        // tri_test for every ray, so it is skipped without changing any result
        const bool degenerate = c.x == d.x && c.y == d.y && c.z == d.z;
        bool fb = !degenerate && tri_test(O, D, tnear, tfar, xyz(c), xyz(d), xyz(b), true, hb);
        if (fa && (!fb || ha.t <= hb.t)) { h = ha; found = true; }
        else if (fb) { h = hb; found = true; }
        if (found && kind == QZ_PRIM_GRIDCELL) {
            const uint32_t cell = float_as_u32(d.w);
            const uint32_t dims = sc.grid_dims[float_as_u32(a.w)];
            h.u = ((float)(cell & 0xffffu) + h.u) / (float)(dims & 0xffffu);
            h.v = ((float)(cell >> 16) + h.v) / (float)(dims >> 16);
        }
    }
    if (!found) return;
    const uint32_t key = prim_key(w2);
    if (h.t < best.t || (h.t == best.t && key < best.key)) {
        best.t = h.t; best.u = h.u; best.v = h.v; best.ng = h.ng; best.prim = slot; best.key = key;
        best.geom_id = float_as_u32(a.w); best.prim_id = float_as_u32(b.w);
    }
}

// The flat kernels (scenes of a few dozen primitives, every ray tests every primitive) keep the
// ray-independent part of Moeller-Trumbore -- the two edges and the normal of each triangle --
// next to the vertices: 8 x 16 bytes per primitive, built once per CTA by the same expressions
// tri_test uses, so every result is unchanged.
struct FlatPrim {
    F4 a;        // v0 of triangle A (sphere: centre) | geomID
    F4 b;        // (sphere: radius in x)             | primID
    F4 e1a, e2a, nga;   // triangle A = (a, b, d) of a quad / (a, b, c) of a triangle; e1a.w = kind | key << 2, e2a.w = grid cell
    F4 c;        // v0 of triangle B = (c, d, b) of a quad; c.w = 1 when triangle B exists
    F4 e1b, e2b; // ngb.xyz is stored in (nga.w, e1b.w, e2b.w)
};

QZ_HD FlatPrim make_flat_prim(const F4& a, const F4& b, const F4& c, const F4& d) {
    FlatPrim f;
    const uint32_t w2 = float_as_u32(c.w);
    const uint32_t kind = prim_kind(w2);
    f.a = a; f.b = b; f.c = c;
    f.c.w = 0.0f;
    V3 e1 = v3(0.0f, 0.0f, 0.0f), e2 = e1, ng = e1, e1b = e1, e2b = e1, ngb = e1;
    if (kind == QZ_PRIM_TRIANGLE) {
        e1 = xyz(a) - xyz(b); e2 = xyz(c) - xyz(a); ng = cross(e2, e1);
    } else if (kind != QZ_PRIM_SPHERE) {
        e1 = xyz(a) - xyz(b); e2 = xyz(d) - xyz(a); ng = cross(e2, e1);                 // (a, b, d)
        const bool degenerate = c.x == d.x && c.y == d.y && c.z == d.z;                  // see prim_test_rec
        if (!degenerate) {
            e1b = xyz(c) - xyz(d); e2b = xyz(b) - xyz(c); ngb = cross(e2b, e1b);         // (c, d, b)
            f.c.w = 1.0f;
        }
    }
    f.e1a.x = e1.x; f.e1a.y = e1.y; f.e1a.z = e1.z; f.e1a.w = c.w;
    f.e2a.x = e2.x; f.e2a.y = e2.y; f.e2a.z = e2.z; f.e2a.w = d.w;
    f.nga.x = ng.x; f.nga.y = ng.y; f.nga.z = ng.z; f.nga.w = ngb.x;
    f.e1b.x = e1b.x; f.e1b.y = e1b.y; f.e1b.z = e1b.z; f.e1b.w = ngb.y;
    f.e2b.x = e2b.x; f.e2b.y = e2b.y; f.e2b.z = e2b.z; f.e2b.w = ngb.z;
    return f;
}

// prim_test_rec on a FlatPrim
QZ_HD void flat_prim_test(const DScene& sc, const FlatPrim& f, uint32_t slot, V3 O, V3 D, float rd2, float tnear, float tfar, Hit& best) {
    const uint32_t w2 = float_as_u32(f.e1a.w);
    const uint32_t kind = prim_kind(w2);
    PrimHit h;
    bool found = false;
    if (kind == QZ_PRIM_SPHERE) {
        found = sphere_test_rd2(O, D, rd2, tnear, tfar, xyz(f.a), f.b.x, h);
    } else if (kind == QZ_PRIM_TRIANGLE) {
        found = tri_test_pre(O, D, tnear, tfar, xyz(f.a), xyz(f.e1a), xyz(f.e2a), xyz(f.nga), false, h);
    } else {
        PrimHit ha, hb;
        bool fa = tri_test_pre(O, D, tnear, tfar, xyz(f.a), xyz(f.e1a), xyz(f.e2a), xyz(f.nga), false, ha);
        bool fb = f.c.w != 0.0f &&
                  tri_test_pre(O, D, tnear, tfar, xyz(f.c), xyz(f.e1b), xyz(f.e2b), v3(f.nga.w, f.e1b.w, f.e2b.w), true, hb);
        if (fa && (!fb || ha.t <= hb.t)) { h = ha; found = true; }
        else if (fb) { h = hb; found = true; }
        if (found && kind == QZ_PRIM_GRIDCELL) {
            const uint32_t cell = float_as_u32(f.e2a.w);
            const uint32_t dims = sc.grid_dims[float_as_u32(f.a.w)];
            h.u = ((float)(cell & 0xffffu) + h.u) / (float)(dims & 0xffffu);
            h.v = ((float)(cell >> 16) + h.v) / (float)(dims >> 16);
        }
    }
    if (!found) return;
    const uint32_t key = prim_key(w2);
    if (h.t < best.t || (h.t == best.t && key < best.key)) {
        best.t = h.t; best.u = h.u; best.v = h.v; best.ng = h.ng; best.prim = slot; best.key = key;
        best.geom_id = float_as_u32(f.a.w); best.prim_id = float_as_u32(f.b.w);
    }
}

// LAZY variant for the flat kernel's closest-hit loop.  A hit costs three IEEE divisions (t, u, v) and a normal, but a ray
// keeps only its closest hit: per candidate only t is computed (the order needs it, rounded exactly as the reference's
// quotient), the barycentric numerators stay as they are, and u, v, Ng, the ids are produced ONCE per ray by
// flat_best_finish() with the very expressions of tri_test_pre / sphere_test_rd2 / flat_prim_test -- same bits.  In the
// lock-step loop the three divisions ran for the few lanes that had just hit something: a tenth of the kernel's
// instructions at 4-8 of 32 lanes (profiles/r02_source_k_step_flat.txt).
struct FlatBest {
    float t, U, V, absDen;
    uint32_t info;   // staged primitive index | 0x100 second triangle of the quad | 0x200 sphere; 0xffffffff = miss
    uint32_t key;
};
#define QZ_FB_FLIP 0x100u
#define QZ_FB_SPHERE 0x200u

// tri_test_pre up to the distance: true = valid hit, t = T / absDen
QZ_HD bool tri_test_lazy(V3 O, V3 D, float tnear, float tfar, V3 v0, V3 e1, V3 e2, V3 ng, float& t, float& Uo, float& Vo, float& absDen_o) {
    V3 C = v0 - O;
    V3 R = cross(C, D);
    float den = dot(ng, D);
    float absDen = fabsf(den);
    uint32_t s = float_as_u32(den) & 0x80000000u;
    float U = xor_sign(dot(R, e2), s);
    float V = xor_sign(dot(R, e1), s);
    if (!(den != 0.0f) || !(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= absDen)) return false;
    float T = xor_sign(dot(ng, C), s);
    if (!(absDen * tnear < T) || !(T <= absDen * tfar)) return false;
    t = T / absDen;
    Uo = U; Vo = V; absDen_o = absDen;
    return true;
}

QZ_HD void flat_prim_test_lazy(const FlatPrim& f, uint32_t slot, V3 O, V3 D, float rd2, float tnear, float tfar, FlatBest& best) {
    const uint32_t w2 = float_as_u32(f.e1a.w);
    const uint32_t kind = prim_kind(w2);
    float t = 0.0f, U = 0.0f, V = 0.0f, absDen = 0.0f;
    uint32_t info = slot;
    bool found = false;
    if (kind == QZ_PRIM_SPHERE) {
        PrimHit h;
        found = sphere_test_rd2(O, D, rd2, tnear, tfar, xyz(f.a), f.b.x, h);
        t = h.t;
        info |= QZ_FB_SPHERE;
    } else {
        float ta, Ua, Va, da, tb, Ub, Vb, db;
        const bool fa = tri_test_lazy(O, D, tnear, tfar, xyz(f.a), xyz(f.e1a), xyz(f.e2a), xyz(f.nga), ta, Ua, Va, da);
        const bool fb = kind != QZ_PRIM_TRIANGLE && f.c.w != 0.0f &&
                        tri_test_lazy(O, D, tnear, tfar, xyz(f.c), xyz(f.e1b), xyz(f.e2b), v3(f.nga.w, f.e1b.w, f.e2b.w), tb, Ub, Vb, db);
        if (fa && (!fb || ta <= tb)) { t = ta; U = Ua; V = Va; absDen = da; found = true; }
        else if (fb) { t = tb; U = Ub; V = Vb; absDen = db; info |= QZ_FB_FLIP; found = true; }
    }
    if (!found) return;
    const uint32_t key = prim_key(w2);
    if (t < best.t || (t == best.t && key < best.key)) {
        best.t = t; best.U = U; best.V = V; best.absDen = absDen; best.info = info; best.key = key;
    }
}

// the full hit record of the winner (prims = the staged FlatPrim list the loop ran over)
QZ_HD void flat_best_finish(const DScene& sc, const FlatPrim* prims, const FlatBest& fb, V3 O, V3 D, float rd2, float tnear, float tfar, Hit& out) {
    out.t = INFINITY; out.u = 0.0f; out.v = 0.0f; out.prim = QZ_NO_HIT; out.key = 0xffffffffu;
    out.ng = v3(0.0f, 0.0f, 0.0f); out.geom_id = QZ_NO_HIT; out.prim_id = 0;
    if (fb.info == 0xffffffffu) return;
    const uint32_t p = fb.info & 0xffu;
    const FlatPrim& f = prims[p];
    out.t = fb.t; out.prim = p; out.key = fb.key;
    out.geom_id = float_as_u32(f.a.w); out.prim_id = float_as_u32(f.b.w);
    if (fb.info & QZ_FB_SPHERE) {
        PrimHit h;
        sphere_test_rd2(O, D, rd2, tnear, tfar, xyz(f.a), f.b.x, h);   // the same call that found it: same t, and its normal
        out.ng = h.ng;
        return;
    }
    const bool flip = (fb.info & QZ_FB_FLIP) != 0u;
    if (flip) {
        out.u = (fb.absDen - fb.U) / fb.absDen;
        out.v = (fb.absDen - fb.V) / fb.absDen;
        out.ng = v3(f.nga.w, f.e1b.w, f.e2b.w);
    } else {
        out.u = fb.U / fb.absDen;
        out.v = fb.V / fb.absDen;
        out.ng = xyz(f.nga);
    }
    if (prim_kind(float_as_u32(f.e1a.w)) == QZ_PRIM_GRIDCELL) {
        const uint32_t cell = float_as_u32(f.e2a.w);
        const uint32_t dims = sc.grid_dims[float_as_u32(f.a.w)];
        out.u = ((float)(cell & 0xffffu) + out.u) / (float)(dims & 0xffffu);
        out.v = ((float)(cell >> 16) + out.v) / (float)(dims >> 16);
    }
}

QZ_HD void prim_test(const DScene& sc, uint32_t slot, V3 O, V3 D, float tnear, float tfar, Hit& best) {
    const F4* rec = sc.prims + (size_t)slot * 4;
    const F4 a = load_f4(rec), b = load_f4(rec + 1), c = load_f4(rec + 2), d = load_f4(rec + 3);
    prim_test_rec(sc, a, b, c, d, slot, O, D, tnear, tfar, best);
}

}  // namespace qz
