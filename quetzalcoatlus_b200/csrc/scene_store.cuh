// Commit path shared by the CUDA library and the test-only host emulation: validates the
// cross-references of a qz_scene_tables, copies the tables into executor memory, builds the
// wide BVH and fills the DScene view the kernels receive.
#pragma once

#include <cstring>
#include <string>
#include <vector>

#include "bvh_build.cuh"
#include "shading.cuh"

namespace qz {

template <class Exec>
struct SceneStore {
    DScene view{};
    BvhBuildResult bvh{};
    bool committed = false;
    std::vector<void*> owned;
    std::vector<int32_t> hot_spectra;   // spectra the shading code evaluates per bounce at the path's wavelengths (sampler.cuh, SAMPLE MEMO)

    template <class T>
    const T* put(Exec& ex, const T* src, size_t n) {
        T* d = ex.template alloc<T>(n ? n : 1);
        if (n) ex.upload(d, src, n);
        owned.push_back(d);
        return d;
    }

    void release(Exec& ex) {
        for (void* p : owned) ex.free(p);
        owned.clear();
        if (bvh.nodes) ex.free(bvh.nodes);
        if (bvh.prims) ex.free(bvh.prims);
        bvh = BvhBuildResult();
        committed = false;
    }

    // returns an empty string on success, else the reason the tables are invalid
    static std::string validate(const qz_scene_tables& t) {
        auto spec_ok = [&](int32_t id) { return id >= 0 && (uint32_t)id < t.n_spectra; };
        for (uint32_t i = 0; i < t.n_spectra; i++) {
            const qz_spectrum& s = t.spectra[i];
            if (s.kind > QZ_SPEC_BLACKBODY) return "spectrum kind out of range";
            if (s.kind == QZ_SPEC_DENSE && (uint64_t)s.offset + s.count > t.n_pool) return "dense spectrum outside the pool";
            if (s.kind == QZ_SPEC_PIECEWISE && (uint64_t)s.offset + 2ull * s.count + 2 > t.n_pool) return "piecewise spectrum outside the pool";
            if (s.kind == QZ_SPEC_RGB_ILLUMINANT && s.aux >= 0) {
                if (!spec_ok(s.aux)) return "illuminant id out of range";
                if (t.spectra[s.aux].kind == QZ_SPEC_RGB_ILLUMINANT) return "nested RGB illuminant spectra are not supported";
            }
        }
        for (uint32_t i = 0; i < t.n_textures; i++) {
            const qz_texture& x = t.textures[i];
            if (x.kind == QZ_TEX_SOLID && !spec_ok(x.a)) return "texture spectrum id out of range";
            if (x.kind == QZ_TEX_DUMMY && (!spec_ok(x.a) || !spec_ok(x.b))) return "texture spectrum id out of range";
            if (x.kind == QZ_TEX_IMAGE && ((uint64_t)x.offset + 3ull * x.width * x.height > t.n_pool || !x.width || !x.height))
                return "image texture outside the pool";
            if (x.kind > QZ_TEX_IMAGE) return "texture kind out of range";
        }
        for (uint32_t i = 0; i < t.n_materials; i++) {
            const qz_material& m = t.materials[i];
            switch (m.kind) {
                case QZ_MAT_DIFFUSE:
                    if (m.a < 0 || (uint32_t)m.a >= t.n_textures) return "material texture id out of range";
                    break;
                case QZ_MAT_CONDUCTOR:
                    if (!spec_ok(m.a) || !spec_ok(m.b)) return "material spectrum id out of range";
                    break;
                case QZ_MAT_DIELECTRIC:
                case QZ_MAT_THIN_DIELECTRIC:
                    if (!spec_ok(m.a)) return "material spectrum id out of range";
                    break;
                case QZ_MAT_MIXED:
                    if (m.count == 0 || m.a < 0 || (uint64_t)m.a + m.count > t.n_mixed_children) return "mixed material children out of range";
                    break;
                default:
                    return "material kind out of range";
            }
        }
        for (uint32_t i = 0; i < t.n_mixed_children; i++)
            if (t.mixed_children[i] < 0 || (uint32_t)t.mixed_children[i] >= t.n_materials) return "mixed child id out of range";
        for (uint32_t i = 0; i < t.n_lights; i++) {
            const qz_light& l = t.lights[i];
            if (l.kind != QZ_LIGHT_POINT && l.kind != QZ_LIGHT_AREA_QUAD && l.kind != QZ_LIGHT_AREA_SPHERE) return "light kind out of range";
            if (!spec_ok(l.spectrum)) return "light spectrum id out of range";
        }
        uint64_t prim_total = 0;
        for (uint32_t i = 0; i < t.n_geometries; i++) {
            const qz_geometry& g = t.geometries[i];
            if (g.material >= 0 && (uint32_t)g.material >= t.n_materials) return "geometry material id out of range";
            if (g.light >= 0 && (uint32_t)g.light >= t.n_lights) return "geometry light id out of range";
            if (g.first_prim != prim_total) return "geometry primitives are not contiguous in geomID order";
            prim_total += g.prim_count;
            if (g.normal_offset >= 0) {
                if (g.nindex_offset < 0 || (uint64_t)g.nindex_offset + g.prim_count > t.n_normal_indices) return "normal indices out of range";
                for (uint64_t k = 0; k < 4ull * g.prim_count; k++) {
                    int32_t ni = t.normal_indices[4ull * g.nindex_offset + k];
                    if (ni < 0 || (uint64_t)g.normal_offset + (uint64_t)ni >= t.n_normals) return "normal index out of range";
                }
            }
        }
        if (prim_total != t.n_prims) return "primitive count does not match the geometries";
        if (t.n_prims >= (1u << 30)) return "too many primitives";
        for (uint32_t i = 0; i < t.n_prims; i++) {
            uint32_t w0, w2;
            std::memcpy(&w0, &t.prims[i].v[0][3], 4);
            std::memcpy(&w2, &t.prims[i].v[2][3], 4);
            if (w0 >= t.n_geometries) return "primitive geomID out of range";
            if (w2 > QZ_PRIM_GRIDCELL) return "primitive kind out of range";
        }
        if (t.bg_spectrum >= 0 && !spec_ok(t.bg_spectrum)) return "background spectrum id out of range";
        if (!t.rgb2spec_z || !t.rgb2spec_coeffs) return "rgb2spec table missing";
        return std::string();
    }

    bool commit(Exec& ex, const qz_scene_tables& t, const SamplerDim* d_sampler_table, const float* d_rho_tab) {
        release(ex);
        view = DScene();
        // Piecewise-linear spectra get a 1-nm bucket table over [360, 831]: bucket b holds
        // lower_bound(knots, 360 + b), so a lookup starts at the right knot instead of bisecting
        // (spectra.cuh).  The table is only built where a forward scan from it provably lands on
        // std::lower_bound's answer: the knots must be partitioned for every integer threshold
        // (sorted tables are; so are the reference's tables with their "..., 916, 831" tails).
        std::vector<qz_spectrum> spectra(t.spectra, t.spectra + t.n_spectra);
        std::vector<uint8_t> accel;
        for (qz_spectrum& sp : spectra) {
            if (sp.kind != QZ_SPEC_PIECEWISE) continue;
            sp.aux = -1;
            if (sp.count == 0 || sp.count > 255) continue;
            const float* l = t.pool + sp.offset;
            uint8_t table[QZ_PW_BUCKETS];
            bool ok = true;
            for (uint32_t bkt = 0; bkt < QZ_PW_BUCKETS && ok; bkt++) {
                const float thr = 360.0f + (float)bkt;
                uint32_t first = 0;
                while (first < sp.count && l[first] < thr) first++;
                for (uint32_t j = first; j < sp.count; j++) if (l[j] < thr) ok = false;
                table[bkt] = (uint8_t)first;
            }
            if (!ok) continue;
            sp.aux = (int32_t)accel.size();
            accel.insert(accel.end(), table, table + QZ_PW_BUCKETS);
        }
        // lights' emission first, then the conductors' eta and k
        hot_spectra.clear();
        auto add_hot = [&](int32_t id) {
            if (id < 0 || hot_spectra.size() >= QZ_MEMO_MAX_HOT) return;
            for (int32_t h : hot_spectra) if (h == id) return;
            hot_spectra.push_back(id);
        };
        for (uint32_t i = 0; i < t.n_lights; i++) add_hot(t.lights[i].spectrum);
        for (uint32_t i = 0; i < t.n_materials; i++)
            if (t.materials[i].kind == QZ_MAT_CONDUCTOR) { add_hot(t.materials[i].a); add_hot(t.materials[i].b); }
        view.spectra = put(ex, spectra.data(), spectra.size());
        view.pw_accel = put(ex, accel.data(), accel.size());
        view.textures = put(ex, t.textures, t.n_textures);
        view.materials = put(ex, t.materials, t.n_materials);
        view.mixed_children = put(ex, t.mixed_children, t.n_mixed_children);
        view.lights = put(ex, t.lights, t.n_lights);
        view.geoms = put(ex, t.geometries, t.n_geometries);
        view.pool = put(ex, t.pool, t.n_pool);
        view.normals = put(ex, t.normals, (size_t)t.n_normals * 3);
        view.nidx = put(ex, t.normal_indices, (size_t)t.n_normal_indices * 4);
        view.grid_dims = put(ex, t.grid_dims, t.n_geometries);
        view.lut_z = put(ex, t.rgb2spec_z, 32);
        view.lut_coeffs = put(ex, t.rgb2spec_coeffs, (size_t)3 * 32 * 32 * 32 * 3);
        const F4* src = put(ex, reinterpret_cast<const F4*>(t.prims), (size_t)t.n_prims * 4);
        ex.mark("upload tables");
        if (!build_wide_bvh(ex, src, t.n_prims, bvh)) return false;
        ex.free(const_cast<F4*>(src));
        owned.pop_back();
        view.prims = bvh.prims;
        view.nodes = bvh.nodes;
        view.n_nodes = bvh.n_nodes;
        view.n_prims = t.n_prims;
        view.n_lights = t.n_lights;
        view.bg_spectrum = t.bg_spectrum;
        view.bg_scale = t.bg_scale;
        view.sampler_table = d_sampler_table;
        view.rho_tab = d_rho_tab;
        view.leaf_prims = nullptr;
        {
            uint32_t* flag = ex.template alloc<uint32_t>(1);
            ex.zero(flag, sizeof(uint32_t));
            owned.push_back(flag);
            view.overflow = flag;
        }
        committed = true;
        return true;
    }
};

}  // namespace qz
