// Scene flattening: turns the open polymorphic object graph the reference API exposes
// (Material / Texture / Spectrum / Light have no type tags -- SURVEY.md section 8.b) into
// the POD tables of include/qz_b200.h.  Every host class implements
// `flatten(qzhost::Flattener&) const` and returns its table id; shared objects
// (shared_ptr<const Spectrum>, caller-owned const Material*) are de-duplicated by address.
#pragma once

#include <cstdint>
#include <unordered_map>
#include <vector>

#include "qz_b200.h"

namespace qzhost {

struct Flattener {
    std::vector<qz_spectrum> spectra;
    std::vector<qz_texture> textures;
    std::vector<qz_material> materials;
    std::vector<int32_t> mixed_children;
    std::vector<qz_light> lights;
    std::vector<qz_geometry> geometries;
    std::vector<qz_prim> prims;
    std::vector<float> pool;
    std::vector<float> normals;
    std::vector<int32_t> normal_indices;
    std::vector<uint32_t> grid_dims;

    std::unordered_map<const void*, int32_t> seen_spectra, seen_materials;

    int32_t add_spectrum(const void* key, const qz_spectrum& s) {
        spectra.push_back(s);
        int32_t id = int32_t(spectra.size()) - 1;
        if (key) seen_spectra[key] = id;
        return id;
    }
    int32_t find_spectrum(const void* key) const {
        auto it = seen_spectra.find(key);
        return it == seen_spectra.end() ? -1 : it->second;
    }
    uint32_t add_pool(const float* p, size_t n) {
        uint32_t off = uint32_t(pool.size());
        pool.insert(pool.end(), p, p + n);
        return off;
    }
};

}  // namespace qzhost
