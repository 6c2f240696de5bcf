// render() -- the hot-path entry (reference render.hpp:8-13, render.cpp:321-397).
// Same signature and blocking behaviour; the work runs on the B200 through qz_render().
#pragma once

#include "camera.hpp"
#include "image.hpp"
#include "scene.hpp"
#include "vec.hpp"

RenderResult render(const Camera& camera, const Scene& world, size_t n_samples, size_t max_bounces);

namespace qzhost {
// B200 additions used by the harness, the bench and the multi-GPU driver
qz_camera flatten_camera(const Camera& camera, std::vector<float>& sensor_storage);
// stats of the most recent render() on this thread
const qz_stats& last_render_stats();
// scene-build times on this thread (secondary metrics): the OBJ text parses of Scene::add_obj since the last
// call of reset_build_times(), and the whole of the last Scene::commit() (flatten + upload + GPU BVH build)
struct BuildTimes { double obj_parse_ms = 0.0, commit_ms = 0.0, bvh_build_ms = 0.0; };
const BuildTimes& build_times();
void reset_build_times();
}  // namespace qzhost
