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
}  // namespace qzhost
