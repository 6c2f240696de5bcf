// Pinhole camera (reference camera.hpp:11-35, camera.cpp:3-31).  Same public fields; the
// frame is computed on the host exactly as the reference does (tanf of half the field of
// view, right = look_at x up, unnormalised ray directions) and handed to the GPU as a
// qz_camera.  cast_ray() is kept for API parity; the render path generates rays on the
// device (csrc/shading.cuh: start_path, called by k_generate and the finish stages of csrc/wf_shade.cuh).
#pragma once

#include <cmath>
#include <cstddef>

#include "color/color.hpp"
#include "ray.hpp"
#include "transform.hpp"
#include "vec.hpp"

class Camera {
public:
    Camera(size_t image_width_, size_t image_height_, float fov, const Transform& transform = Transform::identity(),
           PixelSensor&& sensor_ = PixelSensor::CANON_EOS())
        : image_height(image_height_), image_width(image_width_), sensor(sensor_) {
        float viewport_height = 2.0f * tanf(fov * 0.5f);
        float viewport_width = viewport_height * float(image_width) / float(image_height);
        pos = transform * Pt3(0.0f, 0.0f, 0.0f);
        look_at = transform * Vec3(0.0f, 0.0f, -1.0f);
        up = transform * Vec3(0.0f, 1.0f, 0.0f);
        right = look_at.cross(up);
        Vec3 viewport_u = viewport_width * right;
        Vec3 viewport_v = viewport_height * up;
        pixel_delta_u = viewport_u / float(image_width);
        pixel_delta_v = viewport_v / float(image_height);
        viewport_bottom_left = pos + look_at - viewport_u * 0.5f - viewport_v * 0.5f;
    }

    Ray cast_ray(float u, float v) const { return Ray(pos, viewport_bottom_left + pixel_delta_u * u + pixel_delta_v * v - pos); }

    size_t image_height;
    size_t image_width;
    Pt3 pos;
    Vec3 look_at;
    Vec3 up;
    Vec3 right;
    Vec3 viewport_bottom_left;
    Vec3 pixel_delta_u;
    Vec3 pixel_delta_v;

    PixelSensor sensor;
};
