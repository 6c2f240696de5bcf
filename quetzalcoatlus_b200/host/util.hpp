// Host mirror of the reference's util.hpp:3-5 (ONE_MINUS_EPS, lerp).
#pragma once

const float ONE_MINUS_EPS = float(0x1.fffffep-1);

inline float lerp(float a, float b, float t) { return a + t * (b - a); }
