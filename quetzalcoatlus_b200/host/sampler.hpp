// The reference's Sampler (sampler.hpp:22-78) is consumed only by the integrator; its
// Owen-scrambled Halton sequence is generated on the GPU (csrc/sampler.cuh, bit-exact).
// Host code that wants sample values for tests calls qz_sampler_eval() (include/qz_b200.h).
#pragma once
