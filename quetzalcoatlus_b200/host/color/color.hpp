// umbrella header (reference color/color.hpp)
#pragma once

#include "rgb.hpp"
#include "sensor.hpp"
#include "spectra.hpp"
#include "spectrum.hpp"
#include "spectrum_sample.hpp"
#include "xyz.hpp"
