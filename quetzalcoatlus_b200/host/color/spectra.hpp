// Named spectra (reference color/spectra.hpp:7-38).  The measured tables are DATA loaded
// from quetzalcoatlus_b200/data/spectra_tables.bin (see tools/gen_spectra_tables.py).
#pragma once

#include <memory>
#include <string>

#include "spectrum.hpp"

namespace spectra {

const float CIE_Y_INTEGRAL = 106.856895f;

std::shared_ptr<const DenselySampledSpectrum> X();
std::shared_ptr<const DenselySampledSpectrum> Y();
std::shared_ptr<const DenselySampledSpectrum> Z();

std::shared_ptr<const PiecewiseLinearSpectrum> ILLUM_D65();

std::shared_ptr<const PiecewiseLinearSpectrum> CANON_EOS_R();
std::shared_ptr<const PiecewiseLinearSpectrum> CANON_EOS_G();
std::shared_ptr<const PiecewiseLinearSpectrum> CANON_EOS_B();

std::shared_ptr<const PiecewiseLinearSpectrum> AL_IOR();
std::shared_ptr<const PiecewiseLinearSpectrum> AL_ABSORPTION();
std::shared_ptr<const PiecewiseLinearSpectrum> CU_IOR();
std::shared_ptr<const PiecewiseLinearSpectrum> CU_ABSORPTION();
std::shared_ptr<const PiecewiseLinearSpectrum> GLASS_BK7_IOR();
std::shared_ptr<const PiecewiseLinearSpectrum> GLASS_SF11_IOR();

}  // namespace spectra

namespace qzhost {
// directory holding spectra_tables.bin and coeffs_SRGB_32.dat: $QZ_DATA_DIR, else
// <dir of this shared library>/../data, else the current directory
const std::string& data_dir();
void set_data_dir(const std::string& dir);
}  // namespace qzhost
