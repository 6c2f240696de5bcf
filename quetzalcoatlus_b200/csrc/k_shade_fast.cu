// Shading stages with radiometric values in fused / approximate arithmetic (common.cuh, ARITHMETIC MODES:
// QZ_FAST == 1).  The namespace of the device code is renamed for this unit so that its kernels and functions are
// distinct symbols from the exact unit's.
#define QZ_FAST 1
#define QZL_MODE fast
#define qz qz_fast
#include "k_shade.inc.cuh"
