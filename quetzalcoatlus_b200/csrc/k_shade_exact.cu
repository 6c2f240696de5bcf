// Shading stages in the reference's arithmetic (common.cuh, ARITHMETIC MODES: QZ_FAST == 0).
#define QZ_FAST 0
#define QZL_MODE exact
#include "k_shade.inc.cuh"
