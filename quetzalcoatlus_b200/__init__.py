"""quetzalcoatlus_b200 -- B200-native implementation of quetzalcoatlus's path-tracing hot path.

Layout: ``csrc/`` CUDA kernels + the C ABI (``include/qz_b200.h``), ``host/`` the C++ host
library mirroring the reference's scene API, ``scenes/`` the benchmark scene definitions,
``harness/`` C entry points for ctypes, ``data/`` measured spectra and the RGB->spectrum table.
"""
from .native import load_harness, load_cuda_library, native_paths  # noqa: F401
