"""meshlesshydro_b200 -- B200-native MFV hot path of jammartin/meshlessHydro behind a C ABI.

csrc/       hand-written sm_100a CUDA kernels + the C-ABI layer (include/mlh_gpu.h)
capi.py     ctypes binding of that ABI (no compute, no fallback)
ic.py       synthetic initial conditions of the reference's test cases
host/       C++ mirror of the reference's Domain/Particles/MeshlessScheme/Riemann surface
"""
from . import ic  # noqa: F401

__all__ = ["ic"]
