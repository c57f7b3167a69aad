"""pde_fluid_phi_b200 -- B200-native (sm_100a) spectral-convolution hot path of PDE-Fluid-Phi.

Drop-in mirrors of the reference's operator and model classes whose spectral layers run as
hand-written CUDA kernels behind a C ABI (include/pdephi_b200.h, lib/libpdephi_b200.so).
"""
from ._cabi import PdephiError, lib_path, load as load_library
from .operators import (RationalFourierLayer, SpectralConv3D, StabilityProjection, get_default_precision, get_grid,
                        set_default_precision)
from .models import FNO3D, RationalFNO, RationalFourierOperator3D, StabilityConstraints

__version__ = "0.1.0"
__all__ = ["RationalFourierLayer", "SpectralConv3D", "StabilityProjection", "RationalFourierOperator3D", "FNO3D", "RationalFNO",
           "StabilityConstraints", "get_grid", "set_default_precision", "get_default_precision", "load_library",
           "lib_path", "PdephiError"]
