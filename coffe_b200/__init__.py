"""coffe_b200: B200-native evaluation of COFFE's signal path (2PCF, multipoles, redshift-averaged
multipoles, flat-sky covariance) behind COFFE's own API.  All arithmetic runs in hand-written sm_100a
CUDA kernels inside coffe_b200/csrc/libcoffe_b200.so, reached through the C ABI of include/coffe_b200.h."""
from . import capi  # noqa: F401
from .coffe import Coffe  # noqa: F401
from .engine import CoffeB200Error, Engine, Group  # noqa: F401
from .representation import AverageCovariance, AverageMultipoles, Corrfunc, Covariance, Multipoles  # noqa: F401

__all__ = ["capi", "Coffe", "Engine", "Group", "CoffeB200Error", "Corrfunc", "Multipoles", "AverageMultipoles", "Covariance", "AverageCovariance"]
