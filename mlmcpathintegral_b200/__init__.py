"""mlmcpathintegral_b200 -- B200-native sampler inner loop of eikehmueller/mlmcpathintegral.

Everything computes inside libmlmcpi.so (hand-written sm_100a CUDA kernels behind the
C-ABI of include/mlmcpi.h).  Importing the package loads that library and raises if it
is missing: there is no CPU path."""
from . import _lib  # noqa: F401  (raises ImportError when libmlmcpi.so is absent)
from .api import *  # noqa: F401,F403
from .api import Context, MlmcpiError, MultilevelMC, Sampler, Statistics  # noqa: F401
