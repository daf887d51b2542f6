"""gpflowpilco_b200 — B200-native (sm_100a) moment-matching / pathwise rollout path behind GPflowPILCO's API.

The compute lives in libgpp_b200.so (include/gpp_b200.h); this package is the host-side mirror of the reference
interface for that path (moment_matching, models, dynamics, loops, utils.kernel_expectation).  There is no CPU
implementation: importing works anywhere, computing needs the built library and a CUDA device.
"""
from gpflowpilco_b200 import models, moment_matching  # noqa: F401
from gpflowpilco_b200.moment_matching import GaussianMatch, GaussianMoments  # noqa: F401

__version__ = "0.1.0"
