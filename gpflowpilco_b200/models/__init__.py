from gpflowpilco_b200.models.core import *                    # noqa: F401,F403
from gpflowpilco_b200.models.core import (GPR, SVGP, BijectorChain, Constant, Gaussian, GPModelWrapper, InducingPoints,
                                          InverseLinkWrapper, KernelRegressor, LinearCoregionalization, NormalCDF, Scale,
                                          SeparateIndependent, SeparateIndependentInducingVariables, SharedIndependent,
                                          SharedIndependentInducingVariables, Shift, SquaredExponential, Zero)
