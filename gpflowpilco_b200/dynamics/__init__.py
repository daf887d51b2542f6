from gpflowpilco_b200.dynamics.dynamical_system import DynamicalSystem      # noqa: F401
from gpflowpilco_b200.dynamics.forward_sde import forward_sde               # noqa: F401
from gpflowpilco_b200.dynamics.solvers import Euler, MomentMatchingEuler, foldl, scan   # noqa: F401
