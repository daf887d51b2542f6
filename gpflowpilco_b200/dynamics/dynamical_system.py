"""DynamicalSystem — upstream gpflow_pilco/dynamics/dynamical_system.py:17-90 (minus tf.function compilation)."""
from __future__ import annotations

from typing import Any, Callable

from gpflowpilco_b200.dynamics.forward_sde import forward_sde
from gpflowpilco_b200.dynamics.solvers import Euler


class DynamicalSystem:
  def __init__(self, drift: Callable, diffusion: Callable = None, policy: Callable = None, encoder: Callable = None,
               solver: Callable = None):
    self.drift = drift
    self.diffusion = diffusion
    self.policy = policy
    self.encoder = encoder
    self.solver = Euler() if solver is None else solver

  def forward(self, t: Any, x: Any) -> Any:
    return forward_sde(x, self.drift, self.diffusion, self.policy, self.encoder)

  def solve_forward(self, initial_time, initial_state, solution_times, **kwargs) -> Any:
    return self.solver(func=self.forward, initial_time=initial_time, initial_state=initial_state,
                       solution_times=solution_times, **kwargs)

  def solve_forward_closure(self, initial_time, state_initializer: Callable, solution_times, compile: bool = True, **kwargs):
    def closure(state_initializer=state_initializer):
      return self.solve_forward(initial_time=initial_time, initial_state=state_initializer(), solution_times=solution_times, **kwargs)
    return closure
