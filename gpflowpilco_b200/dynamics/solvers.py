"""Euler / MomentMatchingEuler — same interface as upstream gpflow_pilco/dynamics/solvers.py:48-135 (step(func, t, dt, x) and
__call__(func, initial_time, initial_state, solution_times, callbacks_and_initializers, iterator)); the iterator is a plain
Python fold (upstream passes tf.scan / tf.foldl)."""
from __future__ import annotations

from typing import Any, Callable, List, Tuple

import numpy as np
import torch

from gpflowpilco_b200.moment_matching import GaussianMoments

__all__ = ("Euler", "MomentMatchingEuler", "foldl", "scan")


def foldl(fn, elems, initializer):
  state = initializer
  for t, dt in zip(*elems):
    state = fn(state, (t, dt))
  return state


def _stack_leaves(items):
  """tf.scan stacks the per-step outputs leaf by leaf: a list of (nested) tuples of tensors -> the same nesting of [T, ...] tensors"""
  first = items[0]
  if isinstance(first, (tuple, list)):
    return type(first)(_stack_leaves([it[k] for it in items]) for k in range(len(first)))
  return torch.stack([torch.as_tensor(it) for it in items])


def scan(fn, elems, initializer):
  state, out = initializer, []
  for t, dt in zip(*elems):
    state = fn(state, (t, dt))
    out.append(state)
  return _stack_leaves(out) if out else out


class Euler:
  @classmethod
  def step(cls, func: Callable, t: float, dt: float, x: torch.Tensor) -> torch.Tensor:
    dx_dt, sqrt_cov = func(t, x)
    _x = x + dt * dx_dt
    if sqrt_cov is None:
      return _x
    rvs = torch.randn(_x.shape, dtype=_x.dtype, device=_x.device)
    return _x + ((dt ** 0.5) * sqrt_cov @ rvs.unsqueeze(-1)).squeeze(-1)

  @classmethod
  def __call__(cls, func, initial_time, initial_state, solution_times, callbacks_and_initializers: List[Tuple[Callable, Any]] = None,
               iterator: Callable = scan):
    if callbacks_and_initializers is None:
      initializer = initial_state
    else:
      callbacks, inits = zip(*callbacks_and_initializers)
      initializer = (initial_state,) + tuple(inits)

    def body(carry, elems):
      t, dt = elems
      if callbacks_and_initializers is None:
        return cls.step(func=func, t=t, dt=dt, x=carry)
      state, *cb_args = carry
      new_state = cls.step(func=func, t=t, dt=dt, x=state)
      return (new_state,) + tuple(cb(t, new_state, a) for cb, a in zip(callbacks, cb_args))

    st = np.asarray(solution_times, dtype=np.float64)
    step_sizes = np.concatenate([st[:1] - initial_time, st[1:] - st[:-1]], axis=0)
    return iterator(fn=body, elems=(st, step_sizes), initializer=initializer)


class MomentMatchingEuler(Euler):
  @classmethod
  def step(cls, func, t, dt, x):
    x = GaussianMoments(moments=x, centered=True)
    match_drift, match_noise = func(t, x)
    assert match_noise is None
    mf = match_drift.y.mean()
    Sxf = match_drift.cross_covariance()
    Sff = match_drift.y.covariance()
    _mx = x.mean() + dt * mf
    _Sxx = x.covariance() + dt * (Sxf + Sxf.transpose(-1, -2)) + (dt ** 2) * Sff
    return _mx, _Sxx
