"""MomentMatchingPILCO / PathwisePILCO policy-loss closures — the part of upstream gpflow_pilco/loops/pilco.py that is on
the hot path (:139-303): policy_loss_closure(), _policy_loss_closure(), get_state_initializer().  Environment handling,
model building/fitting and checkpointing (upstream loops/core.py, loops/model_based_rl.py, pilco.py:40-137) are out of
scope: the loop objects here are constructed from already-built models.

When the task has the upstream cart-pole structure — TrigonometricEncoder, InverseLinkWrapper(KernelRegressor(SVGP)) policy
with the Chain[Scale, Shift, NormalCDF] link, multi-output SVGP drift, GaussianObjective — the closure runs the FUSED device
rollout (one C-ABI call for all H steps).  Otherwise the moment-matched closure falls back to the rule-by-rule path
(DynamicalSystem.solve_forward with the foldl iterator and the loss callback, exactly upstream's structure), whose rules
still execute in the CUDA library.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from gpflowpilco_b200 import pathwise as pw
from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
from gpflowpilco_b200.dynamics import DynamicalSystem, MomentMatchingEuler, foldl
from gpflowpilco_b200.models.core import (BijectorChain, Constant, InverseLinkWrapper, KernelRegressor, LinearCoregionalization, SVGP, Zero)
from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
from gpflowpilco_b200.moment_matching.models import DEFAULT_JITTER, svgp_handle
from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm

__all__ = ("EpisodeSpec", "GaussianStateDistribution", "AbstractPILCO", "MomentMatchingPILCO", "PathwisePILCO")


class GaussianStateDistribution:
  """Stand-in for tfd.MultivariateNormalTriL (upstream examples/cartpole_swingup/experiment.py:131-135)."""

  def __init__(self, loc: torch.Tensor, covariance_matrix: torch.Tensor):
    self.loc, self.cov = loc, covariance_matrix

  def mean(self):
    return self.loc

  def covariance(self):
    return self.cov


class EpisodeSpec(NamedTuple):
  state_distrib: GaussianStateDistribution
  horizon: float
  step_size: float
  initial_time: float = 0.0

  @property
  def num_steps(self) -> int:
    return int(np.ceil(self.horizon / self.step_size - 1e-9))


class AbstractPILCO(DynamicalSystem):
  def __init__(self, episode_spec: EpisodeSpec, objective: Callable, drift: Callable, policy: Callable, encoder: Callable = None,
               diffusion: Callable = None, solver: Callable = None):
    assert diffusion is None, NotImplementedError
    super().__init__(drift=drift, diffusion=None, policy=policy, encoder=encoder, solver=solver)
    self.episode_spec = episode_spec
    self.objective = objective

  # ---- structure detection for the fused path -----------------------------------------------------------
  def _policy_params(self) -> Optional[PolicyParams]:
    pol = self.policy
    if not (isinstance(pol, InverseLinkWrapper) and isinstance(pol.invlink, BijectorChain) and isinstance(pol.model, KernelRegressor)):
      return None
    svgp = pol.model.model
    if not isinstance(svgp, SVGP) or len(svgp.latent_kernels()) != 1 or svgp.latent_kernels()[0].active_dims is not None:
      return None
    # the fused kernels take a zero-mean, single-output regressor: anything else goes through the rule-by-rule path, which honours
    # the mean function and every output column (the two closures must return the same loss for the same objects)
    if not _has_zero_mean(svgp.mean_function) or svgp.q_mu.shape[1] != 1:
      return None
    jitter = DEFAULT_JITTER if svgp.kuu_jitter is None else svgp.kuu_jitter
    if not isinstance(jitter, (int, float)):
      jitter = [float(v) for v in jitter]
      if len(jitter) != 1:
        return None
      jitter = jitter[0]
    try:
      scale, shift = pol.invlink.squash_parameters()
    except NotImplementedError:
      return None
    k, Z = svgp.latent_kernels()[0], svgp.latent_inducing()[0]
    De = Z.shape[-1]
    return PolicyParams(Z[None], k.ell(De)[None], k.variance.reshape(1), svgp.q_mu[:, 0][None], whiten=svgp.whiten,
                        jitter=float(jitter), squash_scale=scale, squash_shift=shift)

  def _fusable(self) -> bool:
    return (isinstance(self.encoder, TrigonometricEncoder) and isinstance(self.objective, GaussianObjective) and
            isinstance(self.drift, SVGP) and not isinstance(self.drift.kernel, LinearCoregionalization) and
            all(k.active_dims is None for k in self.drift.latent_kernels()) and self._policy_params() is not None)


class MomentMatchingPILCO(AbstractPILCO):
  def __init__(self, *args, solver: Callable = None, **kwargs):
    super().__init__(*args, solver=MomentMatchingEuler() if solver is None else solver, **kwargs)

  def policy_loss_closure(self, episode_spec: EpisodeSpec = None, state_initializer: Callable = None, **kwargs):
    episode_spec = self.episode_spec if episode_spec is None else episode_spec
    if state_initializer is None:
      state_initializer = self.get_state_initializer(episode_spec.state_distrib)
    solution_times = np.arange(1, 1 + episode_spec.num_steps, dtype=np.float64)          # upstream :186
    return self._policy_loss_closure(state_initializer=state_initializer, initial_time=episode_spec.initial_time,
                                     solution_times=solution_times, **kwargs)

  def _policy_loss_closure(self, state_initializer: Callable, initial_time: float, solution_times, compile: bool = True,
                           fused: bool = True, **kwargs):
    def _accumulate_loss(t, state, loss):
      x = GaussianMoments(moments=state, centered=True)
      if self.encoder is not None:
        x = moment_matching(x, self.encoder).y
      return loss + self.objective(x=x, t=t)

    def _closure():
      mx, Sxx = state_initializer()
      st = np.asarray(solution_times, dtype=np.float64)
      unit_steps = abs(st[0] - initial_time - 1.0) < 1e-12 and (len(st) < 2 or np.allclose(np.diff(st), 1.0))
      if fused and unit_steps and self._fusable():
        pp = self._policy_params()
        if torch.is_grad_enabled() and any(t.requires_grad for t in (pp.Z, pp.lengthscales, pp.q_mu, mx, Sxx)):
          # differentiable closure: upstream's tape.gradient(loss, policy.trainable_variables) (utils/optimizers.py:52-56)
          # maps to gpp_rollout_mm_fwd + gpp_rollout_mm_bwd through the autograd shim
          from gpflowpilco_b200.autograd import rollout_mm_loss
          return rollout_mm_loss(svgp_handle(self.drift, True), pp.Z, pp.lengthscales, pp.variance, pp.q_mu, mx, Sxx, len(st),
                                 self.encoder.active_dims, self.objective.target, self.objective.precis,
                                 squash_scale=pp.squash_scale, squash_shift=pp.squash_shift, whiten=pp.whiten, jitter=pp.jitter)
        res = rollout_mm(svgp_handle(self.drift, True), pp, mx, Sxx, len(st), self.encoder.active_dims,
                         self.objective.target, self.objective.precis)
        return res.loss
      loss = torch.zeros(mx.shape[:-1], dtype=mx.dtype, device=mx.device)
      _, loss = self.solve_forward(iterator=foldl, initial_time=initial_time, initial_state=(mx, Sxx), solution_times=st,
                                   callbacks_and_initializers=((_accumulate_loss, loss),), **kwargs)
      return loss

    return _closure

  def get_state_initializer(self, p: GaussianStateDistribution):
    mx = p.mean().to(torch.float64)[None]
    Sxx = p.covariance().to(torch.float64)[None]
    return lambda: (mx, Sxx)


class PathwisePILCO(AbstractPILCO):
  def policy_loss_closure(self, episode_spec: EpisodeSpec = None, state_initializer: Callable = None, batch_size: int = 128, **kwargs):
    episode_spec = self.episode_spec if episode_spec is None else episode_spec
    self._batch_size = batch_size
    if state_initializer is None:
      state_initializer = self.get_state_initializer(episode_spec.state_distrib, batch_size=batch_size)
    solution_times = np.arange(1, 1 + episode_spec.num_steps, dtype=np.float64)
    return self._policy_loss_closure(state_initializer=state_initializer, initial_time=episode_spec.initial_time,
                                     solution_times=solution_times, **kwargs)

  def _policy_loss_closure(self, state_initializer: Callable, initial_time: float, solution_times, compile: bool = True,
                           num_bases: int = 1024, paths: pw.PackedPaths = None, seed: int = 0, first_particle: int = 0, **kwargs):
    if not self._fusable():
      raise NotImplementedError("PathwisePILCO needs the cart-pole task structure (TrigonometricEncoder, squashed RBF policy, "
                                "SVGP drift, GaussianObjective) on the device path")
    counter = {"calls": 0}

    def _closure():
      state = state_initializer(seed + counter["calls"], first_particle) if _takes_seed(state_initializer) else state_initializer()
      handle = svgp_handle(self.drift, True)
      _paths = paths
      if _paths is None:   # fresh sample paths with each call of the closure (upstream :281-284)
        _paths = pw.generate_paths(handle, state.shape[0], num_bases, seed + counter["calls"], first_particle)
      counter["calls"] += 1
      pp = self._policy_params()
      if torch.is_grad_enabled() and any(t.requires_grad for t in (pp.Z, pp.lengthscales, pp.q_mu, state)):
        # differentiable closure (upstream differentiates it with tape.gradient, utils/optimizers.py:52-56): gradient-mode forward
        # + reverse sweep through the autograd shim, as the moment-matched closure does
        from gpflowpilco_b200.autograd import rollout_pathwise_loss
        return rollout_pathwise_loss(_paths, pp.Z, pp.lengthscales, pp.variance, pp.q_mu, state, len(solution_times),
                                     self.encoder.active_dims, self.objective.target, self.objective.precis,
                                     squash_scale=pp.squash_scale, squash_shift=pp.squash_shift, whiten=pp.whiten, jitter=pp.jitter)
      loss, _, _ = pw.rollout_pathwise(_paths, pp, state, len(solution_times), self.encoder.active_dims,
                                       self.objective.target, self.objective.precis)
      return loss

    return _closure

  def get_state_initializer(self, p: GaussianStateDistribution, batch_size: int = 128):
    def _initializer(seed: int = 0, first_particle: int = 0):   # initial states 1-to-1 with paths (upstream :300-303)
      return pw.draw_initial_states(p.mean().to(torch.float64), p.covariance().to(torch.float64), seed, first_particle, batch_size)
    return _initializer


def _has_zero_mean(mean_function) -> bool:
  """Zero, or a Constant whose value is identically zero (upstream's cart-pole policy: models/svgp.py builds Constant(0)).  The
  device read behind the second case is cached per tensor version, so a closure evaluated in a loop does not synchronise."""
  if isinstance(mean_function, Zero):
    return True
  if not isinstance(mean_function, Constant):
    return False
  c = mean_function.c
  key = (id(c), c._version, c.data_ptr())
  hit = mean_function.__dict__.get("_zero_check")
  if hit is None or hit[0] != key:
    hit = (key, bool((c == 0).all()))
    mean_function.__dict__["_zero_check"] = hit
  return hit[1]


def _takes_seed(fn) -> bool:
  try:
    import inspect
    return len(inspect.signature(fn).parameters) >= 2
  except (TypeError, ValueError):
    return False
