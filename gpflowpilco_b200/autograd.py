"""Gradient shims around the backward entry points of libgpp_b200.so.

Upstream obtains gradients by `tape.gradient(loss, policy.trainable_variables)` (gpflow_pilco/utils/optimizers.py:52-56);
a TensorFlow binding would wrap the same two C calls in `tf.custom_gradient`.  These
`torch.autograd.Function`s are that shim for the tensors this package already uses for device memory:

  mm_predict(handle, m, S)                      differentiable w.r.t. the input moments (gpp_mm_gp_predict_fwd / _bwd)
  rollout_mm_loss(dynamics, policy, m0, S0, …)  differentiable w.r.t. policy Z, lengthscales, q_mu and (m0, S0)
                                                (gpp_rollout_mm_fwd / _bwd)
The C library returns gradients w.r.t. CONSTRAINED values; bijector chain rules stay in the framework.  No arithmetic of the
path happens in torch: beta = Kuu^-1 m and its adjoint are gpp_policy_prepare / gpp_policy_prepare_bwd.
"""
from __future__ import annotations

from typing import Sequence

import torch

from gpflowpilco_b200.ops import GPModelHandle
from gpflowpilco_b200.rollouts import PolicyParams, policy_beta_bwd, rollout_mm, rollout_mm_bwd


class _MMPredict(torch.autograd.Function):
  @staticmethod
  def forward(ctx, m, S, handle, full_output_cov, jitter):
    f1, Sff, cross = handle.predict(m.detach(), S.detach(), full_output_cov=full_output_cov, jitter=jitter)
    ctx.save_for_backward(m.detach(), S.detach())
    ctx.handle, ctx.full = handle, full_output_cov
    return f1, Sff, cross

  @staticmethod
  def backward(ctx, f1_bar, Sff_bar, cross_bar):
    m, S = ctx.saved_tensors
    m_bar, S_bar = ctx.handle.predict_bwd(m, S, f1_bar, Sff_bar, cross_bar, full_output_cov=ctx.full)
    return m_bar, S_bar, None, None, None


def mm_predict(handle: GPModelHandle, m: torch.Tensor, S: torch.Tensor, full_output_cov: bool = True, jitter: float = 0.0):
  """(f1, Sff, cross) of the fused moment-matched GP predict, differentiable w.r.t. (m, S)."""
  return _MMPredict.apply(m, S, handle, full_output_cov, jitter)


class _RolloutMM(torch.autograd.Function):
  @staticmethod
  def forward(ctx, Z, lengthscales, q_mu, m0, S0, dynamics, variance, whiten, jitter, scale, shift, horizon, active_dims, target, W,
              check):
    pol = PolicyParams(Z.detach(), lengthscales.detach(), variance, q_mu.detach(), whiten=whiten, jitter=jitter, squash_scale=scale,
                       squash_shift=shift)
    beta = pol.beta(check=check)                                                           # gpp_policy_prepare
    res = rollout_mm(dynamics, pol, m0.detach(), S0.detach(), horizon, active_dims, target, W, beta=beta, save_for_backward=True,
                     check=check)
    ctx.save_for_backward(beta, res.traj_m, res.traj_S, res.saved)
    ctx.pol, ctx.dynamics, ctx.active_dims, ctx.target, ctx.W, ctx.check = pol, dynamics, tuple(active_dims), target, W, check
    return res.loss

  @staticmethod
  def backward(ctx, loss_bar):
    beta, traj_m, traj_S, saved = ctx.saved_tensors
    Zb, eb, bb, m0b, S0b = rollout_mm_bwd(ctx.dynamics, ctx.pol, beta, traj_m, traj_S, ctx.active_dims, ctx.target, ctx.W,
                                          loss_bar=loss_bar.contiguous(), saved=saved, check=ctx.check)
    qb = policy_beta_bwd(ctx.pol, beta, bb, Zb, eb)                                         # gpp_policy_prepare_bwd
    return Zb, eb, qb, m0b, S0b, None, None, None, None, None, None, None, None, None, None, None


def rollout_mm_loss(dynamics: GPModelHandle, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor,
                    m0: torch.Tensor, S0: torch.Tensor, horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor,
                    cost_W: torch.Tensor, squash_scale: float = 1.0, squash_shift: float = -0.5, whiten: bool = True,
                    jitter: float = 1e-6, check: bool = True) -> torch.Tensor:
  """loss[N] of the moment-matched rollout (upstream MomentMatchingPILCO closure, loops/pilco.py:192-220), differentiable
  w.r.t. the policy parameters Z [R,Mp,De], lengthscales [R,De], q_mu [R,Mp] and the initial moments (m0, S0).
  `check=False` skips the synchronising reads of the not-positive-definite flags (timed loops, CUDA-graph capture);
  `check="defer"` queues them for one `rollouts.raise_deferred()` after the backward pass (no synchronisation in between)."""
  return _RolloutMM.apply(Z, lengthscales, q_mu, m0, S0, dynamics, variance, bool(whiten), float(jitter), float(squash_scale),
                          float(squash_shift), int(horizon), tuple(active_dims), cost_target, cost_W, check if isinstance(check, str) else bool(check))


class _RolloutPathwise(torch.autograd.Function):
  @staticmethod
  def forward(ctx, Z, lengthscales, q_mu, x0, paths, variance, whiten, jitter, scale, shift, horizon, active_dims, target, W):
    from gpflowpilco_b200.pathwise import rollout_pathwise
    pol = PolicyParams(Z.detach(), lengthscales.detach(), variance, q_mu.detach(), whiten=whiten, jitter=jitter, squash_scale=scale,
                       squash_shift=shift)
    beta = pol.beta()
    loss, _, traj, jac = rollout_pathwise(paths, pol, x0.detach(), horizon, active_dims, target, W, beta=beta, save_for_backward=True)
    ctx.save_for_backward(beta, traj, jac)
    ctx.pol, ctx.active_dims, ctx.target, ctx.W = pol, tuple(active_dims), target, W
    return loss

  @staticmethod
  def backward(ctx, loss_bar):
    from gpflowpilco_b200.pathwise import rollout_pathwise_bwd
    beta, traj, jac = ctx.saved_tensors
    Zb, eb, bb, x0b = rollout_pathwise_bwd(ctx.pol, beta, traj, jac, ctx.active_dims, ctx.target, ctx.W, loss_bar=loss_bar.contiguous())
    Zb, eb, bb = Zb[None].contiguous(), eb[None].contiguous(), bb[None].contiguous()
    qb = policy_beta_bwd(ctx.pol, beta, bb, Zb, eb)
    return Zb, eb, qb, x0b, None, None, None, None, None, None, None, None, None, None


def rollout_pathwise_loss(paths, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor, x0: torch.Tensor,
                          horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor,
                          squash_scale: float = 1.0, squash_shift: float = -0.5, whiten: bool = True, jitter: float = 1e-6) -> torch.Tensor:
  """loss[S] of the particle rollout on the function draws `paths` (upstream PathwisePILCO closure, loops/pilco.py:263-298),
  differentiable w.r.t. the (single, shared) policy's Z [1,Mp,De], lengthscales [1,De], q_mu [1,Mp] and the initial states x0."""
  return _RolloutPathwise.apply(Z, lengthscales, q_mu, x0, paths, variance, bool(whiten), float(jitter), float(squash_scale),
                                float(squash_shift), int(horizon), tuple(active_dims), cost_target, cost_W)
