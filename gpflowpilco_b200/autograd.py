"""Gradient shims around the backward entry points of libgpp_b200.so.

Upstream obtains gradients by `tape.gradient(loss, policy.trainable_variables)` (gpflow_pilco/utils/optimizers.py:52-56);
a TensorFlow binding would wrap the same two C calls in `tf.custom_gradient`.  These
`torch.autograd.Function`s are that shim for the tensors this package already uses for device memory:

  mm_predict(handle, m, S)                      differentiable w.r.t. the input moments (gpp_mm_gp_predict_fwd / _bwd)
  rollout_mm_loss(dynamics, policy, m0, S0, …)  differentiable w.r.t. policy Z, lengthscales, q_mu and (m0, S0)
                                                (gpp_rollout_mm_fwd / _bwd)
The C library returns gradients w.r.t. CONSTRAINED values; bijector chain rules stay in the framework.
"""
from __future__ import annotations

from typing import Sequence

import torch

from gpflowpilco_b200.ops import GPModelHandle
from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm, rollout_mm_bwd


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


def policy_beta(Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor, whiten: bool = True,
                jitter: float = 1e-6) -> torch.Tensor:
  """beta_r = Kuu_r^-1 m_r as differentiable torch ops (the policy's 30 x 30 Cholesky, once per closure evaluation; the
  non-differentiable device kernel is gpp_policy_prepare).  Upstream moment_matching/models.py:228-235."""
  Zs = Z / lengthscales[:, None, :]
  d2 = (Zs[:, :, None, :] - Zs[:, None, :, :]).square().sum(-1)
  K = variance[:, None, None] * torch.exp(-0.5 * d2) + jitter * torch.eye(Z.shape[1], dtype=Z.dtype, device=Z.device)
  L = torch.linalg.cholesky(K)
  q = q_mu.unsqueeze(-1)
  if whiten:
    return torch.linalg.solve_triangular(L.transpose(-1, -2), q, upper=True).squeeze(-1)
  return torch.cholesky_solve(q, L).squeeze(-1)


class _RolloutMM(torch.autograd.Function):
  @staticmethod
  def forward(ctx, Z, lengthscales, beta, m0, S0, dynamics, variance, scale, shift, horizon, active_dims, target, W):
    pol = PolicyParams(Z.detach(), lengthscales.detach(), variance, torch.zeros_like(beta), squash_scale=scale, squash_shift=shift)
    res = rollout_mm(dynamics, pol, m0.detach(), S0.detach(), horizon, active_dims, target, W, beta=beta.detach(), save_for_backward=True)
    ctx.save_for_backward(beta.detach(), res.traj_m, res.traj_S, res.saved)
    ctx.pol, ctx.dynamics, ctx.active_dims, ctx.target, ctx.W = pol, dynamics, tuple(active_dims), target, W
    return res.loss

  @staticmethod
  def backward(ctx, loss_bar):
    beta, traj_m, traj_S, saved = ctx.saved_tensors
    Zb, eb, bb, m0b, S0b = rollout_mm_bwd(ctx.dynamics, ctx.pol, beta, traj_m, traj_S, ctx.active_dims, ctx.target, ctx.W,
                                          loss_bar=loss_bar.contiguous(), saved=saved)
    return Zb, eb, bb, m0b, S0b, None, None, None, None, None, None, None, None


def rollout_mm_loss(dynamics: GPModelHandle, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor,
                    m0: torch.Tensor, S0: torch.Tensor, horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor,
                    cost_W: torch.Tensor, squash_scale: float = 1.0, squash_shift: float = -0.5, whiten: bool = True,
                    jitter: float = 1e-6) -> torch.Tensor:
  """loss[N] of the moment-matched rollout (upstream MomentMatchingPILCO closure, loops/pilco.py:192-220), differentiable
  w.r.t. the policy parameters Z [R,Mp,De], lengthscales [R,De], q_mu [R,Mp] and the initial moments (m0, S0)."""
  beta = policy_beta(Z, lengthscales, variance, q_mu, whiten, jitter)
  return _RolloutMM.apply(Z, lengthscales, beta, m0, S0, dynamics, variance, float(squash_scale), float(squash_shift), int(horizon),
                          tuple(active_dims), cost_target, cost_W)


class _RolloutPathwise(torch.autograd.Function):
  @staticmethod
  def forward(ctx, Z, lengthscales, beta, x0, paths, variance, scale, shift, horizon, active_dims, target, W):
    from gpflowpilco_b200.pathwise import rollout_pathwise
    pol = PolicyParams(Z.detach(), lengthscales.detach(), variance, torch.zeros_like(beta), squash_scale=scale, squash_shift=shift)
    loss, _, traj, jac = rollout_pathwise(paths, pol, x0.detach(), horizon, active_dims, target, W, beta=beta.detach(), save_for_backward=True)
    ctx.save_for_backward(beta.detach(), traj, jac)
    ctx.pol, ctx.active_dims, ctx.target, ctx.W = pol, tuple(active_dims), target, W
    return loss

  @staticmethod
  def backward(ctx, loss_bar):
    from gpflowpilco_b200.pathwise import rollout_pathwise_bwd
    beta, traj, jac = ctx.saved_tensors
    Zb, eb, bb, x0b = rollout_pathwise_bwd(ctx.pol, beta, traj, jac, ctx.active_dims, ctx.target, ctx.W, loss_bar=loss_bar.contiguous())
    return Zb[None], eb[None], bb[None], x0b, None, None, None, None, None, None, None, None


def rollout_pathwise_loss(paths, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor, x0: torch.Tensor,
                          horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor,
                          squash_scale: float = 1.0, squash_shift: float = -0.5, whiten: bool = True, jitter: float = 1e-6) -> torch.Tensor:
  """loss[S] of the particle rollout on the function draws `paths` (upstream PathwisePILCO closure, loops/pilco.py:263-298),
  differentiable w.r.t. the (single, shared) policy's Z [1,Mp,De], lengthscales [1,De], q_mu [1,Mp] and the initial states x0."""
  beta = policy_beta(Z, lengthscales, variance, q_mu, whiten, jitter)
  return _RolloutPathwise.apply(Z, lengthscales, beta, x0, paths, variance, float(squash_scale), float(squash_shift), int(horizon),
                                tuple(active_dims), cost_target, cost_W)
