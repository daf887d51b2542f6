"""Tensor-level wrappers of the fused rollout entry points (gpp_policy_prepare, gpp_rollout_mm_fwd, ...)."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from gpflowpilco_b200 import _lib
from gpflowpilco_b200.ops import F64, GPModelHandle, _c, _dev_check, _new_info, _ptr, _stream, raise_if_not_pd


# ---- not-positive-definite flags --------------------------------------------------------------------------
# Every entry point reports a failed factorisation through an asynchronous device flag.  check=True reads it right away (one device
# synchronisation per call); check="defer" queues it, and raise_deferred() reads all queued flags with ONE synchronisation — a
# policy-optimisation step then enqueues policy weights, forward sweep, reverse sweep and policy adjoint back to back instead of
# draining the GPU three times; check=False drops the flag (CUDA-graph capture, timed kernels).
_DEFERRED: list = []


def _finish_check(info: torch.Tensor, what: str, check) -> None:
  if isinstance(check, str):
    if check != "defer":
      raise ValueError("check must be True, False or 'defer'")
    _DEFERRED.append((info, what))
  elif check:
    raise_if_not_pd(info, what)


def raise_deferred() -> None:
  """Read every flag queued by check='defer' (one device synchronisation) and raise for the first failure."""
  global _DEFERRED
  pending, _DEFERRED = _DEFERRED, []
  if not pending:
    return
  dev0 = pending[0][0].device                  # (flags queued from several devices of one process are gathered on the first)
  values = torch.stack([info.reshape(-1)[0].to(dev0) for info, _ in pending]).tolist()
  for v, (_, what) in zip(values, pending):
    if v:
      raise _lib.GppError(-3, f"{what}: a covariance of batch element / parameter set {int(v) - 1} is not positive definite")


@dataclass
class PolicyParams:
  """R sets of SE-ARD kernel-regressor parameters with a scalar output (the upstream RBF policy):
  Z [R,Mp,Dp], lengthscales [R,Dp], variance [R], q_mu [R,Mp]; squashing link scale*(Phi(f)+shift)."""
  Z: torch.Tensor
  lengthscales: torch.Tensor
  variance: torch.Tensor
  q_mu: torch.Tensor
  whiten: bool = True
  jitter: float = 1e-6
  squash_scale: float = 1.0
  squash_shift: float = -0.5

  def __post_init__(self):
    self.Z, self.lengthscales, self.variance, self.q_mu = map(_c, (self.Z, self.lengthscales, self.variance, self.q_mu))
    _dev_check(self.Z, self.lengthscales, self.variance, self.q_mu)
    R, Mp, Dp = self.Z.shape
    if self.lengthscales.shape != (R, Dp) or self.variance.shape != (R,) or self.q_mu.shape != (R, Mp):
      raise ValueError("PolicyParams: expected Z [R,Mp,Dp], lengthscales [R,Dp], variance [R], q_mu [R,Mp]")

  @property
  def shape(self):
    return tuple(self.Z.shape)

  def beta(self, check: bool = True) -> torch.Tensor:
    """beta_r = Kuu_r^-1 m_r (device kernel, one CTA per parameter set)."""
    R, Mp, Dp = self.Z.shape
    out = torch.empty(R, Mp, dtype=F64, device=self.Z.device)
    info = _new_info(self.Z.device)
    with torch.cuda.device(self.Z.device):
      _lib.check(_lib.load().gpp_policy_prepare(R, Mp, Dp, _ptr(self.Z), _ptr(self.lengthscales), _ptr(self.variance), _ptr(self.q_mu),
                                                int(self.whiten), float(self.jitter), _ptr(out), _ptr(info), _stream()))
    _finish_check(info, "policy Kuu (gpp_policy_prepare)", check)
    return out


def policy_beta_bwd(policy: PolicyParams, beta: torch.Tensor, beta_bar: torch.Tensor, Z_bar: torch.Tensor, lengthscales_bar: torch.Tensor):
  """Chain beta_bar through beta = Kuu^-1 m on the device: returns q_mu_bar [R,Mp] and ADDS the Kuu contributions into
  Z_bar / lengthscales_bar (gpp_policy_prepare_bwd)."""
  beta, beta_bar = _c(beta), _c(beta_bar)
  _dev_check(beta, beta_bar, Z_bar, lengthscales_bar)
  if not (Z_bar.is_contiguous() and lengthscales_bar.is_contiguous()):
    raise ValueError("policy_beta_bwd: gradient buffers must be contiguous (they are updated in place)")
  R, Mp, Dp = policy.shape
  q_bar = torch.empty(R, Mp, dtype=F64, device=beta.device)
  with torch.cuda.device(beta.device):
    _lib.check(_lib.load().gpp_policy_prepare_bwd(R, Mp, Dp, _ptr(policy.Z), _ptr(policy.lengthscales), _ptr(policy.variance), _ptr(beta),
                                                  _ptr(beta_bar), int(policy.whiten), float(policy.jitter), _ptr(Z_bar), _ptr(lengthscales_bar),
                                                  _ptr(q_bar), _stream()))
  return q_bar


ROLLOUT_AUTO, ROLLOUT_LEGACY, ROLLOUT_PERSIST = 0, 1, 2


def set_rollout_mode(mode: int) -> int:
  """Select how the moment-matched rollouts run: ROLLOUT_AUTO (default) = the persistent on-device H-loop whenever the model
  fits, ROLLOUT_LEGACY = one launch per stage and step, ROLLOUT_PERSIST = persistent or error.  Returns the previous mode."""
  lib = _lib.load()
  prev = lib.gpp_rollout_mm_get_mode()
  _lib.check(lib.gpp_rollout_mm_set_mode(int(mode)))
  return prev


@dataclass
class MMRolloutResult:
  loss: torch.Tensor                      # [N]
  m_final: torch.Tensor                   # [N,Dx]
  S_final: torch.Tensor                   # [N,Dx,Dx]
  traj_m: Optional[torch.Tensor] = None   # [H+1,N,Dx]
  traj_S: Optional[torch.Tensor] = None   # [H+1,N,Dx,Dx]
  saved: Optional[torch.Tensor] = None    # per-step joint moments / cross terms kept for the backward


def rollout_mm(dynamics: GPModelHandle, policy: PolicyParams, m0: torch.Tensor, S0: torch.Tensor, horizon: int,
               active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor, return_trajectory: bool = False,
               beta: Optional[torch.Tensor] = None, check: bool = True, save_for_backward: bool = False) -> MMRolloutResult:
  """Moment-matched rollout + expected cost for N initial Gaussian states (upstream loops/pilco.py:192-220)."""
  m0, S0, cost_target, cost_W = map(_c, (m0, S0, cost_target, cost_W))
  _dev_check(m0, S0, cost_target, cost_W)
  N, Dx = m0.shape
  na = len(active_dims)
  De = Dx + na
  R, Mp, Dp = policy.shape
  if S0.shape != (N, Dx, Dx):
    raise ValueError("rollout_mm: S0 must be [N,Dx,Dx]")
  if Dp != De:
    raise ValueError(f"rollout_mm: policy input dim {Dp} != encoded state dim {De}")
  if cost_target.shape != (De,) or cost_W.shape != (De, De):
    raise ValueError("rollout_mm: cost target/precision must live in the encoded space")
  if R not in (1, N):
    raise ValueError("rollout_mm: policy must have 1 or N parameter sets")
  dev = m0.device
  dynamics._same_device(m0, S0, cost_target, cost_W, policy.Z)
  if beta is None:
    beta = policy.beta(check=check)
  lib = _lib.load()
  need = lib.gpp_rollout_mm_workspace_bytes(dynamics._h, N, Dx)
  ws = torch.empty(need, dtype=torch.uint8, device=dev)
  loss = torch.empty(N, dtype=F64, device=dev)
  mf = torch.empty(N, Dx, dtype=F64, device=dev)
  Sf = torch.empty(N, Dx, Dx, dtype=F64, device=dev)
  keep = return_trajectory or save_for_backward
  tm = torch.empty(horizon + 1, N, Dx, dtype=F64, device=dev) if keep else None
  tS = torch.empty(horizon + 1, N, Dx, Dx, dtype=F64, device=dev) if keep else None
  info = _new_info(dev)
  act = (ctypes.c_int * max(na, 1))(*active_dims)
  if save_for_backward:
    saved = torch.empty(lib.gpp_rollout_mm_saved_doubles(dynamics._h, N, Dx, int(horizon)), dtype=F64, device=dev)
    with torch.cuda.device(dev):      # the library launches on the current device and is given its current stream
      _lib.check(lib.gpp_rollout_mm_fwd_save(dynamics._h, N, Dx, na, act, R, Mp, _ptr(policy.Z), _ptr(policy.lengthscales),
                                             _ptr(policy.variance), _ptr(beta), float(policy.squash_scale), float(policy.squash_shift),
                                             _ptr(cost_target), _ptr(cost_W), int(horizon), _ptr(m0), _ptr(S0), _ptr(loss), _ptr(tm), _ptr(tS),
                                             _ptr(mf), _ptr(Sf), _ptr(saved), _ptr(ws), ws.numel(), _ptr(info), _stream()))
    _finish_check(info, "rollout_mm", check)
    return MMRolloutResult(loss, mf, Sf, tm, tS, saved)
  with torch.cuda.device(dev):      # the library launches on the current device and is given its current stream
    _lib.check(lib.gpp_rollout_mm_fwd(dynamics._h, N, Dx, na, act, R, Mp, _ptr(policy.Z), _ptr(policy.lengthscales),
                                      _ptr(policy.variance), _ptr(beta), float(policy.squash_scale), float(policy.squash_shift),
                                      _ptr(cost_target), _ptr(cost_W), int(horizon), _ptr(m0), _ptr(S0), _ptr(loss), _ptr(tm), _ptr(tS),
                                      _ptr(mf), _ptr(Sf), _ptr(ws), ws.numel(), _ptr(info), _stream()))
  _finish_check(info, "rollout_mm", check)
  return MMRolloutResult(loss, mf, Sf, tm, tS)


def rollout_mm_bwd(dynamics: GPModelHandle, policy: PolicyParams, beta: torch.Tensor, traj_m: torch.Tensor, traj_S: torch.Tensor,
                   active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor,
                   loss_bar: Optional[torch.Tensor] = None, check: bool = True, saved: Optional[torch.Tensor] = None):
  """Reverse sweep of `rollout_mm` from its stored trajectory: returns (Z_bar [R,Mp,De], lengthscales_bar [R,De] at fixed beta,
  beta_bar [R,Mp], m0_bar [N,Dx], S0_bar [N,Dx,Dx]).  Upstream: tape.gradient through loops/pilco.py:192-220."""
  traj_m, traj_S, cost_target, cost_W, beta, loss_bar = map(_c, (traj_m, traj_S, cost_target, cost_W, beta, loss_bar))
  _dev_check(traj_m, traj_S, cost_target, cost_W, beta, loss_bar)
  H1, N, Dx = traj_m.shape
  na = len(active_dims)
  De = Dx + na
  R, Mp, Dp = policy.shape
  if traj_S.shape != (H1, N, Dx, Dx) or Dp != De or beta.shape != (R, Mp):
    raise ValueError("rollout_mm_bwd: inconsistent trajectory / policy shapes")
  if loss_bar is not None and loss_bar.shape != (N,):
    raise ValueError("rollout_mm_bwd: loss_bar must be [N]")
  dev = traj_m.device
  dynamics._same_device(traj_m, traj_S, cost_target, cost_W, beta, policy.Z)
  lib = _lib.load()
  if saved is not None and (not saved.is_cuda or saved.numel() != lib.gpp_rollout_mm_saved_doubles(dynamics._h, N, Dx, H1 - 1)):
    raise ValueError("rollout_mm_bwd: `saved` does not match this rollout")
  need = lib.gpp_rollout_mm_bwd_workspace_bytes(dynamics._h, N, Dx, Mp, H1 - 1)
  ws = torch.empty(need, dtype=torch.uint8, device=dev)
  Zb = torch.empty(R, Mp, De, dtype=F64, device=dev)
  eb = torch.empty(R, De, dtype=F64, device=dev)
  bb = torch.empty(R, Mp, dtype=F64, device=dev)
  m0b = torch.empty(N, Dx, dtype=F64, device=dev)
  S0b = torch.empty(N, Dx, Dx, dtype=F64, device=dev)
  info = _new_info(dev)
  act = (ctypes.c_int * max(na, 1))(*active_dims)
  with torch.cuda.device(dev):      # the library launches on the current device and is given its current stream
    _lib.check(lib.gpp_rollout_mm_bwd(dynamics._h, N, Dx, na, act, R, Mp, _ptr(policy.Z), _ptr(policy.lengthscales),
                                      _ptr(policy.variance), _ptr(beta), float(policy.squash_scale), float(policy.squash_shift),
                                      _ptr(cost_target), _ptr(cost_W), H1 - 1, _ptr(traj_m), _ptr(traj_S), _ptr(saved), _ptr(loss_bar),
                                      _ptr(Zb), _ptr(eb), _ptr(bb), _ptr(m0b), _ptr(S0b), _ptr(ws), ws.numel(), _ptr(info), _stream()))
  _finish_check(info, "rollout_mm_bwd", check)
  return Zb, eb, bb, m0b, S0b
