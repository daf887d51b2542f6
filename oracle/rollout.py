"""Oracle: one model step (forward_sde), the Euler solvers and the two PILCO policy-loss closures.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Follows
  * forward_sde (tensor path + 4 Gaussian variants)  gpflow_pilco/dynamics/forward_sde.py:23-137
  * Euler / MomentMatchingEuler                      gpflow_pilco/dynamics/solvers.py:48-135
  * DynamicalSystem.forward / solve_forward          gpflow_pilco/dynamics/dynamical_system.py:34-51
  * MomentMatchingPILCO._policy_loss_closure         gpflow_pilco/loops/pilco.py:192-227
  * PathwisePILCO._policy_loss_closure               gpflow_pilco/loops/pilco.py:263-303
The callables ``drift_mm``/``policy_mm`` take GaussianMoments and return a GaussianMatch; ``drift``/``policy``
take tensors.  dt == 1 and solution_times == 1..H as in gpflow_pilco/loops/pilco.py:186.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from oracle.moments import GaussianMatch, GaussianMoments, GaussianObjective, TrigonometricEncoder, mm_encoder

MM = Callable[[GaussianMoments], GaussianMatch]


def forward_sde_tensor(x, drift, policy=None, encoder=None):
  """forward_sde.py:23-31 (noise is None in PILCO, gpflow_pilco/loops/pilco.py:41-42)."""
  e = x if encoder is None else encoder(x)
  eu = e if policy is None else torch.cat([e, policy(e)], -1)
  return drift(eu)


def forward_sde_gauss(x: GaussianMoments, drift_mm: MM, policy_mm: Optional[MM] = None,
                      encoder: Optional[TrigonometricEncoder] = None) -> GaussianMatch:
  """The four Gaussian registrations of forward_sde.py (:34-137), selected by which parts are present."""
  if policy_mm is None and encoder is None:                            # :34-46
    return drift_mm(x)

  if encoder is None:                                                  # :49-69
    mp = policy_mm(x)
    md = drift_mm(mp.joint())
    if md.cross[1]:
      preinv = mp.cross[1]
      # NOTE :61-62 multiplies Cov(x,u)-like [dx,du] by [dx+du, df]; as written that only type-checks when the
      # policy cross is widened to d=(x,u).  We restate the intent: Cov(x,d) Cov(d,d)^{-1} Cov(d,f).
      Sxx = x.covariance()
      Sxd = torch.cat([Sxx, mp.cross_covariance(preinv=False)], -1)
      cross = (Sxd @ md.cross_covariance(preinv=True), False)
    else:
      cross = (md.cross_covariance()[..., :x.ndim, :], False)
    return GaussianMatch(x, md.y, cross)

  me = mm_encoder(x, encoder)
  if policy_mm is None:                                                # :72-92
    md = drift_mm(me.y)
    preinv = me.cross[1]
    Sxe = me.cross_covariance(preinv=preinv)
    return GaussianMatch(x, md.y, (Sxe @ md.cross_covariance(preinv=True), preinv))

  # encoder + policy                                                   # :95-137
  mp = policy_mm(me.y)
  md = drift_mm(mp.joint())
  ndims_x = x.ndim
  ndims_u = mp.y.ndim
  active, inactive = encoder.partition(ndims_x)
  ndims_b = ndims_x - len(active)
  if me.cross[1]:
    Sax = x.covariance()[..., list(active), :]
    Sae = Sax @ me.cross_covariance(preinv=True)
  else:
    Sae = me.cross_covariance()[..., list(active), :]                  # :115-116
  Sau = Sae @ mp.cross_covariance(preinv=True)                         # :117
  order = sorted(zip(active + inactive, range(ndims_x)))               # :121
  perm = [p for _, p in order]
  Sad = torch.cat([Sae, Sau], -1)
  Sd = md.x.covariance()
  lo = Sd.shape[-2] - ndims_b - ndims_u
  Sbd = Sd[..., lo: Sd.shape[-2] - ndims_u, :]                         # :123
  Sxd = torch.cat([Sad, Sbd], -2)[..., perm, :]
  Sxf = Sxd @ md.cross_covariance(preinv=True)
  return GaussianMatch(x, md.y, (Sxf, False))


def mm_euler_step(x: GaussianMoments, match_drift: GaussianMatch, dt: float = 1.0):
  """solvers.py:110-135 (noise branch is dead code in PILCO)."""
  mf = match_drift.y.mean()
  Sxf = match_drift.cross_covariance()
  Sff = match_drift.y.covariance()
  m = x.mean() + dt * mf
  S = x.covariance() + dt * (Sxf + Sxf.transpose(-1, -2)) + (dt ** 2) * Sff
  return m, S


def mm_rollout(m0, S0, horizon: int, drift_mm: MM, policy_mm: Optional[MM],
               encoder: Optional[TrigonometricEncoder], objective: GaussianObjective,
               return_trajectory: bool = False):
  """MomentMatchingPILCO closure: loss[N] = sum_t E[cost(enc(x_t))], t = 1..H  (pilco.py:199-217)."""
  m, S = m0, S0
  loss = torch.zeros(m.shape[:-1], dtype=m.dtype)
  traj = [(m, S)]
  for _ in range(horizon):
    x = GaussianMoments(m, S, True)
    md = forward_sde_gauss(x, drift_mm, policy_mm, encoder)
    m, S = mm_euler_step(x, md)
    xn = GaussianMoments(m, S, True)
    if encoder is not None:
      xn = mm_encoder(xn, encoder).y                                   # pilco.py:203-204
    loss = loss + objective(xn)
    traj.append((m, S))
  return (loss, traj) if return_trajectory else loss


def pathwise_rollout(x0, horizon: int, drift, policy, encoder, objective: GaussianObjective,
                     return_trajectory: bool = False):
  """PathwisePILCO closure body: Euler.step with no diffusion (solvers.py:49-65) + cost callback
  (pilco.py:272-275).  ``drift`` must evaluate function draw s on row s (pilco.py:300-303)."""
  x = x0
  loss = torch.zeros(x.shape[:-1], dtype=x.dtype)
  traj = [x]
  for _ in range(horizon):
    x = x + forward_sde_tensor(x, drift, policy, encoder)
    e = x if encoder is None else encoder(x)
    loss = loss + objective(e)
    traj.append(x)
  return (loss, traj) if return_trajectory else loss
