"""forward_sde dispatcher — one model step, same registrations as upstream gpflow_pilco/dynamics/forward_sde.py:17-137
(tensor path + the four Gaussian variants selected by which of policy / encoder are present; diffusion is None in
PILCO, upstream loops/pilco.py:41-42).  This is the rule-by-rule path; whole rollouts go through the fused kernels
(gpflowpilco_b200.rollouts / .pathwise) which compute exactly this composition on the device."""
from __future__ import annotations

import torch

from gpflowpilco_b200.moment_matching import GaussianMatch, GaussianMoments, moment_matching
from gpflowpilco_b200.utils.dispatch import Dispatcher

NoneType = type(None)
forward_sde = Dispatcher("forward_sde")


@forward_sde.register(torch.Tensor, object, (object, NoneType), (object, NoneType), (object, NoneType))
def _forward_tensor(x, drift, noise, policy, encoder):
  e = x if encoder is None else encoder(x)
  eu = e if policy is None else torch.cat([e, policy(e)], dim=-1)
  return drift(eu), None if noise is None else noise(e)


@forward_sde.register(GaussianMoments, object, (object, NoneType), (object, NoneType), (object, NoneType))
def _forward_gauss(x, drift, noise, policy, encoder):
  assert noise is None, NotImplementedError("diffusion terms are not part of the PILCO path")
  if policy is None and encoder is None:                       # upstream :34-46
    return moment_matching(x, drift), None
  if encoder is None:                                          # upstream :49-69 (intent: Cov(x,d) Cov(d,d)^-1 Cov(d,f))
    mp = moment_matching(x, policy)
    md = moment_matching(mp.joint(), drift)
    Sxd = torch.cat([x.covariance(), mp.cross_covariance(preinv=False)], dim=-1)
    return GaussianMatch(x=x, y=md.y, cross=(Sxd @ md.cross_covariance(preinv=True), False)), None
  me = moment_matching(x, encoder)
  if policy is None:                                           # upstream :72-92
    md = moment_matching(me.y, drift)
    preinv = me.cross[1]
    Sxe = me.cross_covariance(preinv=preinv)
    return GaussianMatch(x=x, y=md.y, cross=(Sxe @ md.cross_covariance(preinv=True), preinv)), None
  mp = moment_matching(me.y, policy)                           # upstream :95-137
  md = moment_matching(mp.joint(), drift)
  ndims_x = x.ndim
  ndims_u = mp.y.ndim
  active, inactive = encoder.get_partition_indices(ndims_x)
  ndims_b = ndims_x - len(active)
  if me.cross[1]:
    Sae = x.covariance()[..., list(active), :] @ me.cross_covariance(preinv=True)
  else:
    Sae = me.cross_covariance()[..., list(active), :]
  Sau = Sae @ mp.cross_covariance(preinv=True)
  perm = [p for _, p in sorted(zip(active + inactive, range(ndims_x)))]
  Sad = torch.cat([Sae, Sau], dim=-1)
  Sd = md.x.covariance()
  lo = Sd.shape[-2] - ndims_b - ndims_u
  Sbd = Sd[..., lo: Sd.shape[-2] - ndims_u, :]
  Sxd = torch.cat([Sad, Sbd], dim=-2)[..., perm, :]
  return GaussianMatch(x=x, y=md.y, cross=(Sxd @ md.cross_covariance(preinv=True), False)), None
