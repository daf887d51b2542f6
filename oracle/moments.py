"""Oracle: moment containers and elementary moment-matching rules.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Follows
  * Moments.covariance (uncentred -> centred)   gpflow_pilco/moment_matching/core.py:69-110
  * GaussianMatch.cross_covariance / joint      gpflow_pilco/moment_matching/gaussian.py:27-63
  * Chain rule                                  gpflow_pilco/moment_matching/gaussian.py:66-83
  * add / sub / mul / matvec / sin / cos / sincos  gpflow_pilco/moment_matching/maths.py:41-176
  * Shift / Scale / NormalCDF bijector rules    gpflow_pilco/moment_matching/bijectors.py:21-69
  * ndtr                                        gpflow_pilco/utils/bvn.py:38-42
  * Encoder rule                                gpflow_pilco/moment_matching/components.py:19-57
  * Encoder / TrigonometricEncoder              gpflow_pilco/components.py:44-75
  * GaussianObjective                           gpflow_pilco/components.py:21-41
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy import special as sps

DTYPE = torch.float64


# ---------------------------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------------------------
@dataclass
class GaussianMoments:
  """(first moment [N,d], second moment [N,d,d]); ``centered`` says whether the second is a covariance."""
  m1: torch.Tensor
  m2: torch.Tensor
  centered: bool = True

  def mean(self) -> torch.Tensor:
    return self.m1

  def covariance(self) -> torch.Tensor:
    if self.centered:
      return self.m2
    return self.m2 - self.m1.unsqueeze(-1) * self.m1.unsqueeze(-2)   # core.py:97

  @property
  def ndim(self) -> int:
    return self.m1.shape[-1]


@dataclass
class GaussianMatch:
  """x -> y push-forward with cross = (tensor [N,dx,dy], preinv) where preinv means the tensor is
  Cov(x,x)^{-1} Cov(x,y) rather than Cov(x,y)   (gaussian.py:27-31)."""
  x: GaussianMoments
  y: GaussianMoments
  cross: Tuple[torch.Tensor, bool]

  def cross_covariance(self, preinv: bool = False) -> torch.Tensor:
    Sxy, is_preinv = self.cross
    if not preinv and is_preinv:
      return self.x.covariance() @ Sxy                                 # gaussian.py:37-38
    if preinv and not is_preinv:
      return torch.cholesky_solve(Sxy, torch.linalg.cholesky(self.x.covariance()))  # :39-46
    return Sxy

  def joint(self) -> GaussianMoments:
    m = torch.cat([self.x.mean(), self.y.mean()], -1)                  # gaussian.py:57-63
    Sxx, Sxy, Syy = self.x.covariance(), self.cross_covariance(False), self.y.covariance()
    S = torch.cat([torch.cat([Sxx, Sxy], -1),
                   torch.cat([Sxy.transpose(-1, -2), Syy], -1)], -2)
    return GaussianMoments(m, S, True)


def mm_chain(x: GaussianMoments, ops: Sequence[Callable[[GaussianMoments], GaussianMatch]]) -> GaussianMatch:
  """Apply ``ops`` right-to-left (like the reference's Chain), multiplying pre-inverted cross terms
  (gaussian.py:66-83)."""
  state, cross, preinv = x, None, None
  for i, op in enumerate(reversed(list(ops))):
    match = op(state)
    state = match.y
    if i:
      cross = cross @ match.cross_covariance(preinv=True)
    else:
      cross, preinv = match.cross
  return GaussianMatch(x, state, (cross, preinv))


# ---------------------------------------------------------------------------------------------
# maths rules
# ---------------------------------------------------------------------------------------------
def _eye_like(x: GaussianMoments) -> torch.Tensor:
  N, d = x.m1.shape
  return torch.eye(d, dtype=DTYPE).expand(N, d, d)


def mm_add(x: GaussianMoments, c) -> GaussianMatch:
  y = GaussianMoments(x.mean() + c, x.covariance(), True)              # maths.py:47-51
  return GaussianMatch(x, y, (_eye_like(x), True))


def mm_sub(x: GaussianMoments, c) -> GaussianMatch:
  y = GaussianMoments(x.mean() - c, x.covariance(), True)              # maths.py:54-58
  return GaussianMatch(x, y, (_eye_like(x), True))


def mm_mul(x: GaussianMoments, c) -> GaussianMatch:
  c = torch.as_tensor(c, dtype=DTYPE)                                   # maths.py:61-78 (keeps centred flag)
  y = GaussianMoments(c * x.m1, (c ** 2) * x.m2, x.centered)
  return GaussianMatch(x, y, (c * _eye_like(x), True))


def mm_matvec(x: GaussianMoments, a: torch.Tensor, adjoint_a: bool = False) -> GaussianMatch:
  A = a.transpose(-1, -2) if adjoint_a else a                          # maths.py:81-94
  y1 = (A @ x.m1.unsqueeze(-1)).squeeze(-1)
  y2 = A @ x.m2 @ A.transpose(-1, -2)
  N = x.m1.shape[0]
  cross = A.transpose(-1, -2).expand(N, *A.transpose(-1, -2).shape[-2:])
  return GaussianMatch(x, GaussianMoments(y1, y2, x.centered), (cross, True))


def _trig_terms(x: GaussianMoments):
  x1, Sxx = x.mean(), x.covariance()
  vx = torch.diagonal(Sxx, dim1=-2, dim2=-1)
  vsum = vx.unsqueeze(-1) + vx.unsqueeze(-2)
  ssum = Sxx + Sxx.transpose(-1, -2)
  A = torch.exp(-0.5 * (vsum + ssum))
  B = torch.exp(-0.5 * (vsum - ssum))
  A_cos_add = A * torch.cos(x1.unsqueeze(-1) + x1.unsqueeze(-2))
  B_cos_sub = B * torch.cos(x1.unsqueeze(-1) - x1.unsqueeze(-2))
  return x1, vx, A, B, A_cos_add, B_cos_sub


def mm_cos(x: GaussianMoments) -> GaussianMatch:
  x1, vx, A, B, Aca, Bcs = _trig_terms(x)                              # maths.py:97-117
  evx = torch.exp(-0.5 * vx)
  y = GaussianMoments(evx * torch.cos(x1), 0.5 * (Bcs + Aca), False)
  return GaussianMatch(x, y, (torch.diag_embed(-torch.sin(x1) * evx), True))


def mm_sin(x: GaussianMoments) -> GaussianMatch:
  x1, vx, A, B, Aca, Bcs = _trig_terms(x)                              # maths.py:120-140
  evx = torch.exp(-0.5 * vx)
  y = GaussianMoments(evx * torch.sin(x1), 0.5 * (Bcs - Aca), False)
  return GaussianMatch(x, y, (torch.diag_embed(torch.cos(x1) * evx), True))


def mm_sincos(x: GaussianMoments) -> GaussianMatch:
  """y = [sin(x), cos(x)]; uncentred second moment, pre-inverted cross (maths.py:143-176)."""
  x1, vx, A, B, Aca, Bcs = _trig_terms(x)
  evx = torch.exp(-0.5 * vx)
  cx, sx = torch.cos(x1), torch.sin(x1)
  c1, s1 = evx * cx, evx * sx
  c2, s2 = 0.5 * (Bcs + Aca), 0.5 * (Bcs - Aca)
  sc_outer = sx.unsqueeze(-1) * cx.unsqueeze(-2)
  sc = 0.5 * (sc_outer * (B + A) - sc_outer.transpose(-1, -2) * (B - A))
  y1 = torch.cat([s1, c1], -1)
  y2 = torch.cat([torch.cat([s2, sc], -1), torch.cat([sc.transpose(-1, -2), c2], -1)], -2)
  cross = torch.cat([torch.diag_embed(c1), torch.diag_embed(-s1)], -1)
  return GaussianMatch(x, GaussianMoments(y1, y2, False), (cross, True))


def sincos(x: torch.Tensor) -> torch.Tensor:
  return torch.cat([torch.sin(x), torch.cos(x)], -1)                   # maths.py:23-24


# ---------------------------------------------------------------------------------------------
# bijector rules
# ---------------------------------------------------------------------------------------------
def ndtr(x: torch.Tensor) -> torch.Tensor:
  return 0.5 * torch.erfc(-x / math.sqrt(2.0))                         # bvn.py:38-42


class _OwensT(torch.autograd.Function):
  """Owen's T(h,a) via scipy (== tfp.math.owens_t) with closed-form partials (App. A.4)."""

  @staticmethod
  def forward(ctx, h, a):
    ctx.save_for_backward(h, a)
    return torch.from_numpy(np.asarray(sps.owens_t(h.detach().numpy(), a.detach().numpy()))).to(DTYPE)

  @staticmethod
  def backward(ctx, g):
    h, a = ctx.saved_tensors
    dT_da = torch.exp(-0.5 * h * h * (1 + a * a)) / (2 * math.pi * (1 + a * a))
    phi = torch.exp(-0.5 * h * h) / math.sqrt(2 * math.pi)
    dT_dh = -0.5 * phi * torch.erf(a * h / math.sqrt(2.0))
    return g * dT_dh, g * dT_da


def owens_t(h: torch.Tensor, a: torch.Tensor) -> torch.Tensor:
  return _OwensT.apply(h, a)


def mm_ndtr(x: GaussianMoments) -> GaussianMatch:
  """y = Phi(x), uncentred second moment (bijectors.py:37-69).  Only the 1-D branch (owens_t, :57-58) and
  the diagonal of the multi-D branch are closed-form here; the off-diagonal multi-D entries use Genz BVN
  (oracle/bvn.py) exactly as :59-63 (lower limit -9)."""
  x1, Sxx = x.mean(), x.covariance()
  vx = torch.diagonal(Sxx, dim1=-2, dim2=-1)
  vw = vx + 1
  isq_vw = torch.rsqrt(vw)
  h = isq_vw * x1
  y1 = ndtr(h)
  if x.ndim == 1:
    y2 = (y1 - 2 * owens_t(h, torch.rsqrt(1 + 2 * vx))).unsqueeze(-1)
  else:
    from oracle.bvn import bvn
    d = x.ndim
    lower = torch.full((d, d), -9.0, dtype=DTYPE)
    upper = h.unsqueeze(-1).expand(*h.shape, d)
    rho = Sxx * isq_vw.unsqueeze(-1) * isq_vw.unsqueeze(-2)
    y2 = bvn(lower, upper, lower, upper.transpose(-1, -2), rho)
  vxy = isq_vw * vx * (2 * math.pi) ** -0.5 * torch.exp(-0.5 * h * h)
  cross = torch.diag_embed(vxy / vx)                                   # :65-66
  return GaussianMatch(x, GaussianMoments(y1, y2, False), (cross, True))


def mm_squash(x: GaussianMoments, scale: float, shift: float = -0.5) -> GaussianMatch:
  """tfb.Chain([Scale(scale), Shift(shift), NormalCDF()])  ->  scale * (Phi(x) + shift)
  (examples/cartpole_swingup/swingup_loops.py:87-90; bijectors.py:21-34)."""
  return mm_chain(x, [lambda s: mm_mul(s, scale), lambda s: mm_add(s, shift), mm_ndtr])


def squash(f: torch.Tensor, scale: float, shift: float = -0.5) -> torch.Tensor:
  return scale * (ndtr(f) + shift)


# ---------------------------------------------------------------------------------------------
# encoder + objective
# ---------------------------------------------------------------------------------------------
class TrigonometricEncoder:
  """[sin(a), cos(a), b]: transform of active dims first, inactive appended (components.py:44-75)."""

  def __init__(self, active_dims: Sequence[int]):
    self.active_dims = tuple(active_dims)

  def partition(self, ndims: int):
    a = tuple(range(ndims)[d] for d in self.active_dims)
    assert len(set(a)) == len(a)
    return a, tuple(sorted(set(range(ndims)) - set(a)))                # components.py:58-67

  def __call__(self, x: torch.Tensor) -> torch.Tensor:
    a, b = self.partition(x.shape[-1])
    out = sincos(x[..., list(a)])
    if len(b):
      out = torch.cat([out, x[..., list(b)]], -1)
    return out

  def out_dims(self, ndims: int) -> int:
    return ndims + len(self.active_dims)


def mm_encoder(x: GaussianMoments, enc: TrigonometricEncoder) -> GaussianMatch:
  """moments of e = [sincos(x_a), x_b] and Cov(x, e) (NOT pre-inverted) (mm/components.py:19-57)."""
  x1, Sxx = x.mean(), x.covariance()
  a, b = (list(t) for t in enc.partition(x1.shape[-1]))
  Sxa = Sxx[..., :, a]
  Saa = Sxa[..., a, :]
  part = mm_sincos(GaussianMoments(x1[..., a], Saa, True))
  iSaa_Say = part.cross_covariance(preinv=True)
  Sxy = Sxa @ iSaa_Say
  y1 = torch.cat([part.y.mean(), x1[..., b]], -1)
  Sxb = Sxx[..., :, b]
  Sbb = Sxb[..., b, :]
  Sby = Sxy[..., b, :]
  Syy = part.y.covariance()
  Syy = torch.cat([torch.cat([Syy, Sby.transpose(-1, -2)], -1), torch.cat([Sby, Sbb], -1)], -2)
  return GaussianMatch(x, GaussianMoments(y1, Syy, True), (torch.cat([Sxy, Sxb], -1), False))


class GaussianObjective:
  """cost(x) = -exp(-1/2 (x-x*)^T W (x-x*)) and its Gaussian expectation (components.py:21-41)."""

  def __init__(self, target: torch.Tensor, precis: torch.Tensor):
    self.target = torch.as_tensor(target, dtype=DTYPE)
    self.precis = torch.as_tensor(precis, dtype=DTYPE)

  def __call__(self, x):
    if isinstance(x, GaussianMoments):
      I = torch.eye(self.precis.shape[-1], dtype=DTYPE)
      IpSW = I + x.covariance() @ self.precis                          # :33
      iSpW = self.precis @ torch.linalg.inv(IpSW)                       # :34
      err = x.mean() - self.target
      dist2 = (err * (iSpW @ err.unsqueeze(-1)).squeeze(-1)).sum(-1)
      return -torch.rsqrt(torch.linalg.det(IpSW)) * torch.exp(-0.5 * dist2)   # :37
    err = x - self.target
    dist2 = (err * (err @ self.precis.T)).sum(-1)
    return -torch.exp(-0.5 * dist2)
