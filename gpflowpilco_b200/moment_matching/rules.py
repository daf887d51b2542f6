"""Moment-matching rules for maths ops, bijectors and encoders — the registrations of upstream
gpflow_pilco/moment_matching/maths.py:41-176, bijectors.py:21-69 and components.py:19-57 on the CUDA path.

Dispatch on functions works like upstream: `register_type(fn)` creates a surrogate type and
`moment_matching(x, fn, ...)` resolves through it (core.py:46-66,134-141).  torch functions stand in for the
tf functions upstream registers (torch.add <-> tf.math.add, ...).
"""
from __future__ import annotations

import torch

from gpflowpilco_b200 import _lib
from gpflowpilco_b200.components import Encoder, encoder_moments, sincos
from gpflowpilco_b200.models.core import BijectorChain, NormalCDF, Scale, Shift
from gpflowpilco_b200.moment_matching.core import (Chain, GaussianMatch, GaussianMoments, Moments, dispatcher, moment_matching,
                                                   register_type)
from gpflowpilco_b200.ops import F64, _c, _dev_check, _ptr, _stream

NumericalTypes = (int, float, torch.Tensor)

_type_identity = register_type(torch.clone, name="torch.identity")
_type_add = register_type(torch.add)
_type_sub = register_type(torch.sub)
_type_mul = register_type(torch.mul)
_type_matvec = register_type(torch.mv)
_type_cos = register_type(torch.cos)
_type_sin = register_type(torch.sin)
_type_sincos = register_type(sincos)


def _eye(x: GaussianMoments) -> torch.Tensor:
  N, d = x.mean().shape
  return torch.eye(d, dtype=F64, device=x.mean().device).expand(N, d, d)


@dispatcher.register(GaussianMoments, _type_identity)
def _mm_gauss_identity(x, _, c=None):
  return GaussianMatch(x=x, y=x, cross=(_eye(x), True))


@dispatcher.register(GaussianMoments, _type_add, NumericalTypes)
def _mm_gauss_add(x, _, c):
  y = GaussianMoments(moments=(x.mean() + c, x.covariance()), centered=True)
  return GaussianMatch(x=x, y=y, cross=(_eye(x), True))


@dispatcher.register(GaussianMoments, _type_sub, NumericalTypes)
def _mm_gauss_sub(x, _, c):
  y = GaussianMoments(moments=(x.mean() - c, x.covariance()), centered=True)
  return GaussianMatch(x=x, y=y, cross=(_eye(x), True))


@dispatcher.register(GaussianMoments, _type_mul, NumericalTypes)
def _mm_gauss_mul(x, _, c):
  c = torch.as_tensor(c, dtype=F64, device=x.mean().device)
  y = GaussianMoments(moments=(c * x[0], (c ** 2) * x[1]), centered=x.centered)
  return GaussianMatch(x=x, y=y, cross=(c * _eye(x), True))


@dispatcher.register(GaussianMoments, _type_matvec, torch.Tensor)
def _mm_gauss_matvec(x, _, a, adjoint_a: bool = False):
  A = a.transpose(-1, -2) if adjoint_a else a
  y1 = (A @ x[0].unsqueeze(-1)).squeeze(-1)
  y2 = A @ x[1] @ A.transpose(-1, -2)
  cross = A.transpose(-1, -2).expand(x[0].shape[0], *A.transpose(-1, -2).shape[-2:])
  return GaussianMatch(x=x, y=GaussianMoments(moments=(y1, y2), centered=x.centered), cross=(cross, True))


def _sincos_moments(x: GaussianMoments):
  d = x.ndim
  me, See, _ = encoder_moments(x.mean(), x.covariance(), tuple(range(d)))
  return me, See


@dispatcher.register(GaussianMoments, _type_sincos)
def _mm_gauss_sincos(x, _):
  me, See = _sincos_moments(x)
  d = x.ndim
  s1, c1 = me[:, :d], me[:, d:]
  cross = torch.cat([torch.diag_embed(c1), torch.diag_embed(-s1)], dim=-1)       # maths.py:173
  return GaussianMatch(x=x, y=GaussianMoments(moments=(me, See), centered=True), cross=(cross, True))


@dispatcher.register(GaussianMoments, _type_sin)
def _mm_gauss_sin(x, _):
  me, See = _sincos_moments(x)
  d = x.ndim
  y = GaussianMoments(moments=(me[:, :d], See[:, :d, :d]), centered=True)
  return GaussianMatch(x=x, y=y, cross=(torch.diag_embed(me[:, d:]), True))           # maths.py:138


@dispatcher.register(GaussianMoments, _type_cos)
def _mm_gauss_cos(x, _):
  me, See = _sincos_moments(x)
  d = x.ndim
  y = GaussianMoments(moments=(me[:, d:], See[:, d:, d:]), centered=True)
  return GaussianMatch(x=x, y=y, cross=(torch.diag_embed(-me[:, :d]), True))          # maths.py:115


# ---- bijectors ------------------------------------------------------------------------------------------
@dispatcher.register(Moments, BijectorChain)
def _mm_chain(x, bijector: BijectorChain, **kwargs):
  return moment_matching(x, Chain(*bijector.bijectors), **kwargs)


@dispatcher.register(Moments, Shift)
def _mm_shift(x, bijector: Shift, **kwargs):
  return moment_matching(x, torch.add, bijector.shift, **kwargs)


@dispatcher.register(Moments, Scale)
def _mm_scale(x, bijector: Scale, **kwargs):
  return moment_matching(x, torch.mul, bijector.scale, **kwargs)


class _SquashND(torch.autograd.Function):
  """gpp_mm_squash_nd with its closed-form reverse mode (gpp_mm_squash_nd_bwd): upstream differentiates the multi-dimensional
  NormalCDF rule through utils/bvn.py with the tape (optimizers.py:52-56)."""

  @staticmethod
  def forward(ctx, mf, Sf):
    lib = _lib.load()
    mf, Sf = _c(mf.detach()), _c(Sf.detach())
    N, A = mf.shape
    mu = torch.empty(N, A, dtype=F64, device=mf.device)
    Su = torch.empty(N, A, A, dtype=F64, device=mf.device)
    gain = torch.empty(N, A, dtype=F64, device=mf.device)
    _lib.check(lib.gpp_mm_squash_nd(N, A, _ptr(mf), _ptr(Sf), 1.0, 0.0, _ptr(mu), _ptr(Su), _ptr(gain), _stream()))
    ctx.save_for_backward(mf, Sf)
    return mu, Su, gain

  @staticmethod
  def backward(ctx, mu_bar, Su_bar, gain_bar):
    lib = _lib.load()
    mf, Sf = ctx.saved_tensors
    N, A = mf.shape
    bars = [None if b is None else _c(b.to(F64)) for b in (mu_bar, Su_bar, gain_bar)]
    mf_bar, Sf_bar = torch.empty_like(mf), torch.empty_like(Sf)
    _lib.check(lib.gpp_mm_squash_nd_bwd(N, A, _ptr(mf), _ptr(Sf), 1.0, 0.0, *(None if b is None else _ptr(b) for b in bars),
                                        _ptr(mf_bar), _ptr(Sf_bar), _stream()))
    return mf_bar, Sf_bar


@dispatcher.register(GaussianMoments, NormalCDF)
def _mm_gauss_ndtr(x, _):
  """y = Phi(x): the 1-D owens_t branch (upstream bijectors.py:37-58) and the multi-dimensional branch with Genz's bivariate
  normal probabilities (bijectors.py:59-63, utils/bvn.py:67-232), both on the device."""
  lib = _lib.load()
  if x.ndim == 1:
    mf, vf = _c(x.mean()[:, 0]), _c(x.covariance()[:, 0, 0])
    _dev_check(mf, vf)
    N = mf.shape[0]
    mu, vu, gain = (torch.empty(N, dtype=F64, device=mf.device) for _ in range(3))
    _lib.check(lib.gpp_mm_squash(N, _ptr(mf), _ptr(vf), 1.0, 0.0, _ptr(mu), _ptr(vu), _ptr(gain), _stream()))
    y = GaussianMoments(moments=(mu[:, None], vu[:, None, None]), centered=True)
    return GaussianMatch(x=x, y=y, cross=(gain[:, None, None], True))
  mf, Sf = _c(x.mean()), _c(x.covariance())
  _dev_check(mf, Sf)
  mu, Su, gain = _SquashND.apply(mf, Sf)
  return GaussianMatch(x=x, y=GaussianMoments(moments=(mu, Su), centered=True), cross=(torch.diag_embed(gain), True))


# ---- encoder --------------------------------------------------------------------------------------------
@dispatcher.register(GaussianMoments, Encoder)
def _mm_gauss_encoder(x, encoder: Encoder, append_inactive: bool = True):
  if encoder.transform is not sincos:
    raise NotImplementedError("only the trigonometric (sincos) encoder is implemented on the device")
  me, See, Cxe = encoder_moments(x.mean(), x.covariance(), encoder.active_dims)
  if not append_inactive:
    k = 2 * len(encoder.active_dims)
    me, See, Cxe = me[:, :k], See[:, :k, :k], Cxe[:, :, :k]
  return GaussianMatch(x=x, y=GaussianMoments(moments=(me, See), centered=True), cross=(Cxe, False))
