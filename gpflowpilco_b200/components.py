"""GaussianObjective, Encoder, TrigonometricEncoder — same surface as upstream gpflow_pilco/components.py:21-75.

Tensors are torch CUDA float64; the moment rules and the costs run in libgpp_b200.so.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Sequence, Tuple

import torch

from gpflowpilco_b200 import _lib
from gpflowpilco_b200.moment_matching.core import GaussianMoments
from gpflowpilco_b200.ops import F64, _c, _dev_check, _ptr, _stream

__all__ = ("GaussianObjective", "Encoder", "TrigonometricEncoder")


def sincos(x: torch.Tensor, axis: int = -1) -> torch.Tensor:
  return torch.cat([torch.sin(x), torch.cos(x)], dim=axis)


class GaussianObjective:
  """cost(x) = -exp(-1/2 (x - x*)^T W (x - x*)); on GaussianMoments its closed-form expectation."""

  def __init__(self, target: torch.Tensor, precis: torch.Tensor):
    self.target = target
    self.precis = precis

  def __call__(self, x, t=None):
    target, W = _c(self.target.to(F64)), _c(self.precis.to(F64))
    De = target.shape[-1]
    lib = _lib.load()
    if isinstance(x, GaussianMoments):
      m, S = _c(x.mean()), _c(x.covariance())
      _dev_check(m, S, target, W)
      out = torch.empty(m.shape[0], dtype=F64, device=m.device)
      _lib.check(lib.gpp_cost_gaussian(m.shape[0], De, _ptr(m), _ptr(S), _ptr(target), _ptr(W), _ptr(out), _stream()))
      return out
    e = _c(x.reshape(-1, De))
    _dev_check(e, target, W)
    out = torch.empty(e.shape[0], dtype=F64, device=e.device)
    _lib.check(lib.gpp_cost_samples(e.shape[0], De, _ptr(e), _ptr(target), _ptr(W), _ptr(out), _stream()))
    return out.reshape(x.shape[:-1])


class Encoder:
  def __init__(self, transform: Callable, active_dims: Sequence[int]):
    self._transform = transform
    self.active_dims = tuple(active_dims)

  def __call__(self, x: torch.Tensor, append_inactive: bool = True) -> torch.Tensor:
    active, inactive = self.get_partition_indices(ndims=x.shape[-1])
    out = self.transform(x[..., list(active)])
    if append_inactive and len(inactive):
      out = torch.cat([out, x[..., list(inactive)]], dim=-1)
    return out

  def get_partition_indices(self, ndims: int) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    indices_x = tuple(range(ndims))
    indices_a = tuple(indices_x[d] for d in self.active_dims)
    assert len(indices_a) == len(set(indices_a))
    return indices_a, tuple(sorted(set(indices_x) - set(indices_a)))

  @property
  def transform(self):
    return self._transform


class TrigonometricEncoder(Encoder):
  def __init__(self, active_dims: Sequence[int]):
    super().__init__(transform=sincos, active_dims=active_dims)


def encoder_moments(m: torch.Tensor, S: torch.Tensor, active_dims: Sequence[int]):
  """(me [N,De], See [N,De,De], Cov(x,e) [N,Dx,De]) of e = [sin x_a, cos x_a, x_b] (gpp_mm_encoder)."""
  m, S = _c(m), _c(S)
  _dev_check(m, S)
  N, Dx = m.shape
  na = len(active_dims)
  De = Dx + na
  me = torch.empty(N, De, dtype=F64, device=m.device)
  See = torch.empty(N, De, De, dtype=F64, device=m.device)
  Cxe = torch.empty(N, Dx, De, dtype=F64, device=m.device)
  act = (ctypes.c_int * max(na, 1))(*active_dims)
  _lib.check(_lib.load().gpp_mm_encoder(N, Dx, na, act, _ptr(m), _ptr(S), _ptr(me), _ptr(See), _ptr(Cxe), _stream()))
  return me, See, Cxe
