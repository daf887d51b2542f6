"""moment_matching() entry point, registry and moment containers — same surface as the reference
(upstream gpflow_pilco/moment_matching/core.py:35-141, gaussian.py:23-83), on torch CUDA float64 tensors.

Semantics preserved: `centered` flag (uncentred second moments are converted lazily, core.py:80-99), the
`(tensor, preinv)` cross term and its conversion rules (gaussian.py:33-51), `joint()` (gaussian.py:53-63),
right-to-left `Chain` with pre-inverted cross products (gaussian.py:66-83), `register_type` for dispatching on
functions (core.py:46-66) and `partial` unwrapping (core.py:129-131).
The containers do a handful of d x d (d <= 8) products with torch on the device; the heavy rules
(GP models, whole rollouts) call the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial
from typing import Any, Callable, Dict, Hashable, Iterable, Tuple, Type

import torch

from gpflowpilco_b200.utils.dispatch import Dispatcher

__all__ = ("ArrayTypes", "Chain", "dispatcher", "get_type", "moment_matching", "Moments", "MomentMatch",
           "register_type", "GaussianMoments", "GaussianMatch")

ArrayTypes = (torch.Tensor,)
dispatcher = Dispatcher("moment_matching")
_custom_types: Dict[Hashable, Type] = {}


def get_type(obj: Hashable) -> Type:
  return _custom_types[obj]


def register_type(obj: Hashable, name: str = None, bases: Tuple = (), dict: Dict = None, exist_ok: bool = False) -> Type:
  if obj in _custom_types and not exist_ok:
    raise ValueError("Attempted to register a preexisting custom type")
  if name is None:
    name = f"{getattr(obj, '__module__', 'obj')}.{getattr(obj, '__name__', repr(obj))}"
  new_type = _custom_types[obj] = type(name, bases, dict or {})
  return new_type


@dataclass
class Moments:
  moments: Tuple[torch.Tensor, torch.Tensor]
  centered: bool = True

  def __getitem__(self, index):
    return self.moments[index]

  def mean(self) -> torch.Tensor:
    return self.moments[0]

  def covariance(self, dense: bool = None) -> torch.Tensor:
    m1, m2 = self.moments[:2]
    if self.centered:
      return m2
    return m2 - m1.unsqueeze(-1) * m1.unsqueeze(-2)

  @property
  def ndim(self) -> int:
    return self.moments[0].shape[-1]

  @property
  def dtype(self):
    return self.moments[0].dtype


class GaussianMoments(Moments):
  pass


@dataclass
class MomentMatch:
  x: Moments
  y: Moments


@dataclass
class GaussianMatch(MomentMatch):
  x: GaussianMoments
  y: GaussianMoments
  cross: Tuple[torch.Tensor, bool] = None

  def cross_covariance(self, dense: bool = None, preinv: bool = False) -> torch.Tensor:
    Sxy, is_preinv = self.cross
    if not preinv and is_preinv:
      return self.x.covariance() @ Sxy
    if preinv and not is_preinv:
      return torch.cholesky_solve(Sxy, torch.linalg.cholesky(self.x.covariance()))
    return Sxy

  def joint(self) -> GaussianMoments:
    m = torch.cat([self.x.mean(), self.y.mean()], -1)
    Sxx, Sxy, Syy = self.x.covariance(), self.cross_covariance(preinv=False), self.y.covariance()
    S = torch.cat([torch.cat([Sxx, Sxy], -1), torch.cat([Sxy.transpose(-1, -2), Syy], -1)], -2)
    return GaussianMoments(moments=(m, S), centered=True)


class Chain(tuple):
  def __new__(cls, *ops: Iterable[Callable]):
    return super().__new__(cls, ops)

  def __call__(self, x):
    for op in reversed(self):
      x = op(x)
    return x


@dispatcher.register(Moments, partial)
def _mm_partial(x: Moments, op: partial):
  return moment_matching(x, op.func, *op.args, **op.keywords)


@dispatcher.register(GaussianMoments, Chain)
def _mm_gauss_chain(x: GaussianMoments, chain: Chain):
  state, preinv, cross = x, None, None
  for i, op in enumerate(reversed(chain)):
    match = moment_matching(state, op)
    state = match.y
    if i:
      cross = cross @ match.cross_covariance(preinv=True)
    else:
      cross, preinv = match.cross
  return GaussianMatch(x=x, y=state, cross=(cross, preinv))


def moment_matching(x: Moments, obj: Any, *args, **kwargs) -> MomentMatch:
  if isinstance(obj, (partial, Chain)):
    return dispatcher(x, obj, *args, **kwargs)
  if isinstance(obj, Hashable) and obj in _custom_types:
    obj = get_type(obj)()
  return dispatcher(x, obj, *args, **kwargs)
