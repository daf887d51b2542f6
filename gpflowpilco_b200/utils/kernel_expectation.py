"""kernel_expectation(p, obj1[, obj2]) — the call signature upstream uses for GPflow's expectation dispatcher
(gpflow_pilco/utils/kernel_expectation.py:25-26 re-export; registrations :72-288):

  kernel_expectation(p, kernel)                      -> eKff  [N] (or [N,L] for multi-output kernels)
  kernel_expectation(p, (kernel, Z))                 -> eKfu  [N,M] (or [N,M,L])
  kernel_expectation(p, (kernel1, Z1), (kernel2, Z2)) -> eKuffu [N,M1,M2] (or [N,L1,M1,L2,M2])

`p` is a Gaussian(mu [N,D], cov [N,D,D]) or DiagonalGaussian(mu, var [N,D]).  Multi-output fan-out follows upstream
:190-247 (only one triangle of kernel pairs is computed, the other is its transpose)."""
from __future__ import annotations

from dataclasses import dataclass

import torch

from gpflowpilco_b200 import ops
from gpflowpilco_b200.models.core import (InducingPoints, MultioutputKernel, SeparateIndependentInducingVariables,
                                          SharedIndependentInducingVariables, SquaredExponential)

__all__ = ("kernel_expectation", "Gaussian", "DiagonalGaussian", "unpack_multioutput")


@dataclass
class Gaussian:
  mu: torch.Tensor
  cov: torch.Tensor


@dataclass
class DiagonalGaussian:
  mu: torch.Tensor
  cov: torch.Tensor      # variances [N,D]


def _moments(p):
  if isinstance(p, DiagonalGaussian):
    return p.mu, torch.diag_embed(p.cov)
  return p.mu, p.cov


def unpack_multioutput(kernel, inducing_variable):
  """(list of latent kernels, list of inducing tensors)  — upstream :41-69."""
  kernels = list(kernel.kernels) if isinstance(kernel, MultioutputKernel) else list(kernel)
  if inducing_variable is None:
    return kernels, None
  if isinstance(inducing_variable, SeparateIndependentInducingVariables):
    zs = [iv.Z for iv in inducing_variable.inducing_variables]
    assert len(zs) == len(kernels)
    return kernels, zs
  if isinstance(inducing_variable, SharedIndependentInducingVariables):
    return kernels, [inducing_variable.inducing_variable.Z] * len(kernels)
  if isinstance(inducing_variable, InducingPoints):
    return kernels, [inducing_variable.Z] * len(kernels)
  return kernels, list(inducing_variable)


def _slice(k: SquaredExponential, mu, cov, Z=None):
  if k.active_dims is None:
    return mu, cov, Z
  idx = list(k.active_dims)
  return mu[..., idx], cov[..., idx, :][..., :, idx], (None if Z is None else Z[..., idx])


def _on_separate_dims(k1: SquaredExponential, k2: SquaredExponential) -> bool:
  """gpflow.kernels.Kernel.on_separate_dims: both kernels name their active dimensions and share none"""
  if k1.active_dims is None or k2.active_dims is None:
    return False
  return not (set(k1.active_dims) & set(k2.active_dims))


def _Z(z):
  return z.Z if isinstance(z, InducingPoints) else z


def _single(p, k1, z1, k2=None, z2=None):
  mu, cov = _moments(p)
  if z1 is None:
    return k1.variance.to(mu.device).expand(mu.shape[0])                     # eKff of a stationary kernel
  if k2 is None:
    m, S, Z = _slice(k1, mu, cov, _Z(z1))
    return ops.ekxz(m, S, Z, k1.ell(m.shape[-1]).to(m.device), float(k1.variance))
  if _on_separate_dims(k1, k2):
    # upstream :85-94: no joint expectation is required when the two kernels read disjoint state dimensions AND those dimensions
    # are independent under p (DiagonalGaussian): E[k1(Z1,x) k2(x,Z2)] = E[k1(Z1,x)] E[k2(x,Z2)]
    if isinstance(p, DiagonalGaussian):
      return _single(p, k1, z1)[:, :, None] * _single(p, k2, z2)[:, None, :]
    raise NotImplementedError("The expectation over two kernels only has an analytical implementation if both kernels "
                              "have the same active features.")
  if (k1.active_dims or None) != (k2.active_dims or None):
    raise NotImplementedError("The expectation over two kernels only has an analytical implementation if both kernels "
                              "have the same active features.")            # upstream :91-94
  m, S, Z1 = _slice(k1, mu, cov, _Z(z1))
  same = (k1 is k2) and (_Z(z1) is _Z(z2))
  if same:
    return ops.ekzxkxz(m, S, Z1, k1.ell(m.shape[-1]).to(m.device), float(k1.variance))
  _, _, Z2 = _slice(k2, mu, cov, _Z(z2))
  D = m.shape[-1]
  return ops.ekzxkxz(m, S, Z1, k1.ell(D).to(m.device), float(k1.variance), Z2, k2.ell(D).to(m.device), float(k2.variance))


def kernel_expectation(p, obj1, obj2=None):
  k1, z1 = obj1 if isinstance(obj1, tuple) else (obj1, None)
  k2, z2 = obj2 if isinstance(obj2, tuple) else (obj2, None)
  multi = isinstance(k1, (MultioutputKernel, list))
  if not multi:
    return _single(p, k1, z1, k2, z2)
  K1, Z1 = unpack_multioutput(k1, z1)
  if Z1 is None:
    return torch.stack([_single(p, k, None) for k in K1], dim=-1)                                  # [N,L]
  if k2 is None:
    return torch.stack([_single(p, k, z) for k, z in zip(K1, Z1)], dim=-1)                         # [N,M,L]
  K2, Z2 = unpack_multioutput(k2, z2)
  symmetric = (k1 is k2) and (z1 is z2)
  blocks = [[None] * len(K2) for _ in K1]
  for i, (ka, za) in enumerate(zip(K1, Z1)):
    for j, (kb, zb) in enumerate(zip(K2, Z2)):
      if symmetric and j < i:
        blocks[i][j] = blocks[j][i].transpose(-1, -2)                                               # adjoint reuse, upstream :238-244
      else:
        blocks[i][j] = _single(p, ka, za, kb, zb)
  return torch.stack([torch.stack(row, dim=-2) for row in blocks], dim=-4)                          # [N,L1,M1,L2,M2]
