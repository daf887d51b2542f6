"""Shared generators for the tests (seeded; mirror upstream tests/utils.py:70-121 semantics)."""
import math

import numpy as np
import torch

DTYPE = torch.float64


def generate_covariance(ndims, sample_shape=(), scale=None, gen=None):
  """Random covariance: eigenvalues -log U, random orthogonal basis, optionally rescaled to
  (scale^2 x correlation)   (upstream tests/utils.py:99-121)."""
  shape = list(sample_shape)
  eig = -torch.log(torch.rand(shape + [1, ndims], dtype=DTYPE, generator=gen))
  U = torch.linalg.svd(torch.randn(shape + [ndims, ndims], dtype=DTYPE, generator=gen))[0]
  sq = torch.sqrt(eig) * U
  cov = sq @ sq.transpose(-1, -2)
  if scale is not None:
    istd = torch.rsqrt(torch.diagonal(cov, dim1=-2, dim2=-1))
    cov = (scale ** 2) * cov * istd[..., None] * istd[..., None, :]
  return cov


def draw_samples_mvn(mu, cov, num, gen=None):
  """[num, *batch, d] draws (upstream tests/utils.py:70-81)."""
  sq = torch.linalg.cholesky(cov)
  rvs = torch.randn([num] + list(cov.shape[:-2]) + [mu.shape[-1]], dtype=DTYPE, generator=gen)
  return mu + (sq @ rvs.unsqueeze(-1)).squeeze(-1)


def empirical_covariance(a, b=None):
  a = a - a.mean(0, keepdim=True)
  b = a if b is None else b - b.mean(0, keepdim=True)
  return torch.einsum("ni,nj->ij", a, b) / (len(a) - 1)


def mc_close(a, b, num_samples):
  """Upstream criterion tests/utils.py:66-67 with rtol = 10/sqrt(S): effectively |a-b| <= 10/sqrt(S) absolute."""
  tol = 10.0 * num_samples ** -0.5
  return bool(torch.all(torch.abs(a - b) <= tol + 1e-8 * torch.abs(b)))


def log_uniform(shape, lo, hi, gen=None):
  return torch.exp(math.log(lo) + (math.log(hi) - math.log(lo)) * torch.rand(shape, dtype=DTYPE, generator=gen))
