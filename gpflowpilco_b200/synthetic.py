"""Seeded synthetic workloads of BASELINE.json's configs (host-side numpy; no GPU, no oracle dependency).

Shapes and distributions follow SURVEY.md §8(d); upstream sources of the constants are cited per function.
All builders return plain dicts of float64 numpy arrays so that the CUDA path, the oracle and the fixtures
consume identical inputs.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def generate_covariance(rng: np.random.Generator, ndims: int, batch: int, scale: float) -> np.ndarray:
  """Random covariances: eigenvalues -log U, Haar-ish orthogonal basis, rescaled to scale^2 x correlation
  (same construction as upstream tests/utils.py:99-121)."""
  eig = -np.log(rng.random((batch, 1, ndims)))
  U = np.linalg.svd(rng.standard_normal((batch, ndims, ndims)))[0]
  sq = np.sqrt(eig) * U
  cov = sq @ np.swapaxes(sq, -1, -2)
  istd = 1.0 / np.sqrt(np.einsum("bii->bi", cov))
  return (scale ** 2) * cov * istd[:, :, None] * istd[:, None, :]


def se_kernel(A: np.ndarray, B: np.ndarray, ell: np.ndarray, var: float) -> np.ndarray:
  d = (A / ell)[:, None, :] - (B / ell)[None, :, :]
  return var * np.exp(-0.5 * (d * d).sum(-1))


def config2_batched_mm_predict(N: int = 8192, M: int = 1000, D: int = 6, E: int = 4, seed: int = 0) -> Dict[str, np.ndarray]:
  """BASELINE config #2: N Gaussian inputs through E independent exact SE-ARD GPs on M training points
  (maths of upstream moment_matching/models.py:44-111 per output)."""
  rng = np.random.default_rng(seed)
  X = rng.random((M, D))
  ell = np.exp(rng.uniform(math.log(0.3), math.log(3.0), size=(E, D)))
  var = np.full(E, 0.89 ** 2)
  noise = 1e-2
  Y = np.empty((M, E))
  for e in range(E):
    K = se_kernel(X, X, ell[e], var[e]) + 1e-10 * np.eye(M)
    Y[:, e] = np.linalg.cholesky(K) @ rng.standard_normal(M) + math.sqrt(noise) * rng.standard_normal(M)
  mean_const = 0.1 * rng.standard_normal(E)
  mu = rng.random((N, D))
  cov = generate_covariance(rng, D, N, 0.1)
  return dict(X=X, Y=Y, lengthscales=ell, variance=var, noise_variance=np.full(E, noise), mean_const=mean_const,
              mu=mu, cov=cov)


def config3_psi2_stress(N: int = 1024, M: int = 2048, D: int = 8, seed: int = 0) -> Dict[str, np.ndarray]:
  """BASELINE config #3: Psi2 with two kernels / two inducing sets (the branch upstream tests/test_kernel_expectation.py:50-93
  exercises), inducing points half near the bulk of the inputs, half uniform (:63-66)."""
  rng = np.random.default_rng(seed)
  mu = rng.standard_normal((N, D))
  cov = generate_covariance(rng, D, N, 0.1)

  def inducing():
    return np.concatenate([math.sqrt(0.1) * rng.standard_normal((M // 2, D)), rng.random((M - M // 2, D))], 0)

  Z1, Z2 = inducing(), inducing()
  ell1 = np.exp(rng.uniform(math.log(0.1), math.log(10.0), size=D))
  ell2 = np.exp(rng.uniform(math.log(0.1), math.log(10.0), size=D))
  return dict(mu=mu, cov=cov, Z1=Z1, Z2=Z2, lengthscales1=ell1, lengthscales2=ell2, variance1=0.89 ** 2, variance2=0.89 ** 2)


def random_svgp(L: int, M: int, D: int, seed: int = 0, whiten: bool = True, P: int | None = None,
                ell_range=(0.5, 2.0), z_scale: float = 1.0) -> Dict[str, np.ndarray]:
  """A generic well-conditioned multi-output SVGP parameter set for parity tests."""
  rng = np.random.default_rng(seed)
  Z = z_scale * rng.standard_normal((L, M, D))
  ell = np.exp(rng.uniform(math.log(ell_range[0]), math.log(ell_range[1]), size=(L, D)))
  var = 0.5 + rng.random(L)
  q_mu = rng.standard_normal((M, L))
  A = 0.3 * rng.standard_normal((L, M, M)) / math.sqrt(M)
  q_sqrt = np.tril(A) + 0.2 * np.eye(M)[None]
  out = dict(Z=Z, lengthscales=ell, variance=var, q_mu=q_mu, q_sqrt=q_sqrt, whiten=whiten,
             mean_const=rng.standard_normal(L if P is None else P))
  if P is not None:
    W = rng.random((P, L))
    out["W"] = W / np.linalg.norm(W, axis=-1, keepdims=True)
  return out
