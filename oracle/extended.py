"""Oracle in EXTENDED precision (numpy longdouble: x87 80-bit, 64-bit mantissa, eps ~ 1.1e-19) — the arbiter of the conditioning study.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Upstream forms Luu^-1 Q Luu^-T by two triangular solves per input (gpflow_pilco/moment_matching/models.py:224-226); the CUDA path
contracts Q with weights derived from Kuu^-1 once per model.  The two are the same algebra and differ in float64 by an
amount that grows with cond(Kuu) (SURVEY §7 "hard parts" 1: 4e-9 at cond 4e7, 3e-6 at cond 2e8).  Which of them is closer to the exact
value cannot be told in float64; here the moment-matched mean / variance of a single-output whitened SVGP (upstream models.py:129-197)
is evaluated with every intermediate — kernel matrix, Cholesky factor, Psi statistics, contractions — in longdouble, so that its own
rounding (cond * 1e-19) is three orders below either float64 form.  numpy's LAPACK bindings stop at float64: the Cholesky factorisation
and the triangular solves are written out.
"""
from __future__ import annotations

import numpy as np

LD = np.longdouble


def cholesky(A: np.ndarray) -> np.ndarray:
  """lower Cholesky factor in the dtype of A (row-oriented Cholesky-Crout, vectorised over the row)"""
  n = A.shape[0]
  L = np.zeros_like(A)
  for j in range(n):
    d = A[j, j] - L[j, :j] @ L[j, :j]
    if not d > 0:
      raise np.linalg.LinAlgError("not positive definite")
    L[j, j] = np.sqrt(d)
    if j + 1 < n:
      L[j + 1:, j] = (A[j + 1:, j] - L[j + 1:, :j] @ L[j, :j]) / L[j, j]
  return L


def solve_lower(L: np.ndarray, B: np.ndarray) -> np.ndarray:
  """L^-1 B by forward substitution"""
  X = np.array(B, dtype=L.dtype, copy=True)
  for i in range(L.shape[0]):
    X[i] = (X[i] - L[i, :i] @ X[:i]) / L[i, i]
  return X


def solve_upper_t(L: np.ndarray, B: np.ndarray) -> np.ndarray:
  """L^-T B by back substitution"""
  X = np.array(B, dtype=L.dtype, copy=True)
  for i in range(L.shape[0] - 1, -1, -1):
    X[i] = (X[i] - L[i + 1:, i] @ X[i + 1:]) / L[i, i]
  return X


def _spd_inverse_logdet(A):
  L = cholesky(A)
  Li = solve_lower(L, np.eye(A.shape[0], dtype=A.dtype))
  return Li.T @ Li, 2 * np.log(np.diag(L)).sum()


def mm_svgp_so_extended(mu, cov, Z, ell, var, q_mu, q_sqrt, jitter=1e-6, dtype=LD):
  """Moment-matched mean f1 [N] and variance Sff [N] (model_uncertainty=True) of a single-output WHITENED SVGP with an SE-ARD kernel,
  in `dtype`, in upstream's triangular-solve association (models.py:129-197).  Inputs are float64 arrays: they are the exact data."""
  mu, cov, Z, ell, q_mu, q_sqrt = (np.asarray(a, dtype=dtype) for a in (mu, cov, Z, ell, q_mu, q_sqrt))
  var, jitter = dtype(var), dtype(jitter)
  N, D = mu.shape
  M = Z.shape[0]
  Zs = Z / ell
  d2 = ((Zs[:, None, :] - Zs[None, :, :]) ** 2).sum(-1)
  Kuu = var * np.exp(-d2 / 2) + jitter * np.eye(M, dtype=dtype)
  Lu = cholesky(Kuu)
  w = q_mu.reshape(M, 1)                                   # whitened: u = Luu w
  R = np.tril(q_sqrt)
  lam = ell ** 2
  f1 = np.zeros(N, dtype=dtype)
  Sff = np.zeros(N, dtype=dtype)
  for n in range(N):
    # Psi1 (GPflow eKxz)
    A1, logdet1 = _spd_inverse_logdet(np.diag(lam) + cov[n])
    dz = Z - mu[n]
    psi1 = var * np.exp((np.log(lam).sum() - logdet1) / 2 - np.einsum("md,de,me->m", dz, A1, dz) / 2)
    # Psi2 (upstream _E, same kernel and features: kernel_expectation.py:96-187)
    A2, logdet2 = _spd_inverse_logdet(np.diag(lam / 2) + cov[n])
    zbar = (Z[:, None, :] + Z[None, :, :]) / 2 - mu[n]
    quad = np.einsum("ijd,de,ije->ij", zbar, A2, zbar)
    psi2 = var * var * np.exp((np.log(lam / 2).sum() - logdet2) / 2 - d2 / 4 - quad / 2)
    # upstream's association: T = Luu^-1 Psi2 Luu^-T, f2 = w^T T w, variance term through T as well (models.py:147-176)
    T = solve_lower(Lu, solve_lower(Lu, psi2).T).T
    a = solve_lower(Lu, psi1.reshape(M, 1))
    f1[n] = (a.T @ w)[0, 0]
    f2 = (w.T @ T @ w)[0, 0]
    Su = R @ R.T
    Sff[n] = f2 - f1[n] ** 2 + var + (T * Su).sum() - np.trace(T)
  return f1, Sff
