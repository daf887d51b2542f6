"""Oracle: decoupled pathwise sampler for a multi-output SVGP (random-Fourier prior + canonical-basis update).

TEST INFRASTRUCTURE — see oracle/__init__.py.

The reference delegates this row to gpflow_sampling >= 0.2 (setup.py:5; NOT in the tree, unpinned):
  drift.generate_paths(num_samples=S, num_bases=F, sample_axis=0)    gpflow_pilco/loops/pilco.py:282-284
  drift.set_temporary_paths(paths); drift(x) == paths(x)            gpflow_pilco/loops/pilco.py:288,
                                                                     gpflow_pilco/models/svgp.py:124-130
so parity for this row is UNPINNED by upstream; this file restates the published algorithm
(Wilson et al. 2020, arXiv:2002.09309 §3) and is the contract the CUDA path is held to:

  phi_{l,i}(x) = sqrt(2 var_l / F) cos(omega_{l,i} . (x / ell_l) + b_{l,i}),  omega ~ N(0,I), b ~ U(0, 2 pi)
  prior_s,l(x) = sum_i w[s,l,i] phi_{l,i}(x),                                 w ~ N(0,I)
  u_{s,l}      = q_mu_l + tril(q_sqrt_l) eps_{s,l}   (then Luu_l u if whitened), eps ~ N(0,I)
  v_{s,l}      = (k_l(Z_l,Z_l) + jitter I)^{-1} (u_{s,l} - Phi_l(Z_l) w_{s,l} - sqrt(jitter) xi_{s,l}),  xi ~ N(0,I)
  f_{s,l}(x)   = c_l + prior_{s,l}(x) + sum_j v[s,l,j] k_l(x, z_{l,j})
  sample_axis=0: row s of x is evaluated on draw s.
All randomness comes from oracle/philox.py streams keyed by the GLOBAL particle index.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from oracle import philox
from oracle.gp_models import DEFAULT_JITTER, SVGPModel
from oracle.psi_stats import DTYPE


@dataclass
class Paths:
  omega: torch.Tensor      # [L,F,D]  unit-normal frequencies (divide inputs by ell)
  phase: torch.Tensor      # [L,F]
  w: torch.Tensor          # [S,L,F]  prior weights
  v: torch.Tensor          # [S,L,M]  canonical-basis (update) weights


def draw_basis(L: int, F: int, D: int, seed: int):
  e = np.arange(L * F * D, dtype=np.uint64).reshape(L, F, D)
  omega = philox.normal(e, philox.STREAM_OMEGA, seed)
  b = 2.0 * np.pi * philox.uniform(np.arange(L * F, dtype=np.uint64).reshape(L, F), philox.STREAM_PHASE, seed)
  return torch.from_numpy(omega), torch.from_numpy(b)


def draw_particle_normals(stream: int, seed: int, first: int, count: int, L: int, K: int) -> torch.Tensor:
  """[count,L,K] normals for global particles first..first+count-1; logical element ((s*L+l)*K+k)."""
  s = np.arange(first, first + count, dtype=np.uint64).reshape(-1, 1, 1)
  l = np.arange(L, dtype=np.uint64).reshape(1, -1, 1)
  k = np.arange(K, dtype=np.uint64).reshape(1, 1, -1)
  return torch.from_numpy(philox.normal((s * np.uint64(L) + l) * np.uint64(K) + k, stream, seed))


def draw_initial_states(m0: torch.Tensor, chol0: torch.Tensor, seed: int, first: int, count: int) -> torch.Tensor:
  """x0_s = m0 + chol(S0) n_s   (p.sample([batch_size]), gpflow_pilco/loops/pilco.py:300-303)."""
  d = m0.shape[-1]
  n = draw_particle_normals(philox.STREAM_X0, seed, first, count, 1, d)[:, 0, :]
  return m0 + n @ chol0.T


def features(model: SVGPModel, omega, phase, x: torch.Tensor) -> torch.Tensor:
  """Phi[l](x): [L,N,F]."""
  out = []
  F = omega.shape[1]
  for l, k in enumerate(model.kernels):
    ls = k.ell(x.shape[-1])
    proj = (x / ls) @ omega[l].T + phase[l]
    out.append(torch.sqrt(2.0 * k.variance / F) * torch.cos(proj))
  return torch.stack(out)


def generate_paths(model: SVGPModel, num_bases: int, seed: int, first: int, count: int,
                   jitter: float = DEFAULT_JITTER) -> Paths:
  L, D, M = model.num_latent, model.Z[0].shape[-1], model.Z[0].shape[0]
  omega, phase = draw_basis(L, num_bases, D, seed)
  w = draw_particle_normals(philox.STREAM_PRIOR_W, seed, first, count, L, num_bases)
  eps = draw_particle_normals(philox.STREAM_U_EPS, seed, first, count, L, M)
  xi = draw_particle_normals(philox.STREAM_UPDATE_XI, seed, first, count, L, M)
  v = torch.empty(count, L, M, dtype=DTYPE)
  for l, (k, Z) in enumerate(zip(model.kernels, model.Z)):
    Kzz = k.K(Z) + jitter * torch.eye(M, dtype=DTYPE)
    Lu = torch.linalg.cholesky(Kzz)
    u = model.q_mu[:, l] + eps[:, l, :] @ torch.tril(model.q_sqrt[l]).T     # [S,M]
    if model.whiten:
      u = u @ Lu.T
    PhiZ = features(model, omega, phase, Z)[l]                               # [M,F]
    err = u - w[:, l, :] @ PhiZ.T - (jitter ** 0.5) * xi[:, l, :]
    v[:, l, :] = torch.cholesky_solve(err.T, Lu).T
  return Paths(omega, phase, w, v)


def evaluate_paths(model: SVGPModel, paths: Paths, x: torch.Tensor) -> torch.Tensor:
  """f[s,l] for row s of x on draw s  ->  [S,L]  (+ constant mean)."""
  S = x.shape[0]
  out = torch.empty(S, model.num_latent, dtype=DTYPE)
  F = paths.omega.shape[1]
  for l, (k, Z) in enumerate(zip(model.kernels, model.Z)):
    ls = k.ell(x.shape[-1])
    proj = (x / ls) @ paths.omega[l].T + paths.phase[l]                       # [S,F]
    phi = torch.sqrt(2.0 * k.variance / F) * torch.cos(proj)
    out[:, l] = (phi * paths.w[:, l, :]).sum(-1) + (k.K(x, Z) * paths.v[:, l, :]).sum(-1)
  if model.mean_const is not None:
    out = out + model.mean_const
  return out
