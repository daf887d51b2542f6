"""Oracle: Gaussian expectations of squared-exponential kernels (Psi statistics).

TEST INFRASTRUCTURE — see oracle/__init__.py.  float64 torch on CPU.

Follows
  * eKzxKxz / ``_E``         : gpflow_pilco/utils/kernel_expectation.py:72-187
  * list fan-out             : gpflow_pilco/utils/kernel_expectation.py:190-247
  * eKff, eKxz (GPflow's)    : not in the tree; restated from the published formula,
                               call sites gpflow_pilco/moment_matching/models.py:61-62,140-141,211-212
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

DTYPE = torch.float64


@dataclass(eq=False)
class SEKernel:
  """Squared-exponential (ARD) kernel parameters; stands in for gpflow.kernels.SquaredExponential."""
  variance: torch.Tensor            # scalar
  lengthscales: torch.Tensor        # [D] (ARD) or scalar
  active_dims: Optional[Tuple[int, ...]] = None

  def __post_init__(self):
    self.variance = torch.as_tensor(self.variance, dtype=DTYPE)
    self.lengthscales = torch.as_tensor(self.lengthscales, dtype=DTYPE)

  def ell(self, ndims: int) -> torch.Tensor:
    ls = self.lengthscales
    return ls if ls.ndim else ls.expand(ndims)

  def slice(self, x: torch.Tensor) -> torch.Tensor:
    return x if self.active_dims is None else x[..., list(self.active_dims)]

  def slice_cov(self, S: torch.Tensor) -> torch.Tensor:
    if self.active_dims is None:
      return S
    idx = list(self.active_dims)
    return S[..., idx, :][..., :, idx]

  def K(self, A: torch.Tensor, B: Optional[torch.Tensor] = None) -> torch.Tensor:
    """k(A, B) = variance * exp(-0.5 |(a-b)/ell|^2)."""
    A = self.slice(A)
    B = A if B is None else self.slice(B)
    ls = self.ell(A.shape[-1])
    d = (A / ls).unsqueeze(-2) - (B / ls).unsqueeze(-3)
    return self.variance * torch.exp(-0.5 * (d * d).sum(-1))


def tri_solve(L: torch.Tensor, B: torch.Tensor, adjoint: bool = False) -> torch.Tensor:
  """L^{-1} B (or L^{-T} B) for lower-triangular L, broadcasting batch dims."""
  if adjoint:
    return torch.linalg.solve_triangular(L.transpose(-1, -2), B, upper=True)
  return torch.linalg.solve_triangular(L, B, upper=False)


def eKff(mu: torch.Tensor, kern: SEKernel) -> torch.Tensor:
  """E[k(x,x)] = variance, shape [N] (stationary kernel)."""
  return kern.variance.expand(mu.shape[0])


def eKxz(mu: torch.Tensor, cov: torch.Tensor, kern: SEKernel, Z: torch.Tensor) -> torch.Tensor:
  """Psi1[n,m] = E_{x~N(mu_n,cov_n)}[k(x, z_m)]   (GPflow's (Gaussian, SE, InducingPoints) rule).

  variance * |Lambda|^{1/2} |Lambda+cov_n|^{-1/2} exp(-1/2 (z_m-mu_n)^T (Lambda+cov_n)^{-1} (z_m-mu_n)).
  """
  m = kern.slice(mu)
  S = kern.slice_cov(cov)
  z = kern.slice(Z)
  ls = kern.ell(m.shape[-1])
  chol = torch.linalg.cholesky(S + torch.diag(ls ** 2))               # [N,D,D]
  diffs = (z.unsqueeze(0) - m.unsqueeze(1)).transpose(-1, -2)          # [N,D,M]
  half = tri_solve(chol, diffs)
  maha = (half * half).sum(-2)                                         # [N,M]
  det_ratio = ls.prod() / torch.diagonal(chol, dim1=-2, dim2=-1).prod(-1)
  return kern.variance * det_ratio[:, None] * torch.exp(-0.5 * maha)


def eKzxKxz(mu: torch.Tensor,
            cov: torch.Tensor,
            kern1: SEKernel,
            Z1: torch.Tensor,
            kern2: Optional[SEKernel] = None,
            Z2: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Psi2[n,i,j] = E[k1(z1_i, x) k2(x, z2_j)], shape [N,M1,M2].

  Restates gpflow_pilco/utils/kernel_expectation.py:96-187 step by step, including its
  same-kernel / same-feature fast branches (:96-97,114-119,145-148,168-174) and the generic
  branch (:175-185).  ``kern2 is None`` / ``Z2 is None`` mean "same object" as in the reference.
  """
  same_kern = kern2 is None or kern2 is kern1
  same_feat = Z2 is None or Z2 is Z1
  if kern2 is None:
    kern2 = kern1
  if Z2 is None:
    Z2 = Z1
  if (kern1.active_dims is not None or kern2.active_dims is not None) and \
     tuple(kern1.active_dims or ()) != tuple(kern2.active_dims or ()):
    # kernel_expectation.py:91-94
    raise NotImplementedError("both kernels must act on the same active features")

  mx = kern1.slice(mu)                                                 # :99
  Sxx = kern1.slice_cov(cov)                                           # :100-103
  N, D = mx.shape

  V1 = kern1.ell(D) ** 2                                               # :109
  z1 = kern1.slice(Z1)
  iV1_z1 = z1 / V1
  V2 = V1 if same_kern else kern2.ell(D) ** 2                          # :114
  z2 = z1 if same_feat else kern2.slice(Z2)
  iV2_z2 = iV1_z1 if (same_kern and same_feat) else z2 / V2

  V = 0.5 * V1 if same_kern else (V1 * V2) / (V1 + V2)                 # :119

  S = Sxx + torch.diag(V)                                              # :125
  L = torch.linalg.cholesky(S)
  half_logdet = torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)
  determinant = torch.sqrt(V.prod()) * torch.exp(-half_logdet)         # :127-130  [N]

  iL_mu = tri_solve(L, mx.unsqueeze(-1))                               # [N,D,1]
  iL_z1 = tri_solve(L, (V * iV1_z1).T.unsqueeze(0).expand(N, -1, -1))  # [N,D,M1]  (:138-142)
  z1_iS_z1 = (iL_z1 ** 2).sum(1)
  z1_iS_mu = (iL_z1.transpose(-1, -2) @ iL_mu).squeeze(-1)
  if same_kern and same_feat:
    iL_z2, z2_iS_z2, z2_iS_mu = iL_z1, z1_iS_z1, z1_iS_mu
  else:
    iL_z2 = tri_solve(L, (V * iV2_z2).T.unsqueeze(0).expand(N, -1, -1))
    z2_iS_z2 = (iL_z2 ** 2).sum(1)
    z2_iS_mu = (iL_z2.transpose(-1, -2) @ iL_mu).squeeze(-1)

  z1_iS_z2 = iL_z1.transpose(-1, -2) @ iL_z2                           # [N,M1,M2]
  mu_iS_mu = (iL_mu ** 2).sum(1).unsqueeze(-1)                         # [N,1,1]
  exp_maha = torch.exp(-0.5 * (mu_iS_mu + 2 * z1_iS_z2                 # :162-165
                               + (z1_iS_z1 - 2 * z1_iS_mu).unsqueeze(-1)
                               + (z2_iS_z2 - 2 * z2_iS_mu).unsqueeze(-2)))

  if same_kern:                                                        # :168-174
    ampl2 = kern1.variance ** 2
    sq_iV = torch.rsqrt(V)
    a, b = sq_iV * z1, sq_iV * z2
    d2 = ((a.unsqueeze(1) - b.unsqueeze(0)) ** 2).sum(-1)
    matrix_term = ampl2 * torch.exp(-0.125 * d2)
  else:                                                                # :175-185
    z1_iV1_z1 = (z1 * iV1_z1).sum(-1)
    z2_iV2_z2 = (z2 * iV2_z2).sum(-1)
    z1_q_z1 = (iV1_z1 * V * iV1_z1).sum(-1)
    z2_q_z2 = (iV2_z2 * V * iV2_z2).sum(-1)
    z1_q_z2 = iV1_z1 @ (V * iV2_z2).T
    matrix_term = kern1.variance * kern2.variance * torch.exp(0.5 * (
        2 * z1_q_z2 + (z1_q_z1 - z1_iV1_z1).unsqueeze(-1) + (z2_q_z2 - z2_iV2_z2).unsqueeze(-2)))

  return determinant.reshape(N, 1, 1) * matrix_term * exp_maha         # :187


# ---------------------------------------------------------------------------------------------
# list-of-kernels fan-out  (kernel_expectation.py:190-247)
# ---------------------------------------------------------------------------------------------
def eKff_list(mu, kernels: Sequence[SEKernel]) -> torch.Tensor:
  """[N,L]  (:190-198)"""
  return torch.stack([eKff(mu, k) for k in kernels], dim=-1)


def eKfu_list(mu, cov, kernels: Sequence[SEKernel], Zs: Sequence[torch.Tensor]) -> torch.Tensor:
  """[N,M,L]  (:201-214)"""
  assert len(kernels) == len(Zs)
  return torch.stack([eKxz(mu, cov, k, z) for k, z in zip(kernels, Zs)], dim=-1)


def eKuffu_list(mu, cov, kernels: Sequence[SEKernel], Zs: Sequence[torch.Tensor]) -> torch.Tensor:
  """[N,L,M,L,M]; only one triangle of (kernel,feature) pairs is computed and the other is its adjoint
  (:217-247; the reference orders pairs by Python hash(), here by index — values are identical)."""
  L = len(kernels)
  blocks = [[None] * L for _ in range(L)]
  for a in range(L):
    for b in range(a, L):
      if a == b:
        blocks[a][b] = eKzxKxz(mu, cov, kernels[a], Zs[a])
      else:
        same_k = kernels[a] is kernels[b]
        blocks[a][b] = eKzxKxz(mu, cov, kernels[a], Zs[a],
                               kernels[a] if same_k else kernels[b], Zs[b])
        blocks[b][a] = blocks[a][b].transpose(-1, -2)
  return torch.stack([torch.stack(row, dim=-2) for row in blocks], dim=-4)
