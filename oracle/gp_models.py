"""Oracle: GP model parameter structs, predict paths and exact moment matching through them.

TEST INFRASTRUCTURE — see oracle/__init__.py.

Follows
  * _mm_gauss_gpr        gpflow_pilco/moment_matching/models.py:44-111
  * _mm_gauss_svgp_so    gpflow_pilco/moment_matching/models.py:129-197
  * _mm_gauss_svgp_mo    gpflow_pilco/moment_matching/models.py:200-299
  * KernelRegressor / InverseLinkWrapper rules   gpflow_pilco/moment_matching/models.py:27-41
  * predict paths        gpflow_pilco/models/core.py:61-71, gpflow_pilco/models/svgp.py:38-39,
                         gpflow_pilco/models/mean_functions.py:28-38; GPflow's SVGP/GPR.predict_f and
                         covariances.Kuu (+ default_jitter 1e-6) are restated.
The triangular-solve ("reference") form is kept on purpose: Luu^{-1} Q Luu^{-T} as at models.py:224-226.
``mm_sparse_reassociated`` is the O(M^2) re-association the CUDA path uses, kept here
to quantify the conditioning gap between the two forms.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from oracle.moments import GaussianMatch, GaussianMoments, mm_chain, mm_squash, squash
from oracle.psi_stats import (DTYPE, SEKernel, eKff, eKff_list, eKfu_list, eKuffu_list, eKxz,
                              eKzxKxz, tri_solve)

DEFAULT_JITTER = 1e-6   # gpflow.config.default_jitter()


@dataclass(eq=False)
class GPRModel:
  """gpflow.models.GPR stand-in (single SE kernel, Gaussian likelihood, constant mean)."""
  kernel: SEKernel
  X: torch.Tensor                 # [M,D]
  Y: torch.Tensor                 # [M,1]
  noise_variance: torch.Tensor
  mean_const: Optional[torch.Tensor] = None    # [1]; None == Zero mean function


@dataclass(eq=False)
class SVGPModel:
  """gpflow.models.SVGP stand-in.

  Single output: ``kernels`` has one entry, ``Z`` one [M,D] tensor.  Multi output (SeparateIndependent or
  LinearCoregionalization with mixing ``W`` [P,L]) : L kernels, L inducing sets
  (gpflow_pilco/models/svgp.py:72-121 builds exactly this)."""
  kernels: List[SEKernel]
  Z: List[torch.Tensor]           # L x [M,D]
  q_mu: torch.Tensor              # [M,L]
  q_sqrt: torch.Tensor            # [L,M,M]
  whiten: bool = True
  mean_const: Optional[torch.Tensor] = None    # [P]
  W: Optional[torch.Tensor] = None             # [P,L] LinearCoregionalization
  multi_output: bool = True

  @property
  def num_latent(self) -> int:
    return len(self.kernels)


def Kuu(kern: SEKernel, Z: torch.Tensor, jitter: float = DEFAULT_JITTER) -> torch.Tensor:
  return kern.K(Z) + jitter * torch.eye(Z.shape[0], dtype=DTYPE)


def _mean_fn(mean_const, N: int):
  return 0.0 if mean_const is None else mean_const.reshape(1, -1).expand(N, -1)


# ---------------------------------------------------------------------------------------------
# predict paths (sample path of the rollout and the Monte-Carlo pin tests)
# ---------------------------------------------------------------------------------------------
def gpr_predict_f(model: GPRModel, Xnew: torch.Tensor):
  """GPflow GPR.predict_f(full_cov=False): mean [N,1], var [N,1]."""
  k = model.kernel
  Kmm = k.K(model.X) + model.noise_variance * torch.eye(model.X.shape[0], dtype=DTYPE)
  Lm = torch.linalg.cholesky(Kmm)
  Kmn = k.K(model.X, Xnew)
  err = model.Y - _mean_fn(model.mean_const, model.X.shape[0])
  A = tri_solve(Lm, Kmn)
  mean = A.T @ tri_solve(Lm, err) + _mean_fn(model.mean_const, Xnew.shape[0])
  var = k.variance - (A * A).sum(0)
  return mean, var.unsqueeze(-1)


def svgp_predict_f(model: SVGPModel, Xnew: torch.Tensor, full_output_cov: bool = False):
  """GPflow SVGP.predict_f(full_cov=False): mean [N,P]; var [N,P] or [N,P,P]."""
  mus, vars_ = [], []
  for l, (k, Z) in enumerate(zip(model.kernels, model.Z)):
    Lu = torch.linalg.cholesky(Kuu(k, Z))
    A = tri_solve(Lu, k.K(Z, Xnew))                                    # [M,N]
    q_mu = model.q_mu[:, l:l + 1]
    q_sqrt = torch.tril(model.q_sqrt[l])
    if not model.whiten:
      A = tri_solve(Lu, A, adjoint=True)
    mus.append((A.T @ q_mu).squeeze(-1))
    LTA = q_sqrt.T @ A
    vars_.append(k.variance - (tri_solve(Lu, k.K(Z, Xnew)) ** 2).sum(0) + (LTA * LTA).sum(0))
  mu = torch.stack(mus, -1)
  var = torch.stack(vars_, -1)
  if model.W is not None:
    mu = mu @ model.W.T
    cov = torch.einsum("pl,nl,ql->npq", model.W, var, model.W)
    var = cov if full_output_cov else torch.diagonal(cov, dim1=-2, dim2=-1)
  elif full_output_cov:
    var = torch.diag_embed(var)
  return mu + _mean_fn(model.mean_const, Xnew.shape[0]), var


def kernel_regressor(model: SVGPModel, x: torch.Tensor) -> torch.Tensor:
  """KernelRegressor.__call__ = predict_f(x)[0]   (models/core.py:61-63)."""
  return svgp_predict_f(model, x)[0]


# ---------------------------------------------------------------------------------------------
# moment matching
# ---------------------------------------------------------------------------------------------
def _cross_preinv(m, S, kern: SEKernel, Z, weights, psi1):
  """Sum_m w[m] psi1[n,m] (S_n + Lambda)^{-1} (z_m - mu_n)  -> [N,D]   (models.py:88-98,172-181)."""
  x1 = kern.slice(m)
  dX = kern.slice(Z).unsqueeze(0) - x1.unsqueeze(-2)                   # [N,M,D]
  Sxx = kern.slice_cov(S)
  D = Sxx.shape[-1]
  Vs = torch.linalg.cholesky(Sxx + torch.diag(kern.ell(D) ** 2))
  iV_dXt = torch.cholesky_solve(dX.transpose(-1, -2), Vs)              # [N,D,M]
  return (weights.reshape(1, 1, -1) * psi1.unsqueeze(-2) * iV_dXt).sum(-1)


def mm_gpr(x: GaussianMoments, model: GPRModel, full_output_cov: bool = True,
           model_uncertainty: bool = True, jitter: float = 0.0) -> GaussianMatch:
  """models.py:44-111.  Single-output."""
  k, X = model.kernel, model.X
  Y = model.Y - _mean_fn(model.mean_const, X.shape[0])                 # :53-56
  m, S = x.mean(), x.covariance()
  N = m.shape[0]
  e_ff = eKff(m, k)                                                    # [N]
  e_fu = eKxz(m, S, k, X)                                              # [N,M]
  e_uffu = eKzxKxz(m, S, k, X)                                         # [N,M,M]

  Kyy = k.K(X) + model.noise_variance * torch.eye(X.shape[0], dtype=DTYPE)   # :66-68
  Lyy = torch.linalg.cholesky(Kyy)
  iL_y = tri_solve(Lyy, Y)                                             # [M,1]
  T = tri_solve(Lyy, tri_solve(Lyy, e_uffu).transpose(-1, -2))         # :71-72
  beta = tri_solve(Lyy, iL_y, adjoint=True)                            # [M,1]
  f1 = e_fu @ beta                                                     # [N,1]

  f2 = (iL_y.T @ (T @ iL_y))                                           # [N,1,1]
  Sff = f2 - f1.unsqueeze(-1) * f1.unsqueeze(-2)
  if model_uncertainty:
    e_cov = e_ff - torch.diagonal(T, dim1=-2, dim2=-1).sum(-1)         # :86
    Sff = Sff + e_cov[:, None, None]

  cross = _cross_preinv(m, S, k, X, beta.squeeze(-1), e_fu).unsqueeze(-1)     # [N,D,1]
  f1 = f1 + _mean_fn(model.mean_const, N)
  Sff = Sff + jitter * torch.eye(1, dtype=DTYPE)
  if not full_output_cov:
    Sff = torch.diag_embed(torch.diagonal(Sff, dim1=-2, dim2=-1))
  return GaussianMatch(x, GaussianMoments(f1, Sff, True), (cross, True))


def mm_svgp(x: GaussianMoments, model: SVGPModel, full_output_cov: bool = True,
            model_uncertainty: bool = True, jitter: float = 0.0) -> GaussianMatch:
  """Dispatch like models.py:114-126."""
  fn = mm_svgp_mo if model.multi_output else mm_svgp_so
  return fn(x, model, full_output_cov, model_uncertainty, jitter)


def mm_svgp_so(x: GaussianMoments, model: SVGPModel, full_output_cov: bool = True,
               model_uncertainty: bool = True, jitter: float = 0.0) -> GaussianMatch:
  """models.py:129-197."""
  assert model.num_latent == 1
  k, Z = model.kernels[0], model.Z[0]
  m, S = x.mean(), x.covariance()
  N = m.shape[0]
  e_ff = eKff(m, k)
  e_fu = eKxz(m, S, k, Z)
  e_uffu = eKzxKxz(m, S, k, Z)

  Luu = torch.linalg.cholesky(Kuu(k, Z))                               # :145-146
  T = tri_solve(Luu, tri_solve(Luu, e_uffu).transpose(-1, -2))         # :147-148
  iL_qmu = model.q_mu                                                  # [M,1]
  iL_qsqrt = torch.tril(model.q_sqrt[0])
  if not model.whiten:
    iL_qmu = tri_solve(Luu, iL_qmu)
    iL_qsqrt = tri_solve(Luu, iL_qsqrt)
  beta = tri_solve(Luu, iL_qmu, adjoint=True)
  f1 = e_fu @ beta

  f2 = iL_qmu.T @ (T @ iL_qmu)
  Sff = f2 - f1.unsqueeze(-1) * f1.unsqueeze(-2)
  if model_uncertainty:
    Li_qcov_LiT = iL_qsqrt @ iL_qsqrt.T
    e_cov = e_ff - torch.diagonal(T, dim1=-2, dim2=-1).sum(-1) + (T * Li_qcov_LiT).sum((-1, -2))
    Sff = Sff + e_cov[:, None, None]

  cross = _cross_preinv(m, S, k, Z, beta.squeeze(-1), e_fu).unsqueeze(-1)
  f1 = f1 + _mean_fn(model.mean_const, N)
  Sff = Sff + jitter * torch.eye(1, dtype=DTYPE)
  return GaussianMatch(x, GaussianMoments(f1, Sff, True), (cross, True))


def mm_svgp_mo(x: GaussianMoments, model: SVGPModel, full_output_cov: bool = True,
               model_uncertainty: bool = True, jitter: float = 0.0) -> GaussianMatch:
  """models.py:200-299 (SeparateIndependent, or LinearCoregionalization when model.W is given)."""
  ks, Zs = model.kernels, model.Z
  L = len(ks)
  m, S = x.mean(), x.covariance()
  N = m.shape[0]
  e_ff = torch.diag_embed(eKff_list(m, ks))                            # [N,L,L]    :210
  e_fu = eKfu_list(m, S, ks, Zs)                                       # [N,M,L]
  e_uffu = eKuffu_list(m, S, ks, Zs)                                   # [N,L,M,L,M]

  Luu = torch.stack([torch.linalg.cholesky(Kuu(k, z)) for k, z in zip(ks, Zs)])   # [L,M,M]  :216-217
  # T[n,a,:,b,:] = Luu_a^{-1} Q_ab Luu_b^{-T}   (the two batched triangular solves of :219-226)
  T = torch.empty_like(e_uffu)
  for a in range(L):
    for b in range(L):
      half = tri_solve(Luu[a], e_uffu[:, a, :, b, :])                  # [N,M,M]
      T[:, a, :, b, :] = tri_solve(Luu[b], half.transpose(-1, -2)).transpose(-1, -2)

  iL_qmu = model.q_mu.T.unsqueeze(-1)                                  # [L,M,1]   :228
  iL_qsqrt = torch.tril(model.q_sqrt)
  if not model.whiten:
    iL_qmu = tri_solve(Luu, iL_qmu)
    iL_qsqrt = tri_solve(Luu, iL_qsqrt)
  beta = tri_solve(Luu, iL_qmu, adjoint=True)                          # [L,M,1]   :235
  f1 = (e_fu.transpose(-1, -2) * beta[..., 0]).sum(-1)                 # [N,L]     :236

  w = iL_qmu[..., 0]                                                   # [L,M]
  blk = torch.stack([T[:, l, :, l, :] for l in range(L)], dim=-3)      # [N,L,M,M]
  coreg = model.W is not None
  if full_output_cov or coreg:
    f2 = torch.einsum("ai,naibj,bj->nab", w, T, w)                     # :245-247
    Sff = f2 - f1.unsqueeze(-1) * f1.unsqueeze(-2)
  else:
    Sff = torch.einsum("li,nlij,lj->nl", w, blk, w) - f1 ** 2          # :250-252

  if model_uncertainty:
    Li_qcov_Lit = iL_qsqrt @ iL_qsqrt.transpose(-1, -2)                # [L,M,M]
    trace = torch.diagonal(blk, dim1=-2, dim2=-1).sum(-1)              # [N,L]
    mm = (blk * Li_qcov_Lit).sum((-1, -2))
    if full_output_cov or coreg:
      Sff = Sff + e_ff + torch.diag_embed(mm - trace)                  # :258-259
    else:
      Sff = Sff + torch.diagonal(e_ff, dim1=-2, dim2=-1) + mm - trace

  cross = torch.stack([_cross_preinv(m, S, ks[l], Zs[l], beta[l, :, 0], e_fu[..., l])
                       for l in range(L)], dim=-1)                     # [N,D,L]   :264-277

  if coreg:                                                            # :279-286
    f1 = f1 @ model.W.T
    cross = cross @ model.W.T
    if full_output_cov:
      Sff = model.W @ Sff @ model.W.T
    else:
      Sff = (model.W * (model.W @ Sff.transpose(-1, -2))).sum(-1)      # [N,P] = diag(W Sff W^T)  (:286)
  f1 = f1 + _mean_fn(model.mean_const, N)                              # :288-289
  if Sff.ndim == 3:
    Sff = Sff + jitter * torch.eye(Sff.shape[-1], dtype=DTYPE)
    if not full_output_cov:
      Sff = torch.diag_embed(torch.diagonal(Sff, dim1=-2, dim2=-1))
  else:
    Sff = torch.diag_embed(Sff + jitter)                               # LinearOperatorDiag, :296
  return GaussianMatch(x, GaussianMoments(f1, Sff, True), (cross, True))


def mm_kernel_regressor(x: GaussianMoments, model: SVGPModel, **kw) -> GaussianMatch:
  """KernelRegressor => model_uncertainty forced False   (models.py:34-41)."""
  assert not kw.pop("model_uncertainty", False)
  return mm_svgp(x, model, model_uncertainty=False, **kw)


def mm_policy(x: GaussianMoments, model: SVGPModel, scale: float, shift: float = -0.5) -> GaussianMatch:
  """InverseLinkWrapper(KernelRegressor(svgp), invlink=Chain[Scale,Shift,NormalCDF])
  -> Chain(invlink, model)   (models.py:27-31)."""
  return mm_chain(x, [lambda s: mm_squash(s, scale, shift), lambda s: mm_kernel_regressor(s, model)])


def policy_sample_path(model: SVGPModel, e: torch.Tensor, scale: float, shift: float = -0.5) -> torch.Tensor:
  """InverseLinkWrapper.__call__ on tensors: invlink(predict_f(e)[0])   (models/core.py:66-71)."""
  return squash(kernel_regressor(model, e), scale, shift)


# ---------------------------------------------------------------------------------------------
# O(M^2) re-association used by the CUDA path — for the conditioning study
# ---------------------------------------------------------------------------------------------
def sparse_weights(model: SVGPModel, model_uncertainty: bool = True):
  """beta_l = Kuu_l^{-1} m_l and C_l = beta_l beta_l^T - B_l with B_l = Kuu^{-1} - Kuu^{-1} S_l Kuu^{-1}."""
  betas, Cs = [], []
  for l, (k, Z) in enumerate(zip(model.kernels, model.Z)):
    Lu = torch.linalg.cholesky(Kuu(k, Z))
    w = model.q_mu[:, l:l + 1]
    R = torch.tril(model.q_sqrt[l])
    if not model.whiten:
      w = tri_solve(Lu, w)
      R = tri_solve(Lu, R)
    beta = tri_solve(Lu, w, adjoint=True)
    C = beta @ beta.T
    if model_uncertainty:
      M = Z.shape[0]
      inner = torch.eye(M, dtype=DTYPE) - R @ R.T
      C = C - tri_solve(Lu, tri_solve(Lu, inner, adjoint=True).T, adjoint=True)
    betas.append(beta[:, 0])
    Cs.append(C)
  return torch.stack(betas), torch.stack(Cs)


def mm_sparse_reassociated(x: GaussianMoments, model: SVGPModel, model_uncertainty: bool = True,
                           jitter: float = 0.0) -> GaussianMatch:
  ks, Zs = model.kernels, model.Z
  L = len(ks)
  m, S = x.mean(), x.covariance()
  N = m.shape[0]
  beta, C = sparse_weights(model, model_uncertainty)
  e_fu = eKfu_list(m, S, ks, Zs)
  f1 = torch.einsum("nml,lm->nl", e_fu, beta)
  Sff = torch.zeros(N, L, L, dtype=DTYPE)
  for a in range(L):
    Qaa = eKzxKxz(m, S, ks[a], Zs[a])
    Sff[:, a, a] = (Qaa * C[a]).sum((-1, -2)) + (ks[a].variance if model_uncertainty else 0.0)
    for b in range(a + 1, L):
      Qab = eKzxKxz(m, S, ks[a], Zs[a], ks[b], Zs[b])
      Sff[:, a, b] = Sff[:, b, a] = torch.einsum("i,nij,j->n", beta[a], Qab, beta[b])
  Sff = Sff - f1.unsqueeze(-1) * f1.unsqueeze(-2) + jitter * torch.eye(L, dtype=DTYPE)
  cross = torch.stack([_cross_preinv(m, S, ks[l], Zs[l], beta[l], e_fu[..., l]) for l in range(L)], -1)
  if model.W is not None:
    f1, cross, Sff = f1 @ model.W.T, cross @ model.W.T, model.W @ Sff @ model.W.T
  f1 = f1 + _mean_fn(model.mean_const, N)
  return GaussianMatch(x, GaussianMoments(f1, Sff, True), (cross, True))
