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


def oracle_svgp(params, model_uncertainty=True):
  """dict from gpflowpilco_b200.synthetic -> oracle.gp_models.SVGPModel."""
  from oracle import gp_models as gm
  from oracle import psi_stats as ps
  L = params["Z"].shape[0]
  ks = [ps.SEKernel(float(params["variance"][l]), torch.as_tensor(params["lengthscales"][l])) for l in range(L)]
  Zs = [torch.as_tensor(params["Z"][l]) for l in range(L)]
  W = params.get("W")
  return gm.SVGPModel(ks, Zs, torch.as_tensor(params["q_mu"]), torch.as_tensor(params["q_sqrt"]),
                      whiten=bool(params["whiten"]), mean_const=torch.as_tensor(params["mean_const"]),
                      W=None if W is None else torch.as_tensor(W))


def cuda_handle(params, model_uncertainty=True, kuu_jitter=1e-6):
  from gpflowpilco_b200 import ops
  dev = torch.device("cuda")
  t = lambda k: None if params.get(k) is None else torch.as_tensor(params[k], dtype=DTYPE, device=dev)
  return ops.GPModelHandle(t("Z"), t("lengthscales"), t("variance"), t("q_mu"), t("q_sqrt"), whiten=bool(params["whiten"]),
                           mean_const=t("mean_const"), W=t("W"), kuu_jitter=kuu_jitter, model_uncertainty=model_uncertainty)


def scaled_close(actual, expected, rel=1e-6, what=""):
  """max |a-e| <= rel * max|e|  — the north-star tolerance (1e-6 relative, FP64 path)."""
  actual = actual.detach().cpu() if hasattr(actual, "cpu") else torch.as_tensor(actual)
  expected = expected.detach().cpu()
  scale = float(expected.abs().max())
  err = float((actual - expected).abs().max())
  assert err <= rel * max(scale, 1e-300), f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rel {err / max(scale, 1e-300):.3e} > {rel:.1e})"
  return err / max(scale, 1e-300)
