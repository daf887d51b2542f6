"""GPU parity of the backward kernels against torch.autograd on the float64 oracle (upstream differentiates the same
maths with TensorFlow autodiff, gpflow_pilco/utils/optimizers.py:52-56; there are no upstream gradient vectors).
Tolerance: 1e-6 relative to the largest reference entry (north-star FP64 tolerance)."""
import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from tests.helpers import DTYPE, cuda_handle, generate_covariance, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _oracle_predict_grads(params, mu, cov, f1_bar, Sff_bar, cross_bar, full_cov, model_uncertainty=True):
  model = oracle_svgp(params)
  mu = mu.clone().requires_grad_(True)
  cov = cov.clone().requires_grad_(True)
  match = gm.mm_svgp_mo(mo.GaussianMoments(mu, cov, True), model, model_uncertainty=model_uncertainty, full_output_cov=full_cov)
  Sff = match.y.covariance()
  if Sff.dim() == 2:
    Sff = torch.diag_embed(Sff)
  s = (match.y.mean() * f1_bar).sum() + (Sff * Sff_bar).sum() + (match.cross[0] * cross_bar).sum()
  gmu, gcov = torch.autograd.grad(s, (mu, cov))
  return gmu, 0.5 * (gcov + gcov.transpose(-1, -2))


@pytest.mark.parametrize("L,M,D,whiten,coreg,full_cov", [(1, 24, 3, True, False, True), (3, 40, 4, False, False, True),
                                                         (4, 64, 6, True, False, True), (2, 300, 5, True, False, True),
                                                         (3, 33, 4, True, True, True), (3, 40, 4, True, False, False)])
def test_mm_gp_predict_bwd_matches_autograd(L, M, D, whiten, coreg, full_cov):
  params = synthetic.random_svgp(L=L, M=M, D=D, seed=5, whiten=whiten, P=(L + 1 if coreg else None))
  gen = torch.Generator().manual_seed(11)
  N = 5
  mu = torch.randn(N, D, dtype=DTYPE, generator=gen) * 0.5
  cov = generate_covariance(D, (N,), 0.3, gen)
  P = params["W"].shape[0] if params.get("W") is not None else L
  f1_bar = torch.randn(N, P, dtype=DTYPE, generator=gen)
  Sff_bar = torch.randn(N, P, P, dtype=DTYPE, generator=gen)
  cross_bar = torch.randn(N, D, P, dtype=DTYPE, generator=gen)
  Sff_bar_eff = Sff_bar if full_cov else torch.diag_embed(torch.diagonal(Sff_bar, dim1=-2, dim2=-1))
  gmu, gcov = _oracle_predict_grads(params, mu, cov, f1_bar, Sff_bar_eff, cross_bar, full_cov)
  h = cuda_handle(params)
  m_bar, S_bar = h.predict_bwd(_dev(mu), _dev(cov), _dev(f1_bar), _dev(Sff_bar), _dev(cross_bar), full_output_cov=full_cov)
  scaled_close(m_bar, gmu, 1e-6, "m_bar")
  scaled_close(S_bar, gcov, 1e-6, "S_bar")


def test_mm_gp_predict_bwd_partial_adjoints_and_fd():
  """None adjoints are zeros; and the result agrees with central finite differences of the CUDA forward itself."""
  params = synthetic.random_svgp(L=2, M=48, D=3, seed=2, whiten=True)
  gen = torch.Generator().manual_seed(3)
  mu = torch.randn(1, 3, dtype=DTYPE, generator=gen) * 0.3
  cov = generate_covariance(3, (1,), 0.2, gen)
  Sff_bar = torch.randn(1, 2, 2, dtype=DTYPE, generator=gen)
  h = cuda_handle(params)
  m_bar, S_bar = h.predict_bwd(_dev(mu), _dev(cov), None, _dev(Sff_bar), None)
  eps = 1e-5
  fd = torch.zeros(3, dtype=DTYPE)
  for d in range(3):
    vals = []
    for sgn in (1, -1):
      mu2 = mu.clone()
      mu2[0, d] += sgn * eps
      _, Sff, _ = h.predict(_dev(mu2), _dev(cov))
      vals.append(float((Sff.cpu() * Sff_bar).sum()))
    fd[d] = (vals[0] - vals[1]) / (2 * eps)
  scaled_close(m_bar[0], fd, 1e-6, "m_bar vs finite differences")
  assert torch.allclose(S_bar, S_bar.transpose(-1, -2))
