"""The reference's own Monte-Carlo tests, run against the CUDA path through the reference-facing API.

upstream tests/test_moment_matching.py:87-264 draws 10^6 samples x ~ N(mx, Sxx), pushes them through model.predict_f and
compares the empirical moments with moment_matching(x, model) at 10/sqrt(S) absolute, and checks full- against
diagonal-covariance outputs at 1e-12.  Here the Monte-Carlo side uses the oracle's restated predict_f (GPflow itself is not
installable) and the closed-form side is `gpflowpilco_b200.moment_matching` on the GPU with the same call signature.
upstream tests/test_kernel_expectation.py:50-93 does the same for the Psi statistics."""
import pytest
import torch

from oracle import gp_models as gm
from oracle import psi_stats as ps
from tests.helpers import DTYPE, draw_samples_mvn, generate_covariance, log_uniform, mc_close
from tests.test_oracle_mc import D_MM, M_MM, N_MM, S_MC, _gen, _mc_estimator

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _check(match_full, match_diag, ref, n):
  mf, Sff, Sxf = ref
  assert mc_close(match_full.y.mean().cpu(), mf, n)
  assert mc_close(match_full.y.covariance().cpu(), Sff, n)
  assert mc_close(match_full.cross_covariance().cpu(), Sxf, n)
  torch.testing.assert_close(match_diag.y.mean(), match_full.y.mean(), rtol=1e-12, atol=0)
  dd = match_diag.y.covariance()
  dd = torch.diagonal(dd, dim1=-2, dim2=-1) if dd.dim() == 3 else dd
  torch.testing.assert_close(dd, torch.diagonal(match_full.y.covariance(), dim1=-2, dim2=-1), rtol=1e-12, atol=0)
  torch.testing.assert_close(match_diag.cross_covariance(), match_full.cross_covariance(), rtol=1e-12, atol=0)


def test_moment_matching_gpr():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = _gen(11)
  k = ps.SEKernel(0.89 ** 2, log_uniform([D_MM], 0.01, 10.0, g))
  c = 1 + torch.randn(1, dtype=DTYPE, generator=g)
  X = torch.rand(M_MM, D_MM, dtype=DTYPE, generator=g)
  Y = 0.89 * torch.randn(M_MM, 1, dtype=DTYPE, generator=g)
  omodel = gm.GPRModel(k, X, Y, torch.tensor(1e-5, dtype=DTYPE), c)
  mx = torch.rand(N_MM, D_MM, dtype=DTYPE, generator=g)
  Sxx = generate_covariance(D_MM, [N_MM], 0.01, g)

  def predict(Xs):
    mu, var = gm.gpr_predict_f(omodel, Xs)
    return mu, var.unsqueeze(-1)

  ref = _mc_estimator(predict, mx, Sxx, S_MC)
  model = M.GPR((_dev(X), _dev(Y)), M.SquaredExponential(_dev(k.variance), _dev(k.lengthscales)), mean_function=M.Constant(_dev(c)),
                noise_variance=1e-5)
  x = GaussianMoments((_dev(mx), _dev(Sxx)), True)
  _check(moment_matching(x, model), moment_matching(x, model, full_output_cov=False), ref, S_MC)


def test_moment_matching_svgp():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = _gen(12)
  k = ps.SEKernel(0.89 ** 2, log_uniform([D_MM], 0.01, 10.0, g))
  Z = torch.rand(M_MM, D_MM, dtype=DTYPE, generator=g)
  q_mu = 0.89 * torch.randn(M_MM, 1, dtype=DTYPE, generator=g)
  q_sqrt = torch.linalg.cholesky(generate_covariance(M_MM, scale=0.89, gen=g)[None])
  c = 1 + torch.randn(1, dtype=DTYPE, generator=g)
  omodel = gm.SVGPModel([k], [Z], q_mu, q_sqrt, whiten=False, mean_const=c, multi_output=False)
  mx = torch.rand(N_MM, D_MM, dtype=DTYPE, generator=g)
  Sxx = generate_covariance(D_MM, [N_MM], 0.01, g)
  ref = _mc_estimator(lambda Xs: gm.svgp_predict_f(omodel, Xs, full_output_cov=True), mx, Sxx, S_MC)
  model = M.SVGP(M.SquaredExponential(_dev(k.variance), _dev(k.lengthscales)), M.InducingPoints(_dev(Z)), _dev(q_mu), _dev(q_sqrt),
                 whiten=False, mean_function=M.Constant(_dev(c)))
  x = GaussianMoments((_dev(mx), _dev(Sxx)), True)
  _check(moment_matching(x, model), moment_matching(x, model, full_output_cov=False), ref, S_MC)


def test_moment_matching_svgp_mo():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = _gen(13)
  Lf, P = 2, 3
  ks, Zs = [], []
  for _ in range(Lf):
    Zs.append(torch.rand(M_MM, D_MM, dtype=DTYPE, generator=g))
    ks.append(ps.SEKernel(0.89 ** 2, log_uniform([D_MM], 0.01, 10.0, g)))
  W = torch.nn.functional.normalize(torch.rand(P, Lf, dtype=DTYPE, generator=g), dim=-1)
  c = 1 + torch.randn(P, dtype=DTYPE, generator=g)
  q_mu = 0.89 * torch.randn(M_MM, Lf, dtype=DTYPE, generator=g)
  q_sqrt = torch.linalg.cholesky(generate_covariance(M_MM, [Lf], 0.89, g))
  omodel = gm.SVGPModel(ks, Zs, q_mu, q_sqrt, whiten=False, mean_const=c, W=W)
  mx = torch.rand(N_MM, D_MM, dtype=DTYPE, generator=g)
  Sxx = generate_covariance(D_MM, [N_MM], 0.01, g)
  ref = _mc_estimator(lambda Xs: gm.svgp_predict_f(omodel, Xs, full_output_cov=True), mx, Sxx, S_MC)
  model = M.SVGP(M.LinearCoregionalization([M.SquaredExponential(_dev(k.variance), _dev(k.lengthscales)) for k in ks], _dev(W)),
                 M.SeparateIndependentInducingVariables([M.InducingPoints(_dev(Z)) for Z in Zs]), _dev(q_mu), _dev(q_sqrt),
                 whiten=False, mean_function=M.Constant(_dev(c)))
  x = GaussianMoments((_dev(mx), _dev(Sxx)), True)
  _check(moment_matching(x, model), moment_matching(x, model, full_output_cov=False), ref, S_MC)


@pytest.mark.parametrize("seed", [1, 2])
def test_expectation_squaredExp(seed):
  """upstream tests/test_kernel_expectation.py:50-93: eKxz and eKzxKxz (two kernels, two inducing sets) vs Monte Carlo."""
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.utils.kernel_expectation import Gaussian, kernel_expectation
  g = _gen(seed)
  D, M1, M2, N, S = 3, 8, 7, 2, 200000
  mu = torch.randn(N, D, dtype=DTYPE, generator=g)
  cov = generate_covariance(D, [N], 0.1, g)
  k1 = ps.SEKernel(0.89 ** 2, log_uniform([D], 0.1, 10.0, g))
  k2 = ps.SEKernel(0.89 ** 2, log_uniform([D], 0.1, 10.0, g))
  Z1 = torch.cat([mu[0] + 0.3 * torch.randn(M1 // 2, D, dtype=DTYPE, generator=g), torch.rand(M1 - M1 // 2, D, dtype=DTYPE, generator=g)])
  Z2 = torch.cat([mu[1] + 0.3 * torch.randn(M2 // 2, D, dtype=DTYPE, generator=g), torch.rand(M2 - M2 // 2, D, dtype=DTYPE, generator=g)])
  X = draw_samples_mvn(mu, cov, S, g)
  K1 = k1.K(X.reshape(-1, D), Z1).reshape(S, N, M1)
  K2 = k2.K(X.reshape(-1, D), Z2).reshape(S, N, M2)
  p = Gaussian(_dev(mu), _dev(cov))
  fk1, fk2 = M.SquaredExponential(_dev(k1.variance), _dev(k1.lengthscales)), M.SquaredExponential(_dev(k2.variance), _dev(k2.lengthscales))
  z1, z2 = M.InducingPoints(_dev(Z1)), M.InducingPoints(_dev(Z2))
  assert mc_close(kernel_expectation(p, (fk1, z1)).cpu(), K1.mean(0), S)
  assert mc_close(kernel_expectation(p, (fk1, z1), (fk2, z2)).cpu(), torch.einsum("sni,snj->nij", K1, K2) / S, S)
