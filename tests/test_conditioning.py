"""Conditioning study (SURVEY §7 "hard parts" 1b; CPU): which float64 association of the moment-matched GP variance is closer to the exact
value as cond(Kuu) grows — upstream's two triangular solves per input (moment_matching/models.py:224-226) or the O(M^2) contraction with
Kuu^-1-derived weights that the CUDA kernels use?  The arbiter is oracle/extended.py: the same quantity with every intermediate in
80-bit longdouble.  tests/test_gpu_conditioning.py repeats the comparison with the CUDA kernels themselves."""
import numpy as np
import pytest
import torch

from oracle import extended as ex
from oracle import gp_models as gm
from oracle import moments as mo
from oracle import psi_stats as ps

pytestmark = pytest.mark.skipif(np.finfo(np.longdouble).eps > 1e-18, reason="numpy longdouble is not extended precision on this platform")


def make_case(spread, jitter, seed=0, M=40, D=2, N=4):
  """M inducing points in a box of side `spread` lengthscales: the smaller the box, the closer Kuu is to singular (cond ~ var M / jitter)"""
  rng = np.random.default_rng(seed)
  Z = spread * rng.random((M, D))
  ell = np.array([1.0, 1.3])[:D]
  var = 0.9
  q_mu = rng.standard_normal((M, 1))
  A = 0.3 * rng.standard_normal((M, M)) / np.sqrt(M)
  q_sqrt = np.tril(A) + 0.2 * np.eye(M)
  mu = spread * (0.2 + 0.6 * rng.random((N, D)))
  B = rng.standard_normal((N, D, D))
  cov = 0.05 * B @ B.transpose(0, 2, 1) + 0.01 * np.eye(D)
  Zs = Z / ell
  d2 = ((Zs[:, None] - Zs[None]) ** 2).sum(-1)
  cond = float(np.linalg.cond(var * np.exp(-0.5 * d2) + jitter * np.eye(M)))
  return dict(Z=Z, ell=ell, var=var, q_mu=q_mu, q_sqrt=q_sqrt, mu=mu, cov=cov, jitter=jitter, cond=cond)


def float64_forms(c):
  """(upstream triangular-solve form, O(M^2) re-associated form) of (f1, Sff) in float64, from the torch oracle"""
  model = gm.SVGPModel([ps.SEKernel(c["var"], torch.as_tensor(c["ell"]))], [torch.as_tensor(c["Z"])], torch.as_tensor(c["q_mu"]),
                       torch.as_tensor(c["q_sqrt"])[None], whiten=True, mean_const=torch.zeros(1, dtype=torch.float64))
  x = mo.GaussianMoments(torch.as_tensor(c["mu"]), torch.as_tensor(c["cov"]), True)
  old = gm.Kuu.__defaults__
  gm.Kuu.__defaults__ = (float(c["jitter"]),)
  try:
    up = gm.mm_svgp_mo(x, model)
    re = gm.mm_sparse_reassociated(x, model)
  finally:
    gm.Kuu.__defaults__ = old
  return ((up.y.mean()[:, 0].numpy(), up.y.covariance()[:, 0, 0].numpy()), (re.y.mean()[:, 0].numpy(), re.y.covariance()[:, 0, 0].numpy()))


def extended(c):
  f1, Sff = ex.mm_svgp_so_extended(c["mu"], c["cov"], c["Z"], c["ell"], c["var"], c["q_mu"], c["q_sqrt"], jitter=c["jitter"])
  return f1, Sff


def test_extended_oracle_agrees_with_float64_when_well_conditioned():
  c = make_case(spread=20.0, jitter=1e-6)
  assert c["cond"] < 1e3
  (f1u, Su), (f1r, Sr) = float64_forms(c)
  f1x, Sx = extended(c)
  for a, b in ((f1u, f1x), (Su, Sx), (f1r, f1x), (Sr, Sx)):
    assert np.max(np.abs(a - b.astype(np.float64))) <= 1e-12 * np.max(np.abs(b.astype(np.float64)))


@pytest.mark.parametrize("spread,jitter", [(2.5, 1e-6), (1.2, 1e-6), (0.8, 1e-6), (1.2, 1e-8), (0.8, 1e-8)])
def test_reassociated_form_is_not_worse_than_upstream_form(spread, jitter):
  """The claim behind the 1e-6 parity bar: against the extended-precision value, the error of the re-associated float64 form stays
  within a small factor of the error of upstream's own float64 form (both are ~ cond(Kuu) * eps); the difference between the two forms
  — what a parity test CUDA-vs-upstream sees — is therefore conditioning noise of upstream's form as much as of ours."""
  c = make_case(spread, jitter)
  (f1u, Su), (f1r, Sr) = float64_forms(c)
  f1x, Sx = extended(c)
  scale = float(np.max(np.abs(Sx)))
  err_up = float(np.max(np.abs(Su - Sx.astype(np.float64)))) / scale
  err_re = float(np.max(np.abs(Sr - Sx.astype(np.float64)))) / scale
  gap = float(np.max(np.abs(Su - Sr))) / scale
  print(f"cond(Kuu) = {c['cond']:.1e}: |Sff - exact| / |Sff|  upstream form {err_up:.1e}, re-associated form {err_re:.1e}; forms differ by {gap:.1e}")
  assert err_re <= 10.0 * max(err_up, 1e-15 * c["cond"])
  assert err_up < 1e-5 and err_re < 1e-5
