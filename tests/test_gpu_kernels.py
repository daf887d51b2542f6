"""GPU parity tests: CUDA path (through the C ABI) vs the oracle on identical seeded inputs.

Tolerance: 1e-6 relative to the largest magnitude of each output tensor (north-star bar for the FP64 path);
most checks are orders of magnitude tighter and assert that too where the conditioning allows.
"""
import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from oracle import psi_stats as ps
from tests.helpers import DTYPE, cuda_handle, generate_covariance, log_uniform, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _inputs(N, D, seed, scale=0.3):
  g = torch.Generator().manual_seed(seed)
  mu = torch.randn(N, D, dtype=DTYPE, generator=g)
  cov = generate_covariance(D, [N], scale, g)
  return mu, cov, g


@pytest.mark.parametrize("D,M,N", [(1, 7, 3), (2, 32, 1), (5, 30, 4), (6, 256, 2), (8, 130, 3)])
def test_ekxz(D, M, N):
  from gpflowpilco_b200 import ops
  mu, cov, g = _inputs(N, D, 100 + D)
  Z = torch.randn(M, D, dtype=DTYPE, generator=g)
  k = ps.SEKernel(0.89 ** 2, log_uniform([D], 0.3, 3.0, g))
  ref = ps.eKxz(mu, cov, k, Z)
  out = ops.ekxz(_dev(mu), _dev(cov), _dev(Z), _dev(k.lengthscales), float(k.variance))
  scaled_close(out, ref, 1e-12, "ekxz")


@pytest.mark.parametrize("D,M1,M2,N", [(1, 5, 4, 2), (2, 32, 32, 1), (4, 16, 33, 3), (6, 70, 64, 2), (8, 129, 100, 2), (3, 140, 257, 2),
                                       (5, 1, 1, 1), (7, 300, 131, 3)])
def test_ekzxkxz_generic(D, M1, M2, N):
  from gpflowpilco_b200 import ops
  mu, cov, g = _inputs(N, D, 200 + D)
  Z1 = torch.randn(M1, D, dtype=DTYPE, generator=g)
  Z2 = torch.randn(M2, D, dtype=DTYPE, generator=g)
  k1 = ps.SEKernel(0.89 ** 2, log_uniform([D], 0.3, 3.0, g))
  k2 = ps.SEKernel(1.3, log_uniform([D], 0.3, 3.0, g))
  ref = ps.eKzxKxz(mu, cov, k1, Z1, k2, Z2)
  out = ops.ekzxkxz(_dev(mu), _dev(cov), _dev(Z1), _dev(k1.lengthscales), float(k1.variance),
                    _dev(Z2), _dev(k2.lengthscales), float(k2.variance))
  scaled_close(out, ref, 1e-11, "ekzxkxz generic")


@pytest.mark.parametrize("D,M,N", [(2, 32, 1), (6, 96, 3), (5, 31, 2)])
def test_ekzxkxz_same(D, M, N):
  from gpflowpilco_b200 import ops
  mu, cov, g = _inputs(N, D, 300 + D)
  Z = torch.randn(M, D, dtype=DTYPE, generator=g)
  k = ps.SEKernel(0.7, log_uniform([D], 0.3, 3.0, g))
  ref = ps.eKzxKxz(mu, cov, k, Z)
  out = ops.ekzxkxz(_dev(mu), _dev(cov), _dev(Z), _dev(k.lengthscales), float(k.variance))
  scaled_close(out, ref, 1e-11, "ekzxkxz same")
  scaled_close(out, out.transpose(-1, -2), 1e-13, "symmetry")


def test_ekzxkxz_far_points_underflow_to_zero():
  """Inducing points hundreds of lengthscales away: entries underflow; result must be exactly 0, not garbage."""
  from gpflowpilco_b200 import ops
  mu, cov, g = _inputs(1, 2, 7)
  Z = torch.tensor([[0.0, 0.0], [500.0, -500.0], [1e4, 1e4]], dtype=DTYPE)
  k = ps.SEKernel(1.0, torch.tensor([0.5, 0.5], dtype=DTYPE))
  out = ops.ekzxkxz(_dev(mu), _dev(cov), _dev(Z), _dev(k.lengthscales), 1.0).cpu()
  ref = ps.eKzxKxz(mu, cov, k, Z)
  assert torch.isfinite(out).all()
  scaled_close(out, ref, 1e-11, "far points")
  assert float(out[0, 2, 2]) == 0.0


def test_not_positive_definite_is_reported():
  from gpflowpilco_b200 import _lib, ops
  mu = torch.zeros(2, 2, dtype=DTYPE)
  cov = torch.stack([torch.eye(2, dtype=DTYPE), torch.tensor([[1.0, 5.0], [5.0, 1.0]], dtype=DTYPE) * 10])
  Z = torch.zeros(3, 2, dtype=DTYPE)
  with pytest.raises(_lib.GppError, match="batch element 1"):
    ops.ekxz(_dev(mu), _dev(cov), _dev(Z), _dev(torch.ones(2, dtype=DTYPE)), 1.0)


@pytest.mark.parametrize("whiten", [True, False])
@pytest.mark.parametrize("unc", [True, False])
def test_model_weights(whiten, unc):
  params = synthetic.random_svgp(L=3, M=40, D=4, seed=5, whiten=whiten)
  beta_ref, C_ref = gm.sparse_weights(oracle_svgp(params), unc)
  beta, C = cuda_handle(params, unc).weights()
  scaled_close(beta, beta_ref, 1e-9, "beta")
  scaled_close(C, C_ref, 1e-9, "C")


@pytest.mark.parametrize("L,M,D,N,whiten,unc,P", [
    (1, 30, 5, 1, True, False, None),      # the cart-pole policy shape (KernelRegressor)
    (4, 64, 6, 3, True, True, None),
    (4, 100, 6, 5, False, True, None),     # M not a multiple of the tile
    (2, 16, 4, 2, False, True, 3),         # upstream tests/test_moment_matching.py:198-264 shape (coregionalised)
    (3, 130, 3, 2, True, False, None),
    (2, 70, 8, 2, True, True, None),
    (1, 9, 1, 4, False, True, None),
    (2, 150, 7, 9, True, True, None),      # D = 7: three k-steps of the exponent inner product
    (2, 40, 2, 3, True, True, None),       # D = 2: a single k-step
])
def test_mm_gp_predict_vs_reference_form(L, M, D, N, whiten, unc, P):
  """CUDA O(M^2) path vs the oracle's triangular-solve form (upstream moment_matching/models.py:200-299)."""
  params = synthetic.random_svgp(L=L, M=M, D=D, seed=11 * L + M, whiten=whiten, P=P)
  mu, cov, _ = _inputs(N, D, 400 + M)
  model = oracle_svgp(params)
  x = mo.GaussianMoments(mu, cov, True)
  ref = gm.mm_svgp_mo(x, model, model_uncertainty=unc, jitter=1e-8)
  h = cuda_handle(params, unc)
  f1, Sff, cross = h.predict(_dev(mu), _dev(cov), jitter=1e-8)
  scaled_close(f1, ref.y.mean(), 1e-8, "f1")
  scaled_close(Sff, ref.y.covariance(), 1e-6, "Sff")
  scaled_close(cross, ref.cross[0], 1e-8, "cross")
  # diag-only variant must equal the diagonal of the full one (upstream tests :127-136 use 1e-12)
  f1d, Sffd, crossd = h.predict(_dev(mu), _dev(cov), full_output_cov=False, jitter=1e-8)
  torch.testing.assert_close(f1d, f1, rtol=1e-12, atol=0)
  torch.testing.assert_close(torch.diagonal(Sffd, dim1=-2, dim2=-1), torch.diagonal(Sff, dim1=-2, dim2=-1), rtol=1e-12, atol=1e-14)
  torch.testing.assert_close(crossd, cross, rtol=1e-12, atol=0)
  off = Sffd - torch.diag_embed(torch.diagonal(Sffd, dim1=-2, dim2=-1))
  assert float(off.abs().max()) == 0.0


def test_mm_gp_predict_elementwise_worst_case():
  """`scaled_close` measures against the LARGEST entry of a tensor.  Once, element by element, at the cart-pole shape (L=4, M=256,
  D=6): every entry must agree with the oracle to 1e-6 of its OWN magnitude unless it is itself below 1e-9 of the tensor's scale
  (entries of Sff that are differences of O(1) terms carry the absolute rounding of those terms)."""
  params = synthetic.random_svgp(L=4, M=256, D=6, seed=5, whiten=True)
  mu, cov, _ = _inputs(16, 6, 905)
  ref = gm.mm_svgp_mo(mo.GaussianMoments(mu, cov, True), oracle_svgp(params), model_uncertainty=True, jitter=1e-8)
  f1, Sff, cross = cuda_handle(params, True).predict(_dev(mu), _dev(cov), jitter=1e-8)
  for name, got, want in (("f1", f1, ref.y.mean()), ("Sff", Sff, ref.y.covariance()), ("cross", cross, ref.cross[0])):
    got = got.cpu()
    scale = float(want.abs().max())
    big = want.abs() > 1e-9 * scale
    rel = ((got - want).abs() / want.abs().clamp_min(1e-300))[big]
    small_abs = (got - want).abs()[~big]
    worst_small = float(small_abs.max()) / scale if small_abs.numel() else 0.0
    print(f"[elementwise] {name}: worst element-wise rel err {float(rel.max()):.2e} over {int(big.sum())} entries "
          f"(smallest |entry| / scale {float(want.abs()[big].min()) / scale:.1e}); {int((~big).sum())} tiny entries, abs err / scale {worst_small:.1e}")
    assert float(rel.max()) <= 1e-6, name
    assert worst_small <= 1e-12, name


def test_mm_gp_predict_gpr():
  """Exact GPR (upstream moment_matching/models.py:44-111, tests/test_moment_matching.py:87-136 sizes)."""
  from gpflowpilco_b200 import ops
  g = torch.Generator().manual_seed(77)
  D, M, N = 4, 16, 2
  k = ps.SEKernel(0.89 ** 2, log_uniform([D], 0.1, 10.0, g))
  X = torch.rand(M, D, dtype=DTYPE, generator=g)
  Y = 0.89 * torch.randn(M, 1, dtype=DTYPE, generator=g)
  c = 1 + torch.randn(1, dtype=DTYPE, generator=g)
  model = gm.GPRModel(k, X, Y, torch.tensor(1e-5, dtype=DTYPE), c)
  mu = torch.rand(N, D, dtype=DTYPE, generator=g)
  cov = generate_covariance(D, [N], 0.01, g)
  ref = gm.mm_gpr(mo.GaussianMoments(mu, cov, True), model)
  h = ops.GPModelHandle(_dev(X)[None], _dev(k.lengthscales)[None], _dev(k.variance.reshape(1)), _dev(Y - c), None,
                        whiten=False, mean_const=_dev(c), kuu_jitter=1e-5, model_uncertainty=True)
  f1, Sff, cross = h.predict(_dev(mu), _dev(cov))
  scaled_close(f1, ref.y.mean(), 1e-7, "f1")
  scaled_close(cross, ref.cross[0], 1e-7, "cross")
  # Sff = f2 - f1^2 + e_cov cancels ~1e4-fold at scale_x = 0.01 with noise 1e-5 (cond(Kyy) ~ 1e6+): compare on the f2 scale
  err = float((Sff.cpu() - ref.y.covariance()).abs().max())
  assert err <= 1e-6 * float(k.variance), err


def test_mm_gp_predict_large_tiles_and_chunks():
  """Enough inputs to take the 128x128-tile, multi-chunk schedule; M=300 leaves ragged edge tiles."""
  L, M, D, N = 2, 300, 6, 700
  params = synthetic.random_svgp(L=L, M=M, D=D, seed=3, whiten=True, z_scale=2.0)
  mu, cov, _ = _inputs(N, D, 9)
  model = oracle_svgp(params)
  ref = gm.mm_sparse_reassociated(mo.GaussianMoments(mu[:40], cov[:40], True), model, jitter=0.0)
  h = cuda_handle(params, True)
  f1, Sff, cross = h.predict(_dev(mu), _dev(cov))
  scaled_close(f1[:40], ref.y.mean(), 1e-10, "f1")
  scaled_close(Sff[:40], ref.y.covariance(), 1e-8, "Sff")
  scaled_close(cross[:40], ref.cross[0], 1e-10, "cross")
  # batch-independence: same inputs evaluated alone (small-tile schedule) give the same numbers
  f1b, Sffb, crossb = h.predict(_dev(mu[690:]), _dev(cov[690:]))
  scaled_close(f1[690:], f1b.cpu(), 1e-12, "f1 batch independence")
  scaled_close(Sff[690:], Sffb.cpu(), 1e-10, "Sff batch independence")


def test_config2_shape_subsample():
  """BASELINE config #2 model (M=1000 exact GPs, E=4, D=6) on 12 of its inputs vs the oracle."""
  from gpflowpilco_b200 import ops
  cfg = synthetic.config2_batched_mm_predict(N=12)
  E = cfg["Y"].shape[1]
  ks = [ps.SEKernel(float(cfg["variance"][e]), torch.as_tensor(cfg["lengthscales"][e])) for e in range(E)]
  X = torch.as_tensor(cfg["X"])
  model = gm.SVGPModel(ks, [X] * E, torch.as_tensor(cfg["Y"] - cfg["mean_const"]), torch.zeros(E, 1000, 1000, dtype=DTYPE),
                       whiten=False, mean_const=torch.as_tensor(cfg["mean_const"]))
  # exact GP == SVGP(q_mu = Y - c, q_sqrt = 0, Kuu jitter = noise): reuse the re-associated oracle with that jitter
  import oracle.gp_models as gmod
  old = gmod.DEFAULT_JITTER
  gmod.DEFAULT_JITTER = 1e-2
  try:
    gmod.Kuu.__defaults__ = (1e-2,)
    ref = gm.mm_sparse_reassociated(mo.GaussianMoments(torch.as_tensor(cfg["mu"]), torch.as_tensor(cfg["cov"]), True), model)
  finally:
    gmod.Kuu.__defaults__ = (old,)
    gmod.DEFAULT_JITTER = old
  h = ops.GPModelHandle(_dev(np.broadcast_to(cfg["X"], (E,) + cfg["X"].shape).copy()), _dev(cfg["lengthscales"]),
                        _dev(cfg["variance"]), _dev(cfg["Y"] - cfg["mean_const"]), None, whiten=False,
                        mean_const=_dev(cfg["mean_const"]), kuu_jitter=list(cfg["noise_variance"]))
  f1, Sff, cross = h.predict(_dev(cfg["mu"]), _dev(cfg["cov"]))
  scaled_close(f1, ref.y.mean(), 1e-7, "f1")
  scaled_close(Sff, ref.y.covariance(), 1e-6, "Sff")
  scaled_close(cross, ref.cross[0], 1e-7, "cross")


def test_empty_and_degenerate_batches():
  """N = 0 / M = 0 are no-ops, a single inducing point and a single input work, odd row lengths take the scalar-store path."""
  from gpflowpilco_b200 import ops
  mu, cov, g = _inputs(2, 3, 5)
  Z = torch.randn(1, 3, dtype=DTYPE, generator=g)
  k = ps.SEKernel(1.1, log_uniform([3], 0.3, 3.0, g))
  out = ops.ekzxkxz(_dev(mu), _dev(cov), _dev(Z), _dev(k.lengthscales), float(k.variance))
  scaled_close(out, ps.eKzxKxz(mu, cov, k, Z), 1e-12, "M = 1")
  empty = ops.ekzxkxz(_dev(mu[:0]), _dev(cov[:0]), _dev(Z), _dev(k.lengthscales), float(k.variance))
  assert empty.shape == (0, 1, 1)
  e1 = ops.ekxz(_dev(mu[:0]), _dev(cov[:0]), _dev(Z), _dev(k.lengthscales), float(k.variance))
  assert e1.shape == (0, 1)
