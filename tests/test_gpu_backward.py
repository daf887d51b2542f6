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
                                                         (3, 33, 4, True, True, True), (3, 40, 4, True, False, False),
                                                         (2, 130, 7, True, False, True), (2, 50, 8, False, False, True),
                                                         (1, 17, 1, True, False, True), (2, 31, 2, True, False, True)])
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


# ---------------------------------------------------------------------------------------------------------
# moment-matched rollout: gradient of the loss w.r.t. the policy parameters and the initial moments
# ---------------------------------------------------------------------------------------------------------
def _oracle_rollout_grads(cfg, dynp, polp_list, m0, S0, H, loss_bar):
  from oracle import psi_stats as ps
  from oracle import rollout as ro
  dyn = oracle_svgp(dynp)
  enc = mo.TrigonometricEncoder(cfg["active_dims"])
  obj = mo.GaussianObjective(cfg["target"], cfg["W"])
  m0 = m0.clone().requires_grad_(True)
  S0 = S0.clone().requires_grad_(True)
  N = m0.shape[0]
  leaves, losses = [], []
  for n in range(N):
    pp = polp_list[0] if len(polp_list) == 1 else polp_list[n]
    if len(polp_list) == 1 and leaves:
      Z, ell, q = leaves[0]
    else:
      Z = torch.as_tensor(pp["Z"][0]).clone().requires_grad_(True)
      ell = torch.as_tensor(pp["lengthscales"][0]).clone().requires_grad_(True)
      q = torch.as_tensor(pp["q_mu"]).clone().requires_grad_(True)
      leaves.append((Z, ell, q))
    pol = gm.SVGPModel([ps.SEKernel(float(pp["variance"][0]), ell)], [Z], q, torch.as_tensor(pp["q_sqrt"]), whiten=bool(pp["whiten"]),
                       mean_const=torch.zeros(1, dtype=DTYPE))
    loss = ro.mm_rollout(m0[n:n + 1], S0[n:n + 1], H, lambda s: gm.mm_svgp(s, dyn),
                         lambda s: gm.mm_policy(s, pol, cfg["squash_scale"], cfg["squash_shift"]), enc, obj)
    losses.append(loss[0])
  total = (torch.stack(losses) * loss_bar).sum()
  flat = [t for trip in leaves for t in trip]
  grads = torch.autograd.grad(total, flat + [m0, S0])
  gS0 = 0.5 * (grads[-1] + grads[-1].transpose(-1, -2))
  trip = [grads[3 * i:3 * i + 3] for i in range(len(leaves))]
  return torch.stack(losses).detach(), trip, grads[-2], gS0


@pytest.mark.parametrize("shared_policy,whiten", [(True, True), (False, True), (True, False)])
def test_rollout_mm_gradients_match_autograd(shared_policy, whiten):
  from gpflowpilco_b200.autograd import rollout_mm_loss
  N, H = 3, 4
  g = torch.Generator().manual_seed(7)
  dynp = synthetic.random_svgp(L=4, M=40, D=6, seed=21, whiten=True, z_scale=1.5)
  dynp["q_mu"] = 0.2 * dynp["q_mu"]
  dynp["mean_const"] = np.zeros(4)
  R = 1 if shared_policy else N
  pols = []
  for r in range(R):
    pp = synthetic.random_svgp(L=1, M=10, D=5, seed=60 + r, whiten=whiten)
    pp["mean_const"] = np.zeros(1)
    pols.append(pp)
  m0 = torch.tensor([0.0, 2.5, 0.0, 0.0], dtype=DTYPE) + 0.3 * torch.randn(N, 4, dtype=DTYPE, generator=g)
  S0 = generate_covariance(4, [N], 0.15, g)
  cfg = dict(active_dims=(1,), target=np.array([0.0, 1.0, 0.0, 0.0, 0.0]), W=synthetic.config1_cartpole(M=8, Mp=4)["W"],
             squash_scale=3.0, squash_shift=-0.5)
  loss_bar = torch.randn(N, dtype=DTYPE, generator=g)
  loss_ref, trip, gm0, gS0 = _oracle_rollout_grads(cfg, dynp, pols, m0, S0, H, loss_bar)

  Z = _dev(np.stack([p["Z"][0] for p in pols])).requires_grad_(True)
  ell = _dev(np.stack([p["lengthscales"][0] for p in pols])).requires_grad_(True)
  q = _dev(np.stack([p["q_mu"][:, 0] for p in pols])).requires_grad_(True)
  var = _dev(np.array([p["variance"][0] for p in pols]))
  m0d, S0d = _dev(m0).requires_grad_(True), _dev(S0).requires_grad_(True)
  loss = rollout_mm_loss(cuda_handle(dynp), Z, ell, var, q, m0d, S0d, H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]),
                         squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"], whiten=whiten)
  scaled_close(loss, loss_ref, 1e-6, "loss")
  (loss * _dev(loss_bar)).sum().backward()
  scaled_close(m0d.grad, gm0, 1e-6, "m0 gradient")
  scaled_close(0.5 * (S0d.grad + S0d.grad.transpose(-1, -2)), gS0, 1e-6, "S0 gradient")
  scaled_close(Z.grad, torch.stack([t[0] for t in trip]), 1e-6, "policy centre gradient")
  scaled_close(ell.grad, torch.stack([t[1] for t in trip]), 1e-6, "policy lengthscale gradient")
  scaled_close(q.grad, torch.stack([t[2][:, 0] for t in trip]), 1e-6, "policy q_mu gradient")


@pytest.mark.parametrize("active_dims", [(0, 2), (), (3, 1, 0)])
def test_rollout_mm_gradients_other_encodings(active_dims):
  """Several / no sincos-encoded dimensions: exercises the pair terms exp(-(v_k + v_l)/2 -+ c_kl) of the encoder rule and of its
  closed-form adjoint (mm_small.cuh), which the cart-pole encoding (one angle) never reaches."""
  from gpflowpilco_b200.autograd import rollout_mm_loss
  N, H, Dx = 2, 3, 4
  na = len(active_dims)
  De = Dx + na
  g = torch.Generator().manual_seed(17)
  rng = np.random.default_rng(4)
  dynp = synthetic.random_svgp(L=Dx, M=24, D=De + 1, seed=31, whiten=True, z_scale=1.5)
  dynp["q_mu"] = 0.2 * dynp["q_mu"]
  dynp["mean_const"] = np.zeros(Dx)
  pp = synthetic.random_svgp(L=1, M=8, D=De, seed=71, whiten=True)
  pp["mean_const"] = np.zeros(1)
  m0 = torch.tensor([0.3, 2.0, -0.4, 0.6], dtype=DTYPE) + 0.3 * torch.randn(N, Dx, dtype=DTYPE, generator=g)
  S0 = generate_covariance(Dx, [N], 0.2, g)
  A = rng.standard_normal((De, De))
  cfg = dict(active_dims=tuple(active_dims), target=rng.standard_normal(De), W=A @ A.T / De, squash_scale=2.0, squash_shift=-0.5)
  loss_bar = torch.randn(N, dtype=DTYPE, generator=g)
  loss_ref, trip, gm0, gS0 = _oracle_rollout_grads(cfg, dynp, [pp], m0, S0, H, loss_bar)
  Z = _dev(pp["Z"][0][None]).requires_grad_(True)
  ell = _dev(pp["lengthscales"][0][None]).requires_grad_(True)
  q = _dev(pp["q_mu"][:, 0][None]).requires_grad_(True)
  var = _dev(np.array([pp["variance"][0]]))
  m0d, S0d = _dev(m0).requires_grad_(True), _dev(S0).requires_grad_(True)
  loss = rollout_mm_loss(cuda_handle(dynp), Z, ell, var, q, m0d, S0d, H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]),
                         squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"], whiten=True)
  scaled_close(loss, loss_ref, 1e-6, "loss")
  (loss * _dev(loss_bar)).sum().backward()
  scaled_close(m0d.grad, gm0, 1e-6, "m0 gradient")
  scaled_close(0.5 * (S0d.grad + S0d.grad.transpose(-1, -2)), gS0, 1e-6, "S0 gradient")
  scaled_close(Z.grad, trip[0][0][None], 1e-6, "policy centre gradient")
  scaled_close(ell.grad, trip[0][1][None], 1e-6, "policy lengthscale gradient")
  scaled_close(q.grad, trip[0][2][:, 0][None], 1e-6, "policy q_mu gradient")


# ---------------------------------------------------------------------------------------------------------
# pathwise particle rollout: gradient w.r.t. the policy parameters and the initial states
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S,F,M,Mp,H", [(40, 64, 20, 6, 3), (300, 128, 40, 10, 4)])
def test_rollout_pathwise_gradients_match_autograd(S, F, M, Mp, H):
  from gpflowpilco_b200.autograd import rollout_pathwise_loss
  from gpflowpilco_b200.pathwise import PackedPaths
  from oracle import pathwise as pw
  from oracle import psi_stats as ps
  from oracle import rollout as ro
  cfg = synthetic.config1_cartpole(M=M, Mp=Mp)
  cfg["policy"]["q_mu"] = 200.0 * cfg["policy"]["q_mu"]
  dyn = oracle_svgp(cfg["dynamics"])
  paths = pw.generate_paths(dyn, F, 3, 0, S)
  x0 = pw.draw_initial_states(torch.as_tensor(cfg["m0"][0]), torch.linalg.cholesky(torch.as_tensor(cfg["S0"][0])), 3, 0, S)
  g = torch.Generator().manual_seed(1)
  loss_bar = torch.randn(S, dtype=DTYPE, generator=g)
  pp = cfg["policy"]
  Z = torch.as_tensor(pp["Z"][0]).clone().requires_grad_(True)
  ell = torch.as_tensor(pp["lengthscales"][0]).clone().requires_grad_(True)
  q = torch.as_tensor(pp["q_mu"]).clone().requires_grad_(True)
  x0r = x0.clone().requires_grad_(True)
  pol = gm.SVGPModel([ps.SEKernel(float(pp["variance"][0]), ell)], [Z], q, torch.as_tensor(pp["q_sqrt"]), whiten=True,
                     mean_const=torch.zeros(1, dtype=DTYPE))
  enc = mo.TrigonometricEncoder(cfg["active_dims"])
  obj = mo.GaussianObjective(cfg["target"], cfg["W"])
  loss_ref = ro.pathwise_rollout(x0r, H, lambda eu: pw.evaluate_paths(dyn, paths, eu),
                                 lambda e: gm.policy_sample_path(pol, e, cfg["squash_scale"], cfg["squash_shift"]), enc, obj)
  gZ, gell, gq, gx0 = torch.autograd.grad((loss_ref * loss_bar).sum(), (Z, ell, q, x0r))

  d = cfg["dynamics"]
  packed = PackedPaths.from_sample_major(_dev(d["Z"]), _dev(d["lengthscales"]), _dev(d["variance"]), _dev(d["mean_const"]),
                                         _dev(paths.omega), _dev(paths.phase), _dev(paths.w), _dev(paths.v))
  Zd = _dev(pp["Z"]).requires_grad_(True)
  elld = _dev(pp["lengthscales"]).requires_grad_(True)
  qd = _dev(pp["q_mu"][:, 0][None]).requires_grad_(True)
  x0d = _dev(x0).requires_grad_(True)
  loss = rollout_pathwise_loss(packed, Zd, elld, _dev(pp["variance"]), qd, x0d, H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]),
                               squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  scaled_close(loss, loss_ref.detach(), 1e-8, "loss")
  (loss * _dev(loss_bar)).sum().backward()
  scaled_close(x0d.grad, gx0, 1e-6, "x0 gradient")
  scaled_close(Zd.grad[0], gZ, 1e-6, "policy centre gradient")
  scaled_close(elld.grad[0], gell, 1e-6, "policy lengthscale gradient")
  scaled_close(qd.grad[0], gq[:, 0], 1e-6, "policy q_mu gradient")


def test_sharded_closures_single_process_and_chunk_invariance():
  """distributed.py on one GPU (no process group): the pathwise mean cost / gradient do not depend on how the particles
  are chunked (the same property that makes them independent of the number of ranks: Philox streams are keyed by the
  global particle index), and the restart-sharded MM closure returns per-restart losses and gradients."""
  from gpflowpilco_b200 import distributed as gd
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  cfg["policy"]["q_mu"] = 100.0 * cfg["policy"]["q_mu"]
  pp = cfg["policy"]
  h = cuda_handle(cfg["dynamics"])
  args = (h, _dev(pp["Z"]), _dev(pp["lengthscales"]), _dev(pp["variance"]), _dev(pp["q_mu"][:, 0][None]), _dev(cfg["m0"][0]), _dev(cfg["S0"][0]))
  kw = dict(total_particles=700, num_bases=64, seed=4, horizon=3, active_dims=cfg["active_dims"], cost_target=_dev(cfg["target"]),
            cost_W=_dev(cfg["W"]), squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  l1, g1 = gd.pathwise_policy_loss_and_grad(*args, **kw, max_particles_per_launch=1024)
  l2, g2 = gd.pathwise_policy_loss_and_grad(*args, **kw, max_particles_per_launch=256)
  assert abs(float(l1) - float(l2)) <= 1e-12 * abs(float(l1))
  for a, b in zip(g1, g2):
    scaled_close(b, a, 1e-10, "chunked vs single-launch gradient")
  R = 3
  g = torch.Generator().manual_seed(0)
  Z = _dev(pp["Z"]).repeat(R, 1, 1) + 0.1 * _dev(torch.randn(R, *pp["Z"].shape[1:], dtype=DTYPE, generator=g))
  ell = _dev(pp["lengthscales"]).repeat(R, 1)
  q = _dev(pp["q_mu"][:, 0][None]).repeat(R, 1)
  var = _dev(pp["variance"]).repeat(R)
  losses, (start, count), grads = gd.mm_restart_losses_and_grads(h, Z, ell, var, q, _dev(cfg["m0"]), _dev(cfg["S0"]), 4, cfg["active_dims"],
                                                                 _dev(cfg["target"]), _dev(cfg["W"]), cfg["squash_scale"], cfg["squash_shift"])
  assert losses.shape == (R,) and (start, count) == (0, R) and grads[0].shape == Z.shape and torch.isfinite(grads[0]).all()
  assert float((losses[1:] - losses[0]).abs().max()) > 0     # different restarts, different losses


@pytest.mark.parametrize("A", [2, 3])
def test_squash_nd_backward_against_finite_differences(A):
  """Multi-dimensional NormalCDF rule (upstream bijectors.py:59-63 through utils/bvn.py): the closed-form reverse mode
  (gpp_mm_squash_nd_bwd: phi * Phi and the bivariate density, no bivariate probability) against Richardson-extrapolated central
  differences of the device forward, every entry of the covariance perturbed on its own as the forward reads it."""
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = torch.Generator().manual_seed(40 + A)
  N = 4
  mf = 0.8 * torch.randn(N, A, dtype=DTYPE, generator=g)
  Sf = generate_covariance(A, [N], 0.6, g)
  w_mu, w_S, w_g = (torch.randn(*s, dtype=DTYPE, generator=g) for s in ((N, A), (N, A, A), (N, A)))

  def value(m, S):
    match = moment_matching(GaussianMoments(moments=(m, S), centered=True), M.NormalCDF())
    gain = torch.diagonal(match.cross[0], dim1=-2, dim2=-1)
    return (match.y.mean() * _dev(w_mu)).sum() + (match.y.covariance() * _dev(w_S)).sum() + (gain * _dev(w_g)).sum()

  m_d, S_d = _dev(mf).requires_grad_(True), _dev(Sf).requires_grad_(True)
  gm_, gS_ = torch.autograd.grad(value(m_d, S_d), (m_d, S_d))

  def fd(which, index, h):
    out = []
    for sgn in (+1, -1):
      m, S = mf.clone(), Sf.clone()
      (m if which == 0 else S)[index] += sgn * h
      out.append(float(value(_dev(m), _dev(S))))
    return (out[0] - out[1]) / (2 * h)

  def richardson(which, index):
    return (4 * fd(which, index, 5e-4) - fd(which, index, 1e-3)) / 3

  fm = torch.tensor([[richardson(0, (n, i)) for i in range(A)] for n in range(N)], dtype=DTYPE)
  fS = torch.tensor([[[richardson(1, (n, i, j)) for j in range(A)] for i in range(A)] for n in range(N)], dtype=DTYPE)
  scaled_close(gm_, fm, 1e-7, "d/d mean")
  scaled_close(gS_, fS, 1e-7, "d/d covariance")
