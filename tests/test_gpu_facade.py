"""GPU tests of the reference-facing Python surface (same names / call structure as upstream gpflow_pilco):
moment_matching rules one at a time, kernel_expectation, forward_sde + MomentMatchingEuler (rule-by-rule path) against
the fused rollout and against the oracle, and the two PILCO policy-loss closures."""
import math

import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from oracle import psi_stats as ps
from oracle import rollout as ro
from tests.helpers import DTYPE, generate_covariance, log_uniform, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _facade_models(cfg):
  from gpflowpilco_b200 import models as M
  d, p = cfg["dynamics"], cfg["policy"]
  L = d["Z"].shape[0]
  drift = M.SVGP(M.SeparateIndependent([M.SquaredExponential(_dev(d["variance"][l]), _dev(d["lengthscales"][l])) for l in range(L)]),
                 M.SeparateIndependentInducingVariables([M.InducingPoints(_dev(d["Z"][l])) for l in range(L)]),
                 _dev(d["q_mu"]), _dev(d["q_sqrt"]), whiten=True, mean_function=M.Constant(_dev(d["mean_const"])))
  pol_svgp = M.SVGP(M.SeparateIndependent([M.SquaredExponential(_dev(p["variance"][0]), _dev(p["lengthscales"][0]))]),
                    M.SeparateIndependentInducingVariables([M.InducingPoints(_dev(p["Z"][0]))]), _dev(p["q_mu"]), _dev(p["q_sqrt"]),
                    whiten=True, mean_function=M.Constant(_dev(p["mean_const"])))
  link = M.BijectorChain([M.Scale(cfg["squash_scale"]), M.Shift(cfg["squash_shift"]), M.NormalCDF()])
  policy = M.InverseLinkWrapper(M.KernelRegressor(pol_svgp), invlink=link)
  return drift, policy


def test_rules_one_at_a_time():
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder, sincos
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  from gpflowpilco_b200 import models as M
  g = torch.Generator().manual_seed(0)
  m = torch.randn(3, 4, dtype=DTYPE, generator=g)
  S = generate_covariance(4, [3], 0.3, g)
  x = GaussianMoments((_dev(m), _dev(S)), True)
  ox = mo.GaussianMoments(m, S, True)
  # encoder (upstream tests/test_components.py:69-104 checks this rule against MC; here against the oracle)
  for active in [(1,), (2, 3), (0, 1, 2, 3)]:
    a = moment_matching(x, TrigonometricEncoder(active))
    b = mo.mm_encoder(ox, mo.TrigonometricEncoder(active))
    scaled_close(a.y.mean(), b.y.mean(), 1e-13, "encoder mean")
    scaled_close(a.y.covariance(), b.y.covariance(), 1e-12, "encoder cov")
    scaled_close(a.cross_covariance(), b.cross_covariance(), 1e-12, "encoder cross")
  # sincos / sin / cos
  a, b = moment_matching(x, sincos), mo.mm_sincos(ox)
  scaled_close(a.y.covariance(), b.y.covariance(), 1e-12, "sincos cov")
  scaled_close(a.cross[0], b.cross[0], 1e-13, "sincos cross")
  scaled_close(moment_matching(x, torch.cos).y.covariance(), mo.mm_cos(ox).y.covariance(), 1e-12, "cos cov")
  scaled_close(moment_matching(x, torch.sin).cross[0], mo.mm_sin(ox).cross[0], 1e-13, "sin cross")
  # squash chain on a 1-D Gaussian
  x1 = GaussianMoments((_dev(m[:, :1]), _dev(S[:, :1, :1])), True)
  link = M.BijectorChain([M.Scale(20 - 1e-5), M.Shift(-0.5), M.NormalCDF()])
  a = moment_matching(x1, link)
  b = mo.mm_squash(mo.GaussianMoments(m[:, :1], S[:, :1, :1], True), 20 - 1e-5)
  scaled_close(a.y.mean(), b.y.mean(), 1e-13, "squash mean")
  scaled_close(a.y.covariance(), b.y.covariance(), 1e-11, "squash var")
  scaled_close(a.cross[0], b.cross[0], 1e-12, "squash cross")
  # objective: expectation and samples (upstream tests/test_components.py:38-66)
  W = torch.linalg.inv(generate_covariance(4, scale=0.5, gen=g))
  t = torch.randn(4, dtype=DTYPE, generator=g)
  obj, oobj = GaussianObjective(_dev(t), _dev(W)), mo.GaussianObjective(t, W)
  scaled_close(obj(x), oobj(ox), 1e-12, "expected cost")
  X = torch.randn(50, 4, dtype=DTYPE, generator=g)
  scaled_close(obj(_dev(X)), oobj(X), 1e-13, "sample cost")


def test_owens_t_against_scipy():
  import ctypes
  from scipy.special import owens_t
  from gpflowpilco_b200 import _lib
  h = np.linspace(-9, 9, 400)
  a = np.random.default_rng(0).uniform(0.01, 1.0, 400)
  out = torch.empty(400, dtype=DTYPE, device="cuda")
  hd, ad = _dev(h), _dev(a)
  _lib.check(_lib.load().gpp_owens_t(400, ctypes.c_void_p(hd.data_ptr()), ctypes.c_void_p(ad.data_ptr()), ctypes.c_void_p(out.data_ptr()), None))
  assert float(np.abs(out.cpu().numpy() - owens_t(h, a)).max()) < 1e-15


def test_kernel_expectation_signatures():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.utils.kernel_expectation import Gaussian, kernel_expectation
  g = torch.Generator().manual_seed(1)
  D, Mz, N = 3, 10, 2
  mu = torch.randn(N, D, dtype=DTYPE, generator=g)
  cov = generate_covariance(D, [N], 0.2, g)
  ks = [ps.SEKernel(0.8, log_uniform([D], 0.5, 2.0, g)), ps.SEKernel(1.2, log_uniform([D], 0.5, 2.0, g))]
  Zs = [torch.randn(Mz, D, dtype=DTYPE, generator=g) for _ in range(2)]
  fk = [M.SquaredExponential(_dev(k.variance), _dev(k.lengthscales)) for k in ks]
  fz = [M.InducingPoints(_dev(z)) for z in Zs]
  p = Gaussian(_dev(mu), _dev(cov))
  scaled_close(kernel_expectation(p, (fk[0], fz[0])), ps.eKxz(mu, cov, ks[0], Zs[0]), 1e-12, "eKxz")
  scaled_close(kernel_expectation(p, (fk[0], fz[0]), (fk[1], fz[1])), ps.eKzxKxz(mu, cov, ks[0], Zs[0], ks[1], Zs[1]), 1e-11, "eKzxKxz")
  mk, mz = M.SeparateIndependent(fk), M.SeparateIndependentInducingVariables(fz)
  scaled_close(kernel_expectation(p, mk), ps.eKff_list(mu, ks), 1e-15, "eKff list")
  scaled_close(kernel_expectation(p, (mk, mz)), ps.eKfu_list(mu, cov, ks, Zs), 1e-12, "eKfu list")
  scaled_close(kernel_expectation(p, (mk, mz), (mk, mz)), ps.eKuffu_list(mu, cov, ks, Zs), 1e-11, "eKuffu list")


def test_kernel_expectation_separate_dims_shortcut():
  """upstream _E (kernel_expectation.py:85-94): kernels on disjoint active dims under a DiagonalGaussian factorise into two Psi1
  blocks; under a full Gaussian the case is refused.  Checked against the oracle's Psi1 outer product and, independently, against
  the JOINT Psi2 kernel with the inactive dimensions switched off by lengthscales of 1e8."""
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.utils.kernel_expectation import DiagonalGaussian, Gaussian, kernel_expectation
  g = torch.Generator().manual_seed(7)
  D, N, M1, M2 = 4, 3, 9, 7
  A, B = (0, 1), (2, 3)
  mu = torch.randn(N, D, dtype=DTYPE, generator=g)
  var = log_uniform([N, D], 0.05, 0.5, g)
  ka, kb = ps.SEKernel(0.7, log_uniform([2], 0.5, 2.0, g)), ps.SEKernel(1.3, log_uniform([2], 0.5, 2.0, g))
  Z1, Z2 = torch.randn(M1, D, dtype=DTYPE, generator=g), torch.randn(M2, D, dtype=DTYPE, generator=g)
  fa = M.SquaredExponential(_dev(ka.variance), _dev(ka.lengthscales), active_dims=A)
  fb = M.SquaredExponential(_dev(kb.variance), _dev(kb.lengthscales), active_dims=B)
  p = DiagonalGaussian(_dev(mu), _dev(var))
  got = kernel_expectation(p, (fa, M.InducingPoints(_dev(Z1))), (fb, M.InducingPoints(_dev(Z2))))
  assert got.shape == (N, M1, M2)
  cov = torch.diag_embed(var)
  ia, ib = list(A), list(B)
  e1 = ps.eKxz(mu[:, ia], cov[:, ia][:, :, ia], ka, Z1[:, ia])
  e2 = ps.eKxz(mu[:, ib], cov[:, ib][:, :, ib], kb, Z2[:, ib])
  scaled_close(got, e1[:, :, None] * e2[:, None, :], 1e-12, "separate dims vs oracle Psi1 x Psi1")
  big = torch.full([2], 1e8, dtype=DTYPE)
  ka_full = ps.SEKernel(ka.variance, torch.cat([ka.lengthscales, big]))
  kb_full = ps.SEKernel(kb.variance, torch.cat([big, kb.lengthscales]))
  joint = kernel_expectation(Gaussian(_dev(mu), _dev(cov)),
                             (M.SquaredExponential(_dev(ka_full.variance), _dev(ka_full.lengthscales)), M.InducingPoints(_dev(Z1))),
                             (M.SquaredExponential(_dev(kb_full.variance), _dev(kb_full.lengthscales)), M.InducingPoints(_dev(Z2))))
  scaled_close(got, joint.cpu(), 1e-10, "separate dims vs the joint Psi2 kernel")
  with pytest.raises(NotImplementedError):
    kernel_expectation(Gaussian(_dev(mu), _dev(cov)), (fa, M.InducingPoints(_dev(Z1))), (fb, M.InducingPoints(_dev(Z2))))


def test_mm_closure_fused_equals_rule_by_rule_and_oracle():
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO
  cfg = synthetic.config1_cartpole(M=48, Mp=10)
  drift, policy = _facade_models(cfg)
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.6, step_size=0.1)
  loop = MomentMatchingPILCO(spec, GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  assert spec.num_steps == 6
  fused = loop.policy_loss_closure()()
  stepwise = loop.policy_loss_closure(fused=False)()
  dyn, pol = oracle_svgp(cfg["dynamics"]), oracle_svgp(cfg["policy"])
  ref = ro.mm_rollout(torch.as_tensor(cfg["m0"]), torch.as_tensor(cfg["S0"]), 6, lambda s: gm.mm_svgp(s, dyn),
                      lambda s: gm.mm_policy(s, pol, cfg["squash_scale"], cfg["squash_shift"]),
                      mo.TrigonometricEncoder(cfg["active_dims"]), mo.GaussianObjective(cfg["target"], cfg["W"]))
  scaled_close(fused, ref, 1e-7, "fused closure vs oracle")
  scaled_close(stepwise, ref, 1e-7, "rule-by-rule closure vs oracle")


def test_policy_sample_path_and_pathwise_closure():
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, PathwisePILCO
  cfg = synthetic.config1_cartpole(M=32, Mp=10)
  drift, policy = _facade_models(cfg)
  g = torch.Generator().manual_seed(2)
  e = torch.randn(9, 5, dtype=DTYPE, generator=g)
  pol = oracle_svgp(cfg["policy"])
  scaled_close(policy(_dev(e)), gm.policy_sample_path(pol, e, cfg["squash_scale"], cfg["squash_shift"]), 1e-10, "policy(e)")
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.3, step_size=0.1)
  loop = PathwisePILCO(spec, GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  closure = loop.policy_loss_closure(batch_size=64, num_bases=64, seed=5)
  l1, l2 = closure(), closure()
  assert l1.shape == (64,) and torch.isfinite(l1).all()
  assert not torch.equal(l1, l2)          # fresh paths per call (upstream loops/pilco.py:281-284)
  again = loop.policy_loss_closure(batch_size=64, num_bases=64, seed=5)()
  assert torch.equal(l1, again)           # same seed -> identical draws


def test_mm_closure_is_differentiable_like_upstream_tape_gradient():
  """loss.backward() on the fused closure (upstream: tape.gradient(loss, policy.trainable_variables), optimizers.py:52-56)
  agrees with central finite differences of the closure itself."""
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO
  cfg = synthetic.config1_cartpole(M=32, Mp=6)
  cfg["policy"]["q_mu"] = 50.0 * cfg["policy"]["q_mu"]          # 1e-3 N(0,1) initial weights give a nearly flat loss
  drift, policy = _facade_models(cfg)
  svgp = policy.model.model
  q_mu = svgp.q_mu.clone().requires_grad_(True)
  svgp.q_mu = q_mu
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.5, step_size=0.1)
  loop = MomentMatchingPILCO(spec, GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  closure = loop.policy_loss_closure()
  loss = closure()
  assert loss.requires_grad
  loss.sum().backward()
  grad = q_mu.grad.clone()
  fd = torch.zeros_like(grad)
  eps = 1e-5
  with torch.no_grad():
    for i in range(q_mu.shape[0]):
      vals = []
      for sgn in (1.0, -1.0):
        q2 = q_mu.detach().clone()
        q2[i, 0] += sgn * eps
        svgp.q_mu = q2
        vals.append(float(closure().sum()))
      fd[i, 0] = (vals[0] - vals[1]) / (2 * eps)
  scaled_close(grad, fd, 1e-6, "d loss / d q_mu vs finite differences")


def test_gradient_descent_minimises_the_mm_policy_loss():
  """The upstream optimisation loop (utils/optimizers.py:46-78: closure under a tape, clip-norm transform, Adam) run on the
  fused differentiable closure: the expected cost goes down and the trained variables are the policy's own tensors."""
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO
  from gpflowpilco_b200.utils.optimizers import GradientDescent, clip_by_global_norm
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  drift, policy = _facade_models(cfg)
  svgp = policy.model.model
  svgp.q_mu = svgp.q_mu.clone().requires_grad_(True)
  Zvar = svgp.latent_inducing()[0].clone().requires_grad_(True)
  svgp.inducing_variable.inducing_variables[0].Z = Zvar
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.8, step_size=0.1)
  loop = MomentMatchingPILCO(spec, GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  closure = loop.policy_loss_closure()
  opt = GradientDescent(step_limit=25, optimizer_factory=lambda vs: torch.optim.Adam(vs, lr=5e-2), transform=clip_by_global_norm(1.0))
  hist = opt.minimize(closure, [svgp.q_mu, Zvar])
  assert len(hist) == 25 and all(math.isfinite(h) for h in hist)
  assert hist[-1] < hist[0] - 1e-6, f"loss did not decrease: {hist[0]} -> {hist[-1]}"


def test_gradient_descent_on_the_pathwise_closure():
  """PathwisePILCO's closure is differentiable (upstream takes tape.gradient of it, utils/optimizers.py:52-56): with trainable policy
  tensors it goes through the gradient-mode forward + reverse sweep, and GradientDescent.minimize lowers the mean particle cost."""
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, PathwisePILCO
  from gpflowpilco_b200.utils.optimizers import GradientDescent, clip_by_global_norm
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  cfg["policy"]["q_mu"] = 100.0 * cfg["policy"]["q_mu"]
  drift, policy = _facade_models(cfg)
  svgp = policy.model.model
  svgp.q_mu = svgp.q_mu.clone().requires_grad_(True)
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.5, step_size=0.1)
  loop = PathwisePILCO(spec, GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  from gpflowpilco_b200 import pathwise as pw
  from gpflowpilco_b200.moment_matching.models import svgp_handle
  paths = pw.generate_paths(svgp_handle(drift, True), 256, 64, seed=3)        # fixed paths: a deterministic objective
  closure = loop.policy_loss_closure(batch_size=256, num_bases=64, seed=3, paths=paths)
  loss = closure()
  assert loss.shape == (256,) and loss.requires_grad
  mean_closure = lambda: closure().mean()[None]
  opt = GradientDescent(step_limit=15, optimizer_factory=lambda vs: torch.optim.Adam(vs, lr=5e-2), transform=clip_by_global_norm(1.0))
  hist = opt.minimize(mean_closure, [svgp.q_mu])
  assert len(hist) == 15 and all(math.isfinite(h) for h in hist)
  assert hist[-1] < hist[0] - 1e-7, f"loss did not decrease: {hist[0]} -> {hist[-1]}"


def test_fused_and_rule_by_rule_closures_agree_when_the_policy_has_a_mean_or_jitter():
  """A policy SVGP with a Constant mean is not what the fused kernels compute: the closure must fall back (and give the same loss as
  fused=False); a non-default kuu_jitter is honoured by the fused path."""
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
  from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  cfg["policy"]["q_mu"] = 100.0 * cfg["policy"]["q_mu"]
  spec = EpisodeSpec(GaussianStateDistribution(_dev(cfg["m0"][0]), _dev(cfg["S0"][0])), horizon=0.4, step_size=0.1)
  obj = GaussianObjective(_dev(cfg["target"]), _dev(cfg["W"]))
  enc = TrigonometricEncoder(cfg["active_dims"])
  # (a) constant mean
  drift, policy = _facade_models(cfg)
  policy.model.model.mean_function = M.Constant(_dev(np.array([0.3])))
  loop = MomentMatchingPILCO(spec, obj, drift, policy, enc)
  assert loop._policy_params() is None and not loop._fusable()
  a, b = loop.policy_loss_closure()(), loop.policy_loss_closure(fused=False)()
  assert torch.equal(a, b)
  drift0, policy0 = _facade_models(cfg)
  base = MomentMatchingPILCO(spec, obj, drift0, policy0, enc).policy_loss_closure()()
  assert float((a - base).abs().max()) > 1e-9          # the mean is not silently dropped
  # (b) non-default jitter
  drift, policy = _facade_models(cfg)
  policy.model.model.kuu_jitter = 1e-3
  loop = MomentMatchingPILCO(spec, obj, drift, policy, enc)
  assert loop._fusable() and loop._policy_params().jitter == 1e-3
  scaled_close(loop.policy_loss_closure()(), loop.policy_loss_closure(fused=False)(), 1e-9, "fused vs rule-by-rule with kuu_jitter = 1e-3")
  assert float((loop.policy_loss_closure()() - base).abs().max()) > 1e-12


def test_active_dims_cross_term_lives_in_the_full_state_and_handle_cache_sees_settings():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = torch.Generator().manual_seed(4)
  D, Mi, N = 5, 12, 3
  act = (0, 2, 3)
  Z = torch.randn(Mi, D, dtype=DTYPE, generator=g)
  kern = M.SquaredExponential(_dev(np.array(0.8)), _dev(np.array([0.9, 1.1, 1.3])), active_dims=act)
  model = M.SVGP(kern, M.InducingPoints(_dev(Z)), _dev(torch.randn(Mi, 1, dtype=DTYPE, generator=g)), None, whiten=True)
  mu = torch.randn(N, D, dtype=DTYPE, generator=g)
  A = torch.randn(N, D, D, dtype=DTYPE, generator=g)
  cov = 0.1 * A @ A.transpose(-1, -2) + 0.05 * torch.eye(D, dtype=DTYPE)
  x = GaussianMoments((_dev(mu), _dev(cov)), True)
  mm = moment_matching(x, model)
  assert mm.cross[0].shape == (N, D, 1)
  inactive = [d for d in range(D) if d not in act]
  assert float(mm.cross[0][:, inactive].abs().max()) == 0.0
  # Cov(x, f) = Sxx (Sxx^-1 Cov(x, f)) through the full state: compare with Monte Carlo-free identity on the active block
  sub = GaussianMoments((_dev(mu[:, list(act)]), _dev(cov[:, list(act)][:, :, list(act)])), True)
  kern_sub = M.SquaredExponential(_dev(np.array(0.8)), _dev(np.array([0.9, 1.1, 1.3])))
  model_sub = M.SVGP(kern_sub, M.InducingPoints(_dev(Z[:, list(act)])), model.q_mu, None, whiten=True)
  ref = moment_matching(sub, model_sub)
  scaled_close(mm.y.mean(), ref.y.mean().cpu(), 1e-12, "mean through active dims")
  scaled_close(mm.cross_covariance()[:, list(act)], ref.cross_covariance().cpu(), 1e-10, "Cov(x_active, f)")
  joint = mm.joint()
  assert joint.mean().shape == (N, D + 1)
  # handle cache: a changed setting (jitter) must rebuild the handle
  f_before = mm.y.covariance().clone()
  model.kuu_jitter = 1e-2
  f_after = moment_matching(x, model).y.covariance()
  assert float((f_after - f_before).abs().max()) > 1e-9
