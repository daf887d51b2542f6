"""GPU parity: fused moment-matched rollout vs the oracle's step-by-step restatement of
forward_sde + MomentMatchingEuler + GaussianObjective (upstream dynamics/forward_sde.py:95-137, dynamics/solvers.py:110-135,
loops/pilco.py:199-217).  Tolerance 1e-6 relative (north star); observed errors are reported in the assertion text."""
import sys

import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from oracle import rollout as ro
from tests.helpers import DTYPE, cuda_handle, generate_covariance, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _policy_params(pol, scale, shift=-0.5):
  from gpflowpilco_b200.rollouts import PolicyParams
  return PolicyParams(_dev(pol["Z"]), _dev(pol["lengthscales"]), _dev(pol["variance"]), _dev(pol["q_mu"][:, 0][None]),
                      whiten=bool(pol["whiten"]), squash_scale=scale, squash_shift=shift)


def _oracle_rollout(cfg, dyn, pol, m0, S0, H, reference_form=True):
  enc = mo.TrigonometricEncoder(cfg["active_dims"])
  obj = mo.GaussianObjective(cfg["target"], cfg["W"])
  drift = (lambda s: gm.mm_svgp(s, dyn)) if reference_form else (lambda s: gm.mm_sparse_reassociated(s, dyn))
  return ro.mm_rollout(m0, S0, H, drift, lambda s: gm.mm_policy(s, pol, cfg["squash_scale"], cfg["squash_shift"]), enc, obj, True)


def test_policy_beta():
  cfg = synthetic.config1_cartpole(M=64)
  pol = oracle_svgp(cfg["policy"])
  beta_ref, _ = gm.sparse_weights(pol, model_uncertainty=False)
  beta = _policy_params(cfg["policy"], 1.0).beta()
  scaled_close(beta[0], beta_ref[0], 1e-9, "policy beta")


def test_cartpole_rollout_config1():
  """BASELINE config #1: cart-pole MM rollout, M=256 dynamics, 30 policy centres, H=30, N=1."""
  from gpflowpilco_b200.rollouts import rollout_mm
  cfg = synthetic.config1_cartpole()
  dyn, pol = oracle_svgp(cfg["dynamics"]), oracle_svgp(cfg["policy"])
  H = cfg["horizon"]
  m0, S0 = torch.as_tensor(cfg["m0"]), torch.as_tensor(cfg["S0"])
  loss_ref, traj = _oracle_rollout(cfg, dyn, pol, m0, S0, H)
  res = rollout_mm(cuda_handle(cfg["dynamics"]), _policy_params(cfg["policy"], cfg["squash_scale"], cfg["squash_shift"]),
                   _dev(m0), _dev(S0), H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]), return_trajectory=True)
  tm_ref = torch.stack([t[0] for t in traj])
  tS_ref = torch.stack([t[1] for t in traj])
  scaled_close(res.traj_m, tm_ref, 1e-6, "trajectory means")
  scaled_close(res.traj_S, tS_ref, 1e-6, "trajectory covariances")
  scaled_close(res.loss, loss_ref, 1e-6, "loss")
  scaled_close(res.m_final, tm_ref[-1], 1e-6, "final mean")


@pytest.mark.parametrize("shared_policy", [True, False])
def test_batched_rollout_random_models(shared_policy):
  """N = 5 initial states, well-conditioned random dynamics; one shared policy or one policy per rollout (restarts)."""
  from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm
  N, H = 5, 6
  g = torch.Generator().manual_seed(5)
  dynp = synthetic.random_svgp(L=4, M=48, D=6, seed=21, whiten=True, z_scale=1.5)
  dynp["q_mu"] = 0.2 * dynp["q_mu"]
  dynp["mean_const"] = np.zeros(4)
  dyn = oracle_svgp(dynp)
  R = 1 if shared_policy else N
  pols = []
  for r in range(R):
    pp = synthetic.random_svgp(L=1, M=12, D=5, seed=40 + r, whiten=(r % 2 == 0))
    pp["mean_const"] = np.zeros(1)
    pols.append(pp)
  m0 = torch.tensor([0.0, 2.5, 0.0, 0.0], dtype=DTYPE) + 0.3 * torch.randn(N, 4, dtype=DTYPE, generator=g)
  S0 = generate_covariance(4, [N], 0.15, g)
  cfg = dict(active_dims=(1,), target=np.array([0.0, 1.0, 0.0, 0.0, 0.0]), W=synthetic.config1_cartpole(M=8, Mp=4)["W"],
             squash_scale=3.0, squash_shift=-0.5)
  losses, finals = [], []
  for n in range(N):
    pol = oracle_svgp(pols[0 if shared_policy else n])
    l, traj = _oracle_rollout(cfg, dyn, pol, m0[n:n + 1], S0[n:n + 1], H)
    losses.append(l)
    finals.append(traj[-1])
  # whiten differs per restart in the non-shared case: the C ABI takes one flag, so prepare beta per set on the host side
  if shared_policy:
    P = PolicyParams(_dev(pols[0]["Z"]), _dev(pols[0]["lengthscales"]), _dev(pols[0]["variance"]), _dev(pols[0]["q_mu"][:, 0][None]),
                     whiten=bool(pols[0]["whiten"]), squash_scale=3.0)
    beta = None
  else:
    P = PolicyParams(_dev(np.concatenate([p["Z"] for p in pols])), _dev(np.concatenate([p["lengthscales"] for p in pols])),
                     _dev(np.concatenate([p["variance"] for p in pols])), _dev(np.stack([p["q_mu"][:, 0] for p in pols])),
                     whiten=True, squash_scale=3.0)
    betas = []
    for p in pols:
      one = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=bool(p["whiten"]))
      betas.append(one.beta())
    beta = torch.cat(betas)
  res = rollout_mm(cuda_handle(dynp), P, _dev(m0), _dev(S0), H, (1,), _dev(cfg["target"]), _dev(cfg["W"]), beta=beta)
  scaled_close(res.loss, torch.cat(losses), 1e-7, "loss")
  scaled_close(res.m_final, torch.cat([f[0] for f in finals]), 1e-7, "final means")
  scaled_close(res.S_final, torch.cat([f[1] for f in finals]), 1e-6, "final covariances")


def test_rollout_is_capturable_in_a_cuda_graph():
  """The C entry points only enqueue kernels / memsets on the caller's stream (no allocation, no synchronisation), so a whole
  H-step rollout can be captured once and replayed: the replay reproduces the eager result bit for bit."""
  from gpflowpilco_b200.rollouts import rollout_mm
  cfg = synthetic.config1_cartpole(M=48, Mp=10)
  h = cuda_handle(cfg["dynamics"])
  pol = _policy_params(cfg["policy"], cfg["squash_scale"], cfg["squash_shift"])
  beta = pol.beta()
  m0, S0, tgt, W = _dev(cfg["m0"]), _dev(cfg["S0"]), _dev(cfg["target"]), _dev(cfg["W"])
  run = lambda: rollout_mm(h, pol, m0, S0, 8, cfg["active_dims"], tgt, W, beta=beta, check=False).loss
  eager = run().clone()
  side = torch.cuda.Stream()
  with torch.cuda.stream(side):
    run()
  torch.cuda.synchronize()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    out = run()
  out.zero_()
  graph.replay()
  torch.cuda.synchronize()
  assert torch.equal(out, eager)


def test_graphed_policy_gradient_replays_the_eager_result():
  """Forward rollout + reverse sweep + policy-weight adjoint captured once (gpflowpilco_b200/graphs.py) and replayed with new
  parameter values: bit-identical to the eager autograd path, for the captured values and for perturbed ones."""
  from gpflowpilco_b200.autograd import rollout_mm_loss
  from gpflowpilco_b200.graphs import GraphedMMPolicyGradient
  cfg = synthetic.config1_cartpole(M=48, Mp=10)
  h = cuda_handle(cfg["dynamics"])
  p = cfg["policy"]
  Z, ell, var, q = _dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None])
  m0, S0, tgt, W = _dev(cfg["m0"]), _dev(cfg["S0"]), _dev(cfg["target"]), _dev(cfg["W"])
  kw = dict(squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  g = GraphedMMPolicyGradient(h, Z, ell, var, q, m0, S0, 6, cfg["active_dims"], tgt, W, **kw)
  gen = torch.Generator().manual_seed(3)
  for trial in range(2):
    Zt = (Z + 0.05 * trial * torch.randn(Z.shape, dtype=torch.float64, generator=gen).to(Z.device)).requires_grad_(True)
    et = (ell * (1.0 + 0.1 * trial)).requires_grad_(True)
    qt = (q + 0.01 * trial).requires_grad_(True)
    loss = rollout_mm_loss(h, Zt, et, var, qt, m0, S0, 6, cfg["active_dims"], tgt, W, **kw)
    ref = torch.autograd.grad(loss.sum(), (Zt, et, qt))
    gl, gg = g(Zt.detach(), et.detach(), qt.detach())
    assert torch.equal(gl, loss.detach())
    for a, b in zip(gg, ref):
      assert torch.equal(a, b)


@pytest.mark.parametrize("R,H,M", [(2, 1, None), (3, 2, None), (15, 3, None), (16, 3, None), (17, 3, None), (33, 4, None), (17, 4, 100),
                                   (5, 7, 40), (129, 2, None)])
def test_persistent_sweeps_match_the_per_stage_path_around_their_thresholds(R, H, M):
  """The persistent forward / backward kernels against the one-launch-per-stage path (same device functions, different
  orchestration) at rollout counts around the switches of the forward hand-over (continuous item stream from 2 rollouts, deferred
  partial reduction from 16), odd counts, one step, several tiles per CTA (M = 100, 40): every output and every gradient."""
  import importlib.util
  import os
  spec = importlib.util.spec_from_file_location("check_persist", os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts", "check_persist.py"))
  cp = importlib.util.module_from_spec(spec)
  argv = sys.argv
  sys.argv = ["check_persist.py", "none"]
  try:
    spec.loader.exec_module(cp)
    from gpflowpilco_b200 import rollouts
    try:
      assert cp.compare(R, H, M) < 1e-7
    finally:
      rollouts.set_rollout_mode(rollouts.ROLLOUT_AUTO)
  finally:
    sys.argv = argv


def test_deferred_not_positive_definite_check():
  """check='defer' queues the asynchronous flags of policy weights / forward sweep / reverse sweep; raise_deferred() reads them with
  one synchronisation: silent for a healthy rollout, GppError for a covariance that is not positive definite."""
  from gpflowpilco_b200 import _lib, rollouts
  from gpflowpilco_b200.autograd import rollout_mm_loss
  from tests.helpers import cuda_handle
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  p = cfg["policy"]
  handle = cuda_handle(cfg["dynamics"])
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
  Z, ell, q = (T(p["Z"]).clone().requires_grad_(True), T(p["lengthscales"]).clone().requires_grad_(True),
               T(p["q_mu"][:, 0][None]).clone().requires_grad_(True))
  args = (3, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]))
  kw = dict(squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"], check="defer")
  loss = rollout_mm_loss(handle, Z, ell, T(p["variance"]), q, T(cfg["m0"]), T(cfg["S0"]), *args, **kw)
  loss.sum().backward()
  assert len(rollouts._DEFERRED) == 3
  rollouts.raise_deferred()
  assert not rollouts._DEFERRED
  ref = rollout_mm_loss(handle, Z, ell, T(p["variance"]), q, T(cfg["m0"]), T(cfg["S0"]), *args, squash_scale=cfg["squash_scale"],
                        squash_shift=cfg["squash_shift"])
  assert torch.equal(ref, loss)
  bad = -10.0 * torch.eye(T(cfg["S0"]).shape[-1], dtype=torch.float64, device="cuda")[None]
  rollout_mm_loss(handle, Z, ell, T(p["variance"]), q, T(cfg["m0"]), bad, *args, **kw)        # no error here: nothing has been read yet
  with pytest.raises(_lib.GppError):
    rollouts.raise_deferred()
  assert not rollouts._DEFERRED
