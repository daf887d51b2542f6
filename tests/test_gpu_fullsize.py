"""GPU checks at BASELINE.json's full sizes, through size-independent properties (the oracle cannot run these sizes):

  config #2  N = 8192 inputs, 1000 training points: batch invariance (any sub-batch gives bit-identical rows: the reduction order
             per input is fixed), the oracle on a handful of rows, symmetric PSD covariances.
  config #3  Psi2 at N = 1024, M = 2048, D = 8 (34 GB): the materialised tensor contracted with weights equals the fused
             contraction kernel's result (checksum of checksums), same-kernel Psi2 is symmetric in (i, j), and a few
             slices match the oracle.
  config #4  pathwise particles: sharding invariance — particles computed in one launch or in shards by global index have
             bit-identical random draws (Philox streams keyed by the global index) and per-particle losses equal to 1e-9
             (the update-weight solve is a cuBLAS TRSM whose rounding depends on the batch shape).
"""
import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import psi_stats as ps
from tests.helpers import DTYPE, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _config2_handle(cfg):
  from gpflowpilco_b200 import ops
  E = cfg["Y"].shape[1]
  return ops.GPModelHandle(_dev(np.broadcast_to(cfg["X"], (E,) + cfg["X"].shape).copy()), _dev(cfg["lengthscales"]), _dev(cfg["variance"]),
                           _dev(cfg["Y"] - cfg["mean_const"]), None, whiten=False, mean_const=_dev(cfg["mean_const"]),
                           kuu_jitter=list(cfg["noise_variance"]))


def test_config2_full_size_properties():
  import bench
  cfg = synthetic.config2_batched_mm_predict()          # N = 8192, M = 1000, D = 6, E = 4
  h = _config2_handle(cfg)
  mu, cov = _dev(cfg["mu"]), _dev(cfg["cov"])
  f1, Sff, cross = h.predict(mu, cov)
  assert torch.isfinite(f1).all() and torch.isfinite(Sff).all() and torch.isfinite(cross).all()
  # batch invariance, bit for bit: rows 4096..4160 recomputed as a batch of 64, and row 8191 alone
  f1b, Sffb, crossb = h.predict(mu[4096:4160].contiguous(), cov[4096:4160].contiguous())
  assert torch.equal(f1b, f1[4096:4160]) and torch.equal(crossb, cross[4096:4160])
  scaled_close(Sffb, Sff[4096:4160], 1e-12, "Sff of a sub-batch (different tile schedule)")
  f1c, _, _ = h.predict(mu[8191:].contiguous(), cov[8191:].contiguous())
  assert torch.equal(f1c, f1[8191:])
  # symmetric, positive semi-definite output covariances
  assert float((Sff - Sff.transpose(-1, -2)).abs().max()) == 0.0
  assert float(torch.linalg.eigvalsh(Sff).min()) > -1e-9
  # the oracle (re-associated form) on 3 rows
  idx = [0, 4100, 8191]
  ref = bench.oracle_predict(bench.oracle_model(cfg), cfg, idx, reference_form=False)
  scaled_close(f1[idx], ref.y.mean(), 1e-6, "mean")
  scaled_close(Sff[idx], ref.y.covariance(), 1e-6, "covariance")
  scaled_close(cross[idx], ref.cross[0], 1e-6, "cross")


def test_config3_full_size_checksum_of_checksums():
  from gpflowpilco_b200 import ops
  cfg = synthetic.config3_psi2_stress()                 # N = 1024, M = 2048, D = 8
  N, M, D = cfg["mu"].shape[0], cfg["Z1"].shape[0], cfg["mu"].shape[1]
  mu, cov = _dev(cfg["mu"]), _dev(cfg["cov"])
  Z1, Z2 = _dev(cfg["Z1"]), _dev(cfg["Z2"])
  l1, l2 = _dev(cfg["lengthscales1"]), _dev(cfg["lengthscales2"])
  g = torch.Generator().manual_seed(0)
  b1 = torch.randn(M, dtype=DTYPE, generator=g)
  b2 = torch.randn(M, dtype=DTYPE, generator=g)
  # (ii) two kernels, two inducing sets: 34 GB materialised, contracted with b1 b2^T by a library GEMV pair
  Q = ops.ekzxkxz(mu, cov, Z1, l1, cfg["variance1"], Z2, l2, cfg["variance2"])
  assert Q.shape == (N, M, M)
  check = torch.einsum("i,nij,j->n", _dev(b1), Q, _dev(b2))
  rows = [0, 511, 1023]
  k1 = ps.SEKernel(cfg["variance1"], torch.as_tensor(cfg["lengthscales1"]))
  k2 = ps.SEKernel(cfg["variance2"], torch.as_tensor(cfg["lengthscales2"]))
  ref = ps.eKzxKxz(torch.as_tensor(cfg["mu"][rows]), torch.as_tensor(cfg["cov"][rows]), k1, torch.as_tensor(cfg["Z1"]), k2, torch.as_tensor(cfg["Z2"]))
  scaled_close(Q[rows], ref, 1e-9, "Psi2 slices vs the oracle")
  scaled_close(check[rows], torch.einsum("i,nij,j->n", b1, ref, b2), 1e-9, "contracted slices")
  del Q
  # the fused contraction kernel on the same problem: a 2-latent model with Z = (Z1, Z2), beta = (b1, b2); the off-diagonal
  # second moment f2[0,1] = b1^T Q b2 shows up in Sff[0,1] + f1[0] f1[1]
  h = ops.GPModelHandle(torch.stack([Z1, Z2]), torch.stack([l1, l2]), _dev([cfg["variance1"], cfg["variance2"]]),
                        torch.stack([_dev(b1), _dev(b2)], 1), None, whiten=False, kuu_jitter=1e-6, model_uncertainty=False)
  beta, _ = h.weights()                                  # Kuu^-1 q_mu: contract the materialised tensor with the SAME weights
  f1, Sff, _ = h.predict(mu, cov)
  Q = ops.ekzxkxz(mu, cov, Z1, l1, cfg["variance1"], Z2, l2, cfg["variance2"])
  f2 = torch.einsum("i,nij,j->n", beta[0], Q, beta[1])
  scaled_close(Sff[:, 0, 1] + f1[:, 0] * f1[:, 1], f2, 1e-9, "fused contraction vs materialised Psi2 (all 1024 inputs)")
  del Q
  # (i) same kernel, same inducing set: symmetric in (i, j)
  Qs = ops.ekzxkxz(mu[:128].contiguous(), cov[:128].contiguous(), Z1, l1, cfg["variance1"])
  assert float((Qs - Qs.transpose(-1, -2)).abs().max()) <= 1e-15 * float(Qs.abs().max())


def test_config4_particles_are_sharding_invariant():
  from gpflowpilco_b200 import ops
  from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths, rollout_pathwise
  from gpflowpilco_b200.rollouts import PolicyParams
  cfg = synthetic.config1_cartpole()
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(_dev(d["Z"]), _dev(d["lengthscales"]), _dev(d["variance"]), _dev(d["q_mu"]), _dev(d["q_sqrt"]), whiten=True,
                             mean_const=_dev(d["mean_const"]))
  policy = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=True,
                        squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  S, F, H = 4096, 4096, 20                               # config #4 shapes (F = 4096 bases, M = 256) on a slice of the 2^20 particles
  first = 2 ** 20 - S                                    # the LAST particles of the 1M-particle job
  m0, S0 = _dev(cfg["m0"][0]), _dev(cfg["S0"][0])

  def run(start, count):
    paths = generate_paths(handle, count, F, seed=9, first_particle=start)
    x0 = draw_initial_states(m0, S0, 9, start, count)
    loss = rollout_pathwise(paths, policy, x0, H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]))[0]
    return loss, paths.w[:, :, :count].clone(), x0

  whole, w_whole, x0_whole = run(first, S)
  shards = [run(first, 1000), run(first + 1000, 2000), run(first + 3000, S - 3000)]
  assert torch.equal(w_whole, torch.cat([s[1] for s in shards], -1))      # prior weights: pure Philox, bit for bit
  assert torch.equal(x0_whole, torch.cat([s[2] for s in shards], 0))      # initial states likewise
  scaled_close(torch.cat([s[0] for s in shards]), whole, 1e-9, "per-particle losses, sharded vs one launch")
  assert torch.isfinite(whole).all() and float(whole.std()) > 0


def _config5_problem(R, seed=5):
  from gpflowpilco_b200 import ops
  cfg = synthetic.config1_cartpole()                     # M = 256 dynamics, 30 policy centres
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(_dev(d["Z"]), _dev(d["lengthscales"]), _dev(d["variance"]), _dev(d["q_mu"]), _dev(d["q_sqrt"]), whiten=True,
                             mean_const=_dev(d["mean_const"]))
  g = torch.Generator().manual_seed(seed)
  Z = _dev(p["Z"]).repeat(R, 1, 1) + 0.3 * torch.randn(R, *p["Z"].shape[1:], dtype=DTYPE, generator=g).cuda()
  ell = _dev(p["lengthscales"]).repeat(R, 1) * torch.exp(torch.empty(R, 1, dtype=DTYPE).uniform_(-0.5, 0.5, generator=g)).cuda()
  q = 0.3 * torch.randn(R, p["Z"].shape[1], dtype=DTYPE, generator=g).cuda()      # large enough for the policy to act
  var = _dev(p["variance"]).repeat(R)
  m0, S0 = _dev(cfg["m0"]).expand(R, -1).contiguous(), _dev(cfg["S0"]).expand(R, -1, -1).contiguous()
  return cfg, handle, Z, ell, var, q, m0, S0


def test_config5_shape_backward_persistent_vs_per_stage_and_directional_differences():
  """BASELINE config #5's per-GPU shape (M = 256, H = 100, 30 policy centres; 4 of the 64 restarts): gpp_rollout_mm_fwd_save +
  gpp_rollout_mm_bwd.  (i) the persistent on-device sweep and the one-launch-per-stage sweep agree; (ii) the gradient agrees with
  central differences of the forward loss along random directions in (Z, lengthscales, q_mu) space — the forward itself is pinned to
  upstream's vectors at H = 30 (tests/test_gpu_golden.py), the backward to upstream's finite differences at H = 5."""
  from gpflowpilco_b200 import rollouts
  from gpflowpilco_b200.autograd import rollout_mm_loss
  R, H = 4, 100
  cfg, handle, Z, ell, var, q, m0, S0 = _config5_problem(R)
  kw = dict(squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  args = (H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]))

  def loss_and_grads(mode):
    prev = rollouts.set_rollout_mode(mode)
    try:
      leaves = [t.clone().requires_grad_(True) for t in (Z, ell, q)]
      loss = rollout_mm_loss(handle, leaves[0], leaves[1], var, leaves[2], m0, S0, *args, **kw)
      loss.sum().backward()
      return loss.detach(), [t.grad for t in leaves]
    finally:
      rollouts.set_rollout_mode(prev)

  loss_p, g_p = loss_and_grads(rollouts.ROLLOUT_PERSIST)
  loss_l, g_l = loss_and_grads(rollouts.ROLLOUT_LEGACY)
  assert torch.isfinite(loss_p).all() and all(torch.isfinite(g).all() for g in g_p)
  scaled_close(loss_p, loss_l, 1e-9, "loss, persistent vs per-stage")
  for name, a, b in zip(("Z", "lengthscales", "q_mu"), g_p, g_l):
    scaled_close(a, b, 1e-7, f"{name} gradient, persistent vs per-stage")
  gen = torch.Generator().manual_seed(3)
  for trial in range(2):
    v = [torch.randn(t.shape, dtype=DTYPE, generator=gen).cuda() * s for t, s in ((Z, 1.0), (ell, 0.2), (q, 0.3))]
    def central(h):
      lp = rollout_mm_loss(handle, Z + h * v[0], ell + h * v[1], var, q + h * v[2], m0, S0, *args, **kw)
      lm = rollout_mm_loss(handle, Z - h * v[0], ell - h * v[1], var, q - h * v[2], m0, S0, *args, **kw)
      return (lp - lm) / (2 * h)                                                 # [R]: the restarts are independent
    # one Richardson step, O(h^4) truncation; the step is large enough for the round-off of a 100-step loss (~1e-10) to stay below 1e-6
    fd = (4.0 * central(1e-4) - central(2e-4)) / 3.0
    an = sum((g * d).reshape(R, -1).sum(-1) for g, d in zip(g_p, v))
    scaled_close(an, fd, 2e-5, f"directional derivative {trial} (H = 100)")


def test_config5_shape_backward_vs_oracle_autograd():
  """Same models at M = 256 with a shorter horizon (H = 12, 2 restarts) against torch.autograd on the oracle (CUDA's O(M^2) association,
  oracle/gp_models.py mm_sparse_reassociated): full-size Psi2 tiles, all 10 kernel pairs, both row blocks of the gradient contraction."""
  from gpflowpilco_b200.autograd import rollout_mm_loss
  from oracle import gp_models as gm
  from oracle import moments as mo
  from oracle import rollout as ro
  R, H = 2, 12
  cfg, handle, Z, ell, var, q, m0, S0 = _config5_problem(R, seed=9)
  d, p = cfg["dynamics"], cfg["policy"]
  dyn = gm.SVGPModel([ps.SEKernel(float(d["variance"][l]), torch.as_tensor(d["lengthscales"][l])) for l in range(4)],
                     [torch.as_tensor(d["Z"][l]) for l in range(4)], torch.as_tensor(d["q_mu"]), torch.as_tensor(d["q_sqrt"]), whiten=True,
                     mean_const=torch.as_tensor(d["mean_const"]))
  enc = mo.TrigonometricEncoder(tuple(cfg["active_dims"]))
  obj = mo.GaussianObjective(torch.as_tensor(cfg["target"]), torch.as_tensor(cfg["W"]))
  ref_loss, ref_g = [], []
  for r in range(R):
    Zr, er, qr = (t[r].cpu().clone().requires_grad_(True) for t in (Z, ell, q))
    pol = gm.SVGPModel([ps.SEKernel(float(var[r]), er)], [Zr], qr[:, None], torch.as_tensor(p["q_sqrt"]), whiten=True,
                       mean_const=torch.zeros(1, dtype=DTYPE))
    loss = ro.mm_rollout(m0[r:r + 1].cpu(), S0[r:r + 1].cpu(), H, lambda s: gm.mm_sparse_reassociated(s, dyn),
                         lambda s: gm.mm_policy(s, pol, cfg["squash_scale"], cfg["squash_shift"]), enc, obj)
    ref_loss.append(loss.detach()[0])
    ref_g.append(torch.autograd.grad(loss.sum(), [Zr, er, qr]))
  leaves = [t.clone().requires_grad_(True) for t in (Z, ell, q)]
  loss = rollout_mm_loss(handle, leaves[0], leaves[1], var, leaves[2], m0, S0, H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]),
                         squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  loss.sum().backward()
  scaled_close(loss, torch.stack(ref_loss), 1e-6, "loss")
  for k, name in enumerate(("Z", "lengthscales", "q_mu")):
    scaled_close(leaves[k].grad, torch.stack([g[k] for g in ref_g]), 1e-6, f"{name} gradient")
