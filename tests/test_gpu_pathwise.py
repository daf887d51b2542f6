"""GPU parity: fused pathwise particle rollout vs the oracle's decoupled-sampler contract (oracle/pathwise.py,
oracle/rollout.py::pathwise_rollout; upstream loops/pilco.py:263-303).  Same draws on both sides (oracle Philox streams).
Tolerance 1e-6 relative; chaotic amplification over long horizons is avoided by short horizons in the tight check."""
import numpy as np
import pytest
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from oracle import pathwise as pw
from oracle import rollout as ro
from tests.helpers import DTYPE, oracle_svgp, scaled_close

pytestmark = pytest.mark.gpu


def _dev(x):
  return torch.as_tensor(x, dtype=DTYPE, device="cuda")


def _setup(S, F, M, Mp, seed=3):
  cfg = synthetic.config1_cartpole(M=M, Mp=Mp)
  dyn, pol = oracle_svgp(cfg["dynamics"]), oracle_svgp(cfg["policy"])
  paths = pw.generate_paths(dyn, F, seed, 0, S)
  x0 = pw.draw_initial_states(torch.as_tensor(cfg["m0"][0]), torch.linalg.cholesky(torch.as_tensor(cfg["S0"][0])), seed, 0, S)
  return cfg, dyn, pol, paths, x0


def _run_cuda(cfg, paths, x0, H, traj=False):
  from gpflowpilco_b200.pathwise import PackedPaths, rollout_pathwise
  from gpflowpilco_b200.rollouts import PolicyParams
  d, p = cfg["dynamics"], cfg["policy"]
  packed = PackedPaths.from_sample_major(_dev(d["Z"]), _dev(d["lengthscales"]), _dev(d["variance"]), _dev(d["mean_const"]),
                                         _dev(paths.omega), _dev(paths.phase), _dev(paths.w), _dev(paths.v))
  P = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=True,
                   squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  return rollout_pathwise(packed, P, _dev(x0), H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]), return_trajectory=traj)


@pytest.mark.parametrize("S,F,M,H", [(7, 64, 20, 3), (300, 128, 48, 5), (256, 256, 64, 2)])
def test_pathwise_rollout_matches_oracle(S, F, M, H):
  cfg, dyn, pol, paths, x0 = _setup(S, F, M, 12)
  enc = mo.TrigonometricEncoder(cfg["active_dims"])
  obj = mo.GaussianObjective(cfg["target"], cfg["W"])
  loss_ref, traj_ref = ro.pathwise_rollout(x0, H, lambda eu: pw.evaluate_paths(dyn, paths, eu),
                                           lambda e: gm.policy_sample_path(pol, e, cfg["squash_scale"], cfg["squash_shift"]), enc, obj, True)
  loss, xf, traj = _run_cuda(cfg, paths, x0, H, traj=True)
  scaled_close(traj, torch.stack(traj_ref), 1e-8, "particle trajectories")
  scaled_close(loss, loss_ref, 1e-8, "losses")
  scaled_close(xf, traj_ref[-1], 1e-8, "final states")


def test_pathwise_mean_loss_close_to_moment_matching():
  """Sanity across the two estimators: the particle average of the cost approaches the moment-matched expected cost for a
  short horizon (both approximate the same expectation; MC error ~ 1/sqrt(S), MM error from Gaussian projection)."""
  from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm
  from tests.helpers import cuda_handle
  S, F, M, H = 4096, 256, 32, 2
  cfg, dyn, pol, paths, x0 = _setup(S, F, M, 12, seed=11)
  loss, _, _ = _run_cuda(cfg, paths, x0, H)
  p = cfg["policy"]
  P = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=True,
                   squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  mm = rollout_mm(cuda_handle(cfg["dynamics"]), P, _dev(cfg["m0"]), _dev(cfg["S0"]), H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]))
  assert abs(float(loss.mean()) - float(mm.loss[0])) < 0.05 * max(abs(float(mm.loss[0])), 1e-3) + 5.0 / np.sqrt(S) * float(loss.std())


def test_chunked_rollout_with_overlapped_generation_matches_serial():
  """rollout_pathwise_chunked (path generation of chunk k+1 on a second stream while chunk k rolls out) returns bit for bit what the
  serial generate -> rollout loop returns: same Philox streams by global particle index, same chunking, same summation order."""
  from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths, rollout_pathwise, rollout_pathwise_chunked
  from gpflowpilco_b200.rollouts import PolicyParams
  from tests.helpers import cuda_handle
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  p = cfg["policy"]
  handle = cuda_handle(cfg["dynamics"])
  P = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=True,
                   squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  m0, S0 = _dev(cfg["m0"][0]), _dev(cfg["S0"][0])
  total, per, F, H, seed, first = 1000, 256, 64, 3, 9, 12345
  args = (H, cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]))
  ref = torch.zeros((), dtype=DTYPE, device="cuda")
  for o in range(0, total, per):
    n = min(per, total - o)
    paths = generate_paths(handle, n, F, seed, first_particle=first + o)
    x0 = draw_initial_states(m0, S0, seed, first + o, n)
    loss, _, _ = rollout_pathwise(paths, P, x0, *args)
    ref = ref + loss.sum()
  for overlap in (True, False):
    got = rollout_pathwise_chunked(handle, P, m0, S0, total, F, seed, *args, first_particle=first, particles_per_launch=per, overlap=overlap)
    torch.cuda.synchronize()
    assert torch.equal(got, ref), (overlap, float(got), float(ref))


def test_mixed_precision_rollout_within_its_stated_tolerance():
  """gpp_rollout_pathwise_fwd_mixed (FP32 Fourier weights + FP32 cosine polynomial; phases, reduction, canonical part, policy, cost
  FP64) against the all-FP64 kernel on the same paths: one step's drift within 5e-6 of the largest drift entry (the tolerance the
  header states), and a 10-step rollout's per-particle loss and final state within 2e-5 (the dynamics amplify the per-step
  difference), mean loss within 2e-6."""
  from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths, rollout_pathwise
  from gpflowpilco_b200.rollouts import PolicyParams
  from tests.helpers import cuda_handle
  cfg = synthetic.config1_cartpole(M=64, Mp=10)
  p = cfg["policy"]
  handle = cuda_handle(cfg["dynamics"])
  P = PolicyParams(_dev(p["Z"]), _dev(p["lengthscales"]), _dev(p["variance"]), _dev(p["q_mu"][:, 0][None]), whiten=True,
                   squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  S, F = 700, 1024
  paths = generate_paths(handle, S, F, seed=5, first_particle=0)
  x0 = draw_initial_states(_dev(cfg["m0"][0]), _dev(cfg["S0"][0]), 5, 0, S)
  args = (cfg["active_dims"], _dev(cfg["target"]), _dev(cfg["W"]))
  _, x1_64, _ = rollout_pathwise(paths, P, x0, 1, *args)
  _, x1_mx, _ = rollout_pathwise(paths, P, x0, 1, *args, mixed_precision=True)
  assert paths.w32 is not None and paths.w32.dtype == torch.float32
  d64, dmx = x1_64 - x0, x1_mx - x0
  step_err = float((dmx - d64).abs().max() / d64.abs().max())
  print(f"[mixed] one-step drift: max |diff| / max |drift| = {step_err:.2e}")
  assert 0.0 < step_err <= 5e-6, step_err
  l64, xf64, _ = rollout_pathwise(paths, P, x0, 10, *args)
  lmx, xfmx, _ = rollout_pathwise(paths, P, x0, 10, *args, mixed_precision=True)
  e_loss = float((lmx - l64).abs().max() / l64.abs().max())
  e_x = float((xfmx - xf64).abs().max() / xf64.abs().max())
  e_mean = float((lmx.mean() - l64.mean()).abs() / l64.mean().abs())
  print(f"[mixed] H=10: per-particle loss {e_loss:.2e}, final state {e_x:.2e}, mean loss {e_mean:.2e}")
  assert e_loss <= 2e-5 and e_x <= 2e-5 and e_mean <= 2e-6, (e_loss, e_x, e_mean)
