"""Two ranks on two GPUs over NCCL (skipped on a single-GPU box; the host logic is covered by the gloo tests on CPU):
the particle-sharded pathwise closure and the restart-sharded MM closure give the single-process results."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _problem(dev):
  from gpflowpilco_b200 import ops, synthetic
  cfg = synthetic.config1_cartpole(M=32, Mp=8)
  cfg["policy"]["q_mu"] = 100.0 * cfg["policy"]["q_mu"]
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
  return cfg, T, handle, p


def _closures(dev):
  from gpflowpilco_b200 import distributed as gd
  cfg, T, handle, p = _problem(dev)
  args = (handle, T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), T(cfg["m0"][0]), T(cfg["S0"][0]))
  kw = dict(total_particles=1000, num_bases=64, seed=4, horizon=3, active_dims=cfg["active_dims"], cost_target=T(cfg["target"]),
            cost_W=T(cfg["W"]), squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  loss, grads = gd.pathwise_policy_loss_and_grad(*args, **kw)
  R = 5
  g = torch.Generator().manual_seed(0)
  Z = T(p["Z"]).repeat(R, 1, 1) + 0.1 * torch.randn(R, *p["Z"].shape[1:], dtype=torch.float64, generator=g).to(dev)
  ell, q, var = T(p["lengthscales"]).repeat(R, 1), T(p["q_mu"][:, 0][None]).repeat(R, 1), T(p["variance"]).repeat(R)
  losses, (start, count), rg = gd.mm_restart_losses_and_grads(handle, Z, ell, var, q, T(cfg["m0"]), T(cfg["S0"]), 4, cfg["active_dims"],
                                                              T(cfg["target"]), T(cfg["W"]), cfg["squash_scale"], cfg["squash_shift"])
  return float(loss), [x.cpu() for x in grads], losses.cpu(), start, count, [x.cpu() for x in rg]


def _worker(rank, ws, port, ret):
  import torch.distributed as dist
  os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
  try:
    ret[rank] = _closures(torch.device("cuda", rank))
  finally:
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_match_one():
  import torch.multiprocessing as mp
  single = _closures(torch.device("cuda", 0))
  ret = mp.Manager().dict()
  mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
  for rank in range(2):
    loss, grads, losses, start, count, rg = ret[rank]
    assert abs(loss - single[0]) <= 1e-12 * abs(single[0])
    for a, b in zip(grads, single[1]):
      assert float((a - b).abs().max()) <= 1e-10 * float(b.abs().max())
    assert float((losses - single[2]).abs().max()) <= 1e-12 * float(single[2].abs().max())     # all-gathered, global order
    for a, b in zip(rg, single[5]):
      assert float((a - b[start:start + count]).abs().max()) <= 1e-10 * float(b.abs().max())
  assert ret[0][0] == ret[1][0]
  assert sorted((ret[0][3], ret[1][3])) == [0, 3]                           # 5 restarts -> blocks [0,3) and [3,5)


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_two_devices_in_one_process():
  """A model handle lives on the device of its tensors; the library's per-device state (shared-memory opt-ins, SM count) and the
  host side's device context follow it even when another device is current.  Inputs from the wrong device are refused."""
  from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm, rollout_mm_bwd
  out = []
  rng = np.random.default_rng(0)
  xin = rng.standard_normal((3, 6))
  A = 0.2 * rng.standard_normal((3, 6, 6))
  Sin = A @ A.transpose(0, 2, 1) + 0.05 * np.eye(6)
  for dev in ("cuda:0", "cuda:1"):
    cfg, T, handle, p = _problem(dev)
    torch.cuda.set_device(0)                       # the handle on cuda:1 is used while cuda:0 is current
    f1, Sff, cross = handle.predict(T(xin), T(Sin))
    pol = PolicyParams(T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), whiten=True,
                       squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
    res = rollout_mm(handle, pol, T(cfg["m0"]), T(cfg["S0"]), 6, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), save_for_backward=True)
    grads = rollout_mm_bwd(handle, pol, pol.beta(), res.traj_m, res.traj_S, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), saved=res.saved)
    out.append([t.cpu() for t in (f1, Sff, cross, res.loss, *grads)])
  for a, b in zip(*out):
    assert torch.equal(a, b)
  cfg, T, handle, p = _problem("cuda:1")
  with pytest.raises(ValueError):
    handle.predict(torch.as_tensor(xin, dtype=torch.float64, device="cuda:0"), torch.as_tensor(Sin, dtype=torch.float64, device="cuda:0"))
