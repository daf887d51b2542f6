"""MM rollout forward + backward timing (config #1: N=1, H=30; config #5 per-GPU share: R restarts, H=100).  Developer tool.
usage: python scripts/prof_rollout.py [R] [H]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import _lib, ops, synthetic
from gpflowpilco_b200.autograd import rollout_mm_loss

R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lib = _lib.load()
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
cfg = synthetic.config1_cartpole()
d, p = cfg["dynamics"], cfg["policy"]
handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
g = torch.Generator().manual_seed(5)
Z = (T(p["Z"]).repeat(R, 1, 1) + 0.3 * torch.randn(R, *p["Z"].shape[1:], dtype=torch.float64, generator=g).to(dev)).requires_grad_(True)
ell = T(p["lengthscales"]).repeat(R, 1).requires_grad_(True)
q = (1e-3 * torch.randn(R, p["Z"].shape[1], dtype=torch.float64, generator=g).to(dev)).requires_grad_(True)
var = T(p["variance"]).repeat(R)
m0, S0 = T(cfg["m0"]).expand(R, -1).contiguous(), T(cfg["S0"]).expand(R, -1, -1).contiguous()
for it in range(3):
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  n0 = lib.gpp_launch_count()
  ev[0].record()
  loss = rollout_mm_loss(handle, Z, ell, var, q, m0, S0, H, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), squash_scale=cfg["squash_scale"],
                         squash_shift=cfg["squash_shift"])
  ev[1].record()
  loss.sum().backward()
  ev[2].record()
  torch.cuda.synchronize()
  print(f"iter {it}: R={R} H={H} forward {ev[0].elapsed_time(ev[1]):.2f} ms, backward {ev[1].elapsed_time(ev[2]):.2f} ms, "
        f"{lib.gpp_launch_count() - n0} launches, {R * H / ev[0].elapsed_time(ev[2]) * 1e3:.0f} rollout-steps/s, loss {float(loss.mean()):.6f}")
