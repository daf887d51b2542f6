"""Gradient-mode pathwise rollout timing (developer tool).  usage: python scripts/prof_pathwise_grad.py [particles] [horizon]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import ops, synthetic
from gpflowpilco_b200.autograd import rollout_pathwise_loss
from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths

S = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
cfg = synthetic.config1_cartpole()
d, p = cfg["dynamics"], cfg["policy"]
handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
paths = generate_paths(handle, S, 4096, seed=0)
x0 = draw_initial_states(T(cfg["m0"][0]), T(cfg["S0"][0]), 0, 0, S)
Z, e, q = [T(x).clone().requires_grad_(True) for x in (p["Z"], p["lengthscales"], p["q_mu"][:, 0][None])]
for it in range(3):
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  ev[0].record()
  loss = rollout_pathwise_loss(paths, Z, e, T(p["variance"]), q, x0, H, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]),
                               squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  ev[1].record()
  loss.sum().backward()
  ev[2].record()
  torch.cuda.synchronize()
  print(f"iter {it}: forward(grad mode) {ev[0].elapsed_time(ev[1]):.1f} ms, backward {ev[1].elapsed_time(ev[2]):.1f} ms, "
        f"{S * H / ev[0].elapsed_time(ev[2]) * 1e3 / 1e6:.2f} M particle-steps/s")
  Z.grad = e.grad = q.grad = None
