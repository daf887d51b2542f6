"""Short pathwise rollout for ncu captures (developer tool; bench.py is the contract).
usage: python scripts/prof_pathwise.py [particles] [horizon] [bases] [mixed]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import _lib, ops, synthetic
from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths, rollout_pathwise
from gpflowpilco_b200.rollouts import PolicyParams

S = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 512
H = int(sys.argv[2]) if len(sys.argv) > 2 else 10
F = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
MIXED = len(sys.argv) > 4 and sys.argv[4] == "mixed"
lib = _lib.load()
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
cfg = synthetic.config1_cartpole()
d, p = cfg["dynamics"], cfg["policy"]
handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True,
                           mean_const=T(d["mean_const"]))
policy = PolicyParams(T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), whiten=True,
                      squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
paths = generate_paths(handle, S, F, seed=0, first_particle=0)
x0 = draw_initial_states(T(cfg["m0"][0]), T(cfg["S0"][0]), 0, 0, S)
beta = policy.beta()
lib.gpp_profile_enable(1)
for it in range(3):
  loss, _, _ = rollout_pathwise(paths, policy, x0, H, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), beta=beta, mixed_precision=MIXED)
  ms = ctypes.c_float()
  lib.gpp_profile_last_ms(ctypes.byref(ms))
  L, M, D = 4, d["Z"].shape[1], 6
  by = (4 if MIXED else 8) * L * F + 8 * L * M + 64
  print(f"iter {it}: {ms.value:.2f} ms, {S * H / ms.value * 1e3 / 1e6:.2f} M particle-steps/s, {by * S * H / ms.value * 1e3 / 1e9:.0f} GB/s algorithmic, "
        f"mean loss {float(loss.mean()):.6f}")
