"""MM rollout forward captured in a CUDA graph (the C entry points only enqueue work on the caller's stream).  Developer tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import ops, synthetic
from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
cfg = synthetic.config1_cartpole()
d, p = cfg["dynamics"], cfg["policy"]
handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
pol = PolicyParams(T(p["Z"]), T(p["lengthscales"]), T(p["variance"]), T(p["q_mu"][:, 0][None]), whiten=True, squash_scale=cfg["squash_scale"],
                   squash_shift=cfg["squash_shift"])
beta = pol.beta()
m0, S0 = T(cfg["m0"]).expand(N, -1).contiguous(), T(cfg["S0"]).expand(N, -1, -1).contiguous()
tgt, W = T(cfg["target"]), T(cfg["W"])
run = lambda: rollout_mm(handle, pol, m0, S0, H, cfg["active_dims"], tgt, W, beta=beta, check=False).loss
for _ in range(3):
  ref = run()
torch.cuda.synchronize()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
  run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
  out = run()
for name, fn in (("eager", run), ("graph replay", g.replay)):
  ts = []
  for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  print(f"{name}: N={N} H={H} {min(ts):.3f} ms")
print("loss eager", float(ref[0]), "graph", float(out[0]))
