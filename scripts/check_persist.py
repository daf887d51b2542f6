"""Persistent (on-device H-loop) vs one-launch-per-stage moment-matched rollouts: agreement and timing.  Developer tool.
usage: python scripts/check_persist.py [quick|time]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import _lib, ops, rollouts, synthetic

lib = _lib.load()
dev = torch.device("cuda")
T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)


def setup(R, M=None, seed=5):
  cfg = synthetic.config1_cartpole() if M is None else synthetic.config1_cartpole(M=M)
  d, p = cfg["dynamics"], cfg["policy"]
  handle = ops.GPModelHandle(T(d["Z"]), T(d["lengthscales"]), T(d["variance"]), T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_const=T(d["mean_const"]))
  g = torch.Generator().manual_seed(seed)
  Z = T(p["Z"]).repeat(R, 1, 1) + 0.3 * torch.randn(R, *p["Z"].shape[1:], dtype=torch.float64, generator=g).to(dev)
  ell = T(p["lengthscales"]).repeat(R, 1)
  q = 1e-3 * torch.randn(R, p["Z"].shape[1], dtype=torch.float64, generator=g).to(dev)
  var = T(p["variance"]).repeat(R)
  pol = rollouts.PolicyParams(Z, ell, var, q, squash_scale=cfg["squash_scale"], squash_shift=cfg["squash_shift"])
  m0, S0 = T(cfg["m0"]).expand(R, -1).contiguous(), T(cfg["S0"]).expand(R, -1, -1).contiguous()
  return cfg, handle, pol, m0, S0


def run(mode, cfg, handle, pol, m0, S0, H, timing=False):
  rollouts.set_rollout_mode(mode)
  beta = pol.beta()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
  n0 = lib.gpp_launch_count()
  ev[0].record()
  res = rollouts.rollout_mm(handle, pol, m0, S0, H, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), beta=beta, save_for_backward=True, check=False)
  ev[1].record()
  grads = rollouts.rollout_mm_bwd(handle, pol, beta, res.traj_m, res.traj_S, cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), saved=res.saved, check=False)
  ev[2].record()
  torch.cuda.synchronize()
  return res, grads, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), lib.gpp_launch_count() - n0


def rel(a, b):
  return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def compare(R, H, M=None):
  args = setup(R, M)
  ref, gref, *_ = run(rollouts.ROLLOUT_LEGACY, *args, H)
  out, gout, *_ = run(rollouts.ROLLOUT_PERSIST, *args, H)
  errs = {"loss": rel(out.loss, ref.loss), "traj_m": rel(out.traj_m, ref.traj_m), "traj_S": rel(out.traj_S, ref.traj_S),
          "m_final": rel(out.m_final, ref.m_final), "S_final": rel(out.S_final, ref.S_final)}   # (`saved` holds uninitialised padding)
  for name, a, b in zip(("Z_bar", "ell_bar", "beta_bar", "m0_bar", "S0_bar"), gout, gref):
    errs[name] = rel(a, b)
  worst = max(errs.values())
  print(f"R={R} H={H} M={M or 256}: worst rel diff persist vs legacy {worst:.2e}  " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)
  return worst


def timing(R, H, reps=3):
  args = setup(R)
  for mode, name in ((rollouts.ROLLOUT_LEGACY, "legacy "), (rollouts.ROLLOUT_PERSIST, "persist")):
    for it in range(reps):
      _, _, f, b, nl = run(mode, *args, H)
      print(f"{name} R={R} H={H} iter {it}: forward {f:.3f} ms, backward {b:.3f} ms, total {f + b:.3f} ms, {nl} launches, "
            f"{R * H / (f + b) * 1e3:.0f} rollout-steps/s", flush=True)


if __name__ == "__main__":
  what = sys.argv[1] if len(sys.argv) > 1 else "quick"
  if what == "quick":
    worst = 0.0
    for R, H, M in ((1, 3, 40), (3, 5, 40), (2, 4, 100), (1, 30, None), (8, 6, None), (64, 4, None), (150, 3, 40)):
      worst = max(worst, compare(R, H, M))
    print("WORST", worst)
    sys.exit(0 if worst < 1e-7 else 1)
  if what == "time":
    timing(1, 30)
    timing(64, 100)


def sweep():
  Rs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else (1, 4, 8, 16, 32, 64, 96, 128)
  for R in Rs:
    args = setup(R)
    rollouts.set_rollout_mode(rollouts.ROLLOUT_PERSIST)
    best = (1e9, 1e9)
    for it in range(3):
      _, _, f, b, nl = run(rollouts.ROLLOUT_PERSIST, *args, 50)
      best = (min(best[0], f), min(best[1], b))
    print(f"persist R={R} H=50: forward {best[0] / 50 * 1e3:.1f} us/step, backward {best[1] / 50 * 1e3:.1f} us/step", flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "sweep":
  sweep()


if len(sys.argv) > 1 and sys.argv[1] == "fuzz":
  # rollout counts around the thresholds of the forward hand-over (urgent / deferred partials at 16, continuous mode from 2), odd
  # counts, one-step and longer horizons, one / several tiles per CTA
  worst = 0.0
  for R, H, M in ((2, 1, None), (3, 2, None), (15, 3, None), (16, 3, None), (17, 3, None), (31, 2, None), (33, 5, None), (100, 2, None),
                  (17, 4, 100), (40, 3, 100), (5, 7, 40), (64, 1, None), (129, 2, None)):
    worst = max(worst, compare(R, H, M))
  print("WORST", worst)
  sys.exit(0 if worst < 1e-7 else 1)
