"""Policy optimisation on the cart-pole swing-up models with the B200 path behind the reference's API.

Mirrors the policy-update half of upstream examples/cartpole_swingup (experiment.py:148-160, train_utils.py:91-135):
build the loop object, take its policy-loss closure, minimise it with clipped Adam.  The environment / data-collection half
is out of scope: the dynamics model here is the synthetic config #1 SVGP (gpflowpilco_b200/synthetic.py).

  python examples/cartpole_policy_optimisation.py [--steps 50] [--graph]

`--graph` evaluates loss and gradients by replaying one captured CUDA graph (gpflowpilco_b200/graphs.py) instead of launching
the ~280 kernels of a forward + backward rollout from Python every step.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gpflowpilco_b200 import models as M
from gpflowpilco_b200 import synthetic
from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder
from gpflowpilco_b200.loops import EpisodeSpec, GaussianStateDistribution, MomentMatchingPILCO
from gpflowpilco_b200.utils.optimizers import GradientDescent, clip_by_global_norm


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--horizon", type=float, default=3.0)
  ap.add_argument("--graph", action="store_true")
  args = ap.parse_args()
  dev = torch.device("cuda")
  T = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)
  cfg = synthetic.config1_cartpole()
  d, p = cfg["dynamics"], cfg["policy"]
  L = d["Z"].shape[0]
  drift = M.SVGP(M.SeparateIndependent([M.SquaredExponential(T(d["variance"][l]), T(d["lengthscales"][l])) for l in range(L)]),
                 M.SeparateIndependentInducingVariables([M.InducingPoints(T(d["Z"][l])) for l in range(L)]),
                 T(d["q_mu"]), T(d["q_sqrt"]), whiten=True, mean_function=M.Constant(T(d["mean_const"])))
  pol = M.SVGP(M.SeparateIndependent([M.SquaredExponential(T(p["variance"][0]), T(p["lengthscales"][0]))]),
               M.SeparateIndependentInducingVariables([M.InducingPoints(T(p["Z"][0]))]), T(p["q_mu"]), T(p["q_sqrt"]), whiten=True,
               mean_function=M.Constant(T(p["mean_const"])))
  link = M.BijectorChain([M.Scale(cfg["squash_scale"]), M.Shift(cfg["squash_shift"]), M.NormalCDF()])
  policy = M.InverseLinkWrapper(M.KernelRegressor(pol), invlink=link)
  # trainable variables of the policy (upstream loops/pilco.py:99-103: centres, q_mu and lengthscales; variance and q_sqrt frozen)
  pol.q_mu = pol.q_mu.clone().requires_grad_(True)
  Zvar = pol.latent_inducing()[0].clone().requires_grad_(True)
  pol.inducing_variable.inducing_variables[0].Z = Zvar
  spec = EpisodeSpec(GaussianStateDistribution(T(cfg["m0"][0]), T(cfg["S0"][0])), horizon=args.horizon, step_size=0.1)
  loop = MomentMatchingPILCO(spec, GaussianObjective(T(cfg["target"]), T(cfg["W"])), drift, policy, TrigonometricEncoder(cfg["active_dims"]))
  closure = loop.policy_loss_closure()
  opt = GradientDescent(step_limit=args.steps, optimizer_factory=lambda vs: torch.optim.Adam(vs, lr=1e-2), transform=clip_by_global_norm(1.0))
  closure().sum().backward()                     # warm-up: CUDA context, lazy module load, handle build
  pol.q_mu.grad = Zvar.grad = None
  t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
  skip = min(5, args.steps - 1)                  # the first optimiser steps pay one-off host costs (torch.optim's lazy imports)
  if args.graph:
    from gpflowpilco_b200.graphs import GraphedMMPolicyGradient
    from gpflowpilco_b200.moment_matching.models import svgp_handle
    k = pol.latent_kernels()[0]
    ell, var = k.ell(Zvar.shape[-1])[None], k.variance.reshape(1)
    m0, S0 = loop.get_state_initializer(spec.state_distrib)()
    g = GraphedMMPolicyGradient(svgp_handle(drift, True), Zvar[None], ell, var, pol.q_mu[:, 0][None], m0, S0, spec.num_steps,
                                cfg["active_dims"], T(cfg["target"]), T(cfg["W"]), squash_scale=cfg["squash_scale"],
                                squash_shift=cfg["squash_shift"])
    adam = torch.optim.Adam([pol.q_mu, Zvar], lr=1e-2)
    clip = clip_by_global_norm(1.0)
    first = None
    for i in range(args.steps):
      if i == skip:
        t0.record()
      loss, (dZ, _, dq) = g(Zvar.detach()[None], ell, pol.q_mu.detach()[:, 0][None], check=False)
      first = loss.clone() if first is None else first          # the graph's output buffer is reused by the next replay
      pol.q_mu.grad, Zvar.grad = clip(dq[0][:, None], dZ[0])
      adam.step()
    t1.record(); torch.cuda.synchronize()
    hist = [float(first.mean()), float(g.loss.mean())]
  else:
    opt.callbacks.append(lambda step, *_: t0.record() if step == skip - 1 else None)
    hist = opt.minimize(closure, [pol.q_mu, Zvar])
    t1.record(); torch.cuda.synchronize()
  print(f"{args.steps} policy-optimisation steps (H = {spec.num_steps}): expected cost {hist[0]:.6f} -> {hist[-1]:.6f}, "
        f"{t0.elapsed_time(t1) / (args.steps - skip):.2f} ms per step (forward + backward + clip + Adam; first {skip} steps not timed)")


if __name__ == "__main__":
  main()
