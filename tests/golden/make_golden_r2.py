"""Round-2 fixtures, generated like make_golden.py: the UNMODIFIED upstream sources (default /root/reference) executed on the numpy
stand-ins of oracle/refshim.  Run from the repo root:   python tests/golden/make_golden_r2.py [reference_root]

  rollout_cfg1.npz          BASELINE config #1 at full size: cart-pole models with M = 256 inducing points per latent, 30 policy centres,
                            H = 30 steps of forward_sde + MomentMatchingEuler + the loss callback (loops/pilco.py:192-220); loss and
                            the trajectory of moments.
  rollout_grads.npz         gradients of upstream's own closure value (H = 5, the models of rollout.npz) w.r.t. the policy centres Z,
                            q_mu, lengthscales and the initial moments (m0, S0): central differences at three step sizes with two Richardson steps
                            (upstream differentiates with tf.GradientTape, utils/optimizers.py:52-56; TensorFlow is absent here, so the
                            pinned quantity is the derivative of the function upstream evaluates).
  forward_sde_variants.npz  one step of forward_sde (dynamics/forward_sde.py:34-92) and two steps of MomentMatchingEuler for the variants
                            without a policy: (no encoder, no policy) and (encoder, no policy).  The (policy, no encoder) registration
                            (:49-69) multiplies a [Dx, Du] block by a [Dx + Du, L] one and cannot run upstream for Du < Dx + Du; it has
                            no fixture (the facade implements the stated intent, checked against the oracle and Monte Carlo).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def fd_gradients(closure, synthetic, np):
  cfg = synthetic.config1_cartpole(M=40, Mp=12)
  d, p = cfg["dynamics"], dict(cfg["policy"])
  p["q_mu"] = 300.0 * p["q_mu"]            # 1e-3 N(0,1) weights leave the policy in the flat middle of the squashing link
  H = 5

  def value(Z=None, q_mu=None, ell=None, m0=None, S0=None):
    pp = dict(p)
    if Z is not None: pp["Z"] = Z
    if q_mu is not None: pp["q_mu"] = q_mu
    if ell is not None: pp["lengthscales"] = ell
    return float(closure(cfg, d, pp, H, cfg["m0"] if m0 is None else m0, cfg["S0"] if S0 is None else S0)[0])

  def fd(name, base, sym=False, h=4e-3):
    """central differences at h, h/2, h/4 with two Richardson steps (error O(h^6)); `h` is relative to max(1, |x|)"""
    g = np.zeros_like(base)
    it = np.nditer(base, flags=["multi_index"])
    for _ in it:
      idx = it.multi_index
      if sym and idx[-1] < idx[-2]:
        continue
      h0 = h * max(1.0, abs(float(base[idx])))

      def D(step):
        e = np.zeros_like(base)
        e[idx] = step
        if sym and idx[-1] != idx[-2]:
          e[idx[:-2] + (idx[-1], idx[-2])] = step
        return (value(**{name: base + e}) - value(**{name: base - e})) / (2 * step)
      d1, d2, d4 = D(h0), D(h0 / 2), D(h0 / 4)
      r1, r2 = (4.0 * d2 - d1) / 3.0, (4.0 * d4 - d2) / 3.0
      g[idx] = (16.0 * r2 - r1) / 15.0
      if sym and idx[-1] != idx[-2]:
        g[idx] *= 0.5                                   # the symmetric perturbation moved two entries: dL = sum_ij Sbar_ij dS_ij
        g[idx[:-2] + (idx[-1], idx[-2])] = g[idx]
    return g

  base_loss = value()
  grads = {"Z": fd("Z", p["Z"]), "q_mu": fd("q_mu", p["q_mu"]), "ell": fd("ell", p["lengthscales"]),
           "m0": fd("m0", cfg["m0"]), "S0": fd("S0", cfg["S0"], sym=True, h=1e-3)}
  # self-check of the differences: another base step must agree
  chk = fd("ell", p["lengthscales"], h=2e-3)
  print("FD self-check (lengthscales, two base steps): rel diff", np.abs(chk - grads["ell"]).max() / np.abs(grads["ell"]).max())
  chk = fd("m0", cfg["m0"], h=2e-3)
  print("FD self-check (m0, two base steps): rel diff", np.abs(chk - grads["m0"]).max() / np.abs(grads["m0"]).max())
  np.savez(os.path.join(OUT, "rollout_grads.npz"), horizon=H, loss=base_loss, g_Z=grads["Z"], g_q_mu=grads["q_mu"], g_ell=grads["ell"],
           g_m0=grads["m0"], g_S0=grads["S0"], pol_Z=p["Z"], pol_ell=p["lengthscales"], pol_var=p["variance"], pol_q_mu=p["q_mu"],
           pol_q_sqrt=p["q_sqrt"], dyn_Z=d["Z"], dyn_ell=d["lengthscales"], dyn_var=d["variance"], dyn_q_mu=d["q_mu"], dyn_q_sqrt=d["q_sqrt"],
           dyn_c=d["mean_const"], m0=cfg["m0"], S0=cfg["S0"], target=cfg["target"], W=cfg["W"], scale=cfg["squash_scale"],
           shift=cfg["squash_shift"], active_dims=np.array(cfg["active_dims"]))
  print("gradients: |g| max", {k: float(np.abs(v).max()) for k, v in grads.items()})


def main(reference_root="/root/reference", only=""):
  from oracle import refshim
  refshim.install(reference_root)
  import gpflow
  import tensorflow as tf
  from gpflow_pilco.components import GaussianObjective, TrigonometricEncoder
  from gpflow_pilco.dynamics.dynamical_system import DynamicalSystem
  from gpflow_pilco.dynamics.forward_sde import forward_sde
  from gpflow_pilco.dynamics.solvers import MomentMatchingEuler
  from gpflow_pilco.models import InverseLinkWrapper, KernelRegressor
  from gpflow_pilco.moment_matching import GaussianMoments, moment_matching
  from tensorflow_probability.python import bijectors as tfb

  from gpflowpilco_b200 import synthetic
  from gpflowpilco_b200.synthetic import generate_covariance

  def A(x):
    return np.asarray(x.to_dense() if hasattr(x, "to_dense") else x)

  SE, IP = gpflow.kernels.SquaredExponential, gpflow.inducing_variables.InducingPoints

  def svgp(Z, ell, var, q_mu, q_sqrt, c, whiten=True):
    L = Z.shape[0]
    return gpflow.models.SVGP(
        kernel=gpflow.kernels.SeparateIndependent([SE(var[l], ell[l]) for l in range(L)]),
        inducing_variable=gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(Z[l]) for l in range(L)]),
        q_mu=q_mu, q_sqrt=q_sqrt, num_latent_gps=L, whiten=whiten, mean_function=gpflow.mean_functions.Constant(c))

  def closure(cfg, d, p, H, m0, S0, record=None):
    """upstream's MM closure (loops/pilco.py:207-217) for the given parameter dictionaries; returns the loss [1]"""
    link = tfb.Chain(bijectors=[tfb.Scale(scale=tf.cast(cfg["squash_scale"], np.float64)), tfb.Shift(shift=tf.cast(cfg["squash_shift"], np.float64)),
                                tfb.NormalCDF()])
    drift = svgp(d["Z"], d["lengthscales"], d["variance"], d["q_mu"], d["q_sqrt"], d["mean_const"])
    pol = svgp(p["Z"], p["lengthscales"], p["variance"], p["q_mu"], p["q_sqrt"], p["mean_const"])
    policy = InverseLinkWrapper(model=KernelRegressor(model=pol), invlink=link)
    encoder = TrigonometricEncoder(active_dims=cfg["active_dims"])
    objective = GaussianObjective(target=tf.convert_to_tensor(cfg["target"]), precis=tf.convert_to_tensor(cfg["W"]))
    system = DynamicalSystem(drift=drift, policy=policy, encoder=encoder, solver=MomentMatchingEuler())

    def accumulate_loss(t, state, loss):          # upstream loops/pilco.py:199-205
      xs = GaussianMoments(moments=state, centered=True)
      xs = moment_matching(xs, encoder).y
      if record is not None:
        record[0].append(A(state[0])); record[1].append(A(state[1]))
      return loss + objective(x=xs, t=t)

    _, loss = system.solve_forward(iterator=tf.foldl, initial_time=0.0, initial_state=(tf.convert_to_tensor(m0), tf.convert_to_tensor(S0)),
                                   solution_times=np.arange(1, 1 + H, dtype=np.float64),
                                   callbacks_and_initializers=((accumulate_loss, tf.zeros([1])),))
    return A(loss)

  # ---------------------------------------------------------------- config #1 at full size
  if only in ("", "cfg1"):
    cfg = synthetic.config1_cartpole()
    d, p = cfg["dynamics"], cfg["policy"]
    H = int(cfg["horizon"])
    rec = ([cfg["m0"]], [cfg["S0"]])
    loss = closure(cfg, d, p, H, cfg["m0"], cfg["S0"], rec)
    np.savez(os.path.join(OUT, "rollout_cfg1.npz"), horizon=H, loss=loss, traj_m=np.stack(rec[0]), traj_S=np.stack(rec[1]),
             note="models = gpflowpilco_b200.synthetic.config1_cartpole() (seeded; M=256, Mp=30)")
    print("config #1: H =", H, "loss =", loss)

  # ---------------------------------------------------------------- finite-difference gradients of the closure value (H = 5, M = 40)
  if only in ("", "grads"):
    fd_gradients(closure, synthetic, np)

  # ---------------------------------------------------------------- forward_sde without a policy
  if only in ("", "variants"):
    rng = np.random.default_rng(31)
    out = {}
    Dx, N = 4, 3
    m4, S4 = rng.standard_normal((N, Dx)), generate_covariance(rng, Dx, N, 0.3)
    out.update(m=m4, S=S4)

    def random_svgp(tag, Din, L, M):
      ell = np.exp(rng.uniform(np.log(0.7), np.log(2.5), (L, Din)))
      var = 0.5 + rng.random(L)
      Z = 1.5 * rng.standard_normal((L, M, Din))
      q_mu = 0.3 * rng.standard_normal((M, L))
      q_sqrt = np.linalg.cholesky(generate_covariance(rng, M, L, 0.4))
      c = 0.1 * rng.standard_normal(L)
      out.update({f"{tag}_Z": Z, f"{tag}_ell": ell, f"{tag}_var": var, f"{tag}_q_mu": q_mu, f"{tag}_q_sqrt": q_sqrt, f"{tag}_c": c})
      return svgp(Z, ell, var, q_mu, q_sqrt, c)

    for tag, Din, encoder in (("plain", 4, None), ("enc", 5, TrigonometricEncoder(active_dims=(1,)))):
      drift = random_svgp(tag, Din, Dx, 16)
      x = GaussianMoments(moments=(tf.convert_to_tensor(m4), tf.convert_to_tensor(S4)), centered=True)
      match, noise = forward_sde(x, drift, None, None, encoder)
      assert noise is None
      out[f"{tag}_mean"], out[f"{tag}_cov"] = A(match.y.mean()), A(match.y.covariance())
      out[f"{tag}_cross"] = A(match.cross_covariance())                    # Cov(x, f)
      system = DynamicalSystem(drift=drift, policy=None, encoder=encoder, solver=MomentMatchingEuler())
      rec_m, rec_S = [], []

      def record_state(t, state, acc):
        rec_m.append(A(state[0])); rec_S.append(A(state[1]))
        return acc
      system.solve_forward(iterator=tf.foldl, initial_time=0.0, initial_state=(tf.convert_to_tensor(m4), tf.convert_to_tensor(S4)),
                           solution_times=np.arange(1, 3, dtype=np.float64), callbacks_and_initializers=((record_state, tf.zeros([1])),))
      out[f"{tag}_euler_m"], out[f"{tag}_euler_S"] = np.stack(rec_m), np.stack(rec_S)
    np.savez(os.path.join(OUT, "forward_sde_variants.npz"), **out)
  print("round-2 golden fixtures written to", OUT)


if __name__ == "__main__":
  main(*sys.argv[1:])
