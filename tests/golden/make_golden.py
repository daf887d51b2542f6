"""Generate tests/golden/*.npz by executing the UNMODIFIED upstream sources (default /root/reference) on the numpy
stand-ins of oracle/refshim.  Run from the repo root:   python tests/golden/make_golden.py [reference_root]

Each fixture stores seeded inputs and the outputs upstream's own functions produced for them:
  psi.npz        kernel_expectation (GPflow-restated eKxz; upstream `_E`: same-kernel, same-feature and generic branches)
  mm_models.npz  moment_matching(x, GPR | SVGP single-output | SVGP SeparateIndependent | SVGP LinearCoregionalization),
                 full and diagonal output covariance   (upstream moment_matching/models.py:44-299)
  rules.npz      sincos / sin / cos rules, TrigonometricEncoder rule, Chain[Scale,Shift,NormalCDF] rule, GaussianObjective
  squash_nd.npz  Chain[Scale,Shift,NormalCDF] on 2-D / 3-D Gaussians: the Genz bivariate-normal branch (bijectors.py:59-63, utils/bvn.py)
  rollout.npz    forward_sde (encoder + squashed RBF policy + SVGP drift) stepped by MomentMatchingEuler with the loss
                 callback of loops/pilco.py:199-205 (5 steps), trajectory of moments and the accumulated loss
The pathwise sampler is a third-party package absent from the tree: no golden vectors for it (parity unpinned).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))


def cov(rng, d, n, scale):
  from gpflowpilco_b200.synthetic import generate_covariance
  return generate_covariance(rng, d, n, scale)


def main(reference_root="/root/reference"):
  from oracle import refshim
  refshim.install(reference_root)
  import gpflow
  import tensorflow as tf
  from gpflow_pilco.components import GaussianObjective, TrigonometricEncoder
  from gpflow_pilco.dynamics.dynamical_system import DynamicalSystem
  from gpflow_pilco.dynamics.solvers import MomentMatchingEuler
  from gpflow_pilco.models import InverseLinkWrapper, KernelRegressor
  from gpflow_pilco.moment_matching import GaussianMoments, moment_matching
  from gpflow_pilco.moment_matching.maths import sincos
  from gpflow_pilco.utils.kernel_expectation import kernel_expectation
  from tensorflow_probability.python import bijectors as tfb

  def A(x):     # dense numpy view of tensors and LinearOperators alike
    return np.asarray(x.to_dense() if hasattr(x, "to_dense") else x)

  SE, IP = gpflow.kernels.SquaredExponential, gpflow.inducing_variables.InducingPoints
  G = gpflow.probability_distributions.Gaussian

  # ---------------------------------------------------------------- psi
  rng = np.random.default_rng(2024)
  out = {}
  for tag, D, M1, M2, N in (("d2", 2, 32, 32, 1), ("d6", 6, 20, 17, 3)):
    mu = rng.standard_normal((N, D))
    S = cov(rng, D, N, 0.1 if D == 2 else 0.3)
    k1 = SE(0.89 ** 2, np.exp(rng.uniform(np.log(0.3), np.log(3.0), D)))
    k2 = SE(1.3, np.exp(rng.uniform(np.log(0.3), np.log(3.0), D)))
    Z1, Z2 = IP(rng.standard_normal((M1, D))), IP(rng.standard_normal((M2, D)))
    p = G(mu, S)
    out.update({f"{tag}_mu": mu, f"{tag}_cov": S, f"{tag}_ell1": A(k1.lengthscales), f"{tag}_ell2": A(k2.lengthscales),
                f"{tag}_var1": A(k1.variance), f"{tag}_var2": A(k2.variance), f"{tag}_Z1": A(Z1.Z), f"{tag}_Z2": A(Z2.Z),
                f"{tag}_eKxz": A(kernel_expectation(p, (k1, Z1))),
                f"{tag}_same": A(kernel_expectation(p, (k1, Z1), (k1, Z1))),
                f"{tag}_samekern": A(kernel_expectation(p, (k1, Z1), (k1, Z2))),
                f"{tag}_generic": A(kernel_expectation(p, (k1, Z1), (k2, Z2)))})
  np.savez(os.path.join(OUT, "psi.npz"), **out)

  # ---------------------------------------------------------------- GP models
  rng = np.random.default_rng(7)
  out = {}

  def record(tag, x, model, **kw):
    mf = moment_matching(x, model, **kw)
    md = moment_matching(x, model, full_output_cov=False, **kw)
    out[f"{tag}_mean"] = A(mf.y.mean())
    out[f"{tag}_cov"] = A(mf.y.covariance())
    out[f"{tag}_cross_pre"] = A(mf.cross[0])
    out[f"{tag}_cross"] = A(mf.cross_covariance())
    out[f"{tag}_diag_cov"] = A(tf.linalg.diag_part(md.y.covariance()))

  D, M, N = 4, 16, 2                      # upstream tests/test_moment_matching.py sizes
  mx, Sxx = rng.random((N, D)), cov(rng, D, N, 0.05)
  x = GaussianMoments(moments=(tf.convert_to_tensor(mx), tf.convert_to_tensor(Sxx)), centered=True)
  out.update(mx=mx, Sxx=Sxx)
  # GPR (:44-111)
  ell = np.exp(rng.uniform(np.log(0.3), np.log(3.0), D))
  X, Y, c = rng.random((M, D)), 0.89 * rng.standard_normal((M, 1)), 1 + rng.standard_normal(1)
  gpr = gpflow.models.GPR(data=(X, Y), kernel=SE(0.89 ** 2, ell), mean_function=gpflow.mean_functions.Constant(c), noise_variance=1e-3)
  out.update(gpr_X=X, gpr_Y=Y, gpr_c=c, gpr_ell=ell, gpr_var=0.89 ** 2, gpr_noise=1e-3)
  record("gpr", x, gpr)
  # SVGP single output (:129-197), not whitened like the upstream test
  Z = rng.random((M, D))
  q_mu = 0.89 * rng.standard_normal((M, 1))
  q_sqrt = np.linalg.cholesky(cov(rng, M, 1, 0.5))
  so = gpflow.models.SVGP(kernel=SE(0.7, ell), q_mu=q_mu, q_sqrt=q_sqrt, inducing_variable=IP(Z), whiten=False,
                          mean_function=gpflow.mean_functions.Constant(c))
  out.update(so_Z=Z, so_q_mu=q_mu, so_q_sqrt=q_sqrt, so_ell=ell, so_var=0.7, so_c=c)
  record("so", x, so)
  record("so_nounc", x, so, model_uncertainty=False)
  # SVGP multi output, LinearCoregionalization 2 latents -> 3 outputs (:200-299), upstream test shape
  Lf, P = 2, 3
  ells = np.exp(rng.uniform(np.log(0.3), np.log(3.0), (Lf, D)))
  Zs = rng.random((Lf, M, D))
  W = rng.random((P, Lf)); W /= np.linalg.norm(W, axis=-1, keepdims=True)
  q_mu2 = 0.89 * rng.standard_normal((M, Lf))
  q_sqrt2 = np.linalg.cholesky(cov(rng, M, Lf, 0.5))
  cm = 1 + rng.standard_normal(P)
  kern = gpflow.kernels.LinearCoregionalization([SE(0.89 ** 2, ells[l]) for l in range(Lf)], W=W)
  ivs = gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(Zs[l]) for l in range(Lf)])
  mo_c = gpflow.models.SVGP(kernel=kern, q_mu=q_mu2, q_sqrt=q_sqrt2, num_latent_gps=Lf, inducing_variable=ivs, whiten=False,
                            mean_function=gpflow.mean_functions.Constant(cm))
  out.update(co_Z=Zs, co_ell=ells, co_var=np.full(Lf, 0.89 ** 2), co_W=W, co_q_mu=q_mu2, co_q_sqrt=q_sqrt2, co_c=cm)
  record("co", x, mo_c)
  # SVGP multi output, SeparateIndependent, whitened, cart-pole dynamics shape (D=6, L=4)
  D6, L4, M6, N6 = 6, 4, 24, 3
  mx6, Sxx6 = rng.standard_normal((N6, D6)), cov(rng, D6, N6, 0.2)
  x6 = GaussianMoments(moments=(tf.convert_to_tensor(mx6), tf.convert_to_tensor(Sxx6)), centered=True)
  ell6 = np.exp(rng.uniform(np.log(0.7), np.log(2.5), (L4, D6)))
  var6 = 0.5 + rng.random(L4)
  Z6 = 1.5 * rng.standard_normal((L4, M6, D6))
  q_mu6 = rng.standard_normal((M6, L4))
  q_sqrt6 = np.linalg.cholesky(cov(rng, M6, L4, 0.4))
  c6 = rng.standard_normal(L4)
  kern6 = gpflow.kernels.SeparateIndependent([SE(var6[l], ell6[l]) for l in range(L4)])
  ivs6 = gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(Z6[l]) for l in range(L4)])
  si = gpflow.models.SVGP(kernel=kern6, q_mu=q_mu6, q_sqrt=q_sqrt6, num_latent_gps=L4, inducing_variable=ivs6, whiten=True,
                          mean_function=gpflow.mean_functions.Constant(c6))
  out.update(si_mx=mx6, si_Sxx=Sxx6, si_Z=Z6, si_ell=ell6, si_var=var6, si_q_mu=q_mu6, si_q_sqrt=q_sqrt6, si_c=c6)
  record("si", x6, si, jitter=1e-8)
  record("si_nounc", x6, si, model_uncertainty=False)
  np.savez(os.path.join(OUT, "mm_models.npz"), **out)

  # ---------------------------------------------------------------- small rules
  rng = np.random.default_rng(11)
  out = {}
  m4, S4 = rng.standard_normal((3, 4)), cov(rng, 4, 3, 0.3)
  x4 = GaussianMoments(moments=(tf.convert_to_tensor(m4), tf.convert_to_tensor(S4)), centered=True)
  out.update(m=m4, S=S4)
  for name, fn in (("sincos", sincos), ("sin", tf.math.sin), ("cos", tf.math.cos)):
    mm = moment_matching(x4, fn)
    out[f"{name}_mean"], out[f"{name}_cov"], out[f"{name}_cross"] = A(mm.y.mean()), A(mm.y.covariance()), A(mm.cross_covariance())
  for tag, active in (("enc1", (1,)), ("enc23", (2, 3))):
    mm = moment_matching(x4, TrigonometricEncoder(active_dims=active))
    out[f"{tag}_mean"], out[f"{tag}_cov"], out[f"{tag}_cross"] = A(mm.y.mean()), A(mm.y.covariance()), A(mm.cross_covariance())
  link = tfb.Chain(bijectors=[tfb.Scale(scale=tf.cast(20 - 1e-5, np.float64)), tfb.Shift(shift=tf.cast(-0.5, np.float64)), tfb.NormalCDF()])
  sq_m, sq_v = rng.standard_normal(5), rng.uniform(0.01, 2.0, 5)
  sq = [moment_matching(GaussianMoments(moments=(tf.convert_to_tensor(sq_m[i:i + 1][None]), tf.convert_to_tensor(sq_v[i:i + 1][None, None])),
                                        centered=True), link) for i in range(5)]       # upstream's 1-D rule is N = 1 only
  out.update(sq_m=sq_m, sq_v=sq_v, sq_mean=np.array([float(A(s.y.mean()).ravel()[0]) for s in sq]),
             sq_var=np.array([float(A(s.y.covariance()).ravel()[0]) for s in sq]),
             sq_cross_pre=np.array([float(A(s.cross_covariance(preinv=True)).ravel()[0]) for s in sq]))
  Wc = np.linalg.inv(cov(rng, 4, 1, 0.7)[0])
  tgt = rng.standard_normal(4)
  obj = GaussianObjective(target=tf.convert_to_tensor(tgt), precis=tf.convert_to_tensor(Wc))
  Xs = rng.standard_normal((20, 4))
  out.update(obj_W=Wc, obj_target=tgt, obj_expected=A(obj(x4)), obj_X=Xs, obj_samples=A(obj(tf.convert_to_tensor(Xs))))
  np.savez(os.path.join(OUT, "rules.npz"), **out)

  # ---------------------------------------------------------------- multi-dimensional squashing link (Genz BVN branch)
  # upstream moment_matching/bijectors.py:59-63 with utils/bvn.py:67-232; one state per call so that upstream's batch-wide choice of
  # the Gauss-Legendre order (bvn.py:221-228) is the per-state choice; cases cover all three orders and the |rho| >= 0.925 branch
  rng = np.random.default_rng(23)
  out = {}
  for tag, Adim, corr in (("a2_weak", 2, 0.1), ("a2_mid", 2, 0.6), ("a2_strong", 2, 0.97), ("a3_mid", 3, 0.5), ("a3_strong", 3, 0.96)):
    n = 3
    ms, Ss, means, covs, pres = [], [], [], [], []
    for i in range(n):
      m = rng.standard_normal(Adim)
      sd = np.exp(rng.uniform(np.log(0.5), np.log(3.0), Adim))     # large variances make rho = S_ij / sqrt((1+v_i)(1+v_j)) approach corr
      C = np.full((Adim, Adim), corr) + (1 - corr) * np.eye(Adim)
      if i == 1 and Adim == 2:
        C[0, 1] = C[1, 0] = -corr
      S = C * sd[:, None] * sd[None, :] * (25.0 if "strong" in tag else 1.0)
      x = GaussianMoments(moments=(tf.convert_to_tensor(m[None]), tf.convert_to_tensor(S[None])), centered=True)
      with np.errstate(all="ignore"):
        mm = moment_matching(x, tfb.Chain(bijectors=[tfb.Scale(scale=tf.cast(3.0, np.float64)), tfb.Shift(shift=tf.cast(-0.5, np.float64)),
                                                     tfb.NormalCDF()]))
      ms.append(m); Ss.append(S); means.append(A(mm.y.mean())[0]); covs.append(A(mm.y.covariance())[0])
      pres.append(A(mm.cross_covariance(preinv=True))[0])
    out.update({f"{tag}_m": np.stack(ms), f"{tag}_S": np.stack(Ss), f"{tag}_mean": np.stack(means), f"{tag}_cov": np.stack(covs),
                f"{tag}_cross_pre": np.stack(pres)})
  out.update(scale=3.0, shift=-0.5)
  np.savez(os.path.join(OUT, "squash_nd.npz"), **out)

  # ---------------------------------------------------------------- rollout (forward_sde + MomentMatchingEuler + loss callback)
  from gpflowpilco_b200 import synthetic
  cfg = synthetic.config1_cartpole(M=40, Mp=12)
  d, p = cfg["dynamics"], cfg["policy"]
  drift = gpflow.models.SVGP(
      kernel=gpflow.kernels.SeparateIndependent([SE(d["variance"][l], d["lengthscales"][l]) for l in range(4)]),
      inducing_variable=gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(d["Z"][l]) for l in range(4)]),
      q_mu=d["q_mu"], q_sqrt=d["q_sqrt"], num_latent_gps=4, whiten=True, mean_function=gpflow.mean_functions.Constant(d["mean_const"]))
  pol_svgp = gpflow.models.SVGP(
      kernel=gpflow.kernels.SeparateIndependent([SE(p["variance"][0], p["lengthscales"][0])]),
      inducing_variable=gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(p["Z"][0])]),
      q_mu=p["q_mu"], q_sqrt=p["q_sqrt"], num_latent_gps=1, whiten=True, mean_function=gpflow.mean_functions.Constant(p["mean_const"]))
  policy = InverseLinkWrapper(model=KernelRegressor(model=pol_svgp), invlink=link)
  encoder = TrigonometricEncoder(active_dims=cfg["active_dims"])
  objective = GaussianObjective(target=tf.convert_to_tensor(cfg["target"]), precis=tf.convert_to_tensor(cfg["W"]))
  system = DynamicalSystem(drift=drift, policy=policy, encoder=encoder, solver=MomentMatchingEuler())
  H = 5
  traj_m, traj_S = [cfg["m0"]], [cfg["S0"]]

  def accumulate_loss(t, state, loss):          # upstream loops/pilco.py:199-205
    xs = GaussianMoments(moments=state, centered=True)
    xs = moment_matching(xs, encoder).y
    traj_m.append(A(state[0])); traj_S.append(A(state[1]))
    return loss + objective(x=xs, t=t)

  _, loss = system.solve_forward(iterator=tf.foldl, initial_time=0.0, initial_state=(tf.convert_to_tensor(cfg["m0"]), tf.convert_to_tensor(cfg["S0"])),
                                 solution_times=np.arange(1, 1 + H, dtype=np.float64),
                                 callbacks_and_initializers=((accumulate_loss, tf.zeros([1])),))
  np.savez(os.path.join(OUT, "rollout.npz"), horizon=H, loss=A(loss), traj_m=np.stack(traj_m), traj_S=np.stack(traj_S),
           dyn_Z=d["Z"], dyn_ell=d["lengthscales"], dyn_var=d["variance"], dyn_q_mu=d["q_mu"], dyn_q_sqrt=d["q_sqrt"], dyn_c=d["mean_const"],
           pol_Z=p["Z"], pol_ell=p["lengthscales"], pol_var=p["variance"], pol_q_mu=p["q_mu"], pol_q_sqrt=p["q_sqrt"],
           m0=cfg["m0"], S0=cfg["S0"], target=cfg["target"], W=cfg["W"], scale=cfg["squash_scale"], shift=cfg["squash_shift"],
           active_dims=np.array(cfg["active_dims"]))
  print("golden fixtures written to", OUT)


if __name__ == "__main__":
  main(*sys.argv[1:])
