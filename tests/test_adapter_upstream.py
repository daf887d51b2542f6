"""gpflowpilco_b200/adapters/upstream.py exercised inside UPSTREAM'S OWN call graph (CPU, needs the upstream sources at /root/reference;
skipped where they are not mounted, e.g. on the GPU box).

The unmodified gpflow_pilco sources are imported on the numpy stand-ins of oracle/refshim (test infrastructure, installed HERE, never by
the package), the adapter registers at upstream's real dispatcher keys and replaces the closure factory, and upstream's forward_sde /
MomentMatchingEuler / loss callback then run with the adapter's rules.  There is no GPU in this container, so the adapter's backend —
the one object that talks to libgpp_b200.so — is swapped for an oracle-backed stand-in with the same two methods: what is tested is
everything between upstream and the C ABI (dispatch keys, object unpacking, keyword semantics, centred / pre-inverted flags, shapes,
closure signature).  The CUDA backend itself is covered by tests/test_gpu_adapter.py through the same functions.
"""
import os

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gpflow_pilco")), reason="upstream sources not mounted")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class OracleBackend:
  """Same interface as adapters.upstream.CudaBackend, answered by the CPU oracle."""

  def __init__(self):
    self.calls = {"predict": 0, "rollout": 0}

  @staticmethod
  def _model(params, kuu_jitter=None):
    from oracle import gp_models as gm
    from oracle import psi_stats as ps
    L = params["Z"].shape[0]
    ks = [ps.SEKernel(float(params["variance"][l]), torch.as_tensor(params["lengthscales"][l])) for l in range(L)]
    M = params["Z"].shape[1]
    q_sqrt = params["q_sqrt"] if params["q_sqrt"] is not None else np.zeros((L, M, M))
    c = params["mean_const"]
    return gm.SVGPModel(ks, [torch.as_tensor(params["Z"][l]) for l in range(L)], torch.as_tensor(params["q_mu"]), torch.as_tensor(q_sqrt),
                        whiten=params["whiten"], mean_const=torch.zeros(L if params["W"] is None else params["W"].shape[0], dtype=torch.float64)
                        if c is None else torch.as_tensor(c), W=None if params["W"] is None else torch.as_tensor(params["W"]))

  def predict(self, params, m, S, full_output_cov, model_uncertainty, jitter, key=None):
    from oracle import gp_models as gm
    from oracle import moments as mo
    self.calls["predict"] += 1
    x = mo.GaussianMoments(torch.as_tensor(np.asarray(m)), torch.as_tensor(np.asarray(S)), True)
    old = gm.Kuu.__defaults__
    if "kuu_jitter" in params:
      gm.Kuu.__defaults__ = (float(params["kuu_jitter"]),)
    try:
      mm = gm.mm_svgp_mo(x, self._model(params), full_output_cov=True, model_uncertainty=model_uncertainty, jitter=jitter)
    finally:
      gm.Kuu.__defaults__ = old
    return mm.y.mean(), mm.y.covariance(), mm.cross[0]

  def rollout(self, dyn_params, policy, m0, S0, horizon, active_dims, target, W, key=None):
    from oracle import gp_models as gm
    from oracle import moments as mo
    from oracle import psi_stats as ps
    from oracle import rollout as ro
    self.calls["rollout"] += 1
    dyn = self._model(dyn_params)
    pol = gm.SVGPModel([ps.SEKernel(float(policy["variance"][0]), torch.as_tensor(policy["lengthscales"][0]))], [torch.as_tensor(policy["Z"][0])],
                       torch.as_tensor(policy["q_mu"][0])[:, None], torch.zeros(1, policy["Z"].shape[1], policy["Z"].shape[1], dtype=torch.float64),
                       whiten=policy["whiten"], mean_const=torch.zeros(1, dtype=torch.float64))
    return ro.mm_rollout(torch.as_tensor(np.asarray(m0)), torch.as_tensor(np.asarray(S0)), horizon, lambda s: gm.mm_svgp(s, dyn),
                         lambda s: gm.mm_policy(s, pol, policy["scale"], policy["shift"]), mo.TrigonometricEncoder(tuple(active_dims)),
                         mo.GaussianObjective(torch.as_tensor(target), torch.as_tensor(W)))


def _load_upstream_loops():
  """Import upstream's gpflow_pilco/loops/{core,model_based_rl,pilco}.py from the mounted sources.  They pull in packages that have nothing
  to do with the rollout (gym, gpflow_sampling, TF checkpoint reader, TFP distributions): those get empty stand-ins, here in the test."""
  import importlib.util
  import sys
  import types
  if "gpflow_pilco.loops.pilco" in sys.modules:
    return

  def stub(name, **attrs):
    m = sys.modules.get(name) or types.ModuleType(name)
    for k, v in attrs.items():
      if not hasattr(m, k):
        setattr(m, k, v)
    sys.modules[name] = m
    return m

  stub("gym", Env=type("Env", (), {}))
  stub("gpflow_sampling")
  stub("gpflow_sampling.sampling")
  stub("gpflow_sampling.sampling.core", AbstractSampler=type("AbstractSampler", (), {}))
  stub("tensorflow_probability.python.distributions", Distribution=type("Distribution", (), {}))
  stub("tensorflow.python.training")
  stub("tensorflow.python.training.py_checkpoint_reader", NewCheckpointReader=lambda *a, **k: None)
  import gpflow
  stub("gpflow.likelihoods", Likelihood=type("Likelihood", (), {}))
  stub("gpflow.utilities", set_trainable=lambda *a, **k: None)
  import tensorflow as tf
  for name, val in (("Module", type("Module", (), {})), ("function", lambda f=None, **k: f if f is not None else (lambda g: g))):
    if not hasattr(tf, name):
      setattr(tf, name, val)
  if not hasattr(tf, "train"):
    tf.train = types.SimpleNamespace(CheckpointManager=type("CheckpointManager", (), {}), Checkpoint=type("Checkpoint", (), {}))
  models = sys.modules["gpflow_pilco.models"]
  for n in ("SVGP", "PathwiseSVGP", "GPR", "PathwiseGPR"):
    if not hasattr(models, n):
      setattr(models, n, type(n, (), {}))
  root = os.path.join(REF, "gpflow_pilco", "loops")
  pkg = types.ModuleType("gpflow_pilco.loops")
  pkg.__path__ = [root]
  sys.modules["gpflow_pilco.loops"] = pkg
  for name in ("core", "model_based_rl", "pilco"):
    spec = importlib.util.spec_from_file_location(f"gpflow_pilco.loops.{name}", os.path.join(root, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"gpflow_pilco.loops.{name}"] = mod
    spec.loader.exec_module(mod)
    setattr(pkg, name, mod)
  sys.modules["gpflow_pilco"].loops = pkg


@pytest.fixture(scope="module")
def upstream_env():
  from oracle import refshim
  refshim.install(REF)
  _load_upstream_loops()
  from gpflowpilco_b200.adapters import upstream as ad
  be = OracleBackend()
  ad.set_backend(be)
  ad.install()
  yield ad, be
  ad.uninstall()
  ad.set_backend(None)


def _A(x):
  return np.asarray(x.to_dense() if hasattr(x, "to_dense") else x)


def _golden_system(g):
  import gpflow
  import tensorflow as tf
  from gpflow_pilco.components import GaussianObjective, TrigonometricEncoder
  from gpflow_pilco.models import InverseLinkWrapper, KernelRegressor
  from tensorflow_probability.python import bijectors as tfb
  SE, IP = gpflow.kernels.SquaredExponential, gpflow.inducing_variables.InducingPoints
  drift = gpflow.models.SVGP(
      kernel=gpflow.kernels.SeparateIndependent([SE(g["dyn_var"][l], g["dyn_ell"][l]) for l in range(4)]),
      inducing_variable=gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(g["dyn_Z"][l]) for l in range(4)]),
      q_mu=g["dyn_q_mu"], q_sqrt=g["dyn_q_sqrt"], num_latent_gps=4, whiten=True, mean_function=gpflow.mean_functions.Constant(g["dyn_c"]))
  pol = gpflow.models.SVGP(
      kernel=gpflow.kernels.SeparateIndependent([SE(g["pol_var"][0], g["pol_ell"][0])]),
      inducing_variable=gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(g["pol_Z"][0])]),
      q_mu=g["pol_q_mu"], q_sqrt=g["pol_q_sqrt"], num_latent_gps=1, whiten=True, mean_function=gpflow.mean_functions.Constant(np.zeros(1)))
  link = tfb.Chain(bijectors=[tfb.Scale(scale=tf.cast(float(g["scale"]), np.float64)), tfb.Shift(shift=tf.cast(float(g["shift"]), np.float64)),
                              tfb.NormalCDF()])
  policy = InverseLinkWrapper(model=KernelRegressor(model=pol), invlink=link)
  encoder = TrigonometricEncoder(active_dims=tuple(int(a) for a in g["active_dims"]))
  objective = GaussianObjective(target=tf.convert_to_tensor(g["target"]), precis=tf.convert_to_tensor(g["W"]))
  return drift, policy, encoder, objective


def test_registered_at_upstreams_dispatch_keys(upstream_env):
  ad, _ = upstream_env
  import gpflow
  from gpflow_pilco.loops import pilco as up_pilco
  from gpflow_pilco.moment_matching import GaussianMoments
  from gpflow_pilco.moment_matching.core import dispatcher
  assert dispatcher.dispatch(GaussianMoments, gpflow.models.SVGP) is ad.mm_gauss_svgp          # moment_matching/models.py:114
  assert dispatcher.dispatch(GaussianMoments, gpflow.models.GPR) is ad.mm_gauss_gpr            # moment_matching/models.py:44
  assert up_pilco.MomentMatchingPILCO._policy_loss_closure is ad.mm_policy_loss_closure       # loops/pilco.py:192


def test_upstream_rollout_runs_on_the_adapters_rules(upstream_env):
  """upstream's DynamicalSystem.solve_forward (forward_sde + MomentMatchingEuler + loss callback, loops/pilco.py:199-217) with the drift
  and policy SVGPs answered by the adapter's registration; result = the vectors upstream produced with its own rules (rollout.npz)."""
  ad, be = upstream_env
  import tensorflow as tf
  from gpflow_pilco.dynamics.dynamical_system import DynamicalSystem
  from gpflow_pilco.dynamics.solvers import MomentMatchingEuler
  from gpflow_pilco.moment_matching import GaussianMoments, moment_matching
  g = np.load(os.path.join(GOLD, "rollout.npz"))
  drift, policy, encoder, objective = _golden_system(g)
  system = DynamicalSystem(drift=drift, policy=policy, encoder=encoder, solver=MomentMatchingEuler())
  traj_m, traj_S = [g["m0"]], [g["S0"]]

  def accumulate_loss(t, state, loss):
    xs = moment_matching(GaussianMoments(moments=state, centered=True), encoder).y
    traj_m.append(_A(state[0])); traj_S.append(_A(state[1]))
    return loss + objective(x=xs, t=t)

  before = be.calls["predict"]
  H = int(g["horizon"])
  _, loss = system.solve_forward(iterator=tf.foldl, initial_time=0.0, initial_state=(tf.convert_to_tensor(g["m0"]), tf.convert_to_tensor(g["S0"])),
                                 solution_times=np.arange(1, 1 + H, dtype=np.float64), callbacks_and_initializers=((accumulate_loss, tf.zeros([1])),))
  assert be.calls["predict"] - before == 2 * H           # drift + policy regressor, every step, through the adapter
  np.testing.assert_allclose(np.stack(traj_m), g["traj_m"], rtol=0, atol=1e-9 * np.abs(g["traj_m"]).max())
  np.testing.assert_allclose(np.stack(traj_S), g["traj_S"], rtol=0, atol=1e-8 * np.abs(g["traj_S"]).max())
  np.testing.assert_allclose(_A(loss), g["loss"], rtol=1e-8)


def test_rule_keywords_and_flags(upstream_env):
  """full_output_cov=False returns upstream's LinearOperatorDiag, model_uncertainty / jitter are passed on, the cross term is flagged
  pre-inverted (moment_matching/models.py:293-299); values against mm_models.npz (coregionalised, not whitened: the upstream test shape)."""
  ad, _ = upstream_env
  import gpflow
  import tensorflow as tf
  from gpflow_pilco.moment_matching import GaussianMoments, moment_matching
  g = np.load(os.path.join(GOLD, "mm_models.npz"))
  SE, IP = gpflow.kernels.SquaredExponential, gpflow.inducing_variables.InducingPoints
  kern = gpflow.kernels.LinearCoregionalization([SE(g["co_var"][l], g["co_ell"][l]) for l in range(2)], W=g["co_W"])
  ivs = gpflow.inducing_variables.SeparateIndependentInducingVariables([IP(g["co_Z"][l]) for l in range(2)])
  model = gpflow.models.SVGP(kernel=kern, q_mu=g["co_q_mu"], q_sqrt=g["co_q_sqrt"], num_latent_gps=2, inducing_variable=ivs, whiten=False,
                             mean_function=gpflow.mean_functions.Constant(g["co_c"]))
  x = GaussianMoments(moments=(tf.convert_to_tensor(g["mx"]), tf.convert_to_tensor(g["Sxx"])), centered=True)
  mm = moment_matching(x, model)
  assert mm.cross[1] is True and mm.y.centered
  np.testing.assert_allclose(_A(mm.y.mean()), g["co_mean"], rtol=1e-9)
  np.testing.assert_allclose(_A(mm.y.covariance()), g["co_cov"], rtol=0, atol=1e-9 * np.abs(g["co_cov"]).max())
  np.testing.assert_allclose(_A(mm.cross[0]), g["co_cross_pre"], rtol=0, atol=1e-9 * np.abs(g["co_cross_pre"]).max())
  np.testing.assert_allclose(_A(mm.cross_covariance()), g["co_cross"], rtol=0, atol=1e-9 * np.abs(g["co_cross"]).max())
  md = moment_matching(x, model, full_output_cov=False)
  assert hasattr(md.y.covariance(), "to_dense") or _A(md.y.covariance()).ndim == 3
  np.testing.assert_allclose(np.diagonal(_A(md.y.covariance()), axis1=-2, axis2=-1), g["co_diag_cov"], rtol=0, atol=1e-9 * np.abs(g["co_cov"]).max())
  # exact GPR through the same handle machinery (Kuu jitter = noise variance)
  gpr = gpflow.models.GPR(data=(g["gpr_X"], g["gpr_Y"]), kernel=SE(float(g["gpr_var"]), g["gpr_ell"]),
                          mean_function=gpflow.mean_functions.Constant(g["gpr_c"]), noise_variance=float(g["gpr_noise"]))
  mg = moment_matching(x, gpr)
  np.testing.assert_allclose(_A(mg.y.mean()), g["gpr_mean"], rtol=1e-8)
  np.testing.assert_allclose(_A(mg.y.covariance()), g["gpr_cov"], rtol=0, atol=1e-8 * np.abs(g["gpr_cov"]).max())


def test_closure_factory_override(upstream_env):
  """MomentMatchingPILCO._policy_loss_closure as replaced by the adapter: same signature and return type; the cart-pole structure takes
  the fused rollout (one backend call for all H steps), any other structure keeps upstream's closure."""
  ad, be = upstream_env
  import tensorflow as tf
  from gpflow_pilco.dynamics.dynamical_system import DynamicalSystem
  from gpflow_pilco.dynamics.solvers import MomentMatchingEuler
  from gpflow_pilco.loops import pilco as up_pilco
  g = np.load(os.path.join(GOLD, "rollout.npz"))
  drift, policy, encoder, objective = _golden_system(g)

  class Loop(DynamicalSystem):                     # the attributes upstream's closure factory reads (loops/pilco.py:40-66,192-220)
    _policy_loss_closure = up_pilco.MomentMatchingPILCO._policy_loss_closure

    def __init__(self, **kw):
      super().__init__(solver=MomentMatchingEuler(), **kw)
      self.objective = objective

  loop = Loop(drift=drift, policy=policy, encoder=encoder)
  init = lambda: (tf.convert_to_tensor(g["m0"]), tf.convert_to_tensor(g["S0"]))
  H = int(g["horizon"])
  before = dict(be.calls)
  closure = loop._policy_loss_closure(state_initializer=init, initial_time=0.0, solution_times=np.arange(1, 1 + H, dtype=np.float64))
  loss = closure()
  assert be.calls["rollout"] == before["rollout"] + 1 and be.calls["predict"] == before["predict"]
  np.testing.assert_allclose(_A(loss), g["loss"], rtol=1e-8)
  # no encoder -> not the fused structure -> upstream's own closure (which still reaches the adapter's rules through the dispatcher)
  assert ad._cartpole_structure(Loop(drift=drift, policy=policy, encoder=None)) is None
