"""Reference-execution shim — TEST INFRASTRUCTURE (see oracle/__init__.py).

`install(reference_root)` registers numpy-backed stand-ins for `tensorflow`, `tensorflow_probability`, `gpflow` (only the
symbols the upstream hot path touches) and a skeleton `gpflow_pilco` package whose sub-modules are then imported from the
UNMODIFIED upstream sources where they lie (default /root/reference).  The upstream package __init__ (which drags in gym,
gpflow_sampling, checkpointing) is bypassed; `gpflow_pilco.models` is reduced to its wrapper classes (models/core.py).

What runs verbatim from upstream:   gpflow_pilco/utils/kernel_expectation.py, gpflow_pilco/moment_matching/*.py,
gpflow_pilco/components.py, gpflow_pilco/dynamics/*.py, gpflow_pilco/models/core.py, gpflow_pilco/utils/bvn.py.
What is ours (third-party restated): GPflow's eKff / eKxz rules, Kuu, kernels, containers; TFP's owens_t (scipy);
TensorFlow ops (numpy).  Used only by tests/golden/make_golden.py.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import numpy as np
from scipy import special as sps

from oracle.refshim import tf_shim

NoneType = type(None)


# ---------------------------------------------------------------------------------------------------------
# multiple dispatch with multipledispatch semantics (exact arity, most specific signature wins)
# ---------------------------------------------------------------------------------------------------------
class Dispatcher:
  def __init__(self, name):
    self.name = name
    self.funcs = {}

  def register(self, *types_):
    def deco(fn):
      def expand(ts):
        if not ts:
          yield ()
          return
        head = ts[0] if isinstance(ts[0], tuple) else (ts[0],)
        for h in head:
          for rest in expand(ts[1:]):
            yield (h,) + rest
      for key in expand(types_):
        self.funcs[key] = fn
      return fn
    return deco

  def dispatch(self, *arg_types):
    best, score_best = None, None
    for key, fn in self.funcs.items():
      if len(key) != len(arg_types):
        continue
      score = 0
      for want, got in zip(key, arg_types):
        if not issubclass(got, want):
          score = None
          break
        score += got.__mro__.index(want) if want in got.__mro__ else len(got.__mro__)
      if score is not None and (score_best is None or score < score_best):
        best, score_best = fn, score
    return best

  def __call__(self, *args, **kwargs):
    fn = self.dispatch(*(type(a) for a in args))
    if fn is None:
      raise NotImplementedError(f"{self.name}: no rule for ({', '.join(type(a).__name__ for a in args)})")
    return fn(*args, **kwargs)


# ---------------------------------------------------------------------------------------------------------
# gpflow stand-ins
# ---------------------------------------------------------------------------------------------------------
T = tf_shim.T
_config = {"float": np.float64, "jitter": 1e-6}


class Kernel:
  pass


class SquaredExponential(Kernel):
  def __init__(self, variance=1.0, lengthscales=1.0, active_dims=None):
    self.variance = T(np.asarray(variance, dtype=np.float64))
    self.lengthscales = T(np.asarray(lengthscales, dtype=np.float64))
    self.active_dims = None if active_dims is None else tuple(active_dims)

  @property
  def ard(self):
    return np.ndim(self.lengthscales) > 0

  def slice(self, X, X2=None):
    if self.active_dims is None:
      return X, X2
    idx = list(self.active_dims)
    return X[..., idx], (None if X2 is None else X2[..., idx])

  def slice_cov(self, cov):
    if self.active_dims is None:
      return cov
    idx = list(self.active_dims)
    return cov[..., idx, :][..., :, idx]

  def on_separate_dims(self, other):
    if self.active_dims is None or other.active_dims is None:
      return False
    return not (set(self.active_dims) & set(other.active_dims))

  def __call__(self, X, X2=None, full_cov=True):
    X, X2 = self.slice(np.asarray(X), None if X2 is None else np.asarray(X2))
    X2 = X if X2 is None else X2
    d = (X / self.lengthscales)[..., :, None, :] - (X2 / self.lengthscales)[..., None, :, :]
    return T(self.variance * np.exp(-0.5 * (d * d).sum(-1)))


class MultioutputKernel(Kernel):
  @property
  def num_latent_gps(self):
    return len(self.kernels)


class SeparateIndependent(MultioutputKernel):
  def __init__(self, kernels):
    self.kernels = list(kernels)


class SharedIndependent(MultioutputKernel):
  def __init__(self, kernel, output_dim):
    self.kernel = kernel
    self.kernels = [kernel] * output_dim


class LinearCoregionalization(MultioutputKernel):
  def __init__(self, kernels, W):
    self.kernels = list(kernels)
    self.W = T(np.asarray(W, dtype=np.float64))


class InducingVariables:
  pass


class InducingPoints(InducingVariables):
  def __init__(self, Z):
    self.Z = T(np.asarray(Z, dtype=np.float64))

  @property
  def variables(self):          # tf.Module.variables of gpflow's InducingPoints: (Z,)
    return (self.Z,)


class SeparateIndependentInducingVariables(InducingVariables):
  def __init__(self, inducing_variable_list):
    self.inducing_variables = list(inducing_variable_list)

  @property
  def variables(self):
    return tuple(iv.Z for iv in self.inducing_variables)


class SharedIndependentInducingVariables(InducingVariables):
  def __init__(self, inducing_variable):
    self.inducing_variable = inducing_variable
    self.inducing_variables = [inducing_variable]


class Gaussian:
  def __init__(self, mu, cov):
    self.mu, self.cov = T(mu), T(cov)


class DiagonalGaussian:
  def __init__(self, mu, cov):
    self.mu, self.cov = T(mu), T(cov)


class Zero:
  def __call__(self, X):
    return T(np.zeros(np.shape(X)[:-1] + (1,)))


class Constant:
  def __init__(self, c=None):
    self.c = T(np.atleast_1d(np.asarray(c, dtype=np.float64)))

  def __call__(self, X):
    return T(np.tile(np.reshape(self.c, (1, -1)), (np.shape(X)[0], 1)))


class GaussianLikelihood:
  def __init__(self, variance=1.0):
    self.variance = T(np.asarray(variance, dtype=np.float64))


class GPModel:
  pass


class GPR(GPModel):
  def __init__(self, data, kernel, mean_function=None, noise_variance=1.0):
    self.data = (T(data[0]), T(data[1]))
    self.kernel = kernel
    self.mean_function = Zero() if mean_function is None else mean_function
    self.likelihood = GaussianLikelihood(noise_variance)


class SVGP(GPModel):
  def __init__(self, kernel, likelihood=None, inducing_variable=None, mean_function=None, num_latent_gps=1, q_mu=None, q_sqrt=None,
               whiten=True, **kw):
    self.kernel, self.likelihood, self.inducing_variable = kernel, likelihood, inducing_variable
    self.mean_function = Zero() if mean_function is None else mean_function
    self.q_mu, self.q_sqrt, self.whiten = T(q_mu), T(q_sqrt), whiten
    self.num_latent_gps = num_latent_gps


# GPflow's expectation dispatcher + the two closed forms the upstream code relies on -------------
expectation_dispatcher = Dispatcher("expectation")


@expectation_dispatcher.register(Gaussian, SquaredExponential, NoneType, NoneType, NoneType)
def _eKff(p, kernel, _, __, ___, nghp=None):
  return T(np.full(np.shape(p.mu)[:-1], float(kernel.variance)))


@expectation_dispatcher.register(DiagonalGaussian, SquaredExponential, NoneType, NoneType, NoneType)
def _eKff_diag(p, kernel, _, __, ___, nghp=None):
  return T(np.full(np.shape(p.mu)[:-1], float(kernel.variance)))


@expectation_dispatcher.register(Gaussian, SquaredExponential, InducingPoints, NoneType, NoneType)
def _eKxz(p, kernel, inducing_variable, _, __, nghp=None):
  Xmu, _ = kernel.slice(np.asarray(p.mu), None)
  Xcov = np.asarray(kernel.slice_cov(p.cov))
  Z, _ = kernel.slice(np.asarray(inducing_variable.Z), None)
  D = Xmu.shape[1]
  ls = np.asarray(kernel.lengthscales) if kernel.ard else np.full(D, float(kernel.lengthscales))
  chol = np.linalg.cholesky(np.diag(ls ** 2) + Xcov)                    # [N,D,D]
  diffs = np.swapaxes(Z[None] - Xmu[:, None, :], -1, -2)                 # [N,D,M]
  half = np.asarray(tf_shim.linalg.triangular_solve(chol, diffs, lower=True))
  maha = (half ** 2).sum(1)
  dets = np.prod(ls) / np.exp(np.log(np.diagonal(chol, axis1=-2, axis2=-1)).sum(-1))
  return T(float(kernel.variance) * dets[:, None] * np.exp(-0.5 * maha))


@expectation_dispatcher.register(DiagonalGaussian, SquaredExponential, InducingPoints, NoneType, NoneType)
def _eKxz_diag(p, kernel, inducing_variable, _, __, nghp=None):
  return _eKxz(Gaussian(p.mu, tf_shim.linalg.diag(p.cov)), kernel, inducing_variable, None, None)


def expectation(p, obj1, obj2=None, nghp=None):
  def unpack(o):
    return o if isinstance(o, tuple) else (o, None)
  k1, f1 = unpack(obj1)
  k2, f2 = unpack(obj2)
  return expectation_dispatcher(p, k1, f1, k2, f2, nghp=nghp)


def square_distance(X, X2):
  X = np.asarray(X)
  X2 = X if X2 is None else np.asarray(X2)
  d = X[..., :, None, :] - X2[..., None, :, :]
  return T((d * d).sum(-1))


def Kuu(inducing_variable, kernel, jitter=0.0):
  if isinstance(kernel, MultioutputKernel):
    ivs = inducing_variable.inducing_variables
    if len(ivs) == 1:
      ivs = ivs * len(kernel.kernels)
    return tf_shim.stack([Kuu(iv, k, jitter=jitter) for iv, k in zip(ivs, kernel.kernels)], axis=0)
  K = np.asarray(kernel(inducing_variable.Z))
  return T(K + jitter * np.eye(K.shape[-1]))


# tfp stand-ins ----------------------------------------------------------------------------------------------
class _Bijector:
  pass


class TfbChain(_Bijector):
  def __init__(self, bijectors=None, **kw):
    self.bijectors = list(bijectors)

  def __call__(self, x):
    for b in reversed(self.bijectors):
      x = b(x)
    return x


class TfbShift(_Bijector):
  def __init__(self, shift=None, **kw):
    self.shift = shift

  def __call__(self, x):
    return T(np.asarray(x) + np.asarray(self.shift))


class TfbScale(_Bijector):
  def __init__(self, scale=None, **kw):
    self.scale = scale

  def __call__(self, x):
    return T(np.asarray(x) * np.asarray(self.scale))


class TfbNormalCDF(_Bijector):
  def __call__(self, x):
    return T(sps.ndtr(np.asarray(x)))


def _module(name, **attrs):
  m = types.ModuleType(name)
  for k, v in attrs.items():
    setattr(m, k, v)
  sys.modules[name] = m
  return m


_installed = {}


def install(reference_root: str = "/root/reference"):
  """Register the stand-in modules and return the upstream `gpflow_pilco` package skeleton."""
  if _installed:
    return _installed["pkg"]
  if not os.path.isdir(os.path.join(reference_root, "gpflow_pilco")):
    raise FileNotFoundError(f"{reference_root}/gpflow_pilco not found: the golden vectors can only be regenerated where the upstream "
                            "sources are mounted")
  tf = tf_shim.build_module()
  sys.modules["tensorflow"] = tf
  _module("tensorflow.linalg", **vars(tf_shim.linalg))
  _module("tensorflow.python")
  _module("tensorflow.python.module")
  _module("tensorflow.python.module.module", camel_to_snake=lambda s: s)

  bij = _module("tensorflow_probability.python.bijectors", Chain=TfbChain, Shift=TfbShift, Scale=TfbScale, NormalCDF=TfbNormalCDF,
                Bijector=_Bijector)
  _module("tensorflow_probability")
  _module("tensorflow_probability.python", bijectors=bij)
  _module("tensorflow_probability.python.math")
  _module("tensorflow_probability.python.math.special", owens_t=lambda h, a: T(sps.owens_t(np.asarray(h), np.asarray(a))))

  kernels = _module("gpflow.kernels", Kernel=Kernel, SquaredExponential=SquaredExponential, MultioutputKernel=MultioutputKernel,
                    SeparateIndependent=SeparateIndependent, SharedIndependent=SharedIndependent,
                    LinearCoregionalization=LinearCoregionalization)
  utilities = _module("gpflow.utilities", Dispatcher=Dispatcher)
  _module("gpflow.utilities.ops", square_distance=square_distance)
  ind = _module("gpflow.inducing_variables", InducingPoints=InducingPoints, InducingVariables=InducingVariables,
                SeparateIndependentInducingVariables=SeparateIndependentInducingVariables,
                SharedIndependentInducingVariables=SharedIndependentInducingVariables)
  pd = _module("gpflow.probability_distributions", Gaussian=Gaussian, DiagonalGaussian=DiagonalGaussian)
  exp_mod = _module("gpflow.expectations", expectation=expectation)
  _module("gpflow.expectations.dispatch", expectation=expectation_dispatcher)
  mf = _module("gpflow.mean_functions", Zero=Zero, Constant=Constant)
  models = _module("gpflow.models", GPR=GPR, SVGP=SVGP, GPModel=GPModel)
  cov = _module("gpflow.covariances", Kuu=Kuu)
  cfg = _module("gpflow.config", default_float=lambda: _config["float"], default_jitter=lambda: _config["jitter"],
                set_default_float=lambda v: None, set_default_jitter=lambda v: _config.__setitem__("jitter", float(v)))
  lik = _module("gpflow.likelihoods", Gaussian=GaussianLikelihood)
  _module("gpflow", kernels=kernels, utilities=utilities, inducing_variables=ind, probability_distributions=pd, expectations=exp_mod,
          mean_functions=mf, models=models, covariances=cov, config=cfg, likelihoods=lik)

  # upstream package skeleton: sub-modules load from the unmodified sources, the heavy __init__ files are bypassed
  root = os.path.join(reference_root, "gpflow_pilco")

  def skeleton(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m

  def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod

  pkg = skeleton("gpflow_pilco", root)
  pkg.utils = skeleton("gpflow_pilco.utils", os.path.join(root, "utils"))
  models_pkg = skeleton("gpflow_pilco.models", os.path.join(root, "models"))
  core = load("gpflow_pilco.models.core", os.path.join(root, "models", "core.py"))
  for n in ("GPModelWrapper", "KernelRegressor", "InverseLinkWrapper"):
    setattr(models_pkg, n, getattr(core, n))
  pkg.models = models_pkg
  pkg.utils.kernel_expectation = importlib.import_module("gpflow_pilco.utils.kernel_expectation")
  pkg.utils.bvn = importlib.import_module("gpflow_pilco.utils.bvn")
  pkg.moment_matching = importlib.import_module("gpflow_pilco.moment_matching")      # upstream's own __init__ (core, gaussian, rules)
  pkg.components = importlib.import_module("gpflow_pilco.components")
  dyn = skeleton("gpflow_pilco.dynamics", os.path.join(root, "dynamics"))
  dyn.forward_sde = importlib.import_module("gpflow_pilco.dynamics.forward_sde")
  dyn.solvers = importlib.import_module("gpflow_pilco.dynamics.solvers")
  dyn.dynamical_system = importlib.import_module("gpflow_pilco.dynamics.dynamical_system")
  pkg.dynamics = dyn
  _installed["pkg"] = pkg
  return pkg
