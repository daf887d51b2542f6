"""Light-weight parameter containers standing in for the GPflow objects the reference's hot path reads.

Only what moment matching / the rollouts consume is modelled (predict paths only):
  gpflow.kernels.SquaredExponential / SeparateIndependent / LinearCoregionalization / SharedIndependent,
  gpflow.inducing_variables.InducingPoints / SeparateIndependentInducingVariables / SharedIndependent...,
  gpflow.models.SVGP / GPR, gpflow.mean_functions.Zero / Constant (upstream models/mean_functions.py:19-38),
  KernelRegressor / InverseLinkWrapper (upstream models/core.py:30-71),
  tfp bijectors Chain / Scale / Shift / NormalCDF as used by upstream examples/cartpole_swingup/swingup_loops.py:85-91.
Attribute names follow GPflow so upstream-style code reads the same.  Tensors are torch float64 (CUDA for compute).
Model fitting (ELBO, initialisers, priors) is out of scope.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

F64 = torch.float64


def _t(x, device=None) -> torch.Tensor:
  t = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=F64)
  t = t.to(F64)
  return t if device is None else t.to(device)


# ---- kernels ------------------------------------------------------------------------------------------
class SquaredExponential:
  def __init__(self, variance=1.0, lengthscales=1.0, active_dims: Optional[Sequence[int]] = None):
    self.variance = _t(variance)
    self.lengthscales = _t(lengthscales)
    self.active_dims = None if active_dims is None else tuple(active_dims)

  @property
  def ard(self) -> bool:
    return self.lengthscales.ndim > 0

  def ell(self, ndims: int) -> torch.Tensor:
    return self.lengthscales if self.ard else self.lengthscales.expand(ndims)


class MultioutputKernel:
  kernels: List[SquaredExponential]

  @property
  def num_latent_gps(self) -> int:
    return len(self.kernels)


class SeparateIndependent(MultioutputKernel):
  def __init__(self, kernels: Sequence[SquaredExponential]):
    self.kernels = list(kernels)


class SharedIndependent(MultioutputKernel):
  def __init__(self, kernel: SquaredExponential, output_dim: int):
    self.kernel = kernel
    self.kernels = [kernel] * output_dim


class LinearCoregionalization(MultioutputKernel):
  def __init__(self, kernels: Sequence[SquaredExponential], W):
    self.kernels = list(kernels)
    self.W = _t(W)


# ---- inducing variables -------------------------------------------------------------------------------
class InducingPoints:
  def __init__(self, Z):
    self.Z = _t(Z)


class SeparateIndependentInducingVariables:
  def __init__(self, inducing_variable_list: Sequence[InducingPoints]):
    self.inducing_variables = list(inducing_variable_list)

  @property
  def variables(self):
    return [iv.Z for iv in self.inducing_variables]


class SharedIndependentInducingVariables:
  def __init__(self, inducing_variable: InducingPoints):
    self.inducing_variable = inducing_variable
    self.inducing_variables = [inducing_variable]


# ---- mean functions -----------------------------------------------------------------------------------
class Zero:
  def __call__(self, X):
    return torch.zeros_like(X[..., :1])


class Constant:
  def __init__(self, c):
    self.c = _t(c).reshape(-1)

  def __call__(self, X):
    return self.c.to(X.device).expand(*X.shape[:-1], self.c.shape[0])


# ---- likelihood stub ----------------------------------------------------------------------------------
class Gaussian:
  def __init__(self, variance=1.0):
    self.variance = _t(variance)


# ---- models -------------------------------------------------------------------------------------------
class _HandleCache:
  """Caches the C-ABI model handle until a parameter tensor is replaced or modified in place."""

  def _param_tensors(self) -> Tuple[torch.Tensor, ...]:
    raise NotImplementedError

  def _settings(self) -> tuple:
    """non-tensor attributes that are baked into a handle (jitter, whitening, kind of mean function, ...)"""
    return ()

  def _signature(self, extra):
    return tuple((id(t), t._version, t.data_ptr()) for t in self._param_tensors()) + tuple(self._settings()) + tuple(extra)

  def cached_handle(self, extra, factory):
    sig = self._signature(extra)
    cache = self.__dict__.setdefault("_handles", {})
    hit = cache.get(extra)
    if hit is not None and hit[0] == sig:
      return hit[1]
    handle = factory()
    cache[extra] = (sig, handle)
    return handle

  def invalidate(self):
    self.__dict__.pop("_handles", None)


class SVGP(_HandleCache):
  """gpflow.models.SVGP parameters (upstream models/svgp.py:33-121 builds these; whiten defaults True like GPflow)."""

  def __init__(self, kernel, inducing_variable, q_mu, q_sqrt=None, whiten: bool = True, mean_function=None,
               likelihood=None, num_latent_gps: Optional[int] = None):
    self.kernel = kernel
    self.inducing_variable = inducing_variable
    self.q_mu = _t(q_mu)
    M, L = self.q_mu.shape
    self.q_sqrt = torch.eye(M, dtype=F64, device=self.q_mu.device).expand(L, M, M).contiguous() if q_sqrt is None else _t(q_sqrt)
    self.whiten = bool(whiten)
    self.mean_function = Zero() if mean_function is None else mean_function
    self.likelihood = likelihood
    self.num_latent_gps = L if num_latent_gps is None else num_latent_gps
    self.kuu_jitter = None      # None -> gpflow default_jitter (1e-6); set per latent to model exact-GP noise on the diagonal

  # unpacking as upstream utils/kernel_expectation.py:41-69
  def latent_kernels(self) -> List[SquaredExponential]:
    return list(self.kernel.kernels) if isinstance(self.kernel, MultioutputKernel) else [self.kernel]

  def latent_inducing(self) -> List[torch.Tensor]:
    iv = self.inducing_variable
    L = len(self.latent_kernels())
    if isinstance(iv, SeparateIndependentInducingVariables):
      zs = [v.Z for v in iv.inducing_variables]
      assert len(zs) == L
      return zs
    if isinstance(iv, SharedIndependentInducingVariables):
      return [iv.inducing_variable.Z] * L
    return [iv.Z] * L

  def _settings(self):
    j = self.kuu_jitter
    j = None if j is None else (tuple(float(v) for v in j) if isinstance(j, (list, tuple)) else float(j))
    lik = getattr(self.likelihood, "variance", None)
    return (j, bool(self.whiten), type(self.mean_function).__name__, type(self.kernel).__name__, type(self.inducing_variable).__name__,
            None if lik is None else float(lik))

  def _param_tensors(self):
    ts = [self.q_mu, self.q_sqrt]
    for k in self.latent_kernels():
      ts += [k.variance, k.lengthscales]
    ts += self.latent_inducing()
    if isinstance(self.kernel, LinearCoregionalization):
      ts.append(self.kernel.W)
    if isinstance(self.mean_function, Constant):
      ts.append(self.mean_function.c)
    return tuple(ts)


class GPR(_HandleCache):
  """gpflow.models.GPR parameters: data (X, Y), one SE kernel, Gaussian likelihood."""

  def __init__(self, data, kernel: SquaredExponential, mean_function=None, noise_variance=1.0):
    self.data = (_t(data[0]), _t(data[1]))
    self.kernel = kernel
    self.mean_function = Zero() if mean_function is None else mean_function
    self.likelihood = Gaussian(noise_variance)

  def _param_tensors(self):
    ts = [self.data[0], self.data[1], self.kernel.variance, self.kernel.lengthscales, self.likelihood.variance]
    if isinstance(self.mean_function, Constant):
      ts.append(self.mean_function.c)
    return tuple(ts)


class GPModelWrapper:
  """Attribute-forwarding wrapper (upstream models/core.py:30-58)."""

  def __init__(self, model, **attrs):
    self.__dict__["_model"] = model
    for k, v in attrs.items():
      self.__dict__[k] = v

  def __getattr__(self, name):
    return getattr(self.__dict__["_model"], name)

  @property
  def model(self):
    return self.__dict__["_model"]


class KernelRegressor(GPModelWrapper):
  """Deterministic kernel regressor = posterior mean of the wrapped SVGP (upstream models/core.py:61-63)."""

  def __call__(self, x, **kwargs):
    from gpflowpilco_b200.models.predict import predict_mean
    return predict_mean(self.model, x)


class InverseLinkWrapper(GPModelWrapper):
  """invlink(model(x))  (upstream models/core.py:66-71)."""

  def __init__(self, model, invlink):
    super().__init__(model=model, invlink=invlink)

  def __call__(self, *args, **kwargs):
    return self.invlink(self.model(*args, **kwargs))


# ---- bijectors (tfp stand-ins; only what the cart-pole policy uses) -------------------------------------
class Bijector:
  pass


class Scale(Bijector):
  def __init__(self, scale):
    self.scale = float(scale)

  def __call__(self, x):
    return self.scale * x


class Shift(Bijector):
  def __init__(self, shift):
    self.shift = float(shift)

  def __call__(self, x):
    return x + self.shift


class NormalCDF(Bijector):
  def __call__(self, x):
    return 0.5 * torch.erfc(-x * 0.7071067811865476)     # ndtr, upstream utils/bvn.py:38-42


class BijectorChain(Bijector):
  """tfb.Chain: applies right-to-left."""

  def __init__(self, bijectors: Sequence[Bijector]):
    self.bijectors = list(bijectors)

  def __call__(self, x):
    for b in reversed(self.bijectors):
      x = b(x)
    return x

  def squash_parameters(self) -> Tuple[float, float]:
    """(scale, shift) if the chain is Scale o Shift o NormalCDF (the upstream policy link), else raises."""
    kinds = [type(b) for b in self.bijectors]
    if kinds != [Scale, Shift, NormalCDF]:
      raise NotImplementedError("only Chain([Scale, Shift, NormalCDF]) is supported on the fused path")
    return self.bijectors[0].scale, self.bijectors[1].shift
