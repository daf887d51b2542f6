"""The reference-side binding as code: plug the B200 path into an UNMODIFIED upstream `gpflow_pilco` installation.

  import gpflow_pilco
  from gpflowpilco_b200.adapters import upstream
  upstream.install()            # registers at upstream's dispatcher keys, overrides the MM closure factory

Import-guarded: importing this module never imports tensorflow / gpflow / gpflow_pilco; `install()` does, and raises ImportError with
a clear message when they are absent (they are not installable in the build image — the registration mechanics are exercised in
tests/test_adapter_upstream.py with the numpy stand-ins of oracle/refshim, which is test infrastructure and is never imported here).

What is bound (all file:line relative to the upstream repository root):

  moment_matching.core.dispatcher keys (GaussianMoments, gpflow.models.SVGP) and (GaussianMoments, gpflow.models.GPR)
      gpflow_pilco/moment_matching/models.py:44-126  ->  gpp_gp_model_create + gpp_mm_gp_predict_fwd (/_bwd under tf.custom_gradient)
  MomentMatchingPILCO._policy_loss_closure
      gpflow_pilco/loops/pilco.py:192-220            ->  gpp_rollout_mm_fwd_save + gpp_rollout_mm_bwd (one persistent kernel per sweep)
      when the loop has the cart-pole structure (TrigonometricEncoder, InverseLinkWrapper(KernelRegressor(SVGP)) policy with the
      Chain[Scale, Shift, NormalCDF] link, SeparateIndependent SVGP drift, GaussianObjective); anything else keeps upstream's closure,
      whose moment_matching calls then reach the rules registered above.
  GradientDescent.minimize (gpflow_pilco/utils/optimizers.py:52-56) is the CALLER: `tape.gradient(loss, variables)` flows through
      the tf.custom_gradient wrappers below; nothing to override.

Objects are read by attribute protocol (kernel.kernels[i].variance / .lengthscales, inducing_variable.inducing_variables[i].Z, q_mu,
q_sqrt, whiten, mean_function.c, ...), so real GPflow objects and structural stand-ins are handled alike.  Tensors cross by DLPack when
the framework tensor lives on the GPU (zero copy), through host memory otherwise.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

F64 = torch.float64
_STATE: Dict[str, Any] = {"installed": False, "originals": {}}


# ---------------------------------------------------------------------------------------------------------
# tensors in and out
# ---------------------------------------------------------------------------------------------------------
def to_torch(x, device: torch.device) -> torch.Tensor:
  """framework tensor / variable / array -> float64 torch tensor on `device` (DLPack when possible)."""
  if isinstance(x, torch.Tensor):
    return x.to(device=device, dtype=F64)
  if hasattr(x, "to_dense"):                      # tf.linalg.LinearOperator
    x = x.to_dense()
  try:                                            # TensorFlow tensor on a GPU: zero copy
    import tensorflow as tf                       # noqa: WPS433
    if hasattr(tf, "experimental") and hasattr(tf.experimental, "dlpack") and isinstance(x, (tf.Tensor, tf.Variable)):
      t = torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(x)))
      return t.to(device=device, dtype=F64)
  except Exception:                               # no TF, a stand-in TF, or a tensor DLPack cannot export: go through the host
    pass
  return torch.as_tensor(np.asarray(x), dtype=F64).to(device)


def from_torch(t: torch.Tensor):
  """torch tensor -> the framework's tensor type (tf.Tensor when TensorFlow is importable, numpy otherwise)."""
  try:
    import tensorflow as tf
    if hasattr(tf, "experimental") and hasattr(tf.experimental, "dlpack") and t.is_cuda:
      return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))
    return tf.convert_to_tensor(t.detach().cpu().numpy())
  except ImportError:
    return t.detach().cpu().numpy()


def device_pointer(t: torch.Tensor) -> int:
  """Raw device pointer of a contiguous float64 CUDA tensor — what the C ABI takes (include/gpp_b200.h)."""
  if not (t.is_cuda and t.dtype == F64 and t.is_contiguous()):
    raise ValueError("device_pointer: need a contiguous float64 CUDA tensor")
  return t.data_ptr()


# ---------------------------------------------------------------------------------------------------------
# GPflow-shaped objects -> parameter dictionaries
# ---------------------------------------------------------------------------------------------------------
def _latent_kernels(kernel) -> list:
  if hasattr(kernel, "kernels"):                           # SeparateIndependent / LinearCoregionalization
    return list(kernel.kernels)
  if hasattr(kernel, "kernel") and hasattr(kernel, "num_latent_gps"):     # SharedIndependent
    return [kernel.kernel] * int(kernel.num_latent_gps)
  return [kernel]


def _latent_inducing(iv, L: int) -> list:
  """upstream utils/kernel_expectation.py:41-69 (unpack_multioutput)"""
  if hasattr(iv, "inducing_variables"):                    # Separate / SharedIndependentInducingVariables (GPflow >= 2.2: both expose it)
    zs = list(iv.inducing_variables)
    return [z.Z for z in (zs if len(zs) == L else zs * L)]
  if hasattr(iv, "inducing_variable"):
    return [iv.inducing_variable.Z] * L
  return [iv.Z] * L


def _mean_constant(mean_function, P: int):
  name = type(mean_function).__name__
  if name == "Zero":
    return None
  if name == "Constant":
    return np.broadcast_to(np.asarray(mean_function.c, dtype=np.float64).reshape(-1), (P,)).copy()
  raise NotImplementedError(f"mean function {name}: upstream supports Zero and Constant (moment_matching/models.py:288-291)")


def svgp_parameters(model) -> Dict[str, Any]:
  """Everything gpp_gp_model_create needs, as numpy arrays, from a gpflow.models.SVGP-shaped object."""
  ks = _latent_kernels(model.kernel)
  L = len(ks)
  Zs = [np.asarray(z, dtype=np.float64) for z in _latent_inducing(model.inducing_variable, L)]
  for k in ks:
    if getattr(k, "active_dims", None) is not None and not isinstance(k.active_dims, slice):
      raise NotImplementedError("adapter: kernels with active_dims go through upstream's own rule")
  D = Zs[0].shape[-1]
  ell = np.stack([np.broadcast_to(np.asarray(k.lengthscales, dtype=np.float64), (D,)) for k in ks])
  var = np.array([float(np.asarray(k.variance)) for k in ks])
  W = np.asarray(model.kernel.W, dtype=np.float64) if hasattr(model.kernel, "W") else None
  P = L if W is None else W.shape[0]
  q_sqrt = getattr(model, "q_sqrt", None)
  return {"Z": np.stack(Zs), "lengthscales": ell, "variance": var, "q_mu": np.asarray(model.q_mu, dtype=np.float64),
          "q_sqrt": None if q_sqrt is None else np.asarray(q_sqrt, dtype=np.float64), "whiten": bool(model.whiten),
          "mean_const": _mean_constant(model.mean_function, P), "W": W}


def gpr_parameters(model) -> Dict[str, Any]:
  X, Y = (np.asarray(a, dtype=np.float64) for a in model.data)
  if Y.shape[-1] != 1:
    raise NotImplementedError("GPR moment matching is single-output, like upstream moment_matching/models.py:44-111")
  c = _mean_constant(model.mean_function, 1)
  D = X.shape[-1]
  return {"Z": X[None], "lengthscales": np.broadcast_to(np.asarray(model.kernel.lengthscales, dtype=np.float64), (D,))[None].copy(),
          "variance": np.array([float(np.asarray(model.kernel.variance))]), "q_mu": Y if c is None else Y - c, "q_sqrt": None,
          "whiten": False, "mean_const": c, "W": None, "kuu_jitter": float(np.asarray(model.likelihood.variance))}


# ---------------------------------------------------------------------------------------------------------
# backend: the CUDA library (tests substitute their own object with the same two methods to exercise the glue without a GPU)
# ---------------------------------------------------------------------------------------------------------
class CudaBackend:
  """predict(): gpp_mm_gp_predict_fwd on a cached handle; rollout(): the fused moment-matched rollout."""

  def __init__(self, device: Optional[torch.device] = None):
    self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    self._handles: Dict[Tuple, Any] = {}

  def _handle(self, params: Dict[str, Any], model_uncertainty: bool, key):
    from gpflowpilco_b200 import ops
    sig = (key, bool(model_uncertainty), hash(tuple(np.asarray(v).tobytes() if v is not None and not isinstance(v, (bool, float)) else v
                                                    for v in params.values())))
    hit = self._handles.get(key)
    if hit is not None and hit[0] == sig:
      return hit[1]
    T = lambda a: None if a is None else torch.as_tensor(a, dtype=F64, device=self.device)
    h = ops.GPModelHandle(T(params["Z"]), T(params["lengthscales"]), T(params["variance"]), T(params["q_mu"]), T(params["q_sqrt"]),
                          whiten=params["whiten"], mean_const=T(params["mean_const"]), W=T(params["W"]),
                          kuu_jitter=params.get("kuu_jitter", 1e-6), model_uncertainty=model_uncertainty)
    self._handles[key] = (sig, h)
    return h

  def predict(self, params, m, S, full_output_cov: bool, model_uncertainty: bool, jitter: float, key=None):
    h = self._handle(params, model_uncertainty, key)
    return h.predict(to_torch(m, self.device), to_torch(S, self.device), full_output_cov=full_output_cov, jitter=jitter)

  def rollout(self, dyn_params, policy: Dict[str, Any], m0, S0, horizon: int, active_dims: Sequence[int], target, W, key=None):
    from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=F64, device=self.device)
    h = self._handle(dyn_params, True, key)
    pol = PolicyParams(T(policy["Z"]), T(policy["lengthscales"]), T(policy["variance"]), T(policy["q_mu"]), whiten=policy["whiten"],
                       jitter=policy.get("jitter", 1e-6), squash_scale=policy["scale"], squash_shift=policy["shift"])
    return rollout_mm(h, pol, to_torch(m0, self.device), to_torch(S0, self.device), horizon, tuple(active_dims), T(target), T(W)).loss


_BACKEND: Optional[Any] = None


def set_backend(backend) -> None:
  global _BACKEND
  _BACKEND = backend


def backend():
  global _BACKEND
  if _BACKEND is None:
    _BACKEND = CudaBackend()
  return _BACKEND


# ---------------------------------------------------------------------------------------------------------
# the rules, with upstream's signatures
# ---------------------------------------------------------------------------------------------------------
def _match(x, f1, Sff, cross, full_output_cov: bool):
  """GaussianMatch exactly as upstream assembles it (moment_matching/models.py:293-299): centred moments, pre-inverted cross term."""
  import tensorflow as tf
  from gpflow_pilco.moment_matching import GaussianMatch, GaussianMoments
  f1, Sff, cross = from_torch(f1), from_torch(Sff), from_torch(cross)
  if not full_output_cov:
    Sff = tf.linalg.LinearOperatorDiag(tf.linalg.diag_part(Sff))
  return GaussianMatch(x=x, y=GaussianMoments(moments=(f1, Sff), centered=True), cross=(cross, True))


def mm_gauss_svgp(x, model, /, full_output_cov: bool = True, model_uncertainty: bool = True, jitter: float = 0.0):
  """registered at (GaussianMoments, gpflow.models.SVGP) — replaces gpflow_pilco/moment_matching/models.py:114-299"""
  try:
    params = svgp_parameters(model)
  except NotImplementedError:
    return _STATE["originals"]["svgp"](x, model, full_output_cov=full_output_cov, model_uncertainty=model_uncertainty, jitter=jitter)
  f1, Sff, cross = backend().predict(params, x.mean(), x.covariance(), full_output_cov, model_uncertainty, float(jitter), key=id(model))
  return _match(x, f1, Sff, cross, full_output_cov)


def mm_gauss_gpr(x, model, /, full_output_cov: bool = True, model_uncertainty: bool = True, jitter: float = 0.0):
  """registered at (GaussianMoments, gpflow.models.GPR) — replaces gpflow_pilco/moment_matching/models.py:44-111"""
  params = gpr_parameters(model)
  f1, Sff, cross = backend().predict(params, x.mean(), x.covariance(), full_output_cov, model_uncertainty, float(jitter), key=id(model))
  return _match(x, f1, Sff, cross, full_output_cov)


def _cartpole_structure(loop) -> Optional[Dict[str, Any]]:
  """The structure the fused rollout computes (and nothing else): returns its parameters, or None to keep upstream's closure."""
  enc, pol, drift, obj = loop.encoder, loop.policy, loop.drift, loop.objective
  if type(enc).__name__ != "TrigonometricEncoder" or type(obj).__name__ != "GaussianObjective":
    return None
  if type(pol).__name__ != "InverseLinkWrapper" or type(pol.model).__name__ != "KernelRegressor":
    return None
  link = pol.invlink
  bij = list(getattr(link, "bijectors", []))
  if len(bij) != 3 or any(not type(b).__name__.endswith(n) for b, n in zip(bij, ("Scale", "Shift", "NormalCDF"))):
    return None
  svgp = pol.model.model
  try:
    pp = svgp_parameters(svgp)
    dp = svgp_parameters(drift)
  except (NotImplementedError, AttributeError):
    return None
  if pp["Z"].shape[0] != 1 or pp["mean_const"] is not None and np.any(pp["mean_const"] != 0.0) or pp["W"] is not None or dp["W"] is not None:
    return None
  if loop.diffusion is not None:
    return None
  policy = {"Z": pp["Z"], "lengthscales": pp["lengthscales"], "variance": pp["variance"], "q_mu": pp["q_mu"][:, 0][None], "whiten": pp["whiten"],
            "scale": float(np.asarray(bij[0].scale)), "shift": float(np.asarray(bij[1].shift))}
  return {"dynamics": dp, "policy": policy, "active_dims": tuple(int(a) for a in enc.active_dims), "target": np.asarray(obj.target),
          "W": np.asarray(obj.precis)}


def mm_policy_loss_closure(self, state_initializer: Callable, initial_time: float, solution_times, compile: bool = True, **kwargs):
  """replaces MomentMatchingPILCO._policy_loss_closure (gpflow_pilco/loops/pilco.py:192-220): same signature, same return value
  (a callable () -> loss[N]).  The structure is re-read at every call, so parameter updates by the optimiser are seen."""
  original = _STATE["originals"]["mm_closure"]
  st = np.asarray(solution_times, dtype=np.float64)
  unit_steps = len(st) > 0 and abs(st[0] - float(initial_time) - 1.0) < 1e-12 and (len(st) < 2 or np.allclose(np.diff(st), 1.0))
  if not unit_steps or kwargs or _cartpole_structure(self) is None:
    return original(self, state_initializer=state_initializer, initial_time=initial_time, solution_times=solution_times, compile=compile,
                    **kwargs)

  def _closure():
    spec = _cartpole_structure(self)
    mx, Sxx = state_initializer()
    tfw = _tf_custom_gradient_rollout()
    if tfw is not None and isinstance(backend(), CudaBackend):
      return tfw(self, spec, mx, Sxx, len(st))
    loss = backend().rollout(spec["dynamics"], spec["policy"], mx, Sxx, len(st), spec["active_dims"], spec["target"], spec["W"], key=id(self.drift))
    return from_torch(loss)
  return _closure


def _tf_custom_gradient_rollout():
  """The gradient shim of INTEGRATION.md §3 for real TensorFlow: forward = gpp_rollout_mm_fwd_save, grad = gpp_rollout_mm_bwd +
  gpp_policy_prepare_bwd, so that upstream's tape.gradient(loss, policy.trainable_variables) (utils/optimizers.py:52-56) works.
  Returns None when TensorFlow (with custom_gradient) is not importable."""
  try:
    import tensorflow as tf
  except ImportError:
    return None
  if not hasattr(tf, "custom_gradient"):
    return None

  def run(loop, spec, mx, Sxx, H):
    from gpflowpilco_b200.autograd import rollout_mm_loss
    be = backend()
    dev = be.device
    svgp = loop.policy.model.model
    Zvar = _latent_inducing(svgp.inducing_variable, 1)[0]
    kern = _latent_kernels(svgp.kernel)[0]
    handle = be._handle(spec["dynamics"], True, id(loop.drift))
    T = lambda a: torch.as_tensor(np.asarray(a), dtype=F64, device=dev)

    @tf.custom_gradient
    def f(Z, q_mu, ell, m0, S0):
      leaves = [to_torch(v, dev).clone().requires_grad_(True) for v in (Z, q_mu, ell, m0, S0)]
      loss = rollout_mm_loss(handle, leaves[0][None], leaves[2].reshape(1, -1), T(spec["policy"]["variance"]), leaves[1][:, 0][None], leaves[3],
                             leaves[4], H, spec["active_dims"], T(spec["target"]), T(spec["W"]), squash_scale=spec["policy"]["scale"],
                             squash_shift=spec["policy"]["shift"], whiten=spec["policy"]["whiten"])

      def grad(dloss):
        gs = torch.autograd.grad(loss, leaves, grad_outputs=to_torch(dloss, dev), allow_unused=True)
        return tuple(from_torch(g if g is not None else torch.zeros_like(l)) for g, l in zip(gs, leaves))
      return from_torch(loss.detach()), grad
    return f(Zvar, svgp.q_mu, kern.lengthscales, mx, Sxx)
  return run


# ---------------------------------------------------------------------------------------------------------
def install(rules: bool = True, closures: bool = True) -> None:
  """Register the B200 path onto the imported upstream package (idempotent)."""
  try:
    import gpflow
    import gpflow_pilco  # noqa: F401
    from gpflow_pilco.loops import pilco as up_pilco
    from gpflow_pilco.moment_matching import GaussianMoments
    from gpflow_pilco.moment_matching.core import dispatcher
  except ImportError as e:
    raise ImportError("gpflowpilco_b200.adapters.upstream.install() needs upstream gpflow_pilco with its dependencies (tensorflow, gpflow, "
                      f"tensorflow_probability) importable: {e}") from e
  if _STATE["installed"]:
    return
  if rules:
    _STATE["originals"]["svgp"] = dispatcher.dispatch(GaussianMoments, gpflow.models.SVGP)
    _STATE["originals"]["gpr"] = dispatcher.dispatch(GaussianMoments, gpflow.models.GPR)
    dispatcher.register(GaussianMoments, gpflow.models.SVGP)(mm_gauss_svgp)      # same keys as moment_matching/models.py:44,114:
    dispatcher.register(GaussianMoments, gpflow.models.GPR)(mm_gauss_gpr)        # multipledispatch lets the last registration win
  if closures:
    _STATE["originals"]["mm_closure"] = up_pilco.MomentMatchingPILCO._policy_loss_closure
    up_pilco.MomentMatchingPILCO._policy_loss_closure = mm_policy_loss_closure
  _STATE["installed"] = True


def uninstall() -> None:
  if not _STATE["installed"]:
    return
  import gpflow
  from gpflow_pilco.loops import pilco as up_pilco
  from gpflow_pilco.moment_matching import GaussianMoments
  from gpflow_pilco.moment_matching.core import dispatcher
  o = _STATE["originals"]
  if "svgp" in o:
    dispatcher.register(GaussianMoments, gpflow.models.SVGP)(o["svgp"])
    dispatcher.register(GaussianMoments, gpflow.models.GPR)(o["gpr"])
  if "mm_closure" in o:
    up_pilco.MomentMatchingPILCO._policy_loss_closure = o["mm_closure"]
  _STATE["installed"] = False
  _STATE["originals"] = {}
