"""Moment-matching rules for GP models: the registrations of upstream gpflow_pilco/moment_matching/models.py
((GaussianMoments, InverseLinkWrapper) :27, KernelRegressor :34, GPR :44, SVGP :114) re-registered on the
CUDA path.  Same keyword contract: full_output_cov=True, model_uncertainty=True, jitter=0.0.
"""
from __future__ import annotations

from functools import partial, update_wrapper

import torch

from gpflowpilco_b200 import ops
from gpflowpilco_b200.models.core import (GPR, SVGP, Constant, InverseLinkWrapper, KernelRegressor,
                                          LinearCoregionalization, Zero)
from gpflowpilco_b200.moment_matching.core import Chain, GaussianMatch, GaussianMoments, dispatcher

DEFAULT_JITTER = 1e-6   # gpflow.config.default_jitter(), used inside Kuu at upstream moment_matching/models.py:145,216


def _mean_const(model, P, device):
  mf = model.mean_function
  if isinstance(mf, Constant):
    c = mf.c.to(device)
    return c.expand(P).contiguous() if c.numel() == 1 else c
  if isinstance(mf, Zero) or mf is None:
    return None
  raise NotImplementedError("only Zero / Constant mean functions are supported (upstream models.py:55-56,188-191,288-291)")


def _active_dims(kernels):
  dims = {k.active_dims for k in kernels}
  if len(dims) != 1:
    raise NotImplementedError("all latent kernels must share the same active_dims on the fused path")
  return dims.pop()


def svgp_handle(model: SVGP, model_uncertainty: bool) -> ops.GPModelHandle:
  def build():
    ks, Zs = model.latent_kernels(), model.latent_inducing()
    dev = model.q_mu.device
    ad = _active_dims(ks)
    Zsel = [z if ad is None else z[:, list(ad)] for z in Zs]
    D = Zsel[0].shape[-1]
    W = model.kernel.W if isinstance(model.kernel, LinearCoregionalization) else None
    P = len(ks) if W is None else W.shape[0]
    return ops.GPModelHandle(torch.stack(Zsel), torch.stack([k.ell(D) for k in ks]), torch.stack([k.variance.reshape(()) for k in ks]),
                             model.q_mu, model.q_sqrt, whiten=model.whiten, mean_const=_mean_const(model, P, dev), W=W,
                             kuu_jitter=DEFAULT_JITTER if getattr(model, 'kuu_jitter', None) is None else model.kuu_jitter,
                             model_uncertainty=model_uncertainty)
  return model.cached_handle(("svgp", bool(model_uncertainty)), build)


def gpr_handle(model: GPR, model_uncertainty: bool) -> ops.GPModelHandle:
  def build():
    X, Y = model.data
    k = model.kernel
    if Y.shape[-1] != 1:
      raise NotImplementedError("GPR moment matching is single-output, like upstream moment_matching/models.py:44-111")
    Xs = X if k.active_dims is None else X[:, list(k.active_dims)]
    c = _mean_const(model, 1, X.device)
    Yc = Y if c is None else Y - c
    return ops.GPModelHandle(Xs[None], k.ell(Xs.shape[-1])[None], k.variance.reshape(1), Yc, None, whiten=False, mean_const=c,
                             kuu_jitter=float(model.likelihood.variance), model_uncertainty=model_uncertainty)
  return model.cached_handle(("gpr", bool(model_uncertainty)), build)


def _predict(x: GaussianMoments, handle: ops.GPModelHandle, active_dims, full_output_cov, jitter, check=True) -> GaussianMatch:
  m, S = x.mean(), x.covariance(dense=True)
  if active_dims is not None:
    idx = list(active_dims)
    ms, Ss = m[..., idx], S[..., idx, :][..., :, idx]
  else:
    ms, Ss = m, S
  f1, Sff, cross = handle.predict(ms, Ss, full_output_cov=full_output_cov, jitter=jitter, check=check)
  if active_dims is not None:
    # GaussianMatch.x is the FULL state: rows of Sxx^-1 Cov(x, f) that belong to inactive dims are exactly zero (f does not depend
    # on them), so the pre-inverted cross term of the sliced problem is scattered into a zero [N, D, P] block.  (Upstream returns
    # the sliced rows, models.py:264-277, and then mis-multiplies them in cross_covariance(preinv=False) / joint().)
    full = torch.zeros(*m.shape[:-1], m.shape[-1], cross.shape[-1], dtype=cross.dtype, device=cross.device)
    full[..., list(active_dims), :] = cross
    cross = full
  return GaussianMatch(x=x, y=GaussianMoments(moments=(f1, Sff), centered=True), cross=(cross, True))


@dispatcher.register(GaussianMoments, SVGP)
def _mm_gauss_svgp(x: GaussianMoments, model: SVGP, full_output_cov: bool = True, model_uncertainty: bool = True,
                   jitter: float = 0.0, check: bool = True):
  """`check=False` skips the synchronising read of the not-positive-definite flag (for timed loops)."""
  h = svgp_handle(model, model_uncertainty)
  return _predict(x, h, _active_dims(model.latent_kernels()), full_output_cov, jitter, check)


@dispatcher.register(GaussianMoments, GPR)
def _mm_gauss_gpr(x: GaussianMoments, model: GPR, full_output_cov: bool = True, model_uncertainty: bool = True,
                  jitter: float = 0.0, check: bool = True):
  h = gpr_handle(model, model_uncertainty)
  return _predict(x, h, model.kernel.active_dims, full_output_cov, jitter, check)


@dispatcher.register(GaussianMoments, KernelRegressor)
def _mm_gauss_kr(x: GaussianMoments, regressor: KernelRegressor, **kwargs):
  uncertainty = kwargs.pop("model_uncertainty", False)
  assert not uncertainty, ValueError("Kernel regressors have no uncertainty.")
  return dispatcher(x, regressor.model, model_uncertainty=False, **kwargs)


@dispatcher.register(GaussianMoments, InverseLinkWrapper)
def _mm_gauss_invlink(x: GaussianMoments, wrapper: InverseLinkWrapper, **kw):
  base = update_wrapper(partial(wrapper.model, **kw), wrapper.model) if kw else wrapper.model
  return dispatcher(x, Chain(wrapper.invlink, base))
