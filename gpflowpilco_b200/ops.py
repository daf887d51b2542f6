"""Tensor-level entry points: torch CUDA float64 tensors in, torch CUDA tensors out, computed by libgpp_b200.so.

torch is used for device memory and streams only; every arithmetic result comes from the C-ABI calls.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from gpflowpilco_b200 import _lib

F64 = torch.float64


def _dev_check(*tensors):
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise RuntimeError("gpflowpilco_b200 ops need CUDA tensors: this path has no CPU implementation")
    if t.dtype != F64:
      raise TypeError(f"expected float64, got {t.dtype}")


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
  return None if t is None else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
  return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _new_info(device) -> torch.Tensor:
  return torch.zeros(1, dtype=torch.int32, device=device)


def raise_if_not_pd(info: torch.Tensor, what: str):
  """Synchronising check of the asynchronous Cholesky flag (1 + first failing batch index)."""
  v = int(info.item())
  if v:
    raise _lib.GppError(-3, f"{what}: input covariance + lengthscale matrix of batch element {v - 1} is not positive definite")


def ekxz(mu, cov, Z, lengthscales, variance: float, check: bool = True) -> torch.Tensor:
  """Psi1 [N,M] (replaces GPflow expectation(p,(kernel,Z)); upstream call sites moment_matching/models.py:62,141,212)."""
  mu, cov, Z, lengthscales = map(_c, (mu, cov, Z, lengthscales))
  _dev_check(mu, cov, Z, lengthscales)
  N, D = mu.shape
  M = Z.shape[0]
  if cov.shape != (N, D, D) or Z.shape[1] != D or lengthscales.shape != (D,):
    raise ValueError("ekxz: inconsistent shapes")
  out = torch.empty(N, M, dtype=F64, device=mu.device)
  info = _new_info(mu.device)
  _lib.check(_lib.load().gpp_ekxz(_ptr(mu), _ptr(cov), N, D, _ptr(Z), M, _ptr(lengthscales), float(variance), _ptr(out),
                                  _ptr(info), _stream()))
  if check:
    raise_if_not_pd(info, "ekxz")
  return out


def ekzxkxz(mu, cov, Z1, lengthscales1, variance1: float, Z2=None, lengthscales2=None, variance2: Optional[float] = None,
            check: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Psi2 [N,M1,M2] (replaces upstream utils/kernel_expectation.py:72-187).  `out`: optional caller-owned result buffer."""
  mu, cov, Z1, lengthscales1, Z2, lengthscales2 = map(_c, (mu, cov, Z1, lengthscales1, Z2, lengthscales2))
  _dev_check(mu, cov, Z1, lengthscales1, Z2, lengthscales2)
  N, D = mu.shape
  M1 = Z1.shape[0]
  M2 = M1 if Z2 is None else Z2.shape[0]
  if cov.shape != (N, D, D) or Z1.shape[1] != D or lengthscales1.shape != (D,):
    raise ValueError("ekzxkxz: inconsistent shapes")
  if (lengthscales2 is None) != (variance2 is None):
    raise ValueError("ekzxkxz: lengthscales2 and variance2 go together")
  if out is None:
    out = torch.empty(N, M1, M2, dtype=F64, device=mu.device)
  elif tuple(out.shape) != (N, M1, M2) or out.dtype != F64 or not out.is_cuda or not out.is_contiguous():
    raise ValueError("ekzxkxz: `out` must be a contiguous CUDA float64 tensor of shape [N,M1,M2]")
  info = _new_info(mu.device)
  _lib.check(_lib.load().gpp_ekzxkxz(_ptr(mu), _ptr(cov), N, D, _ptr(Z1), M1, _ptr(lengthscales1), float(variance1),
                                     _ptr(Z2), M2, _ptr(lengthscales2), float(variance2 or 0.0), _ptr(out), _ptr(info),
                                     _stream()))
  if check:
    raise_if_not_pd(info, "ekzxkxz")
  return out


class GPModelHandle:
  """Owns a gpp_gp_model (cached beta / C weights of one sparse or exact GP)."""

  def __init__(self, Z, lengthscales, variance, q_mu, q_sqrt=None, whiten: bool = True, mean_const=None, W=None,
               kuu_jitter: Sequence[float] | float = 1e-6, model_uncertainty: bool = True):
    Z, lengthscales, variance, q_mu, q_sqrt, mean_const, W = map(_c, (Z, lengthscales, variance, q_mu, q_sqrt, mean_const, W))
    _dev_check(Z, lengthscales, variance, q_mu, q_sqrt, mean_const, W)
    L, M, D = Z.shape
    if lengthscales.shape != (L, D) or variance.shape != (L,) or q_mu.shape != (M, L):
      raise ValueError("GPModelHandle: inconsistent parameter shapes")
    if q_sqrt is not None and q_sqrt.shape != (L, M, M):
      raise ValueError("GPModelHandle: q_sqrt must be [L,M,M]")
    P = L if W is None else W.shape[0]
    if W is not None and W.shape != (P, L):
      raise ValueError("GPModelHandle: W must be [P,L]")
    if mean_const is not None and mean_const.shape != (P,):
      raise ValueError("GPModelHandle: mean_const must be [P]")
    jit = [float(kuu_jitter)] * L if isinstance(kuu_jitter, (int, float)) else [float(j) for j in kuu_jitter]
    self.L, self.M, self.D, self.P = L, M, D, P
    self.device = Z.device
    self._h = ctypes.c_void_p()
    arr = (ctypes.c_double * L)(*jit)
    with torch.cuda.device(self.device):
      _lib.check(_lib.load().gpp_gp_model_create(ctypes.byref(self._h), L, M, D, _ptr(Z), _ptr(lengthscales), _ptr(variance),
                                                 _ptr(q_mu), _ptr(q_sqrt), int(bool(whiten)), _ptr(mean_const), _ptr(W), P,
                                                 arr, int(bool(model_uncertainty)), _stream()))
    self._ws = None
    self._params = (lengthscales, variance, Z, torch.zeros(P, dtype=F64, device=self.device) if mean_const is None else mean_const)

  def parameters(self):
    """(lengthscales [L,D], variance [L], Z [L,M,D], mean_const [P]) as passed at construction."""
    return self._params

  def __del__(self):
    try:                                  # at interpreter shutdown module globals (ctypes, _lib) may already be gone
      h = getattr(self, "_h", None)
      if h is not None and h.value:
        _lib.load().gpp_gp_model_destroy(h)
        self._h = None
    except Exception:
      pass

  def weights(self):
    """(beta [L,M], C [L,M,M]) copies, for tests."""
    beta = torch.empty(self.L, self.M, dtype=F64, device=self.device)
    C = torch.empty(self.L, self.M, self.M, dtype=F64, device=self.device)
    _lib.check(_lib.load().gpp_gp_model_weights(self._h, _ptr(beta), _ptr(C), _stream()))
    return beta, C

  def workspace(self, N: int) -> torch.Tensor:
    need = _lib.load().gpp_mm_gp_predict_workspace_bytes(self._h, N)
    if self._ws is None or self._ws.numel() < need:
      self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
    return self._ws

  def _same_device(self, *tensors):
    """a handle's device arrays live on ONE device: inputs from another GPU of the process are refused, not dereferenced"""
    for t in tensors:
      if t is not None and t.device != self.device:
        raise ValueError(f"tensor on {t.device}, model handle on {self.device}")

  def predict(self, m, S, full_output_cov: bool = True, jitter: float = 0.0, out=None, check: bool = True):
    """(f1 [N,P], Sff [N,P,P], cross [N,D,P] pre-inverted) — upstream moment_matching/models.py:44-299."""
    m, S = _c(m), _c(S)
    _dev_check(m, S)
    N = m.shape[0]
    if m.shape != (N, self.D) or S.shape != (N, self.D, self.D):
      raise ValueError(f"predict: expected m [N,{self.D}] and S [N,{self.D},{self.D}]")
    if out is None:
      f1 = torch.empty(N, self.P, dtype=F64, device=self.device)
      Sff = torch.empty(N, self.P, self.P, dtype=F64, device=self.device)
      cross = torch.empty(N, self.D, self.P, dtype=F64, device=self.device)
    else:
      f1, Sff, cross = out
    ws = self.workspace(N)
    info = _new_info(self.device)
    self._same_device(m, S)
    with torch.cuda.device(self.device):      # the library launches on the CURRENT device / its current stream
      _lib.check(_lib.load().gpp_mm_gp_predict_fwd(self._h, _ptr(m), _ptr(S), N, _ptr(f1), _ptr(Sff), _ptr(cross),
                                                   int(bool(full_output_cov)), float(jitter), _ptr(ws), ws.numel(), _ptr(info),
                                                   _stream()))
    if check:
      raise_if_not_pd(info, "mm_gp_predict")
    return f1, Sff, cross


def _predict_bwd(self, m, S, f1_bar=None, Sff_bar=None, cross_bar=None, full_output_cov: bool = True, check: bool = True):
  """Adjoints (m_bar [N,D], S_bar [N,D,D] symmetric) of `predict` given the output adjoints (None = zero)."""
  m, S, f1_bar, Sff_bar, cross_bar = map(_c, (m, S, f1_bar, Sff_bar, cross_bar))
  _dev_check(m, S, f1_bar, Sff_bar, cross_bar)
  N = m.shape[0]
  if m.shape != (N, self.D) or S.shape != (N, self.D, self.D):
    raise ValueError(f"predict_bwd: expected m [N,{self.D}] and S [N,{self.D},{self.D}]")
  for t, shape, name in ((f1_bar, (N, self.P), "f1_bar"), (Sff_bar, (N, self.P, self.P), "Sff_bar"), (cross_bar, (N, self.D, self.P), "cross_bar")):
    if t is not None and tuple(t.shape) != shape:
      raise ValueError(f"predict_bwd: {name} must have shape {shape}")
  lib = _lib.load()
  need = lib.gpp_mm_gp_predict_bwd_workspace_bytes(self._h, N)
  if getattr(self, "_ws_bwd", None) is None or self._ws_bwd.numel() < need:
    self._ws_bwd = torch.empty(need, dtype=torch.uint8, device=self.device)
  m_bar = torch.empty(N, self.D, dtype=F64, device=self.device)
  S_bar = torch.empty(N, self.D, self.D, dtype=F64, device=self.device)
  info = _new_info(self.device)
  self._same_device(m, S, f1_bar, Sff_bar, cross_bar)
  with torch.cuda.device(self.device):
    _lib.check(lib.gpp_mm_gp_predict_bwd(self._h, _ptr(m), _ptr(S), N, _ptr(f1_bar), _ptr(Sff_bar), _ptr(cross_bar),
                                         int(bool(full_output_cov)), _ptr(m_bar), _ptr(S_bar), _ptr(self._ws_bwd), self._ws_bwd.numel(),
                                         _ptr(info), _stream()))
  if check:
    raise_if_not_pd(info, "mm_gp_predict_bwd")
  return m_bar, S_bar


GPModelHandle.predict_bwd = _predict_bwd
