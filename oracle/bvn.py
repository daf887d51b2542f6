"""Bivariate normal probabilities (Genz) — restatement of upstream gpflow_pilco/utils/bvn.py:67-232 in torch float64.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Used by the multi-dimensional NormalCDF rule
(gpflow_pilco/moment_matching/bijectors.py:59-63)."""
from __future__ import annotations

import math

import torch

DTYPE = torch.float64
_INF = float("inf")


def ndtr(x: torch.Tensor) -> torch.Tensor:
  return 0.5 * torch.erfc(-x / math.sqrt(2.0))                         # bvn.py:38-42


def gauss_legendre(corr: torch.Tensor):
  """Order chosen from the LARGEST |corr| of the whole tensor (bvn.py:221-228); nodes doubled to 1 -+ x (:230-232)."""
  a = corr.abs()
  if bool((a < 0.3).all()):
    x = [0.9324695142031522, 0.6612093864662647, 0.2386191860831970]
    w = [0.1713244923791705, 0.3607615730481384, 0.4679139345726904]
  elif bool((a < 0.75).all()):
    x = [0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802, 0.1252334085114692]
    w = [0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547, 0.2491470458134029]
  else:
    x = [0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
         0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733]
    w = [0.01761400713915212, .04060142980038694, .06267204833410906, 0.08327674157670475, 0.1019301198172404,
         0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259]
  x, w = torch.tensor(x, dtype=DTYPE), torch.tensor(w, dtype=DTYPE)
  return torch.cat([1.0 - x, 1.0 + x]), torch.cat([w, w])


def _bvnu(h: torch.Tensor, k: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
  """bvn.py:105-176 with the default thresholds r = 0.925, hk = asr = -100."""
  hk = h * k
  tp = 2.0 * math.pi
  x, w = gauss_legendre(r)

  def moderate():
    asr = 0.5 * torch.asin(r)
    sn = torch.sin(asr[..., None] * x)
    res = (sn * hk[..., None] - 0.5 * (h ** 2 + k ** 2)[..., None]) / (1 - sn ** 2)
    res = (w * torch.exp(res)).sum(-1)
    return res * asr / tp + ndtr(-h) * ndtr(-k)

  def strong():
    sgn = torch.sign(r)
    _k, _hk = k * sgn, hk * sgn
    _as = 1 - r ** 2
    a = torch.sqrt(_as)
    bs = (h - _k) ** 2
    asr = -0.5 * (bs / _as + _hk)
    c = 0.125 * (4 - _hk)
    d = 0.0125 * (12 - _hk)
    zero = torch.zeros_like(h)
    res = torch.where(asr > -100.0, a * torch.exp(asr) * (1 - c * (bs - _as) * (1 - d * bs) / 3 + c * d * _as ** 2), zero)
    b = torch.sqrt(bs)
    sp = math.sqrt(tp) * ndtr(-b / a)
    res = res - torch.where(_hk > -100.0, torch.exp(-0.5 * _hk) * sp * b * (1 - c * bs * (1 - d * bs) / 3), zero)
    a2 = 0.5 * a
    xs = (a2[..., None] * x) ** 2
    asr2 = -0.5 * (bs[..., None] / xs + _hk[..., None])
    sp2 = 1 + c[..., None] * xs * (1 + 5 * d[..., None] * xs)
    rs = torch.sqrt(1 - xs)
    ep = torch.exp(-0.5 * _hk[..., None] * xs / (1 + rs) ** 2) / rs
    deltas = torch.where(asr2 > -100.0, w * torch.exp(asr2) * (sp2 - ep), torch.zeros_like(asr2))
    partial = (a2 * deltas.sum(-1) - res) / tp
    res = torch.where(r.abs() < 1, partial, zero)
    out = ndtr(-h) - ndtr(-_k) - res                                    # default branch
    out = torch.where(h < 0, ndtr(_k) - ndtr(h) - res, out)
    out = torch.where(h >= _k, -res, out)
    out = torch.where(r > 0, res + ndtr(-torch.maximum(h, _k)), out)
    return out

  with torch.no_grad():
    pass
  res = torch.where(r.abs() < 0.925, moderate(), strong())
  return res.clamp(0.0, 1.0)


def bvnu(dh: torch.Tensor, dk: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
  """P(x > dh, y > dk) for a standard bivariate normal with correlation r (bvn.py:88-101)."""
  dh, dk, r = torch.broadcast_tensors(dh, dk, r)
  safe_h = torch.where(torch.isinf(dh), torch.zeros_like(dh), dh)
  safe_k = torch.where(torch.isinf(dk), torch.zeros_like(dk), dk)
  out = _bvnu(safe_h, safe_k, r)
  out = torch.where(r == 0, ndtr(-dh) * ndtr(-dk), out)
  out = torch.where(dk == -_INF, ndtr(-dh), out)
  out = torch.where(dh == -_INF, ndtr(-dk), out)
  out = torch.where((dh == -_INF) & (dk == -_INF), torch.ones_like(out), out)
  out = torch.where((dh == _INF) | (dk == _INF), torch.zeros_like(out), out)
  return out


def bvn(xl, xu, yl, yu, r) -> torch.Tensor:
  """P(xl < x < xu, yl < y < yu) (bvn.py:67-85)."""
  p = bvnu(xl, yl, r) - bvnu(xu, yl, r) - bvnu(xl, yu, r) + bvnu(xu, yu, r)
  return p.clamp(0.0, 1.0)
