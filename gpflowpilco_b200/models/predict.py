"""Sample-path predict of the wrapped SVGP: posterior mean  Kxz Kuu^-1 m  (GPflow SVGP.predict_f(x)[0], called by
upstream models/core.py:61-63).  k(x, Z) is the zero-covariance limit of Psi1, evaluated by gpp_ekxz."""
from __future__ import annotations

import torch

from gpflowpilco_b200 import ops
from gpflowpilco_b200.models.core import Constant, SVGP


def predict_mean(model: SVGP, x: torch.Tensor) -> torch.Tensor:
  from gpflowpilco_b200.moment_matching.models import svgp_handle
  h = svgp_handle(model, model_uncertainty=False)
  beta, _ = h.weights()
  ell, var, Z, mean = h.parameters()
  flat = x.reshape(-1, x.shape[-1]).contiguous()
  zero = torch.zeros(flat.shape[0], h.D, h.D, dtype=flat.dtype, device=flat.device)
  outs = []
  for l in range(h.L):
    k = ops.ekxz(flat, zero, Z[l], ell[l], float(var[l]), check=False)          # [N,M]
    outs.append(k @ beta[l])
  f = torch.stack(outs, -1)
  W = getattr(model.kernel, "W", None)
  if W is not None:
    f = f @ W.T
  return (f + mean).reshape(*x.shape[:-1], f.shape[-1])
