"""CUDA-graph replay of a whole policy-gradient evaluation.

A moment-matched rollout is a chain of H x ~8 small dependent launches; at one rollout (config #1) the GPU is idle between them
for about as long as it works.  Every C entry point of the path only enqueues work on the caller's stream (no allocation, no
synchronisation), so the forward rollout, the reverse sweep and the policy-weight adjoint can be captured ONCE and replayed with
new parameter values copied into the captured input buffers.  This is the device-side counterpart of upstream's
`tf.function(closure)` compile step (gpflow_pilco/loops/pilco.py:214-218): streams and graphs instead of a tracing compiler.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from gpflowpilco_b200.autograd import rollout_mm_loss
from gpflowpilco_b200.ops import GPModelHandle

__all__ = ("GraphedMMPolicyGradient",)


class GraphedMMPolicyGradient:
  """loss[N] of the moment-matched rollout and the gradients of its SUM w.r.t. the policy's (Z, lengthscales, q_mu), as one graph.

    g = GraphedMMPolicyGradient(dynamics, Z, lengthscales, variance, q_mu, m0, S0, horizon, active_dims, target, W, ...)
    loss, (dZ, dell, dq_mu) = g(Z, lengthscales, q_mu)          # copies the values in, replays, returns the captured outputs

  The returned tensors are the graph's own output buffers: they are overwritten by the next call (clone them to keep them).
  The not-positive-definite flags are not read inside the graph; `check=True` (default) tests the loss for NaN after the replay,
  which is where a failed factorisation shows up.
  """

  def __init__(self, dynamics: GPModelHandle, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor,
               m0: torch.Tensor, S0: torch.Tensor, horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor,
               cost_W: torch.Tensor, squash_scale: float = 1.0, squash_shift: float = -0.5, whiten: bool = True, jitter: float = 1e-6,
               warmup: int = 2):
    if not Z.is_cuda:
      raise ValueError("GraphedMMPolicyGradient: tensors must live on the CUDA device")
    self._vars = tuple(t.detach().clone().requires_grad_(True) for t in (Z, lengthscales, q_mu))
    self._m0, self._S0 = m0.detach().clone(), S0.detach().clone()
    variance, cost_target, cost_W = variance.detach().clone(), cost_target.detach().clone(), cost_W.detach().clone()
    # the captured kernels read these buffers (and the model behind `dynamics`) on every replay: they live as long as the graph
    self._keep = (dynamics, variance, cost_target, cost_W)

    def evaluate():
      loss = rollout_mm_loss(dynamics, self._vars[0], self._vars[1], variance, self._vars[2], self._m0, self._S0, horizon, active_dims,
                             cost_target, cost_W, squash_scale=squash_scale, squash_shift=squash_shift, whiten=whiten, jitter=jitter,
                             check=False)
      grads = torch.autograd.grad(loss.sum(), self._vars)
      return loss.detach(), grads

    # warm-up on a side stream (first-call attribute set-up, allocator state), then capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(max(1, warmup)):
        evaluate()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self.graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self.graph):
      self.loss, self.grads = evaluate()

  def __call__(self, Z: torch.Tensor, lengthscales: torch.Tensor, q_mu: torch.Tensor, m0: torch.Tensor = None, S0: torch.Tensor = None,
               check: bool = True) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    with torch.no_grad():
      for dst, src in zip(self._vars, (Z, lengthscales, q_mu)):
        dst.copy_(src)
      if m0 is not None:
        self._m0.copy_(m0)
      if S0 is not None:
        self._S0.copy_(S0)
    self.graph.replay()
    if check and not bool(torch.isfinite(self.loss).all()):
      raise ArithmeticError("GraphedMMPolicyGradient: non-finite loss (a covariance stopped being positive definite during the rollout)")
    return self.loss, self.grads
