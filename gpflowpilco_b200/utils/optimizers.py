"""GradientDescent — the caller of the policy-loss closures (upstream gpflow_pilco/utils/optimizers.py:22-78), kept because
it defines the gradient contract of the hot path: `loss = closure()` under a tape, `tape.gradient(loss,
variables)` (a VECTOR loss is summed), optional `transform(*grads)` (upstream passes `tf.clip_by_global_norm`), then
`optimizer.apply_gradients`.  Host orchestration only: the arithmetic stays in the closure's CUDA kernels and the
optimiser's element-wise device ops; nothing here loops over data.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

__all__ = ("GradientDescent", "clip_by_global_norm")


def clip_by_global_norm(clip_norm: float) -> Callable:
  """transform(*grads) -> grads scaled so that their global L2 norm is at most `clip_norm`
  (upstream examples/cartpole_swingup/swingup_loops.py:93-98 uses clipnorm = 1.0)."""
  def _transform(*grads):
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    scale = torch.clamp(clip_norm / (norm + 1e-300), max=1.0)
    return tuple(g * scale for g in grads)
  return _transform


class GradientDescent:
  def __init__(self, step_limit: int, optimizer_factory: Optional[Callable[[Sequence[torch.Tensor]], torch.optim.Optimizer]] = None,
               callbacks: Optional[List[Callable]] = None, transform: Optional[Callable] = None, show_progress: bool = False,
               ema_const: float = 0.6):
    self.step_limit = step_limit
    self.optimizer_factory = optimizer_factory or (lambda variables: torch.optim.Adam(variables, lr=1e-3))   # keras Adam default
    self.callbacks = list(callbacks or [])
    self.transform = transform
    self.show_progress = show_progress
    self.ema_const = ema_const
    self.history: List[float] = []

  def minimize(self, closure: Callable[[], torch.Tensor], variables: Sequence[torch.Tensor]):
    variables = list(variables)
    for v in variables:
      if not (v.is_leaf and v.requires_grad):
        raise ValueError("GradientDescent.minimize: variables must be leaf tensors with requires_grad=True")
    optimizer = self.optimizer_factory(variables)
    ema_loss = ema_norm = None
    for step in range(self.step_limit):
      loss = closure()
      grads = torch.autograd.grad(loss.sum(), variables, allow_unused=True)     # tape.gradient sums a vector loss
      grads = tuple(torch.zeros_like(v) if g is None else g for g, v in zip(grads, variables))
      if self.transform is not None:
        grads = self.transform(*grads)
      mean_loss = float(loss.detach().mean())
      self.history.append(mean_loss)
      if self.show_progress:
        norm = float(torch.sqrt(sum((g ** 2).sum() for g in grads)))
        if ema_loss is None or self.ema_const == 1:
          ema_loss, ema_norm = mean_loss, norm
        else:
          ema_loss += self.ema_const * (mean_loss - ema_loss)
          ema_norm += self.ema_const * (norm - ema_norm)
        print(f"step {step}: EMA(loss)={ema_loss:.2e}, EMA(norm)={ema_norm:.2e}")
      for v, g in zip(variables, grads):
        v.grad = g
      optimizer.step()
      optimizer.zero_grad(set_to_none=True)
      for callback in self.callbacks:
        callback(step, loss, tuple(zip(grads, variables)))
    return self.history
