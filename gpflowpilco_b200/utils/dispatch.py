"""Minimal multiple dispatcher standing in for gpflow.utilities.Dispatcher (multipledispatch), which is how the
reference exposes its plug-in surface (moment_matching/core.py:35, dynamics/forward_sde.py:17,
utils/kernel_expectation.py:25-31).  Registration keys are tuples of types (or tuples of alternatives); the
most specific match by MRO distance wins."""
from __future__ import annotations

from typing import Callable, Dict, Tuple


class Dispatcher:
  def __init__(self, name: str):
    self.name = name
    self._table: Dict[Tuple, Callable] = {}

  def register(self, *types):
    def deco(fn):
      def expand(ts):
        if not ts:
          yield ()
          return
        head = ts[0] if isinstance(ts[0], tuple) else (ts[0],)
        for h in head:
          for rest in expand(ts[1:]):
            yield (h,) + rest
      for key in expand(types):
        self._table[key] = fn
      return fn
    return deco

  def dispatch(self, *arg_types):
    best, best_score = None, None
    for key, fn in self._table.items():
      if len(key) > len(arg_types):
        continue
      score = 0
      for want, got in zip(key, arg_types):
        if not issubclass(got, want):
          score = None
          break
        score += got.__mro__.index(want) if want in got.__mro__ else len(got.__mro__)
      if score is None:
        continue
      score -= 1000 * len(key)            # longer signatures are more specific
      if best_score is None or score < best_score:
        best, best_score = fn, score
    return best

  def __call__(self, *args, **kwargs):
    nkey = max((len(k) for k in self._table), default=0)
    fn = self.dispatch(*(type(a) for a in args[:nkey]))
    if fn is None:
      raise NotImplementedError(f"{self.name}: no rule for ({', '.join(type(a).__name__ for a in args)})")
    return fn(*args, **kwargs)
