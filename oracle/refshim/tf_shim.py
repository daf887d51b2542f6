"""numpy-backed stand-in for the TensorFlow calls the upstream hot path makes.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Only semantics are reproduced (shapes, broadcasting, argument names); tensors are float64 numpy arrays.  This is OUR
code: it exists so that the unmodified upstream sources under /root/reference can be executed in this container, where
the real tensorflow is not installable.  Every function follows the documented behaviour of the TF op of the same name.
"""
from __future__ import annotations

import types

import numpy as np
from scipy import linalg as sla
from scipy import special as sps


class TensorShape(tuple):
  @property
  def ndims(self):
    return len(self)

  def as_list(self):
    return list(self)


class Tensor(np.ndarray):
  __array_priority__ = 5.0

  @property
  def shape(self):
    return TensorShape(np.ndarray.shape.__get__(self))

  def numpy(self):
    return np.asarray(self)


class Variable(Tensor):
  pass


def T(x, dtype=None):
  if isinstance(x, LinearOperator):
    return x
  return np.asarray(x, dtype=dtype if dtype is not None else None).view(Tensor)


def _dense(x):
  return x.to_dense() if isinstance(x, LinearOperator) else np.asarray(x)


# ---- linear operators ---------------------------------------------------------------------------------------
class LinearOperator:
  __array_priority__ = 100.0

  def __init__(self, dense):
    self._d = np.asarray(dense, dtype=np.float64)

  @property
  def shape(self):
    return TensorShape(self._d.shape)

  @property
  def dtype(self):
    return self._d.dtype

  def to_dense(self):
    return T(self._d)

  def matmul(self, other):
    if isinstance(other, LinearOperator):
      return LinearOperatorComposition((self, other))
    return T(self._d @ np.asarray(other))

  __matmul__ = matmul

  def __rmatmul__(self, other):
    return T(np.asarray(other) @ self._d)

  def solve(self, rhs):
    return T(np.linalg.solve(self._d, _dense(rhs)))

  def diag_part(self):
    return T(np.diagonal(self._d, axis1=-2, axis2=-1))


class LinearOperatorIdentity(LinearOperator):
  def __init__(self, num_rows, dtype=None, **kw):
    super().__init__(np.eye(int(num_rows)))


class LinearOperatorScaledIdentity(LinearOperator):
  def __init__(self, num_rows, multiplier, **kw):
    m = np.asarray(multiplier, dtype=np.float64)
    super().__init__(m[..., None, None] * np.eye(int(num_rows)))


class LinearOperatorDiag(LinearOperator):
  def __init__(self, diag, **kw):
    d = np.asarray(diag, dtype=np.float64)
    super().__init__(d[..., :, None] * np.eye(d.shape[-1]))


class LinearOperatorComposition(LinearOperator):
  def __init__(self, operators, **kw):
    out = _dense(operators[0])
    for op in operators[1:]:
      out = out @ _dense(op)
    super().__init__(out)


class LinearOperatorLowRankUpdate(LinearOperator):
  def __init__(self, base_operator, u, diag_update=None, v=None, **kw):
    u = np.asarray(u)
    v = u if v is None else np.asarray(v)
    d = np.ones(u.shape[-1]) if diag_update is None else np.asarray(diag_update)
    super().__init__(_dense(base_operator) + (u * d) @ np.swapaxes(v, -1, -2))


# ---- linalg -------------------------------------------------------------------------------------------------
def _tri_solve(L, B, lower=True, adjoint=False):
  L, B = np.asarray(L), np.asarray(B)
  batch = np.broadcast_shapes(L.shape[:-2], B.shape[:-2])
  Lb = np.broadcast_to(L, batch + L.shape[-2:]).reshape((-1,) + L.shape[-2:])
  Bb = np.broadcast_to(B, batch + B.shape[-2:]).reshape((-1,) + B.shape[-2:])
  out = np.empty_like(Bb)
  for i in range(Lb.shape[0]):
    out[i] = sla.solve_triangular(Lb[i], Bb[i], lower=lower, trans=1 if adjoint else 0)
  return T(out.reshape(batch + B.shape[-2:]))


def _matmul(a, b, transpose_a=False, transpose_b=False, adjoint_a=False, adjoint_b=False, **kw):
  a, b = _dense(a), _dense(b)
  if transpose_a or adjoint_a:
    a = np.swapaxes(a, -1, -2)
  if transpose_b or adjoint_b:
    b = np.swapaxes(b, -1, -2)
  return T(a @ b)


def _matvec(a, b, adjoint_a=False, transpose_a=False, **kw):
  a = _dense(a)
  if adjoint_a or transpose_a:
    a = np.swapaxes(a, -1, -2)
  return T((a @ np.asarray(b)[..., None])[..., 0])


def _diag(v, **kw):
  v = np.asarray(v)
  return T(v[..., :, None] * np.eye(v.shape[-1]))


def _diag_part(x, **kw):
  return T(np.diagonal(_dense(x), axis1=-2, axis2=-1).copy())


def _set_diag(x, d, **kw):
  out = np.array(_dense(x), dtype=np.float64, copy=True)
  idx = np.arange(out.shape[-1])
  out[..., idx, idx] = np.asarray(d)
  return T(out)


def _band_part(x, lo, hi):
  x = np.asarray(x)
  assert lo == -1 and hi == 0
  return T(np.tril(x))


def _cholesky_solve(chol, rhs, **kw):
  y = _tri_solve(chol, _dense(rhs), lower=True)
  return _tri_solve(chol, y, lower=True, adjoint=True)


def _l2_normalize(x, axis=None, **kw):
  x = np.asarray(x)
  return T(x / np.sqrt(np.maximum((x * x).sum(axis=axis, keepdims=True), 1e-12)))


linalg = types.SimpleNamespace(
    LinearOperator=LinearOperator, LinearOperatorIdentity=LinearOperatorIdentity, LinearOperatorDiag=LinearOperatorDiag,
    LinearOperatorScaledIdentity=LinearOperatorScaledIdentity, LinearOperatorComposition=LinearOperatorComposition,
    LinearOperatorLowRankUpdate=LinearOperatorLowRankUpdate,
    cholesky=lambda x, **kw: T(np.linalg.cholesky(_dense(x))),
    triangular_solve=lambda matrix, rhs, lower=True, adjoint=False, **kw: _tri_solve(matrix, rhs, lower, adjoint),
    cholesky_solve=_cholesky_solve, matmul=_matmul, matvec=_matvec,
    adjoint=lambda x, **kw: T(np.swapaxes(_dense(x), -1, -2)),
    diag=_diag, diag_part=_diag_part, set_diag=_set_diag, band_part=_band_part,
    trace=lambda x, **kw: T(np.trace(_dense(x), axis1=-2, axis2=-1)),
    inv=lambda x, **kw: T(np.linalg.inv(_dense(x))), det=lambda x, **kw: T(np.linalg.det(_dense(x))),
    svd=lambda x, full_matrices=False, **kw: (lambda u, s, vh: (T(s), T(u), T(np.swapaxes(vh, -1, -2))))(*np.linalg.svd(np.asarray(x), full_matrices=full_matrices)),
    l2_normalize=_l2_normalize, eye=lambda n, dtype=None, **kw: T(np.eye(int(n))),
)


# ---- math ---------------------------------------------------------------------------------------------------
def _un(fn):
  return lambda x, *a, **kw: T(fn(np.asarray(x)))


def add(a, b, **kw):
  return T(np.asarray(a) + np.asarray(b))


def subtract(a, b, **kw):
  return T(np.asarray(a) - np.asarray(b))


def multiply(a, b, **kw):
  return T(np.asarray(a) * np.asarray(b))


math = types.SimpleNamespace(
    add=add, subtract=subtract, multiply=multiply, cos=_un(np.cos), sin=_un(np.sin), exp=_un(np.exp), log=_un(np.log),
    square=_un(np.square), sqrt=_un(np.sqrt), rsqrt=_un(lambda x: 1.0 / np.sqrt(x)), reciprocal=_un(lambda x: 1.0 / x),
    erfc=_un(sps.erfc), erf=_un(sps.erf), l2_normalize=_l2_normalize, abs=_un(np.abs), asin=_un(np.arcsin),
)


def _reduce(fn):
  def f(x, axis=None, keepdims=False, **kw):
    return T(fn(_dense(x), axis=tuple(axis) if isinstance(axis, list) else axis, keepdims=keepdims))
  return f


def concat(values, axis, **kw):
  return T(np.concatenate([_dense(v) for v in values], axis=axis))


def stack(values, axis=0, **kw):
  return T(np.stack([_dense(v) for v in values], axis=axis))


def gather(params, indices, axis=None, **kw):
  return T(np.take(_dense(params), np.asarray(indices, dtype=np.int64), axis=0 if axis is None else axis))


def cast(x, dtype=None, **kw):
  return T(np.asarray(x, dtype=np.float64))


def convert_to_tensor(value=None, dtype=None, **kw):
  return T(np.asarray(value, dtype=np.float64))


def fill(dims, value, **kw):
  return T(np.full(tuple(int(d) for d in np.atleast_1d(dims)), np.asarray(value)))


def transpose(x, perm=None, **kw):
  return T(np.transpose(_dense(x), perm))


def tile(x, multiples, **kw):
  return T(np.tile(np.asarray(x), tuple(int(m) for m in multiples)))


def where(cond, x=None, y=None, **kw):
  return T(np.where(np.asarray(cond), np.asarray(x), np.asarray(y)))


def foldl(fn, elems, initializer=None, **kw):
  state = initializer
  n = len(elems[0]) if isinstance(elems, (tuple, list)) else len(elems)
  for i in range(n):
    e = tuple(x[i] for x in elems) if isinstance(elems, (tuple, list)) else elems[i]
    state = fn(state, e)
  return state


def scan(fn, elems, initializer=None, **kw):
  state, outs = initializer, []
  n = len(elems[0]) if isinstance(elems, (tuple, list)) else len(elems)
  for i in range(n):
    e = tuple(x[i] for x in elems) if isinstance(elems, (tuple, list)) else elems[i]
    state = fn(state, e)
    outs.append(state)
  return outs


class Module:
  def __init__(self, name=None):
    self._name = name
    self._name_scope = None


def function(fn=None, **kw):
  return fn if fn is not None else (lambda f: f)


_rng = np.random.default_rng(0)
random = types.SimpleNamespace(
    normal=lambda shape, dtype=None, **kw: T(_rng.standard_normal(tuple(shape))),
    uniform=lambda shape, minval=0.0, maxval=1.0, dtype=None, **kw: T(_rng.uniform(minval, maxval, tuple(shape))),
    set_seed=lambda s: None,
)


def build_module():
  tf = types.ModuleType("tensorflow")
  tf.Tensor, tf.Variable, tf.Module = Tensor, Variable, Module
  tf.linalg, tf.math, tf.random = linalg, math, random
  tf.float64 = np.float64
  tf.add, tf.subtract, tf.multiply = add, subtract, multiply      # same objects as tf.math.* (upstream registers tf.math.add)
  for name, fn in dict(identity=lambda x, **kw: T(np.array(x, copy=True)), concat=concat, stack=stack, gather=gather, cast=cast,
                       convert_to_tensor=convert_to_tensor, fill=fill, transpose=transpose, tile=tile, where=where, foldl=foldl,
                       scan=scan, function=function, matmul=_matmul,
                       expand_dims=lambda x, axis, **kw: T(np.expand_dims(_dense(x), axis)),
                       squeeze=lambda x, axis=None, **kw: T(np.squeeze(_dense(x), axis=axis)),
                       reduce_sum=_reduce(np.sum), reduce_prod=_reduce(np.prod), reduce_mean=_reduce(np.mean),
                       reduce_all=lambda x, **kw: bool(np.all(np.asarray(x))),
                       shape=lambda x, **kw: TensorShape(np.shape(_dense(x))), reshape=lambda x, s, **kw: T(np.reshape(_dense(x), tuple(int(v) for v in s))),
                       sqrt=_un(np.sqrt), exp=_un(np.exp), square=_un(np.square), abs=_un(np.abs),
                       eye=lambda n, dtype=None, **kw: T(np.eye(int(n))), zeros=lambda s, dtype=None, **kw: T(np.zeros(tuple(s))),
                       zeros_like=lambda x, **kw: T(np.zeros_like(np.asarray(x))), ones=lambda s, dtype=None, **kw: T(np.ones(tuple(s))),
                       einsum=lambda eq, *a, **kw: T(np.einsum(eq, *[np.asarray(v) for v in a])),
                       clip_by_value=lambda x, lo, hi, **kw: T(np.clip(np.asarray(x), lo, hi)),
                       logical_and=np.logical_and, logical_or=np.logical_or, logical_not=np.logical_not, equal=np.equal,
                       constant=lambda v, dtype=None, **kw: T(np.asarray(v, dtype=np.float64 if dtype is None else dtype)),
                       sign=_un(np.sign), maximum=lambda a, b, **kw: T(np.maximum(np.asarray(a), np.asarray(b))),
                       minimum=lambda a, b, **kw: T(np.minimum(np.asarray(a), np.asarray(b))),
                       ).items():
    setattr(tf, name, fn)
  return tf
