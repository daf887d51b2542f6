"""Oracle: Philox4x32-10 counter-based generator (Salmon et al., SC'11) in numpy.

TEST INFRASTRUCTURE — see oracle/__init__.py.

The reference draws its pathwise randomness with tf.random (gpflow_pilco/loops/pilco.py:281-284,300-303 via
gpflow_sampling); a sharded GPU implementation needs draws that depend only on (seed, stream, logical index),
so this repo fixes a counter-based contract:

  element e of stream t  ->  counter (lo32(e>>1), hi32(e>>1), t, 0), key (lo32(seed), hi32(seed))
  the 4x32 output words give two 53-bit uniforms  u = (((x>>5)<<26 | (y>>6)) + 0.5) * 2^-53
  normals : Box-Muller, element e takes r*cos(theta) if e is even else r*sin(theta)
  uniforms: element e takes u1 of counter e (no pairing)

The raw 32-bit words must match the CUDA implementation bit for bit ("bit-exact sample indexing").
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

STREAM_OMEGA, STREAM_PHASE, STREAM_PRIOR_W, STREAM_U_EPS, STREAM_UPDATE_XI, STREAM_X0 = range(6)


def philox4x32(counter_index: np.ndarray, stream: int, seed: int) -> np.ndarray:
  """counter_index: uint64 array [...]; returns uint32 array [...,4]."""
  idx = np.asarray(counter_index, dtype=np.uint64)
  c0 = idx & MASK
  c1 = idx >> np.uint64(32)
  c2 = np.full_like(idx, np.uint64(stream))
  c3 = np.zeros_like(idx)
  k0 = seed & 0xFFFFFFFF
  k1 = (seed >> 32) & 0xFFFFFFFF
  for _ in range(10):
    p0 = M0 * c0
    p1 = M1 * c2
    hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
    hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
    c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
    k0 = (k0 + W0) & 0xFFFFFFFF
    k1 = (k1 + W1) & 0xFFFFFFFF
  return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def _u53(x: np.ndarray, y: np.ndarray) -> np.ndarray:
  hi = (x.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)
  lo = y.astype(np.uint64) >> np.uint64(6)
  return ((hi | lo).astype(np.float64) + 0.5) * (2.0 ** -53)


def normal(elements: np.ndarray, stream: int, seed: int) -> np.ndarray:
  """Standard normals for logical element indices ``elements`` (any shape, uint64)."""
  e = np.asarray(elements, dtype=np.uint64)
  words = philox4x32(e >> np.uint64(1), stream, seed)
  u1 = _u53(words[..., 0], words[..., 1])
  u2 = _u53(words[..., 2], words[..., 3])
  r = np.sqrt(-2.0 * np.log(u1))
  th = 2.0 * np.pi * u2
  return np.where((e & np.uint64(1)) == 0, r * np.cos(th), r * np.sin(th))


def uniform(elements: np.ndarray, stream: int, seed: int) -> np.ndarray:
  e = np.asarray(elements, dtype=np.uint64)
  words = philox4x32(e, stream, seed)
  return _u53(words[..., 0], words[..., 1])
