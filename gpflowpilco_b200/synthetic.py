"""Seeded synthetic workloads of BASELINE.json's configs (host-side numpy; no GPU, no oracle dependency).

Shapes and distributions follow SURVEY.md §8(d); upstream sources of the constants are cited per function.
All builders return plain dicts of float64 numpy arrays so that the CUDA path, the oracle and the fixtures
consume identical inputs.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np


def generate_covariance(rng: np.random.Generator, ndims: int, batch: int, scale: float) -> np.ndarray:
  """Random covariances: eigenvalues -log U, Haar-ish orthogonal basis, rescaled to scale^2 x correlation
  (same construction as upstream tests/utils.py:99-121)."""
  eig = -np.log(rng.random((batch, 1, ndims)))
  U = np.linalg.svd(rng.standard_normal((batch, ndims, ndims)))[0]
  sq = np.sqrt(eig) * U
  cov = sq @ np.swapaxes(sq, -1, -2)
  istd = 1.0 / np.sqrt(np.einsum("bii->bi", cov))
  return (scale ** 2) * cov * istd[:, :, None] * istd[:, None, :]


def se_kernel(A: np.ndarray, B: np.ndarray, ell: np.ndarray, var: float) -> np.ndarray:
  d = (A / ell)[:, None, :] - (B / ell)[None, :, :]
  return var * np.exp(-0.5 * (d * d).sum(-1))


def config2_batched_mm_predict(N: int = 8192, M: int = 1000, D: int = 6, E: int = 4, seed: int = 0) -> Dict[str, np.ndarray]:
  """BASELINE config #2: N Gaussian inputs through E independent exact SE-ARD GPs on M training points
  (maths of upstream moment_matching/models.py:44-111 per output)."""
  rng = np.random.default_rng(seed)
  X = rng.random((M, D))
  ell = np.exp(rng.uniform(math.log(0.3), math.log(3.0), size=(E, D)))
  var = np.full(E, 0.89 ** 2)
  noise = 1e-2
  Y = np.empty((M, E))
  for e in range(E):
    K = se_kernel(X, X, ell[e], var[e]) + 1e-10 * np.eye(M)
    Y[:, e] = np.linalg.cholesky(K) @ rng.standard_normal(M) + math.sqrt(noise) * rng.standard_normal(M)
  mean_const = 0.1 * rng.standard_normal(E)
  mu = rng.random((N, D))
  cov = generate_covariance(rng, D, N, 0.1)
  return dict(X=X, Y=Y, lengthscales=ell, variance=var, noise_variance=np.full(E, noise), mean_const=mean_const,
              mu=mu, cov=cov)


def config3_psi2_stress(N: int = 1024, M: int = 2048, D: int = 8, seed: int = 0) -> Dict[str, np.ndarray]:
  """BASELINE config #3: Psi2 with two kernels / two inducing sets (the branch upstream tests/test_kernel_expectation.py:50-93
  exercises), inducing points half near the bulk of the inputs, half uniform (:63-66)."""
  rng = np.random.default_rng(seed)
  mu = rng.standard_normal((N, D))
  cov = generate_covariance(rng, D, N, 0.1)

  def inducing():
    return np.concatenate([math.sqrt(0.1) * rng.standard_normal((M // 2, D)), rng.random((M - M // 2, D))], 0)

  Z1, Z2 = inducing(), inducing()
  ell1 = np.exp(rng.uniform(math.log(0.1), math.log(10.0), size=D))
  ell2 = np.exp(rng.uniform(math.log(0.1), math.log(10.0), size=D))
  return dict(mu=mu, cov=cov, Z1=Z1, Z2=Z2, lengthscales1=ell1, lengthscales2=ell2, variance1=0.89 ** 2, variance2=0.89 ** 2)


def random_svgp(L: int, M: int, D: int, seed: int = 0, whiten: bool = True, P: int | None = None,
                ell_range=(0.5, 2.0), z_scale: float = 1.0) -> Dict[str, np.ndarray]:
  """A generic well-conditioned multi-output SVGP parameter set for parity tests."""
  rng = np.random.default_rng(seed)
  Z = z_scale * rng.standard_normal((L, M, D))
  ell = np.exp(rng.uniform(math.log(ell_range[0]), math.log(ell_range[1]), size=(L, D)))
  var = 0.5 + rng.random(L)
  q_mu = rng.standard_normal((M, L))
  A = 0.3 * rng.standard_normal((L, M, M)) / math.sqrt(M)
  q_sqrt = np.tril(A) + 0.2 * np.eye(M)[None]
  out = dict(Z=Z, lengthscales=ell, variance=var, q_mu=q_mu, q_sqrt=q_sqrt, whiten=whiten,
             mean_const=rng.standard_normal(L if P is None else P))
  if P is not None:
    W = rng.random((P, L))
    out["W"] = W / np.linalg.norm(W, axis=-1, keepdims=True)
  return out


# ---------------------------------------------------------------------------------------------------------
# config #1: cart-pole swing-up models (upstream examples/cartpole_swingup)
# ---------------------------------------------------------------------------------------------------------
def cartpole_ode(state: np.ndarray, force: float) -> np.ndarray:
  """Time derivative of (x, theta, dx, dtheta) — same physics as upstream gpflow_pilco/envs/cart_pole.py:55-85
  (cart 0.5 kg, friction 0.1; pole 0.5 kg, length 0.5 m; g = 9.81; force clipped to [-10, 10])."""
  g, h, m, M, fr = 9.81, 0.5, 0.5, 0.5, 0.1
  x, a, dx, da = state
  f = float(np.clip(force, -10.0, 10.0))
  s, c = math.sin(a), math.cos(a)
  drag = -fr * dx
  ddx = (f + drag + 0.5 * s * m * (h * da * da + 1.5 * g * c)) / ((M + m) - 0.75 * m * c * c)
  dda = (c * (f + drag + 0.5 * s * m * h * da * da) + (M + m) * g * s) / (2.0 / 3.0 * h * (M + m) - 0.5 * m * h * c * c)
  return np.array([dx, da, ddx, dda])


def cartpole_step(state: np.ndarray, force: float, dt: float = 0.1, substeps: int = 20) -> np.ndarray:
  hh = dt / substeps
  s = state.copy()
  for _ in range(substeps):     # classical RK4 (upstream integrates with scipy solve_ivp; only used to make data)
    k1 = cartpole_ode(s, force)
    k2 = cartpole_ode(s + 0.5 * hh * k1, force)
    k3 = cartpole_ode(s + 0.5 * hh * k2, force)
    k4 = cartpole_ode(s + hh * k3, force)
    s = s + hh / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
  return s


def encode_cartpole(x: np.ndarray) -> np.ndarray:
  """TrigonometricEncoder(active_dims=(1,)): [sin th, cos th, x, dx, dth] (upstream components.py:49-56, swingup_loops.py:44)."""
  return np.stack([np.sin(x[..., 1]), np.cos(x[..., 1]), x[..., 0], x[..., 2], x[..., 3]], -1)


def _kmeans(X: np.ndarray, k: int, rng: np.random.Generator, iters: int = 25) -> np.ndarray:
  C = X[rng.choice(len(X), size=k, replace=False)].copy()
  for _ in range(iters):
    d = ((X[:, None, :] - C[None, :, :]) ** 2).sum(-1)
    lab = d.argmin(1)
    for j in range(k):
      pts = X[lab == j]
      if len(pts):
        C[j] = pts.mean(0)
  return C


def _median_lengthscale(X: np.ndarray) -> float:
  d = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))[np.triu_indices(len(X), 1)]
  return float(np.clip(math.sqrt(0.5) * np.median(d), 0.011, 90.0))     # upstream models/initializers.py:29-50


def config1_cartpole(seed: int = 0, M: int = 256, Mp: int = 30, episodes: int = 9, steps: int = 30,
                     ard_scale: bool = True) -> Dict[str, object]:
  """Dynamics SVGP (4 latents, M inducing, 6 inputs), RBF policy (Mp centres, 5 inputs), task constants of the
  upstream cart-pole example: H = 30 (experiment.py:121-122), m0 = (0, pi, 0, 0), S0 = 0.01 I (experiment.py:131-135),
  squash (20 - 1e-5)(Phi(f) - 1/2) (swingup_loops.py:87-90), cost W (swingup_loops.py:47-54).
  q(u) is the closed-form optimum for a Gaussian likelihood with noise 1e-2 var(y) (no training in this repo)."""
  rng = np.random.default_rng(seed)
  Xs, Us, dXs = [], [], []
  for _ in range(episodes):
    s = np.array([0.0, math.pi, 0.0, 0.0]) + 0.1 * rng.standard_normal(4)
    for _ in range(steps):
      u = rng.uniform(-10.0, 10.0)
      s2 = cartpole_step(s, u)
      Xs.append(s); Us.append(u); dXs.append(s2 - s)
      s = s2
  X = np.array(Xs); U = np.array(Us)[:, None]; dX = np.array(dXs)
  ZU = np.concatenate([encode_cartpole(X), U], -1)                       # [270, 6]
  n, D = ZU.shape
  M = min(M, n)
  ell0 = _median_lengthscale(ZU)
  L = 4
  if ard_scale:    # per-dimension scale so that no input dimension is ignored (stands in for the ELBO fit)
    ell = np.tile(np.maximum(ZU.std(0) * 2.0, 0.05)[None, :], (L, 1))
  else:
    ell = np.full((L, D), ell0)
  var = np.maximum(dX.var(0), 1e-4)
  noise = 1e-2 * var
  Z = _kmeans(ZU, M, rng)
  q_mu = np.zeros((M, L)); q_sqrt = np.zeros((L, M, M))
  for l in range(L):
    Kuu = se_kernel(Z, Z, ell[l], var[l]) + 1e-6 * np.eye(M)
    Kuf = se_kernel(Z, ZU, ell[l], var[l])
    Lu = np.linalg.cholesky(Kuu)
    A = np.linalg.solve(Lu, Kuf) / math.sqrt(noise[l])                   # whitened: Sigma_v = (I + A A^T)^-1
    B = np.eye(M) + A @ A.T
    LB = np.linalg.cholesky(B)
    c = np.linalg.solve(LB, A @ dX[:, l]) / math.sqrt(noise[l])
    q_mu[:, l] = np.linalg.solve(LB.T, c)
    Sv = np.linalg.inv(B)
    q_sqrt[l] = np.linalg.cholesky(0.5 * (Sv + Sv.T) + 1e-12 * np.eye(M))
  dyn = dict(Z=np.tile(Z[None], (L, 1, 1)), lengthscales=ell, variance=var, q_mu=q_mu, q_sqrt=q_sqrt, whiten=True,
             mean_const=np.zeros(L))
  E = encode_cartpole(X)
  pol = dict(Z=_kmeans(E, Mp, rng)[None], lengthscales=np.full((1, 5), _median_lengthscale(E)), variance=np.ones(1),
             q_mu=1e-3 * rng.standard_normal((Mp, 1)), q_sqrt=np.tile(np.eye(Mp)[None], (1, 1, 1)), whiten=True,
             mean_const=np.zeros(1))
  h = 0.5
  W = 16.0 * np.array([[h * h, 0, -h, 0, 0], [0, h * h, 0, 0, 0], [-h, 0, 1, 0, 0], [0, 0, 0, 0, 0], [0, 0, 0, 0, 0]], dtype=np.float64)
  return dict(dynamics=dyn, policy=pol, m0=np.array([[0.0, math.pi, 0.0, 0.0]]), S0=0.01 * np.eye(4)[None],
              target=np.array([0.0, 1.0, 0.0, 0.0, 0.0]), W=W, squash_scale=20.0 - 1e-5, squash_shift=-0.5, horizon=30,
              active_dims=(1,), data=dict(ZU=ZU, dX=dX))
