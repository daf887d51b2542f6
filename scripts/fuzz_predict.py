"""Randomised parity sweep of gpp_mm_gp_predict_fwd / _bwd against the float64 oracle (forward values and torch.autograd gradients)
over random (L, M, D, N, whiten, coregionalisation, covariance mode).  Developer tool; the fixed cases live in tests/.
usage: python scripts/fuzz_predict.py [trials] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gpflowpilco_b200 import synthetic
from oracle import gp_models as gm
from oracle import moments as mo
from tests.helpers import DTYPE, cuda_handle, generate_covariance, oracle_svgp

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rng = np.random.default_rng(seed)
dev = lambda x: torch.as_tensor(x, dtype=DTYPE, device="cuda")
worst = 0.0
for t in range(trials):
  L = int(rng.integers(1, 5))
  D = int(rng.integers(1, 9))
  M = int(rng.choice([1, 2, 3, 7, 8, 9, 31, 64, 65, 127, 128, 129, 200, 257]))
  N = int(rng.choice([1, 2, 3, 5, 17, 40]))
  whiten = bool(rng.integers(0, 2))
  coreg = bool(rng.integers(0, 3) == 0)
  full_cov = bool(rng.integers(0, 4) != 0)
  # keep Kuu reasonably conditioned: M points in D dimensions need a box of side ~ M^(1/D) lengthscales (otherwise the comparison
  # measures the conditioning of Kuu^-1, in the oracle as much as in the kernels)
  z_scale = max(1.0, 0.35 * M ** (1.0 / D))
  params = synthetic.random_svgp(L=L, M=M, D=D, seed=int(rng.integers(1 << 30)), whiten=whiten, P=(L + 1 if coreg else None),
                                 z_scale=z_scale)
  # condition number of the worst latent's Kuu (+ the model's jitter): parity is only meaningful up to ~ cond * eps
  cond = 0.0
  for l in range(L):
    Zl = params["Z"][l] / params["lengthscales"][l]
    d2 = ((Zl[:, None, :] - Zl[None, :, :]) ** 2).sum(-1)
    Kuu = params["variance"][l] * np.exp(-0.5 * d2) + 1e-6 * np.eye(M)
    cond = max(cond, float(np.linalg.cond(Kuu)))
  gen = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
  mu = torch.randn(N, D, dtype=DTYPE, generator=gen) * 0.5 * z_scale
  cov = generate_covariance(D, (N,), 0.3, gen)
  P = params["W"].shape[0] if params.get("W") is not None else L
  f1_bar = torch.randn(N, P, dtype=DTYPE, generator=gen)
  Sff_bar = torch.randn(N, P, P, dtype=DTYPE, generator=gen)
  cross_bar = torch.randn(N, D, P, dtype=DTYPE, generator=gen)
  Sff_bar_eff = Sff_bar if full_cov else torch.diag_embed(torch.diagonal(Sff_bar, dim1=-2, dim2=-1))
  model = oracle_svgp(params)
  mu_r, cov_r = mu.clone().requires_grad_(True), cov.clone().requires_grad_(True)
  match = gm.mm_svgp_mo(mo.GaussianMoments(mu_r, cov_r, True), model, model_uncertainty=True, full_output_cov=full_cov)
  Sff = match.y.covariance()
  if Sff.dim() == 2:
    Sff = torch.diag_embed(Sff)
  s = (match.y.mean() * f1_bar).sum() + (Sff * Sff_bar_eff).sum() + (match.cross[0] * cross_bar).sum()
  gmu, gcov = torch.autograd.grad(s, (mu_r, cov_r))
  gcov = 0.5 * (gcov + gcov.transpose(-1, -2))
  h = cuda_handle(params)
  f1, Sff_c, cross = h.predict(dev(mu), dev(cov), full_output_cov=full_cov)
  m_bar, S_bar = h.predict_bwd(dev(mu), dev(cov), dev(f1_bar), dev(Sff_bar), dev(cross_bar), full_output_cov=full_cov)
  errs = []
  for got, ref in ((f1, match.y.mean()), (Sff_c, Sff), (cross, match.cross[0]), (m_bar, gmu), (S_bar, gcov)):
    ref = ref.detach()
    got = got.cpu()
    if got.shape != ref.shape and got.dim() == 3 and ref.dim() == 3 and not full_cov:
      got = torch.diag_embed(torch.diagonal(got, dim1=-2, dim2=-1))
    errs.append(float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-300)))
  # the same quantities from the oracle's O(M^2) re-association (the algebraic form the kernels use): separates a kernel bug from
  # the sensitivity of the two algebraic forms to cond(Kuu)
  with torch.no_grad():
    alt = gm.mm_sparse_reassociated(mo.GaussianMoments(mu, cov, True), model, model_uncertainty=True)
  Sff_alt = alt.y.covariance()
  if not full_cov:
    Sff_alt = torch.diag_embed(torch.diagonal(Sff_alt, dim1=-2, dim2=-1))
  got = Sff_c.cpu()
  if not full_cov:
    got = torch.diag_embed(torch.diagonal(got, dim1=-2, dim2=-1))
  err_alt = float((got - Sff_alt).abs().max() / max(float(Sff_alt.abs().max()), 1e-300))
  Sff_ref = Sff.detach()
  oracle_gap = float((Sff_ref - Sff_alt).abs().max() / max(float(Sff_alt.abs().max()), 1e-300))   # the two oracle forms against each other
  # 1e-6 (north star) unless the problem itself is ill-conditioned: Sff = sum C o Q cancels from |C| (up to 1e12 for crowded,
  # un-whitened inducing points) down to O(1), so last-bit differences in Q (ours is good to ~|log Q| 4e-16) show up multiplied by
  # |C|; the oracle's two algebraic forms share one Q and therefore agree with each other much better than with anyone else
  with torch.no_grad():
    maxC = float(gm.sparse_weights(model, True)[1].abs().max())
  tol = max(1e-6, 1e-11 * cond, 30.0 * oracle_gap, 1e-13 * maxC / max(1.0, float(Sff_alt.abs().max())))
  ok = max(errs) < tol
  worst = max(worst, max(errs) / tol)
  flag = "" if ok else "   <-- FAIL"
  print(f"trial {t}: L={L} M={M} D={D} N={N} whiten={whiten} coreg={coreg} full_cov={full_cov} cond(Kuu)={cond:.1e}: "
        f"f1 {errs[0]:.1e} Sff {errs[1]:.1e} cross {errs[2]:.1e} m_bar {errs[3]:.1e} S_bar {errs[4]:.1e}; Sff vs re-associated oracle {err_alt:.1e} (oracle forms differ by {oracle_gap:.1e}){flag}")
print(f"worst error / tolerance over {trials} trials: {worst:.2e}  (tolerance = max(1e-6, 1e-11 cond(Kuu), 30 x disagreement of the oracle's two forms, 1e-13 max|C| / max|Sff|))")
sys.exit(0 if worst < 1.0 else 1)
