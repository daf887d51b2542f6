"""GPU parity against the golden vectors produced by the UNMODIFIED upstream sources (tests/golden/*.npz): the CUDA path
through the reference-facing API vs upstream's own outputs.  Tolerance 1e-6 relative (north star), most much tighter."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import DTYPE, scaled_close

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(x):
  return torch.as_tensor(np.asarray(x), dtype=DTYPE, device="cuda")


def load(name):
  return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("tag", ["d2", "d6"])
def test_psi_golden(tag):
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.utils.kernel_expectation import Gaussian, kernel_expectation
  g = load("psi.npz")
  p = Gaussian(_dev(g[f"{tag}_mu"]), _dev(g[f"{tag}_cov"]))
  k1 = M.SquaredExponential(_dev(g[f"{tag}_var1"]), _dev(g[f"{tag}_ell1"]))
  k2 = M.SquaredExponential(_dev(g[f"{tag}_var2"]), _dev(g[f"{tag}_ell2"]))
  Z1, Z2 = M.InducingPoints(_dev(g[f"{tag}_Z1"])), M.InducingPoints(_dev(g[f"{tag}_Z2"]))
  scaled_close(kernel_expectation(p, (k1, Z1)), torch.as_tensor(g[f"{tag}_eKxz"]), 1e-12, "eKxz")
  scaled_close(kernel_expectation(p, (k1, Z1), (k1, Z1)), torch.as_tensor(g[f"{tag}_same"]), 1e-11, "same")
  scaled_close(kernel_expectation(p, (k1, Z1), (k1, Z2)), torch.as_tensor(g[f"{tag}_samekern"]), 1e-11, "same kernel")
  scaled_close(kernel_expectation(p, (k1, Z1), (k2, Z2)), torch.as_tensor(g[f"{tag}_generic"]), 1e-11, "generic")


def _check(mm, mdiag, g, tag):
  scaled_close(mm.y.mean(), torch.as_tensor(g[f"{tag}_mean"]), 1e-8, f"{tag} mean")
  scaled_close(mm.y.covariance(), torch.as_tensor(g[f"{tag}_cov"]), 1e-6, f"{tag} cov")
  scaled_close(mm.cross[0], torch.as_tensor(g[f"{tag}_cross_pre"]), 1e-8, f"{tag} cross (pre-inverted)")
  scaled_close(mm.cross_covariance(), torch.as_tensor(g[f"{tag}_cross"]), 1e-8, f"{tag} cross")
  if mdiag is not None:
    scaled_close(torch.diagonal(mdiag.y.covariance(), dim1=-2, dim2=-1), torch.as_tensor(g[f"{tag}_diag_cov"]), 1e-6, f"{tag} diag")


def test_mm_models_golden():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = load("mm_models.npz")
  x = GaussianMoments((_dev(g["mx"]), _dev(g["Sxx"])), True)
  gpr = M.GPR((_dev(g["gpr_X"]), _dev(g["gpr_Y"])), M.SquaredExponential(_dev(g["gpr_var"]), _dev(g["gpr_ell"])),
              mean_function=M.Constant(_dev(g["gpr_c"])), noise_variance=float(g["gpr_noise"]))
  _check(moment_matching(x, gpr), moment_matching(x, gpr, full_output_cov=False), g, "gpr")
  so = M.SVGP(M.SquaredExponential(_dev(g["so_var"]), _dev(g["so_ell"])), M.InducingPoints(_dev(g["so_Z"])), _dev(g["so_q_mu"]),
              _dev(g["so_q_sqrt"]), whiten=False, mean_function=M.Constant(_dev(g["so_c"])))
  _check(moment_matching(x, so), moment_matching(x, so, full_output_cov=False), g, "so")
  _check(moment_matching(x, so, model_uncertainty=False), None, g, "so_nounc")
  co = M.SVGP(M.LinearCoregionalization([M.SquaredExponential(_dev(g["co_var"][l]), _dev(g["co_ell"][l])) for l in range(2)], _dev(g["co_W"])),
              M.SeparateIndependentInducingVariables([M.InducingPoints(_dev(g["co_Z"][l])) for l in range(2)]), _dev(g["co_q_mu"]),
              _dev(g["co_q_sqrt"]), whiten=False, mean_function=M.Constant(_dev(g["co_c"])))
  _check(moment_matching(x, co), moment_matching(x, co, full_output_cov=False), g, "co")
  x6 = GaussianMoments((_dev(g["si_mx"]), _dev(g["si_Sxx"])), True)
  si = M.SVGP(M.SeparateIndependent([M.SquaredExponential(_dev(g["si_var"][l]), _dev(g["si_ell"][l])) for l in range(4)]),
              M.SeparateIndependentInducingVariables([M.InducingPoints(_dev(g["si_Z"][l])) for l in range(4)]), _dev(g["si_q_mu"]),
              _dev(g["si_q_sqrt"]), whiten=True, mean_function=M.Constant(_dev(g["si_c"])))
  _check(moment_matching(x6, si, jitter=1e-8), moment_matching(x6, si, full_output_cov=False, jitter=1e-8), g, "si")
  _check(moment_matching(x6, si, model_uncertainty=False), None, g, "si_nounc")


def test_rules_golden():
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.components import GaussianObjective, TrigonometricEncoder, sincos
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = load("rules.npz")
  x = GaussianMoments((_dev(g["m"]), _dev(g["S"])), True)
  for name, fn in (("sincos", sincos), ("sin", torch.sin), ("cos", torch.cos)):
    mm = moment_matching(x, fn)
    scaled_close(mm.y.mean(), torch.as_tensor(g[f"{name}_mean"]), 1e-13, name)
    scaled_close(mm.y.covariance(), torch.as_tensor(g[f"{name}_cov"]), 1e-12, name)
    scaled_close(mm.cross_covariance(), torch.as_tensor(g[f"{name}_cross"]), 1e-12, name)
  for tag, active in (("enc1", (1,)), ("enc23", (2, 3))):
    mm = moment_matching(x, TrigonometricEncoder(active))
    scaled_close(mm.y.mean(), torch.as_tensor(g[f"{tag}_mean"]), 1e-13, tag)
    scaled_close(mm.y.covariance(), torch.as_tensor(g[f"{tag}_cov"]), 1e-12, tag)
    scaled_close(mm.cross_covariance(), torch.as_tensor(g[f"{tag}_cross"]), 1e-12, tag)
  link = M.BijectorChain([M.Scale(20 - 1e-5), M.Shift(-0.5), M.NormalCDF()])
  sq = moment_matching(GaussianMoments((_dev(g["sq_m"])[:, None], _dev(g["sq_v"])[:, None, None]), True), link)
  scaled_close(sq.y.mean()[:, 0], torch.as_tensor(g["sq_mean"]), 1e-13, "squash mean")
  scaled_close(sq.y.covariance()[:, 0, 0], torch.as_tensor(g["sq_var"]), 1e-11, "squash var")
  scaled_close(sq.cross_covariance(preinv=True)[:, 0, 0], torch.as_tensor(g["sq_cross_pre"]), 1e-12, "squash cross")
  obj = GaussianObjective(_dev(g["obj_target"]), _dev(g["obj_W"]))
  scaled_close(obj(x), torch.as_tensor(g["obj_expected"]), 1e-12, "expected cost")
  scaled_close(obj(_dev(g["obj_X"])), torch.as_tensor(g["obj_samples"]), 1e-13, "sample cost")


def test_rollout_golden():
  """Upstream forward_sde + MomentMatchingEuler + loss callback (5 steps) vs the fused device rollout."""
  from gpflowpilco_b200 import ops
  from gpflowpilco_b200.rollouts import PolicyParams, rollout_mm
  g = load("rollout.npz")
  h = ops.GPModelHandle(_dev(g["dyn_Z"]), _dev(g["dyn_ell"]), _dev(g["dyn_var"]), _dev(g["dyn_q_mu"]), _dev(g["dyn_q_sqrt"]), whiten=True,
                        mean_const=_dev(g["dyn_c"]))
  P = PolicyParams(_dev(g["pol_Z"]), _dev(g["pol_ell"]), _dev(g["pol_var"]), _dev(g["pol_q_mu"][:, 0][None]), whiten=True,
                   squash_scale=float(g["scale"]), squash_shift=float(g["shift"]))
  res = rollout_mm(h, P, _dev(g["m0"]), _dev(g["S0"]), int(g["horizon"]), tuple(int(a) for a in g["active_dims"]), _dev(g["target"]),
                   _dev(g["W"]), return_trajectory=True)
  scaled_close(res.traj_m, torch.as_tensor(g["traj_m"]), 1e-8, "trajectory means")
  scaled_close(res.traj_S, torch.as_tensor(g["traj_S"]), 1e-6, "trajectory covariances")
  scaled_close(res.loss, torch.as_tensor(g["loss"]), 1e-7, "loss")


def test_squash_nd_golden():
  """Multi-dimensional squashing link on the device (Genz BVN branch) against upstream's outputs (tests/golden/squash_nd.npz),
  through the reference-facing Chain[Scale, Shift, NormalCDF] bijector, batched over the states."""
  from gpflowpilco_b200 import models as M
  from gpflowpilco_b200.moment_matching import GaussianMoments, moment_matching
  g = load("squash_nd.npz")
  link = M.BijectorChain([M.Scale(float(g["scale"])), M.Shift(float(g["shift"])), M.NormalCDF()])
  for tag in ("a2_weak", "a2_mid", "a2_strong", "a3_mid", "a3_strong"):
    x = GaussianMoments((_dev(g[f"{tag}_m"]), _dev(g[f"{tag}_S"])), True)
    mm = moment_matching(x, link)
    scaled_close(mm.y.mean(), torch.as_tensor(g[f"{tag}_mean"]), 1e-9, f"{tag} mean")
    scaled_close(mm.y.covariance(), torch.as_tensor(g[f"{tag}_cov"]), 1e-9, f"{tag} cov")
    scaled_close(mm.cross_covariance(preinv=True), torch.as_tensor(g[f"{tag}_cross_pre"]), 1e-9, f"{tag} cross (pre-inverted)")
