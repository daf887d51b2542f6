"""The CUDA backend of gpflowpilco_b200/adapters/upstream.py on a B200: GPflow-SHAPED objects (plain classes with GPflow's attribute layout —
the real packages cannot be installed in the image, and /root/reference does not exist on the GPU box) are unpacked by the adapter's own
converters and answered by libgpp_b200.so; results against the vectors upstream produced (tests/golden/*.npz)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Zero:
  pass


class Constant:
  def __init__(self, c):
    self.c = np.asarray(c)


def _svgp(Z, ell, var, q_mu, q_sqrt, c, whiten=True, W=None):
  L = Z.shape[0]
  kernel = NS(kernels=[NS(variance=var[l], lengthscales=ell[l], active_dims=None) for l in range(L)])
  if W is not None:
    kernel.W = W
  return NS(kernel=kernel, inducing_variable=NS(inducing_variables=[NS(Z=Z[l]) for l in range(L)]), q_mu=q_mu, q_sqrt=q_sqrt, whiten=whiten,
            mean_function=Zero() if c is None else Constant(c))


def test_cuda_backend_predict_and_rollout_against_upstream_vectors():
  from gpflowpilco_b200.adapters import upstream as ad
  be = ad.CudaBackend()
  # multi-output SVGP with coregionalisation (the upstream test shape), unpacked by the adapter's converter
  g = np.load(os.path.join(GOLD, "mm_models.npz"))
  model = _svgp(g["co_Z"], g["co_ell"], g["co_var"], g["co_q_mu"], g["co_q_sqrt"], g["co_c"], whiten=False, W=g["co_W"])
  params = ad.svgp_parameters(model)
  f1, Sff, cross = be.predict(params, g["mx"], g["Sxx"], True, True, 0.0, key="co")
  np.testing.assert_allclose(f1.cpu().numpy(), g["co_mean"], rtol=1e-8)
  np.testing.assert_allclose(Sff.cpu().numpy(), g["co_cov"], rtol=0, atol=1e-6 * np.abs(g["co_cov"]).max())
  np.testing.assert_allclose(cross.cpu().numpy(), g["co_cross_pre"], rtol=0, atol=1e-7 * np.abs(g["co_cross_pre"]).max())
  # exact GPR
  gpr = NS(data=(g["gpr_X"], g["gpr_Y"]), kernel=NS(variance=g["gpr_var"], lengthscales=g["gpr_ell"]), mean_function=Constant(g["gpr_c"]),
           likelihood=NS(variance=g["gpr_noise"]))
  f1, Sff, _ = be.predict(ad.gpr_parameters(gpr), g["mx"], g["Sxx"], True, True, 0.0, key="gpr")
  np.testing.assert_allclose(f1.cpu().numpy(), g["gpr_mean"], rtol=1e-7)
  np.testing.assert_allclose(Sff.cpu().numpy(), g["gpr_cov"], rtol=0, atol=1e-6 * np.abs(g["gpr_cov"]).max())
  # the fused closure: the structure the adapter extracts from a cart-pole loop object
  r = np.load(os.path.join(GOLD, "rollout.npz"))
  drift = _svgp(r["dyn_Z"], r["dyn_ell"], r["dyn_var"], r["dyn_q_mu"], r["dyn_q_sqrt"], r["dyn_c"])
  pol = _svgp(r["pol_Z"], r["pol_ell"], r["pol_var"], r["pol_q_mu"], r["pol_q_sqrt"], None)
  Scale, Shift, NormalCDF = type("Scale", (), {}), type("Shift", (), {}), type("NormalCDF", (), {})
  sc, sh = Scale(), Shift()
  sc.scale, sh.shift = float(r["scale"]), float(r["shift"])
  KernelRegressor = type("KernelRegressor", (), {})
  InverseLinkWrapper = type("InverseLinkWrapper", (), {})
  kr = KernelRegressor()
  kr.model = pol
  policy = InverseLinkWrapper()
  policy.model, policy.invlink = kr, NS(bijectors=[sc, sh, NormalCDF()])
  TrigonometricEncoder = type("TrigonometricEncoder", (), {})
  GaussianObjective = type("GaussianObjective", (), {})
  enc, obj = TrigonometricEncoder(), GaussianObjective()
  enc.active_dims = tuple(int(a) for a in r["active_dims"])
  obj.target, obj.precis = r["target"], r["W"]
  loop = NS(encoder=enc, policy=policy, drift=drift, objective=obj, diffusion=None)
  spec = ad._cartpole_structure(loop)
  assert spec is not None and spec["active_dims"] == enc.active_dims
  loss = be.rollout(spec["dynamics"], spec["policy"], r["m0"], r["S0"], int(r["horizon"]), spec["active_dims"], spec["target"], spec["W"], key="dyn")
  np.testing.assert_allclose(loss.cpu().numpy(), r["loss"], rtol=1e-7)
  assert ad.device_pointer(loss) == loss.data_ptr()
