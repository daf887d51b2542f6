"""CPU checks of the drop-in boundary: the shared library loads (no GPU needed), exports every function declared in
include/gpp_b200.h, the ctypes table mirrors the header one for one, argument validation works without a device, and the
product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
  text = open(os.path.join(ROOT, "include", "gpp_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(gpp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  from gpflowpilco_b200 import _lib
  lib = _lib.load()
  names = declared_functions()
  assert len(names) >= 25
  for n in names:
    assert hasattr(lib, n), f"{n} declared in include/gpp_b200.h but not exported by libgpp_b200.so"
  assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
  assert lib.gpp_version() >= 100


def test_argument_validation_without_device():
  from gpflowpilco_b200 import _lib
  lib = _lib.load()
  rc = lib.gpp_ekxz(None, None, 1, 2, None, 1, None, 1.0, None, None, None)
  assert rc == -6 and b"null" in lib.gpp_last_error()
  buf = ctypes.c_void_p(1)
  rc = lib.gpp_ekxz(buf, buf, 1, 99, buf, 1, buf, 1.0, buf, None, None)
  assert rc == -2 and b"unsupported" in lib.gpp_last_error()
  with pytest.raises(NotImplementedError):
    _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without CUDA")
def test_no_cpu_fallback():
  from gpflowpilco_b200 import ops
  x = torch.zeros(1, 2, dtype=torch.float64)
  with pytest.raises(RuntimeError, match="no CPU implementation"):
    ops.ekxz(x, torch.eye(2, dtype=torch.float64)[None], torch.zeros(3, 2, dtype=torch.float64), torch.ones(2, dtype=torch.float64), 1.0)


def test_dispatcher_and_episode_spec():
  from gpflowpilco_b200.loops import EpisodeSpec
  from gpflowpilco_b200.utils.dispatch import Dispatcher
  assert EpisodeSpec(None, horizon=3.0, step_size=0.1).num_steps == 30      # upstream experiment.py:121-122 -> H = 30
  d = Dispatcher("t")

  class A: pass
  class B(A): pass

  d.register(A, (int, float))(lambda a, b: "A")
  d.register(B, int)(lambda a, b: "B")
  assert d(B(), 1) == "B" and d(B(), 1.0) == "A" and d(A(), 2) == "A"
  with pytest.raises(NotImplementedError):
    d(A(), "s")


def test_graphed_policy_gradient_refuses_cpu_tensors():
  """The CUDA-graph wrapper is part of the product path: no CPU fallback, it fails loudly on host tensors."""
  import pytest
  import torch
  from gpflowpilco_b200.graphs import GraphedMMPolicyGradient
  z = torch.zeros(1, 4, 5, dtype=torch.float64)
  with pytest.raises(ValueError, match="CUDA"):
    GraphedMMPolicyGradient(None, z, torch.ones(1, 5, dtype=torch.float64), torch.ones(1, dtype=torch.float64),
                            torch.zeros(1, 4, dtype=torch.float64), torch.zeros(1, 4, dtype=torch.float64),
                            torch.eye(4, dtype=torch.float64)[None], 3, (1,), torch.zeros(5, dtype=torch.float64),
                            torch.eye(5, dtype=torch.float64))


def test_dtype_probe_is_a_host_function():
  """gpp_dtype_supported answers without touching a device: FP64 everywhere, the mixed variant only for the pathwise forward rollout"""
  from gpflowpilco_b200 import _lib
  lib = _lib.load()
  assert lib.gpp_dtype_supported(b"gpp_mm_gp_predict_fwd", 0) == 1
  assert lib.gpp_dtype_supported(b"gpp_mm_gp_predict_fwd", 1) == 0
  assert lib.gpp_dtype_supported(b"gpp_rollout_pathwise_fwd", 1) == 1
  assert lib.gpp_dtype_supported(b"gpp_rollout_pathwise_fwd_typed", 1) == 1
  assert lib.gpp_dtype_supported(b"gpp_rollout_pathwise_fwd", 7) == 0
