"""ctypes binding of libgpp_b200.so (include/gpp_b200.h).  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpp_b200.so")

GPP_OK = 0
STATUS_NAMES = {0: "GPP_OK", -1: "GPP_ERR_BAD_SHAPE", -2: "GPP_ERR_UNSUPPORTED", -3: "GPP_ERR_NOT_PD",
                -4: "GPP_ERR_CUDA", -5: "GPP_ERR_WORKSPACE", -6: "GPP_ERR_NULL"}

# symbol -> (restype, argtypes); mirrors include/gpp_b200.h one for one (tests check every declared symbol resolves)
_P = c_void_p
SIGNATURES = {
    "gpp_version": (c_int, []),
    "gpp_last_error": (c_char_p, []),
    "gpp_launch_count": (c_ulonglong, []),
    "gpp_ekxz": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, c_double, _P, _P, _P]),
    "gpp_ekzxkxz": (c_int, [_P, _P, c_int, c_int, _P, c_int, _P, c_double, _P, c_int, _P, c_double, _P, _P, _P]),
    "gpp_gp_model_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, _P, _P, c_int,
                                    POINTER(c_double), c_int, _P]),
    "gpp_gp_model_destroy": (c_int, [_P]),
    "gpp_gp_model_weights": (c_int, [_P, _P, _P, _P]),
    "gpp_mm_encoder": (c_int, [c_int, c_int, c_int, POINTER(c_int), _P, _P, _P, _P, _P, _P]),
    "gpp_mm_squash": (c_int, [c_int, _P, _P, c_double, c_double, _P, _P, _P, _P]),
    "gpp_mm_squash_nd": (c_int, [c_int, c_int, _P, _P, c_double, c_double, _P, _P, _P, _P]),
    "gpp_mm_squash_nd_bwd": (c_int, [c_int, c_int, _P, _P, c_double, c_double, _P, _P, _P, _P, _P, _P]),
    "gpp_cost_gaussian": (c_int, [c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "gpp_cost_samples": (c_int, [c_int, c_int, _P, _P, _P, _P, _P]),
    "gpp_owens_t": (c_int, [c_int, _P, _P, _P, _P]),
    "gpp_policy_prepare": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_double, _P, _P, _P]),
    "gpp_policy_prepare_bwd": (c_int, [c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_double, _P, _P, _P, _P]),
    "gpp_rollout_mm_workspace_bytes": (c_size_t, [_P, c_int, c_int]),
    "gpp_rollout_mm_fwd": (c_int, [_P, c_int, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, _P, c_double, c_double,
                                   _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "gpp_rollout_mm_bwd_workspace_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "gpp_rollout_mm_set_mode": (c_int, [c_int]),
    "gpp_rollout_mm_get_mode": (c_int, []),
    "gpp_rollout_mm_bwd": (c_int, [_P, c_int, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, _P, c_double, c_double,
                                   _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "gpp_rollout_mm_saved_doubles": (c_size_t, [_P, c_int, c_int, c_int]),
    "gpp_rollout_mm_fwd_save": (c_int, [_P, c_int, c_int, c_int, POINTER(c_int), c_int, c_int, _P, _P, _P, _P, c_double, c_double,
                                        _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P, _P]),
    "gpp_pathwise_tile": (c_int, []),
    "gpp_pathwise_particles_per_cta": (c_int, []),
    "gpp_pathwise_pack_basis": (c_int, [c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gpp_rollout_pathwise_fwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                         _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_double, c_double, _P, _P,
                                         _P, _P, _P, _P, _P]),
    "gpp_rollout_pathwise_fwd_mixed": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                               _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_double, c_double, _P, _P,
                                               _P, _P, _P, _P, _P]),
    "gpp_pathwise_weights_f32": (c_int, [c_longlong, _P, _P, _P]),
    "gpp_dtype_supported": (c_int, [c_char_p, c_int]),
    "gpp_rollout_pathwise_fwd_typed": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                               _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_double, c_double, _P, _P,
                                               _P, _P, _P, _P, _P]),
    "gpp_rollout_pathwise_fwd_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int),
                                              _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, c_double, c_double, _P, _P,
                                              _P, _P, _P, _P, _P, _P]),
    "gpp_rollout_pathwise_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "gpp_rollout_pathwise_bwd": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), c_int, _P, _P, c_double,
                                         _P, c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gpp_philox_raw": (c_int, [c_ulonglong, c_int, ctypes.c_uint, c_ulonglong, _P, _P]),
    "gpp_pathwise_draw_basis": (c_int, [c_int, c_int, c_int, c_ulonglong, _P, _P, _P]),
    "gpp_pathwise_draw_x0": (c_int, [c_int, c_ulonglong, c_int, _P, _P, c_ulonglong, _P, _P]),
    "gpp_pathwise_generate_workspace_bytes": (c_size_t, [_P, c_int, c_int]),
    "gpp_pathwise_generate": (c_int, [_P, c_int, c_int, c_ulonglong, c_int, c_int, c_ulonglong, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gpp_profile_enable": (c_int, [c_int]),
    "gpp_profile_last_ms": (c_int, [POINTER(ctypes.c_float)]),
    "gpp_microbench_fp64": (c_int, [c_int, c_int, c_int, _P, _P]),
    "gpp_mm_gp_predict_workspace_bytes": (c_size_t, [_P, c_int]),
    "gpp_mm_gp_predict_fwd": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, c_double, _P, c_size_t, _P, _P]),
    "gpp_mm_gp_predict_bwd_workspace_bytes": (c_size_t, [_P, c_int]),
    "gpp_mm_gp_predict_bwd": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, _P, _P, _P, c_size_t, _P, _P]),
}

_lib = None


class GppError(RuntimeError):
  def __init__(self, code: int, message: str):
    super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
    self.code = code


def load() -> ctypes.CDLL:
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f"{LIB_PATH} is missing. The CUDA library is the only implementation of this path (no CPU fallback): "
        "build it with `python -m gpflowpilco_b200.build` (needs nvcc, targets sm_100a).")
  lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)            # AttributeError here == header/library mismatch
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def check(code: int):
  if code == GPP_OK:
    return
  msg = load().gpp_last_error()
  text = msg.decode() if msg else ""
  if code == -3:
    raise GppError(code, text)        # façade maps this to the same failure class the reference shows (Cholesky)
  if code in (-1, -6):
    raise ValueError(f"{STATUS_NAMES[code]}: {text}")
  if code == -2:
    raise NotImplementedError(f"{STATUS_NAMES[code]}: {text}")
  raise GppError(code, text)
