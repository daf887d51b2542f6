"""Pathwise sample paths on the device: packed layouts + the fused particle rollout.

`PackedPaths` holds function draws of a multi-output SVGP in the layout the kernel streams (particle-minor weights);
`rollout_pathwise` runs S particles x H steps + sample cost in one launch (upstream loops/pilco.py:263-303).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from gpflowpilco_b200 import _lib
from gpflowpilco_b200.ops import F64, _c, _dev_check, _ptr, _stream
from gpflowpilco_b200.rollouts import PolicyParams


GPP_F64, GPP_MIXED_F32_WEIGHTS = 0, 1      # gpp_dtype (include/gpp_b200.h)


def _round_up(x: int, m: int) -> int:
  return (x + m - 1) // m * m


@dataclass
class PackedPaths:
  """Function draws  f_{s,l}(x) = mean_l + sum_i w[l,i,s] phi_{l,i}(x) + sum_j v[l,j,s] k_l(x, z_{l,j})."""
  basis: torch.Tensor       # [L,F,BS]
  zbasis: torch.Tensor      # [L,Mpad,BS]
  w: torch.Tensor           # [L,F,ldS]
  v: torch.Tensor           # [L,Mpad,ldS]
  amp: torch.Tensor         # [L]
  variance: torch.Tensor    # [L]
  inv_lengthscales: torch.Tensor   # [L,D]
  mean_const: torch.Tensor  # [L]
  num_particles: int
  D: int
  w32: Optional[torch.Tensor] = None     # FP32 copy of w for the mixed-precision rollout (made on first use)

  def weights_f32(self) -> torch.Tensor:
    """FP32 copy of the Fourier weights (gpp_pathwise_weights_f32), cached: half the bytes the rollout streams."""
    if self.w32 is None:
      out = torch.empty(self.w.shape, dtype=torch.float32, device=self.w.device)
      _lib.check(_lib.load().gpp_pathwise_weights_f32(self.w.numel(), _ptr(self.w), _ptr(out), _stream()))
      self.w32 = out
    return self.w32

  @staticmethod
  def layout(S: int, F: int, M: int):
    lib = _lib.load()
    tile, ppc = lib.gpp_pathwise_tile(), lib.gpp_pathwise_particles_per_cta()
    return _round_up(S, ppc), _round_up(M, tile), tile

  @classmethod
  def from_sample_major(cls, Z, lengthscales, variance, mean_const, omega, phase, w, v) -> "PackedPaths":
    """Pack draws given in the sampler's natural layout: omega [L,F,D], phase [L,F], w [S,L,F], v [S,L,M] (all CUDA f64)."""
    Z, lengthscales, variance, omega, phase, w, v = map(_c, (Z, lengthscales, variance, omega, phase, w, v))
    _dev_check(Z, lengthscales, variance, omega, phase, w, v)
    L, M, D = Z.shape
    S, _, F = w.shape
    ldS, Mpad, tile = cls.layout(S, F, M)
    if F % tile:
      raise ValueError(f"number of Fourier bases must be a multiple of {tile}")
    dev = Z.device
    BS = (D + 2) & ~1
    basis = torch.empty(L, F, BS, dtype=F64, device=dev)
    zbasis = torch.empty(L, Mpad, BS, dtype=F64, device=dev)
    _lib.check(_lib.load().gpp_pathwise_pack_basis(L, F, M, Mpad, D, _ptr(omega), _ptr(phase), _ptr(Z), _ptr(lengthscales),
                                                   _ptr(basis), _ptr(zbasis), _stream()))
    wp = torch.zeros(L, F, ldS, dtype=F64, device=dev)
    wp[:, :, :S] = w.permute(1, 2, 0)
    vp = torch.zeros(L, Mpad, ldS, dtype=F64, device=dev)
    vp[:, :M, :S] = v.permute(1, 2, 0)
    mean = torch.zeros(L, dtype=F64, device=dev) if mean_const is None else _c(mean_const)
    return cls(basis, zbasis, wp, vp, torch.sqrt(2.0 * variance / F), variance, (1.0 / lengthscales).contiguous(), mean, S, D)


def philox_raw(first_index: int, count: int, stream_id: int, seed: int, device="cuda") -> torch.Tensor:
  """Raw Philox4x32-10 words [count,4] (int64 holding uint32) — bit-exactness hook against oracle/philox.py."""
  out = torch.empty(count, 4, dtype=torch.int32, device=device)
  _lib.check(_lib.load().gpp_philox_raw(int(first_index), int(count), int(stream_id), int(seed), _ptr(out), _stream()))
  return out.to(torch.int64) & 0xFFFFFFFF


def draw_basis(L: int, F: int, D: int, seed: int, device="cuda"):
  omega = torch.empty(L, F, D, dtype=F64, device=device)
  phase = torch.empty(L, F, dtype=F64, device=device)
  _lib.check(_lib.load().gpp_pathwise_draw_basis(L, F, D, int(seed), _ptr(omega), _ptr(phase), _stream()))
  return omega, phase


def draw_initial_states(m0: torch.Tensor, S0: torch.Tensor, seed: int, first_particle: int, count: int) -> torch.Tensor:
  """x0_s = m0 + chol(S0) n_s for global particles first_particle.. (upstream loops/pilco.py:300-303: p.sample([batch]))."""
  m0, chol = _c(m0.reshape(-1)), _c(torch.linalg.cholesky(S0.reshape(m0.numel(), m0.numel())))
  _dev_check(m0, chol)
  x0 = torch.empty(count, m0.numel(), dtype=F64, device=m0.device)
  _lib.check(_lib.load().gpp_pathwise_draw_x0(count, int(first_particle), m0.numel(), _ptr(m0), _ptr(chol), int(seed), _ptr(x0), _stream()))
  return x0


def generate_paths(handle, num_samples: int, num_bases: int, seed: int, first_particle: int = 0, basis=None,
                   out: Optional[PackedPaths] = None) -> PackedPaths:
  """Draw `num_samples` function samples of the GP behind `handle` (ops.GPModelHandle) for global particles
  first_particle.. ; the basis (omega, phase) is shared by all particles and depends on the seed only.
  Mirrors drift.generate_paths(num_samples, num_bases, sample_axis=0) (upstream loops/pilco.py:282-284)."""
  lib = _lib.load()
  L, M, D = handle.L, handle.M, handle.D
  ldS, Mpad, tile = PackedPaths.layout(num_samples, num_bases, M)
  if num_bases % tile:
    raise ValueError(f"num_bases must be a multiple of {tile}")
  dev = handle.device
  omega, phase = draw_basis(L, num_bases, D, seed, dev) if basis is None else basis
  if out is None:
    BS = (D + 2) & ~1
    ell, var, Z, mean = handle.parameters()
    pb = torch.empty(L, num_bases, BS, dtype=F64, device=dev)
    zb = torch.empty(L, Mpad, BS, dtype=F64, device=dev)
    _lib.check(lib.gpp_pathwise_pack_basis(L, num_bases, M, Mpad, D, _ptr(omega), _ptr(phase), _ptr(Z), _ptr(ell), _ptr(pb), _ptr(zb), _stream()))
    out = PackedPaths(pb, zb, torch.empty(L, num_bases, ldS, dtype=F64, device=dev), torch.empty(L, Mpad, ldS, dtype=F64, device=dev),
                      torch.sqrt(2.0 * var / num_bases), var, (1.0 / ell).contiguous(), mean, num_samples, D)
  need = lib.gpp_pathwise_generate_workspace_bytes(handle._h, ldS, num_bases)
  ws = torch.empty(need, dtype=torch.uint8, device=dev)
  _lib.check(lib.gpp_pathwise_generate(handle._h, num_samples, ldS, int(first_particle), num_bases, Mpad, int(seed), _ptr(omega), _ptr(phase),
                                       _ptr(out.w), _ptr(out.v), _ptr(ws), ws.numel(), _stream()))
  out.num_particles = num_samples
  out.omega, out.phase = omega, phase
  return out


def rollout_pathwise(paths: PackedPaths, policy: PolicyParams, x0: torch.Tensor, horizon: int, active_dims: Sequence[int],
                     cost_target: torch.Tensor, cost_W: torch.Tensor, return_trajectory: bool = False,
                     beta: Optional[torch.Tensor] = None, save_for_backward: bool = False, mixed_precision: bool = False):
  """loss [S] (and final states / trajectory) of S particles, particle s evaluated on function draw s.
  With save_for_backward the gradient-mode kernel runs and the result is (loss, x_final, traj, jac).
  `mixed_precision` selects gpp_rollout_pathwise_fwd_mixed (FP32 Fourier weights and cosine polynomial, everything else FP64;
  tolerance stated in include/gpp_b200.h); the default is the all-FP64 kernel."""
  x0, cost_target, cost_W = map(_c, (x0, cost_target, cost_W))
  _dev_check(x0, cost_target, cost_W)
  S, Dx = x0.shape
  if S != paths.num_particles:
    raise ValueError("one initial state per function draw (upstream loops/pilco.py:300-303)")
  L, F, _ = paths.basis.shape
  Mpad = paths.zbasis.shape[1]
  ldS = paths.w.shape[2]
  na = len(active_dims)
  De = Dx + na
  R, Mp, Dp = policy.shape
  if R != 1 or Dp != De:
    raise ValueError("rollout_pathwise: one shared policy with input dim = encoded state dim")
  if beta is None:
    beta = policy.beta()
  alpha = (policy.variance[0] * beta[0]).contiguous()
  pZs = (policy.Z[0] / policy.lengthscales[0]).contiguous()
  pinv = (1.0 / policy.lengthscales[0]).contiguous()
  dev = x0.device
  loss = torch.empty(S, dtype=F64, device=dev)
  xf = torch.empty(S, Dx, dtype=F64, device=dev)
  traj = torch.empty(horizon + 1, S, Dx, dtype=F64, device=dev) if (return_trajectory or save_for_backward) else None
  act = (ctypes.c_int * max(na, 1))(*active_dims)
  if save_for_backward:
    jac = torch.empty(horizon, L * paths.D, ldS, dtype=F64, device=dev)
    _lib.check(_lib.load().gpp_rollout_pathwise_fwd_grad(
        S, ldS, int(horizon), L, F, Mpad, paths.D, Dx, na, act, _ptr(paths.basis), _ptr(paths.zbasis), _ptr(paths.w), _ptr(paths.v),
        _ptr(paths.amp), _ptr(paths.variance), _ptr(paths.inv_lengthscales), _ptr(paths.mean_const), Mp, _ptr(pZs), _ptr(pinv),
        _ptr(alpha), float(policy.squash_scale), float(policy.squash_shift), _ptr(cost_target), _ptr(cost_W), _ptr(x0), _ptr(loss),
        _ptr(xf), _ptr(traj), _ptr(jac), _stream()))
    return loss, xf, traj, jac
  if mixed_precision:
    _lib.check(_lib.load().gpp_rollout_pathwise_fwd_typed(
        GPP_MIXED_F32_WEIGHTS, S, ldS, int(horizon), L, F, Mpad, paths.D, Dx, na, act, _ptr(paths.basis), _ptr(paths.zbasis), _ptr(paths.weights_f32()),
        _ptr(paths.v), _ptr(paths.amp), _ptr(paths.variance), _ptr(paths.inv_lengthscales), _ptr(paths.mean_const), Mp, _ptr(pZs),
        _ptr(pinv), _ptr(alpha), float(policy.squash_scale), float(policy.squash_shift), _ptr(cost_target), _ptr(cost_W), _ptr(x0),
        _ptr(loss), _ptr(xf), _ptr(traj), _stream()))
    return loss, xf, traj
  _lib.check(_lib.load().gpp_rollout_pathwise_fwd(
      S, ldS, int(horizon), L, F, Mpad, paths.D, Dx, na, act, _ptr(paths.basis), _ptr(paths.zbasis), _ptr(paths.w), _ptr(paths.v),
      _ptr(paths.amp), _ptr(paths.variance), _ptr(paths.inv_lengthscales), _ptr(paths.mean_const), Mp, _ptr(pZs), _ptr(pinv),
      _ptr(alpha), float(policy.squash_scale), float(policy.squash_shift), _ptr(cost_target), _ptr(cost_W), _ptr(x0), _ptr(loss),
      _ptr(xf), _ptr(traj), _stream()))
  return loss, xf, traj


def rollout_pathwise_bwd(policy: PolicyParams, beta: torch.Tensor, traj: torch.Tensor, jac: torch.Tensor, active_dims: Sequence[int],
                         cost_target: torch.Tensor, cost_W: torch.Tensor, loss_bar: Optional[torch.Tensor] = None):
  """Reverse sweep of the particle rollout from (traj, jac) saved by rollout_pathwise(save_for_backward=True): returns
  (Z_bar [Mp,De], lengthscales_bar [De] at fixed beta, beta_bar [Mp], x0_bar [S,Dx]), summed over the particles."""
  traj, jac, cost_target, cost_W, beta, loss_bar = map(_c, (traj, jac, cost_target, cost_W, beta, loss_bar))
  _dev_check(traj, jac, cost_target, cost_W, beta, loss_bar)
  H1, S, Dx = traj.shape
  na = len(active_dims)
  De = Dx + na
  D = De + 1
  R, Mp, Dp = policy.shape
  H, LD, ldS = jac.shape
  if R != 1 or Dp != De or H != H1 - 1 or LD != Dx * D:
    raise ValueError("rollout_pathwise_bwd: inconsistent shapes")
  dev = traj.device
  lib = _lib.load()
  need = lib.gpp_rollout_pathwise_bwd_workspace_bytes(S, Mp, De)
  ws = torch.empty(need, dtype=torch.uint8, device=dev)
  Zb = torch.empty(Mp, De, dtype=F64, device=dev)
  eb = torch.empty(De, dtype=F64, device=dev)
  bb = torch.empty(Mp, dtype=F64, device=dev)
  x0b = torch.empty(S, Dx, dtype=F64, device=dev)
  act = (ctypes.c_int * max(na, 1))(*active_dims)
  _lib.check(lib.gpp_rollout_pathwise_bwd(S, ldS, H, Dx, D, Dx, na, act, Mp, _ptr(policy.Z[0].contiguous()), _ptr(policy.lengthscales[0].contiguous()),
                                          float(policy.variance[0]), _ptr(beta[0].contiguous()), float(policy.squash_scale), _ptr(cost_target),
                                          _ptr(cost_W), _ptr(traj), _ptr(jac), _ptr(loss_bar), _ptr(Zb), _ptr(eb), _ptr(bb), _ptr(x0b),
                                          _ptr(ws), ws.numel(), _stream()))
  return Zb, eb, bb, x0b


def rollout_pathwise_chunked(handle, policy: PolicyParams, m0: torch.Tensor, S0: torch.Tensor, total_particles: int, num_bases: int, seed: int,
                             horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor,
                             first_particle: int = 0, particles_per_launch: Optional[int] = None, overlap: bool = True,
                             beta: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Sum of the particle losses of `total_particles` particles (global indices first_particle ...), processed in chunks of one launch
  each: fresh function draws per chunk (upstream loops/pilco.py:281-284), initial states 1-to-1 with the draws (:300-303).
  The weights of all particles would not fit (139 KB per particle at L=4, F=4096, M=256), so a chunk's paths are generated on the
  device right before its rollout; with `overlap` the generation of chunk k+1 (Philox draws + the update-weight solve) runs on a
  second stream while chunk k rolls out — two path buffers, events in both directions, no host synchronisation."""
  lib = _lib.load()
  S = particles_per_launch or lib.gpp_pathwise_particles_per_cta() * 148
  dev = handle.device
  if beta is None:
    beta = policy.beta()
  main = torch.cuda.current_stream(dev)
  side = torch.cuda.Stream(dev) if overlap else main
  chunks = [(first_particle + o, min(S, total_particles - o)) for o in range(0, total_particles, S)]
  bufs = [None, None]
  ev_gen = [torch.cuda.Event(), torch.cuda.Event()]
  ev_roll = [None, None]
  total = torch.zeros((), dtype=F64, device=dev)

  def generate(k):
    first, n = chunks[k]
    b = k % 2
    with torch.cuda.stream(side):
      if ev_roll[b] is not None:
        side.wait_event(ev_roll[b])                 # the rollout that read this buffer two chunks ago has finished
      if n == S and bufs[b] is not None:
        generate_paths(handle, n, num_bases, seed, first_particle=first, out=bufs[b])
      else:
        bufs[b] = generate_paths(handle, n, num_bases, seed, first_particle=first)
      ev_gen[b].record(side)

  side.wait_stream(main)
  generate(0)
  for k, (first, n) in enumerate(chunks):
    if k + 1 < len(chunks):
      generate(k + 1)
    b = k % 2
    main.wait_event(ev_gen[b])
    x0 = draw_initial_states(m0, S0, seed, first, n)
    loss, _, _ = rollout_pathwise(bufs[b], policy, x0, horizon, active_dims, cost_target, cost_W, beta=beta)
    total = total + loss.sum()
    ev_roll[b] = torch.cuda.Event()
    ev_roll[b].record(main)
  main.wait_stream(side)
  return total
