"""Multi-GPU sharding of the rollout path: one process per GPU (`torch.distributed`, NCCL on GPUs / gloo in CPU tests).

The path shards without any data-path collective: independent Gaussian inputs, independent rollouts (policy
restarts / initial states) and independent particles are split into contiguous blocks by GLOBAL index; the only exchange
is the reduction of the expected cost and its policy gradient — `[1 + P]` float64 numbers per closure evaluation
(P = 185 for the cart-pole policy) — or the gather of `loss[R]` for restarts.  The reference has no multi-device code at
all; the partition axes are its implicit batch dimensions (loops/pilco.py:209,223-224,247,300-303).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
  """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_range(total: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous block [start, start + count) of `total` units owned by `rank`; the first total % world ranks get one more."""
  if total < 0 or world_size < 1 or not 0 <= rank < world_size:
    raise ValueError("shard_range: bad arguments")
  base, rem = divmod(total, world_size)
  return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def allreduce_sum_packed(tensors: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
  """One all-reduce (sum) for several small tensors: they are packed into a single flat buffer (cost + gradients =
  1 + P doubles), reduced, and unpacked into tensors of the original shapes.  Identity when not distributed."""
  tensors = list(tensors)
  if not tensors:
    return []
  flat = torch.cat([t.reshape(-1) for t in tensors])
  if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
  out, off = [], 0
  for t in tensors:
    out.append(flat[off:off + t.numel()].reshape(t.shape))
    off += t.numel()
  return out


def allgather_concat(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
  """Concatenate per-rank blocks (sharded along dim 0 by `shard_range(total, ...)`, possibly uneven) on every rank."""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return local
  ws = dist.get_world_size(group)
  counts = [shard_range(total, r, ws)[1] for r in range(ws)]
  width = max(counts)
  pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
  pad[:local.shape[0]] = local
  bufs = [torch.empty_like(pad) for _ in range(ws)]
  dist.all_gather(bufs, pad, group=group)
  return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


def sharded_mean_loss_and_grad(total_units: int, local_fn: Callable[[int, int], Tuple[torch.Tensor, Sequence[torch.Tensor]]],
                               group=None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
  """Mean cost over `total_units` (particles or initial states) and its gradient, identical on every rank.

  `local_fn(start, count)` evaluates this rank's block of GLOBAL unit indices and returns (sum of the unit losses,
  gradients of that sum w.r.t. the shared policy parameters).  One packed all-reduce of [1 + P] doubles follows
  (upstream has a single device: this is the mean over the batch dimension of loops/pilco.py:277-295)."""
  rank, ws = (dist.get_rank(group), dist.get_world_size(group)) if (dist.is_available() and dist.is_initialized()) else (0, 1)
  start, count = shard_range(total_units, rank, ws)
  loss_sum, grads = local_fn(start, count)
  reduced = allreduce_sum_packed([loss_sum.reshape(1)] + list(grads), group)
  scale = 1.0 / float(total_units)
  return reduced[0][0] * scale, [g * scale for g in reduced[1:]]


def sharded_restarts(total_restarts: int, local_fn: Callable[[int, int], Tuple[torch.Tensor, Sequence[torch.Tensor]]], group=None):
  """Policy restarts (independent parameter sets) split over ranks: `local_fn(start, count)` returns (loss [count], per-restart
  gradients [count, ...]); every rank receives the full `loss [R]` (4 KB at R = 512) and keeps its own gradient block."""
  rank, ws = (dist.get_rank(group), dist.get_world_size(group)) if (dist.is_available() and dist.is_initialized()) else (0, 1)
  start, count = shard_range(total_restarts, rank, ws)
  loss, grads = local_fn(start, count)
  return allgather_concat(loss, total_restarts, group), (start, count), list(grads)


# ---------------------------------------------------------------------------------------------------------
# the two closures of the path, sharded
# ---------------------------------------------------------------------------------------------------------
def pathwise_policy_loss_and_grad(handle, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor,
                                  m0: torch.Tensor, S0: torch.Tensor, total_particles: int, num_bases: int, seed: int, horizon: int,
                                  active_dims: Sequence[int], cost_target: torch.Tensor, cost_W: torch.Tensor, squash_scale: float,
                                  squash_shift: float = -0.5, whiten: bool = True, group=None, max_particles_per_launch: int = 148 * 512):
  """Mean pathwise cost of `total_particles` particles and its gradient w.r.t. (Z [1,Mp,De], lengthscales [1,De], q_mu [1,Mp]),
  particles sharded by global index (Philox streams are keyed by it, so the result does not depend on the world size up to
  summation order)."""
  from gpflowpilco_b200.autograd import rollout_pathwise_loss
  from gpflowpilco_b200.pathwise import draw_initial_states, generate_paths

  def local(start: int, count: int):
    params = [t.detach().clone().requires_grad_(True) for t in (Z, lengthscales, q_mu)]
    total = torch.zeros((), dtype=torch.float64, device=Z.device)
    done = 0
    while done < count:   # chunks of one wave of CTAs; weights for all particles at once would not fit (139 KB / particle)
      n = min(max_particles_per_launch, count - done)
      paths = generate_paths(handle, n, num_bases, seed, first_particle=start + done)
      x0 = draw_initial_states(m0, S0, seed, start + done, n)
      loss = rollout_pathwise_loss(paths, params[0], params[1], variance, params[2], x0, horizon, active_dims, cost_target, cost_W,
                                   squash_scale=squash_scale, squash_shift=squash_shift, whiten=whiten)
      s = loss.sum()
      s.backward()
      total = total + s.detach()
      done += n
    return total, [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]

  return sharded_mean_loss_and_grad(total_particles, local, group)


def mm_restart_losses_and_grads(dynamics, Z: torch.Tensor, lengthscales: torch.Tensor, variance: torch.Tensor, q_mu: torch.Tensor,
                                m0: torch.Tensor, S0: torch.Tensor, horizon: int, active_dims: Sequence[int], cost_target: torch.Tensor,
                                cost_W: torch.Tensor, squash_scale: float, squash_shift: float = -0.5, whiten: bool = True, group=None):
  """BASELINE config #5: R policy restarts (Z [R,Mp,De], lengthscales [R,De], variance [R], q_mu [R,Mp]) from one initial
  state distribution (m0 [1,Dx], S0 [1,Dx,Dx]), sharded over ranks; forward + backward of the moment-matched rollout.
  Returns (loss [R] on every rank, (start, count) of this rank's block, its gradients [count, ...] for Z, lengthscales, q_mu)."""
  from gpflowpilco_b200.autograd import rollout_mm_loss
  from gpflowpilco_b200.rollouts import raise_deferred
  R = Z.shape[0]

  def local(start: int, count: int):
    sl = slice(start, start + count)
    params = [t[sl].detach().clone().requires_grad_(True) for t in (Z, lengthscales, q_mu)]
    loss = rollout_mm_loss(dynamics, params[0], params[1], variance[sl].contiguous(), params[2], m0.expand(count, -1).contiguous(),
                           S0.expand(count, -1, -1).contiguous(), horizon, active_dims, cost_target, cost_W,
                           squash_scale=squash_scale, squash_shift=squash_shift, whiten=whiten, check="defer")
    loss.sum().backward()
    raise_deferred()      # the not-positive-definite flags of policy weights, forward and reverse sweep: one synchronisation
    return loss.detach(), [p.grad for p in params]

  return sharded_restarts(R, local, group)
