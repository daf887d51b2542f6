"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous sharding by global index, the packed cost+gradient
all-reduce, the uneven all-gather of per-restart losses, and sharding invariance of the mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpflowpilco_b200.distributed import (allgather_concat, allreduce_sum_packed, shard_range, sharded_mean_loss_and_grad,
                                          sharded_restarts)


def test_shard_range_partitions_exactly():
  for total in (0, 1, 7, 512, 2 ** 20, 1000003):
    for ws in (1, 2, 3, 8):
      blocks = [shard_range(total, r, ws) for r in range(ws)]
      assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
      for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
        assert s0 + c0 == s1
      assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
  with pytest.raises(ValueError):
    shard_range(10, 2, 2)


def test_single_process_is_identity():
  a, b = torch.arange(3.0, dtype=torch.float64), torch.ones(2, 2, dtype=torch.float64)
  ra, rb = allreduce_sum_packed([a, b])
  assert torch.equal(ra, a) and torch.equal(rb, b)
  assert torch.equal(allgather_concat(a, 3), a)


def _unit_loss(idx):   # deterministic per-unit "cost" and "gradient" depending on the GLOBAL index only
  x = idx.to(torch.float64)
  return torch.sin(0.37 * x) + 0.01 * x, torch.stack([torch.cos(0.11 * x), x * 1e-3, torch.ones_like(x)], -1)


def _worker(rank, ws, port, total, ret):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=ws)
  try:
    def local(start, count):
      l, g = _unit_loss(torch.arange(start, start + count))
      return l.sum(), [g.sum(0), g[:, :2].sum(0).reshape(1, 2)]
    mean, grads = sharded_mean_loss_and_grad(total, local)

    def local_r(start, count):
      l, g = _unit_loss(torch.arange(start, start + count))
      return l, [g]
    losses, (start, count), g = sharded_restarts(total, local_r)
    ret[rank] = (float(mean), [x.clone() for x in grads], losses.clone(), start, count, g[0].clone())
  finally:
    dist.destroy_process_group()


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


@pytest.mark.parametrize("total", [7, 64])
def test_world_size_2_matches_single_process(total):
  ws = 2
  mgr = mp.Manager()
  ret = mgr.dict()
  mp.spawn(_worker, args=(ws, _free_port(), total, ret), nprocs=ws, join=True)
  l, g = _unit_loss(torch.arange(total))
  for rank in range(ws):
    mean, grads, losses, start, count, gblock = ret[rank]
    assert abs(mean - float(l.mean())) < 1e-14
    assert torch.allclose(grads[0], g.mean(0), atol=1e-14) and grads[1].shape == (1, 2)
    assert torch.allclose(grads[1], g[:, :2].mean(0).reshape(1, 2), atol=1e-14)
    assert torch.equal(losses, l)                      # uneven all-gather reassembles the global order exactly
    assert (start, count) == shard_range(total, rank, ws)
    assert torch.equal(gblock, g[start:start + count])
  assert ret[0][0] == ret[1][0]                        # every rank holds the same reduced value, bit for bit
