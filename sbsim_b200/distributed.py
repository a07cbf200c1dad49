"""Multi-GPU plumbing: the building batch shards embarrassingly, one process per
GPU, no data-path collective.  The only collective is the
optional all-gather of observations / rewards / step types to every rank (or to
a learner on rank 0), over NCCL on the GPU box and gloo in CPU tests."""

from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
  """Contiguous split: rank r owns envs [lo, hi); sizes differ by at most one."""
  if not 0 <= rank < world_size:
    raise ValueError(f"rank {rank} outside world of {world_size}")
  base, rem = divmod(n_total, world_size)
  lo = rank * base + min(rank, rem)
  hi = lo + base + (1 if rank < rem else 0)
  return lo, hi


def rank_seed(base_seed: int, rank: int) -> int:
  """Workload seed of a rank (config 4: seed 2024 + rank)."""
  return base_seed + rank


class TimeStepGather:
  """Pre-allocated all-gather of (observation, reward, step_type) over equal shards."""

  def __init__(self, obs, reward, step_type, group=None):
    import torch
    import torch.distributed as dist
    self._dist = dist
    self._group = group
    self.world = dist.get_world_size(group)
    w = self.world
    self.obs = torch.empty((w * obs.shape[0],) + tuple(obs.shape[1:]), dtype=obs.dtype,
                           device=obs.device)
    self.reward = torch.empty((w * reward.shape[0],), dtype=reward.dtype, device=reward.device)
    self.step_type = torch.empty((w * step_type.shape[0],), dtype=step_type.dtype,
                                 device=step_type.device)

  def __call__(self, obs, reward, step_type):
    d = self._dist
    d.all_gather_into_tensor(self.obs, obs.contiguous(), group=self._group)
    d.all_gather_into_tensor(self.reward, reward.contiguous(), group=self._group)
    d.all_gather_into_tensor(self.step_type, step_type.contiguous(), group=self._group)
    return self.obs, self.reward, self.step_type


def _selftest(rank: int, world: int, port: int, out_path: str) -> None:
  """Two-process CPU check used by tests/test_distributed_cpu.py (gloo backend)."""
  import os
  import numpy as np
  import torch
  import torch.distributed as dist
  from sbsim_b200 import workloads
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    b, d = 6, 5
    lo, hi = shard_range(world * b, rank, world)
    assert hi - lo == b
    # each rank generates ITS shard of the workload from its own seed
    wl = workloads.randomized(b, seed=rank_seed(2024, rank), n_layouts=2)
    obs = torch.full((b, d), float(rank)) + torch.arange(b)[:, None]
    rew = torch.tensor(wl.weather_low, dtype=torch.float32)
    st = torch.full((b,), rank, dtype=torch.int32)
    gather = TimeStepGather(obs, rew, st)
    o, r, s = gather(obs, rew, st)
    assert o.shape == (world * b, d) and r.shape == (world * b,) and s.shape == (world * b,)
    for k in range(world):
      assert torch.all(s[k * b:(k + 1) * b] == k)
      assert torch.all(o[k * b:(k + 1) * b, 0] == float(k) + torch.arange(b))
    # the max-over-ranks timing reduction bench.py uses
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == float(world)
    if rank == 0:
      np.save(out_path, r.numpy())
    dist.barrier()
  finally:
    dist.destroy_process_group()


if __name__ == "__main__":
  import sys
  _selftest(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
