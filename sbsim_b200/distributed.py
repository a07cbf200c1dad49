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


class PackedTimeStep:
  """One device block [obs B*D | reward B | step_type B (int32) | discount B] with typed views:
  the layout libsbx itself uses for a TimeStep (sbx_api.cu, `d_obs` block), so that a whole
  TimeStep moves in ONE copy or ONE collective."""

  def __init__(self, batch: int, obs_dim: int, device):
    import torch
    self.batch, self.obs_dim = int(batch), int(obs_dim)
    self.block = torch.zeros(self.batch * (self.obs_dim + 3), dtype=torch.float32, device=device)
    self.obs, self.reward, self.step_type, self.discount = self.views(self.block[None])
    self.obs, self.reward = self.obs[0], self.reward[0]
    self.step_type, self.discount = self.step_type[0], self.discount[0]

  def views(self, blocks):
    """Typed views of `blocks` [n, B*(D+3)] -> obs [n,B,D], reward [n,B], step_type [n,B], discount [n,B]."""
    import torch
    b, d = self.batch, self.obs_dim
    obs = blocks[:, :b * d].view(blocks.shape[0], b, d)
    reward = blocks[:, b * d:b * d + b]
    step_type = blocks[:, b * d + b:b * d + 2 * b].view(torch.int32)
    discount = blocks[:, b * d + 2 * b:b * d + 3 * b]
    return obs, reward, step_type, discount


class TimeStepGather:
  """Pre-allocated all-gather of every rank's packed TimeStep block: ONE collective per step
  (ncclAllGather through torch.distributed.all_gather_into_tensor; gloo in the CPU tests)."""

  def __init__(self, packed: PackedTimeStep, group=None):
    import torch
    import torch.distributed as dist
    self._dist = dist
    self._group = group
    self.packed = packed
    self.world = dist.get_world_size(group)
    self.blocks = torch.empty((self.world, packed.block.numel()), dtype=packed.block.dtype,
                              device=packed.block.device)
    self.obs, self.reward, self.step_type, self.discount = packed.views(self.blocks)
    self.bytes_per_rank = packed.block.numel() * 4

  def __call__(self):
    self._dist.all_gather_into_tensor(self.blocks.view(-1), self.packed.block, group=self._group)
    return self.obs, self.reward, self.step_type, self.discount


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
    ts = PackedTimeStep(b, d, "cpu")
    ts.obs.copy_(torch.full((b, d), float(rank)) + torch.arange(b)[:, None])
    ts.reward.copy_(torch.tensor(wl.weather_low, dtype=torch.float32))
    ts.step_type.fill_(rank)
    ts.discount.fill_(0.5 + rank)
    gather = TimeStepGather(ts)
    o, r, s, di = gather()
    assert o.shape == (world, b, d) and r.shape == (world, b) and s.shape == (world, b)
    assert s.dtype == torch.int32
    for k in range(world):
      assert torch.all(s[k] == k) and torch.all(di[k] == 0.5 + k)
      assert torch.all(o[k, :, 0] == float(k) + torch.arange(b))
    r = r.reshape(-1)
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
