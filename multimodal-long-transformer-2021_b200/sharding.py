"""Multi-GPU layout of the attention path: every (batch, head) unit is independent, so the
units are partitioned round-robin over the ranks (one process per GPU) and NO collective runs
on the data path.  torch.distributed is used only for the timing reduction."""

from __future__ import annotations

import torch
import torch.distributed as dist


def partition_units(batch: int, heads: int, rank: int, world: int):
  """(b, h) units owned by `rank`: round-robin over the flattened batch x head index."""
  return [(u // heads, u % heads) for u in range(batch * heads) if u % world == rank]


def _reduce(value: float, op, device=None) -> float:
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
    return float(value)
  t = torch.tensor([value], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=op)
  return float(t.item())


def max_over_ranks(value: float, device=None) -> float:
  return _reduce(value, dist.ReduceOp.MAX, device)


def sum_over_ranks(value: float, device=None) -> float:
  return _reduce(value, dist.ReduceOp.SUM, device)
