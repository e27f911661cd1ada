"""Multi-GPU layout of the attention path: every (batch, head) unit is independent, so the units are
partitioned over the ranks (one process per GPU) and NO collective runs on the data path.
torch.distributed is used only for the barrier and the timing reduction.

``partition_units`` is the round-robin unit map; ``shard_inputs`` turns it into zero-copy views of one
global problem (head slices when the world size divides the head count, batch slices otherwise), which is
what ``bench.py --gpus N`` and the shard-vs-single parity tests use: running the operator on every rank's
views and stitching the outputs back reproduces the single-GPU result bit for bit."""

from __future__ import annotations

import torch
import torch.distributed as dist


def partition_units(batch: int, heads: int, rank: int, world: int):
  """(b, h) units owned by `rank`: round-robin over the flattened batch x head index."""
  return [(u // heads, u % heads) for u in range(batch * heads) if u % world == rank]


def shard_axis(batch: int, heads: int, world: int) -> str:
  """Which tensor axis realises the round-robin partition as a strided view: 'heads' when the world size
  divides the head count (rank r owns heads r, r + world, ... of every batch element -- exactly
  ``partition_units``), else 'batch' (rank r owns batch elements r, r + world, ...)."""
  if world == 1:
    return 'none'
  if heads % world == 0:
    return 'heads'
  if batch % world == 0:
    return 'batch'
  raise ValueError(f'cannot shard {batch} x {heads} (batch x head) units evenly over {world} ranks')


def shard_inputs(x: dict, rank: int, world: int, axis: str):
  """Views (no copies) of one global global-local attention problem for `rank`.  `x` holds the tensors of
  ``synthetic.make_inputs``: q/k/v/d_out ``[B, len, H, d]``, tables ``[R, H, d]`` / ``[R, H]``, descriptors
  ``[B, len]``."""
  if axis == 'none':
    return dict(x)
  out = {}
  for k, v in x.items():
    if not torch.is_tensor(v):
      out[k] = v
    elif axis == 'heads':
      if v.dim() == 4:
        out[k] = v[:, :, rank::world]                  # [B, len, H, d]: stride_h becomes world * d
      elif k.endswith('_emb'):
        out[k] = v[:, rank::world]
      elif k.endswith('_bias'):
        out[k] = v[:, rank::world]
      else:
        out[k] = v                                     # descriptors are per batch element
    else:
      out[k] = v[rank::world] if (v.dim() == 4 or v.dim() == 2 and not k.endswith('_bias')) else v
  return out


def unshard_outputs(parts, world: int, axis: str, like: torch.Tensor):
  """Inverse of the partition for a ``[B, len, H, d]`` output: parts[r] is rank r's result."""
  if axis == 'none':
    return parts[0]
  full = torch.empty_like(like)
  for r, p in enumerate(parts):
    if axis == 'heads':
      full[:, :, r::world] = p
    else:
      full[r::world] = p
  return full


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
