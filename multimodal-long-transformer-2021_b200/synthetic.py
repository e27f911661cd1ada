"""Synthetic inputs of the benchmark shapes (SURVEY.md section 8d / BASELINE.md section 4).

q, k, v ~ N(0,1) (seed 1234 + config index); relative tables ~ N(0, 0.02^2) with a
non-zero bias so the lookup is exercised; valid length per example ~ U[L/2, L]
(seed 99) which yields padding masks; ``sentence_ids[i] = floor(i * G / L)``.
"""

from __future__ import annotations

import dataclasses

import torch


@dataclasses.dataclass
class GlobalLocalShape:
  batch: int
  long_len: int
  global_len: int
  heads: int = 12
  head_dim: int = 64
  local_radius: int = 64
  relative_vocab_size: int = 32
  max_distance: int = 12

  @property
  def tokens(self) -> int:
    return self.batch * self.long_len

  def attended_pairs(self) -> int:
    l, g, r = self.long_len, self.global_len, self.local_radius
    return (l * (2 * r + 1) - r * (r + 1)) + l * g + g * (g + l)

  def flops_fwd(self) -> int:
    """Algorithmic forward FLOPs for the whole batch."""
    d, rv = self.head_dim, self.relative_vocab_size
    per_bh = 4 * d * self.attended_pairs() + 2 * d * rv * (self.long_len + self.global_len)
    return per_bh * self.batch * self.heads

  def bytes_fwd(self, elem_bytes: int = 2) -> int:
    """Algorithmic forward HBM bytes (q,k,v read + out written + stats)."""
    n = self.long_len + self.global_len
    per_bh = 4 * elem_bytes * self.head_dim * n + 8 * n
    return per_bh * self.batch * self.heads


# The configurations BASELINE.json names (index = seed offset).
CONFIGS = {
    'c1_fp32_512': (1, GlobalLocalShape(2, 512, 32)),
    'c3_2048': (3, GlobalLocalShape(32, 2048, 128)),
    'c3_4096': (4, GlobalLocalShape(16, 4096, 256)),
    'c3_8192': (5, GlobalLocalShape(8, 8192, 512)),
}


def make_lengths(shape: GlobalLocalShape, seed: int = 99) -> torch.Tensor:
  gen = torch.Generator().manual_seed(seed)
  lo = shape.long_len // 2
  return torch.randint(lo, shape.long_len + 1, (shape.batch,), generator=gen,
                       dtype=torch.int32)


def make_descriptors(shape: GlobalLocalShape, lengths: torch.Tensor):
  """Compact descriptors: example ids (1 = real token, 0 = padding) and sentence ids."""
  l, g = shape.long_len, shape.global_len
  pos = torch.arange(l, dtype=torch.int32)
  long_example_ids = (pos[None, :] < lengths[:, None]).to(torch.int32)
  sentence_ids = ((pos.to(torch.int64) * g) // l).to(torch.int32)[None, :].expand(
      shape.batch, l).contiguous()
  # A global token is real iff its sentence starts inside the valid prefix.
  gpos = torch.arange(g, dtype=torch.int64)
  first_long = (gpos * l + g - 1) // g
  global_example_ids = (first_long[None, :] < lengths[:, None].to(torch.int64)).to(torch.int32)
  return long_example_ids, global_example_ids, sentence_ids


def make_inputs(shape: GlobalLocalShape, seed: int, dtype=torch.float32,
                device='cpu', pin: bool = False):
  """Returns a dict of host (or device) tensors for one fwd+bwd call."""
  gen = torch.Generator().manual_seed(seed)
  b, l, g, h, d = shape.batch, shape.long_len, shape.global_len, shape.heads, shape.head_dim
  rv = shape.relative_vocab_size

  def normal(*size, std=1.0):
    t = torch.randn(*size, generator=gen, dtype=torch.float32) * std
    t = t.to(dtype)
    if pin and device == 'cpu' and torch.cuda.is_available():
      t = t.pin_memory()
    return t.to(device)

  out = dict(
      long_q=normal(b, l, h, d), long_k=normal(b, l, h, d), long_v=normal(b, l, h, d),
      global_q=normal(b, g, h, d), global_k=normal(b, g, h, d), global_v=normal(b, g, h, d),
      long_emb=normal(rv, h, d, std=0.02), long_bias=normal(rv, h, std=0.02),
      global_emb=normal(rv, h, d, std=0.02), global_bias=normal(rv, h, std=0.02),
      d_long_out=normal(b, l, h, d), d_global_out=normal(b, g, h, d),
  )
  lengths = make_lengths(shape)
  le, ge, sid = make_descriptors(shape, lengths)
  out.update(lengths=lengths, long_example_ids=le.to(device),
             global_example_ids=ge.to(device), sentence_ids=sid.to(device))
  return out
