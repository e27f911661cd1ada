"""Dense fp64 restatement of ETC relative attention (dense and global-local).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED for
attention numerics: the arithmetic is in the absent third-party ``etcmodel``
(un-pinned, "clone master", reference ``src/README.md:9-10``); this file restates
the published algorithm [UPSTREAM-RECALLED] and is anchored on the reference call
sites ``src/modeling/models/mmt_encoder.py:124-135`` (ctor) and ``:220-224``
(call: ``inputs, att_mask, relative_att_ids, training``).

Everything is written on torch CPU tensors so that ``torch.autograd`` supplies
oracle gradients (the reference has no hand-written gradient either: TF autodiff,
``src/tasks/pretraining.py:292-296``).

Semantics.  Tensors are ``[B, len, H, d]``:

    content[b,h,i,j] = sum_c q[b,i,h,c] k[b,j,h,c]
    allrel [b,h,i,p] = sum_c q[b,i,h,c] E[p,h,c] + bias[p,h]
    rel    [b,h,i,j] = allrel[b,h,i,id[b,i,j]] if 0 <= id < R else 0   (one-hot lookup,
                       reference default use_one_hot_lookup=True, src/configs/encoders.py:98)
    s = (content + rel) * SCALE_AFTER_REL(1/sqrt(d)) + NEG * (1 - mask)            (N1, N2)
    p = softmax_j(s);   out[b,i,h,:] = sum_j p v[b,j,h,:]

Chosen behaviour for degenerate rows (N3): the candidate key set of a long row
is {in-range window positions} + {all side keys}; masked candidates get the
additive NEG, non-candidates (outside the band or outside [0, L)) do not exist.
A fully-masked row therefore attends uniformly over its candidates.
"""

from __future__ import annotations

import math

import torch

NEG = -1e9  # N2 [UPSTREAM-RECALLED]: large_compatible_negative for fp32/bf16.


def _all_relative_scores(q, emb, bias):
  # [B,Lq,H,R]
  return torch.einsum('bqhd,rhd->bqhr', q, emb) + bias.transpose(0, 1)


def _lookup(allrel, ids):
  """One-hot lookup: out-of-vocabulary ids contribute 0."""
  r = allrel.shape[-1]
  valid = (ids >= 0) & (ids < r)
  idx = ids.clamp(0, r - 1).long()  # [B,Lq,Lk]
  # allrel [B,Lq,H,R] -> [B,H,Lq,R]; gather along R for every key.
  b, lq, h, _ = allrel.shape
  lk = ids.shape[-1]
  gathered = torch.gather(allrel.permute(0, 2, 1, 3), 3,
                          idx.unsqueeze(1).expand(b, h, lq, lk))
  gathered = gathered.permute(0, 2, 3, 1)  # [B,Lq,Lk,H]
  return gathered * valid.unsqueeze(-1).to(allrel.dtype)


def _dropout(p, keep, rate):
  """Attention-probability dropout AFTER the softmax (``training=True``, reference
  ``src/configs/encoders.py:87-88``, ``src/tasks/pretraining.py:279``): ``keep`` is a boolean
  ``[B,Lq,Lk,H]`` mask supplied by the test (the kernels' counter-based mask restated in
  ``tests/dropout_ref.py``), kept entries are scaled by ``1 / (1 - rate)``."""
  if keep is None:
    return p
  return p * keep.to(p.dtype) / (1.0 - rate)


def relative_scores(q, k, att_mask, relative_att_ids, emb, bias, candidate=None, neg=None):
  """Masked, scaled score tensor [B,Lq,Lk,H] (before softmax)."""
  neg = NEG if neg is None else neg
  d = q.shape[-1]
  s = torch.einsum('bqhd,bkhd->bqkh', q, k)
  if relative_att_ids is not None:
    s = s + _lookup(_all_relative_scores(q, emb, bias), relative_att_ids)
  s = s * (1.0 / math.sqrt(d))  # N1: relative term is scaled too.
  if att_mask is not None:
    # The reference adds NEG in fp32, where |x| < 32 is absorbed (ulp(1e9) = 64):
    # a fully-masked row is then *uniform*.  Reproduce that rounding in fp64.
    masked = (s.to(torch.float32) + neg).to(s.dtype)
    s = torch.where(att_mask.bool().unsqueeze(-1), s, masked)
  if candidate is not None:
    s = s.masked_fill(~candidate.unsqueeze(-1), float('-inf'))
  return s


def qkv_relative_attention(q, k, v, att_mask, relative_att_ids, emb, bias, neg=None, keep=None,
                           rate=0.0):
  """Contract (A): QkvRelativeAttention.call [UPSTREAM-RECALLED] (row a2)."""
  s = relative_scores(q, k, att_mask, relative_att_ids, emb, bias, neg=neg)
  p = _dropout(torch.softmax(s, dim=2), keep, rate)
  return torch.einsum('bqkh,bkhd->bqhd', p, v)


def band_to_dense(x_band, fill):
  """[B,L,2r+1] -> [B,L,L]; column k of row i is key j = i + k - r."""
  b, l, w = x_band.shape
  r = (w - 1) // 2
  out = torch.full((b, l, l), fill, dtype=x_band.dtype)
  for kk in range(w):
    off = kk - r
    lo, hi = max(0, -off), min(l, l - off)
    if lo < hi:
      rows = torch.arange(lo, hi)
      out[:, rows, rows + off] = x_band[:, lo:hi, kk]
  return out


def band_candidates(l, r):
  i = torch.arange(l)[:, None]
  j = torch.arange(l)[None, :]
  return (j - i).abs() <= r  # [L,L]


def qkv_relative_local_attention(q, k, v, att_mask, relative_att_ids, emb, bias,
                                 local_radius, side_k=None, side_v=None,
                                 side_att_mask=None, side_relative_att_ids=None, neg=None,
                                 keep=None, rate=0.0):
  """Long rows of contract (B): QkvRelativeLocalAttention.call (row a3).

  Dense-with-band formulation ("full" att_implementation [UPSTREAM-RECALLED]).
  ``att_mask`` / ``relative_att_ids`` are [B,L,2r+1]; side ones [B,L,G].
  """
  b, l = q.shape[0], q.shape[1]
  cand = band_candidates(l, local_radius).unsqueeze(0).expand(b, l, l)
  mask_d = None if att_mask is None else band_to_dense(att_mask, 0)
  ids_d = None if relative_att_ids is None else band_to_dense(relative_att_ids, -1)
  s = relative_scores(q, k, mask_d, ids_d, emb, bias, candidate=cand, neg=neg)
  vv = v
  if side_k is not None:
    s_side = relative_scores(q, side_k, side_att_mask, side_relative_att_ids,
                             emb, bias, neg=neg)
    s = torch.cat([s, s_side], dim=2)
    vv = torch.cat([v, side_v], dim=1)
  p = _dropout(torch.softmax(s, dim=2), keep, rate)   # keep: [B, L, L + G, H]
  return torch.einsum('bqkh,bkhd->bqhd', p, vv)


def global_rows_attention(gq, gk, gv, lk, lv, g2g_mask, g2g_ids, g2l_mask,
                          g2l_ids, emb, bias, neg=None, keep=None, rate=0.0):
  """Global rows of contract (B): keys = all global (+) all long, one softmax."""
  k = torch.cat([gk, lk], dim=1)
  v = torch.cat([gv, lv], dim=1)
  mask = None
  if g2g_mask is not None:
    mask = torch.cat([g2g_mask, g2l_mask], dim=2)
  ids = None
  if g2g_ids is not None:
    ids = torch.cat([g2g_ids, g2l_ids], dim=2)
  return qkv_relative_attention(gq, k, v, mask, ids, emb, bias, neg=neg, keep=keep, rate=rate)


def fused_global_local_attention(lq, lk, lv, gq, gk, gv, side, long_tables,
                                 global_tables, local_radius, neg=None, keep_long=None,
                                 keep_global=None, rate=0.0):
  """Core of FusedGlobalLocalAttention.call (row a4), projections excluded.

  ``side`` is a dict with the eight l2l/l2g/g2g/g2l mask/id tensors; ``*_tables``
  are ``(emb [R,H,d], bias [R,H])``.  Returns ``(long_out, global_out)``.
  """
  long_out = qkv_relative_local_attention(
      lq, lk, lv, side.get('l2l_att_mask'), side.get('l2l_relative_att_ids'),
      long_tables[0], long_tables[1], local_radius, side_k=gk, side_v=gv,
      side_att_mask=side.get('l2g_att_mask'),
      side_relative_att_ids=side.get('l2g_relative_att_ids'), neg=neg, keep=keep_long, rate=rate)
  global_out = global_rows_attention(
      gq, gk, gv, lk, lv, side.get('g2g_att_mask'),
      side.get('g2g_relative_att_ids'), side.get('g2l_att_mask'),
      side.get('g2l_relative_att_ids'), global_tables[0], global_tables[1], neg=neg,
      keep=keep_global, rate=rate)
  return long_out, global_out
