"""ETC's *blocked* global-local attention, restated in PyTorch-CPU fp32.

TEST INFRASTRUCTURE and the timed CPU baseline ("port") of ``bench.py``
(``cpu_baseline`` and ``--impl reference`` legs).  PARITY UNPINNED (see
``oracle/__init__.py``): the reference's TF-CPU path cannot run in this image
(no TensorFlow/etcmodel), so this is a faithful restatement of the published ETC
algorithm [UPSTREAM-RECALLED], the one BASELINE.md section 3 names:

* ``block_len = r + 1``; the long sequence is padded to a multiple of it;
* K, V are expanded to ``[B, nb, 3*block_len, H, d]`` by concatenating the
  previous / current / next block (zeros at both ends);
* row ``q`` of a block sees columns ``q+1 .. q+2r+1`` of the 3-block axis; masks
  and ids ``[B, L, 2r+1]`` are skewed into that frame, everything else is masked;
* relative scores use the one-hot lookup (``use_one_hot_lookup=True``, reference
  ``src/configs/encoders.py:98``): ``onehot(ids, R) . (q E^T + bias)``;
* ONE softmax over ``[3*block_len (+) G]``; split; contract with 3-block V and
  global V.  Global rows: dense attention over ``[G (+) L]``.

Must agree with ``attention_oracle`` (dense fp64) on non-degenerate rows; checked
in ``tests/test_oracle_attention.py``.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NEG = -1e9


def _onehot_rel(q, emb, bias, ids, r_vocab):
  """rel[...,k,h] = sum_r onehot(ids)[...,k,r] * (q E^T + bias)[...,h,r]."""
  allrel = torch.einsum('...qhd,rhd->...qhr', q, emb) + bias.transpose(0, 1)
  onehot = F.one_hot(ids.clamp(0, r_vocab).long(), r_vocab + 1)[..., :r_vocab]
  onehot = onehot * ((ids >= 0) & (ids < r_vocab)).unsqueeze(-1)
  return torch.einsum('...qkr,...qhr->...qkh', onehot.to(q.dtype), allrel)


def _split_blocks(x, bl):
  b, l = x.shape[0], x.shape[1]
  return x.reshape(b, l // bl, bl, *x.shape[2:])


def _concat_3_blocks(xb):
  """[B,nb,bl,...] -> [B,nb,3*bl,...] with zero blocks at both ends."""
  zeros = torch.zeros_like(xb[:, :1])
  prev = torch.cat([zeros, xb[:, :-1]], dim=1)
  nxt = torch.cat([xb[:, 1:], zeros], dim=1)
  return torch.cat([prev, xb, nxt], dim=2)


def _skew_band_to_3blocks(band, bl, fill):
  """[B,nb,bl,2r+1] -> [B,nb,bl,3*bl]: row q's entries land at q+1..q+2r+1."""
  b, nb, _, w = band.shape
  out = torch.full((b, nb, bl, 3 * bl), fill, dtype=band.dtype)
  cols = torch.arange(w)[None, :] + torch.arange(bl)[:, None] + 1  # [bl,w]
  out.scatter_(3, cols[None, None].expand(b, nb, bl, w), band)
  return out


def local_attention_blocked(q, k, v, att_mask, relative_att_ids, emb, bias,
                            local_radius, side_k=None, side_v=None,
                            side_att_mask=None, side_relative_att_ids=None):
  """Long rows, ETC 'sparse' implementation.  q,k,v [B,L,H,d] fp32."""
  b, l, h, d = q.shape
  r = local_radius
  bl = r + 1
  r_vocab = emb.shape[0]
  pad = (-l) % bl
  if pad:
    q = F.pad(q, (0, 0, 0, 0, 0, pad))
    k = F.pad(k, (0, 0, 0, 0, 0, pad))
    v = F.pad(v, (0, 0, 0, 0, 0, pad))
    att_mask = F.pad(att_mask, (0, 0, 0, pad))
    relative_att_ids = F.pad(relative_att_ids, (0, 0, 0, pad))
  qb = _split_blocks(q, bl)                      # [B,nb,bl,H,d]
  k3 = _concat_3_blocks(_split_blocks(k, bl))    # [B,nb,3bl,H,d]
  v3 = _concat_3_blocks(_split_blocks(v, bl))
  mask3 = _skew_band_to_3blocks(_split_blocks(att_mask, bl), bl, 0)
  ids3 = _skew_band_to_3blocks(_split_blocks(relative_att_ids, bl), bl, -1)
  s = torch.einsum('bnqhd,bnkhd->bnqkh', qb, k3)
  s = s + _onehot_rel(qb, emb, bias, ids3, r_vocab)
  s = s * (1.0 / math.sqrt(d))
  s = s + (1.0 - mask3.to(s.dtype)).unsqueeze(-1) * NEG
  if side_k is not None:
    g = side_k.shape[1]
    if pad:
      side_att_mask = F.pad(side_att_mask, (0, 0, 0, pad))
      side_relative_att_ids = F.pad(side_relative_att_ids, (0, 0, 0, pad))
    ss = torch.einsum('bnqhd,bghd->bnqgh', qb, side_k)
    ss = ss + _onehot_rel(qb, emb, bias,
                          _split_blocks(side_relative_att_ids, bl), r_vocab)
    ss = ss * (1.0 / math.sqrt(d))
    ss = ss + (1.0 - _split_blocks(side_att_mask, bl).to(ss.dtype)).unsqueeze(-1) * NEG
    s = torch.cat([s, ss], dim=3)
  p = torch.softmax(s, dim=3)
  out = torch.einsum('bnqkh,bnkhd->bnqhd', p[:, :, :, :3 * bl], v3)
  if side_k is not None:
    out = out + torch.einsum('bnqgh,bghd->bnqhd', p[:, :, :, 3 * bl:], side_v)
  return out.reshape(b, l + pad, h, d)[:, :l]


def dense_attention_onehot(q, k, v, att_mask, relative_att_ids, emb, bias):
  """QkvRelativeAttention with the one-hot lookup, fp32 (contract A)."""
  d = q.shape[-1]
  s = torch.einsum('bqhd,bkhd->bqkh', q, k)
  s = s + _onehot_rel(q, emb, bias, relative_att_ids, emb.shape[0])
  s = s * (1.0 / math.sqrt(d))
  s = s + (1.0 - att_mask.to(s.dtype)).unsqueeze(-1) * NEG
  p = torch.softmax(s, dim=2)
  return torch.einsum('bqkh,bkhd->bqhd', p, v)


def fused_global_local_blocked(lq, lk, lv, gq, gk, gv, side, long_tables,
                               global_tables, local_radius):
  """Both halves of FusedGlobalLocalAttention's core, ETC's way."""
  long_out = local_attention_blocked(
      lq, lk, lv, side['l2l_att_mask'], side['l2l_relative_att_ids'],
      long_tables[0], long_tables[1], local_radius, side_k=gk, side_v=gv,
      side_att_mask=side['l2g_att_mask'],
      side_relative_att_ids=side['l2g_relative_att_ids'])
  k = torch.cat([gk, lk], dim=1)
  v = torch.cat([gv, lv], dim=1)
  mask = torch.cat([side['g2g_att_mask'], side['g2l_att_mask']], dim=2)
  ids = torch.cat([side['g2g_relative_att_ids'], side['g2l_relative_att_ids']], dim=2)
  global_out = dense_attention_onehot(gq, k, v, mask, ids, global_tables[0],
                                      global_tables[1])
  return long_out, global_out
