"""Oracle self-consistency: dense fp64 == ETC blocked restatement; gradients == finite differences."""
import numpy as np
import pytest
import torch

from oracle import attention_oracle as ao
from oracle import blocked_etc as be
from oracle import feature_oracle as fo


def _inputs(b, l, g, h, d, r, rv, dist, seed, dtype=torch.float64):
  gen = torch.Generator().manual_seed(seed)
  rn = lambda *s, std=1.0: (torch.randn(*s, generator=gen, dtype=torch.float64) * std).to(dtype)
  lengths = torch.randint(max(1, l // 2), l + 1, (b,), generator=gen)
  le = (torch.arange(l)[None] < lengths[:, None]).int().numpy()
  sid = ((torch.arange(l) * g) // l)[None].expand(b, l).int().numpy()
  ge = np.ones((b, g), dtype=np.int32)
  ge[:, -1] = 0
  side = {k: torch.tensor(v) for k, v in
          fo.make_global_local_side_inputs(le, ge, sid, r, dist).items()}
  t = dict(lq=rn(b, l, h, d), lk=rn(b, l, h, d), lv=rn(b, l, h, d),
           gq=rn(b, g, h, d), gk=rn(b, g, h, d), gv=rn(b, g, h, d))
  lt = (rn(rv, h, d, std=0.3), rn(rv, h, std=0.3))
  gt = (rn(rv, h, d, std=0.3), rn(rv, h, std=0.3))
  return t, side, lt, gt


@pytest.mark.parametrize('l,g,r', [(24, 4, 3), (37, 5, 4), (16, 3, 7), (9, 2, 12), (65, 8, 8)])
def test_blocked_equals_dense(l, g, r):
  t, side, lt, gt = _inputs(2, l, g, 3, 8, r, 16, 3, seed=l * 7 + r)
  lo, go = ao.fused_global_local_attention(t['lq'], t['lk'], t['lv'], t['gq'], t['gk'],
                                           t['gv'], side, lt, gt, r)
  f32 = lambda x: x.float()
  lb, gb = be.fused_global_local_blocked(
      *(f32(t[k]) for k in ('lq', 'lk', 'lv', 'gq', 'gk', 'gv')), side,
      tuple(map(f32, lt)), tuple(map(f32, gt)), r)
  assert torch.allclose(lb.double(), lo, atol=2e-5, rtol=1e-5)
  assert torch.allclose(gb.double(), go, atol=2e-5, rtol=1e-5)


def test_out_of_vocabulary_id_contributes_zero():
  gen = torch.Generator().manual_seed(5)
  rn = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
  q, k, v = rn(1, 6, 2, 4), rn(1, 6, 2, 4), rn(1, 6, 2, 4)
  emb, bias = rn(8, 2, 4), rn(8, 2)
  mask = torch.ones(1, 6, 6, dtype=torch.int32)
  ids = torch.tensor(fo.make_relative_att_ids_1d(6, 2))[None]
  out_ref = ao.qkv_relative_attention(q, k, v, mask, ids, emb, bias)
  # replace id 0 by an OOV id; equivalent to zeroing table row 0
  emb0, bias0 = emb.clone(), bias.clone()
  emb0[0] = 0
  bias0[0] = 0
  want = ao.qkv_relative_attention(q, k, v, mask, ids, emb0, bias0)
  for oov in (99, -1, 8):
    ids_oov = torch.where(ids == 0, torch.full_like(ids, oov), ids)
    got = ao.qkv_relative_attention(q, k, v, mask, ids_oov, emb, bias)
    assert torch.allclose(got, want, atol=1e-12)
  assert not torch.allclose(want, out_ref, atol=1e-6)


def test_fully_masked_row_is_uniform_over_candidates():
  t, side, lt, gt = _inputs(1, 10, 2, 1, 4, 2, 8, 2, seed=9)
  side = dict(side)
  side['l2l_att_mask'] = side['l2l_att_mask'].clone()
  side['l2g_att_mask'] = side['l2g_att_mask'].clone()
  side['l2l_att_mask'][0, 0] = 0
  side['l2g_att_mask'][0, 0] = 0
  out = ao.qkv_relative_local_attention(
      t['lq'], t['lk'], t['lv'], side['l2l_att_mask'], side['l2l_relative_att_ids'], *lt, 2,
      side_k=t['gk'], side_v=t['gv'], side_att_mask=side['l2g_att_mask'],
      side_relative_att_ids=side['l2g_relative_att_ids'])
  # row 0 candidates: long keys 0..2 and both global keys
  want = (t['lv'][0, 0:3].sum(0) + t['gv'][0].sum(0)) / 5.0
  assert torch.allclose(out[0, 0], want, atol=1e-9)


def test_oracle_gradients_match_finite_differences():
  t, side, lt, gt = _inputs(1, 7, 2, 2, 4, 2, 8, 2, seed=11)
  leaves = [t[k].clone().requires_grad_() for k in ('lq', 'lk', 'lv', 'gq', 'gk', 'gv')]
  tabs = [x.clone().requires_grad_() for x in (*lt, *gt)]

  def f(*xs):
    lo, go = ao.fused_global_local_attention(*xs[:6], side, (xs[6], xs[7]), (xs[8], xs[9]), 2)
    return lo, go

  assert torch.autograd.gradcheck(f, (*leaves, *tabs), eps=1e-6, atol=1e-6)
