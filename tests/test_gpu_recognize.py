"""Recognition of generator-shaped explicit side inputs (mlt_*_compact_from_explicit): the descriptors it
returns must reproduce every element of the tensors it was given, anything else must be refused, and the
layer stacks must give the same results from explicit tensors as from the descriptors."""
import dataclasses
import os
import sys

import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import layers, ops, synthetic

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_parity import NAMES  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _gl_case(batch=3, long_len=320, global_len=20, radius=16, dist=5, packed=True):
  """Ragged, packed examples: two or three examples per row plus padding."""
  g = torch.Generator().manual_seed(long_len + radius)
  le = torch.zeros(batch, long_len, dtype=torch.int32)
  ge = torch.zeros(batch, global_len, dtype=torch.int32)
  sid = ((torch.arange(long_len, dtype=torch.int64) * global_len) // long_len).to(torch.int32)
  per = long_len // global_len
  for b in range(batch):
    cuts = sorted(torch.randint(1, global_len, (2,), generator=g).tolist()) if packed else [global_len, global_len]
    valid = int(torch.randint(global_len // 2, global_len + 1, (1,), generator=g))
    for s in range(global_len):
      e = 0 if s >= valid else (7 if s < cuts[0] else (3 if s < cuts[1] else 9))
      ge[b, s] = e
      le[b, s * per:(s + 1) * per] = e
  return fu.CompactSideInputs(le.to(DEV), ge.to(DEV), sid[None].expand(batch, -1).contiguous().to(DEV), dist), radius


def test_gl_recognition_round_trip():
  for kw in (dict(), dict(batch=2, long_len=4096, global_len=256, radius=64, dist=12),
             dict(batch=1, long_len=96, global_len=6, radius=100, dist=3, packed=False)):
    compact, radius = _gl_case(**kw)
    explicit = ops.build_gl_side_inputs(compact, radius)
    got = ops.compact_from_explicit_gl(explicit, radius)
    assert got is not None
    assert got.relative_pos_max_distance == compact.relative_pos_max_distance
    again = ops.build_gl_side_inputs(got, radius)
    for k, v in explicit.items():
      assert torch.equal(v, again[k]), k
    # sentence ids are recovered as they were (every long token has its global token here)
    assert torch.equal(got.sentence_ids, compact.sentence_ids)


def test_gl_recognition_of_the_benchmark_descriptors():
  """synthetic.make_descriptors: a padding tail may own no global token; its tokens still see one another."""
  _, shape = synthetic.CONFIGS['c3_4096']
  shape = dataclasses.replace(shape, batch=4)
  x = synthetic.make_inputs(shape, seed=1238, dtype=torch.bfloat16)
  compact = fu.CompactSideInputs(x['long_example_ids'].to(DEV), x['global_example_ids'].to(DEV),
                                 x['sentence_ids'].to(DEV), shape.max_distance)
  explicit = ops.build_gl_side_inputs(compact, shape.local_radius)
  got = ops.compact_from_explicit_gl(explicit, shape.local_radius)
  assert got is not None
  again = ops.build_gl_side_inputs(got, shape.local_radius)
  for k, v in explicit.items():
    assert torch.equal(v, again[k]), k
  # and the operator gives bit-identical results from either set of descriptors
  dev_in = [x[n].to(DEV) for n in NAMES]
  with torch.no_grad():
    a = ops.global_local_attention(*dev_in, local_radius=shape.local_radius, side=compact)
    b = ops.global_local_attention(*dev_in, local_radius=shape.local_radius, side=got)
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_gl_recognition_refuses_anything_else():
  compact, radius = _gl_case()
  explicit = ops.build_gl_side_inputs(compact, radius)
  b, l, w = explicit['l2l_att_mask'].shape
  for k, v in explicit.items():
    bad = dict(explicit)
    t = v.clone()
    # an element the descriptors cannot reproduce: interior, last batch row
    idx = (b - 1, t.shape[1] // 2, t.shape[2] // 2)
    t[idx] = t[idx] + 1
    bad[k] = t
    assert ops.compact_from_explicit_gl(bad, radius) is None, k
  # a padding convention the compact rule does not have: padded long rows see nothing at all
  bad = dict(explicit)
  t = explicit['l2l_att_mask'].clone()
  t[compact.long_example_ids == 0] = 0
  if not torch.equal(t, explicit['l2l_att_mask']):
    bad['l2l_att_mask'] = t
    assert ops.compact_from_explicit_gl(bad, radius) is None
  # missing tensors: nothing to recognise
  bad = dict(explicit)
  bad['g2l_relative_att_ids'] = None
  assert ops.compact_from_explicit_gl(bad, radius) is None


def _dense_case(batch=4, seq=200, packed=True):
  g = torch.Generator().manual_seed(seq)
  eid = torch.zeros(batch, seq, dtype=torch.int32)
  for b in range(batch):
    n = int(torch.randint(seq // 2, seq + 1, (1,), generator=g))
    cut = int(torch.randint(1, n, (1,), generator=g)) if packed else n
    eid[b, :cut] = 5
    eid[b, cut:n] = 2
  return eid.to(DEV)


def test_dense_recognition_1d_and_2d():
  eid = _dense_case()
  # 1-D rule, distance read from the ids; ids given once for the whole batch, as the reference does
  mask, ids = ops.build_dense_side_inputs(eid, 12)
  got = ops.compact_from_explicit_dense(mask, ids[:1])
  assert got is not None and got.max_distance == 12 and got.num_patch_per_row == 0
  m2, i2 = ops.build_dense_side_inputs(got.q_example_ids, got.max_distance)
  assert torch.equal(m2, mask) and torch.equal(i2, ids)
  assert torch.equal(got.q_example_ids, got.k_example_ids)
  # 2-D image + text layout (config 2: 14 x 14 patches), named by the caller
  eid = _dense_case(batch=2, seq=512, packed=False)
  mask, ids = ops.build_dense_side_inputs(eid, 12, num_patch_per_row=14, num_core_layers=3)
  got = ops.compact_from_explicit_dense(mask, ids, 14, 3, 12)
  assert got is not None and got.num_patch_per_row == 14
  m2, i2 = ops.build_dense_side_inputs(got.q_example_ids, 12, 14, 3)
  assert torch.equal(m2, mask) and torch.equal(i2, ids)
  # the 2-D ids are not the 1-D rule's, a wrong layout is refused, and so is one changed element
  assert ops.compact_from_explicit_dense(mask, ids) is None
  assert ops.compact_from_explicit_dense(mask, ids, 14, 2, 12) is None
  for t in (mask, ids):
    u = t.clone()
    u[1, 300, 17] += 1
    args = (u, ids) if t is mask else (mask, u)
    assert ops.compact_from_explicit_dense(*args, 14, 3, 12) is None
  # a mask that is not an equality of labels (causal)
  causal = torch.tril(torch.ones(512, 512, dtype=torch.int32, device=DEV))[None].expand(2, -1, -1).contiguous()
  assert ops.compact_from_explicit_dense(causal, ids, 14, 3, 12) is None


def test_stacks_give_the_same_results_from_explicit_tensors():
  torch.manual_seed(0)
  compact, radius = _gl_case(batch=2, long_len=256, global_len=16, radius=64, dist=12)
  explicit = ops.build_gl_side_inputs(compact, radius)
  stack = layers.GlobalLocalTransformerLayers(128, 128, 2, 2, radius, relative_vocab_size=32).to(DEV).bfloat16().eval()
  xl = torch.randn(2, 256, 128, device=DEV).bfloat16()
  xg = torch.randn(2, 16, 128, device=DEV).bfloat16()
  with torch.no_grad():
    a = stack(xl, xg, compact_side_inputs=compact)
    b = stack(xl, xg, **explicit)                       # recognised once, then the compact kernels
    stack.recognize_side_inputs = False
    c = stack(xl, xg, **explicit)                       # the explicit kernels
  for u, v, w in zip(a, b, c):
    assert torch.equal(u, v)
    assert (u.float() - w.float()).abs().max().item() < 3e-2
  # dense stack, the reference's own signature (mmt_encoder.py:220-224)
  eid = _dense_case(batch=2, seq=256)
  mask, ids = ops.build_dense_side_inputs(eid, 12)
  dense = layers.RelativeTransformerLayers(128, 2, 2, relative_vocab_size=32).to(DEV).bfloat16().eval()
  x = torch.randn(2, 256, 128, device=DEV).bfloat16()
  with torch.no_grad():
    a = dense(x, compact=ops.DenseCompactSideInputs(eid, max_distance=12))
    b = dense(x, att_mask=mask, relative_att_ids=ids)
    dense.recognize_side_inputs = False
    c = dense(x, att_mask=mask, relative_att_ids=ids)
  assert torch.equal(a, b)
  assert (a.float() - c.float()).abs().max().item() < 3e-2


@pytest.mark.parametrize('seed', range(6))
def test_gl_recognition_random_packings(seed):
  """Random packings: 1-4 examples per row of random lengths, each owning a random number of global tokens,
  random padding tail, random radius / distance.  Generator-shaped inputs are always recognised and the
  descriptors always rebuild the very same eight tensors."""
  g = torch.Generator().manual_seed(1000 + seed)
  ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
  batch, long_len, global_len = ri(1, 3), 32 * ri(3, 12), ri(4, 24)
  radius, dist = ri(1, 40), ri(1, 9)
  le = torch.zeros(batch, long_len, dtype=torch.int32)
  ge = torch.zeros(batch, global_len, dtype=torch.int32)
  sid = torch.full((batch, long_len), -1, dtype=torch.int32)
  for b in range(batch):
    n_ex = ri(1, min(4, global_len))
    lcuts = sorted(torch.randperm(long_len - 1, generator=g)[:n_ex].add(1).tolist()) + [long_len]
    gcuts = sorted(torch.randperm(global_len - 1, generator=g)[:n_ex - 1].add(1).tolist()) + [global_len]
    lo = go = 0
    for e in range(n_ex):                     # example e: long tokens [lo, lcuts[e]), global tokens [go, gcuts[e])
      le[b, lo:lcuts[e]] = 11 + 3 * e
      ge[b, go:gcuts[e]] = 11 + 3 * e
      span, ng = lcuts[e] - lo, gcuts[e] - go
      for t in range(span):                   # sentences of the example: its long tokens spread over its global tokens
        sid[b, lo + t] = go + min(ng - 1, t * ng // max(span, 1))
      lo, go = lcuts[e], gcuts[e]
  compact = fu.CompactSideInputs(le.to(DEV), ge.to(DEV), sid.to(DEV), dist)
  explicit = ops.build_gl_side_inputs(compact, radius)
  got = ops.compact_from_explicit_gl(explicit, radius)
  assert got is not None, (batch, long_len, global_len, radius, dist)
  assert got.relative_pos_max_distance == dist
  again = ops.build_gl_side_inputs(got, radius)
  for k, v in explicit.items():
    assert torch.equal(v, again[k]), k


@pytest.mark.parametrize('seed', range(4))
def test_dense_recognition_random_segments(seed):
  """Random segmentations of the sequence (1-5 segments + optional padding), 1-D rule with a random distance and
  the 2-D layout with a random patch grid: recognised, and the descriptors rebuild the very same tensors."""
  g = torch.Generator().manual_seed(2000 + seed)
  ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
  batch, seq = ri(1, 4), ri(40, 300)
  eid = torch.zeros(batch, seq, dtype=torch.int32)
  for b in range(batch):
    cuts = sorted(torch.randperm(seq - 1, generator=g)[:ri(1, 5)].add(1).tolist()) + [seq]
    lo = 0
    for e, hi in enumerate(cuts[:-1] if ri(0, 1) else cuts):      # sometimes the last piece stays padding (id 0)
      eid[b, lo:hi] = 5 + 2 * e
      lo = hi
  eid = eid.to(DEV)
  dist = ri(1, 15)
  mask, ids = ops.build_dense_side_inputs(eid, dist)
  got = ops.compact_from_explicit_dense(mask, ids)
  assert got is not None and got.max_distance == dist
  m2, i2 = ops.build_dense_side_inputs(got.q_example_ids, dist)
  assert torch.equal(m2, mask) and torch.equal(i2, ids)
  npr = ri(2, 6)
  if npr * npr <= seq:
    core = ri(1, 3)
    mask, ids = ops.build_dense_side_inputs(eid, dist, num_patch_per_row=npr, num_core_layers=core)
    got = ops.compact_from_explicit_dense(mask, ids, npr, core, dist)
    assert got is not None
    m2, i2 = ops.build_dense_side_inputs(got.q_example_ids, dist, npr, core)
    assert torch.equal(m2, mask) and torch.equal(i2, ids)
    assert ops.compact_from_explicit_dense(mask, ids, npr, core, dist + 1) is None
