"""GPU parity, round 2: the benchmarked configurations against the oracle, the ABI's `neg`,
attention-probability dropout, broadcast views.

Tolerances are BASELINE.json's north star: bf16 within 2e-2 max-abs against an fp32/fp64 reference,
fp32 within 1e-5 relative.  Gradient tensors whose magnitude exceeds 1 are compared relative to
their maximum (the worst unscaled error per tensor is written to gpurun_out/ for the record).
"""
import dataclasses
import json
import os
import sys

import numpy as np
import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import ops, synthetic, _lib
from oracle import attention_oracle as ao
from oracle import blocked_etc as be
from oracle import feature_oracle as fo

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropout_ref  # noqa: E402
from test_gpu_parity import (BF16_ABS, FP32_REL, NAMES, abs_err, compact_of, oracle_side, rel_err,  # noqa: E402
                             run_cuda_gl, run_oracle_gl)

pytestmark = pytest.mark.gpu


def _record(name, payload):
  out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
  try:
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f'parity_{name}.json'), 'w') as f:
      json.dump(payload, f, indent=1)
  except OSError:
    pass


def run_cuda(x, shape, side, **kw):
  dev = torch.device('cuda')
  dev_in = [x[n].to(dev).requires_grad_() for n in NAMES]
  lo, go = ops.global_local_attention(*dev_in, local_radius=shape.local_radius, side=side, **kw)
  loss = (lo.float() * x['d_long_out'].to(dev).float()).sum() + \
         (go.float() * x['d_global_out'].to(dev).float()).sum()
  loss.backward()
  torch.cuda.synchronize()
  return lo, go, [t.grad for t in dev_in]


# ------------------------------------------------------------------------------------------
# 1. The benchmarked configurations (BASELINE.json configs[2]) against the oracle

def _blocked_reference(x, shape, side):
  """oracle.blocked_etc (ETC's blocked algorithm, fp32, one-hot lookup) fwd + autograd bwd on the CPU."""
  ref_in = [x[n].float().requires_grad_() for n in NAMES]
  rl, rg = be.fused_global_local_blocked(*ref_in[:6], side, (ref_in[6], ref_in[7]), (ref_in[8], ref_in[9]),
                                         shape.local_radius)
  ((rl * x['d_long_out'].float()).sum() + (rg * x['d_global_out'].float()).sum()).backward()
  return rl.detach(), rg.detach(), [t.grad for t in ref_in]


@pytest.mark.parametrize('name', ['c3_2048', 'c3_4096', 'c3_8192'])
def test_benchmarked_configs_match_blocked_oracle(name):
  """L = 2048 / 4096 / 8192, G = L / 16, r = 64, all 12 heads, bf16, fwd + bwd at batch 1 (units are
  independent; test_batch16_slices_bit_identical covers the benchmarked launch geometry) against the
  blocked restatement evaluated in fp32 on the same bf16-rounded inputs."""
  seed_off, shape = synthetic.CONFIGS[name]
  shape = dataclasses.replace(shape, batch=1)
  x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16)
  # explicit tensors for the oracle from the (bit-exact tested) device constructor: the loop oracle is slow here
  side = {k: v.cpu() for k, v in ops.build_gl_side_inputs(compact_of(x, shape), shape.local_radius).items()}
  torch.set_num_threads(max(1, os.cpu_count() or 1))
  rl, rg, rgrads = _blocked_reference(x, shape, side)
  lo, go, grads = run_cuda_gl(x, shape, compact_of(x, shape), impl='tc')
  report = {'long_out': abs_err(lo, rl.double()), 'global_out': abs_err(go, rg.double())}
  assert report['long_out'] < BF16_ABS and report['global_out'] < BF16_ABS, report
  for n, got, want in zip(NAMES, grads, rgrads):
    want = want.double()
    err, mag = abs_err(got, want), want.abs().max().item()
    report['d_' + n] = {'max_abs_err': err, 'max_abs_ref': mag}
    assert err < BF16_ABS * max(1.0, mag), (n, err, mag)
  _record(name, report)


def test_batch16_slices_bit_identical():
  """The benchmarked launch (c3_4096: B = 16, 6144-CTA grids): every batch element of the B = 16 run is
  bit-identical to the same element run alone (deterministic kernels, no cross-unit interaction)."""
  seed_off, shape = synthetic.CONFIGS['c3_4096']
  x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16)
  lo, go, grads = run_cuda_gl(x, shape, compact_of(x, shape), impl='tc')
  one = dataclasses.replace(shape, batch=1)
  per_example = set(NAMES[:6]) | {'long_example_ids', 'global_example_ids', 'sentence_ids', 'd_long_out',
                                  'd_global_out'}
  for b in (0, 7, 15):
    xb = {k: (v[b:b + 1] if k in per_example else v) for k, v in x.items()}
    lob, gob, gb = run_cuda_gl(xb, one, compact_of(xb, one), impl='tc')
    assert torch.equal(lo[b:b + 1], lob) and torch.equal(go[b:b + 1], gob), b
    for n, full, single in zip(NAMES[:6], grads[:6], gb[:6]):
      assert torch.equal(full[b:b + 1], single), (n, b)


def test_tc_repeat_runs_are_bit_identical_including_table_gradients():
  """The persistent kernels hand their tiles out dynamically (whichever CTA is free takes the next one): no
  result may depend on that.  Two runs of the benchmarked shape (batch 4) agree bit for bit in every output
  and gradient, the table gradients (summed over batch and tiles in a fixed order) included."""
  seed_off, shape = synthetic.CONFIGS['c3_4096']
  shape = dataclasses.replace(shape, batch=4)
  x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16)
  side = compact_of(x, shape)
  a = run_cuda_gl(x, shape, side, impl='tc')
  for _ in range(2):
    b = run_cuda_gl(x, shape, side, impl='tc')
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    for n, ga, gb in zip(NAMES, a[2], b[2]):
      assert torch.equal(ga, gb), n


def test_concurrent_callers_on_their_own_streams():
  """Two host threads, each on its own CUDA stream, run forward + backward at the same time (TF's inter-op pool
  does this): the library's per-device state -- side-stream pool, tensor-map cache, the tile counters of the
  persistent kernels -- must not let the calls disturb one another.  Every result equals the serial one."""
  import threading
  seed_off, shape = synthetic.CONFIGS['c3_4096']
  shape = dataclasses.replace(shape, batch=2)
  cases = []
  for k in range(2):
    x = synthetic.make_inputs(shape, seed=77 + k, dtype=torch.bfloat16)
    cases.append((x, compact_of(x, shape), run_cuda_gl(x, shape, compact_of(x, shape), impl='tc')))
  errors = []

  def worker(k):
    try:
      x, side, want = cases[k]
      stream = torch.cuda.Stream()
      with torch.cuda.stream(stream):
        for _ in range(6):
          got = run_cuda_gl(x, shape, side, impl='tc')
          assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
          for n, ga, gb in zip(NAMES, got[2], want[2]):
            assert torch.equal(ga, gb), n
    except BaseException as e:   # noqa: BLE001  (reported in the main thread)
      errors.append((k, repr(e)))

  threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
  for t in threads:
    t.start()
  for t in threads:
    t.join()
  assert not errors, errors


# ------------------------------------------------------------------------------------------
# 2. The ABI's `neg` is honoured by both kernel families

def _masked_rows_case(dtype):
  shape = synthetic.GlobalLocalShape(1, 200, 8, 2, 64, 20, 16, 3)
  x = synthetic.make_inputs(shape, seed=5, dtype=dtype)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).to(dtype)
  side = oracle_side(x, shape)
  for row in (7, 64, 150):
    side['l2l_att_mask'][:, row, :] = 0           # fully masked long rows
    side['l2g_att_mask'][:, row, :] = 0
  side['g2g_att_mask'][:, 3, :] = 0               # and a fully masked global row
  side['g2l_att_mask'][:, 3, :] = 0
  side['l2g_att_mask'][:, 20:40, 2] = 0           # partially masked rows
  return shape, x, side


def _oracle_neg(x, side, shape, neg):
  ref_in = [x[n].double().requires_grad_() for n in NAMES]
  rl, rg = ao.fused_global_local_attention(*ref_in[:6], side, (ref_in[6], ref_in[7]), (ref_in[8], ref_in[9]),
                                           shape.local_radius, neg=neg)
  ((rl * x['d_long_out'].double()).sum() + (rg * x['d_global_out'].double()).sum()).backward()
  return rl.detach(), rg.detach(), [t.grad for t in ref_in]


@pytest.mark.parametrize('neg', [-1e9, -1e4])
def test_tc_honours_neg_on_fully_masked_rows(neg):
  """neg = -1e9: a masked score is the constant itself in fp32 -> a fully-masked row is uniform.
  neg = -1e4 (BERT-style adder): a fully-masked row is softmax(s).  The tcgen05 path must follow the
  ABI's `neg` either way (fwd + bwd), like the oracle evaluated with the same constant."""
  shape, x, side = _masked_rows_case(torch.bfloat16)
  rl, rg, rgrads = _oracle_neg(x, side, shape, neg)
  if neg == -1e4:   # the two constants really differ on the fully-masked rows
    rl9, _, _ = _oracle_neg(x, side, shape, -1e9)
    assert (rl9[:, 7] - rl[:, 7]).abs().max().item() > 0.05
  lo, go, grads = run_cuda(x, shape, {k: v.cuda() for k, v in side.items()}, impl='tc', neg=neg)
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    assert abs_err(got, want) < BF16_ABS * max(1.0, want.abs().max().item()), name


@pytest.mark.parametrize('neg', [-1e9, -1e4])
def test_simt_honours_neg_fp32(neg):
  shape, x, side = _masked_rows_case(torch.float32)
  rl, rg, rgrads = _oracle_neg(x, side, shape, neg)
  lo, go, grads = run_cuda(x, shape, {k: v.cuda() for k, v in side.items()}, impl='simt', neg=neg)
  # neg = -1e4 keeps 1e-3-ulp rounding of (s - 1e4) in fp32 on the fully-masked rows
  tol = FP32_REL if neg == -1e9 else 2e-3
  assert rel_err(lo, rl) < tol and rel_err(go, rg) < tol
  for name, got, want in zip(NAMES, grads, rgrads):
    assert rel_err(got, want) < tol, name


def test_tc_compact_neg_1e4_matches_oracle():
  """Compact descriptors with ragged lengths (masked groups, mask changes inside groups) at neg = -1e4."""
  shape = synthetic.GlobalLocalShape(2, 448, 40, 2, 64, 64, 32, 12)
  x = synthetic.make_inputs(shape, seed=21, dtype=torch.bfloat16)
  side = oracle_side(x, shape)
  rl, rg, rgrads = _oracle_neg(x, side, shape, -1e4)
  lo, go, grads = run_cuda(x, shape, compact_of(x, shape), impl='tc', neg=-1e4)
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    assert abs_err(got, want) < BF16_ABS * max(1.0, want.abs().max().item()), name


# ------------------------------------------------------------------------------------------
# 3. Attention-probability dropout (reference default 0.1: src/configs/encoders.py:87-88)

def _gl_keep(seed, p, shape):
  b, l, g, h = shape.batch, shape.long_len, shape.global_len, shape.heads
  keep_long = torch.from_numpy(dropout_ref.keep_mask(seed, p, b, h, l, l + g, rowset=0))
  keep_glob = torch.from_numpy(dropout_ref.keep_mask(seed, p, b, h, g, g + l, rowset=1))
  return keep_long, keep_glob


def _oracle_dropout(x, side, shape, seed, p):
  keep_long, keep_glob = _gl_keep(seed, p, shape)
  ref_in = [x[n].double().requires_grad_() for n in NAMES]
  rl, rg = ao.fused_global_local_attention(*ref_in[:6], side, (ref_in[6], ref_in[7]), (ref_in[8], ref_in[9]),
                                           shape.local_radius, keep_long=keep_long, keep_global=keep_glob, rate=p)
  ((rl * x['d_long_out'].double()).sum() + (rg * x['d_global_out'].double()).sum()).backward()
  return rl.detach(), rg.detach(), [t.grad for t in ref_in]


@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_dropout_simt_fp32_matches_oracle_with_same_mask(mode):
  shape = synthetic.GlobalLocalShape(2, 200, 8, 3, 64, 20, 32, 12)
  x = synthetic.make_inputs(shape, seed=31)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = x[n] * 10
  side = oracle_side(x, shape)
  seed, p = 0x1234567890ABCDEF, 0.1
  rl, rg, rgrads = _oracle_dropout(x, side, shape, seed, p)
  cuda_side = compact_of(x, shape) if mode == 'compact' else {k: v.cuda() for k, v in side.items()}
  lo, go, grads = run_cuda(x, shape, cuda_side, impl='simt', dropout_p=p, dropout_seed=seed)
  assert rel_err(lo, rl) < FP32_REL and rel_err(go, rg) < FP32_REL
  for name, got, want in zip(NAMES, grads, rgrads):
    assert rel_err(got, want) < FP32_REL, name
  # and dropout really happened: the result differs from the p = 0 result
  lo0, _, _ = run_cuda(x, shape, cuda_side, impl='simt')
  assert (lo0 - lo).abs().max().item() > 1e-3


TC_DROP_SHAPES = [
    (2, 512, 32, 4, 64, 32, 12),
    (1, 200, 8, 2, 64, 32, 12),
    (1, 1024, 64, 2, 100, 64, 30),
    (1, 1100, 40, 2, 64, 32, 12),
]


@pytest.mark.parametrize('dims', TC_DROP_SHAPES, ids=lambda d: 'x'.join(map(str, d)))
@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_dropout_tc_matches_oracle_with_same_mask(dims, mode):
  """Every tcgen05 kernel (forward, query-centric and key-centric backward, all launch configurations)
  regenerates the same keep mask: fwd + bwd against the fp64 oracle that is handed the identical mask
  (tests/dropout_ref.py), p = 0.1."""
  b, l, g, h, r, rv, dist = dims
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=l + 3, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  side = oracle_side(x, shape)
  seed, p = 987654321012345, 0.1
  rl, rg, rgrads = _oracle_dropout(x, side, shape, seed, p)
  cuda_side = compact_of(x, shape) if mode == 'compact' else {k: v.cuda() for k, v in side.items()}
  lo, go, grads = run_cuda(x, shape, cuda_side, impl='tc', dropout_p=p, dropout_seed=seed)
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    assert abs_err(got, want) < BF16_ABS * max(1.0, want.abs().max().item()), name


def test_dropout_zero_is_bit_identical_and_seed_is_replayed():
  shape = synthetic.GlobalLocalShape(2, 320, 16, 2, 64, 64, 32, 12)
  x = synthetic.make_inputs(shape, seed=9, dtype=torch.bfloat16)
  side = compact_of(x, shape)
  a = run_cuda(x, shape, side, impl='tc')
  z = run_cuda(x, shape, side, impl='tc', dropout_p=0.0, dropout_seed=77)
  assert torch.equal(a[0], z[0]) and all(torch.equal(u, v) for u, v in zip(a[2], z[2]))
  d1 = run_cuda(x, shape, side, impl='tc', dropout_p=0.1, dropout_seed=5)
  d2 = run_cuda(x, shape, side, impl='tc', dropout_p=0.1, dropout_seed=5)
  d3 = run_cuda(x, shape, side, impl='tc', dropout_p=0.1, dropout_seed=6)
  assert torch.equal(d1[0], d2[0]) and all(torch.equal(u, v) for u, v in zip(d1[2], d2[2]))
  assert not torch.equal(d1[0], d3[0])
  # keep rate through the op itself: V = 1 makes every output element sum_j keep_ij p_ij / (1 - p)
  ones = dict(x)
  ones['long_v'] = torch.ones_like(x['long_v'])
  ones['global_v'] = torch.ones_like(x['global_v'])
  lo, go, _ = run_cuda(ones, shape, side, impl='tc', dropout_p=0.1, dropout_seed=5)
  assert abs(lo.float().mean().item() - 1.0) < 0.02 and abs(go.float().mean().item() - 1.0) < 0.02
  assert lo.float().std().item() > 0.01   # not all ones: entries were really dropped
  with pytest.raises(ValueError):
    ops.global_local_attention(*[x[n].cuda() for n in NAMES], local_radius=64, side=side, dropout_p=1.0)


def test_dropout_dense_tc_and_simt_match_oracle():
  b, s, h, d, p, seed = 2, 200, 2, 64, 0.1, 424242
  gen = torch.Generator().manual_seed(8)
  q, k, v, do = (torch.randn(b, s, h, d, generator=gen).bfloat16() for _ in range(4))
  emb = (torch.randn(32, h, d, generator=gen) * 0.2).bfloat16()
  bias = (torch.randn(32, h, generator=gen) * 0.2).bfloat16()
  e = (torch.arange(s)[None] < torch.tensor([[200], [150]])).int()
  mask = torch.tensor(fo.make_segmented_att_mask(e.numpy()))
  ids = torch.tensor(fo.make_relative_att_ids_1d(s, 12))[None].expand(b, s, s).contiguous()
  keep = torch.from_numpy(dropout_ref.keep_mask(seed, p, b, h, s, s, rowset=0))
  ref = [t.double().requires_grad_() for t in (q, k, v, emb, bias)]
  ro = ao.qkv_relative_attention(ref[0], ref[1], ref[2], mask, ids, ref[3], ref[4], keep=keep, rate=p)
  (ro * do.double()).sum().backward()
  for impl in ('tc', 'simt'):
    for kwargs in (dict(att_mask=mask.cuda(), relative_att_ids=ids.cuda()),
                   dict(compact=ops.DenseCompactSideInputs(e.cuda(), max_distance=12))):
      dev = [t.cuda().requires_grad_() for t in (q, k, v, emb, bias)]
      out = ops.dense_relative_attention(*dev, impl=impl, dropout_p=p, dropout_seed=seed, **kwargs)
      (out.float() * do.cuda().float()).sum().backward()
      assert abs_err(out, ro.detach()) < BF16_ABS, impl
      for name, got, want in zip('q k v emb bias'.split(), dev, ref):
        assert abs_err(got.grad, want.grad) < BF16_ABS * max(1.0, want.grad.abs().max().item()), (impl, name)


def test_dropout_local_attention_matches_oracle():
  shape = synthetic.GlobalLocalShape(2, 260, 12, 2, 64, 30, 32, 12)
  x = synthetic.make_inputs(shape, seed=13, dtype=torch.bfloat16)
  side = oracle_side(x, shape)
  p, seed = 0.2, 31337
  keep = torch.from_numpy(dropout_ref.keep_mask(seed, p, 2, 2, 260, 260 + 12, rowset=0))
  names = ('long_q', 'long_k', 'long_v', 'global_k', 'global_v', 'long_emb', 'long_bias')
  ref = [x[n].double().requires_grad_() for n in names]
  ro = ao.qkv_relative_local_attention(ref[0], ref[1], ref[2], side['l2l_att_mask'], side['l2l_relative_att_ids'],
                                       ref[5], ref[6], 30, side_k=ref[3], side_v=ref[4],
                                       side_att_mask=side['l2g_att_mask'],
                                       side_relative_att_ids=side['l2g_relative_att_ids'], keep=keep, rate=p)
  (ro * x['d_long_out'].double()).sum().backward()
  for impl in ('tc', 'simt'):
    dev = [x[n].cuda().requires_grad_() for n in names]
    out = ops.local_relative_attention(
        dev[0], dev[1], dev[2], dev[5], dev[6], local_radius=30, side_keys=dev[3], side_values=dev[4],
        compact=ops.LocalCompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(),
                                           x['sentence_ids'].cuda(), 12),
        impl=impl, dropout_p=p, dropout_seed=seed)
    (out.float() * x['d_long_out'].cuda().float()).sum().backward()
    assert abs_err(out, ro.detach()) < BF16_ABS, impl
    for name, got, want in zip(names, dev, ref):
      assert abs_err(got.grad, want.grad) < BF16_ABS * max(1.0, want.grad.abs().max().item()), (impl, name)


# ------------------------------------------------------------------------------------------
# 4. Broadcast (zero-stride) views

def test_broadcast_views_fall_back_to_simt_not_to_garbage():
  """k / v expanded over the batch (stride 0) are not addressable by TMA: 'auto' must route them to the
  SIMT kernels (same result as a materialised copy), 'tc' must refuse; nothing may read ptr + 16 * b."""
  b, s, h, d = 3, 128, 2, 64
  gen = torch.Generator().manual_seed(2)
  q = torch.randn(b, s, h, d, generator=gen).bfloat16().cuda()
  k1 = torch.randn(1, s, h, d, generator=gen).bfloat16().cuda()
  v1 = torch.randn(1, s, h, d, generator=gen).bfloat16().cuda()
  mask = torch.ones(b, s, s, dtype=torch.int32).cuda()
  want = ops.dense_relative_attention(q, k1.expand(b, s, h, d).contiguous(), v1.expand(b, s, h, d).contiguous(),
                                      att_mask=mask, impl='tc')
  got = ops.dense_relative_attention(q, k1.expand(b, s, h, d), v1.expand(b, s, h, d), att_mask=mask, impl='auto')
  assert (got.float() - want.float()).abs().max().item() < BF16_ABS
  with pytest.raises(_lib.MltLibraryError):
    ops.dense_relative_attention(q, k1.expand(b, s, h, d), v1.expand(b, s, h, d), att_mask=mask, impl='tc')
  # a broadcast over an extent of one is fine on the tensor-core path
  one = ops.dense_relative_attention(q[:1], k1.expand(1, s, h, d), v1.expand(1, s, h, d), att_mask=mask[:1], impl='tc')
  assert (one.float() - want[:1].float()).abs().max().item() < BF16_ABS
