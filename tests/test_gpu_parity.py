"""GPU parity: the CUDA path (through the C ABI) against the fp64 oracle.

Tolerances (BASELINE.json north_star): fp32 within 1e-5 relative (measured as
max|x - ref| / max|ref| per tensor), bf16 within 2e-2 max-abs against the fp32/fp64
oracle on O(1)-magnitude outputs; integer constructors bit-exact.
"""
import numpy as np
import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import ops, synthetic
from oracle import attention_oracle as ao
from oracle import feature_oracle as fo

pytestmark = pytest.mark.gpu

FP32_REL = 1e-5
BF16_ABS = 2e-2
NAMES = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb',
         'long_bias', 'global_emb', 'global_bias')


def rel_err(a, ref):
  a = a.detach().double().cpu()
  return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def abs_err(a, ref):
  return (a.detach().double().cpu() - ref).abs().max().item()


def oracle_side(x, shape):
  return {k: torch.tensor(v) for k, v in fo.make_global_local_side_inputs(
      x['long_example_ids'].numpy(), x['global_example_ids'].numpy(),
      x['sentence_ids'].numpy(), shape.local_radius, shape.max_distance).items()}


def run_oracle_gl(x, shape, side):
  ref_in = [x[n].double().requires_grad_() for n in NAMES]
  rl, rg = ao.fused_global_local_attention(*ref_in[:6], side, (ref_in[6], ref_in[7]),
                                           (ref_in[8], ref_in[9]), shape.local_radius)
  ((rl * x['d_long_out'].double()).sum() + (rg * x['d_global_out'].double()).sum()).backward()
  return rl.detach(), rg.detach(), [t.grad for t in ref_in]


def run_cuda_gl(x, shape, side, impl='auto'):
  dev = torch.device('cuda')
  dev_in = [x[n].to(dev).requires_grad_() for n in NAMES]
  lo, go = ops.global_local_attention(*dev_in, local_radius=shape.local_radius, side=side, impl=impl)
  loss = (lo.float() * x['d_long_out'].to(dev).float()).sum() + \
         (go.float() * x['d_global_out'].to(dev).float()).sum()
  loss.backward()
  torch.cuda.synchronize()
  return lo, go, [t.grad for t in dev_in]


def compact_of(x, shape):
  dev = torch.device('cuda')
  return fu.CompactSideInputs(x['long_example_ids'].to(dev), x['global_example_ids'].to(dev),
                              x['sentence_ids'].to(dev), shape.max_distance)


GL_SHAPES = [
    # (B, L, G, H, d, r, R, D)
    (2, 512, 32, 12, 64, 64, 32, 12),   # BASELINE.json configs[0]
    (2, 200, 8, 2, 64, 64, 32, 12),     # L % 64 != 0
    (1, 50, 4, 2, 64, 64, 32, 12),      # L < r
    (2, 130, 5, 3, 32, 7, 20, 3),       # small radius, d = 32
    (1, 96, 6, 1, 128, 16, 32, 12),     # d = 128
]


@pytest.mark.parametrize('dims', GL_SHAPES, ids=lambda d: 'x'.join(map(str, d)))
@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_gl_fp32_matches_oracle(dims, mode):
  b, l, g, h, d, r, rv, dist = dims
  shape = synthetic.GlobalLocalShape(b, l, g, h, d, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=1234 + l)
  # bigger tables than the 0.02 init so that the relative term matters in the comparison
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = x[n] * 10
  side = oracle_side(x, shape)
  rl, rg, rgrads = run_oracle_gl(x, shape, side)
  cuda_side = compact_of(x, shape) if mode == 'compact' else {k: v.cuda() for k, v in side.items()}
  lo, go, grads = run_cuda_gl(x, shape, cuda_side, impl='simt')
  assert rel_err(lo, rl) < FP32_REL
  assert rel_err(go, rg) < FP32_REL
  for name, got, want in zip(NAMES, grads, rgrads):
    assert rel_err(got, want) < FP32_REL, name


@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_gl_bf16_matches_oracle(mode):
  shape = synthetic.GlobalLocalShape(2, 320, 16, 4, 64, 64, 32, 12)
  x = synthetic.make_inputs(shape, seed=77, dtype=torch.bfloat16)
  side = oracle_side(x, shape)
  rl, rg, rgrads = run_oracle_gl(x, shape, side)   # oracle sees the bf16-rounded inputs
  cuda_side = compact_of(x, shape) if mode == 'compact' else {k: v.cuda() for k, v in side.items()}
  lo, go, grads = run_cuda_gl(x, shape, cuda_side)
  assert lo.dtype == torch.bfloat16
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < BF16_ABS * scale, name


def test_gl_partial_side_inputs_and_no_tables():
  # NULL masks = all ones; no tables = no relative term.
  shape = synthetic.GlobalLocalShape(1, 100, 4, 2, 64, 10, 32, 12)
  x = synthetic.make_inputs(shape, seed=3)
  ones = {k: torch.ones_like(v) for k, v in oracle_side(x, shape).items() if 'mask' in k}
  # l2l mask must still exclude out-of-range columns in the oracle's band tensors
  ones['l2l_att_mask'] = torch.tensor(fo.make_local_segmented_att_mask(
      np.ones((1, 100), dtype=np.int32), 10))
  ref_in = [x[n].double() for n in NAMES[:6]]
  rl = ao.qkv_relative_local_attention(ref_in[0], ref_in[1], ref_in[2], ones['l2l_att_mask'], None,
                                       None, None, 10, side_k=ref_in[4], side_v=ref_in[5],
                                       side_att_mask=ones['l2g_att_mask'])
  rg = ao.global_rows_attention(ref_in[3], ref_in[4], ref_in[5], ref_in[1], ref_in[2],
                                ones['g2g_att_mask'], None, ones['g2l_att_mask'], None, None, None)
  dev_in = [x[n].cuda() for n in NAMES[:6]]
  lo, go = ops.global_local_attention(*dev_in, local_radius=10, side=None, impl='simt')
  assert rel_err(lo, rl) < FP32_REL and rel_err(go, rg) < FP32_REL


def test_gl_out_of_vocabulary_ids_and_fully_masked_rows():
  shape = synthetic.GlobalLocalShape(1, 64, 4, 2, 64, 5, 16, 3)
  x = synthetic.make_inputs(shape, seed=5)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = x[n] * 10
  side = oracle_side(x, shape)
  side['l2l_relative_att_ids'][:, :, 0] = 99      # OOV -> contributes 0
  side['g2l_relative_att_ids'][:, 0, :] = -4
  side['l2l_att_mask'][:, 7, :] = 0               # fully masked long row -> uniform
  side['l2g_att_mask'][:, 7, :] = 0
  rl, rg, rgrads = run_oracle_gl(x, shape, side)
  lo, go, grads = run_cuda_gl(x, shape, {k: v.cuda() for k, v in side.items()}, impl='simt')
  assert rel_err(lo, rl) < FP32_REL and rel_err(go, rg) < FP32_REL
  for name, got, want in zip(NAMES, grads, rgrads):
    assert rel_err(got, want) < FP32_REL, name


DENSE_CASES = [
    # (B, S, H, d, R, D, npr, core)
    (2, 96, 3, 64, 32, 12, 0, 0),
    (2, 70, 2, 64, 49, 12, 5, 2),     # 2-D ids: 25 patches + text; ids up to 25+8+25+1 OOV
    (1, 33, 2, 32, 16, 3, 0, 0),
]


@pytest.mark.parametrize('case', DENSE_CASES, ids=lambda c: 'x'.join(map(str, c)))
@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_dense_fp32_matches_oracle(case, mode):
  b, s, h, d, rv, dist, npr, core = case
  gen = torch.Generator().manual_seed(s)
  rn = lambda *sh, std=1.0: torch.randn(*sh, generator=gen) * std
  q, k, v, do = rn(b, s, h, d), rn(b, s, h, d), rn(b, s, h, d), rn(b, s, h, d)
  emb, bias = rn(rv, h, d, std=0.2), rn(rv, h, std=0.2)
  lengths = torch.randint(s // 2, s + 1, (b,), generator=gen)
  e = (torch.arange(s)[None] < lengths[:, None]).int()
  mask = torch.tensor(fo.make_segmented_att_mask(e.numpy()))
  if npr:
    ids = torch.tensor(fo.MmtRelativePositionOracle(npr, core, dist).make_relative_att_ids(s))
  else:
    ids = torch.tensor(fo.make_relative_att_ids_1d(s, dist))
  ids = ids[None].expand(b, s, s).contiguous()
  ref = [t.double().requires_grad_() for t in (q, k, v, emb, bias)]
  ro = ao.qkv_relative_attention(ref[0], ref[1], ref[2], mask, ids, ref[3], ref[4])
  (ro * do.double()).sum().backward()
  dev = [t.cuda().requires_grad_() for t in (q, k, v, emb, bias)]
  if mode == 'compact':
    out = ops.dense_relative_attention(*dev, compact=ops.DenseCompactSideInputs(
        e.cuda(), max_distance=dist, num_patch_per_row=npr, num_core_layers=core), impl='simt')
  else:
    out = ops.dense_relative_attention(*dev, att_mask=mask.cuda(), relative_att_ids=ids.cuda(),
                                       impl='simt')
  (out * do.cuda()).sum().backward()
  torch.cuda.synchronize()
  assert rel_err(out, ro.detach()) < FP32_REL
  for name, got, want in zip('q k v emb bias'.split(), dev, ref):
    assert rel_err(got.grad, want.grad) < FP32_REL, name


def test_dense_bf16_and_strided_views():
  b, s, h, d = 2, 128, 4, 64
  gen = torch.Generator().manual_seed(0)
  qkv = torch.randn(b, s, 3, h, d, generator=gen).bfloat16()   # packed projection output
  q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]           # strided views, d contiguous
  emb = (torch.randn(32, h, d, generator=gen) * 0.1).bfloat16()
  bias = (torch.randn(32, h, generator=gen) * 0.1).bfloat16()
  mask = torch.ones(b, s, s, dtype=torch.int32)
  ids = torch.tensor(fo.make_relative_att_ids_1d(s, 12))[None].expand(b, s, s).contiguous()
  ro = ao.qkv_relative_attention(q.double(), k.double(), v.double(), mask, ids, emb.double(), bias.double())
  out = ops.dense_relative_attention(q.cuda(), k.cuda(), v.cuda(), emb.cuda(), bias.cuda(),
                                     att_mask=mask.cuda(), relative_att_ids=ids.cuda())
  assert abs_err(out, ro) < BF16_ABS


def test_device_side_input_constructors_bit_exact():
  shape = synthetic.GlobalLocalShape(3, 150, 9, 1, 64, 11, 32, 5)
  x = synthetic.make_inputs(shape, seed=1)
  x['sentence_ids'][0, :7] = 8          # irregular sentence assignment
  x['long_example_ids'][1, 40:60] = 2   # packed second example
  got = ops.build_gl_side_inputs(compact_of(x, shape), shape.local_radius)
  want = oracle_side(x, shape)
  assert set(got) == set(want)
  for key in want:
    assert got[key].dtype == torch.int32
    assert torch.equal(got[key].cpu(), want[key]), key
  # dense: mask + 1-D ids, and the reference's golden 2-D matrices
  e = x['long_example_ids'].cuda()
  mask, ids = ops.build_dense_side_inputs(e, max_distance=5)
  assert torch.equal(mask.cpu(), torch.tensor(fo.make_segmented_att_mask(x['long_example_ids'].numpy())))
  assert torch.equal(ids[0].cpu(), torch.tensor(fo.make_relative_att_ids_1d(150, 5)))
  import json, pathlib
  golden = json.loads((pathlib.Path(__file__).parent / 'golden' / 'relative_ids_golden.json').read_text())
  for case in golden['matrices']:
    c = case['ctor']
    s = case['seq_len']
    _, ids2 = ops.build_dense_side_inputs(torch.ones(1, s, dtype=torch.int32).cuda(),
                                          max_distance=c['text_relative_pos_max_distance'],
                                          num_patch_per_row=c['num_patch_per_row'],
                                          num_core_layers=c['num_core_layers'], want_mask=False)
    assert torch.equal(ids2.cpu(), torch.tensor(case['expected'], dtype=torch.int32))
  # larger 2-D case against the oracle
  _, ids3 = ops.build_dense_side_inputs(torch.ones(2, 230, dtype=torch.int32).cuda(), max_distance=12,
                                        num_patch_per_row=14, num_core_layers=2, want_mask=False)
  want3 = torch.tensor(fo.MmtRelativePositionOracle(14, 2, 12).make_relative_att_ids(230))
  assert torch.equal(ids3[1].cpu(), want3)


def test_deterministic_bitwise_repeat():
  shape = synthetic.GlobalLocalShape(2, 256, 16, 2, 64, 64, 32, 12)
  x = synthetic.make_inputs(shape, seed=11)
  side = compact_of(x, shape)
  a = run_cuda_gl(x, shape, side, impl='simt')
  b = run_cuda_gl(x, shape, side, impl='simt')
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
  for ga, gb in zip(a[2], b[2]):
    assert torch.equal(ga, gb)


# ------------------------------------------------------------------------------------------
# tcgen05 path (bf16, d = 64)

TC_SHAPES = [
    # (B, L, G, H, r, R, D)
    (2, 512, 32, 4, 64, 32, 12),
    (1, 200, 8, 2, 64, 32, 12),      # ragged tile / chunk tails
    (1, 50, 4, 2, 64, 32, 12),       # L < r
    (2, 300, 70, 2, 20, 20, 3),      # small radius, R not a multiple of 16, G > 64
    (1, 1024, 64, 2, 100, 64, 30),   # radius > chunk, R = 64: one-warp-set backward configurations
    (1, 1100, 40, 2, 64, 32, 12),    # >= 16 chunks per global tile: the two-warp-set query-centric backward
]
# (Between them the shapes above select every launch configuration of the tcgen05 backward -- slim /
# two CTAs per SM for the long rows and keys, two warp sets for query tiles with many chunks, one warp set
# for relative vocabularies > 32 -- through the library's own shape rules; no debug knobs.)


@pytest.mark.parametrize('dims', TC_SHAPES, ids=lambda d: 'x'.join(map(str, d)))
@pytest.mark.parametrize('mode', ['compact', 'explicit'])
def test_tc_forward_matches_oracle_and_simt(dims, mode):
  b, l, g, h, r, rv, dist = dims
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=l + r, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  side = oracle_side(x, shape)
  rl, rg, rgrads = run_oracle_gl(x, shape, side)
  cuda_side = compact_of(x, shape) if mode == 'compact' else {k: v.cuda() for k, v in side.items()}
  lo, go, grads = run_cuda_gl(x, shape, cuda_side, impl='tc')
  ls, gs, sgrads = run_cuda_gl(x, shape, cuda_side, impl='simt')
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  # the two CUDA paths see identical bf16 inputs: they must agree to bf16 output rounding
  assert abs_err(lo, ls.double().cpu()) < BF16_ABS and abs_err(go, gs.double().cpu()) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < BF16_ABS * scale, name


def test_tc_out_of_vocabulary_ids_and_fully_masked_rows():
  """tcgen05 path: OOV ids contribute 0; a fully-masked row is uniform over its candidates, in the
  forward (lagging-maximum softmax) and in the exponent-form backward (masked elements take 1/l)."""
  shape = synthetic.GlobalLocalShape(1, 200, 8, 2, 64, 20, 16, 3)
  x = synthetic.make_inputs(shape, seed=5, dtype=torch.bfloat16)
  side = oracle_side(x, shape)
  side['l2l_relative_att_ids'][:, :, 0] = 99      # OOV -> contributes 0
  side['g2l_relative_att_ids'][:, 0, :] = -4
  for row in (7, 64, 150):
    side['l2l_att_mask'][:, row, :] = 0           # fully masked long rows -> uniform
    side['l2g_att_mask'][:, row, :] = 0
  side['g2g_att_mask'][:, 3, :] = 0               # and a fully masked global row
  side['g2l_att_mask'][:, 3, :] = 0
  rl, rg, rgrads = run_oracle_gl(x, shape, side)
  lo, go, grads = run_cuda_gl(x, shape, {k: v.cuda() for k, v in side.items()}, impl='tc')
  assert abs_err(lo, rl) < BF16_ABS and abs_err(go, rg) < BF16_ABS
  for name, got, want in zip(NAMES, grads, rgrads):
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < BF16_ABS * scale, name


def _random_tc_case(seed):
  g = torch.Generator().manual_seed(seed)
  ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
  dist = ri(1, 14)
  rv = ri(2 * dist + 4, 40)                      # room for the 1-D ids and the cross ids
  return (ri(1, 2), ri(40, 700), ri(1, 90), ri(1, 3), ri(3, 150), rv, dist)


@pytest.mark.parametrize('seed', range(12))
def test_tc_random_shapes_match_simt(seed):
  """Property test over (B, L, G, H, radius, R, D): the tcgen05 path (planner forms, slim / regular
  launch configurations, lagging-maximum forward, exponent-form backward) against the SIMT path on
  identical bf16 inputs, compact side inputs with ragged lengths."""
  b, l, g, h, r, rv, dist = _random_tc_case(1000 + seed)
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=seed, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  side = compact_of(x, shape)
  lo, go, grads = run_cuda_gl(x, shape, side, impl='tc')
  ls, gs, sgrads = run_cuda_gl(x, shape, side, impl='simt')
  assert abs_err(lo, ls.double().cpu()) < BF16_ABS, (b, l, g, h, r, rv, dist)
  assert abs_err(go, gs.double().cpu()) < BF16_ABS, (b, l, g, h, r, rv, dist)
  for name, got, want in zip(NAMES, grads, sgrads):
    want = want.double().cpu()
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < 2 * BF16_ABS * scale, (name, b, l, g, h, r, rv, dist)


def test_tc_dense_2d_ids():
  b, s, h, d = 2, 230, 2, 64
  gen = torch.Generator().manual_seed(4)
  q, k, v = (torch.randn(b, s, h, d, generator=gen).bfloat16() for _ in range(3))
  emb = (torch.randn(49, h, d, generator=gen) * 0.2).bfloat16()
  bias = (torch.randn(49, h, generator=gen) * 0.2).bfloat16()
  e = (torch.arange(s)[None] < torch.tensor([[230], [180]])).int()
  mask = torch.tensor(fo.make_segmented_att_mask(e.numpy()))
  ids = torch.tensor(fo.MmtRelativePositionOracle(14, 2, 12).make_relative_att_ids(s))[None].expand(b, s, s).contiguous()
  ro = ao.qkv_relative_attention(q.double(), k.double(), v.double(), mask, ids, emb.double(), bias.double())
  for kwargs in (dict(att_mask=mask.cuda(), relative_att_ids=ids.cuda()),
                 dict(compact=ops.DenseCompactSideInputs(e.cuda(), max_distance=12, num_patch_per_row=14,
                                                         num_core_layers=2))):
    out = ops.dense_relative_attention(q.cuda(), k.cuda(), v.cuda(), emb.cuda(), bias.cuda(), impl='tc', **kwargs)
    assert abs_err(out, ro) < BF16_ABS


def test_tc_dense_backward_matches_oracle():
  b, s, h, d = 2, 200, 2, 64
  gen = torch.Generator().manual_seed(8)
  q, k, v, do = (torch.randn(b, s, h, d, generator=gen).bfloat16() for _ in range(4))
  emb = (torch.randn(32, h, d, generator=gen) * 0.2).bfloat16()
  bias = (torch.randn(32, h, generator=gen) * 0.2).bfloat16()
  e = (torch.arange(s)[None] < torch.tensor([[200], [150]])).int()
  mask = torch.tensor(fo.make_segmented_att_mask(e.numpy()))
  ids = torch.tensor(fo.make_relative_att_ids_1d(s, 12))[None].expand(b, s, s).contiguous()
  ref = [t.double().requires_grad_() for t in (q, k, v, emb, bias)]
  ro = ao.qkv_relative_attention(ref[0], ref[1], ref[2], mask, ids, ref[3], ref[4])
  (ro * do.double()).sum().backward()
  for kwargs in (dict(att_mask=mask.cuda(), relative_att_ids=ids.cuda()),
                 dict(compact=ops.DenseCompactSideInputs(e.cuda(), max_distance=12))):
    dev = [t.cuda().requires_grad_() for t in (q, k, v, emb, bias)]
    out = ops.dense_relative_attention(*dev, impl='tc', **kwargs)
    (out.float() * do.cuda().float()).sum().backward()
    assert abs_err(out, ro.detach()) < BF16_ABS
    for name, got, want in zip('q k v emb bias'.split(), dev, ref):
      scale = max(1.0, want.grad.abs().max().item())
      assert abs_err(got.grad, want.grad) < BF16_ABS * scale, name


# ------------------------------------------------------------------------------------------
# Full sweep lengths (BASELINE.json configs[2]): the fp64 oracle is too slow there, so the checks
# are properties that hold at any size, plus the two independent CUDA formulations against each other.

@pytest.mark.parametrize('name', ['c3_2048', 'c3_4096', 'c3_8192'])
def test_full_length_properties(name):
  """At L = 2048 / 4096 / 8192 (G = L / 16, r = 64, 12 heads; batch cut to 1 -- units are independent):
    * rows of the joint softmax sum to one: with V = 1 every output element is 1;
    * gradient mass is conserved: sum_j dV_j (long + global keys) = sum_i dO_i (long + global rows), per
      head and channel, because every softmax row sums to one;
    * the tcgen05 kernels agree with the SIMT kernels (different tiling, different arithmetic order)
      on the outputs and on every gradient."""
  import dataclasses
  seed_off, shape = synthetic.CONFIGS[name]
  shape = dataclasses.replace(shape, batch=1)
  x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16)
  side = compact_of(x, shape)

  lo, go, grads = run_cuda_gl(x, shape, side, impl='tc')
  ls, gs, sgrads = run_cuda_gl(x, shape, side, impl='simt')
  assert abs_err(lo, ls.double().cpu()) < BF16_ABS and abs_err(go, gs.double().cpu()) < BF16_ABS
  for n, got, want in zip(NAMES, grads, sgrads):
    want = want.double().cpu()
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < 2 * BF16_ABS * scale, n

  # gradient mass: sum over keys of dV == sum over rows of dO  (fp32 sums of bf16 tensors)
  d_v = grads[2].float().sum(1) + grads[5].float().sum(1)                       # [B, H, d]
  d_o = x['d_long_out'].float().sum(1) + x['d_global_out'].float().sum(1)
  tol = 2e-2 * d_o.abs().max().item() + 0.5     # bf16 rounding of ~L summands per element
  assert (d_v.cpu() - d_o).abs().max().item() < tol

  # V = 1  ->  out = 1
  ones = dict(x)
  ones['long_v'] = torch.ones_like(x['long_v'])
  ones['global_v'] = torch.ones_like(x['global_v'])
  lo1, go1, _ = run_cuda_gl(ones, shape, side, impl='tc')
  assert (lo1.float() - 1).abs().max().item() < 1e-2 and (go1.float() - 1).abs().max().item() < 1e-2


def test_tc_partial_explicit_side_inputs():
  """tcgen05 EXPL form with some of the eight explicit tensors missing: a missing mask is all ones, missing
  ids contribute no relative term (per block, as in the reference's optional arguments); fwd + bwd vs the
  SIMT kernels on identical bf16 inputs, for an aligned shape (128-bit loads) and a ragged one."""
  for dims in ((2, 256, 32, 2, 64, 32, 12), (1, 210, 9, 2, 64, 32, 12)):
    b, l, g, h, r, rv, dist = dims
    shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
    x = synthetic.make_inputs(shape, seed=l, dtype=torch.bfloat16)
    for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
      x[n] = (x[n].float() * 10).bfloat16()
    full = {k: v.cuda() for k, v in oracle_side(x, shape).items()}
    for drop in (('l2g_att_mask', 'g2g_relative_att_ids'), ('l2l_relative_att_ids', 'g2l_att_mask'),
                 ('l2l_att_mask', 'l2g_relative_att_ids', 'g2g_att_mask', 'g2l_relative_att_ids')):
      side = {k: v for k, v in full.items() if k not in drop}
      lo, go, grads = run_cuda_gl(x, shape, side, impl='tc')
      ls, gs, sgrads = run_cuda_gl(x, shape, side, impl='simt')
      assert abs_err(lo, ls.double().cpu()) < BF16_ABS and abs_err(go, gs.double().cpu()) < BF16_ABS, drop
      for name, got, want in zip(NAMES, grads, sgrads):
        want = want.double().cpu()
        scale = max(1.0, want.abs().max().item())
        assert abs_err(got, want) < 2 * BF16_ABS * scale, (name, drop)


def test_tc_dense_2d_compact_backward_matches_simt_and_explicit():
  """Compact 2-D descriptors on the tcgen05 path (id plane materialised in the workspace, read through the
  EXPL form with the example-id mask) against the SIMT kernels (closed 2-D rule) and against the same ids
  fed explicitly; fwd + bwd, ragged example lengths so that the mask changes inside 32-column groups."""
  b, s, h, d, rv = 2, 300, 2, 64, 49
  gen = torch.Generator().manual_seed(11)
  q, k, v, do = (torch.randn(b, s, h, d, generator=gen).bfloat16() for _ in range(4))
  emb = (torch.randn(rv, h, d, generator=gen) * 0.2).bfloat16()
  bias = (torch.randn(rv, h, generator=gen) * 0.2).bfloat16()
  e = (torch.arange(s)[None] < torch.tensor([[300], [217]])).int().cuda()
  compact = ops.DenseCompactSideInputs(e, max_distance=12, num_patch_per_row=14, num_core_layers=2)
  mask, ids = ops.build_dense_side_inputs(e, 12, num_patch_per_row=14, num_core_layers=2)
  results = []
  for impl, kwargs in (('tc', dict(compact=compact)), ('simt', dict(compact=compact)),
                       ('tc', dict(att_mask=mask, relative_att_ids=ids))):
    dev = [t.cuda().requires_grad_() for t in (q, k, v, emb, bias)]
    out = ops.dense_relative_attention(*dev, impl=impl, **kwargs)
    (out.float() * do.cuda().float()).sum().backward()
    results.append([out.detach()] + [t.grad for t in dev])
  for other in results[1:]:
    for name, got, want in zip('out q k v emb bias'.split(), results[0], other):
      want = want.double().cpu()
      scale = max(1.0, want.abs().max().item())
      assert abs_err(got, want) < 2 * BF16_ABS * scale, name


def test_tc_dense_2d_cross_modality_ids_inside_the_vocabulary():
  """A 2-D layout small enough (6 x 6 patches, distance 3) for image_part_id = 51 / text_part_id = 52 to lie
  inside a relative vocabulary of 64: the cross-modality groups then carry a real per-row constant and a real
  table gradient (at 14 x 14 patches those ids fall outside the vocabulary and contribute nothing).
  tcgen05 path (planner forms: 1-D for text x text, one id per query kind for text x image, the id plane for
  image x image) against the SIMT kernels (closed rule) and the fp64 oracle, fwd + bwd."""
  b, s, h, d, rv, npr, core, dist = 2, 300, 2, 64, 64, 6, 1, 3
  gen = torch.Generator().manual_seed(23)
  q, k, v, do = (torch.randn(b, s, h, d, generator=gen).bfloat16() for _ in range(4))
  emb = (torch.randn(rv, h, d, generator=gen) * 0.2).bfloat16()
  bias = (torch.randn(rv, h, generator=gen) * 0.5).bfloat16()
  e = (torch.arange(s)[None] < torch.tensor([[300], [217]])).int()
  compact = ops.DenseCompactSideInputs(e.cuda(), max_distance=dist, num_patch_per_row=npr, num_core_layers=core)
  ids_host = torch.tensor(fo.MmtRelativePositionOracle(npr, core, dist).make_relative_att_ids(s))
  assert int(ids_host.max()) == npr * npr + 8 + 2 * dist + 2 < rv        # text_part_id is a real table row
  results = []
  for impl in ('tc', 'simt'):
    dev = [t.cuda().requires_grad_() for t in (q, k, v, emb, bias)]
    out = ops.dense_relative_attention(*dev, impl=impl, compact=compact)
    (out.float() * do.cuda().float()).sum().backward()
    results.append([out.detach()] + [t.grad for t in dev])
  ref = [t.double().requires_grad_() for t in (q, k, v, emb, bias)]
  mask = torch.tensor(fo.make_segmented_att_mask(e.numpy()))
  ro = ao.qkv_relative_attention(ref[0], ref[1], ref[2], mask, ids_host[None].expand(b, s, s), ref[3], ref[4])
  (ro * do.double()).sum().backward()
  want_all = [ro.detach()] + [t.grad for t in ref]
  for name, got, simt, want in zip('out q k v emb bias'.split(), results[0], results[1], want_all):
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < 2 * BF16_ABS * scale, name
    assert abs_err(got, simt.double().cpu()) < 2 * BF16_ABS * scale, name


@pytest.mark.parametrize('seed', range(6))
def test_tc_random_shapes_explicit_match_simt(seed):
  """Property test of the EXPL form over (B, L, G, H, radius, R, D): explicit int32 side inputs built by the
  device-side constructor, with random out-of-vocabulary ids and random extra masking sprinkled in, tcgen05
  against SIMT on identical bf16 inputs (fwd + bwd)."""
  b, l, g, h, r, rv, dist = _random_tc_case(2000 + seed)
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=50 + seed, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  side = ops.build_gl_side_inputs(compact_of(x, shape), shape.local_radius)
  gen = torch.Generator(device='cuda').manual_seed(seed)
  for k, t in side.items():
    noise = torch.rand(t.shape, generator=gen, device='cuda')
    if k.endswith('relative_att_ids'):
      t[noise < 0.02] = rv + 3            # out-of-vocabulary: contributes 0
      t[(noise >= 0.02) & (noise < 0.03)] = -2
    else:
      t[noise < 0.05] = 0                 # extra masked pairs
  lo, go, grads = run_cuda_gl(x, shape, side, impl='tc')
  ls, gs, sgrads = run_cuda_gl(x, shape, side, impl='simt')
  assert abs_err(lo, ls.double().cpu()) < BF16_ABS, (b, l, g, h, r, rv, dist)
  assert abs_err(go, gs.double().cpu()) < BF16_ABS, (b, l, g, h, r, rv, dist)
  for name, got, want in zip(NAMES, grads, sgrads):
    want = want.double().cpu()
    scale = max(1.0, want.abs().max().item())
    assert abs_err(got, want) < 2 * BF16_ABS * scale, (name, b, l, g, h, r, rv, dist)
