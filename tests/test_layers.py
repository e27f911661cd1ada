"""Layer mirrors: constructor / call validation on the CPU, numerics against the oracle on the GPU."""
import numpy as np
import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import layers, mmt_encoder, ops
from oracle import attention_oracle as ao
from oracle import feature_oracle as fo


def test_constructor_validation_matches_reference_messages():
  # reference mmt_encoder.py:69-80
  with pytest.raises(ValueError, match='too small'):
    mmt_encoder.MmtEncoder(vocab_size=100, relative_vocab_size=27, relative_pos_max_distance=12)
  with pytest.raises(ValueError, match='must be 0'):
    mmt_encoder.MmtEncoder(vocab_size=100, relative_vocab_size=None, relative_pos_max_distance=3)
  mmt_encoder.MmtEncoder(vocab_size=100, relative_vocab_size=28, relative_pos_max_distance=12,
                         num_hidden_layers=1)
  with pytest.raises(ValueError):
    layers.QkvRelativeLocalAttention(2, 8, local_radius=0)
  with pytest.raises(ValueError):
    layers.RelativeAttention(hidden_size=10, num_heads=3)
  with pytest.raises(ValueError):
    layers.FusedGlobalLocalAttention(16, 32, 2, 4, share_qkv_projections=True)


def test_call_validation_without_gpu():
  att = layers.QkvRelativeAttention(2, 64)   # no relative vocab
  q = torch.zeros(1, 4, 2, 64)
  with pytest.raises(ValueError, match='relative_vocab_size'):
    att(q, q, q, relative_att_ids=torch.zeros(1, 4, 4, dtype=torch.int32))
  loc = layers.QkvRelativeLocalAttention(2, 64, local_radius=2, relative_vocab_size=8)
  with pytest.raises(ValueError, match='side_keys'):
    loc(q, q, q, side_keys=q)
  with pytest.raises(ValueError, match='att_implementation'):
    loc(q, q, q, att_implementation='dense')
  fused = layers.FusedGlobalLocalAttention(128, 128, 2, 4, relative_vocab_size=None)
  with pytest.raises(ValueError, match='relative_vocab_size'):
    fused(torch.zeros(1, 8, 128), torch.zeros(1, 2, 128),
          l2l_relative_att_ids=torch.zeros(1, 8, 9, dtype=torch.int32))


def test_parameter_names_and_sharing():
  fused = layers.FusedGlobalLocalAttention(128, 128, 2, 4, relative_vocab_size=16,
                                           share_qkv_projections=True, share_att_output_projection=True)
  assert fused.global_query_projection is fused.long_query_projection
  assert fused.global_output_projection is fused.long_output_projection
  assert tuple(fused.long_tables.relative_emb_table.shape) == (16, 2, 64)
  assert tuple(fused.global_tables.relative_bias_table.shape) == (16, 2)
  assert float(fused.long_tables.relative_bias_table.abs().sum()) == 0.0   # zeros init
  stack = layers.RelativeTransformerLayers(128, 2, 2, relative_vocab_size=16, use_pre_activation_order=True)
  assert stack.output_layer_norm is not None and len(stack.attention_layers) == 2


@pytest.mark.gpu
def test_local_attention_layer_matches_oracle_fwd_bwd():
  torch.manual_seed(0)
  b, l, g, h, d, r, rv, dist = 2, 150, 6, 2, 64, 9, 16, 3
  layer = layers.QkvRelativeLocalAttention(h, d, local_radius=r, relative_vocab_size=rv).cuda()
  with torch.no_grad():
    layer.relative_emb_table.mul_(10)
    layer.relative_bias_table.normal_(std=0.3)
  q, k, v = (torch.randn(b, l, h, d) for _ in range(3))
  sk, sv = torch.randn(b, g, h, d), torch.randn(b, g, h, d)
  le = (torch.arange(l)[None] < torch.tensor([[150], [97]])).int()
  ge = torch.ones(b, g, dtype=torch.int32)
  sid = (torch.arange(l) * g // l)[None].expand(b, l).int()
  side = {kk: torch.tensor(vv) for kk, vv in fo.make_global_local_side_inputs(
      le.numpy(), ge.numpy(), sid.numpy(), r, dist).items()}
  ref = [t.double().requires_grad_() for t in (q, k, v, sk, sv)]
  emb, bias = layer.relative_emb_table.detach().double().cpu(), layer.relative_bias_table.detach().double().cpu()
  ro = ao.qkv_relative_local_attention(ref[0], ref[1], ref[2], side['l2l_att_mask'],
                                       side['l2l_relative_att_ids'], emb, bias, r, side_k=ref[3],
                                       side_v=ref[4], side_att_mask=side['l2g_att_mask'],
                                       side_relative_att_ids=side['l2g_relative_att_ids'])
  ro.sum().backward()
  for mode in ('explicit', 'compact'):
    dev = [t.cuda().requires_grad_() for t in (q, k, v, sk, sv)]
    if mode == 'explicit':
      out = layer(dev[0], dev[1], dev[2], att_mask=side['l2l_att_mask'].cuda(),
                  relative_att_ids=side['l2l_relative_att_ids'].cuda(), side_keys=dev[3],
                  side_values=dev[4], side_att_mask=side['l2g_att_mask'].cuda(),
                  side_relative_att_ids=side['l2g_relative_att_ids'].cuda())
    else:
      out = layer(dev[0], dev[1], dev[2], side_keys=dev[3], side_values=dev[4],
                  compact=ops.LocalCompactSideInputs(le.cuda(), ge.cuda(), sid.cuda(), dist))
    out.sum().backward()
    rel = lambda a, bb: ((a.detach().double().cpu() - bb).abs().max() / bb.abs().max()).item()
    assert rel(out, ro.detach()) < 1e-5
    for got, want in zip(dev, ref):
      assert rel(got.grad, want.grad) < 1e-5


@pytest.mark.gpu
def test_fused_layer_and_stacks_run_and_agree_between_side_input_modes():
  torch.manual_seed(1)
  b, l, g, hid, h, r = 2, 256, 16, 128, 2, 64
  x, xg = torch.randn(b, l, hid).cuda(), torch.randn(b, g, hid).cuda()
  le = (torch.arange(l)[None] < torch.tensor([[256], [200]])).int().cuda()
  ge = torch.ones(b, g, dtype=torch.int32).cuda()
  sid = (torch.arange(l) * g // l)[None].expand(b, l).int().cuda()
  compact = fu.CompactSideInputs(le, ge, sid, 12)
  explicit = ops.build_gl_side_inputs(compact, r)
  stack = layers.GlobalLocalTransformerLayers(hid, hid, 2, h, r, relative_vocab_size=32,
                                              hidden_dropout_prob=0.0, use_pre_activation_order=True).cuda()
  lo1, go1 = stack(x, xg, compact_side_inputs=compact, training=False)
  lo2, go2 = stack(x, xg, training=False, **explicit)
  assert torch.allclose(lo1, lo2, atol=1e-5) and torch.allclose(go1, go2, atol=1e-5)
  (lo1.sum() + go1.sum()).backward()
  assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in stack.parameters())


@pytest.mark.gpu
def test_encoder_dense_matches_reference_call_contract():
  # the call the reference makes: word_ids, segment_ids, att_mask, relative_att_ids, patch_embeddings
  torch.manual_seed(2)
  b, s, npatch = 2, 64, 16
  enc = mmt_encoder.MmtEncoder(vocab_size=100, hidden_size=128, num_hidden_layers=2,
                               num_attention_heads=2, intermediate_size=256, relative_vocab_size=32,
                               relative_pos_max_distance=12, hidden_dropout_prob=0.0,
                               use_pre_activation_order=True, patch_embedding_size=48).cuda()
  word_ids = torch.randint(0, 100, (b, s)).cuda()
  lengths = torch.tensor([64, 40])
  bp = torch.tensor(fo.breakpoints_from_lengths(lengths.numpy(), s))
  side = fu.make_relative_transformer_side_inputs(bp, fu.RelativePositionGenerator(12), 12)
  patches = torch.randn(b, npatch, 48).cuda()
  out = enc(word_ids, att_mask=side.att_mask.cuda(), relative_att_ids=side.relative_att_ids.cuda(),
            patch_embeddings=patches, training=False)
  assert set(out) == {'sequence_output'} and out['sequence_output'].shape == (b, s, 128)
  e = fu.example_ids_from_breakpoints(bp).cuda()
  out2 = enc(word_ids, patch_embeddings=patches, training=False,
             compact=ops.DenseCompactSideInputs(e, max_distance=12))
  assert torch.allclose(out['sequence_output'], out2['sequence_output'], atol=1e-5)
