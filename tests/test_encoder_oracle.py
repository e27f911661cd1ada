"""Encoder forward (SURVEY next-1, reference src/modeling/models/mmt_encoder.py:166-237) against the fp64
oracle built on attention_oracle: GPU parity for the dense stack fed exactly like the reference ([B,S,S] int32
mask + 2-D relative ids; S = 512 = 2 + 196 patches + text) and for the long-input (global-local) stack.
The oracle consumes the state dict, not the mirror's forward code."""
import numpy as np
import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import mmt_encoder, ops
from oracle import encoder_oracle as eo
from oracle import feature_oracle as fo


def _sd64(model):
  sd = {}
  for k, v in model.state_dict().items():
    sd[k] = v.detach().double().cpu() if v.is_floating_point() else v.detach().cpu()
  return sd


def _dense_case(b, s, npr, hidden, heads, layers_n, pre_ln, seed):
  torch.manual_seed(seed)
  enc = mmt_encoder.MmtEncoder(vocab_size=512, hidden_size=hidden, num_hidden_layers=layers_n,
                               num_attention_heads=heads, intermediate_size=2 * hidden, relative_vocab_size=49,
                               relative_pos_max_distance=12, use_pre_activation_order=pre_ln,
                               patch_embedding_size=96, max_absolute_position_embeddings=s)
  # non-trivial tables / norms so that every term matters
  with torch.no_grad():
    for n, p in enc.named_parameters():
      if 'relative_bias_table' in n or n.endswith('norm.bias') or 'norms' in n and n.endswith('bias'):
        p.normal_(0, 0.1)
      if 'relative_emb_table' in n:
        p.mul_(5)
  gen = torch.Generator().manual_seed(seed + 1)
  word_ids = torch.randint(0, 512, (b, s), generator=gen)
  segment_ids = torch.randint(0, 3, (b, s), generator=gen)
  patches = torch.randn(b, npr * npr, 96, generator=gen)
  lengths = torch.randint(s // 2, s + 1, (b,), generator=gen)
  eid = (torch.arange(s)[None] < lengths[:, None]).int()
  mask = torch.tensor(fo.make_segmented_att_mask(eid.numpy()))
  ids = torch.tensor(fo.MmtRelativePositionOracle(npr, 2, 12).make_relative_att_ids(s))[None].expand(b, s, s).contiguous()
  return enc, word_ids, segment_ids, patches, eid, mask, ids


def test_oracle_embed_matches_reference_layout_on_cpu():
  """CPU: the patch projection lands on positions 2 .. 2 + P (reference :203-217) and the word embedding is
  normalised BEFORE the segment / position / patch terms are added (:191-201)."""
  enc, word_ids, segment_ids, patches, *_ = _dense_case(1, 40, 3, 32, 2, 1, False, 0)
  sd = _sd64(enc)
  with_p = eo.embed(sd, word_ids, segment_ids, patches.double())
  without = eo.embed(sd, word_ids, segment_ids, None)
  diff = (with_p - without).abs().sum(-1)[0]
  assert bool((diff[2:11] > 0).all()) and float(diff[:2].sum()) == 0 and float(diff[11:].sum()) == 0
  got = enc.eval().embed(word_ids, segment_ids, patches)
  assert torch.allclose(got.double(), with_p, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('pre_ln', [False, True])
def test_dense_encoder_forward_matches_fp64_oracle(pre_ln):
  """2 layers, S = 512 = 2 + 196 patches + text, the reference's explicit [B,S,S] side inputs (2-D ids):
  fp32 model on the SIMT kernels against fp64 (1e-4 relative over two layers of LayerNorm / GELU), then the
  bf16 model on the tcgen05 kernels against the same oracle evaluated on the bf16-rounded weights."""
  enc, word_ids, segment_ids, patches, eid, mask, ids = _dense_case(2, 512, 14, 128, 2, 2, pre_ln, 3)
  want = eo.dense_encoder_forward(_sd64(enc), word_ids, mask, ids, 2, 2, pre_ln, segment_ids, patches.double())
  dev = torch.device('cuda')
  enc = enc.to(dev).eval()
  got = enc(word_ids.to(dev), segment_ids.to(dev), att_mask=mask.to(dev), relative_att_ids=ids.to(dev),
            patch_embeddings=patches.to(dev), training=False)['sequence_output']
  err = (got.double().cpu() - want).abs().max().item() / want.abs().max().item()
  assert err < 1e-4, err
  # compact 2-D descriptors give the same result as the explicit tensors
  got_c = enc(word_ids.to(dev), segment_ids.to(dev), patch_embeddings=patches.to(dev), training=False,
              compact=ops.DenseCompactSideInputs(eid.to(dev), max_distance=12, num_patch_per_row=14,
                                                 num_core_layers=2))['sequence_output']
  assert (got_c - got).abs().max().item() < 1e-4
  # bf16 / tcgen05
  enc16 = enc.to(torch.bfloat16)
  want16 = eo.dense_encoder_forward(_sd64(enc16), word_ids, mask, ids, 2, 2, pre_ln, segment_ids,
                                    patches.bfloat16().double())
  got16 = enc16(word_ids.to(dev), segment_ids.to(dev), att_mask=mask.to(dev), relative_att_ids=ids.to(dev),
                patch_embeddings=patches.to(dev).bfloat16(), training=False)['sequence_output']
  err16 = (got16.double().cpu() - want16).abs().max().item()
  assert err16 < 0.1 * max(1.0, want16.abs().max().item()), err16      # two layers of bf16 GEMMs + LayerNorm


@pytest.mark.gpu
def test_global_local_encoder_forward_matches_fp64_oracle():
  """Long-input stack (2 layers, L = 320, 20 global tokens, radius 64), compact descriptors on the device
  against the oracle fed the explicit tensors built by the loop oracle."""
  torch.manual_seed(5)
  l, g, b = 320, 20, 2
  enc = mmt_encoder.MmtEncoder(vocab_size=256, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                               intermediate_size=256, relative_vocab_size=32, relative_pos_max_distance=12,
                               use_pre_activation_order=True, patch_embedding_size=48, local_radius=64,
                               num_global_tokens=g)
  with torch.no_grad():
    for n, p in enc.named_parameters():
      if 'relative_bias_table' in n:
        p.normal_(0, 0.1)
      if 'relative_emb_table' in n:
        p.mul_(5)
  gen = torch.Generator().manual_seed(6)
  word_ids = torch.randint(0, 256, (b, l), generator=gen)
  patches = torch.randn(b, 36, 48, generator=gen)
  lengths = torch.randint(l // 2, l + 1, (b,), generator=gen)
  le = (torch.arange(l)[None] < lengths[:, None]).int()
  ge = torch.ones(b, g, dtype=torch.int32)
  sent = ((torch.arange(l) * g) // l)[None].expand(b, l).int().contiguous()
  side = {k: torch.tensor(v) for k, v in fo.make_global_local_side_inputs(le.numpy(), ge.numpy(), sent.numpy(), 64, 12).items()}
  wl, wg = eo.global_local_encoder_forward(_sd64(enc), word_ids, side, 2, 2, 64, g, True, None, patches.double())
  dev = torch.device('cuda')
  enc = enc.to(dev).eval()
  out = enc(word_ids.to(dev), patch_embeddings=patches.to(dev), training=False,
            compact_side_inputs=fu.CompactSideInputs(le.to(dev), ge.to(dev), sent.to(dev), 12))
  for got, want in ((out['sequence_output'], wl), (out['global_output'], wg)):
    err = (got.double().cpu() - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-4, err
