"""fp64 restatement of the encoder forward (SURVEY next-1), built on ``attention_oracle``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows the reference call
``src/modeling/models/mmt_encoder.py:166-237`` step by step:

  :191-197  word embedding -> LayerNorm(eps 1e-12) -> dropout (inference: identity) + segment embedding
  :199-201  + absolute position embedding (optional)
  :203-217  + Dense(patch_embeddings) padded by 2 leading slots ([CLS], [PATCH]) and a suffix
  :220-224  RelativeTransformerLayers(inputs, att_mask, relative_att_ids, training)

and, inside the stack [UPSTREAM-RECALLED, etcmodel/layers/transformer.py], per layer

  post-LN:  x = LN(x + Att(x));      x = LN(x + FFN(x))
  pre-LN :  x = x + Att(LN(x));      x = x + FFN(LN(x));      final LN
  Att = output_projection(QkvRelativeAttention(q_proj(x), k_proj(x), v_proj(x)));  FFN = Dense-gelu(tanh)-Dense.

The functions take a *state dict* (plain tensors) plus the hyper-parameters, not the mirror modules, so that
the product's forward code is not what is being compared with itself.  PARITY UNPINNED in the sense of
``attention_oracle`` (etcmodel is absent); what this pins is the product stack against an independent fp64
evaluation of the same published algorithm.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import attention_oracle as ao


def _ln(x, sd, prefix, eps=1e-12):
  return F.layer_norm(x, (x.shape[-1],), sd[prefix + '.weight'], sd[prefix + '.bias'], eps)


def _dense(x, sd, prefix):
  b = sd.get(prefix + '.bias')
  return F.linear(x, sd[prefix + '.weight'], b)


def _gelu_tanh(x):
  return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _heads(x, h):
  return x.reshape(x.shape[0], x.shape[1], h, x.shape[2] // h)


def embed(sd, word_ids, segment_ids=None, patch_embeddings=None):
  """reference mmt_encoder.py:186-217 (inference)."""
  if segment_ids is None:
    segment_ids = torch.ones_like(word_ids)

  def lookup(prefix, ids):
    x = sd[prefix + '.table.weight'][ids]
    if prefix + '.projection.weight' in sd:
      x = F.linear(x, sd[prefix + '.projection.weight'])
    return x

  word = _ln(lookup('word_embeddings', word_ids), sd, 'embedding_norm')
  emb = word + lookup('segment_embeddings', segment_ids)
  if 'position_embeddings' in sd:
    emb = emb + sd['position_embeddings'][:emb.shape[1]][None]
  if patch_embeddings is not None:
    pe = _dense(patch_embeddings, sd, 'patch_embedding_projection')
    emb = emb + F.pad(pe, (0, 0, 2, emb.shape[1] - 2 - pe.shape[1]))
  return emb


def dense_encoder_forward(sd, word_ids, att_mask, relative_att_ids, num_layers, num_heads, pre_ln,
                          segment_ids=None, patch_embeddings=None):
  """``MmtEncoder.call`` with the dense stack; ``sd`` = fp64 state dict of the encoder."""
  x = embed(sd, word_ids, segment_ids, patch_embeddings)
  t = 'transformer_layers.'
  for n in range(num_layers):
    a = f'{t}attention_layers.{n}.'

    def att(y):
      q = _heads(_dense(y, sd, a + 'query_projection.linear'), num_heads)
      k = _heads(_dense(y, sd, a + 'key_projection.linear'), num_heads)
      v = _heads(_dense(y, sd, a + 'value_projection.linear'), num_heads)
      emb = sd.get(a + 'qkv_relative_attention.relative_emb_table')
      bias = sd.get(a + 'qkv_relative_attention.relative_bias_table')
      ids = relative_att_ids if emb is not None else None
      o = ao.qkv_relative_attention(q, k, v, att_mask, ids, emb, bias)
      return _dense(o.reshape(o.shape[0], o.shape[1], -1), sd, a + 'output_projection')

    def ffn(y):
      f = f'{t}feed_forward_layers.{n}.'
      return _dense(_gelu_tanh(_dense(y, sd, f + 'inner')), sd, f + 'outer')

    n1, n2 = f'{t}attention_norms.{n}', f'{t}feed_forward_norms.{n}'
    if pre_ln:
      x = x + att(_ln(x, sd, n1))
      x = x + ffn(_ln(x, sd, n2))
    else:
      x = _ln(x + att(x), sd, n1)
      x = _ln(x + ffn(x), sd, n2)
  if pre_ln:
    x = _ln(x, sd, t + 'output_layer_norm')
  return x


def global_local_encoder_forward(sd, word_ids, side, num_layers, num_heads, local_radius, num_global_tokens,
                                 pre_ln, segment_ids=None, patch_embeddings=None):
  """Long-input variant: GlobalLocalTransformerLayers over (embeddings, broadcast global embedding);
  ``side`` = dict of the eight explicit l2l / l2g / g2g / g2l int32 tensors.  Returns (long, global)."""
  xl = embed(sd, word_ids, segment_ids, patch_embeddings)
  xg = sd['global_embeddings'].expand(xl.shape[0], num_global_tokens, -1)
  t = 'transformer_layers.'
  for n in range(num_layers):
    a = f'{t}fused_att_layers.{n}.'

    def att(yl, yg):
      proj = lambda y, name: _heads(_dense(y, sd, a + name + '.linear'), num_heads)
      lq, lk, lv = (proj(yl, f'long_{s}_projection') for s in ('query', 'key', 'value'))
      gq, gk, gv = (proj(yg, f'global_{s}_projection') for s in ('query', 'key', 'value'))
      lt = (sd[a + 'long_tables.relative_emb_table'], sd[a + 'long_tables.relative_bias_table'])
      gt = (sd[a + 'global_tables.relative_emb_table'], sd[a + 'global_tables.relative_bias_table'])
      lo, go = ao.fused_global_local_attention(lq, lk, lv, gq, gk, gv, side, lt, gt, local_radius)
      flat = lambda o: o.reshape(o.shape[0], o.shape[1], -1)
      return (_dense(flat(lo), sd, a + 'long_output_projection'),
              _dense(flat(go), sd, a + 'global_output_projection'))

    def ffn(y, which):
      f = f'{t}{which}_ffn.{n}.'
      return _dense(_gelu_tanh(_dense(y, sd, f + 'inner')), sd, f + 'outer')

    nla, nga = f'{t}long_att_norms.{n}', f'{t}global_att_norms.{n}'
    nlf, ngf = f'{t}long_ffn_norms.{n}', f'{t}global_ffn_norms.{n}'
    if pre_ln:
      al, ag = att(_ln(xl, sd, nla), _ln(xg, sd, nga))
      xl, xg = xl + al, xg + ag
      xl, xg = xl + ffn(_ln(xl, sd, nlf), 'long'), xg + ffn(_ln(xg, sd, ngf), 'global')
    else:
      al, ag = att(xl, xg)
      xl, xg = _ln(xl + al, sd, nla), _ln(xg + ag, sd, nga)
      xl, xg = _ln(xl + ffn(xl, 'long'), sd, nlf), _ln(xg + ffn(xg, 'global'), sd, ngf)
  if pre_ln:
    xl, xg = _ln(xl, sd, t + 'long_output_norm'), _ln(xg, sd, t + 'global_output_norm')
  return xl, xg
