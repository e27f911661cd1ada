"""Host-side mirror of the ETC attention / transformer layers the reference builds on.

The reference reaches its attention through
``etc_layers.RelativeTransformerLayers(...)(inputs, att_mask, relative_att_ids, training)``
(``src/modeling/models/mmt_encoder.py:124-135`` ctor kwargs, ``:220-224`` call).  The classes
here keep the same names, constructor arguments, call arguments and error behaviour
[UPSTREAM-RECALLED for everything inside ``etcmodel``], with PyTorch modules standing in for
Keras layers (TensorFlow is not installable in this image).  Only the attention *core* runs in
the CUDA library (``ops``); projections, layer norm and feed-forward are plain host-framework
ops, unchanged -- exactly the split the north star prescribes.

  ProjectAttentionHeads        Dense -> reshape [B, L, H, d]
  QkvRelativeAttention         contract (A) core, owns relative_emb_table / relative_bias_table
  QkvRelativeLocalAttention    long rows of contract (B)
  RelativeAttention            q/k/v projections + QkvRelativeAttention + output projection
  FusedGlobalLocalAttention    contract (B): call(long_input, global_input, l2l_att_mask, ...)
  RelativeTransformerLayers    the stack the reference instantiates (dense attention)
  GlobalLocalTransformerLayers the long-input stack (global-local attention)

Attention-probability dropout (``att_dropout_prob`` / ``attention_probs_dropout_prob``, reference
default 0.1, ``src/configs/encoders.py:87-88``) runs inside the fused kernels (``ops`` arguments
``dropout_p`` / ``dropout_seed``).

``training`` follows Keras: an explicit ``training=True/False`` applies to THIS call and to every
nested layer (attention-probability dropout, hidden dropout, feed-forward dropout) without changing
the module's own mode; ``training=None`` falls back to the module mode (``.train()`` / ``.eval()``).
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn
import torch.nn.functional as F

from . import ops
from .feature_utils import CompactSideInputs

def resolve_training(module: nn.Module, training) -> bool:
  """Keras ``training`` argument: explicit value wins, ``None`` -> the module's mode."""
  return module.training if training is None else bool(training)


def _trunc_normal_(t: torch.Tensor, std: float):
  return nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std)


def gelu_approximate(x):
  """``tf.keras.activations.gelu(x, approximate=True)`` (reference mmt_encoder.py:53-54)."""
  return F.gelu(x, approximate='tanh')


class ProjectAttentionHeads(nn.Module):
  """Dense projection to ``[B, L, num_heads, size_per_head]`` (heads NOT transposed)."""

  def __init__(self, input_size: int, num_heads: int, size_per_head: int,
               use_bias: bool = True, initializer_range: float = 0.02):
    super().__init__()
    if num_heads < 1:
      raise ValueError('`num_heads` must be positive.')
    if size_per_head < 1:
      raise ValueError('`size_per_head` must be positive.')
    self.num_heads = num_heads
    self.size_per_head = size_per_head
    self.linear = nn.Linear(input_size, num_heads * size_per_head, bias=use_bias)
    _trunc_normal_(self.linear.weight, initializer_range)
    if use_bias:
      nn.init.zeros_(self.linear.bias)

  def forward(self, inputs: torch.Tensor) -> torch.Tensor:
    x = self.linear(inputs)
    return x.view(*x.shape[:-1], self.num_heads, self.size_per_head)


class _RelativeTables(nn.Module):
  """``relative_emb_table [R,H,d]`` (TruncNormal) and ``relative_bias_table [R,H]`` (zeros)."""

  def __init__(self, relative_vocab_size: Optional[int], num_heads: int, size_per_head: int,
               initializer_range: float):
    super().__init__()
    self.relative_vocab_size = relative_vocab_size
    if relative_vocab_size is not None:
      if relative_vocab_size < 1:
        raise ValueError('`relative_vocab_size` must be positive.')
      self.relative_emb_table = nn.Parameter(
          _trunc_normal_(torch.empty(relative_vocab_size, num_heads, size_per_head), initializer_range))
      self.relative_bias_table = nn.Parameter(torch.zeros(relative_vocab_size, num_heads))
    else:
      self.relative_emb_table = None
      self.relative_bias_table = None

  def tables(self, dtype):
    if self.relative_emb_table is None:
      return None, None
    return self.relative_emb_table.to(dtype), self.relative_bias_table.to(dtype)


class QkvRelativeAttention(_RelativeTables):
  """Contract (A) core.  call(queries, keys, values, att_mask, relative_att_ids, training)."""

  def __init__(self, num_heads: int, size_per_head: int, relative_vocab_size: Optional[int] = None,
               att_dropout_prob: float = 0.0, initializer_range: float = 0.02,
               use_one_hot_lookup: bool = False, impl: str = 'auto'):
    super().__init__(relative_vocab_size, num_heads, size_per_head, initializer_range)
    self.att_dropout_prob = att_dropout_prob
    self.use_one_hot_lookup = use_one_hot_lookup  # lookup semantics are identical either way
    self.impl = impl

  def forward(self, queries, keys, values, att_mask=None, relative_att_ids=None, training=None,
              compact: Optional[ops.DenseCompactSideInputs] = None):
    if relative_att_ids is not None and self.relative_vocab_size is None:
      raise ValueError('Cannot use `relative_att_ids` without specifying `relative_vocab_size`.')
    rate = self.att_dropout_prob if resolve_training(self, training) else 0.0
    emb, bias = self.tables(queries.dtype)
    if relative_att_ids is None and compact is None:
      emb = bias = None
    return ops.dense_relative_attention(queries, keys, values, emb, bias, att_mask=att_mask,
                                        relative_att_ids=relative_att_ids, compact=compact,
                                        impl=self.impl, dropout_p=rate)


class QkvRelativeLocalAttention(_RelativeTables):
  """Long rows of contract (B): sliding window (+ optional side keys), one softmax."""

  def __init__(self, num_heads: int, size_per_head: int, local_radius: int,
               relative_vocab_size: Optional[int] = None, att_dropout_prob: float = 0.0,
               initializer_range: float = 0.02, use_one_hot_lookup: bool = False, impl: str = 'auto'):
    super().__init__(relative_vocab_size, num_heads, size_per_head, initializer_range)
    if local_radius < 1:
      raise ValueError('`local_radius` must be positive.')
    self.local_radius = local_radius
    self.att_dropout_prob = att_dropout_prob
    self.impl = impl

  def forward(self, queries, keys, values, att_mask=None, relative_att_ids=None, side_keys=None,
              side_values=None, side_att_mask=None, side_relative_att_ids=None,
              att_implementation: str = 'auto', training=None,
              compact: Optional[ops.LocalCompactSideInputs] = None):
    if (side_keys is None) != (side_values is None):
      raise ValueError('`side_keys` and `side_values` must either both be given or both be None.')
    if side_att_mask is not None and side_keys is None:
      raise ValueError('`side_keys` must be given when `side_att_mask` is.')
    if att_implementation not in ('auto', 'sparse', 'full'):
      raise ValueError('`att_implementation` must be one of ["auto", "sparse", "full"].')
    if relative_att_ids is not None and self.relative_vocab_size is None:
      raise ValueError('Cannot use `relative_att_ids` without specifying `relative_vocab_size`.')
    rate = self.att_dropout_prob if resolve_training(self, training) else 0.0
    emb, bias = self.tables(queries.dtype)
    if relative_att_ids is None and side_relative_att_ids is None and compact is None:
      emb = bias = None
    return ops.local_relative_attention(
        queries, keys, values, emb, bias, local_radius=self.local_radius, att_mask=att_mask,
        relative_att_ids=relative_att_ids, side_keys=side_keys, side_values=side_values,
        side_att_mask=side_att_mask, side_relative_att_ids=side_relative_att_ids, compact=compact,
        impl=self.impl, dropout_p=rate)


class RelativeAttention(nn.Module):
  """Projections + QkvRelativeAttention + output projection (what each dense layer uses)."""

  def __init__(self, hidden_size: int, num_heads: int, total_key_size: Optional[int] = None,
               total_value_size: Optional[int] = None, relative_vocab_size: Optional[int] = None,
               att_dropout_prob: float = 0.0, initializer_range: float = 0.02,
               use_one_hot_lookup: bool = False, impl: str = 'auto'):
    super().__init__()
    total_key_size = hidden_size if total_key_size is None else total_key_size
    total_value_size = hidden_size if total_value_size is None else total_value_size
    if total_key_size % num_heads != 0:
      raise ValueError('`total_key_size` must be a multiple of `num_heads`.')
    if total_value_size % num_heads != 0:
      raise ValueError('`total_value_size` must be a multiple of `num_heads`.')
    self.query_projection = ProjectAttentionHeads(hidden_size, num_heads, total_key_size // num_heads,
                                                  initializer_range=initializer_range)
    self.key_projection = ProjectAttentionHeads(hidden_size, num_heads, total_key_size // num_heads,
                                                initializer_range=initializer_range)
    self.value_projection = ProjectAttentionHeads(hidden_size, num_heads, total_value_size // num_heads,
                                                  initializer_range=initializer_range)
    self.qkv_relative_attention = QkvRelativeAttention(
        num_heads, total_key_size // num_heads, relative_vocab_size, att_dropout_prob,
        initializer_range, use_one_hot_lookup, impl)
    self.output_projection = nn.Linear(total_value_size, hidden_size)
    _trunc_normal_(self.output_projection.weight, initializer_range)
    nn.init.zeros_(self.output_projection.bias)

  def forward(self, from_seq, to_seq=None, att_mask=None, relative_att_ids=None, training=None,
              compact=None):
    to_seq = from_seq if to_seq is None else to_seq
    q = self.query_projection(from_seq)
    k = self.key_projection(to_seq)
    v = self.value_projection(to_seq)
    out = self.qkv_relative_attention(q, k, v, att_mask=att_mask, relative_att_ids=relative_att_ids,
                                      training=training, compact=compact)
    return self.output_projection(out.reshape(*out.shape[:2], -1))


class FusedGlobalLocalAttention(nn.Module):
  """Contract (B).  Same call signature as ETC's layer; returns ``[long_output, global_output]``."""

  def __init__(self, long_hidden_size: int, global_hidden_size: int, num_heads: int, local_radius: int,
               long_total_att_size: Optional[int] = None, global_total_att_size: Optional[int] = None,
               relative_vocab_size: Optional[int] = None, att_dropout_prob: float = 0.0,
               initializer_range: float = 0.02, share_kv_projections: bool = False,
               share_qkv_projections: bool = False, share_att_output_projection: bool = False,
               use_one_hot_lookup: bool = False, impl: str = 'auto'):
    super().__init__()
    long_total = long_hidden_size if long_total_att_size is None else long_total_att_size
    global_total = global_hidden_size if global_total_att_size is None else global_total_att_size
    if long_total % num_heads != 0 or global_total % num_heads != 0:
      raise ValueError('total attention sizes must be multiples of `num_heads`.')
    if long_total != global_total:
      raise ValueError('`long_total_att_size` and `global_total_att_size` must match '
                       '(long and global tokens attend to each other).')
    if (share_kv_projections or share_qkv_projections or share_att_output_projection) and \
       long_hidden_size != global_hidden_size:
      raise ValueError('projection sharing requires `long_hidden_size == global_hidden_size`.')
    if local_radius < 1:
      raise ValueError('`local_radius` must be positive.')
    d = long_total // num_heads
    self.num_heads, self.size_per_head, self.local_radius = num_heads, d, local_radius
    self.att_dropout_prob, self.impl = att_dropout_prob, impl
    mk = lambda n_in: ProjectAttentionHeads(n_in, num_heads, d, initializer_range=initializer_range)
    self.long_query_projection = mk(long_hidden_size)
    self.long_key_projection = mk(long_hidden_size)
    self.long_value_projection = mk(long_hidden_size)
    if share_qkv_projections:
      self.global_query_projection = self.long_query_projection
    else:
      self.global_query_projection = mk(global_hidden_size)
    if share_qkv_projections or share_kv_projections:
      self.global_key_projection = self.long_key_projection
      self.global_value_projection = self.long_value_projection
    else:
      self.global_key_projection = mk(global_hidden_size)
      self.global_value_projection = mk(global_hidden_size)
    self.long_tables = _RelativeTables(relative_vocab_size, num_heads, d, initializer_range)
    self.global_tables = _RelativeTables(relative_vocab_size, num_heads, d, initializer_range)
    self.long_output_projection = nn.Linear(long_total, long_hidden_size)
    if share_att_output_projection:
      self.global_output_projection = self.long_output_projection
    else:
      self.global_output_projection = nn.Linear(global_total, global_hidden_size)
    for lin in {self.long_output_projection, self.global_output_projection}:
      _trunc_normal_(lin.weight, initializer_range)
      nn.init.zeros_(lin.bias)

  def forward(self, long_input, global_input, l2l_att_mask=None, g2g_att_mask=None,
              l2g_att_mask=None, g2l_att_mask=None, l2l_relative_att_ids=None,
              g2g_relative_att_ids=None, l2g_relative_att_ids=None, g2l_relative_att_ids=None,
              att_implementation: str = 'auto', training=None,
              compact_side_inputs: Optional[CompactSideInputs] = None):
    if att_implementation not in ('auto', 'sparse', 'full'):
      raise ValueError('`att_implementation` must be one of ["auto", "sparse", "full"].')
    rate = self.att_dropout_prob if resolve_training(self, training) else 0.0
    lq = self.long_query_projection(long_input)
    lk = self.long_key_projection(long_input)
    lv = self.long_value_projection(long_input)
    gq = self.global_query_projection(global_input)
    gk = self.global_key_projection(global_input)
    gv = self.global_value_projection(global_input)
    if compact_side_inputs is not None:
      side = compact_side_inputs
      use_rel = True
    else:
      side = dict(l2l_att_mask=l2l_att_mask, g2g_att_mask=g2g_att_mask, l2g_att_mask=l2g_att_mask,
                  g2l_att_mask=g2l_att_mask, l2l_relative_att_ids=l2l_relative_att_ids,
                  g2g_relative_att_ids=g2g_relative_att_ids,
                  l2g_relative_att_ids=l2g_relative_att_ids,
                  g2l_relative_att_ids=g2l_relative_att_ids)
      use_rel = any(side[k] is not None for k in side if k.endswith('ids'))
      if use_rel and self.long_tables.relative_vocab_size is None:
        raise ValueError('Cannot use relative ids without specifying `relative_vocab_size`.')
    lemb, lbias = self.long_tables.tables(lq.dtype) if use_rel else (None, None)
    gemb, gbias = self.global_tables.tables(lq.dtype) if use_rel else (None, None)
    lo, go = ops.global_local_attention(lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias,
                                        local_radius=self.local_radius, side=side, impl=self.impl,
                                        dropout_p=rate)
    long_output = self.long_output_projection(lo.reshape(*lo.shape[:2], -1))
    global_output = self.global_output_projection(go.reshape(*go.shape[:2], -1))
    return [long_output, global_output]


class _ResidualFeedForward(nn.Module):

  def __init__(self, hidden_size, intermediate_size, hidden_act, hidden_dropout_prob, initializer_range):
    super().__init__()
    self.inner = nn.Linear(hidden_size, intermediate_size)
    self.outer = nn.Linear(intermediate_size, hidden_size)
    for lin in (self.inner, self.outer):
      _trunc_normal_(lin.weight, initializer_range)
      nn.init.zeros_(lin.bias)
    self.act = hidden_act
    self.hidden_dropout_prob = hidden_dropout_prob

  def forward(self, x, training=None):
    return F.dropout(self.outer(self.act(self.inner(x))), self.hidden_dropout_prob,
                     resolve_training(self, training))


class RelativeTransformerLayers(nn.Module):
  """The stack the reference instantiates (``mmt_encoder.py:124-135``): N x (attention residual
  block + feed-forward residual block), pre- or post-layer-norm.  call(inputs, att_mask,
  relative_att_ids, training) as at ``mmt_encoder.py:220-224``."""

  def __init__(self, hidden_size: int, num_hidden_layers: int, num_attention_heads: int,
               intermediate_size: Optional[int] = None, hidden_act=gelu_approximate,
               hidden_dropout_prob: float = 0.1, attention_probs_dropout_prob: float = 0.1,
               initializer_range: float = 0.02, relative_vocab_size: Optional[int] = None,
               use_pre_activation_order: bool = False, use_one_hot_lookup: bool = False,
               impl: str = 'auto', recognize_side_inputs: bool = True, id_layout_hint=None):
    super().__init__()
    if intermediate_size is None:
      intermediate_size = 4 * hidden_size
    if hidden_size % num_attention_heads != 0:
      raise ValueError('`hidden_size` must be a multiple of `num_attention_heads`.')
    self.use_pre_activation_order = use_pre_activation_order
    # Explicit [B,S,S] side inputs (the reference's signature) are checked ONCE per call of the stack against
    # the compact rules; when every element is reproduced, all layers run from the descriptors (same results,
    # no O(S^2) int32 reads per layer).  `id_layout_hint` = (num_patch_per_row, num_core_layers,
    # max_distance) names the 2-D layout to check against; None = the 1-D rule.
    self.recognize_side_inputs = recognize_side_inputs
    self.id_layout_hint = id_layout_hint
    self.attention_layers = nn.ModuleList([
        RelativeAttention(hidden_size, num_attention_heads, relative_vocab_size=relative_vocab_size,
                          att_dropout_prob=attention_probs_dropout_prob,
                          initializer_range=initializer_range, use_one_hot_lookup=use_one_hot_lookup,
                          impl=impl) for _ in range(num_hidden_layers)])
    self.feed_forward_layers = nn.ModuleList([
        _ResidualFeedForward(hidden_size, intermediate_size, hidden_act, hidden_dropout_prob,
                             initializer_range) for _ in range(num_hidden_layers)])
    self.attention_norms = nn.ModuleList([nn.LayerNorm(hidden_size, eps=1e-12) for _ in range(num_hidden_layers)])
    self.feed_forward_norms = nn.ModuleList([nn.LayerNorm(hidden_size, eps=1e-12) for _ in range(num_hidden_layers)])
    self.hidden_dropout_prob = hidden_dropout_prob
    self.output_layer_norm = nn.LayerNorm(hidden_size, eps=1e-12) if use_pre_activation_order else None

  def forward(self, inputs, att_mask=None, relative_att_ids=None, training=None, compact=None):
    tr = resolve_training(self, training)   # applies to this call only; the module mode is untouched
    drop = lambda t: F.dropout(t, self.hidden_dropout_prob, tr)
    x = inputs
    if (compact is None and self.recognize_side_inputs and att_mask is not None and relative_att_ids is not None
        and att_mask.is_cuda):
      npr, core, dist = self.id_layout_hint or (0, 0, None)
      compact = ops.compact_from_explicit_dense(att_mask, relative_att_ids, npr, core, dist)
      if compact is not None:
        att_mask = relative_att_ids = None
    for att, ffn, n1, n2 in zip(self.attention_layers, self.feed_forward_layers,
                                self.attention_norms, self.feed_forward_norms):
      if self.use_pre_activation_order:
        x = x + drop(att(n1(x), att_mask=att_mask, relative_att_ids=relative_att_ids,
                         training=tr, compact=compact))
        x = x + ffn(n2(x), training=tr)
      else:
        x = n1(x + drop(att(x, att_mask=att_mask, relative_att_ids=relative_att_ids,
                            training=tr, compact=compact)))
        x = n2(x + ffn(x, training=tr))
    if self.output_layer_norm is not None:
      x = self.output_layer_norm(x)
    return x


class GlobalLocalTransformerLayers(nn.Module):
  """Long-input stack: N x (FusedGlobalLocalAttention + feed-forward on long and global tokens)."""

  def __init__(self, long_hidden_size: int, global_hidden_size: int, num_hidden_layers: int,
               num_attention_heads: int, local_radius: int, long_intermediate_size: Optional[int] = None,
               global_intermediate_size: Optional[int] = None, hidden_act=gelu_approximate,
               hidden_dropout_prob: float = 0.1, attention_probs_dropout_prob: float = 0.1,
               initializer_range: float = 0.02, relative_vocab_size: Optional[int] = None,
               share_feed_forward_params: bool = True, share_kv_projections: bool = False,
               share_qkv_projections: bool = True, share_att_output_projection: bool = True,
               use_pre_activation_order: bool = False, use_one_hot_lookup: bool = False,
               impl: str = 'auto', recognize_side_inputs: bool = True):
    super().__init__()
    self.recognize_side_inputs = recognize_side_inputs   # see RelativeTransformerLayers
    self.local_radius = local_radius
    li = 4 * long_hidden_size if long_intermediate_size is None else long_intermediate_size
    gi = 4 * global_hidden_size if global_intermediate_size is None else global_intermediate_size
    self.use_pre_activation_order = use_pre_activation_order
    n = num_hidden_layers
    self.fused_att_layers = nn.ModuleList([
        FusedGlobalLocalAttention(long_hidden_size, global_hidden_size, num_attention_heads, local_radius,
                                  relative_vocab_size=relative_vocab_size,
                                  att_dropout_prob=attention_probs_dropout_prob,
                                  initializer_range=initializer_range,
                                  share_kv_projections=share_kv_projections,
                                  share_qkv_projections=share_qkv_projections,
                                  share_att_output_projection=share_att_output_projection,
                                  use_one_hot_lookup=use_one_hot_lookup, impl=impl) for _ in range(n)])
    self.long_ffn = nn.ModuleList([_ResidualFeedForward(long_hidden_size, li, hidden_act, hidden_dropout_prob,
                                                        initializer_range) for _ in range(n)])
    if share_feed_forward_params:
      if long_hidden_size != global_hidden_size or li != gi:
        raise ValueError('sharing feed-forward parameters requires equal long / global sizes.')
      self.global_ffn = self.long_ffn
    else:
      self.global_ffn = nn.ModuleList([_ResidualFeedForward(global_hidden_size, gi, hidden_act,
                                                            hidden_dropout_prob, initializer_range)
                                       for _ in range(n)])
    ln = lambda size: nn.ModuleList([nn.LayerNorm(size, eps=1e-12) for _ in range(n)])
    self.long_att_norms, self.global_att_norms = ln(long_hidden_size), ln(global_hidden_size)
    self.long_ffn_norms, self.global_ffn_norms = ln(long_hidden_size), ln(global_hidden_size)
    self.hidden_dropout_prob = hidden_dropout_prob
    self.long_output_norm = nn.LayerNorm(long_hidden_size, eps=1e-12) if use_pre_activation_order else None
    self.global_output_norm = nn.LayerNorm(global_hidden_size, eps=1e-12) if use_pre_activation_order else None

  def forward(self, long_input, global_input, l2l_att_mask=None, g2g_att_mask=None, l2g_att_mask=None,
              g2l_att_mask=None, l2l_relative_att_ids=None, g2g_relative_att_ids=None,
              l2g_relative_att_ids=None, g2l_relative_att_ids=None, att_implementation='auto',
              training=None, compact_side_inputs: Optional[CompactSideInputs] = None):
    tr = resolve_training(self, training)   # applies to this call only; the module mode is untouched
    drop = lambda t: F.dropout(t, self.hidden_dropout_prob, tr)
    xl, xg = long_input, global_input
    if compact_side_inputs is None and self.recognize_side_inputs and l2l_att_mask is not None and l2l_att_mask.is_cuda:
      compact_side_inputs = ops.compact_from_explicit_gl(
          dict(l2l_att_mask=l2l_att_mask, l2l_relative_att_ids=l2l_relative_att_ids, l2g_att_mask=l2g_att_mask,
               l2g_relative_att_ids=l2g_relative_att_ids, g2g_att_mask=g2g_att_mask,
               g2g_relative_att_ids=g2g_relative_att_ids, g2l_att_mask=g2l_att_mask,
               g2l_relative_att_ids=g2l_relative_att_ids), self.local_radius)
    side = dict(l2l_att_mask=l2l_att_mask, g2g_att_mask=g2g_att_mask, l2g_att_mask=l2g_att_mask,
                g2l_att_mask=g2l_att_mask, l2l_relative_att_ids=l2l_relative_att_ids,
                g2g_relative_att_ids=g2g_relative_att_ids, l2g_relative_att_ids=l2g_relative_att_ids,
                g2l_relative_att_ids=g2l_relative_att_ids, att_implementation=att_implementation,
                training=tr, compact_side_inputs=compact_side_inputs)
    for idx, att in enumerate(self.fused_att_layers):
      n_la, n_ga = self.long_att_norms[idx], self.global_att_norms[idx]
      n_lf, n_gf = self.long_ffn_norms[idx], self.global_ffn_norms[idx]
      if self.use_pre_activation_order:
        al, ag = att(n_la(xl), n_ga(xg), **side)
        xl, xg = xl + drop(al), xg + drop(ag)
        xl = xl + self.long_ffn[idx](n_lf(xl), training=tr)
        xg = xg + self.global_ffn[idx](n_gf(xg), training=tr)
      else:
        al, ag = att(xl, xg, **side)
        xl, xg = n_la(xl + drop(al)), n_ga(xg + drop(ag))
        xl = n_lf(xl + self.long_ffn[idx](xl, training=tr))
        xg = n_gf(xg + self.global_ffn[idx](xg, training=tr))
    if self.long_output_norm is not None:
      xl, xg = self.long_output_norm(xl), self.global_output_norm(xg)
    return [xl, xg]
