"""Host-side mirror of the reference encoder, the caller of the attention path (SURVEY next-1).

``MmtEncoder`` keeps the constructor arguments, call signature, error behaviour and output
dict of reference ``src/modeling/models/mmt_encoder.py:45-237`` (PyTorch modules stand in for
Keras layers).  Everything except the attention core is ordinary host-framework code
(embedding lookups, LayerNorm, Dense, GELU-tanh): the north star leaves those unchanged.

Additive long-input fields (not in the reference, defaults reproduce it): ``local_radius`` and
``num_global_tokens`` switch the stack to ``GlobalLocalTransformerLayers``; the reference's
dense ``[B,S,S]`` side inputs are still accepted as-is by the dense stack.
"""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import layers
from .feature_utils import RelativePositionGenerator

_NUM_OTHER_RELATIVE_IDS = 3  # reference mmt_encoder.py:26


class EmbeddingLookup(nn.Module):
  """etc_layers.EmbeddingLookup [UPSTREAM-RECALLED]: table (+ projection when sizes differ)."""

  def __init__(self, vocab_size, embedding_size, projection_size=None, initializer_range=0.02):
    super().__init__()
    self.table = nn.Embedding(vocab_size, embedding_size)
    layers._trunc_normal_(self.table.weight, initializer_range)
    self.projection = None
    if projection_size is not None and projection_size != embedding_size:
      self.projection = nn.Linear(embedding_size, projection_size, bias=False)
      layers._trunc_normal_(self.projection.weight, initializer_range)

  def forward(self, ids):
    x = self.table(ids)
    return x if self.projection is None else self.projection(x)


class MmtEncoder(nn.Module):
  """Multimodal transformer encoder (reference mmt_encoder.py:33-237)."""

  def __init__(self, vocab_size: int, segment_vocab_size: int = 16, embedding_size: Optional[int] = None,
               hidden_size: int = 768, num_hidden_layers: int = 12, num_attention_heads: int = 12,
               intermediate_size: int = 3072, inner_activation=layers.gelu_approximate,
               hidden_dropout_prob: float = 0.1, attention_probs_dropout_prob: float = 0.1,
               max_absolute_position_embeddings: Optional[int] = None, relative_vocab_size: int = 32,
               relative_pos_max_distance: int = 12, initializer_range: float = 0.02,
               use_pre_activation_order: bool = False, use_one_hot_lookup: bool = True,
               use_pooler_layer: bool = False, patch_embedding_size: int = 768,
               local_radius: Optional[int] = None, num_global_tokens: int = 0, impl: str = 'auto',
               recognize_side_inputs: bool = True, id_layout_hint=None):
    super().__init__()
    # reference mmt_encoder.py:69-80
    if relative_vocab_size is None:
      if relative_pos_max_distance != 0:
        raise ValueError('`relative_pos_max_distance` must be 0 when `relative_vocab_size` is None.')
    elif relative_vocab_size < (RelativePositionGenerator(relative_pos_max_distance).relative_vocab_size +
                                _NUM_OTHER_RELATIVE_IDS):
      raise ValueError(f'`relative_vocab_size` ({relative_vocab_size}) too small for '
                       f'`relative_pos_max_distance` ({relative_pos_max_distance}')
    if embedding_size is None:
      embedding_size = hidden_size
    self.hidden_size = hidden_size
    self.word_embeddings = EmbeddingLookup(vocab_size, embedding_size, hidden_size, initializer_range)
    self.segment_embeddings = EmbeddingLookup(segment_vocab_size, embedding_size, hidden_size,
                                              initializer_range)
    self.position_embeddings = None
    if max_absolute_position_embeddings is not None:
      self.position_embeddings = nn.Parameter(layers._trunc_normal_(
          torch.empty(max_absolute_position_embeddings, hidden_size), initializer_range))
    self.patch_embedding_projection = nn.Linear(patch_embedding_size, hidden_size)
    layers._trunc_normal_(self.patch_embedding_projection.weight, initializer_range)
    nn.init.zeros_(self.patch_embedding_projection.bias)
    self.embedding_norm = nn.LayerNorm(hidden_size, eps=1e-12)
    self.hidden_dropout_prob = hidden_dropout_prob
    self.local_radius = local_radius
    self.num_global_tokens = num_global_tokens
    common = dict(hidden_act=inner_activation, hidden_dropout_prob=hidden_dropout_prob,
                  attention_probs_dropout_prob=attention_probs_dropout_prob,
                  initializer_range=initializer_range, relative_vocab_size=relative_vocab_size,
                  use_pre_activation_order=use_pre_activation_order,
                  use_one_hot_lookup=use_one_hot_lookup, impl=impl,
                  recognize_side_inputs=recognize_side_inputs)
    # Additive fields (not in the reference): explicit [B,S,S] side inputs are checked once per call against the
    # compact rules and, when reproduced exactly, replaced by descriptors for all layers (layers.py);
    # `id_layout_hint` = (num_patch_per_row, num_core_layers, max_distance) names the 2-D layout of the data
    # pipeline (reference src/feature_utils.py:29-255) to check the ids against.
    if local_radius is None:
      self.transformer_layers = layers.RelativeTransformerLayers(
          hidden_size=hidden_size, num_hidden_layers=num_hidden_layers,
          num_attention_heads=num_attention_heads, intermediate_size=intermediate_size,
          id_layout_hint=id_layout_hint, **common)
    else:
      if num_global_tokens < 1:
        raise ValueError('`num_global_tokens` must be positive when `local_radius` is set.')
      self.global_embeddings = nn.Parameter(layers._trunc_normal_(
          torch.empty(1, hidden_size), initializer_range))
      self.transformer_layers = layers.GlobalLocalTransformerLayers(
          long_hidden_size=hidden_size, global_hidden_size=hidden_size,
          num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
          local_radius=local_radius, long_intermediate_size=intermediate_size,
          global_intermediate_size=intermediate_size, **common)
    self.pooler = None
    if use_pooler_layer:
      self.pooler = nn.Linear(hidden_size, hidden_size)
      layers._trunc_normal_(self.pooler.weight, initializer_range)
      nn.init.zeros_(self.pooler.bias)

  def embed(self, word_ids, segment_ids=None, patch_embeddings=None, training=None):
    """Reference mmt_encoder.py:189-218."""
    if segment_ids is None:
      segment_ids = torch.ones_like(word_ids)
    word = torch.nn.functional.dropout(self.embedding_norm(self.word_embeddings(word_ids)),
                                       self.hidden_dropout_prob, layers.resolve_training(self, training))
    emb = word + self.segment_embeddings(segment_ids)
    if self.position_embeddings is not None:
      emb = emb + self.position_embeddings[:emb.shape[1]].unsqueeze(0)
    if patch_embeddings is not None:
      seq_len, patch_len = emb.shape[1], patch_embeddings.shape[1]
      pe = self.patch_embedding_projection(patch_embeddings.to(emb.dtype))
      # 2 leading slots for [CLS] and [PATCH] (reference :211-217)
      pe = torch.nn.functional.pad(pe, (0, 0, 2, seq_len - 2 - patch_len))
      emb = emb + pe
    return emb

  def forward(self, word_ids, segment_ids=None, att_mask=None, relative_att_ids=None,
              patch_embeddings=None, training=None, compact=None, compact_side_inputs=None):
    # Keras semantics: `training` applies to this call and every nested layer; the module mode is
    # only the default when it is None (reference passes training=True at src/tasks/pretraining.py:279)
    tr = layers.resolve_training(self, training)
    emb = self.embed(word_ids, segment_ids, patch_embeddings, training=tr)
    if self.local_radius is None:
      out = self.transformer_layers(emb, att_mask=att_mask, relative_att_ids=relative_att_ids,
                                    training=tr, compact=compact)
      global_out = None
    else:
      if compact_side_inputs is None:
        raise ValueError('the long-input stack needs `compact_side_inputs`.')
      glob = self.global_embeddings.to(emb.dtype).expand(emb.shape[0], self.num_global_tokens, -1)
      out, global_out = self.transformer_layers(emb, glob, training=tr,
                                                compact_side_inputs=compact_side_inputs)
    outputs = {'sequence_output': out}
    if global_out is not None:
      outputs['global_output'] = global_out
    if self.pooler is not None:
      outputs['pooled_output'] = torch.tanh(self.pooler(out[:, 0]))
    return outputs
