"""Callers on either side of the attention path.

Host-side PyTorch mirrors -- the attention core is the only custom CUDA; everything here is the
unchanged "reference glue" restated so the e2e configurations of BASELINE.json can be run:

  weighted_sparse_categorical_crossentropy_loss   reference src/modeling/losses/
                                                  weighted_sparse_categorical_crossentropy_loss.py:17-43
  MaskedLM / MaskedPP / ClassificationHead        reference src/modeling/layers/masked_patch_prediction_layer.py:23-98,
                                                  official.nlp MaskedLM / ClassificationHead [external]
  MmtPretrainingModel                             reference src/modeling/models/mmt_pretraining_model.py:23-173
  pretraining_losses                              reference src/tasks/pretraining.py:95-140
  PretrainingStep                                 reference src/tasks/pretraining.py:224-298 (micro-batch gradient
                                                  accumulation) + the implicit cross-replica gradient all-reduce of
                                                  optimizer.apply_gradients (:273) as ONE bucketed NCCL all-reduce
  retrieval_scores / recall_at_k                  reference src/tasks/classification.py:256-334,
                                                  src/prediction_helper.py:30-89
"""

from __future__ import annotations

from typing import Dict, List, Optional

import torch
from torch import nn
import torch.distributed as dist
import torch.nn.functional as F

from . import layers


def weighted_sparse_categorical_crossentropy_loss(logits, labels, label_weights, pos_weights=None):
  """sum(label_w * pos_w * CE) / sum(label_w), 0 when no label is active (divide_no_nan)."""
  ce = F.cross_entropy(logits.float().reshape(-1, logits.shape[-1]), labels.reshape(-1).long(),
                       reduction='none').reshape(labels.shape)
  if pos_weights is not None:
    ce = ce * pos_weights.to(ce.dtype)
  lw = label_weights.to(ce.dtype)
  num, den = (lw * ce).sum(), lw.sum()
  return torch.where(den > 0, num / den.clamp_min(1e-30), torch.zeros_like(num))


def gather_indexes(sequence, positions):
  """reference src/tensor_utils.py:27-49: [B,S,H] x [B,M] -> [B*M, H]."""
  b, s, h = sequence.shape
  flat = (positions.long() + torch.arange(b, device=sequence.device)[:, None] * s).reshape(-1)
  return sequence.reshape(b * s, h)[flat]


class MaskedLM(nn.Module):
  """Dense + act + LN, then logits against the (bound) word-embedding table + bias."""

  def __init__(self, embedding_table: nn.Parameter, hidden_size: int, activation=layers.gelu_approximate):
    super().__init__()
    self.embedding_table = embedding_table
    self.dense = nn.Linear(hidden_size, embedding_table.shape[1])
    self.act = activation
    self.layer_norm = nn.LayerNorm(embedding_table.shape[1], eps=1e-12)
    self.bias = nn.Parameter(torch.zeros(embedding_table.shape[0]))

  def forward(self, sequence_data, masked_positions):
    x = gather_indexes(sequence_data, masked_positions)
    x = self.layer_norm(self.act(self.dense(x)))
    logits = x @ self.embedding_table.to(x.dtype).t() + self.bias.to(x.dtype)
    return logits.reshape(masked_positions.shape[0], masked_positions.shape[1], -1)


class MaskedPP(nn.Module):
  """reference masked_patch_prediction_layer.py:74-98: gather -> LN -> Dense -> + bias."""

  def __init__(self, hidden_size: int, output_num_classes: int, activation=None):
    super().__init__()
    self.layer_norm = nn.LayerNorm(hidden_size, eps=1e-12)
    # Keras Dense keeps its own bias; `output_bias` is added on top (reference :62-72, :95-97)
    self.dense = nn.Linear(hidden_size, output_num_classes, bias=True)
    nn.init.zeros_(self.dense.bias)
    self.act = activation
    self.bias = nn.Parameter(torch.zeros(output_num_classes))

  def forward(self, sequence_data, masked_positions):
    x = self.dense(self.layer_norm(gather_indexes(sequence_data, masked_positions)))
    if self.act is not None:
      x = self.act(x)
    x = x + self.bias.to(x.dtype)
    return x.reshape(masked_positions.shape[0], masked_positions.shape[1], -1)


class ClassificationHead(nn.Module):
  """CLS token -> Dense(tanh) -> dropout -> Dense(num_classes)  (TFM ClassificationHead)."""

  def __init__(self, hidden_size: int, num_classes: int, name: str = 'itm', dropout_rate: float = 0.1):
    super().__init__()
    self.name = name
    self.dense = nn.Linear(hidden_size, hidden_size)
    self.dropout_rate = dropout_rate
    self.out_proj = nn.Linear(hidden_size, num_classes)

  def forward(self, sequence_output, training=None):
    x = torch.tanh(self.dense(sequence_output[:, 0]))
    return self.out_proj(F.dropout(x, self.dropout_rate, layers.resolve_training(self, training)))


class MmtPretrainingModel(nn.Module):
  """Encoder + MLM + MPP + optional classification heads; output dict keyed like the reference."""

  def __init__(self, encoder: nn.Module, mpp_output_num_classes: int,
               classification_heads: Optional[List[ClassificationHead]] = None):
    super().__init__()
    self.encoder = encoder
    heads = classification_heads or []
    if len({h.name for h in heads}) != len(heads):
      raise ValueError('Classification heads should have unique names.')
    self.classification_heads = nn.ModuleList(heads)
    self.masked_lm = MaskedLM(encoder.word_embeddings.table.weight, encoder.hidden_size)
    self.masked_pp = MaskedPP(encoder.hidden_size, mpp_output_num_classes)

  def forward(self, word_ids, segment_ids=None, att_mask=None, relative_att_ids=None,
              patch_embeddings=None, mlm_positions=None, mpp_positions=None, training=None, **enc_kwargs):
    outputs = dict(self.encoder(word_ids, segment_ids=segment_ids, att_mask=att_mask,
                                relative_att_ids=relative_att_ids, patch_embeddings=patch_embeddings,
                                training=training, **enc_kwargs))
    seq = outputs['sequence_output']
    if mlm_positions is not None:
      outputs['mlm_logits'] = self.masked_lm(seq, mlm_positions)
    if mpp_positions is not None:
      outputs['mpp_logits'] = self.masked_pp(seq, mpp_positions)
    for head in self.classification_heads:
      outputs[f'{head.name}_logits'] = head(seq, training=training)   # Keras hands `training` to every layer
    return outputs


def pretraining_losses(labels: Dict[str, torch.Tensor], model_outputs: Dict[str, torch.Tensor]):
  """reference src/tasks/pretraining.py:95-140."""
  mlm_w, mpp_w = labels['mlm_label_weights'], labels['mpp_label_weights']
  if 'itm_label_weights' in labels:   # mask MLM / MPP losses on negative pairs
    pos = labels['itm_label_ids'].to(mlm_w.dtype)[:, None]
    mlm_w, mpp_w = mlm_w * pos, mpp_w * pos
  total = weighted_sparse_categorical_crossentropy_loss(model_outputs['mlm_logits'], labels['mlm_label_ids'], mlm_w)
  total = total + weighted_sparse_categorical_crossentropy_loss(model_outputs['mpp_logits'],
                                                                labels['mpp_label_ids'], mpp_w)
  if 'itm_label_weights' in labels:
    total = total + weighted_sparse_categorical_crossentropy_loss(
        model_outputs['itm_logits'], labels['itm_label_ids'], labels['itm_label_weights'])
  return total


class PretrainingStep:
  """One optimizer step: micro-batch accumulation (reference :242-270) then ONE gradient all-reduce
  over NCCL (what `apply_gradients` does implicitly under the TF strategy, :273).

  The gradients of all parameters live in one flat buffer per dtype (every `p.grad` is a view into
  it, as in DDP's bucket views), so the accumulation of the micro-batches writes straight into the
  buffer the collective runs on: no concatenation, no copy back, the all-reduce moves the gradients
  in their own dtype."""

  def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, micro_batch_size: int):
    self.model, self.optimizer, self.micro = model, optimizer, micro_batch_size
    self.params = [p for p in model.parameters() if p.requires_grad]
    self.flat = {}
    by_dtype = {}
    for p in self.params:
      by_dtype.setdefault((p.dtype, p.device), []).append(p)
    for key, ps in by_dtype.items():
      flat = torch.zeros(sum(p.numel() for p in ps), dtype=key[0], device=key[1])
      off = 0
      for p in ps:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
      self.flat[key] = flat

  def __call__(self, inputs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor], **model_kwargs):
    n = inputs['word_ids'].shape[0]
    steps = max(1, n // self.micro)
    for flat in self.flat.values():     # zero in place: the parameters keep their views
      flat.zero_()
    total = torch.zeros((), device=inputs['word_ids'].device)
    for s in range(steps):
      sl = slice(s * self.micro, (s + 1) * self.micro)
      out = self.model(**{k: v[sl] for k, v in inputs.items()}, training=True,
                       **{k: (v.slice(sl) if hasattr(v, 'slice') else v) for k, v in model_kwargs.items()})
      loss = pretraining_losses({k: v[sl] for k, v in labels.items()}, out) / steps
      loss.backward()
      total = total + loss.detach()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
      world = dist.get_world_size()
      works = [dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True)   # NCCL (NVLS on NVSwitch) in
               for flat in self.flat.values()]                              # production, gloo in tests
      for w in works:
        w.wait()
      for flat in self.flat.values():
        flat /= world
    self.optimizer.step()
    return total


@torch.no_grad()
def retrieval_scores(model: nn.Module, batches, head_name: str = 'itm'):
  """softmax(itm_logits)[:, 1] per image-text pair (reference src/tasks/classification.py:286-290)."""
  scores = []
  for inputs in batches:
    out = model(**inputs, training=False)
    scores.append(torch.softmax(out[f'{head_name}_logits'].float(), dim=-1)[:, 1])
  return torch.cat(scores)


def recall_at_k(scores: torch.Tensor, query_ids: torch.Tensor, is_match: torch.Tensor, ks=(1, 5, 10)):
  """Fraction of queries whose top-k scored candidates contain a match
  (reference src/prediction_helper.py:30-89)."""
  out = {}
  for k in ks:
    hits, total = 0, 0
    for q in torch.unique(query_ids):
      sel = query_ids == q
      order = torch.argsort(scores[sel], descending=True)[:k]
      hits += int(is_match[sel][order].any())
      total += 1
    out[f'recall@{k}'] = hits / max(total, 1)
  return out
