"""Callers on either side of the attention path.

Host-side PyTorch mirrors -- the attention core is the only custom CUDA; everything here is the
unchanged "reference glue" restated so the e2e configurations of BASELINE.json can be run:

  weighted_sparse_categorical_crossentropy_loss   reference src/modeling/losses/
                                                  weighted_sparse_categorical_crossentropy_loss.py:17-43
  MaskedLM / MaskedPP / ClassificationHead        reference src/modeling/layers/masked_patch_prediction_layer.py:23-98,
                                                  official.nlp MaskedLM / ClassificationHead [external]
  MmtPretrainingModel                             reference src/modeling/models/mmt_pretraining_model.py:23-173
  pretraining_losses                              reference src/tasks/pretraining.py:95-140
  PretrainingStep                                 reference src/tasks/pretraining.py:224-298 (fp32 micro-batch gradient
                                                  accumulation) + the implicit cross-replica gradient all-reduce of
                                                  optimizer.apply_gradients (:273) as bucketed NCCL all-reduces
                                                  overlapped with the backward pass
  retrieval_scores                                reference src/tasks/classification.py:256-334
  enumerate_image_text_pairs / shard_pairs        reference src/data/retrieval_dataloader.py:188-207
  retrieval_labels                                reference src/data/data_utils.py:744-761
  get_recall_at_k                                 reference src/prediction_helper.py:30-89 (pinned by golden vectors)
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
  """One optimizer step: micro-batch gradient accumulation in fp32 (reference
  src/tasks/pretraining.py:242-270, ``accumulated_grads = zeros_like(var, dtype=float32)``), then the
  cross-replica gradient reduction that ``optimizer.apply_gradients`` performs implicitly under the TF
  distribution strategy (:273) -- here bucketed NCCL all-reduces OVERLAPPED with the backward pass of the
  last micro-batch.

  Layout.  Parameters are taken in reverse registration order (roughly the order their gradients become
  ready).  Every ``p.grad`` is a view into one flat buffer per (dtype, device), as in DDP's bucket views; a
  second flat fp32 buffer of the same layout holds the accumulated gradients (it IS the gradient buffer when
  the parameters are fp32 already).  Buckets are contiguous ranges of ~``bucket_mb`` MB.

  Overlap.  During the last micro-batch a post-accumulate-grad hook counts down each bucket; when a bucket
  is complete its slice is folded into the fp32 accumulator and all-reduced on a side stream while autograd
  keeps producing the gradients of earlier layers.  Before the optimizer runs the step waits for every
  bucket and writes the reduced fp32 gradients back into the parameters' own dtype.

  ``scale_loss`` mirrors the reference switch (:46, :286-296, default False = gradients are SUMMED over
  the replicas, as the TF all-reduce does); True divides the loss by the number of replicas (and, exactly
  like the reference, does not apply the 1 / num_small_steps factor in that branch)."""

  def __init__(self, model: nn.Module, optimizer: torch.optim.Optimizer, micro_batch_size: int,
               scale_loss: bool = False, bucket_mb: float = 50.0, overlap: bool = True):
    self.model, self.optimizer, self.micro = model, optimizer, micro_batch_size
    self.scale_loss, self.overlap, self.bucket_mb = scale_loss, overlap, bucket_mb
    # first guess of the gradient-ready order: reverse registration order.  The first step records the order
    # in which the gradients really arrive and the buffers are laid out again in that order (what DDP's
    # bucket rebuild does): modules that register their parameters by kind (all attention layers, then all
    # feed-forward layers, then all norms) would otherwise put parameters of the first and the last layer
    # into the same bucket, and no bucket would be complete before the end of the backward pass.
    self.params = [p for p in reversed(list(model.parameters())) if p.requires_grad]
    self._armed = False
    self._pending, self._works, self._launched = [], [], []
    self._ready_order, self._ordered = [], False
    self._layout(self.params)
    for p in self.params:
      p.register_post_accumulate_grad_hook(self._on_grad)

  def _layout(self, ordered_params):
    """(Re)builds the flat gradient / fp32 accumulation buffers and the buckets for this parameter order."""
    self.groups = []       # one per (dtype, device): dict(grad=flat, acc=flat fp32, buckets=[(lo, hi)], ...)
    self._slot = {}        # param -> (group index, bucket index)
    by_key = {}
    for p in ordered_params:
      by_key.setdefault((p.dtype, p.device), []).append(p)
    for key, ps in by_key.items():
      n = sum(p.numel() for p in ps)
      grad = torch.zeros(n, dtype=key[0], device=key[1])
      acc = grad if key[0] == torch.float32 else torch.zeros(n, dtype=torch.float32, device=key[1])
      limit = max(1, int(self.bucket_mb * 2**20 / 4))
      buckets, counts, off, lo = [], [], 0, 0
      gi = len(self.groups)
      for p in ps:
        p.grad = grad[off:off + p.numel()].view_as(p)
        self._slot[p] = (gi, len(buckets))
        off += p.numel()
        if len(counts) == len(buckets):
          counts.append(0)
        counts[-1] += 1
        if off - lo >= limit:
          buckets.append((lo, off))
          lo = off
      if off > lo:
        buckets.append((lo, off))
      counts = counts[:len(buckets)]
      stream = torch.cuda.Stream(device=key[1]) if key[1].type == 'cuda' else None
      self.groups.append(dict(grad=grad, acc=acc, buckets=buckets, counts=counts, stream=stream))
    # kept for callers that inspect the flat gradient buffers (tests, checkpoints)
    self.flat = {(g['grad'].dtype, g['grad'].device): g['grad'] for g in self.groups}

  # ---- bucket machinery -----------------------------------------------------------------------
  def _world(self):
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

  def _fold_and_reduce(self, gi: int, bi: int):
    """acc[bucket] += grad[bucket] (fp32) and start the all-reduce of acc[bucket]; on CUDA both run on
    the group's side stream, ordered after the gradient kernels enqueued so far."""
    g = self.groups[gi]
    lo, hi = g['buckets'][bi]
    self._launched[gi][bi] = True

    def body():
      if g['acc'] is not g['grad']:
        g['acc'][lo:hi] += g['grad'][lo:hi]
      if self._world() > 1:
        self._works.append(dist.all_reduce(g['acc'][lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    if g['stream'] is not None:
      g['stream'].wait_stream(torch.cuda.current_stream(g['grad'].device))
      with torch.cuda.stream(g['stream']):
        body()
    else:
      body()

  def _on_grad(self, p):
    if not self._armed:
      return
    if not self._ordered:
      self._ready_order.append(p)
    gi, bi = self._slot[p]
    self._pending[gi][bi] -= 1
    if self._pending[gi][bi] == 0:
      self._fold_and_reduce(gi, bi)

  # ---- the step ---------------------------------------------------------------------------------
  def __call__(self, inputs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor], **model_kwargs):
    n = inputs['word_ids'].shape[0]
    steps = max(1, n // self.micro)
    world = self._world()
    for g in self.groups:     # zero in place: the parameters keep their views
      g['grad'].zero_()
      if g['acc'] is not g['grad']:
        g['acc'].zero_()
    total = torch.zeros((), device=inputs['word_ids'].device)
    for s in range(steps):
      last = s == steps - 1
      sl = slice(s * self.micro, (s + 1) * self.micro)
      out = self.model(**{k: v[sl] for k, v in inputs.items()}, training=True,
                       **{k: (v.slice(sl) if hasattr(v, 'slice') else v) for k, v in model_kwargs.items()})
      loss = pretraining_losses({k: v[sl] for k, v in labels.items()}, out)
      report = loss / steps
      # reference :286-296: scale_loss differentiates loss / num_replicas, else loss / num_small_steps
      objective = loss / world if self.scale_loss else report
      if last and self.overlap:
        self._pending = [list(g['counts']) for g in self.groups]
        self._launched = [[False] * len(g['buckets']) for g in self.groups]
        self._works = []
        self._armed = True
      objective.backward()
      self._armed = False
      total = total + report.detach()
      if not last:
        for g in self.groups:
          if g['acc'] is not g['grad']:
            g['acc'] += g['grad']
            g['grad'].zero_()
    # buckets the hooks did not complete (overlap off, or parameters without a gradient this step)
    if not (self.overlap and self._launched):
      self._launched = [[False] * len(g['buckets']) for g in self.groups]
      self._works = []
    for gi, g in enumerate(self.groups):
      for bi in range(len(g['buckets'])):
        if not self._launched[gi][bi]:
          self._fold_and_reduce(gi, bi)
    for w in self._works:
      w.wait()
    for g in self.groups:
      if g['stream'] is not None:
        torch.cuda.current_stream(g['grad'].device).wait_stream(g['stream'])
      if g['acc'] is not g['grad']:
        g['grad'].copy_(g['acc'])     # reduced fp32 gradients -> the parameters' dtype, for the optimizer
    self._launched = []
    self.optimizer.step()
    if not self._ordered:
      # lay the buffers out again in the observed gradient-ready order (parameters that received no gradient
      # keep their place at the end); the optimizer reads p.grad afresh every step, so the new views are safe
      self._ordered = True
      seen = set(id(p) for p in self._ready_order)
      self._layout(self._ready_order + [p for p in self.params if id(p) not in seen])
      self._ready_order = []
    return total


@torch.no_grad()
def retrieval_scores(model: nn.Module, batches, head_name: str = 'itm'):
  """softmax(itm_logits)[:, 1] per image-text pair (reference src/tasks/classification.py:286-290)."""
  scores = []
  for inputs in batches:
    out = model(**inputs, training=False)   # Keras semantics: inference for every nested layer, heads included
    scores.append(torch.softmax(out[f'{head_name}_logits'].float(), dim=-1)[:, 1])
  return torch.cat(scores)


def enumerate_image_text_pairs(num_images: int, num_texts: int):
  """All image-text combinations in the order the reference's retrieval loader emits them
  (reference src/data/retrieval_dataloader.py:188-195: the TEXT dataset is the outer interleave, the image
  dataset the inner one).  Returns ``(text_index, image_index)`` int64 tensors of length
  ``num_texts * num_images``."""
  text = torch.arange(num_texts).repeat_interleave(num_images)
  image = torch.arange(num_images).repeat(num_texts)
  return text, image


def shard_pairs(n_pairs: int, num_shards: int, shard_id: int):
  """``dataset.shard(num_input_pipelines, input_pipeline_id)`` AFTER the enumeration (reference
  retrieval_dataloader.py:204-207): shard i takes pairs i, i + n, i + 2n, ...  No collective is needed:
  every rank scores its own pairs."""
  if not 0 <= shard_id < num_shards:
    raise ValueError('`shard_id` must lie in [0, num_shards).')
  return torch.arange(shard_id, n_pairs, num_shards)


def retrieval_labels(image_index, gt_image_index, pos_weight: float = 1.0):
  """reference src/data/data_utils.py:744-761: label = (image_index == gt_image_index),
  label_weights = 1 + label * (pos_weight - 1)."""
  label = (image_index == gt_image_index).to(torch.int32)
  return label, label.to(torch.float32) * (pos_weight - 1.0) + 1.0


def get_recall_at_k(image_index, text_index, gt_image_index, output, topks=(1, 3, 5, 10)):
  """Mirror of ``get_recall_at_k_from_dataframe`` (reference src/prediction_helper.py:30-89), pinned to
  the reference function's own outputs by tests/golden/recall_golden.json.  Host-side numpy post-processing
  of the scores (as in the reference): pivot to an [images, texts] score matrix (duplicates averaged,
  missing pairs -1 / not positive), double-argsort ranks, image-to-text and text-to-image recall.
  Returns an ordered dict with the reference's keys (``'i2t @  1'`` ...) and float values."""
  import collections
  import numpy as np
  to_np = lambda t: t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
  img, txt, gt, out = (to_np(t) for t in (image_index, text_index, gt_image_index, output))
  out = out.astype(np.float64)
  u_img, ri = np.unique(img, return_inverse=True)
  u_txt, ci = np.unique(txt, return_inverse=True)
  m, n = len(u_img), len(u_txt)

  def pivot_mean(values, missing):
    total = np.zeros((m, n))
    count = np.zeros((m, n))
    np.add.at(total, (ri, ci), values)
    np.add.at(count, (ri, ci), 1.0)
    return np.where(count > 0, total / np.maximum(count, 1.0), missing)

  score_matrix = pivot_mean(out, -1.0)
  gt_matrix = pivot_mean((img == gt).astype(np.float64), 0.0)

  def rank(x, axis=-1):
    return np.argsort(np.argsort(x, axis=axis), axis=axis)

  i2t_rank = (rank(score_matrix, axis=1) - n) * -1
  t2i_rank = (rank(score_matrix, axis=0) - m) * -1
  recall = collections.OrderedDict()
  for name, rk, axis in (('i2t', i2t_rank, 1), ('t2i', t2i_rank, 0)):
    for k in topks:
      rank_at_gt = rk * gt_matrix
      match = np.clip(np.sum(((rank_at_gt <= k) & (rank_at_gt > 0)).astype(float), axis=axis), 0, 1)
      valid = np.clip(np.sum(gt_matrix, axis=axis), 0, 1)
      recall[f'{name} @ {k:>2}'] = float(np.sum(match) / np.sum(valid)) if np.sum(valid) > 0 else 0.0
  return recall
