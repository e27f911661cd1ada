"""Host-side mirror of the side-input constructors that feed the attention op.

Mirrors, with the same names and argument meaning:

* ``MmtRelativePositionGenerator`` -- reference ``src/feature_utils.py:29-255``.
* ``RelativePositionGenerator``, ``make_segmented_att_mask``,
  ``make_local_segmented_att_mask``, ``make_global_local_transformer_side_inputs``
  -- ``etcmodel.feature_utils`` [UPSTREAM-RECALLED]; reference call sites
  ``src/feature_utils.py:86-87,178-180``, ``src/data/data_utils.py:300-301,322``,
  ``src/modeling/models/mmt_encoder.py:75-77``.
* ``make_relative_transformer_side_inputs`` / ``add_side_input_features`` --
  reference ``src/data/data_utils.py:305-332,335-379``.

Tensors are ``torch`` (vectorised; no Python loops over positions).  These
explicit ``[B,L,*]`` int32 tensors are the *drop-in* interface; the CUDA kernels
can also build the same integers in registers from the compact descriptors
(``CompactSideInputs``) so that nothing of size O(L*(2r+1+G)) ever reaches HBM.
"""

from __future__ import annotations

import dataclasses
from typing import Optional

import torch

_NUM_OTHER_RELATIVE_IDS = 3  # reference src/modeling/models/mmt_encoder.py:26


class RelativePositionGenerator:
  """1-D relative position ids with clipping [UPSTREAM-RECALLED]."""

  def __init__(self, max_distance: int):
    if max_distance < 0:
      raise ValueError('`max_distance` must not be negative.')
    self.max_distance = max_distance

  @property
  def relative_vocab_size(self) -> int:
    return 2 * self.max_distance + 1

  def _ids_from_offsets(self, offsets: torch.Tensor) -> torch.Tensor:
    d = self.max_distance
    pos = offsets.clamp(min=0, max=d)
    neg = d + (-offsets).clamp(min=0, max=d)
    return torch.where(offsets >= 0, pos, neg).to(torch.int32)

  def make_relative_att_ids(self, seq_len: int, batch_size: int = 1,
                            device=None) -> torch.Tensor:
    """int32 [batch_size, seq_len, seq_len]."""
    pos = torch.arange(seq_len, device=device)
    ids = self._ids_from_offsets(pos[None, :] - pos[:, None])
    return ids.unsqueeze(0).expand(batch_size, seq_len, seq_len).contiguous()

  def make_local_relative_att_ids(self, seq_len: int, local_radius: int,
                                  batch_size: int = 1, device=None) -> torch.Tensor:
    """int32 [batch_size, seq_len, 2*local_radius+1]; column k <-> j = i+k-r."""
    if local_radius < 1:
      raise ValueError('`local_radius` must be positive.')
    off = torch.arange(-local_radius, local_radius + 1, device=device)
    ids = self._ids_from_offsets(off)
    return ids[None, None, :].expand(batch_size, seq_len, -1).contiguous()


class MmtRelativePositionGenerator:
  """2-D image + 1-D text relative ids (reference src/feature_utils.py:29-255)."""

  def __init__(self, num_patch_per_row: int, num_core_layers: int,
               text_relative_pos_max_distance: int):
    if num_patch_per_row <= 0:
      raise ValueError('`num_patch_per_row` must be positive.')
    if num_core_layers <= 0:
      raise ValueError('`num_core_layers` must be positive.')
    if text_relative_pos_max_distance < 0:
      raise ValueError('`text_relative_pos_max_distance` must be positive.')
    self._num_patch_per_row = num_patch_per_row
    self._num_core_layers = num_core_layers
    self._core_layer_diameter = num_core_layers * 2 + 1
    text_max_id = text_relative_pos_max_distance * 2 + 1
    self._image_part_id = num_patch_per_row ** 2 + 8 + text_max_id
    self._text_part_id = self._image_part_id + 1
    self._base_tensor = self.create_base_tensor()
    self._text_relative_generator = RelativePositionGenerator(
        text_relative_pos_max_distance)

  def create_base_tensor(self) -> torch.Tensor:
    """(2N+1)^2 table: fine ids in the core, one coarse id per outer direction."""
    n_row = self._num_patch_per_row
    core = self._num_core_layers
    d = self._core_layer_diameter
    side = 2 * n_row + 1
    # Signed offsets of every cell from the centre.
    off = torch.arange(side) - n_row
    dy, dx = torch.meshgrid(off, off, indexing='ij')
    in_core = (dy.abs() <= core) & (dx.abs() <= core)
    fine = ((dy * d + dx) % (d * d)).to(torch.int32)
    # Coarse direction index, clockwise from 'top' (reference :186-255 order).
    vert = torch.sign(dy) * (dy.abs() > core)   # -1 above, +1 below, 0 inside band
    horz = torch.sign(dx) * (dx.abs() > core)
    table = {(-1, 0): 0, (-1, 1): 1, (0, 1): 2, (1, 1): 3,
             (1, 0): 4, (1, -1): 5, (0, -1): 6, (-1, -1): 7}
    coarse = torch.zeros_like(fine)
    for (v, h), k in table.items():
      coarse = torch.where((vert == v) & (horz == h),
                           torch.full_like(coarse, d * d + k), coarse)
    return torch.where(in_core, fine, coarse)

  def make_relative_att_ids(self, seq_len: int, batch_size: int = 1) -> torch.Tensor:
    """int32 [1, seq_len, seq_len] (reference returns batch 1 for the image rows)."""
    n_row = self._num_patch_per_row
    n_img = n_row * n_row
    n_txt = seq_len - n_img
    if n_txt < 0:
      raise ValueError('`seq_len` is shorter than the number of patches.')
    p = torch.arange(n_img)
    px, py = p // n_row, p % n_row
    # ids[row=(x,y), col=(u,v)] = base[n_row - x + u, n_row - y + v]
    rows = n_row - px[:, None] + px[None, :]
    cols = n_row - py[:, None] + py[None, :]
    img = self._base_tensor[rows, cols]
    out = torch.empty((seq_len, seq_len), dtype=torch.int32)
    out[:n_img, :n_img] = img
    out[:n_img, n_img:] = self._text_part_id
    out[n_img:, :n_img] = self._image_part_id
    out[n_img:, n_img:] = self._text_relative_generator.make_relative_att_ids(n_txt)[0]
    return out.unsqueeze(0)


# ---------------------------------------------------------------------------
# Masks


def example_ids_from_breakpoints(breakpoints: torch.Tensor) -> torch.Tensor:
  """``tf.cumsum(reverse=True)`` (reference src/data/data_utils.py:320-321)."""
  return torch.flip(torch.cumsum(torch.flip(breakpoints, [-1]), -1), [-1]).to(torch.int32)


def make_segmented_att_mask(segment_ids: torch.Tensor) -> torch.Tensor:
  """int32 [B,S,S]: 1 where both tokens carry the same id [UPSTREAM-RECALLED]."""
  return (segment_ids[:, :, None] == segment_ids[:, None, :]).to(torch.int32)


def make_local_segmented_att_mask(segment_ids: torch.Tensor,
                                  local_radius: int) -> torch.Tensor:
  """int32 [B,L,2r+1]: in range and same id [UPSTREAM-RECALLED]."""
  b, l = segment_ids.shape
  r = local_radius
  i = torch.arange(l, device=segment_ids.device)[:, None]
  j = i + torch.arange(-r, r + 1, device=segment_ids.device)[None, :]
  ok = (j >= 0) & (j < l)
  gathered = segment_ids[:, j.clamp(0, l - 1)]          # [B,L,2r+1]
  return ((gathered == segment_ids[:, :, None]) & ok[None]).to(torch.int32)


@dataclasses.dataclass
class RelativeTransformerSideInputs:
  """Reference src/data/data_utils.py:45-59."""
  att_mask: Optional[torch.Tensor]
  relative_att_ids: Optional[torch.Tensor]

  def to_dict(self):
    return {k: v for k, v in dataclasses.asdict(self).items() if v is not None}


def make_relative_transformer_side_inputs(long_breakpoints: torch.Tensor,
                                          relative_pos_generator,
                                          relative_pos_max_distance: int):
  """Reference src/data/data_utils.py:305-332."""
  long_example_ids = example_ids_from_breakpoints(long_breakpoints)
  att_mask = make_segmented_att_mask(long_example_ids)
  batch_size, long_seq_len = long_example_ids.shape
  relative_att_ids = None
  if relative_pos_max_distance > 0:
    relative_att_ids = relative_pos_generator.make_relative_att_ids(
        seq_len=long_seq_len, batch_size=batch_size)
  return RelativeTransformerSideInputs(att_mask=att_mask,
                                       relative_att_ids=relative_att_ids)


def add_side_input_features(num_image_wordpieces: int, num_text_wordpieces: int,
                            max_seq_len: int, relative_pos_generator,
                            relative_pos_max_distance: int):
  """Per-example side inputs, reference src/data/data_utils.py:335-379."""
  img_wp, txt_wp = num_image_wordpieces, num_text_wordpieces
  seq_len = img_wp + txt_wp
  position = torch.arange(max_seq_len, dtype=torch.int32)
  img_segment = (position < img_wp).to(torch.int32)
  txt_segment = 2 * ((position > img_wp) & (position < img_wp + txt_wp)).to(torch.int32)
  segment_ids = img_segment + txt_segment
  long_breakpoints = torch.zeros((1, max_seq_len), dtype=torch.int32)
  if 1 <= seq_len <= max_seq_len:
    long_breakpoints[0, seq_len - 1] = 1
  side = make_relative_transformer_side_inputs(
      long_breakpoints, relative_pos_generator, relative_pos_max_distance)
  out = {'segment_ids': segment_ids, 'att_mask': side.att_mask[0]}
  if side.relative_att_ids is not None:
    out['relative_att_ids'] = side.relative_att_ids[0]
  return out


@dataclasses.dataclass
class GlobalLocalTransformerSideInputs:
  """The eight side inputs of FusedGlobalLocalAttention [UPSTREAM-RECALLED]."""
  l2l_att_mask: Optional[torch.Tensor]
  g2g_att_mask: Optional[torch.Tensor]
  l2g_att_mask: Optional[torch.Tensor]
  g2l_att_mask: Optional[torch.Tensor]
  l2l_relative_att_ids: Optional[torch.Tensor]
  g2g_relative_att_ids: Optional[torch.Tensor]
  l2g_relative_att_ids: Optional[torch.Tensor]
  g2l_relative_att_ids: Optional[torch.Tensor]

  def to_dict(self):
    return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)
            if getattr(self, f.name) is not None}


def make_global_local_transformer_side_inputs_from_example_ids(
    long_example_ids: torch.Tensor, global_example_ids: torch.Tensor,
    sentence_ids: torch.Tensor, local_radius: int,
    relative_pos_max_distance: int) -> GlobalLocalTransformerSideInputs:
  """[UPSTREAM-RECALLED]; cross ids = 2D+1 + [long token i is in global g's sentence]."""
  b, l = long_example_ids.shape
  g = global_example_ids.shape[1]
  dev = long_example_ids.device
  l2l_att_mask = make_local_segmented_att_mask(long_example_ids, local_radius)
  g2g_att_mask = make_segmented_att_mask(global_example_ids)
  l2g_att_mask = (long_example_ids[:, :, None] == global_example_ids[:, None, :]).to(torch.int32)
  g2l_att_mask = l2g_att_mask.transpose(1, 2).contiguous()
  l2l_ids = g2g_ids = l2g_ids = g2l_ids = None
  if relative_pos_max_distance > 0:
    gen = RelativePositionGenerator(relative_pos_max_distance)
    l2l_ids = gen.make_local_relative_att_ids(l, local_radius, b, device=dev)
    g2g_ids = gen.make_relative_att_ids(g, b, device=dev)
    same = (sentence_ids[:, :, None] == torch.arange(g, device=dev)[None, None, :])
    l2g_ids = same.to(torch.int32) + gen.relative_vocab_size
    g2l_ids = l2g_ids.transpose(1, 2).contiguous()
  return GlobalLocalTransformerSideInputs(
      l2l_att_mask, g2g_att_mask, l2g_att_mask, g2l_att_mask,
      l2l_ids, g2g_ids, l2g_ids, g2l_ids)


def make_global_local_transformer_side_inputs(
    long_breakpoints: torch.Tensor, global_breakpoints: torch.Tensor,
    sentence_ids: torch.Tensor, local_radius: int,
    relative_pos_max_distance: int) -> GlobalLocalTransformerSideInputs:
  """[UPSTREAM-RECALLED] breakpoints -> example ids -> the eight side inputs."""
  return make_global_local_transformer_side_inputs_from_example_ids(
      example_ids_from_breakpoints(long_breakpoints),
      example_ids_from_breakpoints(global_breakpoints),
      sentence_ids, local_radius, relative_pos_max_distance)


@dataclasses.dataclass
class CompactSideInputs:
  """O(L+G) descriptors from which the kernels rebuild all eight side inputs.

  ``long_example_ids [B,L]``, ``global_example_ids [B,G]``, ``sentence_ids [B,L]``
  (all int32, on the device), plus ``relative_pos_max_distance``.  Bit-exact with
  ``make_global_local_transformer_side_inputs_from_example_ids`` by construction
  (tested on the GPU against the explicit tensors).
  """
  long_example_ids: torch.Tensor
  global_example_ids: torch.Tensor
  sentence_ids: torch.Tensor
  relative_pos_max_distance: int

  def slice(self, sl):
    """Batch slice (micro-batching in the training step)."""
    return CompactSideInputs(self.long_example_ids[sl], self.global_example_ids[sl],
                             self.sentence_ids[sl], self.relative_pos_max_distance)
