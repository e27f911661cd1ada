"""Integer side-input constructors, restated on the CPU with NumPy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PINNED: reproduces the
reference's golden matrices (``src/feature_utils_test.py:64-72,95-108``).

Every function cites the reference ``file:line`` it follows (paths relative to
the reference tree).  ``[UPSTREAM-RECALLED]`` marks rules that live in the
un-vendored ``etcmodel`` package and are restated from the published algorithm.

Written with explicit Python loops on purpose: the product's constructors
(``multimodal-long-transformer-2021_b200/feature_utils.py`` and the in-kernel
closed forms) are vectorised, so agreement is between independent formulations.
"""

from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# Named constants of the upstream scheme (kept in one place so a later
# correction is one line).  mmt_encoder.py:26 reserves three ids after the
# 2D+1 positional ones.
NUM_OTHER_RELATIVE_IDS = 3  # src/modeling/models/mmt_encoder.py:26


def relative_vocab_size_1d(max_distance: int) -> int:
  """[UPSTREAM-RECALLED] RelativePositionGenerator.relative_vocab_size = 2D+1.

  Call site: src/modeling/models/mmt_encoder.py:75-77.
  """
  return 2 * max_distance + 1


def relative_id_1d(offset: int, max_distance: int) -> int:
  """[UPSTREAM-RECALLED] 1-D relative id of key offset ``o = j - i``.

  ``o >= 0 -> min(o, D)``;  ``o < 0 -> D + min(-o, D)``.  Pinned by the text
  blocks of src/feature_utils_test.py:64-72 (rows ``[0,1,2]``, ``[4,0,1]``,
  ``[5,4,0]`` for D=3) and :95-108 (``[10,0,1]``, ``[11,10,0]`` for D=9).
  """
  d = max_distance
  if offset >= 0:
    return min(offset, d)
  return d + min(-offset, d)


def make_relative_att_ids_1d(seq_len: int, max_distance: int) -> np.ndarray:
  """[UPSTREAM-RECALLED] RelativePositionGenerator.make_relative_att_ids.

  Call sites: src/data/data_utils.py:300-301,326-329; src/feature_utils.py:178-180.
  Returns int32 [seq_len, seq_len] (the reference tiles it over batch).
  """
  out = np.zeros((seq_len, seq_len), dtype=np.int32)
  for i in range(seq_len):
    for j in range(seq_len):
      out[i, j] = relative_id_1d(j - i, max_distance)
  return out


def make_local_relative_att_ids(seq_len: int, local_radius: int,
                                max_distance: int) -> np.ndarray:
  """[UPSTREAM-RECALLED] RelativePositionGenerator.make_local_relative_att_ids.

  int32 [seq_len, 2r+1]; column k is key j = i + k - r, so every row is the
  same vector (SURVEY.md section 8-spec "local ids").
  """
  r = local_radius
  row = [relative_id_1d(k - r, max_distance) for k in range(2 * r + 1)]
  return np.tile(np.asarray(row, dtype=np.int32)[None, :], (seq_len, 1))


# --------------------------------------------------------------------------
# 2-D (image) + 1-D (text) generator: src/feature_utils.py:29-255


class MmtRelativePositionOracle:
  """Restates MmtRelativePositionGenerator (src/feature_utils.py:29-255)."""

  def __init__(self, num_patch_per_row: int, num_core_layers: int,
               text_relative_pos_max_distance: int):
    # Argument checks: src/feature_utils.py:61-66.
    if num_patch_per_row <= 0:
      raise ValueError('`num_patch_per_row` must be positive.')
    if num_core_layers <= 0:
      raise ValueError('`num_core_layers` must be positive.')
    if text_relative_pos_max_distance < 0:
      raise ValueError('`text_relative_pos_max_distance` must be positive.')
    self.num_patch_per_row = num_patch_per_row
    self.num_core_layers = num_core_layers
    self.core_layer_diameter = 2 * num_core_layers + 1  # :72
    self.max_distance = text_relative_pos_max_distance
    text_max_id = 2 * text_relative_pos_max_distance + 1  # :75
    # Part ids: src/feature_utils.py:78-82 (8 = number of coarse directions).
    self.image_part_id = num_patch_per_row ** 2 + 8 + text_max_id
    self.text_part_id = self.image_part_id + 1
    self.base_tensor = self._create_base_tensor()

  def _create_base_tensor(self) -> np.ndarray:
    """src/feature_utils.py:89-112 with the direction table of :186-255."""
    r = self.num_core_layers
    d = self.core_layer_diameter
    npr = self.num_patch_per_row
    n = npr - r
    m = npr + r + 1
    side = 2 * npr + 1
    base = np.zeros((side, side), dtype=np.int32)
    # Fine-grained core: ids 0..d*d-1 rolled so that id 0 sits at the centre.
    for a in range(d):
      for b in range(d):
        flat = a * d + b
        base[n + a, n + b] = (flat - (d * r + r)) % (d * d)
    # Eight coarse directions, in the reference's dict order (ids d*d ...).
    # Each entry: (row_begin, row_end, col_begin, col_end) of the filled block.
    blocks = [
        (0, n, n, n + d),        # top
        (0, n, m, side),         # top_right
        (n, n + d, m, side),     # right
        (m, side, m, side),      # right_bottom
        (m, side, n, n + d),     # bottom
        (m, side, 0, n),         # bottom_left
        (n, n + d, 0, n),        # left
        (0, n, 0, n),            # top_left
    ]
    for k, (r0, r1, c0, c1) in enumerate(blocks):
      base[r0:r1, c0:c1] += d * d + k
    return base

  def make_relative_att_ids(self, seq_len: int) -> np.ndarray:
    """src/feature_utils.py:114-184.  Returns int32 [seq_len, seq_len]."""
    npr = self.num_patch_per_row
    n_img = npr * npr
    n_txt = seq_len - n_img
    out = np.zeros((seq_len, seq_len), dtype=np.int32)
    # Image rows: patch (x, y) reads the npr x npr window at (npr-x, npr-y)
    # (:160-170); text columns get text_part_id (:172-175).
    for x in range(npr):
      for y in range(npr):
        row = x * npr + y
        win = self.base_tensor[npr - x:2 * npr - x, npr - y:2 * npr - y]
        out[row, :n_img] = win.reshape(-1)
        out[row, n_img:] = self.text_part_id
    # Text rows: image columns get image_part_id, text block is 1-D (:178-184).
    for i in range(n_txt):
      out[n_img + i, :n_img] = self.image_part_id
      for j in range(n_txt):
        out[n_img + i, n_img + j] = relative_id_1d(j - i, self.max_distance)
    return out


# --------------------------------------------------------------------------
# Masks: src/data/data_utils.py:305-332,350-377


def example_ids_from_breakpoints(breakpoints: np.ndarray) -> np.ndarray:
  """Reverse cumulative sum (src/data/data_utils.py:320-321)."""
  bp = np.asarray(breakpoints)
  out = np.zeros_like(bp)
  for b in range(bp.shape[0]):
    acc = 0
    for i in range(bp.shape[1] - 1, -1, -1):
      acc += int(bp[b, i])
      out[b, i] = acc
  return out


def breakpoints_from_lengths(lengths, max_seq_len: int) -> np.ndarray:
  """one_hot(seq_len - 1) (src/data/data_utils.py:364-366)."""
  lengths = np.asarray(lengths)
  out = np.zeros((lengths.shape[0], max_seq_len), dtype=np.int32)
  for b, n in enumerate(lengths):
    if 1 <= n <= max_seq_len:
      out[b, n - 1] = 1
  return out


def make_segmented_att_mask(example_ids: np.ndarray) -> np.ndarray:
  """[UPSTREAM-RECALLED] make_segmented_att_mask (call: data_utils.py:322).

  ``mask[b,i,j] = (e[b,i] == e[b,j])`` so real<->real and pad<->pad are 1.
  """
  e = np.asarray(example_ids)
  b_sz, s = e.shape
  out = np.zeros((b_sz, s, s), dtype=np.int32)
  for b in range(b_sz):
    for i in range(s):
      for j in range(s):
        out[b, i, j] = 1 if e[b, i] == e[b, j] else 0
  return out


def make_local_segmented_att_mask(example_ids: np.ndarray,
                                  local_radius: int) -> np.ndarray:
  """[UPSTREAM-RECALLED] make_local_segmented_att_mask.

  int32 [B, L, 2r+1]; 1 iff ``0 <= j < L`` and ``e[b,j] == e[b,i]`` with
  ``j = i + k - r``.
  """
  e = np.asarray(example_ids)
  b_sz, l = e.shape
  r = local_radius
  out = np.zeros((b_sz, l, 2 * r + 1), dtype=np.int32)
  for b in range(b_sz):
    for i in range(l):
      for k in range(2 * r + 1):
        j = i + k - r
        if 0 <= j < l and e[b, j] == e[b, i]:
          out[b, i, k] = 1
  return out


def make_global_local_side_inputs(long_example_ids, global_example_ids,
                                  sentence_ids, local_radius: int,
                                  max_distance: int):
  """[UPSTREAM-RECALLED] make_global_local_transformer_side_inputs.

  Cross ids (the reference never builds these): l2g and g2l
  ids are ``2D+1 + [sentence_ids[b,i] == g]`` -- "long token i belongs to global
  token g's sentence" -- offset by the positional vocabulary so they do not
  collide with the l2l / g2g positional ids sharing the same table.
  Returns a dict of int32 arrays.
  """
  le = np.asarray(long_example_ids)
  ge = np.asarray(global_example_ids)
  sid = np.asarray(sentence_ids)
  b_sz, l = le.shape
  g = ge.shape[1]
  d = max_distance
  voc = relative_vocab_size_1d(d)
  l2g_mask = np.zeros((b_sz, l, g), dtype=np.int32)
  l2g_ids = np.zeros((b_sz, l, g), dtype=np.int32)
  for b in range(b_sz):
    for i in range(l):
      for k in range(g):
        l2g_mask[b, i, k] = 1 if le[b, i] == ge[b, k] else 0
        l2g_ids[b, i, k] = voc + (1 if sid[b, i] == k else 0)
  return dict(
      l2l_att_mask=make_local_segmented_att_mask(le, local_radius),
      g2g_att_mask=make_segmented_att_mask(ge),
      l2g_att_mask=l2g_mask,
      g2l_att_mask=np.ascontiguousarray(l2g_mask.transpose(0, 2, 1)),
      l2l_relative_att_ids=np.tile(
          make_local_relative_att_ids(l, local_radius, d)[None], (b_sz, 1, 1)),
      g2g_relative_att_ids=np.tile(
          make_relative_att_ids_1d(g, d)[None], (b_sz, 1, 1)),
      l2g_relative_att_ids=l2g_ids,
      g2l_relative_att_ids=np.ascontiguousarray(l2g_ids.transpose(0, 2, 1)),
  )
