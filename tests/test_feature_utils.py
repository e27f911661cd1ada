"""Host-side constructors (product) vs golden vectors and vs the loop oracle."""
import json
import pathlib

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from oracle import feature_oracle as fo

GOLDEN = json.loads(
    (pathlib.Path(__file__).parent / 'golden' / 'relative_ids_golden.json').read_text())


def test_relative_position_generator_init():
  # mirrors reference src/feature_utils_test.py:25-35
  gen = fu.MmtRelativePositionGenerator(num_patch_per_row=2, num_core_layers=1,
                                        text_relative_pos_max_distance=3)
  assert gen._num_patch_per_row == 2
  assert gen._num_core_layers == 1
  assert gen._core_layer_diameter == 3
  assert gen._image_part_id == 19
  assert gen._text_part_id == 20


def test_relative_position_generator_init_invalid_arguments():
  # mirrors reference src/feature_utils_test.py:37-47 (each call checked separately)
  for args in [(0, 1, 2), (1, 0, 2), (1, 1, -1)]:
    with pytest.raises(ValueError):
      fu.MmtRelativePositionGenerator(*args)


@pytest.mark.parametrize('case', GOLDEN['matrices'], ids=lambda c: c['source'].split('::')[1])
def test_make_relative_att_ids_golden(case):
  # mirrors reference src/feature_utils_test.py:49-110
  gen = fu.MmtRelativePositionGenerator(**case['ctor'])
  got = gen.make_relative_att_ids(case['seq_len'], 1)
  assert got.dtype == torch.int32
  np.testing.assert_array_equal(got.numpy(), np.asarray(case['expected']))


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 6), core=st.integers(1, 3), dist=st.integers(0, 5),
       extra=st.integers(0, 9))
def test_2d_generator_matches_oracle(n, core, dist, extra):
  if core > n:
    core = n
  seq = n * n + extra
  got = fu.MmtRelativePositionGenerator(n, core, dist).make_relative_att_ids(seq)[0]
  want = fo.MmtRelativePositionOracle(n, core, dist).make_relative_att_ids(seq)
  np.testing.assert_array_equal(got.numpy(), want)


@settings(max_examples=25, deadline=None)
@given(l=st.integers(1, 40), r=st.integers(1, 9), dist=st.integers(1, 6),
       g=st.integers(1, 7), seed=st.integers(0, 1000))
def test_global_local_side_inputs_match_oracle(l, r, dist, g, seed):
  rng = np.random.RandomState(seed)
  b = 2
  # packed examples: non-increasing example ids like reverse-cumsum produces
  le = np.sort(rng.randint(0, 3, size=(b, l)), axis=1)[:, ::-1].copy()
  ge = np.sort(rng.randint(0, 3, size=(b, g)), axis=1)[:, ::-1].copy()
  sid = rng.randint(0, g + 1, size=(b, l))
  got = fu.make_global_local_transformer_side_inputs_from_example_ids(
      torch.tensor(le, dtype=torch.int32), torch.tensor(ge, dtype=torch.int32),
      torch.tensor(sid, dtype=torch.int32), r, dist).to_dict()
  want = fo.make_global_local_side_inputs(le, ge, sid, r, dist)
  assert set(got) == set(want)
  for key in want:
    assert got[key].dtype == torch.int32
    np.testing.assert_array_equal(got[key].numpy(), want[key], err_msg=key)


def test_breakpoints_and_dense_side_inputs():
  lengths = [3, 6, 1]
  bp = torch.tensor(fo.breakpoints_from_lengths(lengths, 6))
  e = fu.example_ids_from_breakpoints(bp)
  np.testing.assert_array_equal(e.numpy(), fo.example_ids_from_breakpoints(bp.numpy()))
  gen = fu.RelativePositionGenerator(2)
  side = fu.make_relative_transformer_side_inputs(bp, gen, 2)
  np.testing.assert_array_equal(side.att_mask.numpy(), fo.make_segmented_att_mask(e.numpy()))
  np.testing.assert_array_equal(side.relative_att_ids[1].numpy(),
                                fo.make_relative_att_ids_1d(6, 2))
  assert set(side.to_dict()) == {'att_mask', 'relative_att_ids'}


def test_add_side_input_features_segments():
  # reference src/data/data_utils.py:350-361: image=1, text=2 (position img_wp itself is 0)
  out = fu.add_side_input_features(3, 2, 8, fu.RelativePositionGenerator(2), 2)
  np.testing.assert_array_equal(out['segment_ids'].numpy(), [1, 1, 1, 0, 2, 0, 0, 0])
  assert out['att_mask'].shape == (8, 8) and out['relative_att_ids'].shape == (8, 8)
  assert out['att_mask'][0, 4] == 1 and out['att_mask'][0, 5] == 0
