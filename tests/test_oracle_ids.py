"""Pins the integer-constructor oracle to the reference's golden vectors.

Golden source: reference src/feature_utils_test.py:25-110 (lifted by
tests/golden/make_golden_ids.py into tests/golden/relative_ids_golden.json).
"""
import json
import pathlib

import numpy as np
import pytest

from oracle import feature_oracle as fo

GOLDEN = json.loads(
    (pathlib.Path(__file__).parent / 'golden' / 'relative_ids_golden.json').read_text())


@pytest.mark.parametrize('case', GOLDEN['matrices'], ids=lambda c: c['source'].split('::')[1])
def test_golden_matrices(case):
  gen = fo.MmtRelativePositionOracle(**case['ctor'])
  got = gen.make_relative_att_ids(case['seq_len'])
  np.testing.assert_array_equal(got[None], np.asarray(case['expected'], dtype=np.int32))


def test_init_fields():
  init = GOLDEN['init']
  gen = fo.MmtRelativePositionOracle(**init['ctor'])
  assert gen.core_layer_diameter == init['core_layer_diameter']
  assert gen.image_part_id == init['image_part_id']
  assert gen.text_part_id == init['text_part_id']


def test_invalid_arguments():
  # reference src/feature_utils_test.py:37-47
  with pytest.raises(ValueError):
    fo.MmtRelativePositionOracle(0, 1, 2)
  with pytest.raises(ValueError):
    fo.MmtRelativePositionOracle(1, 0, 2)
  with pytest.raises(ValueError):
    fo.MmtRelativePositionOracle(1, 1, -1)


def test_base_tensor_docstring_examples():
  # The two base tensors printed in reference src/feature_utils_test.py:55-62,82-93.
  small = fo.MmtRelativePositionOracle(2, 1, 3).base_tensor
  np.testing.assert_array_equal(small, [[16, 9, 9, 9, 10], [15, 5, 6, 7, 11],
                                        [15, 8, 0, 1, 11], [15, 2, 3, 4, 11],
                                        [14, 13, 13, 13, 12]])
  # (The 9x9 tensor printed at reference :82-93 is for a 4-patch row; with
  # num_patch_per_row=3 the side is 2*3+1 = 7.  The expected ids at :95-108 are
  # what pin the behaviour, see test_golden_matrices.)
  large = fo.MmtRelativePositionOracle(3, 2, 9).base_tensor
  assert large.shape == (7, 7)
  np.testing.assert_array_equal(large[1:6, 1:6], [[13, 14, 15, 16, 17],
                                                  [18, 19, 20, 21, 22],
                                                  [23, 24, 0, 1, 2],
                                                  [3, 4, 5, 6, 7],
                                                  [8, 9, 10, 11, 12]])
  np.testing.assert_array_equal(large[0], [32, 25, 25, 25, 25, 25, 26])


def test_out_of_vocabulary_part_ids_for_real_config():
  # 224px / 16px -> ids 229/230 with relative_vocab_size 49.
  gen = fo.MmtRelativePositionOracle(14, 2, 12)
  assert (gen.image_part_id, gen.text_part_id) == (229, 230)


def test_1d_rule_and_local_ids():
  ids = fo.make_relative_att_ids_1d(6, 2)
  np.testing.assert_array_equal(ids[0], [0, 1, 2, 2, 2, 2])
  np.testing.assert_array_equal(ids[5], [4, 4, 4, 4, 3, 0])
  loc = fo.make_local_relative_att_ids(5, 3, 2)
  np.testing.assert_array_equal(loc[0], [4, 4, 3, 0, 1, 2, 2])
  assert (loc == loc[0]).all()


def test_masks_from_breakpoints():
  bp = fo.breakpoints_from_lengths([3, 5], 5)
  e = fo.example_ids_from_breakpoints(bp)
  np.testing.assert_array_equal(e, [[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
  m = fo.make_segmented_att_mask(e)
  # real<->real and pad<->pad are 1 (reference data_utils.py:320-322)
  assert m[0, 0, 2] == 1 and m[0, 0, 3] == 0 and m[0, 3, 4] == 1 and m[0, 4, 0] == 0
  lm = fo.make_local_segmented_att_mask(e, 1)
  np.testing.assert_array_equal(lm[0], [[0, 1, 1], [1, 1, 1], [1, 1, 0], [0, 1, 1], [1, 1, 0]])


def test_global_local_side_inputs_shapes_and_cross_ids():
  le = np.array([[1, 1, 1, 1, 0, 0]])
  ge = np.array([[1, 1, 0]])
  sid = np.array([[0, 0, 1, 1, 2, 2]])
  side = fo.make_global_local_side_inputs(le, ge, sid, local_radius=2, max_distance=3)
  assert side['l2l_att_mask'].shape == (1, 6, 5)
  assert side['l2g_att_mask'].shape == (1, 6, 3)
  assert side['g2l_att_mask'].shape == (1, 3, 6)
  assert side['g2g_att_mask'].shape == (1, 3, 3)
  voc = 7
  np.testing.assert_array_equal(side['l2g_relative_att_ids'][0, 0], [voc + 1, voc, voc])
  np.testing.assert_array_equal(side['g2l_relative_att_ids'][0, 1], [voc, voc, voc + 1, voc + 1, voc, voc])
  np.testing.assert_array_equal(side['l2g_att_mask'][0, 4], [0, 0, 1])
