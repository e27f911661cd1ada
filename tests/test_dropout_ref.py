"""Pins tests/dropout_ref.py (numpy) to the keep function the kernels compile (csrc/mlt_common.cuh),
by building that header's __host__ side into a tiny executable.  CPU only."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dropout_ref  # noqa: E402


@pytest.fixture(scope='module')
def host_binary(tmp_path_factory):
  nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(nvcc):
    pytest.skip('nvcc not available')
  out = str(tmp_path_factory.mktemp('dropout') / 'dropout_host')
  subprocess.run([nvcc, '-O1', '-std=c++17', '-o', out, os.path.join(HERE, 'cuda', 'dropout_host.cu')],
                 check=True, capture_output=True, timeout=300)
  return out


@pytest.mark.parametrize('case', [(1234567890123456789, 0.1, 2, 3, 37, 70, 0), (7, 0.5, 1, 2, 64, 33, 1),
                                  (2**63 + 11, 0.999, 1, 1, 9, 300, 0)])
def test_numpy_restatement_matches_compiled_header(host_binary, case):
  seed, p, b, h, rows, cols, rowset = case
  r = subprocess.run([host_binary] + [str(x) for x in case], check=True, capture_output=True, text=True, timeout=60)
  got = np.frombuffer(r.stdout.strip().encode(), dtype=np.uint8).reshape(b, rows, cols, h) == ord('1')
  want = dropout_ref.keep_mask(seed, p, b, h, rows, cols, rowset)
  assert np.array_equal(got, want)


def test_keep_rate_and_independence():
  k = dropout_ref.keep_mask(99, 0.1, 2, 4, 256, 512, 0)
  assert abs(k.mean() - 0.9) < 2e-3
  # heads / batch elements / row sets decorrelated
  assert abs((k[0, :, :, 0] == k[0, :, :, 1]).mean() - (0.81 + 0.01)) < 5e-3
  k1 = dropout_ref.keep_mask(99, 0.1, 2, 4, 256, 512, 1)
  assert abs((k == k1).mean() - 0.82) < 5e-3
