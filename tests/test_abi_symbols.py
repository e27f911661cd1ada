"""CPU checks of the C-ABI boundary: the library loads, exports every declared symbol, the
ctypes mirrors have the C layout, and argument validation returns codes (no compute)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

import mlt_b200  # noqa: F401
from mlt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mlt_attn.h')


@pytest.fixture(scope='module')
def lib():
  if not os.path.exists(_lib.LIB_PATH):
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
  return _lib.load()


def test_every_declared_symbol_is_exported(lib):
  text = open(HEADER).read()
  declared = set(re.findall(r'MLT_API [\w\s\*]+?\b(mlt_\w+)\(', text))
  assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
  for name in declared:
    assert getattr(lib, name) is not None


def test_ctypes_layout_matches_c():
  src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mlt_attn.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(mlt_tensor4), sizeof(mlt_dense_params),
         sizeof(mlt_dense_grads), sizeof(mlt_gl_params), sizeof(mlt_gl_grads),
         sizeof(mlt_id_layout), sizeof(mlt_rel_tables));
  printf("%zu %zu %zu %zu\n", offsetof(mlt_dense_params, q), offsetof(mlt_dense_params, stats),
         offsetof(mlt_dense_params, id_layout), offsetof(mlt_dense_params, workspace_bytes));
  printf("%zu %zu %zu %zu %zu\n", offsetof(mlt_gl_params, long_q), offsetof(mlt_gl_params, long_stats),
         offsetof(mlt_gl_params, side_mode), offsetof(mlt_gl_params, max_distance),
         offsetof(mlt_gl_params, workspace_bytes));
  return 0;
}'''
  with tempfile.TemporaryDirectory() as td:
    cfile = os.path.join(td, 'layout.c')
    open(cfile, 'w').write(src)
    exe = os.path.join(td, 'layout')
    subprocess.run(['gcc', '-std=c99', '-I', os.path.join(ROOT, 'include'), cfile, '-o', exe], check=True)
    lines = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split('\n')
  sizes = list(map(int, lines[0].split()))
  assert sizes == [C.sizeof(_lib.Tensor4), C.sizeof(_lib.DenseParams), C.sizeof(_lib.DenseGrads),
                   C.sizeof(_lib.GlParams), C.sizeof(_lib.GlGrads), C.sizeof(_lib.IdLayout),
                   C.sizeof(_lib.RelTables)]
  d = _lib.DenseParams
  assert list(map(int, lines[1].split())) == [d.q.offset, d.stats.offset, d.id_layout.offset,
                                              d.workspace_bytes.offset]
  g = _lib.GlParams
  assert list(map(int, lines[2].split())) == [g.long_q.offset, g.long_stats.offset, g.side_mode.offset,
                                              g.max_distance.offset, g.workspace_bytes.offset]


def test_validation_codes_without_gpu(lib):
  assert lib.mlt_abi_version() == 1
  assert lib.mlt_gl_attn_fwd(None, None) == -1  # MLT_ERR_NULL
  p = _lib.GlParams()
  p.abi_version = 1
  p.dtype = 7
  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -6  # MLT_ERR_DTYPE
  p.dtype = _lib.MLT_F32
  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -2  # MLT_ERR_SHAPE (all dims zero)
  p.B, p.L, p.G, p.H, p.d, p.R, p.local_radius = 1, 8, 2, 1, 64, 0, 2
  p.dropout_p = 1.0
  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -7  # MLT_ERR_DROPOUT: rate outside [0, 1)
  p.dropout_p = -0.1
  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -7
  p.dropout_p = 0.1                                    # a valid rate passes on to the next check

  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -1  # tensors are NULL
  p.d = 48
  assert lib.mlt_gl_attn_fwd(C.byref(p), None) == -3  # MLT_ERR_UNSUPPORTED head dim
  assert b'dropout' in lib.mlt_strerror(-7)
  assert lib.mlt_gl_workspace_bytes(C.byref(p), 1) > lib.mlt_gl_workspace_bytes(C.byref(p), 0)


def test_recognition_entry_points_validate_before_touching_the_device(lib):
  hint = _lib.IdLayout(0, 0, -1)
  one = C.c_void_p(16)   # never dereferenced: every call below fails validation first
  assert lib.mlt_dense_compact_from_explicit(None, one, 1, 8, hint, one, one, one, None) == -1   # MLT_ERR_NULL
  assert lib.mlt_dense_compact_from_explicit(one, one, 1, 8, hint, one, one, None, None) == -1
  assert lib.mlt_dense_compact_from_explicit(one, one, 0, 8, hint, one, one, one, None) == -2    # MLT_ERR_SHAPE
  assert lib.mlt_dense_compact_from_explicit(one, one, 1, 1, hint, one, one, one, None) == -2    # S < 2: no offset -1
  # a 2-D layout needs its core size and distance, and must fit the sequence
  assert lib.mlt_dense_compact_from_explicit(one, one, 1, 8, _lib.IdLayout(2, 0, 3), one, one, one, None) == -2
  assert lib.mlt_dense_compact_from_explicit(one, one, 1, 8, _lib.IdLayout(2, 1, -1), one, one, one, None) == -2
  assert lib.mlt_dense_compact_from_explicit(one, one, 1, 8, _lib.IdLayout(3, 1, 3), one, one, one, None) == -2
  arr = (C.c_void_p * 8)(*([16] * 8))
  assert lib.mlt_gl_compact_from_explicit(None, 1, 8, 2, 2, one, one, one, one, None) == -1
  assert lib.mlt_gl_compact_from_explicit(C.byref(arr), 1, 8, 2, 2, one, one, one, None, None) == -1
  assert lib.mlt_gl_compact_from_explicit(C.byref(arr), 1, 8, 2, 0, one, one, one, one, None) == -2   # radius < 1
  arr[5] = None                                                                                      # all eight required
  assert lib.mlt_gl_compact_from_explicit(C.byref(arr), 1, 8, 2, 2, one, one, one, one, None) == -1


def test_ops_refuse_cpu_tensors():
  import torch
  from mlt_b200 import ops
  q = torch.zeros(1, 8, 1, 64)
  with pytest.raises(_lib.MltLibraryError):
    ops.dense_relative_attention(q, q, q)
