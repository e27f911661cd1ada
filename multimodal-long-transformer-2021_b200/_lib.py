"""ctypes binding of ``include/mlt_attn.h`` (``libmlt_attn.so``).

The structures below mirror the C declarations field by field.  There is NO
fallback: if the library is missing, loading raises with build instructions.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmlt_attn.so')

MLT_ABI_VERSION = 1
MLT_F32, MLT_BF16 = 0, 1
MLT_SIDE_EXPLICIT, MLT_SIDE_COMPACT = 0, 1
MLT_IMPL_AUTO, MLT_IMPL_SIMT, MLT_IMPL_TC, MLT_IMPL_TC_GENERIC = 0, 1, 2, 3
IMPL = {'auto': MLT_IMPL_AUTO, 'simt': MLT_IMPL_SIMT, 'tc': MLT_IMPL_TC, 'tc_generic': MLT_IMPL_TC_GENERIC}


class Tensor4(C.Structure):
  _fields_ = [('ptr', C.c_void_p), ('stride_b', C.c_int64), ('stride_l', C.c_int64),
              ('stride_h', C.c_int64)]


class RelTables(C.Structure):
  _fields_ = [('emb', C.c_void_p), ('bias', C.c_void_p)]


class IdLayout(C.Structure):
  _fields_ = [('num_patch_per_row', C.c_int32), ('num_core_layers', C.c_int32),
              ('max_distance', C.c_int32)]


class DenseParams(C.Structure):
  _fields_ = [
      ('abi_version', C.c_int32), ('dtype', C.c_int32), ('impl', C.c_int32),
      ('B', C.c_int32), ('Lq', C.c_int32), ('Lk', C.c_int32), ('H', C.c_int32),
      ('d', C.c_int32), ('R', C.c_int32),
      ('scale', C.c_float), ('neg', C.c_float), ('dropout_p', C.c_float),
      ('dropout_seed', C.c_uint64),
      ('q', Tensor4), ('k', Tensor4), ('v', Tensor4), ('out', Tensor4),
      ('stats', C.c_void_p), ('tables', RelTables), ('side_mode', C.c_int32),
      ('att_mask', C.c_void_p), ('relative_att_ids', C.c_void_p),
      ('q_example_ids', C.c_void_p), ('k_example_ids', C.c_void_p),
      ('id_layout', IdLayout), ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
  ]


class DenseGrads(C.Structure):
  _fields_ = [('d_out', Tensor4), ('d_q', Tensor4), ('d_k', Tensor4), ('d_v', Tensor4),
              ('d_emb', C.c_void_p), ('d_bias', C.c_void_p)]


class GlParams(C.Structure):
  _fields_ = [
      ('abi_version', C.c_int32), ('dtype', C.c_int32), ('impl', C.c_int32),
      ('B', C.c_int32), ('L', C.c_int32), ('G', C.c_int32), ('H', C.c_int32),
      ('d', C.c_int32), ('R', C.c_int32), ('local_radius', C.c_int32),
      ('scale', C.c_float), ('neg', C.c_float), ('dropout_p', C.c_float),
      ('dropout_seed', C.c_uint64),
      ('long_q', Tensor4), ('long_k', Tensor4), ('long_v', Tensor4),
      ('global_q', Tensor4), ('global_k', Tensor4), ('global_v', Tensor4),
      ('long_out', Tensor4), ('global_out', Tensor4),
      ('long_stats', C.c_void_p), ('global_stats', C.c_void_p),
      ('long_tables', RelTables), ('global_tables', RelTables),
      ('side_mode', C.c_int32),
      ('l2l_att_mask', C.c_void_p), ('l2l_relative_att_ids', C.c_void_p),
      ('l2g_att_mask', C.c_void_p), ('l2g_relative_att_ids', C.c_void_p),
      ('g2g_att_mask', C.c_void_p), ('g2g_relative_att_ids', C.c_void_p),
      ('g2l_att_mask', C.c_void_p), ('g2l_relative_att_ids', C.c_void_p),
      ('long_example_ids', C.c_void_p), ('global_example_ids', C.c_void_p),
      ('sentence_ids', C.c_void_p), ('max_distance', C.c_int32),
      ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
  ]


class GlGrads(C.Structure):
  _fields_ = [
      ('d_long_out', Tensor4), ('d_global_out', Tensor4),
      ('d_long_q', Tensor4), ('d_long_k', Tensor4), ('d_long_v', Tensor4),
      ('d_global_q', Tensor4), ('d_global_k', Tensor4), ('d_global_v', Tensor4),
      ('d_long_emb', C.c_void_p), ('d_long_bias', C.c_void_p),
      ('d_global_emb', C.c_void_p), ('d_global_bias', C.c_void_p),
  ]


class LocalParams(C.Structure):
  _fields_ = [
      ('abi_version', C.c_int32), ('dtype', C.c_int32), ('impl', C.c_int32),
      ('B', C.c_int32), ('L', C.c_int32), ('G', C.c_int32), ('H', C.c_int32),
      ('d', C.c_int32), ('R', C.c_int32), ('local_radius', C.c_int32),
      ('scale', C.c_float), ('neg', C.c_float), ('dropout_p', C.c_float),
      ('dropout_seed', C.c_uint64),
      ('q', Tensor4), ('k', Tensor4), ('v', Tensor4), ('side_k', Tensor4), ('side_v', Tensor4),
      ('out', Tensor4), ('stats', C.c_void_p), ('tables', RelTables), ('side_mode', C.c_int32),
      ('att_mask', C.c_void_p), ('relative_att_ids', C.c_void_p),
      ('side_att_mask', C.c_void_p), ('side_relative_att_ids', C.c_void_p),
      ('example_ids', C.c_void_p), ('side_example_ids', C.c_void_p), ('sentence_ids', C.c_void_p),
      ('max_distance', C.c_int32), ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t),
  ]


class LocalGrads(C.Structure):
  _fields_ = [('d_out', Tensor4), ('d_q', Tensor4), ('d_k', Tensor4), ('d_v', Tensor4),
              ('d_side_k', Tensor4), ('d_side_v', Tensor4), ('d_emb', C.c_void_p),
              ('d_bias', C.c_void_p)]


class KernelTime(C.Structure):
  _fields_ = [('name', C.c_char * 48), ('ms', C.c_float), ('flops', C.c_double),
              ('bytes', C.c_double)]


# Every symbol include/mlt_attn.h declares (checked by tests/test_abi_symbols.py).
EXPORTS = (
    'mlt_abi_version', 'mlt_strerror', 'mlt_gl_uses_tensor_cores',
    'mlt_dense_uses_tensor_cores', 'mlt_dense_workspace_bytes', 'mlt_gl_workspace_bytes',
    'mlt_dense_rel_attn_fwd', 'mlt_dense_rel_attn_bwd', 'mlt_gl_attn_fwd', 'mlt_gl_attn_bwd',
    'mlt_build_dense_side_inputs', 'mlt_build_gl_side_inputs',
    'mlt_dense_compact_from_explicit', 'mlt_gl_compact_from_explicit',
    'mlt_profile_enable', 'mlt_profile_read', 'mlt_launch_count',
    'mlt_local_workspace_bytes', 'mlt_local_rel_attn_fwd', 'mlt_local_rel_attn_bwd',
)

_lib = None


class MltLibraryError(RuntimeError):
  pass


def load() -> C.CDLL:
  """Loads libmlt_attn.so; raises (never falls back) when it is not built."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise MltLibraryError(
        f'{LIB_PATH} is missing: the CUDA library is not built. Run '
        '`python -c "import __graft_entry__ as g; g.build()"` (or `make` in csrc/). '
        'There is no CPU fallback.')
  lib = C.CDLL(LIB_PATH)
  lib.mlt_abi_version.restype = C.c_int
  lib.mlt_strerror.restype = C.c_char_p
  lib.mlt_strerror.argtypes = [C.c_int]
  for name in ('mlt_gl_uses_tensor_cores',):
    getattr(lib, name).argtypes = [C.POINTER(GlParams)]
    getattr(lib, name).restype = C.c_int
  lib.mlt_dense_uses_tensor_cores.argtypes = [C.POINTER(DenseParams)]
  lib.mlt_dense_uses_tensor_cores.restype = C.c_int
  lib.mlt_dense_workspace_bytes.argtypes = [C.POINTER(DenseParams), C.c_int]
  lib.mlt_dense_workspace_bytes.restype = C.c_size_t
  lib.mlt_gl_workspace_bytes.argtypes = [C.POINTER(GlParams), C.c_int]
  lib.mlt_gl_workspace_bytes.restype = C.c_size_t
  lib.mlt_dense_rel_attn_fwd.argtypes = [C.POINTER(DenseParams), C.c_void_p]
  lib.mlt_dense_rel_attn_bwd.argtypes = [C.POINTER(DenseParams), C.POINTER(DenseGrads), C.c_void_p]
  lib.mlt_gl_attn_fwd.argtypes = [C.POINTER(GlParams), C.c_void_p]
  lib.mlt_gl_attn_bwd.argtypes = [C.POINTER(GlParams), C.POINTER(GlGrads), C.c_void_p]
  lib.mlt_build_dense_side_inputs.argtypes = [
      C.c_void_p, C.c_int32, C.c_int32, IdLayout, C.c_void_p, C.c_void_p, C.c_void_p]
  lib.mlt_build_gl_side_inputs.argtypes = [
      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
      C.c_int32, C.POINTER(C.c_void_p * 8), C.c_void_p]
  lib.mlt_dense_compact_from_explicit.argtypes = [
      C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, IdLayout, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
  lib.mlt_gl_compact_from_explicit.argtypes = [
      C.POINTER(C.c_void_p * 8), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
      C.c_void_p, C.c_void_p, C.c_void_p]
  for name in ('mlt_dense_rel_attn_fwd', 'mlt_dense_rel_attn_bwd', 'mlt_gl_attn_fwd',
               'mlt_gl_attn_bwd', 'mlt_build_dense_side_inputs', 'mlt_build_gl_side_inputs',
               'mlt_dense_compact_from_explicit', 'mlt_gl_compact_from_explicit'):
    getattr(lib, name).restype = C.c_int
  lib.mlt_local_workspace_bytes.argtypes = [C.POINTER(LocalParams), C.c_int]
  lib.mlt_local_workspace_bytes.restype = C.c_size_t
  lib.mlt_local_rel_attn_fwd.argtypes = [C.POINTER(LocalParams), C.c_void_p]
  lib.mlt_local_rel_attn_fwd.restype = C.c_int
  lib.mlt_local_rel_attn_bwd.argtypes = [C.POINTER(LocalParams), C.POINTER(LocalGrads), C.c_void_p]
  lib.mlt_local_rel_attn_bwd.restype = C.c_int
  lib.mlt_profile_enable.argtypes = [C.c_int]
  lib.mlt_profile_read.argtypes = [C.POINTER(KernelTime), C.c_int]
  lib.mlt_launch_count.restype = C.c_longlong
  if lib.mlt_abi_version() != MLT_ABI_VERSION:
    raise MltLibraryError('libmlt_attn.so ABI version mismatch; rebuild it.')
  _lib = lib
  return lib


def check(code: int, what: str):
  if code != 0:
    msg = load().mlt_strerror(code).decode()
    raise MltLibraryError(f'{what} failed with code {code}: {msg}')


def profile_enable(on: bool):
  load().mlt_profile_enable(1 if on else 0)


def profile_read(max_entries: int = 4096):
  """Returns [(name, ms, flops, bytes)] for every launch since the last read."""
  buf = (KernelTime * max_entries)()
  n = load().mlt_profile_read(buf, max_entries)
  return [(buf[i].name.decode(), buf[i].ms, buf[i].flops, buf[i].bytes) for i in range(n)]


def launch_count() -> int:
  return int(load().mlt_launch_count())
