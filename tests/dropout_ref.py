"""Restatement of the kernels' dropout keep function (csrc/mlt_common.cuh: ``mix32``,
``dropout_salt``, ``dropout_row_base``, ``dropout_keep``) in numpy uint32 arithmetic.

TEST INFRASTRUCTURE: the tests build the mask here and hand it to the fp64 oracle, so dropout
results are compared against an independent evaluation with the *identical* mask.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def mix32(x):
  x = np.asarray(x, dtype=np.uint64) & _M
  x ^= x >> np.uint64(16)
  x = (x * np.uint64(0x7feb352d)) & _M
  x ^= x >> np.uint64(15)
  x = (x * np.uint64(0x846ca68b)) & _M
  x ^= x >> np.uint64(16)
  return x


def mix_elem(x):
  x = np.asarray(x, dtype=np.uint64) & _M
  x = (x * np.uint64(0x9e3779b1)) & _M
  x ^= x >> np.uint64(15)
  x = (x * np.uint64(0x85ebca77)) & _M
  return x


def keep_mask(seed, p, batch, heads, rows, cols, rowset):
  """bool [B, rows, cols, H]: keep(b, h, i, col) for one row set (0 = dense / long rows, 1 = global
  rows); ``cols`` indexes the row's concatenated key axis (segment 0 first)."""
  thr = np.uint64(min(int(p * 4294967296.0), 4294967295))
  seed_lo, seed_hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
  bh = (np.arange(batch, dtype=np.uint64)[:, None] * np.uint64(heads) + np.arange(heads, dtype=np.uint64)[None, :])
  inner = mix32((seed_hi + np.uint64(0x9e3779b9) * (np.uint64(2) * bh + np.uint64(rowset + 1))) & _M)
  salt = mix32(seed_lo ^ inner)                                                  # [B, H]
  i = np.arange(rows, dtype=np.uint64)[None, :, None, None]
  c = np.arange(cols, dtype=np.uint64)[None, None, :, None]
  x = (salt[:, None, None, :] + i * np.uint64(0x00010001) + c) & _M
  return mix_elem(x) >= thr
