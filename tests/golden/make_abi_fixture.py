"""Regenerates tests/golden/gl_abi_fixture.bin: inputs and fp64-oracle outputs / gradients of one small
global-local attention problem, as raw little-endian arrays for the torch-free C-ABI test
(tests/cuda/abi_smoke.cu).  Layout: int32 header [B, L, G, H, d, r, R, D], then in order
  float32  long_q long_k long_v [B,L,H,d]   global_q global_k global_v [B,G,H,d]
  float32  long_emb [R,H,d] long_bias [R,H] global_emb global_bias
  float32  d_long_out [B,L,H,d]  d_global_out [B,G,H,d]
  int32    long_example_ids [B,L]  global_example_ids [B,G]  sentence_ids [B,L]
  float32  (oracle) long_out global_out d_long_q d_long_k d_long_v d_global_q d_global_k d_global_v
           d_long_emb d_long_bias d_global_emb d_global_bias
Run from the repo root: python tests/golden/make_abi_fixture.py"""
import os
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import mlt_b200  # noqa: F401,E402
from mlt_b200 import synthetic  # noqa: E402
from oracle import attention_oracle as ao  # noqa: E402
from oracle import feature_oracle as fo  # noqa: E402

NAMES = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias',
         'global_emb', 'global_bias')


def main():
  shape = synthetic.GlobalLocalShape(2, 96, 8, 2, 64, 16, 16, 3)
  x = synthetic.make_inputs(shape, seed=424242)
  for n in NAMES[6:]:
    x[n] = x[n] * 10
  side = {k: torch.tensor(v) for k, v in fo.make_global_local_side_inputs(
      x['long_example_ids'].numpy(), x['global_example_ids'].numpy(), x['sentence_ids'].numpy(),
      shape.local_radius, shape.max_distance).items()}
  ref = [x[n].double().requires_grad_() for n in NAMES]
  lo, go = ao.fused_global_local_attention(*ref[:6], side, (ref[6], ref[7]), (ref[8], ref[9]), shape.local_radius)
  ((lo * x['d_long_out'].double()).sum() + (go * x['d_global_out'].double()).sum()).backward()
  out = pathlib.Path(__file__).with_name('gl_abi_fixture.bin')
  with open(out, 'wb') as f:
    np.asarray([shape.batch, shape.long_len, shape.global_len, shape.heads, shape.head_dim, shape.local_radius,
                shape.relative_vocab_size, shape.max_distance], dtype='<i4').tofile(f)
    for n in NAMES + ('d_long_out', 'd_global_out'):
      x[n].numpy().astype('<f4').tofile(f)
    for n in ('long_example_ids', 'global_example_ids', 'sentence_ids'):
      x[n].numpy().astype('<i4').tofile(f)
    for t in [lo, go] + [r.grad for r in ref]:
      t.detach().numpy().astype('<f4').tofile(f)
  print(f'wrote {out} ({os.path.getsize(out)} bytes)')


if __name__ == '__main__':
  main()
