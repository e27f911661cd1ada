"""gl2 kernels vs the general tcgen05 kernels and SIMT on a few shapes (debug / bring-up probe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mlt_b200  # noqa
from mlt_b200 import synthetic, ops, _lib
from test_gpu_parity import run_cuda_gl, compact_of, NAMES

shapes = [(1, 256, 128, 1, 64, 32, 12), (1, 512, 32, 2, 64, 32, 12), (2, 200, 8, 2, 64, 32, 12), (1, 50, 4, 2, 64, 32, 12),
          (2, 300, 70, 2, 20, 20, 3), (1, 1100, 40, 2, 64, 32, 12), (2, 4096, 256, 12, 64, 32, 12)]
if len(sys.argv) > 1:
  shapes = shapes[:int(sys.argv[1])]
for dims in shapes:
  b, l, g, h, r, rv, dist = dims
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=l + r, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  side = compact_of(x, shape)
  a = run_cuda_gl(x, shape, side, impl='tc')
  c = run_cuda_gl(x, shape, side, impl='tc_generic')
  s = run_cuda_gl(x, shape, side, impl='simt')
  e1 = (a[0].float() - c[0].float()).abs().max().item()
  e2 = (a[0].float() - s[0].float()).abs().max().item()
  eg = max((u.float() - v.float()).abs().max().item() / max(1.0, v.float().abs().max().item()) for u, v in zip(a[2], s[2]))
  print(dims, 'long_out gl2 vs generic %.4g, vs simt %.4g; worst grad (tc path vs simt, scaled) %.4g; finite %s' %
        (e1, e2, eg, bool(torch.isfinite(a[0].float()).all())), flush=True)
