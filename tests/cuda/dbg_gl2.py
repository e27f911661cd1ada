"""gl2 kernels vs the general tcgen05 kernels and SIMT on a few shapes (debug / bring-up probe).
Phase 1: forward only.  Phase 2: forward + backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mlt_b200  # noqa
from mlt_b200 import synthetic, ops, _lib
from test_gpu_parity import run_cuda_gl, compact_of, NAMES

shapes = [(1, 256, 128, 1, 64, 32, 12), (1, 512, 32, 2, 64, 32, 12), (2, 200, 8, 2, 64, 32, 12), (1, 50, 4, 2, 64, 32, 12),
          (2, 300, 70, 2, 64, 30, 3), (1, 1100, 40, 2, 64, 32, 12), (2, 4096, 256, 12, 64, 32, 12)]
phase = sys.argv[1] if len(sys.argv) > 1 else 'both'


def inputs(dims):
  b, l, g, h, r, rv, dist = dims
  shape = synthetic.GlobalLocalShape(b, l, g, h, 64, r, rv, dist)
  x = synthetic.make_inputs(shape, seed=l + r, dtype=torch.bfloat16)
  for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
    x[n] = (x[n].float() * 10).bfloat16()
  return shape, x


if phase in ('fwd', 'both'):
  for dims in shapes:
    shape, x = inputs(dims)
    side = compact_of(x, shape)
    dev_in = [x[n].cuda() for n in NAMES]
    print('fwd start', dims, flush=True)
    outs = {}
    for impl in ('tc', 'tc_generic', 'simt'):
      with torch.no_grad():
        lo, go = ops.global_local_attention(*dev_in, local_radius=shape.local_radius, side=side, impl=impl)
      torch.cuda.synchronize()
      outs[impl] = lo.float()
      outs[impl + '_g'] = go.float()
    print('  long_out gl2 vs generic %.4g, vs simt %.4g, finite %s' % (
        (outs['tc'] - outs['tc_generic']).abs().max().item(), (outs['tc'] - outs['simt']).abs().max().item(),
        bool(torch.isfinite(outs['tc']).all())), flush=True)
    print('  global_out gl2 vs generic %.4g, vs simt %.4g, finite %s' % (
        (outs['tc_g'] - outs['tc_generic_g']).abs().max().item(), (outs['tc_g'] - outs['simt_g']).abs().max().item(),
        bool(torch.isfinite(outs['tc_g']).all())), flush=True)
if phase in ('bwd', 'both'):
  for dims in shapes:
    shape, x = inputs(dims)
    side = compact_of(x, shape)
    print('bwd start', dims, flush=True)
    a = run_cuda_gl(x, shape, side, impl='tc')
    s = run_cuda_gl(x, shape, side, impl='simt')
    errs = {n: (u.float() - v.float()).abs().max().item() / max(1.0, v.float().abs().max().item())
            for n, u, v in zip(NAMES, a[2], s[2])}
    print('  worst grad (scaled) %.4g: %s' % (max(errs.values()), {k: round(v, 4) for k, v in errs.items()}), flush=True)
