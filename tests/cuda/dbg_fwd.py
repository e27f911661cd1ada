"""Debug helper (run on a B200 through gpurun): forward of the tcgen05 path against the SIMT
path, explicit and compact side inputs, with a per-(batch, head, 32-row block) error map."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import mlt_b200  # noqa: F401
from mlt_b200 import ops, synthetic
from mlt_b200.feature_utils import CompactSideInputs
from oracle import feature_oracle as fo

dims = [int(v) for v in sys.argv[1:9]] if len(sys.argv) >= 9 else [2, 320, 16, 4, 64, 64, 32, 12]
seed = int(sys.argv[9]) if len(sys.argv) > 9 else 77
shape = synthetic.GlobalLocalShape(*dims)
x = synthetic.make_inputs(shape, seed=seed, dtype=torch.bfloat16)
for n in ('long_emb', 'long_bias', 'global_emb', 'global_bias'):
  x[n] = (x[n].float() * 10).bfloat16()
names = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias',
         'global_emb', 'global_bias')
dev = [x[n].cuda() for n in names]
compact = CompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(),
                            x['sentence_ids'].cuda(), shape.max_distance)
explicit = {k: torch.tensor(v).cuda() for k, v in fo.make_global_local_side_inputs(
    x['long_example_ids'].numpy(), x['global_example_ids'].numpy(), x['sentence_ids'].numpy(),
    shape.local_radius, shape.max_distance).items()}
for name, side in (('compact', compact), ('explicit', explicit)):
  with torch.no_grad():
    lo_t, go_t = ops.global_local_attention(*dev, local_radius=shape.local_radius, side=side, impl='tc')
    lo_s, go_s = ops.global_local_attention(*dev, local_radius=shape.local_radius, side=side, impl='simt')
  torch.cuda.synchronize()
  for tag, a, r in (('long', lo_t, lo_s), ('global', go_t, go_s)):
    err = (a.float() - r.float()).abs().amax(dim=-1)   # [B, len, H]
    print(name, tag, 'max err', err.max().item())
    if err.max().item() > 2e-2:
      B, Ln, H = err.shape
      for b in range(B):
        for h in range(H):
          blocks = [err[b, s:s + 32, h].max().item() for s in range(0, Ln, 32)]
          print('  b', b, 'h', h, ' '.join('%.2f' % v for v in blocks))
      if tag == 'long':
        e = err[0, :, 0]
        print('rows b0 h0:', ' '.join('%d:%.2f' % (i, v) for i, v in enumerate(e.tolist()) if v > 0.02))
        print('lengths', x['long_example_ids'].sum(dim=1).tolist(), x['global_example_ids'].sum(dim=1).tolist())
