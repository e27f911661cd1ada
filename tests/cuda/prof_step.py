"""Small driver for ncu: a few fwd+bwd steps of the bench workload (batch-reduced)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mlt_b200
from mlt_b200 import ops, synthetic
from mlt_b200.feature_utils import CompactSideInputs
import dataclasses
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
_, shape = synthetic.CONFIGS['c3_4096']
shape = dataclasses.replace(shape, batch=batch)
x = synthetic.make_inputs(shape, seed=1238, dtype=torch.bfloat16)
names = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias', 'global_emb', 'global_bias')
dev = [x[n].cuda().requires_grad_() for n in names]
c = CompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(), x['sentence_ids'].cuda(), shape.max_distance)
for _ in range(steps):
  lo, go = ops.global_local_attention(*dev, local_radius=shape.local_radius, side=c)
  torch.autograd.backward([lo, go], [x['d_long_out'].cuda(), x['d_global_out'].cuda()])
torch.cuda.synchronize()
print('done')
