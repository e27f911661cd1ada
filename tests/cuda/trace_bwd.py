import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import mlt_b200
from mlt_b200 import ops, synthetic, _lib
from mlt_b200.feature_utils import CompactSideInputs
import dataclasses
_, shape = synthetic.CONFIGS['c3_4096']
shape = dataclasses.replace(shape, batch=4)
x = synthetic.make_inputs(shape, seed=1238, dtype=torch.bfloat16)
names = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias', 'global_emb', 'global_bias')
dev = [x[n].cuda().requires_grad_() for n in names]
c = CompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(), x['sentence_ids'].cuda(), shape.max_distance)
for _ in range(2):
  lo, go = ops.global_local_attention(*dev, local_radius=shape.local_radius, side=c)
  torch.autograd.backward([lo, go], [x['d_long_out'].cuda(), x['d_global_out'].cuda()])
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * (6 * 256))()
print('rc', lib.mlt_debug_read_trace_b(buf))
t = np.array(buf[:], dtype=np.int64).reshape(6, 256)
t0 = t[t > 0].min()
for role, name in enumerate(['kv_elementwise', 'bq_elementwise', 'kv_mma', 'bq_mma', 'bq_producer', 'kv_producer']):
  vals = [(i, int(v - t0)) for i, v in enumerate(t[role]) if v > 0]
  print(name, vals)
