"""Timeline of gl2_fwd_long_kernel (library built with make EXTRA=-DMLT_TC_TRACE): clock64 stamps of CTA 3,
softmax warp 0 (SM) and the MMA thread."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mlt_b200  # noqa
from mlt_b200 import synthetic, ops, _lib
from mlt_b200.feature_utils import CompactSideInputs
seed_off, shape = synthetic.CONFIGS['c3_4096']
x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16, device='cuda')
NAMES = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias', 'global_emb', 'global_bias')
compact = CompactSideInputs(x['long_example_ids'], x['global_example_ids'], x['sentence_ids'], shape.max_distance)
with torch.no_grad():
  for _ in range(2):
    ops.global_local_attention(*[x[n] for n in NAMES], local_radius=shape.local_radius, side=compact, impl='auto')
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * (2 * 2048))()
n = (C.c_int * 2)()
lib.mlt_debug_read_trace_gl2f(buf, n)
ev = []
for role in range(2):
  for k in range(n[role]):
    ev.append((buf[role * 2048 + 2 * k], role, buf[role * 2048 + 2 * k + 1]))
ev.sort()
t0 = ev[0][0]
last = {}
lim = int(sys.argv[1]) if len(sys.argv) > 1 else 200
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for t, role, code in ev[skip:skip + lim]:
  print('%8d  %s  %4d   (+%d)' % (t - t0, ['SM ', 'MMA'][role], code, t - last.get(role, t)))
  last[role] = t
