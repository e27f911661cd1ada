"""Per-kernel times of one fwd+bwd step at the headline workload (library profiler hook), quick probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mlt_b200  # noqa
from mlt_b200 import synthetic, ops, _lib
from mlt_b200.feature_utils import CompactSideInputs
name = sys.argv[1] if len(sys.argv) > 1 else 'c3_4096'
impl = sys.argv[2] if len(sys.argv) > 2 else 'auto'
seed_off, shape = synthetic.CONFIGS[name]
x = synthetic.make_inputs(shape, seed=1234 + seed_off, dtype=torch.bfloat16, device='cuda')
NAMES = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias', 'global_emb', 'global_bias')
compact = CompactSideInputs(x['long_example_ids'], x['global_example_ids'], x['sentence_ids'], shape.max_distance)
leaves = [x[n].requires_grad_() for n in NAMES]
def step():
  for t in leaves: t.grad = None
  lo, go = ops.global_local_attention(*leaves, local_radius=shape.local_radius, side=compact, impl=impl)
  torch.autograd.backward([lo, go], [x['d_long_out'], x['d_global_out']])
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): step()
b.record(); torch.cuda.synchronize()
print('step ms', a.elapsed_time(b) / 20)
def fwd():
  with torch.no_grad():
    ops.global_local_attention(*leaves, local_radius=shape.local_radius, side=compact, impl=impl)
for _ in range(3): fwd()
torch.cuda.synchronize()
a.record()
for _ in range(20): fwd()
b.record(); torch.cuda.synchronize()
print('forward only ms', a.elapsed_time(b) / 20)
_lib.profile_enable(True)
for _ in range(3): step()
torch.cuda.synchronize()
recs = _lib.profile_read(); _lib.profile_enable(False)
agg = {}
for n, ms, fl, by in recs:
  agg.setdefault(n, []).append(ms)
for n, v in agg.items(): print('%-28s %.4f ms' % (n, sum(v) / len(v)))
