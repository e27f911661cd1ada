import sys; sys.path.insert(0,'/root/repo')
import torch, mlt_b200
from mlt_b200 import ops, synthetic
from mlt_b200.feature_utils import CompactSideInputs
shape = synthetic.GlobalLocalShape(2, 320, 16, 4, 64, 64, 32, 12)
x = synthetic.make_inputs(shape, seed=77, dtype=torch.bfloat16)
names = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias', 'global_emb', 'global_bias')
dev = [x[n].cuda() for n in names]
c = CompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(), x['sentence_ids'].cuda(), 12)
lo, go = ops.global_local_attention(*dev, local_radius=64, side=c, impl='tc')
torch.cuda.synchronize()
print('ok', lo.float().abs().mean().item())
