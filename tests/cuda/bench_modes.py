"""Times fwd+bwd of the global-local operator with compact vs explicit side inputs, and the dense
operator (config-2 shape, S = 512) with explicit 2-D ids vs compact descriptors.  Development probe."""
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mlt_b200  # noqa: F401
from mlt_b200 import ops, synthetic
from mlt_b200.feature_utils import CompactSideInputs

NAMES = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias',
         'global_emb', 'global_bias')


def timeit(fn, reps=5):
  for _ in range(2):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
_, shape = synthetic.CONFIGS['c3_4096']
shape = dataclasses.replace(shape, batch=batch)
x = synthetic.make_inputs(shape, seed=1238, dtype=torch.bfloat16)
dev = [x[n].cuda().requires_grad_() for n in NAMES]
dlo, dgo = x['d_long_out'].cuda(), x['d_global_out'].cuda()
compact = CompactSideInputs(x['long_example_ids'].cuda(), x['global_example_ids'].cuda(),
                            x['sentence_ids'].cuda(), shape.max_distance)
explicit = ops.build_gl_side_inputs(compact, shape.local_radius)


def gl(side):
  def f():
    lo, go = ops.global_local_attention(*dev, local_radius=shape.local_radius, side=side)
    torch.autograd.backward([lo, go], [dlo, dgo])
  return f


def gl_recognised():
  side = ops.compact_from_explicit_gl(explicit, shape.local_radius)   # per call here; per batch in a stack
  assert side is not None
  gl(side)()


t_rec = timeit(lambda: ops.compact_from_explicit_gl(explicit, shape.local_radius))
print(f'global-local c3_4096 batch {batch}: compact {timeit(gl(compact)):.3f} ms, '
      f'explicit {timeit(gl(explicit)):.3f} ms, explicit recognised per call {timeit(gl_recognised):.3f} ms '
      f'(recognition alone {t_rec:.3f} ms, once per batch in a stack)')

# dense, config-2 shape: S = 512 (196 patches + text), 2-D ids for the patch block
B, S, H, D, R = 32, 512, 12, 64, 32
g = torch.Generator().manual_seed(7)
q, k, v = [torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).cuda().requires_grad_() for _ in range(3)]
emb = (0.02 * torch.randn(R, H, D, generator=g)).to(torch.bfloat16).cuda().requires_grad_()
bias = (0.02 * torch.randn(R, H, generator=g)).to(torch.bfloat16).cuda().requires_grad_()
do = torch.randn(B, S, H, D, generator=g).to(torch.bfloat16).cuda()
lengths = torch.randint(S // 2, S + 1, (B,), generator=g)
eid = (torch.arange(S)[None, :] < lengths[:, None]).to(torch.int32).cuda()
mask, ids = ops.build_dense_side_inputs(eid, 12)


def dense(**kw):
  def f():
    o = ops.dense_relative_attention(q, k, v, emb, bias, **kw)
    o.backward(do)
  return f


print(f'dense B {B} S {S}: explicit 1-D ids {timeit(dense(att_mask=mask, relative_att_ids=ids)):.3f} ms, '
      f'compact {timeit(dense(compact=ops.DenseCompactSideInputs(eid, max_distance=12))):.3f} ms')

# dense, 2-D image + text layout (14 x 14 patches), forward + backward: compact descriptors and the reference's
# explicit tensors (recognised per call here), relative vocabulary 32 and the shipped 49
for R2 in (32, 49):
  emb2 = (0.02 * torch.randn(R2, H, D, generator=g)).to(torch.bfloat16).cuda().requires_grad_()
  bias2 = (0.02 * torch.randn(R2, H, generator=g)).to(torch.bfloat16).cuda().requires_grad_()
  mask2, ids2 = ops.build_dense_side_inputs(eid, 12, num_patch_per_row=14, num_core_layers=2)
  c2d = ops.DenseCompactSideInputs(eid, max_distance=12, num_patch_per_row=14, num_core_layers=2)

  def dense2(**kw):
    def f():
      o = ops.dense_relative_attention(q, k, v, emb2, bias2, **kw)
      o.backward(do)
    return f

  def fwd2(**kw):
    def f():
      with torch.no_grad():
        ops.dense_relative_attention(q, k, v, emb2, bias2, **kw)
    return f

  print(f'dense 2-D ids, R {R2}: compact fwd+bwd {timeit(dense2(compact=c2d)):.3f} ms (fwd {timeit(fwd2(compact=c2d)):.3f}), '
        f'explicit fwd+bwd {timeit(dense2(att_mask=mask2, relative_att_ids=ids2)):.3f} ms, '
        f'1-D compact fwd+bwd {timeit(dense2(compact=ops.DenseCompactSideInputs(eid, max_distance=12))):.3f} ms')

# attention-probability dropout 0.1 (the reference's training default), dense S 512
for label, kw in (('1-D compact', dict(compact=ops.DenseCompactSideInputs(eid, max_distance=12))),
                  ('2-D compact', dict(compact=c2d))):
  def drop_step(p):
    def f():
      o = ops.dense_relative_attention(q, k, v, emb2, bias2, dropout_p=p, dropout_seed=5, **kw)
      o.backward(do)
    return f
  print(f'dense {label}, R {R2}: fwd+bwd without dropout {timeit(drop_step(0.0)):.3f} ms, with dropout 0.1 {timeit(drop_step(0.1)):.3f} ms')
