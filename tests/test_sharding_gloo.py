"""N > 1 path on the CPU: world_size 2 over gloo.  The attention path shards batch x head units
with no data-path collective; ranks only exchange timings.  This test exercises the
host logic bench.py uses: unit partition, barrier, max-over-ranks reduction, value aggregation."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["MLT_ROOT"])
import torch, torch.distributed as dist
import mlt_b200
from mlt_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
units = sharding.partition_units(batch=5, heads=3, rank=rank, world=world)
ms = sharding.max_over_ranks(10.0 + rank)          # slowest rank defines the step time
total = sharding.sum_over_ranks(float(len(units)))
dist.barrier()
# one write per rank: the two processes share the launcher's pipe and print() may split line and newline
sys.stdout.write(json.dumps({"rank": rank, "n": len(units), "first": units[0], "ms": ms, "total": total}) + "\n")
sys.stdout.flush()
dist.destroy_process_group()
'''


def test_two_rank_partition_and_timing_reduction(tmp_path):
  script = tmp_path / 'worker.py'
  script.write_text(WORKER)
  env = dict(os.environ, MLT_ROOT=ROOT)
  for attempt in range(3):      # the probed port can be taken between the probe and the rendezvous: retried
    with socket.socket() as s:
      s.bind(('127.0.0.1', 0))
      port = s.getsockname()[1]
    out = subprocess.run(
        [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
         '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
        capture_output=True, text=True, env=env, timeout=240)
    if out.returncode == 0:
      break
  assert out.returncode == 0, out.stderr[-2000:]
  import json
  import re
  rows = [json.loads(m) for m in re.findall(r'\{[^{}]*\}', out.stdout)]   # tolerant of interleaved output
  assert len(rows) == 2
  assert sorted(r['n'] for r in rows) == [7, 8]          # 15 (b, h) units over 2 ranks
  assert all(r['ms'] == 11.0 for r in rows)              # max over ranks
  assert all(r['total'] == 15.0 for r in rows)


def test_partition_is_a_disjoint_cover():
  sys.path.insert(0, ROOT)
  import mlt_b200  # noqa: F401
  from mlt_b200 import sharding
  for world in (1, 2, 4, 8):
    seen = []
    for rank in range(world):
      seen += sharding.partition_units(batch=3, heads=12, rank=rank, world=world)
    assert sorted(seen) == [(b, h) for b in range(3) for h in range(12)]


def test_shard_views_follow_the_unit_partition():
  """CPU: the strided views of shard_inputs own exactly the (b, h) units partition_units assigns."""
  import torch
  sys.path.insert(0, ROOT)
  import mlt_b200  # noqa: F401
  from mlt_b200 import sharding
  b, h = 3, 12
  tag = (torch.arange(b)[:, None, None, None] * 100 + torch.arange(h)[None, None, :, None]).expand(b, 5, h, 4).float()
  for world in (2, 3, 4):
    axis = sharding.shard_axis(b, h, world)
    assert axis == 'heads'
    for rank in range(world):
      view = sharding.shard_inputs({'long_q': tag}, rank, world, axis)['long_q']
      assert view.data_ptr() == tag[:, :, rank:].data_ptr()          # a view, not a copy
      got = sorted({(int(v) // 100, int(v) % 100) for v in view[:, 0, :, 0].reshape(-1)})
      assert got == sorted(sharding.partition_units(b, h, rank, world))
  assert sharding.shard_axis(8, 12, 8) == 'batch'                     # 8 GPUs, 12 heads: batch elements instead


import pytest  # noqa: E402


@pytest.mark.gpu
def test_sharded_runs_are_bit_identical_to_the_single_run():
  """shard one batch by (b, h) units, run the operator once per rank's views (here
  the ranks run one after the other on the one GPU the test box has; the kernels see exactly what they would
  see on separate GPUs), stitch, and compare with the unsharded run: torch.equal on outputs and q/k/v
  gradients, for a head split and a batch split."""
  import torch
  sys.path.insert(0, ROOT)
  import mlt_b200  # noqa: F401
  from mlt_b200 import feature_utils as fu, ops, sharding, synthetic
  shape = synthetic.GlobalLocalShape(4, 640, 40, 4, 64, 64, 32, 12)
  x = synthetic.make_inputs(shape, seed=17, dtype=torch.bfloat16, device='cuda')
  names = ('long_q', 'long_k', 'long_v', 'global_q', 'global_k', 'global_v', 'long_emb', 'long_bias',
           'global_emb', 'global_bias')

  def run(inp):
    leaves = [inp[n].detach().requires_grad_() for n in names]
    side = fu.CompactSideInputs(inp['long_example_ids'], inp['global_example_ids'], inp['sentence_ids'], 12)
    lo, go = ops.global_local_attention(*leaves, local_radius=64, side=side, impl='tc')
    torch.autograd.backward([lo, go], [inp['d_long_out'], inp['d_global_out']])
    return [lo, go] + [t.grad for t in leaves[:6]]

  full = run(x)
  for world in (2, 4):
    for axis in ('heads', 'batch'):
      parts = [run(sharding.shard_inputs(x, r, world, axis)) for r in range(world)]
      for i, ref in enumerate(full):
        got = sharding.unshard_outputs([p[i] for p in parts], world, axis, ref)
        assert torch.equal(got, ref), (world, axis, i)


@pytest.mark.gpu
def test_two_gpu_shards_match_single_gpu():
  """Same check with two real ranks over NCCL when the box has two GPUs (skipped on a one-GPU box)."""
  import torch
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs')
  worker = r"""
import os, sys
sys.path.insert(0, os.environ['MLT_ROOT'])
import torch, torch.distributed as dist
import mlt_b200
from mlt_b200 import feature_utils as fu, ops, sharding, synthetic
rank = int(os.environ['RANK']); torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
shape = synthetic.GlobalLocalShape(2, 640, 40, 4, 64, 64, 32, 12)
x = synthetic.make_inputs(shape, seed=17, dtype=torch.bfloat16, device='cuda')
def run(inp):
  side = fu.CompactSideInputs(inp['long_example_ids'], inp['global_example_ids'], inp['sentence_ids'], 12)
  return ops.global_local_attention(*[inp[n] for n in ('long_q','long_k','long_v','global_q','global_k','global_v','long_emb','long_bias','global_emb','global_bias')], local_radius=64, side=side, impl='tc')
full = run(x)[0]
mine = run(sharding.shard_inputs(x, rank, 2, 'heads'))[0].contiguous()
parts = [torch.empty_like(mine) for _ in range(2)]
dist.all_gather(parts, mine)          # test-only gather; the data path itself has no collective
ok = torch.equal(sharding.unshard_outputs(parts, 2, 'heads', full), full)
print('SHARD_OK' if ok else 'SHARD_MISMATCH', flush=True)
dist.destroy_process_group()
"""
  import tempfile
  with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, 'w.py')
    open(path, 'w').write(worker)
    with socket.socket() as s:
      s.bind(('127.0.0.1', 0))
      port = s.getsockname()[1]
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), path],
                         capture_output=True, text=True, env=dict(os.environ, MLT_ROOT=ROOT), timeout=600)
  assert out.returncode == 0 and out.stdout.count('SHARD_OK') == 2, out.stdout[-1500:] + out.stderr[-1500:]
