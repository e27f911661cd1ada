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
  for attempt in range(2):      # the probed port can be taken between the probe and the rendezvous: one retry
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
