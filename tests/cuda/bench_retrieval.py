"""Image-text retrieval scoring throughput (BASELINE.json configs[4]): batched pair scoring
`softmax(itm_logits)[:, 1]` with the full-size dense encoder (12 layers, d 768, S = 512 = 2 + 196 patches
+ text, bf16), pairs sharded over the ranks (no collective on the data path), compact 2-D side inputs.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tests/cuda/bench_retrieval.py [pairs per GPU and batch] [batches]

Rank 0 prints one JSON line: pairs/s over all ranks (max-over-ranks device time)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
import mlt_b200  # noqa: F401
from mlt_b200 import mmt_encoder, ops, tasks

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S, NPR, VOCAB = 512, 14, 30522
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)
torch.manual_seed(0)
enc = mmt_encoder.MmtEncoder(vocab_size=VOCAB, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                             intermediate_size=3072, relative_vocab_size=32, relative_pos_max_distance=12,
                             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, patch_embedding_size=768)
model = tasks.MmtPretrainingModel(enc, mpp_output_num_classes=8192,
                                  classification_heads=[tasks.ClassificationHead(768, 2, 'itm', 0.0)])
model = model.to(dev).to(torch.bfloat16).eval()
gen = torch.Generator().manual_seed(7 + rank)
batches = []
for _ in range(NB):
  lengths = torch.randint(S // 2, S + 1, (B,), generator=gen)
  eid = (torch.arange(S)[None] < lengths[:, None]).int().to(dev)
  batches.append({'word_ids': torch.randint(0, VOCAB, (B, S), generator=gen).to(dev),
                  'patch_embeddings': torch.randn(B, NPR * NPR, 768, generator=gen).to(dev).to(torch.bfloat16),
                  'compact': ops.DenseCompactSideInputs(eid, max_distance=12, num_patch_per_row=NPR,
                                                        num_core_layers=2)})


def barrier():
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()


tasks.retrieval_scores(model, batches[:1])
barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
scores = tasks.retrieval_scores(model, batches)
b.record()
barrier()
t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
if world > 1:
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
  ms = t.item()
  print(json.dumps({'probe': 'retrieval_scoring', 'n_gpus': world, 'pairs_per_gpu': B * NB, 'seq_len': S,
                    'ms': ms, 'pairs_per_s': world * B * NB / (ms * 1e-3),
                    'tokens_per_s': world * B * NB * S / (ms * 1e-3),
                    'scores_in_unit_interval': bool(((scores >= 0) & (scores <= 1)).all())}))
if world > 1:
  dist.destroy_process_group()
