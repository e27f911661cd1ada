"""End-to-end pretraining step at full model size (BASELINE.json configs[3]): 12-layer, d = 768 encoder in
long-input mode (L long tokens + L/16 global tokens, radius 64) with the MLM / masked-patch / image-text
matching heads, random-init weights, bf16 parameters, AdamW, micro-batch accumulation and ONE flat NCCL
gradient all-reduce per optimizer step.  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P tests/cuda/bench_pretrain_step.py [L] [batch per GPU] [micro batch]

Prints one JSON line on rank 0: tokens/s over all ranks (max-over-ranks device time), the share of a step
spent inside the attention kernels, and a parameter checksum spread that must be 0 (replicas identical)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
import mlt_b200  # noqa: F401
from mlt_b200 import _lib, feature_utils as fu, mmt_encoder, tasks

L = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
MICRO = int(sys.argv[3]) if len(sys.argv) > 3 else 2
G, NPATCH, VOCAB = L // 16, 196, 30522
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
  dist.init_process_group('nccl', device_id=dev)

torch.manual_seed(0)            # identical initial weights on every rank
enc = mmt_encoder.MmtEncoder(vocab_size=VOCAB, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                             intermediate_size=3072, relative_vocab_size=32, relative_pos_max_distance=12,
                             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                             use_pre_activation_order=True, patch_embedding_size=768, local_radius=64,
                             num_global_tokens=G)
model = tasks.MmtPretrainingModel(enc, mpp_output_num_classes=8192,
                                  classification_heads=[tasks.ClassificationHead(768, 2, 'itm', 0.0)])
model = model.to(dev).to(torch.bfloat16)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01)
step = tasks.PretrainingStep(model, opt, micro_batch_size=MICRO)
nparams = sum(p.numel() for p in model.parameters())

gen = torch.Generator().manual_seed(100 + rank)     # different data per rank
lengths = torch.randint(L // 2, L + 1, (B,), generator=gen)
le = (torch.arange(L)[None] < lengths[:, None]).int()
compact = fu.CompactSideInputs(le.to(dev), torch.ones(B, G, dtype=torch.int32, device=dev),
                               ((torch.arange(L) * G) // L)[None].expand(B, L).int().contiguous().to(dev), 12)
inputs = {'word_ids': torch.randint(0, VOCAB, (B, L), generator=gen).to(dev),
          'patch_embeddings': torch.randn(B, NPATCH, 768, generator=gen).to(dev).to(torch.bfloat16),
          'mlm_positions': torch.randint(NPATCH + 2, L // 2, (B, 64), generator=gen).to(dev),
          'mpp_positions': torch.randint(2, NPATCH + 2, (B, 32), generator=gen).to(dev)}
labels = {'mlm_label_ids': torch.randint(0, VOCAB, (B, 64), generator=gen).to(dev),
          'mlm_label_weights': torch.ones(B, 64, device=dev),
          'mpp_label_ids': torch.randint(0, 8192, (B, 32), generator=gen).to(dev),
          'mpp_label_weights': torch.ones(B, 32, device=dev),
          'itm_label_ids': torch.randint(0, 2, (B,), generator=gen).to(dev),
          'itm_label_weights': torch.ones(B, device=dev)}


def barrier():
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()


for _ in range(2):
  loss = step(inputs, labels, compact_side_inputs=compact)
barrier()
steps = 4
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
  loss = step(inputs, labels, compact_side_inputs=compact)
b.record()
barrier()
t = torch.tensor([a.elapsed_time(b) / steps], device=dev, dtype=torch.float64)
if world > 1:
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = t.item()

# share of the step inside the attention kernels (library profiler serialises them: upper bound)
_lib.profile_enable(True)
step(inputs, labels, compact_side_inputs=compact)
torch.cuda.synchronize()
att_ms = sum(r[1] for r in _lib.profile_read(1 << 16))
_lib.profile_enable(False)

chk = torch.tensor([sum(float(p.detach().float().sum()) for p in model.parameters())], device=dev, dtype=torch.float64)
lo, hi = chk.clone(), chk.clone()
if world > 1:
  dist.all_reduce(lo, op=dist.ReduceOp.MIN)
  dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
  print(json.dumps({'probe': 'pretraining_step', 'n_gpus': world, 'long_len': L, 'global_len': G,
                    'batch_per_gpu': B, 'micro_batch': MICRO, 'params': nparams, 'ms_per_step': ms,
                    'tokens_per_s': world * B * L / (ms * 1e-3), 'attention_kernels_ms': att_ms,
                    'loss': float(loss), 'param_checksum_spread': float(hi - lo)}))
if world > 1:
  dist.destroy_process_group()
