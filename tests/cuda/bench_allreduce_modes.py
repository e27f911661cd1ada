"""PretrainingStep at full model size on N GPUs: overlapped bucketed all-reduce vs one pass after backward."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import mlt_b200  # noqa
from mlt_b200 import feature_utils as fu, mmt_encoder, tasks
L, B, MICRO, G, VOCAB = 4096, 4, 2, 256, 30522
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
if world > 1:
  opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=bool(int(os.environ.get('MLT_NCCL_HIPRI', '0'))))
  dist.init_process_group('nccl', device_id=dev, pg_options=opts)
res = {}
for mode in ('overlap', 'after'):
  torch.manual_seed(0)
  enc = mmt_encoder.MmtEncoder(vocab_size=VOCAB, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                               intermediate_size=3072, relative_vocab_size=32, relative_pos_max_distance=12,
                               hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, use_pre_activation_order=True,
                               patch_embedding_size=768, local_radius=64, num_global_tokens=G)
  model = tasks.MmtPretrainingModel(enc, 8192, [tasks.ClassificationHead(768, 2, 'itm', 0.0)]).to(dev).to(torch.bfloat16)
  opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
  step = tasks.PretrainingStep(model, opt, micro_batch_size=MICRO, overlap=(mode == 'overlap'))
  gen = torch.Generator().manual_seed(100 + rank)
  lengths = torch.randint(L // 2, L + 1, (B,), generator=gen)
  compact = fu.CompactSideInputs((torch.arange(L)[None] < lengths[:, None]).int().to(dev), torch.ones(B, G, dtype=torch.int32, device=dev),
                                 ((torch.arange(L) * G) // L)[None].expand(B, L).int().contiguous().to(dev), 12)
  inputs = {'word_ids': torch.randint(0, VOCAB, (B, L), generator=gen).to(dev),
            'patch_embeddings': torch.randn(B, 196, 768, generator=gen).to(dev).bfloat16(),
            'mlm_positions': torch.randint(198, L // 2, (B, 64), generator=gen).to(dev),
            'mpp_positions': torch.randint(2, 198, (B, 32), generator=gen).to(dev)}
  labels = {'mlm_label_ids': torch.randint(0, VOCAB, (B, 64), generator=gen).to(dev), 'mlm_label_weights': torch.ones(B, 64, device=dev),
            'mpp_label_ids': torch.randint(0, 8192, (B, 32), generator=gen).to(dev), 'mpp_label_weights': torch.ones(B, 32, device=dev),
            'itm_label_ids': torch.randint(0, 2, (B,), generator=gen).to(dev), 'itm_label_weights': torch.ones(B, device=dev)}
  for _ in range(2):
    step(inputs, labels, compact_side_inputs=compact)
  if world > 1: dist.barrier()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(4):
    step(inputs, labels, compact_side_inputs=compact)
  b.record(); torch.cuda.synchronize()
  res[mode] = a.elapsed_time(b) / 4
  chk = torch.tensor([sum(float(p.detach().float().sum()) for p in model.parameters())], device=dev, dtype=torch.float64)
  lo, hi = chk.clone(), chk.clone()
  if world > 1:
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
  res[mode + '_spread'] = float(hi - lo)
  del model, opt, step
if rank == 0:
  print(json.dumps({'n_gpus': world, **res}))
if world > 1:
  dist.destroy_process_group()
