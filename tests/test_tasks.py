"""Task-side mirrors: losses / heads on the CPU, gradient all-reduce over gloo (world size 2),
e2e pretraining step and retrieval scoring on the GPU."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

import mlt_b200  # noqa: F401
from mlt_b200 import feature_utils as fu
from mlt_b200 import mmt_encoder, tasks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_weighted_loss_matches_reference_formula():
  torch.manual_seed(0)
  logits, labels = torch.randn(3, 5, 7), torch.randint(0, 7, (3, 5))
  w = torch.tensor([[1., 1, 0, 0, 0], [1, 0, 0, 0, 0], [0, 0, 0, 0, 0]])
  pos = torch.rand(3, 5)
  got = tasks.weighted_sparse_categorical_crossentropy_loss(logits, labels, w, pos)
  ce = torch.nn.functional.cross_entropy(logits.reshape(-1, 7), labels.reshape(-1), reduction='none').reshape(3, 5)
  assert torch.allclose(got, (w * pos * ce).sum() / w.sum())
  # divide_no_nan: no active label -> 0 (reference :41)
  assert tasks.weighted_sparse_categorical_crossentropy_loss(logits, labels, torch.zeros(3, 5)).item() == 0.0


def test_pretraining_losses_mask_negative_pairs():
  torch.manual_seed(1)
  out = {'mlm_logits': torch.randn(2, 3, 11), 'mpp_logits': torch.randn(2, 4, 8), 'itm_logits': torch.randn(2, 2)}
  labels = {'mlm_label_ids': torch.randint(0, 11, (2, 3)), 'mlm_label_weights': torch.ones(2, 3),
            'mpp_label_ids': torch.randint(0, 8, (2, 4)), 'mpp_label_weights': torch.ones(2, 4),
            'itm_label_ids': torch.tensor([1, 0]), 'itm_label_weights': torch.ones(2)}
  total = tasks.pretraining_losses(labels, out)
  # the negative pair (row 1) must not contribute to MLM / MPP (reference pretraining.py:101-108)
  pos_only = {k: v[:1] for k, v in labels.items() if not k.startswith('itm')}
  mlm = tasks.weighted_sparse_categorical_crossentropy_loss(out['mlm_logits'][:1], pos_only['mlm_label_ids'], pos_only['mlm_label_weights'])
  mpp = tasks.weighted_sparse_categorical_crossentropy_loss(out['mpp_logits'][:1], pos_only['mpp_label_ids'], pos_only['mpp_label_weights'])
  itm = tasks.weighted_sparse_categorical_crossentropy_loss(out['itm_logits'], labels['itm_label_ids'], labels['itm_label_weights'])
  assert torch.allclose(total, mlm + mpp + itm, atol=1e-6)


def test_gather_indexes_and_recall():
  seq = torch.arange(2 * 4 * 3, dtype=torch.float32).reshape(2, 4, 3)
  got = tasks.gather_indexes(seq, torch.tensor([[0, 3], [2, 2]]))
  assert torch.equal(got, torch.stack([seq[0, 0], seq[0, 3], seq[1, 2], seq[1, 2]]))
  # hand-checked retrieval case: 2 texts x 3 images, text 0 <-> image 2, text 1 <-> image 0
  text, image = tasks.enumerate_image_text_pairs(num_images=3, num_texts=2)
  scores = torch.tensor([0.9, 0.1, 0.5, 0.2, 0.8, 0.7], dtype=torch.float64)
  r = tasks.get_recall_at_k(image, text, torch.tensor([2, 0])[text], scores, topks=(1, 2, 3))
  # text -> image: text 0 ranks its image 2nd, text 1 ranks its image 3rd
  assert [r['t2i @  1'], r['t2i @  2'], r['t2i @  3']] == [0.0, 0.5, 1.0]
  # image -> text: image 0 (gt of text 1) ranks it 2nd of 2; image 2 (gt of text 0) ranks it 2nd of 2
  assert [r['i2t @  1'], r['i2t @  2']] == [0.0, 1.0]


WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["MLT_ROOT"])
import torch, torch.distributed as dist
import mlt_b200
from mlt_b200 import tasks
dist.init_process_group("gloo")
rank = dist.get_rank()
torch.manual_seed(0)
class Tiny(torch.nn.Module):
  def __init__(self):
    super().__init__()
    self.emb = torch.nn.Embedding(16, 8); self.mlm = torch.nn.Linear(8, 16); self.mpp = torch.nn.Linear(8, 4); self.itm = torch.nn.Linear(8, 2)
  def forward(self, word_ids, mlm_positions, mpp_positions, training=None):
    x = self.emb(word_ids)
    g = lambda pos: torch.gather(x, 1, pos[..., None].expand(-1, -1, 8))
    return {'mlm_logits': self.mlm(g(mlm_positions)), 'mpp_logits': self.mpp(g(mpp_positions)), 'itm_logits': self.itm(x[:, 0])}
model = Tiny()
init = {k: v.clone() for k, v in model.state_dict().items()}
opt = torch.optim.SGD(model.parameters(), lr=0.1)
step = tasks.PretrainingStep(model, opt, micro_batch_size=2, bucket_mb=0.0001)   # several buckets even for this toy
g = torch.Generator().manual_seed(100 + rank)      # different data per rank
inputs = {'word_ids': torch.randint(0, 16, (4, 6), generator=g), 'mlm_positions': torch.randint(0, 6, (4, 2), generator=g),
          'mpp_positions': torch.randint(0, 6, (4, 2), generator=g)}
labels = {'mlm_label_ids': torch.randint(0, 16, (4, 2), generator=g), 'mlm_label_weights': torch.ones(4, 2),
          'mpp_label_ids': torch.randint(0, 4, (4, 2), generator=g), 'mpp_label_weights': torch.ones(4, 2),
          'itm_label_ids': torch.randint(0, 2, (4,), generator=g), 'itm_label_weights': torch.ones(4)}
nb = sum(len(gr['buckets']) for gr in step.groups)
# reference semantics (scale_loss=False): the applied gradient is the SUM over the replicas of each replica's
# mean-over-micro-batches gradient.  Two steps: the second one runs on the buffers re-laid out in the observed
# gradient-ready order.  Recompute both without the step machinery and compare.
ref_model = Tiny(); ref_model.load_state_dict(init)
for it in range(2):
  loss = step(inputs, labels)
  tot = None
  for sl in (slice(0, 2), slice(2, 4)):
    out = ref_model(**{k: v[sl] for k, v in inputs.items()})
    l = tasks.pretraining_losses({k: v[sl] for k, v in labels.items()}, out) / 2
    gs = torch.autograd.grad(l, list(ref_model.parameters()))
    tot = gs if tot is None else [a + b for a, b in zip(tot, gs)]
  flat_ref = torch.cat([t.reshape(-1) for t in tot])
  dist.all_reduce(flat_ref)
  off = 0
  with torch.no_grad():
    for p_ in ref_model.parameters():
      p_ -= 0.1 * flat_ref[off:off + p_.numel()].view_as(p_)
      off += p_.numel()
assert step._ordered
want = torch.cat([p.detach().reshape(-1) for p in ref_model.parameters()])
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
print(json.dumps({"rank": rank, "loss": float(loss), "checksum": float(flat.double().sum()), "norm": float(flat.norm()),
                  "buckets": nb, "err_vs_manual": float((flat - want).abs().max())}))
dist.destroy_process_group()
'''


def test_gradient_allreduce_keeps_replicas_identical(tmp_path):
  script = tmp_path / 'worker.py'
  script.write_text(WORKER)
  for attempt in range(3):      # the probed port can be taken between the probe and the rendezvous: retried
    with socket.socket() as s:
      s.bind(('127.0.0.1', 0))
      port = s.getsockname()[1]
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                          '--master-addr', '127.0.0.1', '--master-port', str(port), str(script)],
                         capture_output=True, text=True, env=dict(os.environ, MLT_ROOT=ROOT), timeout=240)
    if out.returncode == 0:
      break
  assert out.returncode == 0, out.stderr[-2000:]
  rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{')]
  assert len(rows) == 2
  assert rows[0]['loss'] != rows[1]['loss']                       # different data per replica
  assert abs(rows[0]['checksum'] - rows[1]['checksum']) < 1e-9    # identical parameters after the step
  assert rows[0]['buckets'] >= 3                                  # the bucketed, hook-driven path was exercised
  assert max(r['err_vs_manual'] for r in rows) < 1e-6             # = SGD on the replica-summed gradient


def _long_model(device, layers_n=2):
  enc = mmt_encoder.MmtEncoder(vocab_size=128, hidden_size=128, num_hidden_layers=layers_n, num_attention_heads=2,
                               intermediate_size=256, relative_vocab_size=32, relative_pos_max_distance=12,
                               hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                               use_pre_activation_order=True, patch_embedding_size=48, local_radius=64,
                               num_global_tokens=16)
  return tasks.MmtPretrainingModel(enc, mpp_output_num_classes=64,
                                   classification_heads=[tasks.ClassificationHead(128, 2, 'itm', 0.0)]).to(device)


def _long_batch(b, l, g, device):
  gen = torch.Generator().manual_seed(3)
  lengths = torch.randint(l // 2, l + 1, (b,), generator=gen)
  le = (torch.arange(l)[None] < lengths[:, None]).int()
  compact = fu.CompactSideInputs(le.to(device), torch.ones(b, g, dtype=torch.int32, device=device),
                                 ((torch.arange(l) * g) // l)[None].expand(b, l).int().contiguous().to(device), 12)
  inputs = {'word_ids': torch.randint(0, 128, (b, l), generator=gen).to(device),
            'patch_embeddings': torch.randn(b, 36, 48, generator=gen).to(device),
            'mlm_positions': torch.randint(40, l // 2, (b, 8), generator=gen).to(device),
            'mpp_positions': torch.randint(2, 38, (b, 6), generator=gen).to(device)}
  labels = {'mlm_label_ids': torch.randint(0, 128, (b, 8), generator=gen).to(device),
            'mlm_label_weights': torch.ones(b, 8, device=device),
            'mpp_label_ids': torch.randint(0, 64, (b, 6), generator=gen).to(device),
            'mpp_label_weights': torch.ones(b, 6, device=device),
            'itm_label_ids': torch.randint(0, 2, (b,), generator=gen).to(device),
            'itm_label_weights': torch.ones(b, device=device)}
  return inputs, labels, compact


@pytest.mark.gpu
def test_e2e_pretraining_step_and_retrieval_on_long_inputs():
  torch.manual_seed(0)
  dev = torch.device('cuda')
  model = _long_model(dev)
  inputs, labels, compact = _long_batch(4, 256, 16, dev)
  opt = torch.optim.AdamW(model.parameters(), lr=2e-3, weight_decay=0.01)
  step = tasks.PretrainingStep(model, opt, micro_batch_size=2)
  losses = [float(step(inputs, labels, compact_side_inputs=compact)) for _ in range(6)]
  assert all(torch.isfinite(torch.tensor(losses)))
  assert losses[-1] < losses[0]          # memorises the fixed batch
  scores = tasks.retrieval_scores(model, [dict(inputs, compact_side_inputs=compact)])
  assert scores.shape == (4,) and bool(((scores >= 0) & (scores <= 1)).all())


# ---- retrieval: pair enumeration, sharding, labels, recall@k pinned to the reference function ---------

def test_recall_at_k_matches_reference_function_outputs():
  """tests/golden/recall_golden.json holds outputs of the reference's own get_recall_at_k_from_dataframe
  (src/prediction_helper.py:30-89, executed by tests/golden/make_golden_recall.py), including ties and
  examples that do not share one candidate pool."""
  import json, pathlib
  golden = json.loads((pathlib.Path(__file__).parent / 'golden' / 'recall_golden.json').read_text())
  assert len(golden['cases']) >= 5
  for case in golden['cases']:
    got = tasks.get_recall_at_k(torch.tensor(case['image_index']), torch.tensor(case['text_index']),
                                torch.tensor(case['gt_image_index']), torch.tensor(case['output'], dtype=torch.float64))
    assert list(got) == list(case['expected'])
    for key, want in case['expected'].items():
      assert f'{got[key]:.4f}' == want, (key, got[key], want)


def test_pair_enumeration_sharding_and_labels():
  text, image = tasks.enumerate_image_text_pairs(num_images=3, num_texts=2)
  # text outer, image inner (reference retrieval_dataloader.py:188-195)
  assert text.tolist() == [0, 0, 0, 1, 1, 1] and image.tolist() == [0, 1, 2, 0, 1, 2]
  shards = [tasks.shard_pairs(6, 4, r).tolist() for r in range(4)]
  assert shards == [[0, 4], [1, 5], [2], [3]]                      # dataset.shard semantics
  assert sorted(sum(shards, [])) == list(range(6))                 # a partition: no pair lost or doubled
  gt = torch.tensor([2, 0])[text]
  label, weight = tasks.retrieval_labels(image, gt, pos_weight=4.0)
  assert label.tolist() == [0, 0, 1, 1, 0, 0] and weight.tolist() == [1, 1, 4, 4, 1, 1]
  with pytest.raises(ValueError):
    tasks.shard_pairs(6, 2, 2)
