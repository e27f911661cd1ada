"""Forward of the 12-layer encoder mirror (BASELINE.json configs[1] shape: S = 512 = 2 + 196 patches +
314 text tokens, d = 768, 12 heads, bf16) with the side inputs fed the three ways the operator accepts:
explicit [B,S,S] int32 tensors as the reference feeds them (2-D patch ids), compact 2-D descriptors, and
compact 1-D descriptors.  Reports ms per forward and the share spent inside the attention kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import mlt_b200  # noqa: F401
from mlt_b200 import _lib, mmt_encoder, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S, NPR, D, R = 512, 14, 12, 32
torch.manual_seed(0)
enc = mmt_encoder.MmtEncoder(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                             intermediate_size=3072, relative_vocab_size=R, relative_pos_max_distance=D,
                             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                             patch_embedding_size=768).cuda().to(torch.bfloat16).eval()
word_ids = torch.randint(0, 30522, (B, S)).cuda()
patches = torch.randn(B, NPR * NPR, 768).cuda().to(torch.bfloat16)
lengths = torch.randint(S // 2, S + 1, (B,))
eid = (torch.arange(S)[None, :] < lengths[:, None]).to(torch.int32).cuda()
mask2d, ids2d = ops.build_dense_side_inputs(eid, D, num_patch_per_row=NPR, num_core_layers=2)
cases = {
    'explicit [B,S,S] mask + 2-D ids (reference call), explicit kernels': dict(att_mask=mask2d, relative_att_ids=ids2d),
    'explicit [B,S,S] mask + 2-D ids (reference call), recognised once per forward': dict(
        att_mask=mask2d, relative_att_ids=ids2d),
    'compact 2-D descriptors': dict(compact=ops.DenseCompactSideInputs(eid, max_distance=D, num_patch_per_row=NPR,
                                                                       num_core_layers=2)),
    'compact 1-D descriptors': dict(compact=ops.DenseCompactSideInputs(eid, max_distance=D)),
}


def run(kw):
  with torch.no_grad():
    return enc(word_ids, patch_embeddings=patches, training=False, **kw)['sequence_output']


for name, kw in cases.items():
  enc.transformer_layers.recognize_side_inputs = 'recognised' in name
  enc.transformer_layers.id_layout_hint = (NPR, 2, D)
  for _ in range(2):
    run(kw)
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(5):
    run(kw)
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / 5
  _lib.profile_enable(True)
  run(kw)
  torch.cuda.synchronize()
  recs = _lib.profile_read()
  _lib.profile_enable(False)
  att = sum(r[1] for r in recs)
  print(f'{name}: {ms:.2f} ms per forward (B {B}, S {S}, 12 layers) = {B * S / ms * 1e3 / 1e6:.2f} M tokens/s; '
        f'attention kernels {att:.2f} ms ({len(recs)} launches: {sorted(set(r[0] for r in recs))})')
