"""torch.autograd bindings of the C ABI (``include/mlt_attn.h``).

PyTorch is plumbing only: device memory, the current stream and the autograd
graph.  All arithmetic happens in ``libmlt_attn.so``.  The functions here are
the tested stand-in for the TF custom op of INTEGRATION.md (TensorFlow is not
installable in this image).

Public functions
  ``dense_relative_attention``   -- QkvRelativeAttention.call core
  ``global_local_attention``     -- FusedGlobalLocalAttention core  (rows a3/a4/a7)
  ``build_dense_side_inputs`` / ``build_gl_side_inputs`` -- device constructors (a5/a6)
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from ._lib import (DenseGrads, DenseParams, GlGrads, GlParams, IdLayout, LocalGrads,
                   LocalParams, RelTables, Tensor4)
from .feature_utils import CompactSideInputs

NEG = -1e9  # large_compatible_negative [UPSTREAM-RECALLED], SURVEY.md note N2

_GL_SIDE_KEYS = ('l2l_att_mask', 'l2l_relative_att_ids', 'l2g_att_mask',
                 'l2g_relative_att_ids', 'g2g_att_mask', 'g2g_relative_att_ids',
                 'g2l_att_mask', 'g2l_relative_att_ids')


def _dtype_enum(t: torch.Tensor) -> int:
  if t.dtype == torch.float32:
    return _lib.MLT_F32
  if t.dtype == torch.bfloat16:
    return _lib.MLT_BF16
  raise TypeError(f'unsupported dtype {t.dtype}; use float32 or bfloat16')


def _prep(t: torch.Tensor) -> torch.Tensor:
  """[B, len, H, d] with d contiguous and strides a multiple of 4 elements."""
  if t.dim() != 4:
    raise ValueError(f'expected [B, len, H, d], got {tuple(t.shape)}')
  if not t.is_cuda:
    raise _lib.MltLibraryError('tensors must live on a CUDA device: there is no CPU path')
  ok = t.stride(3) == 1 and all(s % 4 == 0 for s in t.stride()[:3]) and t.data_ptr() % 16 == 0
  return t if ok else t.contiguous()


def _t4(t: torch.Tensor) -> Tensor4:
  return Tensor4(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else t.data_ptr()


def _int32(t: Optional[torch.Tensor], shape, name: str) -> Optional[torch.Tensor]:
  if t is None:
    return None
  if tuple(t.shape) != tuple(shape):
    raise ValueError(f'{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}')
  if not t.is_cuda:
    raise _lib.MltLibraryError(f'{name} must live on the CUDA device')
  return t.to(torch.int32).contiguous()


def _stream(device) -> int:
  return torch.cuda.current_stream(device).cuda_stream


def _same_device(*tensors):
  """All CUDA tensors of one call must live on one device; returns it."""
  dev = None
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise _lib.MltLibraryError('tensors must live on a CUDA device: there is no CPU path')
    if dev is None:
      dev = t.device
    elif t.device != dev:
      raise ValueError(f'all tensors of one call must share a device, got {dev} and {t.device}')
  return dev


def _dropout_args(dropout_p: float, dropout_seed: Optional[int]):
  """Validated (p, seed).  A missing seed is drawn from torch's CPU generator (reproducible under
  ``torch.manual_seed``); the same seed is replayed by the backward pass."""
  p = float(dropout_p)
  if not 0.0 <= p < 1.0:
    raise ValueError(f'dropout_p must lie in [0, 1), got {p}')
  if p == 0.0:
    return 0.0, 0
  if dropout_seed is None:
    dropout_seed = int(torch.randint(0, 2**62, (1,)).item())
  return p, int(dropout_seed) & (2**64 - 1)


def _tables(emb, bias, h, d, dtype, name):
  if emb is None and bias is None:
    return None, None, 0
  if emb is None or bias is None:
    raise ValueError(f'{name}: relative_emb_table and relative_bias_table go together')
  r = emb.shape[0]
  if tuple(emb.shape) != (r, h, d) or tuple(bias.shape) != (r, h):
    raise ValueError(f'{name}: expected emb [R,{h},{d}] and bias [R,{h}], got '
                     f'{tuple(emb.shape)} / {tuple(bias.shape)}')
  return emb.to(dtype).contiguous(), bias.to(dtype).contiguous(), r


# ---------------------------------------------------------------------------
# Contract (B)


class _GlCfg:
  """Non-tensor arguments of one global-local call."""

  def __init__(self, local_radius, side, impl, dropout_p=0.0, dropout_seed=0, neg=NEG):
    self.local_radius = local_radius
    self.side = side
    self.impl = impl
    self.dropout_p = dropout_p
    self.dropout_seed = dropout_seed
    self.neg = neg


def _fill_gl_params(p: GlParams, lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, r_vocab,
                    cfg: _GlCfg, keep: list):
  b, l, h, d = lq.shape
  g = gq.shape[1]
  p.abi_version = _lib.MLT_ABI_VERSION
  p.dtype = _dtype_enum(lq)
  p.impl = _lib.IMPL[cfg.impl]
  p.B, p.L, p.G, p.H, p.d, p.R = b, l, g, h, d, r_vocab
  p.local_radius = cfg.local_radius
  p.scale = 1.0 / math.sqrt(d)
  p.neg = cfg.neg
  p.dropout_p = cfg.dropout_p
  p.dropout_seed = cfg.dropout_seed
  p.long_q, p.long_k, p.long_v = _t4(lq), _t4(lk), _t4(lv)
  p.global_q, p.global_k, p.global_v = _t4(gq), _t4(gk), _t4(gv)
  p.long_tables = RelTables(_ptr(lemb), _ptr(lbias))
  p.global_tables = RelTables(_ptr(gemb), _ptr(gbias))
  side = cfg.side
  if hasattr(side, 'long_example_ids'):  # CompactSideInputs (duck-typed)
    p.side_mode = _lib.MLT_SIDE_COMPACT
    le = _int32(side.long_example_ids, (b, l), 'long_example_ids')
    ge = _int32(side.global_example_ids, (b, g), 'global_example_ids')
    sid = _int32(side.sentence_ids, (b, l), 'sentence_ids')
    keep += [le, ge, sid]
    p.long_example_ids, p.global_example_ids, p.sentence_ids = _ptr(le), _ptr(ge), _ptr(sid)
    p.max_distance = side.relative_pos_max_distance
  else:
    p.side_mode = _lib.MLT_SIDE_EXPLICIT
    side = side or {}
    w = 2 * cfg.local_radius + 1
    shapes = {'l2l': (b, l, w), 'l2g': (b, l, g), 'g2g': (b, g, g), 'g2l': (b, g, l)}
    for key in _GL_SIDE_KEYS:
      t = _int32(side.get(key), shapes[key[:3]], key)
      keep.append(t)
      setattr(p, key, _ptr(t))


class _GlobalLocalAttnFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, cfg):
    lib = _lib.load()
    dev = _same_device(lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias)
    lq, lk, lv, gq, gk, gv = map(_prep, (lq, lk, lv, gq, gk, gv))
    b, l, h, d = lq.shape
    g = gq.shape[1]
    dt = lq.dtype
    ctx.table_dtypes = [None if t is None else t.dtype for t in (lemb, lbias, gemb, gbias)]
    lemb, lbias, r1 = _tables(lemb, lbias, h, d, dt, 'long tables')
    gemb, gbias, r2 = _tables(gemb, gbias, h, d, dt, 'global tables')
    if r1 != r2:
      raise ValueError('long and global relative tables must share relative_vocab_size')
    long_out = torch.empty((b, l, h, d), dtype=dt, device=lq.device)
    global_out = torch.empty((b, g, h, d), dtype=dt, device=lq.device)
    long_stats = torch.empty((b, h, l, 2), dtype=torch.float32, device=lq.device)
    global_stats = torch.empty((b, h, g, 2), dtype=torch.float32, device=lq.device)
    p = GlParams()
    keep = []
    _fill_gl_params(p, lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, r1, cfg, keep)
    p.long_out, p.global_out = _t4(long_out), _t4(global_out)
    p.long_stats, p.global_stats = long_stats.data_ptr(), global_stats.data_ptr()
    nbytes = lib.mlt_gl_workspace_bytes(C.byref(p), 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=lq.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    with torch.cuda.device(dev):
      _lib.check(lib.mlt_gl_attn_fwd(C.byref(p), _stream(dev)), 'mlt_gl_attn_fwd')
    ctx.cfg = cfg
    ctx.r_vocab = r1
    ctx.save_for_backward(lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, long_out,
                          global_out, long_stats, global_stats)
    return long_out, global_out

  @staticmethod
  def backward(ctx, d_long_out, d_global_out):
    lib = _lib.load()
    (lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, long_out, global_out, long_stats,
     global_stats) = ctx.saved_tensors
    b, l, h, d = lq.shape
    r_vocab = ctx.r_vocab
    d_long_out = _prep(d_long_out.to(lq.dtype))
    d_global_out = _prep(d_global_out.to(lq.dtype))
    p = GlParams()
    keep = []
    _fill_gl_params(p, lq, lk, lv, gq, gk, gv, lemb, lbias, gemb, gbias, r_vocab, ctx.cfg, keep)
    p.long_out, p.global_out = _t4(long_out), _t4(global_out)
    p.long_stats, p.global_stats = long_stats.data_ptr(), global_stats.data_ptr()
    nbytes = lib.mlt_gl_workspace_bytes(C.byref(p), 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=lq.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    grads = [torch.empty_like(t) for t in (lq, lk, lv, gq, gk, gv)]
    gr = GlGrads()
    gr.d_long_out, gr.d_global_out = _t4(d_long_out), _t4(d_global_out)
    (gr.d_long_q, gr.d_long_k, gr.d_long_v, gr.d_global_q, gr.d_global_k,
     gr.d_global_v) = map(_t4, grads)
    tab = [None] * 4
    if r_vocab > 0:
      # one flat fp32 buffer, four views: a single conversion launch hands them back in the tables' dtype
      sizes = [r_vocab * h * d, r_vocab * h, r_vocab * h * d, r_vocab * h]
      shapes = [(r_vocab, h, d), (r_vocab, h), (r_vocab, h, d), (r_vocab, h)]
      flat = torch.empty(sum(sizes), dtype=torch.float32, device=lq.device)
      tab = [v.view(sh) for v, sh in zip(flat.split(sizes), shapes)]
      gr.d_long_emb, gr.d_long_bias, gr.d_global_emb, gr.d_global_bias = (
          t.data_ptr() for t in tab)
    with torch.cuda.device(lq.device):
      _lib.check(lib.mlt_gl_attn_bwd(C.byref(p), C.byref(gr), _stream(lq.device)), 'mlt_gl_attn_bwd')
    if r_vocab > 0:   # table gradients are accumulated in fp32; hand them back in the tables' own dtype
      if len(set(ctx.table_dtypes)) == 1:
        if ctx.table_dtypes[0] != torch.float32:
          tab = [v.view(sh) for v, sh in zip(flat.to(ctx.table_dtypes[0]).split(sizes), shapes)]
      else:
        tab = [t.to(dt) for t, dt in zip(tab, ctx.table_dtypes)]
    return (*grads, *tab, None)


def global_local_attention(long_q, long_k, long_v, global_q, global_k, global_v,
                           long_emb=None, long_bias=None, global_emb=None, global_bias=None,
                           *, local_radius: int, side=None, impl: str = 'auto',
                           dropout_p: float = 0.0, dropout_seed: Optional[int] = None,
                           neg: float = NEG):
  """Core of ``FusedGlobalLocalAttention.call`` [UPSTREAM-RECALLED].

  Args:
    long_q/k/v: ``[B, L, H, d]``; global_q/k/v: ``[B, G, H, d]`` (fp32 or bf16, CUDA).
    long_emb/bias, global_emb/bias: relative tables ``[R, H, d]`` / ``[R, H]`` of the
      long-side and global-side cores (``relative_emb_table`` / ``relative_bias_table``).
    local_radius: window radius r (long token i sees long tokens ``|j - i| <= r``).
    side: either a dict with the eight explicit int32 tensors (``l2l_att_mask`` ...
      ``g2l_relative_att_ids``; missing mask = all ones, missing ids = no relative term)
      or a ``CompactSideInputs`` (masks / ids rebuilt inside the kernels).
    impl: ``'auto' | 'simt' | 'tc'``.
    dropout_p / dropout_seed: attention-probability dropout (reference default 0.1 in training,
      ``src/configs/encoders.py:87-88``), applied to the softmax output inside the kernels; the keep
      mask is a counter-based hash of (seed, batch, head, row, key), regenerated in the backward.
    neg: the additive mask constant (ABI field ``neg``).

  Returns ``(long_out [B,L,H,d], global_out [B,G,H,d])``; differentiable w.r.t. the six
  q/k/v tensors and the four tables.
  """
  if local_radius < 1:
    raise ValueError('`local_radius` must be positive.')
  dropout_p, dropout_seed = _dropout_args(dropout_p, dropout_seed)
  cfg = _GlCfg(local_radius, side, impl, dropout_p, dropout_seed, neg)
  return _GlobalLocalAttnFn.apply(long_q, long_k, long_v, global_q, global_k, global_v,
                                  long_emb, long_bias, global_emb, global_bias, cfg)


# ---------------------------------------------------------------------------
# Contract (A)


class DenseCompactSideInputs:
  """``example_ids [B,S]`` (+ optional 2-D layout) for in-kernel mask / id construction."""

  def __init__(self, q_example_ids, k_example_ids=None, max_distance=0,
               num_patch_per_row=0, num_core_layers=0):
    self.q_example_ids = q_example_ids
    self.k_example_ids = q_example_ids if k_example_ids is None else k_example_ids
    self.max_distance = max_distance
    self.num_patch_per_row = num_patch_per_row
    self.num_core_layers = num_core_layers

  def slice(self, sl):
    return DenseCompactSideInputs(self.q_example_ids[sl], self.k_example_ids[sl], self.max_distance,
                                  self.num_patch_per_row, self.num_core_layers)


class _DenseCfg:

  def __init__(self, att_mask, relative_att_ids, compact, impl, dropout_p=0.0, dropout_seed=0, neg=NEG):
    self.att_mask = att_mask
    self.relative_att_ids = relative_att_ids
    self.compact = compact
    self.impl = impl
    self.dropout_p = dropout_p
    self.dropout_seed = dropout_seed
    self.neg = neg


def _fill_dense_params(p: DenseParams, q, k, v, emb, bias, r_vocab, cfg: _DenseCfg, keep):
  b, lq, h, d = q.shape
  lk = k.shape[1]
  p.abi_version = _lib.MLT_ABI_VERSION
  p.dtype = _dtype_enum(q)
  p.impl = _lib.IMPL[cfg.impl]
  p.B, p.Lq, p.Lk, p.H, p.d, p.R = b, lq, lk, h, d, r_vocab
  p.scale = 1.0 / math.sqrt(d)
  p.neg = cfg.neg
  p.dropout_p = cfg.dropout_p
  p.dropout_seed = cfg.dropout_seed
  p.q, p.k, p.v = _t4(q), _t4(k), _t4(v)
  p.tables = RelTables(_ptr(emb), _ptr(bias))
  if cfg.compact is not None:
    c = cfg.compact
    p.side_mode = _lib.MLT_SIDE_COMPACT
    qe = _int32(c.q_example_ids, (b, lq), 'q_example_ids')
    ke = _int32(c.k_example_ids, (b, lk), 'k_example_ids')
    keep += [qe, ke]
    p.q_example_ids, p.k_example_ids = _ptr(qe), _ptr(ke)
    p.id_layout = IdLayout(c.num_patch_per_row, c.num_core_layers, c.max_distance)
  else:
    p.side_mode = _lib.MLT_SIDE_EXPLICIT
    m = _int32(cfg.att_mask, (b, lq, lk), 'att_mask')
    ids = _int32(cfg.relative_att_ids, (b, lq, lk), 'relative_att_ids')
    keep += [m, ids]
    p.att_mask, p.relative_att_ids = _ptr(m), _ptr(ids)


class _DenseRelAttnFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, q, k, v, emb, bias, cfg):
    lib = _lib.load()
    dev = _same_device(q, k, v, emb, bias)
    q, k, v = map(_prep, (q, k, v))
    b, lq, h, d = q.shape
    ctx.table_dtypes = [None if t is None else t.dtype for t in (emb, bias)]
    emb, bias, r_vocab = _tables(emb, bias, h, d, q.dtype, 'tables')
    out = torch.empty((b, lq, h, d), dtype=q.dtype, device=q.device)
    stats = torch.empty((b, h, lq, 2), dtype=torch.float32, device=q.device)
    p = DenseParams()
    keep = []
    _fill_dense_params(p, q, k, v, emb, bias, r_vocab, cfg, keep)
    p.out, p.stats = _t4(out), stats.data_ptr()
    nbytes = lib.mlt_dense_workspace_bytes(C.byref(p), 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    with torch.cuda.device(dev):
      _lib.check(lib.mlt_dense_rel_attn_fwd(C.byref(p), _stream(dev)), 'mlt_dense_rel_attn_fwd')
    ctx.cfg, ctx.r_vocab = cfg, r_vocab
    ctx.save_for_backward(q, k, v, emb, bias, out, stats)
    return out

  @staticmethod
  def backward(ctx, d_out):
    lib = _lib.load()
    q, k, v, emb, bias, out, stats = ctx.saved_tensors
    b, lq, h, d = q.shape
    r_vocab = ctx.r_vocab
    d_out = _prep(d_out.to(q.dtype))
    p = DenseParams()
    keep = []
    _fill_dense_params(p, q, k, v, emb, bias, r_vocab, ctx.cfg, keep)
    p.out, p.stats = _t4(out), stats.data_ptr()
    nbytes = lib.mlt_dense_workspace_bytes(C.byref(p), 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    gr = DenseGrads()
    gr.d_out, gr.d_q, gr.d_k, gr.d_v = _t4(d_out), _t4(dq), _t4(dk), _t4(dv)
    d_emb = d_bias = None
    if r_vocab > 0:
      d_emb = torch.empty((r_vocab, h, d), dtype=torch.float32, device=q.device)
      d_bias = torch.empty((r_vocab, h), dtype=torch.float32, device=q.device)
      gr.d_emb, gr.d_bias = d_emb.data_ptr(), d_bias.data_ptr()
    with torch.cuda.device(q.device):
      _lib.check(lib.mlt_dense_rel_attn_bwd(C.byref(p), C.byref(gr), _stream(q.device)),
                 'mlt_dense_rel_attn_bwd')
    if r_vocab > 0:
      d_emb, d_bias = d_emb.to(ctx.table_dtypes[0]), d_bias.to(ctx.table_dtypes[1])
    return dq, dk, dv, d_emb, d_bias, None


def dense_relative_attention(q, k, v, emb=None, bias=None, att_mask=None,
                             relative_att_ids=None, compact: Optional[DenseCompactSideInputs] = None,
                             impl: str = 'auto', dropout_p: float = 0.0,
                             dropout_seed: Optional[int] = None, neg: float = NEG):
  """``QkvRelativeAttention.call`` core [UPSTREAM-RECALLED].

  ``q [B,Lq,H,d]``, ``k/v [B,Lk,H,d]``; ``att_mask`` / ``relative_att_ids`` int32
  ``[B,Lq,Lk]`` exactly as the reference feeds them
  (``src/modeling/models/mmt_encoder.py:220-224``), or ``compact`` descriptors.
  """
  dropout_p, dropout_seed = _dropout_args(dropout_p, dropout_seed)
  cfg = _DenseCfg(att_mask, relative_att_ids, compact, impl, dropout_p, dropout_seed, neg)
  return _DenseRelAttnFn.apply(q, k, v, emb, bias, cfg)


# ---------------------------------------------------------------------------
# Long rows only (QkvRelativeLocalAttention)


class LocalCompactSideInputs:
  """Compact descriptors for the long rows: example ids of the long / side tokens + sentence ids."""

  def __init__(self, example_ids, side_example_ids=None, sentence_ids=None, max_distance=0):
    self.example_ids = example_ids
    self.side_example_ids = side_example_ids
    self.sentence_ids = sentence_ids
    self.max_distance = max_distance


class _LocalCfg:

  def __init__(self, local_radius, att_mask, relative_att_ids, side_att_mask,
               side_relative_att_ids, compact, impl, dropout_p=0.0, dropout_seed=0, neg=NEG):
    self.dropout_p = dropout_p
    self.dropout_seed = dropout_seed
    self.neg = neg
    self.local_radius = local_radius
    self.att_mask = att_mask
    self.relative_att_ids = relative_att_ids
    self.side_att_mask = side_att_mask
    self.side_relative_att_ids = side_relative_att_ids
    self.compact = compact
    self.impl = impl


def _fill_local_params(p: LocalParams, q, k, v, sk, sv, emb, bias, r_vocab, cfg: _LocalCfg, keep):
  b, l, h, d = q.shape
  g = 0 if sk is None else sk.shape[1]
  p.abi_version = _lib.MLT_ABI_VERSION
  p.dtype = _dtype_enum(q)
  p.impl = _lib.IMPL[cfg.impl]
  p.B, p.L, p.G, p.H, p.d, p.R = b, l, g, h, d, r_vocab
  p.local_radius = cfg.local_radius
  p.scale = 1.0 / math.sqrt(d)
  p.neg = cfg.neg
  p.dropout_p = cfg.dropout_p
  p.dropout_seed = cfg.dropout_seed
  p.q, p.k, p.v = _t4(q), _t4(k), _t4(v)
  if g:
    p.side_k, p.side_v = _t4(sk), _t4(sv)
  p.tables = RelTables(_ptr(emb), _ptr(bias))
  w = 2 * cfg.local_radius + 1
  if cfg.compact is not None:
    c = cfg.compact
    p.side_mode = _lib.MLT_SIDE_COMPACT
    e = _int32(c.example_ids, (b, l), 'example_ids')
    se = _int32(c.side_example_ids, (b, g), 'side_example_ids') if g else None
    sid = _int32(c.sentence_ids, (b, l), 'sentence_ids') if c.sentence_ids is not None else None
    keep += [e, se, sid]
    p.example_ids, p.side_example_ids, p.sentence_ids = _ptr(e), _ptr(se), _ptr(sid)
    p.max_distance = c.max_distance
  else:
    p.side_mode = _lib.MLT_SIDE_EXPLICIT
    m = _int32(cfg.att_mask, (b, l, w), 'att_mask')
    ids = _int32(cfg.relative_att_ids, (b, l, w), 'relative_att_ids')
    sm = _int32(cfg.side_att_mask, (b, l, g), 'side_att_mask') if g else None
    sids = _int32(cfg.side_relative_att_ids, (b, l, g), 'side_relative_att_ids') if g else None
    keep += [m, ids, sm, sids]
    p.att_mask, p.relative_att_ids = _ptr(m), _ptr(ids)
    p.side_att_mask, p.side_relative_att_ids = _ptr(sm), _ptr(sids)


class _LocalRelAttnFn(torch.autograd.Function):

  @staticmethod
  def forward(ctx, q, k, v, sk, sv, emb, bias, cfg):
    lib = _lib.load()
    dev = _same_device(q, k, v, sk, sv, emb, bias)
    q, k, v = map(_prep, (q, k, v))
    if sk is not None:
      sk, sv = _prep(sk), _prep(sv)
    b, l, h, d = q.shape
    ctx.table_dtypes = [None if t is None else t.dtype for t in (emb, bias)]
    emb, bias, r_vocab = _tables(emb, bias, h, d, q.dtype, 'tables')
    out = torch.empty((b, l, h, d), dtype=q.dtype, device=q.device)
    stats = torch.empty((b, h, l, 2), dtype=torch.float32, device=q.device)
    p = LocalParams()
    keep = []
    _fill_local_params(p, q, k, v, sk, sv, emb, bias, r_vocab, cfg, keep)
    p.out, p.stats = _t4(out), stats.data_ptr()
    nbytes = lib.mlt_local_workspace_bytes(C.byref(p), 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    with torch.cuda.device(dev):
      _lib.check(lib.mlt_local_rel_attn_fwd(C.byref(p), _stream(dev)), 'mlt_local_rel_attn_fwd')
    ctx.cfg, ctx.r_vocab = cfg, r_vocab
    ctx.save_for_backward(q, k, v, sk, sv, emb, bias, out, stats)
    return out

  @staticmethod
  def backward(ctx, d_out):
    lib = _lib.load()
    q, k, v, sk, sv, emb, bias, out, stats = ctx.saved_tensors
    b, l, h, d = q.shape
    r_vocab = ctx.r_vocab
    d_out = _prep(d_out.to(q.dtype))
    p = LocalParams()
    keep = []
    _fill_local_params(p, q, k, v, sk, sv, emb, bias, r_vocab, ctx.cfg, keep)
    p.out, p.stats = _t4(out), stats.data_ptr()
    nbytes = lib.mlt_local_workspace_bytes(C.byref(p), 1)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    gr = LocalGrads()
    gr.d_out, gr.d_q, gr.d_k, gr.d_v = _t4(d_out), _t4(dq), _t4(dk), _t4(dv)
    dsk = dsv = None
    if sk is not None:
      dsk, dsv = torch.empty_like(sk), torch.empty_like(sv)
      gr.d_side_k, gr.d_side_v = _t4(dsk), _t4(dsv)
    d_emb = d_bias = None
    if r_vocab > 0:
      d_emb = torch.empty((r_vocab, h, d), dtype=torch.float32, device=q.device)
      d_bias = torch.empty((r_vocab, h), dtype=torch.float32, device=q.device)
      gr.d_emb, gr.d_bias = d_emb.data_ptr(), d_bias.data_ptr()
    with torch.cuda.device(q.device):
      _lib.check(lib.mlt_local_rel_attn_bwd(C.byref(p), C.byref(gr), _stream(q.device)),
                 'mlt_local_rel_attn_bwd')
    if r_vocab > 0:
      d_emb, d_bias = d_emb.to(ctx.table_dtypes[0]), d_bias.to(ctx.table_dtypes[1])
    return dq, dk, dv, dsk, dsv, d_emb, d_bias, None


def local_relative_attention(q, k, v, emb=None, bias=None, *, local_radius: int, att_mask=None,
                             relative_att_ids=None, side_keys=None, side_values=None,
                             side_att_mask=None, side_relative_att_ids=None,
                             compact: Optional[LocalCompactSideInputs] = None, impl: str = 'auto',
                             dropout_p: float = 0.0, dropout_seed: Optional[int] = None,
                             neg: float = NEG):
  """``QkvRelativeLocalAttention.call`` core [UPSTREAM-RECALLED].

  ``q/k/v [B,L,H,d]``; window masks / ids ``[B,L,2r+1]`` (column k <-> key ``i+k-r``);
  optional side keys / values ``[B,G,H,d]`` with ``[B,L,G]`` side masks / ids.  One softmax over
  the window and the side keys.
  """
  if local_radius < 1:
    raise ValueError('`local_radius` must be positive.')
  if (side_keys is None) != (side_values is None):
    raise ValueError('`side_keys` and `side_values` go together.')
  dropout_p, dropout_seed = _dropout_args(dropout_p, dropout_seed)
  cfg = _LocalCfg(local_radius, att_mask, relative_att_ids, side_att_mask, side_relative_att_ids,
                  compact, impl, dropout_p, dropout_seed, neg)
  return _LocalRelAttnFn.apply(q, k, v, side_keys, side_values, emb, bias, cfg)


# ---------------------------------------------------------------------------
# Device-side constructors


def build_dense_side_inputs(example_ids: torch.Tensor, max_distance: int,
                            num_patch_per_row: int = 0, num_core_layers: int = 0,
                            want_mask: bool = True, want_ids: bool = True):
  """``example_ids [B,S]`` -> ``(att_mask, relative_att_ids)`` int32 ``[B,S,S]`` on device."""
  lib = _lib.load()
  e = _int32(example_ids, example_ids.shape, 'example_ids')
  b, s = e.shape
  mask = torch.empty((b, s, s), dtype=torch.int32, device=e.device) if want_mask else None
  ids = torch.empty((b, s, s), dtype=torch.int32, device=e.device) if want_ids else None
  layout = IdLayout(num_patch_per_row, num_core_layers, max_distance)
  with torch.cuda.device(e.device):
    _lib.check(lib.mlt_build_dense_side_inputs(e.data_ptr(), b, s, layout, _ptr(mask), _ptr(ids),
                                               _stream(e.device)), 'mlt_build_dense_side_inputs')
  return mask, ids


def build_gl_side_inputs(compact: CompactSideInputs, local_radius: int):
  """Compact descriptors -> dict of the eight explicit int32 side inputs, on device."""
  lib = _lib.load()
  b, l = compact.long_example_ids.shape
  g = compact.global_example_ids.shape[1]
  le = _int32(compact.long_example_ids, (b, l), 'long_example_ids')
  ge = _int32(compact.global_example_ids, (b, g), 'global_example_ids')
  sid = _int32(compact.sentence_ids, (b, l), 'sentence_ids')
  w = 2 * local_radius + 1
  shapes = {'l2l': (b, l, w), 'l2g': (b, l, g), 'g2g': (b, g, g), 'g2l': (b, g, l)}
  out = {k: torch.empty(shapes[k[:3]], dtype=torch.int32, device=le.device)
         for k in _GL_SIDE_KEYS}
  arr = (C.c_void_p * 8)(*[out[k].data_ptr() for k in _GL_SIDE_KEYS])
  with torch.cuda.device(le.device):
    _lib.check(lib.mlt_build_gl_side_inputs(le.data_ptr(), ge.data_ptr(), sid.data_ptr(), b, l, g,
                                            local_radius, compact.relative_pos_max_distance,
                                            C.byref(arr), _stream(le.device)), 'mlt_build_gl_side_inputs')
  return out


def compact_from_explicit_dense(att_mask: torch.Tensor, relative_att_ids: torch.Tensor,
                                num_patch_per_row: int = 0, num_core_layers: int = 0,
                                max_distance: Optional[int] = None) -> Optional[DenseCompactSideInputs]:
  """Recognition of generator-shaped explicit side inputs of the dense path (the reference's own call
  signature, ``src/modeling/models/mmt_encoder.py:220-224``): returns the compact descriptors when they
  reproduce EVERY element of ``att_mask [B,S,S]`` and ``relative_att_ids [B or 1,S,S]`` exactly, else ``None``.

  Meant to be called once per batch (the tensors are shared by all layers); it reads one int32 flag back,
  i.e. synchronises the stream once.  ``num_patch_per_row > 0`` checks the ids against that 2-D image + text
  layout (``src/feature_utils.py:114-184``; ``max_distance`` required); otherwise the 1-D rule, with the
  distance read from the ids when ``max_distance`` is None.
  """
  lib = _lib.load()
  if att_mask is None or relative_att_ids is None or att_mask.dim() != 3 or att_mask.shape[1] != att_mask.shape[2]:
    return None
  b, s, _ = att_mask.shape
  if s < 2 or relative_att_ids.dim() != 3 or tuple(relative_att_ids.shape[1:]) != (s, s) or \
      relative_att_ids.shape[0] not in (1, b):
    return None
  if num_patch_per_row > 0 and max_distance is None:
    raise ValueError('the 2-D layout needs `max_distance`.')
  mask = _int32(att_mask, (b, s, s), 'att_mask')
  ids = _int32(relative_att_ids.expand(b, s, s), (b, s, s), 'relative_att_ids')
  dev = mask.device
  qe = torch.empty((b, s), dtype=torch.int32, device=dev)
  ke = torch.empty((b, s), dtype=torch.int32, device=dev)
  res = torch.empty(4, dtype=torch.int32, device=dev)
  hint = IdLayout(num_patch_per_row, num_core_layers, -1 if max_distance is None else max_distance)
  with torch.cuda.device(dev):
    _lib.check(lib.mlt_dense_compact_from_explicit(mask.data_ptr(), ids.data_ptr(), b, s, hint, qe.data_ptr(),
                                                   ke.data_ptr(), res.data_ptr(), _stream(dev)),
               'mlt_dense_compact_from_explicit')
  ok, dist = res[:2].tolist()
  if not ok:
    return None
  return DenseCompactSideInputs(qe, ke, dist, num_patch_per_row, num_core_layers)


def compact_from_explicit_gl(side: dict, local_radius: int) -> Optional[CompactSideInputs]:
  """Same for the eight tensors of ``FusedGlobalLocalAttention.call`` [UPSTREAM-RECALLED]: returns
  ``CompactSideInputs`` when the compact rules reproduce every element of all eight, else ``None``."""
  lib = _lib.load()
  if any(side.get(k) is None for k in _GL_SIDE_KEYS):
    return None
  l2l, l2g = side['l2l_att_mask'], side['l2g_att_mask']
  if l2l.dim() != 3 or l2g.dim() != 3 or l2l.shape[2] != 2 * local_radius + 1:
    return None
  b, l, w = l2l.shape
  g = l2g.shape[2]
  shapes = {'l2l': (b, l, w), 'l2g': (b, l, g), 'g2g': (b, g, g), 'g2l': (b, g, l)}
  ts = []
  for k in _GL_SIDE_KEYS:
    t = side[k]
    if t.dim() != 3 or tuple(t.shape[1:]) != shapes[k[:3]][1:] or t.shape[0] not in (1, b):
      return None
    ts.append(_int32(t.expand(*shapes[k[:3]]), shapes[k[:3]], k))
  dev = ts[0].device
  le = torch.empty((b, l), dtype=torch.int32, device=dev)
  ge = torch.empty((b, g), dtype=torch.int32, device=dev)
  sid = torch.empty((b, l), dtype=torch.int32, device=dev)
  res = torch.empty(4, dtype=torch.int32, device=dev)
  arr = (C.c_void_p * 8)(*[t.data_ptr() for t in ts])
  with torch.cuda.device(dev):
    _lib.check(lib.mlt_gl_compact_from_explicit(C.byref(arr), b, l, g, local_radius, le.data_ptr(), ge.data_ptr(),
                                                sid.data_ptr(), res.data_ptr(), _stream(dev)),
               'mlt_gl_compact_from_explicit')
  ok, dist = res[:2].tolist()
  if not ok:
    return None
  return CompactSideInputs(le, ge, sid, dist)
