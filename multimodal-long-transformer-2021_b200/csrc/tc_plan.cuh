// Chunk planner for the tcgen05 kernels.
//
// A score tile is 128 rows x 64 columns; its elementwise work is done per (32-row warp quadrant,
// 32-column group).  For every such pair a *planner warp* (one spare warp of the CTA, running a few
// chunks ahead) decides ONCE which evaluation form applies and publishes the decision in a small
// shared-memory ring, so that the elementwise warps -- the critical resource of these kernels, they
// are bound by the MUFU (ex2) rate -- spend no instructions on classification:
//   DEAD  every pair outside the band / beyond the column range          -> no work, P = 0
//   FAST  all live, mask uniform over the columns, relative term constant per row
//         -> p = ex2(fma(x, scale*log2e, c_row)): ONE FMA per element, nothing written back
//   EDGE  as FAST but the band edge / sequence end cuts the group
//   DIAG  1-D rule near the diagonal: gather from the slot-ordered row table
//   QS    cross block that contains the sentence column of some row      (+ select)
//   KS    cross block that contains columns belonging to some row's sentence (+ select)
//   EXPL  explicit int32 mask / id tensors (the reference's own call signature): the row's 32
//         consecutive entries are loaded with 128-bit accesses where the layout allows
//   GEN   anything else (2-D ids, mixed rules, ...)
// "Row" = the index the thread owns (TMEM lane), "column" = the index that varies inside a chunk.
// In the forward and the query-centric backward rows are queries and columns are keys; in the
// key-centric backward rows are keys and columns are queries (the planner is told which).
// The ring slot also carries the per-column scalars (example id, sentence id) of the chunk, read
// by the rare per-element forms as warp-broadcast LDS.
#pragma once

#include "mlt_common.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace plan {

enum Mode : int { DEAD = 0, FAST, EDGE, DIAG, QS, KS, GEN, EXPL };
// constant relative classes of a group (which per-row constant applies)
// C_MODAL: 2-D image + text layout, text row x image column or image row x text column (one id per row kind)
enum CCls : int { C_NONE = 0, C_POS = 1, C_NEG = 2, C_CROSS = 3, C_MODAL = 4 };

constexpr uint32_t F_MASK_PE = 1u << 16;   // example-id mask changes inside the group

struct ChunkPlan {          // one ring slot (576 B)
  // [row quadrant]: { w0 of group 0, w0 of group 1, ce0 of group 0, ce0 of group 1 } (one LDS.128)
  //   w0  = mode | ccls << 8 | flags
  //   ce0 = example id shared by the group's columns (uniform-mask groups)
  uint32_t q[4][4];
  int32_t ce[64];           // column example ids
  int32_t cs[64];           // column sentence ids (rule KSENT on query-rows / QSENT on key-rows)
};

__host__ __device__ __forceinline__ int slot_of_id(int id, int D, bool perm) {
  if (!perm) return id;
  if (id <= D) return D + id;
  if (id <= 2 * D) return 2 * D - id;
  return id;
}

// Warp-uniform description of one (row set, column segment) block for the planner.
struct PSeg {
  int c_begin, c_end;    // live column range of this tile inside the segment
  int c_len;             // length of the column sequence (loads are clipped to it)
  bool band;
  int radius;
  int mask_rule, id_rule;
  int D, R;
  bool diag_ok;          // the slot-ordered row table matches this segment's 1-D rule
  bool rows_are_keys;    // key-centric pass: offset = row - column instead of column - row
  const int32_t* c_eid;  // per-column example ids  [B, c_len] (or null)
  const int32_t* c_sent; // per-column sentence ids [B, c_len] (or null)
  int64_t c_eid_stride, c_sent_stride;
  bool col_sent;         // sentence ids live on the column side (KS form); else on the row side (QS)
  bool expl_ok;          // the kernel instantiation carries the EXPL form (else explicit tensors take GEN)
  int n_img;             // > 0: the explicit ids are the library's own plane of the 2-D image + text layout
                         // (first n_img positions = patches) and the kernel carries C_MODAL: only
                         // image x image groups read the plane (reference src/feature_utils.py:114-184)
};
// id of a cross-modality pair by the query's kind (rel_id_2d: image_part_id = n_img + 8 + 2 D + 1)
__host__ __device__ __forceinline__ int modal_id(int query_pos, int n_img, int D) {
  return n_img + 8 + 2 * D + 1 + (query_pos >= n_img ? 0 : 1);
}

// Per-tile row-side sentence ranges (QS form): [rs_min[w], rs_max[w]] of quadrant w.
struct RowSent {
  int mn[4], mx[4];
};

// Per-column scalars of one chunk, one pair per lane (columns col0 + lane and col0 + 32 + lane).
struct ColLanes {
  int e0, e1, s0, s1;
};
__device__ __forceinline__ ColLanes load_col_lanes(const PSeg& s, int b, int col0, int lane) {
  ColLanes cl{0, 0, -1, -1};
  const int j0 = col0 + lane, j1 = col0 + 32 + lane;
  if (s.mask_rule == MR_EXAMPLE_ID) {
    if (j0 >= 0 && j0 < s.c_len) cl.e0 = __ldg(s.c_eid + (int64_t)b * s.c_eid_stride + j0);
    if (j1 >= 0 && j1 < s.c_len) cl.e1 = __ldg(s.c_eid + (int64_t)b * s.c_eid_stride + j1);
  }
  if (s.col_sent) {
    if (j0 >= 0 && j0 < s.c_len) cl.s0 = __ldg(s.c_sent + (int64_t)b * s.c_sent_stride + j0);
    if (j1 >= 0 && j1 < s.c_len) cl.s1 = __ldg(s.c_sent + (int64_t)b * s.c_sent_stride + j1);
  }
  return cl;
}

// Whole warp.  col0 = first column of the chunk, row0 = first row of the tile; `cl` = the chunk's
// column scalars (load_col_lanes, issued a chunk ahead by the caller).
__device__ __forceinline__ void plan_chunk(const PSeg& s, const ColLanes& cl, int col0, int row0, const RowSent& rs,
                                           ChunkPlan* out, int lane) {
  const int e0 = cl.e0, e1 = cl.e1, s0 = cl.s0, s1 = cl.s1;
  const int j0 = col0 + lane, j1 = col0 + 32 + lane;
  out->ce[lane] = e0;
  out->ce[32 + lane] = e1;
  out->cs[lane] = s0;
  out->cs[32 + lane] = s1;
  // group-uniform facts
  const int f0 = __shfl_sync(0xffffffffu, e0, 0), f1 = __shfl_sync(0xffffffffu, e1, 0);
  const bool uni0 = __all_sync(0xffffffffu, j0 >= s.c_end || e0 == f0);
  const bool uni1 = __all_sync(0xffffffffu, j1 >= s.c_end || e1 == f1);
  int smin0 = 0x7fffffff, smax0 = -1, smin1 = 0x7fffffff, smax1 = -1;
  if (s.col_sent) {
    smin0 = __reduce_min_sync(0xffffffffu, s0 < 0 ? 0x7fffffff : s0);
    smax0 = __reduce_max_sync(0xffffffffu, s0);
    smin1 = __reduce_min_sync(0xffffffffu, s1 < 0 ? 0x7fffffff : s1);
    smax1 = __reduce_max_sync(0xffffffffu, s1);
  }
  if (lane < 8) {
    const int w = lane >> 1, g = lane & 1;
    const int a = row0 + 32 * w;          // first row of the quadrant
    const int g0 = col0 + 32 * g;         // first column of the group
    // offset = key - query
    int o_min, o_max;
    if (!s.rows_are_keys) {
      o_min = g0 - (a + 31);
      o_max = g0 + 31 - a;
    } else {
      o_min = a - (g0 + 31);
      o_max = a + 31 - g0;
    }
    uint32_t mode, ccls = C_NONE, flags = 0;
    const bool dead = g0 >= s.c_end || (s.band && (o_min > s.radius || o_max < -s.radius));
    // explicit tensors (ids explicit or absent; the mask explicit, absent, or the example-id rule --
    // the latter with an explicit id plane is how compact 2-D ids are served)
    bool expl = s.expl_ok && (s.mask_rule == MR_EXPLICIT || s.id_rule == IDR_EXPLICIT) &&
                (s.id_rule == IDR_EXPLICIT || s.id_rule == IDR_NONE);
    int id_rule = s.id_rule;
    if (expl && s.n_img > 0 && s.mask_rule == MR_EXAMPLE_ID) {
      // text x text pairs follow the 1-D rule, text x image / image x text pairs carry one id per QUERY kind
      // (query text -> image_part_id, query image -> text_part_id); the test is symmetric in rows and columns
      const bool r_txt = a >= s.n_img, r_img = a + 31 < s.n_img;
      const bool c_txt = g0 >= s.n_img, c_img = g0 + 31 < s.n_img;
      if (r_txt && c_txt) {
        id_rule = IDR_1D;
        expl = false;
      } else if ((r_txt && c_img) || (r_img && c_txt)) {
        id_rule = IDR_NONE;
        ccls = C_MODAL;
        expl = false;
      }
    }
    if (dead) {
      mode = DEAD;
    } else if (expl) {
      mode = EXPL;
      if (s.mask_rule == MR_EXAMPLE_ID && !(g ? uni1 : uni0)) flags |= F_MASK_PE;
    } else {
      const bool all_live = (g0 + 31 < s.c_end) && (!s.band || (o_min >= -s.radius && o_max <= s.radius));
      bool gen = false;
      if (s.mask_rule == MR_EXPLICIT) gen = true;
      if (s.mask_rule == MR_EXAMPLE_ID && !(g ? uni1 : uni0)) flags |= F_MASK_PE;
      int rcls = 0;  // 0 const, 1 diag, 2 qs, 3 ks, 4 generic
      switch (id_rule) {
        case IDR_NONE:
          break;
        case IDR_1D:
          if (!s.diag_ok) rcls = 4;
          else if (o_min >= s.D) ccls = C_POS;
          else if (o_max <= -s.D) ccls = C_NEG;
          else rcls = 1;
          break;
        case IDR_CROSS_QSENT:
        case IDR_CROSS_KSENT:
          if (2 * s.D + 2 >= s.R) { rcls = 4; break; }
          ccls = C_CROSS;
          if (s.col_sent) {   // sentence id on the column side: does a column's sentence hit a row?
            const int smin = g ? smin1 : smin0, smax = g ? smax1 : smax0;
            if (!(smax < a || smin > a + 31)) rcls = 3;
          } else {            // sentence id on the row side: does a row's sentence hit a column?
            const int rmn = w == 0 ? rs.mn[0] : (w == 1 ? rs.mn[1] : (w == 2 ? rs.mn[2] : rs.mn[3]));
            const int rmx = w == 0 ? rs.mx[0] : (w == 1 ? rs.mx[1] : (w == 2 ? rs.mx[2] : rs.mx[3]));
            if (!(rmx < g0 || rmn > g0 + 31)) rcls = 2;
          }
          break;
        default:
          rcls = 4;
      }
      if (gen || rcls == 4) {
        mode = GEN;
        flags = 0;
      } else if (rcls == 0) {
        // FAST folds the mask into one per-row constant: a mask that changes inside the group
        // takes the EDGE form (per-element post-pass)
        mode = (all_live && !(flags & F_MASK_PE)) ? FAST : EDGE;
      } else if (!all_live) {
        mode = GEN;
        flags = 0;
      } else {
        mode = rcls == 1 ? DIAG : (rcls == 2 ? QS : KS);
      }
    }
#ifdef MLT_FORCE_GEN
    if (mode != DEAD) { mode = GEN; flags = 0; }
#endif
    out->q[w][g] = mode | (ccls << 8) | flags;
    out->q[w][2 + g] = (uint32_t)(g ? f1 : f0);
  }
}

// ---- explicit side inputs -----------------------------------------------------------------------
// 32 consecutive int32 entries of one row of an explicit mask / id tensor, entry x taken only where
// bit x of `take` is set (others = fill).  `vec` (warp-uniform; implies take == all) = the 32 entries
// of every lane are 16-byte aligned: eight 128-bit loads.
__device__ __forceinline__ void load_row32(const int32_t* p, bool vec, uint32_t take, int fill, int (&out)[32]) {
  if (vec) {
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(p) + x);
      out[4 * x] = v.x;
      out[4 * x + 1] = v.y;
      out[4 * x + 2] = v.z;
      out[4 * x + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int x = 0; x < 32; ++x) out[x] = (take >> x) & 1u ? __ldg(p + x) : fill;
  }
}
// bit x set <=> lo <= x < hi  (0 <= lo, hi <= 32)
__device__ __forceinline__ uint32_t span_bits(int lo, int hi) {
  if (hi <= lo) return 0u;
  const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return upto_hi & ~((1u << lo) - 1u);
}

// ---- per-row relative table ---------------------------------------------------------------------
// allrel = Q.E^T sits in TMEM (lane = row, column = id).  The elementwise warps need it as
// rel_s[slot][row] = (allrel[row, id] + bias[id]) * scale in shared memory (slot order, see
// slot_of_id).  relmeta[id] = (bias[id] * scale, slot * rows-per-tile) is prepared once per tile by
// 64 threads, so that the transposition costs 4 instructions per id and row.
struct RelMeta {
  float bias_scaled;
  int slot_off;     // slot * 128 (element offset of the slot's row vector)
};

// Called by threads t = 0..63 of the elementwise group (followed by a barrier of that group).
__device__ __forceinline__ void rel_meta_init(RelMeta* meta, int t, const __nv_bfloat16* bias, int H, int h,
                                              int R, int D, bool perm, float scale) {
  RelMeta m;
  m.bias_scaled = (t < R) ? __bfloat162float(bias[t * H + h]) * scale : 0.f;
  m.slot_off = (t < R ? slot_of_id(t, D, perm) : t) * 128;   // ids >= R: identity (free slots, value 0)
  meta[t] = m;
}

// Whole warp; `taddr` = lane-selected TMEM address of allrel column 0.  Returns nothing; the
// optional `keep` callback receives (id, value) for callers that also publish the id-ordered row.
template <typename F>
__device__ __forceinline__ void rel_table_build(uint32_t taddr, const RelMeta* meta, float* rel_s, int row,
                                                int rpad, float scale, F&& keep) {
#pragma unroll 1
  for (int c0 = 0; c0 < rpad; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(taddr + c0, v);
    ptx::tmem_wait_ld();
    // all loads first: the compiler cannot prove that rel_s and meta do not alias and would
    // otherwise serialise load -> store -> load
    RelMeta m[16];
#pragma unroll
    for (int x = 0; x < 16; ++x) m[x] = meta[c0 + x];   // warp-broadcast LDS.64
    float val[16];
#pragma unroll
    for (int x = 0; x < 16; ++x) val[x] = fmaf(__uint_as_float(v[x]), scale, m[x].bias_scaled);
#pragma unroll
    for (int x = 0; x < 16; ++x) rel_s[m[x].slot_off + row] = val[x];
    keep(c0, val);
  }
}

// Row-side sentence ranges of the tile (whole warp; rows beyond `len` are ignored).
struct RowSentRaw {
  int v[4];
};
__device__ __forceinline__ RowSentRaw row_sent_load(const int32_t* sent, int64_t stride, int b, int row0, int len,
                                                    int lane) {
  RowSentRaw r;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const int i = row0 + 32 * w + lane;
    r.v[w] = (sent && i < len) ? __ldg(sent + (int64_t)b * stride + i) : -1;
  }
  return r;
}
__device__ __forceinline__ RowSent row_sent_reduce(const RowSentRaw& r) {
  RowSent rs;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    rs.mn[w] = __reduce_min_sync(0xffffffffu, r.v[w] < 0 ? 0x7fffffff : r.v[w]);
    rs.mx[w] = __reduce_max_sync(0xffffffffu, r.v[w]);
  }
  return rs;
}

// The planner warp's loop over the chunks of a tile (up to two column segments).  The global
// loads of chunk c + 1 are in flight while chunk c is classified; the row-side loads overlap the
// first chunk's.  `ring` has NPL slots guarded by pl_full / pl_empty mbarriers.
// `pre(c)` runs (whole warp) at the top of iteration c: kernels whose producer warp doubles as the
// planner issue the chunk's TMA loads there.
struct NoPre {
  __device__ __forceinline__ void operator()(int) const {}
};
template <int NPL, int TN, typename Pre = NoPre>
__device__ __forceinline__ void planner_loop(const PSeg& ps0, const PSeg& ps1, int n0, int n1, int cb0, int cb1,
                                             int b, int row0, const int32_t* row_sent, int64_t row_sent_stride,
                                             int row_len, ChunkPlan* ring, uint64_t* pl_full, uint64_t* pl_empty,
                                             int lane, Pre pre = Pre()) {
  const int nchunks = n0 + n1;
  if (nchunks == 0) return;
  const RowSentRaw raw = row_sent_load(row_sent, row_sent_stride, b, row0, row_len, lane);
  ColLanes cl = (0 < n0) ? load_col_lanes(ps0, b, cb0, lane) : load_col_lanes(ps1, b, cb1, lane);
  const RowSent rs = row_sent_reduce(raw);
#pragma unroll 1
  for (int c = 0; c < nchunks; ++c) {
    pre(c);
    const int sl = c % NPL;
    const bool first = c < n0;
    const int col0 = first ? cb0 + c * TN : cb1 + (c - n0) * TN;
    ColLanes nl{0, 0, -1, -1};
    if (c + 1 < nchunks) {
      if (c + 1 < n0) nl = load_col_lanes(ps0, b, cb0 + (c + 1) * TN, lane);
      else nl = load_col_lanes(ps1, b, cb1 + (c + 1 - n0) * TN, lane);
    }
    if (c >= NPL) ptx::mbar_wait_warp(&pl_empty[sl], ((c / NPL) & 1) ^ 1);
    if (first) plan_chunk(ps0, cl, col0, row0, rs, ring + sl, lane);
    else plan_chunk(ps1, cl, col0, row0, rs, ring + sl, lane);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&pl_full[sl]);
    cl = nl;
  }
}

}  // namespace plan
}  // namespace mlt
