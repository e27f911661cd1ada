// Row-centric score evaluation for the tcgen05 kernels (forward and query-centric backward).
//
// A thread owns one query row i and evaluates, for a group of 32 consecutive keys,
//     t[jj] = (x[jj] + allrel[i, id(i, j)]) * scale (+ neg if masked), -inf if (i, j) is dead,
// where x is the raw Q.K^T accumulator.  The expensive general form (per-element id rule, table
// gather, per-element mask) is only needed on a minority of groups.  Per (warp, group) a
// warp-uniform classifier picks one of:
//   GM_DEAD  every pair outside the band / beyond the key range         -> nothing to do
//   GM_FAST  all live, mask uniform over the keys, relative term constant per row
//            -> t = fma(x, scale, c_row): ONE instruction per element
//   GM_EDGE  as FAST but the band edge / key tail cuts the group          (+ live select)
//   GM_DIAG  1-D rule near the diagonal: gather from the offset-ordered row table
//   GM_QS    l2g group containing some row's own sentence column          (+ select)
//   GM_KS    g2l group containing keys of some row's sentence             (+ shuffle, select)
//   GM_GEN   anything else (explicit int32 side inputs, 2-D ids, mask boundary inside group)
// Key-side scalars (example id, sentence id) live one-per-lane in registers and reach the
// per-element code through warp shuffles: no shared memory, no extra synchronisation.
//
// The per-row relative table rel_s is stored in SLOT order so that the 1-D rule gathers with
// slot = clamp(j - i, -D, D) + D:  slots [0, 2D] hold offsets -D..D, slots > 2D are the id itself.
// (Permutation is only active when 2D + 1 <= R; otherwise slot == id and GM_DIAG is never chosen.)
#pragma once

#include "mlt_common.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace rowscore {

constexpr int TMR = 128;  // rows per tile == stride of the transposed row tables

enum GMode : int { GM_DEAD = 0, GM_FAST, GM_EDGE, GM_DIAG, GM_QS, GM_KS, GM_GEN };

__host__ __device__ __forceinline__ int slot_of_id(int id, int D, bool perm) {
  if (!perm) return id;
  if (id <= D) return D + id;
  if (id <= 2 * D) return 2 * D - id;
  return id;
}

struct SegCtx {        // warp-uniform, one per key segment
  const KeySeg* sg;
  int kb, ke;          // key range of this tile in the segment
  int R, D;            // D: max_distance of THIS segment's id rule
  int pd;              // max_distance the row table was permuted with
  bool perm;           // slot permutation active
  bool band;
  int radius;
  int mask_rule, id_rule;
};

struct RowCtx {        // per thread
  int i, row;
  bool row_ok;
  int q_e, q_sent;
  float relP, relN, relX, relX1;  // rel for offset >= D, <= -D, cross "other", cross "same sentence"
};

struct GroupLanes {    // lane l holds the scalars of key (g0 + l)
  int ke_l, ks_l;
};

__device__ __forceinline__ GroupLanes load_group_lanes(const SegCtx& sc, int b, int g0, int lane) {
  GroupLanes gl{0, -1};
  const Side& sd = sc.sg->side;
  const int j = g0 + lane;
  if (j >= 0 && j < sc.sg->len) {
    if (sc.mask_rule == MR_EXAMPLE_ID) gl.ke_l = __ldg(sd.k_eid + (int64_t)b * sd.k_len + j);
    if (sc.id_rule == IDR_CROSS_KSENT) gl.ks_l = __ldg(sd.sent + (int64_t)b * sd.sent_len + j);
  }
  return gl;
}

// Part 1: global loads of the row's own scalars (issue early, latency overlaps the Q/E loads).
__device__ __forceinline__ void init_row_loads(RowCtx& rc, const SegCtx& sc, int b) {
  const Side& sd = sc.sg->side;
  rc.q_e = 0;
  rc.q_sent = -1;
  if (rc.row_ok && sc.mask_rule == MR_EXAMPLE_ID) rc.q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + rc.i);
  if (rc.row_ok && sc.id_rule == IDR_CROSS_QSENT) rc.q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + rc.i);
}
// Part 2: per-row constants of the relative table (after rel_s has been written).
__device__ __forceinline__ void init_row(RowCtx& rc, const SegCtx& sc, int b, const float* rel_s) {
  auto rel_at = [&](int id) -> float {
    return (id >= 0 && id < sc.R) ? rel_s[slot_of_id(id, sc.pd, sc.perm) * TMR + rc.row] : 0.f;
  };
  rc.relP = rel_at(sc.D);
  rc.relN = rel_at(2 * sc.D);
  rc.relX = rel_at(2 * sc.D + 1);
  rc.relX1 = rel_at(2 * sc.D + 2);
}

struct GroupPlan {
  int mode;
  float cadd;     // per-row additive constant (rel * scale already folded in rel_s) + mask term
  int ccls;       // which constant class the relative term came from: 0 none, 1 +D, 2 -D, 3 cross
  float mrow;     // mask term alone (0 or neg)
  bool mask_pe;   // example-id mask changes inside the group: per-element post-pass (warp-uniform)
};

// a = first row of the warp.  All decisions are warp-uniform (computed from uniform values or
// votes); cadd / mrow are per-thread.
// W = group width (32, or 16 when four threads share a row); the lane-held scalars always cover
// the 32-key window starting at g0 - sub.
template <int W = 32>
__device__ __forceinline__ GroupPlan classify(const SegCtx& sc, const RowCtx& rc, const GroupLanes& gl,
                                              int a, int g0, int lane, float neg, int sub = 0) {
  GroupPlan gp;
  gp.cadd = 0.f;
  gp.ccls = 0;
  gp.mrow = 0.f;
  gp.mask_pe = false;
  const int o_min = g0 - (a + 31), o_max = g0 + (W - 1) - a;
  const bool dead = g0 >= sc.ke || (sc.band && (o_min > sc.radius || o_max < -sc.radius));
  if (dead) {
    gp.mode = GM_DEAD;
    return gp;
  }
  const bool all_live = (g0 + (W - 1) < sc.ke) && (!sc.band || (o_min >= -sc.radius && o_max <= sc.radius));
  // ---- mask ----
  bool mask_uniform = true;
  if (sc.mask_rule == MR_EXPLICIT) {
    mask_uniform = false;
  } else if (sc.mask_rule == MR_EXAMPLE_ID) {
    const int ke0 = __shfl_sync(0xffffffffu, gl.ke_l, sub);
    const bool lane_oob = (g0 - sub + lane >= sc.ke) || (W < 32 && (lane < sub || lane >= sub + W));
    const bool uni = __all_sync(0xffffffffu, lane_oob || gl.ke_l == ke0);
    gp.mask_pe = !uni;
    gp.mrow = uni ? ((rc.q_e == ke0) ? 0.f : neg) : 0.f;
  }
  // ---- relative term ----
  int rcls;  // 0 const, 1 diag, 2 qs, 3 ks, 4 generic
  float relc = 0.f;
  switch (sc.id_rule) {
    case IDR_NONE:
      rcls = 0;
      break;
    case IDR_1D:
      if (!sc.perm || sc.pd != sc.D) {
        rcls = 4;
      } else if (o_min >= sc.D) {
        rcls = 0; relc = rc.relP; gp.ccls = 1;
      } else if (o_max <= -sc.D) {
        rcls = 0; relc = rc.relN; gp.ccls = 2;
      } else {
        rcls = 1;
      }
      break;
    case IDR_CROSS_QSENT: {
      const bool hit = __any_sync(0xffffffffu, rc.q_sent >= g0 && rc.q_sent < g0 + W);
      rcls = hit ? 2 : 0;
      relc = rc.relX;
      gp.ccls = 3;
      break;
    }
    case IDR_CROSS_KSENT: {
      const bool in_grp = (W == 32) || (lane >= sub && lane < sub + W);
      const int smin = __reduce_min_sync(0xffffffffu, (gl.ks_l < 0 || !in_grp) ? 0x7fffffff : gl.ks_l);
      const int smax = __reduce_max_sync(0xffffffffu, in_grp ? gl.ks_l : -1);
      const bool hit = !(smax < a || smin > a + 31);
      rcls = hit ? 3 : 0;
      relc = rc.relX;
      gp.ccls = 3;
      break;
    }
    default:
      rcls = 4;
  }
  gp.cadd = relc + gp.mrow;
  if (!mask_uniform || rcls == 4) {   // explicit mask tensor or a rule without a fast form
    gp.mode = GM_GEN;
    gp.mask_pe = false;
  } else if (rcls == 0) {
    gp.mode = all_live ? GM_FAST : GM_EDGE;
  } else if (!all_live) {
    gp.mode = GM_GEN;
  } else {
    gp.mode = rcls == 1 ? GM_DIAG : (rcls == 2 ? GM_QS : GM_KS);
  }
  return gp;
}

// Generic per-element evaluation (any rule).  Returns the slot (or -1) through `slot`.
__device__ __forceinline__ float score_generic(float x, const SegCtx& sc, const RowCtx& rc,
                                               const GroupLanes& gl, int b, int g0, int jj,
                                               const float* rel_s, float scale, float neg, int& slot,
                                               int sub = 0) {
  const Side& sd = sc.sg->side;
  const int j = g0 + jj;
  const int off = j - rc.i;
  slot = -1;
  // shuffles first: every lane of the warp must take part, live or not
  const int ke_j = __shfl_sync(0xffffffffu, gl.ke_l, sub + jj);
  const int ks_j = __shfl_sync(0xffffffffu, gl.ks_l, sub + jj);
  const bool live = j < sc.ke && (!sc.band || (off <= sc.radius && off >= -sc.radius));
  if (!live) return -INFINITY;
  const int col = sc.band ? off + sc.radius : j;
  bool ok = true;
  int id = -1;
  switch (sc.mask_rule) {
    case MR_EXPLICIT: ok = rc.row_ok ? (__ldg(sd.mask + (int64_t)b * sd.sb + (int64_t)rc.i * sd.sq + col) != 0) : true; break;
    case MR_EXAMPLE_ID: ok = (rc.q_e == ke_j); break;
    default: break;
  }
  switch (sc.id_rule) {
    case IDR_EXPLICIT: id = rc.row_ok ? __ldg(sd.ids + (int64_t)b * sd.sb + (int64_t)rc.i * sd.sq + col) : -1; break;
    case IDR_1D: id = rel_id_1d(off, sc.D); break;
    case IDR_CROSS_QSENT: id = 2 * sc.D + 1 + (rc.q_sent == j ? 1 : 0); break;
    case IDR_CROSS_KSENT: id = 2 * sc.D + 1 + (ks_j == rc.i ? 1 : 0); break;
    case IDR_2D: id = rel_id_2d(rc.i, j, sd.npr, sd.core, sc.D); break;
    default: break;
  }
  float rel = 0.f;
  if (id >= 0 && id < sc.R) {
    slot = slot_of_id(id, sc.pd, sc.perm);
    rel = rel_s[slot * TMR + rc.row];
  }
  float v = fmaf(x, scale, rel);
  if (!ok) v += neg;
  return v;
}

// Scores of one 32-key group, in place (t[OFF .. OFF+32)).  `mode` is warp-uniform.
// The shuffles inside GM_GEN / GM_KS require all 32 lanes to execute this function together.
template <int OFF, int N, int W = 32>
__device__ __forceinline__ void score_group(float (&t)[N], const GroupPlan& gp, const SegCtx& sc,
                                            const RowCtx& rc, const GroupLanes& gl, int b, int g0,
                                            const float* rel_s, float scale, float neg, int sub = 0) {
  switch (gp.mode) {
    case GM_DEAD:
#pragma unroll
      for (int jj = 0; jj < W; ++jj) t[OFF + jj] = -INFINITY;
      break;
    case GM_FAST:
#pragma unroll
      for (int jj = 0; jj < W; ++jj) t[OFF + jj] = fmaf(t[OFF + jj], scale, gp.cadd);
      break;
    case GM_EDGE: {
      // live columns of this row form one interval [jlo, jhi): branch-free per-element test
      const int d0 = g0 - rc.i;
      int jlo = 0, jhi = min(W, sc.ke - g0);
      if (sc.band) {
        jlo = max(jlo, -sc.radius - d0);
        jhi = min(jhi, sc.radius - d0 + 1);
      }
      const unsigned span = (unsigned)max(jhi - jlo, 0);
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const float v = fmaf(t[OFF + jj], scale, gp.cadd);
        t[OFF + jj] = ((unsigned)(jj - jlo) < span) ? v : -INFINITY;
      }
      break;
    }
    case GM_DIAG: {
      const int d0 = g0 - rc.i + sc.D;  // slot = clamp(off, -D, D) + D = clamp(d0 + jj, 0, 2D)
      const float* base = rel_s + rc.row;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int s = min(max(d0 + jj, 0), 2 * sc.D);
        t[OFF + jj] = fmaf(t[OFF + jj], scale, base[s * TMR] + gp.mrow);
      }
      break;
    }
    case GM_QS: {
      const int d0 = rc.q_sent - g0;  // special column index within the group
      const float c0 = rc.relX + gp.mrow, c1 = rc.relX1 + gp.mrow;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) t[OFF + jj] = fmaf(t[OFF + jj], scale, d0 == jj ? c1 : c0);
      break;
    }
    case GM_KS: {
      const float c0 = rc.relX + gp.mrow, c1 = rc.relX1 + gp.mrow;
#pragma unroll
      for (int jj = 0; jj < W; ++jj) {
        const int ks_j = __shfl_sync(0xffffffffu, gl.ks_l, sub + jj);
        t[OFF + jj] = fmaf(t[OFF + jj], scale, ks_j == rc.i ? c1 : c0);
      }
      break;
    }
    default:
      return;  // GM_GEN: already evaluated in place by score_group_generic_tmem()
  }
  if (gp.mask_pe && gp.mode != GM_DEAD) {
#pragma unroll
    for (int jj = 0; jj < W; ++jj) {
      const int ke_j = __shfl_sync(0xffffffffu, gl.ke_l, sub + jj);
      t[OFF + jj] += (ke_j == rc.q_e) ? 0.f : neg;
    }
  }
}

// GM_GEN groups: evaluate the 32 scores in place inside TMEM with a real loop (one copy of the
// generic code instead of 32 unrolled ones -- the kernels must stay instruction-cache resident).
// `taddr` = lane-selected TMEM address of the group's first column.  Whole warp, converged.
__device__ __forceinline__ void score_group_generic_tmem(uint32_t taddr, const SegCtx& sc, const RowCtx& rc,
                                                      const GroupLanes& gl, int b, int g0,
                                                      const float* rel_s, float scale, float neg) {
#pragma unroll 1
  for (int jj = 0; jj < 32; ++jj) {
    const uint32_t raw = ptx::tmem_ld1(taddr + jj);
    ptx::tmem_wait_ld();
    int slot;
    const float v = score_generic(__uint_as_float(raw), sc, rc, gl, b, g0, jj, rel_s, scale, neg, slot);
    ptx::tmem_st1(taddr + jj, __float_as_uint(v));
  }
  ptx::tmem_wait_st();
}

}  // namespace rowscore
}  // namespace mlt
