// tcgen05 backward kernels.  Deterministic: every gradient element is owned by one CTA, no
// atomics.  Two passes recompute the score tile from Q, K and the saved (row max, row sum):
//
//   tc_bwd_q_kernel   query-centric, one 128-row query tile per CTA:
//        S = Q.K_c^T, dP = dO.V_c^T (SS)  ->  ds = p (dp - delta)  ->  dQ += dS.K_c (TS, dS in TMEM)
//        plus the per-row relative-id bins dallrel[i, id] += ds, dQ += dallrel.E, and it publishes
//        rowstat = (m*log2e, 1/l, delta) and allrel*scale for the key-centric pass.
//   tc_bwd_kv_kernel  key-centric, one 128-row key tile per CTA, looping over the query chunks of
//        up to two query sources:  S^T = K.Q_c^T, dP^T = V.dO_c^T (SS)  ->  P^T, dS^T (bf16, TMEM)
//        ->  dV += P^T.dO_c, dK += dS^T.Q_c (TS, B operands MN-major).
//
// Both: 1 CTA / SM, 320 threads = 8 elementwise warps (2 threads per row, 32 columns each -- the
// backward needs no row reductions) + TMA warp + MMA warp, 512 TMEM columns.
#include "tc_api.cuh"

#include <cstdlib>

#include "mlt_common.cuh"
#include "tc_plan.cuh"
#include "tc_ptx.cuh"
#include "tc_rowscore.cuh"

namespace mlt {
namespace {

using namespace ptx;

constexpr int TM = 128;
constexpr int TN = 64;
// NP = threads per row in the elementwise role (each owns 64 / NP columns of a chunk):
// 4 * NP elementwise warps + TMA warp + MMA warp.
template <int NP>
constexpr int nthreads() { return (4 * NP + 2) * 32; }

#ifdef MLT_TC_TRACE
__device__ unsigned long long g_trace_b[3][256];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(role, idx)                                                                  \
  do {                                                                                    \
    if (blockIdx.x == 5 && blockIdx.y == 1 && blockIdx.z == 0 && (idx) < 256) g_trace_b[role][idx] = gtime(); \
  } while (0)
#else
#define TRACE(role, idx) do {} while (0)
#endif
constexpr uint32_t TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <int MR, int IR>
__device__ __forceinline__ void side_ok_id(const Side& sd, int b, int i, int j, int col, int q_e, int k_e,
                                           int q_sent, int k_sent, bool& ok, int& id) {
  ok = true;
  id = -1;
  if (MR == MR_EXPLICIT) ok = __ldg(sd.mask + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col) != 0;
  if (MR == MR_EXAMPLE_ID) ok = (q_e == k_e);
  if (IR == IDR_EXPLICIT) id = __ldg(sd.ids + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col);
  if (IR == IDR_1D) id = rel_id_1d(j - i, sd.max_distance);
  if (IR == IDR_CROSS_QSENT) id = 2 * sd.max_distance + 1 + (q_sent == j ? 1 : 0);
  if (IR == IDR_CROSS_KSENT) id = 2 * sd.max_distance + 1 + (k_sent == i ? 1 : 0);
  if (IR == IDR_2D) id = rel_id_2d(i, j, sd.npr, sd.core, sd.max_distance);
}

// ============================================================================================
// Preprocess: rowstat[b,h,i] = (max * log2e, 1 / sum, delta = sum_c dO * O, 0).  HBM-bound,
// coalesced: 8 lanes x 16 B cover one 128-byte row of dO and of O.
// ============================================================================================
__global__ void __launch_bounds__(256) tc_bwd_prep_kernel(const T4 out, const T4 d_out, const float* stats,
                                                          float4* rowstat, int B, int H, int len, int lp) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t >> 3;   // row index over (b, i, h): h fastest (matches the tensor layout)
  const int sub8 = (int)(t & 7);
  const int64_t total = (int64_t)B * len * H;
  float acc = 0.f;
  int b = 0, i = 0, h = 0;
  const bool ok = r < total;
  if (ok) {
    h = (int)(r % H);
    i = (int)((r / H) % len);
    b = (int)(r / ((int64_t)H * len));
    const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(row_ptr<__nv_bfloat16>(d_out, b, i, h)) + sub8);
    const uint4 o4 = __ldg(reinterpret_cast<const uint4*>(row_ptr<__nv_bfloat16>(out, b, i, h)) + sub8);
    const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, ow[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[y]));
      const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[y]));
      acc = fmaf(gf.x, of.x, acc);
      acc = fmaf(gf.y, of.y, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && sub8 == 0) {
    const float2 st = __ldg(reinterpret_cast<const float2*>(stats) + ((int64_t)(b * H + h) * len + i));
    rowstat[(int64_t)(b * H + h) * lp + i] = make_float4(st.x * LOG2E, 1.f / st.y, acc, 0.f);
  }
}

// ============================================================================================
// Query-centric pass
// ============================================================================================
namespace bq {
constexpr int NST = 4;
constexpr int SM_Q = 0;                        // 16 KB
constexpr int SM_DO = SM_Q + TM * 128;         // 16 KB
constexpr int SM_E = SM_DO + TM * 128;         // 8 KB
constexpr int SM_KV = SM_E + 64 * 128;         // NST x 16 KB
constexpr int SM_REL = SM_KV + NST * 2 * TN * 128;   // [64][128] f32
constexpr int SM_BIN = SM_REL + 64 * TM * 4;         // 2 x [64][128] f32
constexpr int SM_A = SM_BIN + 2 * 64 * TM * 4;      // dallrel^T tile, bf16, [128 rows][64 ids] SW128 (16 KB)
constexpr int SM_BS = SM_A + TM * 128;              // bias partial sums [4 quadrants][64]
constexpr int SM_PLAN = SM_BS + 4 * 64 * 4;         // 4 x ChunkPlan
constexpr int SM_META = SM_PLAN + 4 * (int)sizeof(plan::ChunkPlan);   // 64 x RelMeta
constexpr int SM_BAR = SM_META + 64 * (int)sizeof(plan::RelMeta);
constexpr int SM_ALLOC = SM_BAR + 256 + 1024;
// TMEM columns
constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_REL = 320, T_DE = 384;

struct Bars {
  uint64_t q_full, rel_full;
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t sdp_full[2], ds_full[2], dq_full, dar_full;
  uint64_t pl_full[4], pl_empty[4];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace bq

struct SegRange {
  int kb, ke, n;
};
__device__ __forceinline__ SegRange seg_range(const KeySeg& sg, int i0) {
  SegRange r;
  if (sg.band) {
    r.kb = max(0, i0 - sg.radius);
    r.ke = min(sg.len, i0 + TM + sg.radius);
  } else {
    r.kb = 0;
    r.ke = sg.len;
  }
  r.n = (r.ke - r.kb + TN - 1) / TN;
  return r;
}

struct TcBwdQParams {
  BwdQArgs a;
  int rpad, rw, lp;   // R padded to 16 (MMA N), to 4 (workspace row), rows padded (workspace)
  float4* rowstat;    // ws [B, H, lp] = (m * log2e, 1 / l, delta, 0)
  float* allrel_ws;   // ws [B, H, lp, rw] (allrel * scale)
};

// ---- row-side contexts of the query-centric pass (rows = queries, columns = keys) -------------
struct SegC {          // warp-uniform
  const Side* sd;
  int ke;              // end of the live key range
  bool band;
  int radius;
  int mask_rule, id_rule;
  int D, R, pd;
  bool perm;
};
struct RowC {          // per thread and segment
  int q_e, q_sent;
  float relP, relN, relX, relX1;
};

__device__ __forceinline__ SegC make_segc(const KeySeg& sg, const SegRange& r, int R, int pd, bool perm) {
  SegC sc;
  sc.sd = &sg.side;
  sc.ke = r.ke;
  sc.band = sg.band != 0;
  sc.radius = sg.radius;
  sc.mask_rule = sg.side.mask_rule;
  sc.id_rule = R > 0 ? sg.side.id_rule : IDR_NONE;
  sc.D = sg.side.max_distance;
  sc.R = R;
  sc.pd = pd;
  sc.perm = perm;
  return sc;
}

__device__ __forceinline__ plan::PSeg make_pseg(const KeySeg& sg, const SegRange& r, int R, int pd, bool perm) {
  plan::PSeg s;
  s.c_begin = r.kb;
  s.c_end = r.ke;
  s.c_len = sg.len;
  s.band = sg.band != 0;
  s.radius = sg.radius;
  s.mask_rule = sg.side.mask_rule;
  s.id_rule = R > 0 ? sg.side.id_rule : IDR_NONE;
  s.D = sg.side.max_distance;
  s.R = R;
  s.diag_ok = perm && pd == s.D;
  s.rows_are_keys = false;
  s.c_eid = sg.side.k_eid;
  s.c_eid_stride = sg.side.k_len;
  s.col_sent = (s.id_rule == IDR_CROSS_KSENT);
  s.c_sent = s.col_sent ? sg.side.sent : nullptr;
  s.c_sent_stride = sg.side.sent_len;
  return s;
}

__device__ __forceinline__ void row_loads(RowC& rc, const SegC& sc, int b, int i, bool row_ok) {
  rc.q_e = 0;
  rc.q_sent = -1;
  if (row_ok && sc.mask_rule == MR_EXAMPLE_ID) rc.q_e = __ldg(sc.sd->q_eid + (int64_t)b * sc.sd->q_len + i);
  if (row_ok && sc.id_rule == IDR_CROSS_QSENT) rc.q_sent = __ldg(sc.sd->sent + (int64_t)b * sc.sd->sent_len + i);
}
__device__ __forceinline__ void row_consts(RowC& rc, const SegC& sc, const float* rel_s, int row) {
  auto rel_at = [&](int id) -> float {
    return (id >= 0 && id < sc.R) ? rel_s[plan::slot_of_id(id, sc.pd, sc.perm) * TM + row] : 0.f;
  };
  const bool on = sc.id_rule != IDR_NONE;
  rc.relP = on ? rel_at(sc.D) : 0.f;
  rc.relN = on ? rel_at(2 * sc.D) : 0.f;
  rc.relX = on ? rel_at(2 * sc.D + 1) : 0.f;
  rc.relX1 = on ? rel_at(2 * sc.D + 2) : 0.f;
}

// Generic per-element score (any rule).  Dead pairs return -inf; `slot` = bin slot or -1.
__device__ __forceinline__ float score_generic_q(float x, const SegC& sc, const RowC& rc, int b, int i, int row,
                                                 bool row_ok, int j, int ke_j, int ks_j, const float* rel_s,
                                                 float scale, float neg, int& slot) {
  const Side& sd = *sc.sd;
  const int off = j - i;
  slot = -1;
  const bool live = j < sc.ke && (!sc.band || (off <= sc.radius && off >= -sc.radius));
  if (!live) return -INFINITY;
  const int col = sc.band ? off + sc.radius : j;
  bool ok = true;
  int id = -1;
  switch (sc.mask_rule) {
    case MR_EXPLICIT: ok = row_ok ? (__ldg(sd.mask + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col) != 0) : true; break;
    case MR_EXAMPLE_ID: ok = (rc.q_e == ke_j); break;
    default: break;
  }
  switch (sc.id_rule) {
    case IDR_EXPLICIT: id = row_ok ? __ldg(sd.ids + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col) : -1; break;
    case IDR_1D: id = rel_id_1d(off, sc.D); break;
    case IDR_CROSS_QSENT: id = 2 * sc.D + 1 + (rc.q_sent == j ? 1 : 0); break;
    case IDR_CROSS_KSENT: id = 2 * sc.D + 1 + (ks_j == i ? 1 : 0); break;
    case IDR_2D: id = rel_id_2d(i, j, sd.npr, sd.core, sc.D); break;
    default: break;
  }
  float rel = 0.f;
  if (id >= 0 && id < sc.R) {
    slot = plan::slot_of_id(id, sc.pd, sc.perm);
    rel = rel_s[slot * TM + row];
  }
  float v = fmaf(x, scale, rel);
  if (!ok) v += neg;
  return v;
}

// Query-centric backward.  Each elementwise thread owns (row, one 32-key group) of the chunks its
// warp set serves: NP = 2 threads per row, SETS warp sets on alternate chunks (S / dP are
// double-buffered by chunk parity, so set s owns buffer s).  The evaluation form of every
// (quadrant, group) pair comes from the planner warp (tc_plan.cuh).
constexpr int NP = 2;
constexpr int NPL = 4;   // plan ring slots
template <int SETS>
constexpr int bq_threads() { return (4 * NP * SETS + 3) * 32; }

template <int SETS>
__global__ void __launch_bounds__(bq_threads<SETS>(), 1)
tc_bwd_q_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_v0,
                const __grid_constant__ CUtensorMap map_k1, const __grid_constant__ CUtensorMap map_v1,
                const __grid_constant__ CUtensorMap map_e, const __grid_constant__ TcBwdQParams p) {
  using namespace bq;
  constexpr int W = 32;               // columns per elementwise thread and chunk
  constexpr int NEW = 128 * NP;       // elementwise threads per set
  constexpr int NALL = NEW * SETS;    // all elementwise threads
  constexpr int NB = NP * SETS;       // private bin arrays per row
  constexpr int WP = 4 * NB, WM = 4 * NB + 1, WPL = 4 * NB + 2;   // producer / MMA / planner warp
  constexpr int RB = 128 / NB;        // bin slots per array (host guarantees R <= RB)
  static_assert(SETS == 1 || SETS == 2, "chunk buffers are double-buffered");
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: keeps the shared address space
  // (LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + SM_REL);
  float* bins = reinterpret_cast<float*>(smem + SM_BIN);
  plan::ChunkPlan* plans = reinterpret_cast<plan::ChunkPlan*>(smem + SM_PLAN);
  plan::RelMeta* relmeta = reinterpret_cast<plan::RelMeta*>(smem + SM_META);
  const BwdQArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TM;
  const int R = a.rows.R, rpad = p.rpad;

  if (tid == 0) {
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->rel_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sdp_full[s], 1);
      mbar_init(&bars->ds_full[s], NEW);
    }
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->dar_full, NALL);
    for (int s = 0; s < NPL; ++s) {
      mbar_init(&bars->pl_full[s], 1);
      mbar_init(&bars->pl_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == WM) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SegRange r0 = seg_range(a.seg[0], i0);
  SegRange r1{0, 0, 0};
  if (a.nseg > 1) r1 = seg_range(a.seg[1], i0);
  const int nchunks = r0.n + r1.n;
  const int pd = a.seg[0].side.max_distance;
  const bool perm = (2 * pd + 1 <= R);

  if (warp == WP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->q_full, 2 * TM * 128 + rpad * 128);
      tma_load_4d(smem + SM_Q, &map_q, &bars->q_full, 0, i0, h, b);
      tma_load_4d(smem + SM_DO, &map_do, &bars->q_full, 0, i0, h, b);
      if (rpad) tma_load_4d(smem + SM_E, &map_e, &bars->q_full, 0, 0, h, 0);
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % NST;
        mbar_wait(&bars->kv_empty[st], ((c / NST) & 1) ^ 1);
        const bool first = c < r0.n;
        const int key0 = first ? r0.kb + c * TN : r1.kb + (c - r0.n) * TN;
        uint8_t* ks = smem + SM_KV + st * (2 * TN * 128);
        mbar_arrive_expect_tx(&bars->kv_full[st], 2 * TN * 128);
        tma_load_4d(ks, first ? &map_k0 : &map_k1, &bars->kv_full[st], 0, key0, h, b);
        tma_load_4d(ks + TN * 128, first ? &map_v0 : &map_v1, &bars->kv_full[st], 0, key0, h, b);
      }
    }
  } else if (warp == WM) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_dq = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t q_addr = smem_u32(smem + SM_Q), do_addr = smem_u32(smem + SM_DO);
      mbar_wait(&bars->q_full, 0);
      tc_fence_after_sync();
      if (rpad) {
        const uint32_t idesc_r = make_idesc_bf16(TM, rpad, 0, 0);
        const uint32_t e_addr = smem_u32(smem + SM_E);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_REL, make_smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                  make_smem_desc_sw128(e_addr + kk * 32, 16, 1024), idesc_r, kk > 0);
        umma_commit(&bars->rel_full);
      }
      for (int c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
          const int st = c % NST;
          mbar_wait(&bars->pl_full[c % NPL], (c / NPL) & 1);   // relayed to the elementwise warps by sdp_full
          mbar_wait(&bars->kv_full[st], (c / NST) & 1);
          tc_fence_after_sync();
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
          const uint32_t v_addr = k_addr + TN * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + T_S + (c & 1) * 64, make_smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + T_DP + (c & 1) * 64, make_smem_desc_sw128(do_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(v_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&bars->sdp_full[c & 1]);
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->ds_full[pc & 1], (pc >> 1) & 1);
          tc_fence_after_sync();
          mbar_arrive(&bars->pl_empty[pc % NPL]);   // every elementwise thread is done with plan pc
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + T_DQ, tmem + T_DP + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(k_addr + kk * 2048, 16, 1024), idesc_dq, (pc > 0 || kk > 0));
          umma_commit(&bars->kv_empty[st]);
          if (pc == nchunks - 1) {
            if (rpad) {
              // dQ += dallrel . E : A = dallrel (bf16, TMEM columns of the dead allrel region),
              // B = E tile taken MN-major (K = relative ids)
              mbar_wait(&bars->dar_full, 0);
              tc_fence_after_sync();
              const uint32_t e_addr = smem_u32(smem + SM_E);
              for (int kk = 0; kk < rpad / 16; ++kk)
                umma_ts(tmem + T_DQ, tmem + T_REL + kk * 8, make_smem_desc_sw128(e_addr + kk * 2048, 16, 1024),
                        idesc_dq, 1u);
              if (a.tg_partial) {
                // table-gradient partial of this tile: dE[64 ids x 64] = dallrel^T . Q  (M = 64, both
                // operands MN-major: K = the tile's 128 rows)
                const uint32_t idesc_de = make_idesc_bf16(64, 64, 1, 1);
                const uint32_t a_addr = smem_u32(smem + SM_A);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                  umma_ss(tmem + T_DE, make_smem_desc_sw128(a_addr + kk * 2048, 16, 1024),
                          make_smem_desc_sw128(q_addr + kk * 2048, 16, 1024), idesc_de, kk > 0);
              }
            }
            umma_commit(&bars->dq_full);
          }
        }
      }
    }
  } else if (warp == WPL) {
    // ===================== planner =====================
    const plan::PSeg ps0 = make_pseg(a.seg[0], r0, R, pd, perm);
    const plan::PSeg ps1 = make_pseg(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm);
    const Side* qs_side = nullptr;
    if (ps0.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[0].side;
    if (a.nseg > 1 && ps1.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[1].side;
    const plan::RowSent rs = plan::row_sent_ranges(qs_side ? qs_side->sent : nullptr,
                                                   qs_side ? qs_side->sent_len : 0, b, i0, a.rows.len, lane);
    auto run = [&](const plan::PSeg& ps, int c_begin, int c_end, int kb) {
      for (int c = c_begin; c < c_end; ++c) {
        const int sl = c % NPL;
        if (c >= NPL) mbar_wait_warp(&bars->pl_empty[sl], ((c / NPL) & 1) ^ 1);
        plan::plan_chunk(ps, b, kb + (c - c_begin) * TN, i0, rs, plans + sl, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pl_full[sl]);
      }
    };
    run(ps0, 0, r0.n, r0.kb);
    if (a.nseg > 1) run(ps1, r0.n, nchunks, r1.kb);
  } else {
    // ===================== elementwise warps (2 threads per row) =====================
    if (tid == 0) TRACE(1, 0);
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int set = warp / (4 * NP);     // which chunk parity this warp serves
    const int part = (warp >> 2) % NP;   // which 32-key group of the chunk
    const int bidx = set * NP + part;    // private bin array / output column slice
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    float* bin = bins + bidx * RB * TM;   // slot-ordered, private to (set, part, row)
    for (int x = lane + 32 * quad; x < RB * TM; x += 128) bin[x] = 0.f;
    for (int x = tid; x < TM * 128 / 16; x += NALL) reinterpret_cast<uint4*>(smem + SM_A)[x] = make_uint4(0u, 0u, 0u, 0u);
    const SegC sc0 = make_segc(a.seg[0], r0, R, pd, perm);
    const SegC sc1 = make_segc(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm);
    RowC rc0, rc1;
    row_loads(rc0, sc0, b, i, row_ok);
    row_loads(rc1, sc1, b, i, row_ok);
    // row constants (written by tc_bwd_prep_kernel): p = exp2(t * log2e - m2l), m2l = m*log2e + log2(l)
    float m2l = INFINITY, m2 = 0.f, delta = 0.f;
    const int64_t srow = (int64_t)(b * a.H + h) * a.rows.len + i;
    const int64_t prow = (int64_t)(b * a.H + h) * p.lp + i;
    if (row_ok) {
      const float4 rs4 = __ldg(p.rowstat + prow);
      m2 = rs4.x;
      m2l = rs4.x - __log2f(rs4.y);
      delta = rs4.z;
    }
    if (rpad && tid < 64)
      plan::rel_meta_init(relmeta, tid, reinterpret_cast<const __nv_bfloat16*>(a.rows.bias), a.H, h, R, pd, perm,
                          a.scale);
    if (tid == 0) TRACE(1, 1);
    named_bar_sync(1, NALL);   // relmeta visible
    if (rpad && bidx == 0) {   // warp-uniform: the four (set 0, part 0) warps extract allrel
      mbar_wait_warp(&bars->rel_full, 0);
      if (tid == 0) TRACE(1, 2);
      tc_fence_after_sync();
      float* ws_row = p.allrel_ws + prow * p.rw;
      const int rw = p.rw;
      plan::rel_table_build(tmem + T_REL + lane_sel, relmeta, rel_s, row, rpad, a.scale,
                            [&](int c0, const float (&val)[16]) {
                              if (row_ok) {   // 16 ids = 4 x 16-byte stores (rows padded to a multiple of 4 ids)
#pragma unroll
                                for (int x4 = 0; x4 < 4; ++x4)
                                  if (c0 + 4 * x4 < rw)
                                    *reinterpret_cast<float4*>(ws_row + c0 + 4 * x4) =
                                        make_float4(c0 + 4 * x4 < R ? val[4 * x4] : 0.f, c0 + 4 * x4 + 1 < R ? val[4 * x4 + 1] : 0.f,
                                                    c0 + 4 * x4 + 2 < R ? val[4 * x4 + 2] : 0.f, c0 + 4 * x4 + 3 < R ? val[4 * x4 + 3] : 0.f);
                              }
                            });
    }
    if (tid == 0) TRACE(1, 3);
    named_bar_sync(1, NALL);  // rel_s (written by set 0 / part 0) visible to all; bins zeroed
    if (tid == 0) TRACE(1, 4);
    row_consts(rc0, sc0, rel_s, row);
    row_consts(rc1, sc1, rel_s, row);
    // per-row accumulators of the constant relative classes (flushed into the bins at the end)
    float accP = 0.f, accN = 0.f, accX = 0.f, accX1 = 0.f;
    const float scale2 = a.scale * LOG2E;

    // One call per key segment (inlined twice: no per-field selects inside the chunk loop).
    // Chunks [c_begin, c_end) belong to this segment; this warp set handles c % SETS == set.
    auto run_chunks = [&](const SegC sc, const RowC rc, int c_begin, int c_end, int kb) {
      int c = c_begin + ((set - c_begin) % SETS + SETS) % SETS;
      const bool mre = sc.mask_rule == MR_EXAMPLE_ID;
#pragma unroll 1
      for (; c < c_end; c += SETS) {
        const int g0 = kb + (c - c_begin) * TN + part * W;
        const uint32_t t_s = tmem + T_S + (c & 1) * 64 + lane_sel + part * W;
        const uint32_t t_dp = tmem + T_DP + (c & 1) * 64 + lane_sel + part * W;
        const plan::ChunkPlan* cp = plans + (c % NPL);
        if (tid == 0) TRACE(1, 8 + 3 * c);
        // the MMA warp issued S_c / dP_c only after plan c had been published
        mbar_wait_warp(&bars->sdp_full[c & 1], (c >> 1) & 1);
        if (tid == 0) TRACE(1, 9 + 3 * c);
        tc_fence_after_sync();
        const uint32_t w0 = cp->q[quad][part];
        const int ce0 = (int)cp->q[quad][2 + part];
        const int mode = (int)(w0 & 0xffu);
        const bool mask_pe = (w0 & plan::F_MASK_PE) != 0;
        const bool masked = mre && !mask_pe && (rc.q_e != ce0);
        const float mterm = masked ? a.neg : 0.f;
        const int ccls = (int)((w0 >> 8) & 0xffu);
        const float relc = ccls == plan::C_POS ? rc.relP : (ccls == plan::C_NEG ? rc.relN : (ccls == plan::C_CROSS ? rc.relX : 0.f));
        uint32_t ds_pk[W / 2];
        bool zero = (mode == plan::DEAD);
        // every row of the warp masked for the whole group while holding a real maximum: p == 0 exactly
        if (!zero && mode == plan::FAST && mre && __all_sync(0xffffffffu, masked && m2 > -1e8f)) zero = true;
        if (zero) {
#pragma unroll
          for (int x = 0; x < W / 2; ++x) ds_pk[x] = 0u;
        } else if (mode == plan::FAST) {
          uint32_t v[W], w[W];
          tmem_ld32(t_s, v);
          tmem_ld32(t_dp, w);
          tmem_wait_ld();
          const float cadd = relc + mterm;
          // a masked score is cadd itself (|x * scale| < 32 is absorbed by -1e9 in fp32)
          const float gmul = masked ? 0.f : scale2;
          const float gsub = fmaf(cadd, LOG2E, -m2l);
          float t0 = 0.f, t1 = 0.f;
#pragma unroll
          for (int x = 0; x < W / 2; ++x) {
            const float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), gmul, gsub));
            const float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), gmul, gsub));
            const float d0 = p0 * (__uint_as_float(w[2 * x]) - delta);
            const float d1 = p1 * (__uint_as_float(w[2 * x + 1]) - delta);
            t0 += d0;
            t1 += d1;
            ds_pk[x] = pack_bf16x2(d0, d1);
          }
          const float tot = t0 + t1;
          if (ccls == plan::C_POS) accP += tot;
          else if (ccls == plan::C_NEG) accN += tot;
          else if (ccls == plan::C_CROSS) accX += tot;
        } else {
          float ds[W];
          if (mode == plan::GEN) {
            // real loop, TMEM as dynamically indexed scratch: one copy of the generic code
#pragma unroll 1
            for (int jj = 0; jj < W; ++jj) {
              const uint32_t raw = tmem_ld1(t_s + jj);
              const uint32_t dpr = tmem_ld1(t_dp + jj);
              tmem_wait_ld();
              int slot;
              const float t = score_generic_q(__uint_as_float(raw), sc, rc, b, i, row, row_ok, g0 + jj,
                                              cp->ce[part * W + jj], cp->cs[part * W + jj], rel_s, a.scale, a.neg, slot);
              const float pv = ex2(fmaf(t, LOG2E, -m2l));   // dead: t = -inf -> 0
              const float dsv = (t == -INFINITY) ? 0.f : pv * (__uint_as_float(dpr) - delta);
              if (slot >= 0) bin[slot * TM + row] += dsv;
              __syncwarp();   // score_generic diverges per row; tcgen05.st needs the converged warp
              tmem_st1(t_s + jj, __float_as_uint(dsv));
            }
            tmem_wait_st();
            uint32_t v[W];
            tmem_ld32(t_s, v);
            tmem_wait_ld();
#pragma unroll
            for (int x = 0; x < W; ++x) ds[x] = __uint_as_float(v[x]);
          } else {
            float t[W];
            uint32_t w[W];
            {
              uint32_t v[W];
              tmem_ld32(t_s, v);
              tmem_ld32(t_dp, w);
              tmem_wait_ld();
#pragma unroll
              for (int x = 0; x < W; ++x) t[x] = __uint_as_float(v[x]);
            }
            switch (mode) {
              case plan::EDGE: {
                const int d0 = g0 - i;
                int jlo = 0, jhi = min(W, sc.ke - g0);
                if (sc.band) {
                  jlo = max(jlo, -sc.radius - d0);
                  jhi = min(jhi, sc.radius - d0 + 1);
                }
                const unsigned span = (unsigned)max(jhi - jlo, 0);
                const float cadd = relc + mterm;
#pragma unroll
                for (int jj = 0; jj < W; ++jj) {
                  const float v = fmaf(t[jj], a.scale, cadd);
                  t[jj] = ((unsigned)(jj - jlo) < span) ? v : -INFINITY;
                }
                break;
              }
              case plan::DIAG: {
                const int d0 = g0 - i + sc.D;
                const float* base = rel_s + row;
#pragma unroll
                for (int jj = 0; jj < W; ++jj) {
                  const int sl = min(max(d0 + jj, 0), 2 * sc.D);
                  t[jj] = fmaf(t[jj], a.scale, base[sl * TM] + mterm);
                }
                break;
              }
              case plan::QS: {
                const int d0 = rc.q_sent - g0;
                const float c0 = rc.relX + mterm, c1 = rc.relX1 + mterm;
#pragma unroll
                for (int jj = 0; jj < W; ++jj) t[jj] = fmaf(t[jj], a.scale, d0 == jj ? c1 : c0);
                break;
              }
              default: {   // KS
                const float c0 = rc.relX + mterm, c1 = rc.relX1 + mterm;
#pragma unroll
                for (int jj = 0; jj < W; ++jj) t[jj] = fmaf(t[jj], a.scale, cp->cs[part * W + jj] == i ? c1 : c0);
                break;
              }
            }
            if (mask_pe) {
#pragma unroll
              for (int jj = 0; jj < W; ++jj) t[jj] += (cp->ce[part * W + jj] == rc.q_e) ? 0.f : a.neg;
            }
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int x = 0; x < W; x += 2) {
              const float p0 = ex2(fmaf(t[x], LOG2E, -m2l));   // dead: t = -inf -> 0
              const float p1 = ex2(fmaf(t[x + 1], LOG2E, -m2l));
              ds[x] = p0 * (__uint_as_float(w[x]) - delta);
              ds[x + 1] = p1 * (__uint_as_float(w[x + 1]) - delta);
              t0 += ds[x];
              t1 += ds[x + 1];
            }
            const float tot = t0 + t1;
            // ---- relative-id bins ----
            if (mode == plan::EDGE) {
              if (ccls == plan::C_POS) accP += tot;
              else if (ccls == plan::C_NEG) accN += tot;
              else if (ccls == plan::C_CROSS) accX += tot;
            } else if (mode == plan::DIAG) {
              const int d0 = g0 - i + sc.D;
              float* base = bin + row;
#pragma unroll
              for (int x = 0; x < W; ++x) {
                const int sl = min(max(d0 + x, 0), 2 * sc.D);
                base[sl * TM] += ds[x];
              }
            } else if (mode == plan::QS) {
              const int d0 = rc.q_sent - g0;
              float sp = 0.f;
#pragma unroll
              for (int x = 0; x < W; ++x) sp += (d0 == x) ? ds[x] : 0.f;
              accX1 += sp;
              accX += tot - sp;
            } else {  // KS
              float sp = 0.f;
#pragma unroll
              for (int x = 0; x < W; ++x) sp += (cp->cs[part * W + x] == i) ? ds[x] : 0.f;
              accX1 += sp;
              accX += tot - sp;
            }
          }
#pragma unroll
          for (int x = 0; x < W / 2; ++x) ds_pk[x] = pack_bf16x2(ds[2 * x], ds[2 * x + 1]);
        }
        // each part packs into its OWN column range (the other part may still be reading its inputs)
        tmem_st16(t_dp, ds_pk);
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->ds_full[c & 1]);
        if (tid == 0) TRACE(1, 10 + 3 * c);
      }
    };
    run_chunks(sc0, rc0, 0, r0.n, r0.kb);
    if (a.nseg > 1) run_chunks(sc1, rc1, r0.n, nchunks, r1.kb);
    // flush the constant-class accumulators into this part's bins
    if (R > 0) {
      auto flush = [&](int id, float v) {
        if (id >= 0 && id < R) bin[plan::slot_of_id(id, pd, perm) * TM + row] += v;
      };
      const int dd = sc0.D;   // both segments of a row set share max_distance
      flush(dd, accP);
      flush(2 * dd, accN);
      flush(2 * dd + 1, accX);
      flush(2 * dd + 2, accX1);
    }
    // ---- epilogue: dallrel (summed over parts) -> global + bf16 A-operand for dQ += dallrel.E ----
    if (tid == 0) TRACE(1, 5);
    named_bar_sync(1, NALL);  // all bin arrays complete
    if (rpad) {
      // every thread of the row packs a slice of 8 ids: sum over the NB private bin arrays,
      // publish dallrel (fp32, id order) and write the bf16 A-operand columns for dQ += dallrel.E
#pragma unroll 1
      for (int c0 = 8 * bidx; c0 < rpad; c0 += 8 * NB) {
        uint32_t pk[4];
        float w8[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const int pid = c0 + x;
          float w = 0.f;
          if (pid < R) {
            const int sl = plan::slot_of_id(pid, pd, perm);
#pragma unroll
            for (int pp = 0; pp < NB; ++pp) w += bins[(pp * RB + sl) * TM + row];
          }
          w8[x] = w;
        }
        if (a.tg_partial) {
          // rows beyond the sequence end contribute nothing (their bins may hold garbage-free zeros,
          // but Q rows are zero-filled by TMA anyway)
          const uint4 pk4 = make_uint4(pack_bf16x2(w8[0], w8[1]), pack_bf16x2(w8[2], w8[3]),
                                       pack_bf16x2(w8[4], w8[5]), pack_bf16x2(w8[6], w8[7]));
          *reinterpret_cast<uint4*>(smem + SM_A + row * 128 + ((((c0 >> 3)) ^ (row & 7)) << 4)) = pk4;
          // bias partial: sum over the 32 rows of this warp, one value per id
          float* bs = reinterpret_cast<float*>(smem + SM_BS) + quad * 64 + c0;
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            float v = row_ok ? w8[x] : 0.f;
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (lane == 0) bs[x] = v;
          }
        } else if (row_ok) {
          float* dst = a.dallrel + srow * R + c0;
          if ((R & 3) == 0 && c0 + 8 <= R) {
            *reinterpret_cast<float4*>(dst) = make_float4(w8[0], w8[1], w8[2], w8[3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(w8[4], w8[5], w8[6], w8[7]);
          } else {
#pragma unroll
            for (int x = 0; x < 8; ++x)
              if (c0 + x < R) dst[x] = w8[x];
          }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) pk[x] = pack_bf16x2(w8[2 * x], w8[2 * x + 1]);
        // 4 packed columns; tcgen05.st needs the whole warp: the loop bounds are warp-uniform
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tmem + T_REL + lane_sel + c0 / 2),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                     : "memory");
      }
      tmem_wait_st();
      fence_proxy_async_smem();   // the dallrel^T tile was written through the generic proxy
      tc_fence_before_sync();
      mbar_arrive(&bars->dar_full);
    }
    if (tid == 0) TRACE(1, 6);
    mbar_wait_warp(&bars->dq_full, 0);
    if (tid == 0) TRACE(1, 7);
    tc_fence_after_sync();
    if (rpad && a.tg_partial) {
      const int64_t pidx = ((int64_t)(b * gridDim.x + blockIdx.x) * a.H + h);
      if (bidx == 0) {   // warp-uniform: four warps, lanes with (lane % 32) < 16 hold the 64 id rows
        const int pid = quad * 16 + lane;
#pragma unroll 1
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(tmem + T_DE + lane_sel + hh * 32, v);
          tmem_wait_ld();
          if (lane < 16 && pid < R) {
            float4* dst = reinterpret_cast<float4*>(a.tg_partial + (pidx * R + pid) * 64 + hh * 32);
#pragma unroll
            for (int x = 0; x < 8; ++x)
              dst[x] = make_float4(__uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]),
                                   __uint_as_float(v[4 * x + 2]), __uint_as_float(v[4 * x + 3]));
          }
        }
      }
      named_bar_sync(1, NALL);   // bias sums of all four row quadrants are in smem
      if (tid < R) {
        const float* bs = reinterpret_cast<const float*>(smem + SM_BS);
        a.tg_partial_bias[pidx * R + tid] = (bs[tid] + bs[64 + tid]) + (bs[128 + tid] + bs[192 + tid]);
      }
    }
    constexpr int WO = 64 / NB;   // output columns per thread
    uint32_t dq_raw[WO];
    tmem_ldN(tmem + T_DQ + lane_sel + bidx * WO, dq_raw);
    tmem_wait_ld();
    float dq[WO];
#pragma unroll
    for (int x = 0; x < WO; ++x) dq[x] = __uint_as_float(dq_raw[x]);
    if (row_ok) {
      __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(a.d_q, b, i, h) + bidx * WO;
#pragma unroll
      for (int x = 0; x < WO / 8; ++x) {
        uint4 w;
        w.x = pack_bf16x2(dq[8 * x + 0] * a.scale, dq[8 * x + 1] * a.scale);
        w.y = pack_bf16x2(dq[8 * x + 2] * a.scale, dq[8 * x + 3] * a.scale);
        w.z = pack_bf16x2(dq[8 * x + 4] * a.scale, dq[8 * x + 5] * a.scale);
        w.w = pack_bf16x2(dq[8 * x + 6] * a.scale, dq[8 * x + 7] * a.scale);
        *reinterpret_cast<uint4*>(dst + 8 * x) = w;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WM) tmem_dealloc<TMEM_COLS>(tmem);
}

// ============================================================================================
// Key-centric pass
// ============================================================================================
namespace bk {
constexpr int NST = 3;
constexpr int SM_K = 0;                         // 16 KB
constexpr int SM_V = SM_K + TM * 128;           // 16 KB
constexpr int SM_QD = SM_V + TM * 128;          // NST x (Q 8 KB + dO 8 KB)
constexpr int SM_RS = SM_QD + NST * 2 * TN * 128;       // NST x [64] float4 rowstat
constexpr int SM_RELQ = SM_RS + NST * TN * 16;          // NST x [64][64] f32 allrel rows
constexpr int SM_BAR = SM_RELQ + NST * TN * 64 * 4;
constexpr int SM_ALLOC = SM_BAR + 256 + 1024;
constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320;

struct Bars {
  uint64_t kv_full;
  uint64_t qd_full[NST], qd_empty[NST];
  uint64_t sdp_full[2], pds_full[2], acc_full;
  uint32_t tmem_base;
};
}  // namespace bk

struct TcQuerySource {
  QuerySource q;
  const float4* rowstat;  // [B, H, lp]
  const float* allrel_ws; // [B, H, lp, rw]
  int lp, rw;
};

struct TcBwdKVParams {
  T4 k, v, d_k, d_v;
  int len;
  TcQuerySource src[2];
  int nsrc;
  int B, H;
  float scale, neg;
};

struct SrcRange {
  int ib, ie, n;
};
__device__ __forceinline__ SrcRange src_range(const TcQuerySource& s, int j0) {
  SrcRange r;
  const int lq = s.q.rows.len;
  if (s.q.band) {
    r.ib = max(0, j0 - s.q.radius) & ~3;   // 16-byte aligned bulk copies of the row records
    r.ie = min(lq, j0 + TM + s.q.radius);
  } else {
    r.ib = 0;
    r.ie = lq;
  }
  r.n = (r.ie - r.ib + TN - 1) / TN;
  return r;
}

// ---- column-centric scoring: the thread owns KEY row j, the 32 columns of a group are queries.
// Same warp-uniform classification idea as tc_rowscore.cuh with the roles swapped: query-side
// scalars are lane-held, the relative term comes from the staged per-query rows relq[ii][id].
namespace colscore {
using rowscore::GMode;
using rowscore::GM_DEAD; using rowscore::GM_FAST; using rowscore::GM_EDGE; using rowscore::GM_DIAG;
using rowscore::GM_QS; using rowscore::GM_KS; using rowscore::GM_GEN;

struct SrcCtx {          // warp-uniform, one per query source
  const TcQuerySource* src;
  int ib, ie;            // query range of this key tile
  int R, D, rw;
  bool band;
  int radius;
  int mask_rule, id_rule;
};
struct KeyCtx {          // per thread
  int j, row;
  bool key_ok;
  int k_e, k_sent;
};
struct QLanes {          // lane l holds the scalars of query (g0 + l)
  int qe_l, qs_l;
};
struct Plan {
  int mode;
  int cid;               // id of the constant relative class (-1: none)
  float mrow;
  bool mask_pe;
};

__device__ __forceinline__ QLanes load_q_lanes(const SrcCtx& sc, int b, int g0, int lane) {
  QLanes ql{0, -1};
  const Side& sd = sc.src->q.side;
  const int i = g0 + lane;
  if (i >= 0 && i < sc.src->q.rows.len) {
    if (sc.mask_rule == MR_EXAMPLE_ID) ql.qe_l = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
    if (sc.id_rule == IDR_CROSS_QSENT) ql.qs_l = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
  }
  return ql;
}

__device__ __forceinline__ void init_key(KeyCtx& kc, const SrcCtx& sc, int b) {
  const Side& sd = sc.src->q.side;
  kc.k_e = 0;
  kc.k_sent = -1;
  if (kc.key_ok && sc.mask_rule == MR_EXAMPLE_ID) kc.k_e = __ldg(sd.k_eid + (int64_t)b * sd.k_len + kc.j);
  if (kc.key_ok && sc.id_rule == IDR_CROSS_KSENT) kc.k_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + kc.j);
}

// a = first key row of the warp, g0 = first query of the group.  off = key - query.
template <int W>
__device__ __forceinline__ Plan classify(const SrcCtx& sc, const KeyCtx& kc, const QLanes& ql, int a, int g0,
                                         int lane, float neg, int sub) {
  Plan pl;
  pl.cid = -1;
  pl.mrow = 0.f;
  pl.mask_pe = false;
  const int o_min = a - (g0 + W - 1), o_max = a + 31 - g0;
  const bool dead = g0 >= sc.ie || (sc.band && (o_min > sc.radius || o_max < -sc.radius));
  if (dead) {
    pl.mode = GM_DEAD;
    return pl;
  }
  const bool all_live = (g0 + W - 1 < sc.ie) && (!sc.band || (o_min >= -sc.radius && o_max <= sc.radius));
  bool gen = false;
  if (sc.mask_rule == MR_EXPLICIT) {
    gen = true;
  } else if (sc.mask_rule == MR_EXAMPLE_ID) {
    const int qe0 = __shfl_sync(0xffffffffu, ql.qe_l, sub);
    const bool lane_oob = (g0 - sub + lane >= sc.ie) || (W < 32 && (lane < sub || lane >= sub + W));
    const bool uni = __all_sync(0xffffffffu, lane_oob || ql.qe_l == qe0);
    pl.mask_pe = !uni;
    pl.mrow = uni ? ((kc.k_e == qe0) ? 0.f : neg) : 0.f;
  }
  int rcls = 0;  // 0 const, 1 diag, 2 qs, 3 ks, 4 generic
  switch (sc.id_rule) {
    case IDR_NONE:
      break;
    case IDR_1D:
      if (2 * sc.D + 1 > sc.R) rcls = 4;
      else if (o_min >= sc.D) pl.cid = sc.D;
      else if (o_max <= -sc.D) pl.cid = 2 * sc.D;
      else rcls = 1;
      break;
    case IDR_CROSS_QSENT:
      if (2 * sc.D + 2 >= sc.R) { rcls = 4; break; }
      pl.cid = 2 * sc.D + 1;
      if (__any_sync(0xffffffffu, ql.qs_l >= a && ql.qs_l <= a + 31 && (W == 32 || (lane >= sub && lane < sub + W)))) rcls = 2;
      break;
    case IDR_CROSS_KSENT:
      if (2 * sc.D + 2 >= sc.R) { rcls = 4; break; }
      pl.cid = 2 * sc.D + 1;
      if (__any_sync(0xffffffffu, kc.k_sent >= g0 && kc.k_sent < g0 + W)) rcls = 3;
      break;
    default:
      rcls = 4;
  }
  if (gen || rcls == 4) {
    pl.mode = GM_GEN;
    pl.mask_pe = false;
  } else if (rcls == 0) {
    pl.mode = all_live ? GM_FAST : GM_EDGE;
  } else if (!all_live) {
    pl.mode = GM_GEN;
    pl.mask_pe = false;
  } else {
    pl.mode = rcls == 1 ? GM_DIAG : (rcls == 2 ? GM_QS : GM_KS);
  }
  return pl;
}

// Generic per-element evaluation; returns the score or -inf when (i, j) is dead.
__device__ __forceinline__ float score_generic(float x, const SrcCtx& sc, const KeyCtx& kc, const QLanes& ql,
                                               int b, int g0, int ii, const float* relq, float scale, float neg,
                                               int sub) {
  const Side& sd = sc.src->q.side;
  const int i = g0 + ii;
  const int off = kc.j - i;
  const int qe_i = __shfl_sync(0xffffffffu, ql.qe_l, sub + ii);
  const int qs_i = __shfl_sync(0xffffffffu, ql.qs_l, sub + ii);
  const bool live = kc.key_ok && i < sc.ie && (!sc.band || (off <= sc.radius && off >= -sc.radius));
  if (!live) return -INFINITY;
  const int col = sc.band ? off + sc.radius : kc.j;
  bool ok = true;
  int id = -1;
  switch (sc.mask_rule) {
    case MR_EXPLICIT: ok = __ldg(sd.mask + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col) != 0; break;
    case MR_EXAMPLE_ID: ok = (qe_i == kc.k_e); break;
    default: break;
  }
  switch (sc.id_rule) {
    case IDR_EXPLICIT: id = __ldg(sd.ids + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col); break;
    case IDR_1D: id = rel_id_1d(off, sc.D); break;
    case IDR_CROSS_QSENT: id = 2 * sc.D + 1 + (qs_i == kc.j ? 1 : 0); break;
    case IDR_CROSS_KSENT: id = 2 * sc.D + 1 + (kc.k_sent == i ? 1 : 0); break;
    case IDR_2D: id = rel_id_2d(i, kc.j, sd.npr, sd.core, sc.D); break;
    default: break;
  }
  const float rel = (id >= 0 && id < sc.R) ? relq[ii * sc.rw + id] : 0.f;
  float v = fmaf(x, scale, rel);
  if (!ok) v += neg;
  return v;
}

// Scores of one 32-query group in place.  `mode` is warp-uniform; GM_GEN is handled elsewhere.
template <int W>
__device__ __forceinline__ void score_group(float (&t)[W], const Plan& pl, const SrcCtx& sc, const KeyCtx& kc,
                                            const QLanes& ql, int g0, const float* relq, float scale, float neg,
                                            int sub) {
  const int rw = sc.rw;
  switch (pl.mode) {
    case GM_FAST:
      if (pl.cid >= 0) {
#pragma unroll
        for (int ii = 0; ii < W; ++ii) t[ii] = fmaf(t[ii], scale, relq[ii * rw + pl.cid] + pl.mrow);
      } else {
#pragma unroll
        for (int ii = 0; ii < W; ++ii) t[ii] = fmaf(t[ii], scale, pl.mrow);
      }
      break;
    case GM_EDGE: {
      const int d0 = kc.j - g0;  // off = d0 - ii
      int ilo = 0, ihi = min(W, sc.ie - g0);
      if (sc.band) {
        ilo = max(ilo, d0 - sc.radius);
        ihi = min(ihi, d0 + sc.radius + 1);
      }
      if (!kc.key_ok) ihi = ilo;
      const unsigned span = (unsigned)max(ihi - ilo, 0);
      const int cid = pl.cid >= 0 ? pl.cid : 0;
      const float use = pl.cid >= 0 ? 1.f : 0.f;
#pragma unroll
      for (int ii = 0; ii < W; ++ii) {
        const float v = fmaf(t[ii], scale, fmaf(use, relq[ii * rw + cid], pl.mrow));
        t[ii] = ((unsigned)(ii - ilo) < span) ? v : -INFINITY;
      }
      break;
    }
    case GM_DIAG: {
      const int d0 = kc.j - g0;
#pragma unroll
      for (int ii = 0; ii < W; ++ii) {
        const int o = min(max(d0 - ii, -sc.D), sc.D);
        const int id = o >= 0 ? o : sc.D - o;
        t[ii] = fmaf(t[ii], scale, relq[ii * rw + id] + pl.mrow);
      }
      break;
    }
    case GM_QS: {
#pragma unroll
      for (int ii = 0; ii < W; ++ii) {
        const int qs_i = __shfl_sync(0xffffffffu, ql.qs_l, sub + ii);
        t[ii] = fmaf(t[ii], scale, relq[ii * rw + pl.cid + (qs_i == kc.j ? 1 : 0)] + pl.mrow);
      }
      break;
    }
    case GM_KS: {
      const int sp = kc.k_sent - g0;
#pragma unroll
      for (int ii = 0; ii < W; ++ii)
        t[ii] = fmaf(t[ii], scale, relq[ii * rw + pl.cid + (sp == ii ? 1 : 0)] + pl.mrow);
      break;
    }
    default:
      return;
  }
  if (pl.mask_pe) {
#pragma unroll
    for (int ii = 0; ii < W; ++ii) {
      const int qe_i = __shfl_sync(0xffffffffu, ql.qe_l, sub + ii);
      t[ii] += (qe_i == kc.k_e) ? 0.f : neg;
    }
  }
}

// GM_GEN: real loop, TMEM as scratch.  Leaves p (fp32) in the S^T column and ds in the dP^T column.
template <int W>
__device__ __forceinline__ void group_generic_tmem(uint32_t t_s, uint32_t t_dp, const SrcCtx& sc, const KeyCtx& kc,
                                                   const QLanes& ql, int b, int g0, const float4* rs,
                                                   const float* relq, float scale, float neg, int sub) {
#pragma unroll 1
  for (int ii = 0; ii < W; ++ii) {
    const uint32_t raw = tmem_ld1(t_s + ii);
    const uint32_t dpr = tmem_ld1(t_dp + ii);
    tmem_wait_ld();
    const float t = score_generic(__uint_as_float(raw), sc, kc, ql, b, g0, ii, relq, scale, neg, sub);
    float pv = 0.f, ds = 0.f;
    if (t != -INFINITY) {
      const float4 st = rs[ii];
      pv = ex2(fmaf(t, LOG2E, -st.x)) * st.y;
      ds = pv * (__uint_as_float(dpr) - st.z);
    }
    tmem_st1(t_s + ii, __float_as_uint(pv));
    tmem_st1(t_dp + ii, __float_as_uint(ds));
  }
  tmem_wait_st();
}
}  // namespace colscore

// NP threads per row inside a warp set; SETS warp sets take alternate chunks (S^T / dP^T are
// double-buffered by chunk parity, so set s simply owns buffer s).
template <int NP, int SETS>
__global__ void __launch_bounds__(nthreads<NP * SETS>(), 1)
tc_bwd_kv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_do0,
                 const __grid_constant__ CUtensorMap map_q1, const __grid_constant__ CUtensorMap map_do1,
                 const TcBwdKVParams p) {
  using namespace bk;
  constexpr int W = 64 / NP;
  constexpr int NEW = 128 * NP;                  // elementwise threads per set
  constexpr int WP = 4 * NP * SETS, WM = 4 * NP * SETS + 1;
  static_assert(SETS == 1 || SETS == 2, "chunk buffers are double-buffered");
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: keeps the shared address space
  // (LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TM;

  if (tid == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->qd_full[s], 1);
      mbar_init(&bars->qd_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sdp_full[s], 1);
      mbar_init(&bars->pds_full[s], NEW);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == WM) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SrcRange r0 = src_range(p.src[0], j0);
  SrcRange r1{0, 0, 0};
  if (p.nsrc > 1) r1 = src_range(p.src[1], j0);
  const int nchunks = r0.n + r1.n;

  if (warp == WP) {
    if (elect_one()) {
      mbar_arrive_expect_tx(&bars->kv_full, 2 * TM * 128);
      tma_load_4d(smem + SM_K, &map_k, &bars->kv_full, 0, j0, h, b);
      tma_load_4d(smem + SM_V, &map_v, &bars->kv_full, 0, j0, h, b);
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % NST;
        mbar_wait(&bars->qd_empty[st], ((c / NST) & 1) ^ 1);
        const bool first = c < r0.n;
        const TcQuerySource& src = first ? p.src[0] : p.src[1];
        const int q0 = first ? r0.ib + c * TN : r1.ib + (c - r0.n) * TN;
        const int rw = src.rw;
        uint8_t* qs = smem + SM_QD + st * (2 * TN * 128);
        const int64_t prow = (int64_t)(b * p.H + h) * src.lp + q0;
        const uint32_t rel_bytes = rw > 0 ? TN * rw * 4 : 0;
        mbar_arrive_expect_tx(&bars->qd_full[st], 2 * TN * 128 + TN * 16 + rel_bytes);
        tma_load_4d(qs, first ? &map_q0 : &map_q1, &bars->qd_full[st], 0, q0, h, b);
        tma_load_4d(qs + TN * 128, first ? &map_do0 : &map_do1, &bars->qd_full[st], 0, q0, h, b);
        bulk_g2s(smem + SM_RS + st * TN * 16, src.rowstat + prow, TN * 16, &bars->qd_full[st]);
        if (rel_bytes)
          bulk_g2s(smem + SM_RELQ + st * TN * 64 * 4, src.allrel_ws + prow * rw, rel_bytes, &bars->qd_full[st]);
      }
    }
  } else if (warp == WM) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t k_addr = smem_u32(smem + SM_K), v_addr = smem_u32(smem + SM_V);
      mbar_wait(&bars->kv_full, 0);
      tc_fence_after_sync();
      for (int c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
          const int st = c % NST;
          mbar_wait(&bars->qd_full[st], (c / NST) & 1);
          TRACE(2, 4 * c);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(smem + SM_QD + st * (2 * TN * 128));
          const uint32_t do_addr = q_addr + TN * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // S^T = K . Q_c^T
            umma_ss(tmem + T_S + (c & 1) * 64, make_smem_desc_sw128(k_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(q_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dP^T = V . dO_c^T
            umma_ss(tmem + T_DP + (c & 1) * 64, make_smem_desc_sw128(v_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(do_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&bars->sdp_full[c & 1]);
          TRACE(2, 4 * c + 1);
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->pds_full[pc & 1], (pc >> 1) & 1);
          TRACE(2, 4 * pc + 2);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(smem + SM_QD + st * (2 * TN * 128));
          const uint32_t do_addr = q_addr + TN * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dV += P^T . dO_c
            umma_ts(tmem + T_DV, tmem + T_S + (pc & 1) * 64 + ((16 * kk) / W) * W + ((16 * kk) % W) / 2,
                    make_smem_desc_sw128(do_addr + kk * 2048, 16, 1024), idesc_acc, (pc > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dK += dS^T . Q_c
            umma_ts(tmem + T_DK, tmem + T_DP + (pc & 1) * 64 + ((16 * kk) / W) * W + ((16 * kk) % W) / 2,
                    make_smem_desc_sw128(q_addr + kk * 2048, 16, 1024), idesc_acc, (pc > 0 || kk > 0));
          umma_commit(&bars->qd_empty[st]);
          TRACE(2, 4 * pc + 3);
          if (pc == nchunks - 1) umma_commit(&bars->acc_full);
        }
      }
    }
  } else {
    using namespace colscore;
    const int row = (warp & 3) * 32 + lane;
    const int set = warp / (4 * NP);            // which chunk parity this warp serves
    const int part = (warp >> 2) % NP;
    const int win = (part * W) / 32, sub = (part * W) % 32;
    const int j = j0 + row;
    const bool key_ok = j < p.len;
    const int wrow0 = j0 + (warp & 3) * 32;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    auto make_src = [&](const TcQuerySource& src, const SrcRange& r) {
      SrcCtx sc;
      sc.src = &src;
      sc.ib = r.ib;
      sc.ie = r.ie;
      sc.R = src.q.rows.R;
      sc.D = src.q.side.max_distance;
      sc.rw = src.rw;
      sc.band = src.q.band != 0;
      sc.radius = src.q.radius;
      sc.mask_rule = src.q.side.mask_rule;
      sc.id_rule = sc.R > 0 ? src.q.side.id_rule : IDR_NONE;
      return sc;
    };
    // One call per query source (inlined twice: no per-field selects inside the chunk loop).
    // Chunks [c_begin, c_end) belong to this source; this warp set handles those with c % SETS == set.
    auto run_chunks = [&](const SrcCtx sc, const KeyCtx kc, int c_begin, int c_end, int ib) {
      int c = c_begin + ((set - c_begin) % SETS + SETS) % SETS;
      if (c >= c_end) return;
      QLanes ql = load_q_lanes(sc, b, ib + (c - c_begin) * TN + 32 * win, lane);
#pragma unroll 1
      for (; c < c_end; c += SETS) {
        const int st = c % NST;
        const int g0 = ib + (c - c_begin) * TN + part * W;
        const uint32_t t_s = tmem + T_S + (c & 1) * 64 + lane_sel + part * W;
        const uint32_t t_dp = tmem + T_DP + (c & 1) * 64 + lane_sel + part * W;
        QLanes nl{0, -1};
        if (c + SETS < c_end) nl = load_q_lanes(sc, b, ib + (c + SETS - c_begin) * TN + 32 * win, lane);
        const Plan pl = classify<W>(sc, kc, ql, wrow0, g0, lane, p.neg, sub);
        if (tid == 0) TRACE(0, 4 * c);
        mbar_wait_warp(&bars->qd_full[st], (c / NST) & 1);   // rowstat / allrel rows of this chunk
        mbar_wait_warp(&bars->sdp_full[c & 1], (c >> 1) & 1);
        if (tid == 0) TRACE(0, 4 * c + 1);
        tc_fence_after_sync();
        const float4* rs = reinterpret_cast<const float4*>(smem + SM_RS + st * TN * 16) + part * W;
        const float* relq = reinterpret_cast<const float*>(smem + SM_RELQ + st * TN * 64 * 4) + part * W * sc.rw;
        uint32_t p_pk[W / 2], ds_pk[W / 2];
        if (pl.mode == GM_DEAD) {
#pragma unroll
          for (int x = 0; x < W / 2; ++x) { p_pk[x] = 0u; ds_pk[x] = 0u; }
        } else if (pl.mode == GM_GEN) {
          group_generic_tmem<W>(t_s, t_dp, sc, kc, ql, b, g0, rs, relq, p.scale, p.neg, sub);
          uint32_t v[W];
          tmem_ldN(t_s, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < W / 2; ++x) p_pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
          tmem_ldN(t_dp, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < W / 2; ++x) ds_pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
        } else {
          float t[W];
          uint32_t v[W];
          tmem_ldN(t_s, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < W; ++x) t[x] = __uint_as_float(v[x]);
          score_group<W>(t, pl, sc, kc, ql, g0, relq, p.scale, p.neg, sub);
          tmem_ldN(t_dp, v);
          tmem_wait_ld();
          if (pl.mode == GM_EDGE) {   // dead columns may carry garbage row records: guarded variant
#pragma unroll
            for (int x = 0; x < W / 2; ++x) {
              float pv[2], dsv[2];
#pragma unroll
              for (int y = 0; y < 2; ++y) {
                const int ii = 2 * x + y;
                const float4 r4 = rs[ii];
                const float e = ex2(fmaf(t[ii], LOG2E, -r4.x)) * r4.y;
                const float d = e * (__uint_as_float(v[ii]) - r4.z);
                const bool dead = (t[ii] == -INFINITY);
                pv[y] = dead ? 0.f : e;
                dsv[y] = dead ? 0.f : d;
              }
              p_pk[x] = pack_bf16x2(pv[0], pv[1]);
              ds_pk[x] = pack_bf16x2(dsv[0], dsv[1]);
            }
          } else {
#pragma unroll
            for (int x = 0; x < W / 2; ++x) {
              float pv[2], dsv[2];
#pragma unroll
              for (int y = 0; y < 2; ++y) {
                const int ii = 2 * x + y;
                const float4 r4 = rs[ii];   // (m*log2e, 1/l, delta, -): warp-broadcast LDS.128
                pv[y] = ex2(fmaf(t[ii], LOG2E, -r4.x)) * r4.y;
                dsv[y] = pv[y] * (__uint_as_float(v[ii]) - r4.z);
              }
              p_pk[x] = pack_bf16x2(pv[0], pv[1]);
              ds_pk[x] = pack_bf16x2(dsv[0], dsv[1]);
            }
          }
        }
        if (tid == 0) TRACE(0, 4 * c + 2);
        tmem_stN(t_s, p_pk);
        tmem_stN(t_dp, ds_pk);
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->pds_full[c & 1]);
        if (tid == 0) TRACE(0, 4 * c + 3);
        ql = nl;
      }
    };
    {
      KeyCtx kc;
      kc.j = j; kc.row = row; kc.key_ok = key_ok;
      const SrcCtx sc0 = make_src(p.src[0], r0);
      init_key(kc, sc0, b);
      run_chunks(sc0, kc, 0, r0.n, r0.ib);
      if (p.nsrc > 1) {
        const SrcCtx sc1 = make_src(p.src[1], r1);
        init_key(kc, sc1, b);
        run_chunks(sc1, kc, r0.n, nchunks, r1.ib);
      }
    }
    mbar_wait_warp(&bars->acc_full, 0);
    tc_fence_after_sync();
    constexpr int WO = 64 / (NP * SETS);         // output columns per thread
    const int opart = set * NP + part;
    uint32_t dv_raw[WO], dk_raw[WO];
    tmem_ldN(tmem + T_DV + lane_sel + opart * WO, dv_raw);
    tmem_ldN(tmem + T_DK + lane_sel + opart * WO, dk_raw);
    tmem_wait_ld();
    if (key_ok) {
      __nv_bfloat16* dv = row_ptr_mut<__nv_bfloat16>(p.d_v, b, j, h) + opart * WO;
      __nv_bfloat16* dk = row_ptr_mut<__nv_bfloat16>(p.d_k, b, j, h) + opart * WO;
#pragma unroll
      for (int x = 0; x < WO / 8; ++x) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 0]), __uint_as_float(dv_raw[8 * x + 1]));
        w.y = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 2]), __uint_as_float(dv_raw[8 * x + 3]));
        w.z = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 4]), __uint_as_float(dv_raw[8 * x + 5]));
        w.w = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 6]), __uint_as_float(dv_raw[8 * x + 7]));
        *reinterpret_cast<uint4*>(dv + 8 * x) = w;
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(dk_raw[8 * x + 0]) * p.scale, __uint_as_float(dk_raw[8 * x + 1]) * p.scale);
        u.y = pack_bf16x2(__uint_as_float(dk_raw[8 * x + 2]) * p.scale, __uint_as_float(dk_raw[8 * x + 3]) * p.scale);
        u.z = pack_bf16x2(__uint_as_float(dk_raw[8 * x + 4]) * p.scale, __uint_as_float(dk_raw[8 * x + 5]) * p.scale);
        u.w = pack_bf16x2(__uint_as_float(dk_raw[8 * x + 6]) * p.scale, __uint_as_float(dk_raw[8 * x + 7]) * p.scale);
        *reinterpret_cast<uint4*>(dk + 8 * x) = u;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WM) tmem_dealloc<TMEM_COLS>(tmem);
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }
inline int pad_rows(int len) { return (len + TN - 1) / TN * TN + TN; }
inline int pad4(int r) { return (r + 3) / 4 * 4; }

}  // namespace

// Workspace of one row set for the tcgen05 backward (rowstat + allrel), bytes.
size_t tc_bwd_rows_ws_bytes(int B, int H, int len, int R) {
  const size_t rows = (size_t)B * H * pad_rows(len);
  return align256(rows * sizeof(float4)) + align256(rows * pad4(R > 0 ? R : 0) * sizeof(float) + 256);
}

bool tc_bwd_q_supported(const BwdQArgs& a, int dtype, int d) {
  FwdArgs f{};
  f.rows = a.rows;
  f.seg[0] = a.seg[0];
  f.seg[1] = a.seg[1];
  f.nseg = a.nseg;
  f.out = a.out;
  auto ok = [](const T4& t) {
    return t.ptr && t.sb % 8 == 0 && t.sl % 8 == 0 && t.sh % 8 == 0 && reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0;
  };
  return tc_fwd_args_supported(f, dtype, d) && ok(a.d_out) && ok(a.d_q);
}

static bool g_attr_q = false, g_attr_kv = false;

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace_b(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_b, sizeof(unsigned long long) * 3 * 256);
}
#endif

int tc_launch_bwd_q(const BwdQArgs& a, void* ws, cudaStream_t st) {
  if (!g_attr_q) {
    cudaError_t e = cudaFuncSetAttribute(tc_bwd_q_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bq::SM_ALLOC);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_bwd_q_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bq::SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    g_attr_q = true;
  }
  TcBwdQParams p;
  p.a = a;
  const int R = a.rows.R;
  p.rpad = R > 0 ? (R + 15) / 16 * 16 : 0;
  p.rw = pad4(R);
  p.lp = pad_rows(a.rows.len);
  char* w = reinterpret_cast<char*>(ws);
  const size_t rows = (size_t)a.B * a.H * p.lp;
  p.rowstat = reinterpret_cast<float4*>(w);
  p.allrel_ws = reinterpret_cast<float*>(w + align256(rows * sizeof(float4)));
  CUtensorMap mq, mdo, mk0, mv0, mk1, mv1, me;
  int e = 0;
  e |= make_qkv_tensor_map(&mq, a.rows.q.ptr, a.rows.q.sb, a.rows.q.sl, a.rows.q.sh, a.B, a.rows.len, a.H, TM);
  e |= make_qkv_tensor_map(&mdo, a.d_out.ptr, a.d_out.sb, a.d_out.sl, a.d_out.sh, a.B, a.rows.len, a.H, TM);
  e |= make_qkv_tensor_map(&mk0, a.seg[0].k.ptr, a.seg[0].k.sb, a.seg[0].k.sl, a.seg[0].k.sh, a.B, a.seg[0].len, a.H, TN);
  e |= make_qkv_tensor_map(&mv0, a.seg[0].v.ptr, a.seg[0].v.sb, a.seg[0].v.sl, a.seg[0].v.sh, a.B, a.seg[0].len, a.H, TN);
  const KeySeg& s1 = a.nseg > 1 ? a.seg[1] : a.seg[0];
  e |= make_qkv_tensor_map(&mk1, s1.k.ptr, s1.k.sb, s1.k.sl, s1.k.sh, a.B, s1.len, a.H, TN);
  e |= make_qkv_tensor_map(&mv1, s1.v.ptr, s1.v.sb, s1.v.sl, s1.v.sh, a.B, s1.len, a.H, TN);
  if (p.rpad) {
    e |= make_qkv_tensor_map(&me, a.rows.emb, (int64_t)R * a.H * 64, (int64_t)a.H * 64, 64, 1, R, a.H, p.rpad);
  } else {
    me = mq;
  }
  if (e) return MLT_ERR_UNSUPPORTED;
  {
    const int64_t threads = (int64_t)a.B * a.rows.len * a.H * 8;
    tc_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a.out, a.d_out, a.stats, p.rowstat, a.B, a.H,
                                                                       a.rows.len, p.lp);
    cudaError_t pe = cudaGetLastError();
    if (pe != cudaSuccess) return (int)pe;
  }
  dim3 grid((a.rows.len + TM - 1) / TM, a.H, a.B);
  // chunks per tile: with many chunks (dense global rows) two warp sets on alternate chunks win;
  // with few (long rows: band + G/64) the extra per-tile prologue work does not pay off.
  auto seg_chunks = [](const KeySeg& sg) { return sg.band ? (TM + 2 * sg.radius + TN - 1) / TN : (sg.len + TN - 1) / TN; };
  const int est_chunks = seg_chunks(a.seg[0]) + (a.nseg > 1 ? seg_chunks(a.seg[1]) : 0);
  static const int force_sets = getenv("MLT_BWD_SETS") ? atoi(getenv("MLT_BWD_SETS")) : 0;
  const bool two_sets = force_sets ? force_sets == 2 : est_chunks >= 16;
  if (R <= 32 && two_sets)   // 4 private bin arrays of 32 slots
    tc_bwd_q_kernel<2><<<grid, bq_threads<2>(), bq::SM_ALLOC, st>>>(mq, mdo, mk0, mv0, mk1, mv1, me, p);
  else
    tc_bwd_q_kernel<1><<<grid, bq_threads<1>(), bq::SM_ALLOC, st>>>(mq, mdo, mk0, mv0, mk1, mv1, me, p);
  return (int)cudaGetLastError();
}

int tc_launch_bwd_kv(const BwdKVArgs& a, void* const ws[2], cudaStream_t st) {
  if (!g_attr_kv) {
    cudaError_t e = cudaFuncSetAttribute(tc_bwd_kv_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bk::SM_ALLOC);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(tc_bwd_kv_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bk::SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    g_attr_kv = true;
  }
  TcBwdKVParams p{};
  p.k = a.k; p.v = a.v; p.d_k = a.d_k; p.d_v = a.d_v;
  p.len = a.len;
  p.nsrc = a.nsrc;
  p.B = a.B; p.H = a.H; p.scale = a.scale; p.neg = a.neg;
  CUtensorMap mk, mv, mq[2], mdo[2];
  int e = 0;
  e |= make_qkv_tensor_map(&mk, a.k.ptr, a.k.sb, a.k.sl, a.k.sh, a.B, a.len, a.H, TM);
  e |= make_qkv_tensor_map(&mv, a.v.ptr, a.v.sb, a.v.sl, a.v.sh, a.B, a.len, a.H, TM);
  for (int s = 0; s < 2; ++s) {
    const QuerySource& q = a.src[s < a.nsrc ? s : 0];
    const int lq = q.rows.len;
    TcQuerySource& t = p.src[s];
    t.q = q;
    t.lp = pad_rows(lq);
    t.rw = pad4(q.rows.R);
    char* w = reinterpret_cast<char*>(ws[s < a.nsrc ? s : 0]);
    const size_t rows = (size_t)a.B * a.H * t.lp;
    t.rowstat = reinterpret_cast<const float4*>(w);
    t.allrel_ws = reinterpret_cast<const float*>(w + align256(rows * sizeof(float4)));
    e |= make_qkv_tensor_map(&mq[s], q.rows.q.ptr, q.rows.q.sb, q.rows.q.sl, q.rows.q.sh, a.B, lq, a.H, TN);
    e |= make_qkv_tensor_map(&mdo[s], q.d_out.ptr, q.d_out.sb, q.d_out.sl, q.d_out.sh, a.B, lq, a.H, TN);
  }
  if (e) return MLT_ERR_UNSUPPORTED;
  dim3 grid((a.len + TM - 1) / TM, a.H, a.B);
  auto src_chunks = [](const QuerySource& q) { return q.band ? (TM + 2 * q.radius + TN - 1) / TN : (q.rows.len + TN - 1) / TN; };
  const int est_chunks = src_chunks(a.src[0]) + (a.nsrc > 1 ? src_chunks(a.src[1]) : 0);
  if (est_chunks >= 16)
    tc_bwd_kv_kernel<2, 2><<<grid, nthreads<4>(), bk::SM_ALLOC, st>>>(mk, mv, mq[0], mdo[0], mq[1], mdo[1], p);
  else
    tc_bwd_kv_kernel<4, 1><<<grid, nthreads<4>(), bk::SM_ALLOC, st>>>(mk, mv, mq[0], mdo[0], mq[1], mdo[1], p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
