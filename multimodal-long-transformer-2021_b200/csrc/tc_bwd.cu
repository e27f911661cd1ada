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

#include "mlt_common.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace {

using namespace ptx;

constexpr int TM = 128;
constexpr int TN = 64;
constexpr int NTHREADS = 320;
constexpr int NEW = 256;  // elementwise threads
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
constexpr int SM_BAR = SM_BIN + 2 * 64 * TM * 4;
constexpr int SM_ALLOC = SM_BAR + 256 + 1024;
// TMEM columns
constexpr uint32_t T_S = 0, T_DP = 128, T_DQ = 256, T_REL = 320;

struct Bars {
  uint64_t q_full, rel_full;
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t sdp_full[2], ds_full[2], dq_full;
  uint32_t tmem_base;
};
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

template <int MR, int IR>
__device__ __forceinline__ void bwd_q_chunk(uint32_t (&s_raw)[32], uint32_t (&dp_raw)[32], uint32_t (&ds_pk)[16],
                                            const KeySeg& sg, int b, int i, int row, bool row_ok, int key0,
                                            int ke, int R, const float* rel_s, float* bin, float scale,
                                            float neg, float m2, float linv, float delta) {
  const Side& sd = sg.side;
  int q_e = 0, q_sent = -1;
  if (MR == MR_EXAMPLE_ID && row_ok) q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
  if (IR == IDR_CROSS_QSENT && row_ok) q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    const int j = key0 + jj;
    const int off = j - i;
    const bool live = row_ok && j < ke && (!sg.band || (off <= sg.radius && off >= -sg.radius));
    float ds = 0.f;
    if (live) {
      int k_e = 0, k_sent = -1;
      if (MR == MR_EXAMPLE_ID) k_e = __ldg(sd.k_eid + (int64_t)b * sd.k_len + j);
      if (IR == IDR_CROSS_KSENT) k_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + j);
      bool ok;
      int id;
      side_ok_id<MR, IR>(sd, b, i, j, sg.band ? off + sg.radius : j, q_e, k_e, q_sent, k_sent, ok, id);
      const bool idv = IR != IDR_NONE && id >= 0 && id < R;
      float t = fmaf(__uint_as_float(s_raw[jj]), scale, idv ? rel_s[id * TM + row] : 0.f);
      if (!ok) t += neg;
      const float p = ex2(fmaf(t, LOG2E, -m2)) * linv;
      ds = p * (__uint_as_float(dp_raw[jj]) - delta);
      if (idv) bin[id * TM + row] += ds;
    }
    if (jj & 1) {
      ds_pk[jj >> 1] = pack_bf16x2(__uint_as_float(s_raw[jj - 1]), ds);
    } else {
      s_raw[jj] = __float_as_uint(ds);  // stash the even element until its odd partner is ready
    }
  }
}

template <int MR>
__device__ __forceinline__ void bwd_q_chunk_ir(uint32_t (&s_raw)[32], uint32_t (&dp_raw)[32], uint32_t (&ds_pk)[16],
                                               const KeySeg& sg, int b, int i, int row, bool row_ok, int key0,
                                               int ke, int R, const float* rel_s, float* bin, float scale,
                                               float neg, float m2, float linv, float delta) {
#define MLT_CALL(IRV) bwd_q_chunk<MR, IRV>(s_raw, dp_raw, ds_pk, sg, b, i, row, row_ok, key0, ke, R, rel_s, bin, scale, neg, m2, linv, delta)
  switch (sg.side.id_rule) {
    case IDR_EXPLICIT: MLT_CALL(IDR_EXPLICIT); break;
    case IDR_1D: MLT_CALL(IDR_1D); break;
    case IDR_CROSS_QSENT: MLT_CALL(IDR_CROSS_QSENT); break;
    case IDR_CROSS_KSENT: MLT_CALL(IDR_CROSS_KSENT); break;
    case IDR_2D: MLT_CALL(IDR_2D); break;
    default: MLT_CALL(IDR_NONE); break;
  }
#undef MLT_CALL
}

__global__ void __launch_bounds__(NTHREADS, 1)
tc_bwd_q_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_v0,
                const __grid_constant__ CUtensorMap map_k1, const __grid_constant__ CUtensorMap map_v1,
                const __grid_constant__ CUtensorMap map_e, const TcBwdQParams p) {
  using namespace bq;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + SM_REL);
  float* bins = reinterpret_cast<float*>(smem + SM_BIN);
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
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SegRange r0 = seg_range(a.seg[0], i0);
  SegRange r1{0, 0, 0};
  if (a.nseg > 1) r1 = seg_range(a.seg[1], i0);
  const int nchunks = r0.n + r1.n;

  if (warp == 8) {
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
  } else if (warp == 9) {
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
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + T_DQ, tmem + T_DP + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(k_addr + kk * 2048, 16, 1024), idesc_dq, (pc > 0 || kk > 0));
          umma_commit(&bars->kv_empty[st]);
          if (pc == nchunks - 1) umma_commit(&bars->dq_full);
        }
      }
    }
  } else {
    // ===================== elementwise warps 0-7 =====================
    const int row = (warp & 3) * 32 + lane;
    const int hh = warp >> 2;  // column half
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    float* bin = bins + hh * 64 * TM;
    for (int x = lane + 32 * (warp & 3) + 128 * 0; x < 64 * TM; x += 128) bin[x] = 0.f;  // each half zeroes its bins
    // row constants
    float m2 = 0.f, linv = 0.f, delta = 0.f;
    const int64_t srow = (int64_t)(b * a.H + h) * a.rows.len + i;
    const int64_t prow = (int64_t)(b * a.H + h) * p.lp + i;
    if (row_ok) {
      const uint4* go = reinterpret_cast<const uint4*>(row_ptr<__nv_bfloat16>(a.d_out, b, i, h));
      const uint4* oo = reinterpret_cast<const uint4*>(row_ptr<__nv_bfloat16>(a.out, b, i, h));
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        const uint4 g4 = __ldg(go + x), o4 = __ldg(oo + x);
        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w}, ow[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
        for (int y = 0; y < 4; ++y) {
          const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[y]));
          const float2 of = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ow[y]));
          delta = fmaf(gf.x, of.x, delta);
          delta = fmaf(gf.y, of.y, delta);
        }
      }
      const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats) + srow);
      m2 = st.x * LOG2E;
      linv = 1.f / st.y;
      if (hh == 0) p.rowstat[prow] = make_float4(m2, linv, delta, 0.f);
    }
    if (rpad) {
      mbar_wait_warp(&bars->rel_full, 0);
      tc_fence_after_sync();
      const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.rows.bias);
      for (int c0 = 0; c0 < rpad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + T_REL + lane_sel + c0, v);
        tmem_wait_ld();
        if (hh == 0) {
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int pid = c0 + x;
            if (pid < R) {
              const float val = (__uint_as_float(v[x]) + __bfloat162float(bias[pid * a.H + h])) * a.scale;
              rel_s[pid * TM + row] = val;
              if (row_ok) p.allrel_ws[prow * p.rw + pid] = val;
            }
          }
        }
      }
    }
    named_bar_sync(1, NEW);  // rel_s / bins visible to both halves

    for (int c = 0; c < nchunks; ++c) {
      const bool first = c < r0.n;
      const KeySeg& sg = first ? a.seg[0] : a.seg[1];
      const int key0 = (first ? r0.kb + c * TN : r1.kb + (c - r0.n) * TN) + hh * 32;
      const int ke = first ? r0.ke : r1.ke;
      mbar_wait_warp(&bars->sdp_full[c & 1], (c >> 1) & 1);
      tc_fence_after_sync();
      uint32_t s_raw[32], dp_raw[32], ds_pk[16];
      tmem_ld32(tmem + T_S + (c & 1) * 64 + lane_sel + hh * 32, s_raw);
      tmem_ld32(tmem + T_DP + (c & 1) * 64 + lane_sel + hh * 32, dp_raw);
      tmem_wait_ld();
#define MLT_CALL(MRV) bwd_q_chunk_ir<MRV>(s_raw, dp_raw, ds_pk, sg, b, i, row, row_ok, key0, ke, R, rel_s, bin, a.scale, a.neg, m2, linv, delta)
      switch (sg.side.mask_rule) {
        case MR_EXPLICIT: MLT_CALL(MR_EXPLICIT); break;
        case MR_EXAMPLE_ID: MLT_CALL(MR_EXAMPLE_ID); break;
        default: MLT_CALL(MR_NONE); break;
      }
#undef MLT_CALL
      // each half packs into its OWN column range (the other half may still be reading its inputs)
      tmem_st16(tmem + T_DP + (c & 1) * 64 + lane_sel + hh * 32, ds_pk);
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(&bars->ds_full[c & 1]);
    }
    // ---- epilogue: dq = scale * (dS.K + dallrel.E); publish dallrel ----
    named_bar_sync(1, NEW);  // both halves' bins complete
    mbar_wait_warp(&bars->dq_full, 0);
    tc_fence_after_sync();
    uint32_t dq_raw[32];
    tmem_ld32(tmem + T_DQ + lane_sel + hh * 32, dq_raw);
    tmem_wait_ld();
    float dq[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) dq[x] = __uint_as_float(dq_raw[x]);
    if (R > 0) {
      const float* bin0 = bins;
      const float* bin1 = bins + 64 * TM;
      for (int pid = 0; pid < R; ++pid) {
        const float w = bin0[pid * TM + row] + bin1[pid * TM + row];
        if (hh == 0 && row_ok) a.dallrel[srow * R + pid] = w;
        // E row pid, columns [32 hh, 32 hh + 32): 4 swizzled 16-byte chunks (warp-broadcast reads)
        const uint8_t* erow = smem + SM_E + pid * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 e4 = *reinterpret_cast<const uint4*>(erow + (((hh * 4 + ch) ^ (pid & 7)) << 4));
          const uint32_t ew[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
          for (int y = 0; y < 4; ++y) {
            const float2 ef = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ew[y]));
            dq[ch * 8 + 2 * y] = fmaf(w, ef.x, dq[ch * 8 + 2 * y]);
            dq[ch * 8 + 2 * y + 1] = fmaf(w, ef.y, dq[ch * 8 + 2 * y + 1]);
          }
        }
      }
    }
    if (row_ok) {
      __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(a.d_q, b, i, h) + hh * 32;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
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
  if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem);
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

template <int MR, int IR>
__device__ __forceinline__ void bwd_kv_chunk(uint32_t (&s_raw)[32], uint32_t (&dp_raw)[32], uint32_t (&p_pk)[16],
                                             uint32_t (&ds_pk)[16], const TcQuerySource& src, int b, int j,
                                             bool key_ok, int q0, int ie, const float4* rs, const float* relq,
                                             int rw, float scale, float neg) {
  const Side& sd = src.q.side;
  const int R = src.q.rows.R;
  int k_e = 0, k_sent = -1;
  if (MR == MR_EXAMPLE_ID && key_ok) k_e = __ldg(sd.k_eid + (int64_t)b * sd.k_len + j);
  if (IR == IDR_CROSS_KSENT && key_ok) k_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + j);
  float p_prev = 0.f, ds_prev = 0.f;
#pragma unroll
  for (int ii = 0; ii < 32; ++ii) {
    const int i = q0 + ii;
    const int off = j - i;
    const bool live = key_ok && i < ie && (!src.q.band || (off <= src.q.radius && off >= -src.q.radius));
    float pv = 0.f, ds = 0.f;
    if (live) {
      int q_e = 0, q_sent = -1;
      if (MR == MR_EXAMPLE_ID) q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
      if (IR == IDR_CROSS_QSENT) q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
      bool ok;
      int id;
      side_ok_id<MR, IR>(sd, b, i, j, src.q.band ? off + src.q.radius : j, q_e, k_e, q_sent, k_sent, ok, id);
      const bool idv = IR != IDR_NONE && id >= 0 && id < R;
      float t = fmaf(__uint_as_float(s_raw[ii]), scale, idv ? relq[ii * rw + id] : 0.f);
      if (!ok) t += neg;
      const float4 st = rs[ii];  // (m2, linv, delta, -)
      pv = ex2(fmaf(t, LOG2E, -st.x)) * st.y;
      ds = pv * (__uint_as_float(dp_raw[ii]) - st.z);
    }
    if (ii & 1) {
      p_pk[ii >> 1] = pack_bf16x2(p_prev, pv);
      ds_pk[ii >> 1] = pack_bf16x2(ds_prev, ds);
    } else {
      p_prev = pv;
      ds_prev = ds;
    }
  }
}

template <int MR>
__device__ __forceinline__ void bwd_kv_chunk_ir(uint32_t (&s_raw)[32], uint32_t (&dp_raw)[32], uint32_t (&p_pk)[16],
                                                uint32_t (&ds_pk)[16], const TcQuerySource& src, int b, int j,
                                                bool key_ok, int q0, int ie, const float4* rs, const float* relq,
                                                int rw, float scale, float neg) {
#define MLT_CALL(IRV) bwd_kv_chunk<MR, IRV>(s_raw, dp_raw, p_pk, ds_pk, src, b, j, key_ok, q0, ie, rs, relq, rw, scale, neg)
  switch (src.q.side.id_rule) {
    case IDR_EXPLICIT: MLT_CALL(IDR_EXPLICIT); break;
    case IDR_1D: MLT_CALL(IDR_1D); break;
    case IDR_CROSS_QSENT: MLT_CALL(IDR_CROSS_QSENT); break;
    case IDR_CROSS_KSENT: MLT_CALL(IDR_CROSS_KSENT); break;
    case IDR_2D: MLT_CALL(IDR_2D); break;
    default: MLT_CALL(IDR_NONE); break;
  }
#undef MLT_CALL
}

__global__ void __launch_bounds__(NTHREADS, 1)
tc_bwd_kv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_do0,
                 const __grid_constant__ CUtensorMap map_q1, const __grid_constant__ CUtensorMap map_do1,
                 const TcBwdKVParams p) {
  using namespace bk;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
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
  if (warp == 9) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SrcRange r0 = src_range(p.src[0], j0);
  SrcRange r1{0, 0, 0};
  if (p.nsrc > 1) r1 = src_range(p.src[1], j0);
  const int nchunks = r0.n + r1.n;

  if (warp == 8) {
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
  } else if (warp == 9) {
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
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->pds_full[pc & 1], (pc >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t q_addr = smem_u32(smem + SM_QD + st * (2 * TN * 128));
          const uint32_t do_addr = q_addr + TN * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dV += P^T . dO_c
            umma_ts(tmem + T_DV, tmem + T_S + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(do_addr + kk * 2048, 16, 1024), idesc_acc, (pc > 0 || kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)   // dK += dS^T . Q_c
            umma_ts(tmem + T_DK, tmem + T_DP + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(q_addr + kk * 2048, 16, 1024), idesc_acc, (pc > 0 || kk > 0));
          umma_commit(&bars->qd_empty[st]);
          if (pc == nchunks - 1) umma_commit(&bars->acc_full);
        }
      }
    }
  } else {
    const int row = (warp & 3) * 32 + lane;
    const int hh = warp >> 2;
    const int j = j0 + row;
    const bool key_ok = j < p.len;
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    for (int c = 0; c < nchunks; ++c) {
      const int st = c % NST;
      const bool first = c < r0.n;
      const TcQuerySource& src = first ? p.src[0] : p.src[1];
      const int q0 = (first ? r0.ib + c * TN : r1.ib + (c - r0.n) * TN) + hh * 32;
      const int ie = first ? r0.ie : r1.ie;
      mbar_wait_warp(&bars->qd_full[st], (c / NST) & 1);   // rowstat / allrel rows of this chunk
      mbar_wait_warp(&bars->sdp_full[c & 1], (c >> 1) & 1);
      tc_fence_after_sync();
      const float4* rs = reinterpret_cast<const float4*>(smem + SM_RS + st * TN * 16) + hh * 32;
      const float* relq = reinterpret_cast<const float*>(smem + SM_RELQ + st * TN * 64 * 4) + hh * 32 * src.rw;
      uint32_t s_raw[32], dp_raw[32], p_pk[16], ds_pk[16];
      tmem_ld32(tmem + T_S + (c & 1) * 64 + lane_sel + hh * 32, s_raw);
      tmem_ld32(tmem + T_DP + (c & 1) * 64 + lane_sel + hh * 32, dp_raw);
      tmem_wait_ld();
#define MLT_CALL(MRV) bwd_kv_chunk_ir<MRV>(s_raw, dp_raw, p_pk, ds_pk, src, b, j, key_ok, q0, ie, rs, relq, src.rw, p.scale, p.neg)
      switch (src.q.side.mask_rule) {
        case MR_EXPLICIT: MLT_CALL(MR_EXPLICIT); break;
        case MR_EXAMPLE_ID: MLT_CALL(MR_EXAMPLE_ID); break;
        default: MLT_CALL(MR_NONE); break;
      }
#undef MLT_CALL
      tmem_st16(tmem + T_S + (c & 1) * 64 + lane_sel + hh * 32, p_pk);
      tmem_st16(tmem + T_DP + (c & 1) * 64 + lane_sel + hh * 32, ds_pk);
      tmem_wait_st();
      tc_fence_before_sync();
      mbar_arrive(&bars->pds_full[c & 1]);
    }
    mbar_wait_warp(&bars->acc_full, 0);
    tc_fence_after_sync();
    uint32_t dv_raw[32], dk_raw[32];
    tmem_ld32(tmem + T_DV + lane_sel + hh * 32, dv_raw);
    tmem_ld32(tmem + T_DK + lane_sel + hh * 32, dk_raw);
    tmem_wait_ld();
    if (key_ok) {
      __nv_bfloat16* dv = row_ptr_mut<__nv_bfloat16>(p.d_v, b, j, h) + hh * 32;
      __nv_bfloat16* dk = row_ptr_mut<__nv_bfloat16>(p.d_k, b, j, h) + hh * 32;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
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
  if (warp == 9) tmem_dealloc<TMEM_COLS>(tmem);
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

int tc_launch_bwd_q(const BwdQArgs& a, void* ws, cudaStream_t st) {
  if (!g_attr_q) {
    cudaError_t e = cudaFuncSetAttribute(tc_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bq::SM_ALLOC);
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
  dim3 grid((a.rows.len + TM - 1) / TM, a.H, a.B);
  tc_bwd_q_kernel<<<grid, NTHREADS, bq::SM_ALLOC, st>>>(mq, mdo, mk0, mv0, mk1, mv1, me, p);
  return (int)cudaGetLastError();
}

int tc_launch_bwd_kv(const BwdKVArgs& a, void* const ws[2], cudaStream_t st) {
  if (!g_attr_kv) {
    cudaError_t e = cudaFuncSetAttribute(tc_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bk::SM_ALLOC);
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
  tc_bwd_kv_kernel<<<grid, NTHREADS, bk::SM_ALLOC, st>>>(mk, mv, mq[0], mdo[0], mq[1], mdo[1], p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
