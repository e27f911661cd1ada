// tcgen05 forward kernel: one 128-row query tile per CTA, joint online softmax over up to two
// key segments (band / dense), relative-position scores and masks fused into the score tile.
//
// Roles (192 threads, 2 CTAs per SM, 256 TMEM columns each):
//   warps 0-3  softmax / epilogue: thread t owns query row t == TMEM lane t.  Reads S from TMEM,
//              adds the relative score (gather from a per-row table in smem), applies masks,
//              online softmax in registers, writes P (bf16) back into TMEM, accumulates O in
//              registers from the per-chunk P.V results.
//   warp 4     TMA producer: Q tile, relative-embedding tile, then K/V chunks through a 3-stage
//              mbarrier ring (SWIZZLE_128B tiles, out-of-range rows zero-filled by TMA).
//   warp 5     MMA issuer (one elected lane): S_c = Q.K_c^T (SS), O_c = P_c.V_c (TS: P from TMEM,
//              V MN-major), allrel = Q.E^T.  S is double-buffered so S_{c+1} runs under softmax_c.
//
// TMEM map (columns): [0,64) S0/P0, [64,128) S1/P1 (also allrel before chunk 1), [128,192) O0,
// [192,256) O1.
#include "tc_api.cuh"

#include "mlt_common.cuh"
#include "profile.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace {

using namespace ptx;

constexpr int TM = 128;       // query rows per tile
constexpr int TN = 64;        // keys per chunk
constexpr int NST = 3;        // K/V ring stages
constexpr int NTHREADS = 192;
constexpr uint32_t TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

constexpr int SM_Q = 0;                          // 16 KB
constexpr int SM_E = SM_Q + TM * 128;            // 8 KB  (<= 64 ids x 128 B)
constexpr int SM_KV = SM_E + 64 * 128;           // NST x (K 8 KB + V 8 KB)
constexpr int SM_REL = SM_KV + NST * 2 * TN * 128;   // [64][128] fp32 = 32 KB
constexpr int SM_BAR = SM_REL + 64 * TM * 4;
constexpr int SM_TOTAL = SM_BAR + 256;
constexpr int SM_ALLOC = SM_TOTAL + 1024;        // slack for 1024-B alignment

struct TcFwdParams {
  FwdArgs a;
  int rpad;  // R rounded up to a multiple of 16 (0: no relative term)
};

struct Bars {
  uint64_t q_full, rel_full, rel_done;
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t s_full[2], p_full[2], o_full[2];
  uint32_t tmem_base;
};

struct SegRange {
  int kb, ke, n;  // keys [kb, ke), n chunks of TN
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

// Scores of one chunk for one row: t[jj] = (x + rel) * scale (+ neg if masked), -inf if dead.
template <int MR, int IR>
__device__ __forceinline__ void score_chunk(float (&t)[TN], const KeySeg& sg, int b, int i, int row,
                                            bool row_ok, int key0, int ke, int R, const float* rel_s,
                                            float scale, float neg) {
  const Side& sd = sg.side;
  int q_e = 0, q_sent = -1;
  if (MR == MR_EXAMPLE_ID && row_ok) q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
  if (IR == IDR_CROSS_QSENT && row_ok) q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
  const int32_t* mrow = nullptr;
  const int32_t* irow = nullptr;
  if (MR == MR_EXPLICIT) mrow = sd.mask + (int64_t)b * sd.sb + (int64_t)i * sd.sq;
  if (IR == IDR_EXPLICIT) irow = sd.ids + (int64_t)b * sd.sb + (int64_t)i * sd.sq;
#pragma unroll
  for (int jj = 0; jj < TN; ++jj) {
    const int j = key0 + jj;
    const int off = j - i;
    const bool live = row_ok && j < ke && (!sg.band || (off <= sg.radius && off >= -sg.radius));
    const int col = sg.band ? off + sg.radius : j;
    bool ok = true;
    int id = -1;
    if (live) {
      if (MR == MR_EXPLICIT) ok = __ldg(mrow + col) != 0;
      if (MR == MR_EXAMPLE_ID) ok = (q_e == __ldg(sd.k_eid + (int64_t)b * sd.k_len + j));
      if (IR == IDR_EXPLICIT) id = __ldg(irow + col);
      if (IR == IDR_1D) id = rel_id_1d(off, sd.max_distance);
      if (IR == IDR_CROSS_QSENT) id = 2 * sd.max_distance + 1 + (q_sent == j ? 1 : 0);
      if (IR == IDR_CROSS_KSENT)
        id = 2 * sd.max_distance + 1 + (__ldg(sd.sent + (int64_t)b * sd.sent_len + j) == i ? 1 : 0);
      if (IR == IDR_2D) id = rel_id_2d(i, j, sd.npr, sd.core, sd.max_distance);
    }
    float rel = 0.f;
    if (IR != IDR_NONE && id >= 0 && id < R) rel = rel_s[id * TM + row];
    float v = fmaf(t[jj], scale, rel);  // rel_s already holds allrel * scale
    if (!ok) v += neg;
    t[jj] = live ? v : -INFINITY;
  }
}

template <int MR>
__device__ __forceinline__ void score_chunk_ir(float (&t)[TN], const KeySeg& sg, int b, int i, int row,
                                               bool row_ok, int key0, int ke, int R,
                                               const float* rel_s, float scale, float neg) {
  switch (sg.side.id_rule) {
    case IDR_EXPLICIT: score_chunk<MR, IDR_EXPLICIT>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
    case IDR_1D: score_chunk<MR, IDR_1D>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
    case IDR_CROSS_QSENT: score_chunk<MR, IDR_CROSS_QSENT>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
    case IDR_CROSS_KSENT: score_chunk<MR, IDR_CROSS_KSENT>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
    case IDR_2D: score_chunk<MR, IDR_2D>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
    default: score_chunk<MR, IDR_NONE>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, scale, neg); break;
  }
}

__global__ void __launch_bounds__(NTHREADS, 2)
tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k0,
              const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_k1,
              const __grid_constant__ CUtensorMap map_v1, const __grid_constant__ CUtensorMap map_e,
              const TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + SM_REL);
  const FwdArgs& a = p.a;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * TM;
  const int R = a.rows.R, rpad = p.rpad;

  if (tid == 0) {
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->rel_full, 1);
    mbar_init(&bars->rel_done, 128);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 128);
      mbar_init(&bars->o_full[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<TMEM_COLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SegRange r0 = seg_range(a.seg[0], i0);
  SegRange r1{0, 0, 0};
  if (a.nseg > 1) r1 = seg_range(a.seg[1], i0);
  const int nchunks = r0.n + r1.n;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      prefetch_tensormap(&map_q);
      prefetch_tensormap(&map_k0);
      prefetch_tensormap(&map_v0);
      mbar_arrive_expect_tx(&bars->q_full, TM * 128 + rpad * 128);
      tma_load_4d(smem + SM_Q, &map_q, &bars->q_full, 0, i0, h, b);
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
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t q_addr = smem_u32(smem + SM_Q);
      mbar_wait(&bars->q_full, 0);
      tc_fence_after_sync();
      if (rpad) {
        const uint32_t idesc_r = make_idesc_bf16(TM, rpad, 0, 0);
        const uint32_t e_addr = smem_u32(smem + SM_E);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + 64, make_smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                  make_smem_desc_sw128(e_addr + kk * 32, 16, 1024), idesc_r, kk > 0);
        umma_commit(&bars->rel_full);
      }
      for (int c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
          const int st = c % NST;
          mbar_wait(&bars->kv_full[st], (c / NST) & 1);
          if (c == 1 && rpad) mbar_wait(&bars->rel_done, 0);  // allrel aliases S1
          tc_fence_after_sync();
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + (c & 1) * 64, make_smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&bars->s_full[c & 1]);
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->p_full[pc & 1], (pc >> 1) & 1);
          tc_fence_after_sync();
          const uint32_t v_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128) + TN * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + 128 + (pc & 1) * 64, tmem + (pc & 1) * 64 + kk * 8,
                    make_smem_desc_sw128(v_addr + kk * 2048, 16, 1024), idesc_o, kk > 0);
          umma_commit(&bars->o_full[pc & 1]);
          umma_commit(&bars->kv_empty[st]);
        }
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 0-3) =====================
    const int row = tid;
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    if (rpad) {
      mbar_wait(&bars->rel_full, 0);
      tc_fence_after_sync();
      const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.rows.bias);
      for (int c0 = 0; c0 < rpad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + 64 + lane_sel + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const int pid = c0 + x;
          if (pid < R)
            rel_s[pid * TM + row] =
                (__uint_as_float(v[x]) + __bfloat162float(bias[pid * a.H + h])) * a.scale;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->rel_done);
    }
    float m = -1e30f, l = 0.f, alpha_prev = 0.f;
    float o[64];
#pragma unroll
    for (int x = 0; x < 64; ++x) o[x] = 0.f;

    for (int c = 0; c <= nchunks; ++c) {
      float alpha = 0.f;
      if (c < nchunks) {
        const bool first = c < r0.n;
        const KeySeg& sg = first ? a.seg[0] : a.seg[1];
        const int key0 = first ? r0.kb + c * TN : r1.kb + (c - r0.n) * TN;
        const int ke = first ? r0.ke : r1.ke;
        mbar_wait(&bars->s_full[c & 1], (c >> 1) & 1);
        tc_fence_after_sync();
        float t[TN];
        {
          uint32_t v[32];
          tmem_ld32(tmem + (c & 1) * 64 + lane_sel, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 32; ++x) t[x] = __uint_as_float(v[x]);
          tmem_ld32(tmem + (c & 1) * 64 + lane_sel + 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 32; ++x) t[32 + x] = __uint_as_float(v[x]);
        }
        switch (sg.side.mask_rule) {
          case MR_EXPLICIT: score_chunk_ir<MR_EXPLICIT>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, a.scale, a.neg); break;
          case MR_EXAMPLE_ID: score_chunk_ir<MR_EXAMPLE_ID>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, a.scale, a.neg); break;
          default: score_chunk_ir<MR_NONE>(t, sg, b, i, row, row_ok, key0, ke, R, rel_s, a.scale, a.neg); break;
        }
        float mx = t[0];
#pragma unroll
        for (int x = 1; x < TN; ++x) mx = fmaxf(mx, t[x]);
        const float m_new = fmaxf(m, mx);
        alpha = ex2((m - m_new) * LOG2E);
        const float mb = m_new * LOG2E;
        float lsum = 0.f;
        uint32_t pk[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const float p0 = ex2(fmaf(t[2 * x], LOG2E, -mb));
          const float p1 = ex2(fmaf(t[2 * x + 1], LOG2E, -mb));
          lsum += p0 + p1;
          pk[x] = pack_bf16x2(p0, p1);
        }
        tmem_st32(tmem + (c & 1) * 64 + lane_sel, pk);
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[c & 1]);
        l = l * alpha + lsum;
        m = m_new;
      }
      if (c >= 1) {
        const int pc = c - 1;
        mbar_wait(&bars->o_full[pc & 1], (pc >> 1) & 1);
        tc_fence_after_sync();
        uint32_t v[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          tmem_ld32(tmem + 128 + (pc & 1) * 64 + lane_sel + hh * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 32; ++x) o[hh * 32 + x] = fmaf(o[hh * 32 + x], alpha_prev, __uint_as_float(v[x]));
        }
      }
      alpha_prev = alpha;
    }
    if (row_ok) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(a.out, b, i, h);
#pragma unroll
      for (int x = 0; x < 8; ++x) {
        uint4 w;
        w.x = pack_bf16x2(o[8 * x + 0] * inv, o[8 * x + 1] * inv);
        w.y = pack_bf16x2(o[8 * x + 2] * inv, o[8 * x + 3] * inv);
        w.z = pack_bf16x2(o[8 * x + 4] * inv, o[8 * x + 5] * inv);
        w.w = pack_bf16x2(o[8 * x + 6] * inv, o[8 * x + 7] * inv);
        *reinterpret_cast<uint4*>(dst + 8 * x) = w;
      }
      float2* st = reinterpret_cast<float2*>(a.stats) + ((int64_t)(b * a.H + h) * a.rows.len + i);
      *st = make_float2(m, l);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem);
}

bool t4_ok(const T4& t) {
  return t.ptr && (t.sb % 8 == 0) && (t.sl % 8 == 0) && (t.sh % 8 == 0) &&
         (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0);
}

}  // namespace

bool tc_fwd_args_supported(const FwdArgs& a, int dtype, int d) {
  if (dtype != MLT_BF16 || d != 64 || a.rows.R > 64) return false;
  if (!t4_ok(a.rows.q) || !t4_ok(a.out)) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (!t4_ok(a.seg[s].k) || !t4_ok(a.seg[s].v)) return false;
  if (a.rows.R > 0 && reinterpret_cast<uintptr_t>(a.rows.emb) % 16) return false;
  return get_encode_tiled() != nullptr;
}

int tc_launch_fwd(const FwdArgs& a, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  TcFwdParams p;
  p.a = a;
  p.rpad = a.rows.R > 0 ? (a.rows.R + 15) / 16 * 16 : 0;
  CUtensorMap mq, mk0, mv0, mk1, mv1, me;
  int e = 0;
  e |= make_qkv_tensor_map(&mq, a.rows.q.ptr, a.rows.q.sb, a.rows.q.sl, a.rows.q.sh, a.B, a.rows.len, a.H, TM);
  e |= make_qkv_tensor_map(&mk0, a.seg[0].k.ptr, a.seg[0].k.sb, a.seg[0].k.sl, a.seg[0].k.sh, a.B, a.seg[0].len, a.H, TN);
  e |= make_qkv_tensor_map(&mv0, a.seg[0].v.ptr, a.seg[0].v.sb, a.seg[0].v.sl, a.seg[0].v.sh, a.B, a.seg[0].len, a.H, TN);
  const KeySeg& s1 = a.nseg > 1 ? a.seg[1] : a.seg[0];
  e |= make_qkv_tensor_map(&mk1, s1.k.ptr, s1.k.sb, s1.k.sl, s1.k.sh, a.B, s1.len, a.H, TN);
  e |= make_qkv_tensor_map(&mv1, s1.v.ptr, s1.v.sb, s1.v.sl, s1.v.sh, a.B, s1.len, a.H, TN);
  if (p.rpad) {
    e |= make_qkv_tensor_map(&me, a.rows.emb, (int64_t)a.rows.R * a.H * 64, (int64_t)a.H * 64, 64, 1,
                             a.rows.R, a.H, p.rpad);
  } else {
    me = mq;
  }
  if (e) return MLT_ERR_UNSUPPORTED;
  dim3 grid((a.rows.len + TM - 1) / TM, a.H, a.B);
  tc_fwd_kernel<<<grid, NTHREADS, SM_ALLOC, st>>>(mq, mk0, mv0, mk1, mv1, me, p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
