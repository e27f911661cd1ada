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
#include "tc_rowscore.cuh"

namespace mlt {
namespace {

using namespace ptx;

constexpr int TM = 128;       // query rows per tile
constexpr int TN = 64;        // keys per chunk
constexpr int NST = 3;        // K/V ring stages
constexpr int NTHREADS = 256;   // warpgroup 0: softmax; warpgroup 1: TMA, MMA, 2 idle warps
constexpr uint32_t TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

constexpr int SM_Q = 0;                          // 16 KB
constexpr int SM_E = SM_Q + TM * 128;            // 8 KB  (<= 64 ids x 128 B)
constexpr int SM_KV = SM_E + 64 * 128;           // NST x (K 8 KB + V 8 KB)
constexpr int SM_REL = SM_KV + NST * 2 * TN * 128;   // [64][128] fp32 = 32 KB
constexpr int SM_BAR = SM_REL + 64 * TM * 4;
constexpr int SM_TOTAL = SM_BAR + 256;
constexpr int SM_ALLOC = SM_TOTAL + 1024;        // slack for 1024-B alignment

#ifdef MLT_TC_TRACE
__device__ unsigned long long g_trace[3][256];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(role, idx)                                                                  \
  do {                                                                                    \
    if (blockIdx.x == 5 && blockIdx.y == 1 && blockIdx.z == 0 && (idx) < 256) g_trace[role][idx] = gtime(); \
  } while (0)
#else
#define TRACE(role, idx) do {} while (0)
#endif

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

__device__ __forceinline__ rowscore::SegCtx make_seg_ctx(const KeySeg& sg, const SegRange& r, int R, int pd,
                                                        bool perm) {
  rowscore::SegCtx sc;
  sc.sg = &sg;
  sc.kb = r.kb;
  sc.ke = r.ke;
  sc.R = R;
  sc.D = sg.side.max_distance;
  sc.pd = pd;
  sc.perm = perm;
  sc.band = sg.band != 0;
  sc.radius = sg.radius;
  sc.mask_rule = sg.side.mask_rule;
  sc.id_rule = R > 0 ? sg.side.id_rule : IDR_NONE;
  return sc;
}

__global__ void __launch_bounds__(NTHREADS, 2)
tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k0,
              const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_k1,
              const __grid_constant__ CUtensorMap map_v1, const __grid_constant__ CUtensorMap map_e,
              const TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: keeps the shared address space
  // (LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
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

  if (warp >= 4) {
    // warpgroup 1 (TMA, MMA, 2 idle warps that only complete the warpgroup) gives registers away
    setmaxnreg_dec<40>();
  }
  if (warp >= 6) {
  } else if (warp == 4) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      TRACE(1, 0);
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
        TRACE(1, 2 + c);
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
      TRACE(2, 0);
      mbar_wait(&bars->q_full, 0);
      TRACE(2, 1);
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
          TRACE(2, 2 + 4 * c);
          if (c == 1 && rpad) mbar_wait(&bars->rel_done, 0);  // allrel aliases S1
          tc_fence_after_sync();
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + (c & 1) * 64, make_smem_desc_sw128(q_addr + kk * 32, 16, 1024),
                    make_smem_desc_sw128(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0);
          umma_commit(&bars->s_full[c & 1]);
          TRACE(2, 3 + 4 * c);
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->p_full[pc & 1], (pc >> 1) & 1);
          TRACE(2, 4 + 4 * pc);
          tc_fence_after_sync();
          const uint32_t v_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128) + TN * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + 128 + (pc & 1) * 64, tmem + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    make_smem_desc_sw128(v_addr + kk * 2048, 16, 1024), idesc_o, kk > 0);
          umma_commit(&bars->o_full[pc & 1]);
          umma_commit(&bars->kv_empty[st]);
        }
      }
    }
  } else {
    // ===================== softmax / epilogue (warps 0-3) =====================
    setmaxnreg_inc<216>();
    using namespace rowscore;
    const int row = tid, lane = tid & 31;
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const int wrow0 = i0 + warp * 32;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const int pd = a.seg[0].side.max_distance;
    const bool perm = (2 * pd + 1 <= R);
    if (tid == 0) TRACE(0, 0);
    // ---- early global loads: row scalars, bias lanes, first chunk's key lanes ----
    SegCtx sc0 = make_seg_ctx(a.seg[0], r0, R, pd, perm);
    SegCtx sc1 = make_seg_ctx(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm);
    RowCtx rc0, rc1;
    rc0.i = rc1.i = i;
    rc0.row = rc1.row = row;
    rc0.row_ok = rc1.row_ok = row_ok;
    init_row_loads(rc0, sc0, b);
    init_row_loads(rc1, sc1, b);
    auto chunk_key0 = [&](int c) { return c < r0.n ? r0.kb + c * TN : r1.kb + (c - r0.n) * TN; };
    GroupLanes gl0{0, -1}, gl1{0, -1};
    if (nchunks > 0) {
      const SegCtx& s = 0 < r0.n ? sc0 : sc1;
      gl0 = load_group_lanes(s, b, chunk_key0(0), lane);
      gl1 = load_group_lanes(s, b, chunk_key0(0) + 32, lane);
    }
    float bias_l0 = 0.f, bias_l1 = 0.f;  // lane l holds bias[l], bias[32 + l]
    if (rpad) {
      const __nv_bfloat16* bias = reinterpret_cast<const __nv_bfloat16*>(a.rows.bias);
      if (lane < R) bias_l0 = __bfloat162float(bias[lane * a.H + h]);
      if (lane + 32 < R) bias_l1 = __bfloat162float(bias[(lane + 32) * a.H + h]);
      mbar_wait_warp(&bars->rel_full, 0);
      if (tid == 0) TRACE(0, 1);
      tc_fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < rpad; c0 += 16) {
        {
          uint32_t v[16];
          tmem_ld16(tmem + 64 + lane_sel + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int x = 0; x < 16; ++x) {
            const int pid = c0 + x;
            const float bv = __shfl_sync(0xffffffffu, c0 < 32 ? bias_l0 : bias_l1, pid & 31);
            if (pid < R) rel_s[slot_of_id(pid, pd, perm) * TM + row] = (__uint_as_float(v[x]) + bv) * a.scale;
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->rel_done);
    }
    if (tid == 0) TRACE(0, 2);
    init_row(rc0, sc0, b, rel_s);
    init_row(rc1, sc1, b, rel_s);

    float m = -1e30f, l = 0.f, alpha_prev = 0.f;
    float o[64];
#pragma unroll
    for (int x = 0; x < 64; ++x) o[x] = 0.f;

    auto add_o = [&](int pc) {   // o = o * alpha_prev + O_pc
      mbar_wait_warp(&bars->o_full[pc & 1], (pc >> 1) & 1);
      if (tid == 0) TRACE(0, 7 + 4 * pc);
      tc_fence_after_sync();
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(tmem + 128 + (pc & 1) * 64 + lane_sel + hh * 32, v);
        tmem_wait_ld();
        if (hh == 0) {
#pragma unroll
          for (int x = 0; x < 32; ++x) o[x] = fmaf(o[x], alpha_prev, __uint_as_float(v[x]));
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) o[32 + x] = fmaf(o[32 + x], alpha_prev, __uint_as_float(v[x]));
        }
      }
    };
    int c = 0;   // chunk counter over both segments
#pragma unroll 1
    for (int sgi = 0; sgi < a.nseg; ++sgi) {
      // segment context copied once: no per-field selects inside the chunk loop
      const SegCtx sc = sgi ? sc1 : sc0;
      const RowCtx rc = sgi ? rc1 : rc0;
      const int seg_n = sgi ? r1.n : r0.n;
      const int seg_kb = sgi ? r1.kb : r0.kb;
#pragma unroll 1
      for (int cc = 0; cc < seg_n; ++cc, ++c) {
        const int key0 = seg_kb + cc * TN;
        const uint32_t t_s = tmem + (c & 1) * 64 + lane_sel;
        // prefetch the key-side lane scalars of the next chunk
        GroupLanes nl0{0, -1}, nl1{0, -1};
        if (cc + 1 < seg_n) {
          nl0 = load_group_lanes(sc, b, key0 + TN, lane);
          nl1 = load_group_lanes(sc, b, key0 + TN + 32, lane);
        } else if (sgi == 0 && a.nseg > 1 && r1.n > 0) {
          nl0 = load_group_lanes(sc1, b, r1.kb, lane);
          nl1 = load_group_lanes(sc1, b, r1.kb + 32, lane);
        }
        if (tid == 0) TRACE(0, 4 + 4 * c);
        mbar_wait_warp(&bars->s_full[c & 1], (c >> 1) & 1);
        if (tid == 0) TRACE(0, 5 + 4 * c);
        tc_fence_after_sync();
        // ---- pass 1 (per 32-key group, ONE copy of the code): scores -> TMEM, running max ----
        float mx = -INFINITY;
        uint32_t dead_mask = 0;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const GroupLanes gl = g ? gl1 : gl0;
          const int g0 = key0 + 32 * g;
          const GroupPlan gp = classify(sc, rc, gl, wrow0, g0, lane, a.neg);
          if (gp.mode == GM_DEAD) {
            dead_mask |= 1u << g;
            continue;
          }
          if (gp.mode == GM_GEN) score_group_generic_tmem(t_s + 32 * g, sc, rc, gl, b, g0, rel_s, a.scale, a.neg);
          float t[32];
          {
            uint32_t v[32];
            tmem_ld32(t_s + 32 * g, v);
            tmem_wait_ld();
#pragma unroll
            for (int x = 0; x < 32; ++x) t[x] = __uint_as_float(v[x]);
          }
          score_group<0>(t, gp, sc, rc, gl, b, g0, rel_s, a.scale, a.neg);
#pragma unroll
          for (int x = 0; x < 32; ++x) mx = fmaxf(mx, t[x]);
          if (gp.mode != GM_GEN) {
            uint32_t v[32];
#pragma unroll
            for (int x = 0; x < 32; ++x) v[x] = __float_as_uint(t[x]);
            tmem_st32(t_s + 32 * g, v);
          }
        }
        tmem_wait_st();
        const float m_new = fmaxf(m, mx);
        const float alpha = ex2((m - m_new) * LOG2E);
        const float mb = m_new * LOG2E;
        // ---- pass 2: p = exp2(t * log2e - mb), row sum, P (bf16) back into TMEM ----
        float lsum = 0.f;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          uint32_t pk[16];
          if (dead_mask & (1u << g)) {
#pragma unroll
            for (int x = 0; x < 16; ++x) pk[x] = 0u;
          } else {
            uint32_t v[32];
            tmem_ld32(t_s + 32 * g, v);
            tmem_wait_ld();
#pragma unroll
            for (int x = 0; x < 16; ++x) {
              const float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), LOG2E, -mb));
              const float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), LOG2E, -mb));
              lsum += p0 + p1;
              pk[x] = pack_bf16x2(p0, p1);
            }
          }
          // P of keys [32g, 32g+32) -> packed columns [32g, 32g+16): stays inside this group's own
          // score columns, so the not-yet-read scores of group 1 are never clobbered
          tmem_st16(t_s + 32 * g, pk);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[c & 1]);
        if (tid == 0) TRACE(0, 6 + 4 * c);
        l = l * alpha + lsum;
        m = m_new;
        gl0 = nl0;
        gl1 = nl1;
        if (c >= 1) add_o(c - 1);   // uses alpha_prev = alpha of chunk c-1
        alpha_prev = alpha;
      }
    }
    if (nchunks >= 1) add_o(nchunks - 1);
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
  if (tid == 0) TRACE(0, 3);
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

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 3 * 256);
}
#endif

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
