// gl2: persistent query-centric backward kernel for the long rows of compact global-local attention
// (same specialisation as gl2_fwd.cu: bf16, d = 64, example-id masks, 1-D band ids + sentence cross ids,
// relative vocabulary <= 32, local_radius == 64).
//
//   S = Q.K_c^T, dP = dO.V_c^T (SS)  ->  p = ex2(x * scale*log2e + c_row),  ds = p (dp - delta)
//   ->  dQ += dS.K_c (dS as the TMEM A operand),  per-row relative-id bins dallrel[i, id] += ds,
//   dQ += dallrel.E,  per-tile table-gradient partial dallrel^T.Q,  and the exponent-ready row records the
//   key-centric pass consumes (same workspace format as tc_bwd.cu, so the two families interoperate).
//
// One CTA per SM, 576 threads, persistent over 128-row query tiles:
//   warps 0-15  elementwise: warp w owns row quadrant (w & 3) and the 32-key group (w >> 2) of every
//               128-key chunk (four threads per row: the backward needs no row reductions, and with only
//               ~10 warps per SM these kernels are bound by per-warp latency, not by issue slots)
//   warp 16     TMA producer: Q / dO / E tiles of the NEXT tile while the current one computes; K/V chunks
//               (two interleaved 64-key blocks each, see gl2_fwd.cu) through a 3-stage ring
//   warp 17     MMA issuer.  S runs one chunk ahead in a second buffer; dP_c follows dQ_{c-1}
// TMEM (512 columns): S0 [128] S1 [128] dP / dS [128] dQ [64] allrel / dallrel 2 x [32]; the table-gradient
// tile (M = 64) reuses S1 after the last chunk.  The MMA warp starts the next tile (allrel, S_0, dP_0) as
// soon as this tile's last MMAs are issued, i.e. under the elementwise warps' epilogue.
#include <atomic>
#include "tc_api.cuh"

#include "gl2_geom.cuh"
#include "mlt_common.cuh"
#include "profile.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace gl2 {
namespace bq {

using namespace ptx;

constexpr int NEWARPS = 16;              // elementwise warps
constexpr int NTHREADS = (NEWARPS + 2) * 32;
constexpr int NEW = NEWARPS * 32;        // elementwise threads
constexpr int RSF = 68;                  // record field stride (floats), see tc_bwd.cu

constexpr int SM_Q = 0;                               // [2 bufs] x (Q 16 KB + dO 16 KB)
constexpr int SM_E = SM_Q + 2 * 2 * TM * 128;         // [2 bufs] x 4 KB
constexpr int SM_KV = SM_E + 2 * 32 * 128;            // NST x (K 16 KB + V 16 KB)
constexpr int SM_REL = SM_KV + NST * 2 * TK * 128;    // [32 slots][128 rows] f32, log2 units
constexpr int SM_BIN = SM_REL + 32 * TM * 4;          // [32 slots][128 rows] f32
constexpr int SM_A = SM_BIN + 32 * TM * 4;            // dallrel^T tile, bf16 [128 rows][64 ids] SW128 (16 KB)
constexpr int SM_CMB = SM_A + TM * 128;               // class sums of column groups 1..3: [3][4][128] f32
constexpr int SM_BS = SM_CMB + 3 * 4 * TM * 4;        // bias partial sums [4 quadrants][32]
constexpr int SM_BIAS = SM_BS + 4 * 32 * 4;           // [32] bias * scale * log2e
constexpr int SM_BAR = SM_BIAS + 32 * 4;
constexpr int SM_SCHED = SM_BAR + 256;                // ring of tile descriptors {t, b, h, tile} (16 B each)
constexpr int SM_TOTAL = SM_SCHED + 64;
constexpr int SM_ALLOC = SM_TOTAL + 1024;
static_assert(SM_ALLOC <= 227 * 1024, "shared memory budget");

// allrel / dallrel alternate between two 32-column buffers so that the next tile's allrel MMA can run under
// this tile's epilogue; the table-gradient tile (M = 64) lands in S1, which the next tile needs last
constexpr uint32_t T_S0 = 0, T_S1 = 128, T_DP = 256, T_DQ = 384, T_REL = 448, T_DE = T_S1;

#ifdef MLT_TC_TRACE
__device__ long long g_trace_q[3][1024];
__device__ int g_trace_n[3];
#define GTRACE(role, code)                                                               \
  do {                                                                                   \
    if (blockIdx.x == 3) {                                                               \
      const int n_ = g_trace_n[role];                                                    \
      if (n_ < 511) { g_trace_q[role][2 * n_] = clock64(); g_trace_q[role][2 * n_ + 1] = (code); g_trace_n[role] = n_ + 1; } \
    }                                                                                    \
  } while (0)
#else
#define GTRACE(role, code) do {} while (0)
#endif

struct Params {
  int B, H, L, G, R, D;
  float scale;
  const int32_t* long_eid;
  const int32_t* glob_eid;
  const int32_t* sent;
  const __nv_bfloat16* bias;     // [R, H]
  T4 d_q;
  const float4* rowstat;         // [B, H, lp] = (m * log2e, 1 / l, delta, 0)
  float* rec_ws;                 // row records for the key-centric pass
  float* tg_partial;             // [(b * ntile + tile) * H + h][R][64]
  float* tg_partial_bias;        // [...][R]
  int lp, rw;
  int tiles_per_bh, total_tiles;
  unsigned* sched;   // {next tile after the first gridDim.x, CTAs done}: 0 at launch, reset by the last CTA (see gl2_fwd.cu)
};

constexpr int NSQ = 4;   // ring of tile numbers, producer thread -> the other roles (dynamic distribution, gl2_fwd.cu)
__device__ unsigned g_sched_bq[256][2];
struct Bars {
  uint64_t q_full[2], q_empty[2];
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t rel_full, s_full[2], dp_full, ds_full, dar_full, dq_full, tile_done;
  uint64_t sched_full[NSQ], sched_empty[NSQ];
  uint32_t tmem_base;
};
// The producer thread decodes the tile number once (two integer divisions) and publishes {t, b, h, tile};
// t < 0 ends the loop.
template <bool WARP>
__device__ __forceinline__ int4 sched_take(Bars* bars, const int4* ring, int it, bool lane0) {
  const int sq = it % NSQ;
  mbar_wait(&bars->sched_full[sq], (it / NSQ) & 1);
  const int4 e = ring[sq];
  if (WARP) __syncwarp();   // every lane has read the entry
  if (lane0) mbar_arrive(&bars->sched_empty[sq]);
  return e;
}
static_assert(sizeof(Bars) <= 256, "barrier block");

struct Tile {
  int b, h, i0, tile;
  int nglob;
};
__device__ __forceinline__ Tile make_tile(const Params& p, int t) {
  Tile q;
  const int bh = t / p.tiles_per_bh;
  q.b = bh / p.H;
  q.h = bh - q.b * p.H;
  q.tile = t - bh * p.tiles_per_bh;
  q.i0 = q.tile * TM;
  q.nglob = (p.G + TK - 1) / TK;
  return q;
}
__device__ __forceinline__ Tile tile_of(const Params& p, const int4& e) {   // from a published descriptor
  Tile q;
  q.b = e.y;
  q.h = e.z;
  q.tile = e.w;
  q.i0 = e.w * TM;
  q.nglob = (p.G + TK - 1) / TK;
  return q;
}
// chunk c of a tile: c < 2 band (k0,k2) / (k1,k3), then global-token chunks
__device__ __forceinline__ void chunk_blocks(const Tile& q, int c, int& ka, int& kb) {
  if (c < 2) {
    ka = q.i0 - RAD + c * 64;
    kb = ka + 128;
  } else {
    ka = (c - 2) * TK;
    kb = ka + 64;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
gl2_bwd_q_long_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                      const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                      const __grid_constant__ CUtensorMap map_gk, const __grid_constant__ CUtensorMap map_gv,
                      const __grid_constant__ CUtensorMap map_e, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int4* ring = reinterpret_cast<int4*>(smem + SM_SCHED);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1);
      mbar_init(&bars->q_empty[s], 1);
      mbar_init(&bars->s_full[s], 1);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 1);
    }
    for (int s = 0; s < NSQ; ++s) {
      mbar_init(&bars->sched_full[s], 1);
      mbar_init(&bars->sched_empty[s], NEWARPS + 1);   // the elementwise warps and the MMA thread
    }
    mbar_init(&bars->rel_full, 1);
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->ds_full, NEW);
    mbar_init(&bars->dar_full, NEW);
    mbar_init(&bars->dq_full, 1);
    mbar_init(&bars->tile_done, NEW);
    fence_barrier_init();
  }
  if (warp == NEWARPS + 1) tmem_alloc<512>(&bars->tmem_base);
  // id columns 32..63 of the dallrel^T tile are never written: zero the tile once
  for (int x = tid; x < TM * 128 / 16; x += NTHREADS) reinterpret_cast<uint4*>(smem + SM_A)[x] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (warp == NEWARPS) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int kvc = 0;
      for (int it = 0;; ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->q_empty[buf], ((it >> 1) & 1) ^ 1);
        int t = it == 0 ? (int)blockIdx.x : (int)(gridDim.x + atomicAdd(p.sched, 1u));
        if (t >= p.total_tiles) t = -1;
        const int sq = it % NSQ;
        mbar_wait(&bars->sched_empty[sq], ((it / NSQ) & 1) ^ 1);
        Tile q{};
        if (t >= 0) q = make_tile(p, t);
        ring[sq] = make_int4(t, q.b, q.h, q.tile);
        mbar_arrive(&bars->sched_full[sq]);
        if (t < 0) break;
        mbar_arrive_expect_tx(&bars->q_full[buf], 2 * TM * 128 + 32 * 128);
        uint8_t* qs = smem + SM_Q + buf * 2 * TM * 128;
        tma_load_4d(qs, &map_q, &bars->q_full[buf], 0, q.i0, q.h, q.b);
        tma_load_4d(qs + TM * 128, &map_do, &bars->q_full[buf], 0, q.i0, q.h, q.b);
        tma_load_4d(smem + SM_E + buf * 32 * 128, &map_e, &bars->q_full[buf], 0, 0, q.h, 0);
        const int nc = 2 + q.nglob;
        for (int c = 0; c < nc; ++c, ++kvc) {
          const int st = kvc % NST;
          GTRACE(2, 200 + c);
          mbar_wait(&bars->kv_empty[st], ((kvc / NST) & 1) ^ 1);
          GTRACE(2, 210 + c);
          uint8_t* ks = smem + SM_KV + st * (2 * TK * 128);
          uint8_t* vs = ks + TK * 128;
          mbar_arrive_expect_tx(&bars->kv_full[st], 2 * TK * 128);
          int ka, kb;
          chunk_blocks(q, c, ka, kb);
          const CUtensorMap* mk = c < 2 ? &map_k : &map_gk;
          const CUtensorMap* mv = c < 2 ? &map_v : &map_gv;
          // A box that lies entirely outside the tensor is not issued as such: its start is clamped so that
          // at least one row is in range (the kernel's own geometry marks those keys dead, and what lands
          // in shared memory is finite either way: zero fill or real rows)
          const int klen = (c < 2) ? p.L : p.G;
          const int ca = min(max(ka, -63), klen - 1), cb = min(max(kb, -63), klen - 1);
          tma_load_4d(ks, mk, &bars->kv_full[st], 0, ca, q.h, q.b);              // boxes of 64 keys
          tma_load_4d(ks + 64 * 128, mk, &bars->kv_full[st], 0, cb, q.h, q.b);
          tma_load_4d(vs, mv, &bars->kv_full[st], 0, ca, q.h, q.b);
          tma_load_4d(vs + 64 * 128, mv, &bars->kv_full[st], 0, cb, q.h, q.b);
        }
      }
    }
  } else if (warp == NEWARPS + 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TK, 0, 0);
      const uint32_t idesc_dq = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t idesc_r = make_idesc_bf16(TM, 32, 0, 0);
      const uint32_t idesc_de = make_idesc_bf16(64, 64, 1, 1);
      const uint32_t a_addr = smem_u32(smem + SM_A);
      int it = 0, kv_base = 0;
      uint32_t ds_cnt = 0;
      // tile prologue on the tensor core: allrel = Q.E^T, S_0, dP_0 (issued one tile ahead of the elementwise warps)
      auto kv_stage = [&](int base, int c) { return (base + c) % NST; };
      auto start_tile = [&](int it_, int base) {
        const int buf = it_ & 1;
        mbar_wait(&bars->q_full[buf], (it_ >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t q_addr = smem_u32(smem + SM_Q + buf * 2 * TM * 128), do_addr = q_addr + TM * 128;
        const uint32_t e_addr = smem_u32(smem + SM_E + buf * 32 * 128);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_REL + 32 * (it_ & 1), sdesc(q_addr).at(kk * 32), sdesc(e_addr).at(kk * 32), idesc_r, kk > 0);
        umma_commit(&bars->rel_full);
        const int st = kv_stage(base, 0);
        mbar_wait(&bars->kv_full[st], (base / NST) & 1);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TK * 128)), v_addr = k_addr + TK * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_S0, sdesc(q_addr).at(kk * 32), sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->s_full[0]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_DP, sdesc(do_addr).at(kk * 32), sdesc(v_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->dp_full);
      };
      int4 te = sched_take<false>(bars, ring, 0, true);
      if (te.x >= 0) start_tile(0, 0);
      for (; te.x >= 0; ++it) {
        const Tile q = tile_of(p, te);
        const int buf = it & 1;
        const int nc = 2 + q.nglob;
        const uint32_t q_addr = smem_u32(smem + SM_Q + buf * 2 * TM * 128), do_addr = q_addr + TM * 128;
        const uint32_t e_addr = smem_u32(smem + SM_E + buf * 32 * 128);
        const uint32_t t_rel = tmem + T_REL + 32 * (it & 1);
        auto kv_addr = [&](int c) { return smem_u32(smem + SM_KV + ((kv_base + c) % NST) * (2 * TK * 128)); };
        auto issue_s = [&](int c) {
          mbar_wait(&bars->kv_full[(kv_base + c) % NST], ((kv_base + c) / NST) & 1);
          tc_fence_after_sync();
          const uint32_t k_addr = kv_addr(c);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + ((c & 1) ? T_S1 : T_S0), sdesc(q_addr).at(kk * 32), sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
          umma_commit(&bars->s_full[c & 1]);
        };
        auto issue_dp = [&](int c) {   // kv_full(c) has been waited for by issue_s(c)
          const uint32_t v_addr = kv_addr(c) + TK * 128;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + T_DP, sdesc(do_addr).at(kk * 32), sdesc(v_addr).at(kk * 32), idesc_s, kk > 0);
          umma_commit(&bars->dp_full);
        };
        // S1 holds the previous tile's table-gradient tile until the elementwise warps have read it
        mbar_wait(&bars->tile_done, (it & 1) ^ 1);
        for (int c = 0; c < nc; ++c) {
          if (c + 1 < nc) issue_s(c + 1);
          mbar_wait(&bars->ds_full, ds_cnt & 1);
          ++ds_cnt;
          tc_fence_after_sync();
          // dS_c sits in the chunk's own S buffer (its scores are spent), so the dP columns are free once dS_c has
          // arrived: dP_{c+1} goes first and runs under the dQ MMAs.  (Issuing it as soon as dP_c has been READ
          // would need two-deep dS / dP hand-over barriers: a warp whose group is dead could otherwise arrive
          // twice in one phase.  Tried, deadlocked, not pursued: the remaining dP wait is ~4 % of the samples.)
          if (c + 1 < nc) issue_dp(c + 1);
          const uint32_t k_addr = kv_addr(c);
          const uint32_t t_ds = tmem + ((c & 1) ? T_S1 : T_S0);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)   // dQ += dS_c . K_c  (dS packed per 32-key group at columns [32g, 32g + 16))
            umma_ts(tmem + T_DQ, t_ds + (kk >> 1) * 32 + (kk & 1) * 8, sdesc(k_addr).at(kk * 2048), idesc_dq,
                    (c > 0 || kk > 0));
          umma_commit(&bars->kv_empty[(kv_base + c) % NST]);
        }
        // dQ += dallrel . E ; table-gradient partial dE[64 ids x 64] = dallrel^T . Q (M = 64, K = the tile's rows)
        mbar_wait(&bars->dar_full, it & 1);
        tc_fence_after_sync();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
          umma_ts(tmem + T_DQ, t_rel + kk * 8, sdesc(e_addr).at(kk * 2048), idesc_dq, 1u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tmem + T_DE, sdesc(a_addr).at(kk * 2048), sdesc(q_addr).at(kk * 2048), idesc_de, kk > 0);
        umma_commit(&bars->dq_full);
        umma_commit(&bars->q_empty[buf]);
        kv_base += nc;
        // next tile's prologue under this tile's epilogue: allrel -> the other buffer, S_0 -> S0 (free since the
        // last even chunk), dP_0 -> dP (its dS was consumed by the dQ MMAs above: MMAs execute in issue order)
        te = sched_take<false>(bars, ring, it + 1, true);
        if (te.x >= 0) start_tile(it + 1, kv_base);
      }
    }
  } else {
    // ===================== elementwise warps =====================
    const int quad = warp & 3, hf = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const uint32_t t_base = tmem + lane_sel;
    float* rel_s = reinterpret_cast<float*>(smem + SM_REL);
    float* bins = reinterpret_cast<float*>(smem + SM_BIN);
    float* cmb = reinterpret_cast<float*>(smem + SM_CMB);
    float* bs = reinterpret_cast<float*>(smem + SM_BS);
    float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS);
    uint8_t* a_tile = smem + SM_A;
    const float scale2 = p.scale * LOG2E;
    const int D = p.D, R = p.R;
    const bool d12 = (D == 12);
    uint32_t it = 0, s_par[2] = {0, 0}, dp_par = 0;
    // row scalars of a tile (example id, sentence, forward statistics): loaded one tile ahead, under the
    // epilogue of the previous tile
    struct RowPre { int q_e, q_sent; float4 rs4; };
    auto row_pre = [&](const int4& te_) {
      RowPre r;
      r.q_e = -2; r.q_sent = -1; r.rs4 = make_float4(0.f, 1.f, 0.f, 0.f);
      if (te_.x >= 0) {
        const Tile q_ = tile_of(p, te_);
        const int i_ = q_.i0 + row;
        if (i_ < p.L) {
          r.q_e = __ldg(p.long_eid + (int64_t)q_.b * p.L + i_);
          r.q_sent = __ldg(p.sent + (int64_t)q_.b * p.L + i_);
          r.rs4 = __ldg(p.rowstat + (int64_t)(q_.b * p.H + q_.h) * p.lp + i_);
        }
      }
      return r;
    };
    int4 te = sched_take<true>(bars, ring, 0, lane == 0);
    RowPre pre = row_pre(te);
    for (; te.x >= 0; ++it) {
      const Tile q = tile_of(p, te);
      const int b = q.b, h = q.h;
      const int i = q.i0 + row;
      const bool row_ok = i < p.L;
      const int a_lo = q.i0 + quad * 32;
      if (tid == 0) GTRACE(0, 0);
      const int q_e = pre.q_e, q_sent = pre.q_sent;
      float nm2l = -INFINITY, delta = 0.f;
      if (row_ok) {
        nm2l = __log2f(pre.rs4.y) - pre.rs4.x;     // -(m*log2e + log2 l)
        delta = pre.rs4.z;
      }
      // zero the bins; bias * scale * log2e of this head
      for (int x = 4 * tid; x < 32 * TM; x += 4 * NEW) *reinterpret_cast<float4*>(bins + x) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < 32) bias_s[tid] = tid < R ? __bfloat162float(p.bias[tid * p.H + h]) * scale2 : 0.f;
      named_bar_sync(1, NEW);
      // ---- relative table (log2 units) + row records; the four (half 0) warps extract allrel ----
      const int rs_stride = p.rw + 4;
      float* rec_row = p.rec_ws + (((int64_t)(b * p.H + h) * (p.lp >> 6) + (i >> 6)) * rs_stride) * RSF + (i & 63);
      const uint32_t t_rel = t_base + T_REL + 32 * (it & 1);
      {
        // every column group extracts 8 of the 32 ids (table slots + the records of the key-centric pass)
        if (hf == 0 && row_ok) {
          rec_row[0] = nm2l;
          rec_row[RSF] = -INFINITY;    // masked elements: every row holds a real maximum, p == 0
          rec_row[2 * RSF] = delta;
          rec_row[3 * RSF] = 0.f;
        }
        mbar_wait_warp(&bars->rel_full, it & 1);
        tc_fence_after_sync();
        uint32_t v[8];
        tmem_ld8(t_rel + 8 * hf, v);
        tmem_wait_ld();
#pragma unroll
        for (int x8 = 0; x8 < 8; ++x8) {
          const int x = 8 * hf + x8;
          const float val = x < R ? fmaf(__uint_as_float(v[x8]), scale2, bias_s[x]) : 0.f;
          rel_s[slot_of_id(x, D) * TM + row] = val;
          if (row_ok && x < p.rw) rec_row[(4 + x) * RSF] = val + nm2l;
        }
      }
      named_bar_sync(1, NEW);   // rel_s visible to both column halves
      if (tid == 0) GTRACE(0, 1);
      const float cP = rel_s[(2 * D) * TM + row] + nm2l;
      const float cN = rel_s[0 * TM + row] + nm2l;
      const float cX = rel_s[(2 * D + 1) * TM + row] + nm2l;
      const float cX1 = rel_s[(2 * D + 2) * TM + row] + nm2l;
      const int smin = __reduce_min_sync(0xffffffffu, q_sent < 0 ? 0x7fffffff : q_sent);
      const int smax = __reduce_max_sync(0xffffffffu, q_sent);
      float accP = 0.f, accN = 0.f, accX = 0.f, accX1 = 0.f;

      const int nc = 2 + q.nglob;
#pragma unroll 1
      for (int c = 0; c < nc; ++c) {
        const bool band = c < 2;
        int ka, kb;
        chunk_blocks(q, c, ka, kb);
        const int g0 = ((hf < 2) ? ka : kb) + 32 * (hf & 1);   // this warp's 32-key group of the chunk
        const int klen = band ? p.L : p.G;
        int ceg;
        {
          const int32_t* eids = (band ? p.long_eid + (int64_t)b * p.L : p.glob_eid + (int64_t)b * p.G);
          const int col = g0 + lane;
          ceg = (col >= 0 && col < klen) ? __ldg(eids + col) : -1;
        }
        if (tid == 0) GTRACE(0, 10 + c);
        mbar_wait_warp(&bars->s_full[c & 1], s_par[c & 1]);
        if (tid == 0) GTRACE(0, 20 + c);
        s_par[c & 1] ^= 1;
        tc_fence_after_sync();
        bool dp_ready = false;
        const uint32_t t_s = t_base + ((c & 1) ? T_S1 : T_S0) + 32 * hf;
        const uint32_t t_dp = t_base + T_DP + 32 * hf;
        {
          constexpr int g = 0;
          const int c0 = __shfl_sync(0xffffffffu, ceg, 0);
          const bool uni = __all_sync(0xffffffffu, ceg == c0) && c0 != -1;
          // ---- classify (warp-uniform), same forms as the forward (gl2_fwd.cu) ----
          int fm;
          if (band) {
            const int G0 = g0 - a_lo;
            const bool dead = G0 > RAD + 31 || G0 < -RAD - 31 || g0 >= p.L || g0 + 31 < 0;
            const bool inside = g0 >= 0 && g0 + 31 < p.L;
            fm = dead ? 0 : 7;
            if (!dead && inside && uni && d12) {
              if (G0 == -64) fm = 2;
              else if (G0 == 64) fm = 3;
              else if (G0 == -32) fm = 4;
              else if (G0 == 0) fm = 5;
              else if (G0 == 32) fm = 6;
            }
          } else {
            const bool dead = g0 >= p.G;
            const bool plain = g0 + 31 < p.G && uni && (smax < g0 || smin > g0 + 31);
            fm = dead ? 0 : (plain ? 1 : 8);
          }
          if (fm != 0 && uni && __all_sync(0xffffffffu, q_e != c0)) fm = 0;   // masked for every row of the warp
          uint32_t pk[16];
          if (fm == 0) {
#pragma unroll
            for (int x = 0; x < 16; ++x) pk[x] = 0u;
            if (!dp_ready) {   // keeps this warp inside the chunk: it must not reach the next hand-over early
              mbar_wait_warp(&bars->dp_full, dp_par);
              tc_fence_after_sync();
              dp_ready = true;
            }
          } else {
            float pe[32];
            {
              uint32_t v[32];
              tmem_ld32(t_s + 32 * g, v);
              tmem_wait_ld();
              // pe[jj] = livef(jj) ? ex2(x * scale2 + addf(jj)) : 0
              auto run = [&](auto addf, auto livef) {
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const float tt = ex2(fmaf(__uint_as_float(v[jj]), scale2, addf(jj)));
                  pe[jj] = livef(jj) ? tt : 0.f;
                }
              };
              const float mk = (uni && q_e != c0) ? -INFINITY : 0.f;   // uniform mask: a masked row evaluates to 0
              auto yes = [](int) { return true; };
              switch (fm) {
                case 1: { const float cc = cX + mk; run([&](int) { return cc; }, yes); } break;
                case 2: { const float cc = cN + mk; run([&](int) { return cc; }, [&](int jj) { return jj >= lane; }); } break;
                case 3: { const float cc = cP + mk; run([&](int) { return cc; }, [&](int jj) { return jj <= lane; }); } break;
                case 4: {
                  const float cc = cN + mk, nk = nm2l + mk;
                  const float* tb = rel_s + (12 - 32 - lane) * TM + row;
                  run([&](int jj) { if (jj < 21) return cc; return (jj > lane + 20) ? tb[jj * TM] + nk : cc; }, yes);
                } break;
                case 5: {
                  const float cp = cP + mk, cn = cN + mk, nk = nm2l + mk;
                  const float* tb = rel_s + (12 - lane) * TM + row;
                  const int t0 = 11 - lane;
                  run([&](int jj) { return ((unsigned)(t0 + jj) < 23u) ? tb[jj * TM] + nk : ((jj > lane) ? cp : cn); }, yes);
                } break;
                case 6: {
                  const float cc = cP + mk, nk = nm2l + mk;
                  const float* tb = rel_s + (12 + 32 - lane) * TM + row;
                  run([&](int jj) { if (jj > 10) return cc; return (jj < lane - 20) ? tb[jj * TM] + nk : cc; }, yes);
                } break;
                case 7: {
                  const int d0 = g0 - i;
                  int jlo = max(0, -RAD - d0), jhi = min(32, RAD - d0 + 1);
                  jlo = max(jlo, -g0);
                  jhi = min(jhi, p.L - g0);
                  const unsigned span = (unsigned)max(jhi - jlo, 0);
                  const float* base = rel_s + row;
                  const int sD = d0 + D;
                  run([&](int jj) { return base[min(max(sD + jj, 0), 2 * D) * TM] + nm2l; },
                      [&](int jj) {   // the shuffle is executed by every lane (no short-circuit in front of it)
                        const bool same = __shfl_sync(0xffffffffu, ceg, jj) == q_e;
                        return same && (unsigned)(jj - jlo) < span;
                      });
                } break;
                default: {
                  const int sp = q_sent - g0;
                  const int jhi = min(32, p.G - g0);
                  run([&](int jj) { return sp == jj ? cX1 : cX; },
                      [&](int jj) {
                        const bool same = __shfl_sync(0xffffffffu, ceg, jj) == q_e;
                        return same && jj < jhi;
                      });
                } break;
              }
            }
            __syncwarp();   // per-lane table reads may leave the warp diverged: tcgen05.ld / st are .aligned
            if (!dp_ready) {
              mbar_wait_warp(&bars->dp_full, dp_par);
              tc_fence_after_sync();
              dp_ready = true;
            }
            uint32_t w[32];
            tmem_ld32(t_dp + 32 * g, w);
            tmem_wait_ld();
            float tot0 = 0.f, tot1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              pe[jj] *= (__uint_as_float(w[jj]) - delta);
              pe[jj + 1] *= (__uint_as_float(w[jj + 1]) - delta);
              tot0 += pe[jj];
              tot1 += pe[jj + 1];
              pk[jj >> 1] = pack_bf16x2(pe[jj], pe[jj + 1]);
            }
            const float tot = tot0 + tot1;
            // ---- relative-id bins: constant classes in registers, every interior diagonal slot belongs to
            // exactly one key of the row and is a plain store ----
            switch (fm) {
              case 1: accX += tot; break;
              case 2: accN += tot; break;
              case 3: accP += tot; break;
              case 4: {
                float* tb = bins + (12 - 32 - lane) * TM + row;
                float sw = 0.f;
#pragma unroll
                for (int jj = 21; jj < 32; ++jj)
                  if (jj > lane + 20) { tb[jj * TM] = pe[jj]; sw += pe[jj]; }
                accN += tot - sw;
              } break;
              case 5: {
                float* tb = bins + (12 - lane) * TM + row;
                const int t0 = 11 - lane;
                float sw = 0.f, sp = 0.f;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const bool win = (unsigned)(t0 + jj) < 23u;
                  if (win) tb[jj * TM] = pe[jj];
                  sw += win ? pe[jj] : 0.f;
                  sp += (!win && jj > lane) ? pe[jj] : 0.f;
                }
                accP += sp;
                accN += tot - sw - sp;
              } break;
              case 6: {
                float* tb = bins + (12 + 32 - lane) * TM + row;
                float sw = 0.f;
#pragma unroll
                for (int jj = 0; jj < 11; ++jj)
                  if (jj < lane - 20) { tb[jj * TM] = pe[jj]; sw += pe[jj]; }
                accP += tot - sw;
              } break;
              case 7: {
                const int sD = g0 - i + D;
                float* base = bins + row;
                float sn = 0.f, sp = 0.f;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const int s = sD + jj;                 // dead elements carry ds == 0
                  if (s > 0 && s < 2 * D) base[s * TM] = pe[jj];
                  sn += (s <= 0) ? pe[jj] : 0.f;
                  sp += (s >= 2 * D) ? pe[jj] : 0.f;
                }
                accN += sn;
                accP += sp;
              } break;
              default: {
                const int sp = q_sent - g0;
                float s1 = 0.f;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) s1 += (sp == jj) ? pe[jj] : 0.f;
                accX1 += s1;
                accX += tot - s1;
              } break;
            }
          }
          __syncwarp();   // the per-lane bin stores may leave the warp diverged
          tmem_st16(t_s + 32 * g, pk);   // dS over the group's own (spent) scores
        }
        dp_par ^= 1;
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->ds_full);
        if (tid == 0) GTRACE(0, 30 + c);
      }
      // ---- epilogue ----
      // class sums: the second column half hands its partial sums to the first
      if (hf > 0) {
        float* cm = cmb + (hf - 1) * 4 * TM;
        cm[0 * TM + row] = accP;
        cm[1 * TM + row] = accN;
        cm[2 * TM + row] = accX;
        cm[3 * TM + row] = accX1;
      }
      named_bar_sync(1, NEW);
      if (hf == 0) {
        auto sum3 = [&](int k, float a0) {   // fixed order: deterministic
          return ((a0 + cmb[k * TM + row]) + cmb[(4 + k) * TM + row]) + cmb[(8 + k) * TM + row];
        };
        bins[(2 * D) * TM + row] = sum3(0, accP);       // id D    (offset >= D)
        bins[0 * TM + row] = sum3(1, accN);             // id 2D   (offset <= -D)
        bins[(2 * D + 1) * TM + row] = sum3(2, accX);
        bins[(2 * D + 2) * TM + row] = sum3(3, accX1);
      }
      named_bar_sync(1, NEW);   // bins complete
      if (tid == 0) GTRACE(0, 40);
      const int4 te_next = sched_take<true>(bars, ring, (int)it + 1, lane == 0);
      const RowPre pre_next = row_pre(te_next);
      {
        // each of the four column groups packs 8 ids: bf16 A operand (TMEM) for dQ += dallrel.E, the
        // dallrel^T tile for the table-gradient MMA, and the bias partial sums of the warp's 32 rows
        const int c0 = 8 * hf;
        float w16[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) {
          const int pid = c0 + x;
          w16[x] = (pid < R && row_ok) ? bins[slot_of_id(pid, D) * TM + row] : 0.f;
        }
        uint32_t pk4[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) pk4[x] = pack_bf16x2(w16[2 * x], w16[2 * x + 1]);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_rel + c0 / 2),
                     "r"(pk4[0]), "r"(pk4[1]), "r"(pk4[2]), "r"(pk4[3])
                     : "memory");
        *reinterpret_cast<uint4*>(a_tile + row * 128 + (((c0 >> 3) ^ (row & 7)) << 4)) =
            make_uint4(pk4[0], pk4[1], pk4[2], pk4[3]);
        {
          // bias partial sums of the warp's 32 rows, 8 ids: halving exchange (8 -> 4 -> 2 -> 1 values per lane over
          // lane bits 16, 8, 4), then two plain steps: 9 shuffles instead of 40, fixed order
          const bool b16 = lane & 16, b8 = lane & 8, b4 = lane & 4;
          float a4[4], a2[2];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const float keep = b16 ? w16[x + 4] : w16[x], give = b16 ? w16[x] : w16[x + 4];
            a4[x] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
          }
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            const float keep = b8 ? a4[x + 2] : a4[x], give = b8 ? a4[x] : a4[x + 2];
            a2[x] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
          }
          const float keep = b4 ? a2[1] : a2[0], give = b4 ? a2[0] : a2[1];
          float r = keep + __shfl_xor_sync(0xffffffffu, give, 4);
          r += __shfl_xor_sync(0xffffffffu, r, 2);
          r += __shfl_xor_sync(0xffffffffu, r, 1);
          if ((lane & 3) == 0) bs[quad * 32 + c0 + (b16 ? 4 : 0) + (b8 ? 2 : 0) + (b4 ? 1 : 0)] = r;
        }
        tmem_wait_st();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        mbar_arrive(&bars->dar_full);
      }
      if (tid == 0) GTRACE(0, 41);
      mbar_wait_warp(&bars->dq_full, it & 1);
      if (tid == 0) GTRACE(0, 42);
      tc_fence_after_sync();
      const int64_t pidx = ((int64_t)(b * p.tiles_per_bh + q.tile) * p.H + h);
      uint32_t vq[16];
      {
        // table-gradient tile: M = 64 accumulator rows live in lanes {0-15, 32-47, 64-79, 96-111}; each column
        // group stores 16 of the 64 columns
        const int pid = quad * 16 + lane;
        uint32_t v[16];
        tmem_ld16(t_base + T_DE + hf * 16, v);
        tmem_ld16(t_base + T_DQ + 16 * hf, vq);   // both read-outs in flight together
        tmem_wait_ld();
        if (tid == 0) GTRACE(0, 44);
        if (lane < 16 && pid < R) {
          float4* dst = reinterpret_cast<float4*>(p.tg_partial + (pidx * R + pid) * 64 + hf * 16);
#pragma unroll
          for (int x = 0; x < 4; ++x)
            dst[x] = make_float4(__uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]),
                                 __uint_as_float(v[4 * x + 2]), __uint_as_float(v[4 * x + 3]));
        }
      }
      if (tid == 0) GTRACE(0, 45);
      named_bar_sync(1, NEW);   // bias sums of all four quadrants are in shared memory
      if (tid == 0) GTRACE(0, 46);
      if (tid < R) p.tg_partial_bias[pidx * R + tid] = (bs[tid] + bs[32 + tid]) + (bs[64 + tid] + bs[96 + tid]);
      {
        const uint32_t (&v)[16] = vq;
        if (tid == 0) GTRACE(0, 47);
        if (row_ok) {
          __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(p.d_q, b, i, h) + 16 * hf;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(v[8 * x + 0]) * p.scale, __uint_as_float(v[8 * x + 1]) * p.scale);
            o4.y = pack_bf16x2(__uint_as_float(v[8 * x + 2]) * p.scale, __uint_as_float(v[8 * x + 3]) * p.scale);
            o4.z = pack_bf16x2(__uint_as_float(v[8 * x + 4]) * p.scale, __uint_as_float(v[8 * x + 5]) * p.scale);
            o4.w = pack_bf16x2(__uint_as_float(v[8 * x + 6]) * p.scale, __uint_as_float(v[8 * x + 7]) * p.scale);
            *reinterpret_cast<uint4*>(dst + 8 * x) = o4;
          }
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->tile_done);
      if (tid == 0) GTRACE(0, 43);
      te = te_next;
      pre = pre_next;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NEWARPS + 1) tmem_dealloc<512>(tmem);
  if (tid == 0 && atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {   // every CTA has taken its last tile
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

}  // namespace bq
}  // namespace gl2

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace_gl2q(long long* out, int* n) {
  cudaMemcpyFromSymbol(n, gl2::bq::g_trace_n, sizeof(int) * 3);
  return (int)cudaMemcpyFromSymbol(out, gl2::bq::g_trace_q, sizeof(long long) * 3 * 1024);
}
#endif

bool gl2_bwd_q_long_supported(const BwdQArgs& a, int dtype, int d) {
  FwdArgs f{};
  f.rows = a.rows;
  f.seg[0] = a.seg[0];
  f.seg[1] = a.seg[1];
  f.nseg = a.nseg;
  f.out = a.out;
  f.B = a.B;
  f.H = a.H;
  f.neg = a.neg;
  f.drop = a.drop;
  return a.tg_partial != nullptr && gl2_fwd_long_supported(f, dtype, d);
}

int gl2_launch_bwd_q_long(const BwdQArgs& a, const float4* rowstat, float* rec_ws, int lp, int rw, cudaStream_t st) {
  using namespace gl2;
  static PerDeviceOnce once;
  static int sm_count[64];
  static unsigned* sched_base[64];
  static std::atomic<unsigned> sched_slot{0};
  const int ae = once.run([] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaFuncSetAttribute(bq::gl2_bwd_q_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         bq::SM_ALLOC);
    if (e == cudaSuccess) e = cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base[dev]), bq::g_sched_bq);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    return (int)e;
  });
  if (ae) return ae;
  int dev = 0;
  cudaGetDevice(&dev);
  bq::Params p{};
  p.B = a.B; p.H = a.H; p.L = a.rows.len; p.G = a.seg[1].len; p.R = a.rows.R;
  p.D = a.seg[0].side.max_distance;
  p.scale = a.scale;
  p.long_eid = a.seg[0].side.q_eid;
  p.glob_eid = a.seg[1].side.k_eid;
  p.sent = a.seg[1].side.sent;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.rows.bias);
  p.d_q = a.d_q;
  p.rowstat = rowstat;
  p.rec_ws = rec_ws;
  p.tg_partial = a.tg_partial;
  p.tg_partial_bias = a.tg_partial_bias;
  p.lp = lp;
  p.rw = rw;
  p.tiles_per_bh = (p.L + TM - 1) / TM;
  p.total_tiles = p.tiles_per_bh * a.B * a.H;
  p.sched = sched_base[dev] + 2 * (sched_slot.fetch_add(1, std::memory_order_relaxed) % 256);
  CUtensorMap mq, mdo, mk, mv, mgk, mgv, me;
  int e = 0;
  e |= make_qkv_tensor_map(&mq, a.rows.q.ptr, a.rows.q.sb, a.rows.q.sl, a.rows.q.sh, a.B, p.L, a.H, TM);
  e |= make_qkv_tensor_map(&mdo, a.d_out.ptr, a.d_out.sb, a.d_out.sl, a.d_out.sh, a.B, p.L, a.H, TM);
  e |= make_qkv_tensor_map(&mk, a.seg[0].k.ptr, a.seg[0].k.sb, a.seg[0].k.sl, a.seg[0].k.sh, a.B, p.L, a.H, 64);
  e |= make_qkv_tensor_map(&mv, a.seg[0].v.ptr, a.seg[0].v.sb, a.seg[0].v.sl, a.seg[0].v.sh, a.B, p.L, a.H, 64);
  e |= make_qkv_tensor_map(&mgk, a.seg[1].k.ptr, a.seg[1].k.sb, a.seg[1].k.sl, a.seg[1].k.sh, a.B, p.G, a.H, 64);
  e |= make_qkv_tensor_map(&mgv, a.seg[1].v.ptr, a.seg[1].v.sb, a.seg[1].v.sl, a.seg[1].v.sh, a.B, p.G, a.H, 64);
  e |= make_qkv_tensor_map(&me, a.rows.emb, (int64_t)p.R * a.H * 64, (int64_t)a.H * 64, 64, 1, p.R, a.H, 32);
  if (e) return MLT_ERR_UNSUPPORTED;
  const int grid = p.total_tiles < sm_count[dev] ? p.total_tiles : sm_count[dev];
  bq::gl2_bwd_q_long_kernel<<<grid, bq::NTHREADS, bq::SM_ALLOC, st>>>(mq, mdo, mk, mv, mgk, mgv, me, p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
