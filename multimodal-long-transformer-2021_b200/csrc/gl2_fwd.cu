// gl2: persistent forward kernel for the long rows of compact global-local attention.
//
// Specialisation of the north-star path (FusedGlobalLocalAttention long rows with masks / ids built
// in-kernel from O(L + G) descriptors): bf16, d = 64, example-id masks, 1-D ids in the band, sentence
// cross ids towards the global tokens, relative vocabulary <= 32.  Everything else keeps the general
// kernels of tc_fwd.cu.
//
// Structure (one CTA per SM, 576 threads, persistent over PAIRS of adjacent 128-row query tiles):
//   warps 0-7   softmax group 0: tile A of the pair.  Warp w owns row quadrant (w & 3) == TMEM lanes and the
//               column half ((w >> 2) & 1) of every 128-key chunk: two threads per row.  (With one thread
//               per row the kernel ran at 36 % issue utilisation on ~10 warps per SM: per-warp latency, not
//               issue slots, was the limit.)  The two threads of a row agree on the softmax reference through
//               one shared-memory exchange per chunk.
//   warps 8-15  softmax group 1: tile B of the pair
//   warp 16     TMA producer: Q tiles + relative-embedding tile of the NEXT pair while the current one
//               computes, K/V chunks of 128 keys through a 3-stage ring shared by both tiles (the band
//               chunks of the two tiles overlap; the global-token chunks are common)
//   warp 17     MMA issuer (one elected lane): per tile S = Q.K_c^T (M 128, N 128, K 64) and
//               O += P_c.V_c (P from TMEM), the two tiles ping-pong through the tensor core so that
//               one warpgroup's softmax covers the other's MMAs
// TMEM (512 columns): per warpgroup S / P [128] + O [64] + allrel [32].
//
// Softmax: per 128-key chunk, pass 1 takes the maximum of the raw accumulator over the chunk's live
// groups -- an UPPER bound of the row maximum (U = xmax * scale + max rel) is all a softmax reference
// needs -- and pass 2 evaluates p = ex2(x * scale*log2e + c_row - m) per 32-key group with the form the
// group's geometry calls for.  m only moves when the bound outgrows it by 2^8 (then O is rescaled in
// TMEM).  Every valid row sees its own diagonal key (same example id), so a masked key has weight
// exactly 0 (exp(-1e9 + ...) flushes): masked rows / groups are simply skipped.
#include <atomic>
#include "tc_api.cuh"

#include "gl2_geom.cuh"
#include "mlt_common.cuh"
#include "profile.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace gl2 {

using namespace ptx;

constexpr int NSW = 16;                   // softmax warps: 2 tile slots x 4 row quadrants x 2 column halves
constexpr int NTHREADS = (NSW + 2) * 32;
constexpr float GROW_LOG2 = 8.f; // lazy-rescale threshold (log2 units)

constexpr int SM_Q = 0;                               // [2 bufs][2 tiles] x 16 KB
constexpr int SM_E = SM_Q + 4 * TM * 128;             // [2 bufs] x 4 KB (32 ids x 128 B)
constexpr int SM_KV = SM_E + 2 * 32 * 128;            // NST x (K 16 KB + V 16 KB)
constexpr int SM_REL = SM_KV + NST * 2 * TK * 128;    // [2 tiles][32 slots][128 rows] f32 (log2 units)
constexpr int SM_BIAS = SM_REL + 2 * 32 * TM * 4;     // [2 tiles][32] f32
constexpr int SM_XCH = SM_BIAS + 2 * 32 * 4;          // [2 parities][2 tiles][2 halves][128 rows] f32: pair exchange
constexpr int SM_BAR = SM_XCH + 2 * 2 * 2 * TM * 4;
constexpr int SM_SCHED = SM_BAR + 256;                // ring of pair descriptors {pair, b, h, i0} (16 B each)
constexpr int SM_TOTAL = SM_SCHED + 64;
constexpr int SM_ALLOC = SM_TOTAL + 1024;

constexpr uint32_t T_WG = 224;   // TMEM columns per warpgroup
constexpr uint32_t T_S = 0, T_O = 128, T_REL = 192;

#ifdef MLT_TC_TRACE
// cheap timeline: per-role event log in registers / local arrays, flushed once at kernel end (CTA 3 only)
__device__ long long g_trace_f[2][2048];
__device__ int g_trace_fn[2];
#define FTRACE(buf, n, code) do { if (n < 1000) { buf[2 * n] = clock64(); buf[2 * n + 1] = (code); ++n; } } while (0)
#else
#define FTRACE(buf, n, code) do {} while (0)
#endif

struct Params {
  int B, H, L, G, R, D, radius;
  float scale;
  const int32_t* long_eid;    // [B, L]
  const int32_t* glob_eid;    // [B, G]
  const int32_t* sent;        // [B, L] sentence (global token) of each long token
  const __nv_bfloat16* bias;  // [R, H]
  T4 out;
  float* stats;               // [B, H, L, 2]
  int pairs_per_bh, total_pairs;
  unsigned* sched;            // {next pair after the first gridDim.x, CTAs done}: both 0 at launch, reset by the last CTA
};

// Work distribution: pair number blockIdx.x first, then whatever pair is next when the CTA is ready for one
// (one atomic per pair).  CTAs that start late -- the SM was still running another kernel's blocks, as with
// the global-row kernel that is launched beside this one -- take fewer pairs instead of finishing late.
// The producer thread fetches the pair and publishes it to the other roles through a small ring.
constexpr int NSQ = 4;
__device__ unsigned g_sched_fwd[256][2];

struct Bars {
  uint64_t q_full[2], q_empty[2];
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t rel_full[2], s_full[2], p_full[2], o_full[2], o_empty[2];
  uint64_t sched_full[NSQ], sched_empty[NSQ];
  uint32_t tmem_base;
};
// consumer side of the pair ring: a whole (converged) warp, one arrival -- or a single elected thread.  The
// producer thread decodes the pair number once (two integer divisions) and publishes {pair, b, h, i0}.
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

// Geometry of one pair of tiles, identical in every role.  The kernel is specialised for
// local_radius == 64: a tile's band spans the four 64-key blocks k0..k3 = [i0 - 64, i0 + 192).  A tile's
// two band chunks INTERLEAVE them, (k0, k2) and (k1, k3): with contiguous chunks the four row quadrants
// of a tile carry (4, 3, 2, 1) and (1, 2, 3, 4) live 32-key groups and every chunk waits for its
// slowest quadrant (8 group-times for 5 groups of work); interleaved it is (3, 3, 2, 2) / (2, 2, 3, 3).
// Chunk order of a pair: A(k0,k2), B(k0,k2), A(k1,k3), B(k1,k3), then the global-token chunks (shared).
struct Pair {
  int b, h, i0;        // first row of tile A
  int nglob;           // global-token chunks
  bool has[2];         // tile exists
};
constexpr int NBAND = 4;   // band chunks of a pair
__device__ __forceinline__ Pair make_pair(const Params& p, int pair) {
  Pair q;
  const int bh = pair / p.pairs_per_bh;
  q.b = bh / p.H;
  q.h = bh - q.b * p.H;
  q.i0 = (pair - bh * p.pairs_per_bh) * (2 * TM);
  q.nglob = (p.G + TK - 1) / TK;
  q.has[0] = true;
  q.has[1] = q.i0 + TM < p.L;
  return q;
}
__device__ __forceinline__ Pair pair_of(const Params& p, const int4& e) {   // from a published descriptor
  Pair q;
  q.b = e.y;
  q.h = e.z;
  q.i0 = e.w;
  q.nglob = (p.G + TK - 1) / TK;
  q.has[0] = true;
  q.has[1] = q.i0 + TM < p.L;
  return q;
}
__device__ __forceinline__ bool tile_exists(const Pair& q, int w) { return q.has[w]; }
__device__ __forceinline__ bool uses(const Pair& q, int w, int pc) {
  if (!q.has[w]) return false;
  return pc >= NBAND || (pc & 1) == w;
}
// first keys of the two 64-key blocks of chunk pc
__device__ __forceinline__ void chunk_blocks(const Pair& q, int pc, int& ka, int& kb) {
  if (pc < NBAND) {
    const int t0 = q.i0 + (pc & 1) * TM;          // first row of the tile that uses the chunk
    ka = t0 - RAD + (pc >> 1) * 64;               // k0 or k1
    kb = ka + 128;                                // k2 or k3
  } else {
    ka = (pc - NBAND) * TK;
    kb = ka + 64;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
gl2_fwd_long_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_gk,
                    const __grid_constant__ CUtensorMap map_gv, const __grid_constant__ CUtensorMap map_e,
                    const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int4* ring = reinterpret_cast<int4*>(smem + SM_SCHED);

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1);
      mbar_init(&bars->q_empty[s], 1);
      mbar_init(&bars->rel_full[s], 1);
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 256);
      mbar_init(&bars->o_full[s], 1);
      mbar_init(&bars->o_empty[s], 256);
    }
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->kv_full[s], 1);
      mbar_init(&bars->kv_empty[s], 2);   // two consumers (tiles); the MMA thread arrives for an absent one
    }
    for (int s = 0; s < NSQ; ++s) {
      mbar_init(&bars->sched_full[s], 1);
      mbar_init(&bars->sched_empty[s], NSW + 1);   // the softmax warps and the MMA thread
    }
    fence_barrier_init();
  }
  if (warp == NSW + 1) tmem_alloc<512>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (warp == NSW) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      prefetch_tensormap(&map_q);
      prefetch_tensormap(&map_k);
      prefetch_tensormap(&map_v);
      prefetch_tensormap(&map_gk);
      prefetch_tensormap(&map_gv);
      prefetch_tensormap(&map_e);
      int kvc = 0;
      for (int it = 0;; ++it) {
        const int buf = it & 1;
        mbar_wait(&bars->q_empty[buf], ((it >> 1) & 1) ^ 1);   // taken as late as possible: no pair is held back
        int pair = it == 0 ? (int)blockIdx.x : (int)(gridDim.x + atomicAdd(p.sched, 1u));
        if (pair >= p.total_pairs) pair = -1;
        const int sq = it % NSQ;
        mbar_wait(&bars->sched_empty[sq], ((it / NSQ) & 1) ^ 1);
        Pair q{};
        if (pair >= 0) q = make_pair(p, pair);
        ring[sq] = make_int4(pair, q.b, q.h, q.i0);
        mbar_arrive(&bars->sched_full[sq]);
        if (pair < 0) break;
        mbar_arrive_expect_tx(&bars->q_full[buf], 2 * TM * 128 + 32 * 128);
        tma_load_4d(smem + SM_Q + (buf * 2 + 0) * TM * 128, &map_q, &bars->q_full[buf], 0, q.i0, q.h, q.b);
        tma_load_4d(smem + SM_Q + (buf * 2 + 1) * TM * 128, &map_q, &bars->q_full[buf], 0, q.i0 + TM, q.h, q.b);
        tma_load_4d(smem + SM_E + buf * 32 * 128, &map_e, &bars->q_full[buf], 0, 0, q.h, 0);
        const int npc = NBAND + q.nglob;
        for (int pc = 0; pc < npc; ++pc, ++kvc) {
          const int st = kvc % NST;
          mbar_wait(&bars->kv_empty[st], ((kvc / NST) & 1) ^ 1);
          uint8_t* ks = smem + SM_KV + st * (2 * TK * 128);
          uint8_t* vs = ks + TK * 128;
          mbar_arrive_expect_tx(&bars->kv_full[st], 2 * TK * 128);
          int ka, kb;
          chunk_blocks(q, pc, ka, kb);
          const CUtensorMap* mk = pc < NBAND ? &map_k : &map_gk;
          const CUtensorMap* mv = pc < NBAND ? &map_v : &map_gv;
          // A box that lies entirely outside the tensor is not issued as such: its start is clamped so that
          // at least one row is in range (the kernel's own geometry marks those keys dead, and what lands
          // in shared memory is finite either way: zero fill or real rows)
          const int klen = (pc < NBAND) ? p.L : p.G;
          const int ca = min(max(ka, -63), klen - 1), cb = min(max(kb, -63), klen - 1);
          tma_load_4d(ks, mk, &bars->kv_full[st], 0, ca, q.h, q.b);              // boxes of 64 keys
          tma_load_4d(ks + 64 * 128, mk, &bars->kv_full[st], 0, cb, q.h, q.b);
          tma_load_4d(vs, mv, &bars->kv_full[st], 0, ca, q.h, q.b);
          tma_load_4d(vs + 64 * 128, mv, &bars->kv_full[st], 0, cb, q.h, q.b);
        }
      }
    }
  } else if (warp == NSW + 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TK, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t idesc_r = make_idesc_bf16(TM, 32, 0, 0);
      int it = 0, kv_base = 0;
#ifdef MLT_TC_TRACE
      long long tb1[2048]; int tn1 = 0;
#endif
      uint32_t s_cnt[2] = {0, 0};    // S MMAs issued per tile slot (parity of s_full is kept by the softmax side)
      uint32_t p_cnt[2] = {0, 0};    // P chunks consumed per tile slot
      uint32_t tile_cnt[2] = {0, 0}; // tiles started per slot
      for (;; ++it) {
        const int4 pe = sched_take<false>(bars, ring, it, true);
        if (pe.x < 0) break;
        const Pair q = pair_of(p, pe);
        const int buf = it & 1;
        const int npc = NBAND + q.nglob;
        mbar_wait(&bars->q_full[buf], (it >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t e_addr = smem_u32(smem + SM_E + buf * 32 * 128);
        uint32_t q_addr[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          q_addr[w] = smem_u32(smem + SM_Q + (buf * 2 + w) * TM * 128);
          if (tile_exists(q, w)) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_ss(tmem + w * T_WG + T_REL, sdesc(q_addr[w]).at(kk * 32), sdesc(e_addr).at(kk * 32), idesc_r, kk > 0);
            umma_commit(&bars->rel_full[w]);
          }
        }
        auto stage_of = [&](int pc) { return (kv_base + pc) % NST; };
        auto issue_s = [&](int w, int pc) {
          const int st = stage_of(pc);
          mbar_wait(&bars->kv_full[st], ((kv_base + pc) / NST) & 1);
          tc_fence_after_sync();
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TK * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + w * T_WG + T_S, sdesc(q_addr[w]).at(kk * 32), sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
          umma_commit(&bars->s_full[w]);
          ++s_cnt[w];
        };
        auto next_chunk = [&](int w, int pc) {   // next chunk after pc used by tile w, or npc
          int n = pc + 1;
          while (n < npc && !uses(q, w, n)) ++n;
          return n;
        };
        // prime: first S of each tile
        int first[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          first[w] = next_chunk(w, -1);
          if (first[w] < npc) issue_s(w, first[w]);
        }
        for (int pc = 0; pc < npc; ++pc) {
          const int st = stage_of(pc);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            if (!uses(q, w, pc)) {
              mbar_arrive(&bars->kv_empty[st]);   // absent consumer
              continue;
            }
            FTRACE(tb1, tn1, 100 + 10 * w + pc);
            mbar_wait(&bars->p_full[w], p_cnt[w] & 1);
            FTRACE(tb1, tn1, 200 + 10 * w + pc);
            ++p_cnt[w];
            tc_fence_after_sync();
            if (pc == first[w]) {   // O of the previous tile in this slot has been read out
              mbar_wait(&bars->o_empty[w], (tile_cnt[w] & 1) ^ 1);
              ++tile_cnt[w];
              tc_fence_after_sync();
            }
            const uint32_t v_addr = smem_u32(smem + SM_KV + st * (2 * TK * 128) + TK * 128);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts(tmem + w * T_WG + T_O, tmem + w * T_WG + T_S + (kk >> 1) * 32 + (kk & 1) * 8,
                      sdesc(v_addr).at(kk * 2048), idesc_o, (pc != first[w] || kk > 0));
            umma_commit(&bars->kv_empty[st]);
            const int nx = next_chunk(w, pc);
            if (nx < npc) issue_s(w, nx);
            else umma_commit(&bars->o_full[w]);
          }
        }
        umma_commit(&bars->q_empty[buf]);   // every MMA that reads this pair's Q / E has been issued
        FTRACE(tb1, tn1, 300);
        kv_base += npc;
      }
#ifdef MLT_TC_TRACE
      if (blockIdx.x == 3) { for (int k = 0; k < 2 * tn1; ++k) g_trace_f[1][k] = tb1[k]; g_trace_fn[1] = tn1; }
#endif
    }
  } else {
    // ===================== softmax groups =====================
    const int w = warp >> 3;                 // tile slot
    const int hf = (warp >> 2) & 1;          // column half of every chunk
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const uint32_t t_base = tmem + w * T_WG + lane_sel;
    float* rel_s = reinterpret_cast<float*>(smem + SM_REL) + w * 32 * TM;   // [slot][row], log2 units
    float* bias_s = reinterpret_cast<float*>(smem + SM_BIAS) + w * 32;
    float* xch = reinterpret_cast<float*>(smem + SM_XCH) + w * 2 * TM;      // [parity][tile][half][row]
    const uint32_t bar_group = 1 + w;                  // 256 threads of the tile slot
    const uint32_t bar_pair = 3 + w * 4 + quad;        // the two warps that share 32 rows
    const float scale2 = p.scale * LOG2E;
    const int D = p.D, R = p.R, rad = p.radius;
    uint32_t s_par = 0, tile_par = 0, xpar = 0;   // parities of s_full / per-tile barriers / exchange buffer
#ifdef MLT_TC_TRACE
    long long tb0[2048]; int tn0 = 0;
    const bool tr = (tid == 0);
#define ST(code) do { if (tr) FTRACE(tb0, tn0, code); } while (0)
#else
#define ST(code) do {} while (0)
#endif
    for (int sit = 0;; ++sit) {
      const int4 pe = sched_take<true>(bars, ring, sit, lane == 0);
      if (pe.x < 0) break;
      const Pair q = pair_of(p, pe);
      if (!tile_exists(q, w)) continue;
      const int b = q.b, h = q.h;
      const int ti0 = q.i0 + w * TM;         // first row of this tile
      const int i = ti0 + row;
      const bool row_ok = i < p.L;
      const int a_lo = ti0 + quad * 32;      // first row of the warp
      int q_e = -2, q_sent = -1;
      if (row_ok) {
        q_e = __ldg(p.long_eid + (int64_t)b * p.L + i);
        q_sent = __ldg(p.sent + (int64_t)b * p.L + i);
      }
      // ---- per-tile relative table: rel_s[slot][row] = (allrel[id] * scale + bias[id] * scale) * log2e;
      // each column half builds 16 of the 32 ids
      if (hf == 0 && row < 32) bias_s[row] = row < R ? __bfloat162float(p.bias[row * p.H + h]) * scale2 : 0.f;
      named_bar_sync(bar_group, 256);
      ST(1);
      mbar_wait_warp(&bars->rel_full[w], tile_par);
      ST(2);
      tc_fence_after_sync();
      float relmax = 0.f;   // ids outside [0, R) contribute 0, and so may any key: the bound includes 0
      {
        uint32_t v[32];
        tmem_ld32(t_base + T_REL, v);
        tmem_wait_ld();
#pragma unroll
        for (int x = 0; x < 32; ++x) {
          const float val = x < R ? fmaf(__uint_as_float(v[x]), scale2, bias_s[x]) : 0.f;
          if ((x >> 4) == hf) rel_s[slot_of_id(x, D) * TM + row] = val;
          relmax = fmaxf(relmax, val);
        }
      }
      named_bar_sync(bar_group, 256);   // both halves of the table are visible
      const float cP = rel_s[(2 * D) * TM + row];        // offset >= D   (id D      -> slot 2D)
      const float cN = rel_s[0 * TM + row];              // offset <= -D  (id 2D     -> slot 0)
      const float cX = rel_s[(2 * D + 1) * TM + row];    // cross, other sentence
      const float cX1 = rel_s[(2 * D + 2) * TM + row];   // cross, own sentence
      // sentence range of the warp's rows (which global-token groups hold a special column)
      const int smin = __reduce_min_sync(0xffffffffu, q_sent < 0 ? 0x7fffffff : q_sent);
      const int smax = __reduce_max_sync(0xffffffffu, q_sent);

      float m2 = -INFINITY;   // reference maximum, log2 units (an upper bound of the row maximum so far)
      float l = 0.f;
      const int npc = NBAND + q.nglob;
      const bool d12 = (D == 12);   // the specialised band forms are written for relative_pos_max_distance 12
#pragma unroll 1
      for (int pc = 0; pc < npc; ++pc) {
        if (!uses(q, w, pc)) continue;
        const bool band = pc < NBAND;
        int ka, kb;
        chunk_blocks(q, pc, ka, kb);
        const int klen = band ? p.L : p.G;
        // column example ids, one per lane and group (issued before the wait: latency hidden)
        const int kbase = hf ? kb : ka;   // this warp's 64-key block of the chunk
        int ce[2];
        {
          const int32_t* eids = (band ? p.long_eid + (int64_t)b * p.L : p.glob_eid + (int64_t)b * p.G);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int col = kbase + 32 * g + lane;
            ce[g] = (col >= 0 && col < klen) ? __ldg(eids + col) : -1;
          }
        }
        ST(10 + pc);
        mbar_wait_warp(&bars->s_full[w], s_par);
        ST(20 + pc);
        s_par ^= 1;
        tc_fence_after_sync();
        // ---- classify the four 32-key groups (warp-uniform) ----
        // form: 0 dead | 1 FAST: all live, uniform mask, constant relative term | band, mask uniform, group
        // inside the sequence, by G0 = first key - first row of the warp: 2 (G0 = -64: lower band edge),
        // 3 (+64: upper band edge), 4 (-32), 5 (0: the diagonal), 6 (+32) | 7 general band | 8 general global
        int form[2];
        int ce0[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int g0 = kbase + 32 * g;
          ce0[g] = __shfl_sync(0xffffffffu, ce[g], 0);
          const bool uni = __all_sync(0xffffffffu, ce[g] == ce0[g]) && ce0[g] != -1;
          int f;
          if (band) {
            const int G0 = g0 - a_lo;
            const bool dead = G0 > RAD + 31 || G0 < -RAD - 31 || g0 >= p.L || g0 + 31 < 0;
            const bool inside = g0 >= 0 && g0 + 31 < p.L;
            f = dead ? 0 : 7;
            if (!dead && inside && uni && d12) {
              if (G0 == -64) f = 2;
              else if (G0 == 64) f = 3;
              else if (G0 == -32) f = 4;
              else if (G0 == 0) f = 5;
              else if (G0 == 32) f = 6;
            }
          } else {
            const bool dead = g0 >= p.G;
            const bool plain = g0 + 31 < p.G && uni && (smax < g0 || smin > g0 + 31);
            f = dead ? 0 : (plain ? 1 : 8);
          }
          if (f != 0 && uni && __all_sync(0xffffffffu, q_e != ce0[g])) f = 0;   // masked for every row of the warp
          if (f == 7 || f == 8) f |= uni ? 0 : 16;                              // bit 4: per-key mask
          form[g] = f;
        }
        // ---- pass 1: upper bound of the chunk's row maximum from the raw accumulator ----
        float xmax = -INFINITY;
        const uint32_t t_s = t_base + T_S + 64 * hf;
        if (form[0] != 0 || form[1] != 0) {
          uint32_t v[64];
          tmem_ld32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_wait_ld();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (form[g] == 0) continue;
            float r4[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              const uint32_t* u = v + 32 * g + 8 * y;
              r4[y] = fmaxf(fmaxf(__uint_as_float(u[0]), __uint_as_float(u[1])), __uint_as_float(u[2]));
              r4[y] = fmaxf(fmaxf(r4[y], __uint_as_float(u[3])), __uint_as_float(u[4]));
              r4[y] = fmaxf(fmaxf(r4[y], __uint_as_float(u[5])), __uint_as_float(u[6]));
              r4[y] = fmaxf(r4[y], __uint_as_float(u[7]));
            }
            const float gm = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
            // a row masked against the whole (uniform-mask) group contributes nothing
            if (!(!(form[g] & 16) && q_e != ce0[g])) xmax = fmaxf(xmax, gm);
          }
        }
        // the two threads of a row exchange their bounds: one reference for the whole row
        float* xb = xch + xpar * (2 * 2 * TM);
        xpar ^= 1;
        const float mine = fmaf(xmax, scale2, relmax);
        xb[hf * TM + row] = mine;
        ST(30 + pc);
        named_bar_sync(bar_pair, 64);
        ST(40 + pc);
        const float u2 = fmaxf(mine, xb[(hf ^ 1) * TM + row]);   // -inf when nothing is live for this row
        // ---- reference update, lazy rescale of O ----
        const bool fresh = (m2 == -INFINITY);
        const bool grow = !fresh && (u2 - m2 > GROW_LOG2);
        if (fresh) m2 = u2;
        if (__any_sync(0xffffffffu, grow)) {
          // S_c complete implies P.V of the previous chunk complete (MMAs of one thread execute in order)
          const float f = grow ? ex2(m2 - u2) : 1.f;
          if (grow) m2 = u2;
          l *= f;
          {   // each half rescales its 32 columns of O
            uint32_t v[32];
            tmem_ld32(t_base + T_O + hf * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int x = 0; x < 32; ++x) v[x] = __float_as_uint(__uint_as_float(v[x]) * f);
            tmem_st32(t_base + T_O + hf * 32, v);
          }
          tmem_wait_st();
        }
        // ---- pass 2: probabilities, row sum, P (bf16) back into TMEM ----
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int fm = g == 0 ? form[0] : form[1];
          const int c0 = g == 0 ? ce0[0] : ce0[1];
          const int ceg = g == 0 ? ce[0] : ce[1];
          const int g0 = kbase + 32 * g;
          uint32_t pk[16];
          if (fm == 0) {
#pragma unroll
            for (int x = 0; x < 16; ++x) pk[x] = 0u;
          } else {
            uint32_t v[32];
            tmem_ld32(t_s + 32 * g, v);
            tmem_wait_ld();
            // p = ex2(x * scale2 + addf(jj)) where livef(jj), else 0: four elements per step
            auto run = [&](auto addf, auto livef) {
#pragma unroll
              for (int x = 0; x < 16; x += 2) {
                float e4[4];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                  const int jj = 2 * x + y;
                  const float t = ex2(fmaf(__uint_as_float(v[jj]), scale2, addf(jj)));
                  e4[y] = livef(jj) ? t : 0.f;
                }
                ls0 += e4[0]; ls1 += e4[1]; ls2 += e4[2]; ls3 += e4[3];
                pk[x] = pack_bf16x2(e4[0], e4[1]);
                pk[x + 1] = pack_bf16x2(e4[2], e4[3]);
              }
            };
            const float nm = (q_e != c0) ? -INFINITY : -m2;   // uniform-mask forms: a masked row evaluates to 0
            auto yes = [](int) { return true; };
            switch (fm) {
              case 1: {   // global tokens, other sentences
                const float c = cX + nm;
                run([&](int) { return c; }, yes);
              } break;
              case 2: {   // G0 = -64: offsets -64 - lane + jj: live from jj = lane on, all <= -D
                const float c = cN + nm;
                run([&](int) { return c; }, [&](int jj) { return jj >= lane; });
              } break;
              case 3: {   // G0 = +64: live up to jj = lane, all >= D
                const float c = cP + nm;
                run([&](int) { return c; }, [&](int jj) { return jj <= lane; });
              } break;
              case 4: {   // G0 = -32: offset o = jj - lane - 32 in [-63, -1]; the table matters where o > -12
                const float c = cN + nm;
                const float* tb = rel_s + (12 - 32 - lane) * TM + row;   // slot(o) = o + 12
                run([&](int jj) {
                  if (jj < 21) return c;
                  return (jj > lane + 20) ? tb[jj * TM] + nm : c;
                }, yes);
              } break;
              case 5: {   // G0 = 0: the diagonal group, o = jj - lane
                const float cp = cP + nm, cn = cN + nm;
                const float* tb = rel_s + (12 - lane) * TM + row;
                const int t0 = 11 - lane;
                run([&](int jj) {
                  const bool win = (unsigned)(t0 + jj) < 23u;      // |o| < 12
                  const float far = (jj > lane) ? cp : cn;
                  return win ? tb[jj * TM] + nm : far;
                }, yes);
              } break;
              case 6: {   // G0 = +32: o = jj - lane + 32 in [1, 63]; the table matters where o < 12
                const float c = cP + nm;
                const float* tb = rel_s + (12 + 32 - lane) * TM + row;
                run([&](int jj) {
                  if (jj > 10) return c;
                  return (jj < lane - 20) ? tb[jj * TM] + nm : c;
                }, yes);
              } break;
              case 7: case 7 | 16: {
                // general band group: per-element liveness (band edge, sequence ends), 1-D relative term
                // gathered from the slot-ordered row table, per-key mask when it is not uniform
                const int d0 = g0 - i;   // offset of element jj: d0 + jj
                int jlo = max(0, -rad - d0), jhi = min(32, rad - d0 + 1);
                jlo = max(jlo, -g0);
                jhi = min(jhi, p.L - g0);
                const unsigned span = (unsigned)max(jhi - jlo, 0);
                const float* base = rel_s + row;
                const int sD = d0 + D;
                if (fm == 7) {
                  run([&](int jj) { return base[min(max(sD + jj, 0), 2 * D) * TM] + nm; },
                      [&](int jj) { return (unsigned)(jj - jlo) < span; });
                } else {
                  run([&](int jj) { return base[min(max(sD + jj, 0), 2 * D) * TM] - m2; },
                      [&](int jj) {   // the shuffle is executed by every lane (no short-circuit in front of it)
                        const bool same = __shfl_sync(0xffffffffu, ceg, jj) == q_e;
                        return same && (unsigned)(jj - jlo) < span;
                      });
                }
              } break;
              default: {
                // general global-token group: own-sentence column, key tail, per-key mask
                const int sp = q_sent - g0;
                const int jhi = min(32, p.G - g0);
                if (fm == 8) {
                  const float k0 = cX + nm, k1 = cX1 + nm;
                  run([&](int jj) { return sp == jj ? k1 : k0; }, [&](int jj) { return jj < jhi; });
                } else {
                  const float k0 = cX - m2, k1 = cX1 - m2;
                  run([&](int jj) { return sp == jj ? k1 : k0; },
                      [&](int jj) {
                        const bool same = __shfl_sync(0xffffffffu, ceg, jj) == q_e;
                        return same && jj < jhi;
                      });
                }
              } break;
            }
          }
          // the per-lane table reads of the diagonal forms may leave the warp diverged: tcgen05.st is .aligned
          __syncwarp();
          // P of keys [32g, 32g+32) -> packed columns [32g, 32g+16) of the same group's score columns
          tmem_st16(t_s + 32 * g, pk);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[w]);
        ST(50 + pc);
        l += (ls0 + ls1) + (ls2 + ls3);
      }
      // ---- epilogue: O / l -> out, statistics.  The halves add up their row sums, each stores 32 columns ----
      {
        float* xb = xch + xpar * (2 * 2 * TM);
        xpar ^= 1;
        xb[hf * TM + row] = l;
        named_bar_sync(bar_pair, 64);
        l = hf ? xb[row] + l : l + xb[TM + row];   // same association in both threads
      }
      ST(60);
      mbar_wait_warp(&bars->o_full[w], tile_par);
      ST(61);
      tc_fence_after_sync();
      const float inv = 1.f / l;
      __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(p.out, b, i, h) + hf * 32;
      {
        uint32_t v[32];
        tmem_ld32(t_base + T_O + hf * 32, v);
        tmem_wait_ld();
        tc_fence_before_sync();
        mbar_arrive(&bars->o_empty[w]);   // O has been read: the next tile of this slot may overwrite it
        if (row_ok) {
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(v[8 * x + 0]) * inv, __uint_as_float(v[8 * x + 1]) * inv);
            o4.y = pack_bf16x2(__uint_as_float(v[8 * x + 2]) * inv, __uint_as_float(v[8 * x + 3]) * inv);
            o4.z = pack_bf16x2(__uint_as_float(v[8 * x + 4]) * inv, __uint_as_float(v[8 * x + 5]) * inv);
            o4.w = pack_bf16x2(__uint_as_float(v[8 * x + 6]) * inv, __uint_as_float(v[8 * x + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + 8 * x) = o4;
          }
        }
      }
      if (row_ok && hf == 0) {
        // (reference >= row maximum, sum of exp relative to it): a consistent pair is all the backward needs
        float2* st = reinterpret_cast<float2*>(p.stats) + ((int64_t)(b * p.H + h) * p.L + i);
        *st = make_float2(m2 * (1.f / LOG2E), l);
      }
      tile_par ^= 1;
      ST(62);
    }
#ifdef MLT_TC_TRACE
    if (tr && blockIdx.x == 3) { for (int k = 0; k < 2 * tn0; ++k) g_trace_f[0][k] = tb0[k]; g_trace_fn[0] = tn0; }
#endif
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == NSW + 1) tmem_dealloc<512>(tmem);
  if (tid == 0 && atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {   // every CTA has taken its last pair
    p.sched[0] = 0;
    p.sched[1] = 0;
  }
}

}  // namespace gl2

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace_gl2f(long long* out, int* n) {
  cudaMemcpyFromSymbol(n, gl2::g_trace_fn, sizeof(int) * 2);
  return (int)cudaMemcpyFromSymbol(out, gl2::g_trace_f, sizeof(long long) * 2 * 2048);
}
#endif

// The specialised kernel serves the long rows of a compact global-local problem.
bool gl2_fwd_long_supported(const FwdArgs& a, int dtype, int d) {
  if (!tc_fwd_args_supported(a, dtype, d)) return false;
  if (a.nseg != 2 || !a.seg[0].band || a.seg[1].band) return false;
  const Side& s0 = a.seg[0].side;
  const Side& s1 = a.seg[1].side;
  if (s0.mask_rule != MR_EXAMPLE_ID || s1.mask_rule != MR_EXAMPLE_ID) return false;
  if (s0.id_rule != IDR_1D || s1.id_rule != IDR_CROSS_QSENT) return false;
  const int D = s0.max_distance, R = a.rows.R;
  if (R < 2 * D + 3 || R > 32 || s1.max_distance != D) return false;
  if (a.drop.thr != 0) return false;
  if (!(a.neg < -1e5f)) return false;   // absorbed-mask form only (the reference's -1e9)
  if (a.seg[0].radius != gl2::RAD || a.seg[0].len != a.rows.len) return false;   // fixed-radius chunk schedule
  // the keys of the band are the query sequence; q / k example ids are the same array
  if (s0.q_eid != s0.k_eid || s1.q_eid != s0.q_eid) return false;
  return true;
}

int gl2_launch_fwd_long(const FwdArgs& a, cudaStream_t st) {
  static PerDeviceOnce once;
  static int sm_count[64];
  static unsigned* sched_base[64];
  static std::atomic<unsigned> sched_slot{0};
  const int ae = once.run([] {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e = cudaFuncSetAttribute(gl2::gl2_fwd_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         gl2::SM_ALLOC);
    if (e == cudaSuccess) e = cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base[dev]), gl2::g_sched_fwd);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    return (int)e;
  });
  if (ae) return ae;
  int dev = 0;
  cudaGetDevice(&dev);
  gl2::Params p{};
  p.B = a.B; p.H = a.H; p.L = a.rows.len; p.G = a.seg[1].len; p.R = a.rows.R;
  p.D = a.seg[0].side.max_distance;
  p.radius = a.seg[0].radius;
  p.scale = a.scale;
  p.long_eid = a.seg[0].side.q_eid;
  p.glob_eid = a.seg[1].side.k_eid;
  p.sent = a.seg[1].side.sent;
  p.bias = reinterpret_cast<const __nv_bfloat16*>(a.rows.bias);
  p.out = a.out;
  p.stats = a.stats;
  p.pairs_per_bh = (p.L + 2 * gl2::TM - 1) / (2 * gl2::TM);
  p.total_pairs = p.pairs_per_bh * a.B * a.H;
  // launches that may overlap in time use different counters (256 in rotation, each left at zero by its kernel)
  p.sched = sched_base[dev] + 2 * (sched_slot.fetch_add(1, std::memory_order_relaxed) % 256);
  CUtensorMap mq, mk, mv, mgk, mgv, me;
  int e = 0;
  e |= make_qkv_tensor_map(&mq, a.rows.q.ptr, a.rows.q.sb, a.rows.q.sl, a.rows.q.sh, a.B, p.L, a.H, gl2::TM);
  e |= make_qkv_tensor_map(&mk, a.seg[0].k.ptr, a.seg[0].k.sb, a.seg[0].k.sl, a.seg[0].k.sh, a.B, p.L, a.H, 64);
  e |= make_qkv_tensor_map(&mv, a.seg[0].v.ptr, a.seg[0].v.sb, a.seg[0].v.sl, a.seg[0].v.sh, a.B, p.L, a.H, 64);
  e |= make_qkv_tensor_map(&mgk, a.seg[1].k.ptr, a.seg[1].k.sb, a.seg[1].k.sl, a.seg[1].k.sh, a.B, p.G, a.H, 64);
  e |= make_qkv_tensor_map(&mgv, a.seg[1].v.ptr, a.seg[1].v.sb, a.seg[1].v.sl, a.seg[1].v.sh, a.B, p.G, a.H, 64);
  e |= make_qkv_tensor_map(&me, a.rows.emb, (int64_t)p.R * a.H * 64, (int64_t)a.H * 64, 64, 1, p.R, a.H, 32);
  if (e) return MLT_ERR_UNSUPPORTED;
  const int grid = p.total_pairs < sm_count[dev] ? p.total_pairs : sm_count[dev];
  gl2::gl2_fwd_long_kernel<<<grid, gl2::NTHREADS, gl2::SM_ALLOC, st>>>(mq, mk, mv, mgk, mgv, me, p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
