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
__device__ unsigned long long g_trace_b[6][256];
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
constexpr float LOG2E = 1.4426950408889634f;
// Row records (TcBwdQParams::rec_ws): field stride inside a 64-row block, in floats.  68, not 64: the
// key-centric pass reads field (4 + id) of ONE query from 32 lanes with different ids -- with a
// stride of 64 all of them hit the same shared-memory bank (up to 25-way conflict on every gather of
// the diagonal groups).  68 spreads eight consecutive fields over eight banks (at most 4-way) and
// keeps every field row 16-byte aligned, so the warp-uniform reads of the FAST form stay LDS.128
// (a stride of 65 is conflict-free but turns those into scalar loads: measured slower overall).
constexpr int RSF = 68;

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
  // 32-bit index arithmetic (the host guarantees B * len * H * 8 < 2^32): 64-bit divisions made this
  // kernel issue-bound at half the HBM rate
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t r = t >> 3;   // row index over (b, i, h): h fastest (matches the tensor layout)
  const int sub8 = (int)(t & 7);
  const uint32_t total = (uint32_t)B * (uint32_t)len * (uint32_t)H;
  float acc = 0.f;
  int b = 0, i = 0, h = 0;
  const bool ok = r < total;
  if (ok) {
    const uint32_t bi = r / (uint32_t)H;
    h = (int)(r - bi * (uint32_t)H);
    b = (int)(bi / (uint32_t)len);
    i = (int)(bi - (uint32_t)b * (uint32_t)len);
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
// SLIM: a 256-column / ~105 KB configuration that lets TWO CTAs share an SM (relative vocabulary
// <= 32, one elementwise thread per row, dP single-buffered, S double-buffered through the allrel
// columns that are idle during the chunk loop, two K/V stages, the dallrel^T tile reuses a drained
// K/V stage).  Used for the long-row tiles, which have few chunks: prologue
// (table build), epilogue (dallrel assembly) and the MMA round trips of one CTA are covered by
// the other.
template <bool SLIM>
struct Cfg {
  static constexpr int NST = SLIM ? 2 : 4;
  static constexpr int NSLOT = SLIM ? 32 : 64;            // relative-table / bin slots
  static constexpr int SM_Q = 0;                          // 16 KB
  static constexpr int SM_DO = SM_Q + TM * 128;           // 16 KB
  static constexpr int SM_E = SM_DO + TM * 128;           // NSLOT x 128 B
  static constexpr int SM_KV = SM_E + NSLOT * 128;        // NST x 16 KB
  static constexpr int SM_REL = SM_KV + NST * 2 * TN * 128;    // [NSLOT][128] f32
  static constexpr int SM_BIN = SM_REL + NSLOT * TM * 4;       // slim: [32][128] f32; else 2 x [64][128] f32
  static constexpr int SM_BIN_BYTES = SLIM ? 32 * TM * 4 : 2 * 64 * TM * 4;
  // dallrel^T tile, bf16, [128 rows][64 ids] SW128 (16 KB); slim: a drained K/V stage
  static constexpr int SM_A = SLIM ? SM_KV : SM_BIN + SM_BIN_BYTES;
  static constexpr int SM_BS = SLIM ? SM_BIN + SM_BIN_BYTES : SM_A + TM * 128;   // bias partial sums [4 quadrants][64]
  static constexpr int SM_PLAN = SM_BS + 4 * 64 * 4;         // 4 x ChunkPlan
  static constexpr int SM_META = SM_PLAN + 4 * (int)sizeof(plan::ChunkPlan);   // 64 x RelMeta
  static constexpr int SM_BAR = SM_META + 64 * (int)sizeof(plan::RelMeta);
  static constexpr int SM_ALLOC = SM_BAR + 256 + 1024;
  // TMEM columns
  static constexpr uint32_t TCOLS = SLIM ? 256 : 512;
  static constexpr uint32_t T_S = 0, T_DP = SLIM ? 64 : 128, T_DQ = SLIM ? 128 : 256, T_REL = SLIM ? 192 : 320,
                            T_DE = SLIM ? 0 : 384;   // slim: the table-gradient tile reuses the drained S columns
};

struct Bars {
  uint64_t q_full, rel_full;
  uint64_t kv_full[4], kv_empty[4];
  uint64_t sdp_full[2], ds_full[2], dq_full, dar_full;
  uint64_t pl_full[4], pl_empty[4];
  uint64_t dp_full, rel_done;   // slim: dP of a chunk ready; allrel extracted (its columns become S buffer 1)
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
  float4* rowstat;    // ws [B, H, lp] = (m * log2e, 1 / l, delta, 0)     (tc_bwd_prep_kernel)
  // ws [B, H, lp / 64, rw + 4, 64]: the per-row records the key-centric pass consumes, exponent-ready,
  // stored field-major per block of 64 rows (one contiguous bulk copy per query chunk; in shared
  // memory field f of query x sits at f * RSF + x: immediate offsets, warp-broadcast reads):
  //   field 0      -(m*log2e + log2 l)               exponent term of an element without relative score
  //   field 1      log2 of p for a MASKED element    log2(1/l) on fully-masked rows, -inf otherwise
  //   field 2      delta = sum_c dO*O      field 3: 0
  //   field 4 + id allrel[id]*scale*log2e - (m*log2e + log2 l)
  float* rec_ws;
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
  int col_base;        // dropout counter offset of key 0
  int n_img;           // see plan::PSeg::n_img
};
struct RowC {          // per thread and segment
  int q_e, q_sent;
  float relP, relN, relX, relX1;
  float relM;          // 2-D layout: the row's cross-modality constant (plan::C_MODAL)
};
__device__ __forceinline__ float rel_const(int ccls, const RowC& rc) {
  return ccls == plan::C_POS ? rc.relP
                             : (ccls == plan::C_NEG ? rc.relN
                                                    : (ccls == plan::C_CROSS ? rc.relX : (ccls == plan::C_MODAL ? rc.relM : 0.f)));
}

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
  sc.col_base = sg.col_base;
  sc.n_img = (sg.side.id_rule == IDR_EXPLICIT && sg.side.npr > 0) ? sg.side.npr * sg.side.npr : 0;
  return sc;
}

__device__ __forceinline__ plan::PSeg make_pseg(const KeySeg& sg, const SegRange& r, int R, int pd, bool perm, bool ex) {
  plan::PSeg s;
  // compact 2-D layout served through the library's id plane (abi.cu with_ids_plane keeps npr)
  s.n_img = (ex && sg.side.id_rule == IDR_EXPLICIT && sg.side.npr > 0) ? sg.side.npr * sg.side.npr : 0;
  s.expl_ok = ex;
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
// `meta` = the id -> slot table of the tile (plan::rel_meta_init): one warp-broadcast LDS instead of the
// branchy id -> slot rule in this once-per-tile (cold) code
__device__ __forceinline__ void row_consts(RowC& rc, const SegC& sc, const float* rel_s, const plan::RelMeta* meta,
                                           int row, int i) {
  auto rel_at = [&](int id) -> float {
    return (id >= 0 && id < sc.R) ? rel_s[meta[id].slot_off + row] : 0.f;
  };
  const bool on = sc.id_rule != IDR_NONE;
  rc.relP = on ? rel_at(sc.D) : 0.f;
  rc.relN = on ? rel_at(2 * sc.D) : 0.f;
  rc.relX = on ? rel_at(2 * sc.D + 1) : 0.f;
  rc.relX1 = on ? rel_at(2 * sc.D + 2) : 0.f;
  rc.relM = (on && sc.n_img > 0) ? rel_at(plan::modal_id(i, sc.n_img, sc.D)) : 0.f;
}

// Generic per-element evaluation (any rule): relative term (already * scale), mask and liveness.
__device__ __forceinline__ void eval_generic_q(const SegC& sc, const RowC& rc, int b, int i, int row, bool row_ok,
                                               int j, int ke_j, int ks_j, const float* rel_s, bool& live,
                                               bool& ok, float& rel, int& slot) {
  const Side& sd = *sc.sd;
  const int off = j - i;
  slot = -1;
  ok = true;
  rel = 0.f;
  live = j < sc.ke && (!sc.band || (off <= sc.radius && off >= -sc.radius));
  if (!live) return;
  const int col = sc.band ? off + sc.radius : j;
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
  if (id >= 0 && id < sc.R) {
    slot = plan::slot_of_id(id, sc.pd, sc.perm);
    rel = rel_s[slot * TM + row];
  }
}

// Probabilities in the backward passes are evaluated in exponent form,
//     p = exp2(x * scale*log2e + rel*log2e - (m*log2e + log2 l))          unmasked element
//     p = 1/l on a fully-masked row, 0 otherwise                          masked element
// (in fp32 a masked score is the additive constant itself -- |s| < 32 is absorbed by -1e9 -- so a
// masked element never needs the large numbers; this keeps the folded constants free of
// cancellation).
//
// Query-centric backward.  Each elementwise thread owns (row, one 32-key group) of the chunks its
// warp set serves: NP = 2 threads per row, SETS warp sets on alternate chunks (S / dP are
// double-buffered by chunk parity, so set s owns buffer s).  The evaluation form of every
// (quadrant, group) pair comes from the planner warp (tc_plan.cuh).
constexpr int NPL = 4;   // plan ring slots
template <int SETS, bool SLIM>
constexpr int bq_threads() { return (4 * (SLIM ? 1 : 2) * SETS + 3) * 32; }

// EX: the instantiation that carries the EXPL form (explicit int32 side inputs); the compact
// instantiations stay free of its code and register pressure.
// ABSORB: the literal-`neg` mode is compiled out (|neg| > 1e5 guaranteed by the launcher): the instantiation the
// reference's -1e9 runs on.  Carrying the mode as a run-time switch costs the hot kernels 3 - 7 %.
template <int SETS, bool SLIM, bool EX = false, bool DROP = false, bool ABSORB = false>
__global__ void __launch_bounds__(bq_threads<SETS, SLIM>(), SLIM ? 2 : 1)
tc_bwd_q_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_v0,
                const __grid_constant__ CUtensorMap map_k1, const __grid_constant__ CUtensorMap map_v1,
                const __grid_constant__ CUtensorMap map_e, const __grid_constant__ TcBwdQParams p) {
  using namespace bq;
  using C = Cfg<SLIM>;
  constexpr int NST = C::NST, SM_Q = C::SM_Q, SM_DO = C::SM_DO, SM_E = C::SM_E, SM_KV = C::SM_KV, SM_REL = C::SM_REL,
                SM_BIN = C::SM_BIN, SM_A = C::SM_A, SM_BS = C::SM_BS, SM_PLAN = C::SM_PLAN, SM_META = C::SM_META,
                SM_BAR = C::SM_BAR;
  constexpr uint32_t T_S = C::T_S, T_DP = C::T_DP, T_DQ = C::T_DQ, T_REL = C::T_REL, T_DE = C::T_DE;
  constexpr int NP = SLIM ? 1 : 2;    // elementwise threads per row
  constexpr int W = 32;               // columns per group
  constexpr int NG = 2 / NP;          // groups of a chunk a thread walks through (slim: both)
  constexpr int NEW = 128 * NP;       // elementwise threads per set
  constexpr int NALL = NEW * SETS;    // all elementwise threads
  constexpr int NB = NP * SETS;       // elementwise threads per row over all sets (output / id slices)
  constexpr int WP = 4 * NB, WM = 4 * NB + 1, WPL = 4 * NB + 2;   // producer / MMA / planner warp
  constexpr int RB = SLIM ? 32 : 128 / NB;   // bin slots per array (host guarantees R <= RB)
  static_assert(SETS == 1 || SETS == 2, "chunk buffers are double-buffered");
  static_assert(!SLIM || SETS == 1, "the slim configuration has a single dP buffer");
  // chunk c uses buffer BUF(c) in its PH(c)-th use
  auto BUF = [](int c) { return SLIM ? 0 : (c & 1); };
  auto PH = [](int c) { return SLIM ? (c & 1) : ((c >> 1) & 1); };
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
    mbar_init(&bars->dp_full, 1);
    mbar_init(&bars->rel_done, 128);
    for (int s = 0; s < NPL; ++s) {
      mbar_init(&bars->pl_full[s], 1);
      mbar_init(&bars->pl_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == WM) tmem_alloc<C::TCOLS>(&bars->tmem_base);
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

  auto load_tile = [&]() {
    mbar_arrive_expect_tx(&bars->q_full, 2 * TM * 128 + rpad * 128);
    tma_load_4d(smem + SM_Q, &map_q, &bars->q_full, 0, i0, h, b);
    tma_load_4d(smem + SM_DO, &map_do, &bars->q_full, 0, i0, h, b);
    if (rpad) tma_load_4d(smem + SM_E, &map_e, &bars->q_full, 0, 0, h, 0);
  };
  auto load_chunk = [&](int c) {
    const int st = c % NST;
    TRACE(4, 2 * c);
    mbar_wait(&bars->kv_empty[st], ((c / NST) & 1) ^ 1);
    TRACE(4, 2 * c + 1);
    const bool first = c < r0.n;
    const int key0 = first ? r0.kb + c * TN : r1.kb + (c - r0.n) * TN;
    uint8_t* ks = smem + SM_KV + st * (2 * TN * 128);
    mbar_arrive_expect_tx(&bars->kv_full[st], 2 * TN * 128);
    tma_load_4d(ks, first ? &map_k0 : &map_k1, &bars->kv_full[st], 0, key0, h, b);
    tma_load_4d(ks + TN * 128, first ? &map_v0 : &map_v1, &bars->kv_full[st], 0, key0, h, b);
  };
  auto run_planner = [&](auto pre) {
    const plan::PSeg ps0 = make_pseg(a.seg[0], r0, R, pd, perm, EX);
    const plan::PSeg ps1 = make_pseg(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm, EX);
    const Side* qs_side = nullptr;
    if (ps0.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[0].side;
    if (a.nseg > 1 && ps1.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[1].side;
    plan::planner_loop<NPL, TN>(ps0, ps1, r0.n, r1.n, r0.kb, r1.kb, b, i0, qs_side ? qs_side->sent : nullptr,
                                qs_side ? qs_side->sent_len : 0, a.rows.len, plans, bars->pl_full, bars->pl_empty, lane,
                                pre);
  };
  if (warp == WP) {
    if (elect_one()) {
      load_tile();
      for (int c = 0; c < nchunks; ++c) load_chunk(c);
    }
  } else if (warp == WM) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_dq = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t q_addr = smem_u32(smem + SM_Q), do_addr = smem_u32(smem + SM_DO);
      TRACE(3, 0);
      mbar_wait(&bars->q_full, 0);
      TRACE(3, 1);
      tc_fence_after_sync();
      if (rpad) {
        const uint32_t idesc_r = make_idesc_bf16(TM, rpad, 0, 0);
        const uint32_t e_addr = smem_u32(smem + SM_E);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_REL, sdesc(q_addr).at(kk * 32),
                  sdesc(e_addr).at(kk * 32), idesc_r, kk > 0);
        umma_commit(&bars->rel_full);
      }
      // ---- slim: S runs one chunk ahead ------------------------------------------------------
      // S is double-buffered (buffer 1 = the allrel columns, idle between the table build and the
      // epilogue), dP is not: S_{c+1} is issued as soon as its K tile is there, so that when the
      // elementwise warps hand over dS_c they find S_{c+1} waiting and evaluate its probabilities
      // while dQ_c and dP_{c+1} execute.
      auto issue_s = [&](int c) {
        const int st = c % NST;
        mbar_wait(&bars->pl_full[c % NPL], (c / NPL) & 1);   // relayed to the elementwise warps by sdp_full
        TRACE(3, 4 + 4 * c);
        mbar_wait(&bars->kv_full[st], (c / NST) & 1);
        TRACE(3, 5 + 4 * c);
        if (c == 1 && rpad) mbar_wait(&bars->rel_done, 0);   // allrel has been read out of buffer 1
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + ((c & 1) ? T_REL : T_S), sdesc(q_addr).at(kk * 32),
                  sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->sdp_full[c & 1]);
      };
      auto issue_dp = [&](int c) {   // kv_full(c) has been waited for by issue_s(c)
        const uint32_t v_addr = smem_u32(smem + SM_KV + (c % NST) * (2 * TN * 128)) + TN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_DP, sdesc(do_addr).at(kk * 32),
                  sdesc(v_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->dp_full);
        TRACE(3, 6 + 4 * c);
      };
      auto issue_sdp = [&](int c) {
        const int st = c % NST;
        mbar_wait(&bars->pl_full[c % NPL], (c / NPL) & 1);   // relayed to the elementwise warps by sdp_full
        TRACE(3, 4 + 4 * c);
        mbar_wait(&bars->kv_full[st], (c / NST) & 1);
        TRACE(3, 5 + 4 * c);
        tc_fence_after_sync();
        const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
        const uint32_t v_addr = k_addr + TN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_S + BUF(c) * 64, sdesc(q_addr).at(kk * 32),
                  sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + T_DP + BUF(c) * 64, sdesc(do_addr).at(kk * 32),
                  sdesc(v_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->sdp_full[BUF(c)]);
        TRACE(3, 6 + 4 * c);
      };
      auto issue_dq = [&](int pc) {
        const int st = pc % NST;
        mbar_wait(&bars->ds_full[BUF(pc)], PH(pc));
        TRACE(3, 7 + 4 * pc);
        tc_fence_after_sync();
        mbar_arrive(&bars->pl_empty[pc % NPL]);   // every elementwise thread is done with plan pc
        const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ts(tmem + T_DQ, tmem + T_DP + BUF(pc) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                  sdesc(k_addr).at(kk * 2048), idesc_dq, (pc > 0 || kk > 0));
        umma_commit(&bars->kv_empty[st]);
        if (pc == nchunks - 1) {
          if (rpad) {
            // dQ += dallrel . E : A = dallrel (bf16, TMEM columns of the dead allrel region),
            // B = E tile taken MN-major (K = relative ids)
            mbar_wait(&bars->dar_full, 0);
            tc_fence_after_sync();
            const uint32_t e_addr = smem_u32(smem + SM_E);
            for (int kk = 0; kk < rpad / 16; ++kk)
              umma_ts(tmem + T_DQ, tmem + T_REL + kk * 8, sdesc(e_addr).at(kk * 2048),
                      idesc_dq, 1u);
            if (a.tg_partial) {
              // table-gradient partial of this tile: dE[64 ids x 64] = dallrel^T . Q  (M = 64, both
              // operands MN-major: K = the tile's 128 rows)
              const uint32_t idesc_de = make_idesc_bf16(64, 64, 1, 1);
              const uint32_t a_addr = smem_u32(smem + SM_A + (SLIM ? (nchunks % NST) * (2 * TN * 128) : 0));
#pragma unroll
              for (int kk = 0; kk < 8; ++kk)
                umma_ss(tmem + T_DE, sdesc(a_addr).at(kk * 2048),
                        sdesc(q_addr).at(kk * 2048), idesc_de, kk > 0);
            }
          }
          umma_commit(&bars->dq_full);
        }
      };
      // double-buffered: S/dP of chunk c run ahead of dQ of chunk c-1.  SLIM (single dP buffer): S runs
      // one chunk ahead in its own buffer; the dQ MMAs that read dS are issued before the dP MMAs that
      // overwrite it (MMAs of one thread execute in issue order)
      if (SLIM) {
        if (nchunks > 0) {
          issue_s(0);
          issue_dp(0);
        }
        for (int c = 0; c < nchunks; ++c) {
          if (c + 1 < nchunks) issue_s(c + 1);
          issue_dq(c);                       // waits for dS_c; MMAs execute in issue order
          if (c + 1 < nchunks) issue_dp(c + 1);
        }
      } else {
        for (int c = 0; c <= nchunks; ++c) {
          if (c < nchunks) issue_sdp(c);
          if (c >= 1) issue_dq(c - 1);
        }
      }
    }
  } else if (warp == WPL) {
    // ===================== planner =====================
#ifdef MLT_TC_TRACE
    if (lane == 0) TRACE(4, 30);
    run_planner([&](int c) { if (lane == 0) TRACE(4, 32 + c); });
#else
    run_planner(plan::NoPre());
#endif
  } else {
    // ===================== elementwise warps (NP threads per row) =====================
    if (tid == 0) TRACE(1, 0);
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int set = warp / (4 * NP);     // which chunk parity this warp serves
    const int part0 = (warp >> 2) % NP;  // which 32-key group of the chunk (slim: both, in turn)
    const int bidx = set * NP + part0;   // private bin array / output column slice
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    float* bin = bins + bidx * RB * TM;   // slot-ordered, private to (set, part, row)
    for (int x = 4 * (lane + 32 * quad); x < RB * TM; x += 512)   // 128-bit stores: this is once-per-tile code
      *reinterpret_cast<float4*>(bin + x) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!SLIM)   // slim: the tile lives in a K/V stage and is written in full by the epilogue
      for (int x = tid; x < TM * 128 / 16; x += NALL) reinterpret_cast<uint4*>(smem + SM_A)[x] = make_uint4(0u, 0u, 0u, 0u);
    const SegC sc0 = make_segc(a.seg[0], r0, R, pd, perm);
    const SegC sc1 = make_segc(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm);
    RowC rc0, rc1;
    row_loads(rc0, sc0, b, i, row_ok);
    row_loads(rc1, sc1, b, i, row_ok);
    // row constants (written by tc_bwd_prep_kernel)
    // The ABI's `neg` is honoured literally (see tc_fwd.cu): |neg| > 1e5 -> a masked score is the
    // additive constant itself (per-row constant lpm below); |neg| <= 1e5 ("literal" mode) -> a masked
    // element evaluates like an unmasked one with neg * log2e added to its exponent, and FAST groups
    // with masked rows take the per-element EDGE form.
    const bool lit = ABSORB ? false : fabsf(a.neg) <= 1e5f;
    const bool skip_ok = a.neg < -200.f;
    const float negl2 = a.neg * LOG2E;
    const float real_thr2 = 0.5f * a.neg * LOG2E;
    const uint32_t drow = DROP ? dropout_row_base(dropout_salt(a.drop, (uint32_t)(b * a.H + h)), i0 + (warp & 3) * 32 + lane) : 0u;
    const uint32_t dthr = a.drop.thr;
    const float dinv = a.drop.inv_keep;
    // dP of the element at dropout column `col`: the gradient reaches P only where it was kept
    auto dpv = [&](uint32_t wraw, int col) -> float {
      const float w = __uint_as_float(wraw);
      if constexpr (DROP) return dropout_keep(drow, col, dthr) ? w * dinv : 0.f;
      else return w;
    };
    float nm2l = -INFINITY;   // -(m*log2e + log2 l): rows beyond the end evaluate to p = 0
    float lpm = -INFINITY;    // log2 of p of a masked element
    float delta = 0.f;
    bool real_max = true;     // the row maximum is an unmasked score
    const int64_t srow = (int64_t)(b * a.H + h) * a.rows.len + i;
    const int64_t prow = (int64_t)(b * a.H + h) * p.lp + i;
    if (row_ok) {
      const float4 rs4 = __ldg(p.rowstat + prow);
      const float ll = __log2f(rs4.y);   // log2(1/l)
      nm2l = ll - rs4.x;
      real_max = rs4.x > real_thr2;
      lpm = real_max ? -INFINITY : ll;
      delta = rs4.z;
    }
    const int rs_stride = p.rw + 4;
    // field 0 of this row: block (i / 64), lane (i % 64); consecutive fields are 64 floats apart
    float* rec_row = p.rec_ws + (((int64_t)(b * a.H + h) * (p.lp >> 6) + (i >> 6)) * rs_stride) * RSF + (i & 63);
    if (bidx == 0 && row_ok) {
      rec_row[0] = nm2l;
      rec_row[RSF] = lpm;
      rec_row[2 * RSF] = delta;
      rec_row[3 * RSF] = 0.f;
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
      float* ws_row = rec_row + 4 * RSF;
      const int rw = p.rw;
      plan::rel_table_build(tmem + T_REL + lane_sel, relmeta, rel_s, row, rpad, a.scale,
                            [&](int c0, const float (&val)[16]) {
                              if (row_ok) {   // one coalesced 128-byte store per id and warp
#pragma unroll
                                for (int x = 0; x < 16; ++x)
                                  if (c0 + x < rw) ws_row[(c0 + x) * RSF] = fmaf(val[x], LOG2E, nm2l);
                              }
                            });
      if (SLIM) {   // the allrel columns may now be overwritten (S buffer 1)
        tc_fence_before_sync();
        mbar_arrive(&bars->rel_done);
      }
    }
    if (tid == 0) TRACE(1, 3);
    named_bar_sync(1, NALL);  // rel_s (written by set 0 / part 0) visible to all; bins zeroed
    if (tid == 0) TRACE(1, 4);
    row_consts(rc0, sc0, rel_s, relmeta, row, i);
    row_consts(rc1, sc1, rel_s, relmeta, row, i);
    // per-row accumulators of the constant relative classes (flushed into the bins at the end)
    float accP = 0.f, accN = 0.f, accX = 0.f, accX1 = 0.f, accM = 0.f;
    const float scale2 = a.scale * LOG2E;

    // One call per key segment (inlined twice: no per-field selects inside the chunk loop).
    // Chunks [c_begin, c_end) belong to this segment; this warp set handles c % SETS == set.
    auto run_chunks = [&](const SegC sc, const RowC rc, int c_begin, int c_end, int kb) {
      int c = c_begin + ((set - c_begin) % SETS + SETS) % SETS;
      const bool mre = sc.mask_rule == MR_EXAMPLE_ID;
#pragma unroll 1
      for (; c < c_end; c += SETS) {
        const plan::ChunkPlan* cp = plans + (c % NPL);
        if (tid == 0) TRACE(1, 8 + 3 * c);
        // the MMA warp issued S_c / dP_c only after plan c had been published.  Slim: this barrier
        // covers S_c alone (buffer c & 1); dP_c follows behind dp_full (wait_dp below).
        if (SLIM) mbar_wait_warp(&bars->sdp_full[c & 1], (c >> 1) & 1);
        else mbar_wait_warp(&bars->sdp_full[BUF(c)], PH(c));
        if (tid == 0) TRACE(1, 9 + 3 * c);
        tc_fence_after_sync();
        bool dp_ready = !SLIM;
        auto wait_dp = [&]() {   // before the first access (read or write) to the chunk's dP columns
          if (!dp_ready) {
            mbar_wait_warp(&bars->dp_full, c & 1);
            tc_fence_after_sync();
            dp_ready = true;
          }
        };
#pragma unroll 1
        for (int pg = 0; pg < NG; ++pg) {
        const int part = NG > 1 ? pg : part0;
        const int g0 = kb + (c - c_begin) * TN + part * W;
        const uint32_t t_s = tmem + (SLIM ? ((c & 1) ? T_REL : T_S) : T_S + BUF(c) * 64) + lane_sel + part * W;
        const uint32_t t_dp = tmem + T_DP + BUF(c) * 64 + lane_sel + part * W;
        const uint32_t w0 = cp->q[quad][part];
        const int ce0 = (int)cp->q[quad][2 + part];
        int mode = (int)(w0 & 0xffu);
        const bool mask_pe = (w0 & plan::F_MASK_PE) != 0;
        const bool masked = mre && !mask_pe && (rc.q_e != ce0);
        const int ccls = (int)((w0 >> 8) & 0xffu);
        const float relc = rel_const(ccls, rc);
        const int dcol = sc.col_base + g0;   // dropout counter of the group's first key
        uint32_t ds_pk[W / 2];
        bool zero = (mode == plan::DEAD);
        // every row of the warp masked for the whole group while holding a real maximum: p == 0 exactly
        if (!zero && mode == plan::FAST && mre && skip_ok && __all_sync(0xffffffffu, masked && real_max)) zero = true;
        if (lit && !zero && mode == plan::FAST && __any_sync(0xffffffffu, masked)) mode = plan::EDGE;
        if (zero) {
#pragma unroll
          for (int x = 0; x < W / 2; ++x) ds_pk[x] = 0u;
        } else if (mode == plan::FAST) {
          uint32_t v[W], w[W];
          tmem_ld32(t_s, v);
          if (!SLIM) tmem_ld32(t_dp, w);
          tmem_wait_ld();
          const float gmul = masked ? 0.f : scale2;
          const float gsub = masked ? lpm : fmaf(relc, LOG2E, nm2l);
          float t0 = 0.f, t1 = 0.f;
          if (SLIM) {   // probabilities first: dP of the chunk may still be executing
#pragma unroll
            for (int x = 0; x < W; ++x) v[x] = __float_as_uint(ex2(fmaf(__uint_as_float(v[x]), gmul, gsub)));
            wait_dp();
            tmem_ld32(t_dp, w);
            tmem_wait_ld();
          }
#pragma unroll
          for (int x = 0; x < W / 2; ++x) {
            const float p0 = SLIM ? __uint_as_float(v[2 * x]) : ex2(fmaf(__uint_as_float(v[2 * x]), gmul, gsub));
            const float p1 = SLIM ? __uint_as_float(v[2 * x + 1]) : ex2(fmaf(__uint_as_float(v[2 * x + 1]), gmul, gsub));
            const float d0 = p0 * (dpv(w[2 * x], dcol + 2 * x) - delta);
            const float d1 = p1 * (dpv(w[2 * x + 1], dcol + 2 * x + 1) - delta);
            t0 += d0;
            t1 += d1;
            ds_pk[x] = pack_bf16x2(d0, d1);
          }
          const float tot = t0 + t1;
          if (ccls == plan::C_POS) accP += tot;
          else if (ccls == plan::C_NEG) accN += tot;
          else if (ccls == plan::C_CROSS) accX += tot;
          else if (EX && ccls == plan::C_MODAL) accM += tot;
        } else {
          float ds[W];
          if (mode == plan::GEN) {
            // real loop, TMEM as dynamically indexed scratch: one copy of the generic code
            wait_dp();
#pragma unroll 1
            for (int jj = 0; jj < W; ++jj) {
              const uint32_t raw = tmem_ld1(t_s + jj);
              const uint32_t dpr = tmem_ld1(t_dp + jj);
              tmem_wait_ld();
              int slot;
              bool live, ok;
              float rel;
              eval_generic_q(sc, rc, b, i, row, row_ok, g0 + jj, cp->ce[part * W + jj], cp->cs[part * W + jj], rel_s,
                             live, ok, rel, slot);
              const float eu = fmaf(__uint_as_float(raw), scale2, fmaf(rel, LOG2E, nm2l));
              const float pv = ex2(ok ? eu : (lit ? eu + negl2 : lpm));
              const float dsv = live ? pv * (dpv(dpr, dcol + jj) - delta) : 0.f;
              if (live && slot >= 0) bin[slot * TM + row] += dsv;
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
            // e[jj] = exponent of element jj (log2 units); dead elements get -inf, masked ones lpm
            float e[W];
            uint32_t w[W];
            {
              uint32_t v[W];
              tmem_ld32(t_s, v);
              if (!SLIM) tmem_ld32(t_dp, w);
              tmem_wait_ld();
#pragma unroll
              for (int x = 0; x < W; ++x) e[x] = __uint_as_float(v[x]);
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
                const float c2 = fmaf(relc, LOG2E, nm2l);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) {
                  const float v = fmaf(e[jj], scale2, c2);
                  e[jj] = ((unsigned)(jj - jlo) < span) ? v : -INFINITY;
                }
                break;
              }
              case plan::DIAG: {
                const int d0 = g0 - i + sc.D;
                const float* base = rel_s + row;
#pragma unroll
                for (int jj = 0; jj < W; ++jj) {
                  const int sl = min(max(d0 + jj, 0), 2 * sc.D);
                  e[jj] = fmaf(e[jj], scale2, fmaf(base[sl * TM], LOG2E, nm2l));
                }
                break;
              }
              case plan::QS: {
                const int d0 = rc.q_sent - g0;
                const float c0 = fmaf(rc.relX, LOG2E, nm2l), c1 = fmaf(rc.relX1, LOG2E, nm2l);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) e[jj] = fmaf(e[jj], scale2, d0 == jj ? c1 : c0);
                break;
              }
              case plan::EXPL: if constexpr (EX) {
                // explicit int32 tensors: this row's 32 consecutive entries (band: column k = j - i + r)
                const Side& sd = *sc.sd;
                const int d0 = g0 - i;
                int jlo = 0, jhi = min(W, sc.ke - g0);
                if (sc.band) {
                  jlo = max(jlo, -sc.radius - d0);
                  jhi = min(jhi, sc.radius - d0 + 1);
                }
                const uint32_t live = plan::span_bits(jlo, jhi);
                const uint32_t take = row_ok ? live : 0u;
                const int64_t eoff = (int64_t)b * sd.sb + (int64_t)i * sd.sq + (sc.band ? d0 + sc.radius : g0);
                const bool vec = !sc.band && ((sd.sb | sd.sq) & 3) == 0 &&
                                 __all_sync(0xffffffffu, take == 0xffffffffu);
                if (sc.id_rule == IDR_EXPLICIT) {
                  int id[W];
                  plan::load_row32(sd.ids + eoff, vec && (reinterpret_cast<uintptr_t>(sd.ids) & 15) == 0, take, -1, id);
#pragma unroll
                  for (int jj = 0; jj < W; ++jj) {
                    float rel = 0.f;
                    if ((unsigned)id[jj] < (unsigned)sc.R) rel = rel_s[relmeta[id[jj]].slot_off + row];
                    e[jj] = fmaf(e[jj], scale2, fmaf(rel, LOG2E, nm2l));
                  }
                } else {
#pragma unroll
                  for (int jj = 0; jj < W; ++jj) e[jj] = fmaf(e[jj], scale2, nm2l);
                }
                if (sc.mask_rule == MR_EXPLICIT) {
                  int ok[W];
                  plan::load_row32(sd.mask + eoff, vec && (reinterpret_cast<uintptr_t>(sd.mask) & 15) == 0, take, 1, ok);
#pragma unroll
                  for (int jj = 0; jj < W; ++jj) e[jj] = ok[jj] != 0 ? e[jj] : (lit ? e[jj] + negl2 : lpm);
                }
                if (live != 0xffffffffu) {
#pragma unroll
                  for (int jj = 0; jj < W; ++jj) e[jj] = (live >> jj) & 1u ? e[jj] : -INFINITY;
                }
              } break;
              default: {   // KS
                const float c0 = fmaf(rc.relX, LOG2E, nm2l), c1 = fmaf(rc.relX1, LOG2E, nm2l);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) e[jj] = fmaf(e[jj], scale2, cp->cs[part * W + jj] == i ? c1 : c0);
                break;
              }
            }
            if (mask_pe) {
#pragma unroll
              for (int jj = 0; jj < W; ++jj)
                e[jj] = (cp->ce[part * W + jj] == rc.q_e || e[jj] == -INFINITY) ? e[jj] : (lit ? e[jj] + negl2 : lpm);
            } else if (mre && __any_sync(0xffffffffu, masked)) {
#pragma unroll
              for (int jj = 0; jj < W; ++jj) e[jj] = (masked && e[jj] != -INFINITY) ? (lit ? e[jj] + negl2 : lpm) : e[jj];
            }
            float t0 = 0.f, t1 = 0.f;
            if (SLIM) {   // probabilities first: dP of the chunk may still be executing
#pragma unroll
              for (int x = 0; x < W; ++x) e[x] = ex2(e[x]);
              wait_dp();
              tmem_ld32(t_dp, w);
              tmem_wait_ld();
            }
#pragma unroll
            for (int x = 0; x < W; x += 2) {
              const float p0 = SLIM ? e[x] : ex2(e[x]);
              const float p1 = SLIM ? e[x + 1] : ex2(e[x + 1]);
              ds[x] = p0 * (dpv(w[x], dcol + x) - delta);
              ds[x + 1] = p1 * (dpv(w[x + 1], dcol + x + 1) - delta);
              t0 += ds[x];
              t1 += ds[x + 1];
            }
            const float tot = t0 + t1;
            // ---- relative-id bins ----
            if (mode == plan::EDGE) {
              if (ccls == plan::C_POS) accP += tot;
              else if (ccls == plan::C_NEG) accN += tot;
              else if (ccls == plan::C_CROSS) accX += tot;
              else if (EX && ccls == plan::C_MODAL) accM += tot;
            } else if (mode == plan::DIAG) {
              // slot = clamp(d0 + x, 0, 2D).  The clamped ends are the constant classes (offset <= -D,
              // offset >= D): register accumulators.  Every interior slot belongs to exactly one key
              // of the row (offset = slot - D), so it is written once: a plain store, no
              // read-modify-write chain through shared memory.
              const int d0 = g0 - i + sc.D;
              float* base = bin + row;
              float eN0 = 0.f, eN1 = 0.f, eP0 = 0.f, eP1 = 0.f;
#pragma unroll
              for (int x = 0; x < W; x += 2) {
                const int s0 = d0 + x, s1 = d0 + x + 1;
                eN0 += (s0 <= 0) ? ds[x] : 0.f;
                eN1 += (s1 <= 0) ? ds[x + 1] : 0.f;
                eP0 += (s0 >= 2 * sc.D) ? ds[x] : 0.f;
                eP1 += (s1 >= 2 * sc.D) ? ds[x + 1] : 0.f;
                if (s0 > 0 && s0 < 2 * sc.D) base[s0 * TM] = ds[x];
                if (s1 > 0 && s1 < 2 * sc.D) base[s1 * TM] = ds[x + 1];
              }
              accN += eN0 + eN1;
              accP += eP0 + eP1;
            } else if (mode == plan::QS) {
              const int d0 = rc.q_sent - g0;
              float sp = 0.f;
#pragma unroll
              for (int x = 0; x < W; ++x) sp += (d0 == x) ? ds[x] : 0.f;
              accX1 += sp;
              accX += tot - sp;
            } else if (EX && mode == plan::EXPL) {
              if (sc.id_rule == IDR_EXPLICIT) {
                // ids re-read (L1 hits) instead of kept live across the exponentials; entries not taken
                // read back as -1.  Read-modify-write in column order: same association as the generic form.
                const Side& sd = *sc.sd;
                const int d0 = g0 - i;
                int jlo = 0, jhi = min(W, sc.ke - g0);
                if (sc.band) {
                  jlo = max(jlo, -sc.radius - d0);
                  jhi = min(jhi, sc.radius - d0 + 1);
                }
                const uint32_t take = row_ok ? plan::span_bits(jlo, jhi) : 0u;
                const int64_t eoff = (int64_t)b * sd.sb + (int64_t)i * sd.sq + (sc.band ? d0 + sc.radius : g0);
                const bool vec = !sc.band && ((sd.sb | sd.sq) & 3) == 0 && (reinterpret_cast<uintptr_t>(sd.ids) & 15) == 0 &&
                                 __all_sync(0xffffffffu, take == 0xffffffffu);
                int id[W];
                plan::load_row32(sd.ids + eoff, vec, take, -1, id);
#pragma unroll
                for (int x = 0; x < W; ++x)
                  if ((unsigned)id[x] < (unsigned)sc.R) bin[relmeta[id[x]].slot_off + row] += ds[x];
              }
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
        // each group packs into its OWN column range (the other group's inputs are still unread)
        wait_dp();   // (dead groups reach this point without having touched dP)
        tmem_st16(t_dp, ds_pk);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->ds_full[BUF(c)]);
        if (tid == 0) TRACE(1, 10 + 3 * c);
      }
    };
    // (one inlined copy per segment measured faster than a single copy behind a segment loop here:
    // 0.227 vs 0.240 ms on the global rows, no difference on the long rows)
    run_chunks(sc0, rc0, 0, r0.n, r0.kb);
    if (a.nseg > 1) run_chunks(sc1, rc1, r0.n, nchunks, r1.kb);
    // flush the constant-class accumulators into this part's bins
    if (R > 0) {
      auto flush = [&](int id, float v) {
        if (id >= 0 && id < R) bin[relmeta[id].slot_off + row] += v;
      };
      const int dd = sc0.D;   // both segments of a row set share max_distance
      flush(dd, accP);
      flush(2 * dd, accN);
      flush(2 * dd + 1, accX);
      flush(2 * dd + 2, accX1);
      if (EX && sc0.n_img > 0) flush(plan::modal_id(i, sc0.n_img, dd), accM);
    }
    // ---- epilogue: dallrel (summed over parts) -> global + bf16 A-operand for dQ += dallrel.E ----
    if (tid == 0) TRACE(1, 5);
    named_bar_sync(1, NALL);  // all bin arrays complete
    // slim: the dallrel^T tile reuses the K/V stage the last chunk did NOT use (its reader, the dQ
    // MMA of chunk n-2, completed before S/dP of the last chunk were seen)
    uint8_t* a_tile = smem + SM_A + (SLIM ? (nchunks % NST) * (2 * TN * 128) : 0);
    if (rpad) {
      if (SLIM && a.tg_partial) {   // id columns beyond rpad: zeros (non-slim: the tile was cleared up front)
        for (int c0 = rpad; c0 < 64; c0 += 8)
          *reinterpret_cast<uint4*>(a_tile + row * 128 + ((((c0 >> 3)) ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      }
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
            const int so = relmeta[pid].slot_off;   // slot * TM (warp-broadcast LDS instead of the branchy id -> slot rule)
#pragma unroll
            for (int pp = 0; pp < NB; ++pp) w += bins[pp * RB * TM + so + row];
          }
          w8[x] = w;
        }
        if (a.tg_partial) {
          // rows beyond the sequence end contribute nothing (their bins may hold garbage-free zeros,
          // but Q rows are zero-filled by TMA anyway)
          const uint4 pk4 = make_uint4(pack_bf16x2(w8[0], w8[1]), pack_bf16x2(w8[2], w8[3]),
                                       pack_bf16x2(w8[4], w8[5]), pack_bf16x2(w8[6], w8[7]));
          *reinterpret_cast<uint4*>(a_tile + row * 128 + ((((c0 >> 3)) ^ (row & 7)) << 4)) = pk4;
          // bias partial: sum over the 32 rows of this warp, one value per id
          float* bs = reinterpret_cast<float*>(smem + SM_BS) + quad * 64 + c0;
          {
            // 8 values per lane -> 8 sums over the 32 lanes with 9 shuffles: every exchange step
            // halves the number of values a lane carries (lane bits 4, 3, 2 select the id)
            const bool h4 = (lane & 16) != 0, h3 = (lane & 8) != 0, h2 = (lane & 4) != 0;
            float b4[4], b2[2];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float lo = row_ok ? w8[k] : 0.f, hi = row_ok ? w8[k + 4] : 0.f;
              b4[k] = (h4 ? hi : lo) + __shfl_xor_sync(0xffffffffu, h4 ? lo : hi, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k)
              b2[k] = (h3 ? b4[k + 2] : b4[k]) + __shfl_xor_sync(0xffffffffu, h3 ? b4[k] : b4[k + 2], 8);
            float r = (h2 ? b2[1] : b2[0]) + __shfl_xor_sync(0xffffffffu, h2 ? b2[0] : b2[1], 4);
            r += __shfl_xor_sync(0xffffffffu, r, 2);
            r += __shfl_xor_sync(0xffffffffu, r, 1);
            if ((lane & 3) == 0) bs[(h4 ? 4 : 0) + (h3 ? 2 : 0) + (h2 ? 1 : 0)] = r;
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
    constexpr int WL = WO > 32 ? 32 : WO;   // ... read in pieces of at most 32 columns
#pragma unroll 1
    for (int hh = 0; hh < WO / WL; ++hh) {
      uint32_t dq_raw[WL];
      tmem_ldN(tmem + T_DQ + lane_sel + bidx * WO + hh * WL, dq_raw);
      tmem_wait_ld();
      if (row_ok) {
        __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(a.d_q, b, i, h) + bidx * WO + hh * WL;
#pragma unroll
        for (int x = 0; x < WL / 8; ++x) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(dq_raw[8 * x + 0]) * a.scale, __uint_as_float(dq_raw[8 * x + 1]) * a.scale);
          w.y = pack_bf16x2(__uint_as_float(dq_raw[8 * x + 2]) * a.scale, __uint_as_float(dq_raw[8 * x + 3]) * a.scale);
          w.z = pack_bf16x2(__uint_as_float(dq_raw[8 * x + 4]) * a.scale, __uint_as_float(dq_raw[8 * x + 5]) * a.scale);
          w.w = pack_bf16x2(__uint_as_float(dq_raw[8 * x + 6]) * a.scale, __uint_as_float(dq_raw[8 * x + 7]) * a.scale);
          *reinterpret_cast<uint4*>(dst + 8 * x) = w;
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WM) tmem_dealloc<C::TCOLS>(tmem);
}

// ============================================================================================
// Key-centric pass
// ============================================================================================
namespace bk {
constexpr int NST = 3;
// SLIM: a 256-column / ~111 KB configuration that lets TWO CTAs share an SM (relative vocabulary
// <= 32, S^T / dP^T single-buffered, the producer warp doubles as the planner).  The long-key tiles
// have few chunks, so start-up and drain of one CTA are hidden by the other instead of by
// double buffering.
template <bool SLIM>
struct Cfg {
  static constexpr int RSMAX = (SLIM ? 32 : 64) + 4;     // floats per row record (header + ids)
  static constexpr int SM_K = 0;                         // 16 KB
  static constexpr int SM_V = SM_K + TM * 128;           // 16 KB
  static constexpr int SM_QD = SM_V + TM * 128;          // NST x (Q 8 KB + dO 8 KB)
  static constexpr int SM_REC = SM_QD + NST * 2 * TN * 128;      // NST x [64] row records (see TcBwdQParams::rec_ws)
  static constexpr int SM_PLAN = SM_REC + NST * RSF * RSMAX * 4;  // NPL x ChunkPlan
  static constexpr int SM_BAR = SM_PLAN + NPL * (int)sizeof(plan::ChunkPlan);
  static constexpr int SM_ALLOC = SM_BAR + 256 + 1024;
  static constexpr uint32_t TCOLS = SLIM ? 256 : 512;
  static constexpr uint32_t T_S = 0, T_DP = SLIM ? 64 : 128, T_DV = SLIM ? 128 : 256, T_DK = SLIM ? 192 : 320;
};

struct Bars {
  uint64_t kv_full;
  uint64_t qd_full[NST], qd_empty[NST];
  uint64_t sdp_full[2], pds_full[2], acc_full;
  uint64_t pl_full[NPL], pl_empty[NPL];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");
}  // namespace bk

struct TcQuerySource {
  QuerySource q;
  const float* rec_ws;    // [B, H, lp, rw + 4] row records written by the query-centric pass
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
    r.ib = max(0, j0 - s.q.radius) & ~63;  // query chunks coincide with the 64-row record blocks
    r.ie = min(lq, j0 + TM + s.q.radius);
  } else {
    r.ib = 0;
    r.ie = lq;
  }
  r.n = (r.ie - r.ib + TN - 1) / TN;
  return r;
}

// ---- key-centric contexts (rows = keys, columns = queries) -------------------------------------
struct SrcC {            // warp-uniform, one per query source
  const Side* sd;
  int ie;                // end of the live query range
  int R, D, rw;
  bool band;
  int radius;
  int mask_rule, id_rule;
  int col_base;          // dropout: position of this key set on the source rows' key axis
  Dropout drop;          // dropout descriptor of the source's row set
  int n_img;             // see plan::PSeg::n_img
};
struct KeyC {            // per thread and source
  int k_e, k_sent;
};

__device__ __forceinline__ SrcC make_srcc(const TcQuerySource& src, const SrcRange& r) {
  SrcC sc;
  sc.sd = &src.q.side;
  sc.ie = r.ie;
  sc.R = src.q.rows.R;
  sc.D = src.q.side.max_distance;
  sc.rw = src.rw + 4;   // record stride (floats)
  sc.n_img = (src.q.side.id_rule == IDR_EXPLICIT && src.q.side.npr > 0) ? src.q.side.npr * src.q.side.npr : 0;
  sc.band = src.q.band != 0;
  sc.radius = src.q.radius;
  sc.mask_rule = src.q.side.mask_rule;
  sc.id_rule = sc.R > 0 ? src.q.side.id_rule : IDR_NONE;
  sc.col_base = src.q.col_base;
  sc.drop = src.q.drop;
  return sc;
}

__device__ __forceinline__ plan::PSeg make_kv_pseg(const TcQuerySource& src, const SrcRange& r, bool ex) {
  plan::PSeg s;
  s.expl_ok = ex;
  const Side& sd = src.q.side;
  s.n_img = (ex && sd.id_rule == IDR_EXPLICIT && sd.npr > 0) ? sd.npr * sd.npr : 0;
  s.c_begin = r.ib;
  s.c_end = r.ie;
  s.c_len = src.q.rows.len;
  s.band = src.q.band != 0;
  s.radius = src.q.radius;
  s.mask_rule = sd.mask_rule;
  s.id_rule = src.q.rows.R > 0 ? sd.id_rule : IDR_NONE;
  s.D = sd.max_distance;
  s.R = src.q.rows.R;
  s.diag_ok = (2 * s.D + 1 <= s.R);   // the key-centric DIAG form gathers by id: no slot table involved
  s.rows_are_keys = true;
  s.c_eid = sd.q_eid;
  s.c_eid_stride = sd.q_len;
  // QSENT: the sentence id belongs to the query = column side; KSENT: to the key = row side
  s.col_sent = (s.id_rule == IDR_CROSS_QSENT);
  s.c_sent = s.col_sent ? sd.sent : nullptr;
  s.c_sent_stride = sd.sent_len;
  return s;
}

__device__ __forceinline__ void key_loads(KeyC& kc, const SrcC& sc, int b, int j, bool key_ok) {
  kc.k_e = 0;
  kc.k_sent = -1;
  if (key_ok && sc.mask_rule == MR_EXAMPLE_ID) kc.k_e = __ldg(sc.sd->k_eid + (int64_t)b * sc.sd->k_len + j);
  if (key_ok && sc.id_rule == IDR_CROSS_KSENT) kc.k_sent = __ldg(sc.sd->sent + (int64_t)b * sc.sd->sent_len + j);
}

// Generic per-element evaluation: liveness, mask, relative id (or -1).
__device__ __forceinline__ void eval_generic_kv(const SrcC& sc, const KeyC& kc, int b, int i, int j, bool key_ok,
                                                int qe_i, int qs_i, bool& live, bool& ok, int& id) {
  const Side& sd = *sc.sd;
  const int off = j - i;
  ok = true;
  id = -1;
  live = key_ok && i < sc.ie && (!sc.band || (off <= sc.radius && off >= -sc.radius));
  if (!live) return;
  const int col = sc.band ? off + sc.radius : j;
  switch (sc.mask_rule) {
    case MR_EXPLICIT: ok = __ldg(sd.mask + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col) != 0; break;
    case MR_EXAMPLE_ID: ok = (qe_i == kc.k_e); break;
    default: break;
  }
  switch (sc.id_rule) {
    case IDR_EXPLICIT: id = __ldg(sd.ids + (int64_t)b * sd.sb + (int64_t)i * sd.sq + col); break;
    case IDR_1D: id = rel_id_1d(off, sc.D); break;
    case IDR_CROSS_QSENT: id = 2 * sc.D + 1 + (qs_i == j ? 1 : 0); break;
    case IDR_CROSS_KSENT: id = 2 * sc.D + 1 + (kc.k_sent == i ? 1 : 0); break;
    case IDR_2D: id = rel_id_2d(i, j, sd.npr, sd.core, sc.D); break;
    default: break;
  }
  if (id >= sc.R) id = -1;
}

template <int NP, int SETS, bool SLIM>
constexpr int bk_threads() { return (4 * NP * SETS + (SLIM ? 2 : 3)) * 32; }

// NP threads per key row inside a warp set (each owns W = 64 / NP query columns of a chunk); SETS
// warp sets take alternate chunks (S^T / dP^T are double-buffered by chunk parity, so set s owns
// buffer s).  A planner warp classifies the (quadrant, group) pairs of each chunk (tc_plan.cuh); the
// per-query constants arrive exponent-ready in the row records the query-centric pass published
// (TcBwdQParams::rec_ws), so the common element costs
//     p = ex2(fma(x, scale*log2e, rec[i][4 + id])),  ds = p * (dp - rec[i][2]).
template <int NP, int SETS, bool SLIM, bool EX = false, bool DROP = false, bool ABSORB = false>
__global__ void __launch_bounds__(bk_threads<NP, SETS, SLIM>(), SLIM ? 2 : 1)
tc_bwd_kv_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                 const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_do0,
                 const __grid_constant__ CUtensorMap map_q1, const __grid_constant__ CUtensorMap map_do1,
                 const __grid_constant__ TcBwdKVParams p) {
  using namespace bk;
  using C = Cfg<SLIM>;
  constexpr int SM_K = C::SM_K, SM_V = C::SM_V, SM_QD = C::SM_QD, SM_REC = C::SM_REC, SM_PLAN = C::SM_PLAN,
                SM_BAR = C::SM_BAR, RSMAX = C::RSMAX;
  constexpr uint32_t T_S = C::T_S, T_DP = C::T_DP, T_DV = C::T_DV, T_DK = C::T_DK;
  constexpr int W = 64 / NP;                     // query columns per elementwise thread and chunk
  constexpr int WS = SLIM ? 16 : W;              // ... processed in sub-slices of WS (register budget)
  constexpr int NEW = 128 * NP;                  // elementwise threads per set
  constexpr int WP = 4 * NP * SETS, WM = WP + 1, WH = SLIM ? WP : WP + 2;   // SLIM: producer == planner
  static_assert(SETS == 1 || SETS == 2, "chunk buffers are double-buffered");
  static_assert(!SLIM || SETS == 1, "the slim configuration has a single S^T / dP^T buffer");
  // chunk c uses buffer BUF(c) in its PH(c)-th use
  auto BUF = [](int c) { return SLIM ? 0 : (c & 1); };
  auto PH = [](int c) { return SLIM ? (c & 1) : ((c >> 1) & 1); };
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: keeps the shared address space
  // (LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  plan::ChunkPlan* plans = reinterpret_cast<plan::ChunkPlan*>(smem + SM_PLAN);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * TM;

  if (tid == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&bars->qd_full[s], 1);
      mbar_init(&bars->qd_empty[s], 1);
    }
    for (int s = 0; s < NPL; ++s) {
      mbar_init(&bars->pl_full[s], 1);
      mbar_init(&bars->pl_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->sdp_full[s], 1);
      mbar_init(&bars->pds_full[s], NEW);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == WM) tmem_alloc<C::TCOLS>(&bars->tmem_base);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  const SrcRange r0 = src_range(p.src[0], j0);
  SrcRange r1{0, 0, 0};
  if (p.nsrc > 1) r1 = src_range(p.src[1], j0);
  const int nchunks = r0.n + r1.n;

  // One query chunk: Q and dO tiles plus the 64-row record block (one elected lane).
  auto load_chunk = [&](int c) {
    const int st = c % NST;
    TRACE(5, 2 * c);
    mbar_wait(&bars->qd_empty[st], ((c / NST) & 1) ^ 1);
    TRACE(5, 2 * c + 1);
    const bool first = c < r0.n;
    const TcQuerySource& src = first ? p.src[0] : p.src[1];
    const int q0 = first ? r0.ib + c * TN : r1.ib + (c - r0.n) * TN;
    const int rs = src.rw + 4;
    uint8_t* qs = smem + SM_QD + st * (2 * TN * 128);
    const int64_t prow = ((int64_t)(b * p.H + h) * (src.lp >> 6) + (q0 >> 6)) * RSF;   // record block of q0
    const uint32_t rec_bytes = RSF * rs * 4;
    mbar_arrive_expect_tx(&bars->qd_full[st], 2 * TN * 128 + rec_bytes);
    tma_load_4d(qs, first ? &map_q0 : &map_q1, &bars->qd_full[st], 0, q0, h, b);
    tma_load_4d(qs + TN * 128, first ? &map_do0 : &map_do1, &bars->qd_full[st], 0, q0, h, b);
    bulk_g2s(smem + SM_REC + st * RSF * RSMAX * 4, src.rec_ws + prow * rs, rec_bytes, &bars->qd_full[st]);
  };
  auto run_planner = [&](auto pre) {
    const plan::PSeg ps0 = make_kv_pseg(p.src[0], r0, EX);
    const plan::PSeg ps1 = make_kv_pseg(p.nsrc > 1 ? p.src[1] : p.src[0], r1, EX);
    const Side* ks_side = nullptr;   // row-side (key) sentence ids: rule KSENT
    if (ps0.id_rule == IDR_CROSS_KSENT) ks_side = &p.src[0].q.side;
    if (p.nsrc > 1 && ps1.id_rule == IDR_CROSS_KSENT) ks_side = &p.src[1].q.side;
    plan::planner_loop<NPL, TN>(ps0, ps1, r0.n, r1.n, r0.ib, r1.ib, b, j0, ks_side ? ks_side->sent : nullptr,
                                ks_side ? ks_side->sent_len : 0, p.len, plans, bars->pl_full, bars->pl_empty, lane,
                                pre);
  };

  if (warp == WP) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->kv_full, 2 * TM * 128);
      tma_load_4d(smem + SM_K, &map_k, &bars->kv_full, 0, j0, h, b);
      tma_load_4d(smem + SM_V, &map_v, &bars->kv_full, 0, j0, h, b);
    }
    if (SLIM) {
      // producer and planner in one warp: the chunk's loads go out, then the chunk is classified
      run_planner([&](int c) {
        if (lane == 0) load_chunk(c);
        __syncwarp();
      });
    } else if (lane == 0) {
      for (int c = 0; c < nchunks; ++c) load_chunk(c);
    }
  } else if (warp == WM) {
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_bf16(TM, TN, 0, 0);
      const uint32_t idesc_acc = make_idesc_bf16(TM, 64, 0, 1);
      const uint32_t k_addr = smem_u32(smem + SM_K), v_addr = smem_u32(smem + SM_V);
      mbar_wait(&bars->kv_full, 0);
      tc_fence_after_sync();
      auto issue_sdp = [&](int c) {
        const int st = c % NST;
        mbar_wait(&bars->pl_full[c % NPL], (c / NPL) & 1);   // relayed to the elementwise warps by sdp_full
        mbar_wait(&bars->qd_full[st], (c / NST) & 1);
        TRACE(2, 4 * c);
        tc_fence_after_sync();
        const uint32_t q_addr = smem_u32(smem + SM_QD + st * (2 * TN * 128));
        const uint32_t do_addr = q_addr + TN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // S^T = K . Q_c^T
          umma_ss(tmem + T_S + BUF(c) * 64, sdesc(k_addr).at(kk * 32),
                  sdesc(q_addr).at(kk * 32), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dP^T = V . dO_c^T
          umma_ss(tmem + T_DP + BUF(c) * 64, sdesc(v_addr).at(kk * 32),
                  sdesc(do_addr).at(kk * 32), idesc_s, kk > 0);
        umma_commit(&bars->sdp_full[BUF(c)]);
        TRACE(2, 4 * c + 1);
      };
      auto issue_acc = [&](int pc) {
        const int st = pc % NST;
        mbar_wait(&bars->pds_full[BUF(pc)], PH(pc));
        TRACE(2, 4 * pc + 2);
        tc_fence_after_sync();
        mbar_arrive(&bars->pl_empty[pc % NPL]);   // every elementwise thread is done with plan pc
        const uint32_t q_addr = smem_u32(smem + SM_QD + st * (2 * TN * 128));
        const uint32_t do_addr = q_addr + TN * 128;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dV += P^T . dO_c
          umma_ts(tmem + T_DV, tmem + T_S + BUF(pc) * 64 + ((16 * kk) / W) * W + ((16 * kk) % W) / 2,
                  sdesc(do_addr).at(kk * 2048), idesc_acc, (pc > 0 || kk > 0));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dK += dS^T . Q_c
          umma_ts(tmem + T_DK, tmem + T_DP + BUF(pc) * 64 + ((16 * kk) / W) * W + ((16 * kk) % W) / 2,
                  sdesc(q_addr).at(kk * 2048), idesc_acc, (pc > 0 || kk > 0));
        umma_commit(&bars->qd_empty[st]);
        TRACE(2, 4 * pc + 3);
        if (pc == nchunks - 1) umma_commit(&bars->acc_full);
      };
      // double-buffered: S/dP of chunk c run ahead of the accumulation of chunk c-1; single buffer
      // (SLIM): the accumulation MMAs that read P^T / dS^T are issued first (tensor-core MMAs of one
      // thread execute in issue order, so the overwrite cannot overtake the read)
      for (int c = 0; c <= nchunks; ++c) {
        if (SLIM) {
          if (c >= 1) issue_acc(c - 1);
          if (c < nchunks) issue_sdp(c);
        } else {
          if (c < nchunks) issue_sdp(c);
          if (c >= 1) issue_acc(c - 1);
        }
      }
    }
  } else if (!SLIM && warp == WH) {
    // ===================== planner =====================
    run_planner(plan::NoPre());
  } else {
    // ===================== elementwise warps =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int set = warp / (4 * NP);            // which chunk parity this warp serves
    const int part = (warp >> 2) % NP;
    const int grp = (part * W) / 32;            // which 32-column group of the chunk the slice lies in
    const int j = j0 + row;
    const bool key_ok = j < p.len;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const float scale2 = p.scale * LOG2E;
    // `neg` honoured literally (see tc_fwd.cu / the query-centric pass): in literal mode a masked element
    // evaluates like an unmasked one with neg * log2e added to its exponent.
    const bool lit = ABSORB ? false : fabsf(p.neg) <= 1e5f;
    const float negl2 = p.neg * LOG2E;
    // One call per query source (inlined twice: no per-field selects inside the chunk loop).
    // Chunks [c_begin, c_end) belong to this source; this warp set handles those with c % SETS == set.
    auto run_chunks = [&](const SrcC sc, const KeyC kc, int c_begin, int c_end, int ib) {
      int c = c_begin + ((set - c_begin) % SETS + SETS) % SETS;
      const bool mre = sc.mask_rule == MR_EXAMPLE_ID;
      // dropout: keep(query i, this key) = hash(salt + i * 0x10001 + col); per thread the key is fixed
      const uint32_t dthr = sc.drop.thr;
      const float dinv = sc.drop.inv_keep;
      const uint32_t dkey = DROP ? dropout_salt(sc.drop, (uint32_t)(b * p.H + h)) + (uint32_t)(sc.col_base + j) : 0u;
      auto keep_at = [&](int i) -> bool { return mix_elem(dkey + (uint32_t)i * 0x00010001U) >= dthr; };
#pragma unroll 1
      for (; c < c_end; c += SETS) {
        const int st = c % NST;
        const int g0w = ib + (c - c_begin) * TN + part * W;   // first query of this thread's slice
        const uint32_t t_sw = tmem + T_S + BUF(c) * 64 + lane_sel + part * W;
        const uint32_t t_dpw = tmem + T_DP + BUF(c) * 64 + lane_sel + part * W;
        const plan::ChunkPlan* cp = plans + (c % NPL);
        if (tid == 0) TRACE(0, 4 * c);
        // the MMA warp issued S^T_c / dP^T_c only after plan c had been published and the row
        // records of the chunk had landed
        mbar_wait_warp(&bars->sdp_full[BUF(c)], PH(c));
        if (tid == 0) TRACE(0, 4 * c + 1);
        tc_fence_after_sync();
        // record fields of this thread's query slice: field f of query x at rec[f * RSF + x]
        const float* recw = reinterpret_cast<const float*>(smem + SM_REC + st * RSF * RSMAX * 4) + part * W;
        const int32_t* cew = cp->ce + part * W;
        const int32_t* csw = cp->cs + part * W;
        const uint32_t w0 = cp->q[quad][grp];
        const int ce0 = (int)cp->q[quad][2 + grp];
        int mode = (int)(w0 & 0xffu);
        const bool mask_pe = (w0 & plan::F_MASK_PE) != 0;
        const bool masked = mre && !mask_pe && (kc.k_e != ce0);
        if (lit && mode == plan::FAST && __any_sync(0xffffffffu, masked)) mode = plan::EDGE;
        const int ccls = (int)((w0 >> 8) & 0xffu);
        // record field of the group's constant class
        int coff = ccls == plan::C_POS ? 4 + sc.D : (ccls == plan::C_NEG ? 4 + 2 * sc.D : (ccls == plan::C_CROSS ? 5 + 2 * sc.D : 0));
        if (EX && ccls == plan::C_MODAL) {   // the group's queries are all text or all image
          const int idm = plan::modal_id(g0w, sc.n_img, sc.D);
          coff = idm < sc.R ? 4 + idm : 0;
        }
        // the thread's W columns are processed in sub-slices of WS (SLIM: 16, to stay within the
        // register budget of two CTAs per SM); packed results of sub-slice hh land in the first half
        // of the columns already consumed
#pragma unroll 1
        for (int hh = 0; hh < W / WS; ++hh) {
          const int g0 = g0w + hh * WS;
          const uint32_t t_s = t_sw + hh * WS, t_dp = t_dpw + hh * WS;
          const float* rec = recw + hh * WS;
          const int32_t* ce = cew + hh * WS;
          const int32_t* cs = csw + hh * WS;
          uint32_t p_pk[WS / 2], ds_pk[WS / 2];
          if (mode == plan::DEAD) {
  #pragma unroll
            for (int x = 0; x < WS / 2; ++x) { p_pk[x] = 0u; ds_pk[x] = 0u; }
          } else if (mode == plan::GEN) {
            // real loop, TMEM as dynamically indexed scratch: one copy of the generic code.
            // Leaves p (fp32) in the S^T column and ds in the dP^T column.
  #pragma unroll 1
            for (int ii = 0; ii < WS; ++ii) {
              const uint32_t raw = tmem_ld1(t_s + ii);
              const uint32_t dpr = tmem_ld1(t_dp + ii);
              tmem_wait_ld();
              bool live, ok;
              int id;
              eval_generic_kv(sc, kc, b, g0 + ii, j, key_ok, ce[ii], cs[ii], live, ok, id);
              float pv = 0.f, ds = 0.f;
              if (live) {
                const float* r = rec + ii;
                const float eu = fmaf(__uint_as_float(raw), scale2, r[(id >= 0 ? 4 + id : 0) * RSF]);
                pv = ex2(ok ? eu : (lit ? eu + negl2 : r[RSF]));
                float dpx = __uint_as_float(dpr);
                if constexpr (DROP) {
                  const bool kp = keep_at(g0 + ii);
                  dpx = kp ? dpx * dinv : 0.f;
                  ds = pv * (dpx - r[2 * RSF]);
                  pv = kp ? pv : 0.f;    // dV takes the dropped-out probabilities (1 / (1 - p) joins the epilogue)
                } else {
                  ds = pv * (dpx - r[2 * RSF]);
                }
              }
              __syncwarp();   // the generic evaluation diverges per row; tcgen05.st needs the converged warp
              tmem_st1(t_s + ii, __float_as_uint(pv));
              tmem_st1(t_dp + ii, __float_as_uint(ds));
            }
            tmem_wait_st();
            uint32_t v[WS];
            tmem_ldN(t_s, v);
            tmem_wait_ld();
  #pragma unroll
            for (int x = 0; x < WS / 2; ++x) p_pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
            tmem_ldN(t_dp, v);
            tmem_wait_ld();
  #pragma unroll
            for (int x = 0; x < WS / 2; ++x) ds_pk[x] = pack_bf16x2(__uint_as_float(v[2 * x]), __uint_as_float(v[2 * x + 1]));
          } else {
            uint32_t v[WS], w[WS];
            tmem_ldN(t_s, v);
            tmem_ldN(t_dp, w);
            tmem_wait_ld();
            // unmasked: exponent = x * scale2 + field(4 + id); masked: the per-query constant field 1
            const float* dl = rec + 2 * RSF;
            if (mode == plan::FAST) {
              const float gmul = masked ? 0.f : scale2;
              const float* cc = rec + (masked ? 1 : coff) * RSF;
  #pragma unroll
              for (int x = 0; x < WS / 2; ++x) {
                const float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), gmul, cc[2 * x]));
                const float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), gmul, cc[2 * x + 1]));
                if constexpr (DROP) {
                  const bool k0 = keep_at(g0 + 2 * x), k1 = keep_at(g0 + 2 * x + 1);
                  p_pk[x] = pack_bf16x2(k0 ? p0 : 0.f, k1 ? p1 : 0.f);
                  ds_pk[x] = pack_bf16x2(p0 * ((k0 ? __uint_as_float(w[2 * x]) * dinv : 0.f) - dl[2 * x]),
                                         p1 * ((k1 ? __uint_as_float(w[2 * x + 1]) * dinv : 0.f) - dl[2 * x + 1]));
                } else {
                p_pk[x] = pack_bf16x2(p0, p1);
                ds_pk[x] = pack_bf16x2(p0 * (__uint_as_float(w[2 * x]) - dl[2 * x]), p1 * (__uint_as_float(w[2 * x + 1]) - dl[2 * x + 1]));
                }
              }
            } else {
              // 1. exponent per element, in place (one code copy per form: `mode` is warp-uniform)
              float t[WS];
              const int d0 = j - g0;   // offset(key - query) = d0 - x
              if (mode == plan::EDGE) {
                const float* cc = rec + coff * RSF;
  #pragma unroll
                for (int x = 0; x < WS; ++x) t[x] = fmaf(__uint_as_float(v[x]), scale2, cc[x]);
              } else if (mode == plan::DIAG) {
  #pragma unroll
                for (int x = 0; x < WS; ++x) {
                  const int o = min(max(d0 - x, -sc.D), sc.D);
                  t[x] = fmaf(__uint_as_float(v[x]), scale2, rec[(4 + (o >= 0 ? o : sc.D - o)) * RSF + x]);
                }
              } else if (EX && mode == plan::EXPL) {
                // explicit int32 tensors [B, Lq, W]: entry of (query g0 + x, key j); consecutive keys
                // (= lanes) are adjacent in memory, so every load is one coalesced line per warp
                const Side& sd = *sc.sd;
                int ilo = 0, ihi = min(WS, sc.ie - g0);
                if (sc.band) {
                  ilo = max(ilo, d0 - sc.radius);
                  ihi = min(ihi, d0 + sc.radius + 1);
                }
                if (!key_ok) ihi = ilo;
                const uint32_t live = plan::span_bits(ilo, ihi);
                const int64_t est = sd.sq - (sc.band ? 1 : 0);   // next query, same key
                const int64_t eoff = (int64_t)b * sd.sb + (int64_t)g0 * sd.sq + (sc.band ? d0 + sc.radius : j);
                const bool has_i = sc.id_rule == IDR_EXPLICIT, has_m = sc.mask_rule == MR_EXPLICIT;
#pragma unroll
                for (int x = 0; x < WS; ++x) {
                  const bool lv = (live >> x) & 1u;
                  int id = -1, ok = 1;
                  if (lv && has_i) id = __ldg(sd.ids + eoff + x * est);
                  if (lv && has_m) ok = __ldg(sd.mask + eoff + x * est);
                  const int f = (unsigned)id < (unsigned)sc.R ? 4 + id : 0;
                  const float eu = fmaf(__uint_as_float(v[x]), scale2, rec[f * RSF + x]);
                  t[x] = ok != 0 ? eu : (lit ? eu + negl2 : rec[RSF + x]);
                }
              } else {
                const float* c0 = rec + (5 + 2 * sc.D) * RSF;
                if (mode == plan::QS) {          // row-side sentence: key j belongs to query (k_sent)
                  const int sp = kc.k_sent - g0;
  #pragma unroll
                  for (int x = 0; x < WS; ++x) t[x] = fmaf(__uint_as_float(v[x]), scale2, c0[(sp == x ? RSF : 0) + x]);
                } else {                         // KS: column-side sentence: query x's sentence is key j
  #pragma unroll
                  for (int x = 0; x < WS; ++x) t[x] = fmaf(__uint_as_float(v[x]), scale2, c0[(cs[x] == j ? RSF : 0) + x]);
                }
              }
              // 2. masked elements take the per-query constant (field 1)
              const float* lp = rec + RSF;
              if (mask_pe) {
  #pragma unroll
                for (int x = 0; x < WS; ++x) t[x] = (ce[x] != kc.k_e) ? (lit ? t[x] + negl2 : lp[x]) : t[x];
              } else if (mre && __any_sync(0xffffffffu, masked)) {
  #pragma unroll
                for (int x = 0; x < WS; ++x) t[x] = masked ? (lit ? t[x] + negl2 : lp[x]) : t[x];
              }
              // 3. probabilities; EDGE / EXPL: dead columns may carry garbage records -> select, not multiply
              if (mode == plan::EDGE || (EX && mode == plan::EXPL)) {
                int ilo = 0, ihi = min(WS, sc.ie - g0);
                if (sc.band) {
                  ilo = max(ilo, d0 - sc.radius);
                  ihi = min(ihi, d0 + sc.radius + 1);
                }
                if (!key_ok) ihi = ilo;
                const unsigned span = (unsigned)max(ihi - ilo, 0);
  #pragma unroll
                for (int x = 0; x < WS / 2; ++x) {
                  const bool l0 = (unsigned)(2 * x - ilo) < span, l1 = (unsigned)(2 * x + 1 - ilo) < span;
                  float p0 = l0 ? ex2(t[2 * x]) : 0.f, p1 = l1 ? ex2(t[2 * x + 1]) : 0.f;
                  float w0f = __uint_as_float(w[2 * x]), w1f = __uint_as_float(w[2 * x + 1]);
                  bool k0 = true, k1 = true;
                  if constexpr (DROP) {
                    k0 = keep_at(g0 + 2 * x);
                    k1 = keep_at(g0 + 2 * x + 1);
                    w0f = k0 ? w0f * dinv : 0.f;
                    w1f = k1 ? w1f * dinv : 0.f;
                  }
                  const float e0 = l0 ? p0 * (w0f - dl[2 * x]) : 0.f;
                  const float e1 = l1 ? p1 * (w1f - dl[2 * x + 1]) : 0.f;
                  p_pk[x] = pack_bf16x2(k0 ? p0 : 0.f, k1 ? p1 : 0.f);
                  ds_pk[x] = pack_bf16x2(e0, e1);
                }
              } else {
  #pragma unroll
                for (int x = 0; x < WS / 2; ++x) {
                  const float p0 = ex2(t[2 * x]), p1 = ex2(t[2 * x + 1]);
                  if constexpr (DROP) {
                    const bool k0 = keep_at(g0 + 2 * x), k1 = keep_at(g0 + 2 * x + 1);
                    p_pk[x] = pack_bf16x2(k0 ? p0 : 0.f, k1 ? p1 : 0.f);
                    ds_pk[x] = pack_bf16x2(p0 * ((k0 ? __uint_as_float(w[2 * x]) * dinv : 0.f) - dl[2 * x]),
                                           p1 * ((k1 ? __uint_as_float(w[2 * x + 1]) * dinv : 0.f) - dl[2 * x + 1]));
                  } else {
                  p_pk[x] = pack_bf16x2(p0, p1);
                  ds_pk[x] = pack_bf16x2(p0 * (__uint_as_float(w[2 * x]) - dl[2 * x]), p1 * (__uint_as_float(w[2 * x + 1]) - dl[2 * x + 1]));
                  }
                }
              }
            }
          }
          tmem_stN(t_sw + hh * (WS / 2), p_pk);
          tmem_stN(t_dpw + hh * (WS / 2), ds_pk);
        }
        if (tid == 0) TRACE(0, 4 * c + 2);
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->pds_full[BUF(c)]);
        if (tid == 0) TRACE(0, 4 * c + 3);
      }
    };
    if constexpr (SLIM) {
      // ONE copy of the chunk loop, source context built once per source outside it: the smaller code
      // wins on the few-chunk tiles (long keys 0.522 -> 0.498 ms) ...
#pragma unroll 1
      for (int si = 0; si < p.nsrc; ++si) {
        KeyC kc;
        const SrcC sc = make_srcc(p.src[si], si ? r1 : r0);
        key_loads(kc, sc, b, j, key_ok);
        run_chunks(sc, kc, si ? r0.n : 0, si ? nchunks : r0.n, si ? r1.ib : r0.ib);
      }
    } else {
      // ... and loses on the 68-chunk tiles of the global keys (0.232 -> 0.255 ms): one inlined copy per source
      KeyC kc;
      const SrcC sc0 = make_srcc(p.src[0], r0);
      key_loads(kc, sc0, b, j, key_ok);
      run_chunks(sc0, kc, 0, r0.n, r0.ib);
      if (p.nsrc > 1) {
        const SrcC sc1 = make_srcc(p.src[1], r1);
        key_loads(kc, sc1, b, j, key_ok);
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
        const float vs = DROP ? p.src[0].q.drop.inv_keep : 1.f;   // every source shares the dropout rate
        w.x = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 0]) * vs, __uint_as_float(dv_raw[8 * x + 1]) * vs);
        w.y = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 2]) * vs, __uint_as_float(dv_raw[8 * x + 3]) * vs);
        w.z = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 4]) * vs, __uint_as_float(dv_raw[8 * x + 5]) * vs);
        w.w = pack_bf16x2(__uint_as_float(dv_raw[8 * x + 6]) * vs, __uint_as_float(dv_raw[8 * x + 7]) * vs);
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
  if (warp == WM) tmem_dealloc<C::TCOLS>(tmem);
}

inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }
inline int pad_rows(int len) { return (len + TN - 1) / TN * TN + TN; }
inline int pad4(int r) { return (r + 3) / 4 * 4; }

}  // namespace

// Workspace of one row set for the tcgen05 backward (rowstat + row records), bytes.
size_t tc_bwd_rows_ws_bytes(int B, int H, int len, int R) {
  const size_t rows = (size_t)B * H * pad_rows(len);
  // rowstat [rows] float4 + row records [rows / 64][R4 + 4][RSF] floats
  return align256(rows * sizeof(float4)) + align256(rows / 64 * RSF * (pad4(R > 0 ? R : 0) + 4) * sizeof(float) + 256);
}

bool tc_bwd_q_supported(const BwdQArgs& a, int dtype, int d) {
  FwdArgs f{};
  f.rows = a.rows;
  f.seg[0] = a.seg[0];
  f.seg[1] = a.seg[1];
  f.nseg = a.nseg;
  f.out = a.out;
  f.B = a.B;
  f.H = a.H;
  return tc_fwd_args_supported(f, dtype, d) && tc_t4_ok(a.d_out, a.B, a.rows.len, a.H) &&
         tc_t4_ok(a.d_q, a.B, a.rows.len, a.H);
}

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace_b(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace_b, sizeof(unsigned long long) * 6 * 256);
}
#endif

namespace {
// launch configuration -> kernel instantiation (EX: explicit int32 side inputs, DROP: dropout on)
template <int SETS, bool SLIM>
struct BqLaunch {
  template <bool EX, bool DROP>
  static void go(dim3 grid, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mdo, const CUtensorMap& mk0,
                 const CUtensorMap& mv0, const CUtensorMap& mk1, const CUtensorMap& mv1, const CUtensorMap& me,
                 const TcBwdQParams& p) {
    tc_bwd_q_kernel<SETS, SLIM, EX, DROP><<<grid, bq_threads<SETS, SLIM>(), bq::Cfg<SLIM>::SM_ALLOC, st>>>(
        mq, mdo, mk0, mv0, mk1, mv1, me, p);
  }
  static bool absorbed(const TcBwdQParams& p) { return !(fabsf(p.a.neg) <= 1e5f); }
  static void run(bool ex, bool dr, dim3 grid, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mdo,
                  const CUtensorMap& mk0, const CUtensorMap& mv0, const CUtensorMap& mk1, const CUtensorMap& mv1,
                  const CUtensorMap& me, const TcBwdQParams& p) {
    // the two-warp-set configuration does not carry the EXPL form (19 warps: 96 registers per thread, the form
    // spilled 1.9 KB there); the launcher sends explicit side inputs to the one-warp-set configuration
    if constexpr (SETS == 1) {
      if (ex && dr) return go<true, true>(grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
      if (ex) return go<true, false>(grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
    }
    if (dr) go<false, true>(grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
    else if (absorbed(p))
      tc_bwd_q_kernel<SETS, SLIM, false, false, true><<<grid, bq_threads<SETS, SLIM>(), bq::Cfg<SLIM>::SM_ALLOC, st>>>(
          mq, mdo, mk0, mv0, mk1, mv1, me, p);
    else go<false, false>(grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
  }
  static cudaError_t attrs() {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kernel) {
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bq::Cfg<SLIM>::SM_ALLOC);
    };
    set(tc_bwd_q_kernel<SETS, SLIM, false, false>);
    set(tc_bwd_q_kernel<SETS, SLIM, false, true>);
    if constexpr (SETS == 1) {
      set(tc_bwd_q_kernel<SETS, SLIM, true, false>);
      set(tc_bwd_q_kernel<SETS, SLIM, true, true>);
    }
    set(tc_bwd_q_kernel<SETS, SLIM, false, false, true>);
    return e;
  }
};
template <int NP, int SETS, bool SLIM>
struct BkLaunch {
  template <bool EX, bool DROP>
  static void go(dim3 grid, cudaStream_t st, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap* mq,
                 const CUtensorMap* mdo, const TcBwdKVParams& p) {
    tc_bwd_kv_kernel<NP, SETS, SLIM, EX, DROP><<<grid, bk_threads<NP, SETS, SLIM>(), bk::Cfg<SLIM>::SM_ALLOC, st>>>(
        mk, mv, mq[0], mdo[0], mq[1], mdo[1], p);
  }
  static void run(bool ex, bool dr, dim3 grid, cudaStream_t st, const CUtensorMap& mk, const CUtensorMap& mv,
                  const CUtensorMap* mq, const CUtensorMap* mdo, const TcBwdKVParams& p) {
    if (ex && dr) go<true, true>(grid, st, mk, mv, mq, mdo, p);
    else if (ex) go<true, false>(grid, st, mk, mv, mq, mdo, p);
    else if (dr) go<false, true>(grid, st, mk, mv, mq, mdo, p);
    else if (!(fabsf(p.neg) <= 1e5f))
      tc_bwd_kv_kernel<NP, SETS, SLIM, false, false, true>
          <<<grid, bk_threads<NP, SETS, SLIM>(), bk::Cfg<SLIM>::SM_ALLOC, st>>>(mk, mv, mq[0], mdo[0], mq[1], mdo[1], p);
    else go<false, false>(grid, st, mk, mv, mq, mdo, p);
  }
  static cudaError_t attrs() {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kernel) {
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bk::Cfg<SLIM>::SM_ALLOC);
    };
    set(tc_bwd_kv_kernel<NP, SETS, SLIM, false, false>);
    set(tc_bwd_kv_kernel<NP, SETS, SLIM, true, false>);
    set(tc_bwd_kv_kernel<NP, SETS, SLIM, false, true>);
    set(tc_bwd_kv_kernel<NP, SETS, SLIM, true, true>);
    set(tc_bwd_kv_kernel<NP, SETS, SLIM, false, false, true>);
    return e;
  }
};
}  // namespace

int tc_launch_bwd_q(const BwdQArgs& a, void* ws, cudaStream_t st, bool allow_gl2) {
  static PerDeviceOnce once;
  const int ae = once.run([] {
    cudaError_t e = BqLaunch<2, false>::attrs();
    if (e == cudaSuccess) e = BqLaunch<1, false>::attrs();
    if (e == cudaSuccess) e = BqLaunch<1, true>::attrs();
    return (int)e;
  });
  if (ae) return ae;
  TcBwdQParams p;
  p.a = a;
  const int R = a.rows.R;
  p.rpad = R > 0 ? (R + 15) / 16 * 16 : 0;
  p.rw = pad4(R);
  p.lp = pad_rows(a.rows.len);
  char* w = reinterpret_cast<char*>(ws);
  const size_t rows = (size_t)a.B * a.H * p.lp;
  p.rowstat = reinterpret_cast<float4*>(w);
  p.rec_ws = reinterpret_cast<float*>(w + align256(rows * sizeof(float4)));
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
    if (threads >= (int64_t)1 << 32) return MLT_ERR_UNSUPPORTED;   // 32-bit row index in the kernel
    tc_bwd_prep_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a.out, a.d_out, a.stats, p.rowstat, a.B, a.H,
                                                                       a.rows.len, p.lp);
    cudaError_t pe = cudaGetLastError();
    if (pe != cudaSuccess) return (int)pe;
  }
  // compact global-local long rows: the persistent specialised kernel (gl2_bwd_q.cu), same workspace format
  if (allow_gl2 && gl2_bwd_q_long_supported(a, MLT_BF16, 64))
    return gl2_launch_bwd_q_long(a, p.rowstat, p.rec_ws, p.lp, p.rw, st);
  dim3 grid((a.rows.len + TM - 1) / TM, a.H, a.B);
  // chunks per tile: with many chunks (dense global rows) two warp sets on alternate chunks win;
  // with few (long rows: band + G/64) the extra per-tile prologue work does not pay off.
  auto seg_chunks = [](const KeySeg& sg) { return sg.band ? (TM + 2 * sg.radius + TN - 1) / TN : (sg.len + TN - 1) / TN; };
  const int est_chunks = seg_chunks(a.seg[0]) + (a.nseg > 1 ? seg_chunks(a.seg[1]) : 0);
  // 1: one warp set (relative vocabulary > 32), 2: two warp sets on alternate chunks (many chunks:
  // dense global rows), 3: slim, two CTAs per SM (few chunks: long rows = band + G/64).  (A slim
  // variant with two elementwise threads per row was measured slower, 0.75 vs 0.66 ms on the c3_4096
  // long rows, and removed: the tile is bound by the quadrant imbalance of the band chunks and by the
  // per-warp latency chains, not by the number of elementwise warps.)
  int cfg = est_chunks >= 16 ? 2 : 3;
  if (R > 32) cfg = 1;   // slim and two-set bins hold 32 slots
  // explicit int32 side inputs: the instantiations that carry the EXPL form
  const bool ex = side_is_explicit(a.seg[0].side) || (a.nseg > 1 && side_is_explicit(a.seg[1].side));
  const bool dr = a.drop.thr != 0;
  // explicit side inputs (and the 2-D id plane): the one-warp-set configuration.  Against two warp sets it
  // measures equal (explicit global-local step 3.06 vs 3.08 ms) and is free of their 1.9 KB of spills; against
  // the slim configuration it is faster from 8 chunks on (dense S 512 with the 2-D plane, fwd + bwd 1.06 -> 0.93 ms)
  if (ex && est_chunks >= 8) cfg = 1;
  if (cfg == 3) BqLaunch<1, true>::run(ex, dr, grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
  else if (cfg == 2) BqLaunch<2, false>::run(ex, dr, grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
  else BqLaunch<1, false>::run(ex, dr, grid, st, mq, mdo, mk0, mv0, mk1, mv1, me, p);
  return (int)cudaGetLastError();
}

int tc_launch_bwd_kv(const BwdKVArgs& a, void* const ws[2], cudaStream_t st) {
  static PerDeviceOnce once;
  const int ae = once.run([] {
    cudaError_t e = BkLaunch<4, 1, false>::attrs();
    if (e == cudaSuccess) e = BkLaunch<2, 1, true>::attrs();
    return (int)e;
  });
  if (ae) return ae;
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
    t.rec_ws = reinterpret_cast<const float*>(w + align256(rows * sizeof(float4)));
    e |= make_qkv_tensor_map(&mq[s], q.rows.q.ptr, q.rows.q.sb, q.rows.q.sl, q.rows.q.sh, a.B, lq, a.H, TN);
    e |= make_qkv_tensor_map(&mdo[s], q.d_out.ptr, q.d_out.sb, q.d_out.sl, q.d_out.sh, a.B, lq, a.H, TN);
  }
  if (e) return MLT_ERR_UNSUPPORTED;
  dim3 grid((a.len + TM - 1) / TM, a.H, a.B);
  auto src_chunks = [](const QuerySource& q) { return q.band ? (TM + 2 * q.radius + TN - 1) / TN : (q.rows.len + TN - 1) / TN; };
  const int est_chunks = src_chunks(a.src[0]) + (a.nsrc > 1 ? src_chunks(a.src[1]) : 0);
  // Few chunks per tile (long keys: band + G/64): start-up and drain dominate -> the slim
  // configuration with two CTAs per SM (3).  Otherwise one warp set with four threads per row (1): for the
  // many chunks of the global keys it measures 5 % faster than two warp sets on alternate chunks
  // (0.214 vs 0.226 ms on c3_4096, the same ratio at L = 2048 and 8192), and it is the one configuration
  // that holds relative vocabularies > 32.
  const bool slim_ok = p.src[0].rw <= 32 && p.src[1].rw <= 32;
  const bool ex = side_is_explicit(a.src[0].side) || (a.nsrc > 1 && side_is_explicit(a.src[1].side));
  // (explicit side inputs: the slim configuration only below 8 chunks, as in the query-centric pass)
  const int cfg = est_chunks >= (ex ? 8 : 16) ? 1 : (slim_ok ? 3 : 1);
  const bool dr = a.src[0].drop.thr != 0;
  if (cfg == 3) BkLaunch<2, 1, true>::run(ex, dr, grid, st, mk, mv, mq, mdo, p);
  else BkLaunch<4, 1, false>::run(ex, dr, grid, st, mk, mv, mq, mdo, p);
  return (int)cudaGetLastError();
}

}  // namespace mlt
