// tcgen05 forward kernel: one 128-row query tile per CTA, joint online softmax over up to two
// key segments (band / dense), relative-position scores and masks fused into the score tile.
//
// Roles (256 threads, 2 CTAs per SM, 256 TMEM columns each):
//   warps 0-3  softmax: thread t owns query row t == TMEM lane t.  The loop is bound by the MUFU
//              (ex2) rate of the SM (16 / clk, measured by tests/cuda/tmem_bw.cu), so everything
//              else is kept off these warps: per 32-key group they execute
//                pass 1   row maximum of the raw Q.K^T accumulator          (0.5 instr / element)
//                pass 2   p = ex2(fma(x, scale*log2e, c_row)), sum, pack    (3.5 instr / element)
//              with c_row = (rel + mask)*log2e - max*log2e folded per row; the group's evaluation
//              form comes ready-made from the planner warp (tc_plan.cuh).
//   warp 4     TMA producer: Q tile, relative-embedding tile, K/V chunks through a 3-stage ring.
//   warp 5     MMA issuer (one elected lane): S_c = Q.K_c^T (SS), O += P_c.V_c (TS: P from TMEM,
//              V MN-major), allrel = Q.E^T.  S is double-buffered: S_{c+1} runs under softmax_c.
//   warp 6     planner: classifies the (quadrant, group) pairs of each chunk a few chunks ahead.
//   warp 7     idle (completes the warpgroup).
//
// O stays in TMEM for the whole tile and is accumulated by the tensor core (FlashAttention-4
// style lazy rescaling): the running maximum used for the exponentials only moves when the true
// maximum outgrows it by more than 2^RESCALE_LOG2, in which case the owning warp rescales its O
// rows in place.  The true maximum is tracked separately and published in the row statistics.
//
// TMEM map (columns): [0,64) S0/P0, [64,128) S1/P1, [128,192) O, [192,256) allrel.
#include "tc_api.cuh"

#include "mlt_common.cuh"
#include "profile.cuh"
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace mlt {
namespace {

using namespace ptx;

constexpr int TM = 128;       // query rows per tile
constexpr int TN = 64;        // keys per chunk
constexpr int NST = 3;        // K/V ring stages
constexpr int NPL = 4;        // plan ring slots
constexpr int NTHREADS = 256;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t T_O = 128, T_REL = 192;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float M_INIT = -1e30f;         // "no live key so far"
constexpr float RESCALE_LOG2 = 32.f;     // lazy-rescale threshold (log2 units)

constexpr int SM_Q = 0;                          // 16 KB
constexpr int SM_E = SM_Q + TM * 128;            // 8 KB  (<= 64 ids x 128 B)
constexpr int SM_KV = SM_E + 64 * 128;           // NST x (K 8 KB + V 8 KB)
constexpr int SM_REL = SM_KV + NST * 2 * TN * 128;   // [64 slots][128 rows] fp32 = 32 KB
constexpr int SM_PLAN = SM_REL + 64 * TM * 4;        // NPL x ChunkPlan
constexpr int SM_META = SM_PLAN + NPL * (int)sizeof(plan::ChunkPlan);   // 64 x RelMeta
constexpr int SM_BAR = SM_META + 64 * (int)sizeof(plan::RelMeta);
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
  uint64_t q_full, rel_full;
  uint64_t kv_full[NST], kv_empty[NST];
  uint64_t s_full[2], p_full[2], o_full[2];
  uint64_t pl_full[NPL], pl_empty[NPL];
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 256, "barrier block");

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

__device__ __forceinline__ plan::PSeg make_pseg(const KeySeg& sg, const SegRange& r, int R, int pd, bool perm, bool ex) {
  plan::PSeg s;
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
  // compact 2-D layout served through the library's id plane (abi.cu with_ids_plane keeps npr)
  s.n_img = (ex && sg.side.id_rule == IDR_EXPLICIT && sg.side.npr > 0) ? sg.side.npr * sg.side.npr : 0;
  return s;
}

// ---- softmax-side contexts -------------------------------------------------------------------
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
  float relM;          // 2-D layout: text row -> id of an image column, image row -> id of a text column
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
  sc.col_base = sg.col_base;
  sc.n_img = (sg.side.id_rule == IDR_EXPLICIT && sg.side.npr > 0) ? sg.side.npr * sg.side.npr : 0;
  return sc;
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
  rc.relM = 0.f;
  if (on && sc.n_img > 0) rc.relM = rel_at(plan::modal_id(i, sc.n_img, sc.D));
}
__device__ __forceinline__ float rel_const(int ccls, const RowC& rc) {
  return ccls == plan::C_POS ? rc.relP
                             : (ccls == plan::C_NEG ? rc.relN
                                                    : (ccls == plan::C_CROSS ? rc.relX : (ccls == plan::C_MODAL ? rc.relM : 0.f)));
}

// Generic per-element score (any rule).  Dead pairs return -inf.
__device__ __forceinline__ float score_generic(float x, const SegC& sc, const RowC& rc, int b, int i, int row,
                                               bool row_ok, int j, int ke_j, int ks_j, const float* rel_s,
                                               float scale, float neg) {
  const Side& sd = *sc.sd;
  const int off = j - i;
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
  if (id >= 0 && id < sc.R) rel = rel_s[plan::slot_of_id(id, sc.pd, sc.perm) * TM + row];
  float v = fmaf(x, scale, rel);
  if (!ok) v += neg;
  return v;
}

// EX: the instantiation that carries the EXPL form (explicit int32 side inputs); the compact
// instantiation stays free of its code and register pressure.
// ABSORB: literal-`neg` mode compiled out (see tc_bwd.cu)
template <bool EX, bool DROP, bool ABSORB = false>
__global__ void __launch_bounds__(NTHREADS, 2)
tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k0,
              const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_k1,
              const __grid_constant__ CUtensorMap map_v1, const __grid_constant__ CUtensorMap map_e,
              const __grid_constant__ TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the __shared__ array: keeps the shared address space
  // (LDS/STS with 32-bit addresses instead of generic LD/ST with 64-bit address math)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars* bars = reinterpret_cast<Bars*>(smem + SM_BAR);
  float* rel_s = reinterpret_cast<float*>(smem + SM_REL);
  plan::ChunkPlan* plans = reinterpret_cast<plan::ChunkPlan*>(smem + SM_PLAN);
  plan::RelMeta* relmeta = reinterpret_cast<plan::RelMeta*>(smem + SM_META);
  const FwdArgs& a = p.a;
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
      mbar_init(&bars->s_full[s], 1);
      mbar_init(&bars->p_full[s], 128);
      mbar_init(&bars->o_full[s], 1);
    }
    for (int s = 0; s < NPL; ++s) {
      mbar_init(&bars->pl_full[s], 1);
      mbar_init(&bars->pl_empty[s], 1);
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
  const int pd = a.seg[0].side.max_distance;
  const bool perm = (2 * pd + 1 <= R);

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
      TRACE(2, 0);
      mbar_wait(&bars->q_full, 0);
      TRACE(2, 1);
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
      for (int c = 0; c <= nchunks; ++c) {
        if (c < nchunks) {
          const int st = c % NST;
          mbar_wait(&bars->pl_full[c % NPL], (c / NPL) & 1);   // relayed to the softmax warps by s_full
          mbar_wait(&bars->kv_full[st], (c / NST) & 1);
          tc_fence_after_sync();
          const uint32_t k_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + (c & 1) * 64, sdesc(q_addr).at(kk * 32),
                    sdesc(k_addr).at(kk * 32), idesc_s, kk > 0);
          umma_commit(&bars->s_full[c & 1]);
          TRACE(2, 2 + 2 * c);
        }
        if (c >= 1) {
          const int pc = c - 1, st = pc % NST;
          mbar_wait(&bars->p_full[pc & 1], (pc >> 1) & 1);
          tc_fence_after_sync();
          mbar_arrive(&bars->pl_empty[pc % NPL]);   // every softmax thread is done with plan pc
          const uint32_t v_addr = smem_u32(smem + SM_KV + st * (2 * TN * 128) + TN * 128);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ts(tmem + T_O, tmem + (pc & 1) * 64 + (kk >> 1) * 32 + (kk & 1) * 8,
                    sdesc(v_addr).at(kk * 2048), idesc_o, (pc > 0 || kk > 0));
          umma_commit(&bars->o_full[pc & 1]);
          umma_commit(&bars->kv_empty[st]);
          TRACE(2, 3 + 2 * pc);
        }
      }
    }
  } else if (warp == 6) {
    // ===================== planner =====================
    const plan::PSeg ps0 = make_pseg(a.seg[0], r0, R, pd, perm, EX);
    const plan::PSeg ps1 = make_pseg(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm, EX);
    // row-side sentence ranges (only the QSENT rule needs them; both segments share the array)
    const Side* qs_side = nullptr;
    if (ps0.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[0].side;
    if (a.nseg > 1 && ps1.id_rule == IDR_CROSS_QSENT) qs_side = &a.seg[1].side;
    plan::planner_loop<NPL, TN>(ps0, ps1, r0.n, r1.n, r0.kb, r1.kb, b, i0, qs_side ? qs_side->sent : nullptr,
                                qs_side ? qs_side->sent_len : 0, a.rows.len, plans, bars->pl_full, bars->pl_empty, lane);
  } else if (warp < 4) {
    // ===================== softmax / epilogue (warps 0-3) =====================
    const int row = tid;
    const int i = i0 + row;
    const bool row_ok = i < a.rows.len;
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    if (tid == 0) TRACE(0, 0);
    const SegC sc0 = make_segc(a.seg[0], r0, R, pd, perm);
    const SegC sc1 = make_segc(a.nseg > 1 ? a.seg[1] : a.seg[0], r1, R, pd, perm);
    RowC rc0, rc1;
    row_loads(rc0, sc0, b, i, row_ok);
    row_loads(rc1, sc1, b, i, row_ok);
    if (rpad) {
      // allrel (+ bias) * scale -> slot-ordered per-row table in shared memory
      if (tid < 64)
        plan::rel_meta_init(relmeta, tid, reinterpret_cast<const __nv_bfloat16*>(a.rows.bias), a.H, h, R, pd, perm,
                            a.scale);
      named_bar_sync(1, 128);
      mbar_wait_warp(&bars->rel_full, 0);
      if (tid == 0) TRACE(0, 1);
      tc_fence_after_sync();
      plan::rel_table_build(tmem + T_REL + lane_sel, relmeta, rel_s, row, rpad, a.scale, [](int, const float (&)[16]) {});
    }
    row_consts(rc0, sc0, rel_s, relmeta, row, i);   // a thread only reads its own column of rel_s: no barrier
    row_consts(rc1, sc1, rel_s, relmeta, row, i);

    if (tid == 0) TRACE(0, 2);
    const float scale2 = a.scale * LOG2E;
    // The ABI's `neg` is honoured literally.  |neg| > 1e5 (the reference's -1e9): a masked score is
    // the additive constant itself in fp32, which the FAST form folds into a per-row constant.
    // |neg| <= 1e5 ("literal" mode): a masked score still carries x * scale + rel, so FAST groups with
    // masked rows take the per-element EDGE form.  Masked groups are skipped (p == 0 exactly) only when
    // neg is negative enough for exp(neg + 64) to flush to zero, against a real (unmasked) maximum.
    const bool lit = ABSORB ? false : fabsf(a.neg) <= 1e5f;
    const bool skip_ok = a.neg < -200.f;
    const float real_thr = 0.5f * a.neg;
    const uint32_t drow = DROP ? dropout_row_base(dropout_salt(a.drop, (uint32_t)(b * a.H + h)), i) : 0u;
    const uint32_t dthr = a.drop.thr;
    float m = M_INIT;        // maximum the exponentials are taken against (lags the true maximum)
    float m_true = M_INIT;   // running maximum over the chunks that took the two-pass path
    float l = 0.f;
    bool seeded = false;     // warp-uniform: every row of the warp has a reference maximum
    int c = 0;   // chunk counter over both segments
#pragma unroll 1
    for (int sgi = 0; sgi < a.nseg; ++sgi) {
      // segment context copied once: no per-field selects inside the chunk loop
      const SegC sc = sgi ? sc1 : sc0;
      const RowC rc = sgi ? rc1 : rc0;
      const int seg_n = sgi ? r1.n : r0.n;
      const int seg_kb = sgi ? r1.kb : r0.kb;
      const bool mre = sc.mask_rule == MR_EXAMPLE_ID;
#pragma unroll 1
      for (int cc = 0; cc < seg_n; ++cc, ++c) {
        const int key0 = seg_kb + cc * TN;
        const uint32_t t_s = tmem + (c & 1) * 64 + lane_sel;
        const plan::ChunkPlan* cp = plans + (c % NPL);
        if (tid == 0) TRACE(0, 4 + 3 * c);
        // the MMA warp issued S_c only after plan c had been published: one wait covers both
        mbar_wait_warp(&bars->s_full[c & 1], (c >> 1) & 1);
        const uint4 pq = *reinterpret_cast<const uint4*>(&cp->q[warp][0]);
        const uint2 pw = make_uint2(pq.x, pq.y);
        const int2 pe = make_int2((int)pq.z, (int)pq.w);
        if (tid == 0) TRACE(0, 5 + 3 * c);
        tc_fence_after_sync();
        // ---- optimistic single pass ----
        // Once every row of the warp holds a reference maximum m, a chunk whose groups are all FAST
        // (or dead) is evaluated straight away against m: p = ex2(x * mul + c_row - m * log2e).  Any
        // consistent reference works for a softmax; it only must not overflow.  A row sum that is not
        // a finite number below 2^100 (a score outgrew m by 2^100, e.g. the first unmasked score
        // after masked ones) sends the warp to the exact two-pass path below: nothing has been
        // written yet, so S is still intact.
        {
          const int md0 = (int)(pw.x & 0xffu), md1 = (int)(pw.y & 0xffu);
          // (KS groups -- cross blocks in which some column's sentence is one of the warp's rows -- ride
          // along: the same single FMA per element with the per-row constant selected per column; a mask
          // that changes inside the group keeps the two-pass path.  Measured: global rows 0.198 -> 0.178 ms.
          // The QS form of the long rows does not: it made the long-row tiles 6 % slower.)
          auto single_pass_ok = [&](int md, uint32_t w0) {
            return md == plan::DEAD || md == plan::FAST || (md == plan::KS && !(w0 & plan::F_MASK_PE));
          };
          const bool litblock = lit && mre &&
              __any_sync(0xffffffffu, (md0 != plan::DEAD && rc.q_e != pe.x) || (md1 != plan::DEAD && rc.q_e != pe.y));
          const bool fastable = seeded && single_pass_ok(md0, pw.x) && single_pass_ok(md1, pw.y) && !litblock;
          if (fastable) {
            const float mb = m * LOG2E;
            uint32_t pk0[16], pk1[16];
            float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
            auto group = [&](int g, uint32_t (&pk)[16]) {
              const uint32_t w0 = g ? pw.y : pw.x;
              const bool masked = mre && (rc.q_e != (g ? pe.y : pe.x));   // FAST: mask uniform over the keys
              bool dead = (w0 & 0xffu) == plan::DEAD;
              if (!dead && mre && skip_ok && __all_sync(0xffffffffu, masked && m > real_thr)) dead = true;   // p == 0 exactly
              if (dead) {
#pragma unroll
                for (int x = 0; x < 16; ++x) pk[x] = 0u;
                return;
              }
              const int ccls = (int)((w0 >> 8) & 0xffu);
              const float relc = rel_const(ccls, rc);
              const float gmul = masked ? 0.f : scale2;
              const float gsub = fmaf(relc + (masked ? a.neg : 0.f), LOG2E, -mb);
              const int gmode = (int)(w0 & 0xffu);
              uint32_t v[32];
              tmem_ld32(t_s + 32 * g, v);
              tmem_wait_ld();
              const int dcol = sc.col_base + key0 + 32 * g;   // dropout counter of the group's first key
              auto body = [&](auto subf) {
#pragma unroll
                for (int x = 0; x < 16; x += 2) {
                  float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), gmul, subf(2 * x)));
                  float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), gmul, subf(2 * x + 1)));
                  float p2 = ex2(fmaf(__uint_as_float(v[2 * x + 2]), gmul, subf(2 * x + 2)));
                  float p3 = ex2(fmaf(__uint_as_float(v[2 * x + 3]), gmul, subf(2 * x + 3)));
                  ls0 += p0;
                  ls1 += p1;
                  ls2 += p2;
                  ls3 += p3;
                  if constexpr (DROP) {   // the normaliser is taken before dropout; 1 / (1 - p) joins 1 / l
                    p0 = dropout_keep(drow, dcol + 2 * x, dthr) ? p0 : 0.f;
                    p1 = dropout_keep(drow, dcol + 2 * x + 1, dthr) ? p1 : 0.f;
                    p2 = dropout_keep(drow, dcol + 2 * x + 2, dthr) ? p2 : 0.f;
                    p3 = dropout_keep(drow, dcol + 2 * x + 3, dthr) ? p3 : 0.f;
                  }
                  pk[x] = pack_bf16x2(p0, p1);
                  pk[x + 1] = pack_bf16x2(p2, p3);
                }
              };
              if (gmode == plan::FAST) {   // warp-uniform: the common form stays select-free
                body([&](int) { return gsub; });
              } else {                     // KS: a column whose sentence is this row takes relX1 instead of relX
                const float gsub1 = fmaf(rc.relX1 + (masked ? a.neg : 0.f), LOG2E, -mb);
                const int32_t* kcs = cp->cs + 32 * g;
                body([&](int jj) { return kcs[jj] == i ? gsub1 : gsub; });
              }
            };
            group(0, pk0);
            group(1, pk1);
            const float lsum = (ls0 + ls1) + (ls2 + ls3);
            if (!__any_sync(0xffffffffu, !(lsum < 1.2676506e30f))) {   // 2^100; false for inf / NaN too
              tmem_st16(t_s, pk0);
              tmem_st16(t_s + 32, pk1);
              tmem_wait_st();
              tc_fence_before_sync();
              mbar_arrive(&bars->p_full[c & 1]);
              if (tid == 0) TRACE(0, 6 + 3 * c);
              l += lsum;
              continue;
            }
          }
        }
        // ---- pass 1: row maximum; FAST groups leave the raw accumulator in place ----
        float mx = -INFINITY;
        uint32_t dead_mask = 0;
        float mul0 = 0.f, mul1 = 0.f, add0 = 0.f, add1 = 0.f;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const uint32_t w0 = g ? pw.y : pw.x;
          int mode = (int)(w0 & 0xffu);
          if (mode == plan::DEAD) {
            dead_mask |= 1u << g;
            continue;
          }
          const int g0 = key0 + 32 * g;
          const bool mask_pe = (w0 & plan::F_MASK_PE) != 0;
          const bool masked = mre && !mask_pe && (rc.q_e != (g ? pe.y : pe.x));
          // literal mode: a masked score keeps x * scale + rel -> per-element form (EDGE with a full span)
          if (lit && mode == plan::FAST && __any_sync(0xffffffffu, masked)) mode = plan::EDGE;
          const float mterm = masked ? a.neg : 0.f;
          const int ccls = (int)((w0 >> 8) & 0xffu);
          const float relc = rel_const(ccls, rc);
          float gmul, gadd, gmax;
          if (mode == plan::FAST) {
            // every row of the warp masked here and already holding a real maximum: p == 0 exactly
            if (mre && skip_ok && __all_sync(0xffffffffu, masked && m > real_thr)) {
              dead_mask |= 1u << g;
              continue;
            }
            uint32_t v[32];
            tmem_ld32(t_s + 32 * g, v);
            tmem_wait_ld();
            float r4[4];
#pragma unroll
            for (int y = 0; y < 4; ++y)
              r4[y] = fmaxf(fmaxf(__uint_as_float(v[8 * y]), __uint_as_float(v[8 * y + 1])), __uint_as_float(v[8 * y + 2]));
#pragma unroll
            for (int y = 0; y < 4; ++y) {
              r4[y] = fmaxf(fmaxf(r4[y], __uint_as_float(v[8 * y + 3])), __uint_as_float(v[8 * y + 4]));
              r4[y] = fmaxf(fmaxf(r4[y], __uint_as_float(v[8 * y + 5])), __uint_as_float(v[8 * y + 6]));
              r4[y] = fmaxf(r4[y], __uint_as_float(v[8 * y + 7]));
            }
            const float mraw = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
            const float cadd = relc + mterm;
            // a masked score is cadd itself: |x * scale| < 32 is absorbed by -1e9 in fp32
            gmax = masked ? cadd : fmaf(mraw, a.scale, cadd);
            gmul = masked ? 0.f : scale2;
            gadd = cadd;
          } else {
            if (mode == plan::GEN) {
              // real loop, TMEM as dynamically indexed scratch: one copy of the generic code
#pragma unroll 1
              for (int jj = 0; jj < 32; ++jj) {
                const uint32_t raw = tmem_ld1(t_s + 32 * g + jj);
                tmem_wait_ld();
                const float tv = score_generic(__uint_as_float(raw), sc, rc, b, i, row, row_ok, g0 + jj,
                                               cp->ce[32 * g + jj], cp->cs[32 * g + jj], rel_s, a.scale, a.neg);
                __syncwarp();   // score_generic diverges per row; tcgen05.st needs the converged warp
                tmem_st1(t_s + 32 * g + jj, __float_as_uint(tv));
              }
              tmem_wait_st();
            }
            float t[32];
            {
              uint32_t v[32];
              tmem_ld32(t_s + 32 * g, v);
              tmem_wait_ld();
#pragma unroll
              for (int x = 0; x < 32; ++x) t[x] = __uint_as_float(v[x]);
            }
            switch (mode) {
              case plan::EDGE: {
                // live columns of this row form one interval [jlo, jhi)
                const int d0 = g0 - i;
                int jlo = 0, jhi = min(32, sc.ke - g0);
                if (sc.band) {
                  jlo = max(jlo, -sc.radius - d0);
                  jhi = min(jhi, sc.radius - d0 + 1);
                }
                const unsigned span = (unsigned)max(jhi - jlo, 0);
                const float cadd = relc + mterm;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const float v = fmaf(t[jj], a.scale, cadd);
                  t[jj] = ((unsigned)(jj - jlo) < span) ? v : -INFINITY;
                }
                break;
              }
              case plan::DIAG: {
                const int d0 = g0 - i + sc.D;  // slot = clamp(off, -D, D) + D = clamp(d0 + jj, 0, 2D)
                const float* base = rel_s + row;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const int s = min(max(d0 + jj, 0), 2 * sc.D);
                  t[jj] = fmaf(t[jj], a.scale, base[s * TM] + mterm);
                }
                break;
              }
              case plan::QS: {
                const int d0 = rc.q_sent - g0;  // special column index within the group
                const float c0 = rc.relX + mterm, c1 = rc.relX1 + mterm;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) t[jj] = fmaf(t[jj], a.scale, d0 == jj ? c1 : c0);
                break;
              }
              case plan::KS: {
                const float c0 = rc.relX + mterm, c1 = rc.relX1 + mterm;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) t[jj] = fmaf(t[jj], a.scale, cp->cs[32 * g + jj] == i ? c1 : c0);
                break;
              }
              case plan::EXPL: if constexpr (EX) {
                // explicit int32 tensors: this row's 32 consecutive entries (band: column k = j - i + r)
                const Side& sd = *sc.sd;
                const int d0 = g0 - i;
                int jlo = 0, jhi = min(32, sc.ke - g0);
                if (sc.band) {
                  jlo = max(jlo, -sc.radius - d0);
                  jhi = min(jhi, sc.radius - d0 + 1);
                }
                const uint32_t live = plan::span_bits(jlo, jhi);
                const uint32_t take = row_ok ? live : 0u;   // rows beyond the end: no loads (unmasked, no id)
                const int64_t eoff = (int64_t)b * sd.sb + (int64_t)i * sd.sq + (sc.band ? d0 + sc.radius : g0);
                const bool vec = !sc.band && ((sd.sb | sd.sq) & 3) == 0 &&
                                 __all_sync(0xffffffffu, take == 0xffffffffu);
                if (sc.id_rule == IDR_EXPLICIT) {
                  int id[32];
                  plan::load_row32(sd.ids + eoff, vec && (reinterpret_cast<uintptr_t>(sd.ids) & 15) == 0, take, -1, id);
#pragma unroll
                  for (int jj = 0; jj < 32; ++jj) {
                    float rel = 0.f;
                    if ((unsigned)id[jj] < (unsigned)sc.R) rel = rel_s[relmeta[id[jj]].slot_off + row];
                    t[jj] = fmaf(t[jj], a.scale, rel + mterm);   // mterm: example-id mask uniform over the group
                  }
                } else {
#pragma unroll
                  for (int jj = 0; jj < 32; ++jj) t[jj] = fmaf(t[jj], a.scale, mterm);
                }
                if (sc.mask_rule == MR_EXPLICIT) {
                  int ok[32];
                  plan::load_row32(sd.mask + eoff, vec && (reinterpret_cast<uintptr_t>(sd.mask) & 15) == 0, take, 1, ok);
#pragma unroll
                  for (int jj = 0; jj < 32; ++jj) t[jj] += ok[jj] != 0 ? 0.f : a.neg;
                }
                if (live != 0xffffffffu) {
#pragma unroll
                  for (int jj = 0; jj < 32; ++jj) t[jj] = (live >> jj) & 1u ? t[jj] : -INFINITY;
                }
              } break;
              default:
                break;  // GEN: evaluated above
            }
            if (mask_pe) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) t[jj] += (cp->ce[32 * g + jj] == rc.q_e) ? 0.f : a.neg;
            }
            float r4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int x = 0; x < 32; x += 8) {
#pragma unroll
              for (int y = 0; y < 4; ++y) r4[y] = fmaxf(fmaxf(r4[y], t[x + 2 * y]), t[x + 2 * y + 1]);
            }
            gmax = fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3]));
            if (mode != plan::GEN) {
              uint32_t v[32];
#pragma unroll
              for (int x = 0; x < 32; ++x) v[x] = __float_as_uint(t[x]);
              tmem_st32(t_s + 32 * g, v);
            }
            gmul = LOG2E;
            gadd = 0.f;
          }
          mx = fmaxf(mx, gmax);
          if (g == 0) {
            mul0 = gmul;
            add0 = gadd;
          } else {
            mul1 = gmul;
            add1 = gadd;
          }
        }
        tmem_wait_st();
        // ---- running maximum with lazy rescaling of O (in TMEM) ----
        m_true = fmaxf(m_true, mx);
        const bool fresh = (m == M_INIT);   // nothing live so far: O row == 0 and l == 0
        const bool grow = !fresh && (m_true - m) * LOG2E > RESCALE_LOG2;
        if (fresh) m = m_true;
        if (__any_sync(0xffffffffu, grow)) {
          // P.V of the previous chunk must have landed before O is touched.  (o_full alternates
          // between two barriers: P.V of chunk c-3 is known complete once S_c has been seen, so
          // the barrier of chunk c-1 is at most one phase away and the parity wait is exact.)
          mbar_wait_warp(&bars->o_full[(c - 1) & 1], ((c - 1) >> 1) & 1);
          tc_fence_after_sync();
          const float f = grow ? ex2((m - m_true) * LOG2E) : 1.f;
          if (grow) m = m_true;
          l *= f;
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t v[32];
            tmem_ld32(tmem + T_O + lane_sel + hh * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int x = 0; x < 32; ++x) v[x] = __float_as_uint(__uint_as_float(v[x]) * f);
            tmem_st32(tmem + T_O + lane_sel + hh * 32, v);
          }
          tmem_wait_st();
        }
        const float mb = m * LOG2E;
        // ---- pass 2: p = exp2(x * mul + add - mb), row sum, P (bf16) back into TMEM ----
        float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          uint32_t pk[16];
          if (dead_mask & (1u << g)) {
#pragma unroll
            for (int x = 0; x < 16; ++x) pk[x] = 0u;
          } else {
            const float gmul = g ? mul1 : mul0;
            // exponent = x * mul + add * log2e - mb with the product taken exactly (one FMA): every
            // form then yields the same value for a score equal to the running maximum
            const float gsub = fmaf(g ? add1 : add0, LOG2E, -mb);
            uint32_t v[32];
            tmem_ld32(t_s + 32 * g, v);
            tmem_wait_ld();
            const int dcol = sc.col_base + key0 + 32 * g;
#pragma unroll
            for (int x = 0; x < 16; x += 2) {
              float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), gmul, gsub));
              float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), gmul, gsub));
              float p2 = ex2(fmaf(__uint_as_float(v[2 * x + 2]), gmul, gsub));
              float p3 = ex2(fmaf(__uint_as_float(v[2 * x + 3]), gmul, gsub));
              ls0 += p0;
              ls1 += p1;
              ls2 += p2;
              ls3 += p3;
              if constexpr (DROP) {
                p0 = dropout_keep(drow, dcol + 2 * x, dthr) ? p0 : 0.f;
                p1 = dropout_keep(drow, dcol + 2 * x + 1, dthr) ? p1 : 0.f;
                p2 = dropout_keep(drow, dcol + 2 * x + 2, dthr) ? p2 : 0.f;
                p3 = dropout_keep(drow, dcol + 2 * x + 3, dthr) ? p3 : 0.f;
              }
              pk[x] = pack_bf16x2(p0, p1);
              pk[x + 1] = pack_bf16x2(p2, p3);
            }
          }
          // P of keys [32g, 32g+32) -> packed columns [32g, 32g+16): stays inside this group's own
          // score columns, so the not-yet-read scores of group 1 are never clobbered
          tmem_st16(t_s + 32 * g, pk);
        }
        tmem_wait_st();
        tc_fence_before_sync();
        mbar_arrive(&bars->p_full[c & 1]);
        if (tid == 0) TRACE(0, 6 + 3 * c);
        l += (ls0 + ls1) + (ls2 + ls3);
        seeded = __all_sync(0xffffffffu, m != M_INIT);
      }
    }
    // ---- epilogue: O / l ----
    if (nchunks >= 1) {
      mbar_wait_warp(&bars->o_full[(nchunks - 1) & 1], ((nchunks - 1) >> 1) & 1);
      tc_fence_after_sync();
    }
    const float inv = (DROP ? a.drop.inv_keep : 1.f) / l;
    __nv_bfloat16* dst = row_ptr_mut<__nv_bfloat16>(a.out, b, i, h);
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v[32];
      tmem_ld32(tmem + T_O + lane_sel + hh * 32, v);
      tmem_wait_ld();
      if (row_ok) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * x + 0]) * inv, __uint_as_float(v[8 * x + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(v[8 * x + 2]) * inv, __uint_as_float(v[8 * x + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(v[8 * x + 4]) * inv, __uint_as_float(v[8 * x + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(v[8 * x + 6]) * inv, __uint_as_float(v[8 * x + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + hh * 32 + 8 * x) = w;
        }
      }
    }
    if (row_ok) {
      // statistics against the TRUE maximum: (m_true, sum exp(t - m_true)).  The exponentials were
      // taken against mb = fl(m * log2e); r = m * log2e - mb (exact product) is the offset every term
      // carries -- up to +-64 when the maximum is a masked score (~ -1e9) -- and is divided out.
      float2* st = reinterpret_cast<float2*>(a.stats) + ((int64_t)(b * a.H + h) * a.rows.len + i);
      const float r = fmaf(m, LOG2E, -(m * LOG2E));
      *st = make_float2(m_true, l * ex2((m - m_true) * LOG2E - r));
    }
    if (tid == 0) TRACE(0, 3);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

// A [B, len, H, 64] bf16 view the TMA path can address: 16-byte base and strides; a zero (broadcast)
// stride is only representable over an extent of 1 (tensor maps need non-zero strides), so broadcast
// views over more than one element take the SIMT kernels, which handle them.
bool tc_t4_ok(const T4& t, int B, int len, int H) {
  if (!t.ptr || reinterpret_cast<uintptr_t>(t.ptr) % 16) return false;
  if (t.sb % 8 || t.sl % 8 || t.sh % 8) return false;
  if ((t.sb == 0 && B > 1) || (t.sl == 0 && len > 1) || (t.sh == 0 && H > 1)) return false;
  return true;
}

bool tc_fwd_args_supported(const FwdArgs& a, int dtype, int d) {
  if (dtype != MLT_BF16 || d != 64 || a.rows.R > 64) return false;
  if (!tc_t4_ok(a.rows.q, a.B, a.rows.len, a.H) || !tc_t4_ok(a.out, a.B, a.rows.len, a.H)) return false;
  for (int s = 0; s < a.nseg; ++s)
    if (!tc_t4_ok(a.seg[s].k, a.B, a.seg[s].len, a.H) || !tc_t4_ok(a.seg[s].v, a.B, a.seg[s].len, a.H)) return false;
  if (a.rows.R > 0 && reinterpret_cast<uintptr_t>(a.rows.emb) % 16) return false;
  return get_encode_tiled() != nullptr;
}

#ifdef MLT_TC_TRACE
extern "C" __attribute__((visibility("default"))) int mlt_debug_read_trace(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 3 * 256);
}
#endif

int tc_launch_fwd(const FwdArgs& a, cudaStream_t st) {
  // the shared-memory opt-in is a per-device function attribute: once per device, thread-safe
  static PerDeviceOnce once;
  const int ae = once.run([] {
    cudaError_t e = cudaSuccess;
    auto set = [&](auto kernel) {
      if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_ALLOC);
    };
    set(tc_fwd_kernel<false, false>);
    set(tc_fwd_kernel<true, false>);
    set(tc_fwd_kernel<false, true>);
    set(tc_fwd_kernel<true, true>);
    set(tc_fwd_kernel<false, false, true>);
    return (int)e;
  });
  if (ae) return ae;
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
  const bool ex = side_is_explicit(a.seg[0].side) || (a.nseg > 1 && side_is_explicit(a.seg[1].side));
  const bool dr = a.drop.thr != 0;
  auto launch = [&](auto kernel) { kernel<<<grid, NTHREADS, SM_ALLOC, st>>>(mq, mk0, mv0, mk1, mv1, me, p); };
  if (ex && dr) launch(tc_fwd_kernel<true, true>);
  else if (ex) launch(tc_fwd_kernel<true, false>);
  else if (dr) launch(tc_fwd_kernel<false, true>);
  else if (!(fabsf(a.neg) <= 1e5f)) launch(tc_fwd_kernel<false, false, true>);
  else launch(tc_fwd_kernel<false, false>);
  return (int)cudaGetLastError();
}

}  // namespace mlt
