// CUDA-core (fp32 accumulate) kernels for relative / global-local attention.
//
// Role: (1) the fp32 parity path (1e-5 relative against the fp64 oracle needs fp32 FMAs, not
// tensor-core bf16), (2) the generic-shape path (any d in {32,64,128}, any R <= 64, explicit
// int32 side inputs), (3) the on-device cross-check for the tcgen05 kernels.  All kernels are
// deterministic: every output element is owned by exactly one thread pair, no atomics.
//
// Mapping: a pair of adjacent lanes owns one row (query row in the row-centric kernels, key
// row in the key-centric kernel); each lane holds half of the head dimension in registers and
// the two partial dot products are combined with one shuffle.  Key (or query) rows are staged
// through shared memory as fp32 in chunks of KC rows and read back as warp-broadcast LDS.128.
//
// Semantics follow SURVEY.md 8-spec: s = (q.k + allrel[id]) * scale (+ neg if masked), one
// joint online softmax over all segments; stats = (row max, row sum).

#include "mlt_common.cuh"

namespace mlt {
namespace {

constexpr int ROWS = 64;   // rows owned by one block
constexpr int NT = 128;    // threads per block (2 per row)
constexpr int KC = 32;     // staged rows per chunk
constexpr int RMAX = 64;   // largest supported relative vocabulary

template <typename T, int D>
__device__ __forceinline__ void stage_rows(float* dst, const T4& t, int b, int h, int row0,
                                           int nrows, int len) {
  // dst[KC][D] <- rows [row0, row0 + nrows) of t (zero outside [0, len)); 4 elements / access.
  constexpr int V = D / 4;
  for (int idx = threadIdx.x; idx < KC * V; idx += NT) {
    const int r = idx / V, c4 = idx % V;
    const int row = row0 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows && row >= 0 && row < len) v = load4<T>(row_ptr<T>(t, b, row, h) + 4 * c4);
    *reinterpret_cast<float4*>(dst + r * D + 4 * c4) = v;
  }
}

template <int DH>
__device__ __forceinline__ float dot_half(const float (&x)[DH], const float* y) {
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(y + 4 * c);
    acc = fmaf(x[4 * c + 0], v.x, acc);
    acc = fmaf(x[4 * c + 1], v.y, acc);
    acc = fmaf(x[4 * c + 2], v.z, acc);
    acc = fmaf(x[4 * c + 3], v.w, acc);
  }
  return acc;
}

template <int DH>
__device__ __forceinline__ void axpy_half(float (&acc)[DH], float a, const float* y) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(y + 4 * c);
    acc[4 * c + 0] = fmaf(a, v.x, acc[4 * c + 0]);
    acc[4 * c + 1] = fmaf(a, v.y, acc[4 * c + 1]);
    acc[4 * c + 2] = fmaf(a, v.z, acc[4 * c + 2]);
    acc[4 * c + 3] = fmaf(a, v.w, acc[4 * c + 3]);
  }
}

template <typename T, int DH>
__device__ __forceinline__ void load_half(float (&x)[DH], const T* p) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 v = load4<T>(p + 4 * c);
    x[4 * c + 0] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
  }
}

template <typename T, int DH>
__device__ __forceinline__ void store_half(T* p, const float (&x)[DH], float mul) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c)
    store4<T>(p + 4 * c, make_float4(x[4 * c] * mul, x[4 * c + 1] * mul, x[4 * c + 2] * mul,
                                     x[4 * c + 3] * mul));
}

// allrel[p] = q . E[p, h, :] + bias[p, h] for the block's rows, transposed into rel_s[R][ROWS].
template <typename T, int D>
__device__ __forceinline__ void compute_allrel(const RowSet& rows, int h, int H,
                                               const float (&q)[D / 2], float* stage,
                                               float* rel_s) {
  const int half = threadIdx.x & 1, rl = threadIdx.x >> 1;
  const T4 e4{const_cast<void*>(rows.emb), 0, (int64_t)H * D, (int64_t)D};
  for (int r0 = 0; r0 < rows.R; r0 += KC) {
    const int n = min(KC, rows.R - r0);
    __syncthreads();
    stage_rows<T, D>(stage, e4, 0, h, r0, n, rows.R);
    __syncthreads();
    for (int rr = 0; rr < n; ++rr) {
      float acc = dot_half<D / 2>(q, stage + rr * D + half * (D / 2));
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (half == 0)
        rel_s[(r0 + rr) * ROWS + rl] =
            acc + to_f32<T>(reinterpret_cast<const T*>(rows.bias)[(r0 + rr) * H + h]);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void stage_key_scalars(const Side& sd, int b, int c0, int n, int* ke_s,
                                                  int* ksent_s) {
  if (threadIdx.x < KC) {
    const int j = c0 + threadIdx.x;
    int ke = 0, ks = -1;
    if ((int)threadIdx.x < n) {
      if (sd.mask_rule == MR_EXAMPLE_ID) ke = __ldg(sd.k_eid + (int64_t)b * sd.k_len + j);
      if (sd.id_rule == IDR_CROSS_KSENT) ks = __ldg(sd.sent + (int64_t)b * sd.sent_len + j);
    }
    ke_s[threadIdx.x] = ke;
    ksent_s[threadIdx.x] = ks;
  }
}

// ------------------------------------------------------------------------------------------
// Forward: grid (ceil(Lq / ROWS), H, B).
template <typename T, int D>
__global__ void __launch_bounds__(NT) fwd_kernel(const FwdArgs a) {
  constexpr int DH = D / 2;
  const int tid = threadIdx.x, half = tid & 1, rl = tid >> 1;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * ROWS;
  const int i = i0 + rl;
  const bool row_ok = i < a.rows.len;
  const int R = a.rows.R;
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;
  float* vs = ks + KC * D;
  float* rel_s = vs + KC * D;  // [R][ROWS]
  int* ke_s = reinterpret_cast<int*>(rel_s + RMAX * ROWS);
  int* ksent_s = ke_s + KC;

  float q[DH];
  if (row_ok) {
    load_half<T, DH>(q, row_ptr<T>(a.rows.q, b, i, h) + half * DH);
  } else {
#pragma unroll
    for (int c = 0; c < DH; ++c) q[c] = 0.f;
  }
  if (R > 0) compute_allrel<T, D>(a.rows, h, a.H, q, ks, rel_s);

  float m = -INFINITY, l = 0.f;
  float o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.f;

  const int wrow0 = i0 + (tid >> 5) * 16;  // first row of this warp
  const uint32_t dthr = a.drop.thr;
  const uint32_t drow = dropout_row_base(dropout_salt(a.drop, (uint32_t)(b * a.H + h)), i);
  for (int sgi = 0; sgi < a.nseg; ++sgi) {
    const KeySeg& sg = a.seg[sgi];
    const Side& sd = sg.side;
    int jb = 0, je = sg.len;
    if (sg.band) {
      jb = max(0, i0 - sg.radius);
      je = min(sg.len, i0 + ROWS + sg.radius);
    }
    int q_e = 0, q_sent = -1;
    if (row_ok && sd.mask_rule == MR_EXAMPLE_ID) q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
    if (row_ok && sd.id_rule == IDR_CROSS_QSENT) q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
    for (int c0 = jb; c0 < je; c0 += KC) {
      const int n = min(KC, je - c0);
      __syncthreads();
      stage_rows<T, D>(ks, sg.k, b, h, c0, n, sg.len);
      stage_rows<T, D>(vs, sg.v, b, h, c0, n, sg.len);
      stage_key_scalars(sd, b, c0, n, ke_s, ksent_s);
      __syncthreads();
      if (sg.band && (c0 + n - 1 < wrow0 - sg.radius || c0 > wrow0 + 15 + sg.radius)) continue;
      for (int jj = 0; jj < n; ++jj) {
        const int j = c0 + jj;
        float acc = dot_half<DH>(q, ks + jj * D + half * DH);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        const int off = j - i;
        const bool cand = row_ok && (!sg.band || (off <= sg.radius && off >= -sg.radius));
        if (cand) {
          bool ok;
          int id;
          side_eval(sd, b, i, j, sg.band ? off + sg.radius : j, q_e, ke_s[jj], q_sent,
                    ksent_s[jj], ok, id);
          const float rel = (id >= 0 && id < R) ? rel_s[id * ROWS + rl] : 0.f;
          float s = (acc + rel) * a.scale;
          if (!ok) s += a.neg;
          if (s > m) {
            const float alpha = __expf(m - s);
            l *= alpha;
#pragma unroll
            for (int c = 0; c < DH; ++c) o[c] *= alpha;
            m = s;
          }
          const float p = __expf(s - m);
          l += p;   // the softmax normaliser is taken BEFORE dropout
          const bool keep = dthr == 0 || dropout_keep(drow, sg.col_base + j, dthr);
          axpy_half<DH>(o, keep ? p : 0.f, vs + jj * D + half * DH);
        }
      }
    }
  }
  if (row_ok) {
    store_half<T, DH>(row_ptr_mut<T>(a.out, b, i, h) + half * DH, o, (dthr ? a.drop.inv_keep : 1.f) / l);
    if (half == 0) {
      float2* st = reinterpret_cast<float2*>(a.stats) + ((int64_t)(b * a.H + h) * a.rows.len + i);
      *st = make_float2(m, l);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Backward, query-centric: dq, delta, allrel, dallrel.  grid (ceil(Lq / ROWS), H, B).
template <typename T, int D>
__global__ void __launch_bounds__(NT) bwd_q_kernel(const BwdQArgs a) {
  constexpr int DH = D / 2;
  const int tid = threadIdx.x, half = tid & 1, rl = tid >> 1;
  const int b = blockIdx.z, h = blockIdx.y, i0 = blockIdx.x * ROWS;
  const int i = i0 + rl;
  const bool row_ok = i < a.rows.len;
  const int R = a.rows.R;
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;
  float* vs = ks + KC * D;
  float* rel_s = vs + KC * D;          // [R][ROWS]
  float* drel_s = rel_s + RMAX * ROWS;  // [R][ROWS]
  int* ke_s = reinterpret_cast<int*>(drel_s + RMAX * ROWS);
  int* ksent_s = ke_s + KC;

  float q[DH], go[DH], dq[DH];
  float delta = 0.f, m = 0.f, linv = 0.f;
#pragma unroll
  for (int c = 0; c < DH; ++c) { q[c] = 0.f; go[c] = 0.f; dq[c] = 0.f; }
  const int64_t srow = (int64_t)(b * a.H + h) * a.rows.len + i;
  if (row_ok) {
    load_half<T, DH>(q, row_ptr<T>(a.rows.q, b, i, h) + half * DH);
    load_half<T, DH>(go, row_ptr<T>(a.d_out, b, i, h) + half * DH);
    float ov[DH];
    load_half<T, DH>(ov, row_ptr<T>(a.out, b, i, h) + half * DH);
#pragma unroll
    for (int c = 0; c < DH; ++c) delta = fmaf(go[c], ov[c], delta);
    const float2 st = __ldg(reinterpret_cast<const float2*>(a.stats) + srow);
    m = st.x;
    linv = 1.f / st.y;
  }
  delta += __shfl_xor_sync(0xffffffffu, delta, 1);
  if (row_ok && half == 0) a.delta[srow] = delta;
  for (int idx = tid; idx < RMAX * ROWS; idx += NT) drel_s[idx] = 0.f;
  if (R > 0) {
    compute_allrel<T, D>(a.rows, h, a.H, q, ks, rel_s);
    if (row_ok) {
      // publish allrel for the key-centric pass: this lane writes R/2 (rounded) entries
      const int per = (R + 1) / 2;
      for (int p = half * per; p < min(R, (half + 1) * per); ++p)
        a.allrel[srow * R + p] = rel_s[p * ROWS + rl];
    }
  }
  __syncthreads();

  const int wrow0 = i0 + (tid >> 5) * 16;
  const uint32_t dthr = a.drop.thr;
  const uint32_t drow = dropout_row_base(dropout_salt(a.drop, (uint32_t)(b * a.H + h)), i);
  for (int sgi = 0; sgi < a.nseg; ++sgi) {
    const KeySeg& sg = a.seg[sgi];
    const Side& sd = sg.side;
    int jb = 0, je = sg.len;
    if (sg.band) {
      jb = max(0, i0 - sg.radius);
      je = min(sg.len, i0 + ROWS + sg.radius);
    }
    int q_e = 0, q_sent = -1;
    if (row_ok && sd.mask_rule == MR_EXAMPLE_ID) q_e = __ldg(sd.q_eid + (int64_t)b * sd.q_len + i);
    if (row_ok && sd.id_rule == IDR_CROSS_QSENT) q_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + i);
    for (int c0 = jb; c0 < je; c0 += KC) {
      const int n = min(KC, je - c0);
      __syncthreads();
      stage_rows<T, D>(ks, sg.k, b, h, c0, n, sg.len);
      stage_rows<T, D>(vs, sg.v, b, h, c0, n, sg.len);
      stage_key_scalars(sd, b, c0, n, ke_s, ksent_s);
      __syncthreads();
      if (sg.band && (c0 + n - 1 < wrow0 - sg.radius || c0 > wrow0 + 15 + sg.radius)) continue;
      for (int jj = 0; jj < n; ++jj) {
        const int j = c0 + jj;
        float acc = dot_half<DH>(q, ks + jj * D + half * DH);
        float dp = dot_half<DH>(go, vs + jj * D + half * DH);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        dp += __shfl_xor_sync(0xffffffffu, dp, 1);
        const int off = j - i;
        const bool cand = row_ok && (!sg.band || (off <= sg.radius && off >= -sg.radius));
        if (cand) {
          bool ok;
          int id;
          side_eval(sd, b, i, j, sg.band ? off + sg.radius : j, q_e, ke_s[jj], q_sent,
                    ksent_s[jj], ok, id);
          const bool idv = id >= 0 && id < R;
          const float rel = idv ? rel_s[id * ROWS + rl] : 0.f;
          float s = (acc + rel) * a.scale;
          if (!ok) s += a.neg;
          const float p = __expf(s - m) * linv;
          if (dthr) dp = dropout_keep(drow, sg.col_base + j, dthr) ? dp * a.drop.inv_keep : 0.f;
          const float ds = p * (dp - delta);
          axpy_half<DH>(dq, ds, ks + jj * D + half * DH);
          if (idv && half == 0) drel_s[id * ROWS + rl] += ds;
        }
      }
    }
  }
  __syncthreads();
  // dq += sum_p dallrel[p] * E[p, h, :]; publish dallrel.
  if (R > 0) {
    const T4 e4{const_cast<void*>(a.rows.emb), 0, (int64_t)a.H * D, (int64_t)D};
    for (int r0 = 0; r0 < R; r0 += KC) {
      const int n = min(KC, R - r0);
      __syncthreads();
      stage_rows<T, D>(ks, e4, 0, h, r0, n, R);
      __syncthreads();
      for (int rr = 0; rr < n; ++rr)
        axpy_half<DH>(dq, drel_s[(r0 + rr) * ROWS + rl], ks + rr * D + half * DH);
    }
    if (row_ok) {
      const int per = (R + 1) / 2;
      for (int p = half * per; p < min(R, (half + 1) * per); ++p)
        a.dallrel[srow * R + p] = drel_s[p * ROWS + rl];
    }
  }
  if (row_ok) store_half<T, DH>(row_ptr_mut<T>(a.d_q, b, i, h) + half * DH, dq, a.scale);
}

// ------------------------------------------------------------------------------------------
// Backward, key-centric: dk, dv of one key set from up to two query sources.
// grid (ceil(Lk / ROWS), H, B).
template <typename T, int D>
__global__ void __launch_bounds__(NT) bwd_kv_kernel(const BwdKVArgs a) {
  constexpr int DH = D / 2;
  const int tid = threadIdx.x, half = tid & 1, rl = tid >> 1;
  const int b = blockIdx.z, h = blockIdx.y, j0 = blockIdx.x * ROWS;
  const int j = j0 + rl;
  const bool key_ok = j < a.len;
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                    // [KC][D]
  float* gos = qs + KC * D;          // [KC][D]
  float* rel_q = gos + KC * D;       // [KC][RMAX]
  float* m_s = rel_q + KC * RMAX;    // [KC]
  float* linv_s = m_s + KC;
  float* delta_s = linv_s + KC;
  int* qe_s = reinterpret_cast<int*>(delta_s + KC);
  int* qsent_s = qe_s + KC;

  float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) { k[c] = 0.f; v[c] = 0.f; dk[c] = 0.f; dv[c] = 0.f; }
  if (key_ok) {
    load_half<T, DH>(k, row_ptr<T>(a.k, b, j, h) + half * DH);
    load_half<T, DH>(v, row_ptr<T>(a.v, b, j, h) + half * DH);
  }
  const int wrow0 = j0 + (tid >> 5) * 16;
  for (int si = 0; si < a.nsrc; ++si) {
    const QuerySource& qs_src = a.src[si];
    const Side& sd = qs_src.side;
    const int R = qs_src.rows.R;
    const int lq = qs_src.rows.len;
    int ib = 0, ie = lq;
    if (qs_src.band) {
      ib = max(0, j0 - qs_src.radius);
      ie = min(lq, j0 + ROWS + qs_src.radius);
    }
    const uint32_t dthr = qs_src.drop.thr;
    const uint32_t dsalt = dropout_salt(qs_src.drop, (uint32_t)(b * a.H + h));
    int k_e = 0, k_sent = -1;
    if (key_ok && sd.mask_rule == MR_EXAMPLE_ID) k_e = __ldg(sd.k_eid + (int64_t)b * sd.k_len + j);
    if (key_ok && sd.id_rule == IDR_CROSS_KSENT) k_sent = __ldg(sd.sent + (int64_t)b * sd.sent_len + j);
    for (int c0 = ib; c0 < ie; c0 += KC) {
      const int n = min(KC, ie - c0);
      __syncthreads();
      stage_rows<T, D>(qs, qs_src.rows.q, b, h, c0, n, lq);
      stage_rows<T, D>(gos, qs_src.d_out, b, h, c0, n, lq);
      const int64_t srow0 = (int64_t)(b * a.H + h) * lq + c0;
      if (tid < KC) {
        float mm = 0.f, li = 0.f, de = 0.f;
        int qe = 0, qsn = -1;
        if (tid < n) {
          const float2 st = __ldg(reinterpret_cast<const float2*>(qs_src.stats) + srow0 + tid);
          mm = st.x;
          li = 1.f / st.y;
          de = __ldg(qs_src.delta + srow0 + tid);
          if (sd.mask_rule == MR_EXAMPLE_ID) qe = __ldg(sd.q_eid + (int64_t)b * sd.q_len + c0 + tid);
          if (sd.id_rule == IDR_CROSS_QSENT) qsn = __ldg(sd.sent + (int64_t)b * sd.sent_len + c0 + tid);
        }
        m_s[tid] = mm; linv_s[tid] = li; delta_s[tid] = de; qe_s[tid] = qe; qsent_s[tid] = qsn;
      }
      for (int idx = tid; idx < n * R; idx += NT)
        rel_q[(idx / R) * RMAX + idx % R] = __ldg(qs_src.allrel + srow0 * R + idx);
      __syncthreads();
      if (qs_src.band && (c0 + n - 1 < wrow0 - qs_src.radius || c0 > wrow0 + 15 + qs_src.radius))
        continue;
      for (int ii = 0; ii < n; ++ii) {
        const int i = c0 + ii;
        float acc = dot_half<DH>(k, qs + ii * D + half * DH);
        float dp = dot_half<DH>(v, gos + ii * D + half * DH);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        dp += __shfl_xor_sync(0xffffffffu, dp, 1);
        const int off = j - i;
        const bool cand = key_ok && (!qs_src.band || (off <= qs_src.radius && off >= -qs_src.radius));
        if (cand) {
          bool ok;
          int id;
          side_eval(sd, b, i, j, qs_src.band ? off + qs_src.radius : j, qe_s[ii], k_e, qsent_s[ii],
                    k_sent, ok, id);
          const float rel = (id >= 0 && id < R) ? rel_q[ii * RMAX + id] : 0.f;
          float s = (acc + rel) * a.scale;
          if (!ok) s += a.neg;
          const float p = __expf(s - m_s[ii]) * linv_s[ii];
          float pd = p;   // dropped-out probability (what multiplied V in the forward)
          if (dthr) {
            const bool keep = dropout_keep(dropout_row_base(dsalt, i), qs_src.col_base + j, dthr);
            pd = keep ? p * qs_src.drop.inv_keep : 0.f;
            dp = keep ? dp * qs_src.drop.inv_keep : 0.f;
          }
          const float ds = p * (dp - delta_s[ii]);
          axpy_half<DH>(dv, pd, gos + ii * D + half * DH);
          axpy_half<DH>(dk, ds, qs + ii * D + half * DH);
        }
      }
    }
  }
  if (key_ok) {
    store_half<T, DH>(row_ptr_mut<T>(a.d_k, b, j, h) + half * DH, dk, a.scale);
    store_half<T, DH>(row_ptr_mut<T>(a.d_v, b, j, h) + half * DH, dv, 1.f);
  }
}

// ------------------------------------------------------------------------------------------
// Table gradients: d_emb[p,h,:] = scale * sum_{b,i} dallrel[b,h,i,p] q[b,i,h,:],
//                  d_bias[p,h]  = scale * sum_{b,i} dallrel[b,h,i,p].
// Stage 1: grid (nchunk, H, B), one partial per 128-row chunk.  Stage 2: fixed-order sum.
constexpr int TG_ROWS = 128;

template <typename T, int D>
__global__ void __launch_bounds__(NT) table_grad_partial_kernel(const TableGradArgs a) {
  constexpr int GROUPS = NT / D > 0 ? NT / D : 1;   // rho groups handled in parallel
  constexpr int CPT = D > NT ? D / NT : 1;          // columns per thread (D = 128 -> 1)
  constexpr int NACC = RMAX / GROUPS;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int R = a.R;
  __shared__ __align__(16) float qs[KC * D];
  __shared__ __align__(16) float ds[KC * RMAX];
  const int c = tid % D;
  const int g = tid / D;
  float acc[NACC][CPT];
#pragma unroll
  for (int x = 0; x < NACC; ++x)
#pragma unroll
    for (int y = 0; y < CPT; ++y) acc[x][y] = 0.f;
  float bacc = 0.f;
  const int row_begin = chunk * TG_ROWS;
  const int row_end = min(a.len, row_begin + TG_ROWS);
  for (int r0 = row_begin; r0 < row_end; r0 += KC) {
    const int n = min(KC, row_end - r0);
    __syncthreads();
    stage_rows<T, D>(qs, a.q, b, h, r0, n, a.len);
    const int64_t srow0 = (int64_t)(b * a.H + h) * a.len + r0;
    for (int idx = tid; idx < KC * RMAX; idx += NT) {
      const int rr = idx / RMAX, p = idx % RMAX;
      ds[idx] = (rr < n && p < R) ? __ldg(a.dallrel + (srow0 + rr) * R + p) : 0.f;
    }
    __syncthreads();
    if (g < GROUPS) {
      // this thread owns ids [g * NACC, (g + 1) * NACC): contiguous, read as warp-broadcast float4
      for (int rr = 0; rr < n; ++rr) {
        const float qv = qs[rr * D + c];
        const float4* w4 = reinterpret_cast<const float4*>(ds + rr * RMAX + g * NACC);
#pragma unroll
        for (int x4 = 0; x4 < NACC / 4; ++x4) {
          const float4 w = w4[x4];
          acc[4 * x4 + 0][0] = fmaf(w.x, qv, acc[4 * x4 + 0][0]);
          acc[4 * x4 + 1][0] = fmaf(w.y, qv, acc[4 * x4 + 1][0]);
          acc[4 * x4 + 2][0] = fmaf(w.z, qv, acc[4 * x4 + 2][0]);
          acc[4 * x4 + 3][0] = fmaf(w.w, qv, acc[4 * x4 + 3][0]);
        }
      }
    }
    if (tid < RMAX)
      for (int rr = 0; rr < n; ++rr) bacc += ds[rr * RMAX + tid];
  }
  const int64_t pidx = ((int64_t)(b * a.nchunk + chunk) * a.H + h);
  if (g < GROUPS) {
#pragma unroll
    for (int x = 0; x < NACC; ++x) {
      const int p = g * NACC + x;
      if (p < R) {
#pragma unroll
        for (int y = 0; y < CPT; ++y) a.partial[(pidx * R + p) * D + c + y * NT] = acc[x][y];
      }
    }
  }
  if (tid < R) a.partial_bias[pidx * R + tid] = bacc;
}

// Stage 2 of the table gradients: d_emb[p, h, :] = scale * sum_k partial[k, h, p, :] (k over batch x row chunks),
// d_bias likewise.  One block per (p, h): 64 lanes across d (one coalesced 256-byte row per k), 8 slices of k
// with independent loads in flight, combined through shared memory in a fixed order (deterministic).
// HBM-read-bound: the partials are read exactly once.
constexpr int TGR_X = 64, TGR_Y = 8;
__global__ void __launch_bounds__(TGR_X * TGR_Y) table_grad_reduce_kernel(const TableGradArgs a) {
  __shared__ float red[TGR_Y][TGR_X];
  __shared__ float redb[TGR_Y];
  const int D = a.d;
  const int h = blockIdx.x % a.H, p = blockIdx.x / a.H;   // output order [R][H]
  const int x = threadIdx.x, y = threadIdx.y;
  const int np = a.B * a.nchunk;
  const int64_t kstride = (int64_t)a.H * a.R * D;
  for (int c0 = 0; c0 < D; c0 += TGR_X) {
    const int c = c0 + x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < D) {
      const float* src = a.partial + ((int64_t)h * a.R + p) * D + c;
      int k = y;
      for (; k + 3 * TGR_Y < np; k += 4 * TGR_Y) {
        s0 += __ldg(src + (int64_t)k * kstride);
        s1 += __ldg(src + (int64_t)(k + TGR_Y) * kstride);
        s2 += __ldg(src + (int64_t)(k + 2 * TGR_Y) * kstride);
        s3 += __ldg(src + (int64_t)(k + 3 * TGR_Y) * kstride);
      }
      for (; k < np; k += TGR_Y) s0 += __ldg(src + (int64_t)k * kstride);
    }
    red[y][x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (y == 0 && c < D) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < TGR_Y; ++j) s += red[j][x];
      a.d_emb[((int64_t)p * a.H + h) * D + c] = s * a.scale;
    }
    __syncthreads();
  }
  // bias: lane x of slice y takes k = y * 64 + x, y * 64 + x + 512, ...
  {
    float s = 0.f;
    for (int k = y * TGR_X + x; k < np; k += TGR_X * TGR_Y) s += __ldg(a.partial_bias + ((int64_t)k * a.H + h) * a.R + p);
    red[y][x] = s;
    __syncthreads();
    if (x == 0) {
      float t = 0.f;
      for (int j = 0; j < TGR_X; ++j) t += red[y][j];
      redb[y] = t;
    }
    __syncthreads();
    if (x == 0 && y == 0) {
      float t = 0.f;
      for (int j = 0; j < TGR_Y; ++j) t += redb[j];
      a.d_bias[p * a.H + h] = t * a.scale;
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int D>
constexpr size_t fwd_smem() { return (2 * KC * D + RMAX * ROWS) * sizeof(float) + 2 * KC * sizeof(int); }
template <int D>
constexpr size_t bwd_q_smem() { return (2 * KC * D + 2 * RMAX * ROWS) * sizeof(float) + 2 * KC * sizeof(int); }
template <int D>
constexpr size_t bwd_kv_smem() { return (2 * KC * D + KC * RMAX + 3 * KC) * sizeof(float) + 2 * KC * sizeof(int); }

template <typename K>
cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename T, int D>
cudaError_t launch_fwd_t(const FwdArgs& a, cudaStream_t st) {
  cudaError_t e = set_smem(fwd_kernel<T, D>, fwd_smem<D>());
  if (e != cudaSuccess) return e;
  dim3 grid((a.rows.len + ROWS - 1) / ROWS, a.H, a.B);
  fwd_kernel<T, D><<<grid, NT, fwd_smem<D>(), st>>>(a);
  return cudaGetLastError();
}
template <typename T, int D>
cudaError_t launch_bwd_q_t(const BwdQArgs& a, cudaStream_t st) {
  cudaError_t e = set_smem(bwd_q_kernel<T, D>, bwd_q_smem<D>());
  if (e != cudaSuccess) return e;
  dim3 grid((a.rows.len + ROWS - 1) / ROWS, a.H, a.B);
  bwd_q_kernel<T, D><<<grid, NT, bwd_q_smem<D>(), st>>>(a);
  return cudaGetLastError();
}
template <typename T, int D>
cudaError_t launch_bwd_kv_t(const BwdKVArgs& a, cudaStream_t st) {
  cudaError_t e = set_smem(bwd_kv_kernel<T, D>, bwd_kv_smem<D>());
  if (e != cudaSuccess) return e;
  dim3 grid((a.len + ROWS - 1) / ROWS, a.H, a.B);
  bwd_kv_kernel<T, D><<<grid, NT, bwd_kv_smem<D>(), st>>>(a);
  return cudaGetLastError();
}
template <typename T, int D>
cudaError_t launch_table_grad_t(const TableGradArgs& a, cudaStream_t st) {
  dim3 grid(a.nchunk, a.H, a.B);
  table_grad_partial_kernel<T, D><<<grid, NT, 0, st>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  table_grad_reduce_kernel<<<a.R * a.H, dim3(TGR_X, TGR_Y), 0, st>>>(a);
  return cudaGetLastError();
}

#define MLT_DISPATCH(FN, ARGS)                                                   \
  do {                                                                           \
    if (dtype == 0) {                                                            \
      if (d == 32) return FN<float, 32>(ARGS, st);                               \
      if (d == 64) return FN<float, 64>(ARGS, st);                               \
      if (d == 128) return FN<float, 128>(ARGS, st);                             \
    } else {                                                                     \
      if (d == 32) return FN<__nv_bfloat16, 32>(ARGS, st);                       \
      if (d == 64) return FN<__nv_bfloat16, 64>(ARGS, st);                       \
      if (d == 128) return FN<__nv_bfloat16, 128>(ARGS, st);                     \
    }                                                                            \
    return cudaErrorInvalidValue;                                                \
  } while (0)

}  // namespace

bool simt_supports_head_dim(int d) { return d == 32 || d == 64 || d == 128; }
int simt_table_grad_chunks(int len) { return (len + TG_ROWS - 1) / TG_ROWS; }

cudaError_t simt_launch_fwd(const FwdArgs& a, int dtype, int d, cudaStream_t st) {
  MLT_DISPATCH(launch_fwd_t, a);
}
cudaError_t simt_launch_bwd_q(const BwdQArgs& a, int dtype, int d, cudaStream_t st) {
  MLT_DISPATCH(launch_bwd_q_t, a);
}
cudaError_t simt_launch_bwd_kv(const BwdKVArgs& a, int dtype, int d, cudaStream_t st) {
  MLT_DISPATCH(launch_bwd_kv_t, a);
}
cudaError_t simt_launch_table_grad_reduce(const TableGradArgs& a, cudaStream_t st) {
  table_grad_reduce_kernel<<<a.R * a.H, dim3(TGR_X, TGR_Y), 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t simt_launch_table_grad(const TableGradArgs& a, int dtype, cudaStream_t st) {
  const int d = a.d;
  MLT_DISPATCH(launch_table_grad_t, a);
}

}  // namespace mlt
