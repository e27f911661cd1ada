// Bring-up probe for the tcgen05 building blocks (run on a B200 through gpurun).
// Validates, against a CPU reference, each MMA form the attention kernels depend on:
//   T1  S = Q . K^T      A, B K-major SWIZZLE_128B tiles loaded by 4-D TMA from [B,len,H,64]
//   T2  O = P . V        A = P (bf16) in TMEM written by tcgen05.st, B = V MN-major SW128
//   T3  O = P . V        A = P in shared memory (K-major SW128, written by threads)
//   T4  dV-style         A = P^T taken from smem MN-major (A MN-major), B = dO MN-major
// Prints PASS/FAIL per test with the max abs error.
#define MLT_TC_DEBUG_TIMEOUT 1
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../../multimodal-long-transformer-2021_b200/csrc/tc_ptx.cuh"

using namespace mlt;
using namespace mlt::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

struct ProbeArgs {
  float* out;          // [128, 64]
  const float* p_host; // [128, 64] P values (fp32) for T2/T3
  int row0_q, row0_k, h, b;
  int mode;            // test id
  uint32_t v_lbo, v_sbo, v_kstep;  // descriptor params for MN-major operand
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_q,
                                                    const __grid_constant__ CUtensorMap map_k,
                                                    const __grid_constant__ CUtensorMap map_v,
                                                    ProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* q_s = smem;               // 128 x 128 B = 16 KB
  uint8_t* k_s = smem + 16384;       // 64 x 128 B  = 8 KB
  uint8_t* v_s = smem + 24576;       // 64 x 128 B  = 8 KB
  uint8_t* p_s = smem + 32768;       // 128 x 128 B = 16 KB (P in smem, K-major SW128)
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t t_acc = tmem;          // 64 columns fp32 accumulator
  const uint32_t t_p = tmem + 64;       // 32 columns: P as packed bf16x2
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_load, (a.mode == 4 ? 0 : 16384) + 8192 + 8192);
    if (a.mode != 4) tma_load_4d(q_s, &map_q, &bar_load, 0, a.row0_q, a.h, a.b);
    tma_load_4d(k_s, &map_k, &bar_load, 0, a.row0_k, a.h, a.b);
    tma_load_4d(v_s, &map_v, &bar_load, 0, a.row0_k, a.h, a.b);
  }
  if (a.mode == 5) {
    uint32_t z[32];
    for (int c = 0; c < 32; ++c) z[c] = 0x7fc00000u;  // NaN marker: untouched lanes stay NaN
    tmem_st32(t_acc + lane_sel, z);
    tmem_st32(t_acc + lane_sel + 32, z);
    tmem_wait_st();
  }
  // P (row = tid, 64 key columns)
  float prow[64];
  for (int c = 0; c < 64; ++c) prow[c] = a.p_host[tid * 64 + c];
  if (a.mode == 2) {
    uint32_t pk[32];
    for (int c = 0; c < 32; ++c) pk[c] = pack_bf16x2(prow[2 * c], prow[2 * c + 1]);
    tmem_st32(t_p + lane_sel, pk);
    tmem_wait_st();
  }
  if (a.mode == 3) {
    // K-major SW128: row r at r*128 B, 16-byte chunk c stored at chunk (c ^ (r & 7)).
    for (int c = 0; c < 8; ++c) {
      uint4 v;
      v.x = pack_bf16x2(prow[8 * c + 0], prow[8 * c + 1]);
      v.y = pack_bf16x2(prow[8 * c + 2], prow[8 * c + 3]);
      v.z = pack_bf16x2(prow[8 * c + 4], prow[8 * c + 5]);
      v.w = pack_bf16x2(prow[8 * c + 6], prow[8 * c + 7]);
      *reinterpret_cast<uint4*>(p_s + tid * 128 + ((c ^ (tid & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();
  }
  if (a.mode == 4) {
    // A^T stored [k = 64 rows][m = 128] as two SW128 blocks of [64 rows x 128 B] (m halves):
    // A[m, k] = P[m][k]  ->  element (k, m) at block(m / 64) + k * 128 + swizzled 16-B chunk.
    const int m = tid, hb = m >> 6, mm = m & 63;
    for (int kk = 0; kk < 64; ++kk) {
      uint8_t* dst = q_s + hb * 8192 + kk * 128 + (((mm >> 3) ^ (kk & 7)) << 4) + (mm & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16(prow[kk]);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  mbar_wait(&bar_load, 0);
  if (tid == 0) {
    tc_fence_after_sync();
    if (a.mode == 1) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc_sw128(smem_u32(q_s) + kk * 32, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(k_s) + kk * 32, 16, 1024);
        umma_ss(t_acc, da, db, idesc, kk > 0);
      }
    } else if (a.mode == 2) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
      for (int kk = 0; kk < 4; ++kk) {   // 64 keys = 4 x K16
        uint64_t db = make_smem_desc_sw128(smem_u32(v_s) + kk * a.v_kstep, a.v_lbo, a.v_sbo);
        umma_ts(t_acc, t_p + kk * 8, db, idesc, kk > 0);
      }
    } else if (a.mode == 3) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc_sw128(smem_u32(p_s) + kk * 32, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(v_s) + kk * a.v_kstep, a.v_lbo, a.v_sbo);
        umma_ss(t_acc, da, db, idesc, kk > 0);
      }
    } else if (a.mode == 5) {
      // M = 64 accumulator layout probe: D[64 x 64] = Q[0:64] . K^T
      const uint32_t idesc = make_idesc_bf16(64, 64, 0, 0);
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc_sw128(smem_u32(q_s) + kk * 32, 16, 1024);
        uint64_t db = make_smem_desc_sw128(smem_u32(k_s) + kk * 32, 16, 1024);
        umma_ss(t_acc, da, db, idesc, kk > 0);
      }
    } else if (a.mode == 4) {
      // D[128 x 64] = A . V with A MN-major (M contiguous, 2 atoms along M at stride LBO) from smem.
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t da = make_smem_desc_sw128(smem_u32(q_s) + kk * 2048, a.v_lbo, a.v_sbo);
        uint64_t db = make_smem_desc_sw128(smem_u32(v_s) + kk * 2048, 16, 1024);
        umma_ss(t_acc, da, db, idesc, kk > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after_sync();
  uint32_t r[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(t_acc + lane_sel + half * 32, r);
    tmem_wait_ld();
    for (int c = 0; c < 32; ++c) a.out[tid * 64 + half * 32 + c] = __uint_as_float(r[c]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  const int B = 2, LEN = 300, H = 3, D = 64;
  std::vector<__nv_bfloat16> q(B * LEN * H * D), k(q.size()), v(q.size());
  srand(1);
  auto rnd = [] { return (rand() % 2001 - 1000) / 1000.0f; };
  for (size_t i = 0; i < q.size(); ++i) { q[i] = __float2bfloat16(rnd()); k[i] = __float2bfloat16(rnd()); v[i] = __float2bfloat16(rnd()); }
  std::vector<float> p(128 * 64);
  for (auto& x : p) x = bf(fabsf(rnd()));
  __nv_bfloat16 *dq, *dk, *dv; float *dp, *dout;
  CK(cudaMalloc(&dq, q.size() * 2)); CK(cudaMalloc(&dk, q.size() * 2)); CK(cudaMalloc(&dv, q.size() * 2));
  CK(cudaMalloc(&dp, p.size() * 4)); CK(cudaMalloc(&dout, 128 * 64 * 4));
  CK(cudaMemcpy(dq, q.data(), q.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, k.data(), q.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, v.data(), q.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp, p.data(), p.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap mq, mk, mv;
  const int64_t sb = (int64_t)LEN * H * D, sl = H * D, sh = D;
  int e1 = make_qkv_tensor_map(&mq, dq, sb, sl, sh, B, LEN, H, 128);
  int e2 = make_qkv_tensor_map(&mk, dk, sb, sl, sh, B, LEN, H, 64);
  int e3 = make_qkv_tensor_map(&mv, dv, sb, sl, sh, B, LEN, H, 64);
  printf("tensor map encode: %d %d %d\n", e1, e2, e3);
  if (e1 || e2 || e3) return 3;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
  const int b = 1, h = 2;
  auto at = [&](const std::vector<__nv_bfloat16>& t, int row, int c) -> float {
    if (row < 0 || row >= LEN) return 0.f;
    return __bfloat162float(t[((size_t)(b * LEN + row) * H + h) * D + c]);
  };
  std::vector<float> out(128 * 64), ref(128 * 64);
  auto run = [&](ProbeArgs a, const char* name) -> double {
    a.out = dout; a.p_host = dp; a.h = h; a.b = b;
    CK(cudaMemset(dout, 0xff, 128 * 64 * 4));
    probe_kernel<<<1, 128, 49152 + 1024>>>(mq, mk, mv, a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: kernel failed: %s\n", name, cudaGetErrorString(e)); exit(4); }
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double err = 0;
    for (size_t i = 0; i < out.size(); ++i) { double d = fabs((double)out[i] - ref[i]); if (!(d <= err)) err = (d != d) ? 1e30 : d; }
    return err;
  };
  // T1: includes out-of-bounds rows at both ends (row0_q = 200 -> rows 300..327 OOB; row0_k = -20)
  for (int variant = 0; variant < 2; ++variant) {
    ProbeArgs a{}; a.mode = 1; a.row0_q = variant ? 200 : 64; a.row0_k = variant ? -20 : 130;
    for (int i = 0; i < 128; ++i) for (int j = 0; j < 64; ++j) {
      double s = 0; for (int c = 0; c < D; ++c) s += (double)at(q, a.row0_q + i, c) * at(k, a.row0_k + j, c);
      ref[i * 64 + j] = (float)s;
    }
    double err = run(a, "T1");
    printf("T1 S=Q.K^T (variant %d, OOB rows %s): max abs err %.3e -> %s\n", variant, variant ? "yes" : "no", err, err < 1e-3 ? "PASS" : "FAIL");
  }
  // T2/T3: O = P . V for several MN-major descriptor encodings
  struct Enc { uint32_t lbo, sbo, kstep; } encs[] = {{16, 1024, 2048}, {1024, 1024, 2048}, {128, 1024, 2048}, {2048, 1024, 2048}, {1024, 2048, 2048}, {16, 128, 2048}};
  for (int mode = 2; mode <= 3; ++mode)
    for (auto& en : encs) {
      ProbeArgs a{}; a.mode = mode; a.row0_q = 0; a.row0_k = 37; a.v_lbo = en.lbo; a.v_sbo = en.sbo; a.v_kstep = en.kstep;
      for (int i = 0; i < 128; ++i) for (int c = 0; c < D; ++c) {
        double s = 0; for (int j = 0; j < 64; ++j) s += (double)p[i * 64 + j] * at(v, a.row0_k + j, c);
        ref[i * 64 + c] = (float)s;
      }
      double err = run(a, mode == 2 ? "T2" : "T3");
      printf("T%d O=P.V (P in %s; V MN-major lbo=%u sbo=%u kstep=%u): max abs err %.3e -> %s\n", mode, mode == 2 ? "TMEM" : "smem", en.lbo, en.sbo, en.kstep, err, err < 1e-3 ? "PASS" : "FAIL");
    }
  // T4: O = P . V with A = P^T-stored-in-smem taken MN-major (the dS^T -> dQ form of the backward)
  struct EncA { uint32_t lbo, sbo; } encas[] = {{8192, 1024}, {1024, 8192}, {16, 1024}, {8192, 128}};
  for (auto& en : encas) {
    ProbeArgs a{}; a.mode = 4; a.row0_q = 0; a.row0_k = 37; a.v_lbo = en.lbo; a.v_sbo = en.sbo;
    for (int i = 0; i < 128; ++i) for (int c = 0; c < D; ++c) {
      double s = 0; for (int j = 0; j < 64; ++j) s += (double)p[i * 64 + j] * at(v, a.row0_k + j, c);
      ref[i * 64 + c] = (float)s;
    }
    double err = run(a, "T4");
    printf("T4 O=P.V (A MN-major smem, 2 M-atoms: lbo=%u sbo=%u): max abs err %.3e -> %s\n", en.lbo, en.sbo, err, err < 1e-3 ? "PASS" : "FAIL");
  }
  {
    ProbeArgs a{}; a.mode = 5; a.row0_q = 64; a.row0_k = 130;
    std::vector<float> r64(64 * 64);
    for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) {
      double s2 = 0; for (int c = 0; c < D; ++c) s2 += (double)at(q, a.row0_q + i, c) * at(k, a.row0_k + j, c);
      r64[i * 64 + j] = (float)s2;
    }
    std::fill(ref.begin(), ref.end(), 0.f);
    run(a, "T5");
    // which TMEM lane holds row i?
    int lane_of[64]; bool ok5 = true;
    for (int i = 0; i < 64; ++i) {
      lane_of[i] = -1;
      for (int l = 0; l < 128; ++l) {
        double e = 0; for (int j = 0; j < 64; ++j) { double d = fabs((double)out[l * 64 + j] - r64[i * 64 + j]); if (!(d <= e)) e = (d != d) ? 1e30 : d; }
        if (e < 1e-3) { lane_of[i] = l; break; }
      }
      if (lane_of[i] < 0) ok5 = false;
    }
    printf("T5 M=64 accumulator layout: row->lane:");
    for (int i = 0; i < 64; i += 8) printf(" %d->%d", i, lane_of[i]);
    printf(" ... 15->%d 16->%d 31->%d 32->%d 63->%d -> %s\n", lane_of[15], lane_of[16], lane_of[31], lane_of[32], lane_of[63], ok5 ? "PASS" : "FAIL");
  }
  printf("probe done\n");
  return 0;
}
