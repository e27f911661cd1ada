// Micro-benchmarks for the per-SM limits the softmax design depends on (run on a B200 through
// gpurun): tcgen05.ld / tcgen05.st throughput, ex2 (MUFU) throughput, FFMA throughput, and a
// mixed "softmax-like" loop.  One CTA per SM on `nblk` SMs; cycles via clock64 around the loop.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "../../multimodal-long-transformer-2021_b200/csrc/tc_ptx.cuh"

using namespace mlt;
using namespace mlt::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}

// mode: 0 ld32+wait each   1 2x ld32 then wait   2 ld64+wait   3 st32+wait   4 ex2 only
//       5 ffma only        6 softmax-like: ld32, fma, max, ex2, add, pack, st16
//       7 as 6 without any TMEM traffic (registers only)   8 ld16 + wait
__global__ void __launch_bounds__(512) bw_kernel(int mode, int iters, unsigned long long* cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t colbase = ((warp >> 2) * 128) & 511;   // each warpgroup its own 128 columns
  float acc = threadIdx.x * 1e-9f, acc2 = 0.f;
  uint32_t xacc = 0;
  // initialise TMEM so that loads return finite values
  {
    uint32_t z[32];
#pragma unroll
    for (int x = 0; x < 32; ++x) z[x] = __float_as_uint(0.001f * x);
    for (int c = 0; c < 128; c += 32) tmem_st32(tmem + lane_sel + colbase + c, z);
    tmem_wait_st();
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t ta = tmem + lane_sel + colbase + (it & 1) * 64;
    if (mode == 0) {
      uint32_t v[32];
      tmem_ld32(ta, v);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 32; ++x) xacc ^= v[x];
    } else if (mode == 1) {
      uint32_t v[32], w[32];
      tmem_ld32(ta, v);
      tmem_ld32(ta + 32, w);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 32; ++x) xacc ^= v[x] ^ w[x];
    } else if (mode == 2) {
      uint32_t v[64];
      tmem_ld64(ta, v);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 64; ++x) xacc ^= v[x];
    } else if (mode == 3) {
      uint32_t v[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) v[x] = xacc + x + it;
      tmem_st32(ta, v);
      tmem_wait_st();
    } else if (mode == 4) {
#pragma unroll
      for (int x = 0; x < 32; ++x) acc += ex2(acc2 + (float)x * 1e-3f - acc * 1e-3f);
    } else if (mode == 5) {
      float a[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) a[x] = fmaf(acc, 1.0001f + x, acc2);
#pragma unroll
      for (int x = 0; x < 32; ++x) acc = fmaf(a[x], 1e-9f, acc);
    } else if (mode == 6 || mode == 7) {
      uint32_t v[32];
      if (mode == 6) {
        tmem_ld32(ta, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int x = 0; x < 32; ++x) v[x] = __float_as_uint(acc * (x + 1));
      }
      float mx = -1e30f;
#pragma unroll
      for (int x = 0; x < 32; ++x) mx = fmaxf(mx, __uint_as_float(v[x]));
      const float mb = mx * 0.18f;
      uint32_t pk[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) {
        const float p0 = ex2(fmaf(__uint_as_float(v[2 * x]), 0.18f, -mb));
        const float p1 = ex2(fmaf(__uint_as_float(v[2 * x + 1]), 0.18f, -mb));
        acc2 += p0 + p1;
        pk[x] = pack_bf16x2(p0, p1);
      }
      if (mode == 6) {
        tmem_st16(ta, pk);
        tmem_wait_st();
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) xacc ^= pk[x];
      }
      acc = acc2 * 1e-20f;
    } else if (mode == 8) {
      uint32_t v[16];
      tmem_ld16(ta, v);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 16; ++x) xacc ^= v[x];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
  if (xacc == 0x12345678u || acc == 1.2345f) sink[threadIdx.x] = acc + acc2 + xacc;
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
  unsigned long long* cyc;
  float* sink;
  CK(cudaMalloc(&cyc, 148 * sizeof(unsigned long long)));
  CK(cudaMalloc(&sink, 1024 * sizeof(float)));
  const char* names[] = {"ld32+wait", "2x ld32, wait", "ld64+wait", "st32+wait", "ex2 x32", "ffma x64",
                         "softmax-like (tmem)", "softmax-like (regs)", "ld16+wait"};
  const int elems[] = {32, 64, 64, 32, 32, 64, 32, 32, 16};
  const int iters = 2000;
  for (int mode = 0; mode <= 8; ++mode) {
    for (int nw : {4, 8, 16}) {
      bw_kernel<<<1, nw * 32>>>(mode, iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      bw_kernel<<<1, nw * 32>>>(mode, iters, cyc, sink);
      CK(cudaDeviceSynchronize());
      unsigned long long c;
      CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
      const double per_it = (double)c / iters;
      const double el_per_cyc = (double)elems[mode] * 32 * nw / per_it;
      printf("%-22s warps %2d: %8.1f cyc/iter  -> %7.1f elements/cyc/SM (%7.1f B/cyc)\n", names[mode], nw,
             per_it, el_per_cyc, el_per_cyc * 4);
    }
  }
  return 0;
}
