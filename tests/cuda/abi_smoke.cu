// Torch-free exercise of the C ABI (include/mlt_attn.h): cudaMalloc buffers -> mlt_gl_attn_fwd / bwd
// (fp32 on the SIMT kernels, then bf16 on the tcgen05 kernels) -> compare with the fp64-oracle fixture
// tests/golden/gl_abi_fixture.bin (written by tests/golden/make_abi_fixture.py).  Nothing but the CUDA
// runtime and libmlt_attn.so is linked: this is what a non-Python host (the TF shim, a C++ server) sees.
//
//   nvcc -std=c++17 -arch=sm_100a tests/cuda/abi_smoke.cu -Iinclude -L<pkg> -lmlt_attn -o abi_smoke
//   ./abi_smoke tests/golden/gl_abi_fixture.bin        -> prints one "name max_abs_err max_abs_ref" line per tensor
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mlt_attn.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      std::exit(2);                                                                    \
    }                                                                                  \
  } while (0)

struct Host {
  std::vector<float> f;
};

static std::vector<float> read_f32(FILE* fp, size_t n) {
  std::vector<float> v(n);
  if (std::fread(v.data(), 4, n, fp) != n) { std::fprintf(stderr, "short fixture\n"); std::exit(2); }
  return v;
}
static std::vector<int32_t> read_i32(FILE* fp, size_t n) {
  std::vector<int32_t> v(n);
  if (std::fread(v.data(), 4, n, fp) != n) { std::fprintf(stderr, "short fixture\n"); std::exit(2); }
  return v;
}

// device buffer holding `h` as fp32 or bf16
static void* upload(const std::vector<float>& h, bool bf16) {
  void* d = nullptr;
  if (!bf16) {
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  } else {
    std::vector<__nv_bfloat16> t(h.size());
    for (size_t i = 0; i < h.size(); ++i) t[i] = __float2bfloat16(h[i]);
    CK(cudaMalloc(&d, h.size() * 2));
    CK(cudaMemcpy(d, t.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  return d;
}
static std::vector<float> download(const void* d, size_t n, bool bf16) {
  std::vector<float> h(n);
  if (!bf16) {
    CK(cudaMemcpy(h.data(), d, n * 4, cudaMemcpyDeviceToHost));
  } else {
    std::vector<__nv_bfloat16> t(n);
    CK(cudaMemcpy(t.data(), d, n * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; ++i) h[i] = __bfloat162float(t[i]);
  }
  return h;
}
static mlt_tensor4 view(void* p, int len, int H, int d) {
  return mlt_tensor4{p, (int64_t)len * H * d, (int64_t)H * d, (int64_t)d};
}
static int report(const char* tag, const char* name, const std::vector<float>& got, const std::vector<float>& want,
                  double tol_rel) {
  double err = 0, mag = 0;
  for (size_t i = 0; i < got.size(); ++i) {
    err = std::fmax(err, std::fabs((double)got[i] - want[i]));
    mag = std::fmax(mag, std::fabs((double)want[i]));
  }
  const bool ok = err <= tol_rel * std::fmax(1.0, mag) && std::isfinite(err);
  std::printf("%s %s max_abs_err %.3e max_abs_ref %.3e %s\n", tag, name, err, mag, ok ? "ok" : "FAIL");
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: abi_smoke <fixture.bin>\n"); return 2; }
  FILE* fp = std::fopen(argv[1], "rb");
  if (!fp) { std::perror("fixture"); return 2; }
  const std::vector<int32_t> hd = read_i32(fp, 8);
  const int B = hd[0], L = hd[1], G = hd[2], H = hd[3], d = hd[4], r = hd[5], R = hd[6], D = hd[7];
  const size_t nl = (size_t)B * L * H * d, ng = (size_t)B * G * H * d, ne = (size_t)R * H * d, nb = (size_t)R * H;
  const size_t in_sizes[12] = {nl, nl, nl, ng, ng, ng, ne, nb, ne, nb, nl, ng};
  std::vector<std::vector<float>> in(12);
  for (int i = 0; i < 12; ++i) in[i] = read_f32(fp, in_sizes[i]);
  const std::vector<int32_t> le = read_i32(fp, (size_t)B * L), ge = read_i32(fp, (size_t)B * G), se = read_i32(fp, (size_t)B * L);
  const size_t out_sizes[12] = {nl, ng, nl, nl, nl, ng, ng, ng, ne, nb, ne, nb};
  const char* out_names[12] = {"long_out", "global_out", "d_long_q", "d_long_k", "d_long_v", "d_global_q",
                               "d_global_k", "d_global_v", "d_long_emb", "d_long_bias", "d_global_emb", "d_global_bias"};
  std::vector<std::vector<float>> want(12);
  for (int i = 0; i < 12; ++i) want[i] = read_f32(fp, out_sizes[i]);
  std::fclose(fp);
  if (mlt_abi_version() != MLT_ABI_VERSION) { std::fprintf(stderr, "ABI version mismatch\n"); return 2; }

  int32_t *d_le, *d_ge, *d_se;
  CK(cudaMalloc(&d_le, le.size() * 4)); CK(cudaMemcpy(d_le, le.data(), le.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_ge, ge.size() * 4)); CK(cudaMemcpy(d_ge, ge.data(), ge.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_se, se.size() * 4)); CK(cudaMemcpy(d_se, se.data(), se.size() * 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  int failures = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const bool bf16 = pass == 1;
    const char* tag = bf16 ? "bf16/tcgen05" : "fp32/simt";
    const size_t es = bf16 ? 2 : 4;
    void* din[12];
    for (int i = 0; i < 12; ++i) din[i] = upload(in[i], bf16);
    void *lo, *go, *dq[6];
    float *lstats, *gstats, *tab[4];
    CK(cudaMalloc(&lo, nl * es)); CK(cudaMalloc(&go, ng * es));
    for (int i = 0; i < 6; ++i) CK(cudaMalloc(&dq[i], (i < 3 ? nl : ng) * es));
    CK(cudaMalloc(&lstats, (size_t)B * H * L * 2 * 4)); CK(cudaMalloc(&gstats, (size_t)B * H * G * 2 * 4));
    CK(cudaMalloc(&tab[0], ne * 4)); CK(cudaMalloc(&tab[1], nb * 4)); CK(cudaMalloc(&tab[2], ne * 4)); CK(cudaMalloc(&tab[3], nb * 4));

    mlt_gl_params p = {};
    p.abi_version = MLT_ABI_VERSION;
    p.dtype = bf16 ? MLT_BF16 : MLT_F32;
    p.impl = bf16 ? MLT_IMPL_TC : MLT_IMPL_SIMT;
    p.B = B; p.L = L; p.G = G; p.H = H; p.d = d; p.R = R; p.local_radius = r;
    p.scale = 1.0f / std::sqrt((float)d);
    p.neg = -1e9f;
    p.long_q = view(din[0], L, H, d); p.long_k = view(din[1], L, H, d); p.long_v = view(din[2], L, H, d);
    p.global_q = view(din[3], G, H, d); p.global_k = view(din[4], G, H, d); p.global_v = view(din[5], G, H, d);
    p.long_out = view(lo, L, H, d); p.global_out = view(go, G, H, d);
    p.long_stats = lstats; p.global_stats = gstats;
    p.long_tables = {din[6], din[7]}; p.global_tables = {din[8], din[9]};
    p.side_mode = MLT_SIDE_COMPACT;
    p.long_example_ids = d_le; p.global_example_ids = d_ge; p.sentence_ids = d_se; p.max_distance = D;
    size_t nws = mlt_gl_workspace_bytes(&p, 1);
    void* ws;
    CK(cudaMalloc(&ws, nws));
    p.workspace = ws; p.workspace_bytes = nws;
    if (bf16 && !mlt_gl_uses_tensor_cores(&p)) { std::fprintf(stderr, "tcgen05 path not selected\n"); return 2; }
    int rc = mlt_gl_attn_fwd(&p, st);
    if (rc != MLT_OK) { std::fprintf(stderr, "mlt_gl_attn_fwd: %s\n", mlt_strerror(rc)); return 2; }
    mlt_gl_grads g = {};
    g.d_long_out = view(din[10], L, H, d); g.d_global_out = view(din[11], G, H, d);
    g.d_long_q = view(dq[0], L, H, d); g.d_long_k = view(dq[1], L, H, d); g.d_long_v = view(dq[2], L, H, d);
    g.d_global_q = view(dq[3], G, H, d); g.d_global_k = view(dq[4], G, H, d); g.d_global_v = view(dq[5], G, H, d);
    g.d_long_emb = tab[0]; g.d_long_bias = tab[1]; g.d_global_emb = tab[2]; g.d_global_bias = tab[3];
    rc = mlt_gl_attn_bwd(&p, &g, st);
    if (rc != MLT_OK) { std::fprintf(stderr, "mlt_gl_attn_bwd: %s\n", mlt_strerror(rc)); return 2; }
    CK(cudaStreamSynchronize(st));
    const double tol = bf16 ? 2e-2 : 1e-5;   // north-star tolerances (bf16: inputs rounded to bf16 as well)
    failures += report(tag, out_names[0], download(lo, nl, bf16), want[0], bf16 ? 3e-2 : tol);
    failures += report(tag, out_names[1], download(go, ng, bf16), want[1], bf16 ? 3e-2 : tol);
    for (int i = 0; i < 6; ++i) failures += report(tag, out_names[2 + i], download(dq[i], i < 3 ? nl : ng, bf16), want[2 + i], bf16 ? 3e-2 : tol);
    for (int i = 0; i < 4; ++i) failures += report(tag, out_names[8 + i], download(tab[i], i % 2 ? nb : ne, false), want[8 + i], bf16 ? 3e-2 : tol);
    // error behaviour across the boundary: codes, never exceptions
    mlt_gl_params bad = p;
    bad.workspace_bytes = 16;
    if (mlt_gl_attn_bwd(&bad, &g, st) != MLT_ERR_WORKSPACE) { std::printf("%s workspace check FAIL\n", tag); ++failures; }
    bad = p;
    bad.dropout_p = 1.5f;
    if (mlt_gl_attn_fwd(&bad, st) != MLT_ERR_DROPOUT) { std::printf("%s dropout check FAIL\n", tag); ++failures; }
    for (int i = 0; i < 12; ++i) CK(cudaFree(din[i]));
    CK(cudaFree(lo)); CK(cudaFree(go)); CK(cudaFree(ws)); CK(cudaFree(lstats)); CK(cudaFree(gstats));
    for (int i = 0; i < 6; ++i) CK(cudaFree(dq[i]));
    for (int i = 0; i < 4; ++i) CK(cudaFree(tab[i]));
  }
  std::printf("launches %lld failures %d\n", mlt_launch_count(), failures);
  return failures ? 1 : 0;
}
