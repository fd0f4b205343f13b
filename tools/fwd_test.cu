// Standalone forward harness (no torch): calls the C ABI on random inputs, checks sampled rows
// against a double-precision host computation, and times the kernel with CUDA events.
//   nvcc -O2 -std=c++17 -o build/fwd_test tools/fwd_test.cu -Iinclude -Lffpa-attn_b200/ffpa_attn -lffpa_b200
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <algorithm>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "ffpa_b200.h"

struct Case { int B, Hq, Hkv, Nq, Nkv, D, causal, bf16, time_iters; };

static uint32_t rng_state = 12345;
static float frand() { rng_state = rng_state * 1664525u + 1013904223u; return ((rng_state >> 8) & 0xFFFF) / 65536.f * 2.f - 1.f; }
static float gauss() { float s = 0; for (int i = 0; i < 6; ++i) s += frand(); return s * 0.70710678f * 1.0f; }  // ~N(0, ~0.7)

template <typename T> static float tof(T x);
template <> float tof<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> float tof<__half>(__half x) { return __half2float(x); }
template <typename T> static T fromf(float x);
template <> __nv_bfloat16 fromf<__nv_bfloat16>(float x) { return __float2bfloat16(x); }
template <> __half fromf<__half>(float x) { return __float2half(x); }

template <typename T>
int run(const Case& c) {
  const size_t nq = (size_t)c.B * c.Hq * c.Nq * c.D, nk = (size_t)c.B * c.Hkv * c.Nkv * c.D;
  std::vector<T> hq(nq), hk(nk), hv(nk);
  for (auto& x : hq) x = fromf<T>(gauss());
  for (auto& x : hk) x = fromf<T>(gauss());
  for (auto& x : hv) x = fromf<T>(gauss());
  T *dq, *dk, *dv, *dout; float* dlse;
  cudaMalloc(&dq, nq * 2); cudaMalloc(&dk, nk * 2); cudaMalloc(&dv, nk * 2); cudaMalloc(&dout, nq * 2);
  cudaMalloc(&dlse, (size_t)c.B * c.Hq * c.Nq * 4);
  cudaMemcpy(dq, hq.data(), nq * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, hk.data(), nk * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dv, hv.data(), nk * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, nq * 2);
  cudaMemset(dlse, 0xff, (size_t)c.B * c.Hq * c.Nq * 4);

  ffpa_fwd_params p{};
  p.q = dq; p.k = dk; p.v = dv; p.o = dout; p.lse = dlse;
  auto set = [&](int64_t* s, int H, int N) { s[0] = (int64_t)H * N * c.D; s[1] = (int64_t)N * c.D; s[2] = c.D; s[3] = 1; };
  set(p.q_stride, c.Hq, c.Nq); set(p.k_stride, c.Hkv, c.Nkv); set(p.v_stride, c.Hkv, c.Nkv); set(p.o_stride, c.Hq, c.Nq);
  p.batch = c.B; p.heads_q = c.Hq; p.heads_kv = c.Hkv; p.seqlen_q = c.Nq; p.seqlen_kv = c.Nkv; p.head_dim = c.D;
  p.dtype = c.bf16 ? FFPA_DTYPE_BF16 : FFPA_DTYPE_F16; p.causal = c.causal;
  p.softmax_scale = 1.f / sqrtf((float)c.D);
  int rc = ffpa_b200_fwd(&p, nullptr);
  if (rc) { printf("  fwd rc=%d: %s\n", rc, ffpa_b200_last_error()); return 1; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<T> ho(nq); std::vector<float> hl((size_t)c.B * c.Hq * c.Nq);
  cudaMemcpy(ho.data(), dout, nq * 2, cudaMemcpyDeviceToHost);
  cudaMemcpy(hl.data(), dlse, hl.size() * 4, cudaMemcpyDeviceToHost);

  // sampled verification
  const int group = c.Hq / c.Hkv;
  double max_err = 0, max_lse_err = 0; int nan_cnt = 0;
  std::vector<int> rows;
  if (c.Nq <= 512) for (int r = 0; r < c.Nq; ++r) rows.push_back(r);
  else { int cand[] = {0, 1, 63, 64, 65, 127, 128, 129, 191, 192, 255, 256, c.Nq / 2 - 1, c.Nq / 2, c.Nq - 130, c.Nq - 129, c.Nq - 128, c.Nq - 65, c.Nq - 64, c.Nq - 2, c.Nq - 1}; for (int r : cand) if (r >= 0 && r < c.Nq) rows.push_back(r); }
  std::vector<int> bhs;
  { int tot = c.B * c.Hq; if (tot <= 4) for (int i = 0; i < tot; ++i) bhs.push_back(i); else { bhs = {0, 1, tot / 2, tot - 1}; } }
  std::vector<double> s(c.Nkv), o(c.D);
  for (int bh : bhs) {
    const int b = bh / c.Hq, h = bh % c.Hq, hk_ = h / group;
    const T* K = hk.data() + ((size_t)b * c.Hkv + hk_) * c.Nkv * c.D;
    const T* V = hv.data() + ((size_t)b * c.Hkv + hk_) * c.Nkv * c.D;
    for (int r : rows) {
      const T* Q = hq.data() + (((size_t)b * c.Hq + h) * c.Nq + r) * c.D;
      const int lim = c.causal ? std::min(c.Nkv - 1, r + c.Nkv - c.Nq) : c.Nkv - 1;
      double mx = -1e300;
      for (int k = 0; k <= lim; ++k) { double a = 0; for (int d = 0; d < c.D; ++d) a += (double)tof(Q[d]) * tof(K[(size_t)k * c.D + d]); s[k] = a * p.softmax_scale; mx = std::max(mx, s[k]); }
      double sum = 0; for (int k = 0; k <= lim; ++k) { s[k] = exp(s[k] - mx); sum += s[k]; }
      std::fill(o.begin(), o.end(), 0.0);
      for (int k = 0; k <= lim; ++k) { const double w = s[k] / sum; for (int d = 0; d < c.D; ++d) o[d] += w * tof(V[(size_t)k * c.D + d]); }
      const T* O = ho.data() + (((size_t)b * c.Hq + h) * c.Nq + r) * c.D;
      for (int d = 0; d < c.D; ++d) { float g = tof(O[d]); if (g != g) { nan_cnt++; continue; } max_err = std::max(max_err, fabs((double)g - o[d])); }
      const float gl = hl[((size_t)b * c.Hq + h) * c.Nq + r];
      if (gl != gl) nan_cnt++; else max_lse_err = std::max(max_lse_err, fabs((double)gl - (mx + log(sum))));
    }
  }
  const bool ok = nan_cnt == 0 && max_err < 2e-2 && max_lse_err < 1e-2;
  printf("  check: rows=%zu x bh=%zu  max|dO|=%.3e  max|dLSE|=%.3e  nan=%d  -> %s\n", rows.size(), bhs.size(), max_err, max_lse_err, nan_cnt, ok ? "OK" : "FAIL");

  if (c.time_iters > 0 && ok) {
    for (int i = 0; i < 3; ++i) ffpa_b200_fwd(&p, nullptr);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < c.time_iters; ++i) ffpa_b200_fwd(&p, nullptr);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= c.time_iters;
    double pairs = c.causal ? ((double)c.Nq * (c.Nkv - c.Nq) + (double)c.Nq * (c.Nq + 1) / 2) : (double)c.Nq * c.Nkv;
    double flops = 4.0 * c.B * c.Hq * c.D * pairs;
    printf("  time: %.4f ms  -> %.1f TFLOPS\n", ms, flops / ms * 1e-9);
  }
  cudaFree(dq); cudaFree(dk); cudaFree(dv); cudaFree(dout); cudaFree(dlse);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  std::vector<Case> cases = {
      {1, 2, 2, 128, 128, 512, 0, 1, 0},
      {1, 2, 2, 512, 512, 512, 0, 1, 0},
      {1, 2, 2, 512, 512, 320, 0, 1, 0},
      {1, 2, 2, 512, 512, 64, 0, 1, 0},
      {1, 2, 2, 512, 512, 128, 0, 1, 0},
      {1, 2, 2, 512, 512, 256, 0, 1, 0},
      {1, 2, 2, 512, 512, 384, 0, 1, 0},
      {1, 2, 2, 512, 512, 448, 0, 0, 0},
      {2, 4, 2, 500, 700, 512, 0, 1, 0},
      {1, 4, 1, 512, 512, 512, 1, 1, 0},
      {1, 2, 2, 300, 1000, 320, 1, 0, 0},
      {1, 2, 2, 1, 513, 512, 0, 1, 0},
      {1, 32, 32, 8192, 8192, 512, 0, 1, 10},
      {1, 32, 8, 8192, 8192, 512, 0, 1, 10},
      {1, 32, 32, 8192, 8192, 512, 1, 1, 10},
      {1, 32, 32, 8192, 8192, 320, 0, 1, 10},
      {1, 32, 32, 8192, 8192, 256, 0, 1, 10},
      {1, 32, 32, 8192, 8192, 128, 0, 1, 10},
      {1, 2, 2, 512, 512, 768, 0, 1, 0},
      {1, 2, 2, 300, 400, 1024, 1, 1, 0},
      {1, 2, 2, 512, 512, 640, 0, 0, 0},
      {1, 2, 2, 260, 512, 896, 0, 1, 0},
      {1, 32, 32, 8192, 8192, 768, 0, 1, 5},
      {1, 32, 32, 8192, 8192, 1024, 0, 1, 5},
      {1, 32, 8, 4096, 4096, 512, 1, 1, 10},
  };
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int fails = 0;
  for (size_t i = 0; i < cases.size(); ++i) {
    if (only >= 0 && (int)i != only) continue;
    const Case& c = cases[i];
    printf("[%zu] B=%d Hq=%d Hkv=%d Nq=%d Nkv=%d D=%d causal=%d %s\n", i, c.B, c.Hq, c.Hkv, c.Nq, c.Nkv, c.D, c.causal, c.bf16 ? "bf16" : "fp16");
    fflush(stdout);
    int r = c.bf16 ? run<__nv_bfloat16>(c) : run<__half>(c);
    if (r == 2) { printf("aborting after CUDA error\n"); return 2; }
    fails += r;
  }
  printf("FWD_TEST %s (%d failing)\n", fails ? "FAIL" : "PASS", fails);
  return fails ? 1 : 0;
}
