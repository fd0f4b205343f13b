// Standalone probe for the sm_100a primitives the attention kernels are built from.
// Verifies, on a real B200, (1) TMA SWIZZLE_128B loads + K-major / MN-major UMMA shared-memory
// descriptors, (2) the instruction descriptor, (3) tcgen05.commit -> mbarrier, (4) the TMEM
// accumulator layout for cta_group::1 (M=128) and cta_group::2 (M=128 and M=256), by running one
// GEMM tile C[M,N] = A[M,K] * B^T and comparing a full TMEM dump against a host reference.
// Every wait is bounded, so a mis-programmed descriptor reports an error instead of hanging.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "../ffpa-attn_b200/csrc/sm100_ptx.cuh"

struct ProbeParams {
  int m_cta;    // A rows held by each CTA
  int n_total;  // MMA N
  int k_total;  // K extent (multiple of 64)
  int b_mn;     // 1: B is MN-major (global [K][N], N contiguous), 0: K-major (global [N][K])
  int m_instr;  // instruction M (128 / 256)
  int dump_cols;
};

__device__ __forceinline__ bool wait_bounded(uint32_t bar, uint32_t parity) {
  for (int i = 0; i < (1 << 22); ++i)
    if (ptx::mbar_try_wait_cluster(bar, parity)) return true;
  return false;
}

template <int CG>
__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             float* __restrict__ out, int* __restrict__ err, ProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B align (SWIZZLE_128B requirement)
  uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = (CG == 2) ? ptx::cluster_ctarank() : 0u;
  const uint32_t bar_full = ptx::smem_u32(&bars[0]);
  const uint32_t bar_done = ptx::smem_u32(&bars[1]);

  const int n_cta = p.n_total / CG;
  const uint32_t a_box_bytes = p.m_cta * 128;
  const uint32_t a_bytes = a_box_bytes * (p.k_total / 64);
  const uint32_t b_box_bytes = p.b_mn ? p.k_total * 128 : n_cta * 128;
  const uint32_t b_boxes = p.b_mn ? n_cta / 64 : p.k_total / 64;
  const uint32_t b_bytes = b_box_bytes * b_boxes;
  const uint32_t sA = base;
  const uint32_t sB = base + a_bytes;

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_full, 1);
    ptx::mbar_init(bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) {
    ptx::tmem_alloc<CG>(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 4) {
    // ---- TMA (one lane in each CTA) ----
    if (ptx::elect_one()) {
      const uint32_t lead_full = (CG == 2) ? ptx::mapa(bar_full, 0) : bar_full;
      if (rank == 0) ptx::mbar_expect_tx(bar_full, (a_bytes + b_bytes) * CG);
      for (int kb = 0; kb < p.k_total / 64; ++kb) {
        if (CG == 2) ptx::tma_load_2d_2sm(sA + kb * a_box_bytes, &map_a, lead_full, kb * 64, rank * p.m_cta);
        else ptx::tma_load_2d(sA + kb * a_box_bytes, &map_a, bar_full, kb * 64, 0);
      }
      for (uint32_t bb = 0; bb < b_boxes; ++bb) {
        int c0, c1;
        if (p.b_mn) { c0 = rank * n_cta + bb * 64; c1 = 0; }
        else        { c0 = bb * 64;                c1 = rank * n_cta; }
        if (CG == 2) ptx::tma_load_2d_2sm(sB + bb * b_box_bytes, &map_b, lead_full, c0, c1);
        else ptx::tma_load_2d(sB + bb * b_box_bytes, &map_b, bar_full, c0, c1);
      }
    }
    __syncwarp();
    // ---- MMA (leader CTA only) ----
    if (rank == 0) {
      bool ok = wait_bounded(bar_full, 0);
      if (!ok && ptx::lane_id() == 0) atomicOr(err, 1);
      ptx::tc_fence_after();
      if (ok && ptx::elect_one()) {
        const uint32_t idesc = ptx::make_idesc(1, 1, 0, p.b_mn ? 1 : 0, p.m_instr, p.n_total);
        for (int k = 0; k < p.k_total / 16; ++k) {
          uint64_t ad = ptx::make_smem_desc_sw128(sA + (k >> 2) * a_box_bytes + (k & 3) * 32, 16, 1024);
          uint64_t bd;
          if (p.b_mn) bd = ptx::make_smem_desc_sw128(sB + k * 2048, b_box_bytes, 1024);
          else        bd = ptx::make_smem_desc_sw128(sB + (k >> 2) * b_box_bytes + (k & 3) * 32, 16, 1024);
          ptx::umma_f16_ss<CG>(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
        }
        if (CG == 2) ptx::umma_commit_mc<CG>(bar_done, 0x3);
        else ptx::umma_commit<CG>(bar_done);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    bool ok = wait_bounded(bar_done, 0);
    if (!ok && ptx::lane_id() == 0) atomicOr(err, 2 << rank);
    ptx::tc_fence_after();
    const uint32_t lane = warp * 32 + ptx::lane_id();
    for (int c = 0; c < p.dump_cols; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_x32(tmem + ((warp * 32u) << 16) + c, r);
      ptx::tmem_wait_ld();
      float* dst = out + ((size_t)rank * 128 + lane) * p.dump_cols + c;
      for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
    }
  }
  ptx::tc_fence_before();
  if (CG == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 4) ptx::tmem_dealloc<CG>(tmem, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Case { const char* name; int cg, m_instr, n, k, b_mn; };

int run_case(const Case& cs) {
  const int M = cs.m_instr, N = cs.n, K = cs.k;
  const int m_cta = M / cs.cg;
  printf("=== %s: cta_group::%d M=%d N=%d K=%d B=%s-major\n", cs.name, cs.cg, M, N, K, cs.b_mn ? "MN" : "K");
  int fails = 0;
  for (int mode = 0; mode < 2; ++mode) {  // 0: layout-encoding inputs, 1: random inputs
    std::vector<float> A(M * K, 0.f), B(N * K, 0.f);  // B[n][k] logical
    if (mode == 0) {
      for (int m = 0; m < M; ++m) { A[m * K + 0] = (float)m; A[m * K + 1] = 1.f; }
      for (int n = 0; n < N; ++n) { B[n * K + 0] = 256.f; B[n * K + 1] = (float)n; }
    } else {
      srand(1234);
      for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
      for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
    }
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16(A[i]);
    if (cs.b_mn) { for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[k * N + n] = __float2bfloat16(B[n * K + k]); }
    else         { for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16(B[i]); }
    __nv_bfloat16 *dA, *dB; float* dOut; int* dErr;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2);
    const int dump_cols = 256;
    cudaMalloc(&dOut, 2 * 128 * dump_cols * 4); cudaMalloc(&dErr, 4);
    cudaMemset(dOut, 0xff, 2 * 128 * dump_cols * 4); cudaMemset(dErr, 0, 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);

    CUtensorMap ma, mb;
    { uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}; uint64_t str[1] = {(uint64_t)K * 2}; uint32_t box[2] = {64, (uint32_t)m_cta};
      if (!tmap::encode_sw128(&ma, dA, 2, 2, dims, str, box)) { printf("encode A failed\n"); return 1; } }
    if (cs.b_mn) { uint64_t dims[2] = {(uint64_t)N, (uint64_t)K}; uint64_t str[1] = {(uint64_t)N * 2}; uint32_t box[2] = {64, (uint32_t)K};
      if (!tmap::encode_sw128(&mb, dB, 2, 2, dims, str, box)) { printf("encode B failed\n"); return 1; } }
    else { uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}; uint64_t str[1] = {(uint64_t)K * 2}; uint32_t box[2] = {64, (uint32_t)(N / cs.cg)};
      if (!tmap::encode_sw128(&mb, dB, 2, 2, dims, str, box)) { printf("encode B failed\n"); return 1; } }

    ProbeParams p{m_cta, N, K, cs.b_mn, M, dump_cols};
    size_t smem = (size_t)m_cta * K * 2 + (size_t)(N / cs.cg) * K * 2 + 2048;
    cudaError_t e;
    if (cs.cg == 1) {
      cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      probe_kernel<1><<<1, 192, smem>>>(ma, mb, dOut, dErr, p);
    } else {
      cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim = {2, 1, 1};
      cfg.attrs = at; cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, probe_kernel<2>, ma, mb, dOut, dErr, p);
      if (e != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(e));
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 1; }
    int herr = 0; cudaMemcpy(&herr, dErr, 4, cudaMemcpyDeviceToHost);
    std::vector<float> out(2 * 128 * dump_cols);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
    if (herr) { printf("  mode %d: WAIT TIMEOUT flags=0x%x\n", mode, herr); fails++; }

    // reference
    std::vector<float> C(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; C[m * N + n] = (float)s; }
    // hypotheses: 0 = plain (lane=m within CTA's m_cta rows [needs m_cta==128], col=n)
    //             1 = folded (lane l: m = rank*64 + l%64, n = (l/64)*(N/2) + col)
    int best = -1;
    for (int hyp = 0; hyp < 2; ++hyp) {
      long bad = 0, cnt = 0; double maxerr = 0;
      for (int r = 0; r < cs.cg; ++r) for (int l = 0; l < 128; ++l) {
        int ncols = hyp == 0 ? N : N / 2;
        for (int c = 0; c < ncols && c < dump_cols; ++c) {
          int m, n;
          if (hyp == 0) { if (m_cta != 128) { bad = -1; break; } m = r * 128 + l; n = c; }
          else { if (m_cta != 64) { bad = -1; break; } m = r * 64 + (l % 64); n = (l / 64) * (N / 2) + c; }
          float got = out[((size_t)r * 128 + l) * dump_cols + c];
          float ref = C[m * N + n];
          double er = fabs((double)got - ref);
          if (!(er <= 1e-2 + 1e-3 * fabs(ref))) bad++;
          if (er > maxerr) maxerr = er;
          cnt++;
        }
        if (bad < 0) break;
      }
      if (bad >= 0) printf("  mode %d hyp %d: mismatches %ld / %ld  maxerr %.4g\n", mode, hyp, bad, cnt, maxerr);
      if (bad == 0) best = hyp;
    }
    if (best < 0) {
      fails++;
      if (mode == 0) {
        printf("  raw decode (cta,lane,col)->(m,n):\n");
        int lanes[] = {0, 1, 15, 16, 31, 32, 63, 64, 65, 96, 127};
        for (int r = 0; r < cs.cg; ++r) for (int l : lanes) {
          printf("   cta%d lane%3d:", r, l);
          int cols[] = {0, 1, 2, 31, 32, 63, 64, 127, 128, 255};
          for (int c : cols) { float v = out[((size_t)r * 128 + l) * dump_cols + c]; int iv = (int)v; if (v != v) printf(" c%d=nan", c); else printf(" c%d=(%d,%d)", c, iv / 256, iv % 256); }
          printf("\n");
        }
      }
    } else printf("  mode %d: OK with hypothesis %d\n", mode, best);
    cudaFree(dA); cudaFree(dB); cudaFree(dOut); cudaFree(dErr);
  }
  return fails;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device: %s sm_%d%d SMs=%d smem/blk optin=%zu\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  Case cases[] = {
      {"T1 1cta K-major", 1, 128, 128, 128, 0},
      {"T2 1cta MN-major N256", 1, 128, 256, 128, 1},
      {"T3 2cta M128 K-major N128", 2, 128, 128, 128, 0},
      {"T4 2cta M128 MN-major N256", 2, 128, 256, 128, 1},
      {"T5 2cta M256 K-major N128", 2, 256, 128, 128, 0},
      {"T6 2cta M128 K-major N128 K512", 2, 128, 128, 512, 0},
  };
  int fails = 0;
  for (auto& c : cases) fails += run_case(c);
  printf("PROBE %s (%d failing sub-cases)\n", fails ? "FAIL" : "PASS", fails);
  return fails ? 1 : 0;
}
