// B200 (sm_100a) backward, "stash" path for large head dims: dK / dV as plain tcgen05 GEMMs over score
// tiles the dQ kernel has already produced.
//
// The three-kernel backward recomputes S (and dP) in every kernel: 8 GEMM passes for the 5 the math needs
// (the reference's SM100 backend pays the same, /root/reference/src/ffpa_attn/cute/_ffpa_bwd_sm100.py:301-484).
// At large head dims a GEMM pass costs far more than moving an [Nq x Nkv] 16-bit tile through HBM, so the dQ
// kernel (which owns S, dP, P and dS anyway) stashes P_drop and dS as 16-bit tiles
//     stash[b * Hq + hq][q tile][64-key block][128 q][64 k]   (tile-major: every TMA box is one contiguous
//                                                              8-16 KB run; exactly the values its own MMA consumes)
// and this kernel computes
//     dV[keys, d] = sum_{heads of the group} sum_q P^T[keys, q]  dO[q, d]
//     dK[keys, d] = scale * sum_{...}        sum_q dS^T[keys, q] Q[q, d]
// (math: /root/reference/src/ffpa_attn/triton/_ffpa_bwd.py:692-855) with 5 GEMM passes in total.
//
// Tile: a 2-CTA cluster owns 256 keys (tcgen05.mma cta_group::2, M = 256: CTA r holds keys [128 r, 128 r + 128)
// in TMEM lanes 0..127), the accumulator [128 x D] fp32 fills the CTA's TMEM (512 columns at D = 512) -- no
// S / dP buffers are needed any more, which is what makes M = 256 possible: per streamed query tile a CTA
// loads 32 KB of stash (A, MN-major: keys contiguous) + 64 KB of dO / Q (B, MN-major) for 2048 MMA cycles,
// i.e. 47 B/clk/SM from L2 instead of the 62 B/clk/SM of the 128-row kernels (L2 -> SM fabric cap ~43).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace bwd {

constexpr int kGemmStages = 2;
constexpr int kGemmABytes = 32768;                 // [128 queries x 128 keys of this CTA]
constexpr int kGemmBBytes = 65536;                 // [128 queries x up to 256 head dims of this CTA]
constexpr int kGemmStageBytes = kGemmABytes + kGemmBBytes;
constexpr int kGemmSmem = kGemmStages * kGemmStageBytes;
constexpr int kGemmThreads = 320;

struct GemmBarriers {
  uint64_t full[kGemmStages], empty[kGemmStages];
  uint64_t acc_full, acc_empty;
};

struct GemmItem { int kb, bh, pass, first, count, T; };

__device__ __forceinline__ GemmItem decode_gemm_item(const BwdGemmParams& p, int item, int group) {
  GemmItem it;
  it.kb = item % p.n_kblocks;
  const int rest = item / p.n_kblocks;
  it.pass = rest % p.n_pass;     // 512-wide output slab (head dims > 512: the accumulator is one TMEM)
  it.bh = rest / p.n_pass;
  const int tq = (p.seqlen_q + 127) >> 7;
  int first = 0;
  if (p.causal) {
    const int qmin = it.kb * 256 - (p.seqlen_kv - p.seqlen_q);  // first query that sees the block's first key
    first = qmin > 0 ? (qmin >> 7) : 0;
    if (first > tq) first = tq;
  }
  it.first = first;
  it.count = tq - first;
  it.T = it.count * group;
  return it;
}

__device__ __forceinline__ int next_gemm_item(const BwdGemmParams& p, uint32_t cluster, uint32_t nclusters, uint32_t k) {
  if (p.sched != nullptr) return (k < (uint32_t)p.sched_stride) ? __ldg(p.sched + (size_t)cluster * p.sched_stride + k) : -1;
  const uint32_t item = cluster + k * nclusters;
  return item < (uint32_t)p.n_items ? (int)item : -1;
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
ffpa_bwd_gemm_kernel(const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_b,
                     const BwdGemmParams p) {
  constexpr int CG = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ GemmBarriers bars;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t cluster = blockIdx.x >> 1;
  const uint32_t nclusters = gridDim.x >> 1;
  auto bar = [](uint64_t& b) { return ptx::smem_u32(&b); };
  auto sA = [&](uint32_t stage) { return smem_base + stage * kGemmStageBytes; };
  auto sB = [&](uint32_t stage) { return smem_base + stage * kGemmStageBytes + kGemmABytes; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGemmStages; ++i) { ptx::mbar_init(bar(bars.full[i]), 1); ptx::mbar_init(bar(bars.empty[i]), 1); }
    ptx::mbar_init(bar(bars.acc_full), 1);
    ptx::mbar_init(bar(bars.acc_empty), 2 * 8);   // 8 epilogue warps of both CTAs
    ptx::fence_mbar_init();
  }
  if (warp == 9 && ptx::elect_one()) { ptx::prefetch_tmap(&map_t); ptx::prefetch_tmap(&map_b); }
  if (warp == 8) {
    ptx::tmem_alloc<CG>(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int group = p.heads_q / p.heads_kv;
  const int D = p.head_dim;
  // output slab of pass `ps`: head dims [512 ps, 512 ps + w); N = 256 slices (the last may be 128)
  auto slab_w = [&](int ps) { return (D - 512 * ps) > 512 ? 512 : (D - 512 * ps); };
  auto slice_n = [&](int w, int s) { return (w - 256 * s) > 128 ? 256 : 128; };

  if (warp == 9) {
    // =========================================== TMA producer (both CTAs) =======================
    if (ptx::elect_one()) {
      uint32_t rc = 0;
      const uint64_t pol = ptx::l2_policy_evict_first();   // stash tiles are read once (per slab pass)
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_gemm_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const GemmItem it = decode_gemm_item(p, item_s, group);
        const int hk = it.bh % p.heads_kv, b = it.bh / p.heads_kv;
        const int key0 = it.kb * 256 + 128 * (int)rank;
        const int w = slab_w(it.pass), n_slices = (w + 255) >> 8, d_base = 512 * it.pass;
        uint32_t b_bytes = 0;
        for (int s = 0; s < n_slices; ++s) b_bytes += (slice_n(w, s) / 128) * 16384;
        for (int step = 0; step < it.T; ++step, ++rc) {
          const int gi = step / it.count, ci = it.first + step % it.count;
          const int hq = hk * group + gi;
          const uint32_t stage = rc % kGemmStages, n = rc / kGemmStages;
          ptx::mbar_wait(bar(bars.empty[stage]), (n & 1) ^ 1);
          if (rank == 0) ptx::mbar_expect_tx(bar(bars.full[stage]), 2 * (kGemmABytes + b_bytes));
          const uint32_t l_full = ptx::mapa(bar(bars.full[stage]), 0);
          // A: stash tile, this CTA's 128 keys as two 64-key boxes of 128 query rows (each box is one
          // contiguous 16 KB block of the tile-major stash: [b * Hq + h][query tile][64-key block][128][64])
          const int blk = ci * (p.nk_pad >> 6) + (key0 >> 6), bhq = b * p.heads_q + hq;
          ptx::tma_load_4d_2sm_hint(sA(stage), &map_t, l_full, 0, 0, blk, bhq, pol);
          ptx::tma_load_4d_2sm_hint(sA(stage) + 16384, &map_t, l_full, 0, 0, blk + 1, bhq, pol);
          // B: dO / Q rows of the query tile, this CTA's half of every N slice as 64-wide boxes
          for (int s = 0; s < n_slices; ++s) {
            const int ns = slice_n(w, s), nb = ns / 128;
            for (int bx = 0; bx < nb; ++bx)
              ptx::tma_load_4d_2sm(sB(stage) + s * 32768 + bx * 16384, &map_b, l_full,
                                   d_base + 256 * s + (ns / 2) * (int)rank + 64 * bx, ci * 128, hq, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // =========================================== MMA issuer (leader CTA) ========================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc256 = ptx::make_idesc(fmt, fmt, 1, 1, 256, 256);
      constexpr uint32_t idesc128 = ptx::make_idesc(fmt, fmt, 1, 1, 256, 128);
      uint32_t rc = 0, itc = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_gemm_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const GemmItem it = decode_gemm_item(p, item_s, group);
        if (it.T <= 0) continue;
        const int w = slab_w(it.pass), n_slices = (w + 255) >> 8;
        // accumulator free? (epilogue of the previous item has drained TMEM)
        ptx::mbar_wait_cluster(bar(bars.acc_empty), (itc & 1) ^ 1);
        ptx::tc_fence_after();
        for (int step = 0; step < it.T; ++step, ++rc) {
          const uint32_t stage = rc % kGemmStages, n = rc / kGemmStages;
          ptx::mbar_wait(bar(bars.full[stage]), n & 1);
          ptx::tc_fence_after();
          for (int s = 0; s < n_slices; ++s) {
            const uint32_t idesc = slice_n(w, s) == 256 ? idesc256 : idesc128;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {   // 16 queries per instruction
              const uint64_t ad = ptx::make_smem_desc_sw128(sA(stage) + kk * 2048, 16384, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(sB(stage) + s * 32768 + kk * 2048, 16384, 1024);
              ptx::umma_f16_ss<CG>(tmem + 256 * s, ad, bd, idesc, (step > 0 || kk > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit_mc<CG>(bar(bars.empty[stage]), 0x3);
        }
        ptx::umma_commit_mc<CG>(bar(bars.acc_full), 0x3);
        ++itc;
      }
    }
    __syncwarp();
  } else {
    // =========================================== epilogue warps =================================
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t wg = warp >> 2;   // the two warpgroups take alternate 32-column chunks
    const uint32_t l_acc_empty = ptx::mapa(bar(bars.acc_empty), 0);
    uint32_t itc = 0;
    for (uint32_t kidx = 0;; ++kidx) {
      const int item_s = next_gemm_item(p, cluster, nclusters, kidx);
      if (item_s < 0) break;
      const GemmItem it = decode_gemm_item(p, item_s, group);
      const int hk = it.bh % p.heads_kv, b = it.bh / p.heads_kv;
      const int key = it.kb * 256 + 128 * (int)rank + (int)(warp & 3) * 32 + (int)ptx::lane_id();
      const bool row_ok = key < p.seqlen_kv;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) +
                      2 * ((int64_t)b * p.out_stride[0] + (int64_t)hk * p.out_stride[1] + (int64_t)key * p.out_stride[2]);
      const int w = slab_w(it.pass), d_base = 512 * it.pass;
      if (it.T <= 0) {   // no query sees these keys: zero gradient
        if (row_ok)
          for (int d = d_base + (int)wg * 8; d < d_base + w; d += 16) *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(0, 0, 0, 0);
        continue;
      }
      ptx::mbar_wait(bar(bars.acc_full), itc & 1);
      ptx::tc_fence_after();
      for (int c0 = (int)wg * 32; c0 < w; c0 += 64) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + lane_base + c0, r);
        ptx::tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int d = d_base + c0 + 8 * v;
            if (d < D) {
              uint32_t w[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float a = __uint_as_float(r[8 * v + 2 * u]) * p.mul;
                const float c = __uint_as_float(r[8 * v + 2 * u + 1]) * p.mul;
                w[u] = BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
              }
              *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (ptx::lane_id() == 0) ptx::mbar_arrive_cluster(l_acc_empty);
      ++itc;
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 8) ptx::tmem_dealloc<CG>(tmem, 512);
}

template <bool BF16>
int launch_bwd_gemm(const CUtensorMap& map_t, const CUtensorMap& map_b, const BwdGemmParams& kp, int nclusters,
                    cudaStream_t stream) {
  auto kern = ffpa_bwd_gemm_kernel<BF16>;
  static bool attr_set[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  dev_id = (dev_id >= 0 && dev_id < 64) ? dev_id : 0;
  if (!attr_set[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaFuncSetAttribute(bwd gemm smem=%d): %s", kGemmSmem, cudaGetErrorString(e));
    attr_set[dev_id] = true;
  }
  kern<<<dim3(2 * nclusters), dim3(kGemmThreads), kGemmSmem, stream>>>(map_t, map_b, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "backward gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

}  // namespace bwd
}  // namespace ffpa
