// B200 (sm_100a) forward, second O slab of head dims in (768, 1024] as a GEMM over stashed P tiles.
//
// O[64 rows x D] fp32 fills TMEM at D = 1024, so the forward runs per 512-wide O slab; recomputing S for the
// second slab costs a full Q K^T pass (1.5x the MMA work of the operator). Instead pass 0 of ffpa_fwd_kernel
// stores what its own P V MMA consumed -- the 16-bit P tiles (tile-major, TMA stores from the P buffers), the O
// rescale factor of every (row, KV tile) of its lazy-rescale online softmax, and 1 / rowsum -- and this kernel
// computes   O[:, 512:D] = (sum_tiles rescale-replayed P_tile V_tile[:, 512:D]) / rowsum.
// Numerically identical to what a second softmax pass would produce (same P bits, same rescale schedule).
//
// Tile: 2-CTA cluster = 256 query rows (tcgen05.mma cta_group::2, M = 256; CTA r owns query tile 2 blk + r, rows in
// TMEM lanes 0..127), accumulator [128 x (D - 512)] fp32, P tiles as the K-major A operand, V as the MN-major B
// operand (N = 256 instructions), 2 x 96 KB stages. The eight non-MMA warps check each tile's rescale factors
// (all 1 on almost every tile), scale the accumulator rows when one is not, and release the tile's MMAs.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace replay {

constexpr int kStages = 2;
constexpr int kABytes = 32768;    // P: [128 query rows x 128 keys]
constexpr int kBBytes = 65536;    // V: [128 keys x up to 256 head dims of this CTA]
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSmem = kStages * kStageBytes;
constexpr int kThreads = 320;

struct Barriers {
  uint64_t full[kStages], empty[kStages];
  uint64_t corr[2];   // tile checked / accumulator rescaled: its MMAs may be issued
  uint64_t acc_full, acc_empty;
};

struct Item { int qb, bh, T; };

__device__ __forceinline__ Item decode_item(const FwdReplayParams& p, int item) {
  Item it;
  it.qb = item % p.n_qblocks;
  it.bh = item / p.n_qblocks;
  int tc = (p.seqlen_kv + 127) >> 7;
  if (p.causal) {   // same rule as pass 0 in stash mode: both query tiles of the block walk the odd tile's range
    const int lim = (((it.qb * 256 + 128) + 127 + (p.seqlen_kv - p.seqlen_q)) >> 7) + 1;
    tc = lim < tc ? lim : tc;
  }
  it.T = tc < 1 ? 1 : tc;
  return it;
}

__device__ __forceinline__ int next_item(const FwdReplayParams& p, uint32_t cluster, uint32_t nclusters, uint32_t k) {
  if (p.sched != nullptr) return (k < (uint32_t)p.sched_stride) ? __ldg(p.sched + (size_t)cluster * p.sched_stride + k) : -1;
  const uint32_t item = cluster + k * nclusters;
  return item < (uint32_t)p.n_items ? (int)item : -1;
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ffpa_fwd_replay_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_v,
                       const FwdReplayParams p) {
  constexpr int CG = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Barriers bars;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t cluster = blockIdx.x >> 1;
  const uint32_t nclusters = gridDim.x >> 1;
  auto bar = [](uint64_t& b) { return ptx::smem_u32(&b); };
  auto sA = [&](uint32_t stage) { return smem_base + stage * kStageBytes; };
  auto sB = [&](uint32_t stage) { return smem_base + stage * kStageBytes + kABytes; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { ptx::mbar_init(bar(bars.full[i]), 1); ptx::mbar_init(bar(bars.empty[i]), 1); }
    for (int i = 0; i < 2; ++i) ptx::mbar_init(bar(bars.corr[i]), 2 * 8);
    ptx::mbar_init(bar(bars.acc_full), 1);
    ptx::mbar_init(bar(bars.acc_empty), 2 * 8);
    ptx::fence_mbar_init();
  }
  if (warp == 9 && ptx::elect_one()) { ptx::prefetch_tmap(&map_p); ptx::prefetch_tmap(&map_v); }
  if (warp == 8) {
    ptx::tmem_alloc<CG>(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  const int group = p.heads_q / p.heads_kv;
  const int D = p.head_dim, d_base = 512, w = D - 512;              // second slab: head dims [512, D)
  const int n_slices = (w + 255) >> 8;
  auto slice_n = [&](int s) { return (w - 256 * s) > 128 ? 256 : 128; };
  const int nk64 = p.nk_pad >> 6, nk128 = p.nk_pad >> 7;

  if (warp == 9) {
    // =========================================== TMA producer (both CTAs) =======================
    if (ptx::elect_one()) {
      uint32_t rc = 0;
      const uint64_t pol = ptx::l2_policy_evict_first();   // P tiles are read exactly once
      uint32_t b_bytes = 0;
      for (int s = 0; s < n_slices; ++s) b_bytes += (slice_n(s) / 128) * 16384;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const Item it = decode_item(p, item_s);
        const int h = it.bh % p.heads_q, b = it.bh / p.heads_q, hk = h / group;
        const int mt = 2 * it.qb + (int)rank;   // this CTA's query tile
        for (int ti = 0; ti < it.T; ++ti, ++rc) {
          const uint32_t stage = rc % kStages, n = rc / kStages;
          ptx::mbar_wait(bar(bars.empty[stage]), (n & 1) ^ 1);
          if (rank == 0) ptx::mbar_expect_tx(bar(bars.full[stage]), 2 * (kABytes + b_bytes));
          const uint32_t l_full = ptx::mapa(bar(bars.full[stage]), 0);
          // A: this CTA's [128 rows x 128 keys] P tile = two contiguous 16 KB blocks of the tile-major stash
          const int blk = mt * nk64 + 2 * ti;
          ptx::tma_load_4d_2sm_hint(sA(stage), &map_p, l_full, 0, 0, blk, it.bh, pol);
          ptx::tma_load_4d_2sm_hint(sA(stage) + 16384, &map_p, l_full, 0, 0, blk + 1, it.bh, pol);
          // B: V rows of the KV tile, this CTA's half of every N slice of the second slab as 64-wide boxes
          for (int s = 0; s < n_slices; ++s) {
            const int ns = slice_n(s), nb = ns / 128;
            for (int bx = 0; bx < nb; ++bx)
              ptx::tma_load_4d_2sm(sB(stage) + s * 32768 + bx * 16384, &map_v, l_full,
                                   d_base + 256 * s + (ns / 2) * (int)rank + 64 * bx, ti * 128, hk, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // =========================================== MMA issuer (leader CTA) ========================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc256 = ptx::make_idesc(fmt, fmt, 0, 1, 256, 256);
      constexpr uint32_t idesc128 = ptx::make_idesc(fmt, fmt, 0, 1, 256, 128);
      uint32_t rc = 0, itc = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const Item it = decode_item(p, item_s);
        ptx::mbar_wait_cluster(bar(bars.acc_empty), (itc & 1) ^ 1);   // epilogue of the previous item drained TMEM
        ptx::tc_fence_after();
        for (int ti = 0; ti < it.T; ++ti, ++rc) {
          const uint32_t stage = rc % kStages, n = rc / kStages;
          ptx::mbar_wait_cluster(bar(bars.corr[rc & 1]), (rc >> 1) & 1);   // rescale replay of this tile done
          ptx::mbar_wait(bar(bars.full[stage]), n & 1);
          ptx::tc_fence_after();
          for (int s = 0; s < n_slices; ++s) {
            const uint32_t idesc = slice_n(s) == 256 ? idesc256 : idesc128;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {   // 16 keys per instruction; P box = [128 rows x 64 keys], K-major
              const uint64_t ad = ptx::make_smem_desc_sw128(sA(stage) + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(sB(stage) + s * 32768 + kk * 2048, 16384, 1024);
              ptx::umma_f16_ss<CG>(tmem + 256 * s, ad, bd, idesc, (ti > 0 || kk > 0) ? 1u : 0u);
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
    // =========================================== rescale replay + epilogue warps ================
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t wg = warp >> 2;   // the two warpgroups take alternate 32-column chunks of a row
    const int rrow = (int)(warp & 3) * 32 + (int)ptx::lane_id();   // row inside this CTA's query tile
    const uint32_t l_acc_empty = ptx::mapa(bar(bars.acc_empty), 0);
    const uint32_t l_corr0 = ptx::mapa(bar(bars.corr[0]), 0), l_corr1 = ptx::mapa(bar(bars.corr[1]), 0);
    uint32_t rc = 0, itc = 0;
    for (uint32_t kidx = 0;; ++kidx) {
      const int item_s = next_item(p, cluster, nclusters, kidx);
      if (item_s < 0) break;
      const Item it = decode_item(p, item_s);
      const int h = it.bh % p.heads_q, b = it.bh / p.heads_q;
      const int mt = 2 * it.qb + (int)rank;
      const int gq = mt * 128 + rrow;
      const bool row_ok = gq < p.seqlen_q;
      const float* frow = p.stash_f + (((int64_t)it.bh * p.n_mt_even + mt) * nk128) * 128 + rrow;
      for (int ti = 0; ti < it.T; ++ti, ++rc) {
        const float factor = __ldg(frow + (int64_t)ti * 128);
        // never run more than one tile ahead of the MMA issuer (same wait as the producer's)
        const uint32_t stage = rc % kStages, n = rc / kStages;
        ptx::mbar_wait(bar(bars.empty[stage]), (n & 1) ^ 1);
        if (ti > 0 && __any_sync(0xffffffffu, factor != 1.f)) {
          // the accumulator may only be touched once the MMAs of tile ti - 1 have retired
          const uint32_t ps = (rc - 1) % kStages, pn = (rc - 1) / kStages;
          ptx::mbar_wait(bar(bars.empty[ps]), pn & 1);
          ptx::tc_fence_after();
          for (int c0 = (int)wg * 32; c0 < w; c0 += 64) {
            uint32_t r[32];
            ptx::tmem_ld_x32(tmem + lane_base + c0, r);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * factor);
            ptx::tmem_st_x32(tmem + lane_base + c0, r);
          }
          ptx::tmem_wait_st();
          ptx::tc_fence_before();
        }
        __syncwarp();
        if (ptx::lane_id() == 0) ptx::mbar_arrive_cluster((rc & 1) ? l_corr1 : l_corr0);
      }
      // ---------------- epilogue: O[:, 512:D] = acc / rowsum ----------------
      ptx::mbar_wait(bar(bars.acc_full), itc & 1);
      ptx::tc_fence_after();
      const float inv = p.stash_inv[(int64_t)it.bh * p.n_mt_even * 128 + gq];
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                      2 * ((int64_t)b * p.o_stride[0] + (int64_t)h * p.o_stride[1] + (int64_t)gq * p.o_stride[2]);
      for (int c0 = (int)wg * 32; c0 < w; c0 += 64) {
        uint32_t r[32];
        ptx::tmem_ld_x32(tmem + lane_base + c0, r);
        ptx::tmem_wait_ld();
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int d = d_base + c0 + 8 * v;
            if (d < D) {
              uint32_t wv[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float a = __uint_as_float(r[8 * v + 2 * u]) * inv;
                const float c = __uint_as_float(r[8 * v + 2 * u + 1]) * inv;
                wv[u] = BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
              }
              *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
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
int launch_fwd_replay(const CUtensorMap& map_p, const CUtensorMap& map_v, const FwdReplayParams& kp, int nclusters,
                      cudaStream_t stream) {
  auto kern = ffpa_fwd_replay_kernel<BF16>;
  static bool attr_set[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  dev_id = (dev_id >= 0 && dev_id < 64) ? dev_id : 0;
  if (!attr_set[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaFuncSetAttribute(fwd replay smem=%d): %s", kSmem, cudaGetErrorString(e));
    attr_set[dev_id] = true;
  }
  kern<<<dim3(2 * nclusters), dim3(kThreads), kSmem, stream>>>(map_p, map_v, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "forward replay launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

}  // namespace replay
}  // namespace ffpa
