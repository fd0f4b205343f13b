// B200 (sm_100a) FP8 attention forward: tcgen05.mma kind::f8f6f4 (e4m3 x e4m3 -> fp32 TMEM).
//
// Semantics follow the reference's FP8 path (per-block scales, P quantised with the V scale folded
// in; /root/reference/csrc/cuffpa/cute/fp8/quantize_fp8.cuh:56-168, fp8_pscale.cuh:11-76,
// cute/fp8/sm_120/split_d.cuh:26-43) re-designed for sm_100a:
//   Q8 = e4m3(Q / qs), qs = amax(128-row block) / 448     (same for K8/ks, V8/vs)
//   S  = (Q8 K8^T) * qs * ks[tile] * scale                 -> online softmax (lazy threshold 4.0)
//   P8 = e4m3(P * (vs[tile] / vref) * 28),  vref = max_tile vs,  28 = 448 / 2^4 (lazy-rescale slack)
//   O  = (sum_tiles P8 V8) * vref / 28 / l
// One fused pre-pass (quantize_e4m3_kernel) writes Q8/K8/V8 + scales once; the attention kernel
// streams 1-byte tiles (half the L2->SMEM bytes of the bf16 kernel). V stays row-major [keys x d]
// and is consumed as an MN-major B operand (the reference pre-transposes V for its mma.sync path).
// Kernel structure (2-CTA cluster, lane-folded TMEM accumulators, warp roles, mbarrier pipelines)
// is the one of ffpa_fwd_sm100.cuh; byte geometry of the tiles is identical (128-byte swizzled
// rows), only the element count per row doubles (128 e4m3 per row, MMA K = 32).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace fp8 {

constexpr int kSmemLimit = 232448;
constexpr float kLazyThreshold = 4.0f;   // /root/reference/csrc/cuffpa/common.cuh:14-18 (fp8)
constexpr float kPScale = 28.0f;         // 448 / 2^kLazyThreshold

struct Fp8KernelParams {
  void* o;
  float* lse;
  int64_t lse_bh_stride;
  int64_t o_stride[3];
  const float* qs;     // [B, Hq,  TQ]
  const float* ks;     // [B, Hkv, TK]
  const float* vs;     // [B, Hkv, TK]
  const float* vref;   // [B, Hkv]
  const float* qkm;    // [B, Hq, Nq]  q . mean_seq(K) (smooth-K LSE correction) or nullptr
  const float* vsum;   // [B, Hkv, D]  column sums of V over the sequence (smooth-V: added back as mean) or nullptr
  const float* vamax;  // [B, Hkv, D]  per-channel |V| maxima (per-channel V scales, applied in the epilogue) or nullptr
  int tq, tk;
  int batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int causal;
  float scale_log2;
  int n_mtiles, n_items;
};

template <int NB>  // number of 128-wide head-dim boxes (head_dim padded to NB*128)
struct Fp8Cfg {
  static constexpr int HD = NB * 128;
  static constexpr int DVP = ((HD + 255) / 256) * 256;
  static constexpr int NSLICE = DVP / 256;
  static constexpr int O_COLS = DVP / 2;
  static constexpr int KST = (NB + 1) / 2;         // 16 KB K stages ([64 keys x 256 d]) per KV tile
  static constexpr int S_BASE = O_COLS;            // S stages follow the O accumulator in TMEM
  static constexpr int Q_BYTES = NB * 8192;
  // S (TMEM) / P (SMEM) pipeline depth = every 64-column stage that fits behind O: 4 at head dims 257..512, 6 at
  // head dims <= 256. Deeper than the number of softmax warpgroups matters: with depth == warpgroups the QK MMA of a
  // warpgroup's NEXT tile can only be issued once that warpgroup has finished its current one (same S stage), so
  // every warpgroup idles for a QK + PV + signalling round trip per tile (31 % of all warp samples sat in the
  // s_full wait at C4, profiles/r02_fp8_c4_ncu.md).
  static constexpr int KSTG = (512 - O_COLS) / 64 > 6 ? 6 : (512 - O_COLS) / 64;
  static constexpr int P_BYTES = KSTG * 8192;
  static constexpr int kBudget = kSmemLimit - 8192;   // static smem: barriers, exchange buffers (4 softmax warpgroups)
  static constexpr int kAvail = (kBudget - Q_BYTES - P_BYTES) / 16384;
  static constexpr int NVS = (kAvail / 2) > 6 ? 6 : (kAvail / 2);    // 16 KB V stages ([128 keys x 128 d])
  static constexpr int NKS = (kAvail - NVS) > 8 ? 8 : (kAvail - NVS);
  static constexpr int SMEM_DYN = Q_BYTES + P_BYTES + (NKS + NVS) * 16384;
  static_assert(O_COLS <= 256, "fp8 kernel supports head_dim <= 512");
  static_assert(NKS >= 2 && NVS >= 2, "not enough shared memory");
};

struct Barriers {
  uint64_t q_full, q_empty;
  uint64_t k_full[8], k_empty[8];
  uint64_t v_full[6], v_empty[6];
  uint64_t s_full[6];
  uint64_t p_full[6], p_empty[6];
  uint64_t m_full[4];
};

__device__ __forceinline__ int num_kv_tiles(const Fp8KernelParams& p, int q0) {
  int tc = (p.seqlen_kv + 127) >> 7;
  if (p.causal) {
    int lim = ((q0 + 127 + (p.seqlen_kv - p.seqlen_q)) >> 7) + 1;
    tc = lim < tc ? lim : tc;
  }
  return tc < 1 ? 1 : tc;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)),
        "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
// two floats -> two e4m3 bytes (lo at the lower address)
__device__ __forceinline__ uint32_t pack_e4m3x2(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  return pack_e4m3x2(a, b) | (pack_e4m3x2(c, d) << 16);
}

// NWG softmax warpgroups (4 warps each) take KV tiles round robin (tile i -> warpgroup i % NWG); warps NWG*4 / NWG*4+1
// are the MMA issuer / TMA producer. NWG = 4 keeps four warps per scheduler in four different phases of four different
// tiles: the softmax of one tile is a chain of dependent waits (S ready -> TMEM load -> max -> SMEM exchange -> running
// max of the previous tile -> exp2 -> P store), and with two warps per scheduler the issue slots sat idle 59 % of the
// time while both waited (profiles/r02_fp8_c4_ncu.md).
template <int NB, bool OUT_BF16, int NWG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((4 * NWG + 3) * 32, 1)
ffpa_fwd_fp8_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const Fp8KernelParams p) {
  using Cfg = Fp8Cfg<NB>;
  // Two MMA issuers: at small head dims a KV tile is only 512 tensor-pipe cycles (12 MMA instructions), and ONE thread
  // that waits on k_full, issues 8 QK MMAs, commits, waits on p_full / v_full, issues 4 PV MMAs and commits again could
  // not keep up: tensor pipe 53 % busy while the softmax warps spent 24 % of all samples waiting for S
  // (profiles/r02_fp8_c4_ncu.md). The QK issuer never blocks on the softmax of an older tile except for the S stage it
  // is about to overwrite.
  constexpr int kMmaWarp = 4 * NWG, kTmaWarp = 4 * NWG + 1, kPvWarp = 4 * NWG + 2;
  static_assert(NWG == 2 || NWG == 4, "softmax warpgroups: 2 or 4");
  static_assert(NWG <= Cfg::KSTG, "one S / P stage per warpgroup in flight");
  constexpr int CG = 2;
  constexpr uint32_t KS = Cfg::KSTG;
  constexpr int LA = Cfg::KSTG - 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Barriers bars;
  __shared__ float xchw[NWG][2][2][64];  // row-max exchange between the two lane halves: [wg][parity][kh][row]
  __shared__ float mval[4][64];        // running row max, published per tile (ring of 4: a lagging warp of the
                                       // consumer warpgroup is at most one of its own tiles behind)
  __shared__ float xl[2 * NWG][64];    // end-of-item row-sum exchange
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sQ = smem_base;
  const uint32_t sP = sQ + Cfg::Q_BYTES;
  const uint32_t sK = sP + Cfg::P_BYTES;
  const uint32_t sV = sK + Cfg::NKS * 16384;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t cluster = blockIdx.x >> 1;
  const uint32_t nclusters = gridDim.x >> 1;
  auto bar = [](uint64_t& b) { return ptx::smem_u32(&b); };

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar(bars.q_full), 1);
    ptx::mbar_init(bar(bars.q_empty), 1);
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(bar(bars.k_full[i]), 1); ptx::mbar_init(bar(bars.k_empty[i]), 1); }
    for (int i = 0; i < 6; ++i) { ptx::mbar_init(bar(bars.v_full[i]), 1); ptx::mbar_init(bar(bars.v_empty[i]), 1); }
    for (int i = 0; i < 6; ++i) {
      ptx::mbar_init(bar(bars.s_full[i]), 1);
      ptx::mbar_init(bar(bars.p_full[i]), 8);  // 4 warps of one warpgroup x 2 CTAs
      ptx::mbar_init(bar(bars.p_empty[i]), 1);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(bar(bars.m_full[i]), 2);  // the two kh == 0 warps of the publishing warpgroup
    ptx::fence_mbar_init();
  }
  if (warp == kTmaWarp && ptx::elect_one()) {
    ptx::prefetch_tmap(&map_q);
    ptx::prefetch_tmap(&map_k);
    ptx::prefetch_tmap(&map_v);
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc<CG>(ptx::smem_u32(&tmem_slot), 512);
    ptx::tmem_relinquish<CG>();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int group = p.heads_q / p.heads_kv;

  if (warp == kTmaWarp) {
    // =========================================== TMA producer ===================================
    if (ptx::elect_one()) {
      uint32_t kc = 0, vc = 0, it = 0;
      const uint32_t l_q_full = ptx::mapa(bar(bars.q_full), 0);
      for (uint32_t item = cluster; item < (uint32_t)p.n_items; item += nclusters, ++it) {
        const int mt = item % p.n_mtiles;
        const int bh = item / p.n_mtiles;
        const int h = bh % p.heads_q, b = bh / p.heads_q;
        const int hk = h / group;
        const int q0 = mt * 128;
        const int T = num_kv_tiles(p, q0);
        ptx::mbar_wait(bar(bars.q_empty), (it & 1) ^ 1);
        if (rank == 0) ptx::mbar_expect_tx(bar(bars.q_full), 2 * Cfg::Q_BYTES);
#pragma unroll
        for (int jb = 0; jb < NB; ++jb)
          ptx::tma_load_4d_2sm(sQ + jb * 8192, &map_q, l_q_full, jb * 128, q0 + 64 * (int)rank, h, b);
        for (int step = 0; step < T + LA; ++step) {
          if (step < T) {
            const int kv0 = step * 128;
#pragma unroll
            for (int ks = 0; ks < Cfg::KST; ++ks) {
              const uint32_t stage = kc % Cfg::NKS, n = kc / Cfg::NKS;
              ptx::mbar_wait(bar(bars.k_empty[stage]), (n & 1) ^ 1);
              const int nb = (NB - 2 * ks) >= 2 ? 2 : 1;
              if (rank == 0) ptx::mbar_expect_tx(bar(bars.k_full[stage]), 2 * nb * 8192);
              const uint32_t l_full = ptx::mapa(bar(bars.k_full[stage]), 0);
              for (int bx = 0; bx < nb; ++bx)
                ptx::tma_load_4d_2sm(sK + stage * 16384 + bx * 8192, &map_k, l_full, (2 * ks + bx) * 128,
                                     kv0 + 64 * (int)rank, hk, b);
              ++kc;
            }
          }
          if (step >= LA) {
            const int kv0 = (step - LA) * 128;
#pragma unroll
            for (int s = 0; s < Cfg::NSLICE; ++s) {
              const uint32_t stage = vc % Cfg::NVS, n = vc / Cfg::NVS;
              ptx::mbar_wait(bar(bars.v_empty[stage]), (n & 1) ^ 1);
              if (rank == 0) ptx::mbar_expect_tx(bar(bars.v_full[stage]), 2 * 16384);
              const uint32_t l_full = ptx::mapa(bar(bars.v_full[stage]), 0);
              ptx::tma_load_4d_2sm(sV + stage * 16384, &map_v, l_full, 256 * s + 128 * (int)rank, kv0, hk, b);
              ++vc;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // =========================================== QK MMA issuer (leader CTA) =====================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc_qk = ptx::make_idesc(0, 0, 0, 0, 128, 128);   // e4m3 x e4m3, K-major both
      uint32_t kc = 0, it = 0, g = 0;
      for (uint32_t item = cluster; item < (uint32_t)p.n_items; item += nclusters, ++it) {
        const int mt = item % p.n_mtiles;
        const int T = num_kv_tiles(p, mt * 128);
        ptx::mbar_wait(bar(bars.q_full), it & 1);
        ptx::tc_fence_after();
        for (int step = 0; step < T; ++step, ++g) {
          const uint32_t sbuf = g % KS;
          const uint32_t d_tmem = tmem + Cfg::S_BASE + 64 * sbuf;
          // S stage reuse: the softmax of tile g - KS has read it (and written its P) once p_full[sbuf] completed
          if (g >= KS) {
            ptx::mbar_wait_cluster(bar(bars.p_full[sbuf]), ((g - KS) / KS) & 1);
            ptx::tc_fence_after();
          }
#pragma unroll
          for (int ks = 0; ks < Cfg::KST; ++ks) {
            const uint32_t stage = kc % Cfg::NKS, n = kc / Cfg::NKS;
            ptx::mbar_wait(bar(bars.k_full[stage]), n & 1);
            ptx::tc_fence_after();
            const int nb = (NB - 2 * ks) >= 2 ? 2 : 1;
            for (int bx = 0; bx < nb; ++bx) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {  // 32 e4m3 = 32 bytes per MMA K step
                const uint64_t ad = ptx::make_smem_desc_sw128(sQ + (2 * ks + bx) * 8192 + k4 * 32, 16, 1024);
                const uint64_t bd = ptx::make_smem_desc_sw128(sK + stage * 16384 + bx * 8192 + k4 * 32, 16, 1024);
                ptx::umma_f8_ss<CG>(d_tmem, ad, bd, idesc_qk, (ks | bx | k4) != 0 ? 1u : 0u);
              }
            }
            ptx::umma_commit_mc<CG>(bar(bars.k_empty[stage]), 0x3);
            ++kc;
          }
          ptx::umma_commit_mc<CG>(bar(bars.s_full[sbuf]), 0x3);
          if (step == T - 1) ptx::umma_commit_mc<CG>(bar(bars.q_empty), 0x3);
        }
      }
    }
    __syncwarp();
  } else if (warp == kPvWarp) {
    // =========================================== PV MMA issuer (leader CTA) =====================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc_pv = ptx::make_idesc(0, 0, 0, 1, 128, 256);   // B = V, MN-major
      uint32_t vc = 0, gp = 0;
      for (uint32_t item = cluster; item < (uint32_t)p.n_items; item += nclusters) {
        const int mt = item % p.n_mtiles;
        const int T = num_kv_tiles(p, mt * 128);
        for (int step = 0; step < T; ++step, ++gp) {
          const uint32_t pbuf = gp % KS;
          ptx::mbar_wait_cluster(bar(bars.p_full[pbuf]), (gp / KS) & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int s = 0; s < Cfg::NSLICE; ++s) {
            const uint32_t stage = vc % Cfg::NVS, n = vc / Cfg::NVS;
            ptx::mbar_wait(bar(bars.v_full[stage]), n & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 32 keys per MMA: 4 atoms of 8 key-lines
              const uint64_t ad = ptx::make_smem_desc_sw128(sP + pbuf * 8192 + kk * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(sV + stage * 16384 + kk * 4096, 16384, 1024);
              ptx::umma_f8_ss<CG>(tmem + 128 * s, ad, bd, idesc_pv, (step > 0 || kk > 0) ? 1u : 0u);
            }
            ptx::umma_commit_mc<CG>(bar(bars.v_empty[stage]), 0x3);
            ++vc;
          }
          ptx::umma_commit_mc<CG>(bar(bars.p_empty[pbuf]), 0x3);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== softmax / correction / epilogue ================
    // NWG warpgroups work on DIFFERENT KV tiles (tile i -> warpgroup i % NWG), each thread owning one TMEM lane
    // (row, 64-key half) of its tile. Tiles i .. i+NWG-1 are therefore in different phases of
    // load / max / exp2 / pack, so the MUFU pipe of one overlaps the ALU work and the waits of the others (with
    // column-split warpgroups all sit in the same phase of the same tile, see
    // profiles/r01_fwd_small_d_ncu.md). The only cross-warpgroup dependency is the running row max,
    // published through SMEM + an mbarrier right after a tile's max is known.
    // The scores are read from TMEM twice -- once for the max, once (in two 32-column halves) for exp2 / pack -- so a
    // thread never holds more than 32 of them: that is what lets 4 warpgroups fit the register file.
    const uint32_t t = threadIdx.x;
    const uint32_t wg = t >> 7;               // tiles i with i % NWG == wg
    const uint32_t lane128 = t & 127;
    const uint32_t row = lane128 & 63;
    const uint32_t kh = lane128 >> 6;         // 64-key half of the KV tile / column half of O
    const uint32_t rgrp = warp & 1;           // rows 0-31 / 32-63
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t l_p_full0 = ptx::mapa(bar(bars.p_full[0]), 0);
    const float NEG_INF = -INFINITY;
    uint32_t g = 0;                           // global tile counter at the start of the item
    uint32_t pub[4] = {0, 0, 0, 0};           // completed publications of m_full[0..3]
    uint32_t uw = 0;                          // tiles processed by this warpgroup (exchange buffer parity)
    for (uint32_t item = cluster; item < (uint32_t)p.n_items; item += nclusters) {
      const int mt = item % p.n_mtiles;
      const int bh = item / p.n_mtiles;
      const int h = bh % p.heads_q, b = bh / p.heads_q;
      const int hk = h / group;
      const int q0 = mt * 128;
      const int T = num_kv_tiles(p, q0);
      const int gq = q0 + 64 * (int)rank + (int)row;
      const int causal_lim = gq + (p.seqlen_kv - p.seqlen_q);
      const float qs_c = p.qs[((int64_t)b * p.heads_q + h) * p.tq + mt] * p.scale_log2;
      const float* ksp = p.ks + ((int64_t)b * p.heads_kv + hk) * p.tk;
      const float* vsp = p.vs + ((int64_t)b * p.heads_kv + hk) * p.tk;
      const float vref = p.vref[(int64_t)b * p.heads_kv + hk];
      const float pc_base = vref > 0.f ? kPScale / vref : 0.f;
      float m_l = NEG_INF, l = 0.f;           // partial row sum of this thread, relative to m_l

      for (int i = (int)wg; i < T; i += NWG, ++uw) {
        const uint32_t gi = g + (uint32_t)i;
        const uint32_t sbuf = gi % KS;
        const uint32_t xb = uw & 1;
        const float mul = qs_c * __ldg(ksp + i);
        const float pc = pc_base * __ldg(vsp + i);
        const uint32_t s_addr = tmem + lane_base + Cfg::S_BASE + 64 * sbuf;
        const int key0 = i * 128 + 64 * (int)kh;
        const bool tail = (i * 128 + 128 > p.seqlen_kv);
        const bool diag = p.causal && (i * 128 + 127 > q0 + (p.seqlen_kv - p.seqlen_q));
        const bool masked = tail || diag;
        const int lim = p.causal ? (causal_lim < p.seqlen_kv - 1 ? causal_lim : p.seqlen_kv - 1) : p.seqlen_kv - 1;
        ptx::mbar_wait(bar(bars.s_full[sbuf]), (gi / KS) & 1);
        ptx::tc_fence_after();
        // ---- pass 1: row max of this thread's 64 raw scores, 32 at a time
        float tmax = NEG_INF;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sr[32];
          ptx::tmem_ld_x32(s_addr + 32 * half, sr);
          ptx::tmem_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(sr[j]);
          if (masked) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (key0 + 32 * half + j > lim) x[j] = NEG_INF;
          }
          float mx0 = fmax3(x[0], x[1], x[2]), mx1 = fmax3(x[3], x[4], x[5]);
          float mx2 = fmax3(x[6], x[7], x[8]), mx3 = fmax3(x[9], x[10], x[11]);
          mx0 = fmax3(mx0, x[12], x[13]); mx1 = fmax3(mx1, x[14], x[15]);
          mx2 = fmax3(mx2, x[16], x[17]); mx3 = fmax3(mx3, x[18], x[19]);
          mx0 = fmax3(mx0, x[20], x[21]); mx1 = fmax3(mx1, x[22], x[23]);
          mx2 = fmax3(mx2, x[24], x[25]); mx3 = fmax3(mx3, x[26], x[27]);
          mx0 = fmax3(mx0, x[28], x[29]); mx1 = fmax3(mx1, x[30], x[31]);
          tmax = fmaxf(tmax, fmaxf(fmax3(mx0, mx1, mx2), mx3));
        }
        tmax *= mul;   // scaled (log2) domain
        xchw[wg][xb][kh][row] = tmax;
        ptx::named_bar_sync(1 + 2 * wg + rgrp, 64);
        tmax = fmaxf(tmax, xchw[wg][xb][kh ^ 1][row]);
        // running max of tile i-1, published by the previous warpgroup
        float m_prev = NEG_INF;
        if (i > 0) {
          const uint32_t par = (uint32_t)(i - 1) & 3u;
          const uint32_t cnt = pub[par] + (uint32_t)((i - 1) >> 2);
          ptx::mbar_wait(bar(bars.m_full[par]), cnt & 1);
          m_prev = mval[par][row];
        }
        const float m_new = fmaxf(m_prev, tmax);
        const bool upd = (m_new - m_prev) > kLazyThreshold;
        const float m_use = upd ? m_new : m_prev;
        const bool need_rescale = upd && (m_prev != NEG_INF);
        if (kh == 0) mval[i & 3][row] = m_use;
        __syncwarp();
        if (kh == 0 && ptx::lane_id() == 0) ptx::mbar_arrive(bar(bars.m_full[i & 3]));
        const float m_safe = (m_use == NEG_INF) ? 0.f : m_use;
        float factor = 1.f;
        if (need_rescale) factor = exp2f(m_prev - m_use);
        const float cadd = (pc > 0.f ? log2f(pc) : 0.f) - m_safe;
        const float2 mul2 = make_float2(mul, mul), c2 = make_float2(cadd, cadd);
        // ---- pass 2: exp2 / row sum / e4m3 pack, 32 scores at a time, straight into the P tile
        ptx::mbar_wait(bar(bars.p_empty[sbuf]), ((gi / KS) & 1) ^ 1);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        // P8 tile: one 128-byte row per query row; this thread owns bytes [64 kh, 64 kh + 64)
        const uint32_t prow = sP + sbuf * 8192 + row * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t sr[32];
          ptx::tmem_ld_x32(s_addr + 32 * half, sr);
          ptx::tmem_wait_ld();
          float x[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(sr[j]);
          if (masked) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (key0 + 32 * half + j > lim) x[j] = NEG_INF;
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float2 a0 = ffma2(make_float2(x[j], x[j + 1]), mul2, c2);
            const float2 a1 = ffma2(make_float2(x[j + 2], x[j + 3]), mul2, c2);
            const float2 e0 = make_float2(exp2f(a0.x), exp2f(a0.y));
            const float2 e1 = make_float2(exp2f(a1.x), exp2f(a1.y));
            acc0 = fadd2(acc0, e0);
            acc1 = fadd2(acc1, e1);
            pk[j >> 2] = pack_e4m3x4(e0.x, e0.y, e1.x, e1.y);
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t addr = prow + (((4 * kh + 2 * half + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
          }
        }
        acc0 = fadd2(acc0, acc1);
        const float lsum = (acc0.x + acc0.y) * (pc > 0.f ? 1.f / pc : 0.f);
        // partial row sum, re-expressed relative to the max in effect
        if (m_l != m_use) l *= (m_l == NEG_INF) ? 0.f : exp2f(m_l - m_use);
        l += lsum;
        m_l = m_use;

        if (__any_sync(0xffffffffu, need_rescale)) {
          // O may only be touched once PV of tile i-1 has retired; this warpgroup scales all columns
          ptx::mbar_wait(bar(bars.p_empty[(gi - 1) % KS]), ((gi - 1) / KS) & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < Cfg::O_COLS; c0 += 32) {
            uint32_t orr[32];
            ptx::tmem_ld_x32(tmem + lane_base + c0, orr);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * factor);
            ptx::tmem_st_x32(tmem + lane_base + c0, orr);
          }
          ptx::tmem_wait_st();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::lane_id() == 0) ptx::mbar_arrive_cluster(l_p_full0 + 8u * sbuf);
      }

      // ---------------- epilogue ----------------
      {
        // final running max: published with the last tile
        const uint32_t parl = (uint32_t)(T - 1) & 3u;
        const uint32_t cntl = pub[parl] + (uint32_t)((T - 1) >> 2);
        ptx::mbar_wait(bar(bars.m_full[parl]), cntl & 1);
        const float m_fin = mval[parl][row];
#pragma unroll
        for (int j = 0; j < 4; ++j) pub[j] += (T > j) ? (uint32_t)((T - 1 - j) / 4 + 1) : 0u;
        const float l_fin = (m_l == NEG_INF) ? 0.f : l * exp2f(m_l - m_fin);
        xl[wg * 2 + kh][row] = l_fin;
        ptx::named_bar_sync(1 + 2 * NWG + rgrp, 64 * NWG);   // all warpgroups: nobody enters the next item before m_fin is read
        float l_tot = (xl[0][row] + xl[1][row]) + (xl[2][row] + xl[3][row]);
        if constexpr (NWG == 4) l_tot += (xl[4][row] + xl[5][row]) + (xl[6][row] + xl[7][row]);
        const uint32_t gl = g + (uint32_t)T - 1;
        ptx::mbar_wait(bar(bars.p_empty[gl % KS]), (gl / KS) & 1);
        ptx::tc_fence_after();
        const float inv = l_tot > 0.f ? (vref / kPScale) / l_tot : 0.f;
        const bool row_ok = gq < p.seqlen_q;
        // smooth-V: the kernel saw V - mean_seq(V); rows of P sum to 1, so O = P (V - mean) + mean
        // (reference knob fp8_smooth_v, functional.py:247). Rows without a visible key stay 0.
        const float* vmean = (p.vsum != nullptr && l_tot > 0.f) ? p.vsum + ((int64_t)b * p.heads_kv + h / (p.heads_q / p.heads_kv)) * p.head_dim : nullptr;
        const float inv_nkv = 1.f / (float)p.seqlen_kv;
        // per-channel V scales (reference knob fp8_v_quant_method="per_channel"): V8[k, d] = V[k, d] / cs[d], block scales
        // are 1, so O[d] is scaled by cs[d] = amax[d] / 448 here
        const float* vcs = p.vamax != nullptr ? p.vamax + ((int64_t)b * p.heads_kv + h / (p.heads_q / p.heads_kv)) * p.head_dim : nullptr;
        uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                        2 * ((int64_t)b * p.o_stride[0] + (int64_t)h * p.o_stride[1] + (int64_t)gq * p.o_stride[2]);
#pragma unroll
        for (int s = 0; s < Cfg::NSLICE; ++s) {
#pragma unroll 1
          for (int c0 = (int)wg * (128 / NWG); c0 < (int)(wg + 1) * (128 / NWG); c0 += 32) {
            uint32_t orr[32];
            ptx::tmem_ld_x32(tmem + lane_base + 128 * s + c0, orr);
            ptx::tmem_wait_ld();
            const int d0 = 256 * s + 128 * (int)kh + c0;
            if (row_ok) {
#pragma unroll
              for (int v = 0; v < 4; ++v) {
                const int d = d0 + 8 * v;
                if (d < p.head_dim) {
                  uint32_t w[4];
#pragma unroll
                  for (int u = 0; u < 4; ++u) {
                    float a = __uint_as_float(orr[8 * v + 2 * u]) * inv;
                    float c = __uint_as_float(orr[8 * v + 2 * u + 1]) * inv;
                    if (vcs != nullptr) {
                      a *= fmaxf(__ldg(vcs + d + 2 * u), 1e-12f) * (1.f / 448.f);
                      c *= fmaxf(__ldg(vcs + d + 2 * u + 1), 1e-12f) * (1.f / 448.f);
                    }
                    if (vmean != nullptr) { a = fmaf(__ldg(vmean + d + 2 * u), inv_nkv, a); c = fmaf(__ldg(vmean + d + 2 * u + 1), inv_nkv, c); }
                    w[u] = OUT_BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
                  }
                  *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(w[0], w[1], w[2], w[3]);
                }
              }
            }
          }
        }
        if (p.lse != nullptr && wg == 0 && kh == 0 && row_ok) {
          // smooth-K: the kernel saw S' = Q (K - mean)^T, i.e. S shifted by the row constant q . mean;
          // softmax and O are invariant, the LSE is shifted back (reference: cute/launch.cuh:602-635)
          const int64_t ridx = ((int64_t)b * p.heads_q + h) * p.seqlen_q + gq;
          const float corr = p.qkm != nullptr ? p.qkm[ridx] * (p.scale_log2 * 0.6931471805599453f) : 0.f;
          const float lse = (l_tot > 0.f) ? (m_fin + log2f(l_tot)) * 0.6931471805599453f + corr : NEG_INF;
          p.lse[((int64_t)b * p.heads_q + h) * p.lse_bh_stride + gq] = lse;
        }
        ptx::tc_fence_before();
        // xl is reused by the next item: all four writers must be past their reads first
        ptx::named_bar_sync(1 + 2 * NWG + rgrp, 64 * NWG);
        g += (uint32_t)T;
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == kMmaWarp) ptx::tmem_dealloc<CG>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// fused pre-pass: per-(b, h, 128-row block) amax -> scale = amax / 448, e4m3 rows padded to dpad
// (multiple of 16 bytes), scales, and the per-(b, h) maximum V scale.
// One 256-thread block per (tensor, b, h, block). (reference: quantize_fp8.cuh:67-168)
// ------------------------------------------------------------------------------------------------
struct QuantArgs {
  const void* src[3];      // Q, K, V (16-bit)
  uint8_t* dst[3];         // Q8, K8, V8  [B, H, N, dpad]
  float* scale[3];         // [B, H, T]
  float* vref;             // [B, Hkv]
  int64_t stride[3][3];    // element strides (b, h, n) of the sources
  int heads[3], seqlen[3], tiles[3];
  int64_t first_block[4];  // prefix sums of blocks per tensor
  int batch, head_dim, dpad;
  const float* ksum;       // [B, Hkv, D] column sums of K over the sequence (smooth-K) or nullptr
  const float* vsum;       // [B, Hkv, D] column sums of V over the sequence (smooth-V) or nullptr
  const float* vamax;      // [B, Hkv, D] per-channel maxima of |V - mean| (per-channel V quantisation) or nullptr
  float* qkm;              // [B, Hq, Nq] q . mean_seq(K), written by the Q blocks under smooth-K, or nullptr
};

// One 256-thread block per (tensor, b, h, 128-row block). The tile (<= 128 x 512 x 2 B) is read TWICE -- once for
// the amax, once for the conversion -- instead of being parked in registers in between: the second read hits L2 (a
// block re-reads the 64-128 KB it has just pulled in; all resident blocks together hold a few tens of MB of the
// 126 MB L2), HBM traffic stays one read + one write per element, and without a 32-register cache several blocks fit
// one SM, so the loads of one overlap the reduction / stores of the others. The register-cached 1024-thread version
// (one block per SM, load -> reduce -> store strictly in sequence) ran at 2.1 TB/s = 31 % of the HBM peak at C4
// (profiles/r02_fp8_c4_ncu.md).
// For Q under smooth-K the block also emits qkm[row] = q_row . mean_seq(K) (the LSE correction), which used to be a
// separate pass over Q.
constexpr int kQuantThreads = 256;
constexpr int kQuantUnroll = 8;   // 16-byte loads in flight per thread
template <bool BF16>
__global__ void __launch_bounds__(kQuantThreads, 4) quantize_e4m3_kernel(const QuantArgs a) {
  __shared__ float red[kQuantThreads / 32];
  __shared__ float rowdot[128];
  const int64_t blk = blockIdx.x;
  const int which = blk >= a.first_block[2] ? 2 : (blk >= a.first_block[1] ? 1 : 0);
  const int64_t local = blk - a.first_block[which];
  const int T = a.tiles[which], H = a.heads[which], N = a.seqlen[which];
  const int tile = (int)(local % T);
  const int h = (int)((local / T) % H);
  const int b = (int)(local / ((int64_t)T * H));
  const int D = a.head_dim, dpad = a.dpad;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(a.src[which]) +
                       2 * ((int64_t)b * a.stride[which][0] + (int64_t)h * a.stride[which][1]);
  const int64_t rs = a.stride[which][2];
  const int r0 = tile * 128;
  const int rows = (N - r0) < 128 ? (N - r0) : 128;
  const int vec_per_row = D / 8;               // 8 elements (16 B) per vector; D % 8 == 0
  const int nvec = rows * vec_per_row;         // <= 128 * 64
  // smooth-K / smooth-V: quantise X - mean_seq(X) (per (b, h) and channel)
  const float* km = (which == 1 && a.ksum != nullptr) ? a.ksum + ((int64_t)b * H + h) * D
                  : (which == 2 && a.vsum != nullptr) ? a.vsum + ((int64_t)b * H + h) * D : nullptr;
  const float inv_n = 1.f / (float)N;
  // Q under smooth-K: dot of every query row with mean_seq(K) of its KV head
  const float* qk_mean = (which == 0 && a.ksum != nullptr && a.qkm != nullptr)
                             ? a.ksum + ((int64_t)b * a.heads[1] + h / (a.heads[0] / a.heads[1])) * D : nullptr;
  if (qk_mean != nullptr && threadIdx.x < 128) rowdot[threadIdx.x] = 0.f;
  auto expand = [&](const uint4& raw, int c, float* f) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (BF16) { f[2 * u] = __uint_as_float(w[u] << 16); f[2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u); }
      else { const __half2 hh = *reinterpret_cast<const __half2*>(&w[u]); f[2 * u] = __low2float(hh); f[2 * u + 1] = __high2float(hh); }
    }
    if (km != nullptr) {
      const float4 k0 = __ldg(reinterpret_cast<const float4*>(km + 8 * c)), k1 = __ldg(reinterpret_cast<const float4*>(km + 8 * c + 4));
      f[0] -= k0.x * inv_n; f[1] -= k0.y * inv_n; f[2] -= k0.z * inv_n; f[3] -= k0.w * inv_n;
      f[4] -= k1.x * inv_n; f[5] -= k1.y * inv_n; f[6] -= k1.z * inv_n; f[7] -= k1.w * inv_n;
    }
  };
  // vector i of the tile -> (row, 8-element column group). When the block's 256 threads cover whole rows
  // (vec_per_row divides 256: head dims 64 / 128 / 256 / 512) a thread keeps its column and steps rows by a constant:
  // no integer division in the loops.
  const bool aligned = (kQuantThreads % vec_per_row) == 0;
  const int t_r = (int)threadIdx.x / vec_per_row, t_c = (int)threadIdx.x - t_r * vec_per_row;
  const int r_step = aligned ? kQuantThreads / vec_per_row : 0;
  auto rc_of = [&](int i, int step, int& r, int& c) {
    if (aligned) { r = t_r + step * r_step; c = t_c; }
    else { r = i / vec_per_row; c = i - r * vec_per_row; }
  };
  auto load = [&](int r, int c) {
    return __ldg(reinterpret_cast<const uint4*>(src + 2 * ((int64_t)(r0 + r) * rs + 8 * c)));
  };
  // |x| maximum of 8 packed 16-bit floats without unpacking: clear the sign bits, then compare as unsigned 16-bit
  // integers (non-negative IEEE values order like their bit patterns); exact, so the scale equals the fp32 path's
  auto amax_bits = [&](const uint4& raw, uint32_t acc) {
    acc = __vmaxu2(acc, raw.x & 0x7fff7fffu); acc = __vmaxu2(acc, raw.y & 0x7fff7fffu);
    acc = __vmaxu2(acc, raw.z & 0x7fff7fffu); acc = __vmaxu2(acc, raw.w & 0x7fff7fffu);
    return acc;
  };
  // ---- pass 1: amax of the tile
  float amax = 0.f;
  uint32_t abits = 0u;   // packed 16-bit running maxima (used when no mean is subtracted)
  for (int i0 = threadIdx.x, st = 0; i0 < nvec; i0 += kQuantThreads * kQuantUnroll, st += kQuantUnroll) {
    uint4 buf[kQuantUnroll];
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const int i = i0 + u * kQuantThreads;
      int r, c;
      rc_of(i, st + u, r, c);
      if (i < nvec) buf[u] = load(r, c);
    }
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const int i = i0 + u * kQuantThreads;
      if (i < nvec) {
        if (km == nullptr) {
          abits = amax_bits(buf[u], abits);
        } else {
          int r, c;
          rc_of(i, st + u, r, c);
          float f[8];
          expand(buf[u], c, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(f[e]));
        }
      }
    }
  }
  if (km == nullptr) {
    const uint32_t m16 = max(abits & 0xffffu, abits >> 16);
    amax = BF16 ? __uint_as_float(m16 << 16) : __half2float(__ushort_as_half((unsigned short)m16));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = red[0];
#pragma unroll
  for (int i = 1; i < kQuantThreads / 32; ++i) amax = fmaxf(amax, red[i]);
  // per-channel V: every channel has its own scale (applied in the attention epilogue), block scale = 1
  const float* cam = (which == 2 && a.vamax != nullptr) ? a.vamax + ((int64_t)b * H + h) * D : nullptr;
  // IEEE division (not the --use_fast_math reciprocal): s = amax / 448 and 1 / s are what the reference's quantiser
  // computes (/root/reference/csrc/cuffpa/cute/fp8/quantize_fp8.cuh:140-144), so the e4m3 tiles match it bit for bit
  const float scale = cam != nullptr ? 1.f : __fdiv_rn(fmaxf(amax, 1e-12f), 448.f);
  const float inv = __fdiv_rn(1.f, scale);
  if (threadIdx.x == 0) {
    a.scale[which][((int64_t)b * H + h) * T + tile] = scale;
    if (which == 2) atomicMax(reinterpret_cast<unsigned int*>(a.vref + (int64_t)b * H + h), __float_as_uint(scale));
  }
  // ---- pass 2: the same vectors again (L2), converted and stored
  uint8_t* dst = a.dst[which] + (((int64_t)b * H + h) * N + r0) * dpad;
  for (int i0 = threadIdx.x, st = 0; i0 < nvec; i0 += kQuantThreads * kQuantUnroll, st += kQuantUnroll) {
    uint4 buf[kQuantUnroll];
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const int i = i0 + u * kQuantThreads;
      int r, c;
      rc_of(i, st + u, r, c);
      if (i < nvec) buf[u] = load(r, c);
    }
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      const int i = i0 + u * kQuantThreads;
      int r, c;
      rc_of(i, st + u, r, c);
      float qd = 0.f;   // this vector's share of q_row . mean_seq(K)
      if (i < nvec) {
        float f[8];
        expand(buf[u], c, f);
        if (qk_mean != nullptr) {
          const float4 k0 = __ldg(reinterpret_cast<const float4*>(qk_mean + 8 * c)), k1 = __ldg(reinterpret_cast<const float4*>(qk_mean + 8 * c + 4));
          qd = f[0] * k0.x;
          qd = fmaf(f[1], k0.y, qd); qd = fmaf(f[2], k0.z, qd); qd = fmaf(f[3], k0.w, qd);
          qd = fmaf(f[4], k1.x, qd); qd = fmaf(f[5], k1.y, qd); qd = fmaf(f[6], k1.z, qd); qd = fmaf(f[7], k1.w, qd);
        }
        if (cam != nullptr) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] *= 448.f / fmaxf(__ldg(cam + 8 * c + e), 1e-12f);
        }
        const float2 inv2 = make_float2(inv, inv), z2 = make_float2(-0.f, -0.f);   // x * inv + (-0) keeps the sign of a zero product
        const float2 g0 = ffma2(make_float2(f[0], f[1]), inv2, z2), g1 = ffma2(make_float2(f[2], f[3]), inv2, z2);
        const float2 g2 = ffma2(make_float2(f[4], f[5]), inv2, z2), g3 = ffma2(make_float2(f[6], f[7]), inv2, z2);
        uint2 o;
        o.x = pack_e4m3x4(g0.x, g0.y, g1.x, g1.y);
        o.y = pack_e4m3x4(g2.x, g2.y, g3.x, g3.y);
        *reinterpret_cast<uint2*>(dst + (int64_t)r * dpad + 8 * c) = o;
      }
      if (qk_mean != nullptr) {   // block-uniform branch: every lane takes part in the shuffles
        if (aligned) {
          // the lanes of a row segment (min(32, vec_per_row) consecutive lanes) hold pieces of the same row: butterfly
          // sum inside the segment, one shared-memory add per segment (a per-lane atomicAdd on one address is a
          // 32-way serialised CAS loop: it took more time than the HBM traffic of the whole kernel)
          const int w = vec_per_row < 32 ? vec_per_row : 32;
          for (int sft = w >> 1; sft > 0; sft >>= 1) qd += __shfl_xor_sync(0xffffffffu, qd, sft);
          if (i < nvec && ((int)(threadIdx.x & 31) & (w - 1)) == 0) atomicAdd(&rowdot[r], qd);
        } else if (i < nvec) {
          atomicAdd(&rowdot[r], qd);
        }
      }
    }
  }
  if (dpad > D) {  // zero the padding bytes [D, dpad) (dpad - D is 0 or 8)
    for (int r = threadIdx.x; r < rows; r += kQuantThreads)
      *reinterpret_cast<uint2*>(dst + (int64_t)r * dpad + D) = make_uint2(0u, 0u);
  }
  if (qk_mean != nullptr) {
    __syncthreads();
    if ((int)threadIdx.x < rows)
      a.qkm[((int64_t)b * H + h) * N + r0 + threadIdx.x] = rowdot[threadIdx.x] / (float)a.seqlen[1];
  }
}

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (BF16) { f[2 * u] = __uint_as_float(w[u] << 16); f[2 * u + 1] = __uint_as_float(w[u] & 0xffff0000u); }
    else { const __half2 hh = *reinterpret_cast<const __half2*>(&w[u]); f[2 * u] = __low2float(hh); f[2 * u + 1] = __high2float(hh); }
  }
}

// column sums of K over a chunk of rows: warp w takes rows r0+w, r0+w+8, ...; lane l the 8 channels
// [256 j + 8 l, +8) (16-byte loads); warps are combined through shared memory, chunks through atomics.
template <bool BF16>
__global__ void __launch_bounds__(256) k_colsum_kernel(const void* __restrict__ k, float* __restrict__ ksum, int64_t s0,
                                                       int64_t s1, int64_t s2, int H, int N, int D, int rows_per_block) {
  __shared__ float part[8][256];
  const int nchunk = (N + rows_per_block - 1) / rows_per_block;
  const int chunk = blockIdx.x % nchunk;
  const int h = (blockIdx.x / nchunk) % H;
  const int b = blockIdx.x / (nchunk * H);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(k) + 2 * ((int64_t)b * s0 + (int64_t)h * s1);
  const int r0 = chunk * rows_per_block;
  const int r1 = (r0 + rows_per_block) < N ? (r0 + rows_per_block) : N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d0 = 0; d0 < D; d0 += 256) {
    const int d = d0 + 8 * lane;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (d < D) {
#pragma unroll 4
      for (int r = r0 + warp; r < r1; r += 8) {
        float f[8];
        unpack8<BF16>(*reinterpret_cast<const uint4*>(src + 2 * ((int64_t)r * s2 + d)), f);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += f[u];
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) part[warp][8 * lane + u] = acc[u];
    __syncthreads();
    const int dd = d0 + threadIdx.x;
    if (dd < D) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
      atomicAdd(ksum + ((int64_t)b * H + h) * D + dd, t);
    }
    __syncthreads();
  }
}

// per-channel maxima of |V - mean| over the sequence (same decomposition as k_colsum_kernel; non-negative floats
// order like their bit patterns, so chunks combine with an integer atomicMax)
template <bool BF16>
__global__ void __launch_bounds__(256) v_colamax_kernel(const void* __restrict__ v, const float* __restrict__ vsum,
                                                        float* __restrict__ vamax, int64_t s0, int64_t s1, int64_t s2, int H,
                                                        int N, int D, int rows_per_block) {
  __shared__ float part[8][256];
  const int nchunk = (N + rows_per_block - 1) / rows_per_block;
  const int chunk = blockIdx.x % nchunk;
  const int h = (blockIdx.x / nchunk) % H;
  const int b = blockIdx.x / (nchunk * H);
  const uint8_t* src = reinterpret_cast<const uint8_t*>(v) + 2 * ((int64_t)b * s0 + (int64_t)h * s1);
  const float* mean = vsum != nullptr ? vsum + ((int64_t)b * H + h) * D : nullptr;
  const float inv_n = 1.f / (float)N;
  const int r0 = chunk * rows_per_block;
  const int r1 = (r0 + rows_per_block) < N ? (r0 + rows_per_block) : N;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d0 = 0; d0 < D; d0 += 256) {
    const int d = d0 + 8 * lane;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (d < D) {
      float mu[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (mean != nullptr) {
#pragma unroll
        for (int u = 0; u < 8; ++u) mu[u] = mean[d + u] * inv_n;
      }
#pragma unroll 4
      for (int r = r0 + warp; r < r1; r += 8) {
        float f[8];
        unpack8<BF16>(*reinterpret_cast<const uint4*>(src + 2 * ((int64_t)r * s2 + d)), f);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaxf(acc[u], fabsf(f[u] - mu[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) part[warp][8 * lane + u] = acc[u];
    __syncthreads();
    const int dd = d0 + threadIdx.x;
    if (dd < D) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t = fmaxf(t, part[w][threadIdx.x]);
      atomicMax(reinterpret_cast<unsigned int*>(vamax + ((int64_t)b * H + h) * D + dd), __float_as_uint(t));
    }
    __syncthreads();
  }
}

template <int NB, bool OUT_BF16, int NWG>
int launch_fp8_variant(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                              const Fp8KernelParams& kp, int nclusters, cudaStream_t stream) {
  using Cfg = Fp8Cfg<NB>;
  auto kern = ffpa_fwd_fp8_kernel<NB, OUT_BF16, NWG>;
  // the opt-in shared-memory size is a per-device function attribute
  static bool attr_set[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  dev_id = (dev_id >= 0 && dev_id < 64) ? dev_id : 0;
  if (!attr_set[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_DYN);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaFuncSetAttribute(fp8 smem=%d): %s", Cfg::SMEM_DYN, cudaGetErrorString(e));
    attr_set[dev_id] = true;
  }
  kern<<<dim3(2 * nclusters), dim3((4 * NWG + 3) * 32), Cfg::SMEM_DYN, stream>>>(mq, mk, mv, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "fp8 forward launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

}  // namespace fp8
}  // namespace ffpa
