// B200 (sm_100a) attention backward: three tcgen05 kernels built from one skeleton.
//
// Math (reference: /root/reference/src/ffpa_attn/triton/_ffpa_bwd.py:236-306, 692-855; the split
// into dQ / dK / dV launches follows what the reference's SM100 backend does,
// /root/reference/src/ffpa_attn/cute/_ffpa_bwd_sm100.py:301-484):
//   delta = rowsum(dO * O);  P = exp(scale*S - LSE);  dP = dO V^T;  dS = P * (dP - delta)
//   dQ = scale * dS K;  dK = scale * dS^T Q;  dV = P^T dO;  GQA: dK/dV sum over the head group.
//
// Skeleton ("stationary rows x streamed column tiles", 2-CTA cluster owns 128 stationary rows):
//   GEMM1  S  = A1 * B1^T          A1 resident (K-major), B1 streamed (K-major)      -> TMEM
//   GEMM2  dP = A2 * B2^T          (dQ / dK kinds only)                              -> TMEM
//   elementwise warps:  T = f(S, dP, stats)  -> bf16/fp16 -> SMEM (K-major A operand)
//   GEMM3  ACC += T * B3           B3 streamed as an MN-major operand (row-major [cols x d])
//   kind dQ: rows = queries, cols = keys   A1=Q  B1=K  A2=dO B2=V  T=dS    B3=K
//   kind dK: rows = keys,    cols = queries A1=K  B1=Q  A2=V  B2=dO T=dS^T  B3=Q
//   kind dV: rows = keys,    cols = queries A1=K  B1=Q              T=P^T   B3=dO
// TMEM (D=512): ACC 256 columns (lane folded, 4 N=128 slices), S 2x64, dP 2x64 = 512.
// head_dim in (512, 1024] (LARGE): the accumulator no longer fits next to S / dP, so every item runs
// as two slab passes (each pass = one virtual work item owning half of the output columns and
// recomputing S / dP), and only A1 stays resident in SMEM: A2 is streamed through the ring next to
// B2, one 64-wide head-dim box of each per stage ("Split-D" in its pure form: O(1) SMEM in D).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace bwd {

constexpr int kSoftmaxWarps = 8;
constexpr int kMmaWarp = 8;
constexpr int kTmaWarp = 9;
constexpr int kStoreWarp = 10;   // stash path: TMA-stores the dS tiles from shared memory (dQ kind)
constexpr int kThreads = 352;
constexpr int kSmemLimit = 232448;

constexpr int kKindDQ = 0;
constexpr int kKindDK = 1;
constexpr int kKindDV = 2;


template <int NQK, int KIND>
struct BwdCfg {
  static constexpr bool HAS_DP = (KIND != kKindDV);
  static constexpr int HD = NQK * 64;
  static constexpr int DVP = ((HD + 127) / 128) * 128;
  static constexpr bool LARGE = NQK > 8;                   // head_dim > 512: slab passes + streamed A2
  static constexpr int NPASS = LARGE ? 2 : 1;
  static constexpr int SLAB = LARGE ? ((DVP / 2 + 127) / 128) * 128 : DVP;  // output columns per pass
  static constexpr int NSL = SLAB / 128;                   // N=128 slices of the accumulator
  static constexpr int ACC_COLS = SLAB / 2;
  static constexpr int KST = (NQK + 1) / 2;                // 16 KB K-major stages per streamed tile and operand
  static constexpr int S_BASE = 256, DP_BASE = 384;
  static constexpr int A_BYTES = NQK * 8192;
  static constexpr int NA = (HAS_DP && !LARGE) ? 2 : 1;    // resident operands
  // ring stages one step consumes before the B3 slices (keeps N=256 stage pairs on even indices)
  static constexpr int PRE = KST + (HAS_DP ? (LARGE ? NQK : KST) : 0);
  static constexpr int T_BYTES = 2 * 16384;
  static constexpr int kBudget = kSmemLimit - 3072;
  static constexpr int kRaw = (kBudget - NA * A_BYTES - T_BYTES) / 16384;
  static constexpr int NST = kRaw > 12 ? 12 : kRaw;        // unified ring of 16 KB stages
  static constexpr int SMEM_DYN = NA * A_BYTES + T_BYTES + NST * 16384;
  static_assert(NST >= 3, "not enough shared memory for the streaming ring");
  // WIDE: accumulate with N=256 MMAs (a slice = two consecutive ring stages): T is re-read from SMEM 2x
  // instead of 4x per tile. With N=128 everywhere the operand fetch needs 125 B/clk of the 128 B/clk SMEM
  // port (A 2 KB + B 2 KB per 32-cycle MMA), which is what limits these kernels.
  static constexpr bool WIDE = (NSL % 2 == 0) && (NST % 2 == 0) && (PRE % 2 == 0);
  static_assert(!LARGE || NQK % 2 == 0, "head_dim > 512 is rounded up to a multiple of 128");
  static_assert(ACC_COLS <= 256, "backward kernels support head_dim <= 1024");
};

struct Barriers {
  uint64_t a_full, a_empty;
  uint64_t r_full[12], r_empty[12];
  uint64_t s_full[2];
  uint64_t t_full[2], t_empty[2];
  uint64_t t_written[2], t_stored[2];   // stash path: T tile complete in this CTA / drained by the store warp
};

__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t ctr) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t o0 = c0, o2 = c2;
    c0 = __umulhi(0xCD9E8D57u, o2) ^ c1 ^ k0;
    c2 = __umulhi(0xD2511F53u, o0) ^ c3 ^ k1;
    c1 = 0xCD9E8D57u * o2;
    c3 = 0xD2511F53u * o0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// streamed-tile range of an item
struct TileRange { int first, count; };

template <int KIND>
__device__ __forceinline__ TileRange col_tiles(int causal, int nq, int nkv, int r0, bool stash = false) {
  if (KIND == kKindDQ) {  // rows = queries starting at r0; columns = KV tiles
    int tc = (nkv + 127) >> 7;
    if (causal) {
      int lim = ((r0 + 127 + (nkv - nq)) >> 7) + 1;
      // stash path: the GEMM-only dK/dV kernel consumes 256-key blocks, so visit KV tiles in pairs (the
      // extra tile is fully masked and stores zeros)
      if (stash) { lim = (lim + 1) & ~1; tc = (tc + 1) & ~1; }
      tc = lim < tc ? lim : tc;
    }
    return {0, tc < 1 ? 1 : tc};
  } else {  // rows = keys starting at r0; columns = query tiles (per head of the group)
    const int tq = (nq + 127) >> 7;
    int first = 0;
    if (causal) {
      const int qmin = r0 - (nkv - nq);  // first query row that sees key r0
      first = qmin > 0 ? (qmin >> 7) : 0;
      if (first > tq) first = tq;
    }
    return {first, tq - first};
  }
}

// One work item = 128 stationary rows x a contiguous chunk of the streamed (head, column-tile)
// sequence. With n_chunks > 1 (few, long items: e.g. GQA + causal dK/dV) partial accumulators are
// added into an fp32 buffer with atomics and converted afterwards.
// Packed variable-length mode (p.cu_q != nullptr): the item carries the lengths and token offsets of its
// sequence; row tiles past the end of a short sequence are empty items (n = 0, nothing stored).
struct Item { int rt, bh, s0, n, pass; TileRange tr; int nq, nkv, qoff, koff, bt; };

template <int KIND>
__device__ __forceinline__ Item decode_item(const BwdKernelParams& p, int item, int n_inner) {
  Item it;
  it.rt = item % p.n_rtiles;
  const int rest = item / p.n_rtiles;
  const int chunk = rest % p.n_chunks;
  const int rest2 = rest / p.n_chunks;
  it.pass = rest2 % p.n_pass;   // output-column slab (head_dim > 512); adjacent in the item order
  it.bh = rest2 / p.n_pass;
  const int b = it.bh / ((KIND == kKindDQ) ? p.heads_q : p.heads_kv);
  if (p.cu_q != nullptr) {
    // offsets are clamped to the packed extents so a malformed cu_seqlens can never address past the tensors
    it.qoff = min(max(__ldg(p.cu_q + b), 0), p.total_q);
    it.nq = max(min(__ldg(p.cu_q + b + 1), p.total_q) - it.qoff, 0);
    it.koff = min(max(__ldg(p.cu_k + b), 0), p.total_k);
    it.nkv = max(min(__ldg(p.cu_k + b + 1), p.total_k) - it.koff, 0);
    it.bt = 0;
    if (it.rt * 128 >= ((KIND == kKindDQ) ? it.nq : it.nkv)) { it.tr = {0, 0}; it.s0 = 0; it.n = 0; return it; }
  } else {
    it.nq = p.seqlen_q; it.nkv = p.seqlen_kv; it.qoff = 0; it.koff = 0; it.bt = b;
  }
  it.tr = col_tiles<KIND>(p.causal, it.nq, it.nkv, it.rt * 128, KIND == kKindDQ && p.stash_ds != nullptr);
  const int tfull = it.tr.count * n_inner;
  if (p.n_chunks == 1) { it.s0 = 0; it.n = tfull; }
  else {
    it.s0 = chunk * p.chunk_len;
    int n = tfull - it.s0;
    n = n < 0 ? 0 : n;
    it.n = n < p.chunk_len ? n : p.chunk_len;
  }
  return it;
}

__device__ __forceinline__ int next_item(const BwdKernelParams& p, uint32_t cluster, uint32_t nclusters, uint32_t k) {
  if (p.sched != nullptr) return (k < (uint32_t)p.sched_stride) ? __ldg(p.sched + (size_t)cluster * p.sched_stride + k) : -1;
  const uint32_t item = cluster + k * nclusters;
  return item < (uint32_t)p.n_items ? (int)item : -1;
}

template <int NQK, bool BF16, int KIND, bool GENERAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
ffpa_bwd_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_a2,
                const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                const __grid_constant__ CUtensorMap map_b3, const __grid_constant__ CUtensorMap map_st,
                const __grid_constant__ CUtensorMap map_sp, const BwdKernelParams p) {
  using Cfg = BwdCfg<NQK, KIND>;
  constexpr int CG = 2;
  constexpr bool HAS_DP = Cfg::HAS_DP;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Barriers bars;
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();
  const uint32_t sA1 = smem_base;
  const uint32_t sA2 = sA1 + Cfg::A_BYTES;                  // only when HAS_DP
  const uint32_t sT = sA1 + Cfg::NA * Cfg::A_BYTES;
  const uint32_t sR = sT + Cfg::T_BYTES;

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t cluster = blockIdx.x >> 1;
  const uint32_t nclusters = gridDim.x >> 1;
  auto bar = [](uint64_t& b) { return ptx::smem_u32(&b); };

  if (threadIdx.x == 0) {
    ptx::mbar_init(bar(bars.a_full), 1);
    ptx::mbar_init(bar(bars.a_empty), 1);
    for (int i = 0; i < 12; ++i) { ptx::mbar_init(bar(bars.r_full[i]), 1); ptx::mbar_init(bar(bars.r_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar(bars.s_full[i]), 1);
      ptx::mbar_init(bar(bars.t_full[i]), 2 * kSoftmaxWarps);
      ptx::mbar_init(bar(bars.t_empty[i]), 1);
      ptx::mbar_init(bar(bars.t_written[i]), kSoftmaxWarps);
      ptx::mbar_init(bar(bars.t_stored[i]), 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kTmaWarp && ptx::elect_one()) {
    ptx::prefetch_tmap(&map_a1); ptx::prefetch_tmap(&map_b1); ptx::prefetch_tmap(&map_b3);
    if (HAS_DP) { ptx::prefetch_tmap(&map_a2); ptx::prefetch_tmap(&map_b2); }
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
  // heads: dQ kind iterates query heads; dK/dV kinds iterate KV heads and loop the group inside
  const int heads_it = (KIND == kKindDQ) ? p.heads_q : p.heads_kv;
  const int n_inner = (KIND == kKindDQ) ? 1 : group;

  if (warp == kTmaWarp) {
    // =========================================== TMA producer ===================================
    if (ptx::elect_one()) {
      uint32_t rc = 0, it = 0;
      const uint32_t l_a_full = ptx::mapa(bar(bars.a_full), 0);
      auto load_kmajor = [&](const CUtensorMap* m, int c_row0, int hh, int bb) {
        // KST stages of [64 rows(this CTA) x 128 d]
#pragma unroll
        for (int ks = 0; ks < Cfg::KST; ++ks) {
          const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
          ptx::mbar_wait(bar(bars.r_empty[stage]), (n & 1) ^ 1);
          const int nb = (NQK - 2 * ks) >= 2 ? 2 : 1;
          if (rank == 0) ptx::mbar_expect_tx(bar(bars.r_full[stage]), 2 * nb * 8192);
          const uint32_t l_full = ptx::mapa(bar(bars.r_full[stage]), 0);
          for (int bx = 0; bx < nb; ++bx)
            ptx::tma_load_4d_2sm(sR + stage * 16384 + bx * 8192, m, l_full, (2 * ks + bx) * 64, c_row0 + 64 * (int)rank, hh, bb);
          ++rc;
        }
      };
      // LARGE: one stage = [A2 box | B2 box] of the same 64-wide head-dim box (64 rows of this CTA each)
      auto load_pair = [&](const CUtensorMap* ma, int a_row0, int ha_, const CUtensorMap* mb, int c_row0, int hh, int bb) {
        for (int jb = 0; jb < NQK; ++jb) {
          const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
          ptx::mbar_wait(bar(bars.r_empty[stage]), (n & 1) ^ 1);
          if (rank == 0) ptx::mbar_expect_tx(bar(bars.r_full[stage]), 2 * 16384);
          const uint32_t l_full = ptx::mapa(bar(bars.r_full[stage]), 0);
          ptx::tma_load_4d_2sm(sR + stage * 16384, ma, l_full, jb * 64, a_row0 + 64 * (int)rank, ha_, bb);
          ptx::tma_load_4d_2sm(sR + stage * 16384 + 8192, mb, l_full, jb * 64, c_row0 + 64 * (int)rank, hh, bb);
          ++rc;
        }
      };
      auto load_mnmajor = [&](const CUtensorMap* m, int c_row0, int hh, int bb, int d_base) {
        // NSL stages of [128 rows x 64 d(this CTA's half of the 128-wide slice)]
#pragma unroll
        for (int s = 0; s < Cfg::NSL; ++s) {
          const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
          ptx::mbar_wait(bar(bars.r_empty[stage]), (n & 1) ^ 1);
          if (rank == 0) ptx::mbar_expect_tx(bar(bars.r_full[stage]), 2 * 16384);
          const uint32_t l_full = ptx::mapa(bar(bars.r_full[stage]), 0);
          // WIDE: stage pair (s & ~1, s | 1) = N=256 slice s/2, this CTA's 128 columns as two 64-wide boxes
          const int dcol = Cfg::WIDE ? 256 * (s >> 1) + 128 * (int)rank + 64 * (s & 1) : 128 * s + 64 * (int)rank;
          ptx::tma_load_4d_2sm(sR + stage * 16384, m, l_full, d_base + dcol, c_row0, hh, bb);
          ++rc;
        }
      };
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const Item itm = decode_item<KIND>(p, item_s, n_inner);
        const int bh = itm.bh;
        const int hs = bh % heads_it, b = itm.bt;  // head of the stationary operand; batch coordinate of the maps
        // token rows of the stationary (r0) and streamed (c_off + tile * 128) operands
        const int r0 = ((KIND == kKindDQ) ? itm.qoff : itm.koff) + itm.rt * 128;
        const int c_off = (KIND == kKindDQ) ? itm.koff : itm.qoff;
        const TileRange tr = itm.tr;
        const int T = itm.n;
        if (T <= 0) continue;
        // stationary operands
        ptx::mbar_wait(bar(bars.a_empty), (it & 1) ^ 1);
        if (rank == 0) ptx::mbar_expect_tx(bar(bars.a_full), 2 * Cfg::NA * Cfg::A_BYTES);
        const int ha = hs;  // A tensors are indexed by their own head (q head for dQ, kv head for dK/dV)
#pragma unroll
        for (int jb = 0; jb < NQK; ++jb) {
          ptx::tma_load_4d_2sm(sA1 + jb * 8192, &map_a1, l_a_full, jb * 64, r0 + 64 * (int)rank, ha, b);
          if (Cfg::NA == 2) ptx::tma_load_4d_2sm(sA2 + jb * 8192, &map_a2, l_a_full, jb * 64, r0 + 64 * (int)rank, ha, b);
        }
        for (int step = 0; step <= T; ++step) {
          if (step < T) {
            const int sa = itm.s0 + step;
            const int gi = sa / tr.count, ci = tr.first + sa % tr.count;
            const int hb = (KIND == kKindDQ) ? hs / group : hs * group + gi;  // head of the streamed operands
            load_kmajor(&map_b1, c_off + ci * 128, hb, b);
            if (HAS_DP) {
              if constexpr (Cfg::LARGE) load_pair(&map_a2, r0, ha, &map_b2, c_off + ci * 128, hb, b);
              else load_kmajor(&map_b2, c_off + ci * 128, hb, b);
            }
          }
          if (step >= 1) {
            const int st = itm.s0 + step - 1;
            const int gi = st / tr.count, ci = tr.first + st % tr.count;
            const int hb = (KIND == kKindDQ) ? hs / group : hs * group + gi;
            load_mnmajor(&map_b3, c_off + ci * 128, hb, b, itm.pass * Cfg::SLAB);
          }
        }
        ++it;
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // =========================================== MMA issuer (leader CTA) ========================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_s = ptx::make_idesc(fmt, fmt, 0, 0, 128, 128);
      constexpr uint32_t idesc_acc = ptx::make_idesc(fmt, fmt, 0, 1, 128, 128);
      uint32_t rc = 0, it = 0, g = 0, gp = 0;
      const bool stash_mode = (KIND == kKindDQ) && p.stash_ds != nullptr;
      auto gemm_kmajor = [&](uint32_t sA, uint32_t d_tmem) {
#pragma unroll
        for (int ks = 0; ks < Cfg::KST; ++ks) {
          const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
          ptx::mbar_wait(bar(bars.r_full[stage]), n & 1);
          ptx::tc_fence_after();
          const int nb = (NQK - 2 * ks) >= 2 ? 2 : 1;
          for (int bx = 0; bx < nb; ++bx) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t ad = ptx::make_smem_desc_sw128(sA + (2 * ks + bx) * 8192 + k4 * 32, 16, 1024);
              const uint64_t bd = ptx::make_smem_desc_sw128(sR + stage * 16384 + bx * 8192 + k4 * 32, 16, 1024);
              ptx::umma_f16_ss<CG>(d_tmem, ad, bd, idesc_s, (ks | bx | k4) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit_mc<CG>(bar(bars.r_empty[stage]), 0x3);
          ++rc;
        }
      };
      auto gemm_pair = [&](uint32_t d_tmem) {  // LARGE: A2 and B2 boxes streamed side by side
        for (int jb = 0; jb < NQK; ++jb) {
          const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
          ptx::mbar_wait(bar(bars.r_full[stage]), n & 1);
          ptx::tc_fence_after();
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t ad = ptx::make_smem_desc_sw128(sR + stage * 16384 + k4 * 32, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc_sw128(sR + stage * 16384 + 8192 + k4 * 32, 16, 1024);
            ptx::umma_f16_ss<CG>(d_tmem, ad, bd, idesc_s, (jb | k4) != 0 ? 1u : 0u);
          }
          ptx::umma_commit_mc<CG>(bar(bars.r_empty[stage]), 0x3);
          ++rc;
        }
      };
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const Item itm = decode_item<KIND>(p, item_s, n_inner);
        const int T = itm.n;
        if (T <= 0) continue;
        ptx::mbar_wait(bar(bars.a_full), it & 1);
        ptx::tc_fence_after();
        for (int step = 0; step <= T; ++step) {
          if (step < T) {
            const uint32_t sbuf = g & 1;
            gemm_kmajor(sA1, tmem + Cfg::S_BASE + 64 * sbuf);
            if (HAS_DP) {
              if constexpr (Cfg::LARGE) gemm_pair(tmem + Cfg::DP_BASE + 64 * sbuf);
              else gemm_kmajor(sA2, tmem + Cfg::DP_BASE + 64 * sbuf);
            }
            ptx::umma_commit_mc<CG>(bar(bars.s_full[sbuf]), 0x3);
            if (step == T - 1) ptx::umma_commit_mc<CG>(bar(bars.a_empty), 0x3);
            ++g;
          }
          if (step >= 1) {
            // stash path: T is single-buffered (buffer 0; buffer 1 stages the P tile for its TMA store)
            const uint32_t tbuf = stash_mode ? 0u : (gp & 1);
            ptx::mbar_wait_cluster(bar(bars.t_full[tbuf]), (stash_mode ? gp : (gp >> 1)) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int s = 0; s < Cfg::NSL; ++s) {
              if constexpr (Cfg::WIDE) {
                if (s & 1) continue;  // handled together with the even stage
                const uint32_t st = rc % Cfg::NST, n = rc / Cfg::NST;
                ptx::mbar_wait(bar(bars.r_full[st]), n & 1);
                ptx::mbar_wait(bar(bars.r_full[st + 1]), n & 1);
                ptx::tc_fence_after();
                constexpr uint32_t idesc_acc2 = ptx::make_idesc(fmt, fmt, 0, 1, 128, 256);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                  const uint64_t ad = ptx::make_smem_desc_sw128(sT + tbuf * 16384 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                  const uint64_t bd = ptx::make_smem_desc_sw128(sR + st * 16384 + kk * 2048, 16384, 1024);
                  ptx::umma_f16_ss<CG>(tmem + 64 * s, ad, bd, idesc_acc2, (step > 1 || kk > 0) ? 1u : 0u);
                }
                ptx::umma_commit_mc<CG>(bar(bars.r_empty[st]), 0x3);
                ptx::umma_commit_mc<CG>(bar(bars.r_empty[st + 1]), 0x3);
                rc += 2;
                continue;
              }
              const uint32_t stage = rc % Cfg::NST, n = rc / Cfg::NST;
              ptx::mbar_wait(bar(bars.r_full[stage]), n & 1);
              ptx::tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t ad = ptx::make_smem_desc_sw128(sT + tbuf * 16384 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                const uint64_t bd = ptx::make_smem_desc_sw128(sR + stage * 16384 + kk * 2048, 16384, 1024);
                ptx::umma_f16_ss<CG>(tmem + 64 * s, ad, bd, idesc_acc, (step > 1 || kk > 0) ? 1u : 0u);
              }
              ptx::umma_commit_mc<CG>(bar(bars.r_empty[stage]), 0x3);
              ++rc;
            }
            ptx::umma_commit_mc<CG>(bar(bars.t_empty[tbuf]), 0x3);
            ++gp;
          }
        }
        ++it;
      }
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // =========================================== stash store warp (dQ kind) =====================
    // The dS tile the elementwise warps wrote for the MMA ([64 rows x 128 keys] of this CTA, two SW128 boxes)
    // goes to the stash with TMA stores straight from shared memory: no register traffic, no LSU work.
    if (KIND == kKindDQ && p.stash_ds != nullptr && ptx::elect_one()) {
      ptx::prefetch_tmap(&map_st);
      ptx::prefetch_tmap(&map_sp);
      const uint64_t pol = ptx::l2_policy_evict_first();   // written once, read once by the GEMM kernels: keep K/V in L2
      uint32_t g = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const Item itm = decode_item<KIND>(p, item_s, n_inner);
        if (itm.n <= 0 || itm.pass != 0) continue;   // head dims > 512: the second slab pass recomputes, pass 0 stores
        const int hs = itm.bh % heads_it, b = itm.bh / heads_it;
        for (int i = 0; i < itm.n; ++i, ++g) {       // g counts stored tiles only
          const int ci = itm.tr.first + (itm.s0 + i) % itm.tr.count;
          ptx::mbar_wait(bar(bars.t_written[0]), g & 1);
          // stash layout: [b * Hq + h][query tile][64-key block][128 queries][64 keys] -- every box this CTA
          // stores is one contiguous 8 KB run (row-major [Nq, Nk] scatters 128-byte pieces 2 Nk bytes apart,
          // which caps the DRAM write rate near 1.2 TB/s and starves the operand loads)
          const int blk = itm.rt * (p.nk_pad >> 6) + 2 * ci, bhq = b * p.heads_q + hs, qh = 64 * (int)rank;
          ptx::tma_store_4d_hint(&map_st, sT, 0, qh, blk, bhq, pol);                        // dS (T buffer 0)
          ptx::tma_store_4d_hint(&map_st, sT + 8192, 0, qh, blk + 1, bhq, pol);
          ptx::tma_store_4d_hint(&map_sp, sT + 16384, 0, qh, blk, bhq, pol);               // P_drop (staging buffer)
          ptx::tma_store_4d_hint(&map_sp, sT + 16384 + 8192, 0, qh, blk + 1, bhq, pol);
          ptx::bulk_commit_group();
          ptx::bulk_wait_group_read0();
          ptx::mbar_arrive(bar(bars.t_stored[0]));
        }
      }
      ptx::bulk_wait_group0();
    }
    __syncwarp();
  } else {
    // =========================================== elementwise warps + epilogue ===================
    const uint32_t t = threadIdx.x;
    const uint32_t lane128 = t & 127;
    const uint32_t row = lane128 & 63;
    const uint32_t kh = lane128 >> 6;   // 64-column half of the streamed tile / column half of ACC
    const uint32_t ch = t >> 7;         // 32-column half inside the S stage
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t l_t_full0 = ptx::mapa(bar(bars.t_full[0]), 0);
    const uint32_t l_t_full1 = ptx::mapa(bar(bars.t_full[1]), 0);
    uint32_t g = 0, gs = 0;   // tiles processed / tiles handed to the stash store warp
    for (uint32_t kidx = 0;; ++kidx) {
      const int item_s = next_item(p, cluster, nclusters, kidx);
      if (item_s < 0) break;
      const Item itm = decode_item<KIND>(p, item_s, n_inner);
      const int bh = itm.bh;
      const int hs = bh % heads_it, b = bh / heads_it;
      const int r0 = itm.rt * 128;
      const TileRange tr = itm.tr;
      const int T = itm.n;
      const int grow = r0 + 64 * (int)rank + (int)row;  // stationary row (query or key) inside its sequence
      const int seq_q = itm.nq, seq_kv = itm.nkv;
      const int off = seq_kv - seq_q;
      const bool row_ok = grow < ((KIND == kKindDQ) ? seq_q : seq_kv);
      const int tok0 = (KIND == kKindDQ) ? itm.qoff : itm.koff;  // token offset of the output rows (packed mode)
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.out) +
                      2 * ((int64_t)itm.bt * p.out_stride[0] + (int64_t)hs * p.out_stride[1] + (int64_t)(tok0 + grow) * p.out_stride[2]);
      if (T <= 0) {
        // nothing contributes: gradient is zero (already so in the pre-zeroed fp32 buffer of split mode)
        if (row_ok && p.out32 == nullptr && itm.pass == 0) {
          for (int d = (int)(kh * 2 + ch) * 8; d < p.head_dim; d += 32)
            *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      float row_lse2 = 0.f, row_delta = 0.f;
      if (KIND == kKindDQ) {
        const int64_t so = ((int64_t)b * p.heads_q + hs) * p.nq_pad + grow;  // grow < nq_pad always
        row_lse2 = p.lse2[so];
        row_delta = p.delta[so];
      }
      for (int i = 0; i < T; ++i, ++g) {
        const uint32_t sbuf = g & 1;
        const int sa = itm.s0 + i;
        const int gi = sa / tr.count, ci = tr.first + sa % tr.count;
        const int col0 = ci * 128 + 64 * (int)kh + 32 * (int)ch;  // first global column of this thread
        // column statistics (dK / dV kinds): issue the loads before waiting for the MMA
        float4 c_lse[8], c_dl[8];
        if (KIND != kKindDQ) {
          const int hq = hs * group + gi;
          const int64_t so = ((int64_t)b * p.heads_q + hq) * p.nq_pad + col0;
          const float4* pl = reinterpret_cast<const float4*>(p.lse2 + so);
          const float4* pd = reinterpret_cast<const float4*>(p.delta + so);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            c_lse[j] = __ldg(pl + j);
            if (HAS_DP) c_dl[j] = __ldg(pd + j);
          }
        }
        ptx::mbar_wait(bar(bars.s_full[sbuf]), (g >> 1) & 1);
        ptx::tc_fence_after();
        uint32_t sr[32], dr[32];
        ptx::tmem_ld_x32(tmem + lane_base + Cfg::S_BASE + 64 * sbuf + 32 * ch, sr);
        if (HAS_DP) ptx::tmem_ld_x32(tmem + lane_base + Cfg::DP_BASE + 64 * sbuf + 32 * ch, dr);
        ptx::tmem_wait_ld();
        // visibility: key <= query + off (causal), key < Nkv; padded / empty query rows carry lse2=+inf
        int lim_lo = 0, lim_hi = 0x7fffffff;  // visible columns: lim_lo <= col <= lim_hi
        if (KIND == kKindDQ) {
          lim_hi = seq_kv - 1;
          if (p.causal) { const int cl = grow + off; lim_hi = cl < lim_hi ? cl : lim_hi; }
        } else {
          if (p.causal) lim_lo = grow - off;
        }
        // A query that sees at most ONE key has an analytically zero dS row (dP == delta when P == 1): write the
        // exact zero instead of the ~1 ulp residue of two differently ordered fp32 reductions (the reference's
        // sm_100 kernels make the same promise, /root/reference/tests/test_ffpa_cute_sm100.py:1000-1050).
        // single_hi: queries <= single_hi are such rows (causal: q + off <= 0; otherwise all when Nkv == 1).
        // (not when a gradient flows in through the LSE output: then dS = P * dLSE on such a row)
        const int single_hi = !p.zero_single ? -0x7fffffff : p.causal ? -off : (seq_kv <= 1 ? 0x7fffffff : -1);
        const bool any_single = (KIND == kKindDQ) ? (grow <= single_hi) : (col0 <= single_hi);
        uint32_t pk[16];
        uint32_t pp[16];  // stash path: packed P_drop, stored to global after the tile has been handed to the MMA
        const bool stash = (KIND == kKindDQ) && p.stash_ds != nullptr;   // launch-wide: T single-buffered
        const bool stash_w = stash && itm.pass == 0;                     // this item's tiles are stored
        // GENERAL: additive bias, Philox dropout replay, dBias output (dQ kind). Query / key of
        // element jj: dQ kind (q = grow, key = col), dK/dV kinds (q = col, key = grow).
        const int hq_cur = (KIND == kKindDQ) ? hs : hs * group + gi;
        const float inv_keep = GENERAL ? 1.f / (1.f - p.dropout_p) : 1.f;
        float dsv[(GENERAL && KIND == kKindDQ) ? 32 : 1];   // fp32 dS of this thread's 32 columns (dBias)
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float e[2], pv[2] = {0.f, 0.f};
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int jj = j + u;
            float l2, dl;
            if (KIND == kKindDQ) { l2 = row_lse2; dl = row_delta; }
            else {
              const float4 a = c_lse[jj >> 2];
              l2 = (jj & 3) == 0 ? a.x : (jj & 3) == 1 ? a.y : (jj & 3) == 2 ? a.z : a.w;
              if (HAS_DP) { const float4 d4 = c_dl[jj >> 2]; dl = (jj & 3) == 0 ? d4.x : (jj & 3) == 1 ? d4.y : (jj & 3) == 2 ? d4.z : d4.w; }
              else dl = 0.f;
            }
            const int col = col0 + jj;
            const int qi = (KIND == kKindDQ) ? grow : col;
            const int ki = (KIND == kKindDQ) ? col : grow;
            float xs = fmaf(__uint_as_float(sr[jj]), p.scale_log2, -l2);
            float mult = 1.f;
            if constexpr (GENERAL) {
              const bool inb = qi < seq_q && ki < seq_kv;
              if (p.bias_kind != 0 && inb) {
                const int64_t bo = (int64_t)b * p.bias_stride[0] + (int64_t)hq_cur * p.bias_stride[1] +
                                   (int64_t)qi * p.bias_stride[2] + (int64_t)ki * p.bias_stride[3];
                float bv;
                if (p.bias_kind == 1) bv = reinterpret_cast<const float*>(p.bias)[bo];
                else if (BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[bo]);
                else bv = __half2float(reinterpret_cast<const __half*>(p.bias)[bo]);
                xs = fmaf(bv, 1.4426950408889634f, xs);
              }
              if (p.dropout_p > 0.f && inb) {
                const uint64_t eo = p.philox_offset +
                    ((uint64_t)((int64_t)b * p.heads_q + hq_cur) * (uint64_t)p.seqlen_q + (uint64_t)qi) * (uint64_t)p.seqlen_kv + (uint64_t)ki;
                const uint4 r4 = philox4x32_10(p.philox_seed, eo >> 2);
                const uint32_t sel = (uint32_t)(eo & 3);
                const uint32_t rv = sel == 0 ? r4.x : sel == 1 ? r4.y : sel == 2 ? r4.z : r4.w;
                const float uni = ((float)rv + 1.0f) * 2.3283064365386963e-10f;
                mult = (uni > p.dropout_p) ? inv_keep : 0.f;
              }
            }
            float pe = exp2f(xs);
            if (col < lim_lo || col > lim_hi) pe = 0.f;
            if (HAS_DP) {
              float ds = pe * (__uint_as_float(dr[jj]) * mult - dl);
              if (any_single && qi <= single_hi) ds = 0.f;
              e[u] = ds;
              if (KIND == kKindDQ) pv[u] = pe * mult;   // P_drop (what dV consumes), only used by the stash path
              if constexpr (GENERAL && KIND == kKindDQ) dsv[jj] = ds;
            } else {
              e[u] = pe * mult;
            }
          }
          pk[j >> 1] = BF16 ? ptx::pack_bf16x2(e[0], e[1]) : ptx::pack_f16x2(e[0], e[1]);
          if (KIND == kKindDQ) pp[j >> 1] = BF16 ? ptx::pack_bf16x2(pv[0], pv[1]) : ptx::pack_f16x2(pv[0], pv[1]);
        }
        // dBias = dS reduced over the dims the bias broadcasts over, accumulated straight into the bias-shaped
        // fp32 buffer: a bias without a query dim is first summed over the warp's 32 rows (butterfly
        // reduce-scatter, 31 shuffles: lane L ends up with column col0 + L), then one atomic per column.
        if constexpr (GENERAL && KIND == kKindDQ) {
          if (p.dbias != nullptr && itm.pass == 0) {
            float* dbase = p.dbias + (int64_t)b * p.dbias_stride[0] + (int64_t)hq_cur * p.dbias_stride[1];
            const bool red_bh = (p.dbias_stride[0] == 0 && p.batch > 1) || (p.dbias_stride[1] == 0 && p.heads_q > 1);
            if (p.dbias_stride[2] != 0 || seq_q == 1) {
              const bool red = red_bh || (p.dbias_stride[3] == 0 && seq_kv > 1);
              if (grow < seq_q) {
                float* drow = dbase + (int64_t)grow * p.dbias_stride[2];
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                  const int ki = col0 + jj;
                  if (ki < seq_kv) {
                    if (red) atomicAdd(drow + (int64_t)ki * p.dbias_stride[3], dsv[jj]);
                    else drow[(int64_t)ki * p.dbias_stride[3]] = dsv[jj];
                  }
                }
              }
            } else {
              const uint32_t ln = threadIdx.x & 31;
#pragma unroll
              for (int w = 16; w >= 1; w >>= 1) {
                const bool hi = (ln & w) != 0;
#pragma unroll
                for (int i2 = 0; i2 < w; ++i2) {
                  const float keep = hi ? dsv[i2 + w] : dsv[i2];
                  const float send = hi ? dsv[i2] : dsv[i2 + w];
                  dsv[i2] = keep + __shfl_xor_sync(0xffffffffu, send, w);
                }
              }
              const int ki = col0 + (int)ln;
              if (ki < seq_kv) atomicAdd(dbase + (int64_t)ki * p.dbias_stride[3], dsv[0]);
            }
          }
        }
        // stash path: T single-buffered (buffer 0), buffer 1 stages P_drop; both leave through the store warp
        const uint32_t tb = stash ? 0u : sbuf;
        ptx::mbar_wait(bar(bars.t_empty[tb]), ((stash ? g : (g >> 1)) & 1) ^ 1);
        // store warp has drained both buffers of the most recent stored tile (gs = tiles stored so far)
        if (stash && gs > 0) ptx::mbar_wait(bar(bars.t_stored[0]), (gs - 1) & 1);
        if (stash_w) {
          const uint32_t prow = sT + 16384 + kh * 8192 + row * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t addr = prow + (((4 * ch + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pp[4 * c]),
                         "r"(pp[4 * c + 1]), "r"(pp[4 * c + 2]), "r"(pp[4 * c + 3])
                         : "memory");
          }
        }
        {
          const uint32_t prow = sT + tb * 16384 + kh * 8192 + row * 128;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t addr = prow + (((4 * ch + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (ptx::lane_id() == 0) {
          ptx::mbar_arrive_cluster(tb ? l_t_full1 : l_t_full0);
          if (stash_w) ptx::mbar_arrive(bar(bars.t_written[0]));
        }
        if (stash_w) ++gs;
      }
      // ---------------- epilogue: ACC (x scale) -> global ----------------
      {
        const uint32_t gl = g - 1;
        const bool stash_e = (KIND == kKindDQ) && p.stash_ds != nullptr;
        ptx::mbar_wait(bar(bars.t_empty[stash_e ? 0u : (gl & 1)]), (stash_e ? gl : (gl >> 1)) & 1);
        ptx::tc_fence_after();
        const float mulo = (KIND == kKindDV) ? 1.f : p.scale;
#pragma unroll
        for (int s = 0; s < Cfg::NSL; ++s) {
          uint32_t orr[32];
          ptx::tmem_ld_x32(tmem + lane_base + 64 * s + 32 * ch, orr);
          ptx::tmem_wait_ld();
          // TMEM column 64 s + 32 ch + j of lane half kh: N=128 slices -> d = 128 s + 64 kh + 32 ch + j;
          // N=256 slices (WIDE) -> column c = 64 (s & 1) + 32 ch + j of slice s/2 -> d = 256 (s/2) + 128 kh + c
          const int d0 = itm.pass * Cfg::SLAB +
                         (Cfg::WIDE ? 256 * (s >> 1) + 128 * (int)kh + 64 * (s & 1) + 32 * (int)ch
                                    : 128 * s + 64 * (int)kh + 32 * (int)ch);
          if (row_ok && p.out32 != nullptr) {
            // split mode: accumulate this chunk's partial result (fp32, [B, H, rows, D] contiguous)
            float* dst = p.out32 + (((int64_t)b * heads_it + hs) * p.out_rows + grow) * (int64_t)p.head_dim;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (d0 + j < p.head_dim) atomicAdd(dst + d0 + j, __uint_as_float(orr[j]) * mulo);
          } else if (row_ok) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const int d = d0 + 8 * v;
              if (d < p.head_dim) {
                uint32_t w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float a = __uint_as_float(orr[8 * v + 2 * u]) * mulo;
                  const float c = __uint_as_float(orr[8 * v + 2 * u + 1]) * mulo;
                  w[u] = BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
                }
                *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        ptx::tc_fence_before();
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == kMmaWarp) ptx::tmem_dealloc<CG>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// preprocess: delta = rowsum(dO * O), lse2 = LSE * log2(e) (+inf when the row saw no key), padded
// to a multiple of 128 rows per (b, h).  One warp per row.  (reference: _ffpa_bwd.py:236-306)
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void bwd_preprocess_kernel(const void* __restrict__ o, const void* __restrict__ d_o,
                                      const float* __restrict__ lse, float* __restrict__ lse2,
                                      float* __restrict__ delta, int64_t os0, int64_t os1, int64_t os2,
                                      int64_t ds0, int64_t ds1, int64_t ds2, int B, int H, int Nq,
                                      int nq_pad, int D, const int* __restrict__ cu_q, int total_q,
                                      const float* __restrict__ dlse) {
  const int warps_per_block = blockDim.x >> 5;
  const int64_t rowid = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t total = (int64_t)B * H * nq_pad;
  if (rowid >= total) return;
  const int lane = threadIdx.x & 31;
  const int q = (int)(rowid % nq_pad);
  const int64_t bh = rowid / nq_pad;
  const int h = (int)(bh % H), b = (int)(bh / H);
  // packed mode: sequence b owns tokens [cu_q[b], cu_q[b+1]); LSE is [H, total_q]
  int64_t tok = q, bt = b, lse_idx = bh * Nq + q;
  if (cu_q != nullptr) {
    const int q0 = min(max(__ldg(cu_q + b), 0), total_q);
    Nq = min(__ldg(cu_q + b + 1), total_q) - q0;
    tok = (int64_t)q0 + q; bt = 0; lse_idx = (int64_t)h * total_q + tok;
  }
  if (q >= Nq) {
    if (lane == 0) { lse2[rowid] = INFINITY; delta[rowid] = 0.f; }
    return;
  }
  const uint8_t* po = reinterpret_cast<const uint8_t*>(o) + 2 * (bt * os0 + h * os1 + tok * os2);
  const uint8_t* pd = reinterpret_cast<const uint8_t*>(d_o) + 2 * (bt * ds0 + h * ds1 + tok * ds2);
  float acc = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    const uint4 a = *reinterpret_cast<const uint4*>(po + 2 * d);
    const uint4 c = *reinterpret_cast<const uint4*>(pd + 2 * d);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float a0, a1, c0, c1;
      if (BF16) {
        a0 = __uint_as_float(aw[u] << 16); a1 = __uint_as_float(aw[u] & 0xffff0000u);
        c0 = __uint_as_float(cw[u] << 16); c1 = __uint_as_float(cw[u] & 0xffff0000u);
      } else {
        const __half2 ha = *reinterpret_cast<const __half2*>(&aw[u]);
        const __half2 hc = *reinterpret_cast<const __half2*>(&cw[u]);
        a0 = __low2float(ha); a1 = __high2float(ha); c0 = __low2float(hc); c1 = __high2float(hc);
      }
      acc = fmaf(a0, c0, acc);
      acc = fmaf(a1, c1, acc);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    const float l = lse[lse_idx];
    lse2[rowid] = (l == -INFINITY) ? INFINITY : l * 1.4426950408889634f;
    // dLSE: d lse_i / d s_ij = p_ij, so dS = P (dP - delta + dlse): fold it into delta
    delta[rowid] = dlse != nullptr ? acc - dlse[lse_idx] : acc;
  }
}


template <int NQK, bool BF16, int KIND, bool GENERAL>
static int launch_bwd_variant(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1,
                              const CUtensorMap& b2, const CUtensorMap& b3, const CUtensorMap& st, const CUtensorMap& sp, const BwdKernelParams& kp,
                              int nclusters, cudaStream_t stream) {
  using Cfg = BwdCfg<NQK, KIND>;
  auto kern = ffpa_bwd_kernel<NQK, BF16, KIND, GENERAL>;
  // the opt-in shared-memory size is a per-device function attribute
  static bool attr_set[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  dev_id = (dev_id >= 0 && dev_id < 64) ? dev_id : 0;
  if (!attr_set[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_DYN);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaFuncSetAttribute(bwd smem=%d): %s", Cfg::SMEM_DYN, cudaGetErrorString(e));
    attr_set[dev_id] = true;
  }
  kern<<<dim3(2 * nclusters), dim3(kThreads), Cfg::SMEM_DYN, stream>>>(a1, a2, b1, b2, b3, st, sp, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "backward launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

template <bool BF16, int KIND, bool GENERAL>
static int dispatch_bwd_nqk(int nqk, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1,
                            const CUtensorMap& b2, const CUtensorMap& b3, const CUtensorMap& st, const CUtensorMap& sp, const BwdKernelParams& kp,
                            int nclusters, cudaStream_t stream) {
  switch (nqk) {
    case 1: return launch_bwd_variant<1, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 2: return launch_bwd_variant<2, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 3: return launch_bwd_variant<3, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 4: return launch_bwd_variant<4, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 5: return launch_bwd_variant<5, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 6: return launch_bwd_variant<6, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 7: return launch_bwd_variant<7, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 8: return launch_bwd_variant<8, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    // head_dim > 512: the launcher rounds the box count up to an even number (TMA zero fill)
    case 10: return launch_bwd_variant<10, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 12: return launch_bwd_variant<12, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 14: return launch_bwd_variant<14, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    case 16: return launch_bwd_variant<16, BF16, KIND, GENERAL>(a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    default: return set_error(FFPA_ERR_UNSUPPORTED, "backward supports head_dim <= 1024");
  }
}

// kind: 0 dQ, 1 dK, 2 dV
template <bool BF16>
int dispatch_bwd_dtype(int nqk, int kind, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1,
                       const CUtensorMap& b2, const CUtensorMap& b3, const CUtensorMap& st, const CUtensorMap& sp, const BwdKernelParams& kp, int nclusters,
                       cudaStream_t stream) {
  const bool general = kp.bias_kind != 0 || kp.dropout_p > 0.f || kp.dbias != nullptr;
  if (general) {
    if (kind == kKindDQ) return dispatch_bwd_nqk<BF16, kKindDQ, true>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    if (kind == kKindDK) return dispatch_bwd_nqk<BF16, kKindDK, true>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
    return dispatch_bwd_nqk<BF16, kKindDV, true>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
  }
  if (kind == kKindDQ) return dispatch_bwd_nqk<BF16, kKindDQ, false>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
  if (kind == kKindDK) return dispatch_bwd_nqk<BF16, kKindDK, false>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
  return dispatch_bwd_nqk<BF16, kKindDV, false>(nqk, a1, a2, b1, b2, b3, st, sp, kp, nclusters, stream);
}

template <bool BF16>
int launch_preprocess(const ffpa_bwd_params& a, float* lse2, float* delta, int nq_pad, cudaStream_t stream) {
  const int64_t rows = (int64_t)a.batch * a.heads_q * nq_pad;
  const int wpb = 8;
  const int64_t blocks = (rows + wpb - 1) / wpb;
  bwd_preprocess_kernel<BF16><<<dim3((unsigned)blocks), dim3(wpb * 32), 0, stream>>>(
      a.o, a.d_o, a.lse, lse2, delta, a.o_stride[0], a.o_stride[1], a.o_stride[2], a.do_stride[0],
      a.do_stride[1], a.do_stride[2], a.batch, a.heads_q, a.seqlen_q, nq_pad, a.head_dim, a.cu_seqlens_q, a.total_q, a.d_lse);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "backward preprocess launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

}  // namespace bwd
}  // namespace ffpa
