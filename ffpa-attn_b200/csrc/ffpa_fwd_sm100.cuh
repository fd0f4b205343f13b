// B200 (sm_100a) Split-D attention forward: tcgen05 MMA into TMEM, TMA loads, 2-CTA cluster,
// warp-specialised persistent kernel.
//
// Replaces the reference's forward kernels behind the same boundary:
//   /root/reference/csrc/cuffpa/launch.cuh:61-606               (launcher contract / checks)
//   /root/reference/csrc/cuffpa/native/sm_80/split_d.cuh:85-777 (Split-D algorithm)
//   /root/reference/csrc/cuffpa/native/prefill.cuh:252-1172     (numerics: mask, online softmax
//                                                                 with lazy rescale, dropout, LSE)
// Design:
//   * a 2-CTA cluster owns 128 query rows of one (batch, head); CTA r owns rows [64r, 64r+64).
//   * S = Q K^T  : tcgen05.mma cta_group::2, M=128 N=128, K = head_dim split into 64-wide boxes
//                  ("Split-D"); each CTA stages its 64 Q rows and its 64 keys of the KV tile.
//   * O += P V   : tcgen05.mma cta_group::2, M=128 N=256 (head-dim slices), K = 128 keys;
//                  each CTA stages its P rows and its 128 head-dim columns of V (MN-major).
//   * accumulators are "lane folded": TMEM lane l of CTA r holds row 64r + l%64, column half l/64.
//   * warps 0-7 softmax/correction/epilogue (two warpgroups, each owning half of the S columns),
//     warp 8 MMA issue (leader CTA) + TMEM alloc, warp 9 TMA producer.
//     mbarrier pipelines: K ring, V ring, S (2 stages), P (2 stages).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cmath>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

#ifndef FFPA_ALT_MAX_HD
#define FFPA_ALT_MAX_HD 256
#endif
// Epilogue variant. 0 (default): the softmax warps store O themselves, staged in the idle P ring and written with TMA
// bulk stores. 1: four extra warps (512 threads, setmaxnreg re-allocation) read the finished O tile out of TMEM -- parked
// as a 16-bit copy in spare TMEM columns where there are any -- so the softmax warps go straight from the last KV tile
// of an item to the first of the next one; they have no shared memory left to stage in and must use per-thread global
// stores, whose LSU traffic costs more than the decoupling gains: measured 3-4 % SLOWER than variant 0 on causal /
// short-sequence shapes, equal at C2 (profiles/r02_fwd_epilogue.md). Kept as a build option (-DFFPA_EPI_WARPS=1).
#ifndef FFPA_EPI_WARPS
#define FFPA_EPI_WARPS 0
#endif
#ifndef FFPA_UNIFIED_MIN_HD
#define FFPA_UNIFIED_MIN_HD 512
#endif

namespace ffpa {

constexpr int kSmemLimit = 232448;

// softmax flavour
constexpr int kModeFast = 0;     // no bias, scale > 0, no dropout
constexpr int kModeGeneral = 1;  // additive bias and/or scale <= 0
constexpr int kModeDropout = 2;  // general + Philox dropout

template <int NQK>
struct FwdCfg {
  static constexpr int HD = NQK * 64;                     // padded head dim for QK^T
  static constexpr int DVP = ((HD + 127) / 128) * 128;    // padded head dim for PV / O
  // O must stay TMEM-resident next to the two S stages (128 columns): at most 384 columns = 768
  // head dims per pass. Wider heads run two passes over 512-wide O slabs, recomputing S.
  static constexpr int NPASS = DVP > 768 ? 2 : 1;
  static constexpr int DSLAB = NPASS == 1 ? DVP : 512;    // head dims of O per pass (last pass may be narrower)
  static constexpr int O_COLS = DSLAB / 2;                // TMEM columns of O per CTA (lane folded)
  static constexpr int NSLICE = (DSLAB + 255) / 256;      // max PV N-slices per pass (256 wide, last may be 128)
  static constexpr int KST = (NQK + 1) / 2;               // 16 KB K stages per KV tile
  static constexpr int S_BASE = O_COLS > 256 ? 384 : 256; // TMEM column of S[0]; S[1] = +64
  static constexpr int Q_BYTES = NQK * 8192;
  // depth of the S (TMEM) / P (SMEM) pipeline = tiles whose softmax may be in flight. The
  // S -> softmax -> P -> PV chain costs ~2000 cycles of latency per tile; with depth k the tensor
  // pipe stays busy when (T_mma + 2000) / k <= T_mma, i.e. k = 2 suffices at D = 512 (T_mma = 2048)
  // but small heads need 3-4 (TMEM: O_COLS + 64 k <= 512).
  static constexpr int KSTG = HD <= 256 ? 4 : (HD <= 384 ? 3 : 2);
  // softmax warps: 8 (two per scheduler, each thread owns 32 keys of a row) when the tensor pipe is the
  // limiter, 16 (four per scheduler, 16 keys per thread) for small heads where the SIMT side is
  // (profiles/r01_fwd_small_d_ncu.md)
  // Measured (profiles/r01_fwd_timings_v5.log): 16 warps are 5-8 % SLOWER at D <= 256 than 8 -- the SIMT
  // limit there is pipe throughput (MUFU.EX2 + conversions), not latency hiding -- so 8 is used everywhere;
  // the kernel stays generic in NSW.
  static constexpr int NSW = 8;
  // ALT: the two softmax warpgroups take ALTERNATE KV tiles (each thread a whole 64-key lane of its
  // tile) instead of splitting the columns of one tile; tiles i and i+1 are then in different phases
  // (load / max / exp2 / pack) so MUFU and ALU work overlap. Pays off when the SIMT side is the limiter
  // (small heads; the FP8 kernel always uses it). Needs NSW == 8.
  static constexpr bool ALT = HD <= FFPA_ALT_MAX_HD;
  static constexpr int CQ = ALT ? 1 : NSW / 4;   // column groups of the 64-column S stage
  static constexpr int CPT = 64 / CQ;            // S columns (keys) per thread and tile
  static constexpr int MMA_WARP = NSW, TMA_WARP = NSW + 1;
  // two-slab head dims carry one more warp: it TMA-stores the P tiles for the replay path (see FwdKernelParams)
  static constexpr int STORE_WARP = NSW + 2;
  static constexpr int V_WARP = NSW + 2;   // separate-ring head dims (<= 512): a second TMA producer warp streams V
  // epilogue warps: warps 12..15 (a whole warpgroup, TMEM lane quarter = warp % 4); warps 8..11 = MMA, TMA, V / P-store,
  // idle. Register file: 512 threads start with 128 registers; the two softmax warpgroups grow to 168, the producer /
  // MMA warpgroup shrinks to 88 and the epilogue warpgroup to 80 (2 x 128 x 168 + 128 x 88 + 128 x 80 = 64512).
  static constexpr bool EPIW = FFPA_EPI_WARPS != 0;
  static constexpr int EPI_WARP0 = 12;
  static constexpr int THREADS = EPIW ? 512 : (NSW + 2 + ((DVP > 768 || HD <= FFPA_UNIFIED_MIN_HD) ? 1 : 0)) * 32;
  static_assert(!EPIW || NSW == 8, "the epilogue warpgroup layout assumes 8 softmax warps");
  static constexpr int P_BYTES = KSTG * 16384;
  // Wide heads keep Q resident (96-128 KB), leaving too little for separate K and V rings; they use
  // ONE ring of 16 KB stages shared by K stages and V slices (a slice = 2 consecutive stages), so
  // loads can run a full ring ahead regardless of operand. Needs even stage counts per tile.
  static constexpr bool UNIFIED = HD > FFPA_UNIFIED_MIN_HD;
  static constexpr bool K_DUMMY = UNIFIED && (KST % 2 == 1);   // pad the K stages of a tile to an even count
  static constexpr int NVS = HD > 768 ? 1 : 2;           // 32 KB V stages (separate-ring mode)
  static constexpr int NVS_ALLOC = UNIFIED ? 0 : NVS;
  static constexpr int kBudget = kSmemLimit - ((NSW == 16 || ALT) ? 5120 : 3072);  // static smem (barriers + exchange), 1 KB aligned
  static constexpr int kNksRaw = (kBudget - Q_BYTES - P_BYTES - NVS_ALLOC * 32768) / 16384;
  static constexpr int kNksCap = kNksRaw > 8 ? 8 : kNksRaw;
  static constexpr int NKS = UNIFIED ? (kNksCap & ~1) : kNksCap;   // 16 KB K stages (or unified ring stages)
  static constexpr int SMEM_DYN = Q_BYTES + P_BYTES + NKS * 16384 + NVS_ALLOC * 32768;
  static_assert(NKS >= 2, "not enough shared memory for the K ring");
  static_assert(!UNIFIED || (NKS % 2 == 0 && NKS >= 4), "unified ring needs an even number of stages");
  static_assert(S_BASE + 64 * KSTG <= 512 && S_BASE >= O_COLS, "O does not fit TMEM next to S");
  // TMEM columns left over next to O and the S ring: when they hold the 16-bit copy of O (O_COLS / 2 columns) the
  // epilogue warps first "park" the normalised tile there -- TMEM to TMEM, ~1000 cycles -- and release O for the next
  // item's first PV MMA before they start the slow part, the global stores. -1: no room (head dims 320-384, 576-768).
  static constexpr int PARK_BASE = (512 - (S_BASE + 64 * KSTG) >= O_COLS / 2) ? (S_BASE + 64 * KSTG)
                                   : ((S_BASE - O_COLS >= O_COLS / 2) ? O_COLS : -1);
  // width of the O slab of pass `pass`, and N of its slice s
  __host__ __device__ static constexpr int slab_w(int pass) { return (DVP - pass * DSLAB) >= DSLAB ? DSLAB : (DVP - pass * DSLAB); }
  __host__ __device__ static constexpr int slice_n(int w, int s) { return (w - 256 * s) >= 256 ? 256 : 128; }
};

struct Barriers {
  uint64_t q_full, q_empty;
  uint64_t k_full[8], k_empty[8];
  uint64_t v_full[3], v_empty[3];
  uint64_t s_full[4];
  uint64_t p_full[4], p_empty[4];
  uint64_t m_full[4];
  uint64_t p_written[4], p_stored[4];   // replay path: P tile complete in this CTA / drained by the store warp
  uint64_t o_ready, inv_taken, o_free;  // epilogue warps: 1 / row sum published, ... consumed, O read out of TMEM (leader CTA)
};

__device__ __forceinline__ int num_kv_tiles(int causal, int nq, int nkv, int q0) {
  int tc = (nkv + 127) >> 7;
  if (causal) {
    int lim = ((q0 + 127 + (nkv - nq)) >> 7) + 1;
    tc = lim < tc ? lim : tc;
  }
  return tc < 1 ? 1 : tc;
}

// k-th work item of this cluster: from the balanced schedule table when present (causal), else
// static round robin. Every role of both CTAs evaluates the same sequence.
__device__ __forceinline__ int next_item(const FwdKernelParams& p, uint32_t cluster, uint32_t nclusters, uint32_t k) {
  if (p.sched != nullptr) return (k < (uint32_t)p.sched_stride) ? __ldg(p.sched + (size_t)cluster * p.sched_stride + k) : -1;
  const uint32_t item = cluster + k * nclusters;
  return item < (uint32_t)p.n_items ? (int)item : -1;
}

// decoded work item. With kv_splits > 1 (few query tiles, long KV: decode-like shapes) an item covers only
// the KV tiles [tbeg, tbeg + T) and writes an fp32 partial (O_s, LSE_s) that merge_splits_kernel combines.
// Packed variable-length mode (p.cu_q != nullptr; reference API ffpa_attn_varlen_func,
// /root/reference/src/ffpa_attn/ffpa_attn_interface.py:192-279): `batch` counts sequences, the tensors are
// [total tokens, H, D] (tensor-map batch extent 1), and the item carries its sequence's lengths and token
// offsets read from cu_seqlens; query tiles past the end of a short sequence are empty items.
struct FwdItem { int mt, pass, bh, split, tbeg, T; int nq, nkv, qoff, koff, bt; };
template <int NPASS>
__device__ __forceinline__ FwdItem decode_fwd_item(const FwdKernelParams& p, uint32_t item) {
  FwdItem it;
  it.mt = item % p.n_mtiles;
  uint32_t rest = item / p.n_mtiles;
  it.pass = rest % p.n_pass;
  rest /= p.n_pass;
  it.split = rest % p.kv_splits;
  it.bh = rest / p.kv_splits;
  if (p.cu_q != nullptr) {
    const int b = it.bh / p.heads_q;
    // offsets are clamped to the packed extents so a malformed cu_seqlens can never address past the tensors
    it.qoff = min(max(__ldg(p.cu_q + b), 0), p.total_q);
    it.nq = min(__ldg(p.cu_q + b + 1), p.total_q) - it.qoff;
    it.koff = min(max(__ldg(p.cu_k + b), 0), p.total_k);
    it.nkv = max(min(__ldg(p.cu_k + b + 1), p.total_k) - it.koff, 0);
    it.bt = 0;
    if (it.mt * 128 >= it.nq) { it.tbeg = 0; it.T = 0; return it; }
  } else {
    it.nq = p.seqlen_q; it.nkv = p.seqlen_kv; it.qoff = 0; it.koff = 0; it.bt = it.bh / p.heads_q;
  }
  // replay path: the second-slab kernel consumes 256-row blocks, so both query tiles of a block visit the same
  // KV tiles (the extra one of the even tile is fully masked and stores P = 0)
  const int ttot = num_kv_tiles(p.causal, it.nq, it.nkv, p.stash_p != nullptr ? ((it.mt * 128) | 128) : it.mt * 128);
  if (p.kv_splits == 1) { it.tbeg = 0; it.T = ttot; }
  else {
    const int per = (ttot + p.kv_splits - 1) / p.kv_splits;
    it.tbeg = it.split * per;
    int n = ttot - it.tbeg;
    n = n < 0 ? 0 : n;
    it.T = n < per ? n : per;
  }
  return it;
}

__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t ctr) {
  // Philox-4x32-10, counter = (ctr_lo, ctr_hi, 0, 0), key = seed. Same generator as
  // /root/reference/csrc/cuffpa/native/prefill.cuh:398-422 (and curand / torch SDPA).
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
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

#ifdef FFPA_TRACE
// development aid (never built into the shipped library): per-cluster, per-item clock64 stamps of the pipeline roles
__device__ unsigned long long* g_fwd_trace = nullptr;
#ifdef FFPA_TRACE_SETTER
extern "C" void ffpa_dbg_set_fwd_trace(unsigned long long* ptr) { cudaMemcpyToSymbol(g_fwd_trace, &ptr, sizeof(ptr)); }
#endif
#define FFPA_STAMP(kidx_, ev_)                                                                              \
  do {                                                                                                      \
    if (g_fwd_trace != nullptr && rank == 0 && (kidx_) < 64)                                                \
      g_fwd_trace[((size_t)cluster * 64 + (kidx_)) * 16 + (ev_)] = (unsigned long long)clock64();            \
  } while (0)
#else
#define FFPA_STAMP(kidx_, ev_) do {} while (0)
#endif

template <int NQK, bool BF16, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FwdCfg<NQK>::THREADS, 1)
ffpa_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_sp,
                const __grid_constant__ CUtensorMap map_o, const FwdKernelParams p) {
  using Cfg = FwdCfg<NQK>;
  constexpr int CG = 2;
  constexpr uint32_t KS = Cfg::KSTG;   // S/P pipeline depth
  constexpr int LA = Cfg::KSTG - 1;    // QK runs LA tiles ahead of PV
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Barriers bars;
  constexpr int kMmaWarp = Cfg::MMA_WARP, kTmaWarp = Cfg::TMA_WARP, kSoftmaxWarps = Cfg::NSW;
  constexpr int CQ = Cfg::CQ, CPT = Cfg::CPT;
  __shared__ float xch[2][Cfg::ALT ? 4 : 2 * Cfg::CQ][64];  // row-max exchange: [parity][kh*CQ+ch][row]  (ALT: [parity][wg*2+kh][row])
  __shared__ float mval[Cfg::ALT ? 4 : 1][Cfg::ALT ? 64 : 1];   // ALT: running row max published per tile (ring of 4)
  constexpr bool ALT = Cfg::ALT;
  __shared__ uint32_t tmem_slot;
  __shared__ float epi_inv[Cfg::EPIW ? 64 : 1];   // 1 / row sum of the item handed to the epilogue warps
  __shared__ float epi_lse[Cfg::EPIW ? 64 : 1];   // ... and its LSE: a global store issued by a softmax thread would make
                                                  // the generic->async proxy fence of its next P tile wait behind the
                                                  // epilogue warps' store traffic (profiles/r02_fwd_epilogue.md)

  const uint32_t smem_base = ptx::smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();  // SWIZZLE_128B tiles need 1 KB alignment
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
    for (int i = 0; i < 3; ++i) { ptx::mbar_init(bar(bars.v_full[i]), 1); ptx::mbar_init(bar(bars.v_empty[i]), 1); }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(bar(bars.s_full[i]), 1);
      ptx::mbar_init(bar(bars.p_full[i]), Cfg::ALT ? kSoftmaxWarps : 2 * kSoftmaxWarps);  // softmax warps (of one tile) of both CTAs
      ptx::mbar_init(bar(bars.m_full[i]), 2);
      ptx::mbar_init(bar(bars.p_empty[i]), 1);
      ptx::mbar_init(bar(bars.p_written[i]), kSoftmaxWarps);
      ptx::mbar_init(bar(bars.p_stored[i]), 1);
    }
    ptx::mbar_init(bar(bars.o_ready), 2);     // softmax warps 0 and 1 (one writer per row)
    ptx::mbar_init(bar(bars.inv_taken), 4);   // epilogue warps of this CTA
    ptx::mbar_init(bar(bars.o_free), 8);      // epilogue warps of both CTAs
    ptx::fence_mbar_init();
  }
  if (warp == kTmaWarp && ptx::elect_one()) {
    ptx::prefetch_tmap(&map_q);
    ptx::prefetch_tmap(&map_k);
    ptx::prefetch_tmap(&map_v);
    if (p.o_tma) ptx::prefetch_tmap(&map_o);
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

  if (Cfg::EPIW && warp >= (uint32_t)Cfg::EPI_WARP0) {
    // =========================================== epilogue warps =================================
    // O (fp32, TMEM) x 1 / row sum -> 16 bit -> global. Per item: wait for the row sums (o_ready) and for the last
    // PV MMA (p_empty of the last tile), read O, hand TMEM back (o_free, awaited by the MMA issuer before the first
    // PV MMA of the next item), store. Every lane owns another row, so a store instruction is 32 LSU wavefronts: slow,
    // but no longer between two items of the softmax warps.
    if constexpr (Cfg::EPIW) {
      ptx::setmaxnreg_dec<80>();
      const uint32_t lane128 = (warp & 3u) * 32u + ptx::lane_id();
      const uint32_t row = lane128 & 63u, kh = lane128 >> 6;
      const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
      const uint32_t l_o_free = ptx::mapa(bar(bars.o_free), 0);
      uint32_t g = 0, it = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const FwdItem fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)item_s);
        if (fi.T <= 0) continue;
        const uint32_t gl = g + (uint32_t)fi.T - 1;
        g += (uint32_t)fi.T;
        const int h = fi.bh % p.heads_q, b = fi.bh / p.heads_q;
        const int gq = fi.mt * 128 + 64 * (int)rank + (int)row;
        const bool row_ok = gq < fi.nq;
        const int dv0 = fi.pass * Cfg::DSLAB, dvw = Cfg::slab_w(fi.pass);
        uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                        2 * ((int64_t)fi.bt * p.o_stride[0] + (int64_t)h * p.o_stride[1] + (int64_t)(fi.qoff + gq) * p.o_stride[2]);
        const bool o_al32 = (reinterpret_cast<uintptr_t>(orow) & 31u) == 0;
        ptx::mbar_wait(bar(bars.o_ready), it & 1);
        const float inv = epi_inv[row];
        const float lse = epi_lse[row];
        __syncwarp();
        if (ptx::lane_id() == 0) ptx::mbar_arrive(bar(bars.inv_taken));
        if (kh == 0 && row_ok) {
          if (Cfg::NPASS == 2 && p.stash_p != nullptr) p.stash_inv[(int64_t)fi.bh * p.n_mt_even * 128 + gq] = inv;
          if ((p.lse != nullptr || p.kv_splits > 1) && fi.pass == 0) {
            if (p.kv_splits > 1) p.part_lse[(((int64_t)fi.split * p.batch + b) * p.heads_q + h) * p.seqlen_q + gq] = lse;
            else if (p.cu_q != nullptr) p.lse[(int64_t)h * p.total_q + fi.qoff + gq] = lse;   // [Hq, total_q]
            else p.lse[((int64_t)b * p.heads_q + h) * p.lse_bh_stride + gq] = lse;
          }
        }
        ptx::mbar_wait(bar(bars.p_empty[gl % KS]), (gl / KS) & 1);
        ptx::tc_fence_after();
        if (threadIdx.x == Cfg::EPI_WARP0 * 32) FFPA_STAMP(kidx, 5);
        // 32 head dims (16 packed registers) of this row -> global
        auto store32 = [&](const uint32_t* w, int d0) {
          if (!row_ok) return;
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int d = d0 + 16 * v;
            if (o_al32 && d + 16 <= p.head_dim) {
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + 2 * d), "r"(w[8 * v]), "r"(w[8 * v + 1]),
                           "r"(w[8 * v + 2]), "r"(w[8 * v + 3]), "r"(w[8 * v + 4]), "r"(w[8 * v + 5]), "r"(w[8 * v + 6]), "r"(w[8 * v + 7])
                           : "memory");
            } else {
              if (d < p.head_dim) *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(w[8 * v], w[8 * v + 1], w[8 * v + 2], w[8 * v + 3]);
              if (d + 8 < p.head_dim)
                *reinterpret_cast<uint4*>(orow + 2 * d + 16) = make_uint4(w[8 * v + 4], w[8 * v + 5], w[8 * v + 6], w[8 * v + 7]);
            }
          }
        };
        auto release_o = [&]() {
          ptx::tc_fence_before();
          __syncwarp();
          if (ptx::lane_id() == 0) ptx::mbar_arrive_cluster(l_o_free);
          if (threadIdx.x == Cfg::EPI_WARP0 * 32) FFPA_STAMP(kidx, 8);
        };
        constexpr uint32_t kPark = Cfg::PARK_BASE >= 0 ? (uint32_t)Cfg::PARK_BASE : 0u;
        if (Cfg::PARK_BASE >= 0 && p.kv_splits == 1) {
          // phase 1: park the normalised 16-bit tile in the spare TMEM columns, release O
#pragma unroll
          for (int s = 0; s < Cfg::NSLICE; ++s) {
            if (256 * s >= dvw) break;
            const int half = Cfg::slice_n(dvw, s) / 2;   // columns of this slice per lane: 128 or 64
#pragma unroll 1
            for (int c = 0; c < half; c += 32) {
              uint32_t orr[32], w[16];
              ptx::tmem_ld_x32(tmem + lane_base + 128 * s + c, orr);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int u = 0; u < 16; ++u) {
                const float a = __uint_as_float(orr[2 * u]) * inv, c2 = __uint_as_float(orr[2 * u + 1]) * inv;
                w[u] = BF16 ? ptx::pack_bf16x2(a, c2) : ptx::pack_f16x2(a, c2);
              }
              ptx::tmem_st_x16(tmem + lane_base + kPark + (128 * s + c) / 2, w);
            }
          }
          ptx::tmem_wait_st();
          release_o();
          // phase 2: parked tile -> global
#pragma unroll
          for (int s = 0; s < Cfg::NSLICE; ++s) {
            if (256 * s >= dvw) break;
            const int half = Cfg::slice_n(dvw, s) / 2;
#pragma unroll 1
            for (int c = 0; c < half; c += 64) {
              uint32_t w[32];
              ptx::tmem_ld_x32(tmem + lane_base + kPark + (128 * s + c) / 2, w);
              ptx::tmem_wait_ld();
              store32(w, dv0 + 256 * s + half * (int)kh + c);
              store32(w + 16, dv0 + 256 * s + half * (int)kh + c + 32);
            }
          }
          ptx::tc_fence_before();   // the parked copy is rewritten by the next item's phase 1 (same thread, program order)
        } else {
          // no spare TMEM columns (or fp32 partials of a KV split): O is released after its last load
#pragma unroll
          for (int s = 0; s < Cfg::NSLICE; ++s) {
            if (256 * s >= dvw) break;
            const int half = Cfg::slice_n(dvw, s) / 2;
            const bool last_slice = (256 * (s + 1) >= dvw);
#pragma unroll 1
            for (int c = 0; c < half; c += 32) {
              uint32_t orr[32];
              ptx::tmem_ld_x32(tmem + lane_base + 128 * s + c, orr);
              ptx::tmem_wait_ld();
              if (last_slice && c + 32 >= half) release_o();
              const int d0 = dv0 + 256 * s + half * (int)kh + c;
              if (p.kv_splits > 1) {
                // fp32 partial of this KV split, normalised by its own row sum
                if (row_ok) {
                  float* po = p.part_o + ((((int64_t)fi.split * p.batch + b) * p.heads_q + h) * p.seqlen_q + gq) * (int64_t)p.head_dim;
#pragma unroll
                  for (int v = 0; v < 8; ++v)
                    if (d0 + 4 * v < p.head_dim)
                      *reinterpret_cast<float4*>(po + d0 + 4 * v) =
                          make_float4(__uint_as_float(orr[4 * v]) * inv, __uint_as_float(orr[4 * v + 1]) * inv,
                                      __uint_as_float(orr[4 * v + 2]) * inv, __uint_as_float(orr[4 * v + 3]) * inv);
                }
              } else {
                uint32_t w[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                  const float a = __uint_as_float(orr[2 * u]) * inv, c2 = __uint_as_float(orr[2 * u + 1]) * inv;
                  w[u] = BF16 ? ptx::pack_bf16x2(a, c2) : ptx::pack_f16x2(a, c2);
                }
                store32(w, d0);
              }
            }
          }
        }
        if (threadIdx.x == Cfg::EPI_WARP0 * 32) FFPA_STAMP(kidx, 6);
        ++it;
      }
    }
  } else if (warp >= (uint32_t)kSoftmaxWarps) {
  if constexpr (Cfg::EPIW) ptx::setmaxnreg_dec<88>();
  if (warp == kTmaWarp) {
    // =========================================== TMA producer: Q and K (and V on the shared ring) ===========
    // Separate rings (head dims <= 512): this thread streams Q and K only and the V warp below streams V, so neither
    // stream ever queues behind a ring slot of the other. With ONE thread doing both, the next item's Q and first K
    // tile were requested only after the last V tile of the current item had found a free slot (= after its
    // second-to-last PV MMA), and every item started with ~3000 idle tensor-pipe cycles (profiles/r02_fwd_epilogue.md).
    // Shared ring (wider heads): the stream order is fixed by the ring, but the next item's Q is requested before the
    // last V tile, whose slot cannot free before Q's own (q_empty fires when the last QK MMA completes).
    if (ptx::elect_one()) {
      uint32_t kc = 0, it = 0;
      bool q_ahead = false;
      const uint32_t l_q_full = ptx::mapa(bar(bars.q_full), 0);
      auto load_q = [&](const FwdItem& f, uint32_t itn) {
        ptx::mbar_wait(bar(bars.q_empty), (itn & 1) ^ 1);
        if (rank == 0) ptx::mbar_expect_tx(bar(bars.q_full), 2 * Cfg::Q_BYTES);
#pragma unroll
        for (int jb = 0; jb < NQK; ++jb)
          ptx::tma_load_4d_2sm(sQ + jb * 8192, &map_q, l_q_full, jb * 64, f.qoff + f.mt * 128 + 64 * (int)rank, f.bh % p.heads_q, f.bt);
      };
      for (uint32_t kidx = 0;; ++kidx, ++it) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const uint32_t item = (uint32_t)item_s;
        const FwdItem fi = decode_fwd_item<Cfg::NPASS>(p, item);
        const int pass = fi.pass, bh = fi.bh;
        const int h = bh % p.heads_q, b = fi.bt;   // b: batch coordinate of the tensor maps
        const int hk = h / group;
        const int T = fi.T, tbeg = fi.tbeg;
        if (T <= 0) { --it; continue; }   // empty KV split: no barrier traffic (the for-increment re-adds 1)
        const int dv0 = pass * Cfg::DSLAB, dvw = Cfg::slab_w(pass);
        if (q_ahead) q_ahead = false;
        else load_q(fi, it);
        FFPA_STAMP(kidx, 7);
        for (int step = 0; step < T + (Cfg::UNIFIED ? LA : 0); ++step) {
          if (step < T) {
            const int kv0 = fi.koff + (tbeg + step) * 128;
#pragma unroll
            for (int ks = 0; ks < Cfg::KST; ++ks) {
              const uint32_t stage = kc % Cfg::NKS, n = kc / Cfg::NKS;
              ptx::mbar_wait(bar(bars.k_empty[stage]), (n & 1) ^ 1);
              const int nb = (NQK - 2 * ks) >= 2 ? 2 : 1;
              if (rank == 0) ptx::mbar_expect_tx(bar(bars.k_full[stage]), 2 * nb * 8192);
              const uint32_t l_full = ptx::mapa(bar(bars.k_full[stage]), 0);
              for (int bx = 0; bx < nb; ++bx)
                ptx::tma_load_4d_2sm(sK + stage * 16384 + bx * 8192, &map_k, l_full, (2 * ks + bx) * 64,
                                     kv0 + 64 * (int)rank, hk, b);
              ++kc;
            }
            if constexpr (Cfg::K_DUMMY) {  // keep V slices on even ring stages
              const uint32_t st = kc % Cfg::NKS, n = kc / Cfg::NKS;
              ptx::mbar_wait(bar(bars.k_empty[st]), (n & 1) ^ 1);
              if (rank == 0) ptx::mbar_arrive(bar(bars.k_full[st]));
              ++kc;
            }
          }
          if constexpr (Cfg::UNIFIED) {
            if (step == T) {
              // first drain step: request the next non-empty item's Q before the last V tiles
              for (uint32_t kn = kidx + 1;; ++kn) {
                const int nx = next_item(p, cluster, nclusters, kn);
                if (nx < 0) break;
                const FwdItem fn = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)nx);
                if (fn.T <= 0) continue;
                load_q(fn, it + 1);
                q_ahead = true;
                break;
              }
            }
            if (step >= LA) {
              const int kv0 = fi.koff + (tbeg + step - LA) * 128;
#pragma unroll
              for (int s = 0; s < Cfg::NSLICE; ++s) {
                if (256 * s >= dvw) break;
                // slice = stages (st, st+1) of the shared ring; kc is the shared counter
                const uint32_t st = kc % Cfg::NKS, n = kc / Cfg::NKS;
                const int nsu = Cfg::slice_n(dvw, s);
                for (int bx = 0; bx < 2; ++bx) {
                  ptx::mbar_wait(bar(bars.k_empty[st + bx]), (n & 1) ^ 1);
                  if (bx < nsu / 128) {
                    if (rank == 0) ptx::mbar_expect_tx(bar(bars.k_full[st + bx]), 2 * 16384);
                    ptx::tma_load_4d_2sm(sK + (st + bx) * 16384, &map_v, ptx::mapa(bar(bars.k_full[st + bx]), 0),
                                         dv0 + 256 * s + (nsu / 2) * (int)rank + 64 * bx, kv0, hk, b);
                  } else if (rank == 0) {
                    ptx::mbar_arrive(bar(bars.k_full[st + bx]));  // unused half of a 128-wide slice
                  }
                }
                kc += 2;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (!Cfg::UNIFIED && warp == Cfg::V_WARP) {
    // =========================================== TMA producer: V (separate rings) ===========================
    if (ptx::elect_one()) {
      uint32_t vc = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const FwdItem fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)item_s);
        if (fi.T <= 0) continue;
        const int hk = (fi.bh % p.heads_q) / group, b = fi.bt;
        const int dv0 = fi.pass * Cfg::DSLAB, dvw = Cfg::slab_w(fi.pass);
        for (int step = 0; step < fi.T; ++step) {
          const int kv0 = fi.koff + (fi.tbeg + step) * 128;
#pragma unroll
          for (int s = 0; s < Cfg::NSLICE; ++s) {
            if (256 * s >= dvw) break;
            const uint32_t stage = vc % Cfg::NVS, n = vc / Cfg::NVS;
            ptx::mbar_wait(bar(bars.v_empty[stage]), (n & 1) ^ 1);
            const int ns = Cfg::slice_n(dvw, s);
            const int nb = ns / 128;  // 64-wide boxes this CTA loads
            if (rank == 0) ptx::mbar_expect_tx(bar(bars.v_full[stage]), 2 * nb * 16384);
            const uint32_t l_full = ptx::mapa(bar(bars.v_full[stage]), 0);
            for (int bx = 0; bx < nb; ++bx)
              ptx::tma_load_4d_2sm(sV + stage * 32768 + bx * 16384, &map_v, l_full,
                                   dv0 + 256 * s + (ns / 2) * (int)rank + 64 * bx, kv0, hk, b);
            ++vc;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // =========================================== MMA issuer (leader CTA) ========================
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t fmt = BF16 ? 1u : 0u;
      constexpr uint32_t idesc_qk = ptx::make_idesc(fmt, fmt, 0, 0, 128, 128);
      uint32_t kc = 0, vc = 0, it = 0, g = 0, gp = 0;
      for (uint32_t kidx = 0;; ++kidx, ++it) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const uint32_t item = (uint32_t)item_s;
        const FwdItem fi = decode_fwd_item<Cfg::NPASS>(p, item);
        const int dvw = Cfg::slab_w(fi.pass);
        const int T = fi.T;
        if (T <= 0) { --it; continue; }
        ptx::mbar_wait(bar(bars.q_full), it & 1);
        ptx::tc_fence_after();
        FFPA_STAMP(kidx, 0);
        for (int step = 0; step < T + LA; ++step) {
          if (step < T) {
            const uint32_t sbuf = g % KS;
            const uint32_t d_tmem = tmem + Cfg::S_BASE + 64 * sbuf;
#pragma unroll
            for (int ks = 0; ks < Cfg::KST; ++ks) {
              const uint32_t stage = kc % Cfg::NKS, n = kc / Cfg::NKS;
              ptx::mbar_wait(bar(bars.k_full[stage]), n & 1);
              ptx::tc_fence_after();
              const int nb = (NQK - 2 * ks) >= 2 ? 2 : 1;
              for (int bx = 0; bx < nb; ++bx) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                  const uint64_t ad = ptx::make_smem_desc_sw128(sQ + (2 * ks + bx) * 8192 + k4 * 32, 16, 1024);
                  const uint64_t bd = ptx::make_smem_desc_sw128(sK + stage * 16384 + bx * 8192 + k4 * 32, 16, 1024);
                  ptx::umma_f16_ss<CG>(d_tmem, ad, bd, idesc_qk, (ks | bx | k4) != 0 ? 1u : 0u);
                }
              }
              ptx::umma_commit_mc<CG>(bar(bars.k_empty[stage]), 0x3);
              ++kc;
            }
            if constexpr (Cfg::K_DUMMY) {
              const uint32_t st = kc % Cfg::NKS, n = kc / Cfg::NKS;
              ptx::mbar_wait(bar(bars.k_full[st]), n & 1);
              ptx::umma_commit_mc<CG>(bar(bars.k_empty[st]), 0x3);
              ++kc;
            }
            ptx::umma_commit_mc<CG>(bar(bars.s_full[sbuf]), 0x3);
            if (step == T - 1) ptx::umma_commit_mc<CG>(bar(bars.q_empty), 0x3);
            ++g;
          }
          if (step >= LA) {
            const uint32_t pbuf = gp % KS;
            if constexpr (Cfg::EPIW) {
              // the first PV MMA of an item overwrites O: the epilogue warps must have read the previous item's tile
              if (step == LA) ptx::mbar_wait_cluster(bar(bars.o_free), (it & 1) ^ 1);
            }
            ptx::mbar_wait_cluster(bar(bars.p_full[pbuf]), (gp / KS) & 1);
            ptx::tc_fence_after();
            if (step == LA) FFPA_STAMP(kidx, 1);
            if (step == T + LA - 1) FFPA_STAMP(kidx, 2);
#pragma unroll
            for (int s = 0; s < Cfg::NSLICE; ++s) {
              if (256 * s >= dvw) break;
              if constexpr (Cfg::UNIFIED) {
                const uint32_t st = kc % Cfg::NKS, n = kc / Cfg::NKS;
                ptx::mbar_wait(bar(bars.k_full[st]), n & 1);
                ptx::mbar_wait(bar(bars.k_full[st + 1]), n & 1);
                ptx::tc_fence_after();
                const uint32_t idesc_pv = ptx::make_idesc(fmt, fmt, 0, 1, 128, Cfg::slice_n(dvw, s));
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                  const uint64_t ad = ptx::make_smem_desc_sw128(sP + pbuf * 16384 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                  const uint64_t bd = ptx::make_smem_desc_sw128(sK + st * 16384 + kk * 2048, 16384, 1024);
                  ptx::umma_f16_ss<CG>(tmem + 128 * s, ad, bd, idesc_pv, (step > LA || kk > 0) ? 1u : 0u);
                }
                ptx::umma_commit_mc<CG>(bar(bars.k_empty[st]), 0x3);
                ptx::umma_commit_mc<CG>(bar(bars.k_empty[st + 1]), 0x3);
                kc += 2;
                continue;
              }
              const uint32_t stage = vc % Cfg::NVS, n = vc / Cfg::NVS;
              ptx::mbar_wait(bar(bars.v_full[stage]), n & 1);
              ptx::tc_fence_after();
              const int ns = Cfg::slice_n(dvw, s);
              const uint32_t idesc_pv = ptx::make_idesc(fmt, fmt, 0, 1, 128, ns);
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t ad = ptx::make_smem_desc_sw128(sP + pbuf * 16384 + (kk >> 2) * 8192 + (kk & 3) * 32, 16, 1024);
                const uint64_t bd = ptx::make_smem_desc_sw128(sV + stage * 32768 + kk * 2048, 16384, 1024);
                ptx::umma_f16_ss<CG>(tmem + 128 * s, ad, bd, idesc_pv, (step > LA || kk > 0) ? 1u : 0u);
              }
              ptx::umma_commit_mc<CG>(bar(bars.v_empty[stage]), 0x3);
              ++vc;
            }
            ptx::umma_commit_mc<CG>(bar(bars.p_empty[pbuf]), 0x3);
            ++gp;
          }
        }
      }
    }
    __syncwarp();
  } else if (Cfg::NPASS == 2 && warp == Cfg::STORE_WARP) {
    // =========================================== P store warp (replay path) =====================
    if (p.stash_p != nullptr && ptx::elect_one()) {
      ptx::prefetch_tmap(&map_sp);
      const uint64_t pol = ptx::l2_policy_evict_first();
      uint32_t g = 0;
      for (uint32_t kidx = 0;; ++kidx) {
        const int item_s = next_item(p, cluster, nclusters, kidx);
        if (item_s < 0) break;
        const FwdItem fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)item_s);
        if (fi.T <= 0) continue;
        for (int i = 0; i < fi.T; ++i, ++g) {
          const uint32_t sbuf = g % KS;
          const int blk = fi.mt * (p.nk_pad >> 6) + 2 * (fi.tbeg + i);
          ptx::mbar_wait(bar(bars.p_written[sbuf]), (g / KS) & 1);
          ptx::tma_store_4d_hint(&map_sp, sP + sbuf * 16384, 0, 64 * (int)rank, blk, fi.bh, pol);
          ptx::tma_store_4d_hint(&map_sp, sP + sbuf * 16384 + 8192, 0, 64 * (int)rank, blk + 1, fi.bh, pol);
          ptx::bulk_commit_group();
          ptx::bulk_wait_group_read0();
          ptx::mbar_arrive(bar(bars.p_stored[sbuf]));
        }
      }
      ptx::bulk_wait_group0();
    }
    __syncwarp();
  }
  } else {
    // =========================================== softmax / correction / epilogue ================
    if constexpr (Cfg::EPIW) ptx::setmaxnreg_inc<168>();
    // NSW*32 threads: TMEM lane = t % 128; warpgroup ch = t / 128 owns S columns [CPT*ch, CPT*ch+CPT).
    const uint32_t t = threadIdx.x;
    const uint32_t lane128 = t & 127;
    const uint32_t row = lane128 & 63;        // row inside this CTA's 64
    const uint32_t kh = lane128 >> 6;         // 64-key half of the KV tile / column half of O
    const uint32_t wgi = t >> 7;              // warpgroup index
    const uint32_t ch = ALT ? 0u : wgi;       // column group inside the S stage (column-split mode)
    const uint32_t slot = ALT ? (wgi * 2 + kh) : (kh * CQ + ch);
    const uint32_t rgrp = warp & 1;           // warps sharing rows: {0,2,4,6} / {1,3,5,7}
    const uint32_t lane_base = ((warp & 3) * 32u) << 16;
    const uint32_t l_p_full0 = ptx::mapa(bar(bars.p_full[0]), 0);
    const float NEG_INF = -INFINITY;
    // softmax domain: fast mode works on raw scores (mul = scale*log2e), general on scaled+biased
    const float mul = (MODE == kModeFast) ? p.scale_log2 : 1.0f;
    uint32_t g = 0;                      // global tile counter at the start of the item (ALT) / running (column split)
    uint32_t uw = 0;                     // ALT: tiles processed by this warpgroup
    uint32_t pub[4] = {0, 0, 0, 0};      // ALT: completed publications of m_full[0..3]
    uint32_t itn = 0;                    // non-empty items finished (phase of the epilogue-warp barriers)
    bool epi_pending = false;            // a bulk store of the last epilogue may still be reading the P ring
    // the next item is looked up and decoded (schedule-table load, integer divisions: ~1000 cycles) while this item's
    // last PV MMA drains, not between the epilogue and the first softmax of the next item where nothing hides it
    int ahead_s = next_item(p, cluster, nclusters, 0);
    FwdItem ahead_fi{};
    if (ahead_s >= 0) ahead_fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)ahead_s);
    for (uint32_t kidx = 0;; ++kidx) {
      const int item_s = ahead_s;
      if (item_s < 0) break;
      const FwdItem fi = ahead_fi;
      const int mt = fi.mt, pass = fi.pass, bh = fi.bh;
      const int h = bh % p.heads_q, b = bh / p.heads_q;
      const int q0 = mt * 128;
      const int T = fi.T, tbeg = fi.tbeg;
      const int dv0 = pass * Cfg::DSLAB, dvw = Cfg::slab_w(pass);
      const int gq = q0 + 64 * (int)rank + (int)row;  // query row of this thread inside its sequence
      const int seq_q = fi.nq, seq_kv = fi.nkv;       // lengths of this item's sequence
      const int causal_lim = gq + (seq_kv - seq_q);  // last visible key when causal
      float m = NEG_INF, l = 0.f;   // ALT: m = the max this thread's partial sum l is expressed against
      if (T <= 0) {
        // empty KV split (causal rows that end before this split starts): contributes nothing
        if (p.part_lse != nullptr && wgi == 0 && kh == 0 && gq < seq_q)
          p.part_lse[(((int64_t)fi.split * p.batch + b) * p.heads_q + h) * p.seqlen_q + gq] = NEG_INF;
        ahead_s = next_item(p, cluster, nclusters, kidx + 1);
        if (ahead_s >= 0) ahead_fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)ahead_s);
        continue;
      }

      for (int i = ALT ? (int)wgi : 0; i < T; i += ALT ? 2 : 1) {
        const uint32_t gi = ALT ? g + (uint32_t)i : g;   // global index of this tile
        const uint32_t sbuf = gi % KS;   // S / P pipeline stage
        const uint32_t xb = ALT ? (uw & 1) : (gi & 1);   // row-max exchange buffer
        ptx::mbar_wait(bar(bars.s_full[sbuf]), (gi / KS) & 1);
        ptx::tc_fence_after();
        if (t == 0 && i == 0) FFPA_STAMP(kidx, 3);
        uint32_t sr[CPT];
        if constexpr (CPT == 64) {
          ptx::tmem_ld_x32(tmem + lane_base + Cfg::S_BASE + 64 * sbuf, sr);
          ptx::tmem_ld_x32(tmem + lane_base + Cfg::S_BASE + 64 * sbuf + 32, sr + 32);
        } else if constexpr (CPT == 32) ptx::tmem_ld_x32(tmem + lane_base + Cfg::S_BASE + 64 * sbuf + CPT * ch, sr);
        else ptx::tmem_ld_x16(tmem + lane_base + Cfg::S_BASE + 64 * sbuf + CPT * ch, sr);
        ptx::tmem_wait_ld();
        float x[CPT];
        const int ti = tbeg + i;   // absolute KV tile index
        const int key0 = ti * 128 + 64 * (int)kh + CPT * (int)ch;
        if constexpr (MODE == kModeFast) {
#pragma unroll
          for (int j = 0; j < CPT; ++j) x[j] = __uint_as_float(sr[j]);
        } else {
          if (p.bias_kind == 0) {
#pragma unroll
            for (int j = 0; j < CPT; ++j) x[j] = __uint_as_float(sr[j]) * p.scale_log2;
          } else {
            const int64_t boff = (int64_t)b * p.bias_stride[0] + (int64_t)h * p.bias_stride[1] +
                                 (int64_t)(gq < seq_q ? gq : 0) * p.bias_stride[2];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
              const int key = key0 + j;
              float bv = 0.f;
              if (key < seq_kv) {
                const int64_t bi = boff + (int64_t)key * p.bias_stride[3];   // k-stride 1, or 0 for a [.., 1] bias
                if (p.bias_kind == 1) bv = reinterpret_cast<const float*>(p.bias)[bi];
                else if (BF16) bv = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.bias)[bi]);
                else bv = __half2float(reinterpret_cast<const __half*>(p.bias)[bi]);
              }
              x[j] = fmaf(__uint_as_float(sr[j]), p.scale_log2, bv * 1.4426950408889634f);
            }
          }
        }
        const bool tail = (ti * 128 + 128 > seq_kv);
        const bool diag = p.causal && (ti * 128 + 127 > q0 + (seq_kv - seq_q));
        if (tail || diag) {
          const int lim = p.causal ? (causal_lim < seq_kv - 1 ? causal_lim : seq_kv - 1) : seq_kv - 1;
#pragma unroll
          for (int j = 0; j < CPT; ++j)
            if (key0 + j > lim) x[j] = NEG_INF;
        }
        // local max: independent chains of 3-input max
        float mx0 = fmax3(x[0], x[1], x[2]), mx1 = fmax3(x[3], x[4], x[5]);
        float mx2 = fmax3(x[6], x[7], x[8]), mx3 = fmax3(x[9], x[10], x[11]);
        mx0 = fmax3(mx0, x[12], x[13]); mx1 = fmax3(mx1, x[14], x[15]);
        if constexpr (CPT >= 32) {
          mx2 = fmax3(mx2, x[16], x[17]); mx3 = fmax3(mx3, x[18], x[19]);
          mx0 = fmax3(mx0, x[20], x[21]); mx1 = fmax3(mx1, x[22], x[23]);
          mx2 = fmax3(mx2, x[24], x[25]); mx3 = fmax3(mx3, x[26], x[27]);
          mx0 = fmax3(mx0, x[28], x[29]); mx1 = fmax3(mx1, x[30], x[31]);
        }
        if constexpr (CPT == 64) {
#pragma unroll
          for (int j = 32; j < 64; j += 8) {
            mx0 = fmax3(mx0, x[j], x[j + 1]); mx1 = fmax3(mx1, x[j + 2], x[j + 3]);
            mx2 = fmax3(mx2, x[j + 4], x[j + 5]); mx3 = fmax3(mx3, x[j + 6], x[j + 7]);
          }
        }
        float tmax = fmaxf(fmax3(mx0, mx1, mx2), mx3);
        xch[xb][slot][row] = tmax;
        if (epi_pending) {
          // O tiles of the previous item were staged in the P ring (epilogue): their bulk stores must have read them
          // before any warp that meets this one at the barrier below writes P
          if (ptx::lane_id() == 0) ptx::bulk_wait_group_read0();
          epi_pending = false;
        }
        if constexpr (ALT) {
          ptx::named_bar_sync(1 + 2 * wgi + rgrp, 64);      // the two lane halves of this warpgroup's rows
          tmax = fmaxf(tmax, xch[xb][slot ^ 1][row]);
          // running max after tile i-1, published by the other warpgroup
          if (i > 0) {
            const uint32_t par = (uint32_t)(i - 1) & 3u;
            const uint32_t cnt = pub[par] + (uint32_t)((i - 1) >> 2);
            ptx::mbar_wait(bar(bars.m_full[par]), cnt & 1);
          }
        } else {
          ptx::named_bar_sync(1 + rgrp, kSoftmaxWarps * 16);
          tmax = fmaxf(fmax3(xch[xb][0][row], xch[xb][1][row], xch[xb][2][row]), xch[xb][3][row]);
          if constexpr (CQ == 4)
            tmax = fmaxf(tmax, fmaxf(fmax3(xch[xb][4][row], xch[xb][5][row], xch[xb][6][row]), xch[xb][7][row]));
        }
        // ALT: m_run = running max after the previous tile (from the other warpgroup); else the local state
        const float m_run = ALT ? ((i > 0) ? mval[(i - 1) & 3][row] : NEG_INF) : m;
        // lazy rescale: keep the stale max while the true max is < 8 (log2 units) above it
        // (/root/reference/csrc/cuffpa/native/prefill.cuh:719-738, common.cuh:14-18)
        const float m_new = fmaxf(m_run, tmax);
        const bool upd = (m_new - m_run) * mul > 8.0f;  // true for -inf -> finite; false for -inf -> -inf
        const float m_use = upd ? m_new : m_run;
        const bool need_rescale = upd && (m_run != NEG_INF);
        if constexpr (ALT) {
          if (kh == 0) mval[i & 3][row] = m_use;
          __syncwarp();
          if (kh == 0 && ptx::lane_id() == 0) ptx::mbar_arrive(bar(bars.m_full[i & 3]));
        }
        const float m_safe = (m_use == NEG_INF) ? 0.f : m_use;
        const float neg_mc = -m_safe * mul;
        float factor = 1.f;
        if (need_rescale) factor = exp2f((m_run - m_use) * mul);
        uint32_t pk[CPT / 2];
        float lsum;
        if constexpr (MODE == kModeDropout) {
          // keep iff u > p; the row sum uses the un-dropped probabilities
          // (/root/reference/csrc/cuffpa/native/prefill.cuh:506-546)
          const float inv_keep = 1.f / (1.f - p.dropout_p);
          const uint64_t ebase = p.philox_offset +
              ((uint64_t)((int64_t)b * p.heads_q + h) * (uint64_t)seq_q + (uint64_t)(gq < seq_q ? gq : 0)) * (uint64_t)seq_kv +
              (uint64_t)key0;
          lsum = 0.f;
#pragma unroll
          for (int j2 = 0; j2 < CPT; j2 += 2) {
            float pv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float e = exp2f(fmaf(x[j2 + u], mul, neg_mc));
              lsum += e;
              const uint64_t eo = ebase + (uint64_t)(j2 + u);
              const uint4 r4 = philox4x32_10(p.philox_seed, eo >> 2);
              const uint32_t sel = (uint32_t)(eo & 3);
              const uint32_t rv = sel == 0 ? r4.x : sel == 1 ? r4.y : sel == 2 ? r4.z : r4.w;
              const float uni = ((float)rv + 1.0f) * 2.3283064365386963e-10f;
              pv[u] = (uni > p.dropout_p) ? e * inv_keep : 0.f;
            }
            pk[j2 >> 1] = BF16 ? ptx::pack_bf16x2(pv[0], pv[1]) : ptx::pack_f16x2(pv[0], pv[1]);
          }
        } else {
          const float2 mul2 = make_float2(mul, mul), nm2 = make_float2(neg_mc, neg_mc);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < CPT; j += 4) {
            const float2 a0 = ffma2(make_float2(x[j], x[j + 1]), mul2, nm2);
            const float2 a1 = ffma2(make_float2(x[j + 2], x[j + 3]), mul2, nm2);
            const float2 e0 = make_float2(exp2f(a0.x), exp2f(a0.y));
            const float2 e1 = make_float2(exp2f(a1.x), exp2f(a1.y));
            acc0 = fadd2(acc0, e0);
            acc1 = fadd2(acc1, e1);
            pk[j >> 1] = BF16 ? ptx::pack_bf16x2(e0.x, e0.y) : ptx::pack_f16x2(e0.x, e0.y);
            pk[(j >> 1) + 1] = BF16 ? ptx::pack_bf16x2(e1.x, e1.y) : ptx::pack_f16x2(e1.x, e1.y);
          }
          acc0 = fadd2(acc0, acc1);
          lsum = acc0.x + acc0.y;
        }
        if constexpr (ALT) {
          // this thread's partial row sum, re-expressed against the max now in effect
          if (m != m_use) l *= (m == NEG_INF) ? 0.f : exp2f((m - m_use) * mul);
          l += lsum;
        } else {
          l = l * factor + lsum;
        }
        m = m_use;

        // P buffer free? (PV of tile g-2 retired)
        ptx::mbar_wait(bar(bars.p_empty[sbuf]), ((gi / KS) & 1) ^ 1);
        const bool stash = Cfg::NPASS == 2 && p.stash_p != nullptr;
        if (stash) {
          ptx::mbar_wait(bar(bars.p_stored[sbuf]), ((gi / KS) & 1) ^ 1);   // ... and drained by the store warp
          // O rescale factor of this (row, tile): what the second-slab GEMM replays (1 on most tiles)
          if (kh == 0 && ch == 0)
            p.stash_f[(((int64_t)bh * p.n_mt_even + mt) * (p.nk_pad >> 7) + ti) * 128 + 64 * (int)rank + (int)row] = factor;
        }
        {
          const uint32_t prow = sP + sbuf * 16384 + kh * 8192 + row * 128;
#pragma unroll
          for (int c = 0; c < CPT / 8; ++c) {
            const uint32_t addr = prow + ((((CPT / 8) * ch + c) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]),
                         "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3])
                         : "memory");
          }
        }
        if (__any_sync(0xffffffffu, need_rescale)) {
          // O may only be touched once PV of tile g-1 has retired; each warpgroup scales half the columns
          ptx::mbar_wait(bar(bars.p_empty[(gi - 1) % KS]), ((gi - 1) / KS) & 1);
          ptx::tc_fence_after();
          constexpr int RC = CPT == 16 ? 16 : 32;   // columns per TMEM round trip
#pragma unroll 1
          for (int c0 = (int)ch * (dvw / 2 / CQ); c0 < (int)(ch + 1) * (dvw / 2 / CQ); c0 += RC) {
            uint32_t orr[RC];
            if constexpr (RC == 32) ptx::tmem_ld_x32(tmem + lane_base + c0, orr);
            else ptx::tmem_ld_x16(tmem + lane_base + c0, orr);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < RC; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * factor);
            if constexpr (RC == 32) ptx::tmem_st_x32(tmem + lane_base + c0, orr);
            else ptx::tmem_st_x16(tmem + lane_base + c0, orr);
          }
          ptx::tmem_wait_st();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (t == 0 && i == 0) FFPA_STAMP(kidx, 4);
        if (ptx::lane_id() == 0) {
          ptx::mbar_arrive_cluster(l_p_full0 + 8u * sbuf);  // p_full[] is contiguous
          if (stash) ptx::mbar_arrive(bar(bars.p_written[sbuf]));
        }
        if constexpr (ALT) ++uw; else ++g;
      }

      // ---------------- epilogue: O / l -> global, LSE ----------------
      ahead_s = next_item(p, cluster, nclusters, kidx + 1);
      if (ahead_s >= 0) ahead_fi = decode_fwd_item<Cfg::NPASS>(p, (uint32_t)ahead_s);
      {
        float l_tot;
        uint32_t gl;
        if constexpr (ALT) {
          // final running max: published with the last tile
          const uint32_t parl = (uint32_t)(T - 1) & 3u;
          const uint32_t cntl = pub[parl] + (uint32_t)((T - 1) >> 2);
          ptx::mbar_wait(bar(bars.m_full[parl]), cntl & 1);
          const float m_fin = mval[parl][row];
#pragma unroll
          for (int j = 0; j < 4; ++j) pub[j] += (T > j) ? (uint32_t)((T - 1 - j) / 4 + 1) : 0u;
          const float l_fin = (m == NEG_INF) ? 0.f : l * exp2f((m - m_fin) * mul);
          m = m_fin;
          gl = g + (uint32_t)T - 1;
          g += (uint32_t)T;
          // the exchange buffer of parity 0/1 may still be read by a slower partner of the last tiles:
          // a dedicated 128-thread barrier pair brackets its reuse for the row sums
          ptx::named_bar_sync(5 + rgrp, 128);
          xch[0][slot][row] = l_fin;
          ptx::named_bar_sync(5 + rgrp, 128);
          l_tot = (xch[0][0][row] + xch[0][1][row]) + (xch[0][2][row] + xch[0][3][row]);
          ptx::named_bar_sync(5 + rgrp, 128);
          if constexpr (!Cfg::EPIW) {
            ptx::mbar_wait(bar(bars.p_empty[gl % KS]), (gl / KS) & 1);
            ptx::tc_fence_after();
          }
        } else {
        gl = g - 1;
        // row-sum exchange reuses the max-exchange buffer of the last tile: every thread of the row group must have
        // finished reading it. Without epilogue warps the wait for PV(gl) (needed to read O anyway) implies that
        // (p_full precedes p_empty); with them the softmax warps do not wait for the MMA: one more named barrier.
        if constexpr (Cfg::EPIW) {
          ptx::named_bar_sync(1 + rgrp, kSoftmaxWarps * 16);
        } else {
          ptx::mbar_wait(bar(bars.p_empty[gl % KS]), (gl / KS) & 1);
          ptx::tc_fence_after();
        }
        float (*xl)[64] = xch[gl & 1];
        xl[slot][row] = l;
        ptx::named_bar_sync(1 + rgrp, kSoftmaxWarps * 16);
        l_tot = (xl[0][row] + xl[1][row]) + (xl[2][row] + xl[3][row]);
        if constexpr (CQ == 4) l_tot += (xl[4][row] + xl[5][row]) + (xl[6][row] + xl[7][row]);
        }
        if (t == 0) FFPA_STAMP(kidx, 5);
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        const bool row_ok = gq < seq_q;
        if (!Cfg::EPIW && Cfg::NPASS == 2 && p.stash_p != nullptr && wgi == 0 && kh == 0)
          p.stash_inv[(int64_t)bh * p.n_mt_even * 128 + gq] = inv;
        if constexpr (Cfg::EPIW) {
          // hand the tile to the epilogue warps: 1 / row sum and LSE per row (one writer per row: warps 0 and 1)
          if (wgi == 0 && kh == 0) {
            ptx::mbar_wait(bar(bars.inv_taken), (itn & 1) ^ 1);   // the previous item's values have been read
            epi_inv[row] = inv;
            // natural-log LSE; rows without any visible key: O = 0, LSE = -inf
            epi_lse[row] = (l_tot > 0.f) ? (m * mul + log2f(l_tot)) * 0.6931471805599453f : NEG_INF;
            __syncwarp();
            if (ptx::lane_id() == 0) ptx::mbar_arrive(bar(bars.o_ready));
          }
          ++itn;
        }
        if constexpr (!Cfg::EPIW) {
        uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                        2 * ((int64_t)fi.bt * p.o_stride[0] + (int64_t)h * p.o_stride[1] + (int64_t)(fi.qoff + gq) * p.o_stride[2]);
        const bool o_al32 = (reinterpret_cast<uintptr_t>(orow) & 31u) == 0;   // 32-byte stores need it (strides are only 16-byte granular)
        // TMA-store path. In the direct path every lane writes another row, so each store instruction costs 32 LSU
        // wavefronts and the epilogue of a 128 x 512 tile keeps the LSU busy for ~4000 cycles -- longer than the QK MMA
        // of the next item's first tile that should hide it (profiles/r02_fwd_epilogue.md). Here a warp stages its
        // 32 rows x 64 head dims (4 KB, 128-byte swizzle) in a piece of the P ring and one lane issues a bulk store.
        // The P ring is idle: every PV of this item has retired (p_empty above), and the next writers of the piece a
        // warp borrows are warps that meet this one at a named barrier (row-max exchange) before they write P.
        //   column-split mode: piece (stage j/2, key half j%2, row group) with j = 2 * warpgroup + lane half
        //   alternate-tile mode: a stage of the parity this warpgroup itself writes next
        // Packed mode: rows past the end of a sequence belong to the next one, so partial row blocks go direct.
        constexpr bool kTmaEpi = (CPT != 16) && Cfg::NSW == 8;
        const int r0 = q0 + 64 * (int)rank + 32 * (int)rgrp;   // first row of this warp inside its sequence
        const bool use_tma = kTmaEpi && p.o_tma != 0 && p.kv_splits == 1 && (p.cu_q == nullptr || r0 + 32 <= seq_q);
        const uint32_t epi_j = ALT ? (((g + wgi) & 1u) * 2u + kh) : (wgi * 2u + kh);
        const uint32_t epi_base = sP + (epi_j >> 1) * 16384u + (epi_j & 1u) * 8192u + rgrp * 4096u;
        if (Cfg::NPASS == 2 && p.stash_p != nullptr && use_tma) {
          // replay path: the store warp may still be draining the last P tiles of this item out of the ring
#pragma unroll
          for (uint32_t st = 0; st < KS; ++st) ptx::mbar_wait(bar(bars.p_stored[(g + st) % KS]), (((g + st) / KS) & 1) ^ 1);
        }
#pragma unroll
        for (int s = 0; s < Cfg::NSLICE; ++s) {
          if (256 * s >= dvw) break;
          const int ns = Cfg::slice_n(dvw, s);
          constexpr int EG = ALT ? 2 : CQ;            // warpgroups sharing the columns of a lane
          constexpr int EC = CPT == 16 ? 16 : 32;     // columns per TMEM load
          const uint32_t eg = ALT ? wgi : ch;
          const int part = ns / 2 / EG;  // columns of this slice handled by each warpgroup
          if constexpr (kTmaEpi) {
            if (use_tma && part == 64) {
              const int dbox = dv0 + 256 * s + 128 * (int)kh + 64 * (int)eg;
              if (dbox >= p.head_dim) continue;   // padding columns only
              const uint32_t srow = epi_base + (row & 31u) * 128u;
              uint32_t w[32];   // 64 head dims of this row, packed
              {
                uint32_t orr[64];
                ptx::tmem_ld_x32(tmem + lane_base + 128 * s + 64 * (int)eg, orr);
                ptx::tmem_ld_x32(tmem + lane_base + 128 * s + 64 * (int)eg + 32, orr + 32);
                ptx::tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                  const float a = __uint_as_float(orr[2 * u]) * inv, c = __uint_as_float(orr[2 * u + 1]) * inv;
                  w[u] = BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
                }
              }
              if (t == 0) FFPA_STAMP(kidx, 8 + 3 * s);
              if (epi_pending) {
                // the piece is still being read by the bulk store of the previous slice (or of the previous item, when
                // no KV tile was processed since); the wait overlaps the TMEM loads and packing above. Measured: storing
                // the second slice with per-thread stores instead costs MORE (LSU back-pressure), see r02_fwd_epilogue.md
                if (ptx::lane_id() == 0) ptx::bulk_wait_group_read0();
                __syncwarp();
              }
              if (t == 0) FFPA_STAMP(kidx, 9 + 3 * s);
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const uint32_t addr = srow + (((uint32_t)v ^ (row & 7u)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * v]), "r"(w[4 * v + 1]), "r"(w[4 * v + 2]),
                             "r"(w[4 * v + 3])
                             : "memory");
              }
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (ptx::lane_id() == 0) {
                ptx::tma_store_4d(&map_o, epi_base, dbox, fi.qoff + r0, h, fi.bt);
                ptx::bulk_commit_group();
              }
              if (t == 0) FFPA_STAMP(kidx, 10 + 3 * s);
              epi_pending = true;   // the piece goes back to the P ring: waited for before the next P tile is written
              continue;
            }
          }
#pragma unroll 1
          for (int c0 = (int)eg * part; c0 < (int)(eg + 1) * part; c0 += EC) {
            uint32_t orr[EC];
            if constexpr (EC == 32) ptx::tmem_ld_x32(tmem + lane_base + 128 * s + c0, orr);
            else ptx::tmem_ld_x16(tmem + lane_base + 128 * s + c0, orr);
            ptx::tmem_wait_ld();
            const int d0 = dv0 + 256 * s + (ns / 2) * (int)kh + c0;
            if (row_ok) {
              if (p.kv_splits > 1) {
                // fp32 partial of this KV split, normalised by its own row sum
                float* po = p.part_o + ((((int64_t)fi.split * p.batch + b) * p.heads_q + h) * p.seqlen_q + gq) * (int64_t)p.head_dim;
#pragma unroll
                for (int v = 0; v < EC / 4; ++v)
                  if (d0 + 4 * v < p.head_dim)
                    *reinterpret_cast<float4*>(po + d0 + 4 * v) =
                        make_float4(__uint_as_float(orr[4 * v]) * inv, __uint_as_float(orr[4 * v + 1]) * inv,
                                    __uint_as_float(orr[4 * v + 2]) * inv, __uint_as_float(orr[4 * v + 3]) * inv);
              } else {
                uint32_t w[EC / 2];
#pragma unroll
                for (int u = 0; u < EC / 2; ++u) {
                  const float a = __uint_as_float(orr[2 * u]) * inv, c = __uint_as_float(orr[2 * u + 1]) * inv;
                  w[u] = BF16 ? ptx::pack_bf16x2(a, c) : ptx::pack_f16x2(a, c);
                }
#ifndef FFPA_DBG_SKIP_O_STORE
#pragma unroll
                for (int v = 0; v < EC / 16; ++v) {
                  const int d = d0 + 16 * v;
                  // every lane writes another row, so a store instruction costs one LSU wavefront per lane whatever its
                  // width: one 32-byte store (a whole sector) per 16 head dims where the row allows it, else two of 16
                  if (o_al32 && d + 16 <= p.head_dim) {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(orow + 2 * d), "r"(w[8 * v]),
                                 "r"(w[8 * v + 1]), "r"(w[8 * v + 2]), "r"(w[8 * v + 3]), "r"(w[8 * v + 4]), "r"(w[8 * v + 5]),
                                 "r"(w[8 * v + 6]), "r"(w[8 * v + 7])
                                 : "memory");
                  } else {
                    if (d < p.head_dim) *reinterpret_cast<uint4*>(orow + 2 * d) = make_uint4(w[8 * v], w[8 * v + 1], w[8 * v + 2], w[8 * v + 3]);
                    if (d + 8 < p.head_dim)
                      *reinterpret_cast<uint4*>(orow + 2 * d + 16) = make_uint4(w[8 * v + 4], w[8 * v + 5], w[8 * v + 6], w[8 * v + 7]);
                  }
                }
#endif
              }
            }
          }
        }
        }  // !EPIW
        if (!Cfg::EPIW && (p.lse != nullptr || p.kv_splits > 1) && wgi == 0 && kh == 0 && row_ok && pass == 0) {
          // natural-log LSE; rows without any visible key: O = 0, LSE = -inf
          const float lse = (l_tot > 0.f) ? (m * mul + log2f(l_tot)) * 0.6931471805599453f : NEG_INF;
          if (p.kv_splits > 1) p.part_lse[(((int64_t)fi.split * p.batch + b) * p.heads_q + h) * p.seqlen_q + gq] = lse;
          else if (p.cu_q != nullptr) p.lse[(int64_t)h * p.total_q + fi.qoff + gq] = lse;   // [Hq, total_q]
          else p.lse[((int64_t)b * p.heads_q + h) * p.lse_bh_stride + gq] = lse;
        }
        ptx::tc_fence_before();
        if (t == 0) FFPA_STAMP(kidx, 6);
      }
    }
  }

  if (!Cfg::EPIW && warp < (uint32_t)kSoftmaxWarps && ptx::lane_id() == 0) ptx::bulk_wait_group0();   // O tiles stored by the epilogue
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == kMmaWarp) ptx::tmem_dealloc<CG>(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// KV-split merge: O = sum_s w_s O_s / sum_s w_s with w_s = exp(LSE_s - max LSE), LSE = max + log(sum w_s)
// (the split-KV stage 2 of the reference's decode path, csrc/cuffpa/native/sm_80/split_kv.cuh:330-455).
// One warp per (b, h, q) row.
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void merge_splits_kernel(const float* __restrict__ part_o, const float* __restrict__ part_lse,
                                    void* __restrict__ o, float* __restrict__ lse, int64_t lse_bh_stride, int64_t os0,
                                    int64_t os1, int64_t os2, int B, int H, int Nq, int D, int S) {
  const int64_t rowid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t rows = (int64_t)B * H * Nq;
  if (rowid >= rows) return;
  const int lane = threadIdx.x & 31;
  const int q = (int)(rowid % Nq);
  const int h = (int)((rowid / Nq) % H);
  const int b = (int)(rowid / ((int64_t)Nq * H));
  float mx = -INFINITY;
  for (int s = 0; s < S; ++s) mx = fmaxf(mx, part_lse[(int64_t)s * rows + rowid]);
  float den = 0.f;
  for (int s = 0; s < S; ++s) {
    const float l = part_lse[(int64_t)s * rows + rowid];
    den += (l == -INFINITY) ? 0.f : __expf(l - mx);
  }
  const float inv = den > 0.f ? 1.f / den : 0.f;
  uint8_t* orow = reinterpret_cast<uint8_t*>(o) + 2 * (b * os0 + h * os1 + (int64_t)q * os2);
  for (int d = lane * 2; d < D; d += 64) {
    float a0 = 0.f, a1 = 0.f;
    for (int s = 0; s < S; ++s) {
      const float l = part_lse[(int64_t)s * rows + rowid];
      if (l == -INFINITY) continue;
      const float w = __expf(l - mx) * inv;
      const float2 v = *reinterpret_cast<const float2*>(part_o + ((int64_t)s * rows + rowid) * D + d);
      a0 = fmaf(w, v.x, a0);
      a1 = fmaf(w, v.y, a1);
    }
    *reinterpret_cast<uint32_t*>(orow + 2 * d) = BF16 ? ptx::pack_bf16x2(a0, a1) : ptx::pack_f16x2(a0, a1);
  }
  if (lse != nullptr && lane == 0) lse[((int64_t)b * H + h) * lse_bh_stride + q] = den > 0.f ? mx + __logf(den) : -INFINITY;
}

template <bool BF16>
int launch_merge_splits(const float* part_o, const float* part_lse, void* o, float* lse, int64_t lse_bh_stride,
                        const int64_t* ostride, int B, int H, int Nq, int D, int S, cudaStream_t stream) {
  const int64_t rows = (int64_t)B * H * Nq;
  const int wpb = 4;
  merge_splits_kernel<BF16><<<dim3((unsigned)((rows + wpb - 1) / wpb)), dim3(wpb * 32), 0, stream>>>(
      part_o, part_lse, o, lse, lse_bh_stride, ostride[0], ostride[1], ostride[2], B, H, Nq, D, S);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "split merge launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

// ------------------------------------------------------------------------------------------------
// host launcher pieces (instantiated per dtype in ffpa_fwd_bf16.cu / ffpa_fwd_f16.cu)
// ------------------------------------------------------------------------------------------------
template <int NQK, bool BF16, int MODE>
static int launch_variant(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& msp, const CUtensorMap& mo,
                          const FwdKernelParams& kp, int nclusters, cudaStream_t stream) {
  using Cfg = FwdCfg<NQK>;
  auto kern = ffpa_fwd_kernel<NQK, BF16, MODE>;
  // the opt-in shared-memory size is a per-device function attribute
  static bool attr_set[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  dev_id = (dev_id >= 0 && dev_id < 64) ? dev_id : 0;
  if (!attr_set[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_DYN);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_DYN, cudaGetErrorString(e));
    attr_set[dev_id] = true;
  }
  kern<<<dim3(2 * nclusters), dim3(Cfg::THREADS), Cfg::SMEM_DYN, stream>>>(mq, mk, mv, msp, mo, kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "forward launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return FFPA_OK;
}

template <bool BF16, int MODE>
static int dispatch_nqk(int nqk, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& msp, const CUtensorMap& mo,
                        const FwdKernelParams& kp, int nclusters, cudaStream_t stream) {
  switch (nqk) {
    case 1: return launch_variant<1, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 2: return launch_variant<2, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 3: return launch_variant<3, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 4: return launch_variant<4, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 5: return launch_variant<5, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 6: return launch_variant<6, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 7: return launch_variant<7, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 8: return launch_variant<8, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 9: return launch_variant<9, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 10: return launch_variant<10, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 11: return launch_variant<11, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 12: return launch_variant<12, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 13: return launch_variant<13, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 14: return launch_variant<14, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 15: return launch_variant<15, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    case 16: return launch_variant<16, BF16, MODE>(mq, mk, mv, msp, mo, kp, nclusters, stream);
    default: return set_error(FFPA_ERR_UNSUPPORTED, "head_dim > 1024 not supported");
  }
}

template <bool BF16>
int dispatch_fwd_dtype(int nqk, int mode, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& msp, const CUtensorMap& mo,
                       const FwdKernelParams& kp, int nclusters, cudaStream_t stream) {
  if (mode == kModeFast) return dispatch_nqk<BF16, kModeFast>(nqk, mq, mk, mv, msp, mo, kp, nclusters, stream);
  if (mode == kModeGeneral) return dispatch_nqk<BF16, kModeGeneral>(nqk, mq, mk, mv, msp, mo, kp, nclusters, stream);
  return dispatch_nqk<BF16, kModeDropout>(nqk, mq, mk, mv, msp, mo, kp, nclusters, stream);
}

}  // namespace ffpa
