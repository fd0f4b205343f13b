// Host launcher of the forward: tensor-map construction (passed as __grid_constant__ kernel
// parameters -- no per-launch cudaMalloc/cudaMemcpy as in
// /root/reference/csrc/cuffpa/native/launch.cuh:503-509), persistent grid sizing, mode selection.
#include <cstdlib>
#include <vector>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {

template <bool BF16>
int dispatch_fwd_dtype(int nqk, int mode, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                       const CUtensorMap& msp, const CUtensorMap& mo, const FwdKernelParams& kp, int nclusters, cudaStream_t stream);
extern template int dispatch_fwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                             const CUtensorMap&, const CUtensorMap&, const FwdKernelParams&, int, cudaStream_t);
extern template int dispatch_fwd_dtype<false>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                              const CUtensorMap&, const CUtensorMap&, const FwdKernelParams&, int, cudaStream_t);
namespace replay {
template <bool BF16>
int launch_fwd_replay(const CUtensorMap& map_p, const CUtensorMap& map_v, const FwdReplayParams& kp, int nclusters,
                      cudaStream_t stream);
extern template int launch_fwd_replay<true>(const CUtensorMap&, const CUtensorMap&, const FwdReplayParams&, int, cudaStream_t);
extern template int launch_fwd_replay<false>(const CUtensorMap&, const CUtensorMap&, const FwdReplayParams&, int, cudaStream_t);
}  // namespace replay

template <bool BF16>
int launch_merge_splits(const float* part_o, const float* part_lse, void* o, float* lse, int64_t lse_bh_stride,
                        const int64_t* ostride, int B, int H, int Nq, int D, int S, cudaStream_t stream);
extern template int launch_merge_splits<true>(const float*, const float*, void*, float*, int64_t, const int64_t*, int, int, int, int, int, cudaStream_t);
extern template int launch_merge_splits<false>(const float*, const float*, void*, float*, int64_t, const int64_t*, int, int, int, int, int, cudaStream_t);

// KV splits for decode-like shapes: few query tiles (items) and a long KV sequence leave most of the 74
// clusters idle and one cluster streaming all of K/V of a head; splitting the KV range restores the
// parallelism (HBM-bound regime; reference: split-KV decode, csrc/cuffpa/native/launch.cuh:17-67).
int fwd_kv_splits(int batch, int heads_q, int seqlen_q, int seqlen_kv, int head_dim) {
  const int nqk = (head_dim + 63) / 64;
  const int dvp = ((nqk * 64 + 127) / 128) * 128;
  if (dvp > 768) return 1;  // two-pass head dims keep the direct path
  const long long items = (long long)((seqlen_q + 127) / 128) * batch * heads_q;
  const int ncl = sm_count() / 2;
  const int tk = (seqlen_kv + 127) / 128;
  if (items * 2 > ncl || tk < 8) return 1;
  long long s = ncl / items;
  if (s > tk / 4) s = tk / 4;
  if (s > 32) s = 32;
  return s < 2 ? 1 : (int)s;
}

// Replay path for head dims > 768 (ffpa_fwd_replay_sm100.cuh): pass 0 stores its 16-bit P tiles, the per-tile O
// rescale factors and 1 / rowsum so the second O slab is a GEMM instead of a second Q K^T + softmax pass.
// The scratch is O(Nq * Nkv) per head, so it is planned from the bytes on offer (`avail`: the caller's cap when
// sizing, the granted workspace at launch): the whole problem if it fits, else (batch element, KV-head range) chunks
// that run one after the other through the same scratch (each must still fill the machine), else the two-pass
// kernel (no scratch). FFPA_FWD_REPLAY=0 disables the path.
struct ReplayPlan {
  uint64_t p_bytes = 0, f_bytes = 0, inv_bytes = 0;
  int n_mt_even = 0, nk_pad = 0;
  int chunk_hkv = 0;
  bool chunked = false;
  uint64_t total() const { return p_bytes + f_bytes + inv_bytes; }
};
static ReplayPlan replay_sizes(uint64_t bh, int seqlen_q, int seqlen_kv) {
  ReplayPlan pl;
  pl.n_mt_even = (((seqlen_q + 127) / 128) + 1) & ~1;
  pl.nk_pad = (seqlen_kv + 255) / 256 * 256;
  pl.p_bytes = bh * pl.n_mt_even * 128ull * pl.nk_pad * 2;
  pl.f_bytes = (bh * pl.n_mt_even * (uint64_t)(pl.nk_pad / 128) * 128 * 4 + 255) / 256 * 256;
  pl.inv_bytes = (bh * pl.n_mt_even * 128 * 4 + 255) / 256 * 256;
  return pl;
}
static ReplayPlan replay_plan(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim, uint64_t avail) {
  const int nqk = (head_dim + 63) / 64, dvp = ((nqk * 64 + 127) / 128) * 128;
  if (dvp <= 768 || heads_kv <= 0) return ReplayPlan{};
  if (env_off("FFPA_FWD_REPLAY")) return ReplayPlan{};
  ReplayPlan whole = replay_sizes((uint64_t)batch * heads_q, seqlen_q, seqlen_kv);
  if (whole.total() + 256 <= avail) { whole.chunk_hkv = heads_kv; return whole; }
  const int group = heads_q / heads_kv, n_mt = (seqlen_q + 127) / 128;
  for (int hc = heads_kv; hc >= 1; --hc) {
    if (hc == heads_kv && batch == 1) continue;   // that is the whole problem
    ReplayPlan pl = replay_sizes((uint64_t)hc * group, seqlen_q, seqlen_kv);
    if (pl.total() + 256 > avail) continue;
    if ((int64_t)hc * group * n_mt < sm_count() / 2) break;   // a chunk must still fill the machine
    pl.chunk_hkv = hc; pl.chunked = true;
    return pl;
  }
  return ReplayPlan{};
}

uint64_t fwd_split_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim, uint64_t cap_bytes) {
  const ReplayPlan rp = replay_plan(batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim, cap_bytes);
  if (rp.total() > 0) return rp.total() + 256;
  const int s = fwd_kv_splits(batch, heads_q, seqlen_q, seqlen_kv, head_dim);
  if (s <= 1) return 0;
  const uint64_t rows = (uint64_t)batch * heads_q * seqlen_q;
  return (uint64_t)s * rows * ((uint64_t)head_dim + 1) * sizeof(float) + 256;
}

static bool make_map(CUtensorMap* m, const void* base, const int64_t* stride, int B, int H, int N,
                     int D, uint32_t box_d, uint32_t box_n) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)N, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)stride[2] * 2, (uint64_t)stride[1] * 2, (uint64_t)stride[0] * 2};
  // size-1 dims may carry arbitrary strides; give TMA the packed one (multiple of 16 bytes)
  const uint64_t packed[3] = {(uint64_t)D * 2, (uint64_t)D * 2 * N, (uint64_t)D * 2 * N * H};
  for (int i = 0; i < 3; ++i)
    if (dims[i + 1] == 1) str[i] = packed[i];
  uint32_t box[4] = {box_d, box_n, 1, 1};
  return tmap::encode_sw128(m, const_cast<void*>(base), 2, 4, dims, str, box);
}

int launch_fwd_sm100(const ffpa_fwd_params& a, cudaStream_t stream) {
  const int D = a.head_dim;
  const int nqk = (D + 63) / 64;
  CUtensorMap mq, mk, mv;
  // packed variable-length mode: one [total tokens, H, D] tensor per operand (map batch extent 1)
  const bool varlen = a.cu_seqlens_q != nullptr;
  const int mb = varlen ? 1 : a.batch, mnq = varlen ? a.total_q : a.seqlen_q, mnk = varlen ? a.total_k : a.seqlen_kv;
  if (!make_map(&mq, a.q, a.q_stride, mb, a.heads_q, mnq, D, 64, 64) ||
      !make_map(&mk, a.k, a.k_stride, mb, a.heads_kv, mnk, D, 64, 64) ||
      !make_map(&mv, a.v, a.v_stride, mb, a.heads_kv, mnk, D, 64, 128))
    return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed (strides must be multiples of 8 elements, base 16-byte aligned)");

  FwdKernelParams kp{};
  // O store map: [64 head dims x 32 rows] boxes, one per softmax warp and slice (epilogue of the forward kernel).
  // Optional: an O the tensor-map encoder refuses is written with per-thread stores instead.
  CUtensorMap mo = mq;
  kp.o_tma = make_map(&mo, a.o, a.o_stride, mb, a.heads_q, mnq, D, 64, 32) ? 1 : 0;
  kp.o = a.o;
  kp.lse = a.lse;
  kp.lse_bh_stride = a.lse_bh_stride > 0 ? a.lse_bh_stride : a.seqlen_q;
  kp.bias = a.bias;
  for (int i = 0; i < 3; ++i) kp.o_stride[i] = a.o_stride[i];
  for (int i = 0; i < 4; ++i) kp.bias_stride[i] = a.bias_stride[i];
  kp.batch = a.batch; kp.heads_q = a.heads_q; kp.heads_kv = a.heads_kv;
  kp.seqlen_q = a.seqlen_q; kp.seqlen_kv = a.seqlen_kv; kp.head_dim = D;
  kp.causal = a.causal; kp.bias_kind = a.bias_kind;
  kp.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  kp.dropout_p = a.dropout_p;
  kp.philox_seed = a.philox_seed; kp.philox_offset = a.philox_offset;
  kp.n_mtiles = (a.seqlen_q + 127) / 128;
  const int dvp = ((nqk * 64 + 127) / 128) * 128;
  const int npass = dvp > 768 ? 2 : 1;  // must match FwdCfg<NQK>::NPASS
  kp.kv_splits = 1;
  kp.part_o = nullptr;
  kp.part_lse = nullptr;
  kp.cu_q = a.cu_seqlens_q;
  kp.cu_k = a.cu_seqlens_k;
  kp.total_q = a.total_q;
  kp.total_k = a.total_k;
  if (!varlen) {
    const int sp = fwd_kv_splits(a.batch, a.heads_q, a.seqlen_q, a.seqlen_kv, D);
    const uint64_t need = fwd_split_workspace_bytes(a.batch, a.heads_q, a.heads_kv, a.seqlen_q, a.seqlen_kv, D, 0);
    if (sp > 1 && a.workspace != nullptr && a.workspace_bytes >= need && (reinterpret_cast<uintptr_t>(a.workspace) & 15u) == 0) {
      const uint64_t rows = (uint64_t)a.batch * a.heads_q * a.seqlen_q;
      kp.kv_splits = sp;
      kp.part_o = static_cast<float*>(a.workspace);
      kp.part_lse = kp.part_o + (uint64_t)sp * rows * D;
    }
  }
  // replay path (head dims > 768): one softmax pass that stores P / rescale factors / 1/rowsum, then a GEMM
  const bool ws_ok = a.workspace != nullptr && (reinterpret_cast<uintptr_t>(a.workspace) & 255u) == 0;
  const ReplayPlan rp = (varlen || !ws_ok) ? ReplayPlan{}
                                           : replay_plan(a.batch, a.heads_q, a.heads_kv, a.seqlen_q, a.seqlen_kv, D, a.workspace_bytes);
  if (rp.chunked && a.bias_kind == FFPA_BIAS_NONE && !(a.dropout_p > 0.f)) {
    // the stash of the whole problem does not fit the scratch: (batch element, KV-head range) chunks, one after the
    // other through the same scratch (stream order serialises them); each chunk is an ordinary dense sub-problem
    const int group = a.heads_q / a.heads_kv;
    auto off = [](const void* p, int64_t elems) { return static_cast<const void*>(static_cast<const uint8_t*>(p) + 2 * elems); };
    const int64_t lse_bh = a.lse_bh_stride > 0 ? a.lse_bh_stride : a.seqlen_q;
    for (int b = 0; b < a.batch; ++b)
      for (int hk0 = 0; hk0 < a.heads_kv; hk0 += rp.chunk_hkv) {
        const int hc = (a.heads_kv - hk0) < rp.chunk_hkv ? (a.heads_kv - hk0) : rp.chunk_hkv;
        const int hq0 = hk0 * group;
        ffpa_fwd_params s = a;
        s.batch = 1; s.heads_kv = hc; s.heads_q = hc * group;
        s.q = off(a.q, b * a.q_stride[0] + hq0 * a.q_stride[1]);
        s.o = const_cast<void*>(off(a.o, b * a.o_stride[0] + hq0 * a.o_stride[1]));
        s.k = off(a.k, b * a.k_stride[0] + hk0 * a.k_stride[1]);
        s.v = off(a.v, b * a.v_stride[0] + hk0 * a.v_stride[1]);
        if (a.lse) s.lse = a.lse + ((int64_t)b * a.heads_q + hq0) * lse_bh;
        s.lse_bh_stride = lse_bh;
        if (int rc = launch_fwd_sm100(s, stream)) return rc;
      }
    return FFPA_OK;
  }
  const bool use_replay = rp.total() > 0 && !rp.chunked;
  kp.stash_p = nullptr; kp.stash_f = nullptr; kp.stash_inv = nullptr;
  kp.nk_pad = rp.nk_pad; kp.n_mt_even = rp.n_mt_even;
  kp.n_pass = npass;
  CUtensorMap msp = mq, mpl = mq;   // P store map ([64 x 64] boxes) / load map ([128 x 64] boxes) over the tile-major stash
  if (use_replay) {
    uint8_t* ws = static_cast<uint8_t*>(a.workspace);
    kp.stash_p = ws;
    kp.stash_f = reinterpret_cast<float*>(ws + rp.p_bytes);
    kp.stash_inv = reinterpret_cast<float*>(ws + rp.p_bytes + rp.f_bytes);
    kp.n_pass = 1;
    const int sblocks = rp.n_mt_even * (rp.nk_pad / 64);
    const int64_t sstr[4] = {(int64_t)sblocks * 8192, 8192, 64, 1};
    if (!make_map(&msp, kp.stash_p, sstr, a.batch * a.heads_q, sblocks, 128, 64, 64, 64) ||
        !make_map(&mpl, kp.stash_p, sstr, a.batch * a.heads_q, sblocks, 128, 64, 64, 128))
      return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the forward replay stash");
  }
  kp.n_items = kp.n_mtiles * kp.n_pass * kp.kv_splits * a.batch * a.heads_q;

  int nclusters = sm_count() / 2;
  if (nclusters > kp.n_items) nclusters = kp.n_items;
  kp.sched = nullptr;
  kp.sched_stride = 0;
  if (a.causal && kp.n_items > nclusters && kp.kv_splits == 1 && !varlen) {
    // causal items differ in length: balance them over the persistent clusters (greedy LPT over a
    // head-major, longest-first order; table cached on the device per shape)
    std::vector<int> cost((size_t)kp.n_items);
    const int off = a.seqlen_kv - a.seqlen_q, tc = (a.seqlen_kv + 127) / 128;
    for (int it = 0; it < kp.n_items; ++it) {
      const int mt = it % kp.n_mtiles;
      int t = (((use_replay ? (mt * 128) | 128 : mt * 128) + 127 + off) >> 7) + 1;
      t = t < tc ? t : tc;
      cost[it] = (t < 1 ? 1 : t) * 16 + 24;  // tiles + fixed per-item overhead (prologue/epilogue ~1.5 tiles)
    }
    kp.sched = get_schedule(cost.data(), kp.n_items, nclusters, &kp.sched_stride, stream);
  }
  int mode = 0;  // fast
  if (a.dropout_p > 0.f) mode = 2;
  else if (a.bias_kind != FFPA_BIAS_NONE || !(a.softmax_scale > 0.f)) mode = 1;
  int rc = (a.dtype == FFPA_DTYPE_BF16) ? dispatch_fwd_dtype<true>(nqk, mode, mq, mk, mv, msp, mo, kp, nclusters, stream)
                                        : dispatch_fwd_dtype<false>(nqk, mode, mq, mk, mv, msp, mo, kp, nclusters, stream);
  if (rc) return rc;
  if (use_replay) {
    FwdReplayParams gp{};
    gp.o = a.o;
    for (int i = 0; i < 3; ++i) gp.o_stride[i] = a.o_stride[i];
    gp.stash_f = kp.stash_f; gp.stash_inv = kp.stash_inv;
    gp.batch = a.batch; gp.heads_q = a.heads_q; gp.heads_kv = a.heads_kv;
    gp.seqlen_q = a.seqlen_q; gp.seqlen_kv = a.seqlen_kv; gp.head_dim = D; gp.causal = a.causal;
    gp.nk_pad = rp.nk_pad; gp.n_mt_even = rp.n_mt_even;
    gp.n_qblocks = rp.n_mt_even / 2;
    gp.n_items = gp.n_qblocks * a.batch * a.heads_q;
    int ncl = sm_count() / 2;
    if (ncl > gp.n_items) ncl = gp.n_items;
    gp.sched = nullptr; gp.sched_stride = 0;
    if (a.causal && gp.n_items > ncl) {
      std::vector<int> cost((size_t)gp.n_items);
      const int off = a.seqlen_kv - a.seqlen_q, tc = (a.seqlen_kv + 127) / 128;
      for (int it = 0; it < gp.n_items; ++it) {
        int t = ((((it % gp.n_qblocks) * 256 + 128) + 127 + off) >> 7) + 1;
        t = t < tc ? t : tc;
        cost[it] = (t < 1 ? 1 : t) * 16 + 24;
      }
      gp.sched = get_schedule(cost.data(), gp.n_items, ncl, &gp.sched_stride, stream);
    }
    return (a.dtype == FFPA_DTYPE_BF16) ? replay::launch_fwd_replay<true>(mpl, mv, gp, ncl, stream)
                                        : replay::launch_fwd_replay<false>(mpl, mv, gp, ncl, stream);
  }
  if (kp.kv_splits == 1) return rc;
  return (a.dtype == FFPA_DTYPE_BF16)
             ? launch_merge_splits<true>(kp.part_o, kp.part_lse, a.o, a.lse, kp.lse_bh_stride, a.o_stride, a.batch, a.heads_q, a.seqlen_q, D, kp.kv_splits, stream)
             : launch_merge_splits<false>(kp.part_o, kp.part_lse, a.o, a.lse, kp.lse_bh_stride, a.o_stride, a.batch, a.heads_q, a.seqlen_q, D, kp.kv_splits, stream);
}

}  // namespace ffpa
