// Host launcher of the forward: tensor-map construction (passed as __grid_constant__ kernel
// parameters -- no per-launch cudaMalloc/cudaMemcpy as in
// /root/reference/csrc/cuffpa/native/launch.cuh:503-509), persistent grid sizing, mode selection.
#include <vector>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {

template <bool BF16>
int dispatch_fwd_dtype(int nqk, int mode, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                       const FwdKernelParams& kp, int nclusters, cudaStream_t stream);
extern template int dispatch_fwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                             const FwdKernelParams&, int, cudaStream_t);
extern template int dispatch_fwd_dtype<false>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                              const FwdKernelParams&, int, cudaStream_t);

template <bool BF16>
int launch_merge_splits(const float* part_o, const float* part_lse, void* o, float* lse, const int64_t* ostride, int B,
                        int H, int Nq, int D, int S, cudaStream_t stream);
extern template int launch_merge_splits<true>(const float*, const float*, void*, float*, const int64_t*, int, int, int, int, int, cudaStream_t);
extern template int launch_merge_splits<false>(const float*, const float*, void*, float*, const int64_t*, int, int, int, int, int, cudaStream_t);

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

uint64_t fwd_split_workspace_bytes(int batch, int heads_q, int seqlen_q, int seqlen_kv, int head_dim) {
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
  kp.o = a.o;
  kp.lse = a.lse;
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
    const uint64_t need = fwd_split_workspace_bytes(a.batch, a.heads_q, a.seqlen_q, a.seqlen_kv, D);
    if (sp > 1 && a.workspace != nullptr && a.workspace_bytes >= need && (reinterpret_cast<uintptr_t>(a.workspace) & 15u) == 0) {
      const uint64_t rows = (uint64_t)a.batch * a.heads_q * a.seqlen_q;
      kp.kv_splits = sp;
      kp.part_o = static_cast<float*>(a.workspace);
      kp.part_lse = kp.part_o + (uint64_t)sp * rows * D;
    }
  }
  kp.n_items = kp.n_mtiles * npass * kp.kv_splits * a.batch * a.heads_q;

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
      int t = ((mt * 128 + 127 + off) >> 7) + 1;
      t = t < tc ? t : tc;
      cost[it] = (t < 1 ? 1 : t) * 16 + 24;  // tiles + fixed per-item overhead (prologue/epilogue ~1.5 tiles)
    }
    kp.sched = get_schedule(cost.data(), kp.n_items, nclusters, &kp.sched_stride, stream);
  }
  int mode = 0;  // fast
  if (a.dropout_p > 0.f) mode = 2;
  else if (a.bias_kind != FFPA_BIAS_NONE || !(a.softmax_scale > 0.f)) mode = 1;
  int rc = (a.dtype == FFPA_DTYPE_BF16) ? dispatch_fwd_dtype<true>(nqk, mode, mq, mk, mv, kp, nclusters, stream)
                                        : dispatch_fwd_dtype<false>(nqk, mode, mq, mk, mv, kp, nclusters, stream);
  if (rc || kp.kv_splits == 1) return rc;
  return (a.dtype == FFPA_DTYPE_BF16)
             ? launch_merge_splits<true>(kp.part_o, kp.part_lse, a.o, a.lse, a.o_stride, a.batch, a.heads_q, a.seqlen_q, D, kp.kv_splits, stream)
             : launch_merge_splits<false>(kp.part_o, kp.part_lse, a.o, a.lse, a.o_stride, a.batch, a.heads_q, a.seqlen_q, D, kp.kv_splits, stream);
}

}  // namespace ffpa
