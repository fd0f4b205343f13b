// Host launcher of the backward: preprocess (delta, lse2) + dQ + dK + dV kernels on one stream.
// Implements the symbol the reference leaves as a thrower
// (/root/reference/csrc/cuffpa/ffpa_api.cc:242-263).
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace bwd {
template <bool BF16>
int dispatch_bwd_dtype(int nqk, int kind, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1,
                       const CUtensorMap& b2, const CUtensorMap& b3, const BwdKernelParams& kp, int nclusters,
                       cudaStream_t stream);
template <bool BF16>
int launch_preprocess(const ffpa_bwd_params& a, float* lse2, float* delta, int nq_pad, cudaStream_t stream);
extern template int dispatch_bwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                             const CUtensorMap&, const CUtensorMap&, const BwdKernelParams&, int, cudaStream_t);
extern template int dispatch_bwd_dtype<false>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                              const CUtensorMap&, const CUtensorMap&, const BwdKernelParams&, int, cudaStream_t);
extern template int launch_preprocess<true>(const ffpa_bwd_params&, float*, float*, int, cudaStream_t);
extern template int launch_preprocess<false>(const ffpa_bwd_params&, float*, float*, int, cudaStream_t);
}  // namespace bwd

static bool make_map4(CUtensorMap* m, const void* base, const int64_t* stride, int B, int H, int N, int D,
                      uint32_t box_d, uint32_t box_n) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)N, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)stride[2] * 2, (uint64_t)stride[1] * 2, (uint64_t)stride[0] * 2};
  const uint64_t packed[3] = {(uint64_t)D * 2, (uint64_t)D * 2 * N, (uint64_t)D * 2 * N * H};
  for (int i = 0; i < 3; ++i)
    if (dims[i + 1] == 1) str[i] = packed[i];
  uint32_t box[4] = {box_d, box_n, 1, 1};
  return tmap::encode_sw128(m, const_cast<void*>(base), 2, 4, dims, str, box);
}

uint64_t bwd_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim) {
  (void)heads_kv; (void)seqlen_kv; (void)head_dim;
  const uint64_t nq_pad = ((uint64_t)seqlen_q + 127) / 128 * 128;
  return 2ull * batch * heads_q * nq_pad * sizeof(float);
}

int launch_bwd_sm100(const ffpa_bwd_params& a, cudaStream_t stream) {
  const int D = a.head_dim, nqk = (D + 63) / 64;
  const int nq_pad = (a.seqlen_q + 127) / 128 * 128;
  const uint64_t need = bwd_workspace_bytes(a.batch, a.heads_q, a.heads_kv, a.seqlen_q, a.seqlen_kv, D);
  if (!a.workspace || a.workspace_bytes < need)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "backward workspace too small: need %llu bytes", (unsigned long long)need);
  float* lse2 = static_cast<float*>(a.workspace);
  float* delta = lse2 + (size_t)a.batch * a.heads_q * nq_pad;
  const bool bf16 = a.dtype == FFPA_DTYPE_BF16;

  int rc = bf16 ? bwd::launch_preprocess<true>(a, lse2, delta, nq_pad, stream)
                : bwd::launch_preprocess<false>(a, lse2, delta, nq_pad, stream);
  if (rc) return rc;

  CUtensorMap q_km, q_mn, k_km, k_mn, v_km, do_km, do_mn;
  const int B = a.batch, Hq = a.heads_q, Hkv = a.heads_kv, Nq = a.seqlen_q, Nkv = a.seqlen_kv;
  if (!make_map4(&q_km, a.q, a.q_stride, B, Hq, Nq, D, 64, 64) || !make_map4(&q_mn, a.q, a.q_stride, B, Hq, Nq, D, 64, 128) ||
      !make_map4(&k_km, a.k, a.k_stride, B, Hkv, Nkv, D, 64, 64) || !make_map4(&k_mn, a.k, a.k_stride, B, Hkv, Nkv, D, 64, 128) ||
      !make_map4(&v_km, a.v, a.v_stride, B, Hkv, Nkv, D, 64, 64) ||
      !make_map4(&do_km, a.d_o, a.do_stride, B, Hq, Nq, D, 64, 64) || !make_map4(&do_mn, a.d_o, a.do_stride, B, Hq, Nq, D, 64, 128))
    return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed in backward");

  bwd::BwdKernelParams kp{};
  kp.lse2 = lse2; kp.delta = delta; kp.nq_pad = nq_pad;
  kp.batch = B; kp.heads_q = Hq; kp.heads_kv = Hkv; kp.seqlen_q = Nq; kp.seqlen_kv = Nkv; kp.head_dim = D;
  kp.causal = a.causal; kp.scale = a.softmax_scale; kp.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  kp.bias = a.bias_kind != FFPA_BIAS_NONE ? a.bias : nullptr;
  kp.bias_kind = a.bias_kind;
  for (int i = 0; i < 4; ++i) kp.bias_stride[i] = a.bias_stride[i];
  kp.dropout_p = a.dropout_p; kp.philox_seed = a.philox_seed; kp.philox_offset = a.philox_offset;
  kp.dbias = a.d_bias;
  const int max_clusters = sm_count() / 2;
  auto run = [&](int kind, void* out, const int64_t* ostride, int rows, int heads, const CUtensorMap& a1,
                 const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2, const CUtensorMap& b3) {
    kp.out = out;
    for (int i = 0; i < 3; ++i) kp.out_stride[i] = ostride[i];
    kp.n_rtiles = (rows + 127) / 128;
    kp.n_items = kp.n_rtiles * B * heads;
    const int ncl = kp.n_items < max_clusters ? kp.n_items : max_clusters;
    return bf16 ? bwd::dispatch_bwd_dtype<true>(nqk, kind, a1, a2, b1, b2, b3, kp, ncl, stream)
                : bwd::dispatch_bwd_dtype<false>(nqk, kind, a1, a2, b1, b2, b3, kp, ncl, stream);
  };
  if ((rc = run(0, a.dq, a.dq_stride, Nq, Hq, q_km, do_km, k_km, v_km, k_mn))) return rc;
  if ((rc = run(1, a.dk, a.dk_stride, Nkv, Hkv, k_km, v_km, q_km, do_km, q_mn))) return rc;
  if ((rc = run(2, a.dv, a.dv_stride, Nkv, Hkv, k_km, k_km, q_km, q_km, do_mn))) return rc;
  return FFPA_OK;
}

}  // namespace ffpa
