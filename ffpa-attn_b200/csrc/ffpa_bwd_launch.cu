// Host launcher of the backward: preprocess (delta, lse2) + dQ + dK + dV kernels on one stream.
// Implements the symbol the reference leaves as a thrower
// (/root/reference/csrc/cuffpa/ffpa_api.cc:242-263).
#include <cstdlib>
#include <vector>
#include "ffpa_internal.h"
#include "sm100_ptx.cuh"

namespace ffpa {
namespace bwd {
template <bool BF16>
int dispatch_bwd_dtype(int nqk, int kind, const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1,
                       const CUtensorMap& b2, const CUtensorMap& b3, const CUtensorMap& st, const CUtensorMap& sp,
                       const BwdKernelParams& kp, int nclusters, cudaStream_t stream);
template <bool BF16>
int launch_preprocess(const ffpa_bwd_params& a, float* lse2, float* delta, int nq_pad, cudaStream_t stream);
template <bool BF16>
int launch_bwd_gemm(const CUtensorMap& map_t, const CUtensorMap& map_b, const BwdGemmParams& kp, int nclusters,
                    cudaStream_t stream);
extern template int launch_bwd_gemm<true>(const CUtensorMap&, const CUtensorMap&, const BwdGemmParams&, int, cudaStream_t);
extern template int launch_bwd_gemm<false>(const CUtensorMap&, const CUtensorMap&, const BwdGemmParams&, int, cudaStream_t);
extern template int dispatch_bwd_dtype<true>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                             const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const BwdKernelParams&, int, cudaStream_t);
extern template int dispatch_bwd_dtype<false>(int, int, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&,
                                              const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const BwdKernelParams&, int, cudaStream_t);
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

static inline uint64_t align256(uint64_t x) { return (x + 255) / 256 * 256; }

// few KV-stationary items (B*Hkv*ceil(Nkv/128) < 8 clusters' worth): dK/dV may need chunked items
// with fp32 accumulation buffers; reserve them.
static bool may_split_kv(int batch, int heads_kv, int seqlen_kv) {
  const int64_t kv_items = (int64_t)batch * heads_kv * ((seqlen_kv + 127) / 128);
  return kv_items < 8ll * (sm_count() / 2);
}

uint64_t bwd_workspace_bytes_min(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim) {
  const uint64_t nq_pad = ((uint64_t)seqlen_q + 127) / 128 * 128;
  uint64_t total = align256(2ull * batch * heads_q * nq_pad * sizeof(float));
  if (may_split_kv(batch, heads_kv, seqlen_kv))
    total += 2 * align256((uint64_t)batch * heads_kv * seqlen_kv * head_dim * sizeof(float));
  return total;
}

// Stash path (ffpa_bwd_gemm_sm100.cuh): two 16-bit [B, Hq, nq_pad, nk_pad] score buffers (P_drop and dS).
// It pays when a GEMM pass over the head dim costs more than moving the N x N tile through HBM: head dims
// >= 384 (above 512 the dQ kernel stores on the first of its two slab passes and the GEMM-only kernel runs one
// pass per 512-wide output slab). The plan is a pure function of the workspace bytes on offer (`avail`): at
// sizing time that is the caller's cap (the torch binding derives it from free device memory), at launch the
// bytes actually granted -- so any workspace >= the minimum is valid and more memory means fewer GEMM passes.
// A problem whose buffers do not fit is cut into (batch element, KV-head range) chunks that run one after the
// other through the same buffers; if not even a machine-filling chunk fits, the three recompute kernels run
// (O(N) memory). FFPA_BWD_STASH=0 disables the path; how much memory to offer is the binder's policy
// (csrc/ffpa_torch_binding.cpp: at most half of the free memory, FFPA_BWD_STASH_MAX_GB as an upper bound).
struct StashPlan {
  uint64_t one = 0;     // bytes of ONE score buffer (of a chunk when chunked); 0 = path does not apply
  int chunk_hkv = 0;    // KV heads per chunk
  bool chunked = false;
  uint64_t need = 0;    // workspace bytes this plan needs (minimum scratch included)
};
static StashPlan stash_plan(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim, uint64_t avail) {
  StashPlan pl;
  if (head_dim < 384 || head_dim > 1024 || heads_kv <= 0) return pl;
  if (env_off("FFPA_BWD_STASH")) return pl;
  const uint64_t nq_pad = ((uint64_t)seqlen_q + 127) / 128 * 128, nk_pad = ((uint64_t)seqlen_kv + 255) / 256 * 256;
  const int group = heads_q / heads_kv;
  const uint64_t per_hkv = (uint64_t)group * nq_pad * nk_pad * 2;   // multiple of 256 bytes
  const uint64_t total = (uint64_t)batch * heads_kv * per_hkv;
  const uint64_t whole = bwd_workspace_bytes_min(batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim);
  if (whole + 2 * total <= avail) {
    pl.one = total; pl.chunk_hkv = heads_kv; pl.need = whole + 2 * total;
    return pl;
  }
  // largest KV-head count whose sub-problem (batch 1) fits; a chunk must still fill the machine
  for (int hc = heads_kv; hc >= 1; --hc) {
    const uint64_t sub = bwd_workspace_bytes_min(1, hc * group, hc, seqlen_q, seqlen_kv, head_dim) + 2 * hc * per_hkv;
    if (sub > avail) continue;
    if ((int64_t)hc * group * (int64_t)(nq_pad / 128) < sm_count() / 2) break;
    if (hc == heads_kv && batch == 1) break;   // would be the unchunked plan, which did not fit
    pl.one = (uint64_t)hc * per_hkv; pl.chunk_hkv = hc; pl.chunked = true;
    pl.need = sub > whole ? sub : whole;
    return pl;
  }
  return pl;
}

uint64_t bwd_workspace_bytes(int batch, int heads_q, int heads_kv, int seqlen_q, int seqlen_kv, int head_dim, uint64_t cap_bytes) {
  const uint64_t whole = bwd_workspace_bytes_min(batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim);
  if (cap_bytes <= whole) return whole;
  const StashPlan pl = stash_plan(batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim, cap_bytes);
  return pl.one == 0 ? whole : pl.need;
}

template <bool BF16>
__global__ void f32_to_16_kernel(const float* __restrict__ src, void* __restrict__ dst, int64_t s0, int64_t s1,
                                 int64_t s2, int H, int N, int D, int64_t total_vec) {
  // src [B, H, N, D] contiguous fp32 -> dst with element strides (s0, s1, s2, 1); 8 elements per thread
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8;
    const int d = (int)(e % D);
    const int64_t r = e / D;
    const int n = (int)(r % N);
    const int h = (int)((r / N) % H);
    const int64_t b = r / ((int64_t)N * H);
    const float4 a = *reinterpret_cast<const float4*>(src + e);
    const float4 c = *reinterpret_cast<const float4*>(src + e + 4);
    uint4 o;
    if (BF16) { o.x = ptx::pack_bf16x2(a.x, a.y); o.y = ptx::pack_bf16x2(a.z, a.w); o.z = ptx::pack_bf16x2(c.x, c.y); o.w = ptx::pack_bf16x2(c.z, c.w); }
    else { o.x = ptx::pack_f16x2(a.x, a.y); o.y = ptx::pack_f16x2(a.z, a.w); o.z = ptx::pack_f16x2(c.x, c.y); o.w = ptx::pack_f16x2(c.z, c.w); }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(dst) + 2 * (b * s0 + h * s1 + (int64_t)n * s2 + d)) = o;
  }
}

int launch_bwd_sm100(const ffpa_bwd_params& a, cudaStream_t stream) {
  const int D = a.head_dim;
  // head_dim > 512: two output-slab passes per item; box count rounded up to even (TMA zero-fills)
  const int nqk = D > 512 ? (D + 127) / 128 * 2 : (D + 63) / 64;
  const int n_pass = D > 512 ? 2 : 1;
  const int nq_pad = (a.seqlen_q + 127) / 128 * 128;
  // plan from the bytes actually granted (any size >= the minimum is valid; more memory = fewer GEMM passes)
  const StashPlan plan = (a.cu_seqlens_q || !a.workspace)
                             ? StashPlan{}
                             : stash_plan(a.batch, a.heads_q, a.heads_kv, a.seqlen_q, a.seqlen_kv, D, a.workspace_bytes);
  if (plan.chunked && a.bias_kind == FFPA_BIAS_NONE && !(a.dropout_p > 0.f) && a.d_bias == nullptr) {
    // stash buffers bounded by FFPA_BWD_STASH_MAX_GB: run (batch element, KV-head range) chunks one after the other
    // through the same scratch (stream order serialises them); each chunk is an ordinary dense sub-problem
    const int group = a.heads_q / a.heads_kv;
    {
      auto off = [](const void* p, int64_t elems) { return static_cast<const void*>(static_cast<const uint8_t*>(p) + 2 * elems); };
      for (int b = 0; b < a.batch; ++b)
        for (int hk0 = 0; hk0 < a.heads_kv; hk0 += plan.chunk_hkv) {
          const int hc = (a.heads_kv - hk0) < plan.chunk_hkv ? (a.heads_kv - hk0) : plan.chunk_hkv;
          const int hq0 = hk0 * group;
          ffpa_bwd_params s = a;
          s.batch = 1; s.heads_kv = hc; s.heads_q = hc * group;
          s.q = off(a.q, b * a.q_stride[0] + hq0 * a.q_stride[1]);
          s.o = off(a.o, b * a.o_stride[0] + hq0 * a.o_stride[1]);
          s.d_o = off(a.d_o, b * a.do_stride[0] + hq0 * a.do_stride[1]);
          s.dq = const_cast<void*>(off(a.dq, b * a.dq_stride[0] + hq0 * a.dq_stride[1]));
          s.k = off(a.k, b * a.k_stride[0] + hk0 * a.k_stride[1]);
          s.v = off(a.v, b * a.v_stride[0] + hk0 * a.v_stride[1]);
          s.dk = const_cast<void*>(off(a.dk, b * a.dk_stride[0] + hk0 * a.dk_stride[1]));
          s.dv = const_cast<void*>(off(a.dv, b * a.dv_stride[0] + hk0 * a.dv_stride[1]));
          s.lse = a.lse + ((int64_t)b * a.heads_q + hq0) * a.seqlen_q;
          if (a.d_lse) s.d_lse = a.d_lse + ((int64_t)b * a.heads_q + hq0) * a.seqlen_q;
          if (int rc = launch_bwd_sm100(s, stream)) return rc;
        }
      return FFPA_OK;
    }
  }
  const uint64_t need = align256(2ull * a.batch * a.heads_q * nq_pad * sizeof(float));  // lse2 + delta (split buffers optional)
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
  // packed variable-length mode: [total tokens, H, D] operands (map batch extent 1); Nq / Nkv are the maxima
  const bool varlen = a.cu_seqlens_q != nullptr;
  const int mb = varlen ? 1 : B, mq = varlen ? a.total_q : Nq, mk = varlen ? a.total_k : Nkv;
  if (!make_map4(&q_km, a.q, a.q_stride, mb, Hq, mq, D, 64, 64) || !make_map4(&q_mn, a.q, a.q_stride, mb, Hq, mq, D, 64, 128) ||
      !make_map4(&k_km, a.k, a.k_stride, mb, Hkv, mk, D, 64, 64) || !make_map4(&k_mn, a.k, a.k_stride, mb, Hkv, mk, D, 64, 128) ||
      !make_map4(&v_km, a.v, a.v_stride, mb, Hkv, mk, D, 64, 64) ||
      !make_map4(&do_km, a.d_o, a.do_stride, mb, Hq, mq, D, 64, 64) || !make_map4(&do_mn, a.d_o, a.do_stride, mb, Hq, mq, D, 64, 128))
    return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed in backward");

  // optional fp32 accumulation buffers for chunked dK / dV items
  float* dk32 = nullptr;
  float* dv32 = nullptr;
  {
    const uint64_t base = align256(2ull * a.batch * a.heads_q * nq_pad * sizeof(float));
    const uint64_t one = align256((uint64_t)a.batch * a.heads_kv * a.seqlen_kv * D * sizeof(float));
    if (!varlen && may_split_kv(a.batch, a.heads_kv, a.seqlen_kv) && a.workspace_bytes >= base + 2 * one) {
      dk32 = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + base);
      dv32 = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + base + one);
    }
  }

  bwd::BwdKernelParams kp{};
  kp.lse2 = lse2; kp.delta = delta; kp.nq_pad = nq_pad;
  kp.batch = B; kp.heads_q = Hq; kp.heads_kv = Hkv; kp.seqlen_q = Nq; kp.seqlen_kv = Nkv; kp.head_dim = D;
  kp.causal = a.causal; kp.scale = a.softmax_scale; kp.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  kp.bias = a.bias_kind != FFPA_BIAS_NONE ? a.bias : nullptr;
  kp.bias_kind = a.bias_kind;
  for (int i = 0; i < 4; ++i) kp.bias_stride[i] = a.bias_stride[i];
  kp.dropout_p = a.dropout_p; kp.philox_seed = a.philox_seed; kp.philox_offset = a.philox_offset;
  kp.dbias = a.d_bias;
  kp.zero_single = a.d_lse == nullptr;
  for (int i = 0; i < 4; ++i) kp.dbias_stride[i] = a.d_bias_stride[i];
  if (a.d_bias) {
    // the dQ kernel accumulates into the bias-shaped buffer (atomics wherever a dim is reduced): start from zero
    const int ext[4] = {a.batch, a.heads_q, a.seqlen_q, a.seqlen_kv};
    bool reduced = false;
    uint64_t elems = 1;
    for (int i = 0; i < 4; ++i) {
      if (a.d_bias_stride[i] == 0) reduced = reduced || ext[i] > 1;
      else elems *= (uint64_t)ext[i];
    }
    if (reduced) {
      cudaError_t e = cudaMemsetAsync(a.d_bias, 0, elems * sizeof(float), stream);
      if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync(d_bias): %s", cudaGetErrorString(e));
    }
  }
  kp.cu_q = a.cu_seqlens_q;
  kp.cu_k = a.cu_seqlens_k;
  kp.total_q = a.total_q; kp.total_k = a.total_k;
  // stash path: the dQ kernel writes P_drop / dS tiles, dK and dV become plain GEMMs over them
  const uint64_t stash_one = (varlen || plan.chunked) ? 0 : plan.one;
  const uint64_t stash_at = bwd_workspace_bytes_min(B, Hq, Hkv, Nq, Nkv, D);
  const bool use_stash = stash_one > 0 && a.workspace_bytes >= stash_at + 2 * stash_one;
  const int nk_pad = (Nkv + 255) / 256 * 256;
  void* stash_p = use_stash ? static_cast<uint8_t*>(a.workspace) + stash_at : nullptr;
  void* stash_ds = use_stash ? static_cast<uint8_t*>(a.workspace) + stash_at + stash_one : nullptr;
  // tile-major stash: [B * Hq][query tile][64-key block][128 queries][64 keys] as a 4-D tensor
  // (64, 128, blocks, B * Hq); a [64 x 64] store box / [128 x 64] load box is one contiguous run
  const int sblocks = (nq_pad / 128) * (nk_pad / 64);
  const int64_t sstr[3] = {(int64_t)sblocks * 8192, 8192, 64};
  CUtensorMap st_store, sp_store;   // [64 rows x 64 keys] boxes: what one CTA's T buffer holds per key half
  if (use_stash && (!make_map4(&st_store, stash_ds, sstr, B * Hq, sblocks, 128, 64, 64, 64) ||
                    !make_map4(&sp_store, stash_p, sstr, B * Hq, sblocks, 128, 64, 64, 64)))
    return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the backward stash store map");
  const int max_clusters = sm_count() / 2;
  const int off = Nkv - Nq;
  auto run = [&](int kind, void* out, const int64_t* ostride, int rows, int heads, float* acc32,
                 const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& b1, const CUtensorMap& b2,
                 const CUtensorMap& b3) -> int {
    kp.out = out;
    for (int i = 0; i < 3; ++i) kp.out_stride[i] = ostride[i];
    kp.n_rtiles = (rows + 127) / 128;
    kp.out_rows = rows;
    kp.stash_p = kind == 0 ? stash_p : nullptr;
    kp.stash_ds = kind == 0 ? stash_ds : nullptr;
    kp.nk_pad = nk_pad;
    // streamed tiles per item (same rule as col_tiles<KIND> in the kernel)
    const int group = Hq / Hkv;
    const int tq = (Nq + 127) / 128, tk = (Nkv + 127) / 128;
    std::vector<int> tfull(kp.n_rtiles);
    int tmax = 0;
    long long tsum = 0;
    for (int rt = 0; rt < kp.n_rtiles; ++rt) {
      int t;
      if (kind == 0) {
        t = tk;
        if (a.causal) {
          int lim = ((rt * 128 + 127 + off) >> 7) + 1;
          if (use_stash) { lim = (lim + 1) & ~1; t = (t + 1) & ~1; }   // same pairing rule as col_tiles<dQ>
          t = lim < t ? lim : t;
        }
        t = t < 1 ? 1 : t;
      } else {
        int first = 0;
        if (a.causal) { const int qmin = rt * 128 - off; first = qmin > 0 ? (qmin >> 7) : 0; first = first > tq ? tq : first; }
        t = (tq - first) * group;
      }
      tfull[rt] = t;
      tmax = t > tmax ? t : tmax;
      tsum += t;
    }
    const long long nbh = (long long)B * heads;
    const double avg = (double)tsum * nbh * n_pass / max_clusters;  // tiles per cluster if perfectly balanced
    kp.n_chunks = 1;
    kp.chunk_len = tmax;
    kp.out32 = nullptr;
    if (kind != 0 && acc32 != nullptr && tmax > avg / 2) {
      int len = (int)(avg / 4);
      len = len < 4 ? 4 : len;
      if (len < tmax) {
        kp.chunk_len = len;
        kp.n_chunks = (tmax + len - 1) / len;
        kp.out32 = acc32;
        cudaError_t e = cudaMemsetAsync(acc32, 0, (size_t)B * heads * rows * D * sizeof(float), stream);
        if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
      }
    }
    kp.n_pass = n_pass;
    kp.n_items = kp.n_rtiles * kp.n_chunks * (int)nbh * n_pass;
    const int ncl = kp.n_items < max_clusters ? kp.n_items : max_clusters;
    kp.sched = nullptr;
    kp.sched_stride = 0;
    if ((a.causal || kp.n_chunks > 1) && kp.n_items > ncl && !varlen) {
      std::vector<int> cost((size_t)kp.n_items);
      for (int it = 0; it < kp.n_items; ++it) {
        const int rt = it % kp.n_rtiles;
        const int chunk = (it / kp.n_rtiles) % kp.n_chunks;
        int n = tfull[rt] - chunk * kp.chunk_len;
        n = n < 0 ? 0 : (n < kp.chunk_len ? n : kp.chunk_len);
        cost[it] = n > 0 ? n * 16 + 24 : 1;
      }
      kp.sched = get_schedule(cost.data(), kp.n_items, ncl, &kp.sched_stride, stream);
    }
    const CUtensorMap& st = (kind == 0 && use_stash) ? st_store : a1;   // dS / P store maps (dQ kind, stash path)
    const CUtensorMap& sp = (kind == 0 && use_stash) ? sp_store : a1;
    int r = bf16 ? bwd::dispatch_bwd_dtype<true>(nqk, kind, a1, a2, b1, b2, b3, st, sp, kp, ncl, stream)
                 : bwd::dispatch_bwd_dtype<false>(nqk, kind, a1, a2, b1, b2, b3, st, sp, kp, ncl, stream);
    if (r) return r;
    if (kp.out32 != nullptr) {
      const int64_t total_vec = (int64_t)B * heads * rows * D / 8;
      const int blocks = (int)((total_vec + 255) / 256 < 4096 ? (total_vec + 255) / 256 : 4096);
      if (bf16) f32_to_16_kernel<true><<<blocks, 256, 0, stream>>>(acc32, out, ostride[0], ostride[1], ostride[2], heads, rows, D, total_vec);
      else f32_to_16_kernel<false><<<blocks, 256, 0, stream>>>(acc32, out, ostride[0], ostride[1], ostride[2], heads, rows, D, total_vec);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "fp32->16 convert launch failed: %s", cudaGetErrorString(e));
      count_launch();
    }
    return FFPA_OK;
  };
  if ((rc = run(0, a.dq, a.dq_stride, Nq, Hq, nullptr, q_km, do_km, k_km, v_km, k_mn))) return rc;
  if (use_stash) {
    CUtensorMap st_p, st_ds;
    if (!make_map4(&st_p, stash_p, sstr, B * Hq, sblocks, 128, 64, 64, 128) ||
        !make_map4(&st_ds, stash_ds, sstr, B * Hq, sblocks, 128, 64, 64, 128))
      return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the backward stash");
    bwd::BwdGemmParams gp{};
    gp.batch = B; gp.heads_q = Hq; gp.heads_kv = Hkv; gp.seqlen_q = Nq; gp.seqlen_kv = Nkv; gp.head_dim = D;
    gp.causal = a.causal;
    gp.n_kblocks = nk_pad / 256;
    gp.nk_pad = nk_pad;
    gp.n_pass = D > 512 ? 2 : 1;
    gp.n_items = gp.n_kblocks * B * Hkv * gp.n_pass;
    const int ncl = gp.n_items < max_clusters ? gp.n_items : max_clusters;
    gp.sched = nullptr;
    gp.sched_stride = 0;
    if (a.causal && gp.n_items > ncl) {
      const int tq = (Nq + 127) / 128, group = Hq / Hkv;
      std::vector<int> cost((size_t)gp.n_items);
      for (int it = 0; it < gp.n_items; ++it) {
        const int qmin = (it % gp.n_kblocks) * 256 - off;
        int first = qmin > 0 ? (qmin >> 7) : 0;
        first = first > tq ? tq : first;
        const int t = (tq - first) * group;
        cost[it] = t > 0 ? t * 16 + 24 : 1;
      }
      gp.sched = get_schedule(cost.data(), gp.n_items, ncl, &gp.sched_stride, stream);
    }
    auto gemm = [&](const CUtensorMap& mt, const CUtensorMap& mb, void* out, const int64_t* ostride, float mul) -> int {
      gp.out = out;
      for (int i = 0; i < 3; ++i) gp.out_stride[i] = ostride[i];
      gp.mul = mul;
      return bf16 ? bwd::launch_bwd_gemm<true>(mt, mb, gp, ncl, stream) : bwd::launch_bwd_gemm<false>(mt, mb, gp, ncl, stream);
    };
    if ((rc = gemm(st_p, do_mn, a.dv, a.dv_stride, 1.f))) return rc;
    return gemm(st_ds, q_mn, a.dk, a.dk_stride, a.softmax_scale);
  }
  if ((rc = run(1, a.dk, a.dk_stride, Nkv, Hkv, dk32, k_km, v_km, q_km, do_km, q_mn))) return rc;
  if ((rc = run(2, a.dv, a.dv_stride, Nkv, Hkv, dv32, k_km, k_km, q_km, q_km, do_mn))) return rc;
  return FFPA_OK;
}

}  // namespace ffpa
