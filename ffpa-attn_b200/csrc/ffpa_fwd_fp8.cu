// FP8 forward: host launcher (workspace carving, fused quantisation pre-pass, tensor maps,
// dispatch). Entry: launch_fwd_fp8_sm100 (called from ffpa_b200_fwd when the call resolves to FFPA_IMPL_CUTE_TMA_FP8).
#include "ffpa_fwd_fp8_sm100.cuh"

namespace ffpa {
namespace fp8 {
#define FFPA_FP8_EXTERN(NB, BF, W) \
  extern template int launch_fp8_variant<NB, BF, W>(const CUtensorMap&, const CUtensorMap&, const CUtensorMap&, const Fp8KernelParams&, int, cudaStream_t);
FFPA_FP8_EXTERN(1, true, 2) FFPA_FP8_EXTERN(2, true, 2) FFPA_FP8_EXTERN(3, true, 2) FFPA_FP8_EXTERN(4, true, 2)
FFPA_FP8_EXTERN(1, true, 4) FFPA_FP8_EXTERN(2, true, 4) FFPA_FP8_EXTERN(3, true, 4) FFPA_FP8_EXTERN(4, true, 4)
FFPA_FP8_EXTERN(1, false, 2) FFPA_FP8_EXTERN(2, false, 2) FFPA_FP8_EXTERN(3, false, 2) FFPA_FP8_EXTERN(4, false, 2)
FFPA_FP8_EXTERN(1, false, 4) FFPA_FP8_EXTERN(2, false, 4) FFPA_FP8_EXTERN(3, false, 4) FFPA_FP8_EXTERN(4, false, 4)
#undef FFPA_FP8_EXTERN
}  // namespace fp8

static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

struct Fp8Layout {
  int dpad, tq, tk;
  uint64_t off_q8, off_k8, off_v8, off_qs, off_ks, off_vs, off_vref, off_ksum, off_vsum, off_vamax, off_qkm, total;
};

static Fp8Layout fp8_layout(int B, int Hq, int Hkv, int Nq, int Nkv, int D) {
  Fp8Layout L{};
  L.dpad = (D + 15) / 16 * 16;
  L.tq = (Nq + 127) / 128;
  L.tk = (Nkv + 127) / 128;
  uint64_t o = 0;
  L.off_q8 = o; o = align_up(o + (uint64_t)B * Hq * Nq * L.dpad, 256);
  L.off_k8 = o; o = align_up(o + (uint64_t)B * Hkv * Nkv * L.dpad, 256);
  L.off_v8 = o; o = align_up(o + (uint64_t)B * Hkv * Nkv * L.dpad, 256);
  L.off_qs = o; o = align_up(o + (uint64_t)B * Hq * L.tq * 4, 256);
  L.off_ks = o; o = align_up(o + (uint64_t)B * Hkv * L.tk * 4, 256);
  L.off_vs = o; o = align_up(o + (uint64_t)B * Hkv * L.tk * 4, 256);
  L.off_vref = o; o = align_up(o + (uint64_t)B * Hkv * 4, 256);
  L.off_ksum = o; o = align_up(o + (uint64_t)B * Hkv * D * 4, 256);
  L.off_vsum = o; o = align_up(o + (uint64_t)B * Hkv * D * 4, 256);
  L.off_vamax = o; o = align_up(o + (uint64_t)B * Hkv * D * 4, 256);
  L.off_qkm = o; o = align_up(o + (uint64_t)B * Hq * Nq * 4, 256);
  L.total = o;
  return L;
}

uint64_t fwd_fp8_workspace_bytes(int B, int Hq, int Hkv, int Nq, int Nkv, int D) {
  return fp8_layout(B, Hq, Hkv, Nq, Nkv, D).total;
}

static bool make_map8(CUtensorMap* m, const void* base, int B, int H, int N, int dpad, uint32_t box_n) {
  uint64_t dims[4] = {(uint64_t)dpad, (uint64_t)N, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)dpad, (uint64_t)dpad * N, (uint64_t)dpad * N * H};
  uint32_t box[4] = {128, box_n, 1, 1};
  return tmap::encode_sw128(m, const_cast<void*>(base), 1, 4, dims, str, box);
}

// softmax warpgroups per CTA: 4 (tiles round robin over four warpgroups) unless FFPA_FP8_NWG=2 asks for the
// two-warpgroup variant (A/B knob; read once per process)
template <bool OUT_BF16, int NWG>
static int dispatch_nb(int nb, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                       const fp8::Fp8KernelParams& kp, int ncl, cudaStream_t s) {
  switch (nb) {
    case 1: return fp8::launch_fp8_variant<1, OUT_BF16, NWG>(mq, mk, mv, kp, ncl, s);
    case 2: return fp8::launch_fp8_variant<2, OUT_BF16, NWG>(mq, mk, mv, kp, ncl, s);
    case 3: return fp8::launch_fp8_variant<3, OUT_BF16, NWG>(mq, mk, mv, kp, ncl, s);
    case 4: return fp8::launch_fp8_variant<4, OUT_BF16, NWG>(mq, mk, mv, kp, ncl, s);
    default: return set_error(FFPA_ERR_UNSUPPORTED, "FP8 forward supports head_dim <= 512");
  }
}

int launch_fwd_fp8_sm100(const ffpa_fwd_params& a, int fp8_bits, cudaStream_t stream) {
  const int B = a.batch, Hq = a.heads_q, Hkv = a.heads_kv, Nq = a.seqlen_q, Nkv = a.seqlen_kv, D = a.head_dim;
  if (D > 512) return set_error(FFPA_ERR_UNSUPPORTED, "FP8 forward supports head_dim <= 512 (got %d)", D);
  if (a.bias_kind != FFPA_BIAS_NONE || a.dropout_p > 0.f)
    return set_error(FFPA_ERR_UNSUPPORTED, "FP8 forward does not implement attn bias / dropout");
  if (!(a.softmax_scale > 0.f)) return set_error(FFPA_ERR_UNSUPPORTED, "FP8 forward needs softmax_scale > 0");
  const Fp8Layout L = fp8_layout(B, Hq, Hkv, Nq, Nkv, D);
  if (!a.workspace || a.workspace_bytes < L.total)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "FP8 forward workspace too small: need %llu bytes (ffpa_b200_fwd_workspace_bytes)",
                     (unsigned long long)L.total);
  if (reinterpret_cast<uintptr_t>(a.workspace) & 255u)
    return set_error(FFPA_ERR_INVALID_ARGUMENT, "FP8 forward workspace must be 256-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(a.workspace);

  fp8::QuantArgs qa{};
  qa.src[0] = a.q; qa.src[1] = a.k; qa.src[2] = a.v;
  qa.dst[0] = ws + L.off_q8; qa.dst[1] = ws + L.off_k8; qa.dst[2] = ws + L.off_v8;
  qa.scale[0] = reinterpret_cast<float*>(ws + L.off_qs);
  qa.scale[1] = reinterpret_cast<float*>(ws + L.off_ks);
  qa.scale[2] = reinterpret_cast<float*>(ws + L.off_vs);
  qa.vref = reinterpret_cast<float*>(ws + L.off_vref);
  for (int i = 0; i < 3; ++i) { qa.stride[0][i] = a.q_stride[i]; qa.stride[1][i] = a.k_stride[i]; qa.stride[2][i] = a.v_stride[i]; }
  qa.heads[0] = Hq; qa.heads[1] = Hkv; qa.heads[2] = Hkv;
  qa.seqlen[0] = Nq; qa.seqlen[1] = Nkv; qa.seqlen[2] = Nkv;
  qa.tiles[0] = L.tq; qa.tiles[1] = L.tk; qa.tiles[2] = L.tk;
  qa.first_block[0] = 0;
  qa.first_block[1] = (int64_t)B * Hq * L.tq;
  qa.first_block[2] = qa.first_block[1] + (int64_t)B * Hkv * L.tk;
  qa.first_block[3] = qa.first_block[2] + (int64_t)B * Hkv * L.tk;
  qa.batch = B; qa.head_dim = D; qa.dpad = L.dpad;
  cudaError_t e = cudaMemsetAsync(qa.vref, 0, (size_t)B * Hkv * 4, stream);
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  // smooth-K (fp8_bits bit 1; the reference's default, functional.py:246): quantise K - mean_seq(K) and shift
  // the LSE back by scale * q . mean
  const bool smooth_k = (fp8_bits & 2) != 0;
  float* ksum = reinterpret_cast<float*>(ws + L.off_ksum);
  float* qkm = reinterpret_cast<float*>(ws + L.off_qkm);
  qa.ksum = nullptr;
  if (smooth_k) {
    e = cudaMemsetAsync(ksum, 0, (size_t)B * Hkv * D * 4, stream);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int rpb = 256;
    const unsigned nblk = (unsigned)((Nkv + rpb - 1) / rpb) * Hkv * B;
    if (a.dtype == FFPA_DTYPE_BF16)
      fp8::k_colsum_kernel<true><<<nblk, 256, 0, stream>>>(a.k, ksum, a.k_stride[0], a.k_stride[1], a.k_stride[2], Hkv, Nkv, D, rpb);
    else
      fp8::k_colsum_kernel<false><<<nblk, 256, 0, stream>>>(a.k, ksum, a.k_stride[0], a.k_stride[1], a.k_stride[2], Hkv, Nkv, D, rpb);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "smooth-K pre-pass launch failed: %s", cudaGetErrorString(e));
    count_launch();
    qa.qkm = qkm;   // q . mean_seq(K) per query row: emitted by the Q blocks of the quantiser
    qa.ksum = ksum;
  }
  // smooth-V (fp8_bits bit 2; reference knob fp8_smooth_v): quantise V - mean_seq(V), add the mean back to O
  const bool smooth_v = (fp8_bits & 4) != 0;
  float* vsum = reinterpret_cast<float*>(ws + L.off_vsum);
  qa.vsum = nullptr;
  if (smooth_v) {
    e = cudaMemsetAsync(vsum, 0, (size_t)B * Hkv * D * 4, stream);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int rpb = 256;
    const unsigned nblk = (unsigned)((Nkv + rpb - 1) / rpb) * Hkv * B;
    if (a.dtype == FFPA_DTYPE_BF16)
      fp8::k_colsum_kernel<true><<<nblk, 256, 0, stream>>>(a.v, vsum, a.v_stride[0], a.v_stride[1], a.v_stride[2], Hkv, Nkv, D, rpb);
    else
      fp8::k_colsum_kernel<false><<<nblk, 256, 0, stream>>>(a.v, vsum, a.v_stride[0], a.v_stride[1], a.v_stride[2], Hkv, Nkv, D, rpb);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "smooth-V pre-pass launch failed: %s", cudaGetErrorString(e));
    count_launch();
    qa.vsum = vsum;
  }
  // per-channel V scales (fp8_bits bit 3; reference knob fp8_v_quant_method="per_channel")
  const bool v_per_channel = (fp8_bits & 8) != 0;
  float* vamax = reinterpret_cast<float*>(ws + L.off_vamax);
  qa.vamax = nullptr;
  if (v_per_channel) {
    e = cudaMemsetAsync(vamax, 0, (size_t)B * Hkv * D * 4, stream);
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
    const int rpb = 256;
    const unsigned nblk = (unsigned)((Nkv + rpb - 1) / rpb) * Hkv * B;
    if (a.dtype == FFPA_DTYPE_BF16)
      fp8::v_colamax_kernel<true><<<nblk, 256, 0, stream>>>(a.v, qa.vsum, vamax, a.v_stride[0], a.v_stride[1], a.v_stride[2], Hkv, Nkv, D, rpb);
    else
      fp8::v_colamax_kernel<false><<<nblk, 256, 0, stream>>>(a.v, qa.vsum, vamax, a.v_stride[0], a.v_stride[1], a.v_stride[2], Hkv, Nkv, D, rpb);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "per-channel V pre-pass launch failed: %s", cudaGetErrorString(e));
    count_launch();
    qa.vamax = vamax;
  }
  if (a.dtype == FFPA_DTYPE_BF16)
    fp8::quantize_e4m3_kernel<true><<<dim3((unsigned)qa.first_block[3]), dim3(fp8::kQuantThreads), 0, stream>>>(qa);
  else
    fp8::quantize_e4m3_kernel<false><<<dim3((unsigned)qa.first_block[3]), dim3(fp8::kQuantThreads), 0, stream>>>(qa);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FFPA_ERR_CUDA, "fp8 quantise launch failed: %s", cudaGetErrorString(e));
  count_launch();

  CUtensorMap mq, mk, mv;
  if (!make_map8(&mq, qa.dst[0], B, Hq, Nq, L.dpad, 64) || !make_map8(&mk, qa.dst[1], B, Hkv, Nkv, L.dpad, 64) ||
      !make_map8(&mv, qa.dst[2], B, Hkv, Nkv, L.dpad, 128))
    return set_error(FFPA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the fp8 tensors");

  fp8::Fp8KernelParams kp{};
  kp.o = a.o; kp.lse = a.lse;
  kp.lse_bh_stride = a.lse_bh_stride > 0 ? a.lse_bh_stride : Nq;
  for (int i = 0; i < 3; ++i) kp.o_stride[i] = a.o_stride[i];
  kp.qs = qa.scale[0]; kp.ks = qa.scale[1]; kp.vs = qa.scale[2]; kp.vref = qa.vref;
  kp.qkm = smooth_k ? qkm : nullptr;
  kp.vsum = smooth_v ? vsum : nullptr;
  kp.vamax = v_per_channel ? vamax : nullptr;
  kp.tq = L.tq; kp.tk = L.tk;
  kp.batch = B; kp.heads_q = Hq; kp.heads_kv = Hkv; kp.seqlen_q = Nq; kp.seqlen_kv = Nkv; kp.head_dim = D;
  kp.causal = a.causal;
  kp.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  kp.n_mtiles = L.tq;
  kp.n_items = kp.n_mtiles * B * Hq;
  int ncl = sm_count() / 2;
  if (ncl > kp.n_items) ncl = kp.n_items;
  const int nb = (D + 127) / 128;
  const bool two = env_gb("FFPA_FP8_NWG", 4.0) == 2.0;
  if (two)
    return a.dtype == FFPA_DTYPE_BF16 ? dispatch_nb<true, 2>(nb, mq, mk, mv, kp, ncl, stream)
                                      : dispatch_nb<false, 2>(nb, mq, mk, mv, kp, ncl, stream);
  return a.dtype == FFPA_DTYPE_BF16 ? dispatch_nb<true, 4>(nb, mq, mk, mv, kp, ncl, stream)
                                    : dispatch_nb<false, 4>(nb, mq, mk, mv, kp, ncl, stream);
}

}  // namespace ffpa
