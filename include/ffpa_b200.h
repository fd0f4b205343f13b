/* ffpa_b200.h -- C ABI of the B200-native (sm_100a) Split-D attention engine.
 *
 * This is the drop-in boundary for the reference's native backend. The reference binds its CUDA
 * backend through a pybind11 module `ffpa_attn._C`
 *   (/root/reference/csrc/cuffpa/ffpa_api.cc:86-96  ffpa_attn_forward,
 *    /root/reference/csrc/cuffpa/ffpa_api.cc:242-246 ffpa_attn_backward (a stub that throws),
 *    /root/reference/csrc/cuffpa/backend.h:6-27      set/get_cuda_backend_impl)
 * whose arguments are torch tensors. The entry points below carry exactly the same information
 * as plain device pointers + sizes + strides, so they can be bound from pybind/torch (see
 * INTEGRATION.md), ctypes (ffpa-attn_b200/ffpa_attn/cuda/_C.py) or any other FFI.
 *
 * Conventions (identical to the reference contract):
 *   Q  [B, Hq,  Nq,  D]   K,V [B, Hkv, Nkv, D]   O like Q   LSE fp32 [B, Hq, Nq] natural log
 *   dtype fp16 or bf16, last dim contiguous (stride 1), other strides in ELEMENTS and
 *   multiples of 8 (16 bytes, a TMA requirement); Hq % Hkv == 0 (GQA: kv_head = q_head / group);
 *   causal is bottom-right aligned (key k visible to row r iff k <= r + Nkv - Nq), needs Nkv>=Nq;
 *   attn bias is additive, applied after scaling, broadcast through zero strides, fp32 or the
 *   Q dtype, mutually exclusive with causal; dropout uses Philox-4x32-10 keyed by (seed, offset)
 *   with element index ((b*Hq+h)*Nq+q)*Nkv+k, keep iff u > p, scale 1/(1-p).
 *   All work is enqueued on `stream` (a cudaStream_t); no host synchronisation.
 *
 * Every function returns 0 on success, a negative FFPA_ERR_* code otherwise;
 * ffpa_b200_last_error() returns a thread-local message for the last failure.
 */
#ifndef FFPA_B200_H_
#define FFPA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFPA_B200_ABI_VERSION 2

enum {
  FFPA_OK = 0,
  FFPA_ERR_INVALID_ARGUMENT = -1, /* shape / dtype / stride contract violated (TORCH_CHECK class) */
  FFPA_ERR_UNSUPPORTED = -2,      /* valid request this build does not implement               */
  FFPA_ERR_CUDA = -3,             /* CUDA runtime / driver error                                */
  FFPA_ERR_NO_DEVICE = -4         /* device is not sm_100                                       */
};

enum { FFPA_DTYPE_F16 = 0, FFPA_DTYPE_BF16 = 1 };
enum { FFPA_BIAS_NONE = 0, FFPA_BIAS_F32 = 1, FFPA_BIAS_QDTYPE = 2 };

/* mirrors ffpa::CudaBackendImpl (/root/reference/csrc/cuffpa/backend.h:6-14); advisory here:
 * 0..4 -> the sm_100a bf16/fp16 kernel, 5 -> FP8 kernel, 6 -> unsupported. */
enum {
  FFPA_IMPL_AUTO = 0, FFPA_IMPL_NATIVE = 1, FFPA_IMPL_TMA = 2, FFPA_IMPL_CUTE = 3,
  FFPA_IMPL_CUTE_TMA = 4, FFPA_IMPL_CUTE_TMA_FP8 = 5, FFPA_IMPL_CUTE_TMA_FP4 = 6
};

typedef struct ffpa_fwd_params {
  /* tensors (device pointers) */
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;        /* [B, Hq, Nq] contiguous fp32; may be NULL (not written) */
  const void* bias;  /* NULL when bias_kind == FFPA_BIAS_NONE */
  /* element strides, order (b, h, n, d); d-stride must be 1 */
  int64_t q_stride[4];
  int64_t k_stride[4];
  int64_t v_stride[4];
  int64_t o_stride[4];
  int64_t bias_stride[4]; /* (b, h, q, k); 0 on broadcast dims; k-stride must be 1 */
  /* sizes */
  int32_t batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int32_t dtype;     /* FFPA_DTYPE_* */
  int32_t bias_kind; /* FFPA_BIAS_* */
  int32_t causal;    /* 0 / 1 */
  int32_t fp8;       /* 0: fp16/bf16 MMA; bit 0: per-tile e4m3 quantised MMA (FFPA_IMPL_CUTE_TMA_FP8);
                        bit 1: smooth-K (quantise K - mean_seq(K), LSE corrected; the reference's default);
                        bit 2: smooth-V (quantise V - mean_seq(V), mean added back to O);
                        bit 3: per-channel V scales instead of per-128-row-block ones */
  float softmax_scale;
  float dropout_p;
  uint64_t philox_seed;
  uint64_t philox_offset;
  /* scratch for the FP8 path (quantised copies + scales): ffpa_b200_fwd_workspace_bytes() bytes,
   * 256-byte aligned device memory; ignored (may be NULL) when fp8 == 0 */
  void* workspace;
  uint64_t workspace_bytes;
  /* Packed variable-length mode (ABI 2; replaces the reference's ffpa_attn_varlen_func backend,
   * /root/reference/src/ffpa_attn/ffpa_attn_interface.py:192-279). All NULL / 0 for the dense layout.
   * When cu_seqlens_q != NULL: q/o are [total_q, Hq, D] and k/v [total_k, Hkv, D] (stride[1] = head,
   * stride[2] = token, stride[0] ignored); batch = number of sequences; seqlen_q / seqlen_kv = the MAXIMUM
   * per-sequence lengths; cu_seqlens_*: int32 DEVICE arrays of batch + 1 token offsets (never read on the
   * host: no synchronisation); LSE is [Hq, total_q]; causal is bottom-right aligned per sequence (rows
   * that see no key give O = 0, LSE = -inf). One launch for the whole batch. No bias / dropout / fp8. */
  const int32_t* cu_seqlens_q;
  const int32_t* cu_seqlens_k;
  int32_t total_q, total_k;
} ffpa_fwd_params;

typedef struct ffpa_bwd_params {
  const void* q;
  const void* k;
  const void* v;
  const void* o;
  const float* lse;
  const void* d_o;
  void* dq;
  void* dk;
  void* dv;
  int64_t q_stride[4], k_stride[4], v_stride[4], o_stride[4], do_stride[4];
  int64_t dq_stride[4], dk_stride[4], dv_stride[4];
  int32_t batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int32_t dtype;
  int32_t causal;
  float softmax_scale;
  /* scratch: fp32 workspace of ffpa_b200_bwd_workspace_bytes() bytes (device) */
  void* workspace;
  uint64_t workspace_bytes;
  /* optional pieces of the forward that must be replayed (all zero / NULL when unused):
   * additive bias (same conventions as ffpa_fwd_params), dropout (same Philox seed/offset as the
   * forward), and d_bias: fp32 [B, Hq, Nq, Nkv] contiguous, receives dS = P*(dP - delta) per
   * score (the caller reduces over the bias' broadcast dims); NULL = not wanted. */
  const void* bias;
  int64_t bias_stride[4];
  int32_t bias_kind;
  float dropout_p;
  uint64_t philox_seed;
  uint64_t philox_offset;
  float* d_bias;
  /* packed variable-length mode, same conventions as ffpa_fwd_params (lse is [Hq, total_q]); no bias /
   * dropout / d_bias */
  const int32_t* cu_seqlens_q;
  const int32_t* cu_seqlens_k;
  int32_t total_q, total_k;
  /* optional gradient of the loss w.r.t. the LSE output (same shape as lse; NULL = zero): dS gains P * dLSE,
   * folded into delta by the preprocess kernel (reference: cute/_bwd_preprocess.py:6-15) */
  const float* d_lse;
} ffpa_bwd_params;

/* replaces ffpa_attn_forward (/root/reference/csrc/cuffpa/ffpa_api.cc:86-239) */
int ffpa_b200_fwd(const ffpa_fwd_params* p, void* stream);

/* scratch bytes ffpa_b200_fwd needs for these sizes (0 unless fp8 != 0) */
uint64_t ffpa_b200_fwd_workspace_bytes(int32_t batch, int32_t heads_q, int32_t heads_kv,
                                       int32_t seqlen_q, int32_t seqlen_kv, int32_t head_dim,
                                       int32_t fp8);

/* replaces ffpa_attn_backward (/root/reference/csrc/cuffpa/ffpa_api.cc:242-263, a thrower there) */
int ffpa_b200_bwd(const ffpa_bwd_params* p, void* stream);
/* Scratch sizes of the backward. `_bytes` is the RECOMMENDED size: for head dims 384..512 it includes the
 * two 16-bit [B, Hq, Nq_pad, Nk_pad] score buffers of the stash path (dQ kernel stores P / dS tiles, dK and
 * dV run as plain GEMMs over them: 5 GEMM passes instead of 8). `_bytes_min` is the REQUIRED size; given
 * less than the recommended size the three recompute kernels run instead (O(N) memory). The packed
 * variable-length mode only ever needs the minimum. */
uint64_t ffpa_b200_bwd_workspace_bytes(int32_t batch, int32_t heads_q, int32_t heads_kv,
                                       int32_t seqlen_q, int32_t seqlen_kv, int32_t head_dim);
uint64_t ffpa_b200_bwd_workspace_bytes_min(int32_t batch, int32_t heads_q, int32_t heads_kv,
                                           int32_t seqlen_q, int32_t seqlen_kv, int32_t head_dim);

/* replaces set_cuda_backend_impl / get_cuda_backend_impl (ffpa_api.cc:272-282, backend.h:16-25) */
int ffpa_b200_set_backend_impl(int32_t impl);
int32_t ffpa_b200_get_backend_impl(void);

/* capability flags = the module attributes of ffpa_api.cc:283-305 */
int32_t ffpa_b200_fwd_available(void);
int32_t ffpa_b200_bwd_available(void);
int32_t ffpa_b200_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t ffpa_b200_launch_count(void);

const char* ffpa_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FFPA_B200_H_ */
