/* ffpa_b200.h -- C ABI of the B200-native (sm_100a) Split-D attention engine.
 *
 * This is the drop-in boundary for the reference's native backend. The reference binds its CUDA
 * backend through a pybind11 module `ffpa_attn._C`
 *   (/root/reference/csrc/cuffpa/ffpa_api.cc:86-96  ffpa_attn_forward,
 *    /root/reference/csrc/cuffpa/ffpa_api.cc:242-246 ffpa_attn_backward (a stub that throws),
 *    /root/reference/csrc/cuffpa/backend.h:6-27      set/get_cuda_backend_impl)
 * whose arguments are torch tensors. The entry points below carry exactly the same information
 * as plain device pointers + sizes + strides, so they can be bound from pybind/torch
 * (ffpa-attn_b200/csrc/ffpa_torch_binding.cpp builds the real `ffpa_attn._C`),
 * ctypes (tests/test_host.py) or any other FFI.  Everything that changes results lives BEHIND this
 * boundary -- FP8 hybrid staging, workspace planning, bias broadcast, dBias reduction -- so every
 * binder gets the same behaviour; the binder only supplies device memory.
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

#define FFPA_B200_ABI_VERSION 3

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

/* FP8 knob codes of the reference's op signature (/root/reference/src/ffpa_attn/functional.py:46-67) */
enum { FFPA_QUANT_PER_BLOCK = 0, FFPA_QUANT_PER_CHANNEL = 1, FFPA_QUANT_PER_THREAD = 2 };
enum { FFPA_PV_ACC_F16 = 0, FFPA_PV_ACC_F32 = 1 };
enum { FFPA_QK_MM_FP8 = 0, FFPA_QK_MM_INT8 = 1 };

typedef struct ffpa_fwd_params {
  /* tensors (device pointers) */
  const void* q;
  const void* k;
  const void* v;
  void* o;
  float* lse;        /* fp32 [B, Hq, Nq] (rows of one (b, h) contiguous, see lse_bh_stride); NULL = not written */
  const void* bias;  /* NULL when bias_kind == FFPA_BIAS_NONE */
  /* element strides, order (b, h, n, d); d-stride must be 1 */
  int64_t q_stride[4];
  int64_t k_stride[4];
  int64_t v_stride[4];
  int64_t o_stride[4];
  int64_t bias_stride[4]; /* (b, h, q, k); 0 on broadcast dims; k-stride must be 1, or 0 for a [.., .., .., 1] bias */
  /* sizes */
  int32_t batch, heads_q, heads_kv, seqlen_q, seqlen_kv, head_dim;
  int32_t dtype;     /* FFPA_DTYPE_* */
  int32_t bias_kind; /* FFPA_BIAS_* */
  int32_t causal;    /* 0 / 1 */
  int32_t impl;      /* kernel family of THIS call: FFPA_IMPL_AUTO = follow the calling thread's hint
                        (ffpa_b200_set_backend_impl); 1..4 -> fp16/bf16 tcgen05 kernel; FFPA_IMPL_CUTE_TMA_FP8 -> per-tile
                        e4m3 quantised kernel (fp8_* fields below); FFPA_IMPL_CUTE_TMA_FP4 -> FFPA_ERR_UNSUPPORTED */
  float softmax_scale;
  float dropout_p;
  uint64_t philox_seed;
  uint64_t philox_offset;
  /* scratch: ffpa_b200_fwd_workspace_bytes_p(params, cap) bytes of 256-byte aligned device memory (FP8 copies and
   * scales, KV-split partials of decode-like shapes, the replay stash of head dims > 768); may be NULL when
   * that function returns 0. Given less than asked for, paths that need no scratch run where they exist
   * (no KV split, two-pass instead of replay); the FP8 path fails with FFPA_ERR_INVALID_ARGUMENT. */
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
  /* ABI 3: the FP8 arguments of ffpa_attn_forward (/root/reference/csrc/cuffpa/ffpa_api.cc:86-96), read when the
   * call resolves to FFPA_IMPL_CUTE_TMA_FP8. Implemented: smooth-K, smooth-V (needs per-channel V, as in the
   * reference), per-block Q/K, per-block or per-channel V, f32 PV accumulation (TMEM), e4m3 QK, hybrid.
   * per_thread Q/K, int8 QK and the f16 PV accumulator are sm_120 mma.sync variants: FFPA_ERR_UNSUPPORTED, naming the knob. */
  int32_t fp8_smooth_k, fp8_smooth_v;
  int32_t fp8_q_quant_method, fp8_k_quant_method, fp8_v_quant_method; /* FFPA_QUANT_* */
  int32_t fp8_pv_acc_type;                                              /* FFPA_PV_ACC_* */
  int32_t fp8_qk_mm_type;                                               /* FFPA_QK_MM_* */
  /* fp8_hybrid != 0: query rows [0, n_early) run on the fp16/bf16 kernel, rows [n_early, Nq) on the FP8 kernel
   * (/root/reference/csrc/cuffpa/launch.cuh:341-374), both as zero-copy row views inside this one call; honoured
   * for causal and non-causal calls alike; n_early must be a positive multiple of 128 (ignored when >= Nq). */
  int32_t fp8_hybrid, fp8_hybrid_n_early;
  /* elements between the LSE rows of consecutive (b, h) pairs; 0 = seqlen_q (contiguous [B, Hq, Nq]) */
  int64_t lse_bh_stride;
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
  /* scratch: at least ffpa_b200_bwd_workspace_bytes_min_p() bytes, 256-byte aligned device memory */
  void* workspace;
  uint64_t workspace_bytes;
  /* optional pieces of the forward that must be replayed (all zero / NULL when unused):
   * additive bias (same conventions as ffpa_fwd_params), dropout (same Philox seed/offset as the
   * forward), and d_bias: fp32 buffer of the BIAS' OWN (broadcast) shape [1|B, 1|Hq, 1|Nq, 1|Nkv], contiguous,
   * receives dBias = sum over the broadcast dims of P*(dP - delta); the reduction happens inside the dQ kernel
   * (warp-level column sums + fp32 atomics), the library zero-fills the buffer first; NULL = not wanted. */
  const void* bias;
  int64_t bias_stride[4];
  int32_t bias_kind;
  float dropout_p;
  uint64_t philox_seed;
  uint64_t philox_offset;
  float* d_bias;
  int64_t d_bias_stride[4]; /* (b, h, q, k) element strides of d_bias; 0 on the dims the bias broadcasts over */
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

/* Scratch bytes ffpa_b200_fwd wants for exactly this call (pointers are not read; impl / fp8_* / hybrid / sizes
 * are). The FP8 buffers are REQUIRED and returned whatever the cap; the optional scratch -- KV-split partials, and
 * for head dims > 768 the O(Nq*Nkv)-per-head replay stash -- is planned to fit `cap_bytes`: the whole problem, else
 * (batch, KV-head) chunks through one scratch, else none (two-pass kernel). The launcher plans from the bytes it is
 * actually given, so any size is valid. */
uint64_t ffpa_b200_fwd_workspace_bytes_p(const ffpa_fwd_params* p, uint64_t cap_bytes);

/* replaces ffpa_attn_backward (/root/reference/csrc/cuffpa/ffpa_api.cc:242-263, a thrower there) */
int ffpa_b200_bwd(const ffpa_bwd_params* p, void* stream);
/* Scratch sizes of the backward for exactly this call. `_min` is the REQUIRED size (O(N): lse2 + delta, plus fp32
 * dK/dV accumulators for shapes with few KV-stationary items). `_p(params, cap)` is the RECOMMENDED size, never
 * above max(cap, min): for head dims 384..1024 it adds the two 16-bit score buffers of the stash path (dQ kernel
 * stores P / dS tiles, dK and dV run as plain GEMMs over them: 5 GEMM passes instead of 8), chunked over
 * (batch, KV-head range) when the whole problem does not fit under `cap`. cap == 0 asks for the minimum. The
 * launcher plans from the bytes it is actually given, so any size >= `_min` is valid. */
uint64_t ffpa_b200_bwd_workspace_bytes_p(const ffpa_bwd_params* p, uint64_t cap_bytes);
uint64_t ffpa_b200_bwd_workspace_bytes_min_p(const ffpa_bwd_params* p);

/* replaces set_cuda_backend_impl / get_cuda_backend_impl (ffpa_api.cc:272-282, backend.h:16-25). The hint is
 * THREAD-LOCAL here (the reference's process-global atomic races between threads that use different
 * backends); a call with params.impl != FFPA_IMPL_AUTO does not read it at all. */
int ffpa_b200_set_backend_impl(int32_t impl);
int32_t ffpa_b200_get_backend_impl(void);

/* capability flags = the module attributes of ffpa_api.cc:283-305 */
int32_t ffpa_b200_fwd_available(void);
int32_t ffpa_b200_bwd_available(void);
int32_t ffpa_b200_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t ffpa_b200_launch_count(void);

const char* ffpa_b200_last_error(void);
/* The tuning variables FFPA_FWD_REPLAY, FFPA_FWD_REPLAY_MAX_GB and FFPA_BWD_STASH are read once per process
 * (not on every launch); call this after changing them at run time. */
void ffpa_b200_refresh_env(void);

#ifdef __cplusplus
}
#endif
#endif /* FFPA_B200_H_ */
