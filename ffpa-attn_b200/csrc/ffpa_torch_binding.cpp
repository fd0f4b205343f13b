// ffpa_attn._C -- the PyTorch C++ extension of the package, B200 edition.
//
// Exports exactly the pybind surface of the reference's native module
//   /root/reference/csrc/cuffpa/ffpa_api.cc:265-306   (PYBIND11_MODULE)
//   /root/reference/csrc/cuffpa/ffpa_api.cc:86-96     ffpa_attn_forward, all 24 arguments honoured
//   /root/reference/csrc/cuffpa/ffpa_api.cc:242-246   ffpa_attn_backward (a thrower there, real here)
//   /root/reference/csrc/cuffpa/backend.h:6-27        set/get_cuda_backend_impl
// so the UNCHANGED reference Python package (src/ffpa_attn, import site cuda/__init__.py:6-25) runs on the
// sm_100a kernels when this module is dropped into it (tests/test_dropin_gpu.py does exactly that).
// It is C++ only: every result-changing decision (kernel family, FP8 knobs, hybrid staging, workspace plan,
// bias broadcast, dBias reduction) is made behind the C ABI of include/ffpa_b200.h in libffpa_b200.so; this
// file converts tensors to pointers / strides, takes scratch from the torch caching allocator on the current
// stream and maps error codes to the reference's exception classes (TORCH_CHECK -> RuntimeError,
// std::invalid_argument -> ValueError, unsupported -> NotImplementedError).
// Extras beyond the reference surface (used by this repo's own package): ffpa_attn_backward_ex (bias /
// dropout replay, dBias, dLSE), ffpa_attn_varlen_forward / _backward, launch_count, abi_version.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDACachingAllocator.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime_api.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "ffpa_b200.h"

namespace {

using torch::Tensor;

[[noreturn]] void raise_code(int rc) {
  const std::string msg = std::string("ffpa_attn._C: ") + ffpa_b200_last_error();
  if (rc == FFPA_ERR_UNSUPPORTED) {
    TORCH_CHECK_NOT_IMPLEMENTED(false, msg);
  }
  TORCH_CHECK(false, msg);
}

int dtype_code(const Tensor& t, const char* what) {
  if (t.scalar_type() == at::kHalf) return FFPA_DTYPE_F16;
  if (t.scalar_type() == at::kBFloat16) return FFPA_DTYPE_BF16;
  // std::invalid_argument -> ValueError, as ffpa_api.cc:235-237
  throw std::invalid_argument(std::string("ffpa_attn: ") + what + ".dtype must be torch.float16 or torch.bfloat16");
}

void check_cuda_same(const Tensor& ref, std::initializer_list<const Tensor*> ts) {
  TORCH_CHECK(ref.is_cuda(), "ffpa_attn._C: all tensors must be CUDA tensors (there is no CPU path)");
  for (const Tensor* t : ts) {
    if (!t->defined() || t->numel() == 0) continue;
    TORCH_CHECK(t->is_cuda(), "ffpa_attn._C: all tensors must be CUDA tensors (there is no CPU path)");
    TORCH_CHECK(t->device() == ref.device(), "ffpa_attn._C: all tensors must live on the same device");
  }
}

void fill4(int64_t* dst, const Tensor& t) {
  for (int i = 0; i < 4; ++i) dst[i] = t.stride(i);
}
// packed [T, H, D] -> (b, h, n, d) element strides of the C ABI (batch stride unused)
void fill_thd(int64_t* dst, const Tensor& t) {
  dst[0] = 0; dst[1] = t.stride(1); dst[2] = t.stride(0); dst[3] = t.stride(2);
}

// attn bias contract of native/launch.cuh:277-290: 4-D [1|B, 1|Hq, 1|Nq, 1|Nkv], fp32 or Q's dtype, last dim
// contiguous, broadcast dims stride 0. Returns the tensor to keep alive (possibly a contiguous copy).
Tensor bias_fields(const Tensor& attn_bias, const Tensor& Q, const Tensor& K, const void** ptr, int32_t* kind, int64_t* stride) {
  *ptr = nullptr; *kind = FFPA_BIAS_NONE;
  for (int i = 0; i < 4; ++i) stride[i] = 0;
  if (!attn_bias.defined() || attn_bias.numel() == 0) return attn_bias;
  TORCH_CHECK(attn_bias.dim() == 4, "ffpa_attn: attn_bias must be 4-D [1|B, 1|Hq, 1|Nq, 1|Nkv]");
  const int64_t want[4] = {Q.size(0), Q.size(1), Q.size(2), K.size(2)};
  for (int i = 0; i < 4; ++i)
    TORCH_CHECK(attn_bias.size(i) == 1 || attn_bias.size(i) == want[i], "ffpa_attn: attn_bias dim ", i, " must be 1 or ", want[i]);
  if (attn_bias.scalar_type() == at::kFloat) *kind = FFPA_BIAS_F32;
  else if (attn_bias.scalar_type() == Q.scalar_type()) *kind = FFPA_BIAS_QDTYPE;
  else TORCH_CHECK(false, "ffpa_attn: attn_bias dtype must be float32 or match Q");
  Tensor b = attn_bias;
  if (b.size(3) != 1 && b.stride(3) != 1) b = b.contiguous();
  for (int i = 0; i < 4; ++i) stride[i] = b.size(i) == 1 ? 0 : b.stride(i);
  *ptr = b.data_ptr();
  return b;
}

Tensor alloc_bytes(uint64_t n, const Tensor& like) {
  return torch::empty({(int64_t)(n ? n : 16)}, torch::TensorOptions().dtype(torch::kUInt8).device(like.device()));
}

void fill_fwd_sizes(ffpa_fwd_params& p, const Tensor& Q, const Tensor& K) {
  p.batch = (int32_t)Q.size(0); p.heads_q = (int32_t)Q.size(1); p.seqlen_q = (int32_t)Q.size(2); p.head_dim = (int32_t)Q.size(3);
  p.heads_kv = (int32_t)K.size(1); p.seqlen_kv = (int32_t)K.size(2);
}

// Scratch policy: the stash path (5 GEMM passes instead of 8) trades O(Nq*Nkv) scratch for time. The scratch is
// BOUNDED: at most FFPA_BWD_STASH_MAX_GB (default 2.5 GiB -- measured on B200, profiles/r02_bwd_cap_sweep.md: a problem
// cut into (batch, KV-head) chunks that fit 2.25 GiB runs within 1-4 % of the unbounded stash, B=4 C2 960 vs 959
// TFLOP/s) and never more than half of (free device memory + the allocator's cached blocks). `min_workspace`
// (CUDABackend(bwd_min_workspace=True)) or FFPA_BWD_STASH=0 asks for the O(N) plan (three recompute kernels), and an
// allocation failure falls back to it.
// Only consulted when a call actually has optional O(Nq*Nkv) scratch to size (callers check that first): the driver
// query behind cudaMemGetInfo is slow and serialises with copies in flight -- calling it on every forward cost the
// chunk-pipelined host entry its copy / compute overlap.
uint64_t scratch_cap(const Tensor& like, const char* env_name) {
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 0; }
  uint64_t cached = 0;
  try {
    const auto st = c10::cuda::CUDACachingAllocator::getDeviceStats(like.device().index());
    const int64_t r = st.reserved_bytes[0].current, a = st.allocated_bytes[0].current;
    if (r > a) cached = (uint64_t)(r - a);
  } catch (...) {
  }
  uint64_t cap = ((uint64_t)free_b + cached) / 2;
  const char* g = std::getenv(env_name);
  const double gb = (g ? atof(g) : 2.5) * 1073741824.0;
  if (gb >= 0 && (double)cap > gb) cap = (uint64_t)gb;
  return cap;
}

// ---------------------------------------------------------------------------------------------
// ffpa_attn_forward -- signature of ffpa_api.cc:86-96; writes O and softmax_lse in place
// ---------------------------------------------------------------------------------------------
void ffpa_attn_forward(Tensor Q, Tensor K, Tensor V, Tensor attn_bias, Tensor O, Tensor softmax_lse, int64_t stages,
                       int64_t acc, int64_t causal, double softmax_scale, double dropout_p, int64_t philox_seed,
                       int64_t philox_offset, bool fp8_smooth_k, bool fp8_smooth_v, int64_t fp8_q_quant_method,
                       int64_t fp8_k_quant_method, int64_t fp8_v_quant_method, int64_t fp8_pv_acc_type,
                       int64_t fp8_qk_mm_type, bool fp8_hybrid, int64_t fp8_hybrid_n_early, bool fp4_hybrid,
                       int64_t fp4_hybrid_n_early) {
  (void)stages;  // the sm_100a pipeline depth is a compile-time property of each head-dim variant
  (void)fp4_hybrid; (void)fp4_hybrid_n_early;  // FP4 is refused by the library (backend hint 6)
  check_cuda_same(Q, {&K, &V, &O, &attn_bias, &softmax_lse});
  TORCH_CHECK(Q.dim() == 4 && K.dim() == 4 && V.dim() == 4, "ffpa_attn_forward: Q/K/V must be 4-D [B, H, N, D]");
  const int dt = dtype_code(Q, "Q");
  TORCH_CHECK(K.scalar_type() == Q.scalar_type() && V.scalar_type() == Q.scalar_type() && O.scalar_type() == Q.scalar_type(),
              "ffpa_attn_forward: Q/K/V/O must share one dtype");
  if (acc == 0) throw std::invalid_argument("ffpa_attn: fp16 MMA acc (acc=0) does not exist on sm_100a: tcgen05 accumulates in fp32 (TMEM)");
  if (acc != 1) throw std::invalid_argument("ffpa_attn: acc must be 0 (f16) or 1 (f32)");
  TORCH_CHECK(O.sizes() == Q.sizes(), "ffpa_attn_forward: O must have the shape of Q");
  TORCH_CHECK(K.sizes() == V.sizes() && K.size(0) == Q.size(0) && K.size(3) == Q.size(3),
              "ffpa_attn_forward: K/V must be [B, Hkv, Nkv, D] matching Q's B and D");
  if (Q.numel() == 0) return;                                   // B == 0 or Nq == 0: nothing to compute
  c10::cuda::OptionalCUDAGuard guard(Q.device());
  if (K.size(2) == 0) {                                         // no keys: O = 0, LSE = -inf (empty-row convention)
    O.zero_();
    if (softmax_lse.numel()) softmax_lse.fill_(-std::numeric_limits<float>::infinity());
    return;
  }
  // the reference kernels index dense row-major tensors (native/sm_80/split_d.cuh:137-142); strides are honoured
  // here through the tensor maps, only the head dim must have unit stride
  if (Q.stride(3) != 1) Q = Q.contiguous();
  if (K.stride(3) != 1) K = K.contiguous();
  if (V.stride(3) != 1) V = V.contiguous();
  TORCH_CHECK(O.stride(3) == 1, "ffpa_attn_forward: O must have unit stride on the head dim");

  ffpa_fwd_params p{};
  p.q = Q.data_ptr(); p.k = K.data_ptr(); p.v = V.data_ptr(); p.o = O.data_ptr();
  if (softmax_lse.numel() > 0) {
    TORCH_CHECK(softmax_lse.scalar_type() == at::kFloat && softmax_lse.dim() == 3 && softmax_lse.size(0) == Q.size(0) &&
                    softmax_lse.size(1) == Q.size(1) && softmax_lse.size(2) == Q.size(2) && softmax_lse.is_contiguous(),
                "ffpa_attn_forward: softmax_lse must be contiguous fp32 [B, Hq, Nq]");
    p.lse = softmax_lse.data_ptr<float>();
  }
  fill4(p.q_stride, Q); fill4(p.k_stride, K); fill4(p.v_stride, V); fill4(p.o_stride, O);
  Tensor bias_keep = bias_fields(attn_bias, Q, K, &p.bias, &p.bias_kind, p.bias_stride);
  fill_fwd_sizes(p, Q, K);
  p.dtype = dt;
  p.causal = causal != 0;
  p.impl = FFPA_IMPL_AUTO;   // = this thread's set_cuda_backend_impl() hint, as ffpa_api.cc:124 reads it
  p.softmax_scale = (float)softmax_scale;
  p.dropout_p = (float)dropout_p;
  p.philox_seed = (uint64_t)philox_seed;
  p.philox_offset = (uint64_t)philox_offset;
  p.fp8_smooth_k = fp8_smooth_k; p.fp8_smooth_v = fp8_smooth_v;
  p.fp8_q_quant_method = (int32_t)fp8_q_quant_method; p.fp8_k_quant_method = (int32_t)fp8_k_quant_method;
  p.fp8_v_quant_method = (int32_t)fp8_v_quant_method;
  p.fp8_pv_acc_type = (int32_t)fp8_pv_acc_type; p.fp8_qk_mm_type = (int32_t)fp8_qk_mm_type;
  p.fp8_hybrid = fp8_hybrid; p.fp8_hybrid_n_early = (int32_t)fp8_hybrid_n_early;

  Tensor ws;
  // optional O(Nq*Nkv) scratch (replay stash of head dims > 768) is bounded like the backward's: FFPA_FWD_REPLAY_MAX_GB
  // (default 2.5 GiB) and half of the free memory; chunked over heads when the whole problem does not fit
  const uint64_t required = ffpa_b200_fwd_workspace_bytes_p(&p, 0);
  uint64_t want = ffpa_b200_fwd_workspace_bytes_p(&p, ~0ull);
  if (want > required && want > (64ull << 20)) want = ffpa_b200_fwd_workspace_bytes_p(&p, scratch_cap(Q, "FFPA_FWD_REPLAY_MAX_GB"));
  if (want > 0) {
    try {
      ws = alloc_bytes(want, Q);
      p.workspace = ws.data_ptr(); p.workspace_bytes = want;
    } catch (const c10::OutOfMemoryError&) {
      // optional scratch (KV-split partials, replay stash): the library runs the scratch-free path; the FP8
      // path needs its buffers and reports the shortfall itself
      p.workspace = nullptr; p.workspace_bytes = 0;
    }
  }
  const int rc = ffpa_b200_fwd(&p, at::cuda::getCurrentCUDAStream(Q.device().index()).stream());
  if (rc != 0) raise_code(rc);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
void backward_impl(const Tensor& Q, const Tensor& K, const Tensor& V, const Tensor& O, const Tensor& softmax_lse, const Tensor& dO,
                   Tensor& dQ, Tensor& dK, Tensor& dV, int64_t causal, double softmax_scale, const Tensor& attn_bias,
                   double dropout_p, int64_t philox_seed, int64_t philox_offset, const Tensor& d_bias, const Tensor& d_lse,
                   bool min_workspace) {
  check_cuda_same(Q, {&K, &V, &O, &softmax_lse, &dO, &dQ, &dK, &dV, &attn_bias, &d_bias, &d_lse});
  TORCH_CHECK(Q.dim() == 4 && K.dim() == 4 && V.dim() == 4, "ffpa_attn_backward: Q/K/V must be 4-D [B, H, N, D]");
  const int dt = dtype_code(Q, "Q");
  if (Q.numel() == 0 || K.numel() == 0) { dQ.zero_(); dK.zero_(); dV.zero_(); return; }
  c10::cuda::OptionalCUDAGuard guard(Q.device());
  ffpa_bwd_params p{};
  p.q = Q.data_ptr(); p.k = K.data_ptr(); p.v = V.data_ptr(); p.o = O.data_ptr();
  TORCH_CHECK(softmax_lse.scalar_type() == at::kFloat && softmax_lse.is_contiguous() && softmax_lse.dim() == 3 &&
                  softmax_lse.size(0) == Q.size(0) && softmax_lse.size(1) == Q.size(1) && softmax_lse.size(2) == Q.size(2),
              "ffpa_attn_backward: softmax_lse must be contiguous fp32 [B, Hq, Nq]");
  p.lse = softmax_lse.data_ptr<float>();
  p.d_o = dO.data_ptr(); p.dq = dQ.data_ptr(); p.dk = dK.data_ptr(); p.dv = dV.data_ptr();
  for (const Tensor* t : std::initializer_list<const Tensor*>{&Q, &K, &V, &O, &dO, &dQ, &dK, &dV})
    TORCH_CHECK(t->dim() == 4 && t->stride(3) == 1, "ffpa_attn_backward: all tensors need unit stride on the head dim");
  fill4(p.q_stride, Q); fill4(p.k_stride, K); fill4(p.v_stride, V); fill4(p.o_stride, O); fill4(p.do_stride, dO);
  fill4(p.dq_stride, dQ); fill4(p.dk_stride, dK); fill4(p.dv_stride, dV);
  p.batch = (int32_t)Q.size(0); p.heads_q = (int32_t)Q.size(1); p.seqlen_q = (int32_t)Q.size(2); p.head_dim = (int32_t)Q.size(3);
  p.heads_kv = (int32_t)K.size(1); p.seqlen_kv = (int32_t)K.size(2);
  p.dtype = dt; p.causal = causal != 0; p.softmax_scale = (float)softmax_scale;
  Tensor bias_keep = bias_fields(attn_bias, Q, K, &p.bias, &p.bias_kind, p.bias_stride);
  p.dropout_p = (float)dropout_p; p.philox_seed = (uint64_t)philox_seed; p.philox_offset = (uint64_t)philox_offset;
  if (d_bias.defined() && d_bias.numel() > 0) {
    TORCH_CHECK(p.bias_kind != FFPA_BIAS_NONE, "ffpa_attn_backward: d_bias requested without attn_bias");
    TORCH_CHECK(d_bias.scalar_type() == at::kFloat && d_bias.is_contiguous() && d_bias.sizes() == bias_keep.sizes(),
                "ffpa_attn_backward: d_bias must be contiguous fp32 with the shape of attn_bias");
    p.d_bias = d_bias.data_ptr<float>();
    for (int i = 0; i < 4; ++i) p.d_bias_stride[i] = d_bias.size(i) == 1 ? 0 : d_bias.stride(i);
  }
  if (d_lse.defined() && d_lse.numel() > 0) {
    TORCH_CHECK(d_lse.scalar_type() == at::kFloat && d_lse.is_contiguous() && d_lse.sizes() == softmax_lse.sizes(),
                "ffpa_attn_backward: d_lse must be contiguous fp32 with the shape of softmax_lse");
    p.d_lse = d_lse.data_ptr<float>();
  }
  const uint64_t need_min = ffpa_b200_bwd_workspace_bytes_min_p(&p);
  uint64_t want = min_workspace ? need_min : ffpa_b200_bwd_workspace_bytes_p(&p, ~0ull);
  if (want > need_min && want > (64ull << 20)) want = ffpa_b200_bwd_workspace_bytes_p(&p, scratch_cap(Q, "FFPA_BWD_STASH_MAX_GB"));
  Tensor ws;
  try {
    ws = alloc_bytes(want, Q);
  } catch (const c10::OutOfMemoryError&) {
    if (want == need_min) throw;
    want = need_min;
    ws = alloc_bytes(want, Q);
  }
  p.workspace = ws.data_ptr(); p.workspace_bytes = want;
  const int rc = ffpa_b200_bwd(&p, at::cuda::getCurrentCUDAStream(Q.device().index()).stream());
  if (rc != 0) raise_code(rc);
}

// signature of ffpa_api.cc:242-246
void ffpa_attn_backward(Tensor Q, Tensor K, Tensor V, Tensor O, Tensor softmax_lse, Tensor dO, Tensor dQ, Tensor dK, Tensor dV,
                        int64_t stages, int64_t causal, double softmax_scale) {
  (void)stages;
  backward_impl(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, causal, softmax_scale, Tensor(), 0.0, 0, 0, Tensor(), Tensor(), false);
}

void ffpa_attn_backward_ex(Tensor Q, Tensor K, Tensor V, Tensor O, Tensor softmax_lse, Tensor dO, Tensor dQ, Tensor dK, Tensor dV,
                           int64_t stages, int64_t causal, double softmax_scale, c10::optional<Tensor> attn_bias,
                           double dropout_p, int64_t philox_seed, int64_t philox_offset, c10::optional<Tensor> d_bias,
                           c10::optional<Tensor> d_lse, bool min_workspace) {
  (void)stages;
  backward_impl(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, causal, softmax_scale, attn_bias.value_or(Tensor()), dropout_p,
                philox_seed, philox_offset, d_bias.value_or(Tensor()), d_lse.value_or(Tensor()), min_workspace);
}

// ---------------------------------------------------------------------------------------------
// packed variable-length entries (C ABI: cu_seqlens_* set); backend of ffpa_attn_varlen_func
// (/root/reference/src/ffpa_attn/ffpa_attn_interface.py:192-279)
// ---------------------------------------------------------------------------------------------
void check_varlen(const Tensor& Q, const Tensor& K, const Tensor& V, const Tensor& cu_q, const Tensor& cu_k) {
  TORCH_CHECK(Q.dim() == 3 && K.dim() == 3 && V.dim() == 3, "ffpa_attn varlen: q/k/v must be packed [T, H, D] tensors");
  TORCH_CHECK(K.sizes() == V.sizes() && K.size(2) == Q.size(2), "ffpa_attn varlen: k and v must share [T_k, H_kv, D] and q's head dim");
  for (const Tensor* cu : std::initializer_list<const Tensor*>{&cu_q, &cu_k})
    TORCH_CHECK(cu->scalar_type() == at::kInt && cu->dim() == 1 && cu->is_contiguous() && cu->device() == Q.device(),
                "ffpa_attn varlen: cu_seqlens must be contiguous int32 1-D tensors on q's device");
  TORCH_CHECK(cu_q.numel() == cu_k.numel() && cu_q.numel() >= 2, "ffpa_attn varlen: cu_seqlens_q / cu_seqlens_k must both have B + 1 entries");
}

void ffpa_attn_varlen_forward(Tensor Q, Tensor K, Tensor V, Tensor O, Tensor softmax_lse, Tensor cu_q, Tensor cu_k,
                              int64_t max_q, int64_t max_k, int64_t causal, double softmax_scale) {
  check_cuda_same(Q, {&K, &V, &O, &softmax_lse, &cu_q, &cu_k});
  check_varlen(Q, K, V, cu_q, cu_k);
  const int dt = dtype_code(Q, "q");
  TORCH_CHECK(K.scalar_type() == Q.scalar_type() && V.scalar_type() == Q.scalar_type() && O.scalar_type() == Q.scalar_type() &&
                  O.sizes() == Q.sizes(), "ffpa_attn varlen: Q/K/V/O must share one dtype and O the shape of Q");
  for (const Tensor* t : std::initializer_list<const Tensor*>{&Q, &K, &V, &O}) TORCH_CHECK(t->stride(2) == 1, "ffpa_attn varlen: unit stride on the head dim is required");
  c10::cuda::OptionalCUDAGuard guard(Q.device());
  ffpa_fwd_params p{};
  p.q = Q.data_ptr(); p.k = K.data_ptr(); p.v = V.data_ptr(); p.o = O.data_ptr();
  if (softmax_lse.numel() > 0) {
    TORCH_CHECK(softmax_lse.scalar_type() == at::kFloat && softmax_lse.is_contiguous() && softmax_lse.dim() == 2 &&
                    softmax_lse.size(0) == Q.size(1) && softmax_lse.size(1) == Q.size(0),
                "ffpa_attn varlen: softmax_lse must be contiguous fp32 [Hq, T_q]");
    p.lse = softmax_lse.data_ptr<float>();
  }
  fill_thd(p.q_stride, Q); fill_thd(p.k_stride, K); fill_thd(p.v_stride, V); fill_thd(p.o_stride, O);
  p.batch = (int32_t)cu_q.numel() - 1; p.heads_q = (int32_t)Q.size(1); p.heads_kv = (int32_t)K.size(1); p.head_dim = (int32_t)Q.size(2);
  p.seqlen_q = (int32_t)max_q; p.seqlen_kv = (int32_t)max_k;
  p.total_q = (int32_t)Q.size(0); p.total_k = (int32_t)K.size(0);
  p.cu_seqlens_q = cu_q.data_ptr<int32_t>(); p.cu_seqlens_k = cu_k.data_ptr<int32_t>();
  p.dtype = dt; p.causal = causal != 0; p.softmax_scale = (float)softmax_scale;
  p.impl = FFPA_IMPL_NATIVE;
  const int rc = ffpa_b200_fwd(&p, at::cuda::getCurrentCUDAStream(Q.device().index()).stream());
  if (rc != 0) raise_code(rc);
}

void ffpa_attn_varlen_backward(Tensor Q, Tensor K, Tensor V, Tensor O, Tensor softmax_lse, Tensor dO, Tensor dQ, Tensor dK, Tensor dV,
                               Tensor cu_q, Tensor cu_k, int64_t max_q, int64_t max_k, int64_t causal, double softmax_scale,
                               c10::optional<Tensor> d_lse_opt) {
  check_cuda_same(Q, {&K, &V, &O, &softmax_lse, &dO, &dQ, &dK, &dV, &cu_q, &cu_k});
  check_varlen(Q, K, V, cu_q, cu_k);
  const int dt = dtype_code(Q, "q");
  c10::cuda::OptionalCUDAGuard guard(Q.device());
  ffpa_bwd_params p{};
  p.q = Q.data_ptr(); p.k = K.data_ptr(); p.v = V.data_ptr(); p.o = O.data_ptr();
  TORCH_CHECK(softmax_lse.scalar_type() == at::kFloat && softmax_lse.is_contiguous(), "ffpa_attn varlen backward: softmax_lse must be contiguous fp32 [Hq, T_q]");
  p.lse = softmax_lse.data_ptr<float>();
  p.d_o = dO.data_ptr(); p.dq = dQ.data_ptr(); p.dk = dK.data_ptr(); p.dv = dV.data_ptr();
  for (const Tensor* t : std::initializer_list<const Tensor*>{&Q, &K, &V, &O, &dO, &dQ, &dK, &dV})
    TORCH_CHECK(t->dim() == 3 && t->stride(2) == 1, "ffpa_attn varlen backward: all tensors need unit stride on the head dim");
  fill_thd(p.q_stride, Q); fill_thd(p.k_stride, K); fill_thd(p.v_stride, V); fill_thd(p.o_stride, O); fill_thd(p.do_stride, dO);
  fill_thd(p.dq_stride, dQ); fill_thd(p.dk_stride, dK); fill_thd(p.dv_stride, dV);
  p.batch = (int32_t)cu_q.numel() - 1; p.heads_q = (int32_t)Q.size(1); p.heads_kv = (int32_t)K.size(1); p.head_dim = (int32_t)Q.size(2);
  p.seqlen_q = (int32_t)max_q; p.seqlen_kv = (int32_t)max_k;
  p.total_q = (int32_t)Q.size(0); p.total_k = (int32_t)K.size(0);
  p.cu_seqlens_q = cu_q.data_ptr<int32_t>(); p.cu_seqlens_k = cu_k.data_ptr<int32_t>();
  p.dtype = dt; p.causal = causal != 0; p.softmax_scale = (float)softmax_scale;
  const Tensor d_lse = d_lse_opt.value_or(Tensor());
  if (d_lse.defined() && d_lse.numel() > 0) {
    TORCH_CHECK(d_lse.scalar_type() == at::kFloat && d_lse.is_contiguous() && d_lse.sizes() == softmax_lse.sizes(),
                "ffpa_attn varlen backward: d_lse must be contiguous fp32 [Hq, T_q]");
    p.d_lse = d_lse.data_ptr<float>();
  }
  const uint64_t need = ffpa_b200_bwd_workspace_bytes_min_p(&p);
  Tensor ws = alloc_bytes(need, Q);
  p.workspace = ws.data_ptr(); p.workspace_bytes = need;
  const int rc = ffpa_b200_bwd(&p, at::cuda::getCurrentCUDAStream(Q.device().index()).stream());
  if (rc != 0) raise_code(rc);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("ffpa_attn_forward", &ffpa_attn_forward,
        "FFPA unified prefill attention dispatch (fp16/bf16, causal, softmax_scale, outputs softmax_lse [B, Nh_q, Nq] "
        "float32) on the sm_100a tcgen05 kernels");
  m.def("ffpa_attn_backward", &ffpa_attn_backward, "FFPA native CUDA backward (dQ, dK, dV written in place)");
  m.def(
      "set_cuda_backend_impl",
      [](int impl) {
        const int rc = ffpa_b200_set_backend_impl(impl);
        if (rc != 0) raise_code(rc);
      },
      "Set CUDA backend implementation hint (0=AUTO, 1=NATIVE, 2=TMA, 3=CUTE, 4=CUTE_TMA, 5=CUTE_TMA_FP8, 6=CUTE_TMA_FP4)");
  m.def("get_cuda_backend_impl", []() { return (int)ffpa_b200_get_backend_impl(); },
        "Get current CUDA backend implementation hint");
  m.attr("CUDA_FWD_AVAILABLE") = py::bool_(ffpa_b200_fwd_available() != 0);
  m.attr("CUDA_AVAILABLE") = py::bool_(ffpa_b200_fwd_available() != 0);
  m.attr("F16_ACC_AVAILABLE") = py::bool_(false);        // tcgen05 accumulates in fp32 (TMEM)
  m.attr("CUDA_TMA_AVAILABLE") = py::bool_(true);        // every kernel here is TMA-fed
  m.attr("CUDA_CUTE_TMA_AVAILABLE") = py::bool_(false);  // no CuTe / CUTLASS code in this build
  m.attr("CUDA_BWD_AVAILABLE") = py::bool_(ffpa_b200_bwd_available() != 0);
  // ---- extras (not part of the reference surface) ----
  m.def("ffpa_attn_backward_ex", &ffpa_attn_backward_ex, py::arg("Q"), py::arg("K"), py::arg("V"), py::arg("O"),
        py::arg("softmax_lse"), py::arg("dO"), py::arg("dQ"), py::arg("dK"), py::arg("dV"), py::arg("stages"),
        py::arg("causal"), py::arg("softmax_scale"), py::arg("attn_bias") = py::none(), py::arg("dropout_p") = 0.0,
        py::arg("philox_seed") = 0, py::arg("philox_offset") = 0, py::arg("d_bias") = py::none(),
        py::arg("d_lse") = py::none(), py::arg("min_workspace") = false,
        "backward with bias / dropout replay, bias gradient (fp32, bias-shaped, reduced in-kernel) and dLSE");
  m.def("ffpa_attn_varlen_forward", &ffpa_attn_varlen_forward, "packed [T, H, D] forward, one launch for the batch");
  m.def("ffpa_attn_varlen_backward", &ffpa_attn_varlen_backward, py::arg("Q"), py::arg("K"), py::arg("V"), py::arg("O"),
        py::arg("softmax_lse"), py::arg("dO"), py::arg("dQ"), py::arg("dK"), py::arg("dV"), py::arg("cu_seqlens_q"),
        py::arg("cu_seqlens_k"), py::arg("max_seqlen_q"), py::arg("max_seqlen_k"), py::arg("causal"),
        py::arg("softmax_scale"), py::arg("d_lse") = py::none(), "packed [T, H, D] backward");
  m.def("launch_count", []() { return (uint64_t)ffpa_b200_launch_count(); }, "kernels launched by libffpa_b200.so since load");
  m.def("refresh_env", []() { ffpa_b200_refresh_env(); }, "re-read the FFPA_* tuning variables (cached per process)");
  m.attr("ABI_VERSION") = py::int_(ffpa_b200_abi_version());
}
