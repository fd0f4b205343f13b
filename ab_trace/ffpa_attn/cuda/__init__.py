"""CUDA backend shim: registers ``torch.ops.ffpa_attn._fwd_cuda`` (and ``_bwd_cuda``) on top of
the native binding ``ffpa_attn._C`` (a real PyTorch C++ extension, csrc/ffpa_torch_binding.cpp, with the
reference module's pybind surface), with the reference's op schema so ``torch.compile`` sees the same op
(/root/reference/src/ffpa_attn/cuda/__init__.py:28-35, 57-171; wrappers _ffpa_fwd.py:6-62,
_ffpa_bwd.py:6-26).  The only backend behind the op is the sm_100a kernel family; a missing extension is an
ImportError, never a fallback."""
from __future__ import annotations

import enum

import torch

try:
  from .. import _C as _cuda_ext
except ImportError as exc:  # no CPU / Triton / SDPA route exists: fail loudly
  raise ImportError(
    "ffpa_attn._C (the sm_100a PyTorch extension) is not built. Run `python -c 'import __graft_entry__ as g; "
    f"g.build()'` (or `make -C ffpa-attn_b200/csrc`). There is no fallback backend. Original error: {exc}") from exc

CUDA_FWD_AVAILABLE = _cuda_ext.CUDA_FWD_AVAILABLE
CUDA_AVAILABLE = _cuda_ext.CUDA_AVAILABLE
CUDA_BWD_AVAILABLE = _cuda_ext.CUDA_BWD_AVAILABLE
F16_ACC_AVAILABLE = _cuda_ext.F16_ACC_AVAILABLE
CUDA_TMA_AVAILABLE = _cuda_ext.CUDA_TMA_AVAILABLE
CUDA_CUTE_TMA_AVAILABLE = _cuda_ext.CUDA_CUTE_TMA_AVAILABLE


class CudaBackendImpl(enum.IntEnum):
  """Mirror of ffpa::CudaBackendImpl (/root/reference/csrc/cuffpa/backend.h:6-14)."""
  AUTO = 0
  NATIVE = 1
  TMA = 2
  CUTE = 3
  CUTE_TMA = 4
  CUTE_TMA_FP8 = 5
  CUTE_TMA_FP4 = 6


def set_cuda_backend_impl(impl: CudaBackendImpl) -> None:
  _cuda_ext.set_cuda_backend_impl(int(impl))


def get_cuda_backend_impl() -> CudaBackendImpl:
  return CudaBackendImpl(_cuda_ext.get_cuda_backend_impl())


def launch_count() -> int:
  """Kernels launched by libffpa_b200.so since load (bench.py reports it as ``gpu_launches``)."""
  return int(_cuda_ext.launch_count())


_OP_NAMESPACE = "ffpa_attn"

torch.library.define(
  f"{_OP_NAMESPACE}::_fwd_cuda",
  "(Tensor q, Tensor k, Tensor v, Tensor attn_bias, int stages, int acc, int causal, "
  "float softmax_scale, float dropout_p, int philox_seed, int philox_offset, "
  "bool fp8_smooth_k, bool fp8_smooth_v, int fp8_q_quant_method, int fp8_k_quant_method, "
  "int fp8_v_quant_method, int fp8_pv_acc_type, int fp8_qk_mm_type, "
  "bool fp8_hybrid, int fp8_hybrid_n_early, "
  "bool fp4_hybrid, int fp4_hybrid_n_early) -> "
  "(Tensor o, Tensor softmax_lse)",
)


@torch.library.impl(f"{_OP_NAMESPACE}::_fwd_cuda", "CUDA")
def _fwd_cuda_torch_op(Q, K, V, attn_bias, stages, acc, causal, softmax_scale, dropout_p,
                       philox_seed, philox_offset, fp8_smooth_k, fp8_smooth_v, fp8_q_quant_method,
                       fp8_k_quant_method, fp8_v_quant_method, fp8_pv_acc_type, fp8_qk_mm_type,
                       fp8_hybrid, fp8_hybrid_n_early, fp4_hybrid, fp4_hybrid_n_early):
  O = torch.empty_like(Q, memory_format=torch.contiguous_format)  # noqa: E741
  # exact Nq, not padded (/root/reference/src/ffpa_attn/cuda/__init__.py:100-110)
  softmax_lse = torch.empty(Q.size(0), Q.size(1), Q.size(2), dtype=torch.float32, device=Q.device)
  _cuda_ext.ffpa_attn_forward(
    Q, K, V, attn_bias, O, softmax_lse, stages, acc, causal, softmax_scale, dropout_p, philox_seed,
    philox_offset, fp8_smooth_k, fp8_smooth_v, fp8_q_quant_method, fp8_k_quant_method,
    fp8_v_quant_method, fp8_pv_acc_type, fp8_qk_mm_type, fp8_hybrid, fp8_hybrid_n_early,
    fp4_hybrid, fp4_hybrid_n_early)
  return O, softmax_lse


@torch.library.register_fake(f"{_OP_NAMESPACE}::_fwd_cuda")
def _fwd_cuda_fake(Q, K, V, attn_bias, stages, acc, causal, softmax_scale, dropout_p, philox_seed,
                   philox_offset, fp8_smooth_k, fp8_smooth_v, fp8_q_quant_method,
                   fp8_k_quant_method, fp8_v_quant_method, fp8_pv_acc_type, fp8_qk_mm_type,
                   fp8_hybrid, fp8_hybrid_n_early, fp4_hybrid, fp4_hybrid_n_early):
  O = torch.empty_like(Q, memory_format=torch.contiguous_format)  # noqa: E741
  softmax_lse = Q.new_empty(Q.size(0), Q.size(1), Q.size(2), dtype=torch.float32)
  return O, softmax_lse


# ffpa_attn::_bwd_cuda -- the reference has no CUDA-C++ backward (cuda/_ffpa_bwd.py:6-26 raises);
# the symbol ffpa_attn_backward keeps its reference signature and is real here.
torch.library.define(
  f"{_OP_NAMESPACE}::_bwd_cuda",
  "(Tensor q, Tensor k, Tensor v, Tensor o, Tensor softmax_lse, Tensor d_o, int stages, "
  "int causal, float softmax_scale) -> (Tensor dq, Tensor dk, Tensor dv)",
)


class _Flag:
  """Per-thread switch read by the backward ops (set by functional._FFPAAttnFunc.backward from
  ``CUDABackend.bwd_min_workspace``): True = O(N)-memory recompute kernels, no score stash."""

  def __init__(self):
    import threading
    self._tls = threading.local()

  def get(self) -> bool:
    return bool(getattr(self._tls, "v", False))

  def set(self, v: bool) -> None:
    self._tls.v = bool(v)


_BWD_MIN_WORKSPACE = _Flag()


@torch.library.impl(f"{_OP_NAMESPACE}::_bwd_cuda", "CUDA")
def _bwd_cuda_torch_op(Q, K, V, O, softmax_lse, dO, stages, causal, softmax_scale):
  dQ = torch.empty_like(Q, memory_format=torch.contiguous_format)
  dK = torch.empty_like(K, memory_format=torch.contiguous_format)
  dV = torch.empty_like(V, memory_format=torch.contiguous_format)
  if _BWD_MIN_WORKSPACE.get():
    _cuda_ext.ffpa_attn_backward_ex(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, stages, causal, softmax_scale,
                                    min_workspace=True)
  else:
    _cuda_ext.ffpa_attn_backward(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, stages, causal, softmax_scale)
  return dQ, dK, dV


@torch.library.register_fake(f"{_OP_NAMESPACE}::_bwd_cuda")
def _bwd_cuda_fake(Q, K, V, O, softmax_lse, dO, stages, causal, softmax_scale):
  return (torch.empty_like(Q, memory_format=torch.contiguous_format),
          torch.empty_like(K, memory_format=torch.contiguous_format),
          torch.empty_like(V, memory_format=torch.contiguous_format))


# extended backward: replays bias / dropout and optionally returns the bias gradient (reduced over the
# bias' broadcast dims, cast to its dtype) -- the math of the reference's Triton backward
# (/root/reference/src/ffpa_attn/triton/_ffpa_bwd.py:692-855) on the sm_100a kernels.
torch.library.define(
  f"{_OP_NAMESPACE}::_bwd_cuda_ex",
  "(Tensor q, Tensor k, Tensor v, Tensor o, Tensor softmax_lse, Tensor d_o, Tensor attn_bias, int stages, "
  "int causal, float softmax_scale, float dropout_p, int philox_seed, int philox_offset, bool bias_grad) -> "
  "(Tensor dq, Tensor dk, Tensor dv, Tensor dbias)",
)


@torch.library.impl(f"{_OP_NAMESPACE}::_bwd_cuda_ex", "CUDA")
def _bwd_cuda_ex_torch_op(Q, K, V, O, softmax_lse, dO, attn_bias, stages, causal, softmax_scale, dropout_p,
                          philox_seed, philox_offset, bias_grad):
  dQ = torch.empty_like(Q, memory_format=torch.contiguous_format)
  dK = torch.empty_like(K, memory_format=torch.contiguous_format)
  dV = torch.empty_like(V, memory_format=torch.contiguous_format)
  has_bias = attn_bias.numel() > 0
  d32 = None
  if bias_grad and has_bias:
    # bias-shaped fp32 accumulator: the dQ kernel reduces dS over the bias' broadcast dims itself (never a
    # [B, Hq, Nq, Nkv] buffer for a [B, 1, 1, Nkv] bias); the library zero-fills it
    d32 = torch.empty(attn_bias.shape, dtype=torch.float32, device=Q.device)
  _cuda_ext.ffpa_attn_backward_ex(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, stages, causal, softmax_scale,
                                  attn_bias=attn_bias if has_bias else None, dropout_p=dropout_p,
                                  philox_seed=philox_seed, philox_offset=philox_offset, d_bias=d32,
                                  min_workspace=_BWD_MIN_WORKSPACE.get())
  dbias = d32.to(attn_bias.dtype) if d32 is not None else Q.new_empty(0)
  return dQ, dK, dV, dbias


@torch.library.register_fake(f"{_OP_NAMESPACE}::_bwd_cuda_ex")
def _bwd_cuda_ex_fake(Q, K, V, O, softmax_lse, dO, attn_bias, stages, causal, softmax_scale, dropout_p,
                      philox_seed, philox_offset, bias_grad):
  dbias = torch.empty_like(attn_bias) if (bias_grad and attn_bias.numel() > 0) else Q.new_empty(0)
  return (torch.empty_like(Q, memory_format=torch.contiguous_format),
          torch.empty_like(K, memory_format=torch.contiguous_format),
          torch.empty_like(V, memory_format=torch.contiguous_format), dbias)


# packed variable-length ops (reference: ffpa_attn::_varlen_fwd_cute / _varlen_bwd_cute registered at
# /root/reference/src/ffpa_attn/cute/__init__.py:466-571; here one sm_100a launch set for the whole batch,
# cu_seqlens never leave the device)
torch.library.define(
  f"{_OP_NAMESPACE}::_varlen_fwd_cuda",
  "(Tensor q, Tensor k, Tensor v, Tensor cu_seqlens_q, Tensor cu_seqlens_k, int max_seqlen_q, int max_seqlen_k, "
  "int causal, float softmax_scale) -> (Tensor o, Tensor softmax_lse)",
)


@torch.library.impl(f"{_OP_NAMESPACE}::_varlen_fwd_cuda", "CUDA")
def _varlen_fwd_cuda_torch_op(Q, K, V, cu_q, cu_k, max_q, max_k, causal, softmax_scale):
  O = torch.empty_like(Q, memory_format=torch.contiguous_format)  # noqa: E741
  softmax_lse = torch.empty(Q.size(1), Q.size(0), dtype=torch.float32, device=Q.device)
  _cuda_ext.ffpa_attn_varlen_forward(Q, K, V, O, softmax_lse, cu_q, cu_k, max_q, max_k, causal, softmax_scale)
  return O, softmax_lse


@torch.library.register_fake(f"{_OP_NAMESPACE}::_varlen_fwd_cuda")
def _varlen_fwd_cuda_fake(Q, K, V, cu_q, cu_k, max_q, max_k, causal, softmax_scale):
  return (torch.empty_like(Q, memory_format=torch.contiguous_format),
          Q.new_empty(Q.size(1), Q.size(0), dtype=torch.float32))


torch.library.define(
  f"{_OP_NAMESPACE}::_varlen_bwd_cuda",
  "(Tensor q, Tensor k, Tensor v, Tensor o, Tensor softmax_lse, Tensor d_o, Tensor cu_seqlens_q, Tensor cu_seqlens_k, "
  "int max_seqlen_q, int max_seqlen_k, int causal, float softmax_scale, Tensor? d_lse=None) -> (Tensor dq, Tensor dk, Tensor dv)",
)


@torch.library.impl(f"{_OP_NAMESPACE}::_varlen_bwd_cuda", "CUDA")
def _varlen_bwd_cuda_torch_op(Q, K, V, O, softmax_lse, dO, cu_q, cu_k, max_q, max_k, causal, softmax_scale, d_lse=None):
  # tokens outside every sequence (none when cu_seqlens spans the tensors) keep a zero gradient
  dQ = torch.zeros_like(Q, memory_format=torch.contiguous_format)
  dK = torch.zeros_like(K, memory_format=torch.contiguous_format)
  dV = torch.zeros_like(V, memory_format=torch.contiguous_format)
  _cuda_ext.ffpa_attn_varlen_backward(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, cu_q, cu_k, max_q, max_k,
                                      causal, softmax_scale, d_lse=d_lse)
  return dQ, dK, dV


@torch.library.register_fake(f"{_OP_NAMESPACE}::_varlen_bwd_cuda")
def _varlen_bwd_cuda_fake(Q, K, V, O, softmax_lse, dO, cu_q, cu_k, max_q, max_k, causal, softmax_scale, d_lse=None):
  return (torch.empty_like(Q, memory_format=torch.contiguous_format),
          torch.empty_like(K, memory_format=torch.contiguous_format),
          torch.empty_like(V, memory_format=torch.contiguous_format))


def _ffpa_attn_forward_cuda(Q, K, V, O, attn_bias, stages, acc, causal, softmax_scale,
                            dropout_p=0.0, philox_seed=0, philox_offset=0, fp8_smooth_k=True,
                            fp8_smooth_v=False, fp8_q_quant_method=0, fp8_k_quant_method=0,
                            fp8_v_quant_method=0, fp8_pv_acc_type=1, fp8_qk_mm_type=0,
                            fp8_hybrid=False, fp8_hybrid_n_early=256, fp4_hybrid=False,
                            fp4_hybrid_n_early=256):
  """Python wrapper with the reference's argument order (cuda/_ffpa_fwd.py:6-62).
  ``O`` is accepted for signature compatibility; the op allocates its own output."""
  del O
  if attn_bias is None:
    attn_bias = Q.new_empty(0)
  return torch.ops.ffpa_attn._fwd_cuda(
    Q, K, V, attn_bias, int(stages) if stages is not None else 0, int(acc), int(causal),
    float(softmax_scale), float(dropout_p), int(philox_seed), int(philox_offset),
    bool(fp8_smooth_k), bool(fp8_smooth_v), int(fp8_q_quant_method), int(fp8_k_quant_method),
    int(fp8_v_quant_method), int(fp8_pv_acc_type), int(fp8_qk_mm_type), bool(fp8_hybrid),
    int(fp8_hybrid_n_early), bool(fp4_hybrid), int(fp4_hybrid_n_early))


def _ffpa_attn_backward_cuda(Q, K, V, O, softmax_lse, dO, stages, causal, softmax_scale):
  return torch.ops.ffpa_attn._bwd_cuda(Q, K, V, O, softmax_lse, dO,
                                       int(stages) if stages is not None else 0, int(causal),
                                       float(softmax_scale))
