"""Host-side semantics of ``ffpa_attn_func``: backend config object, input validation and
normalisation, dropout RNG reservation and the autograd Function.

Mirrors the behaviour (names, argument meaning, error classes) of
/root/reference/src/ffpa_attn/functional.py for the CUDA-backend path:
  CUDABackend                      :217-373
  FFPAAttnMeta.from_kwargs         :610-652
  FFPAAttnMeta.normalize_inputs    :726-849
  FFPAAttnMeta.normalize_attn_mask :851-911
  _reserve_large_d_dropout_rng     :518-540
  _FFPAAttnFunc.forward / backward :964-1079 / :1081-1172
with one deliberate difference: there is a single backend (the sm_100a kernels in
libffpa_b200.so).  No Triton, no CuTe-DSL, no SDPA fallback, no CPU path -- every shape the
reference would hand to aten SDPA (D <= 256, 8 <= Nq < 512, Nkv < 512; functional.py:676-724) is
served by the same kernel here, and CPU tensors raise.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

from .cuda import (
  _BWD_MIN_WORKSPACE,
  CUDA_BWD_AVAILABLE,
  CudaBackendImpl,
  _ffpa_attn_backward_cuda,
  _ffpa_attn_forward_cuda,
  set_cuda_backend_impl,
)

_ACC_F16, _ACC_F32 = 0, 1
# codes of the reference's op signature (/root/reference/src/ffpa_attn/functional.py:46-67)
_QUANT_CODE = {"per_block": 0, "per_channel": 1, "per_thread": 2}
_PV_ACC_CODE = {"f16": 0, "f32": 1}
_QK_MM_TYPE_CODE = {"fp8": 0, "int8": 1}


@dataclass
class AttentionMeta:
  is_causal: bool = False
  dropout_p: float = 0.0
  scale: float | None = None
  is_grad_enabled: bool = True


@dataclass
class CUDABackend:
  """Configuration object of the CUDA backend (reference: functional.py:217-264).

  All reference fields are accepted so existing call sites keep working; on B200 they select
  between two kernels only: the fp16/bf16 tcgen05 kernel (default) and the FP8 kernel
  (``enable_fp8=True``).  ``stages``/``enable_tma``/``enable_cute``/``enable_ws`` describe properties every
  sm_100a kernel already has (TMA-fed, warp-specialised, static pipeline depth per head dim) and change
  nothing.  Knobs that would change numerics but select sm_120 ``mma.sync`` variants this build does not have
  (``fp8_q/k_quant_method="per_thread"``, ``fp8_qk_mm_type="int8"``, ``fp8_pv_acc_type="f16"``) raise
  ``NotImplementedError`` naming the knob -- they are never silently dropped.
  ``bwd_min_workspace`` (B200 extra): True forces the O(N)-memory backward (three recompute kernels);
  by default the backward may use an O(Nq*Nkv) score stash taken from at most half of the memory that is
  free anyway (see csrc/ffpa_torch_binding.cpp), falling back to the O(N) plan when allocation fails.
  """
  name: str = "cuda"
  acc: str = "f32"
  stages: int | None = None
  enable_tma: bool | None = None
  enable_cute: bool | None = None
  enable_ws: bool = False
  enable_fp8: bool = False
  enable_fp4: bool = False
  fp8_smooth_k: bool = True
  fp8_smooth_v: bool = False
  fp8_q_quant_method: str = "per_block"
  fp8_k_quant_method: str = "per_block"
  fp8_v_quant_method: str = "per_block"
  fp8_pv_acc_type: str = "f32"
  fp8_qk_mm_type: str = "fp8"
  fp8_hybrid: bool | None = None
  fp8_hybrid_n_early: int = 256
  fp4_hybrid: bool | None = None
  fp4_hybrid_n_early: int = 256
  is_causal: bool = False
  forward: bool = True
  backward: bool = True  # unlike the reference (functional.py:266-268) a CUDA backward exists
  bwd_min_workspace: bool = False

  def __post_init__(self) -> None:
    if self.name != "cuda":
      raise ValueError(f"CUDABackend.name must be 'cuda', got {self.name!r}")
    if self.acc not in ("f16", "f32"):
      raise ValueError(f"acc must be 'f16' or 'f32', got {self.acc!r}")
    if self.acc == "f16":
      # reference: ValueError when the f16-acc kernels were not compiled (functional.py:274-278)
      raise ValueError("CUDABackend(acc='f16') is unavailable: tcgen05 accumulates in fp32 (TMEM)")
    if self.enable_fp8 and self.enable_fp4:
      raise ValueError("enable_fp8 and enable_fp4 are mutually exclusive")
    if self.enable_fp4:
      raise NotImplementedError("the NVFP4 path (sm_120 block-scaled mma) is out of scope on sm_100a")
    if self.fp8_q_quant_method not in ("per_block", "per_thread"):
      raise ValueError(f"fp8_q_quant_method must be 'per_block' or 'per_thread', got {self.fp8_q_quant_method!r}")
    if self.fp8_k_quant_method not in ("per_block", "per_thread"):
      raise ValueError(f"fp8_k_quant_method must be 'per_block' or 'per_thread', got {self.fp8_k_quant_method!r}")
    if self.fp8_v_quant_method not in ("per_block", "per_channel"):
      raise ValueError(f"fp8_v_quant_method must be 'per_block' or 'per_channel', got {self.fp8_v_quant_method!r}")
    if self.fp8_pv_acc_type not in _PV_ACC_CODE:
      raise ValueError(f"fp8_pv_acc_type must be 'f32' or 'f16', got {self.fp8_pv_acc_type!r}")
    if self.fp8_qk_mm_type not in _QK_MM_TYPE_CODE:
      raise ValueError(f"fp8_qk_mm_type must be 'fp8' or 'int8', got {self.fp8_qk_mm_type!r}")
    if self.fp8_smooth_v and self.fp8_v_quant_method != "per_channel":
      # reference: functional.py:300-302
      raise ValueError("fp8_smooth_v requires fp8_v_quant_method='per_channel'")
    if self.enable_fp8:
      # result-changing knobs without an sm_100a implementation are refused, not ignored (the native layer
      # refuses them too for callers that bypass this class)
      if "per_thread" in (self.fp8_q_quant_method, self.fp8_k_quant_method):
        raise NotImplementedError(
          "CUDABackend: fp8_q_quant_method / fp8_k_quant_method='per_thread' is not implemented on sm_100a "
          "(per-thread scales follow the mma.sync fragment layout); use 'per_block'")
      if self.fp8_qk_mm_type == "int8":
        raise NotImplementedError("CUDABackend: fp8_qk_mm_type='int8' is not implemented on sm_100a; use 'fp8'")
      if self.fp8_pv_acc_type == "f16":
        raise NotImplementedError(
          "CUDABackend: fp8_pv_acc_type='f16' is not implemented on sm_100a (TMEM accumulators are fp32); use 'f32'")

  @property
  def acc_code(self) -> int:
    return _ACC_F16 if self.acc == "f16" else _ACC_F32

  @property
  def impl_hint(self) -> CudaBackendImpl:
    """Backend hint handed to the native layer (reference: functional.py:138-155)."""
    if self.enable_fp8:
      return CudaBackendImpl.CUTE_TMA_FP8
    if self.enable_cute and self.enable_tma:
      return CudaBackendImpl.CUTE_TMA
    if self.enable_cute:
      return CudaBackendImpl.CUTE
    if self.enable_tma:
      return CudaBackendImpl.TMA
    if self.enable_tma is False and self.enable_cute is False:
      return CudaBackendImpl.NATIVE
    return CudaBackendImpl.AUTO


Backend = CUDABackend


def _coerce_backend(backend, *, source: str) -> CUDABackend:
  if isinstance(backend, CUDABackend):
    return backend
  if isinstance(backend, str):
    if backend.lower() == "cuda":
      return CUDABackend()
    raise NotImplementedError(
      f"ffpa_attn_func: {source}={backend!r} is not available in the B200 build; the only backend "
      "is 'cuda' (hand-written sm_100a kernels). There is no Triton / CuTe-DSL / SDPA route.")
  raise TypeError(
    f"ffpa_attn_func: {source} must be a str or Backend instance, got {type(backend).__name__}")


def _validate_attn_mask_shape(attn_mask, batch, nheads_q, seqlen_q, seqlen_k) -> None:
  if attn_mask.dim() not in (2, 3, 4):
    raise ValueError(f"ffpa_attn_func: attn_mask must be 2-D, 3-D or 4-D, got {attn_mask.dim()}-D")
  full = (batch, nheads_q, seqlen_q, seqlen_k)
  if attn_mask.dim() == 2:
    want = {2: (seqlen_q, seqlen_k)}[2]
    dims = list(zip(attn_mask.shape, want))
  elif attn_mask.dim() == 3:
    dims = list(zip(attn_mask.shape, (batch, seqlen_q, seqlen_k)))
  else:
    dims = list(zip(attn_mask.shape, full))
  for got, exp in dims:
    if got != 1 and got != exp:
      raise ValueError(
        f"ffpa_attn_func: attn_mask shape {tuple(attn_mask.shape)} is not broadcastable to {full}")


@dataclass
class FFPAAttnMeta:
  """Non-tensor options carried through the autograd Function (reference: functional.py:593-608)."""
  attn_meta: AttentionMeta = field(default_factory=AttentionMeta)
  forward_meta: CUDABackend = field(default_factory=CUDABackend)
  backward_meta: CUDABackend = field(default_factory=CUDABackend)

  @classmethod
  def from_kwargs(cls, **kwargs) -> "FFPAAttnMeta":
    """Pops ``backend`` / ``forward_backend`` / ``backward_backend``; any other keyword is a
    TypeError (reference: functional.py:610-652)."""
    backend = kwargs.pop("backend", None)
    fwd = kwargs.pop("forward_backend", None)
    bwd = kwargs.pop("backward_backend", None)
    if kwargs:
      unexpected = ", ".join(sorted(kwargs))
      raise TypeError(f"ffpa_attn_func() got unexpected keyword argument(s): {unexpected}")
    fwd = None if fwd is None else _coerce_backend(fwd, source="forward_backend")
    bwd = None if bwd is None else _coerce_backend(bwd, source="backward_backend")
    if fwd is None and bwd is None and backend is not None:
      fwd = bwd = _coerce_backend(backend, source="backend")
    return cls(forward_meta=fwd or CUDABackend(), backward_meta=bwd or CUDABackend())

  def fallback(self, query, key, attn_mask, dropout_p) -> bool:
    """The reference delegates small-D / short-sequence shapes to aten SDPA here
    (functional.py:676-724). The B200 kernel covers them itself, so nothing ever falls back."""
    return False

  def normalize_inputs(self, query, key, value, attn_mask, dropout_p, is_causal, scale, enable_gqa):
    """Validation with the reference's error classes (functional.py:726-849)."""
    if not 0.0 <= dropout_p <= 1.0:
      raise ValueError(f"ffpa_attn_func: dropout_p must be in [0, 1], got {dropout_p}")
    if dropout_p >= 1.0:
      raise ValueError("ffpa_attn_func: dropout_p=1.0 is not supported by SDPA fused kernels")
    if attn_mask is not None and is_causal:
      raise RuntimeError("ffpa_attn_func: explicit attn_mask should not be set when is_causal=True")
    if attn_mask is not None and attn_mask.dtype == torch.bool and attn_mask.requires_grad:
      raise TypeError("ffpa_attn_func: boolean attn_mask cannot require gradients")
    self.attn_meta.is_causal = bool(is_causal)
    self.attn_meta.dropout_p = float(dropout_p)
    self.attn_meta.is_grad_enabled = torch.is_grad_enabled()
    self.forward_meta.is_causal = bool(is_causal)
    # *_hybrid=None means "auto": on when causal + the matching quant path, to protect the precision of the early
    # rows (reference: functional.py:781-794); explicit True / False is honoured as given
    if self.forward_meta.fp8_hybrid is None:
      self.forward_meta.fp8_hybrid = bool(self.forward_meta.enable_fp8 and is_causal)
    if self.forward_meta.fp4_hybrid is None:
      self.forward_meta.fp4_hybrid = False
    if query.dtype not in (torch.float16, torch.bfloat16):
      raise TypeError(f"ffpa_attn_func only supports fp16/bf16, got {query.dtype}")
    if key.dtype != query.dtype or value.dtype != query.dtype:
      raise TypeError("ffpa_attn_func: query/key/value must share one dtype")
    if query.dim() != 4 or key.dim() != 4 or value.dim() != 4:
      raise ValueError("query/key/value must be 4-D [B, H, N, D] tensors")
    if query.size(0) != key.size(0) or query.size(0) != value.size(0):
      raise ValueError("query/key/value must share the same batch size")
    if key.size(1) != value.size(1):
      raise ValueError(f"key and value must share the same num_heads, got Nh_k={key.size(1)}, Nh_v={value.size(1)}")
    if query.size(1) % key.size(1) != 0:
      raise ValueError(
        "query num_heads must be an integer multiple of key/value num_heads (GQA/MQA), "
        f"got Nh_q={query.size(1)}, Nh_kv={key.size(1)}")
    if key.size(2) != value.size(2):
      raise ValueError(f"key and value must share the same seqlen, got Nk={key.size(2)}, Nv={value.size(2)}")
    if query.size(3) != key.size(3) or query.size(3) != value.size(3):
      raise ValueError("query/key/value must share the same head dim")
    if not enable_gqa and query.size(1) != key.size(1):
      raise ValueError(
        f"enable_gqa=False but query num_heads ({query.size(1)}) != key/value num_heads "
        f"({key.size(1)}). Set enable_gqa=True or use matching head counts.")
    if is_causal and key.size(2) < query.size(2):
      raise ValueError(
        "is_causal=True requires Nkv >= Nq (queries are aligned to the KV tail), "
        f"got Nq={query.size(2)}, Nkv={key.size(2)}")
    if query.size(3) % 8 != 0:
      raise ValueError(f"head dim must be a multiple of 8, got {query.size(3)}")
    if query.size(3) > 1024:
      raise NotImplementedError(f"head dim {query.size(3)} > 1024 is not supported")
    if query.device.type != "cuda":
      raise RuntimeError(
        "ffpa_attn_func: tensors must be CUDA tensors on an sm_100 device; this build has no CPU "
        "or SDPA fallback")
    if scale is None:
      self.attn_meta.scale = 1.0 / math.sqrt(query.size(-1))
    else:
      self.attn_meta.scale = float(scale)
    return self

  def normalize_attn_mask(self, query, key, attn_mask):
    """SDPA mask -> additive 4-D bias; bool True = keep -> 0 / -inf in q.dtype; 2-D / 3-D masks
    become 4-D views; last dim contiguous (reference: functional.py:851-911)."""
    if attn_mask is None:
      return None
    if attn_mask.device != query.device:
      raise TypeError(
        f"ffpa_attn_func: attn_mask must be on the same device as query, got {attn_mask.device} and {query.device}")
    if attn_mask.dtype not in (torch.bool, torch.float32, query.dtype):
      raise TypeError(
        "ffpa_attn_func: attn_mask dtype must be bool, torch.float32, or match query dtype, "
        f"got attn_mask.dtype={attn_mask.dtype} and query.dtype={query.dtype}")
    batch, nheads_q, seqlen_q, _ = query.shape
    _validate_attn_mask_shape(attn_mask, batch, nheads_q, seqlen_q, key.size(2))
    if attn_mask.dtype == torch.bool:
      zero = torch.zeros((), dtype=query.dtype, device=query.device)
      neg_inf = torch.full((), float("-inf"), dtype=query.dtype, device=query.device)
      attn_bias = torch.where(attn_mask, zero, neg_inf)
    else:
      attn_bias = attn_mask
    if attn_bias.dim() == 2:
      attn_bias = attn_bias.view(1, 1, attn_bias.size(0), attn_bias.size(1))
    elif attn_bias.dim() == 3:
      attn_bias = attn_bias.view(attn_bias.size(0), 1, attn_bias.size(1), attn_bias.size(2))
    if attn_bias.stride(-1) != 1:
      attn_bias = attn_bias.contiguous()
    return attn_bias

  def normalize(self, query, key, value, attn_mask, dropout_p, is_causal, scale, enable_gqa):
    self.normalize_inputs(query, key, value, attn_mask, dropout_p, is_causal, scale, enable_gqa)
    return self, query, key, value, self.normalize_attn_mask(query, key, attn_mask)


def _reserve_large_d_dropout_rng(q: torch.Tensor, k: torch.Tensor, dropout_p: float) -> torch.Tensor:
  """Reserve one Philox output per logical score [B, Hq, Nq, Nkv], rounded up to 4, from the CUDA
  generator; returns CPU int64 [seed, offset] (reference: functional.py:518-540)."""
  if dropout_p <= 0.0:
    return torch.empty(0, dtype=torch.int64)
  seed = int(torch.cuda.initial_seed())
  offset = int(torch.cuda._get_rng_state_offset())
  attn_elems = q.size(0) * q.size(1) * q.size(2) * k.size(2)
  torch.cuda._set_rng_state_offset(offset + ((attn_elems + 3) // 4) * 4)
  return torch.tensor([seed, offset], dtype=torch.int64)


class _FFPAAttnFunc(torch.autograd.Function):
  """fwd -> torch.ops.ffpa_attn._fwd_cuda ; bwd -> torch.ops.ffpa_attn._bwd_cuda.
  Saved tensors keep the reference tuple (q, k, v, O, lse, rng_state, unused), LSE fp32 natural log
  [B, Hq, Nq] and rng_state CPU int64 [seed, offset] (functional.py:1066-1077)."""

  @staticmethod
  def forward(ctx, q, k, v, attn_bias, meta: FFPAAttnMeta):
    is_grad = meta.attn_meta.is_grad_enabled and any(
      x.requires_grad for x in (q, k, v, attn_bias) if x is not None)
    fm = meta.forward_meta
    set_cuda_backend_impl(fm.impl_hint)
    rng_state = _reserve_large_d_dropout_rng(q, k, meta.attn_meta.dropout_p)
    O, lse = _ffpa_attn_forward_cuda(  # noqa: E741
      q, k, v, None, attn_bias, fm.stages, fm.acc_code, int(meta.attn_meta.is_causal),
      meta.attn_meta.scale, meta.attn_meta.dropout_p,
      int(rng_state[0].item()) if rng_state.numel() else 0,
      int(rng_state[1].item()) if rng_state.numel() else 0,
      fm.fp8_smooth_k, fm.fp8_smooth_v, _QUANT_CODE[fm.fp8_q_quant_method],
      _QUANT_CODE[fm.fp8_k_quant_method], _QUANT_CODE[fm.fp8_v_quant_method],
      _PV_ACC_CODE[fm.fp8_pv_acc_type], _QK_MM_TYPE_CODE[fm.fp8_qk_mm_type],
      bool(fm.fp8_hybrid), fm.fp8_hybrid_n_early, bool(fm.fp4_hybrid), fm.fp4_hybrid_n_early)
    if is_grad:
      unused = torch.empty(0, dtype=torch.uint8, device=q.device)
      ctx.save_for_backward(q.contiguous(), k.contiguous(), v.contiguous(), O.contiguous(), lse,
                            rng_state, unused)
      ctx.attn_bias = attn_bias
      ctx.meta = meta
    return O

  @staticmethod
  def backward(ctx, d_o):
    q, k, v, O, lse, rng_state, _unused = ctx.saved_tensors  # noqa: E741
    meta: FFPAAttnMeta = ctx.meta
    if not CUDA_BWD_AVAILABLE:
      raise NotImplementedError("the sm_100a backward kernels are not built into libffpa_b200.so")
    if q.size(-1) > 1024:
      raise NotImplementedError("ffpa_attn backward supports head_dim <= 1024 on sm_100a")
    bias = ctx.attn_bias
    p_drop = meta.attn_meta.dropout_p
    stages = meta.backward_meta.stages
    _BWD_MIN_WORKSPACE.set(bool(getattr(meta.backward_meta, "bwd_min_workspace", False)))
    if bias is None and p_drop <= 0.0:
      dq, dk, dv = _ffpa_attn_backward_cuda(
        q, k, v, O, lse, d_o.contiguous(), stages, int(meta.attn_meta.is_causal), meta.attn_meta.scale)
      return dq, dk, dv, None, None
    # bias and/or dropout: replay them in the backward kernels; dBias = P * (dP - delta)
    # (reference math: triton/_ffpa_bwd.py:692-855; returned tuple: functional.py:1081-1172)
    want_dbias = bias is not None and bias.requires_grad
    seed = int(rng_state[0].item()) if rng_state.numel() else 0
    offset = int(rng_state[1].item()) if rng_state.numel() else 0
    dq, dk, dv, dbias = torch.ops.ffpa_attn._bwd_cuda_ex(
      q, k, v, O, lse, d_o.contiguous(), bias if bias is not None else q.new_empty(0),
      int(stages) if stages is not None else 0, int(meta.attn_meta.is_causal), float(meta.attn_meta.scale),
      float(p_drop), seed, offset, bool(want_dbias))
    return dq, dk, dv, (dbias if want_dbias else None), None


@torch._dynamo.disable
def _ffpa_apply(q, k, v, attn_bias, meta):
  return _FFPAAttnFunc.apply(q, k, v, attn_bias, meta)


class FFPAAttnFunc:
  """Callable facade with the reference's name (functional.py:1195-1216)."""

  @staticmethod
  def apply(q, k, v, attn_bias, meta):
    return _ffpa_apply(q, k, v, attn_bias, meta)
