"""Public API: ``ffpa_attn_func`` with the signature of
``torch.nn.functional.scaled_dot_product_attention`` (reference:
/root/reference/src/ffpa_attn/ffpa_attn_interface.py:71-189).

query [B, Hq, Nq, D], key/value [B, Hkv, Nkv, D], fp16/bf16 CUDA tensors; self/cross attention,
GQA/MQA (``enable_gqa=True``), bottom-right aligned causal (``Nkv >= Nq``), bool or additive
``attn_mask`` broadcastable to [B, Hq, Nq, Nkv], dropout (Philox, SDPA-compatible), optional
``scale``.  ``backend`` / ``forward_backend`` / ``backward_backend`` keywords are accepted with the
value ``"cuda"`` or a :class:`CUDABackend`; anything else raises -- there is one backend.
"""
from __future__ import annotations

import torch

from .functional import FFPAAttnFunc, FFPAAttnMeta, _coerce_backend


def ffpa_attn_func(
  query: torch.Tensor,
  key: torch.Tensor,
  value: torch.Tensor,
  attn_mask: torch.Tensor | None = None,
  dropout_p: float = 0.0,
  is_causal: bool = False,
  scale: float | None = None,
  enable_gqa: bool = False,
  **kwargs: object,
) -> torch.Tensor:
  meta = FFPAAttnMeta.from_kwargs(**kwargs)
  meta, query, key, value, attn_bias = meta.normalize(
    query, key, value, attn_mask, dropout_p, is_causal, scale, enable_gqa)
  return FFPAAttnFunc.apply(query, key, value, attn_bias, meta)


_VARLEN_UNSUPPORTED = ("window_size", "softcap", "sink", "attention_mask", "attn_mask", "block_mask", "score_mod",
                       "aux_tensors", "seqused_k", "block_table", "num_splits", "alibi_slopes")


def ffpa_attn_varlen_func(
  q: torch.Tensor,
  k: torch.Tensor,
  v: torch.Tensor,
  cu_seqlens_q: torch.Tensor,
  cu_seqlens_k: torch.Tensor | None,
  max_seqlen_q: int,
  max_seqlen_k: int,
  *,
  dropout_p: float = 0.0,
  softmax_scale: float | None = None,
  causal: bool = False,
  enable_gqa: bool = False,
  return_lse: bool = False,
  **kwargs: object,
):
  """Packed-THD variable-length attention with the reference's signature
  (/root/reference/src/ffpa_attn/ffpa_attn_interface.py:192-279; flash_attn_varlen_func style):
  ``q`` [T_q, Hq, D], ``k``/``v`` [T_k, Hkv, D], int32 ``cu_seqlens_*`` of length B+1 starting at 0,
  lower-right causal per sequence, LSE ``[Hq, T_q]`` fp32 when ``return_lse``.

  B200 build: ONE launch set for the whole packed batch (``torch.ops.ffpa_attn._varlen_fwd_cuda`` /
  ``_varlen_bwd_cuda`` -> C ABI 2 -> the same sm_100a kernels as the dense path, whose work items read
  their sequence's token range from ``cu_seqlens`` on the device). ``cu_seqlens`` are never read on the
  host -- no synchronisation, CUDA-graph capturable -- exactly like the reference, which validates only
  dtypes and shapes (/root/reference/src/ffpa_attn/cute/__init__.py:466-571); ``max_seqlen_*`` size the
  grid, so sequences longer than the stated maximum are a caller error. Forward and backward support every
  head dim of the dense path (8..1024).
  """
  # flash-attn style callers pass the DEFAULTS of options this path does not implement; like the reference's
  # _check_supported_options (/root/reference/src/ffpa_attn/cute/__init__.py:107-118) accept those and refuse
  # anything that would change the result
  neutral = {"window_size": (None, (None, None), (-1, -1)), "softcap": (None, 0.0, 0)}
  for name in _VARLEN_UNSUPPORTED:
    val = kwargs.pop(name, None)
    if val is None:
      continue
    if name in neutral and (tuple(val) if isinstance(val, (tuple, list)) else val) in neutral[name]:
      continue
    raise NotImplementedError(f"ffpa_attn_varlen_func: option {name}={val!r} is not supported")
  for name in ("backend", "forward_backend", "backward_backend"):
    val = kwargs.pop(name, None)
    if val is not None:
      _coerce_backend(val, source=name)   # 'cuda' / CUDABackend; other strings NotImplementedError, other types TypeError
  if kwargs:
    raise TypeError(f"ffpa_attn_varlen_func() got unexpected keyword argument(s): {', '.join(sorted(kwargs))}")
  if dropout_p != 0.0:
    raise NotImplementedError("ffpa_attn_varlen_func: dropout_p must be 0.0")
  if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
    raise ValueError("q/k/v must be packed THD tensors [T, H, D]")
  if q.dtype not in (torch.float16, torch.bfloat16):
    raise TypeError(f"ffpa_attn_varlen_func only supports fp16/bf16, got {q.dtype}")
  if k.dtype != q.dtype or v.dtype != q.dtype:
    raise TypeError("ffpa_attn_varlen_func: q/k/v must share one dtype")
  if k.shape != v.shape or k.size(2) != q.size(2):
    raise ValueError("k and v must share [T_k, H_kv, D] and q's head dim")
  if cu_seqlens_k is None:
    cu_seqlens_k = cu_seqlens_q
  for name, cu in (("cu_seqlens_q", cu_seqlens_q), ("cu_seqlens_k", cu_seqlens_k)):
    if cu.dtype != torch.int32:
      raise TypeError(f"{name} must be int32, got {cu.dtype}")
    if cu.dim() != 1 or cu.numel() < 2:
      raise ValueError(f"{name} must be a 1-D tensor of length B+1")
  if cu_seqlens_q.numel() != cu_seqlens_k.numel():
    raise ValueError("cu_seqlens_q and cu_seqlens_k must describe the same batch size")
  if not enable_gqa and q.size(1) != k.size(1):
    raise ValueError("enable_gqa=False but H_q != H_kv")
  if q.size(1) % k.size(1) != 0:
    raise ValueError("H_q must be an integer multiple of H_kv")
  if q.size(2) % 8 != 0 or q.size(2) > 1024:
    raise NotImplementedError(f"ffpa_attn_varlen_func supports head_dim % 8 == 0 and <= 1024, got {q.size(2)}")
  for t in (q, k, v, cu_seqlens_q, cu_seqlens_k):
    if t.device.type != "cuda":
      raise RuntimeError("ffpa_attn_varlen_func: all tensors must be CUDA tensors (there is no CPU / SDPA fallback path)")
  if q.size(0) == 0 or k.size(0) == 0:
    out = torch.zeros_like(q)
    lse = torch.full((q.size(1), q.size(0)), float("-inf"), dtype=torch.float32, device=q.device)
    return (out, lse) if return_lse else out
  scale = float(softmax_scale) if softmax_scale is not None else q.size(-1) ** -0.5
  out, lse = _FFPAVarlenFunc.apply(q, k, v, cu_seqlens_q.contiguous(), cu_seqlens_k.contiguous(),
                                   int(max_seqlen_q), int(max_seqlen_k), bool(causal), scale)
  return (out, lse) if return_lse else out


class _FFPAVarlenFunc(torch.autograd.Function):
  """Autograd glue of the packed path (reference: FFPAAttnVarlenFunc, functional.py:1218-1250)."""

  @staticmethod
  def forward(ctx, q, k, v, cu_q, cu_k, max_q, max_k, causal, scale):
    qc, kc, vc = (t if t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 else t.contiguous()
                  for t in (q, k, v))
    out, lse = torch.ops.ffpa_attn._varlen_fwd_cuda(qc, kc, vc, cu_q, cu_k, max_q, max_k, int(causal), scale)
    ctx.save_for_backward(qc, kc, vc, out, lse, cu_q, cu_k)
    ctx.args = (max_q, max_k, int(causal), scale)
    ctx.set_materialize_grads(False)   # d_lse stays None unless the caller differentiates through the LSE
    return out, lse

  @staticmethod
  def backward(ctx, d_o, d_lse):
    q, k, v, out, lse, cu_q, cu_k = ctx.saved_tensors
    max_q, max_k, causal, scale = ctx.args
    if d_o is None:
      d_o = torch.zeros_like(out)
    # the LSE output is differentiable (reference: cute/_bwd_preprocess.py:6-15): dS gains P * dLSE
    d_lse = d_lse.float().contiguous() if d_lse is not None else None
    dq, dk, dv = torch.ops.ffpa_attn._varlen_bwd_cuda(q, k, v, out, lse, d_o.contiguous(), cu_q, cu_k,
                                                     max_q, max_k, causal, scale, d_lse)
    return dq, dk, dv, None, None, None, None, None, None
