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

from .functional import FFPAAttnFunc, FFPAAttnMeta


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

  B200 build: every sequence is a zero-copy strided view handed to the dense sm_100a kernels (the tensor
  maps honour the THD strides), so forward AND backward (autograd) work for every head dim the dense
  path supports; the price is one launch set per sequence and one host read of ``cu_seqlens``.
  A single-launch varlen kernel is listed as a next step in DESIGN.md.
  """
  for name in _VARLEN_UNSUPPORTED:
    if name in kwargs and kwargs[name] is not None:
      raise NotImplementedError(f"ffpa_attn_varlen_func: option {name!r} is not supported")
  backend_kw = {k_: kwargs.pop(k_) for k_ in ("backend", "forward_backend", "backward_backend") if k_ in kwargs}
  for name in _VARLEN_UNSUPPORTED:
    kwargs.pop(name, None)
  if kwargs:
    raise TypeError(f"ffpa_attn_varlen_func() got unexpected keyword argument(s): {', '.join(sorted(kwargs))}")
  if dropout_p != 0.0:
    raise NotImplementedError("ffpa_attn_varlen_func: dropout_p must be 0.0")
  if q.dtype not in (torch.float16, torch.bfloat16):
    raise TypeError(f"ffpa_attn_varlen_func only supports fp16/bf16, got {q.dtype}")
  if q.dim() != 3 or k.dim() != 3 or v.dim() != 3:
    raise ValueError("q/k/v must be packed THD tensors [T, H, D]")
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
  cq, ck = cu_seqlens_q.tolist(), cu_seqlens_k.tolist()
  if cq[0] != 0 or ck[0] != 0 or cq[-1] != q.size(0) or ck[-1] != k.size(0):
    raise ValueError("cu_seqlens must start at 0 and end at the packed token count")
  if any(b < a for a, b in zip(cq, cq[1:])) or any(b < a for a, b in zip(ck, ck[1:])):
    raise ValueError("cu_seqlens must be non-decreasing")
  scale = softmax_scale if softmax_scale is not None else q.size(-1) ** -0.5

  out = torch.empty_like(q)
  lse = torch.full((q.size(1), q.size(0)), float("-inf"), dtype=torch.float32, device=q.device) if return_lse else None
  outs = []
  for b in range(len(cq) - 1):
    nq, nk = cq[b + 1] - cq[b], ck[b + 1] - ck[b]
    if nq == 0:
      continue
    if nk == 0:
      outs.append((b, None))
      continue
    if causal and nk < nq:
      raise ValueError(f"causal varlen attention requires Nkv >= Nq per sequence (sequence {b}: {nq} vs {nk})")
    qb = q[cq[b]:cq[b + 1]].transpose(0, 1).unsqueeze(0)   # [1, H, n, D] view over the THD storage
    kb = k[ck[b]:ck[b + 1]].transpose(0, 1).unsqueeze(0)
    vb = v[ck[b]:ck[b + 1]].transpose(0, 1).unsqueeze(0)
    if return_lse and not torch.is_grad_enabled():
      from .cuda import _ffpa_attn_forward_cuda

      ob, lb = _ffpa_attn_forward_cuda(qb, kb, vb, None, None, 0, 1, int(causal), scale)
      lse[:, cq[b]:cq[b + 1]] = lb[0]
    else:
      ob = ffpa_attn_func(qb, kb, vb, is_causal=causal, scale=scale, enable_gqa=enable_gqa, **backend_kw)
      if return_lse:
        from .cuda import _ffpa_attn_forward_cuda

        with torch.no_grad():
          _, lb = _ffpa_attn_forward_cuda(qb.detach(), kb.detach(), vb.detach(), None, None, 0, 1, int(causal), scale)
        lse[:, cq[b]:cq[b + 1]] = lb[0]
    outs.append((b, ob[0].transpose(0, 1)))
  if torch.is_grad_enabled() and any(t.requires_grad for t in (q, k, v)):
    pieces = []
    for b in range(len(cq) - 1):
      nq = cq[b + 1] - cq[b]
      if nq == 0:
        continue
      match = [o for bb, o in outs if bb == b]
      pieces.append(match[0] if match and match[0] is not None else q.new_zeros(nq, q.size(1), q.size(2)))
    out = torch.cat(pieces, dim=0) if pieces else out
  else:
    out.zero_()
    for b, ob in outs:
      if ob is not None:
        out[cq[b]:cq[b + 1]] = ob
  return (out, lse) if return_lse else out
