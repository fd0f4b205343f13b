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
