"""CPU oracle for the ffpa_attn hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU in fp64/fp32, the algorithm the reference implements for
``ffpa_attn_func`` forward and backward.  It exists so that the CUDA path can be checked against
an independent implementation.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
(``ffpa-attn_b200/ffpa_attn``) never does, and has no CPU route at all.

Parity status: PINNED.  ``tests/golden/*.npz`` were produced by importing the reference package
itself in the build container (``oracle/make_golden.py``; ``ffpa_attn_func(..., backend="sdpa")``
on CPU, the reference's own CPU-runnable route, /root/reference/src/ffpa_attn/
ffpa_attn_interface.py:163-176) and ``tests/test_oracle.py`` checks every function here against
them; the Philox generator is pinned by the Random123 known-answer vectors.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
  "attention_fwd",
  "attention_bwd",
  "philox4x32_10",
  "dropout_keep_mask",
  "attn_flops",
  "sdpa_cpu",
]


# --------------------------------------------------------------------------------------------
# Philox-4x32-10 (csrc/cuffpa/native/prefill.cuh:398-422); same generator as curand / SDPA.
# --------------------------------------------------------------------------------------------
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: tuple[int, int]) -> np.ndarray:
  """counter: uint32 array [..., 4]; key: (k0, k1). Returns uint32 array [..., 4]."""
  c = np.asarray(counter, dtype=np.uint64).copy()
  k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
  c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
  for _ in range(10):
    p0 = _M0 * c0  # 64-bit products
    p1 = _M1 * c2
    hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
    hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
    n0 = hi1 ^ c1 ^ np.uint64(k0)
    n2 = hi0 ^ c3 ^ np.uint64(k1)
    c0, c1, c2, c3 = n0, lo1, n2, lo0
    k0 = (k0 + _W0) & 0xFFFFFFFF
    k1 = (k1 + _W1) & 0xFFFFFFFF
  return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def dropout_keep_mask(batch: int, heads_q: int, seqlen_q: int, seqlen_kv: int, p: float, seed: int,
                      offset: int) -> np.ndarray:
  """Boolean keep mask [B, Hq, Nq, Nkv].

  element index e = offset + ((b*Hq + h)*Nq + q)*Nkv + k; uint32 = lane (e & 3) of
  Philox(key=seed, counter=(e >> 2, 0, 0, 0)); u = (uint32 + 1) * 2**-32; keep iff u > p
  (csrc/cuffpa/native/prefill.cuh:424-452, 506-546).
  """
  n = batch * heads_q * seqlen_q * seqlen_kv
  e = np.arange(n, dtype=np.uint64) + np.uint64(offset)
  quad = e >> np.uint64(2)
  ctr = np.zeros((n, 4), dtype=np.uint64)
  ctr[:, 0] = quad & _MASK32
  ctr[:, 1] = quad >> np.uint64(32)
  r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
  lane = (e & np.uint64(3)).astype(np.int64)
  u32 = r[np.arange(n), lane]
  u = (u32.astype(np.float32) + np.float32(1.0)) * np.float32(2.3283064365386963e-10)
  return (u > np.float32(p)).reshape(batch, heads_q, seqlen_q, seqlen_kv)


# --------------------------------------------------------------------------------------------
# forward / backward restatement
# --------------------------------------------------------------------------------------------
def _as64(x) -> np.ndarray:
  try:
    import torch

    if isinstance(x, torch.Tensor):
      return x.detach().to(torch.float64).cpu().numpy()
  except ImportError:  # pragma: no cover
    pass
  return np.asarray(x, dtype=np.float64)


def _scores(q, k, bias, causal, scale):
  """scale * Q K^T + bias with bottom-right causal mask and GQA head mapping.

  kv_head = q_head // (Hq/Hkv)             (csrc/cuffpa/native/sm_80/split_d.cuh:135-136)
  key k visible to row r iff k <= r + Nkv-Nq (csrc/cuffpa/native/prefill.cuh:320-349)
  bias is added after scaling               (csrc/cuffpa/native/prefill.cuh:548-555)
  """
  B, Hq, Nq, D = q.shape
  Hkv, Nkv = k.shape[1], k.shape[2]
  g = Hq // Hkv
  kk = np.repeat(k, g, axis=1)
  s = np.einsum("bhqd,bhkd->bhqk", q, kk) * scale
  if bias is not None:
    s = s + np.broadcast_to(bias, s.shape)
  if causal:
    r = np.arange(Nq)[:, None]
    c = np.arange(Nkv)[None, :]
    s = np.where(c <= r + (Nkv - Nq), s, -np.inf)
  return s


def attention_fwd(q, k, v, bias=None, causal=False, scale=None, dropout_p=0.0, philox_seed=0,
                  philox_offset=0):
  """Returns (O [B,Hq,Nq,D] float64, LSE [B,Hq,Nq] float64 natural log).

  softmax over keys, dropout applied to the normalised probabilities with 1/(1-p) rescale while
  the row sum / LSE ignore dropout (csrc/cuffpa/native/prefill.cuh:506-546, 746-762);
  rows without a visible key give O = 0, LSE = -inf
  (src/ffpa_attn/cute/_fwd_d512_sm100.py:2635-2646).
  """
  q, k, v = _as64(q), _as64(k), _as64(v)
  if bias is not None:
    bias = _as64(bias)
  B, Hq, Nq, D = q.shape
  Hkv, Nkv = k.shape[1], k.shape[2]
  if scale is None:
    scale = 1.0 / math.sqrt(D)  # src/ffpa_attn/functional.py:844-847
  s = _scores(q, k, bias, causal, scale)
  m = s.max(axis=-1, keepdims=True)
  m_safe = np.where(np.isfinite(m), m, 0.0)
  e = np.exp(s - m_safe)
  l = e.sum(axis=-1, keepdims=True)
  with np.errstate(divide="ignore", invalid="ignore"):
    p = np.where(l > 0, e / l, 0.0)
    lse = np.where(l[..., 0] > 0, m_safe[..., 0] + np.log(l[..., 0]), -np.inf)
  if dropout_p > 0.0:
    keep = dropout_keep_mask(B, Hq, Nq, Nkv, dropout_p, philox_seed, philox_offset)
    p = np.where(keep, p / (1.0 - dropout_p), 0.0)
  vv = np.repeat(v, Hq // Hkv, axis=1)
  o = np.einsum("bhqk,bhkd->bhqd", p, vv)
  return o, lse


def attention_bwd(q, k, v, d_o, bias=None, causal=False, scale=None, dropout_p=0.0, philox_seed=0,
                  philox_offset=0):
  """Returns (dQ, dK, dV, dBias_full[B,Hq,Nq,Nkv]) in float64.

  delta = rowsum(dO * O); P = exp(scale*S + bias - LSE); dP = dO V^T (x dropout multiplier);
  dS = P * (dP - delta); dQ = scale * dS K; dK = scale * dS^T Q; dV = P_drop^T dO;
  dBias = dS; GQA sums dK/dV over the group (src/ffpa_attn/triton/_ffpa_bwd.py:236-306, 692-855).
  """
  q, k, v, d_o = _as64(q), _as64(k), _as64(v), _as64(d_o)
  if bias is not None:
    bias = _as64(bias)
  B, Hq, Nq, D = q.shape
  Hkv, Nkv = k.shape[1], k.shape[2]
  g = Hq // Hkv
  if scale is None:
    scale = 1.0 / math.sqrt(D)
  s = _scores(q, k, bias, causal, scale)
  m = s.max(axis=-1, keepdims=True)
  m_safe = np.where(np.isfinite(m), m, 0.0)
  e = np.exp(s - m_safe)
  l = e.sum(axis=-1, keepdims=True)
  with np.errstate(divide="ignore", invalid="ignore"):
    p = np.where(l > 0, e / l, 0.0)
  mult = np.ones_like(p)
  if dropout_p > 0.0:
    keep = dropout_keep_mask(B, Hq, Nq, Nkv, dropout_p, philox_seed, philox_offset)
    mult = np.where(keep, 1.0 / (1.0 - dropout_p), 0.0)
  pd = p * mult
  kk = np.repeat(k, g, axis=1)
  vv = np.repeat(v, g, axis=1)
  o = np.einsum("bhqk,bhkd->bhqd", pd, vv)
  delta = (d_o * o).sum(axis=-1, keepdims=True)
  dp = np.einsum("bhqd,bhkd->bhqk", d_o, vv) * mult
  ds = p * (dp - delta)
  dq = np.einsum("bhqk,bhkd->bhqd", ds, kk) * scale
  dk_full = np.einsum("bhqk,bhqd->bhkd", ds, q) * scale
  dv_full = np.einsum("bhqk,bhqd->bhkd", pd, d_o)
  dk = dk_full.reshape(B, Hkv, g, Nkv, D).sum(axis=2)
  dv = dv_full.reshape(B, Hkv, g, Nkv, D).sum(axis=2)
  return dq, dk, dv, ds


def attn_flops(batch: int, heads_q: int, seqlen_q: int, seqlen_kv: int, head_dim: int,
               causal: bool = False, mode: str = "fwd") -> float:
  """Dominant-GEMM FLOPs: fwd 4*B*Hq*D*valid_pairs, bwd 2.5x (src/ffpa_attn/cli/_flops.py:36-76)."""
  if causal:
    off = seqlen_kv - seqlen_q
    pairs = sum(min(seqlen_kv, r + off + 1) for r in range(seqlen_q)) if seqlen_q < 4096 else (
      seqlen_q * off + seqlen_q * (seqlen_q + 1) // 2)
  else:
    pairs = seqlen_q * seqlen_kv
  f = 4.0 * batch * heads_q * head_dim * pairs
  return f * (2.5 if mode == "bwd" else 3.5 if mode == "fwd_bwd" else 1.0)


def sdpa_cpu(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, enable_gqa=False):
  """The reference's own CPU-runnable route: ``ffpa_attn_func(..., backend="sdpa")`` hands the
  call to aten SDPA unchanged (src/ffpa_attn/ffpa_attn_interface.py:163-176). Used as the timed
  CPU baseline in bench.py and as a second opinion in tests."""
  import torch

  return torch._C._nn.scaled_dot_product_attention(
    q, k, v, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
    enable_gqa=enable_gqa)
