"""CPU oracle for the FP8 forward -- TEST INFRASTRUCTURE ONLY (same rules as attention_oracle.py: only
``tests/`` may import it; the product has no CPU route).

Restates the reference's QUANTISED numerics for ``CUDABackend(enable_fp8=True)`` so the sm_100a FP8 kernel is
checked against what the reference computes, not merely against exact attention (paths relative to
/root/reference/csrc/cuffpa/cute/fp8):

  quantize_per_block   quantize_fp8.cuh:67-168   s = amax_block / 448 per (b, h, 128-row block); inv_s = 1/s
                                                 (0 when s == 0); y = e4m3_rn_satfinite(x * inv_s); smooth-K
                                                 subtracts the sequence mean BEFORE the amax (:106-116, 151-156)
  kv_mean              smooth_k.cuh:61-137       fp32 column sums / Nkv, emitted in the input dtype (used by the
                                                 quantiser) and in fp32 (used by the LSE correction)
  fp8_attention_fwd    sm_120/split_d.cuh:26-43  S = (Q8 K8^T) qs ks scale in the log2 domain; online softmax per
                       fp8_pscale.cuh:11-76      128-key tile; P8 = e4m3(P vs 448) ("Mode B", fixed p_scale = 1/448:
                       sm_120/split_d.cuh:753-762 (P vs 448) @ (V / vs) = 448 (P @ V)); row sum from the fp32 P;
                                                 O = acc / 448 / l; LSE = (m + log2 l) ln2 + scale qs dot(Q8_row, km)

Parity status: PARTIALLY PINNED.  The reference's FP8 kernels are sm_120-only (tests gate on
``major == 12``, /root/reference/tests/test_ffpa_fp8.py:23-33) and ship no golden vectors, so this restatement
cannot be run against reference outputs here or on the B200 box.  It is pinned on what can be: the e4m3
rounding against ``torch.float8_e4m3fn`` known answers, the reference's own acceptance bounds versus exact fp32
attention (``tests/test_ffpa_fp8.py:71,86``: O 4e-2 dense / 1e-1 causal, LSE 5e-2) and algebraic identities
(smooth-K leaves O unchanged up to quantisation; vs cancels).  tests/test_oracle.py holds those checks.

Known, deliberate differences of the sm_100a kernel that the GPU tests bound instead of hiding:
  * P8 = e4m3(P (vs / vref) 28) with vref = max_tile vs, O = acc vref / 28 / l (28 = 448 / 2^4 leaves head room
    for the lazy-rescale threshold): the same fixed-scale scheme with a different constant, so individual P8
    roundings differ from the formula above;
  * smooth-K subtracts the fp32 mean (the reference rounds the mean to the input dtype first), and the LSE
    correction uses dot(Q_row, km) with the unquantised Q row;
  * the lazy rescale keeps a stale row max for up to 2^4 growth (the reference does too, common.cuh:14-18); this
    oracle always uses the exact running max.
"""
from __future__ import annotations

import math

import numpy as np
import torch

E4M3_MAX = 448.0
BLOCK = 128


def e4m3_round(x: np.ndarray) -> np.ndarray:
  """Round-to-nearest-even to e4m3 (fn variant: no inf, max 448) with saturation, as ``__nv_fp8_e4m3(float)``
  (``__NV_SATFINITE``) does.  Returns the rounded values as float32."""
  t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).clamp(-E4M3_MAX, E4M3_MAX)
  return t.to(torch.float8_e4m3fn).to(torch.float32).numpy()


def e4m3_bits(x: np.ndarray) -> np.ndarray:
  """The e4m3 byte codes of ``e4m3_round(x)`` (for bit-level comparison with the kernel's quantised tiles)."""
  t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).clamp(-E4M3_MAX, E4M3_MAX)
  return t.to(torch.float8_e4m3fn).view(torch.uint8).numpy()


def kv_mean(k: torch.Tensor) -> tuple[np.ndarray, np.ndarray]:
  """Sequence mean of K [B, H, N, D] per (b, h, channel): (mean rounded to k.dtype as float32, mean fp32).
  smooth_k.cuh:61-137 (fp32 accumulate, one division by Nkv)."""
  s = k.float().sum(dim=2, dtype=torch.float32) / float(k.size(2))
  return s.to(k.dtype).float().numpy(), s.numpy()


def quantize_per_block(x: np.ndarray, mean: np.ndarray | None = None, block: int = BLOCK):
  """x [B, H, N, D] float32 (already holding the bf16/fp16 values); mean [B, H, D] or None.
  Returns (x8 float32 [B, H, N, D] = the e4m3 values, scale float32 [B, H, ceil(N/block)]).
  quantize_fp8.cuh:67-168."""
  B, H, N, D = x.shape
  xs = x.astype(np.float32)
  if mean is not None:
    xs = xs - mean.astype(np.float32)[:, :, None, :]
  T = (N + block - 1) // block
  x8 = np.zeros_like(xs)
  scale = np.zeros((B, H, T), dtype=np.float32)
  for t in range(T):
    blk = xs[:, :, t * block:(t + 1) * block]
    amax = np.abs(blk).max(axis=(2, 3)).astype(np.float32)
    s = (amax / np.float32(E4M3_MAX)).astype(np.float32)
    inv = np.where(s == 0, np.float32(0), np.float32(1) / np.where(s == 0, np.float32(1), s)).astype(np.float32)
    x8[:, :, t * block:(t + 1) * block] = e4m3_round(blk * inv[:, :, None, None])
    scale[:, :, t] = s
  return x8, scale


def fp8_attention_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False,
                      scale: float | None = None, smooth_k: bool = True, mean_in_input_dtype: bool = True):
  """Quantised forward following the reference's FP8 contract (see module docstring).
  q [B, Hq, Nq, D], k / v [B, Hkv, Nkv, D] fp16/bf16 CPU tensors.  Returns (O float64 [B, Hq, Nq, D],
  LSE float64 [B, Hq, Nq], aux dict with q8 / k8 / v8 / scales / km).  Per-block Q/K/V scales, fixed P scale."""
  B, Hq, Nq, D = q.shape
  Hkv, Nkv = k.size(1), k.size(2)
  g = Hq // Hkv
  sc = float(scale) if scale is not None else 1.0 / math.sqrt(D)
  qf, kf, vf = (t.float().numpy() for t in (q, k, v))
  km_dt, km32 = kv_mean(k) if smooth_k else (None, None)
  km_q = (km_dt if mean_in_input_dtype else km32) if smooth_k else None
  q8, qs = quantize_per_block(qf)
  k8, ks = quantize_per_block(kf, km_q)
  v8, vs = quantize_per_block(vf)
  off = Nkv - Nq
  O = np.zeros((B, Hq, Nq, D), dtype=np.float64)
  LSE = np.full((B, Hq, Nq), -np.inf, dtype=np.float64)
  Tk = (Nkv + BLOCK - 1) // BLOCK
  rows = np.arange(Nq)
  for b in range(B):
    for h in range(Hq):
      hk = h // g
      Q8 = q8[b, h].astype(np.float64)
      qs_row = qs[b, h][rows // BLOCK].astype(np.float64)                  # [Nq]
      m = np.full(Nq, -np.inf)
      l = np.zeros(Nq)
      acc = np.zeros((Nq, D))
      for t in range(Tk):
        k0, k1 = t * BLOCK, min((t + 1) * BLOCK, Nkv)
        S = (Q8 @ k8[b, hk, k0:k1].astype(np.float64).T) * (qs_row[:, None] * float(ks[b, hk, t]) * sc * math.log2(math.e))
        if causal:
          S = np.where(np.arange(k0, k1)[None, :] <= (rows[:, None] + off), S, -np.inf)
        m_new = np.maximum(m, S.max(axis=1))
        safe = np.where(np.isfinite(m_new), m_new, 0.0)
        alpha = np.where(np.isfinite(m), np.exp2(m - safe), 0.0)
        P = np.exp2(S - safe[:, None])                                         # fp32 in the kernel
        P8 = e4m3_round((P * float(vs[b, hk, t]) * E4M3_MAX).astype(np.float32)).astype(np.float64)
        acc = acc * alpha[:, None] + P8 @ v8[b, hk, k0:k1].astype(np.float64)
        l = l * alpha + P.sum(axis=1)
        m = m_new
      ok = l > 0
      O[b, h][ok] = acc[ok] / E4M3_MAX / l[ok][:, None]
      lse = np.where(ok, (np.where(ok, m, 0.0) + np.log2(np.where(ok, l, 1.0))) * math.log(2.0), -np.inf)
      if smooth_k:
        # scale * qs * dot(Q8_row, km_f32)   (sm_120/split_d.cuh:753-762, smooth_k.cuh:8-16)
        lse = np.where(ok, lse + sc * qs_row * (Q8 @ km32[b, hk].astype(np.float64)), lse)
      LSE[b, h] = lse
  aux = {"q8": q8, "k8": k8, "v8": v8, "qs": qs, "ks": ks, "vs": vs, "km": km32}
  return O, LSE, aux
