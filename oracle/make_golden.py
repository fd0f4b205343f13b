"""Generate tests/golden/*.npz by running the REFERENCE package on the CPU.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The reference's CPU-runnable route is ``ffpa_attn_func(..., backend="sdpa")``
(/root/reference/src/ffpa_attn/ffpa_attn_interface.py:163-176 -> aten SDPA); BASELINE.json config 1
(B=1, H=2, N=512, D=320, bf16) is exactly that call.  Inputs follow the reference tests' recipe
(``torch.manual_seed(0)`` + ``randn``, /root/reference/tests/test_ffpa_fwd.py:116-121).  Inputs and
outputs are stored as raw uint16/float32 arrays so the fixtures do not depend on the torch RNG.
Backward fixtures come from autograd through the same reference call with ``loss = (out*dO).sum()``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
  # name, B, Hq, Hkv, Nq, Nkv, D, dtype, causal, mask_kind, with_bwd
  ("c1_self_b1h2n512d320_bf16", 1, 2, 2, 512, 512, 320, "bf16", False, None, False),
  ("self_b1h2n128d512_bf16_bwd", 1, 2, 2, 128, 128, 512, "bf16", False, None, True),
  ("causal_b1h2n192d128_f16_bwd", 1, 2, 2, 192, 192, 128, "f16", True, None, True),
  ("gqa_b2h4kv2_nq130_nkv257_d64_bf16_bwd", 2, 4, 2, 130, 257, 64, "bf16", False, None, True),
  ("boolmask_b1h2n128d64_f16", 1, 2, 2, 128, 160, 64, "f16", False, "bool", False),
  ("addmask_b2h2n96d128_bf16_bwd", 2, 2, 2, 96, 200, 128, "bf16", False, "add", True),
]


def _raw16(t: torch.Tensor) -> np.ndarray:
  return t.detach().contiguous().view(torch.int16).numpy().view(np.uint16)


def main() -> None:
  sys.path.insert(0, REF_SRC)
  import ffpa_attn as ref  # the reference package itself

  os.makedirs(OUT_DIR, exist_ok=True)
  for name, B, Hq, Hkv, Nq, Nkv, D, dt, causal, mask_kind, with_bwd in CASES:
    dtype = torch.bfloat16 if dt == "bf16" else torch.float16
    torch.manual_seed(0)
    q = torch.randn(B, Hq, Nq, D, dtype=dtype)
    k = torch.randn(B, Hkv, Nkv, D, dtype=dtype)
    v = torch.randn(B, Hkv, Nkv, D, dtype=dtype)
    d_o = torch.randn(B, Hq, Nq, D, dtype=dtype)
    mask = None
    if mask_kind == "bool":
      mask = torch.rand(1, 1, Nq, Nkv) > 0.3
      mask[..., 0] = True  # every row keeps at least one key
    elif mask_kind == "add":
      mask = (torch.randn(B, 1, Nq, Nkv) * 0.5).to(torch.float32)
    # the reference's non-square causal is bottom-right aligned; aten SDPA's is top-left, so the
    # reference tests build an explicit mask for that case (tests/test_ffpa_fwd.py:99-103). All
    # causal fixtures here are square, where both agree.
    assert not (causal and Nq != Nkv)
    save = {"q": _raw16(q), "k": _raw16(k), "v": _raw16(v), "d_o": _raw16(d_o),
            "meta": np.array([B, Hq, Hkv, Nq, Nkv, D, int(dt == "bf16"), int(causal)], dtype=np.int64)}
    if mask is not None:
      save["mask"] = mask.numpy()
    if with_bwd:
      # fp32 leaves so the fixture's gradients carry no bf16 rounding of the reference's own
      q32, k32, v32 = (t.float().requires_grad_(True) for t in (q, k, v))
      out = ref.ffpa_attn_func(q32, k32, v32, attn_mask=mask, is_causal=causal, enable_gqa=Hq != Hkv,
                               backend="sdpa")
      (out * d_o.float()).sum().backward()
      save["o_f32"] = out.detach().numpy()
      save["dq"] = q32.grad.numpy()
      save["dk"] = k32.grad.numpy()
      save["dv"] = v32.grad.numpy()
    out16 = ref.ffpa_attn_func(q, k, v, attn_mask=mask, is_causal=causal, enable_gqa=Hq != Hkv,
                               backend="sdpa")
    save["o"] = _raw16(out16)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, **save)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
  main()
