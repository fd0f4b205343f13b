#!/usr/bin/env python
"""Backward time vs scratch budget: C2 / C3 backward with the binder's stash cap (FFPA_BWD_STASH_MAX_GB) swept from
0 (O(N) recompute kernels) upwards. Shows what the score stash buys per byte; prints one JSON line per point."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch  # noqa: E402

import ffpa_attn  # noqa: E402
from ffpa_attn import _C  # noqa: E402
from ffpa_attn.cuda import _ffpa_attn_forward_cuda  # noqa: E402


def timeit(fn, n=5, warm=2):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / n


for name, (B, Hq, Hkv, N, D, causal) in {"c2": (1, 32, 32, 8192, 512, False), "c3": (1, 32, 8, 4096, 512, True),
                                           "b4_c2": (4, 32, 32, 8192, 512, False)}.items():
  torch.manual_seed(0)
  q = torch.randn(B, Hq, N, D, dtype=torch.bfloat16, device="cuda")
  k = torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device="cuda")
  v = torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device="cuda")
  d_o = torch.randn_like(q)
  sc = D ** -0.5
  o, lse = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), sc)
  dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
  pairs = N * (N + 1) // 2 if causal else N * N
  fl = 2.5 * 4.0 * B * Hq * D * pairs
  for cap in ("0", "0.3", "0.6", "1.2", "2.4", "4.8", "9.6", "1000"):
    os.environ["FFPA_BWD_STASH_MAX_GB"] = cap
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    n0 = _C.launch_count()
    _C.ffpa_attn_backward(q, k, v, o, lse, d_o, dq, dk, dv, 0, int(causal), sc)
    torch.cuda.synchronize()
    launches = _C.launch_count() - n0
    ws = torch.cuda.max_memory_allocated() - base
    ms = timeit(lambda: _C.ffpa_attn_backward(q, k, v, o, lse, d_o, dq, dk, dv, 0, int(causal), sc))
    print(json.dumps({"case": name, "cap_gb": float(cap), "scratch_gb": round(ws / 2 ** 30, 3), "launches": int(launches),
                      "ms": round(ms, 3), "tflops": round(fl / ms * 1e-9, 1)}), flush=True)
  del q, k, v, d_o, o, lse, dq, dk, dv
  torch.cuda.empty_cache()
