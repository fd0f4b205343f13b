#!/usr/bin/env python
"""Does the reference's own GPU code run on this box?  Imports the UNMODIFIED reference package from
baseline/_ref (pip-installed by tools/install_reference.sh, git-ignored) and tries its CuTe-DSL tcgen05
D=512 kernels and its Triton backend on the BASELINE shapes, both timing protocols (CUDA events and
the reference's wall-clock loop, /root/reference/src/ffpa_attn/cli/_runner_fwd.py:84-103).
Writes gpurun_out/probe_reference_gpu.json; every failure is recorded as text, never swallowed."""
from __future__ import annotations

import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import torch  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "probe_reference_gpu.json")
res: dict = {"torch": torch.__version__, "cases": {}}


def flops(B, H, Nq, Nkv, D, causal):
  pairs = (Nq * (Nkv - Nq) + Nq * (Nq + 1) // 2) if causal else Nq * Nkv
  return 4.0 * B * H * D * pairs


def time_events(fn, warm, iters):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(iters):
    fn()
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / iters


def time_wall(fn, warm=2, iters=10):
  for _ in range(warm):
    fn()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(iters):
    fn()
  torch.cuda.synchronize()
  return (time.perf_counter() - t0) / iters * 1e3


def main():
  try:
    import ffpa_attn
    res["ffpa_attn_file"] = ffpa_attn.__file__
  except Exception:
    res["import_error"] = traceback.format_exc()
    return
  for mod in ("cutlass", "quack", "triton"):
    try:
      m = __import__(mod)
      res[mod] = getattr(m, "__version__", "?")
    except Exception as e:  # noqa: BLE001
      res[mod] = f"import failed: {e}"
  dev = torch.device("cuda:0")
  dt = torch.bfloat16
  cases = [
    # name, backend, B,H,Hkv,N,D,causal, do_bwd
    ("c2_cutedsl", "cutedsl", 1, 32, 32, 8192, 512, False, True),
    ("c3_cutedsl", "cutedsl", 1, 32, 8, 4096, 512, True, True),
    ("c2_causal_cutedsl", "cutedsl", 1, 32, 32, 8192, 512, True, False),
    ("c2_triton", "triton", 1, 32, 32, 8192, 512, False, True),
    ("d320_triton", "triton", 1, 32, 32, 8192, 320, False, False),
    ("d768_triton", "triton", 1, 32, 32, 8192, 768, False, False),
    ("d1024_triton", "triton", 1, 32, 32, 8192, 1024, False, False),
    ("d320_cutedsl", "cutedsl", 1, 32, 32, 8192, 320, False, False),
    ("d768_cutedsl", "cutedsl", 1, 32, 32, 8192, 768, False, False),
  ]
  only = os.environ.get("PROBE_ONLY")
  for name, backend, B, H, Hkv, N, D, causal, do_bwd in cases:
    if only and only not in name:
      continue
    c: dict = {"backend": backend, "shape": [B, H, Hkv, N, D], "causal": causal}
    res["cases"][name] = c
    try:
      torch.manual_seed(0)
      q = torch.randn(B, H, N, D, dtype=dt, device=dev, requires_grad=do_bwd)
      k = torch.randn(B, Hkv, N, D, dtype=dt, device=dev, requires_grad=do_bwd)
      v = torch.randn(B, Hkv, N, D, dtype=dt, device=dev, requires_grad=do_bwd)
      kw = dict(is_causal=causal, enable_gqa=H != Hkv, backend=backend)
      t0 = time.perf_counter()
      with torch.no_grad():
        o = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
      torch.cuda.synchronize()
      c["first_call_s"] = time.perf_counter() - t0
      # correctness on one head vs fp32 SDPA
      g = H // Hkv
      ref = torch.nn.functional.scaled_dot_product_attention(
        q[:, :1].detach().float(), k[:, :1].detach().float(), v[:, :1].detach().float(), is_causal=causal)
      c["max_abs_err_head0"] = float((o[:, :1].float() - ref).abs().max())
      f = flops(B, H, N, N, D, causal)
      with torch.no_grad():
        fwd = lambda: ffpa_attn.ffpa_attn_func(q, k, v, **kw)  # noqa: E731
        ms = time_events(fwd, 3, 10)
        c["fwd_ms_events"] = ms
        c["fwd_tflops_events"] = f / ms * 1e-9
        ms = time_wall(fwd)
        c["fwd_ms_wall"] = ms
        c["fwd_tflops_wall"] = f / ms * 1e-9
      if do_bwd:
        t0 = time.perf_counter()
        o = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
        do = torch.randn_like(o)
        o.backward(do, retain_graph=True)
        torch.cuda.synchronize()
        c["first_bwd_s"] = time.perf_counter() - t0
        bwd = lambda: o.backward(do, retain_graph=True)  # noqa: E731
        ms = time_events(bwd, 2, 5)
        c["bwd_ms_events"] = ms
        c["bwd_tflops_events"] = 2.5 * f / ms * 1e-9
      del q, k, v, o
      torch.cuda.empty_cache()
    except Exception:
      c["error"] = traceback.format_exc()[-3000:]
    json.dump(res, open(OUT, "w"), indent=1)


if __name__ == "__main__":
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  try:
    main()
  finally:
    json.dump(res, open(OUT, "w"), indent=1)
    print(json.dumps(res, indent=1)[:6000])
