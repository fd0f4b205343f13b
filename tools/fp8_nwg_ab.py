"""A/B of the FP8 softmax layout: FFPA_FP8_NWG=2 (two warpgroups, alternate tiles) vs 4 (four warpgroups, round robin),
interleaved in one process (refresh_env between arms), 8 warm-up + 20 timed launches per arm and round; bf16 kernel as context."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn

def t(fn, n=20):
  for _ in range(8): fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n

for (B, H, N, D, causal) in ((4, 32, 8192, 256, False), (1, 32, 8192, 512, False), (1, 32, 8192, 128, False), (2, 32, 8192, 256, True)):
  torch.manual_seed(0)
  q, k, v = (torch.randn(B, H, N, D, dtype=torch.bfloat16, device="cuda") * 0.5 for _ in range(3))
  f = 4.0 * B * H * D * (N * (N + 1) // 2 if causal else N * N)
  be = ffpa_attn.CUDABackend(enable_fp8=True)
  ref = None
  for rnd in range(3):
    for nwg in ("2", "4"):
      os.environ["FFPA_FP8_NWG"] = nwg
      ffpa_attn._C.refresh_env()
      out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, forward_backend=be)
      if ref is None: ref = out
      d = (out.float() - ref.float()).abs().max().item()
      ms = t(lambda: ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, forward_backend=be))
      print(f"B{B} H{H} N{N} D{D} causal={causal} round {rnd} NWG={nwg}  {ms:7.3f} ms  {f / ms * 1e-9:7.1f} TFLOP/s  maxdiff_vs_first={d:.2e}", flush=True)
  ms = t(lambda: ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal))
  print(f"B{B} H{H} N{N} D{D} causal={causal} bf16 kernel      {ms:7.3f} ms  {f / ms * 1e-9:7.1f} TFLOP/s", flush=True)
