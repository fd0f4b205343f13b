"""One arm of a two-library A/B of the FP8 forward: FFPA_AB_PKG names the directory holding the ``ffpa_attn`` package to load
(default: this repo's). Run alternately from a shell loop on ONE box; prints per-case median / min of per-launch CUDA-event times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.environ.get("FFPA_AB_PKG", os.path.join(ROOT, "ffpa-attn_b200")))
import torch, ffpa_attn

def t(fn, n=30):
  for _ in range(10): fn()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
  ev[0].record()
  for i in range(n):
    fn(); ev[i + 1].record()
  torch.cuda.synchronize()
  ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
  return ts[n // 2], ts[0]

tag = os.environ.get("FFPA_AB_TAG", "new")
for (B, H, N, D, causal) in ((4, 32, 8192, 256, False), (1, 32, 8192, 512, False), (1, 32, 8192, 128, False), (2, 32, 8192, 256, True)):
  torch.manual_seed(0)
  q, k, v = (torch.randn(B, H, N, D, dtype=torch.bfloat16, device="cuda") * 0.5 for _ in range(3))
  f = 4.0 * B * H * D * (N * (N + 1) // 2 if causal else N * N)
  be = ffpa_attn.CUDABackend(enable_fp8=True)
  o = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, forward_backend=be)
  med, mn = t(lambda: ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, forward_backend=be))
  print(f"{tag} B{B} N{N} D{D} causal={int(causal)}  median {med:7.3f} ms {f / med * 1e-9:7.1f} TFLOP/s   min {mn:7.3f} ms {f / mn * 1e-9:7.1f} TFLOP/s"
        f"   checksum {o.float().abs().sum().item():.6e}", flush=True)
