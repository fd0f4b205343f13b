"""A/B timing inside one process: bf16 vs FP8 (smooth-K off/on) at BASELINE config 4."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn

def t(fn, n=10):
  for _ in range(3): fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize()
  return a.elapsed_time(b) / n

B, H, N, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (4, 32, 8192, 256)))
torch.manual_seed(0)
q, k, v = (torch.randn(B, H, N, D, dtype=torch.bfloat16, device="cuda") * 0.5 for _ in range(3))
f = 4.0 * B * H * D * N * N
on = ffpa_attn.CUDABackend(enable_fp8=True, fp8_smooth_k=True)
off = ffpa_attn.CUDABackend(enable_fp8=True, fp8_smooth_k=False)
for rnd in range(2):
  for name, fn in (("bf16", lambda: ffpa_attn.ffpa_attn_func(q, k, v)),
                   ("fp8 smooth-K off", lambda: ffpa_attn.ffpa_attn_func(q, k, v, forward_backend=off)),
                   ("fp8 smooth-K on", lambda: ffpa_attn.ffpa_attn_func(q, k, v, forward_backend=on))):
    ms = t(fn)
    print(f"round {rnd} {name:18s} {ms:7.3f} ms  {f / ms * 1e-9:7.1f} TFLOP/s", flush=True)
