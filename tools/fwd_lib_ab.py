"""One arm of a two-library A/B of the 16-bit forward: FFPA_AB_PKG names the directory holding the ``ffpa_attn`` package to
load (default: this repo's). Run alternately from a shell loop on ONE box; prints median / min of per-launch CUDA-event times
and a checksum of the output so that arms can be compared for equality."""
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
sel = os.environ.get("FFPA_AB_CASES")   # e.g. "0,1,5": indices into the case list
for ci, (B, Hq, Hkv, N, D, causal) in enumerate(((1, 32, 32, 8192, 512, False), (1, 32, 8, 4096, 512, True), (1, 32, 32, 8192, 320, False), (1, 32, 32, 8192, 1024, False),
                                   (1, 32, 32, 8192, 256, False), (8, 32, 32, 1024, 512, True))):
  if sel and str(ci) not in sel.split(","): continue
  torch.manual_seed(0)
  q = torch.randn(B, Hq, N, D, dtype=torch.bfloat16, device="cuda")
  k, v = (torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device="cuda") for _ in range(2))
  f = 4.0 * B * Hq * D * (N * (N + 1) // 2 if causal else N * N)
  kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)
  o = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
  med, mn = t(lambda: ffpa_attn.ffpa_attn_func(q, k, v, **kw))
  print(f"{tag} B{B} H{Hq}/{Hkv} N{N} D{D} causal={int(causal)}  median {med:7.3f} ms {f / med * 1e-9:7.1f} TFLOP/s   min {mn:7.3f} ms {f / mn * 1e-9:7.1f} TFLOP/s"
        f"   checksum {o.float().abs().sum().item():.6e}", flush=True)
