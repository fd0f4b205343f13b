"""Per-item overhead of the forward kernel: non-causal D=512 bf16 at equal FLOPs, sequence length varied, so the number of
KV tiles per work item T changes while the item count x T stays constant. time ~ items * (T + x) * t_tile."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn

def tmin(fn, n=30):
  for _ in range(8): fn()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
  ev[0].record()
  for i in range(n):
    fn(); ev[i + 1].record()
  torch.cuda.synchronize()
  ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
  return ts[0], ts[n // 2]

D = int(sys.argv[1]) if len(sys.argv) > 1 else 512
rows = []
for B, H, N in ((16, 32, 512), (8, 32, 1024), (4, 32, 2048), (2, 32, 4096), (1, 32, 8192), (1, 16, 16384), (1, 8, 32768)):
  torch.manual_seed(0)
  q, k, v = (torch.randn(B, H, N, D, dtype=torch.bfloat16, device="cuda") for _ in range(3))
  f = 4.0 * B * H * D * N * N
  mn, md = tmin(lambda: ffpa_attn.ffpa_attn_func(q, k, v))
  items, T = B * H * (N // 128), N // 128
  rows.append((T, items, mn))
  print(f"B{B} H{H} N{N}: T={T:4d} items={items:5d}  min {mn:7.3f} ms ({f / mn * 1e-9:6.0f} TF)  median {md:7.3f} ms ({f / md * 1e-9:6.0f} TF)   us/item/cluster-slot {mn * 1e3 / (items / 74):7.2f}", flush=True)
  del q, k, v
# least squares: mn = a * items * T + b * items  ->  x = b / a tiles
import numpy as np
A = np.array([[it * T, it] for T, it, _ in rows], dtype=float); y = np.array([m for _, _, m in rows])
(a, b), *_ = np.linalg.lstsq(A, y, rcond=None)
print(f"fit: {a * 74 * 1e3:.3f} us per KV tile and cluster, per-item overhead = {b / a:.2f} KV tiles ({b * 74 * 1e3:.2f} us)")
