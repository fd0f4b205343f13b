"""Tiny driver for ncu: two backward calls at BASELINE config 2's shape (B=1 H=32 N=8192 D=512 bf16, non-causal)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn
torch.manual_seed(0)
H = int(os.environ.get("PROF_H", "32"))
q, k, v = (torch.randn(1, H, 8192, 512, dtype=torch.bfloat16, device="cuda", requires_grad=True) for _ in range(3))
out = ffpa_attn.ffpa_attn_func(q, k, v)
d_o = torch.randn_like(out)
for _ in range(2):
  out.backward(d_o, retain_graph=True)
torch.cuda.synchronize()
