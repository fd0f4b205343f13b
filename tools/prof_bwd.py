"""Tiny driver for ncu: one forward + backward at BASELINE config 3 (GQA causal N=4096 D=512)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn
torch.manual_seed(0)
q = torch.randn(1, 32, 4096, 512, dtype=torch.bfloat16, device="cuda", requires_grad=True)
k = torch.randn(1, 8, 4096, 512, dtype=torch.bfloat16, device="cuda", requires_grad=True)
v = torch.randn(1, 8, 4096, 512, dtype=torch.bfloat16, device="cuda", requires_grad=True)
for _ in range(2):
  out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=True, enable_gqa=True)
  out.backward(torch.randn_like(out))
torch.cuda.synchronize()
