"""Tiny driver for ncu: forward at D=1024 (B=1 H=8 N=8192 bf16), replay path then two-pass path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch, ffpa_attn
torch.manual_seed(0)
q, k, v = (torch.randn(1, 8, 8192, 1024, dtype=torch.bfloat16, device="cuda") for _ in range(3))
for _ in range(2):
  ffpa_attn.ffpa_attn_func(q, k, v)
os.environ["FFPA_FWD_REPLAY"] = "0"
for _ in range(2):
  ffpa_attn.ffpa_attn_func(q, k, v)
torch.cuda.synchronize()
