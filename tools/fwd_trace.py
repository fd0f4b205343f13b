"""Development aid: run one forward with a library built with -DFFPA_TRACE (see ffpa_fwd_sm100.cuh) and dump the per-cluster,
per-item clock64 stamps. Usage: FFPA_AB_PKG=<dir with the traced ffpa_attn package> python tools/fwd_trace.py B Hq Hkv N D causal out.npy"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = os.environ.get("FFPA_AB_PKG", os.path.join(ROOT, "ffpa-attn_b200"))
sys.path.insert(0, pkg)
import numpy as np, torch, ffpa_attn
B, Hq, Hkv, N, D, causal = (int(x) for x in sys.argv[1:7])
lib = ctypes.CDLL(os.path.join(pkg, "ffpa_attn", "libffpa_b200.so"))
torch.manual_seed(0)
q = torch.randn(B, Hq, N, D, dtype=torch.bfloat16, device="cuda")
k, v = (torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device="cuda") for _ in range(2))
kw = dict(is_causal=bool(causal), enable_gqa=Hq != Hkv)
for _ in range(5): ffpa_attn.ffpa_attn_func(q, k, v, **kw)
buf = torch.zeros(74 * 64 * 16, dtype=torch.int64, device="cuda")
lib.ffpa_dbg_set_fwd_trace.argtypes = [ctypes.c_void_p]
lib.ffpa_dbg_set_fwd_trace(buf.data_ptr())
torch.cuda.synchronize()
ffpa_attn.ffpa_attn_func(q, k, v, **kw)
torch.cuda.synchronize()
np.save(sys.argv[7], buf.cpu().numpy().reshape(74, 64, 16))
print("saved", sys.argv[7])
