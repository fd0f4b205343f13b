"""Secondary timings (CUDA events, device-resident inputs): forward / backward TFLOPS for the
BASELINE.json configs other than the headline one. Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ffpa-attn_b200"))
import torch  # noqa: E402

import ffpa_attn  # noqa: E402

CASES = [
  # name, B, Hq, Hkv, Nq, Nkv, D, causal
  ("c2_self_d512", 1, 32, 32, 8192, 8192, 512, False),
  ("c2_causal_d512", 1, 32, 32, 8192, 8192, 512, True),
  ("gqa_d512", 1, 32, 8, 8192, 8192, 512, False),
  ("c3_gqa_causal_n4096_d512", 1, 32, 8, 4096, 4096, 512, True),
  ("cross_nq1024_d512", 1, 32, 32, 1024, 8192, 512, False),
  ("nonaligned_n8191_h8_d512", 1, 8, 8, 8191, 8191, 512, False),
  ("self_n16384_d512", 1, 32, 32, 16384, 16384, 512, False),
  ("d320", 1, 32, 32, 8192, 8192, 320, False),
  ("d384", 1, 32, 32, 8192, 8192, 384, False),
  ("d256", 1, 32, 32, 8192, 8192, 256, False),
  ("d128", 1, 32, 32, 8192, 8192, 128, False),
  ("d768", 1, 32, 32, 8192, 8192, 768, False),
  ("d1024", 1, 32, 32, 8192, 8192, 1024, False),
  ("c4_b4_d256", 4, 32, 32, 8192, 8192, 256, False),
  ("decode_nq1_n8192_d512", 1, 32, 32, 1, 8192, 512, False),
  ("decode_nq1_b8_gqa_n16384_d512", 8, 32, 8, 1, 16384, 512, False),
]


def flops(B, Hq, Nq, Nkv, D, causal):
  pairs = (Nq * (Nkv - Nq) + Nq * (Nq + 1) // 2) if causal else Nq * Nkv
  return 4.0 * B * Hq * D * pairs


def timeit(fn, iters):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(iters):
    fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / iters


def main():
  only = sys.argv[1:] or None
  dev = "cuda"
  for name, B, Hq, Hkv, Nq, Nkv, D, causal in CASES:
    if only and name not in only:
      continue
    torch.manual_seed(42)
    q = torch.randn(B, Hq, Nq, D, dtype=torch.bfloat16, device=dev)
    k = torch.randn(B, Hkv, Nkv, D, dtype=torch.bfloat16, device=dev)
    v = torch.randn(B, Hkv, Nkv, D, dtype=torch.bfloat16, device=dev)
    kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)
    f = flops(B, Hq, Nq, Nkv, D, causal)
    ms_f = timeit(lambda: ffpa_attn.ffpa_attn_func(q, k, v, **kw), 10)
    rec = {"case": name, "fwd_ms": ms_f, "fwd_tflops": f / ms_f * 1e-9}
    if D > 768:
      os.environ["FFPA_FWD_REPLAY"] = "0"   # A/B: second O slab as a full second softmax pass
      ms_2 = timeit(lambda: ffpa_attn.ffpa_attn_func(q, k, v, **kw), 10)
      del os.environ["FFPA_FWD_REPLAY"]
      rec.update({"fwd_two_pass_ms": ms_2, "fwd_two_pass_tflops": f / ms_2 * 1e-9})
    if name.startswith("decode"):
      kv_bytes = 2 * k.numel() * 2  # K and V are each read once: the HBM roofline of decode
      rec.update({"kv_gbs": kv_bytes / ms_f * 1e-6, "hbm_peak_gbs": 6580.9})
    if not name.startswith("decode"):
      qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
      out = ffpa_attn.ffpa_attn_func(qg, kg, vg, **kw)
      d_o = torch.randn_like(out)

      def bwd():
        out.backward(d_o, retain_graph=True)

      ms_b = timeit(bwd, 5)
      rec.update({"bwd_ms": ms_b, "bwd_tflops": 2.5 * f / ms_b * 1e-9})
      if 384 <= D <= 1024:
        # A/B: the same backward with the minimum workspace (three recompute kernels, O(N) memory)
        from ffpa_attn import _C
        from ffpa_attn.cuda import _ffpa_attn_forward_cuda

        o2, lse2 = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), D ** -0.5)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ms_r = timeit(lambda: _C.ffpa_attn_backward_ex(q, k, v, o2, lse2, d_o, dq, dk, dv, 0, int(causal), D ** -0.5,
                                                       min_workspace=True), 5)
        ms_s = timeit(lambda: _C.ffpa_attn_backward(q, k, v, o2, lse2, d_o, dq, dk, dv, 0, int(causal), D ** -0.5), 5)
        rec.update({"bwd_recompute_ms": ms_r, "bwd_recompute_tflops": 2.5 * f / ms_r * 1e-9,
                    "bwd_stash_ms": ms_s, "bwd_stash_tflops": 2.5 * f / ms_s * 1e-9})
    if name.startswith("c4") or name in ("c2_self_d512", "c2_causal_d512", "d320"):
      be = ffpa_attn.CUDABackend(enable_fp8=True)
      ms8 = timeit(lambda: ffpa_attn.ffpa_attn_func(q, k, v, forward_backend=be, **kw), 10)
      rec.update({"fp8_fwd_ms": ms8, "fp8_fwd_tflops": f / ms8 * 1e-9, "fp8_includes": "quantise pre-pass + attention"})
    print(json.dumps(rec), flush=True)
  if only is None or "varlen" in only:
    varlen_case(dev)


def varlen_case(dev):
  """Packed variable-length batch (one launch set): 24 sequences of 256..4096 tokens, H=16, D=512, causal."""
  import random

  random.seed(0)
  lens = [random.choice([256, 384, 512, 1024, 1536, 2048, 4096]) for _ in range(24)]
  cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
  T, H, D = int(sum(lens)), 16, 512
  torch.manual_seed(42)
  q, k, v = (torch.randn(T, H, D, dtype=torch.bfloat16, device=dev) for _ in range(3))
  f = sum(4.0 * H * D * (n * (n + 1) // 2) for n in lens)
  fn = lambda: ffpa_attn.ffpa_attn_varlen_func(q, k, v, cu, cu, max(lens), max(lens), causal=True)  # noqa: E731
  ms_f = timeit(fn, 10)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_varlen_func(qg, kg, vg, cu, cu, max(lens), max(lens), causal=True)
  d_o = torch.randn_like(out)
  ms_b = timeit(lambda: out.backward(d_o, retain_graph=True), 5)
  print(json.dumps({"case": "varlen_24seq_h16_d512_causal", "tokens": T, "fwd_ms": ms_f, "fwd_tflops": f / ms_f * 1e-9,
                    "bwd_ms": ms_b, "bwd_tflops": 2.5 * f / ms_b * 1e-9}), flush=True)


if __name__ == "__main__":
  main()
