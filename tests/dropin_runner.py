"""Runs INSIDE a subprocess whose sys.path starts with a temp copy of the UNMODIFIED reference package
(baseline/_ref/ffpa_attn) into which this repo's built ``_C*.so`` + ``libffpa_b200.so`` were dropped
(tests/test_dropin_gpu.py builds that directory). Exercises the reference's own public API with
``forward_backend="cuda"`` and prints one JSON dict of measurements; the test asserts on it.
Oracle = oracle/attention_oracle.py (fp64 numpy), never the product."""
import json
import os
import sys
import traceback

import numpy as np
import torch

import ffpa_attn
from ffpa_attn.functional import CUDABackend

sys.path.insert(0, os.environ["FFPA_REPO_ROOT"])
from oracle import attention_oracle as orc  # noqa: E402

DEV = "cuda"
res = {"package_file": ffpa_attn.__file__}


def mk(B, Hq, Hkv, Nq, Nkv, D, dtype=torch.bfloat16, seed=0, amp=1.0):
  g = torch.Generator().manual_seed(seed)
  q = (torch.randn(B, Hq, Nq, D, generator=g) * amp).to(dtype).to(DEV)
  k = (torch.randn(B, Hkv, Nkv, D, generator=g) * amp).to(dtype).to(DEV)
  v = (torch.randn(B, Hkv, Nkv, D, generator=g) * amp).to(dtype).to(DEV)
  return q, k, v


def err(out, ref):
  return float(np.abs(out.float().cpu().numpy() - ref).max())


def case(name):
  def deco(fn):
    try:
      n0 = C.launch_count()
      r = fn() or {}
      torch.cuda.synchronize()
      r["launches"] = int(C.launch_count() - n0)
      res[name] = r
    except Exception:
      res[name] = {"error": traceback.format_exc()[-2500:]}
    return fn
  return deco


import ffpa_attn.cuda as cuda_mod  # noqa: E402
from ffpa_attn import _C as C  # noqa: E402

res["native_module"] = C.__file__
res["cuda_fwd_available"] = bool(cuda_mod.CUDA_FWD_AVAILABLE)
res["import_error"] = repr(cuda_mod._CUDA_IMPORT_ERROR)


@case("bf16_plain_d512")
def _():
  q, k, v = mk(1, 4, 4, 1024, 1024, 512)
  out = ffpa_attn.ffpa_attn_func(q, k, v, forward_backend="cuda")
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  return {"err": err(out, ref)}


@case("fp16_causal_gqa_d320")
def _():
  q, k, v = mk(2, 4, 2, 640, 900, 320, torch.float16, seed=1)
  out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=True, enable_gqa=True, forward_backend="cuda")
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=True)
  return {"err": err(out, ref)}


@case("bool_mask_d512")
def _():
  q, k, v = mk(1, 2, 2, 600, 700, 512, seed=2)
  m = torch.rand(1, 1, 600, 700, generator=torch.Generator().manual_seed(3)) > 0.3
  m[..., 0] = True
  out = ffpa_attn.ffpa_attn_func(q, k, v, attn_mask=m.to(DEV), forward_backend="cuda")
  bias = np.where(m.numpy(), 0.0, -np.inf)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=bias)
  return {"err": err(out, ref)}


@case("key_padding_additive_d512")
def _():
  q, k, v = mk(2, 2, 2, 512, 640, 512, seed=4)
  b = (torch.randn(2, 1, 1, 640, generator=torch.Generator().manual_seed(5)) * 2).float()
  out = ffpa_attn.ffpa_attn_func(q, k, v, attn_mask=b.to(DEV), forward_backend="cuda")
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=b.double().numpy())
  return {"err": err(out, ref)}


@case("dropout_d512")
def _():
  q, k, v = mk(1, 2, 2, 512, 512, 512, seed=6)
  torch.cuda.manual_seed(123)
  seed, offset = int(torch.cuda.initial_seed()), int(torch.cuda._get_rng_state_offset())
  out = ffpa_attn.ffpa_attn_func(q, k, v, dropout_p=0.25, forward_backend="cuda")
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), dropout_p=0.25, philox_seed=seed, philox_offset=offset)
  return {"err": err(out, ref), "rng_advanced": int(torch.cuda._get_rng_state_offset()) - offset}


@case("fp8_d256_small_d_env")
def _():
  # BASELINE config 4 reaches the CUDA backend only with FFPA_CUDA_ALLOW_SMALL_D=1 (functional.py:99-105,717-724)
  q, k, v = mk(1, 4, 4, 1024, 1024, 256, seed=7, amp=0.5)
  be = CUDABackend(forward=True, backward=False, enable_fp8=True)
  out = ffpa_attn.ffpa_attn_func(q, k, v, forward_backend=be)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  o16 = ffpa_attn.ffpa_attn_func(q, k, v, forward_backend="cuda")
  return {"err": err(out, ref), "differs_from_16bit": bool((out != o16).any())}


@case("fp8_causal_hybrid_d512")
def _():
  q, k, v = mk(1, 2, 2, 1024, 1024, 512, seed=8, amp=0.5)
  be = CUDABackend(forward=True, backward=False, enable_fp8=True)   # fp8_hybrid=None -> auto on (causal)
  out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=True, forward_backend=be)
  o16 = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=True, forward_backend="cuda")
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=True)
  return {"err": err(out, ref), "early_rows_max_diff_vs_16bit": float((out[:, :, :256].float() - o16[:, :, :256].float()).abs().max()),
          "late_rows_differ": bool((out[:, :, 256:] != o16[:, :, 256:]).any()),
          "early_err": err(out[:, :, :256], ref[:, :, :256]), "o16_early_err": err(o16[:, :, :256], ref[:, :, :256])}


@case("fp8_unsupported_knob_raises")
def _():
  q, k, v = mk(1, 2, 2, 512, 512, 512, seed=9)
  be = CUDABackend(forward=True, backward=False, enable_fp8=True, fp8_qk_mm_type="int8")
  try:
    ffpa_attn.ffpa_attn_func(q, k, v, forward_backend=be)
    return {"raised": None}
  except Exception as e:  # noqa: BLE001
    return {"raised": type(e).__name__, "msg": str(e)[:200]}


def grads(fn_kwargs, q, k, v, d_o):
  q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(q, k, v, **fn_kwargs)
  out.backward(d_o)
  return out.detach(), q.grad, k.grad, v.grad


def grad_errs(q, k, v, d_o, got, causal):
  wq, wk, wv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), causal=causal)
  return {n: err(g, w) / max(1.0, float(np.abs(w).max())) for n, g, w in (("dq", got[1], wq), ("dk", got[2], wk), ("dv", got[3], wv))}


@case("our_forward_plus_reference_sdpa_backward_d512")
def _():
  # saved-tensor contract (functional.py:1066-1077): O in q.dtype, LSE fp32 natural log [B,Hq,Nq]
  q, k, v = mk(1, 2, 2, 512, 512, 512, seed=10)
  d_o = torch.randn_like(q)
  got = grads(dict(forward_backend="cuda", backward_backend="sdpa"), q, k, v, d_o)
  return grad_errs(q, k, v, d_o, got, False)


@case("our_forward_plus_reference_triton_backward_d320")
def _():
  q, k, v = mk(1, 2, 2, 512, 512, 320, seed=11)
  d_o = torch.randn_like(q)
  got = grads(dict(is_causal=True, forward_backend="cuda"), q, k, v, d_o)   # default backward = Triton
  return grad_errs(q, k, v, d_o, got, True)


# ---- parity against the reference's OWN GPU backends on identical inputs, in this very process -----------------
def _ref_vs_ours(backend, B, Hq, Hkv, Nq, Nkv, D, causal, dtype=torch.bfloat16, mask=False, dropout=0.0, bwd=False, seed=20):
  q, k, v = mk(B, Hq, Hkv, Nq, Nkv, D, dtype, seed=seed)
  kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)
  if mask:
    kw["attn_mask"] = (torch.randn(B, 1, Nq, Nkv, generator=torch.Generator().manual_seed(seed + 1)) * 2).to(dtype).to(DEV)
  if dropout:
    kw["dropout_p"] = dropout
  out = {}
  d_o = torch.randn_like(q)
  torch.manual_seed(5)
  if bwd:
    r_o, r_dq, r_dk, r_dv = grads(dict(backend=backend, **kw), q, k, v, d_o)
  else:
    with torch.no_grad():
      r_o = ffpa_attn.ffpa_attn_func(q, k, v, backend=backend, **kw)
  torch.manual_seed(5)   # same Philox reservation for both backends (functional.py:518-540)
  n0 = C.launch_count()
  with torch.no_grad():
    o = ffpa_attn.ffpa_attn_func(q, k, v, forward_backend="cuda", **kw)
  out["launches"] = int(C.launch_count() - n0)
  out["o_err"] = float((o.float() - r_o.float()).abs().max())
  out["o_cos"] = float(torch.nn.functional.cosine_similarity(o.float().flatten(), r_o.float().flatten(), dim=0))
  if bwd:
    # our native backward (the symbol the reference leaves as a thrower), called on our forward's O / LSE
    from ffpa_attn.cuda import _ffpa_attn_forward_cuda
    sc = 1.0 / (D ** 0.5)
    cuda_mod.set_cuda_backend_impl(cuda_mod.CudaBackendImpl.NATIVE)
    o2, lse = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), sc)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    C.ffpa_attn_backward(q, k, v, o2, lse.contiguous(), d_o, dq, dk, dv, 0, int(causal), sc)
    for n, a, b_ in (("dq", dq, r_dq), ("dk", dk, r_dk), ("dv", dv, r_dv)):
      out[n + "_rel"] = float((a.float() - b_.float()).abs().max() / (b_.float().abs().max() + 1e-30))
      out[n + "_cos"] = float(torch.nn.functional.cosine_similarity(a.float().flatten(), b_.float().flatten(), dim=0))
  return out


for _name, _args in (
    ("vs_triton_d320_fwd_bwd", dict(backend="triton", B=1, Hq=4, Hkv=4, Nq=1024, Nkv=1024, D=320, causal=False, bwd=True)),
    ("vs_triton_d320_causal_gqa_fwd_bwd", dict(backend="triton", B=2, Hq=4, Hkv=2, Nq=768, Nkv=768, D=320, causal=True, bwd=True)),
    ("vs_triton_d320_mask", dict(backend="triton", B=1, Hq=2, Hkv=2, Nq=512, Nkv=640, D=320, causal=False, mask=True)),
    ("vs_triton_d320_dropout", dict(backend="triton", B=1, Hq=2, Hkv=2, Nq=512, Nkv=512, D=320, causal=False, dropout=0.2, dtype=torch.float16)),
    ("vs_triton_d512_fwd", dict(backend="triton", B=1, Hq=2, Hkv=2, Nq=1024, Nkv=1024, D=512, causal=False)),
    ("vs_cutedsl_d320_fwd_bwd", dict(backend="cutedsl", B=1, Hq=4, Hkv=4, Nq=1024, Nkv=1024, D=320, causal=True, bwd=True)),
    ("vs_cutedsl_d768_fwd", dict(backend="cutedsl", B=1, Hq=2, Hkv=2, Nq=1024, Nkv=1024, D=768, causal=False)),
    ("vs_cutedsl_d512_sm100_fwd_bwd", dict(backend="cutedsl", B=1, Hq=2, Hkv=2, Nq=1024, Nkv=1024, D=512, causal=True, bwd=True)),
):
  case(_name)(lambda _a=_args: _ref_vs_ours(**_a))

_out_dir = os.path.join(os.environ["FFPA_REPO_ROOT"], "gpurun_out")
if os.path.isdir(_out_dir):   # scratch copy of the measurements (the test asserts on the line below)
  json.dump(res, open(os.path.join(_out_dir, "dropin_results.json"), "w"), indent=1)
print("DROPIN_JSON " + json.dumps(res))
