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
  return {"err": err(out, ref), "early_rows_bit_equal_16bit": bool(torch.equal(out[:, :, :256], o16[:, :, :256])),
          "late_rows_differ": bool((out[:, :, 256:] != o16[:, :, 256:]).any())}


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


print("DROPIN_JSON " + json.dumps(res))
