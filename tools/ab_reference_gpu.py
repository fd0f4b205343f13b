#!/usr/bin/env python
"""Same-box, same-process A/B of this repo's kernels against the reference's own GPU backends.

Builds the drop-in directory (the UNMODIFIED reference package from baseline/_ref + this repo's compiled
``_C`` / ``libffpa_b200.so``) in a temp dir, re-executes itself with that directory first on sys.path, and
times, interleaved on one GPU through the reference's public ``ffpa_attn_func``:

  ours          forward_backend="cuda"  (-> ffpa_attn._C -> libffpa_b200.so; backward: _C.ffpa_attn_backward)
  ref_cutedsl   backend="cutedsl"       (the reference's tcgen05 D=512 kernels / SM80-generic for other D)
  ref_triton    backend="triton"

with both protocols: CUDA events (warm-up 3, 10 iterations) and the reference's own wall-clock loop
(warm-up 2, 10 iterations, one trailing synchronize: /root/reference/src/ffpa_attn/cli/_runner_fwd.py:84-103).
TFLOPS use the reference formula (cli/_flops.py:36-76). An arm that cannot run on this box (the image ships
cutlass-dsl 4.5 while the reference pins 4.6; Triton 3.6 exceeds TMEM at D >= 512) is reported with its exception
text -- never dropped. Prints one line ``AB_JSON {...}``; bench.py embeds it as ``also.reference_gpu``."""
from __future__ import annotations

import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PKG = os.path.join(ROOT, "baseline", "_ref", "ffpa_attn")
OURS = os.path.join(ROOT, "ffpa-attn_b200", "ffpa_attn")

CASES = {
  # name: (B, Hq, Hkv, N, D, causal, bwd)
  "c2_self_d512": (1, 32, 32, 8192, 512, False, True),
  "c3_gqa_causal_d512": (1, 32, 8, 4096, 512, True, True),
  "d320_self": (1, 32, 32, 8192, 320, False, False),
  "d768_self": (1, 32, 32, 8192, 768, False, False),
  "d1024_self": (1, 32, 32, 8192, 1024, False, False),
}


def make_dropin_dir() -> str:
  d = tempfile.mkdtemp(prefix="ffpa_dropin_")
  dst = os.path.join(d, "ffpa_attn")
  shutil.copytree(REF_PKG, dst, ignore=shutil.ignore_patterns("__pycache__"))
  shutil.copy(glob.glob(os.path.join(OURS, "_C*.so"))[0], dst)
  shutil.copy(os.path.join(OURS, "libffpa_b200.so"), dst)
  return d


def parent(argv) -> int:
  if not os.path.isdir(REF_PKG):
    print("AB_JSON " + json.dumps({"unavailable": "baseline/_ref/ffpa_attn is absent (tools/install_reference.sh)"}))
    return 0
  d = make_dropin_dir()
  try:
    env = dict(os.environ, PYTHONPATH=d, FFPA_AB_CHILD="1", FFPA_CUDA_ALLOW_SMALL_D="1")
    return subprocess.call([sys.executable, os.path.abspath(__file__)] + argv, env=env, cwd=d)
  finally:
    shutil.rmtree(d, ignore_errors=True)


def child(argv) -> int:
  import argparse

  import torch

  ap = argparse.ArgumentParser()
  ap.add_argument("--cases", default=",".join(CASES))
  ap.add_argument("--budget-s", type=float, default=170.0)
  ap.add_argument("--device", type=int, default=0)
  args = ap.parse_args(argv)
  t_start = time.time()
  torch.cuda.set_device(args.device)
  dev = torch.device("cuda", args.device)
  import ffpa_attn
  from ffpa_attn import _C as C
  import ffpa_attn.cuda as cuda_mod

  try:
    import pynvml
    pynvml.nvmlInit()
    nv = pynvml.nvmlDeviceGetHandleByIndex(args.device)
    sm_clock = lambda: int(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))  # noqa: E731
  except Exception:  # noqa: BLE001
    sm_clock = lambda: None  # noqa: E731

  res = {"package": ffpa_attn.__file__, "native": C.__file__, "protocols": {"events": "warmup 3, iters 10, median of per-call CUDA-event times, best of 2 interleaved rounds", "wall": "warmup 2, iters 10, one sync (reference _time_fn)"},
         "cases": {}}

  def flops(B, H, N, D, causal):
    pairs = N * (N + 1) // 2 if causal else N * N
    return 4.0 * B * H * D * pairs

  def t_events(fn, warm=3, iters=10):
    """median of per-call CUDA-event times (robust against a single slow call on a warm, power-capped GPU)"""
    for _ in range(warm):
      fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
      fn()
      ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[iters // 2]

  def t_wall(fn, warm=2, iters=10):
    for _ in range(warm):
      fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
      fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3

  for name in args.cases.split(","):
    if name not in CASES:
      continue
    B, Hq, Hkv, N, D, causal, do_bwd = CASES[name]
    out = {"shape": {"B": B, "Hq": Hq, "Hkv": Hkv, "N": N, "D": D, "causal": causal}, "arms": {}}
    res["cases"][name] = out
    if time.time() - t_start > args.budget_s:
      out["skipped"] = "time budget of the default bench run exhausted"
      continue
    torch.manual_seed(42)
    q = torch.randn(B, Hq, N, D, dtype=torch.bfloat16, device=dev)
    k = torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device=dev)
    v = torch.randn(B, Hkv, N, D, dtype=torch.bfloat16, device=dev)
    d_o = torch.randn_like(q)
    kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)
    f = flops(B, Hq, N, D, causal)
    sc = D ** -0.5

    def mk_arm(arm):
      if arm == "ours":
        fwd = lambda: ffpa_attn.ffpa_attn_func(q, k, v, forward_backend="cuda", **kw)  # noqa: E731
        bwd = None
        if do_bwd:
          from ffpa_attn.cuda import _ffpa_attn_forward_cuda
          cuda_mod.set_cuda_backend_impl(cuda_mod.CudaBackendImpl.NATIVE)
          o2, lse = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), sc)
          lse = lse.contiguous()
          dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
          bwd = lambda: C.ffpa_attn_backward(q, k, v, o2, lse, d_o, dq, dk, dv, 0, int(causal), sc)  # noqa: E731
        return fwd, bwd
      backend = arm[len("ref_"):]
      fwd = lambda: ffpa_attn.ffpa_attn_func(q, k, v, backend=backend, **kw)  # noqa: E731
      bwd = None
      if do_bwd:
        qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
        o = ffpa_attn.ffpa_attn_func(qg, kg, vg, backend=backend, **kw)
        bwd = lambda: o.backward(d_o, retain_graph=True)  # noqa: E731
      return fwd, bwd

    fns = {}
    for arm in ("ours", "ref_cutedsl", "ref_triton"):
      a = {}
      out["arms"][arm] = a
      try:
        t0 = time.perf_counter()
        with torch.no_grad():
          fwd, _ = mk_arm(arm)
          o = fwd()
        torch.cuda.synchronize()
        a["first_call_s"] = round(time.perf_counter() - t0, 3)
        fwd, bwd = mk_arm(arm)
        if bwd is not None:
          bwd()
          torch.cuda.synchronize()
        fns[arm] = (fwd, bwd)
        if arm != "ours" and "ours" in fns:
          with torch.no_grad():
            a["max_abs_diff_vs_ours"] = float((o.float() - fns["ours"][0]().float()).abs().max())
      except Exception:  # noqa: BLE001
        a["error"] = traceback.format_exc().strip().splitlines()[-1][:400]
      torch.cuda.empty_cache()
    # interleaved timing: two rounds over the arms that run, best of the rounds per arm and protocol
    for rnd in range(2):
      for arm, (fwd, bwd) in fns.items():
        a = out["arms"][arm]
        with torch.no_grad():
          for proto, tf in (("events", t_events), ("wall", t_wall)):
            ms = tf(fwd)
            key = f"fwd_ms_{proto}"
            a[key] = min(a.get(key, 1e30), ms)
        if bwd is not None:
          for proto, tf in (("events", lambda fn: t_events(fn, 2, 5)), ("wall", lambda fn: t_wall(fn, 2, 5))):
            ms = tf(bwd)
            key = f"bwd_ms_{proto}"
            a[key] = min(a.get(key, 1e30), ms)
        a["sm_mhz_after"] = sm_clock()
    for arm in fns:
      a = out["arms"][arm]
      for proto in ("events", "wall"):
        a[f"fwd_tflops_{proto}"] = f / a[f"fwd_ms_{proto}"] * 1e-9
        if f"bwd_ms_{proto}" in a:
          a[f"bwd_tflops_{proto}"] = 2.5 * f / a[f"bwd_ms_{proto}"] * 1e-9
    ours = out["arms"].get("ours", {})
    for arm in ("ref_cutedsl", "ref_triton"):
      a = out["arms"][arm]
      if "fwd_ms_events" in a and "fwd_ms_events" in ours:
        a["ours_speedup_fwd_events"] = a["fwd_ms_events"] / ours["fwd_ms_events"]
        a["ours_speedup_fwd_wall"] = a["fwd_ms_wall"] / ours["fwd_ms_wall"]
        if "bwd_ms_events" in a and "bwd_ms_events" in ours:
          a["ours_speedup_bwd_events"] = a["bwd_ms_events"] / ours["bwd_ms_events"]
    del q, k, v, d_o, fns
    torch.cuda.empty_cache()
  res["elapsed_s"] = round(time.time() - t_start, 1)
  print("AB_JSON " + json.dumps(res))
  return 0


if __name__ == "__main__":
  sys.exit(child(sys.argv[1:]) if os.environ.get("FFPA_AB_CHILD") == "1" else parent(sys.argv[1:]))
