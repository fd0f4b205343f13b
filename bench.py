#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric): attention forward TFLOPS at
B=1 (per GPU), H=32, N=8192, D=512, bf16 -- BASELINE.json configs[1].

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --gpus N ...            # the reference's CPU route, timed

A "step" is one forward pass of the hot path over one batch of synthetic (seeded randn) input.
Multi-GPU = batch sharding: every rank runs the same per-GPU workload on its own batch element,
no collective on the data path -> "scaling": "weak".
TFLOPS use the reference's dominant-GEMM formula 4*B*Hq*D*pairs
(/root/reference/src/ffpa_attn/cli/_flops.py:36-54).

JSON keys beyond the base contract:
  roofline      tensor-core bound; achieved = algorithmic FLOPs per launch / mean per-launch
                duration from CUDA events on the launching stream; peak = MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU route (aten SDPA, what ffpa_attn_func(backend="sdpa") runs on
                CPU tensors) timed on this box's host cores on a bounded head-sample of the workload.
  e2e           same metric through the public API with HOST (pinned) q/k/v and the output copied
                back, both copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ffpa-attn_b200")
for _p in (ROOT, PKG):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import torch  # noqa: E402

WORKLOADS = {
  # name: (B per GPU, Hq, Hkv, Nq, Nkv, D, causal)
  "c2_self_fwd_b1h32n8192d512": (1, 32, 32, 8192, 8192, 512, False),
  "c3_gqa_causal_fwd_hq32hkv8n4096d512": (1, 32, 8, 4096, 4096, 512, True),
  "d320_self_fwd": (1, 32, 32, 8192, 8192, 320, False),
  "d256_self_fwd": (1, 32, 32, 8192, 8192, 256, False),
  "d128_self_fwd": (1, 32, 32, 8192, 8192, 128, False),
}
DEFAULT_WORKLOAD = "c2_self_fwd_b1h32n8192d512"


def flops_of(B, Hq, Nq, Nkv, D, causal):
  pairs = (Nq * (Nkv - Nq) + Nq * (Nq + 1) // 2) if causal else Nq * Nkv
  return 4.0 * B * Hq * D * pairs


def measured_peaks():
  path = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(path):
    try:
      d = json.load(open(path))
      return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
              "src": "MEASURED_PEAKS.json"}
    except Exception:
      pass
  return {"burst": 1590.0, "sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled during the timed region."""

  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index: int):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    try:
      self.proc = subprocess.Popen(
        ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
        stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.time(), line.strip()))

  def stop(self, t0: float, t1: float):
    if not self.proc:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.12)
    self.proc.terminate()
    sm, smax, reasons, power = [], None, set(), []
    for ts, line in self.rows:
      f = [x.strip() for x in line.split(",")]
      if len(f) < 8:
        continue
      try:
        clk, mx = float(f[1]), float(f[2])
      except ValueError:
        continue
      smax = mx
      if t0 - 0.05 <= ts <= t1 + 0.05:
        sm.append(clk)
        try:
          power.append(float(f[3]))
        except ValueError:
          pass
        for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
          if val.lower().startswith("active"):
            reasons.add(name)
    if not sm:  # timed region shorter than the sampling period: use every sample we have
      for ts, line in self.rows:
        f = [x.strip() for x in line.split(",")]
        try:
          sm.append(float(f[1]))
        except (ValueError, IndexError):
          pass
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
            "power_w_max": max(power) if power else None, "samples": len(sm)}


def cpu_sdpa_sample(Hq, Hkv, Nq, Nkv, D, causal, budget_s, max_heads=None):
  """Time the reference's CPU route (aten SDPA on host tensors) on a head-sample of the workload.
  Returns (tflops, heads_used, seconds, threads)."""
  from oracle import attention_oracle as orc

  # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use the host's cores
  want = max(1, (os.cpu_count() or 2) // 2)
  if torch.get_num_threads() < want:
    torch.set_num_threads(want)
  threads = torch.get_num_threads()
  group = Hq // Hkv
  torch.manual_seed(42)

  def run(hq):
    hkv = max(1, hq // group)
    q = torch.randn(1, hq, Nq, D, dtype=torch.bfloat16)
    k = torch.randn(1, hkv, Nkv, D, dtype=torch.bfloat16)
    v = torch.randn(1, hkv, Nkv, D, dtype=torch.bfloat16)
    t0 = time.perf_counter()
    orc.sdpa_cpu(q, k, v, is_causal=causal, enable_gqa=hq != hkv)
    return time.perf_counter() - t0

  h0 = group  # smallest sample that keeps whole KV heads
  t_probe = run(h0)
  heads = h0
  if t_probe < budget_s / 2:
    heads = int(min(Hq, max(h0, (budget_s / max(t_probe, 1e-3)) * h0)))
    heads = max(h0, (heads // group) * group)
    if max_heads:
      heads = min(heads, max_heads)
  secs = run(heads) if heads != h0 else t_probe
  return flops_of(1, heads, Nq, Nkv, D, causal) / secs * 1e-12, heads, secs, threads


def dist_setup(n_gpus):
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  return rank, local, world


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

  rank, local, world = dist_setup(args.gpus)
  B, Hq, Hkv, Nq, Nkv, D, causal = WORKLOADS[args.workload]
  step_flops = flops_of(B, Hq, Nq, Nkv, D, causal)
  config = {"workload": args.workload, "batch_per_gpu": B, "heads_q": Hq, "heads_kv": Hkv, "seqlen_q": Nq,
            "seqlen_kv": Nkv, "head_dim": D, "causal": causal, "sharding": f"batch x{world}, no collective",
            "l2": "inputs (1.07 GB at C2) exceed the 126 MB L2; no explicit flush"}

  # ------------------------------------------------------------------ reference arm (CPU) ----
  if args.impl == "reference":
    if rank != 0:
      return 0
    # each step = a bounded head-sample of the workload through the reference's CPU route
    per_step_budget = 1.5
    tf, heads, secs, threads = cpu_sdpa_sample(Hq, Hkv, Nq, Nkv, D, causal, per_step_budget)
    times = []
    from oracle import attention_oracle as orc

    group = Hq // Hkv
    torch.manual_seed(42)
    q = torch.randn(1, heads, Nq, D, dtype=torch.bfloat16)
    k = torch.randn(1, max(1, heads // group), Nkv, D, dtype=torch.bfloat16)
    v = torch.randn_like(k)
    for i in range(args.warmup + args.steps):
      t0 = time.perf_counter()
      orc.sdpa_cpu(q, k, v, is_causal=causal, enable_gqa=heads != k.size(1))
      dt = time.perf_counter() - t0
      if i >= args.warmup:
        times.append(dt)
    sample_flops = flops_of(1, heads, Nq, Nkv, D, causal)
    total = sum(times)
    val = sample_flops * len(times) / total * 1e-12
    sample = f"{heads} of {Hq} heads per step (B=1, Nq={Nq}, Nkv={Nkv}, D={D}, bf16), aten SDPA on host"
    line = {
      "impl": "reference", "metric": "attn_fwd_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
      "config": config,
      "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample},
      "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0

  # ------------------------------------------------------------------ our arm (B200) ---------
  if not torch.cuda.is_available():
    print(json.dumps({"error": "no CUDA device: this benchmark has no CPU path for --impl ours"}))
    return 1
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  use_dist = world > 1
  if use_dist:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
  import __graft_entry__ as ge

  ge.build()
  import ffpa_attn

  torch.manual_seed(42 + rank)
  dt = torch.bfloat16
  q = torch.randn(B, Hq, Nq, D, dtype=dt, device=dev)
  k = torch.randn(B, Hkv, Nkv, D, dtype=dt, device=dev)
  v = torch.randn(B, Hkv, Nkv, D, dtype=dt, device=dev)
  kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)

  def step():
    return ffpa_attn.ffpa_attn_func(q, k, v, **kw)

  def barrier():
    if use_dist:
      dist.barrier()
    torch.cuda.synchronize()

  for _ in range(args.warmup):
    out = step()
  barrier()

  # ---- device-resident timing: K steps bracketed by barrier + synchronize ----
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
    time.sleep(0.15)
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
  launches0 = ffpa_attn._C.launch_count()
  barrier()
  t_wall0 = time.time()
  ev[0].record()
  for i in range(args.steps):
    out = step()
    ev[i + 1].record()
  barrier()
  t_wall1 = time.time()
  launches = ffpa_attn._C.launch_count() - launches0
  total_ms = ev[0].elapsed_time(ev[-1])
  per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
  clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
  tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
  if use_dist:
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
  total_ms_max = float(tmax.item())
  ms_per_step = total_ms_max / args.steps
  value = step_flops * world / (ms_per_step * 1e-3) * 1e-12

  # ---- correctness spot check inside the bench (rank 0): sampled rows vs the oracle ----
  max_abs_err = None
  if rank == 0:
    from oracle import attention_oracle as orc
    import numpy as np

    rows = [0, Nq // 2, Nq - 1]
    g = Hq // Hkv
    qs = q[:1, :1, rows].cpu()
    if causal:
      off = Nkv - Nq
      bias = np.where(np.arange(Nkv)[None, :] <= (np.array(rows)[:, None] + off), 0.0, -np.inf)[None, None]
    else:
      bias = None
    ref, _ = orc.attention_fwd(qs, k[:1, :1].cpu(), v[:1, :1].cpu(), bias=bias)
    max_abs_err = float(np.abs(out[0, 0, rows].float().cpu().numpy() - ref[0, 0]).max())

  # ---- end-to-end: host (pinned) buffers in, output back to host, copies inside the timed region ----
  e2e = None
  if not args.no_e2e:
    hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
    ho = torch.empty(q.shape, dtype=dt).pin_memory()
    n_e2e = max(3, min(args.steps, 10))

    def e2e_step():
      # public host-buffer call: head-chunked copy-in / kernel / copy-out pipeline, result complete
      # in `ho` (host) when it returns
      ffpa_attn.ffpa_attn_host_func(hq, hk, hv, out=ho, **kw)

    for _ in range(2):
      e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_e2e):
      e2e_step()
    e1.record()
    barrier()
    t_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if use_dist:
      dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.item()) / n_e2e
    esz = q.element_size()
    e2e = {"value": step_flops * world / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s",
           "h2d_bytes_per_step": int((q.numel() + k.numel() + v.numel()) * esz),
           "d2h_bytes_per_step": int(q.numel() * esz), "ms_per_step": e2e_ms, "steps": n_e2e}

  if rank != 0:
    if use_dist:
      dist.destroy_process_group()
    return 0

  # ---- roofline of the dominant (only) kernel ----
  peaks = measured_peaks()
  timed_s = total_ms * 1e-3
  peak_kind = "burst" if timed_s < 1.0 else "sustained"
  peak = peaks[peak_kind]
  mean_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
  achieved = step_flops / (mean_launch_ms * 1e-3) * 1e-12
  traffic = None
  tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
  if os.path.exists(tpath):
    try:
      traffic = json.load(open(tpath)).get(args.workload)
    except Exception:
      traffic = None
  roofline = {"bound": "tensor", "kernel": "ffpa_fwd_kernel", "achieved": achieved, "peak": peak,
              "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
              "peak_source": f"{peaks['src']} bf16 {peak_kind} (timed region {timed_s:.3f} s)",
              "flops_per_launch": step_flops,
              "algorithmic_hbm_bytes_per_launch": int(2 * (2 * q.numel() + k.numel() + v.numel())),
              "launch_ms_mean": mean_launch_ms, "launch_ms_min": min(per_launch_ms)}

  peaks = measured_peaks()
  # ---- secondary numbers of the same metric family (BASELINE.json: "attn TFLOPS (fwd, bwd)") ----
  also = None
  if world == 1 and not args.no_e2e:
    also = {}
    def _t(fn, n):
      for _ in range(2):
        fn()
      torch.cuda.synchronize()
      a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(n):
        fn()
      b_.record()
      torch.cuda.synchronize()
      return a.elapsed_time(b_) / n
    for name in (args.workload, "c3_gqa_causal_fwd_hq32hkv8n4096d512"):
      b2, hq2, hkv2, nq2, nkv2, d2, c2 = WORKLOADS[name]
      torch.manual_seed(7)
      qg = torch.randn(b2, hq2, nq2, d2, dtype=dt, device=dev, requires_grad=True)
      kg = torch.randn(b2, hkv2, nkv2, d2, dtype=dt, device=dev, requires_grad=True)
      vg = torch.randn(b2, hkv2, nkv2, d2, dtype=dt, device=dev, requires_grad=True)
      kw2 = dict(is_causal=c2, enable_gqa=hq2 != hkv2)
      f2 = flops_of(b2, hq2, nq2, nkv2, d2, c2)
      ms_f = _t(lambda: ffpa_attn.ffpa_attn_func(qg.detach(), kg.detach(), vg.detach(), **kw2), 10)
      o2 = ffpa_attn.ffpa_attn_func(qg, kg, vg, **kw2)
      do2 = torch.randn_like(o2)
      ms_b = _t(lambda: o2.backward(do2, retain_graph=True), 5)
      also[name] = {"fwd_ms": ms_f, "fwd_tflops": f2 / ms_f * 1e-9, "bwd_ms": ms_b,
                    "bwd_tflops": 2.5 * f2 / ms_b * 1e-9, "bwd_flops_rule": "2.5 x fwd (reference _flops.py:57-76)"}
      del qg, kg, vg, o2, do2
    # BASELINE config 4: FP8 forward, B=4 H=32 N=8192 D=256 (quantise pre-pass + attention timed together)
    try:
      torch.manual_seed(11)
      q8 = torch.randn(4, 32, 8192, 256, dtype=dt, device=dev) * 0.5
      k8 = torch.randn(4, 32, 8192, 256, dtype=dt, device=dev) * 0.5
      v8 = torch.randn(4, 32, 8192, 256, dtype=dt, device=dev) * 0.5
      f8 = flops_of(4, 32, 8192, 8192, 256, False)
      be = ffpa_attn.CUDABackend(enable_fp8=True)
      ms8 = _t(lambda: ffpa_attn.ffpa_attn_func(q8, k8, v8, forward_backend=be), 10)
      ms16 = _t(lambda: ffpa_attn.ffpa_attn_func(q8, k8, v8), 10)
      also["c4_fp8_fwd_b4h32n8192d256"] = {"fp8_ms": ms8, "fp8_tflops": f8 / ms8 * 1e-9, "bf16_ms": ms16,
                                          "bf16_tflops": f8 / ms16 * 1e-9,
                                          "fp8_peak_tflops": 2 * peaks["burst"], "fp8_frac": f8 / ms8 * 1e-9 / (2 * peaks["burst"])}
      del q8, k8, v8
    except Exception as e:  # noqa: BLE001
      also["c4_fp8_fwd_b4h32n8192d256"] = {"error": str(e)}

  # ---- CPU baseline on this box's host cores (bounded sample) ----
  cpu = None
  if not args.no_cpu_baseline and world == 1:
    tf, heads, secs, threads = cpu_sdpa_sample(Hq, Hkv, Nq, Nkv, D, causal, budget_s=12.0)
    cpu = {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port",
           "sample": f"{heads} of {Hq} heads of the same workload, one pass, {secs:.2f} s, aten SDPA bf16 on host "
                     f"({os.cpu_count()} logical CPUs)"}

  line = {
    "metric": "attn_fwd_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": (value / world / 1456.0) if args.workload == DEFAULT_WORKLOAD else None,
    "dtype": "bf16", "data": "synthetic", "config": config, "roofline": roofline,
    "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
    "max_abs_err_vs_oracle": max_abs_err, "also": also,
    "reference_published": {"value": 1456.0, "unit": "TFLOP/s", "where": "bench/README.md:132 (CuTe-DSL tcgen05, B200)",
                            "ratio": value / world / 1456.0 if args.workload == DEFAULT_WORKLOAD else None},
  }
  print(json.dumps(line))
  if use_dist:
    dist.destroy_process_group()
  return 0


if __name__ == "__main__":
  sys.exit(main())
