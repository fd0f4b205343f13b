#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json metric): attention TFLOPS on synthetic (B, H, N, D) tensors.

    python bench.py --gpus N --steps K --warmup W              # this repo's sm_100a kernels, default = C2 forward
    python bench.py --workload d1024_self_fwd ...              # any BASELINE config as a first-class workload
    python bench.py --impl reference --gpus N ...              # the reference's CPU route, timed

Default workload = BASELINE.json configs[1]: forward, B=1 (per GPU), H=32, N=8192, D=512, bf16. The other BASELINE
configs are first-class workloads with their own metric / roofline: c2_bwd, c3_fwd_bwd (GQA causal, config 3),
c4_fp8_fwd (config 4), d320 / d768 / d1024_self_fwd (config 5). A "step" is one pass of the workload's hot path over
one batch of synthetic (seeded randn) input. Multi-GPU = batch sharding: every rank runs the same per-GPU workload
on its own batch element, no collective on the data path -> "scaling": "weak".
TFLOPS use the reference's dominant-GEMM formula 4*B*Hq*D*pairs, backward = 2.5 x forward
(/root/reference/src/ffpa_attn/cli/_flops.py:36-76).

JSON keys beyond the base contract:
  roofline      tensor-core bound; achieved = algorithmic FLOPs per step / mean per-step duration from CUDA events on
                the launching stream; peak = MEASURED_PEAKS.json bf16 (burst for a timed region < 1 s, else sustained;
                FP8: 2 x that, stated); traffic = ncu dram bytes per launch from profiles/roofline_traffic.json when that
                capture was taken from the kernel sources being run (hash match), else null.
  sustained     the same step repeated for >= 1 s (>= 400 steps at C2) right after the K timed steps: TFLOP/s, fraction
                of the SUSTAINED measured peak, SM clock -- the K-step number is a burst.
  per_rank_ms   each rank's own time for the K steps (the headline uses the max).
  cpu_baseline  the reference's CPU route (aten SDPA, what ffpa_attn_func(backend="sdpa") runs on CPU tensors) timed on
                this box's host cores on a bounded head-sample of the workload.
  e2e           same metric through the public host-buffer call with pinned HOST tensors in and out, copies inside the
                timed region; plus the measured raw concurrent pinned H2D rate of the same bytes (platform ceiling).
  also          secondary workloads in the same run, and reference_gpu: the same-box A/B against the reference's own
                GPU backends (tools/ab_reference_gpu.py) -- measured numbers or the captured exception.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
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

# name: B per GPU, Hq, Hkv, Nq, Nkv, D, causal, kind (fwd | bwd | fwd_bwd | fp8_fwd)
WORKLOADS = {
  "c2_self_fwd_b1h32n8192d512": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=512, causal=False, kind="fwd"),
  "c2_bwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=512, causal=False, kind="bwd"),
  "c3_fwd_bwd": dict(B=1, Hq=32, Hkv=8, Nq=4096, Nkv=4096, D=512, causal=True, kind="fwd_bwd"),
  "c3_gqa_causal_fwd_hq32hkv8n4096d512": dict(B=1, Hq=32, Hkv=8, Nq=4096, Nkv=4096, D=512, causal=True, kind="fwd"),
  "c4_fp8_fwd": dict(B=4, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=256, causal=False, kind="fp8_fwd"),
  "d320_self_fwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=320, causal=False, kind="fwd"),
  "d768_self_fwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=768, causal=False, kind="fwd"),
  "d1024_self_fwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=1024, causal=False, kind="fwd"),
  "d256_self_fwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=256, causal=False, kind="fwd"),
  "d128_self_fwd": dict(B=1, Hq=32, Hkv=32, Nq=8192, Nkv=8192, D=128, causal=False, kind="fwd"),
}
DEFAULT_WORKLOAD = "c2_self_fwd_b1h32n8192d512"
METRIC = {"fwd": "attn_fwd_tflops", "fp8_fwd": "attn_fwd_tflops", "bwd": "attn_bwd_tflops", "fwd_bwd": "attn_fwd_bwd_tflops"}
FLOP_MULT = {"fwd": 1.0, "fp8_fwd": 1.0, "bwd": 2.5, "fwd_bwd": 3.5}
KERNELS = {"fwd": "ffpa_fwd_kernel", "fp8_fwd": "quantize_e4m3_kernel + ffpa_fwd_fp8_kernel",
           "bwd": "bwd_preprocess + ffpa_bwd_kernel<dQ> + dK / dV kernels", "fwd_bwd": "ffpa_fwd_kernel + backward launch set"}


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


def kernel_source_hash(files) -> str:
  """sha of the sources of one kernel: a stored ncu traffic figure is only quoted for the sources it was captured
  from (tools/update_roofline_traffic.py writes the record together with the file list)."""
  h = hashlib.sha256()
  for f in sorted(files):
    h.update(open(os.path.join(PKG, "csrc", f), "rb").read())
  return h.hexdigest()[:16]


def measured_traffic(workload: str):
  path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
  try:
    rec = json.load(open(path)).get(workload)
  except Exception:
    return None, "no capture"
  if not isinstance(rec, dict):
    return None, "no capture for this workload"
  if rec.get("kernel_source_sha16") != kernel_source_hash(rec.get("kernel_sources", [])):
    return None, f"capture {rec.get('capture')} is from other kernel sources (sha {rec.get('kernel_source_sha16')})"
  return rec.get("dram_bytes_per_launch"), rec.get("capture")


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

  def window(self, t0: float, t1: float):
    sm, smax, reasons, power = [], None, set(), []
    for ts, line in list(self.rows):
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
    if not sm:  # window shorter than the sampling period: use every sample we have
      for ts, line in list(self.rows):
        f = [x.strip() for x in line.split(",")]
        try:
          sm.append(float(f[1]))
        except (ValueError, IndexError):
          pass
    sm.sort()
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None, "sm_max_mhz": smax,
            "reasons": sorted(reasons), "power_w_max": max(power) if power else None, "samples": len(sm)}

  def stop(self):
    if self.proc:
      time.sleep(0.12)
      self.proc.terminate()


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


def reference_arm(args, wl, config):
  """--impl reference: the reference's own CPU implementation of the path (aten SDPA, the route
  ffpa_attn_func(backend="sdpa") takes on CPU tensors), all host threads, a bounded head-sample per step."""
  Hq, Hkv, Nq, Nkv, D, causal = wl["Hq"], wl["Hkv"], wl["Nq"], wl["Nkv"], wl["D"], wl["causal"]
  tf, heads, secs, threads = cpu_sdpa_sample(Hq, Hkv, Nq, Nkv, D, causal, 1.5)
  from oracle import attention_oracle as orc

  group = Hq // Hkv
  torch.manual_seed(42)
  q = torch.randn(1, heads, Nq, D, dtype=torch.bfloat16)
  k = torch.randn(1, max(1, heads // group), Nkv, D, dtype=torch.bfloat16)
  v = torch.randn_like(k)
  times = []
  for i in range(args.warmup + args.steps):
    t0 = time.perf_counter()
    orc.sdpa_cpu(q, k, v, is_causal=causal, enable_gqa=heads != k.size(1))
    dt = time.perf_counter() - t0
    if i >= args.warmup:
      times.append(dt)
  sample_flops = flops_of(1, heads, Nq, Nkv, D, causal)
  total = sum(times)
  val = sample_flops * len(times) / total * 1e-12
  sample = f"{heads} of {Hq} heads per step (B=1, Nq={Nq}, Nkv={Nkv}, D={D}, bf16, forward), aten SDPA on host"
  return {
    "impl": "reference", "metric": METRIC[wl["kind"]], "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
    "config": config,
    "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample},
    "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }


class Workload:
  """Device tensors + the step closure of one workload (public API only)."""

  def __init__(self, wl, dev, seed):
    import ffpa_attn

    self.wl, self.dev = wl, dev
    B, Hq, Hkv, Nq, Nkv, D = (wl[x] for x in ("B", "Hq", "Hkv", "Nq", "Nkv", "D"))
    kind = wl["kind"]
    dt = torch.bfloat16
    amp = 0.5 if kind == "fp8_fwd" else 1.0   # /root/reference/tests/test_ffpa_fp8.py:63-65
    torch.manual_seed(seed)
    self.q = torch.randn(B, Hq, Nq, D, dtype=dt, device=dev) * amp
    self.k = torch.randn(B, Hkv, Nkv, D, dtype=dt, device=dev) * amp
    self.v = torch.randn(B, Hkv, Nkv, D, dtype=dt, device=dev) * amp
    self.kw = dict(is_causal=wl["causal"], enable_gqa=Hq != Hkv)
    if kind == "fp8_fwd":
      self.kw["forward_backend"] = ffpa_attn.CUDABackend(enable_fp8=True)
    self.flops = flops_of(B, Hq, Nq, Nkv, D, wl["causal"]) * FLOP_MULT[kind]
    self.fn = ffpa_attn.ffpa_attn_func
    if kind in ("bwd", "fwd_bwd"):
      self.d_o = torch.randn(B, Hq, Nq, D, dtype=dt, device=dev)
      self.qg, self.kg, self.vg = (t.detach().requires_grad_(True) for t in (self.q, self.k, self.v))
      if kind == "bwd":
        self.out = self.fn(self.qg, self.kg, self.vg, **self.kw)

  def step(self):
    kind = self.wl["kind"]
    if kind in ("fwd", "fp8_fwd"):
      with torch.no_grad():
        return self.fn(self.q, self.k, self.v, **self.kw)
    if kind == "bwd":
      return torch.autograd.grad(self.out, (self.qg, self.kg, self.vg), self.d_o, retain_graph=True)
    out = self.fn(self.qg, self.kg, self.vg, **self.kw)
    return torch.autograd.grad(out, (self.qg, self.kg, self.vg), self.d_o)

  def algorithmic_hbm_bytes(self):
    es = self.q.element_size()
    fwd = es * (2 * self.q.numel() + self.k.numel() + self.v.numel())
    if self.wl["kind"] in ("fwd", "fp8_fwd"):
      return int(fwd)
    bwd = es * (4 * self.q.numel() + 2 * self.k.numel() + 2 * self.v.numel())   # Q,K,V,O,dO in; dQ,dK,dV out
    return int(bwd if self.wl["kind"] == "bwd" else fwd + bwd)


def time_steps(step, n, warm):
  """(mean, median, min) ms over n steps, one CUDA event pair per step. The secondary lines quote the MEDIAN: they run
  back to back on a GPU that is already at its power cap, and a single slow step (clock change) moves a 5-10 step
  mean by tens of percent."""
  for _ in range(warm):
    step()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
  ev[0].record()
  for i in range(n):
    step()
    ev[i + 1].record()
  torch.cuda.synchronize()
  ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
  return ev[0].elapsed_time(ev[n]) / n, ts[n // 2], ts[0]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads and the reference-GPU A/B")
  ap.add_argument("--no-ab", action="store_true", help="skip only the reference-GPU A/B subprocess")
  ap.add_argument("--sustain-seconds", type=float, default=1.0)
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
  # watchdog: a run that has not finished after 20 minutes (normal: 2-3) dumps every thread's stack and exits instead
  # of hanging its caller (FFPA_BENCH_WATCHDOG_S overrides, 0 disables)
  wd = float(os.environ.get("FFPA_BENCH_WATCHDOG_S", "1200"))
  if wd > 0:
    import faulthandler
    faulthandler.dump_traceback_later(wd, exit=True)

  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  wl = WORKLOADS[args.workload]
  kind = wl["kind"]
  in_bytes = 2 * (wl["B"] * wl["Hq"] * wl["Nq"] * wl["D"] + 2 * wl["B"] * wl["Hkv"] * wl["Nkv"] * wl["D"])
  config = {"workload": args.workload, "kind": kind, "batch_per_gpu": wl["B"], "heads_q": wl["Hq"], "heads_kv": wl["Hkv"],
            "seqlen_q": wl["Nq"], "seqlen_kv": wl["Nkv"], "head_dim": wl["D"], "causal": wl["causal"],
            "sharding": f"batch x{world}, no collective",
            "l2": f"inputs ({in_bytes / 1e6:.0f} MB) exceed the 126 MB L2; no explicit flush"}

  if args.impl == "reference":
    if rank != 0:
      return 0
    print(json.dumps(reference_arm(args, wl, config)))
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
  from ffpa_attn import host as ffpa_host

  # pinned host buffers of the e2e leg should live on the GPU's NUMA node: bind before anything is pinned
  numa = ffpa_host.bind_to_gpu_numa_node(dev)

  W = Workload(wl, dev, 42 + rank)
  step = W.step

  def barrier():
    if use_dist:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(x: float) -> float:
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if use_dist:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  for _ in range(args.warmup):
    out = step()
  barrier()

  # ---- device-resident timing: K steps bracketed by barrier + synchronize ----
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
    time.sleep(0.15)
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
  launches0 = ffpa_attn.cuda.launch_count()
  barrier()
  t_wall0 = time.time()
  ev[0].record()
  for i in range(args.steps):
    out = step()
    ev[i + 1].record()
  barrier()
  t_wall1 = time.time()
  launches = ffpa_attn.cuda.launch_count() - launches0
  total_ms = ev[0].elapsed_time(ev[-1])
  per_step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
  ms_per_step = max_over_ranks(total_ms) / args.steps
  value = W.flops * world / (ms_per_step * 1e-3) * 1e-12
  per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
  per_rank[rank] = total_ms / args.steps
  if use_dist:
    dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
  per_rank_ms = [round(float(x), 4) for x in per_rank.tolist()]

  # ---- the reference's own protocol next to the CUDA-event number: wall clock around the public call, warm-up 2,
  # ---- 10 iterations, ONE trailing synchronize (/root/reference/src/ffpa_attn/cli/_runner_fwd.py:84-103)
  for _ in range(2):
    out = step()
  torch.cuda.synchronize()
  tw0 = time.perf_counter()
  for _ in range(10):
    out = step()
  torch.cuda.synchronize()
  wall_ms = (time.perf_counter() - tw0) / 10 * 1e3
  reference_protocol = {"ms_per_step": wall_ms, "value": W.flops / (wall_ms * 1e-3) * 1e-12, "unit": "TFLOP/s (this rank)",
                        "protocol": "wall clock, warm-up 2, iters 10, one trailing synchronize (cli/_runner_fwd.py:84-103)"}

  # ---- sustained: the same step back to back for >= sustain_seconds (a K-step region of ~60 ms is a burst) ----
  sustained = None
  if args.sustain_seconds > 0:
    n_sus = max(int(math.ceil(args.sustain_seconds * 1e3 / max(ms_per_step, 1e-3))), 50)
    n_sus = min(n_sus, 4000)
    barrier()
    ts0 = time.time()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n_sus):
      out = step()
    b.record()
    barrier()
    ts1 = time.time()
    sus_ms = max_over_ranks(a.elapsed_time(b)) / n_sus
    sustained = {"steps": n_sus, "ms_per_step": sus_ms, "seconds": sus_ms * n_sus * 1e-3,
                 "value": W.flops * world / (sus_ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "window": (ts0, ts1)}
  clocks = None
  if rank == 0:
    sampler.stop()
    clocks = sampler.window(t_wall0, t_wall1)
    if sustained is not None:
      sustained["clocks"] = sampler.window(*sustained.pop("window"))
  elif sustained is not None:
    sustained.pop("window")

  # ---- correctness spot check inside the bench (rank 0): sampled rows vs the oracle (forward kinds) ----
  max_abs_err = None
  if rank == 0 and kind in ("fwd", "fp8_fwd"):
    from oracle import attention_oracle as orc
    import numpy as np

    Nq, Nkv = wl["Nq"], wl["Nkv"]
    rows = [0, Nq // 2, Nq - 1]
    if wl["causal"]:
      bias = np.where(np.arange(Nkv)[None, :] <= (np.array(rows)[:, None] + (Nkv - Nq)), 0.0, -np.inf)[None, None]
    else:
      bias = None
    ref, _ = orc.attention_fwd(W.q[:1, :1, rows].cpu(), W.k[:1, :1].cpu(), W.v[:1, :1].cpu(), bias=bias)
    max_abs_err = float(np.abs(out[0, 0, rows].float().cpu().numpy() - ref[0, 0]).max())

  # ---- end-to-end: host (pinned) buffers in, result back to host, copies inside the timed region ----
  e2e = None
  if not args.no_e2e:
    e2e = run_e2e(W, ffpa_attn, dev, world, use_dist, barrier, max_over_ranks, min(args.steps, 10))
    if e2e is not None:
      e2e["numa_binding"] = numa

  if rank != 0:
    if use_dist:
      dist.destroy_process_group()
    return 0

  # ---- roofline of the step's kernels ----
  peaks = measured_peaks()
  mult = 2.0 if kind == "fp8_fwd" else 1.0
  timed_s = total_ms * 1e-3
  peak_kind = "burst" if timed_s < 1.0 else "sustained"
  peak = peaks[peak_kind] * mult
  mean_ms = sum(per_step_ms) / len(per_step_ms)
  achieved = W.flops / (mean_ms * 1e-3) * 1e-12
  traffic, traffic_src = measured_traffic(args.workload)
  roofline = {"bound": "tensor", "kernel": KERNELS[kind], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
              "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
              "peak_source": f"{peaks['src']} bf16 {peak_kind} (timed region {timed_s:.3f} s)" +
                             (" x 2 for e4m3 operands (no measured fp8 peak on this pool)" if mult == 2.0 else ""),
              "flops_per_step": W.flops, "algorithmic_hbm_bytes_per_step": W.algorithmic_hbm_bytes(),
              "launch_ms_mean": mean_ms, "launch_ms_min": min(per_step_ms)}
  if sustained is not None:
    sustained["peak"] = peaks["sustained"] * mult
    sustained["frac"] = sustained["value"] / world / sustained["peak"]
    sustained["peak_source"] = f"{peaks['src']} bf16 sustained" + (" x 2 (fp8)" if mult == 2.0 else "")

  # ---- secondary workloads + same-box A/B against the reference's GPU backends ----
  also = None
  if world == 1 and not args.no_also:
    del out
    also = run_also(args, dev, peaks)

  # ---- CPU baseline on this box's host cores (bounded sample) ----
  cpu = None
  if not args.no_cpu_baseline and world == 1:
    tf, heads, secs, threads = cpu_sdpa_sample(wl["Hq"], wl["Hkv"], wl["Nq"], wl["Nkv"], wl["D"], wl["causal"], budget_s=12.0)
    cpu = {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port",
           "sample": f"{heads} of {wl['Hq']} heads of the same workload (forward), one pass, {secs:.2f} s, aten SDPA bf16 on "
                     f"host ({os.cpu_count()} logical CPUs)"}

  published = 1456.0 if args.workload == DEFAULT_WORKLOAD else None
  line = {
    "metric": METRIC[kind], "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
    "vs_baseline": (value / world / published) if published else None,
    "dtype": "fp8_e4m3" if kind == "fp8_fwd" else "bf16", "data": "synthetic", "config": config, "roofline": roofline,
    "sustained": sustained, "per_rank_ms": per_rank_ms, "reference_protocol": reference_protocol,
    "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
    "max_abs_err_vs_oracle": max_abs_err, "also": also,
    "reference_published": {"value": published, "unit": "TFLOP/s", "where": "bench/README.md:132 (CuTe-DSL tcgen05, B200, "
                            "someone else's box; the same-box measurement is also.reference_gpu)",
                            "ratio": value / world / published} if published else None,
  }
  print(json.dumps(line))
  if use_dist:
    dist.destroy_process_group()
  return 0


def run_e2e(W, ffpa_attn, dev, world, use_dist, barrier, max_over_ranks, n_e2e):
  """Through the public host-buffer API: pinned q/k/v -> device -> kernels -> result back in host memory, all inside
  the timed region. Forward kinds use ffpa_attn_host_func (chunk-pipelined copy-in / kernel / copy-out); backward
  kinds copy q, k, v, dO in, run the step and copy dQ, dK, dV out on one stream."""
  wl = W.wl
  kind = wl["kind"]
  n_e2e = max(3, n_e2e)
  esz = W.q.element_size()
  hq, hk, hv = (t.detach().cpu().pin_memory() for t in (W.q, W.k, W.v))
  if kind in ("fp8_fwd", "bwd", "fwd_bwd"):
    return _e2e_simple(W, ffpa_attn, dev, world, barrier, max_over_ranks, n_e2e, (hq, hk, hv))
  ho = torch.empty(W.q.shape, dtype=W.q.dtype).pin_memory()
  kw = {k: v for k, v in W.kw.items()}
  chunks = int(os.environ.get("FFPA_E2E_CHUNKS", "8"))

  def e2e_step():
    ffpa_attn.ffpa_attn_host_func(hq, hk, hv, out=ho, chunks=chunks, **kw)

  for _ in range(2):
    e2e_step()
  # CPU time to ISSUE one call (no wait for the GPU): what the host side adds per step
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  ffpa_attn.ffpa_attn_host_func(hq, hk, hv, out=ho, chunks=chunks, sync=False, **kw)
  issue_ms = (time.perf_counter() - t0) * 1e3
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n_e2e):
    e2e_step()
  e1.record()
  barrier()
  e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
  h2d = int((W.q.numel() + W.k.numel() + W.v.numel()) * esz)
  d2h = int(W.q.numel() * esz)
  # platform ceiling: the same bytes as plain pinned copies on every rank at once (no kernels): inputs host->device
  # on one stream while an output-sized buffer goes device->host on another (what the pipelined call overlaps)
  dq, dk, dv = (torch.empty_like(t) for t in (W.q, W.k, W.v))
  do_ = torch.empty_like(W.q)
  side = torch.cuda.Stream(dev)
  def raw():
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
      ho.copy_(do_, non_blocking=True)
    dq.copy_(hq, non_blocking=True); dk.copy_(hk, non_blocking=True); dv.copy_(hv, non_blocking=True)
    torch.cuda.current_stream(dev).wait_stream(side)
  raw()
  barrier()
  r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  r0.record()
  for _ in range(3):
    raw()
  r1.record()
  barrier()
  raw_ms = max_over_ranks(r0.elapsed_time(r1)) / 3
  raw_gbs = h2d / (raw_ms * 1e-3) * 1e-9
  gbs = h2d / (e2e_ms * 1e-3) * 1e-9
  return {"value": W.flops * world / (e2e_ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
          "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": n_e2e, "chunks": chunks, "host_issue_ms": issue_ms,
          "h2d_gbs_per_gpu": gbs, "raw_concurrent_h2d_gbs_per_gpu": raw_gbs, "raw_copies_ms": raw_ms,
          "raw_copies": "all ranks at once: q, k, v pinned host->device + an output-sized device->host copy on a second stream",
          "bound": "pcie / host memory (the step runs at >= 85 % of the raw concurrent pinned-copy rate of its inputs)"
                   if gbs >= 0.85 * raw_gbs else "host pipeline (below the raw pinned-copy rate measured in this run)"}


def _e2e_simple(W, ffpa_attn, dev, world, barrier, max_over_ranks, n_e2e, hqkv):
  kind = W.wl["kind"]
  hq, hk, hv = hqkv
  esz = W.q.element_size()
  bwd = kind in ("bwd", "fwd_bwd")
  hdo = W.d_o.cpu().pin_memory() if bwd else None
  outs_host = None

  def e2e_step():
    nonlocal outs_host
    q = hq.to(dev, non_blocking=True); k = hk.to(dev, non_blocking=True); v = hv.to(dev, non_blocking=True)
    if not bwd:
      with torch.no_grad():
        res = (ffpa_attn.ffpa_attn_func(q, k, v, **W.kw),)
    else:
      d_o = hdo.to(dev, non_blocking=True)
      q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
      o = ffpa_attn.ffpa_attn_func(q, k, v, **W.kw)
      res = torch.autograd.grad(o, (q, k, v), d_o)
    if outs_host is None:
      outs_host = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res]
    for h, r in zip(outs_host, res):
      h.copy_(r, non_blocking=True)
    torch.cuda.current_stream().synchronize()

  for _ in range(2):
    e2e_step()
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n_e2e):
    e2e_step()
  e1.record()
  barrier()
  ms = max_over_ranks(e0.elapsed_time(e1)) / n_e2e
  h2d = int((hq.numel() + hk.numel() + hv.numel() + (hdo.numel() if bwd else 0)) * esz)
  d2h = int(sum(h.numel() for h in outs_host) * esz)
  flops = W.flops if kind != "bwd" else W.flops * 3.5 / 2.5   # the e2e backward call has to run its forward too
  return {"value": flops * world / (ms * 1e-3) * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
          "ms_per_step": ms, "steps": n_e2e, "h2d_gbs_per_gpu": h2d / (ms * 1e-3) * 1e-9,
          "note": "single-stream copy-in, kernels, copy-out" + ("; forward + backward FLOPs (3.5 x) counted" if kind == "bwd" else "")}


def run_also(args, dev, peaks):
  """Secondary workloads of the same metric family in the same process, each with its own roofline fraction, then
  the same-box A/B against the reference's GPU backends (subprocess: both packages register the same torch ops)."""
  import ffpa_attn

  try:
    import pynvml
    pynvml.nvmlInit()
    _nv = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
    sm_clock = lambda: int(pynvml.nvmlDeviceGetClockInfo(_nv, pynvml.NVML_CLOCK_SM))  # noqa: E731
  except Exception:  # noqa: BLE001
    sm_clock = lambda: None  # noqa: E731
  also = {"protocol": "per workload, back to back in one process: 5 warm-up steps, 10 timed steps (6 for backward kinds), one CUDA "
                      "event pair per step; value = FLOPs / MEDIAN step time (ms_mean / ms_min alongside); sm_mhz_after = SM "
                      "clock right after the timed steps (the GPU is warm: later lines run at lower clocks than the headline)"}
  names = [n for n in ("c2_bwd", "c3_fwd_bwd", "c3_gqa_causal_fwd_hq32hkv8n4096d512", "c4_fp8_fwd", "d320_self_fwd",
                       "d768_self_fwd", "d1024_self_fwd") if n != args.workload]
  for name in names:
    wl = WORKLOADS[name]
    try:
      W = Workload(wl, dev, 7)
      ms_mean, ms, ms_min = time_steps(W.step, 10 if wl["kind"] in ("fwd", "fp8_fwd") else 6, 5)
      mhz = sm_clock()
      mult = 2.0 if wl["kind"] == "fp8_fwd" else 1.0
      rec = {"metric": METRIC[wl["kind"]], "ms_per_step": ms, "ms_mean": ms_mean, "ms_min": ms_min,
             "value": W.flops / ms * 1e-9, "value_best_step": W.flops / ms_min * 1e-9, "unit": "TFLOP/s",
             "peak": peaks["burst"] * mult, "frac": W.flops / ms * 1e-9 / (peaks["burst"] * mult), "sm_mhz_after": mhz}
      if name == "c2_bwd":
        # the O(N)-memory backward (three recompute kernels) next to the default (score stash from free memory)
        be = ffpa_attn.CUDABackend(bwd_min_workspace=True)
        o = ffpa_attn.ffpa_attn_func(W.qg, W.kg, W.vg, backend=be)
        _, ms_r, _ = time_steps(lambda: torch.autograd.grad(o, (W.qg, W.kg, W.vg), W.d_o, retain_graph=True), 6, 5)
        rec["min_workspace_ms"] = ms_r
        rec["min_workspace_value"] = W.flops / ms_r * 1e-9
        del o
      if name == "c4_fp8_fwd":
        W.kw.pop("forward_backend")
        _, ms16, _ = time_steps(W.step, 10, 5)
        rec["bf16_kernel_ms_same_inputs"] = ms16
        rec["fp8_speedup_over_bf16_kernel"] = ms16 / ms
      also[name] = rec
      del W
      torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
      also[name] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
  if not args.no_ab:
    try:
      p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ab_reference_gpu.py"), "--device", str(dev.index or 0)],
                         capture_output=True, text=True, timeout=420)
      lines = [l for l in p.stdout.splitlines() if l.startswith("AB_JSON ")]
      also["reference_gpu"] = json.loads(lines[-1][8:]) if lines else {
        "error": f"A/B subprocess rc={p.returncode}: {(p.stderr or p.stdout)[-600:]}"}
    except Exception as e:  # noqa: BLE001
      also["reference_gpu"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
  return also


if __name__ == "__main__":
  sys.exit(main())
