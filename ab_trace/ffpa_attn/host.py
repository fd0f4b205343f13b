"""Host-buffer entry point: attention over q/k/v that live in (pinned) host memory.

The operator is independent per (batch, KV-head group) -- the same fact ``sharding.py`` uses across
GPUs -- so a call whose operands start on the host does not have to wait for the whole host->device
copy: the work is cut into head chunks and chunk c+1 is copied in while chunk c runs on the sm_100a
kernel and chunk c-1 is copied out (three streams, PCIe is full duplex).  Every byte of q, k and v
crosses the bus once per call and every byte of the output comes back; with D=512 the kernel is
5-6x faster than the bus, so the call runs at the host->device copy rate instead of the sum of
copy-in + kernel + copy-out.

This is what ``bench.py`` times as ``e2e`` (reference call site: the same ``ffpa_attn_func`` signature,
/root/reference/src/ffpa_attn/ffpa_attn_interface.py:71-189, with host tensors moved by the caller).
Forward only; no autograd (host tensors carry no graph).
"""
from __future__ import annotations

import torch

from .ffpa_attn_interface import ffpa_attn_func


def _units(batch: int, heads_kv: int, chunks: int):
  """Contiguous (batch index, kv-head slice) work units, at most ``chunks`` per batch element."""
  per_b = max(1, min(heads_kv, chunks))
  base, rem = divmod(heads_kv, per_b)
  out = []
  for b in range(batch):
    lo = 0
    for i in range(per_b):
      hi = lo + base + (1 if i < rem else 0)
      out.append((b, lo, hi))
      lo = hi
  return out


def ffpa_attn_host_func(
  query: torch.Tensor,
  key: torch.Tensor,
  value: torch.Tensor,
  out: torch.Tensor | None = None,
  *,
  is_causal: bool = False,
  scale: float | None = None,
  enable_gqa: bool = False,
  chunks: int = 8,
  device: torch.device | str | None = None,
  sync: bool = True,
  **kwargs: object,
) -> torch.Tensor:
  """``ffpa_attn_func`` for host-resident operands.

  ``query`` [B, Hq, Nq, D], ``key`` / ``value`` [B, Hkv, Nkv, D]: CPU fp16/bf16 tensors (pinned memory
  gives asynchronous copies; pageable memory works but serialises).  ``out``: optional CPU tensor of
  query's shape (pinned for an asynchronous copy back); allocated pinned when omitted.  ``chunks``:
  pieces per batch element along the KV-head axis (whole GQA groups).  Returns ``out``; with
  ``sync=True`` (default) the data is complete on return, otherwise the caller must synchronise
  the current stream of ``device`` before reading it.
  """
  if kwargs.get("attn_mask") is not None or kwargs.get("dropout_p", 0.0) != 0.0:
    raise NotImplementedError("ffpa_attn_host_func: attn_mask / dropout are not supported on the chunked host path")
  if query.is_cuda or key.is_cuda or value.is_cuda:
    raise ValueError("ffpa_attn_host_func takes host tensors; use ffpa_attn_func for device tensors")
  if query.dim() != 4 or key.dim() != 4 or value.dim() != 4:
    raise ValueError("query/key/value must be [B, H, N, D]")
  if key.shape != value.shape:
    raise ValueError("key and value must have the same shape")
  B, Hq, Nq, D = query.shape
  Hkv = key.size(1)
  if Hq != Hkv and not enable_gqa:
    raise ValueError("Hq != Hkv needs enable_gqa=True")
  if Hq % Hkv != 0:
    raise ValueError("Hq must be a multiple of Hkv")
  group = Hq // Hkv
  if not torch.cuda.is_available():
    raise RuntimeError("ffpa_attn_host_func needs a CUDA device: there is no CPU fallback")
  dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
  if out is None:
    out = torch.empty(query.shape, dtype=query.dtype).pin_memory()
  elif out.is_cuda or out.shape != query.shape or out.dtype != query.dtype:
    raise ValueError("out must be a host tensor with query's shape and dtype")

  with torch.cuda.device(dev):
    main = torch.cuda.current_stream(dev)
    s_in, s_out = _side_streams(dev)
    # Full-size device staging, allocated on the caller's stream: chunks are slices of it, so there
    # is no buffer recycling inside the call and the allocator's stream ordering stays trivial.
    dq = torch.empty(query.shape, dtype=query.dtype, device=dev)
    dk = torch.empty(key.shape, dtype=key.dtype, device=dev)
    dv = torch.empty(value.shape, dtype=value.dtype, device=dev)
    s_in.wait_stream(main)
    s_out.wait_stream(main)
    for (b, lo, hi) in _units(B, Hkv, chunks):
      qs = slice(lo * group, hi * group)
      with torch.cuda.stream(s_in):
        dk[b, lo:hi].copy_(key[b, lo:hi], non_blocking=True)
        dv[b, lo:hi].copy_(value[b, lo:hi], non_blocking=True)
        dq[b, qs].copy_(query[b, qs], non_blocking=True)
        ev_in = s_in.record_event()
      main.wait_event(ev_in)
      with torch.no_grad():
        o = ffpa_attn_func(dq[b:b + 1, qs], dk[b:b + 1, lo:hi], dv[b:b + 1, lo:hi], is_causal=is_causal,
                           scale=scale, enable_gqa=enable_gqa, **kwargs)
      ev_cmp = main.record_event()
      o.record_stream(s_out)  # allocated on the caller's stream, read by the copy-out stream
      with torch.cuda.stream(s_out):
        s_out.wait_event(ev_cmp)
        out[b:b + 1, qs].copy_(o, non_blocking=True)
    main.wait_stream(s_out)
    main.wait_stream(s_in)
    if sync:
      main.synchronize()
  return out


_STREAMS: dict = {}


def _side_streams(dev: torch.device):
  key = (dev.type, dev.index)
  if key not in _STREAMS:
    _STREAMS[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
  return _STREAMS[key]


def bind_to_gpu_numa_node(device: torch.device | str | int | None = None) -> dict | None:
  """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that host buffers pinned
  afterwards (first touch) are local to the GPU's PCIe root: with one process per GPU all landing on whatever
  node the launcher left them on, every host->device copy crosses the socket interconnect and the copies of
  different ranks contend for it. Reads only sysfs; returns {"node", "cpus", "pci"} or None when the topology
  is not exposed (then nothing is changed)."""
  import os

  dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
  cpus, how, node = set(), None, None
  try:
    props = torch.cuda.get_device_properties(dev)
    bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    if node >= 0:
      for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
      how = "sysfs"
  except Exception:  # noqa: BLE001
    pass
  if not cpus:
    # containers often hide the PCI device's numa_node (-1); NVML still knows the GPU's ideal CPU set
    # (the "CPU Affinity" column of `nvidia-smi topo -m`)
    try:
      import pynvml

      pynvml.nvmlInit()
      idx = dev.index if dev.index is not None else torch.cuda.current_device()
      visible = os.environ.get("CUDA_VISIBLE_DEVICES")
      if visible:
        idx = int(visible.split(",")[idx]) if visible.split(",")[idx].isdigit() else idx
      h = pynvml.nvmlDeviceGetHandleByIndex(idx)
      words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
      for w, mask in enumerate(words):
        for bit in range(64):
          if (int(mask) >> bit) & 1:
            cpus.add(64 * w + bit)
      how = "nvml"
    except Exception:  # noqa: BLE001  (no NVML: leave the affinity alone)
      return None
  try:
    allowed = os.sched_getaffinity(0) & cpus
    if not allowed:
      return None
    os.sched_setaffinity(0, allowed)
    return {"via": how, "node": node, "cpus": len(allowed)}
  except Exception:  # noqa: BLE001
    return None
