"""CPU tests for the multi-GPU path: batch/head sharding partitions the work exactly, and the
world_size=2 gloo run (one process per rank, as bench.py under torchrun) reproduces the unsharded
result and the max-over-ranks timing reduction. Uses the CPU oracle as the per-shard operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffpa_attn.sharding import shard_for_rank
from oracle import attention_oracle as orc


@pytest.mark.parametrize("B,Hq,Hkv,world", [(8, 32, 8, 8), (8, 32, 8, 2), (3, 4, 2, 2), (1, 32, 8, 8), (1, 32, 8, 4),
                                              (2, 8, 4, 8), (1, 32, 32, 2), (5, 2, 1, 4)])
def test_shards_partition_the_work(B, Hq, Hkv, world):
  seen = np.zeros((B, Hq), dtype=int)
  seen_kv = np.zeros((B, Hkv), dtype=int)
  g = Hq // Hkv
  for r in range(world):
    s = shard_for_rank(B, Hq, Hkv, r, world)
    seen[s.batch, s.heads_q] += 1
    seen_kv[s.batch, s.heads_kv] += 1
    assert s.heads_q.start == s.heads_kv.start * g and s.heads_q.stop == s.heads_kv.stop * g
  assert (seen == 1).all() and (seen_kv == 1).all()


def test_unsplittable_work_raises():
  with pytest.raises(ValueError):
    shard_for_rank(1, 4, 1, 0, 2)  # one KV head cannot be split
  with pytest.raises(ValueError):
    shard_for_rank(3, 8, 8, 0, 4)  # batch < world and world % batch != 0


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, out_path):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  torch.manual_seed(0)  # same global tensors on every rank; each takes its shard
  q = torch.randn(1, 4, 48, 32)
  k = torch.randn(1, 2, 64, 32)
  v = torch.randn(1, 2, 64, 32)
  sh = shard_for_rank(1, 4, 2, rank, world)
  ql, kl, vl = sh.apply(q, k, v)
  o, _ = orc.attention_fwd(ql, kl, vl, causal=True)
  # timing reduction exactly as bench.py does it: max over ranks
  t = torch.tensor([10.0 + rank], dtype=torch.float64)
  dist.all_reduce(t, op=dist.ReduceOp.MAX)
  outs = [None] * world
  dist.all_gather_object(outs, (sh.heads_q.start, sh.heads_q.stop, o))
  if rank == 0:
    full = np.zeros((1, 4, 48, 32))
    for lo, hi, part in outs:
      full[:, lo:hi] = part
    np.save(out_path, np.concatenate([full.reshape(-1), [t.item()]]))
  dist.barrier()
  dist.destroy_process_group()


def test_world_size_2_gloo_matches_unsharded(tmp_path):
  out = str(tmp_path / "o.npy")
  mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
  got = np.load(out)
  torch.manual_seed(0)
  q = torch.randn(1, 4, 48, 32)
  k = torch.randn(1, 2, 64, 32)
  v = torch.randn(1, 2, 64, 32)
  ref, _ = orc.attention_fwd(q, k, v, causal=True)
  assert np.allclose(got[:-1], ref.reshape(-1), atol=1e-12)
  assert got[-1] == 11.0  # max over ranks


def test_bench_reference_arm_runs_on_rank_0_only():
  """bench.py --impl reference under a multi-rank launch: rank 0 alone times the CPU route and prints the line,
  every other rank exits 0 without work (the contract of the reference arm)."""
  import json
  import subprocess
  import sys

  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", FFPA_BENCH_WATCHDOG_S="0")
  p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                      "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
  assert p.returncode == 0 and p.stdout.strip() == ""
  env["RANK"] = env["LOCAL_RANK"] = "0"
  p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                      "--warmup", "0", "--workload", "d320_self_fwd"], env=env, capture_output=True, text=True, timeout=300)
  assert p.returncode == 0, p.stderr[-500:]
  line = json.loads(p.stdout.strip().splitlines()[-1])
  assert line["impl"] == "reference" and line["metric"] == "attn_fwd_tflops" and line["value"] > 0
  assert line["config"]["workload"] == "d320_self_fwd" and line["cpu_baseline"]["kind"] == "port"
  assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
