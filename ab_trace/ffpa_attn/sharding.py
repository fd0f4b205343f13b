"""Batch / head sharding for multi-GPU runs.

Attention is independent per (batch, query head); KV heads are shared only inside a GQA group.
So an N-GPU run needs no collective on the data path: rank r takes a contiguous slice of the batch,
or -- when the batch is smaller than the world -- a slice of query heads made of whole GQA groups
together with the matching KV heads.  The only cross-rank step is the timing reduction (max over
ranks), done by the caller with torch.distributed.  Pure index arithmetic: works on any device.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
  batch: slice
  heads_q: slice
  heads_kv: slice

  def apply(self, q, k, v):
    return (q[self.batch, self.heads_q], k[self.batch, self.heads_kv], v[self.batch, self.heads_kv])


def _even(n: int, parts: int, idx: int) -> slice:
  base, rem = divmod(n, parts)
  lo = idx * base + min(idx, rem)
  return slice(lo, lo + base + (1 if idx < rem else 0))


def shard_for_rank(batch: int, heads_q: int, heads_kv: int, rank: int, world: int) -> Shard:
  """Slice of (batch, q heads, kv heads) owned by ``rank``; slices over all ranks partition the work.
  Raises ValueError when the work cannot be split without cutting a GQA group."""
  if not 0 <= rank < world:
    raise ValueError(f"rank {rank} outside world {world}")
  if heads_q % heads_kv != 0:
    raise ValueError("heads_q must be a multiple of heads_kv")
  if batch >= world:
    return Shard(_even(batch, world, rank), slice(0, heads_q), slice(0, heads_kv))
  # fewer batch elements than ranks: split (batch x kv-head groups) units
  group = heads_q // heads_kv
  units = batch * heads_kv
  if units < world:
    raise ValueError(f"cannot shard B={batch}, Hkv={heads_kv} over {world} ranks without splitting a GQA group")
  if world % batch != 0:
    raise ValueError(f"world {world} must be a multiple of batch {batch} when batch < world")
  per_b = world // batch
  b, sub = divmod(rank, per_b)
  kv = _even(heads_kv, per_b, sub)
  return Shard(slice(b, b + 1), slice(kv.start * group, kv.stop * group), kv)
