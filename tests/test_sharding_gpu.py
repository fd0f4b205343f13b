"""GPU side of the multi-GPU story: ONE global problem is cut by ffpa_attn.sharding into
per-rank shards, every shard goes through the sm_100a kernels (forward and backward), and the gather must equal
the unsharded run bit for bit -- there is no collective to get wrong, only index arithmetic and strided views.
On a single GPU the ranks run one after the other on that device; with >= 2 GPUs visible a second test puts each
shard on its own device (the driver's round-end box has one GPU, so that one usually skips)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _global(B, Hq, Hkv, N, D, seed=0):
  g = torch.Generator().manual_seed(seed)
  mk = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)  # noqa: E731
  return mk(B, Hq, N, D), mk(B, Hkv, N, D), mk(B, Hkv, N, D), mk(B, Hq, N, D)


def _run(q, k, v, d_o, dev, causal):
  import ffpa_attn

  q, k, v = (t.to(dev).requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, enable_gqa=q.size(1) != k.size(1))
  out.backward(d_o.to(dev))
  return tuple(t.detach().cpu() for t in (out, q.grad, k.grad, v.grad))


@pytest.mark.parametrize("B,Hq,Hkv,world", [(4, 8, 2, 4), (1, 8, 4, 4), (2, 8, 8, 8), (3, 4, 2, 2)])
@pytest.mark.parametrize("causal", [False, True])
def test_sharded_ranks_reproduce_the_unsharded_result(B, Hq, Hkv, world, causal):
  from ffpa_attn.sharding import shard_for_rank

  q, k, v, d_o = _global(B, Hq, Hkv, 384, 512)
  full = _run(q, k, v, d_o, "cuda:0", causal)
  got = [torch.zeros_like(t) for t in full]
  for rank in range(world):
    sh = shard_for_rank(B, Hq, Hkv, rank, world)
    ql, kl, vl = sh.apply(q, k, v)              # strided views of the global tensors: no copies on the host side
    o, dq, dk, dv = _run(ql, kl, vl, d_o[sh.batch, sh.heads_q], "cuda:0", causal)
    got[0][sh.batch, sh.heads_q] = o
    got[1][sh.batch, sh.heads_q] = dq
    got[2][sh.batch, sh.heads_kv] = dk
    got[3][sh.batch, sh.heads_kv] = dv
  for name, a, b in zip(("O", "dQ", "dK", "dV"), got, full):
    assert torch.equal(a, b), name


def test_shards_on_separate_devices_reproduce_the_unsharded_result():
  from ffpa_attn.sharding import shard_for_rank

  world = torch.cuda.device_count()
  if world < 2:
    pytest.skip("needs >= 2 visible GPUs (covered on one device by the test above)")
  world = min(world, 8)
  B, Hq, Hkv = world, 8, 4
  q, k, v, d_o = _global(B, Hq, Hkv, 512, 512, seed=1)
  full = _run(q, k, v, d_o, "cuda:0", True)
  for rank in range(world):
    sh = shard_for_rank(B, Hq, Hkv, rank, world)
    ql, kl, vl = sh.apply(q, k, v)
    o, dq, dk, dv = _run(ql, kl, vl, d_o[sh.batch, sh.heads_q], f"cuda:{rank}", True)
    assert torch.equal(o, full[0][sh.batch, sh.heads_q]) and torch.equal(dq, full[1][sh.batch, sh.heads_q])
    assert torch.equal(dk, full[2][sh.batch, sh.heads_kv]) and torch.equal(dv, full[3][sh.batch, sh.heads_kv])
