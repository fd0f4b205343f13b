"""GPU tests of the packed-THD varlen entry (reference API: ffpa_attn_interface.py:192-279; semantics
cute/__init__.py:466-571: per-sequence lower-right causal, LSE [Hq, T_q])."""
import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _pack(lens_q, lens_k, Hq, Hkv, D, dtype, seed=0):
  torch.manual_seed(seed)
  cq = torch.tensor([0] + list(np.cumsum(lens_q)), dtype=torch.int32, device=DEV)
  ck = torch.tensor([0] + list(np.cumsum(lens_k)), dtype=torch.int32, device=DEV)
  q = torch.randn(int(cq[-1]), Hq, D).to(dtype).to(DEV)
  k = torch.randn(int(ck[-1]), Hkv, D).to(dtype).to(DEV)
  v = torch.randn(int(ck[-1]), Hkv, D).to(dtype).to(DEV)
  return q, k, v, cq, ck


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [128, 512])
def test_varlen_forward_lse_and_backward(causal, D):
  import ffpa_attn

  lens_q, lens_k = [130, 1, 257, 64], [200, 77, 257, 300]
  q, k, v, cq, ck = _pack(lens_q, lens_k, 4, 2, D, torch.bfloat16)
  out, lse = ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, max(lens_q), max(lens_k), causal=causal,
                                             enable_gqa=True, return_lse=True)
  assert out.shape == q.shape and lse.shape == (4, q.size(0)) and lse.dtype == torch.float32
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  d_o = torch.randn_like(q)
  og = ffpa_attn.ffpa_attn_varlen_func(qg, kg, vg, cq, ck, max(lens_q), max(lens_k), causal=causal, enable_gqa=True)
  og.backward(d_o)
  torch.cuda.synchronize()
  cql, ckl = cq.tolist(), ck.tolist()
  for b in range(len(lens_q)):
    sq, sk = slice(cql[b], cql[b + 1]), slice(ckl[b], ckl[b + 1])
    qb = q[sq].transpose(0, 1)[None].cpu()
    kb, vb = k[sk].transpose(0, 1)[None].cpu(), v[sk].transpose(0, 1)[None].cpu()
    ref, lref = orc.attention_fwd(qb, kb, vb, causal=causal)
    got = out[sq].transpose(0, 1)[None].float().cpu().numpy()
    assert np.abs(got - ref).max() < 2e-2
    assert np.abs(lse[:, sq].cpu().numpy() - lref[0]).max() < 2e-4
    assert np.abs(og[sq].transpose(0, 1)[None].detach().float().cpu().numpy() - ref).max() < 2e-2
    dob = d_o[sq].transpose(0, 1)[None].cpu()
    rq, rk, rv, _ = orc.attention_bwd(qb, kb, vb, dob, causal=causal)
    for got_g, want in ((qg.grad[sq].transpose(0, 1)[None], rq), (kg.grad[sk].transpose(0, 1)[None], rk),
                        (vg.grad[sk].transpose(0, 1)[None], rv)):
      err = np.abs(got_g.float().cpu().numpy() - want).max()
      assert err < 1e-1 * max(1.0, np.abs(want).max())


def test_varlen_validation_errors():
  import ffpa_attn

  q, k, v, cq, ck = _pack([64, 64], [64, 64], 2, 2, 64, torch.bfloat16)
  with pytest.raises(TypeError):
    ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq.long(), ck, 64, 64)
  with pytest.raises(NotImplementedError):
    ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, 64, 64, dropout_p=0.1)
  with pytest.raises(NotImplementedError):
    ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, 64, 64, window_size=(8, 8))
  with pytest.raises(ValueError):
    ffpa_attn.ffpa_attn_varlen_func(q, k[:, :1], v[:, :1], cq, ck, 64, 64)
  with pytest.raises(ValueError):  # batch-size mismatch between the two offset vectors
    ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq[:-1], ck, 64, 64)
  # like the reference, cu_seqlens VALUES are never read on the host; a malformed vector is clamped to the
  # packed extent on the device instead of addressing past the tensors
  bad = cq.clone()
  bad[-1] += 64
  out = ffpa_attn.ffpa_attn_varlen_func(q, k, v, bad, ck, 128, 64)
  torch.cuda.synchronize()
  assert torch.isfinite(out.float()).all()


def test_varlen_single_launch_many_sequences_and_graph_capture():
  """One forward launch and four backward launches for the whole packed batch (no per-sequence dispatch,
  no host read of cu_seqlens): launch counter + CUDA-graph capture + empty sequences + D = 1024."""
  import ffpa_attn

  lens_q = [5, 0, 129, 300, 64, 17, 256, 1]
  lens_k = [5, 40, 129, 0, 200, 17, 512, 9]
  q, k, v, cq, ck = _pack(lens_q, lens_k, 4, 4, 256, torch.float16, seed=2)
  n0 = ffpa_attn._C.launch_count()
  out, lse = ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, 300, 512, causal=False, return_lse=True)
  torch.cuda.synchronize()
  assert ffpa_attn._C.launch_count() - n0 == 1
  cql, ckl = cq.tolist(), ck.tolist()
  for b in range(len(lens_q)):
    sq, sk = slice(cql[b], cql[b + 1]), slice(ckl[b], ckl[b + 1])
    if lens_q[b] == 0:
      continue
    if lens_k[b] == 0:   # no key: O = 0, LSE = -inf (SM100 convention, _fwd_d512_sm100.py:2635-2646)
      assert out[sq].float().abs().max().item() == 0.0 and torch.isinf(lse[:, sq]).all()
      continue
    ref, lref = orc.attention_fwd(q[sq].transpose(0, 1)[None].cpu(), k[sk].transpose(0, 1)[None].cpu(),
                                  v[sk].transpose(0, 1)[None].cpu())
    assert np.abs(out[sq].transpose(0, 1)[None].float().cpu().numpy() - ref).max() < 1e-2
    assert np.abs(lse[:, sq].cpu().numpy() - lref[0]).max() < 2e-4
  # graph capture: nothing on the path synchronises
  g = torch.cuda.CUDAGraph()
  static_out = None
  s = torch.cuda.Stream()
  s.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(s):
    ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, 300, 512)
  torch.cuda.current_stream().wait_stream(s)
  with torch.cuda.graph(g):
    static_out = ffpa_attn.ffpa_attn_varlen_func(q, k, v, cq, ck, 300, 512)
  g.replay()
  torch.cuda.synchronize()
  assert torch.equal(static_out, out)


@pytest.mark.parametrize("D", [64, 320, 1024])
def test_varlen_causal_cross_lengths_and_head_dims(D):
  """Per-sequence bottom-right causal incl. Nkv < Nq sequences (leading rows see no key -> O = 0, zero
  gradients), forward + backward in packed mode."""
  import ffpa_attn

  lens_q, lens_k = [100, 260, 33], [228, 260, 20]
  q, k, v, cq, ck = _pack(lens_q, lens_k, 2, 1, D, torch.bfloat16, seed=4)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  d_o = torch.randn_like(q)
  n0 = ffpa_attn._C.launch_count()
  og = ffpa_attn.ffpa_attn_varlen_func(qg, kg, vg, cq, ck, 260, 260, causal=True, enable_gqa=True)
  og.backward(d_o)
  torch.cuda.synchronize()
  assert ffpa_attn._C.launch_count() - n0 == 5   # fwd + preprocess + dQ + dK + dV
  cql, ckl = cq.tolist(), ck.tolist()
  for b in range(len(lens_q)):
    sq, sk = slice(cql[b], cql[b + 1]), slice(ckl[b], ckl[b + 1])
    qb, kb, vb = (t.transpose(0, 1)[None].cpu() for t in (q[sq], k[sk], v[sk]))
    ref, _ = orc.attention_fwd(qb, kb, vb, causal=True)
    assert np.abs(og[sq].transpose(0, 1)[None].detach().float().cpu().numpy() - ref).max() < 2e-2
    rq, rk, rv, _ = orc.attention_bwd(qb, kb, vb, d_o[sq].transpose(0, 1)[None].cpu(), causal=True)
    for got_g, want in ((qg.grad[sq].transpose(0, 1)[None], rq), (kg.grad[sk].transpose(0, 1)[None], rk),
                        (vg.grad[sk].transpose(0, 1)[None], rv)):
      err = np.abs(got_g.float().cpu().numpy() - want).max()
      assert err < 1e-1 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("causal,Hkv,chunks", [(False, 4, 8), (True, 2, 2), (False, 4, 3)])
def test_host_buffer_entry_matches_device_call_and_oracle(causal, Hkv, chunks):
  """ffpa_attn_host_func: pinned host q/k/v in, pinned host output back; head-chunked pipeline must be
  bit-identical to the one-shot device call (same kernel, same per-head arithmetic)."""
  import ffpa_attn

  torch.manual_seed(3)
  B, Hq, N, D = 2, 4, 384, 256
  hq = torch.randn(B, Hq, N, D).to(torch.bfloat16).pin_memory()
  hk = torch.randn(B, Hkv, N, D).to(torch.bfloat16).pin_memory()
  hv = torch.randn(B, Hkv, N, D).to(torch.bfloat16).pin_memory()
  ho = torch.zeros(B, Hq, N, D, dtype=torch.bfloat16).pin_memory()
  ret = ffpa_attn.ffpa_attn_host_func(hq, hk, hv, out=ho, is_causal=causal, enable_gqa=Hq != Hkv, chunks=chunks)
  assert ret is ho
  dense = ffpa_attn.ffpa_attn_func(hq.to(DEV), hk.to(DEV), hv.to(DEV), is_causal=causal, enable_gqa=Hq != Hkv)
  assert torch.equal(dense.cpu(), ho)
  ref, _ = orc.attention_fwd(hq, hk, hv, causal=causal)
  assert np.abs(ho.float().numpy() - ref).max() < 1e-2
  # pageable inputs and an allocated output also work
  out2 = ffpa_attn.ffpa_attn_host_func(hq.clone(), hk.clone(), hv.clone(), is_causal=causal, enable_gqa=Hq != Hkv)
  assert torch.equal(out2, ho)


@pytest.mark.parametrize("causal", [False, True])
def test_varlen_lse_output_is_differentiable(causal):
  """The LSE returned by the packed entry carries a gradient (reference: cute/_bwd_preprocess.py:6-15):
  dS gains P * dLSE, folded into delta by the preprocess kernel. Checked against fp32 autograd."""
  import ffpa_attn

  lens = [200, 77]
  q, k, v, cq, ck = _pack(lens, lens, 2, 2, 128, torch.bfloat16, seed=9)
  d_o = torch.randn_like(q)
  w = torch.randn(2, q.size(0), device=DEV)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out, lse = ffpa_attn.ffpa_attn_varlen_func(qg, kg, vg, cq, ck, max(lens), max(lens), causal=causal, return_lse=True)
  ((out.float() * d_o.float()).sum() + (lse * w).sum()).backward()
  torch.cuda.synchronize()
  cql = cq.tolist()
  for b in range(len(lens)):
    s = slice(cql[b], cql[b + 1])
    q32, k32, v32 = (t[s].float().transpose(0, 1).detach().requires_grad_(True) for t in (q, k, v))   # [H, n, D]
    sc = (q32 @ k32.transpose(-1, -2)) * 128 ** -0.5
    if causal:
      n = sc.size(-1)
      sc = sc.masked_fill(~torch.ones(n, n, dtype=torch.bool, device=DEV).tril(), float("-inf"))
    ref_lse = torch.logsumexp(sc, dim=-1)
    ref_out = torch.softmax(sc, dim=-1) @ v32
    ((ref_out * d_o[s].float().transpose(0, 1)).sum() + (ref_lse * w[:, s]).sum()).backward()
    assert (lse[:, s] - ref_lse).abs().max().item() < 2e-4
    for got, want, name in ((qg.grad[s], q32.grad, "dQ"), (kg.grad[s], k32.grad, "dK"), (vg.grad[s], v32.grad, "dV")):
      err = (got.float().transpose(0, 1) - want).abs().max().item()
      assert err < 5e-2 * max(1.0, want.abs().max().item()), f"{name} seq {b}: {err}"
  # LSE alone (no gradient through out) also works: d_o is materialised as zeros
  qg2 = q.clone().requires_grad_(True)
  _, lse2 = ffpa_attn.ffpa_attn_varlen_func(qg2, k, v, cq, ck, max(lens), max(lens), causal=causal, return_lse=True)
  (lse2 * w).sum().backward()
  assert torch.isfinite(qg2.grad.float()).all() and qg2.grad.float().abs().max().item() > 0
