"""GPU parity tests for the backward (dQ, dK, dV) through the public API / autograd ->
torch.ops.ffpa_attn._bwd_cuda -> C ABI -> sm_100a kernels. Oracle: oracle/attention_oracle.py
(pinned on the reference's own autograd results in tests/golden/*_bwd.npz); full-size configs use a
plain fp32 torch restatement on the GPU. Tolerances follow the reference's backward tests
(/root/reference/tests/test_ffpa_bwd.py:38-45, 975-985): fp16 1e-2, bf16 5e-2, causal bf16 large 1e-1;
and its analytic KATs (/root/reference/tests/test_ffpa_cute_sm100.py:1026-1050)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _tol(dtype):
  return 5e-2 if dtype == torch.bfloat16 else 1e-2


def _grads(q, k, v, d_o, **kw):
  import ffpa_attn

  q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
  n0 = ffpa_attn._C.launch_count()
  out = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
  out.backward(d_o)
  torch.cuda.synchronize()
  assert ffpa_attn._C.launch_count() - n0 >= 5, "expected fwd + preprocess + dQ + dK + dV launches"
  return out.detach(), q.grad, k.grad, v.grad


def _mk(B, Hq, Hkv, Nq, Nkv, D, dtype, seed=0):
  torch.manual_seed(seed)
  q = torch.randn(B, Hq, Nq, D).to(dtype).to(DEV)
  k = torch.randn(B, Hkv, Nkv, D).to(dtype).to(DEV)
  v = torch.randn(B, Hkv, Nkv, D).to(dtype).to(DEV)
  d_o = torch.randn(B, Hq, Nq, D).to(dtype).to(DEV)
  return q, k, v, d_o


def _cmp(got, want, tol, name):
  g = got.float().cpu().numpy()
  assert np.isfinite(g).all(), f"{name}: non-finite gradient"
  err = np.abs(g - want).max()
  bound = tol * max(1.0, float(np.abs(want).max()))
  assert err < bound, f"{name}: max-abs-err {err:.3e} >= {bound:.3e}"


def _check_against_oracle(q, k, v, d_o, causal, dtype, enable_gqa=False, tol=None):
  _, dq, dk, dv = _grads(q, k, v, d_o, is_causal=causal, enable_gqa=enable_gqa)
  rq, rk, rv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), causal=causal)
  tol = tol or _tol(dtype)
  _cmp(dq, rq, tol, "dQ")
  _cmp(dk, rk, tol, "dK")
  _cmp(dv, rv, tol, "dV")


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*_bwd.npz")))


@pytest.mark.parametrize("path", [p for p in GOLDEN if "mask" not in p],
                         ids=[os.path.basename(p)[:-4] for p in GOLDEN if "mask" not in p])
def test_backward_matches_reference_golden(path):
  z = np.load(path)
  B, Hq, Hkv, Nq, Nkv, D, is_bf16, causal = [int(x) for x in z["meta"]]
  dt = torch.bfloat16 if is_bf16 else torch.float16
  t = lambda n, s: torch.from_numpy(z[n].view(np.int16).copy()).view(dt).reshape(s).to(DEV)  # noqa: E731
  q, k, v = t("q", (B, Hq, Nq, D)), t("k", (B, Hkv, Nkv, D)), t("v", (B, Hkv, Nkv, D))
  d_o = t("d_o", (B, Hq, Nq, D))
  _, dq, dk, dv = _grads(q, k, v, d_o, is_causal=bool(causal), enable_gqa=Hq != Hkv)
  for got, name in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
    _cmp(got, z[name].astype(np.float64), _tol(dt), name)


@pytest.mark.parametrize("D", [64, 128, 256, 320, 512])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_backward_headdims(D, dtype):
  q, k, v, d_o = _mk(1, 2, 2, 256, 256, D, dtype)
  _check_against_oracle(q, k, v, d_o, False, dtype)


@pytest.mark.parametrize("Nq,Nkv", [(128, 128), (130, 257), (1, 300), (7, 129), (384, 200), (500, 1000)])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_shapes_and_tails(Nq, Nkv, causal):
  if causal and Nkv < Nq:
    pytest.skip("causal needs Nkv >= Nq")
  q, k, v, d_o = _mk(2, 2, 2, Nq, Nkv, 320, torch.float16, seed=1)
  _check_against_oracle(q, k, v, d_o, causal, torch.float16)


@pytest.mark.parametrize("Hq,Hkv", [(4, 2), (8, 1), (6, 3)])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_gqa(Hq, Hkv, causal):  # tests/test_ffpa_bwd.py:991-997
  q, k, v, d_o = _mk(1, Hq, Hkv, 256, 384, 512, torch.bfloat16, seed=2)
  _check_against_oracle(q, k, v, d_o, causal, torch.bfloat16, enable_gqa=True, tol=1e-1 if causal else None)


def test_backward_single_key_kat():
  """s_k == 1: softmax is identically 1 -> dQ = dK = 0 exactly-ish, dV = sum over rows of dO."""
  q, k, v, d_o = _mk(1, 2, 2, 200, 1, 128, torch.bfloat16, seed=3)
  _, dq, dk, dv = _grads(q, k, v, d_o)
  assert dq.float().abs().max().item() < 1e-2
  assert dk.float().abs().max().item() < 5e-2
  want = d_o.float().sum(dim=2, keepdim=True)
  assert (dv.float() - want).abs().max().item() < 5e-2 * max(1.0, want.abs().max().item())


def test_backward_linear_in_dO():
  """Size-independent property: gradients are linear in dO."""
  q, k, v, d_o = _mk(1, 2, 2, 300, 300, 256, torch.float16, seed=4)
  _, a_q, a_k, a_v = _grads(q, k, v, d_o)
  _, b_q, b_k, b_v = _grads(q, k, v, 2 * d_o)
  for a, b in ((a_q, b_q), (a_k, b_k), (a_v, b_v)):
    assert (2 * a.float() - b.float()).abs().max().item() < 2e-2 * max(1.0, b.float().abs().max().item())


def _torch_ref_grads(q, k, v, d_o, causal):
  q32, k32, v32 = (t.float().detach().requires_grad_(True) for t in (q, k, v))
  g = q.size(1) // k.size(1)
  kk, vv = k32.repeat_interleave(g, dim=1), v32.repeat_interleave(g, dim=1)
  s = (q32 @ kk.transpose(-1, -2)) * q.size(-1) ** -0.5
  if causal:
    Nq, Nkv = q.size(2), k.size(2)
    m = torch.arange(Nkv, device=q.device)[None, :] <= (torch.arange(Nq, device=q.device)[:, None] + (Nkv - Nq))
    s = s.masked_fill(~m, float("-inf"))
  o = torch.softmax(s, dim=-1) @ vv
  o.backward(d_o.float())
  return q32.grad, k32.grad, v32.grad


def test_c3_full_size_gqa_causal_backward():
  """BASELINE config 3: Hq=32 Hkv=8 N=4096 D=512 causal bf16, forward + backward. Reference grads in
  fp32 torch on the GPU for two KV heads (8 query heads)."""
  q, k, v, d_o = _mk(1, 32, 8, 4096, 4096, 512, torch.bfloat16, seed=42)
  _, dq, dk, dv = _grads(q, k, v, d_o, is_causal=True, enable_gqa=True)
  for hk in (0, 7):
    hs = slice(4 * hk, 4 * hk + 4)
    rq, rk, rv = _torch_ref_grads(q[:, hs], k[:, hk:hk + 1], v[:, hk:hk + 1], d_o[:, hs], True)
    for got, want, name in ((dq[:, hs], rq, "dQ"), (dk[:, hk:hk + 1], rk, "dK"), (dv[:, hk:hk + 1], rv, "dV")):
      err = (got.float() - want).abs().max().item()
      bound = 1e-1 * max(1.0, want.abs().max().item())  # tests/test_ffpa_bwd.py:975-985 (causal bf16, D>=512)
      assert err < bound, f"{name} kv-head {hk}: {err:.3e} >= {bound:.3e}"
      cos = torch.nn.functional.cosine_similarity(got.float().flatten(), want.flatten(), dim=0).item()
      assert cos > 0.999, f"{name} kv-head {hk}: cosine {cos}"  # tests/test_ffpa_cute_sm100.py:822-842


@pytest.mark.parametrize("kind", ["add_full_f32", "add_bcast_qdtype", "key_bias", "bool"])
def test_backward_with_attn_mask_and_bias_grad(kind):
  """mask-grad path (/root/reference/tests/test_ffpa_bwd.py mask-grad cases; math
  triton/_ffpa_bwd.py:692-855): dQ/dK/dV with an additive bias, and dBias = P*(dP - delta)
  reduced over the bias' broadcast dims."""
  import ffpa_attn

  B, H, Nq, Nkv, D = 2, 2, 130, 200, 128
  q, k, v, d_o = _mk(B, H, H, Nq, Nkv, D, torch.bfloat16, seed=5)
  g = torch.Generator().manual_seed(9)
  if kind == "add_full_f32":
    bias = torch.randn(B, H, Nq, Nkv, generator=g)
  elif kind == "add_bcast_qdtype":
    bias = torch.randn(1, H, Nq, Nkv, generator=g).to(torch.bfloat16)
  elif kind == "key_bias":
    bias = torch.randn(B, 1, 1, Nkv, generator=g)
  else:
    bias = torch.rand(1, 1, Nq, Nkv, generator=g) > 0.3
    bias[..., 0] = True
  bias = bias.to(DEV)
  want_grad = bias.dtype != torch.bool
  if want_grad:
    bias.requires_grad_(True)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, attn_mask=bias)
  out.backward(d_o)
  torch.cuda.synchronize()
  add = torch.where(bias, 0.0, float("-inf")) if bias.dtype == torch.bool else bias.detach().float()
  rq, rk, rv, rds = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), bias=add.cpu().double().numpy())
  _cmp(qg.grad, rq, 5e-2, "dQ")
  _cmp(kg.grad, rk, 5e-2, "dK")
  _cmp(vg.grad, rv, 5e-2, "dV")
  if want_grad:
    red = tuple(i for i in range(4) if bias.size(i) == 1 and rds.shape[i] != 1)
    want = rds.sum(axis=red, keepdims=True) if red else rds
    assert bias.grad is not None and bias.grad.shape == bias.shape
    _cmp(bias.grad, want, 5e-2, "dBias")


@pytest.mark.parametrize("p_drop", [0.2])
@pytest.mark.parametrize("causal", [False, True])
def test_backward_dropout_replay(p_drop, causal):
  """Dropout replay in the backward (tests/test_ffpa_bwd.py:790-897): the Philox mask of the
  forward is regenerated from the saved (seed, offset)."""
  import ffpa_attn

  q, k, v, d_o = _mk(1, 2, 2, 200, 264, 256, torch.float16, seed=6)
  torch.cuda.manual_seed(1234)
  seed = int(torch.cuda.initial_seed())
  offset = int(torch.cuda._get_rng_state_offset())
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, dropout_p=p_drop, is_causal=causal)
  out.backward(d_o)
  torch.cuda.synchronize()
  ro, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=causal, dropout_p=p_drop, philox_seed=seed, philox_offset=offset)
  assert np.abs(out.detach().float().cpu().numpy() - ro).max() < 4e-2
  rq, rk, rv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), causal=causal, dropout_p=p_drop,
                                    philox_seed=seed, philox_offset=offset)
  _cmp(qg.grad, rq, 2e-2, "dQ")
  _cmp(kg.grad, rk, 2e-2, "dK")
  _cmp(vg.grad, rv, 2e-2, "dV")


@pytest.mark.parametrize("D", [520, 576, 640, 768, 896, 1024])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_backward_large_headdims(D, dtype):
  """head_dim in (512, 1024] (north star: headdim 64-1024 forward and backward): two output-slab
  passes per item, A2 streamed through the ring (csrc/ffpa_bwd_sm100.cuh, BwdCfg::LARGE)."""
  q, k, v, d_o = _mk(1, 2, 2, 256, 384, D, dtype, seed=10)
  _check_against_oracle(q, k, v, d_o, False, dtype)


@pytest.mark.parametrize("D", [640, 1024])
@pytest.mark.parametrize("Nq,Nkv", [(130, 257), (3, 300), (500, 1000)])
def test_backward_large_headdims_causal_gqa_tails(D, Nq, Nkv):
  q, k, v, d_o = _mk(2, 4, 2, Nq, Nkv, D, torch.bfloat16, seed=11)
  _check_against_oracle(q, k, v, d_o, True, torch.bfloat16, enable_gqa=True, tol=1e-1)
  _check_against_oracle(q, k, v, d_o, False, torch.bfloat16, enable_gqa=True)


def test_backward_large_headdim_bias_and_dropout():
  """GENERAL variants at D=768: additive bias + dBias, then dropout replay."""
  import ffpa_attn

  B, H, Nq, Nkv, D = 1, 2, 130, 200, 768
  q, k, v, d_o = _mk(B, H, H, Nq, Nkv, D, torch.bfloat16, seed=12)
  bias = torch.randn(B, H, Nq, Nkv, generator=torch.Generator().manual_seed(3)).to(DEV).requires_grad_(True)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  ffpa_attn.ffpa_attn_func(qg, kg, vg, attn_mask=bias).backward(d_o)
  torch.cuda.synchronize()
  rq, rk, rv, rds = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), bias=bias.detach().cpu().double().numpy())
  for got, want, name in ((qg.grad, rq, "dQ"), (kg.grad, rk, "dK"), (vg.grad, rv, "dV"), (bias.grad, rds, "dBias")):
    _cmp(got, want, 5e-2, name)
  torch.cuda.manual_seed(4321)
  seed, offset = int(torch.cuda.initial_seed()), int(torch.cuda._get_rng_state_offset())
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  ffpa_attn.ffpa_attn_func(qg, kg, vg, dropout_p=0.1).backward(d_o)
  torch.cuda.synchronize()
  rq, rk, rv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), dropout_p=0.1, philox_seed=seed, philox_offset=offset)
  for got, want, name in ((qg.grad, rq, "dQ"), (kg.grad, rk, "dK"), (vg.grad, rv, "dV")):
    _cmp(got, want, 5e-2, name)


def test_backward_large_headdim_full_length_sampled_head():
  """BASELINE config 5's largest head dim in the backward: N=4096 D=1024 bf16 against fp32 torch on the GPU."""
  q, k, v, d_o = _mk(1, 2, 2, 4096, 4096, 1024, torch.bfloat16, seed=13)
  _, dq, dk, dv = _grads(q, k, v, d_o, is_causal=True)
  rq, rk, rv = _torch_ref_grads(q[:, 1:], k[:, 1:], v[:, 1:], d_o[:, 1:], True)
  for got, want, name in ((dq[:, 1:], rq, "dQ"), (dk[:, 1:], rk, "dK"), (dv[:, 1:], rv, "dV")):
    err = (got.float() - want).abs().max().item()
    assert err < 1e-1 * max(1.0, want.abs().max().item()), f"{name}: {err}"
    cos = torch.nn.functional.cosine_similarity(got.float().flatten(), want.flatten(), dim=0).item()
    assert cos > 0.999, f"{name}: cosine {cos}"


def _raw_backward(q, k, v, d_o, causal, min_workspace):
  """forward + backward through the native binding, choosing the backward path by workspace size"""
  import ffpa_attn
  from ffpa_attn import _C
  from ffpa_attn.cuda import _ffpa_attn_forward_cuda

  scale = q.size(-1) ** -0.5
  out, lse = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), scale)
  dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
  n0 = ffpa_attn._C.launch_count()
  _C.refresh_env()
  _C.ffpa_attn_backward_ex(q, k, v, out, lse, d_o, dq, dk, dv, 0, int(causal), scale, min_workspace=min_workspace)
  torch.cuda.synchronize()
  return dq, dk, dv, ffpa_attn._C.launch_count() - n0


@pytest.mark.parametrize("D", [384, 448, 512, 520, 640, 1024])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 4, 2, 300, 300), (2, 2, 2, 130, 700), (1, 3, 1, 1000, 1000), (1, 2, 2, 5, 257)])
def test_backward_stash_path_matches_recompute_path_and_oracle(D, causal, shape):
  """Head dims 384..1024: the dQ kernel stashes P / dS tiles (on the first slab pass when D > 512) and dK / dV run as GEMMs over them
  (csrc/ffpa_bwd_gemm_sm100.cuh, 3 launches after the preprocess); with the minimum workspace the three
  recompute kernels run. Both must agree with each other and with the oracle (GQA, tails, odd tile counts,
  bottom-right causal with Nkv > Nq)."""
  B, Hq, Hkv, Nq, Nkv = shape
  q, k, v, d_o = _mk(B, Hq, Hkv, Nq, Nkv, D, torch.bfloat16, seed=21)
  sq, sk, sv, n_stash = _raw_backward(q, k, v, d_o, causal, False)
  rq, rk, rv, n_rec = _raw_backward(q, k, v, d_o, causal, True)
  assert n_stash == 4 and n_rec >= 4   # preprocess + dQ + 2 GEMMs  vs  preprocess + dQ + dK + dV (+ converts)
  for a, b_, name in ((sq, rq, "dQ"), (sk, rk, "dK"), (sv, rv, "dV")):
    assert torch.isfinite(a.float()).all(), name
    err = (a.float() - b_.float()).abs().max().item()
    assert err < 2e-2 * max(1.0, b_.float().abs().max().item()), f"{name}: stash vs recompute {err}"
  wq, wk, wv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), causal=causal)
  tol = 1e-1 if causal else 5e-2
  _cmp(sq, wq, tol, "dQ")
  _cmp(sk, wk, tol, "dK")
  _cmp(sv, wv, tol, "dV")


def test_backward_stash_chunked_by_memory_cap(monkeypatch):
  """FFPA_BWD_STASH_MAX_GB bounds the score buffers: a larger problem runs as (batch element, KV-head range)
  chunks through the same scratch. Results must equal the unchunked run bit for bit (same kernels, same
  per-head arithmetic) and the workspace must respect the cap."""
  from ffpa_attn import _C

  B, Hq, Hkv, N, D = 2, 16, 8, 2048, 384
  q, k, v, d_o = _mk(B, Hq, Hkv, N, N, D, torch.bfloat16, seed=23)
  full_q, full_k, full_v, n_full = _raw_backward(q, k, v, d_o, True, False)
  assert n_full == 4
  import ctypes

  import capi
  lib = capi.load()
  sizes = capi.bwd_sizes(B, Hq, Hkv, N, N, D)
  ws_full = int(lib.ffpa_b200_bwd_workspace_bytes_p(ctypes.byref(sizes), 1 << 40))
  monkeypatch.setenv("FFPA_BWD_STASH_MAX_GB", "0.15")   # the binder's upper bound on the scratch it offers
  ws_cap = int(lib.ffpa_b200_bwd_workspace_bytes_p(ctypes.byref(sizes), int(0.15 * 2 ** 30)))
  assert ws_cap <= 0.15 * 2 ** 30 < ws_full
  cq, ck, cv, n_chunked = _raw_backward(q, k, v, d_o, True, False)
  # 2 batch elements x ceil(8 / hc) KV-head chunks, 4 launches each (hc = the largest head count whose buffers fit)
  assert n_chunked % (4 * B) == 0 and n_chunked // (4 * B) >= 2, n_chunked
  for a, b_ in ((cq, full_q), (ck, full_k), (cv, full_v)):
    assert torch.equal(a, b_)
  monkeypatch.setenv("FFPA_BWD_STASH_MAX_GB", "0.001")   # not even one KV head fits: recompute kernels
  rq, rk, rv, _ = _raw_backward(q, k, v, d_o, True, False)
  for a, b_ in ((rq, full_q), (rk, full_k), (rv, full_v)):
    assert (a.float() - b_.float()).abs().max().item() < 2e-2 * max(1.0, b_.float().abs().max().item())


def test_backward_stash_path_with_bias_and_dropout():
  """GENERAL variant of the stashing dQ kernel: P_drop (dropout applied) feeds dV, dS feeds dK."""
  import ffpa_attn

  B, H, Nq, Nkv, D = 1, 2, 200, 264, 512
  q, k, v, d_o = _mk(B, H, H, Nq, Nkv, D, torch.float16, seed=22)
  bias = torch.randn(B, 1, Nq, Nkv, generator=torch.Generator().manual_seed(5)).to(DEV).requires_grad_(True)
  torch.cuda.manual_seed(99)
  seed, offset = int(torch.cuda.initial_seed()), int(torch.cuda._get_rng_state_offset())
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  ffpa_attn.ffpa_attn_func(qg, kg, vg, attn_mask=bias, dropout_p=0.15).backward(d_o)
  torch.cuda.synchronize()
  rq, rk, rv, rds = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), bias=bias.detach().cpu().double().numpy(),
                                      dropout_p=0.15, philox_seed=seed, philox_offset=offset)
  for got, want, name in ((qg.grad, rq, "dQ"), (kg.grad, rk, "dK"), (vg.grad, rv, "dV")):
    _cmp(got, want, 2e-2, name)
  _cmp(bias.grad, rds.sum(axis=1, keepdims=True), 2e-2, "dBias")


def test_backward_rejects_unsupported():
  import ffpa_attn

  q = torch.randn(1, 1, 128, 1032, dtype=torch.bfloat16, device=DEV)
  with pytest.raises((NotImplementedError, ValueError, RuntimeError)):
    ffpa_attn.ffpa_attn_func(q, q, q)


@pytest.mark.parametrize("Nq", [1, 2, 3, 4, 7])
def test_backward_decode_like_query_lengths(Nq):  # tests/test_ffpa_bwd.py:637-640
  q, k, v, d_o = _mk(2, 4, 2, Nq, 700, 512, torch.bfloat16, seed=8)
  _check_against_oracle(q, k, v, d_o, False, torch.bfloat16, enable_gqa=True)
  _check_against_oracle(q, k, v, d_o, True, torch.bfloat16, enable_gqa=True, tol=1e-1)


def test_backward_long_sequence_sampled_head():
  """tests/test_ffpa_bwd.py:900-907 go up to (1, 16, 16384, 512); here one head of N=8192 D=512 against an
  fp32 torch restatement on the GPU (tolerance of the reference for large bf16: 1e-1)."""
  q, k, v, d_o = _mk(1, 2, 2, 8192, 8192, 512, torch.bfloat16, seed=9)
  _, dq, dk, dv = _grads(q, k, v, d_o)
  rq, rk, rv = _torch_ref_grads(q[:, :1], k[:, :1], v[:, :1], d_o[:, :1], False)
  for got, want, name in ((dq[:, :1], rq, "dQ"), (dk[:, :1], rk, "dK"), (dv[:, :1], rv, "dV")):
    err = (got.float() - want).abs().max().item()
    assert err < 1e-1 * max(1.0, want.abs().max().item()), f"{name}: {err}"
    cos = torch.nn.functional.cosine_similarity(got.float().flatten(), want.flatten(), dim=0).item()
    assert cos > 0.999, f"{name}: cosine {cos}"


def test_bias_broadcast_over_keys_is_a_softmax_invariant():
  """An additive bias of shape [1, 1, Nq, 1] (broadcast over KEYS, stride 0 on the last dim) shifts every score of a
  row by the same constant: O, dQ, dK, dV must equal the unbiased run, the LSE moves by the constant, and the bias
  gradient -- sum_k dS, reduced in the dQ kernel -- is analytically zero. Exercises the stride-0 key dim of the bias
  in forward and backward and the in-kernel dBias reduction over keys."""
  import ffpa_attn
  from ffpa_attn.cuda import _ffpa_attn_forward_cuda

  q, k, v, d_o = _mk(1, 2, 2, 200, 300, 128, torch.bfloat16, seed=41)
  bias = (torch.randn(1, 1, 200, 1, generator=torch.Generator().manual_seed(3)) * 3).to(DEV).requires_grad_(True)
  qg, kg, vg = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, attn_mask=bias)
  out.backward(d_o)
  q0, k0, v0 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
  out0 = ffpa_attn.ffpa_attn_func(q0, k0, v0)
  out0.backward(d_o)
  assert (out.float() - out0.float()).abs().max().item() < 4e-3
  for a, b_, name in ((qg.grad, q0.grad, "dQ"), (kg.grad, k0.grad, "dK"), (vg.grad, v0.grad, "dV")):
    assert (a.float() - b_.float()).abs().max().item() < 2e-2 * max(1.0, b_.float().abs().max().item()), name
  assert bias.grad.shape == bias.shape
  scale_ds = float((qg.grad.float().abs().max()))
  assert bias.grad.abs().max().item() < 2e-2 * max(1.0, scale_ds), bias.grad.abs().max().item()
  _, lse_b = _ffpa_attn_forward_cuda(q, k, v, None, bias.detach(), 0, 1, 0, 128 ** -0.5)
  _, lse_0 = _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, 0, 128 ** -0.5)
  assert (lse_b - lse_0 - bias.detach()[0, 0, :, 0][None, None]).abs().max().item() < 2e-3
