"""Full-tensor parity at the BASELINE sizes and the reference's own SM100-grade bounds.

Round-1 checked the big configs on a handful of rows; here EVERY output element of C2 / C3 / C4 / C5 is compared
with a plain fp32 restatement evaluated on the GPU head by head (fp32 matmuls with TF32 off; the restatement itself
is checked against oracle/attention_oracle.py on a small case below), and the D = 512 kernels are held to the
bounds the reference holds ITS sm_100 kernels to (/root/reference/tests/test_ffpa_cute_sm100.py:875-890 shapes,
:916-920 `max|err| <= max|V| * 2^-8` and LSE < 2e-4, :822-842 / :978-1023 backward cosine > 0.999 and
rel-max < 3e-2 in aggregate and per head, :1026-1050 single-key KAT).  Dropout is compared with torch's
SDPA-efficient kernel under the same seed, the reference's own dropout test
(/root/reference/tests/test_ffpa_fwd.py:343-414)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"
_BF16_OPERAND_ULP = 2 ** -8     # test_ffpa_cute_sm100.py:916
_LSE_ABS_TOL = 2e-4             # test_ffpa_cute_sm100.py:920


@pytest.fixture(autouse=True)
def _fp32_matmuls():
  old = torch.backends.cuda.matmul.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = False
  yield
  torch.backends.cuda.matmul.allow_tf32 = old


def _mk(B, Hq, Hkv, Nq, Nkv, D, dtype=torch.bfloat16, seed=0, amp=1.0, with_do=False):
  g = torch.Generator(device=DEV).manual_seed(seed)
  q = (torch.randn(B, Hq, Nq, D, generator=g, device=DEV) * amp).to(dtype)
  k = (torch.randn(B, Hkv, Nkv, D, generator=g, device=DEV) * amp).to(dtype)
  v = (torch.randn(B, Hkv, Nkv, D, generator=g, device=DEV) * amp).to(dtype)
  if with_do:
    return q, k, v, torch.randn(B, Hq, Nq, D, generator=g, device=DEV).to(dtype)
  return q, k, v


def _mask(nq, nkv):
  i = torch.arange(nq, device=DEV)[:, None]
  j = torch.arange(nkv, device=DEV)[None, :]
  return j <= i + (nkv - nq)


def ref_head(q, k, v, causal, scale, d_o=None, row_chunk=4096):
  """fp32 restatement for ONE head on the GPU: q [Nq, D], k / v [Nkv, D] -> (O fp32, LSE fp32[, dQ, dK, dV]).
  Empty rows: O = 0, LSE = -inf (declared contract, not SDPA's backend-defined behaviour)."""
  qf, kf, vf = q.float(), k.float(), v.float()
  nq, nkv = qf.size(0), kf.size(0)
  O = torch.empty_like(qf)
  lse = torch.empty(nq, device=DEV, dtype=torch.float32)
  if d_o is not None:
    dq = torch.empty_like(qf)
    dk = torch.zeros_like(kf)
    dv = torch.zeros_like(vf)
    dof = d_o.float()
  for lo in range(0, nq, row_chunk):
    hi = min(lo + row_chunk, nq)
    s = (qf[lo:hi] @ kf.T) * scale
    if causal:
      j = torch.arange(nkv, device=DEV)[None, :]
      i = torch.arange(lo, hi, device=DEV)[:, None]
      s = s.masked_fill(j > i + (nkv - nq), float("-inf"))
    l = torch.logsumexp(s, dim=-1)
    empty = l == float("-inf")
    p = torch.exp(s - torch.where(empty, torch.zeros_like(l), l)[:, None])
    p = p.masked_fill(empty[:, None], 0.0)
    o = p @ vf
    O[lo:hi] = o
    lse[lo:hi] = l
    if d_o is not None:
      dp = dof[lo:hi] @ vf.T
      delta = (dof[lo:hi] * o).sum(-1, keepdim=True)
      ds = p * (dp - delta) * scale
      dq[lo:hi] = ds @ kf
      dk += ds.T @ qf[lo:hi]
      dv += p.T @ dof[lo:hi]
  if d_o is not None:
    return O, lse, dq, dk, dv
  return O, lse


def _cos(a, b):
  return F.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _run_fwd(q, k, v, causal, backend=None):
  import ffpa_attn
  from ffpa_attn.cuda import _ffpa_attn_forward_cuda

  if backend is not None:
    out = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=causal, enable_gqa=q.size(1) != k.size(1), forward_backend=backend)
    return out, None
  ffpa_attn.set_cuda_backend_impl(ffpa_attn.CudaBackendImpl.AUTO)
  return _ffpa_attn_forward_cuda(q, k, v, None, None, 0, 1, int(causal), 1.0 / math.sqrt(q.size(-1)))


def _check_fwd_all_heads(q, k, v, causal, o, lse, o_tol, lse_tol, cos_min, tag):
  B, Hq, Nq, D = q.shape
  g = Hq // k.size(1)
  scale = 1.0 / math.sqrt(D)
  worst = {"o": 0.0, "lse": 0.0, "cos": 1.0}
  for b in range(B):
    for h in range(Hq):
      ro, rl = ref_head(q[b, h], k[b, h // g], v[b, h // g], causal, scale)
      finite = torch.isfinite(rl)
      if lse is not None:
        assert torch.all(lse[b, h][~finite] == float("-inf")), (tag, b, h)
        e_l = (lse[b, h][finite] - rl[finite]).abs().max().item()
        assert e_l < lse_tol, (tag, "LSE", b, h, e_l)
        worst["lse"] = max(worst["lse"], e_l)
      assert torch.all(o[b, h][~finite] == 0), (tag, b, h)
      e_o = (o[b, h].float() - ro).abs().max().item()
      assert e_o <= o_tol, (tag, "O", b, h, e_o, o_tol)
      c = _cos(o[b, h], ro)
      assert c > cos_min, (tag, "cosine", b, h, c)
      worst["o"], worst["cos"] = max(worst["o"], e_o), min(worst["cos"], c)
  return worst


def test_gpu_restatement_matches_the_pinned_oracle():
  """The per-head fp32 GPU restatement used below equals oracle/attention_oracle.py (pinned on the reference's
  goldens) on a small causal GQA case, forward and backward."""
  q, k, v, d_o = _mk(1, 2, 1, 130, 257, 64, torch.float16, seed=5, with_do=True)
  ro, rl = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=True)
  wq, wk, wv, _ = orc.attention_bwd(q.cpu(), k.cpu(), v.cpu(), d_o.cpu(), causal=True)
  dk_sum, dv_sum = 0, 0
  for h in range(2):
    O, lse, dq, dk, dv = ref_head(q[0, h], k[0, 0], v[0, 0], True, 1 / 8.0, d_o[0, h], row_chunk=64)
    assert np.abs(O.cpu().numpy() - ro[0, h]).max() < 2e-6
    assert np.abs(lse.cpu().numpy() - rl[0, h]).max() < 2e-6
    assert np.abs(dq.cpu().numpy() - wq[0, h]).max() < 2e-5
    dk_sum, dv_sum = dk_sum + dk, dv_sum + dv
  assert np.abs(dk_sum.cpu().numpy() - wk[0, 0]).max() < 5e-5
  assert np.abs(dv_sum.cpu().numpy() - wv[0, 0]).max() < 5e-5


# ---------------------------------------------------------------------------------------------------------------
# the reference's SM100 shape sweep at D = 512, its bounds
# ---------------------------------------------------------------------------------------------------------------
_SHAPES = [
  (1, 256, 256, 2, 2, False), (1, 256, 256, 2, 2, True), (2, 128, 128, 4, 4, True), (1, 384, 384, 4, 2, True),
  (1, 200, 200, 2, 2, True), (1, 128, 320, 2, 2, True), (1, 256, 256, 8, 1, True), (1, 1024, 1024, 8, 8, True),
  (1, 2048, 2048, 8, 8, True), (1, 4096, 4096, 8, 8, True), (1, 2048, 2048, 8, 1, True), (1, 2048, 2048, 8, 8, False),
  (2, 1024, 1024, 4, 2, True),
]   # (1, 320, 128, ...) of the reference list has s_q > s_k with causal, which this API rejects (functional.py:838-842)


@pytest.mark.parametrize("b,s_q,s_k,h_q,h_kv,causal", _SHAPES)
def test_d512_forward_meets_reference_sm100_bounds(b, s_q, s_k, h_q, h_kv, causal):
  q, k, v = _mk(b, h_q, h_kv, s_q, s_k, 512, seed=s_q * 31 + s_k)
  o, lse = _run_fwd(q, k, v, causal)
  bound = v.float().abs().max().item() * _BF16_OPERAND_ULP
  _check_fwd_all_heads(q, k, v, causal, o, lse, bound, _LSE_ABS_TOL, 0.99999, "sm100-sweep")


@pytest.mark.parametrize("b,s_q,s_k,h_q,h_kv,causal", _SHAPES)
@pytest.mark.parametrize("min_ws", [False, True], ids=["stash", "recompute"])
def test_d512_backward_meets_reference_sm100_bounds(b, s_q, s_k, h_q, h_kv, causal, min_ws):
  import ffpa_attn

  q, k, v, d_o = _mk(b, h_q, h_kv, s_q, s_k, 512, seed=s_q * 31 + s_k, with_do=True)
  be = ffpa_attn.CUDABackend(bwd_min_workspace=min_ws)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, is_causal=causal, enable_gqa=h_q != h_kv, backend=be)
  out.backward(d_o)
  _check_bwd_all_heads(q, k, v, d_o, causal, qg.grad, kg.grad, vg.grad, "sm100-sweep")


def _check_bwd_all_heads(q, k, v, d_o, causal, dq, dk, dv, tag, cos_min=0.999, rel_tol=3e-2):
  B, Hq, Nq, D = q.shape
  Hkv = k.size(1)
  g = Hq // Hkv
  scale = 1.0 / math.sqrt(D)
  for b in range(B):
    for hk in range(Hkv):
      rk = torch.zeros(k.size(2), D, device=DEV)
      rv = torch.zeros(k.size(2), D, device=DEV)
      for h in range(hk * g, (hk + 1) * g):
        _, _, rq, dk_h, dv_h = ref_head(q[b, h], k[b, hk], v[b, hk], causal, scale, d_o[b, h])
        rk += dk_h
        rv += dv_h
        # rows with <= 1 visible key have an analytically zero dS row (test_ffpa_cute_sm100.py:1000-1012)
        keep = torch.ones(Nq, dtype=torch.bool, device=DEV)
        if causal:
          keep = (torch.arange(Nq, device=DEV) + (k.size(2) - Nq) + 1) > 1
          if (~keep).any():
            assert dq[b, h][~keep].abs().max() == 0, (tag, "dQ of single-key rows", b, h)
        _lane(dq[b, h][keep], rq[keep], (tag, "dQ", b, h), cos_min, rel_tol)
      _lane(dk[b, hk], rk, (tag, "dK", b, hk), cos_min, rel_tol)
      _lane(dv[b, hk], rv, (tag, "dV", b, hk), cos_min, rel_tol)


def _lane(got, want, where, cos_min, rel_tol):
  assert torch.isfinite(got).all(), where
  if want.abs().max() == 0:
    assert got.abs().max() == 0, where
    return
  assert got.abs().max() > 0, (*where, "no gradient at all")
  c = _cos(got, want)
  assert c > cos_min, (*where, "cosine", c)
  rel = ((got.float() - want).abs().max() / (want.abs().max() + 1e-30)).item()
  assert rel < rel_tol, (*where, "rel-max", rel)


def test_single_visible_key_backward_is_analytically_zero():
  """s_k == 1: dS == 0 exactly, so dK = dQ = 0 and dV = sum_s dO (test_ffpa_cute_sm100.py:1026-1050)."""
  import ffpa_attn

  q, k, v, d_o = _mk(1, 2, 2, 128, 1, 512, seed=17, with_do=True)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  ffpa_attn.ffpa_attn_func(qg, kg, vg).backward(d_o)
  assert kg.grad.abs().max() == 0 and qg.grad.abs().max() == 0
  assert _cos(vg.grad, d_o.float().sum(dim=2, keepdim=True)) > 0.999


# ---------------------------------------------------------------------------------------------------------------
# BASELINE configs, every element
# ---------------------------------------------------------------------------------------------------------------
def test_c2_forward_every_head_full_tensor():
  """BASELINE config 2 (B=1, H=32, N=8192, D=512, bf16): all 32 heads x 8192 rows x 512 columns + LSE."""
  q, k, v = _mk(1, 32, 32, 8192, 8192, 512, seed=42)
  o, lse = _run_fwd(q, k, v, False)
  bound = v.float().abs().max().item() * _BF16_OPERAND_ULP
  w = _check_fwd_all_heads(q, k, v, False, o, lse, bound, _LSE_ABS_TOL, 0.99999, "C2")
  assert w["o"] <= 1e-2   # the north-star bound (max-abs-err <= 1e-2 vs SDPA)


@pytest.mark.parametrize("min_ws", [False, True], ids=["stash", "recompute"])
def test_c2_backward_every_head_full_tensor(min_ws):
  import ffpa_attn

  q, k, v, d_o = _mk(1, 32, 32, 8192, 8192, 512, seed=43, with_do=True)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, backend=ffpa_attn.CUDABackend(bwd_min_workspace=min_ws))
  out.backward(d_o)
  del out
  _check_bwd_all_heads(q, k, v, d_o, False, qg.grad, kg.grad, vg.grad, "C2")


@pytest.mark.parametrize("min_ws", [False, True], ids=["stash", "recompute"])
def test_c3_gqa_causal_forward_backward_every_head(min_ws):
  """BASELINE config 3: Hq=32, Hkv=8, N=4096, D=512, causal, bf16."""
  import ffpa_attn

  q, k, v, d_o = _mk(1, 32, 8, 4096, 4096, 512, seed=44, with_do=True)
  o, lse = _run_fwd(q, k, v, True)
  bound = v.float().abs().max().item() * _BF16_OPERAND_ULP
  _check_fwd_all_heads(q, k, v, True, o, lse, bound, _LSE_ABS_TOL, 0.99999, "C3")
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  out = ffpa_attn.ffpa_attn_func(qg, kg, vg, is_causal=True, enable_gqa=True,
                                 backend=ffpa_attn.CUDABackend(bwd_min_workspace=min_ws))
  assert torch.equal(out, o)
  out.backward(d_o)
  _check_bwd_all_heads(q, k, v, d_o, True, qg.grad, kg.grad, vg.grad, "C3")


def test_c4_fp8_every_batch_and_head_full_tensor():
  """BASELINE config 4 (B=4, H=32, N=8192, D=256, FP8): all 128 (b, h) pairs against exact fp32 attention at the
  reference's FP8 tolerance (inputs randn * 0.5, O 4e-2, /root/reference/tests/test_ffpa_fp8.py:63-71)."""
  import ffpa_attn

  q, k, v = _mk(4, 32, 32, 8192, 8192, 256, seed=45, amp=0.5)
  o, _ = _run_fwd(q, k, v, False, backend=ffpa_attn.CUDABackend(enable_fp8=True))
  w = _check_fwd_all_heads(q, k, v, False, o, None, 4e-2, None, 0.999, "C4-fp8")
  assert w["o"] < 4e-2


@pytest.mark.parametrize("D", [320, 768, 1024])
@pytest.mark.parametrize("replay", [True, False], ids=["replay", "two-pass"])
def test_c5_headdim_sweep_forward_every_head(D, replay, monkeypatch):
  """BASELINE config 5 at full size (H=32, N=8192): D = 320 / 768 / 1024 (512 is C2). D = 1024 runs both the
  replay path (4.3 GB P stash + GEMM for the second O slab) and the two-pass path."""
  import ffpa_attn

  if D != 1024 and not replay:
    pytest.skip("the replay / two-pass split only exists above D = 768")
  if not replay:
    monkeypatch.setenv("FFPA_FWD_REPLAY", "0")
  ffpa_attn._C.refresh_env()
  q, k, v = _mk(1, 32, 32, 8192, 8192, D, seed=46 + D)
  n0 = ffpa_attn._C.launch_count()
  o, lse = _run_fwd(q, k, v, False)
  torch.cuda.synchronize()
  if D == 1024:
    # replay: (softmax pass + GEMM) per chunk -- the 4.3 GB stash is cut into head chunks that fit the default
    # 2.5 GiB scratch bound (FFPA_FWD_REPLAY_MAX_GB); two-pass: one launch
    n = ffpa_attn._C.launch_count() - n0
    assert (n >= 2 and n % 2 == 0) if replay else n == 1, n
  bound = v.float().abs().max().item() * _BF16_OPERAND_ULP
  _check_fwd_all_heads(q, k, v, False, o, lse, bound, _LSE_ABS_TOL, 0.99999, f"C5-D{D}")


@pytest.mark.parametrize("D", [768, 1024])
def test_c5_large_headdim_backward_full_size(D):
  """D = 768 / 1024 backward at H=32, N=8192 (stash path over two output slabs): 8 of the 32 heads are compared
  element-wise (the fp32 restatement of one D = 1024 head costs ~0.7 TFLOP), every head is checked finite and
  non-zero."""
  import ffpa_attn

  q, k, v, d_o = _mk(1, 32, 32, 8192, 8192, D, seed=50 + D, with_do=True)
  qg, kg, vg = (t.clone().requires_grad_(True) for t in (q, k, v))
  ffpa_attn.ffpa_attn_func(qg, kg, vg).backward(d_o)
  scale = 1.0 / math.sqrt(D)
  for t in (qg.grad, kg.grad, vg.grad):
    assert torch.isfinite(t).all()
    assert (t.float().abs().amax(dim=(0, 2, 3)) > 0).all()
  for h in range(0, 32, 4):
    _, _, rq, rk, rv = ref_head(q[0, h], k[0, h], v[0, h], False, scale, d_o[0, h])
    _lane(qg.grad[0, h], rq, ("dQ", D, h), 0.999, 3e-2)
    _lane(kg.grad[0, h], rk, ("dK", D, h), 0.999, 3e-2)
    _lane(vg.grad[0, h], rv, ("dV", D, h), 0.999, 3e-2)


# ---------------------------------------------------------------------------------------------------------------
# dropout vs torch's SDPA-efficient kernel under the same seed (the reference's own test)
# ---------------------------------------------------------------------------------------------------------------
def _sdpa_efficient(q, k, v, **kw):
  from torch.nn.attention import SDPBackend, sdpa_kernel

  try:
    with sdpa_kernel(SDPBackend.EFFICIENT_ATTENTION):
      return F.scaled_dot_product_attention(q, k, v, **kw)
  except RuntimeError as e:
    pytest.skip(f"torch's SDPA-efficient kernel does not run this case on this box: {str(e).splitlines()[0][:200]}")


@pytest.mark.parametrize("shape,p,mask", [((1, 2, 512, 512, 512), 0.25, False), ((1, 2, 512, 512, 320), 0.2, True),
                                          ((1, 2, 1, 4096, 512), 0.2, False), ((2, 3, 300, 700, 128), 0.1, False)])
def test_dropout_matches_sdpa_efficient_same_seed(shape, p, mask):
  """test_ffpa_fwd.py:356-414: torch.manual_seed(s); ours; torch.manual_seed(s); SDPA-efficient; tol 4e-2. The two
  draw the same Philox stream (seed, offset reserved from the CUDA generator; element index ((b Hq + h) Nq + q) Nkv + k)."""
  import ffpa_attn

  B, H, Nq, Nkv, D = shape
  q, k, v = _mk(B, H, H, Nq, Nkv, D, torch.float16, seed=3)
  attn_mask = (torch.randn(1, 1, 1, Nkv, device=DEV, dtype=q.dtype) * 0.125) if mask else None
  kw = dict(dropout_p=p, scale=1.0 / math.sqrt(D))
  torch.manual_seed(0)
  out = ffpa_attn.ffpa_attn_func(q, k, v, attn_mask=attn_mask, **kw)
  torch.manual_seed(0)
  ref = _sdpa_efficient(q, k, v, attn_mask=attn_mask, **kw)
  torch.testing.assert_close(out, ref, atol=4e-2, rtol=4e-2)
  # and the masks are the same mask, not merely similar statistics: without dropout the outputs differ visibly
  torch.manual_seed(0)
  plain = ffpa_attn.ffpa_attn_func(q, k, v, attn_mask=attn_mask, scale=kw["scale"])
  assert (plain.float() - ref.float()).abs().max() > 5 * (out.float() - ref.float()).abs().max()
