"""GPU parity tests (run with -m gpu on a B200). Every call goes through the public API ->
torch.ops.ffpa_attn._fwd_cuda -> ffpa_attn._C -> the C ABI -> sm_100a kernels, and is compared with
the CPU oracle (oracle/attention_oracle.py) on the same seeded inputs, with the committed golden
fixtures produced by the reference package, and -- at full size -- through sampled rows and
size-independent properties. Shapes/tolerances follow the reference's own test lists
(/root/reference/tests/test_ffpa_fwd.py:32-45,106-113,1086-1120,1158-1173,1226-1249;
 /root/reference/tests/test_ffpa_cute_sm100.py:916-975)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _tol(dtype):
  # tests/test_ffpa_fwd.py:106-113
  return 2e-2 if dtype == torch.bfloat16 else 1e-2


def _mk(B, Hq, Hkv, Nq, Nkv, D, dtype, seed=0, amp=1.0):
  torch.manual_seed(seed)
  q = (torch.randn(B, Hq, Nq, D) * amp).to(dtype).to(DEV)
  k = (torch.randn(B, Hkv, Nkv, D) * amp).to(dtype).to(DEV)
  v = (torch.randn(B, Hkv, Nkv, D) * amp).to(dtype).to(DEV)
  return q, k, v


def _run(q, k, v, **kw):
  import ffpa_attn

  n0 = ffpa_attn._C.launch_count()
  out = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
  torch.cuda.synchronize()
  assert ffpa_attn._C.launch_count() > n0, "native kernel did not launch"
  return out


def _lse(q, k, v, causal=False, bias=None, scale=None):
  import ffpa_attn.cuda as fc

  if bias is None:
    bias = q.new_empty(0)
  scale = scale if scale is not None else q.size(-1) ** -0.5
  o, lse = torch.ops.ffpa_attn._fwd_cuda(q, k, v, bias, 0, 1, int(causal), scale, 0.0, 0, 0, True, False,
                                         0, 0, 0, 1, 0, False, 256, False, 256)
  torch.cuda.synchronize()
  return o, lse


def _check(out, ref, tol, what=""):
  err = float(np.abs(out.float().cpu().numpy() - ref).max())
  assert np.isfinite(out.float().cpu().numpy()).all(), f"{what}: non-finite output"
  assert err < tol, f"{what}: max-abs-err {err:.3e} >= {tol}"
  return err


# --------------------------------------------------------------------------------------------
# golden fixtures from the reference package
# --------------------------------------------------------------------------------------------
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_forward_matches_reference_golden(path):
  z = np.load(path)
  B, Hq, Hkv, Nq, Nkv, D, is_bf16, causal = [int(x) for x in z["meta"]]
  dt = torch.bfloat16 if is_bf16 else torch.float16
  t = lambda n, s: torch.from_numpy(z[n].view(np.int16).copy()).view(dt).reshape(s).to(DEV)  # noqa: E731
  q, k, v = t("q", (B, Hq, Nq, D)), t("k", (B, Hkv, Nkv, D)), t("v", (B, Hkv, Nkv, D))
  mask = torch.from_numpy(z["mask"]).to(DEV) if "mask" in z.files else None
  out = _run(q, k, v, attn_mask=mask, is_causal=bool(causal), enable_gqa=Hq != Hkv)
  ref = torch.from_numpy(z["o"].view(np.int16).copy()).view(dt).reshape(B, Hq, Nq, D).double().numpy()
  _check(out, ref, _tol(dt), "golden")
  if "o_f32" in z.files:  # fp32 run of the reference: only our own output rounding remains
    _check(out, z["o_f32"].astype(np.float64), _tol(dt) / 2, "golden-fp32")


# --------------------------------------------------------------------------------------------
# oracle parity: shapes from the reference lists, sized so the oracle runs in seconds
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [64, 128, 192, 256, 320, 384, 448, 512])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_self_attention_headdims(D, dtype):
  q, k, v = _mk(1, 2, 2, 512, 512, D, dtype)
  out = _run(q, k, v)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  _check(out, ref, _tol(dtype), f"D={D}")


@pytest.mark.parametrize("D", [72, 104, 200, 328, 504])
def test_headdim_multiple_of_8_not_64(D):
  q, k, v = _mk(1, 2, 2, 257, 300, D, torch.bfloat16)
  out = _run(q, k, v)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  _check(out, ref, 2e-2, f"D={D}")


@pytest.mark.parametrize("N", [1, 17, 33, 63, 65, 100, 127, 129, 200, 1000, 2047])
def test_boundary_seqlens(N):  # tests/test_ffpa_fwd.py:1086-1093
  q, k, v = _mk(1, 2, 2, N, N, 320, torch.bfloat16)
  out = _run(q, k, v)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  _check(out, ref, 2e-2, f"N={N}")


@pytest.mark.parametrize("Nq,Nkv", [(1, 512), (7, 1000), (128, 1024), (512, 777), (1000, 64), (300, 8191 // 8)])
def test_cross_attention(Nq, Nkv):  # tests/test_ffpa_fwd.py:1110-1120 (scaled)
  q, k, v = _mk(2, 2, 2, Nq, Nkv, 512, torch.bfloat16)
  out = _run(q, k, v)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  _check(out, ref, 2e-2)


@pytest.mark.parametrize("Hq,Hkv", [(8, 2), (8, 1), (6, 3), (4, 4)])
@pytest.mark.parametrize("causal", [False, True])
def test_gqa_mqa(Hq, Hkv, causal):  # tests/test_ffpa_fwd.py:1158-1173
  q, k, v = _mk(1, Hq, Hkv, 384, 384, 512, torch.bfloat16)
  out = _run(q, k, v, is_causal=causal, enable_gqa=True)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=causal)
  _check(out, ref, 2e-2)


@pytest.mark.parametrize("Nq,Nkv,D", [(512, 512, 512), (300, 300, 320), (256, 700, 512), (129, 1000, 256),
                                      (640, 640, 128), (1, 200, 512)])
def test_causal_bottom_right(Nq, Nkv, D):  # tests/test_ffpa_fwd.py:1226-1249, 1289-1320
  q, k, v = _mk(1, 2, 2, Nq, Nkv, D, torch.float16)
  out = _run(q, k, v, is_causal=True)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=True)
  _check(out, ref, 1e-2)
  if Nq == Nkv:  # equals SDPA's top-left convention only when square
    sd = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), is_causal=True)
    assert (out.float() - sd).abs().max().item() < 1e-2


@pytest.mark.parametrize("kind", ["bool2d", "bool4d", "add_f32", "add_qdtype", "key_padding", "per_head"])
def test_attn_mask(kind):  # tests/test_ffpa_fwd.py:225-247, 291-336
  B, H, Nq, Nkv, D = 2, 3, 200, 333, 320
  q, k, v = _mk(B, H, H, Nq, Nkv, D, torch.bfloat16)
  g = torch.Generator().manual_seed(5)
  if kind == "bool2d":
    m = torch.rand(Nq, Nkv, generator=g) > 0.4
    m[:, 0] = True
  elif kind == "bool4d":
    m = torch.rand(B, H, Nq, Nkv, generator=g) > 0.4
    m[..., 0] = True
  elif kind == "add_f32":
    m = torch.randn(B, 1, Nq, Nkv, generator=g)
  elif kind == "add_qdtype":
    m = torch.randn(1, H, Nq, Nkv, generator=g).to(torch.bfloat16)
  elif kind == "key_padding":
    m = torch.ones(B, 1, 1, Nkv, dtype=torch.bool)
    m[0, ..., 250:] = False
    m[1, ..., 100:] = False
  else:
    m = torch.randn(1, H, 1, 1, generator=g).expand(1, H, 1, Nkv).contiguous()
  out = _run(q, k, v, attn_mask=m.to(DEV))
  bias = torch.where(m, 0.0, float("-inf")) if m.dtype == torch.bool else m.float()
  if bias.dim() == 2:
    bias = bias.view(1, 1, Nq, Nkv)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=bias.double().numpy())
  _check(out, ref, 2e-2, kind)


def test_fully_masked_rows_give_zero_and_neg_inf_lse():  # tests/test_ffpa_cute_sm100.py:959-966
  q, k, v = _mk(1, 2, 2, 130, 200, 512, torch.bfloat16)
  bias = torch.zeros(1, 1, 130, 200)
  bias[0, 0, 5, :] = float("-inf")
  bias[0, 0, 129, :] = float("-inf")
  o, lse = _lse(q, k, v, bias=bias.to(DEV))
  assert (o[:, :, 5] == 0).all() and (o[:, :, 129] == 0).all()
  assert torch.isinf(lse[:, :, 5]).all() and (lse[:, :, 5] < 0).all()
  ref, lref = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=bias.double().numpy())
  _check(o, ref, 2e-2)
  fin = np.isfinite(lref)
  assert np.abs(lse.cpu().numpy()[fin] - lref[fin]).max() < 2e-4


@pytest.mark.parametrize("causal", [False, True])
def test_lse_matches_oracle(causal):  # LSE abs err, tests/test_ffpa_cute_sm100.py:968-975
  q, k, v = _mk(2, 2, 2, 300, 500, 512, torch.bfloat16)
  _, lse = _lse(q, k, v, causal=causal)
  _, lref = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=causal)
  assert lse.shape == (2, 2, 300) and lse.dtype == torch.float32
  assert np.abs(lse.cpu().numpy() - lref).max() < 2e-4


def test_explicit_scale_and_large_amplitude():
  q, k, v = _mk(1, 2, 2, 256, 256, 256, torch.float16, amp=3.0)
  out = _run(q, k, v, scale=0.05)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), scale=0.05)
  _check(out, ref, 3e-2)


def test_non_contiguous_inputs_honour_strides():
  torch.manual_seed(3)
  big = torch.randn(2, 4, 300, 3 * 512, dtype=torch.bfloat16, device=DEV)
  q, k, v = big[..., :512], big[..., 512:1024], big[..., 1024:]
  out = _run(q, k, v)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu())
  _check(out, ref, 2e-2)
  qt = torch.randn(2, 300, 4, 512, dtype=torch.bfloat16, device=DEV).transpose(1, 2)  # BNHD view
  out = _run(qt, k, v)
  ref, _ = orc.attention_fwd(qt.cpu(), k.cpu(), v.cpu())
  _check(out, ref, 2e-2)


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_matches_philox_oracle(p):
  """Dropout replay: same Philox stream as the oracle (and as SDPA-efficient / tl.randint4x),
  element index ((b*Hq+h)*Nq+q)*Nkv+k (csrc/cuffpa/native/prefill.cuh:424-452)."""
  q, k, v = _mk(1, 2, 2, 130, 203, 320, torch.bfloat16)
  seed, offset = 1234567, 40
  o, lse = torch.ops.ffpa_attn._fwd_cuda(q, k, v, q.new_empty(0), 0, 1, 0, 320 ** -0.5, p, seed, offset, True,
                                         False, 0, 0, 0, 1, 0, False, 256, False, 256)
  torch.cuda.synchronize()
  ref, lref = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), dropout_p=p, philox_seed=seed, philox_offset=offset)
  _check(o, ref, 4e-2, "dropout")  # tests/test_ffpa_fwd.py:343-414 tolerance
  assert np.abs(lse.cpu().numpy() - lref).max() < 2e-4


def test_dropout_through_public_api_advances_generator():
  import ffpa_attn

  q, k, v = _mk(1, 2, 2, 128, 128, 128, torch.bfloat16)
  torch.cuda.manual_seed(99)
  off0 = torch.cuda._get_rng_state_offset()
  a = ffpa_attn.ffpa_attn_func(q, k, v, dropout_p=0.3)
  off1 = torch.cuda._get_rng_state_offset()
  assert off1 - off0 == 1 * 2 * 128 * 128
  torch.cuda.manual_seed(99)
  b = ffpa_attn.ffpa_attn_func(q, k, v, dropout_p=0.3)
  assert torch.equal(a, b)
  c = ffpa_attn.ffpa_attn_func(q, k, v, dropout_p=0.3)
  assert not torch.equal(a, c)


# --------------------------------------------------------------------------------------------
# full-size configs (BASELINE.json): sampled rows + size-independent properties
# --------------------------------------------------------------------------------------------
def _sampled_rows_check(q, k, v, out, rows, heads, causal, tol):
  for (b, h) in heads:
    g = q.size(1) // k.size(1)
    qs = q[b:b + 1, h:h + 1, rows].cpu()
    ks, vs = k[b:b + 1, h // g:h // g + 1].cpu(), v[b:b + 1, h // g:h // g + 1].cpu()
    if causal:
      off = k.size(2) - q.size(2)
      bias = np.where(np.arange(k.size(2))[None, :] <= (np.array(rows)[:, None] + off), 0.0, -np.inf)[None, None]
      ref, _ = orc.attention_fwd(qs, ks, vs, bias=bias)
    else:
      ref, _ = orc.attention_fwd(qs, ks, vs)
    got = out[b, h, rows].float().cpu().numpy()
    err = np.abs(got - ref[0, 0]).max()
    assert err < tol, f"(b={b},h={h}) max-abs-err {err:.3e}"


def test_c2_full_size_self_attention_d512():
  """BASELINE config 2: B=1 H=32 N=8192 D=512 bf16; north-star bound max-abs-err <= 1e-2."""
  q, k, v = _mk(1, 32, 32, 8192, 8192, 512, torch.bfloat16, seed=42)
  out = _run(q, k, v)
  rows = [0, 1, 63, 64, 127, 128, 4095, 4096, 8000, 8191]
  _sampled_rows_check(q, k, v, out, rows, [(0, 0), (0, 17), (0, 31)], False, 1e-2)
  # property: permuting the keys/values together leaves the output unchanged (up to rounding)
  perm = torch.randperm(8192, device=DEV)
  out_p = _run(q[:, :2], k[:, :2, perm], v[:, :2, perm])
  assert (out_p.float() - out[:, :2].float()).abs().max().item() < 4e-3
  # property: V -> constant c gives O == c
  vc = torch.full_like(v[:, :2], 0.5)
  oc = _run(q[:, :2], k[:, :2], vc)
  assert (oc.float() - 0.5).abs().max().item() < 4e-3


def test_c3_full_size_gqa_causal_d512():
  """BASELINE config 3 forward: Hq=32 Hkv=8 N=4096 D=512 causal bf16."""
  q, k, v = _mk(1, 32, 8, 4096, 4096, 512, torch.bfloat16, seed=42)
  out = _run(q, k, v, is_causal=True, enable_gqa=True)
  rows = [0, 1, 64, 127, 128, 129, 2047, 2048, 4000, 4095]
  _sampled_rows_check(q, k, v, out, rows, [(0, 0), (0, 5), (0, 31)], True, 1e-2)
  # row 0 only sees key 0 -> O[0] == V[0]
  assert torch.equal(out[0, :, 0], v[0].repeat_interleave(4, dim=0)[:, 0])


@pytest.mark.parametrize("D", [320, 512])
def test_c5_headdim_sweep_sampled(D):
  q, k, v = _mk(1, 4, 4, 8192, 8192, D, torch.bfloat16, seed=42)
  out = _run(q, k, v)
  _sampled_rows_check(q, k, v, out, [0, 100, 4097, 8191], [(0, 0), (0, 3)], False, 1e-2)


# the 11 trailing fp8 / fp4 arguments of ffpa_attn_forward at their reference defaults (cuda/_ffpa_fwd.py:6-30)
_FP8_DEFAULTS = (True, False, 0, 0, 0, 1, 0, False, 256, False, 256)


def test_errors_raised_by_native_layer():
  import ffpa_attn._C as C

  q, k, v = _mk(1, 2, 2, 64, 64, 64, torch.bfloat16)
  o = torch.empty_like(q)
  lse = torch.empty(1, 2, 64, dtype=torch.float32, device=DEV)
  with pytest.raises(RuntimeError):  # causal with Nkv < Nq  (launch.cuh:79-129 TORCH_CHECK class)
    C.ffpa_attn_forward(q, k[:, :, :32], v[:, :, :32], q.new_empty(0), o, lse, 0, 1, 1, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  with pytest.raises(ValueError):  # dtype
    C.ffpa_attn_forward(q.float(), k.float(), v.float(), q.new_empty(0), o.float(), lse, 0, 1, 0, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  with pytest.raises(ValueError):  # acc = f16 does not exist on sm_100a (std::invalid_argument class, ffpa_api.cc:180-205)
    C.ffpa_attn_forward(q, k, v, q.new_empty(0), o, lse, 0, 0, 0, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  with pytest.raises(RuntimeError):  # bias + causal
    C.ffpa_attn_forward(q, k, v, torch.zeros(1, 1, 64, 64, device=DEV), o, lse, 0, 1, 1, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  # zero-sized problems return without a launch (B == 0, Nq == 0) or give the empty-row convention (Nkv == 0)
  n0 = C.launch_count()
  C.ffpa_attn_forward(q[:0], k[:0], v[:0], q.new_empty(0), o[:0], q.new_empty(0), 0, 1, 0, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  C.ffpa_attn_forward(q, k[:, :, :0], v[:, :, :0], q.new_empty(0), o, lse, 0, 1, 0, 0.125, 0.0, 0, 0, *_FP8_DEFAULTS)
  assert C.launch_count() == n0
  assert torch.all(o == 0) and torch.all(lse == float("-inf"))


def test_torch_compile_sees_the_op():  # tests/test_ffpa_compile.py:53-71
  import ffpa_attn

  q, k, v = _mk(1, 2, 2, 256, 256, 128, torch.bfloat16)
  f = torch.compile(lambda a, b, c: ffpa_attn.ffpa_attn_func(a, b, c, is_causal=True) * 2, fullgraph=False)
  eager = ffpa_attn.ffpa_attn_func(q, k, v, is_causal=True) * 2
  assert torch.equal(f(q, k, v), eager)


def test_monkey_patched_sdpa_reaches_the_kernel_without_recursion():
  """Documented drop-in: F.scaled_dot_product_attention = ffpa_attn_func
  (/root/reference/README.md:53-59, tests/test_monkey_patch.py:104-136). The reference recurses into
  aten for the shapes it does not serve; here every shape is served by the sm_100a kernel."""
  import torch.nn.functional as F

  import ffpa_attn

  q, k, v = _mk(1, 4, 4, 300, 300, 512, torch.bfloat16)
  orig = F.scaled_dot_product_attention
  ref = orig(q.float(), k.float(), v.float(), is_causal=True)
  F.scaled_dot_product_attention = ffpa_attn.ffpa_attn_func
  try:
    n0 = ffpa_attn._C.launch_count()
    out = F.scaled_dot_product_attention(q, k, v, is_causal=True)
    small = F.scaled_dot_product_attention(q[..., :64].contiguous(), k[..., :64].contiguous(), v[..., :64].contiguous())
    torch.cuda.synchronize()
    assert ffpa_attn._C.launch_count() - n0 == 2
  finally:
    F.scaled_dot_product_attention = orig
  assert (out.float() - ref).abs().max().item() < 2e-2
  assert small.shape == (1, 4, 300, 64)


@pytest.mark.parametrize("D,causal", [(768, False), (1024, True), (640, True), (896, False)])
def test_large_headdims(D, causal):  # tests/test_ffpa_fwd.py:1226-1249 includes D=1024 causal
  q, k, v = _mk(1, 2, 2, 384, 520, D, torch.bfloat16)
  out = _run(q, k, v, is_causal=causal)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=causal)
  _check(out, ref, 2e-2, f"D={D}")


@pytest.mark.parametrize("D", [776, 832, 896, 1024])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(1, 2, 2, 384, 520), (2, 4, 2, 300, 300), (1, 2, 1, 1000, 1300), (1, 2, 2, 5, 257)])
def test_replay_path_matches_two_pass_path_and_oracle(D, causal, shape, monkeypatch):
  """Head dims > 768 (two O slabs): by default pass 0 stores its P tiles / rescale factors / 1/rowsum and
  the second slab is a GEMM that replays them (csrc/ffpa_fwd_replay_sm100.cuh, 2 launches); with
  FFPA_FWD_REPLAY=0 the second slab is a full second softmax pass (1 launch). Same P bits, same rescale
  schedule -> the two paths must agree to rounding, and both with the oracle (GQA, tails, odd query-tile
  counts, bottom-right causal)."""
  import ffpa_attn

  B, Hq, Hkv, Nq, Nkv = shape
  q, k, v = _mk(B, Hq, Hkv, Nq, Nkv, D, torch.bfloat16, seed=31)
  kw = dict(is_causal=causal, enable_gqa=Hq != Hkv)
  n0 = ffpa_attn._C.launch_count()
  o_replay, lse_replay = _lse(q, k, v, causal)
  n_replay = ffpa_attn._C.launch_count() - n0
  monkeypatch.setenv("FFPA_FWD_REPLAY", "0")
  ffpa_attn._C.refresh_env()   # the library caches its tuning variables per process
  n0 = ffpa_attn._C.launch_count()
  o_two, lse_two = _lse(q, k, v, causal)
  n_two = ffpa_attn._C.launch_count() - n0
  assert n_replay == 2 and n_two == 1
  assert torch.equal(lse_replay, lse_two)
  assert torch.equal(o_replay[..., :512], o_two[..., :512])   # first slab: the same kernel pass
  assert (o_replay.float() - o_two.float()).abs().max().item() < 4e-3
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=causal)
  _check(o_replay, ref, 2e-2, f"replay D={D}")
  del kw


def test_replay_path_chunked_by_scratch_bound(monkeypatch):
  """FFPA_FWD_REPLAY_MAX_GB bounds the O(Nq*Nkv) replay stash: a larger problem runs as (batch element, KV-head
  range) chunks through one scratch and must equal the unchunked run bit for bit; below a machine-filling chunk
  the two-pass kernel runs (same LSE and first slab bits, second slab to rounding)."""
  import ffpa_attn

  q, k, v = _mk(2, 8, 4, 2048, 2048, 1024, torch.bfloat16, seed=33)
  n0 = ffpa_attn._C.launch_count()
  o_full, lse_full = _lse(q, k, v, True)
  assert ffpa_attn._C.launch_count() - n0 == 2
  monkeypatch.setenv("FFPA_FWD_REPLAY_MAX_GB", "0.06")   # whole stash: 134 MB; 3 KV heads (6 query heads): 50 MB
  n0 = ffpa_attn._C.launch_count()
  o_c, lse_c = _lse(q, k, v, True)
  assert ffpa_attn._C.launch_count() - n0 == 2 * 2 * 2   # 2 batch elements x KV-head chunks (3, 1) x 2 launches
  assert torch.equal(o_c, o_full) and torch.equal(lse_c, lse_full)
  monkeypatch.setenv("FFPA_FWD_REPLAY_MAX_GB", "0.001")
  n0 = ffpa_attn._C.launch_count()
  o_t, lse_t = _lse(q, k, v, True)
  assert ffpa_attn._C.launch_count() - n0 == 1
  assert torch.equal(lse_t, lse_full) and torch.equal(o_t[..., :512], o_full[..., :512])
  assert (o_t.float() - o_full.float()).abs().max().item() < 4e-3


def test_replay_path_with_bias_dropout_fp16_and_large_amplitude():
  """The replayed P is whatever pass 0 fed its MMA: bias and dropout included; large score amplitudes make
  the row max move often, i.e. many rescale events to replay."""
  import ffpa_attn

  q, k, v = _mk(1, 2, 2, 300, 700, 1024, torch.float16, seed=32, amp=2.0)
  bias = (torch.randn(1, 2, 300, 700, generator=torch.Generator().manual_seed(6)) * 3).to(DEV)
  out = _run(q, k, v, attn_mask=bias)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=bias.cpu().double().numpy())
  _check(out, ref, 1e-2, "bias, amp 2")
  torch.cuda.manual_seed(77)
  seed, offset = int(torch.cuda.initial_seed()), int(torch.cuda._get_rng_state_offset())
  out = _run(q, k, v, dropout_p=0.2, is_causal=True)
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), causal=True, dropout_p=0.2, philox_seed=seed, philox_offset=offset)
  _check(out, ref, 4e-2, "dropout causal")


@pytest.mark.parametrize("H", [8, 16, 48])
@pytest.mark.parametrize("D", [64, 192, 320, 576, 640])
def test_dispatch_smoke_heads_by_headdim(H, D):  # tests/test_ffpa_fwd.py:44-45 (H x D dispatch grid)
  q, k, v = _mk(1, H, H, 160, 160, D, torch.float16, seed=H + D)
  out = _run(q, k, v)
  hs = [0, H // 2, H - 1]
  ref, _ = orc.attention_fwd(q[:, hs].cpu(), k[:, hs].cpu(), v[:, hs].cpu())
  _check(out[:, hs], ref, 1e-2, f"H={H} D={D}")


@pytest.mark.parametrize("Nq,Nkv", [(8191, 8192), (1, 4096), (4095, 5000)])
def test_cross_attention_long_sampled(Nq, Nkv):  # tests/test_ffpa_fwd.py:1110-1120
  q, k, v = _mk(1, 2, 2, Nq, Nkv, 512, torch.bfloat16, seed=7)
  out = _run(q, k, v)
  rows = sorted({0, Nq // 3, Nq - 1})
  _sampled_rows_check(q, k, v, out, rows, [(0, 0), (0, 1)], False, 1e-2)
  outc = _run(q, k, v, is_causal=True)
  _sampled_rows_check(q, k, v, outc, rows, [(0, 1)], True, 1e-2)


@pytest.mark.parametrize("Nq,Nkv,Hq,Hkv,causal", [(1, 8192, 32, 32, False), (1, 4096, 16, 4, False), (4, 5000, 8, 8, True),
                                                   (7, 3000, 4, 2, True), (130, 4096, 2, 2, True)])
def test_decode_like_shapes_use_kv_splits(Nq, Nkv, Hq, Hkv, causal):
  """Few query rows, long KV (tests/test_ffpa_fwd.py:1110-1120 has (1, 4096)): the launcher splits the KV
  range over clusters and merges fp32 partials (reference: split-KV decode, native/sm_80/split_kv.cuh)."""
  import ffpa_attn

  q, k, v = _mk(1, Hq, Hkv, Nq, Nkv, 512, torch.bfloat16, seed=3)
  n0 = ffpa_attn._C.launch_count()
  o, lse = _lse(q, k, v, causal=causal)
  assert ffpa_attn._C.launch_count() - n0 == 2, "expected split kernel + merge kernel"
  hs = [0, Hq - 1]
  g = Hq // Hkv
  for h in hs:
    ref, lref = orc.attention_fwd(q[:, h:h + 1].cpu(), k[:, h // g:h // g + 1].cpu(), v[:, h // g:h // g + 1].cpu(), causal=causal)
    assert np.abs(o[:, h:h + 1].float().cpu().numpy() - ref).max() < 1e-2
    assert np.abs(lse[:, h:h + 1].cpu().numpy() - lref).max() < 2e-4


def test_kv_split_with_bias_and_dropout_matches_oracle():
  q, k, v = _mk(1, 2, 2, 3, 2048, 256, torch.bfloat16, seed=4)
  bias = torch.randn(1, 2, 3, 2048)
  out = _run(q, k, v, attn_mask=bias.to(DEV))
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), bias=bias.double().numpy())
  _check(out, ref, 2e-2, "split+bias")
  seed, offset = 99, 12
  o, _ = torch.ops.ffpa_attn._fwd_cuda(q, k, v, q.new_empty(0), 0, 1, 0, 256 ** -0.5, 0.25, seed, offset, True, False,
                                       0, 0, 0, 1, 0, False, 256, False, 256)
  torch.cuda.synchronize()
  ref, _ = orc.attention_fwd(q.cpu(), k.cpu(), v.cpu(), dropout_p=0.25, philox_seed=seed, philox_offset=offset)
  _check(o, ref, 4e-2, "split+dropout")


def test_cuda_graph_capture_and_side_stream():
  """The launcher does no host synchronisation and launches on the caller's current stream
  (reference contract: native/launch.cuh:303-304), so it can be captured into a CUDA graph and run on a side
  stream. (Causal shapes upload their schedule table on first use: warm up once before capturing.)"""
  import ffpa_attn

  q, k, v = _mk(2, 4, 2, 384, 640, 512, torch.bfloat16, seed=12)
  kw = dict(is_causal=True, enable_gqa=True)
  eager = ffpa_attn.ffpa_attn_func(q, k, v, **kw)  # warm-up: smem attribute, schedule table
  torch.cuda.synchronize()
  s = torch.cuda.Stream()
  with torch.cuda.stream(s):
    side = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
  s.synchronize()
  assert torch.equal(side, eager)
  g = torch.cuda.CUDAGraph()
  with torch.cuda.graph(g):
    captured = ffpa_attn.ffpa_attn_func(q, k, v, **kw)
  captured.zero_()
  g.replay()
  torch.cuda.synchronize()
  assert torch.equal(captured, eager)


def test_large_batch_many_items():
  q, k, v = _mk(48, 4, 4, 256, 256, 128, torch.float16, seed=13)
  out = _run(q, k, v, is_causal=True)
  ref, _ = orc.attention_fwd(q[[0, 47]].cpu(), k[[0, 47]].cpu(), v[[0, 47]].cpu(), causal=True)
  _check(out[[0, 47]], ref, 1e-2)
