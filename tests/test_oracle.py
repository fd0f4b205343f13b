"""CPU tests: the oracle (oracle/attention_oracle.py) against the golden fixtures produced by the
reference package itself (oracle/make_golden.py), plus known-answer tests."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc


def _load(path):
  z = np.load(path)
  B, Hq, Hkv, Nq, Nkv, D, is_bf16, causal = [int(x) for x in z["meta"]]
  dt = torch.bfloat16 if is_bf16 else torch.float16

  def t(name, shape):
    return torch.from_numpy(z[name].view(np.int16).copy()).view(dt).reshape(shape)

  q = t("q", (B, Hq, Nq, D))
  k = t("k", (B, Hkv, Nkv, D))
  v = t("v", (B, Hkv, Nkv, D))
  d_o = t("d_o", (B, Hq, Nq, D))
  o = t("o", (B, Hq, Nq, D))
  mask = torch.from_numpy(z["mask"]) if "mask" in z.files else None
  return dict(z=z, q=q, k=k, v=v, d_o=d_o, o=o, mask=mask, causal=bool(causal), dtype=dt)


def _bias_of(mask):
  if mask is None:
    return None
  if mask.dtype == torch.bool:
    return torch.where(mask, 0.0, float("-inf")).double().numpy()
  return mask.double().numpy()


GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_golden_present():
  assert len(GOLDEN) >= 6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_forward_matches_reference(path):
  g = _load(path)
  o, lse = orc.attention_fwd(g["q"], g["k"], g["v"], bias=_bias_of(g["mask"]), causal=g["causal"])
  ref = g["o"].double().numpy()
  # the reference output is rounded to bf16/fp16 and computed by aten with 16-bit probabilities:
  # tolerance = the reference tests' own (tests/test_ffpa_fwd.py:106-113)
  tol = 2e-2 if g["dtype"] == torch.bfloat16 else 1e-2
  assert np.abs(o - ref).max() < tol
  if "o_f32" in g["z"].files:  # fp32 reference run: tight
    assert np.abs(o - g["z"]["o_f32"]).max() < 2e-5
  assert np.isfinite(lse).all()


@pytest.mark.parametrize("path", [p for p in GOLDEN if p.endswith("_bwd.npz")],
                         ids=[os.path.basename(p)[:-4] for p in GOLDEN if p.endswith("_bwd.npz")])
def test_oracle_backward_matches_reference(path):
  g = _load(path)
  dq, dk, dv, _ = orc.attention_bwd(g["q"], g["k"], g["v"], g["d_o"], bias=_bias_of(g["mask"]),
                                    causal=g["causal"])
  z = g["z"]
  for got, name in ((dq, "dq"), (dk, "dk"), (dv, "dv")):
    ref = z[name].astype(np.float64)
    assert np.abs(got - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), name


def test_c1_config_matches_torch_sdpa():
  """BASELINE config 1: B=1 H=2 N=512 D=320 bf16 through SDPA on the host CPU."""
  torch.manual_seed(0)
  q, k, v = (torch.randn(1, 2, 512, 320, dtype=torch.bfloat16) for _ in range(3))
  ref = orc.sdpa_cpu(q, k, v).double().numpy()
  o, _ = orc.attention_fwd(q, k, v)
  assert np.abs(o - ref).max() < 2e-2


def test_philox_known_answers():
  # Random123 kat_vectors, philox4x32 10 rounds
  kat = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
  ]
  for ctr, key, want in kat:
    got = orc.philox4x32_10(np.array([ctr], dtype=np.uint64), key)[0]
    assert tuple(int(x) for x in got) == want


def test_dropout_mask_statistics_and_fwd_consistency():
  keep = orc.dropout_keep_mask(1, 2, 64, 96, 0.25, seed=1234, offset=8)
  assert abs(keep.mean() - 0.75) < 0.02
  # offset shifts the stream by elements
  keep2 = orc.dropout_keep_mask(1, 2, 64, 96, 0.25, seed=1234, offset=8 + 96)
  assert (keep.reshape(-1)[96:] == keep2.reshape(-1)[:-96]).all()
  torch.manual_seed(1)
  q, k, v = (torch.randn(1, 2, 64, 32) for _ in range(3))
  o0, lse0 = orc.attention_fwd(q, k[:, :, :64], v[:, :, :64])
  o1, lse1 = orc.attention_fwd(q, k[:, :, :64], v[:, :, :64], dropout_p=0.25, philox_seed=7)
  assert np.allclose(lse0, lse1)  # LSE ignores dropout
  assert not np.allclose(o0, o1)


def test_empty_rows_and_causal_kats():
  torch.manual_seed(2)
  q = torch.randn(1, 1, 4, 8)
  k = torch.randn(1, 1, 6, 8)
  v = torch.randn(1, 1, 6, 8)
  bias = np.zeros((1, 1, 4, 6))
  bias[0, 0, 1, :] = -np.inf  # row 1 sees nothing
  o, lse = orc.attention_fwd(q, k, v, bias=bias)
  assert (o[0, 0, 1] == 0).all() and lse[0, 0, 1] == -np.inf
  # bottom-right causal: row r sees keys <= r + 2
  o, _ = orc.attention_fwd(q, k, v, causal=True)
  o_row0, _ = orc.attention_fwd(q[:, :, :1], k[:, :, :3], v[:, :, :3])
  assert np.allclose(o[0, 0, 0], o_row0[0, 0, 0])
  # s_k == 1: dK = dQ = 0, dV = sum dO  (tests/test_ffpa_cute_sm100.py:1026-1050)
  d_o = torch.randn(1, 1, 4, 8)
  dq, dk, dv, _ = orc.attention_bwd(q, k[:, :, :1], v[:, :, :1], d_o)
  assert np.abs(dq).max() < 1e-12 and np.abs(dk).max() < 1e-12
  assert np.allclose(dv[0, 0, 0], d_o.double().numpy().sum(axis=2)[0, 0])


def test_flops_formula():
  # /root/reference/tests/test_perf_tflops.py:17-57 style KATs
  assert orc.attn_flops(1, 32, 8192, 8192, 512) == 4.0 * 32 * 512 * 8192 * 8192
  assert orc.attn_flops(1, 1, 4, 4, 8, causal=True) == 4.0 * 8 * 10
  assert orc.attn_flops(1, 1, 2, 4, 8, causal=True) == 4.0 * 8 * (3 + 4)
  assert orc.attn_flops(2, 4, 16, 16, 64, mode="bwd") == 2.5 * orc.attn_flops(2, 4, 16, 16, 64)


# ---- FP8 oracle (oracle/fp8_oracle.py): pinned on e4m3 known answers, the reference's own acceptance bounds
# ---- versus exact attention (/root/reference/tests/test_ffpa_fp8.py:56-86, 238-251) and algebraic identities
def test_e4m3_rounding_known_answers():
  from oracle import fp8_oracle as f8

  x = np.array([0.0, 1.0, 1.0625, 1.1875, 447.0, 448.0, 449.0, 1e6, -1e6, 2.0 ** -9, 2.0 ** -10, 0.3, -0.3, 17.0, 25.0],
               dtype=np.float32)
  #            ties-to-even: 1.0625 -> 1.0 (mantissa step 0.125), 1.1875 -> 1.25; saturation at 448; subnormal step 2^-9
  want = np.array([0.0, 1.0, 1.0, 1.25, 448.0, 448.0, 448.0, 448.0, -448.0, 2.0 ** -9, 0.0, 0.3125, -0.3125, 16.0, 24.0],
                  dtype=np.float32)
  assert np.array_equal(f8.e4m3_round(x), want)
  assert f8.e4m3_bits(np.array([448.0, -448.0, 1.0, 0.0], dtype=np.float32)).tolist() == [0x7E, 0xFE, 0x38, 0x00]


def test_fp8_quantiser_contract():
  """quantize_fp8.cuh:67-168: one scale per (b, h, 128-row block) = amax / 448, the block's largest magnitude maps
  to +-448 exactly, an all-zero block gets scale 0 and zeros (inv_s = 0 branch), tails use the rows that exist."""
  from oracle import fp8_oracle as f8

  rng = np.random.default_rng(0)
  x = rng.standard_normal((2, 3, 300, 64)).astype(np.float32)
  x[0, 0, 128:256] = 0.0
  x8, s = f8.quantize_per_block(x)
  assert s.shape == (2, 3, 3)
  assert s[0, 0, 1] == 0.0 and np.all(x8[0, 0, 128:256] == 0)
  for t in range(3):
    blk = x[1, 2, t * 128:(t + 1) * 128]
    assert np.isclose(s[1, 2, t], np.abs(blk).max() / 448.0, rtol=1e-6)
    assert np.abs(x8[1, 2, t * 128:(t + 1) * 128]).max() == 448.0
  deq = x8 * np.repeat(s, 128, axis=2)[:, :, :300, None]
  assert np.abs(deq - x).max() <= np.abs(x).max() / 16 + 1e-6   # e4m3: 3 mantissa bits -> relative step 2^-4 at most


@pytest.mark.parametrize("causal,smooth_k", [(False, True), (False, False), (True, True)])
def test_fp8_oracle_meets_reference_bounds(causal, smooth_k):
  """The quantised restatement stays inside the bounds the reference holds its FP8 kernels to versus exact fp32
  attention on randn * 0.5 inputs: O 4e-2 dense / 1e-1 causal, LSE 5e-2 (tests/test_ffpa_fp8.py:63-65, 71, 86, 251)."""
  from oracle import fp8_oracle as f8

  torch.manual_seed(0)
  B, Hq, Hkv, Nq, Nkv, D = 1, 4, 2, 256, 384, 128
  q = (torch.randn(B, Hq, Nq, D) * 0.5).to(torch.bfloat16)
  k = (torch.randn(B, Hkv, Nkv, D) * 0.5 + 0.75).to(torch.bfloat16)   # a channel mean smooth-K can remove
  v = (torch.randn(B, Hkv, Nkv, D) * 0.5).to(torch.bfloat16)
  o8, lse8, aux = f8.fp8_attention_fwd(q, k, v, causal=causal, smooth_k=smooth_k)
  ref, lse = orc.attention_fwd(q, k, v, causal=causal)
  assert np.abs(o8 - ref).max() < (1e-1 if causal else 4e-2)
  assert np.abs(lse8 - lse).max() < 5e-2
  assert aux["q8"].shape == (B, Hq, Nq, D) and aux["ks"].shape == (B, Hkv, 3)


def test_fp8_smooth_k_identities():
  """Softmax is shift invariant: subtracting mean_seq(K) changes neither O nor (after the q.km correction) the LSE
  beyond quantisation noise, and it shrinks the K scales when K has a large common offset (smooth_k.cuh:8-16)."""
  from oracle import fp8_oracle as f8

  torch.manual_seed(1)
  q = (torch.randn(1, 2, 128, 64) * 0.5).to(torch.float16)
  k = (torch.randn(1, 2, 256, 64) * 0.25 + 1.5).to(torch.float16)
  v = (torch.randn(1, 2, 256, 64) * 0.5).to(torch.float16)
  o_s, lse_s, a_s = f8.fp8_attention_fwd(q, k, v, smooth_k=True)
  o_n, lse_n, a_n = f8.fp8_attention_fwd(q, k, v, smooth_k=False)
  ref, lse = orc.attention_fwd(q, k, v)
  assert a_s["ks"].max() < 0.5 * a_n["ks"].max()
  assert np.abs(o_s - ref).max() < np.abs(o_n - ref).max() + 1e-3
  # the correction term is scale * qs * dot(Q8_row, km): the QUANTISED query against a mean of magnitude 1.5 over 64
  # channels, so its own rounding error is what bounds the LSE here (the sm_100a kernel uses the unquantised row)
  assert np.abs(o_s - ref).max() < 4e-2 and np.abs(lse_s - lse).max() < 1e-1
