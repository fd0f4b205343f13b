"""CPU tests for the boundary: the C-ABI library loads and exports every symbol declared in
include/ffpa_b200.h, the ctypes structs match the C layout, and the host-side mirror of the
reference API validates inputs with the reference's error classes. No kernel is launched."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffpa_b200.h")
LIB = os.path.join(ROOT, "ffpa-attn_b200", "ffpa_attn", "libffpa_b200.so")


@pytest.fixture(scope="module")
def lib():
  if not os.path.exists(LIB):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.build()
  return ctypes.CDLL(LIB)


def _declared_symbols():
  src = open(HEADER).read()
  return sorted(set(re.findall(r"\b(ffpa_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
  syms = _declared_symbols()
  for s in ("ffpa_b200_fwd", "ffpa_b200_bwd", "ffpa_b200_set_backend_impl", "ffpa_b200_get_backend_impl",
            "ffpa_b200_last_error", "ffpa_b200_launch_count", "ffpa_b200_bwd_workspace_bytes",
            "ffpa_b200_fwd_workspace_bytes"):
    assert s in syms


def test_library_exports_every_declared_symbol(lib):
  for s in _declared_symbols():
    assert hasattr(lib, s), f"libffpa_b200.so does not export {s}"


def test_abi_version_and_flags(lib):
  lib.ffpa_b200_abi_version.restype = ctypes.c_int32
  assert lib.ffpa_b200_abi_version() == 2
  assert lib.ffpa_b200_fwd_available() == 1


def test_ctypes_struct_layout_matches_c():
  import ffpa_attn._C as C

  prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "ffpa_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(ffpa_fwd_params), offsetof(ffpa_fwd_params, bias_stride),
         offsetof(ffpa_fwd_params, batch), offsetof(ffpa_fwd_params, softmax_scale),
         offsetof(ffpa_fwd_params, philox_seed), offsetof(ffpa_fwd_params, philox_offset),
         offsetof(ffpa_fwd_params, workspace_bytes), offsetof(ffpa_fwd_params, cu_seqlens_q),
         offsetof(ffpa_fwd_params, total_k));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(ffpa_bwd_params), offsetof(ffpa_bwd_params, batch),
         offsetof(ffpa_bwd_params, softmax_scale), offsetof(ffpa_bwd_params, workspace),
         offsetof(ffpa_bwd_params, bias_kind), offsetof(ffpa_bwd_params, d_bias),
         offsetof(ffpa_bwd_params, cu_seqlens_k), offsetof(ffpa_bwd_params, total_q));
  printf("%zu\n", offsetof(ffpa_bwd_params, d_lse));
  return 0;
}
"""
  with tempfile.TemporaryDirectory() as d:
    src = os.path.join(d, "t.c")
    open(src, "w").write(prog)
    exe = os.path.join(d, "t")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
    out = subprocess.check_output([exe]).decode().split()
  F, B = C._FwdParams, C._BwdParams
  want = [ctypes.sizeof(F), F.bias_stride.offset, F.batch.offset, F.softmax_scale.offset,
          F.philox_seed.offset, F.philox_offset.offset, F.workspace_bytes.offset,
          F.cu_seqlens_q.offset, F.total_k.offset,
          ctypes.sizeof(B), B.batch.offset, B.softmax_scale.offset, B.workspace.offset,
          B.bias_kind.offset, B.d_bias.offset, B.cu_seqlens_k.offset, B.total_q.offset, B.d_lse.offset]
  assert [int(x) for x in out] == want


def test_backend_hint_roundtrip(lib):
  import ffpa_attn

  ffpa_attn.set_cuda_backend_impl(ffpa_attn.CudaBackendImpl.TMA)
  assert ffpa_attn.get_cuda_backend_impl() == ffpa_attn.CudaBackendImpl.TMA
  ffpa_attn.set_cuda_backend_impl(ffpa_attn.CudaBackendImpl.AUTO)
  with pytest.raises(RuntimeError):
    ffpa_attn._C.set_cuda_backend_impl(99)


def test_c_abi_rejects_bad_arguments_without_a_gpu(lib):
  """Argument validation happens before any CUDA work except the device probe; on a box with no
  GPU every call must fail loudly (no silent CPU path)."""
  import ffpa_attn._C as C

  p = C._FwdParams()
  rc = C._lib.ffpa_b200_fwd(ctypes.byref(p), None)
  assert rc < 0
  assert len(C._lib.ffpa_b200_last_error()) > 0
  assert C._lib.ffpa_b200_fwd(None, None) == -1


# ---- host-side API semantics (reference: tests/test_ffpa_fwd.py:162-177, 1146-1152, 1199-1215) ----
def _qkv(B=1, Hq=2, Hkv=2, Nq=16, Nkv=16, D=64, dtype=torch.bfloat16):
  q = torch.randn(B, Hq, Nq, D, dtype=dtype)
  k = torch.randn(B, Hkv, Nkv, D, dtype=dtype)
  v = torch.randn(B, Hkv, Nkv, D, dtype=dtype)
  return q, k, v


def test_unknown_kwarg_is_type_error():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv()
  with pytest.raises(TypeError, match="unexpected keyword"):
    ffpa_attn_func(q, k, v, not_a_kwarg=1)


def test_fp32_inputs_are_type_error():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv(dtype=torch.float32)
  with pytest.raises(TypeError, match="fp16/bf16"):
    ffpa_attn_func(q, k, v)


def test_gqa_needs_opt_in_and_divisibility():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv(Hq=4, Hkv=2)
  with pytest.raises(ValueError, match="enable_gqa"):
    ffpa_attn_func(q, k, v)
  q, k, v = _qkv(Hq=3, Hkv=2)
  with pytest.raises(ValueError, match="integer multiple"):
    ffpa_attn_func(q, k, v, enable_gqa=True)


def test_causal_requires_nkv_ge_nq_and_excludes_mask():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv(Nq=32, Nkv=16)
  with pytest.raises(ValueError, match="Nkv >= Nq"):
    ffpa_attn_func(q, k, v, is_causal=True)
  q, k, v = _qkv()
  with pytest.raises(RuntimeError, match="attn_mask"):
    ffpa_attn_func(q, k, v, attn_mask=torch.ones(16, 16, dtype=torch.bool), is_causal=True)


def test_dropout_range_and_shapes():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv()
  with pytest.raises(ValueError):
    ffpa_attn_func(q, k, v, dropout_p=1.0)
  with pytest.raises(ValueError):
    ffpa_attn_func(q, k, v, dropout_p=-0.1)
  with pytest.raises(ValueError, match="4-D"):
    ffpa_attn_func(q[0], k[0], v[0])


def test_cpu_tensors_fail_loudly_no_fallback():
  from ffpa_attn import ffpa_attn_func

  q, k, v = _qkv()
  with pytest.raises(RuntimeError, match="no CPU"):
    ffpa_attn_func(q, k, v)


def test_host_func_units_partition_kv_heads_and_no_cpu_fallback():
  from ffpa_attn import ffpa_attn_host_func
  from ffpa_attn.host import _units

  for B, Hkv, chunks in [(1, 32, 8), (2, 8, 3), (3, 1, 8), (1, 5, 16)]:
    units = _units(B, Hkv, chunks)
    for b in range(B):
      mine = [(lo, hi) for (bb, lo, hi) in units if bb == b]
      assert mine[0][0] == 0 and mine[-1][1] == Hkv
      assert all(a[1] == c[0] and a[0] < a[1] for a, c in zip(mine, mine[1:]))
      assert len(mine) == min(Hkv, chunks)
  q, k, v = _qkv()
  if not torch.cuda.is_available():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
      ffpa_attn_host_func(q, k, v)
  with pytest.raises(NotImplementedError):
    ffpa_attn_host_func(q, k, v, dropout_p=0.1)
  with pytest.raises(ValueError, match="enable_gqa"):
    ffpa_attn_host_func(q, k[:, :1], v[:, :1])


def test_backend_kwarg_accepts_only_cuda():
  from ffpa_attn import CUDABackend, FFPAAttnMeta

  m = FFPAAttnMeta.from_kwargs(backend="cuda")
  assert isinstance(m.forward_meta, CUDABackend)
  m = FFPAAttnMeta.from_kwargs(forward_backend=CUDABackend(enable_fp8=True))
  assert m.forward_meta.impl_hint.name == "CUTE_TMA_FP8"
  for name in ("triton", "cutedsl", "sdpa"):
    with pytest.raises(NotImplementedError):
      FFPAAttnMeta.from_kwargs(backend=name)
  with pytest.raises(ValueError):
    CUDABackend(acc="f16")
  with pytest.raises(TypeError):
    FFPAAttnMeta.from_kwargs(backend=3)


def test_mask_normalisation_matches_reference_rules():
  from ffpa_attn import FFPAAttnMeta

  meta = FFPAAttnMeta()
  q, k, _ = _qkv(B=2, Nq=8, Nkv=12)
  m2 = torch.rand(8, 12) > 0.5
  b = meta.normalize_attn_mask(q, k, m2)
  assert b.shape == (1, 1, 8, 12) and b.dtype == q.dtype
  assert set(b.unique().tolist()) <= {0.0, float("-inf")}
  m3 = torch.randn(2, 8, 12)
  assert meta.normalize_attn_mask(q, k, m3).shape == (2, 1, 8, 12)
  with pytest.raises(ValueError, match="broadcastable"):
    meta.normalize_attn_mask(q, k, torch.zeros(3, 1, 8, 12))
  with pytest.raises(TypeError, match="dtype"):
    meta.normalize_attn_mask(q, k, torch.zeros(8, 12, dtype=torch.float64))


def test_fake_op_registered_for_compile():
  import ffpa_attn  # noqa: F401

  q = torch.empty(1, 2, 16, 64, dtype=torch.bfloat16, device="meta")
  o, lse = torch.ops.ffpa_attn._fwd_cuda(q, q, q, q.new_empty(0), 0, 1, 0, 0.125, 0.0, 0, 0, True, False,
                                         0, 0, 0, 0, 0, False, 256, False, 256)
  assert o.shape == q.shape and lse.shape == (1, 2, 16) and lse.dtype == torch.float32


# ---- packed variable-length entry, workspace plans, hybrid resolution (host logic only) ----
def test_varlen_host_validation_and_no_cpu_fallback():
  """Reference checks (cute/__init__.py:466-571): dtypes and shapes only, cu_seqlens values are never read on
  the host; CPU tensors must fail loudly."""
  from ffpa_attn import ffpa_attn_varlen_func

  q = torch.randn(32, 2, 64, dtype=torch.bfloat16)
  k = torch.randn(32, 2, 64, dtype=torch.bfloat16)
  cu = torch.tensor([0, 16, 32], dtype=torch.int32)
  with pytest.raises(TypeError, match="int32"):
    ffpa_attn_varlen_func(q, k, k, cu.long(), cu, 16, 16)
  with pytest.raises(NotImplementedError):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, dropout_p=0.1)
  with pytest.raises(NotImplementedError, match="softcap"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, softcap=1.0)
  with pytest.raises(TypeError, match="unexpected"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, bogus=1)
  with pytest.raises(ValueError, match="enable_gqa"):
    ffpa_attn_varlen_func(q, k[:, :1], k[:, :1], cu, cu, 16, 16)
  with pytest.raises(ValueError, match="same batch"):
    ffpa_attn_varlen_func(q, k, k, cu, cu[:-1], 16, 16)
  with pytest.raises(ValueError, match="THD"):
    ffpa_attn_varlen_func(q[None], k[None], k[None], cu, cu, 16, 16)
  with pytest.raises(TypeError, match="fp16/bf16"):
    ffpa_attn_varlen_func(q.float(), k.float(), k.float(), cu, cu, 16, 16)
  with pytest.raises(ValueError, match="only backend"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, backend="cutedsl")
  with pytest.raises(RuntimeError, match="no CPU"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16)


def test_backward_workspace_plan(lib, monkeypatch):
  """ffpa_b200_bwd_workspace_bytes = required scratch + the two 16-bit score buffers of the stash path for head
  dims 384..1024, bounded by FFPA_BWD_STASH_MAX_GB (KV-head chunks beyond it); _min = required scratch only."""
  f, m = lib.ffpa_b200_bwd_workspace_bytes, lib.ffpa_b200_bwd_workspace_bytes_min
  for fn in (f, m):
    fn.argtypes = [ctypes.c_int32] * 6
    fn.restype = ctypes.c_uint64
  monkeypatch.delenv("FFPA_BWD_STASH_MAX_GB", raising=False)
  monkeypatch.delenv("FFPA_BWD_STASH", raising=False)
  B, H, N, D = 1, 32, 8192, 512
  stash = 2 * B * H * N * N * 2
  assert f(B, H, H, N, N, D) == m(B, H, H, N, N, D) + stash
  assert f(B, H, H, N, N, 256) == m(B, H, H, N, N, 256)            # small heads: recompute kernels
  assert f(B, H, H, N, N, 1024) == m(B, H, H, N, N, 1024) + stash   # large heads: stash on the first slab pass
  assert f(1, 4, 2, 130, 257, 512) == m(1, 4, 2, 130, 257, 512) + 2 * 4 * 256 * 512 * 2   # padded to 128 x 256
  # 4 batch elements would need 34 GB: chunked per batch element under the default 20 GB cap
  assert f(4, H, H, N, N, D) <= 20 * 2 ** 30 + m(4, H, H, N, N, D)
  assert f(4, H, H, N, N, D) >= stash
  monkeypatch.setenv("FFPA_BWD_STASH", "0")
  assert f(B, H, H, N, N, D) == m(B, H, H, N, N, D)


def test_forward_workspace_plan(lib, monkeypatch):
  """Forward scratch: KV-split partials for decode-like shapes, the replay stash for head dims > 768, FP8 copies."""
  f = lib.ffpa_b200_fwd_workspace_bytes
  f.argtypes = [ctypes.c_int32] * 7
  f.restype = ctypes.c_uint64
  monkeypatch.delenv("FFPA_FWD_REPLAY", raising=False)
  monkeypatch.delenv("FFPA_FWD_REPLAY_MAX_GB", raising=False)
  assert f(1, 32, 32, 8192, 8192, 512, 0) == 0
  assert f(1, 32, 32, 8192, 8192, 768, 0) == 0
  need = f(1, 32, 32, 8192, 8192, 1024, 0)
  assert 32 * 8192 * 8192 * 2 <= need <= 32 * 8192 * 8192 * 2 * 1.02   # P tiles + factors + 1/rowsum
  assert f(1, 32, 32, 1, 8192, 512, 0) > 0     # decode: KV-split partials
  assert f(1, 32, 32, 8192, 8192, 256, 1) > 3 * 32 * 8192 * 256   # FP8: e4m3 copies of Q, K, V + scales
  monkeypatch.setenv("FFPA_FWD_REPLAY", "0")
  assert f(1, 32, 32, 8192, 8192, 1024, 0) == 0


def test_fp8_hybrid_auto_resolution():
  """*_hybrid=None resolves to (enable_fp8 and is_causal) in normalize_inputs (reference functional.py:781-794)."""
  from ffpa_attn import CUDABackend, FFPAAttnMeta

  q, k, v = _qkv()
  for causal, fp8, want in ((True, True, True), (False, True, False), (True, False, False)):
    meta = FFPAAttnMeta.from_kwargs(forward_backend=CUDABackend(enable_fp8=fp8))
    with pytest.raises(RuntimeError, match="no CPU"):   # CPU tensors are rejected after the switches are resolved
      meta.normalize_inputs(q, k, v, None, 0.0, causal, None, False)
    assert meta.forward_meta.fp8_hybrid is want
  meta = FFPAAttnMeta.from_kwargs(forward_backend=CUDABackend(enable_fp8=True, fp8_hybrid=False))
  with pytest.raises(RuntimeError, match="no CPU"):
    meta.normalize_inputs(q, k, v, None, 0.0, True, None, False)
  assert meta.forward_meta.fp8_hybrid is False


def test_varlen_fake_ops_registered_for_compile():
  """torch.compile sees the packed ops through their fake (shape-only) implementations."""
  import ffpa_attn  # noqa: F401

  q = torch.empty(48, 4, 128, dtype=torch.float16, device="meta")
  k = torch.empty(64, 2, 128, dtype=torch.float16, device="meta")
  cu = torch.empty(3, dtype=torch.int32, device="meta")
  o, lse = torch.ops.ffpa_attn._varlen_fwd_cuda(q, k, k, cu, cu, 32, 40, 1, 0.1)
  assert o.shape == q.shape and o.dtype == q.dtype
  assert lse.shape == (4, 48) and lse.dtype == torch.float32
  dq, dk, dv = torch.ops.ffpa_attn._varlen_bwd_cuda(q, k, k, o, lse, o, cu, cu, 32, 40, 1, 0.1, None)
  assert dq.shape == q.shape and dk.shape == k.shape and dv.shape == k.shape
  dq, dk, dv = torch.ops.ffpa_attn._varlen_bwd_cuda(q, k, k, o, lse, o, cu, cu, 32, 40, 1, 0.1, lse)
  assert dq.shape == q.shape
