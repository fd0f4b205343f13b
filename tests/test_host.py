"""CPU tests for the boundary: the C-ABI library loads and exports every symbol declared in
include/ffpa_b200.h, the ctypes structs (tests/capi.py) match the C layout, the PyTorch extension
``ffpa_attn._C`` exports the reference module's pybind surface (ffpa_api.cc:265-306), and the host-side mirror
of the reference API validates inputs with the reference's error classes. No kernel is launched."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import capi  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ffpa_b200.h")
LIB = os.path.join(ROOT, "ffpa-attn_b200", "ffpa_attn", "libffpa_b200.so")


@pytest.fixture(scope="module")
def lib():
  if not os.path.exists(LIB):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.build()
  return capi.load()


def _declared_symbols():
  src = open(HEADER).read()
  return sorted(set(re.findall(r"\b(ffpa_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
  syms = _declared_symbols()
  for s in ("ffpa_b200_fwd", "ffpa_b200_bwd", "ffpa_b200_set_backend_impl", "ffpa_b200_get_backend_impl",
            "ffpa_b200_last_error", "ffpa_b200_launch_count", "ffpa_b200_bwd_workspace_bytes_p",
            "ffpa_b200_bwd_workspace_bytes_min_p", "ffpa_b200_fwd_workspace_bytes_p"):
    assert s in syms


def test_library_exports_every_declared_symbol(lib):
  for s in _declared_symbols():
    assert hasattr(lib, s), f"libffpa_b200.so does not export {s}"


def test_abi_version_and_flags(lib):
  lib.ffpa_b200_abi_version.restype = ctypes.c_int32
  assert lib.ffpa_b200_abi_version() == 3
  assert lib.ffpa_b200_fwd_available() == 1


def test_ctypes_struct_layout_matches_c():
  C = capi

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
  printf("%zu %zu %zu %zu %zu ", offsetof(ffpa_fwd_params, impl), offsetof(ffpa_fwd_params, fp8_smooth_k),
         offsetof(ffpa_fwd_params, fp8_qk_mm_type), offsetof(ffpa_fwd_params, fp8_hybrid_n_early),
         offsetof(ffpa_fwd_params, lse_bh_stride));
  printf("%zu ", offsetof(ffpa_bwd_params, d_bias_stride));
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
  F, B = C.FwdParams, C.BwdParams
  want = [ctypes.sizeof(F), F.bias_stride.offset, F.batch.offset, F.softmax_scale.offset,
          F.philox_seed.offset, F.philox_offset.offset, F.workspace_bytes.offset,
          F.cu_seqlens_q.offset, F.total_k.offset,
          F.impl.offset, F.fp8_smooth_k.offset, F.fp8_qk_mm_type.offset, F.fp8_hybrid_n_early.offset,
          F.lse_bh_stride.offset, B.d_bias_stride.offset,
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
  p = capi.FwdParams()
  rc = lib.ffpa_b200_fwd(ctypes.byref(p), None)
  assert rc < 0
  assert len(lib.ffpa_b200_last_error()) > 0
  assert lib.ffpa_b200_fwd(None, None) == -1
  assert lib.ffpa_b200_bwd(None, None) == -1


def test_torch_extension_exports_the_reference_pybind_surface():
  """ffpa_attn._C is a compiled PyTorch extension (not a Python shim) with the names, attributes and argument
  counts of /root/reference/csrc/cuffpa/ffpa_api.cc:265-306 (forward: 24 positional arguments, :86-96;
  backward: 12, :242-246)."""
  import ffpa_attn._C as C

  assert C.__file__.endswith(".so")
  for name in ("ffpa_attn_forward", "ffpa_attn_backward", "set_cuda_backend_impl", "get_cuda_backend_impl"):
    assert callable(getattr(C, name))
  for attr, want in (("CUDA_FWD_AVAILABLE", True), ("CUDA_AVAILABLE", True), ("F16_ACC_AVAILABLE", False),
                     ("CUDA_TMA_AVAILABLE", True), ("CUDA_CUTE_TMA_AVAILABLE", False), ("CUDA_BWD_AVAILABLE", True)):
    assert getattr(C, attr) is want
  assert C.ABI_VERSION == 3
  doc = C.ffpa_attn_forward.__doc__.split("->")[0]
  assert doc.count("arg") == 24, doc
  assert C.ffpa_attn_backward.__doc__.split("->")[0].count("arg") == 12
  t = torch.zeros(1, 1, 8, 64, dtype=torch.bfloat16)
  e = torch.empty(0)
  with pytest.raises(RuntimeError, match="CUDA tensors"):   # CPU tensors: no fallback path
    C.ffpa_attn_forward(t, t, t, e, t.clone(), e, 0, 1, 0, 0.125, 0.0, 0, 0, True, False, 0, 0, 0, 1, 0, False, 256, False, 256)


def test_backend_hint_is_thread_local(lib):
  """The reference keeps one process-global atomic (backend.h:16-25): two threads using different backends
  race. Here the hint set by a thread is visible to that thread only."""
  import threading

  import ffpa_attn._C as C

  C.set_cuda_backend_impl(5)
  seen = []
  t = threading.Thread(target=lambda: seen.append(C.get_cuda_backend_impl()))
  t.start()
  t.join()
  assert seen == [0] and C.get_cuda_backend_impl() == 5
  C.set_cuda_backend_impl(0)


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
                                         0, 0, 0, 1, 0, False, 256, False, 256)
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
  with pytest.raises(NotImplementedError, match="only backend"):   # same class as the dense entry
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, backend="cutedsl")
  with pytest.raises(TypeError):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, backend=3)
  # the defaults flash-attn style callers pass are accepted (reference: cute/__init__.py:107-118) ...
  with pytest.raises(RuntimeError, match="no CPU"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, window_size=(None, None), softcap=0.0)
  # ... anything that would change the result is refused
  with pytest.raises(NotImplementedError, match="softcap"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, softcap=30.0)
  with pytest.raises(NotImplementedError, match="window_size"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16, window_size=(128, 0))
  with pytest.raises(RuntimeError, match="no CPU"):
    ffpa_attn_varlen_func(q, k, k, cu, cu, 16, 16)


def test_backward_workspace_plan(lib):
  """_min_p = required O(N) scratch; _p(params, cap) = recommended size under `cap`: for head dims 384..1024 it
  adds the two 16-bit score buffers of the stash path, cut into (batch, KV-head) chunks when they do not fit,
  and degrades to the minimum when not even a machine-filling chunk fits. cap == 0 asks for the minimum."""
  f, m = lib.ffpa_b200_bwd_workspace_bytes_p, lib.ffpa_b200_bwd_workspace_bytes_min_p
  big = 1 << 40

  def F(*a, cap=big):
    return f(ctypes.byref(capi.bwd_sizes(*a)), cap)

  def M(*a):
    return m(ctypes.byref(capi.bwd_sizes(*a)))

  B, H, N, D = 1, 32, 8192, 512
  stash = 2 * B * H * N * N * 2
  assert F(B, H, H, N, N, D) == M(B, H, H, N, N, D) + stash
  assert F(B, H, H, N, N, 256) == M(B, H, H, N, N, 256)            # small heads: recompute kernels
  assert F(B, H, H, N, N, 1024) == M(B, H, H, N, N, 1024) + stash   # large heads: stash on the first slab pass
  assert F(1, 4, 2, 130, 257, 512) == M(1, 4, 2, 130, 257, 512) + 2 * 4 * 256 * 512 * 2   # padded to 128 x 256
  assert F(B, H, H, N, N, D, cap=0) == M(B, H, H, N, N, D)
  # 4 batch elements need 34 GB: under a 20 GB cap the plan is chunked (one batch element at a time)
  cap = 20 << 30
  got = F(4, H, H, N, N, D, cap=cap)
  assert stash <= got <= cap
  # a cap of 3 GB still fits a machine-filling chunk of KV heads; 100 MB does not -> minimum
  got = F(B, H, H, N, N, D, cap=3 << 30)
  assert M(B, H, H, N, N, D) < got <= 3 << 30
  assert F(B, H, H, N, N, D, cap=100 << 20) == M(B, H, H, N, N, D)
  # the plan is monotone in the cap
  sizes = [F(2, 16, 4, 4096, 4096, 512, cap=c << 30) for c in (0, 1, 2, 4, 8, 64)]
  assert sizes == sorted(sizes)


def test_forward_workspace_plan(lib, monkeypatch):
  """Forward scratch: KV-split partials for decode-like shapes, the replay stash for head dims > 768, FP8 copies,
  and for FP8 hybrid the larger of its two stages."""
  f = lib.ffpa_b200_fwd_workspace_bytes_p

  def F(*a, cap=1 << 40, **kw):
    return f(ctypes.byref(capi.fwd_sizes(*a, **kw)), cap)

  assert F(1, 32, 32, 8192, 8192, 512) == 0
  assert F(1, 32, 32, 8192, 8192, 768) == 0
  need = F(1, 32, 32, 8192, 8192, 1024)
  if os.environ.get("FFPA_FWD_REPLAY", "1") != "0":
    assert 32 * 8192 * 8192 * 2 <= need <= 32 * 8192 * 8192 * 2 * 1.02   # P tiles + factors + 1/rowsum
  if os.environ.get("FFPA_FWD_REPLAY", "1") != "0":
    # bounded scratch: 2.5 GiB holds a machine-filling chunk of heads, 100 MB holds nothing -> two-pass kernel
    assert 0 < F(1, 32, 32, 8192, 8192, 1024, cap=int(2.5 * 2 ** 30)) <= 2.5 * 2 ** 30
    assert F(1, 32, 32, 8192, 8192, 1024, cap=100 << 20) == 0
  assert F(1, 32, 32, 1, 8192, 512) > 0     # decode: KV-split partials
  fp8 = F(1, 32, 32, 8192, 8192, 256, impl=5, cap=0)   # required scratch: returned whatever the cap
  assert fp8 > 3 * 32 * 8192 * 256           # FP8: e4m3 copies of Q, K, V + scales
  hyb = F(1, 32, 32, 8192, 8192, 256, impl=5, causal=1, fp8_hybrid=1, fp8_hybrid_n_early=256)
  assert 0 < hyb <= fp8                      # stage 2 quantises 256 fewer query rows
  assert F(1, 32, 32, 8192, 8192, 256, impl=5, fp8_q_quant_method=2) == 0   # refused knob: nothing to plan


def test_unsupported_fp8_knobs_are_refused_by_name(lib):
  """per_thread Q/K scales, int8 QK and the f16 PV accumulator select sm_120 variants: the native layer refuses
  them (FFPA_ERR_UNSUPPORTED = -2, knob named) instead of ignoring them; so does CUDABackend."""
  from ffpa_attn import CUDABackend

  base = dict(q=4096, k=4096, v=4096, o=4096, dtype=1, impl=5)   # never dereferenced: the knob is refused before any launch
  for kw, word in ((dict(fp8_q_quant_method=2, fp8_k_quant_method=2), b"per_thread"), (dict(fp8_qk_mm_type=1), b"int8"),
                   (dict(fp8_pv_acc_type=0), b"f16")):
    p = capi.fwd_sizes(1, 2, 2, 256, 256, 128, **base)
    for i in range(4):
      p.q_stride[i] = p.k_stride[i] = p.v_stride[i] = p.o_stride[i] = (1 if i == 3 else 128)
    for k, v in kw.items():
      setattr(p, k, v)
    rc = lib.ffpa_b200_fwd(ctypes.byref(p), None)
    # without a GPU the device probe fails first (-4); with one the knob is refused (-2)
    assert rc in (-2, -4)
    if rc == -2:
      assert word in lib.ffpa_b200_last_error()
  for kw in (dict(fp8_q_quant_method="per_thread"), dict(fp8_qk_mm_type="int8"), dict(fp8_pv_acc_type="f16")):
    with pytest.raises(NotImplementedError):
      CUDABackend(enable_fp8=True, **kw)
    CUDABackend(**kw)   # harmless while FP8 is off, as in the reference
  with pytest.raises(ValueError, match="per_channel"):
    CUDABackend(enable_fp8=True, fp8_smooth_v=True)


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
