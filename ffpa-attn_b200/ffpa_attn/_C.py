"""``ffpa_attn._C`` -- the native binding of the package, B200 edition.

Exposes the same surface as the reference's pybind11 module
(/root/reference/csrc/cuffpa/ffpa_api.cc:265-306, imported at
/root/reference/src/ffpa_attn/cuda/__init__.py:6-25):

    ffpa_attn_forward(Q, K, V, attn_bias, O, softmax_lse, stages, acc, causal, softmax_scale,
                      dropout_p, philox_seed, philox_offset, <11 fp8/fp4 knobs>) -> None
    ffpa_attn_backward(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, stages, causal, softmax_scale)
    set_cuda_backend_impl(int) / get_cuda_backend_impl() -> int
    CUDA_FWD_AVAILABLE, CUDA_AVAILABLE, F16_ACC_AVAILABLE, CUDA_TMA_AVAILABLE,
    CUDA_CUTE_TMA_AVAILABLE, CUDA_BWD_AVAILABLE

but is a thin ctypes shim over the C ABI in ``include/ffpa_b200.h`` (``libffpa_b200.so``, built
in-tree by ``__graft_entry__.build()``).  There is exactly one backend -- the hand-written sm_100a
kernels -- and no fallback: if the library is missing, importing this module raises.
Errors follow the reference's convention (ffpa_api.cc:180-205, launch.cuh:79-129):
contract violations -> RuntimeError (TORCH_CHECK), dtype/acc -> ValueError.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libffpa_b200.so")

if not os.path.exists(_LIB_PATH):
  raise ImportError(
    f"ffpa_attn._C: {_LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; "
    "g.build()'` (or `make -C ffpa-attn_b200/csrc`). There is no CPU / Triton / SDPA fallback."
  )

_lib = ctypes.CDLL(_LIB_PATH)


class _FwdParams(ctypes.Structure):
  _fields_ = [
    ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("o", ctypes.c_void_p),
    ("lse", ctypes.c_void_p), ("bias", ctypes.c_void_p),
    ("q_stride", ctypes.c_int64 * 4), ("k_stride", ctypes.c_int64 * 4),
    ("v_stride", ctypes.c_int64 * 4), ("o_stride", ctypes.c_int64 * 4),
    ("bias_stride", ctypes.c_int64 * 4),
    ("batch", ctypes.c_int32), ("heads_q", ctypes.c_int32), ("heads_kv", ctypes.c_int32),
    ("seqlen_q", ctypes.c_int32), ("seqlen_kv", ctypes.c_int32), ("head_dim", ctypes.c_int32),
    ("dtype", ctypes.c_int32), ("bias_kind", ctypes.c_int32), ("causal", ctypes.c_int32),
    ("fp8", ctypes.c_int32),
    ("softmax_scale", ctypes.c_float), ("dropout_p", ctypes.c_float),
    ("philox_seed", ctypes.c_uint64), ("philox_offset", ctypes.c_uint64),
    ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_uint64),
    ("cu_seqlens_q", ctypes.c_void_p), ("cu_seqlens_k", ctypes.c_void_p),
    ("total_q", ctypes.c_int32), ("total_k", ctypes.c_int32),
  ]


class _BwdParams(ctypes.Structure):
  _fields_ = [
    ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p), ("o", ctypes.c_void_p),
    ("lse", ctypes.c_void_p), ("d_o", ctypes.c_void_p),
    ("dq", ctypes.c_void_p), ("dk", ctypes.c_void_p), ("dv", ctypes.c_void_p),
    ("q_stride", ctypes.c_int64 * 4), ("k_stride", ctypes.c_int64 * 4),
    ("v_stride", ctypes.c_int64 * 4), ("o_stride", ctypes.c_int64 * 4),
    ("do_stride", ctypes.c_int64 * 4), ("dq_stride", ctypes.c_int64 * 4),
    ("dk_stride", ctypes.c_int64 * 4), ("dv_stride", ctypes.c_int64 * 4),
    ("batch", ctypes.c_int32), ("heads_q", ctypes.c_int32), ("heads_kv", ctypes.c_int32),
    ("seqlen_q", ctypes.c_int32), ("seqlen_kv", ctypes.c_int32), ("head_dim", ctypes.c_int32),
    ("dtype", ctypes.c_int32), ("causal", ctypes.c_int32),
    ("softmax_scale", ctypes.c_float),
    ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_uint64),
    ("bias", ctypes.c_void_p), ("bias_stride", ctypes.c_int64 * 4), ("bias_kind", ctypes.c_int32),
    ("dropout_p", ctypes.c_float), ("philox_seed", ctypes.c_uint64), ("philox_offset", ctypes.c_uint64),
    ("d_bias", ctypes.c_void_p),
    ("cu_seqlens_q", ctypes.c_void_p), ("cu_seqlens_k", ctypes.c_void_p),
    ("total_q", ctypes.c_int32), ("total_k", ctypes.c_int32),
    ("d_lse", ctypes.c_void_p),
  ]


_lib.ffpa_b200_fwd.argtypes = [ctypes.POINTER(_FwdParams), ctypes.c_void_p]
_lib.ffpa_b200_fwd.restype = ctypes.c_int
_lib.ffpa_b200_bwd.argtypes = [ctypes.POINTER(_BwdParams), ctypes.c_void_p]
_lib.ffpa_b200_bwd.restype = ctypes.c_int
_lib.ffpa_b200_fwd_workspace_bytes.argtypes = [ctypes.c_int32] * 7
_lib.ffpa_b200_fwd_workspace_bytes.restype = ctypes.c_uint64
_lib.ffpa_b200_bwd_workspace_bytes.argtypes = [ctypes.c_int32] * 6
_lib.ffpa_b200_bwd_workspace_bytes.restype = ctypes.c_uint64
_lib.ffpa_b200_bwd_workspace_bytes_min.argtypes = [ctypes.c_int32] * 6
_lib.ffpa_b200_bwd_workspace_bytes_min.restype = ctypes.c_uint64
_lib.ffpa_b200_set_backend_impl.argtypes = [ctypes.c_int32]
_lib.ffpa_b200_set_backend_impl.restype = ctypes.c_int
_lib.ffpa_b200_get_backend_impl.restype = ctypes.c_int32
_lib.ffpa_b200_fwd_available.restype = ctypes.c_int32
_lib.ffpa_b200_bwd_available.restype = ctypes.c_int32
_lib.ffpa_b200_abi_version.restype = ctypes.c_int32
_lib.ffpa_b200_launch_count.restype = ctypes.c_uint64
_lib.ffpa_b200_last_error.restype = ctypes.c_char_p

ABI_VERSION = int(_lib.ffpa_b200_abi_version())
if ABI_VERSION != 2:
  raise ImportError(f"ffpa_attn._C: libffpa_b200.so ABI {ABI_VERSION} != 2")

# module attributes of the reference binding (ffpa_api.cc:283-305)
CUDA_FWD_AVAILABLE = bool(_lib.ffpa_b200_fwd_available())
CUDA_AVAILABLE = CUDA_FWD_AVAILABLE
CUDA_BWD_AVAILABLE = bool(_lib.ffpa_b200_bwd_available())
F16_ACC_AVAILABLE = False       # tcgen05 accumulates in fp32 (TMEM); there is no f16-acc variant
CUDA_TMA_AVAILABLE = True       # every kernel here is TMA-fed
CUDA_CUTE_TMA_AVAILABLE = False  # no CuTe / CUTLASS code in this build

_ERR_INVALID, _ERR_UNSUPPORTED, _ERR_CUDA, _ERR_NO_DEVICE = -1, -2, -3, -4


def _raise(code: int) -> None:
  msg = (_lib.ffpa_b200_last_error() or b"").decode()
  if code == _ERR_UNSUPPORTED:
    raise NotImplementedError(f"ffpa_attn._C: {msg}")
  raise RuntimeError(f"ffpa_attn._C: {msg}")


def _dtype_code(t: torch.Tensor) -> int:
  if t.dtype == torch.float16:
    return 0
  if t.dtype == torch.bfloat16:
    return 1
  # std::invalid_argument -> ValueError in the reference binding (ffpa_api.cc:235-237)
  raise ValueError(f"ffpa_attn_forward only supports fp16/bf16 tensors, got {t.dtype}")


def _check_cuda(*ts: torch.Tensor) -> None:
  dev = ts[0].device
  for t in ts:
    if t.device.type != "cuda":
      raise RuntimeError("ffpa_attn._C: all tensors must be CUDA tensors (no CPU path exists)")
    if t.device != dev:
      raise RuntimeError("ffpa_attn._C: all tensors must live on the same device")


def _strides4(t: torch.Tensor):
  return (ctypes.c_int64 * 4)(*[int(s) for s in t.stride()])


def _bias_fields(attn_bias, Q, K):
  """(tensor kept alive, bias_kind, strides[4], data_ptr) following native/launch.cuh:277-290:
  4-D [1|B, 1|Hq, 1|Nq, 1|Nkv], fp32 or Q's dtype, last dim contiguous, broadcast dims stride 0."""
  if attn_bias is None or attn_bias.numel() == 0:
    return None, 0, (ctypes.c_int64 * 4)(0, 0, 0, 0), None
  _check_cuda(Q, attn_bias)
  if attn_bias.dim() != 4:
    raise RuntimeError("ffpa_attn: attn_bias must be 4-D [1|B, 1|Hq, 1|Nq, 1|Nkv]")
  want = (Q.size(0), Q.size(1), Q.size(2), K.size(2))
  for i in range(4):
    if attn_bias.size(i) not in (1, want[i]):
      raise RuntimeError(f"ffpa_attn: attn_bias dim {i} must be 1 or {want[i]}")
  if attn_bias.dtype == torch.float32:
    kind = 1
  elif attn_bias.dtype == Q.dtype:
    kind = 2
  else:
    raise RuntimeError("ffpa_attn: attn_bias dtype must be float32 or match Q")
  if attn_bias.size(3) == 1 and K.size(2) != 1:
    attn_bias = attn_bias.expand(-1, -1, -1, K.size(2)).contiguous()
  elif attn_bias.stride(3) != 1:
    raise RuntimeError("ffpa_attn: attn_bias last dim must be contiguous")
  bs = [0 if attn_bias.size(i) == 1 else int(attn_bias.stride(i)) for i in range(3)] + [1]
  return attn_bias, kind, (ctypes.c_int64 * 4)(*bs), attn_bias.data_ptr()


def launch_count() -> int:
  """Kernels launched by the library since load (bench.py reports it as ``gpu_launches``)."""
  return int(_lib.ffpa_b200_launch_count())


def ffpa_attn_forward(Q, K, V, attn_bias, O, softmax_lse, stages, acc, causal, softmax_scale,
                      dropout_p, philox_seed, philox_offset, fp8_smooth_k=True, fp8_smooth_v=False,
                      fp8_q_quant_method=0, fp8_k_quant_method=0, fp8_v_quant_method=0,
                      fp8_pv_acc_type=0, fp8_qk_mm_type=0, fp8_hybrid=False, fp8_hybrid_n_early=256,
                      fp4_hybrid=False, fp4_hybrid_n_early=256) -> None:
  """Writes ``O`` and ``softmax_lse`` in place (caller-allocated), like ffpa_api.cc:86-239.

  ``stages`` is accepted for signature compatibility (the sm_100a pipeline depth is static);
  ``acc`` must be 1 (f32). ``attn_bias.numel() == 0`` means "no bias".
  """
  _check_cuda(Q, K, V, O)
  if Q.dim() != 4 or K.dim() != 4 or V.dim() != 4:
    raise RuntimeError("ffpa_attn_forward: Q/K/V must be 4-D [B, H, N, D]")
  if K.dtype != Q.dtype or V.dtype != Q.dtype or O.dtype != Q.dtype:
    raise RuntimeError("ffpa_attn_forward: Q/K/V/O must share one dtype")
  dt = _dtype_code(Q)
  if int(acc) == 0:
    raise ValueError("acc='f16' is not available: tcgen05 accumulates in fp32")
  if O.shape != Q.shape:
    raise RuntimeError("ffpa_attn_forward: O must have the shape of Q")
  if K.shape != V.shape or K.size(0) != Q.size(0) or K.size(3) != Q.size(3):
    raise RuntimeError("ffpa_attn_forward: K/V must be [B, Hkv, Nkv, D] matching Q's B and D")
  # the reference kernels index with dense row-major offsets (native/sm_80/split_d.cuh:137-142);
  # we honour strides through the tensor maps but need unit stride on D.
  if Q.stride(3) != 1:
    Q = Q.contiguous()
  if K.stride(3) != 1:
    K = K.contiguous()
  if V.stride(3) != 1:
    V = V.contiguous()
  if O.stride(3) != 1:
    raise RuntimeError("ffpa_attn_forward: O must have unit stride on the head dim")

  # FP8 hybrid (reference: csrc/cuffpa/launch.cuh:341-374): the first n_early query rows run on the fp16/bf16
  # kernel (they see few keys, so quantisation error weighs most there), rows [n_early, Nq) on the FP8 kernel.
  # Bottom-right causal alignment makes both stages ordinary calls on zero-copy row views: stage 1 = the early
  # rows against the keys they can see, stage 2 = the late rows against all keys.
  if int(_lib.ffpa_b200_get_backend_impl()) == 5 and fp8_hybrid and int(causal) and Q.size(2) > int(fp8_hybrid_n_early) > 0:
    n_early = int(fp8_hybrid_n_early)
    if n_early % 128 != 0:
      raise RuntimeError("ffpa_attn: fp8_hybrid_n_early must be multiple of 128")
    nk_early = n_early + K.size(2) - Q.size(2)
    want_lse = softmax_lse is not None and softmax_lse.numel() > 0
    lse_e = torch.empty(Q.size(0), Q.size(1), n_early, dtype=torch.float32, device=Q.device) if want_lse else None
    lse_l = torch.empty(Q.size(0), Q.size(1), Q.size(2) - n_early, dtype=torch.float32, device=Q.device) if want_lse else None
    common = (stages, acc, causal, softmax_scale, dropout_p, philox_seed, philox_offset, fp8_smooth_k, fp8_smooth_v,
              fp8_q_quant_method, fp8_k_quant_method, fp8_v_quant_method, fp8_pv_acc_type, fp8_qk_mm_type)
    _lib.ffpa_b200_set_backend_impl(0)
    try:
      ffpa_attn_forward(Q[:, :, :n_early], K[:, :, :nk_early], V[:, :, :nk_early], attn_bias, O[:, :, :n_early], lse_e,
                        *common, False, n_early, fp4_hybrid, fp4_hybrid_n_early)
    finally:
      _lib.ffpa_b200_set_backend_impl(5)
    ffpa_attn_forward(Q[:, :, n_early:], K, V, attn_bias, O[:, :, n_early:], lse_l, *common, False, n_early,
                      fp4_hybrid, fp4_hybrid_n_early)
    if want_lse:
      softmax_lse[:, :, :n_early].copy_(lse_e)
      softmax_lse[:, :, n_early:].copy_(lse_l)
    return

  p = _FwdParams()
  p.q, p.k, p.v, p.o = Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr()
  if softmax_lse is not None and softmax_lse.numel() > 0:
    if softmax_lse.dtype != torch.float32 or not softmax_lse.is_contiguous() or \
        tuple(softmax_lse.shape) != (Q.size(0), Q.size(1), Q.size(2)):
      raise RuntimeError("ffpa_attn_forward: softmax_lse must be contiguous fp32 [B, Hq, Nq]")
    p.lse = softmax_lse.data_ptr()
  else:
    p.lse = None
  p.q_stride, p.k_stride, p.v_stride, p.o_stride = _strides4(Q), _strides4(K), _strides4(V), _strides4(O)
  _bias_keep, p.bias_kind, p.bias_stride, p.bias = _bias_fields(attn_bias, Q, K)
  p.batch, p.heads_q, p.seqlen_q, p.head_dim = Q.size(0), Q.size(1), Q.size(2), Q.size(3)
  p.heads_kv, p.seqlen_kv = K.size(1), K.size(2)
  p.dtype = dt
  p.causal = int(causal)
  # bit 0: FP8 path (backend hint CUTE_TMA_FP8); bit 1: smooth-K (fp8_smooth_k, default True as in the reference)
  # fp8_v_quant_method: 0 per_block, 1 per_channel (codes of functional._QUANT_CODE)
  p.fp8 = (1 | (2 if fp8_smooth_k else 0) | (4 if fp8_smooth_v else 0) | (8 if int(fp8_v_quant_method) == 1 else 0)) \
      if int(_lib.ffpa_b200_get_backend_impl()) == 5 else 0
  p.softmax_scale = float(softmax_scale)
  p.dropout_p = float(dropout_p)
  p.philox_seed = int(philox_seed) & 0xFFFFFFFFFFFFFFFF
  p.philox_offset = int(philox_offset) & 0xFFFFFFFFFFFFFFFF
  ws = None
  if not p.fp8:
    # decode-like shapes (few query tiles, long KV) use KV splits: fp32 partials live in this scratch
    nbytes = int(_lib.ffpa_b200_fwd_workspace_bytes(p.batch, p.heads_q, p.heads_kv, p.seqlen_q,
                                                     p.seqlen_kv, p.head_dim, 0))
    if nbytes > 0:
      ws = torch.empty(nbytes, dtype=torch.uint8, device=Q.device)
      p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
  if p.fp8:
    # FP8 path (backend hint CUTE_TMA_FP8): scratch for the e4m3 copies of Q/K/V and their scales.
    # The fp8_* knobs of the reference signature select sm_120 variants (per-thread scales, int8 QK,
    # smooth-K/V, hybrid early rows); the sm_100a kernel implements per-block e4m3 only.
    nbytes = int(_lib.ffpa_b200_fwd_workspace_bytes(p.batch, p.heads_q, p.heads_kv, p.seqlen_q,
                                                     p.seqlen_kv, p.head_dim, 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=Q.device)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
  with torch.cuda.device(Q.device):
    stream = torch.cuda.current_stream(Q.device).cuda_stream
    rc = _lib.ffpa_b200_fwd(ctypes.byref(p), ctypes.c_void_p(stream))
  if rc != 0:
    _raise(rc)


def ffpa_attn_backward(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, stages, causal, softmax_scale,
                       attn_bias=None, dropout_p=0.0, philox_seed=0, philox_offset=0, d_bias=None,
                       min_workspace=False, d_lse=None) -> None:
  """Positional signature of ffpa_api.cc:242-246 (a thrower in the reference); real here.
  Keyword extras replay what the forward applied: additive ``attn_bias``, dropout (same Philox
  seed/offset) and ``d_bias`` -- an fp32 [B, Hq, Nq, Nkv] buffer that receives dS per score."""
  _check_cuda(Q, K, V, O, dO, dQ, dK, dV, softmax_lse)
  dt = _dtype_code(Q)
  p = _BwdParams()
  p.q, p.k, p.v, p.o = Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr()
  p.lse, p.d_o = softmax_lse.data_ptr(), dO.data_ptr()
  p.dq, p.dk, p.dv = dQ.data_ptr(), dK.data_ptr(), dV.data_ptr()
  for name, t in (("q_stride", Q), ("k_stride", K), ("v_stride", V), ("o_stride", O),
                  ("do_stride", dO), ("dq_stride", dQ), ("dk_stride", dK), ("dv_stride", dV)):
    if t.stride(3) != 1:
      raise RuntimeError("ffpa_attn_backward: all tensors need unit stride on the head dim")
    setattr(p, name, _strides4(t))
  p.batch, p.heads_q, p.seqlen_q, p.head_dim = Q.size(0), Q.size(1), Q.size(2), Q.size(3)
  p.heads_kv, p.seqlen_kv = K.size(1), K.size(2)
  p.dtype = dt
  p.causal = int(causal)
  p.softmax_scale = float(softmax_scale)
  # recommended size: includes the score stash of the 5-GEMM path for head dims 384..512; with
  # ``min_workspace`` (O(N) memory) the three recompute kernels run instead
  ws_fn = _lib.ffpa_b200_bwd_workspace_bytes_min if min_workspace else _lib.ffpa_b200_bwd_workspace_bytes
  nbytes = int(ws_fn(p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim))
  ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=Q.device)
  p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
  _bias_keep, p.bias_kind, p.bias_stride, p.bias = _bias_fields(attn_bias, Q, K)
  p.dropout_p = float(dropout_p)
  p.philox_seed = int(philox_seed) & 0xFFFFFFFFFFFFFFFF
  p.philox_offset = int(philox_offset) & 0xFFFFFFFFFFFFFFFF
  if d_bias is not None:
    if d_bias.dtype != torch.float32 or not d_bias.is_contiguous() or \
        tuple(d_bias.shape) != (Q.size(0), Q.size(1), Q.size(2), K.size(2)):
      raise RuntimeError("ffpa_attn_backward: d_bias must be contiguous fp32 [B, Hq, Nq, Nkv]")
    p.d_bias = d_bias.data_ptr()
  else:
    p.d_bias = None
  if d_lse is not None:
    if d_lse.dtype != torch.float32 or not d_lse.is_contiguous() or d_lse.shape != softmax_lse.shape:
      raise RuntimeError("ffpa_attn_backward: d_lse must be contiguous fp32 with the shape of softmax_lse")
    p.d_lse = d_lse.data_ptr()
  with torch.cuda.device(Q.device):
    stream = torch.cuda.current_stream(Q.device).cuda_stream
    rc = _lib.ffpa_b200_bwd(ctypes.byref(p), ctypes.c_void_p(stream))
  if rc != 0:
    _raise(rc)


def _strides_thd(t: torch.Tensor):
  """[T, H, D] packed tensor -> (b, h, n, d) element strides of the C ABI (batch stride unused)."""
  return (ctypes.c_int64 * 4)(0, int(t.stride(1)), int(t.stride(0)), int(t.stride(2)))


def _check_varlen(Q, K, V, cu_q, cu_k):
  if Q.dim() != 3 or K.dim() != 3 or V.dim() != 3:
    raise RuntimeError("ffpa_attn varlen: q/k/v must be packed [T, H, D] tensors")
  if K.shape != V.shape or K.size(2) != Q.size(2):
    raise RuntimeError("ffpa_attn varlen: k and v must share [T_k, H_kv, D] and q's head dim")
  for cu in (cu_q, cu_k):
    if cu.dtype != torch.int32 or cu.dim() != 1 or not cu.is_contiguous() or cu.device != Q.device:
      raise RuntimeError("ffpa_attn varlen: cu_seqlens must be contiguous int32 1-D tensors on q's device")
  if cu_q.numel() != cu_k.numel() or cu_q.numel() < 2:
    raise RuntimeError("ffpa_attn varlen: cu_seqlens_q / cu_seqlens_k must both have B + 1 entries")


def ffpa_attn_varlen_forward(Q, K, V, O, softmax_lse, cu_seqlens_q, cu_seqlens_k, max_seqlen_q, max_seqlen_k,
                             causal, softmax_scale) -> None:
  """Packed variable-length forward in ONE launch (C ABI 2, ``ffpa_fwd_params.cu_seqlens_*``): Q/O
  [T_q, Hq, D], K/V [T_k, Hkv, D], LSE fp32 [Hq, T_q]. ``cu_seqlens`` stay on the device (no host sync);
  ``max_seqlen_*`` size the grid. Backend of ``ffpa_attn_varlen_func``
  (/root/reference/src/ffpa_attn/ffpa_attn_interface.py:192-279)."""
  _check_cuda(Q, K, V, O, cu_seqlens_q, cu_seqlens_k)
  _check_varlen(Q, K, V, cu_seqlens_q, cu_seqlens_k)
  if K.dtype != Q.dtype or V.dtype != Q.dtype or O.dtype != Q.dtype or O.shape != Q.shape:
    raise RuntimeError("ffpa_attn varlen: Q/K/V/O must share one dtype and O the shape of Q")
  for t in (Q, K, V, O):
    if t.stride(2) != 1:
      raise RuntimeError("ffpa_attn varlen: unit stride on the head dim is required")
  p = _FwdParams()
  p.q, p.k, p.v, p.o = Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr()
  if softmax_lse is not None and softmax_lse.numel() > 0:
    if softmax_lse.dtype != torch.float32 or not softmax_lse.is_contiguous() or \
        tuple(softmax_lse.shape) != (Q.size(1), Q.size(0)):
      raise RuntimeError("ffpa_attn varlen: softmax_lse must be contiguous fp32 [Hq, T_q]")
    p.lse = softmax_lse.data_ptr()
  p.q_stride, p.k_stride, p.v_stride, p.o_stride = _strides_thd(Q), _strides_thd(K), _strides_thd(V), _strides_thd(O)
  p.bias_stride = (ctypes.c_int64 * 4)(0, 0, 0, 0)
  p.batch, p.heads_q, p.heads_kv, p.head_dim = cu_seqlens_q.numel() - 1, Q.size(1), K.size(1), Q.size(2)
  p.seqlen_q, p.seqlen_kv = int(max_seqlen_q), int(max_seqlen_k)
  p.total_q, p.total_k = Q.size(0), K.size(0)
  p.cu_seqlens_q, p.cu_seqlens_k = cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr()
  p.dtype = _dtype_code(Q)
  p.causal = int(causal)
  p.softmax_scale = float(softmax_scale)
  with torch.cuda.device(Q.device):
    stream = torch.cuda.current_stream(Q.device).cuda_stream
    rc = _lib.ffpa_b200_fwd(ctypes.byref(p), ctypes.c_void_p(stream))
  if rc != 0:
    _raise(rc)


def ffpa_attn_varlen_backward(Q, K, V, O, softmax_lse, dO, dQ, dK, dV, cu_seqlens_q, cu_seqlens_k,
                              max_seqlen_q, max_seqlen_k, causal, softmax_scale, d_lse=None) -> None:
  """Packed variable-length backward (preprocess + dQ + dK + dV launches for the whole batch)."""
  _check_cuda(Q, K, V, O, dO, dQ, dK, dV, softmax_lse, cu_seqlens_q, cu_seqlens_k)
  _check_varlen(Q, K, V, cu_seqlens_q, cu_seqlens_k)
  p = _BwdParams()
  p.q, p.k, p.v, p.o = Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr()
  p.lse, p.d_o = softmax_lse.data_ptr(), dO.data_ptr()
  p.dq, p.dk, p.dv = dQ.data_ptr(), dK.data_ptr(), dV.data_ptr()
  for name, t in (("q_stride", Q), ("k_stride", K), ("v_stride", V), ("o_stride", O),
                  ("do_stride", dO), ("dq_stride", dQ), ("dk_stride", dK), ("dv_stride", dV)):
    if t.stride(2) != 1:
      raise RuntimeError("ffpa_attn varlen backward: all tensors need unit stride on the head dim")
    setattr(p, name, _strides_thd(t))
  p.bias_stride = (ctypes.c_int64 * 4)(0, 0, 0, 0)
  p.batch, p.heads_q, p.heads_kv, p.head_dim = cu_seqlens_q.numel() - 1, Q.size(1), K.size(1), Q.size(2)
  p.seqlen_q, p.seqlen_kv = int(max_seqlen_q), int(max_seqlen_k)
  p.total_q, p.total_k = Q.size(0), K.size(0)
  p.cu_seqlens_q, p.cu_seqlens_k = cu_seqlens_q.data_ptr(), cu_seqlens_k.data_ptr()
  p.dtype = _dtype_code(Q)
  p.causal = int(causal)
  p.softmax_scale = float(softmax_scale)
  if d_lse is not None:
    if d_lse.dtype != torch.float32 or not d_lse.is_contiguous() or d_lse.shape != softmax_lse.shape:
      raise RuntimeError("ffpa_attn varlen backward: d_lse must be contiguous fp32 [Hq, T_q]")
    p.d_lse = d_lse.data_ptr()
  nbytes = int(_lib.ffpa_b200_bwd_workspace_bytes_min(p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim))
  ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=Q.device)
  p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
  with torch.cuda.device(Q.device):
    stream = torch.cuda.current_stream(Q.device).cuda_stream
    rc = _lib.ffpa_b200_bwd(ctypes.byref(p), ctypes.c_void_p(stream))
  if rc != 0:
    _raise(rc)


def set_cuda_backend_impl(impl: int) -> None:
  rc = _lib.ffpa_b200_set_backend_impl(int(impl))
  if rc != 0:
    _raise(rc)


def get_cuda_backend_impl() -> int:
  return int(_lib.ffpa_b200_get_backend_impl())
