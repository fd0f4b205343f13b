"""ctypes view of the C ABI (include/ffpa_b200.h, ABI 3) used by the tests: struct mirrors + prototypes.
The product path binds the same ABI from C++ (ffpa-attn_b200/csrc/ffpa_torch_binding.cpp); this file is the
"any other FFI" example of INTEGRATION.md and lets the CPU suite check layouts, planning and argument contracts
without torch in the loop."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "ffpa-attn_b200", "ffpa_attn", "libffpa_b200.so")

i32, i64, u64, f32, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p


class FwdParams(ctypes.Structure):
  _fields_ = [
    ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp), ("bias", vp),
    ("q_stride", i64 * 4), ("k_stride", i64 * 4), ("v_stride", i64 * 4), ("o_stride", i64 * 4),
    ("bias_stride", i64 * 4),
    ("batch", i32), ("heads_q", i32), ("heads_kv", i32), ("seqlen_q", i32), ("seqlen_kv", i32), ("head_dim", i32),
    ("dtype", i32), ("bias_kind", i32), ("causal", i32), ("impl", i32),
    ("softmax_scale", f32), ("dropout_p", f32), ("philox_seed", u64), ("philox_offset", u64),
    ("workspace", vp), ("workspace_bytes", u64),
    ("cu_seqlens_q", vp), ("cu_seqlens_k", vp), ("total_q", i32), ("total_k", i32),
    ("fp8_smooth_k", i32), ("fp8_smooth_v", i32),
    ("fp8_q_quant_method", i32), ("fp8_k_quant_method", i32), ("fp8_v_quant_method", i32),
    ("fp8_pv_acc_type", i32), ("fp8_qk_mm_type", i32),
    ("fp8_hybrid", i32), ("fp8_hybrid_n_early", i32),
    ("lse_bh_stride", i64),
  ]


class BwdParams(ctypes.Structure):
  _fields_ = [
    ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp), ("d_o", vp), ("dq", vp), ("dk", vp), ("dv", vp),
    ("q_stride", i64 * 4), ("k_stride", i64 * 4), ("v_stride", i64 * 4), ("o_stride", i64 * 4),
    ("do_stride", i64 * 4), ("dq_stride", i64 * 4), ("dk_stride", i64 * 4), ("dv_stride", i64 * 4),
    ("batch", i32), ("heads_q", i32), ("heads_kv", i32), ("seqlen_q", i32), ("seqlen_kv", i32), ("head_dim", i32),
    ("dtype", i32), ("causal", i32), ("softmax_scale", f32),
    ("workspace", vp), ("workspace_bytes", u64),
    ("bias", vp), ("bias_stride", i64 * 4), ("bias_kind", i32),
    ("dropout_p", f32), ("philox_seed", u64), ("philox_offset", u64),
    ("d_bias", vp), ("d_bias_stride", i64 * 4),
    ("cu_seqlens_q", vp), ("cu_seqlens_k", vp), ("total_q", i32), ("total_k", i32),
    ("d_lse", vp),
  ]


def load():
  lib = ctypes.CDLL(LIB_PATH)
  lib.ffpa_b200_fwd.argtypes = [ctypes.POINTER(FwdParams), vp]
  lib.ffpa_b200_fwd.restype = ctypes.c_int
  lib.ffpa_b200_bwd.argtypes = [ctypes.POINTER(BwdParams), vp]
  lib.ffpa_b200_bwd.restype = ctypes.c_int
  lib.ffpa_b200_fwd_workspace_bytes_p.argtypes = [ctypes.POINTER(FwdParams), u64]
  lib.ffpa_b200_fwd_workspace_bytes_p.restype = u64
  lib.ffpa_b200_bwd_workspace_bytes_p.argtypes = [ctypes.POINTER(BwdParams), u64]
  lib.ffpa_b200_bwd_workspace_bytes_p.restype = u64
  lib.ffpa_b200_bwd_workspace_bytes_min_p.argtypes = [ctypes.POINTER(BwdParams)]
  lib.ffpa_b200_bwd_workspace_bytes_min_p.restype = u64
  lib.ffpa_b200_set_backend_impl.argtypes = [i32]
  lib.ffpa_b200_set_backend_impl.restype = ctypes.c_int
  lib.ffpa_b200_get_backend_impl.restype = i32
  lib.ffpa_b200_fwd_available.restype = i32
  lib.ffpa_b200_bwd_available.restype = i32
  lib.ffpa_b200_abi_version.restype = i32
  lib.ffpa_b200_launch_count.restype = u64
  lib.ffpa_b200_last_error.restype = ctypes.c_char_p
  return lib


def fwd_sizes(B, Hq, Hkv, Nq, Nkv, D, **kw):
  p = FwdParams()
  p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim = B, Hq, Hkv, Nq, Nkv, D
  p.fp8_smooth_k, p.fp8_pv_acc_type = 1, 1
  for k, v in kw.items():
    setattr(p, k, v)
  return p


def bwd_sizes(B, Hq, Hkv, Nq, Nkv, D, **kw):
  p = BwdParams()
  p.batch, p.heads_q, p.heads_kv, p.seqlen_q, p.seqlen_kv, p.head_dim = B, Hq, Hkv, Nq, Nkv, D
  for k, v in kw.items():
    setattr(p, k, v)
  return p
