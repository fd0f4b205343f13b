"""ffpa_attn -- B200-native drop-in for the ``ffpa_attn`` package's attention operator.

Importing this package loads ``libffpa_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/ffpa_b200.h``) and registers ``torch.ops.ffpa_attn._fwd_cuda`` / ``_bwd_cuda``.
"""
from .cuda import CudaBackendImpl, get_cuda_backend_impl, set_cuda_backend_impl
from .ffpa_attn_interface import ffpa_attn_func, ffpa_attn_varlen_func
from .functional import CUDABackend, FFPAAttnFunc, FFPAAttnMeta
from .host import ffpa_attn_host_func

__all__ = [
  "ffpa_attn_func",
  "ffpa_attn_varlen_func",
  "ffpa_attn_host_func",
  "CUDABackend",
  "FFPAAttnFunc",
  "FFPAAttnMeta",
  "CudaBackendImpl",
  "get_cuda_backend_impl",
  "set_cuda_backend_impl",
]
__version__ = "0.1.0+b200"
