#!/usr/bin/env python
"""Builds the PyTorch C++ extension ``ffpa_attn._C`` (csrc/ffpa_torch_binding.cpp, C++ only) in-tree:
``ffpa-attn_b200/ffpa_attn/_C<EXT_SUFFIX>`` linked against ``libffpa_b200.so`` next to it (rpath $ORIGIN).
Same include / library / ABI flags ``torch.utils.cpp_extension.CppExtension`` would pass, as one explicit
g++ command so that ``__graft_entry__.build()`` and ``make`` need no setuptools build directory."""
import os
import subprocess
import sys
import sysconfig

import torch
from torch.utils import cpp_extension as ce

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "ffpa_attn")
SRC = os.path.join(HERE, "ffpa_torch_binding.cpp")
OUT = os.path.join(PKG, "_C" + sysconfig.get_config_var("EXT_SUFFIX"))


def out_path() -> str:
  return OUT


def up_to_date() -> bool:
  if not os.path.exists(OUT):
    return False
  t = os.path.getmtime(OUT)
  deps = [SRC, os.path.join(HERE, "..", "..", "include", "ffpa_b200.h"), os.path.join(PKG, "libffpa_b200.so"), __file__]
  return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False) -> str:
  if not force and up_to_date():
    return OUT
  cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
  inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include"),
                              os.path.join(HERE, "..", "..", "include")]
  libdirs = ce.library_paths() + [os.path.join(cuda_home, "lib64")]
  cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wno-attributes",
         "-DTORCH_EXTENSION_NAME=_C", "-DTORCH_API_INCLUDE_EXTENSION_H",
         f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
  cmd += [f"-I{p}" for p in inc]
  cmd += [SRC, "-o", OUT]
  cmd += [f"-L{p}" for p in libdirs] + [f"-L{PKG}"]
  cmd += ["-lffpa_b200", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
          "-Wl,-rpath,$ORIGIN"] + [f"-Wl,-rpath,{p}" for p in ce.library_paths()]
  subprocess.check_call(cmd)
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv))
