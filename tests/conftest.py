import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ffpa-attn_b200")
for p in (ROOT, PKG):
  if p not in sys.path:
    sys.path.insert(0, p)


def pytest_configure(config):
  config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
  import torch

  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason="no CUDA device in this container")
  for item in items:
    if "gpu" in item.keywords:
      item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
  return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _fresh_tuning_env():
  """libffpa_b200.so caches the FFPA_* tuning variables per process; tests that monkeypatch them call
  ``_C.refresh_env()`` themselves, and this makes the restored environment visible to the next test."""
  yield
  mod = sys.modules.get("ffpa_attn._C")
  if mod is not None and hasattr(mod, "refresh_env"):
    mod.refresh_env()
