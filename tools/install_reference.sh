#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box with gpurun).
# /root/reference is read-only, so the install runs from a copy under /tmp (pure-Python wheel: the optional
# CUDAExtension is off by default, setup.py "Default: pure Python / Triton-only build").
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
TMP=$(mktemp -d)
(cd "$SRC" && tar --exclude=./third_party --exclude=./docs -cf - .) | (cd "$TMP" && tar xf -)
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
