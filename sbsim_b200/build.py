"""Builds libsbx.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "sbx_api.cu")
DEPS = sorted(glob.glob(os.path.join(_HERE, "csrc", "*.cu")) + glob.glob(os.path.join(_HERE, "csrc", "*.cuh"))) + [
    os.path.join(os.path.dirname(_HERE), "include", "sbx.h")]
OUT = os.path.join(_HERE, "lib", "libsbx.so")
STAMP = OUT + ".srchash"      # hash of the sources + flags the library was built from (travels with it)

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # one IEEE rounding per fp32/fp64 op, as the reference's eager arithmetic has
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-shared", "-Xcompiler", "-fPIC",
]


def source_hash() -> str:
  h = hashlib.sha256()
  for d in DEPS:
    h.update(os.path.basename(d).encode())
    with open(d, "rb") as f:
      h.update(f.read())
  h.update(" ".join(NVCC_FLAGS + os.environ.get("NVCC_EXTRA", "").split()).encode())
  return h.hexdigest()


def needs_build() -> bool:
  """True unless libsbx.so was built from exactly these sources and flags (content hash, not mtime)."""
  if not os.path.exists(OUT) or not os.path.exists(STAMP):
    return True
  with open(STAMP) as f:
    return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
  if not force and not needs_build():
    return OUT
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  nvcc = os.environ.get("NVCC", "nvcc")
  extra = os.environ.get("NVCC_EXTRA", "").split()
  out = os.environ.get("SBX_LIB_OUT") or OUT     # variant builds for A/B profiling
  cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
  res = subprocess.run(cmd, capture_output=True, text=True)
  if res.returncode != 0:
    sys.stderr.write(res.stdout + res.stderr)
    raise RuntimeError("nvcc failed building libsbx.so")
  if verbose:
    sys.stderr.write(res.stderr)
  if out == OUT:
    with open(STAMP, "w") as f:
      f.write(source_hash() + "\n")
  return out


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
