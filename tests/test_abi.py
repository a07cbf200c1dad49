"""The C-ABI library loads without a GPU and exports every symbol include/sbx.h declares."""

import ctypes
import os
import re

import pytest

from sbsim_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
  text = open(os.path.join(ROOT, "include", "sbx.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(sbx_[a-z_]+)\s*\(", text)))


def test_header_functions_match_binding_list():
  assert _declared_functions() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_every_declared_symbol():
  lib = _lib.load()
  for name in _declared_functions():
    assert hasattr(lib, name), name


def test_struct_layouts_match_library():
  lib = _lib.load()
  ver, cb, ib = ctypes.c_int32(), ctypes.c_size_t(), ctypes.c_size_t()
  assert lib.sbx_abi_info(ctypes.byref(ver), ctypes.byref(cb), ctypes.byref(ib)) == 0
  assert ver.value == _lib.ABI_VERSION
  assert cb.value == ctypes.sizeof(_lib.SbxConfig)
  assert ib.value == ctypes.sizeof(_lib.SbxInfo)


def test_header_constants_match_binding():
  text = open(os.path.join(ROOT, "include", "sbx.h")).read()
  defs = dict(re.findall(r"#define\s+(SBX_[A-Z_]+)\s+(-?\(?-?\w+\)?)", text))

  def val(name):
    return int(defs[name].strip("()"), 0)

  assert val("SBX_MAX_ACTIONS") == _lib.MAX_ACTIONS
  assert val("SBX_MAX_HIST_BINS") == _lib.MAX_HIST_BINS
  assert val("SBX_N_DEVICE_FIELDS") == _lib.N_DEVICE_FIELDS
  assert val("SBX_DESC_DIFFUSER") == _lib.DESC_DIFFUSER
  assert val("SBX_ZONE_NONE") == _lib.ZONE_NONE
  fields = dict(re.findall(r"(SBX_F_[A-Z_0-9]+)\s*=\s*(\d+)", text))
  for name, (fid, _) in _lib.FIELDS.items():
    assert int(fields["SBX_F_" + name.upper()]) == fid, name
  diag = dict(re.findall(r"(SBX_DIAG_[A-Z_]+)\s*=\s*(\d+)", text))
  for name, idx in _lib.DIAG.items():
    assert int(diag["SBX_DIAG_" + name.upper()]) == idx, name
  assert int(diag["SBX_DIAG_N"]) == _lib.DIAG_N
  # sbx_set_option knobs
  opts = dict(re.findall(r"#define\s+(SBX_OPT_[A-Z0-9_]+)\s+(\d+)", text))
  assert len(opts) == 4
  for name, v in opts.items():
    assert int(v) == getattr(_lib, name[len("SBX_"):]), name


def test_create_fails_loudly_without_cuda_or_with_bad_config():
  """No CPU fallback: creating a handle either works on a GPU or raises."""
  cfg = _lib.SbxConfig()
  cfg.abi_version = _lib.ABI_VERSION + 7
  with pytest.raises(_lib.SbxLibraryError, match="abi_version"):
    _lib.Handle(cfg)


def test_no_oracle_in_product():
  """Nothing under sbsim_b200/ or integration/ may import or execute the oracle."""
  walks = [w for pkg in ("sbsim_b200", "integration") for w in os.walk(os.path.join(ROOT, pkg))]
  for dirpath, _, files in walks:
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        src = open(os.path.join(dirpath, f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
        assert "refshim" not in src, f
